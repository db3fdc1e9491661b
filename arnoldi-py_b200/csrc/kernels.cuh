// kernels.cuh -- argument blocks and launcher prototypes of the n-length kernels.
#pragma once

#include "common.cuh"
#include "state.cuh"

namespace ab200 {

// ---------------------------------------------------------------- orthogonalisation
struct OrthoArgs {
  const cplx* U;     // basis, column-major, un-normalised columns (V_i = scale[i] * U_i)
  cplx* w;           // vector being orthogonalised (column j+1 of the basis, or a scratch vector)
  int64_t n;         // 16-byte elements per column: local rows, or ceil(rows / 2) pairs when real
  int64_t ld;        // leading dimension of U in 16-byte elements
  int real;          // basis stored real: an element is two consecutive real rows
  int ncols;         // c = number of basis columns to orthogonalise against
  int p1_col0;       // pass 1 only: this launch covers columns [p1_col0, p1_col0 + p1_ncols)
  int p1_ncols;      //   (c > 128 is swept in column groups of at most 128)
  int j;             // Arnoldi step (H column) -- ncols - 1 inside an expansion
  int round;         // 1 or 2 (DGKS repeat)
  int accumulate;    // h += (round 2) instead of h =
  int finalize;      // 1: write H[j+1, j], scale[j+1], breakdown flag when the step ends and count
                     //    an Arnoldi step; 2: the same without counting (a column normalised
                     //    outside an expansion); 0: nothing
  int grid_cap;      // capacity (in blocks) of the partial buffers
  int stages;        // cp.async staging depth of the fused sweep (0 = automatic)
  int fused_r;       // rows pairs per lane and chunk in the fused sweep (0 = automatic)
  double tol;        // breakdown threshold (absolute, ortho.py:107)
  double eta;        // DGKS factor (ortho.py:101)
  double* scale;     // [max_dim + 1]
  cplx* hcol;        // column j of the device copy of H (max_dim + 1 entries)
  cplx* coef;        // [max_dim + 1] pass-2 coefficients scale[i] * h_i of the current round
  cplx* part;        // [(max_dim + 1) * grid_cap] per-block partial dot products
  double* npart;     // [grid_cap] per-block partial norms
  unsigned* ticket;  // last-block ticket
  StepCtl* ctl;
  int* step_flag;    // nullptr, or where to record that this step ran the second round
  PeerComm comm;
};

cudaError_t launch_cgs_pass1(const OrthoArgs& a, int num_sms, cudaStream_t st, int grid_mult);
cudaError_t launch_cgs_fused(const OrthoArgs& a, int num_sms, cudaStream_t st, int grid_mult,
                             int variant, int fused_ct);
cudaError_t launch_cgs_pass2(const OrthoArgs& a, int num_sms, cudaStream_t st, int grid_mult);
cudaError_t launch_mgs_step(const OrthoArgs& a, int i, int num_sms, cudaStream_t st,
                            int grid_mult);

cudaError_t launch_peer_barrier(const PeerComm& pc, StepCtl* ctl, int real_mode, cudaStream_t st);

// ---------------------------------------------------------------- SpMV
struct SpmvArgs {
  const void* indptr;      // [n + 1], 32- or 64-bit, relative to the local block
  const int32_t* indices;  // [nnz] column ids: < n_local -> local x, else ghost[id - n_local]
  const void* values;      // [nnz] float64 or complex128
  const int64_t* rowblk;   // [nblocks + 1] first row of each nnz tile
  const void* x;           // local part of the input vector: complex128, or float64 when real
  const void* ghost;       // halo entries received from peers (nullptr on one GPU), same type
  void* y;                 // output (local rows), same type
  int real;                // vectors are float64 (basis stored real)
  const double* xscale;    // nullptr or pointer to the lazy scale of x
  int64_t n;               // local rows
  int64_t n_local_cols;    // column ids below this are local
  int nblocks;
  int tile;                // nnz staged per block iteration
  int threads;             // block size: 128 or 256
  int long_rows;           // some row has more than 16 entries: warp-per-row path compiled in
  int variant;             // 0 = bulk-copy (TMA) pipeline when no long rows; 1 = plain tile kernel;
                           // 2 = cp.async streaming kernel (round-1 default, kept for A/B)
  int num_sms;
  const StepCtl* ctl;      // nullptr for the stand-alone entry point
  // bulk-copy pipeline shape (spmv_bulk_kernel)
  int stages;              // shared-memory ring depth
  int rp_cap;              // row pointers staged per tile (multiple of 4)
  int bps;                 // resident blocks per SM to launch (0 = what fits)
  // skewed rows with local columns (spmv_ring_kernel): the x entries around the diagonal live in
  // a shared-memory ring
  int window;              // use the ring kernel
  int win_cap;             // ring capacity in float64 entries (power of two; complex: half as many)
  int win_half;            // entries kept on each side of a round's rows
  int ring_warps;          // consumer warps per block
  int contig;              // tile kernel: persistent blocks over contiguous tile ranges (L1 reuse of x)
  // halo read straight from the owners' HBM (multi-GPU "pull"): ghost entry g of rank q lives at
  // peer_col[q][ghost_off[g]]; entries [seg_start[q], seg_start[q+1]) belong to rank q
  int direct_halo;
  int nranks;
  const void* peer_col[kMaxRanks];
  int64_t seg_start[kMaxRanks + 1];
  const int64_t* ghost_off;
};

cudaError_t launch_spmv(const SpmvArgs& a, int indptr_bits, int value_kind, cudaStream_t st);
size_t spmv_ring_smem(int win_cap, int nwarps, int x_bytes);
// fraction (in 1/1024ths, over a sample of the rows) of the entries whose column lies within
// `half` of their row: decides whether staging an x window in shared memory pays
cudaError_t launch_spmv_locality(const void* indptr, int indptr_bits, const int32_t* indices, int64_t n,
                                 int64_t n_local_cols, int64_t half, unsigned long long* out2,
                                 cudaStream_t st);
// builds rowblk[b] = first row whose indptr >= b * tile  (b = 0..nblocks), on device
cudaError_t launch_spmv_maxrow(const void* indptr, int indptr_bits, int64_t n, int* out,
                               cudaStream_t st);
cudaError_t launch_spmv_plan(const void* indptr, int indptr_bits, int64_t n, int64_t nnz, int tile,
                             int nblocks, int64_t* rowblk, cudaStream_t st);

// ---------------------------------------------------------------- halo (multi-GPU)
struct HaloArgs {
  const cplx* peer_base[kMaxRanks];  // peer_base[q] = rank q's basis (IPC mapped; own pointer for q == rank)
  int64_t peer_ld[kMaxRanks];        // leading dimension of rank q's basis
  int64_t seg_start[kMaxRanks + 1];  // ghost entries [seg_start[q], seg_start[q+1]) are owned by rank q
  const int64_t* src_off;            // [nghost] row offset inside the owner's block
  cplx* ghost;                       // [nghost] destination
  int64_t nghost;
  int col;                           // basis column to fetch
  int real;                          // columns hold float64 (column stride = peer_ld doubles... see kernel)
  int nranks;
  const StepCtl* ctl;
};
cudaError_t launch_halo_gather(const HaloArgs& a, int num_sms, cudaStream_t st);

struct HaloPushArgs {
  const cplx* U_col;                      // my column j (un-normalised; the receiver applies scale[j])
  const int64_t* send_idx;                // [nsend] local row of every entry some peer needs
  int64_t send_ptr[kMaxRanks + 1];        // entries [send_ptr[r], send_ptr[r+1]) go to rank r
  int64_t dst_off[kMaxRanks];             // where my block starts in rank r's ghost buffer
  cplx* peer_ghost[kMaxRanks];            // IPC-mapped ghost buffers
  unsigned long long* peer_hflags[kMaxRanks];  // IPC-mapped "halo delivered" flags [kMaxRanks]
  int64_t nsend;
  unsigned long long seq;
  unsigned* ticket;
  int rank, nranks;
  int real;                               // entries are float64
  const StepCtl* ctl;
};
cudaError_t launch_halo_push(const HaloPushArgs& a, int num_sms, cudaStream_t st);
cudaError_t launch_halo_wait(const unsigned long long* hflags, unsigned need_mask,
                             unsigned long long seq, StepCtl* ctl, cudaStream_t st);

// ---------------------------------------------------------------- restart + helpers
struct RestartArgs {
  int copy_tail;        // 1: also U[:, p] = scale_m * U[:, m] (a restart); 0: plain U[:, :p] = U[:, :m] q
  cplx* U;              // basis (first of the m input columns), updated in place
  int64_t n, ld;
  int m, p;
  const cplx* q;        // device copy of Q[:, :p], row i pre-multiplied by scale[i]; layout [i * p + k]
  double scale_m;       // scale of column m (applied while it is copied to column p)
  int real;             // basis stored real: n, ld count pairs of rows and q is real (.x)
};
cudaError_t launch_restart(const RestartArgs& a, int num_sms, cudaStream_t st, int variant);

// y = (*scale) * x over n 16-byte elements (device-operator path: the operator sees v_j itself)
cudaError_t launch_scaled_copy(const cplx* x, cplx* y, int64_t n, const double* scale, int real,
                               int num_sms, cudaStream_t st);

cudaError_t launch_pack_real(const cplx* src, double* dst, int64_t n, int num_sms, cudaStream_t st);
cudaError_t launch_unpack_real(const double* src, cplx* dst, int64_t n, double scale, int num_sms,
                               cudaStream_t st);

// U[:, col] *= scale[col]; scale[col] = 1   for col in [col0, col0 + ncols)
cudaError_t launch_materialize(cplx* U, int64_t n, int64_t ld, int col0, int ncols, double* scale,
                               int num_sms, cudaStream_t st);

}  // namespace ab200
