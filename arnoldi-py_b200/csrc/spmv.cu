// spmv.cu -- CSR sparse matrix x vector  (the reference's `A @ V[:, j]`, decomposition.py:57-58,
// which lands in scipy's csr_matvec).  Values float64 or complex128; vectors complex128, or
// float64 while the basis is stored real.
//
// Work decomposition, common to every kernel here ("nnz tiles with row-aligned edges"): the nnz
// range is cut into tiles of `tile` entries; tile b starts at the first row whose indptr >=
// b * tile (rowblk[], built once per matrix on the device), so every tile owns whole rows and
// about the same number of non-zeros no matter how skewed the row lengths are (the merge-path
// split, snapped to row boundaries).  Rows are summed in STORED ORDER with separate multiply and
// add -- the operation order of scipy's csr_matvec -- wherever one thread sums a row, so those
// rows are bit-identical to scipy.
//
//   spmv_bulk_kernel    rows of <= 16 entries (mark, the Laplacians): a producer warp moves each
//                       tile into a shared-memory ring with bulk copies (TMA) + mbarriers, one
//                       consumer thread per row.  The default for such operators.
//   spmv_stream_kernel  the same with cp.async staging and block barriers (round 1; A/B only).
//   spmv_ring2_kernel   skewed rows, float64 vectors: one warp per tile, strips staged with cp.async,
//                       x gathered from a shared-memory ring that slides along the diagonal.
//   spmv_ring_kernel    its predecessor (strips in registers, 15 warps; spmv_variant = 4): x gathered from a
//                       shared-memory ring that slides along the diagonal.
//   spmv_tile_kernel    everything else: one block per tile, x gathered from global memory, a
//                       warp for segments > 16 entries, rows longer than a tile carried through y.
#include "kernels.cuh"

namespace ab200 {

constexpr int kShortRow = 16;  // segments up to this long are summed by one thread, in stored order

template <typename T>
struct ValOps;
template <>
struct ValOps<double> {
  // (a + 0i) * x, rounded like numpy/scipy's complex multiply with a zero imaginary part
  static __device__ __forceinline__ cplx mul(double a, cplx x) {
    return make_double2(__dmul_rn(a, x.x), __dmul_rn(a, x.y));
  }
};
template <>
struct ValOps<cplx> {
  // (ar*xr - ai*xi) + i (ar*xi + ai*xr), each product and sum rounded separately (no FMA)
  static __device__ __forceinline__ cplx mul(cplx a, cplx x) {
    return make_double2(__dsub_rn(__dmul_rn(a.x, x.x), __dmul_rn(a.y, x.y)),
                        __dadd_rn(__dmul_rn(a.x, x.y), __dmul_rn(a.y, x.x)));
  }
};
__device__ __forceinline__ cplx cadd_rn(cplx a, cplx b) {
  return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y));
}
// real-storage mode: x and y are float64 (A must be float64 too); same operation order, so the
// result equals the real part of what the complex path computes, bit for bit
__device__ __forceinline__ double cadd_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double vmul(double a, double x) { return __dmul_rn(a, x); }
__device__ __forceinline__ cplx vmul(double a, cplx x) { return ValOps<double>::mul(a, x); }
__device__ __forceinline__ cplx vmul(cplx a, cplx x) { return ValOps<cplx>::mul(a, x); }
__device__ __forceinline__ double vmul(cplx a, double x) { return 0.0 * a.x * x; }  // never used
__device__ __forceinline__ double ld_ro(const double* p) { return __ldg(p); }
__device__ __forceinline__ double cscale(double a, double s) { return a * s; }
template <typename XT> __device__ __forceinline__ XT xzero();
template <> __device__ __forceinline__ double xzero<double>() { return 0.0; }
template <> __device__ __forceinline__ cplx xzero<cplx>() { return make_double2(0.0, 0.0); }

template <typename IdxT>
__global__ void spmv_maxrow_kernel(const IdxT* __restrict__ indptr, int64_t n, int* __restrict__ out) {
  int m = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t len = (int64_t)indptr[r + 1] - (int64_t)indptr[r];
    const int l = len > 0x7fffffff ? 0x7fffffff : (int)len;
    m = l > m ? l : m;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const int t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);  // integer max: order-independent
}

template <typename IdxT>
__global__ void spmv_plan_kernel(const IdxT* __restrict__ indptr, int64_t n, int64_t nnz, int tile,
                                 int nblocks, int64_t* __restrict__ rowblk) {
  // rowblk[b] = first row of tile b; rowblk[nblocks + 1 + b] = indptr of that row
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nblocks) return;
  int64_t* __restrict__ nnzblk = rowblk + nblocks + 1;
  if (b == nblocks) {
    rowblk[b] = n;
    nnzblk[b] = nnz;
    return;
  }
  const int64_t target = (int64_t)b * tile;
  // lower_bound over indptr[0..n]: first row r with indptr[r] >= target
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)indptr[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  rowblk[b] = lo;
  nnzblk[b] = (int64_t)indptr[lo];
}

// LONG = false: the matrix has no row longer than kShortRow, the warp-per-row machinery is
// compiled out (lean kernel for banded operators: mark, Laplacians).
template <typename IdxT, typename ValT, typename XT, int kSpmvThreads, bool LONG>
__global__ void __launch_bounds__(kSpmvThreads) spmv_tile_kernel(SpmvArgs a) {
  if (a.ctl != nullptr && a.ctl->stop) return;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  ValT* sval = reinterpret_cast<ValT*>(smem_raw);
  int32_t* scol = reinterpret_cast<int32_t*>(smem_raw + (size_t)(a.tile + 4) * sizeof(ValT));
  __shared__ int s_nlong;
  int64_t* s_long_row = reinterpret_cast<int64_t*>(
      smem_raw + (size_t)(a.tile + 4) * (sizeof(ValT) + sizeof(int32_t)) + 8);  // [tile / 16 + 1]

  const IdxT* __restrict__ indptr = static_cast<const IdxT*>(a.indptr);
  const ValT* __restrict__ values = static_cast<const ValT*>(a.values);
  const int32_t* __restrict__ indices = a.indices;
  const XT* __restrict__ x = static_cast<const XT*>(a.x);
  const XT* __restrict__ ghost = static_cast<const XT*>(a.ghost);
  XT* __restrict__ yout = static_cast<XT*>(a.y);
  const int64_t nloc = a.n_local_cols;
  const double xs = a.xscale ? *a.xscale : 1.0;
  const int tid = threadIdx.x;
  const int tile = a.tile;
  const unsigned long long l2pol = l2_policy_evict_first();  // CSR streams must not evict x from L2

  // a.contig: each block walks a CONTIGUOUS range of tiles, so the x entries its SM gathers
  // slide along the band of the operator and stay in L1 (otherwise: one tile per block)
  const int per = a.contig ? (a.nblocks + (int)gridDim.x - 1) / (int)gridDim.x : 1;
  const int bfirst = a.contig ? (int)blockIdx.x * per : (int)blockIdx.x;
  const int bstep = a.contig ? 1 : (int)gridDim.x;
  const int blast = a.contig ? (bfirst + per < a.nblocks ? bfirst + per : a.nblocks) : a.nblocks;
  for (int b = bfirst; b < blast; b += bstep) {
    const int64_t r0 = a.rowblk[b], r1 = a.rowblk[b + 1];
    if (r0 >= r1) continue;
    const int64_t s = (int64_t)indptr[r0], e = (int64_t)indptr[r1];
    // at least one iteration so that empty rows get y = 0
    for (int64_t cs = s; cs == s || cs < e; cs += tile) {
      const int64_t ce = (cs + tile < e) ? cs + tile : e;
      const int cnt = (int)(ce - cs);
      // stage from the 4-entry-aligned address at or below cs so the 128-bit path always applies;
      // the (at most 3) leading entries belong to the previous tile and are never read back
      const int off = (int)(cs & 3);
      const int64_t ca = cs - off;
      const int tot = cnt + off;
      __syncthreads();  // previous iteration's readers are done with smem
      if (tid == 0) s_nlong = 0;
      {
        const int nvec = tot >> 2;
        const int4* c4 = reinterpret_cast<const int4*>(indices + ca);
        for (int k = tid; k < nvec; k += kSpmvThreads)
          reinterpret_cast<int4*>(scol)[k] = ld_stream_ef(c4 + k, l2pol);
        for (int k = (nvec << 2) + tid; k < tot; k += kSpmvThreads) scol[k] = indices[ca + k];
        const double2* v2 = reinterpret_cast<const double2*>(values + ca);
        if (sizeof(ValT) == 8) {
          for (int k = tid; k < (tot >> 1); k += kSpmvThreads)
            reinterpret_cast<double2*>(sval)[k] = ld_stream_ef(v2 + k, l2pol);
          if ((tot & 1) && tid == 0) sval[tot - 1] = values[ca + tot - 1];
        } else {
          for (int k = tid; k < tot; k += kSpmvThreads)
            reinterpret_cast<double2*>(sval)[k] = ld_stream_ef(v2 + k, l2pol);
        }
      }
      __syncthreads();
      // ---- one thread per row, stored order
      for (int64_t row = r0 + tid; row < r1; row += kSpmvThreads) {
        const int64_t rs = (int64_t)indptr[row], re = (int64_t)indptr[row + 1];
        if (re <= cs && !(rs == re && cs == s)) continue;  // finished in an earlier tile
        if (rs >= ce && rs != re) continue;                // starts in a later tile
        const int64_t lo = rs > cs ? rs : cs;
        const int64_t hi = re < ce ? re : ce;
        const int len = (int)(hi - lo);
        if (LONG && len > kShortRow) {  // handed to a whole warp below
          const int slot = atomicAdd(&s_nlong, 1);
          s_long_row[slot] = row;
          continue;
        }
        XT acc = (rs < cs) ? yout[row] : xzero<XT>();
        int k = (int)(lo - ca);
        const int kend = (int)(hi - ca);

        for (; k + 4 <= kend; k += 4) {
          XT xv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t col = scol[k + u];
            xv[u] = (col < nloc) ? ld_ro(x + col) : ld_ro(ghost + (col - nloc));
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) acc = cadd_rn(acc, vmul(sval[k + u], xv[u]));
        }

        for (; k < kend; ++k) {
          const int64_t col = scol[k];
          const XT xv = (col < nloc) ? ld_ro(x + col) : ld_ro(ghost + (col - nloc));
          acc = cadd_rn(acc, vmul(sval[k], xv));
        }
        if (re <= ce) acc = cscale(acc, xs);  // row complete: apply the lazy scale of x
        yout[row] = acc;
      }
      __syncthreads();
      // ---- longer segments: one warp per row, lanes stride the segment, fixed-order butterfly
      const int nlong = LONG ? s_nlong : 0;
      for (int l = (tid >> 5); l < nlong; l += kSpmvThreads / kWarp) {
        const int64_t row = s_long_row[l];
        const int64_t rs = (int64_t)indptr[row], re = (int64_t)indptr[row + 1];
        const int64_t lo = rs > cs ? rs : cs;
        const int64_t hi = re < ce ? re : ce;
        XT acc = xzero<XT>();
        for (int k = (int)(lo - ca) + (tid & 31); k < (int)(hi - ca); k += kWarp) {
          const int64_t col = scol[k];
          const XT xv = (col < nloc) ? ld_ro(x + col) : ld_ro(ghost + (col - nloc));
          acc = cadd_rn(acc, vmul(sval[k], xv));
        }
        acc = warp_sum(acc);
        if ((tid & 31) == 0) {
          XT t = (rs < cs) ? cadd_rn(yout[row], acc) : acc;
          if (re <= ce) t = cscale(t, xs);
          yout[row] = t;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ streaming kernel
// For matrices whose rows all have at most kShortRow entries (mark, the Laplacians).
// Persistent blocks walk the tiles; the NEXT tile's values, column ids and row pointers are
// staged with cp.async (LDGSTS: no register round trip) while the current tile's rows are
// summed, so the four dependent memory round trips of the plain kernel (tile bounds -> row
// pointers -> staging -> gathers) are taken off the critical path.  Same arithmetic, same
// order: y is bit-identical to scipy's csr_matvec.
__device__ __forceinline__ void cpa16(void* dst, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cpa_small(void* dst, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(src), "n"(BYTES) : "memory");
}

template <typename IdxT, typename ValT, typename XT, int THREADS>
__global__ void __launch_bounds__(THREADS) spmv_stream_kernel(SpmvArgs a) {
  if (a.ctl != nullptr && a.ctl->stop) return;
  constexpr int RP = 2 * THREADS + 2;  // row pointers staged per tile (more rows: read from global)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int cap = a.tile + kShortRow + 8;  // entries per stage (a tile overshoots by < one row)
  const size_t stage_bytes =
      ((size_t)cap * (sizeof(ValT) + sizeof(int32_t)) + (size_t)RP * sizeof(IdxT) + 15) / 16 * 16;
  const IdxT* __restrict__ indptr = static_cast<const IdxT*>(a.indptr);
  const ValT* __restrict__ values = static_cast<const ValT*>(a.values);
  const int32_t* __restrict__ indices = a.indices;
  const XT* __restrict__ x = static_cast<const XT*>(a.x);
  const XT* __restrict__ ghost = static_cast<const XT*>(a.ghost);
  XT* __restrict__ yout = static_cast<XT*>(a.y);
  const int64_t* __restrict__ rowblk = a.rowblk;
  const int64_t* __restrict__ nnzblk = a.rowblk + a.nblocks + 1;
  const int64_t nloc = a.n_local_cols;
  const double xs = a.xscale ? *a.xscale : 1.0;
  const int tid = threadIdx.x;

  auto stage = [&](int b, int st) {
    unsigned char* base = smem_raw + (size_t)st * stage_bytes;
    ValT* sval = reinterpret_cast<ValT*>(base);
    int32_t* scol = reinterpret_cast<int32_t*>(base + (size_t)cap * sizeof(ValT));
    IdxT* srow = reinterpret_cast<IdxT*>(base + (size_t)cap * (sizeof(ValT) + sizeof(int32_t)));
    const int64_t r0 = rowblk[b], r1 = rowblk[b + 1];
    const int64_t s = nnzblk[b], e = nnzblk[b + 1];
    const int64_t ca = s & ~(int64_t)3;
    const int tot = (int)(e - ca);
    const int nvec = (tot + 3) >> 2;  // the arrays are padded by 4 entries: reading past e is safe
    for (int k = tid; k < nvec; k += THREADS) cpa16(scol + 4 * k, indices + ca + 4 * k);
    if (sizeof(ValT) == 8) {
      for (int k = tid; k < 2 * nvec; k += THREADS) cpa16(sval + 2 * k, values + ca + 2 * k);
    } else {
      for (int k = tid; k < 4 * nvec; k += THREADS) cpa16(sval + k, values + ca + k);
    }
    int64_t nr = r1 - r0 + 1;
    if (nr > RP) nr = RP;
    for (int k = tid; k < (int)nr; k += THREADS) cpa_small<sizeof(IdxT)>(srow + k, indptr + r0 + k);
  };

  int b = blockIdx.x;
  if (b < a.nblocks) stage(b, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  int st = 0;
  for (; b < a.nblocks; b += gridDim.x, st ^= 1) {
    const int bn = b + gridDim.x;
    if (bn < a.nblocks) stage(bn, st ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();  // this tile's stage is complete and visible to every thread
    unsigned char* base = smem_raw + (size_t)st * stage_bytes;
    const ValT* sval = reinterpret_cast<const ValT*>(base);
    const int32_t* scol = reinterpret_cast<const int32_t*>(base + (size_t)cap * sizeof(ValT));
    const IdxT* srow = reinterpret_cast<const IdxT*>(base + (size_t)cap * (sizeof(ValT) + sizeof(int32_t)));
    const int64_t r0 = rowblk[b], r1 = rowblk[b + 1];
    const int64_t ca = nnzblk[b] & ~(int64_t)3;
    for (int64_t rl = tid; rl < r1 - r0; rl += THREADS) {
      const int64_t rs = rl + 1 < RP ? (int64_t)srow[rl] : (int64_t)indptr[r0 + rl];
      const int64_t re = rl + 1 < RP ? (int64_t)srow[rl + 1] : (int64_t)indptr[r0 + rl + 1];
      XT acc = xzero<XT>();
      int k = (int)(rs - ca);
      const int kend = (int)(re - ca);
      for (; k + 4 <= kend; k += 4) {
        XT xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t col = scol[k + u];
          xv[u] = (col < nloc) ? ld_ro(x + col) : ld_ro(ghost + (col - nloc));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc = cadd_rn(acc, vmul(sval[k + u], xv[u]));
      }
      for (; k < kend; ++k) {
        const int64_t col = scol[k];
        const XT xv = (col < nloc) ? ld_ro(x + col) : ld_ro(ghost + (col - nloc));
        acc = cadd_rn(acc, vmul(sval[k], xv));
      }
      yout[r0 + rl] = cscale(acc, xs);
    }
    __syncthreads();  // everyone is done with this stage before it is refilled
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ bulk-copy (TMA) pipeline
// Same decomposition and the same arithmetic as spmv_stream_kernel (rows of <= kShortRow
// entries, one thread per row, stored order, separate multiply and add: bit-identical to
// scipy's csr_matvec), different data movement.  ncu on the cp.async version: 367 warp
// instructions per 32 rows, three quarters of them address arithmetic and LDGSTS issue for
// the staging -- the kernel was issue-bound at 0.68 of the copy bandwidth.  Here a producer
// warp moves each tile with three bulk copies (cp.async.bulk -> UBLKCP: values, column ids,
// row pointers) into a ring of shared-memory stages and signals an mbarrier per stage
// (complete_tx); the consumer warps only wait, gather and accumulate, and hand the stage back
// through a second mbarrier.  No block-wide barrier in the loop: warps drift apart by up to
// `stages - 1` tiles.
struct TileHdr {   // written by the producer before it arms the stage's barrier
  int64_t ca;      // nnz index staged at slot 0 (4-entry aligned, <= first entry of the tile)
  int64_t r0;      // first row of the tile
  int nrows;
  int rpoff;       // slot of row r0's pointer in the staged row-pointer window
  int nrp;         // row pointers staged
  int pad;
};

// coherent load of a halo entry straight from its owner's HBM over NVLink
__device__ __forceinline__ double ld_peer(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ cplx ld_peer(const cplx* p) {
  cplx v;
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
template <typename XT>
__device__ __forceinline__ XT halo_entry(const SpmvArgs& a, int64_t g) {
  if (!a.direct_halo) return ld_ro(static_cast<const XT*>(a.ghost) + g);
  int q = 0;
#pragma unroll
  for (int r = 1; r < kMaxRanks; ++r)
    if (r < a.nranks && g >= a.seg_start[r]) q = r;
  return ld_peer(static_cast<const XT*>(a.peer_col[q]) + a.ghost_off[g]);
}

template <typename IdxT, typename ValT, typename XT, bool HALO>
__global__ void __launch_bounds__(288) spmv_bulk_kernel(SpmvArgs a) {
  if (a.ctl != nullptr && a.ctl->stop) return;
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  const int nthr = blockDim.x - kWarp;  // consumer threads; the last warp is the producer
  const int ncw = nthr >> 5;
  const int S = a.stages;
  const int cap = a.tile + kShortRow + 8;  // entries per stage (a tile overshoots by < one row)
  const int rpc = a.rp_cap;
  const size_t off_col = (size_t)cap * sizeof(ValT);
  const size_t off_row = off_col + (size_t)cap * sizeof(int32_t);
  const size_t off_hdr = off_row + (size_t)rpc * sizeof(IdxT);
  const size_t stage_bytes = (off_hdr + sizeof(TileHdr) + 127) / 128 * 128;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(bulk_smem + (size_t)S * stage_bytes);
  unsigned long long* empty = full + S;

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, ncw);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const IdxT* __restrict__ indptr = static_cast<const IdxT*>(a.indptr);
  const int ntiles = (a.nblocks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid >= nthr) {
    // ---------------- producer warp: lane 0 issues, the others leave
    if (tid != nthr) return;
    const ValT* __restrict__ values = static_cast<const ValT*>(a.values);
    const int32_t* __restrict__ indices = a.indices;
    const int64_t* __restrict__ rowblk = a.rowblk;
    const int64_t* __restrict__ nnzblk = a.rowblk + a.nblocks + 1;
    constexpr int RA = 16 / (int)sizeof(IdxT);  // row pointers per 16 bytes
    int st = 0;
    unsigned ph = 0;
    for (int i = 0, b = blockIdx.x; i < ntiles; ++i, b += gridDim.x) {
      if (i >= S) mbar_wait(empty + st, ph ^ 1u);  // the consumers released the previous use
      unsigned char* base = bulk_smem + (size_t)st * stage_bytes;
      const int64_t r0 = rowblk[b], r1 = rowblk[b + 1];
      const int64_t s0 = nnzblk[b], e0 = nnzblk[b + 1];
      const int64_t ca = s0 & ~(int64_t)3;
      const unsigned ncopy = (unsigned)((e0 - ca + 3) & ~(int64_t)3);  // arrays carry >= 4 spare entries
      const int64_t r0a = r0 & ~(int64_t)(RA - 1);
      int64_t nrp = r1 - r0a + 1;
      if (nrp > rpc) nrp = rpc;
      nrp = (nrp + RA - 1) & ~(int64_t)(RA - 1);  // indptr carries >= 8 spare entries
      TileHdr* h = reinterpret_cast<TileHdr*>(base + off_hdr);
      h->ca = ca;
      h->r0 = r0;
      h->nrows = (int)(r1 - r0);
      h->rpoff = (int)(r0 - r0a);
      h->nrp = (int)nrp;
      const unsigned bv = ncopy * (unsigned)sizeof(ValT), bc = ncopy * 4u,
                     br = (unsigned)nrp * (unsigned)sizeof(IdxT);
      mbar_expect_tx(full + st, bv + bc + br);  // release: the header is visible with the phase
      if (ncopy) {
        bulk_g2s(base, values + ca, bv, full + st);
        bulk_g2s(base + off_col, indices + ca, bc, full + st);
      }
      bulk_g2s(base + off_row, indptr + r0a, br, full + st);
      if (++st == S) st = 0, ph ^= 1u;
    }
    return;
  }

  // ---------------- consumer warps
  const XT* __restrict__ x = static_cast<const XT*>(a.x);
  XT* __restrict__ yout = static_cast<XT*>(a.y);
  const int nloc = a.n_local_cols > 0x7fffffff ? 0x7fffffff : (int)a.n_local_cols;
  const double xs = a.xscale ? *a.xscale : 1.0;
  int st = 0;
  unsigned ph = 0;
  for (int i = 0; i < ntiles; ++i) {
    mbar_wait(full + st, ph);
    const unsigned char* base = bulk_smem + (size_t)st * stage_bytes;
    const ValT* sval = reinterpret_cast<const ValT*>(base);
    const int32_t* scol = reinterpret_cast<const int32_t*>(base + off_col);
    const IdxT* srow = reinterpret_cast<const IdxT*>(base + off_row);
    const TileHdr h = *reinterpret_cast<const TileHdr*>(base + off_hdr);
    const IdxT ca = (IdxT)h.ca;
    for (int rl = tid; rl < h.nrows; rl += nthr) {
      IdxT rs, re;
      const int slot = h.rpoff + rl;
      if (slot + 1 < h.nrp) {
        rs = srow[slot];
        re = srow[slot + 1];
      } else {  // more rows than staged pointers (many empty rows): read them from global
        rs = indptr[h.r0 + rl];
        re = indptr[h.r0 + rl + 1];
      }
      int k = (int)(rs - ca);
      const int kend = (int)(re - ca);
      XT acc = xzero<XT>();
      for (; k + 4 <= kend; k += 4) {
        XT xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int col = scol[k + u];
          if (HALO)
            xv[u] = (col < nloc) ? ld_ro(x + col) : halo_entry<XT>(a, (int64_t)col - nloc);
          else
            xv[u] = ld_ro(x + col);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc = cadd_rn(acc, vmul(sval[k + u], xv[u]));
      }
      for (; k < kend; ++k) {
        const int col = scol[k];
        XT xv;
        if (HALO)
          xv = (col < nloc) ? ld_ro(x + col) : halo_entry<XT>(a, (int64_t)col - nloc);
        else
          xv = ld_ro(x + col);
        acc = cadd_rn(acc, vmul(sval[k], xv));
      }
      yout[h.r0 + rl] = cscale(acc, xs);
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(empty + st);  // release: this warp is done with the stage
    if (++st == S) st = 0, ph ^= 1u;
  }
}

// ------------------------------------------------------------------ skewed rows: sliding x ring in shared memory
// For operators whose rows are long and uneven (BASELINE config 4: Pareto row lengths, columns in
// a band around the diagonal plus a sparse far tail) the CSR streaming roofline is not what limits
// the tile kernel: every gathered 8/16-byte x entry costs its own L1 tag look-up and a 32-byte
// sector through L2 (ncu: 10 sectors per request, L1 hit rate 7%), about one entry per SM clock.
// This kernel gathers from SHARED memory instead:
//   * a block owns a CONTIGUOUS range of the plan's nnz tiles (row-aligned, ~equal non-zeros), so
//     the x entries it can touch around the diagonal slide forward monotonically.  They live in a
//     power-of-two ring (entry c at slot c & (W-1)); a producer thread extends the ring with bulk
//     copies (cp.async.bulk + mbarrier) one round AHEAD of the consumers, so every x entry is read
//     from L2/HBM about once per block and the copy overlaps the arithmetic;
//   * a round gives every consumer warp one tile.  The warp works alone (no block barrier):
//     lanes run ALONG THE ENTRIES -- 128-bit loads of 4 column ids and 4 values per lane, all
//     in flight together -- turn them into products val * x (x from the ring; the few entries
//     outside it, and ghost entries, are gathered from global memory with the loads issued
//     together) and park the products in the warp's own strip of shared memory; then lanes run
//     ALONG THE ROWS and sum each row's products in STORED ORDER (separate multiply and add =
//     scipy's csr_matvec order: rows of up to 32 entries are bit-identical to scipy); longer
//     segments are summed by the whole warp (fixed-order butterfly);
//   * work per warp is balanced in non-zeros, not rows (the merge-path split, snapped to rows);
//     a row longer than a tile is walked by its warp in strips, carried through y.
// Rounds are ordered by mbarriers only: ring_full[t % depth] (the producer's copies for round t
// have landed) and round_done[t % depth] (every consumer warp finished round t).  The warps may
// drift up to depth - 1 rounds apart -- a tile that holds a 2000-entry row takes its warp four
// strips while the next three warps find their tiles empty -- and the producer only overwrites
// ring slots below the window of the oldest round still in flight.
constexpr int kRingIters = 5;                   // 128-entry groups per strip
constexpr int kRingStrip = 128 * kRingIters;    // products a warp parks per strip
constexpr int kRingLongSeg = 32;                // segments longer than this are summed by the warp
constexpr int kRingRowGroups = 5;               // groups of 32 row pointers fetched ahead per tile

constexpr int kRingDepthLog = 4;
constexpr int kRingDepth = 1 << kRingDepthLog;   // capacity of the hand-shake arrays
// rounds the consumer warps may drift apart, per kernel (measured, power-law n = 1e7): the
// register-staged kernel 0.92 ms at 8 and 0.88 ms at 16; the cp.async kernel 0.73 ms at 8 and
// 1.14 ms at 16 (its rounds are shorter: a deep drift truncates the leading rounds' windows)
constexpr int kRing1DepthLog = 4;
constexpr int kRing2DepthLog = 3;
struct RingHdr {
  int lo[kRingDepth], hi[kRingDepth];   // x entries [lo, hi) valid in the ring for round t (slot t % depth)
  int slot[kRingDepth];                 // ring slot of entry lo
};

struct RingTile {
  int64_t r0, r1, k0, k1;   // rows and entries of a tile (r0 >= r1: nothing to do)
};

// The producer thread of the ring kernels: keeps the shared-memory ring of x up to kRingDepth
// rounds ahead of the consumer warps (see spmv_ring_kernel).
template <typename XT, int DLOG>
__device__ __forceinline__ void ring_producer(const SpmvArgs& a, XT* ring, RingHdr* hdr,
                                              unsigned long long* ring_full, unsigned long long* round_done,
                                              int b0, int b1, int ncw, int rounds, int W, int nloc) {
  constexpr int DEPTH = 1 << DLOG;
  constexpr int DM = DEPTH - 1;
  const int64_t* __restrict__ rowblk = a.rowblk;
  const XT* __restrict__ x = static_cast<const XT*>(a.x);
  constexpr int XA = 16 / (int)sizeof(XT);                 // x entries per 16 bytes
  const int nloc_al = (nloc + XA - 1) & ~(XA - 1);          // columns are padded to 16 entries
  const int H = a.win_half;
  int cur_lo = 0, whi = 0;   // lower bound of the previous round's window; entries loaded so far
  int hi_slot = 0, lo_slot = 0;   // ring slots of entries whi and cur_lo
  for (int t = 0; t < rounds; ++t) {
    // the consumer warps may be anywhere in rounds (t - depth, t): round t's copies may be
    // issued once every warp has left round t - depth, and overwrite only slots below the
    // lower bound of round t - depth + 1
    if (t >= DEPTH) mbar_wait(round_done + (t & DM), (unsigned)((t - DEPTH) >> DLOG) & 1u);
    const int bt = b0 + t * ncw;
    const int bt1 = bt + ncw < b1 ? bt + ncw : b1;
    const int64_t R0 = rowblk[bt], R1 = rowblk[bt1];
    int64_t want_lo = R0 - H;
    if (want_lo < 0) want_lo = 0;
    if (want_lo > nloc_al) want_lo = nloc_al;
    want_lo &= ~(int64_t)(XA - 1);
    int64_t want_hi = R1 + H;
    if (want_hi > nloc_al) want_hi = nloc_al;
    want_hi = (want_hi + XA - 1) & ~(int64_t)(XA - 1);
    int new_hi;
    if (t == 0) {
      whi = cur_lo = (int)want_lo;
      hi_slot = lo_slot = 0;
      new_hi = (int)(want_hi < want_lo + W ? want_hi : want_lo + W);
    } else {
      const int oldest = t - DEPTH + 1 > 0 ? t - DEPTH + 1 : 0;
      const int64_t cap = (int64_t)hdr->lo[oldest & DM] + W;   // that round still reads from its lo on
      const int64_t h = want_hi < cap ? want_hi : cap;
      new_hi = h > whi ? (int)h : whi;
    }
    int next_lo = (int)want_lo;
    if (next_lo < new_hi - W) next_lo = new_hi - W;
    if (next_lo > new_hi) next_lo = new_hi;
    if (next_lo < cur_lo) next_lo = cur_lo;
    lo_slot += next_lo - cur_lo;   // < 2 W: one conditional subtraction
    if (lo_slot >= W) lo_slot -= W;
    hdr->lo[t & DM] = next_lo;
    hdr->hi[t & DM] = new_hi;
    hdr->slot[t & DM] = lo_slot;
    const int cnt = new_hi - whi;   // multiple of XA, <= W
    mbar_expect_tx(ring_full + (t & DM), (unsigned)cnt * (unsigned)sizeof(XT));  // release: hdr visible
    if (cnt > 0) {
      const int first = cnt < W - hi_slot ? cnt : W - hi_slot;
      bulk_g2s(ring + hi_slot, x + whi, (unsigned)first * (unsigned)sizeof(XT), ring_full + (t & DM));
      if (cnt > first)
        bulk_g2s(ring, x + whi + first, (unsigned)(cnt - first) * (unsigned)sizeof(XT), ring_full + (t & DM));
      hi_slot += cnt;
      if (hi_slot >= W) hi_slot -= W;
    }
    whi = new_hi;
    cur_lo = next_lo;
  }
}

template <typename IdxT, typename ValT, typename XT, int MAXW>
__global__ void __launch_bounds__((MAXW + 1) * 32, 1) spmv_ring_kernel(SpmvArgs a) {
  if (a.ctl != nullptr && a.ctl->stop) return;
  extern __shared__ __align__(128) unsigned char ring_smem[];
  const int ncw = (int)(blockDim.x >> 5) - 1;   // consumer warps; the last warp is the producer
  const int W = a.win_cap;                       // ring capacity in x entries (multiple of 16)
  XT* ring = reinterpret_cast<XT*>(ring_smem);
  XT* prod_all = ring + W;
  unsigned char* tail = reinterpret_cast<unsigned char*>(prod_all + (size_t)ncw * kRingStrip);
  RingHdr* hdr = reinterpret_cast<RingHdr*>(tail);
  unsigned long long* ring_full = reinterpret_cast<unsigned long long*>(tail + sizeof(RingHdr));  // [depth]
  unsigned long long* round_done = ring_full + kRingDepth;                                         // [depth]
  constexpr int DLOG = kRing1DepthLog;
  constexpr int DM = (1 << DLOG) - 1;

  // contiguous share of the tiles for this block
  const int per = (a.nblocks + (int)gridDim.x - 1) / (int)gridDim.x;
  const int b0 = (int)blockIdx.x * per;
  const int b1 = b0 + per < a.nblocks ? b0 + per : a.nblocks;
  if (b0 >= b1) return;
  const int rounds = (b1 - b0 + ncw - 1) / ncw;

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i <= DM; ++i) {
      mbar_init(ring_full + i, 1);
      mbar_init(round_done + i, ncw);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int64_t* __restrict__ rowblk = a.rowblk;
  const int64_t* __restrict__ nnzblk = a.rowblk + a.nblocks + 1;
  const XT* __restrict__ x = static_cast<const XT*>(a.x);
  const int nloc = a.n_local_cols > 0x7fffffff ? 0x7fffffff : (int)a.n_local_cols;
  const int warp = tid >> 5, lane = tid & 31;

  if (warp == ncw) {
    // ---------------- producer (lane 0): keeps the ring ahead of the consumers
    if (lane == 0) ring_producer<XT, DLOG>(a, ring, hdr, ring_full, round_done, b0, b1, ncw, rounds, W, nloc);
    return;
  }

  // ---------------- consumer warps
  const IdxT* __restrict__ indptr = static_cast<const IdxT*>(a.indptr);
  const ValT* __restrict__ values = static_cast<const ValT*>(a.values);
  const int32_t* __restrict__ indices = a.indices;
  const XT* __restrict__ ghost = static_cast<const XT*>(a.ghost);
  XT* __restrict__ yout = static_cast<XT*>(a.y);
  XT* prod = prod_all + (size_t)warp * kRingStrip;
  const double xs = a.xscale ? *a.xscale : 1.0;
  const unsigned long long l2pol = l2_policy_evict_first();  // CSR streams must not evict x from L2
  constexpr int VQ = (int)sizeof(ValT) / 4;   // 16-byte loads that hold 4 values

  auto load_tile = [&](int tt) {
    RingTile T = {0, 0, 0, 0};
    const int b = b0 + tt * ncw + warp;
    if (tt < rounds && b < b1) {
      T.r0 = rowblk[b], T.r1 = rowblk[b + 1];
      T.k0 = nnzblk[b], T.k1 = nnzblk[b + 1];
    }
    return T;
  };

  // registers of the strip in flight: 4 column ids + 4 values per lane and 128-entry group, and
  // (first strip of a tile) the tile's first kRingRowGroups x 32 row pointers, one per lane
  int4 cc[kRingIters];
  double2 vv[kRingIters][VQ];
  IdxT rp[kRingRowGroups], nrp[kRingRowGroups];
  auto issue_loads = [&](int64_t ca, int64_t k1) {
    const int tot = (int)((ca + kRingStrip < k1 ? ca + kRingStrip : k1) - ca);
#pragma unroll
    for (int it = 0; it < kRingIters; ++it) {
      const int g = it * kWarp + lane;
      if (4 * g < tot) {
        cc[it] = ld_stream_ef(reinterpret_cast<const int4*>(indices + ca) + g, l2pol);
#pragma unroll
        for (int q = 0; q < VQ; ++q)
          vv[it][q] = ld_stream_ef(reinterpret_cast<const double2*>(values + ca) + (size_t)g * VQ + q, l2pol);
      }
    }
  };
  auto issue_rowptr = [&](const RingTile& T) {
#pragma unroll
    for (int q = 0; q < kRingRowGroups; ++q) {
      int64_t row = T.r0 + q * kWarp + lane;
      if (row > T.r1) row = T.r1;
      nrp[q] = indptr[row];
    }
  };
  int t = 0;
  RingTile cur = load_tile(0), nxt = load_tile(1);
  // rounds in which this warp has no rows still take part in the hand-shake, in order
  auto skip_idle = [&]() {
    while (t < rounds && cur.r0 >= cur.r1) {
      mbar_wait(ring_full + (t & DM), (unsigned)(t >> DLOG) & 1u);
      __syncwarp();
      if (lane == 0) mbar_arrive(round_done + (t & DM));
      ++t;
      cur = nxt;
      nxt = load_tile(t + 1);
    }
  };
  skip_idle();
  int64_t ca = cur.k0 & ~(int64_t)3, cs = cur.k0;
  if (t < rounds) {
    issue_loads(ca, cur.k1);
    issue_rowptr(cur);
  }
  int wlo = 0, wslot = 0;
  unsigned wlen = 0;
  while (t < rounds) {
    // strips of kRingStrip entries starting at the 4-entry-aligned address at or below k0 (the few
    // leading entries belong to the previous tile: computed, never read back)
    const int64_t ce = ca + kRingStrip < cur.k1 ? ca + kRingStrip : cur.k1;
    const int tot = (int)(ce - ca);
    const bool first = (cs == cur.k0);
    if (first) {
      mbar_wait(ring_full + (t & DM), (unsigned)(t >> DLOG) & 1u);
      wlo = hdr->lo[t & DM];
      wslot = hdr->slot[t & DM];
      int whi = hdr->hi[t & DM];
      if (whi > nloc) whi = nloc;   // the padding entry after an odd-length column is not x
      wlen = whi > wlo ? (unsigned)(whi - wlo) : 0u;
#pragma unroll
      for (int q = 0; q < kRingRowGroups; ++q) rp[q] = nrp[q];
    }
    // ---- products, lanes along the entries.  Entries outside the ring (the far tail, ghost
    //      columns) are gathered from global memory; so that nothing branches and the four
    //      gathers of a group are in flight together, EVERY entry issues one: in-ring entries
    //      read the same dummy address (one broadcast sector per request, an L1 hit).  Measured
    //      against two alternatives on the power-law operator (n = 1e7, profiles/r02_spmv_ring.md):
    //      listing the out-of-ring entries in shared memory and gathering them one per lane
    //      (+17%), and one predicated gather per group of four issued ahead for all groups (+11%)
#pragma unroll
    for (int it = 0; it < kRingIters; ++it) {
      const int g = it * kWarp + lane;
      if (4 * g < tot) {
        const int col[4] = {cc[it].x, cc[it].y, cc[it].z, cc[it].w};
        XT xg[4];
        unsigned off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          off[u] = (unsigned)(col[u] - wlo);
          const XT* src = (col[u] < nloc) ? x + col[u] : ghost + (col[u] - nloc);
          xg[u] = ld_ro(off[u] < wlen ? x : src);
        }
        XT pr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          XT xv = xg[u];
          if (off[u] < wlen) {
            int sl = (int)off[u] + wslot;
            if (sl >= W) sl -= W;
            xv = ring[sl];
          }
          if constexpr (sizeof(ValT) == 8) {
            const double2 two = vv[it][u >> 1];
            pr[u] = vmul((u & 1) ? two.y : two.x, xv);
          } else {
            pr[u] = vmul(vv[it][u], xv);
          }
        }
        if constexpr (sizeof(XT) == 8) {
          double2* dst = reinterpret_cast<double2*>(prod + 4 * g);
          dst[0] = make_double2(pr[0], pr[1]);
          dst[1] = make_double2(pr[2], pr[3]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) prod[4 * g + u] = pr[u];
        }
      }
    }
    // ---- the next strip's entries go in flight before this strip's rows are summed
    const bool more = ca + kRingStrip < cur.k1;
    bool prefetched = false;
    if (more) {
      issue_loads(ca + kRingStrip, cur.k1);
      prefetched = true;
    } else if (t + 1 < rounds && nxt.r0 < nxt.r1) {
      issue_loads(nxt.k0 & ~(int64_t)3, nxt.k1);
      issue_rowptr(nxt);
      prefetched = true;
    }
    __syncwarp();
    // ---- rows, lanes along the rows, stored order
    int grp = 0;
    for (int64_t rb = cur.r0; rb < cur.r1; rb += kWarp, ++grp) {
      const int64_t row = rb + lane;
      const bool act = row < cur.r1;
      int64_t rs, re;
      if (grp < kRingRowGroups - 1) {
        // row pointers fetched ahead: lane l holds indptr[rb + l]; its row ends where the next begins
        IdxT mine = rp[0], next0 = rp[1];
#pragma unroll
        for (int q = 1; q < kRingRowGroups - 1; ++q)
          if (grp == q) mine = rp[q], next0 = rp[q + 1];
        IdxT up = __shfl_down_sync(0xffffffffu, mine, 1);
        const IdxT wrap = __shfl_sync(0xffffffffu, next0, 0);
        if (lane == kWarp - 1) up = wrap;
        rs = act ? (int64_t)mine : 0;
        re = act ? (int64_t)up : 0;
      } else {
        rs = act ? (int64_t)indptr[row] : 0;
        re = act ? (int64_t)indptr[row + 1] : 0;
      }
      const int64_t lo = rs > cs ? rs : cs;
      const int64_t hi = re < ce ? re : ce;
      const bool has = act && (hi > lo || (rs == re && first));
      const int kb = (int)(lo - ca), ke = (int)(hi - ca);
      const bool lng = has && (ke - kb > kRingLongSeg);
      XT acc = xzero<XT>();
      if (has && rs < cs) acc = yout[row];
      if (has && !lng)
        for (int k = kb; k < ke; ++k) acc = cadd_rn(acc, prod[k]);
      unsigned m = __ballot_sync(0xffffffffu, lng);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int sb = __shfl_sync(0xffffffffu, kb, src);
        const int se = __shfl_sync(0xffffffffu, ke, src);
        XT part = xzero<XT>();
        for (int k = sb + lane; k < se; k += kWarp) part = cadd_rn(part, prod[k]);
        part = warp_sum(part);
        if (lane == src) acc = cadd_rn(acc, part);
      }
      if (has) yout[row] = (re <= ce) ? cscale(acc, xs) : acc;
    }
    __syncwarp();   // the strip of products may be overwritten
    if (more) {
      ca += kRingStrip;
      cs = ca;
    } else {
      if (lane == 0) mbar_arrive(round_done + (t & DM));
      ++t;
      cur = nxt;
      nxt = load_tile(t + 1);
      if (!prefetched) {
        skip_idle();
        if (t < rounds) {
          issue_loads(cur.k0 & ~(int64_t)3, cur.k1);
          issue_rowptr(cur);
        }
      }
      ca = cur.k0 & ~(int64_t)3;
      cs = cur.k0;
    }
  }
}

// ------------------------------------------------------------------ ring kernel, strips staged with cp.async
// Same decomposition and arithmetic as spmv_ring_kernel (float64 values and vectors); what
// differs is where a strip of the CSR stream waits.  There the 4 column ids + 4 values per lane
// and 128-entry group sit in REGISTERS while in flight (60 of the 128 registers a thread may
// use with 16 warps per SM), and ncu shows the kernel latency-bound at 4 warps per scheduler with
// every unit under 40 % (profiles/r02_spmv_ring.md).  Here a warp stages its strip into its own
// shared-memory buffer with cp.async (LDGSTS: no register holds data in flight), forms the
// products in place (a product overwrites its value) and sums the rows from there; a thread
// needs < 64 registers, so up to 31 consumer warps + the producer fit one SM.  The strip is
// shorter (kRing2Strip entries: the buffers share the SM with the ring) and a warp's next strip
// is requested only when its rows are summed -- the other 30 warps cover that round trip.
constexpr int kRing2Iters = 3;
constexpr int kRing2Strip = 128 * kRing2Iters;   // 384 entries: 4.5 KB per warp
constexpr int kRing2RowGroups = 3;

__device__ __forceinline__ void cpa16_hint(void* dst, const void* src, unsigned long long pol) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "l"(pol)
               : "memory");
}

template <typename IdxT>
__global__ void __launch_bounds__(1024, 1) spmv_ring2_kernel(SpmvArgs a) {
  if (a.ctl != nullptr && a.ctl->stop) return;
  extern __shared__ __align__(128) unsigned char ring_smem[];
  const int ncw = (int)(blockDim.x >> 5) - 1;   // consumer warps; the last warp is the producer
  const int W = a.win_cap;
  double* ring = reinterpret_cast<double*>(ring_smem);
  unsigned char* strips = reinterpret_cast<unsigned char*>(ring + W);
  constexpr size_t kStripBytes = (size_t)kRing2Strip * 12;   // values (8 B) then column ids (4 B)
  unsigned char* tail = strips + (size_t)ncw * kStripBytes;
  RingHdr* hdr = reinterpret_cast<RingHdr*>(tail);
  unsigned long long* ring_full = reinterpret_cast<unsigned long long*>(tail + sizeof(RingHdr));  // [depth]
  unsigned long long* round_done = ring_full + kRingDepth;                                         // [depth]
  constexpr int DLOG = kRing2DepthLog;
  constexpr int DM = (1 << DLOG) - 1;

  const int per = (a.nblocks + (int)gridDim.x - 1) / (int)gridDim.x;
  const int b0 = (int)blockIdx.x * per;
  const int b1 = b0 + per < a.nblocks ? b0 + per : a.nblocks;
  if (b0 >= b1) return;
  const int rounds = (b1 - b0 + ncw - 1) / ncw;

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i <= DM; ++i) {
      mbar_init(ring_full + i, 1);
      mbar_init(round_done + i, ncw);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int64_t* __restrict__ rowblk = a.rowblk;
  const int64_t* __restrict__ nnzblk = a.rowblk + a.nblocks + 1;
  const double* __restrict__ x = static_cast<const double*>(a.x);
  const int nloc = a.n_local_cols > 0x7fffffff ? 0x7fffffff : (int)a.n_local_cols;
  const int warp = tid >> 5, lane = tid & 31;

  if (warp == ncw) {
    if (lane == 0) ring_producer<double, DLOG>(a, ring, hdr, ring_full, round_done, b0, b1, ncw, rounds, W, nloc);
    return;
  }

  // ---------------- consumer warps
  const IdxT* __restrict__ indptr = static_cast<const IdxT*>(a.indptr);
  const double* __restrict__ values = static_cast<const double*>(a.values);
  const int32_t* __restrict__ indices = a.indices;
  const double* __restrict__ ghost = static_cast<const double*>(a.ghost);
  double* __restrict__ yout = static_cast<double*>(a.y);
  double* sval = reinterpret_cast<double*>(strips + (size_t)warp * kStripBytes);   // values, then products
  int32_t* scol = reinterpret_cast<int32_t*>(sval + kRing2Strip);
  const double xs = a.xscale ? *a.xscale : 1.0;
  const unsigned long long l2pol = l2_policy_evict_first();

  auto load_tile = [&](int tt) {
    RingTile T = {0, 0, 0, 0};
    const int b = b0 + tt * ncw + warp;
    if (tt < rounds && b < b1) {
      T.r0 = rowblk[b], T.r1 = rowblk[b + 1];
      T.k0 = nnzblk[b], T.k1 = nnzblk[b + 1];
    }
    return T;
  };
  // request the strip that starts at the 4-aligned entry ca (cp.async, 16 bytes per request)
  auto request = [&](int64_t ca, int64_t k1) {
    const int tot = (int)((ca + kRing2Strip < k1 ? ca + kRing2Strip : k1) - ca);
    const int ngrp = (tot + 3) >> 2;   // groups of 4 entries; the arrays carry 8 spare entries
    for (int g = lane; g < ngrp; g += kWarp) {
      cpa16_hint(scol + 4 * g, indices + ca + 4 * g, l2pol);
      cpa16_hint(sval + 4 * g, values + ca + 4 * g, l2pol);
      cpa16_hint(sval + 4 * g + 2, values + ca + 4 * g + 2, l2pol);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  IdxT rp[kRing2RowGroups];
  auto request_rowptr = [&](const RingTile& T) {
#pragma unroll
    for (int q = 0; q < kRing2RowGroups; ++q) {
      int64_t row = T.r0 + q * kWarp + lane;
      if (row > T.r1) row = T.r1;
      rp[q] = indptr[row];
    }
  };

  int t = 0;
  RingTile cur = load_tile(0), nxt = load_tile(1);
  auto skip_idle = [&]() {
    while (t < rounds && cur.r0 >= cur.r1) {
      mbar_wait(ring_full + (t & DM), (unsigned)(t >> DLOG) & 1u);
      __syncwarp();
      if (lane == 0) mbar_arrive(round_done + (t & DM));
      ++t;
      cur = nxt;
      nxt = load_tile(t + 1);
    }
  };
  skip_idle();
  int64_t ca = cur.k0 & ~(int64_t)3, cs = cur.k0;
  if (t < rounds) {
    request(ca, cur.k1);
    request_rowptr(cur);
  }
  int wlo = 0, wslot = 0;
  unsigned wlen = 0;
  while (t < rounds) {
    const int64_t ce = ca + kRing2Strip < cur.k1 ? ca + kRing2Strip : cur.k1;
    const int tot = (int)(ce - ca);
    const bool first = (cs == cur.k0);
    if (first) {
      mbar_wait(ring_full + (t & DM), (unsigned)(t >> DLOG) & 1u);
      wlo = hdr->lo[t & DM];
      wslot = hdr->slot[t & DM];
      int whi = hdr->hi[t & DM];
      if (whi > nloc) whi = nloc;
      wlen = whi > wlo ? (unsigned)(whi - wlo) : 0u;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- products in place, lanes along the entries (x from the ring; every entry also issues
    //      one global gather, at a dummy address when the ring has it -- see spmv_ring_kernel)
#pragma unroll
    for (int it = 0; it < kRing2Iters; ++it) {
      const int g = it * kWarp + lane;
      if (4 * g < tot) {
        const int4 c4 = *reinterpret_cast<const int4*>(scol + 4 * g);
        const double2 va = *reinterpret_cast<const double2*>(sval + 4 * g);
        const double2 vb = *reinterpret_cast<const double2*>(sval + 4 * g + 2);
        const int col[4] = {c4.x, c4.y, c4.z, c4.w};
        const double av[4] = {va.x, va.y, vb.x, vb.y};
        double xg[4];
        unsigned off[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          off[u] = (unsigned)(col[u] - wlo);
          const double* src = (col[u] < nloc) ? x + col[u] : ghost + (col[u] - nloc);
          xg[u] = ld_ro(off[u] < wlen ? x : src);
        }
        double pr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          double xv = xg[u];
          if (off[u] < wlen) {
            int sl = (int)off[u] + wslot;
            if (sl >= W) sl -= W;
            xv = ring[sl];
          }
          pr[u] = vmul(av[u], xv);
        }
        *reinterpret_cast<double2*>(sval + 4 * g) = make_double2(pr[0], pr[1]);
        *reinterpret_cast<double2*>(sval + 4 * g + 2) = make_double2(pr[2], pr[3]);
      }
    }
    __syncwarp();
    // ---- rows, lanes along the rows, stored order
    const double* prod = sval;
    int grp = 0;
    for (int64_t rb = cur.r0; rb < cur.r1; rb += kWarp, ++grp) {
      const int64_t row = rb + lane;
      const bool act = row < cur.r1;
      int64_t rs, re;
      if (grp < kRing2RowGroups - 1) {
        IdxT mine = rp[0], next0 = rp[1];
#pragma unroll
        for (int q = 1; q < kRing2RowGroups - 1; ++q)
          if (grp == q) mine = rp[q], next0 = rp[q + 1];
        IdxT up = __shfl_down_sync(0xffffffffu, mine, 1);
        const IdxT wrap = __shfl_sync(0xffffffffu, next0, 0);
        if (lane == kWarp - 1) up = wrap;
        rs = act ? (int64_t)mine : 0;
        re = act ? (int64_t)up : 0;
      } else {
        rs = act ? (int64_t)indptr[row] : 0;
        re = act ? (int64_t)indptr[row + 1] : 0;
      }
      const int64_t lo = rs > cs ? rs : cs;
      const int64_t hi = re < ce ? re : ce;
      const bool has = act && (hi > lo || (rs == re && first));
      const int kb = (int)(lo - ca), ke = (int)(hi - ca);
      const bool lng = has && (ke - kb > kRingLongSeg);
      double acc = 0.0;
      if (has && rs < cs) acc = yout[row];
      if (has && !lng)
        for (int k = kb; k < ke; ++k) acc = cadd_rn(acc, prod[k]);
      unsigned m = __ballot_sync(0xffffffffu, lng);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int sb = __shfl_sync(0xffffffffu, kb, src);
        const int se = __shfl_sync(0xffffffffu, ke, src);
        double part = 0.0;
        for (int k = sb + lane; k < se; k += kWarp) part = cadd_rn(part, prod[k]);
        part = warp_sum(part);
        if (lane == src) acc = cadd_rn(acc, part);
      }
      if (has) yout[row] = (re <= ce) ? cscale(acc, xs) : acc;
    }
    __syncwarp();   // the strip buffer may be refilled
    if (ca + kRing2Strip < cur.k1) {
      ca += kRing2Strip;
      cs = ca;
      request(ca, cur.k1);
    } else {
      if (lane == 0) mbar_arrive(round_done + (t & DM));
      ++t;
      cur = nxt;
      nxt = load_tile(t + 1);
      skip_idle();
      ca = cur.k0 & ~(int64_t)3;
      cs = cur.k0;
      if (t < rounds) {
        request(ca, cur.k1);
        request_rowptr(cur);
      }
    }
  }
}

cudaError_t launch_spmv_maxrow(const void* indptr, int indptr_bits, int64_t n, int* out,
                               cudaStream_t st) {
  int64_t grid = (n + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  if (indptr_bits == 32)
    spmv_maxrow_kernel<int32_t><<<(int)grid, 256, 0, st>>>(static_cast<const int32_t*>(indptr), n, out);
  else
    spmv_maxrow_kernel<int64_t><<<(int)grid, 256, 0, st>>>(static_cast<const int64_t*>(indptr), n, out);
  return cudaGetLastError();
}

template <typename IdxT>
__global__ void spmv_locality_kernel(const IdxT* __restrict__ indptr, const int32_t* __restrict__ indices,
                                     int64_t n, int64_t nloc, int64_t half,
                                     unsigned long long* __restrict__ out2) {
  unsigned long long in = 0, tot = 0;
  // every 16th row, a warp per sampled row
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = wid * 16; r < n; r += nw * 16) {
    const int64_t rs = (int64_t)indptr[r], re = (int64_t)indptr[r + 1];
    for (int64_t k = rs + lane; k < re; k += 32) {
      const int64_t c = indices[k];
      const int64_t d = c - r;
      tot += 1;
      if (c < nloc && d >= -half && d <= half) in += 1;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    in += __shfl_xor_sync(0xffffffffu, in, o);
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
  }
  if (lane == 0 && tot) {  // integer sums: order-independent
    atomicAdd(out2, in);
    atomicAdd(out2 + 1, tot);
  }
}
cudaError_t launch_spmv_locality(const void* indptr, int indptr_bits, const int32_t* indices, int64_t n,
                                 int64_t n_local_cols, int64_t half, unsigned long long* out2,
                                 cudaStream_t st) {
  if (indptr_bits == 32)
    spmv_locality_kernel<int32_t><<<148 * 4, 256, 0, st>>>(static_cast<const int32_t*>(indptr), indices, n,
                                                           n_local_cols, half, out2);
  else
    spmv_locality_kernel<int64_t><<<148 * 4, 256, 0, st>>>(static_cast<const int64_t*>(indptr), indices, n,
                                                           n_local_cols, half, out2);
  return cudaGetLastError();
}

cudaError_t launch_spmv_plan(const void* indptr, int indptr_bits, int64_t n, int64_t nnz, int tile,
                             int nblocks, int64_t* rowblk, cudaStream_t st) {
  const int threads = 256;
  const int grid = (nblocks + 1 + threads - 1) / threads;
  if (indptr_bits == 32)
    spmv_plan_kernel<int32_t><<<grid, threads, 0, st>>>(static_cast<const int32_t*>(indptr), n, nnz,
                                                        tile, nblocks, rowblk);
  else
    spmv_plan_kernel<int64_t><<<grid, threads, 0, st>>>(static_cast<const int64_t*>(indptr), n, nnz,
                                                        tile, nblocks, rowblk);
  return cudaGetLastError();
}

template <typename IdxT, typename ValT, typename XT, int THREADS, bool LONG>
static cudaError_t launch_spmv_ttl(const SpmvArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)(a.tile + 4) * (sizeof(ValT) + sizeof(int32_t)) + 16 +
                      sizeof(int64_t) * (size_t)(a.tile / 16 + 2);
  static PerDeviceOnce once;
  if (once.first_use()) {
    cudaFuncSetAttribute(spmv_tile_kernel<IdxT, ValT, XT, THREADS, LONG>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  int grid = a.nblocks;  // one tile per block; blocks are small and many per SM
  if (a.contig) {
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, spmv_tile_kernel<IdxT, ValT, XT, THREADS, LONG>, THREADS, smem) != cudaSuccess || occ < 1)
      occ = 1;
    if (a.bps > 0 && a.bps < occ) occ = a.bps;
    const int64_t g = (int64_t)a.num_sms * occ;
    if (g < grid) grid = (int)g;
  }
  spmv_tile_kernel<IdxT, ValT, XT, THREADS, LONG><<<grid, THREADS, smem, st>>>(a);
  return cudaGetLastError();
}
template <typename IdxT, typename ValT, typename XT, int THREADS>
static cudaError_t launch_spmv_stream(const SpmvArgs& a, cudaStream_t st) {
  const int cap = a.tile + kShortRow + 8;
  const size_t stage_bytes = ((size_t)cap * (sizeof(ValT) + sizeof(int32_t)) +
                              (size_t)(2 * THREADS + 2) * sizeof(IdxT) + 15) / 16 * 16;
  const size_t smem = 2 * stage_bytes;
  static int occ = 0;
  static PerDeviceOnce once;
  if (once.first_use()) {
    cudaFuncSetAttribute(spmv_stream_kernel<IdxT, ValT, XT, THREADS>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, spmv_stream_kernel<IdxT, ValT, XT, THREADS>, THREADS, smem) != cudaSuccess ||
        occ < 1)
      occ = 1;
  }
  int64_t grid = (int64_t)a.num_sms * occ;
  if (grid > a.nblocks) grid = a.nblocks;
  spmv_stream_kernel<IdxT, ValT, XT, THREADS><<<(int)grid, THREADS, smem, st>>>(a);
  return cudaGetLastError();
}
size_t spmv_bulk_stage_bytes(int tile, int rp_cap, int val_bytes, int idx_bytes) {
  const size_t cap = (size_t)tile + kShortRow + 8;
  return (cap * (val_bytes + 4) + (size_t)rp_cap * idx_bytes + sizeof(TileHdr) + 127) / 128 * 128;
}

template <typename IdxT, typename ValT, typename XT, bool HALO>
static cudaError_t launch_spmv_bulk_h(const SpmvArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)a.stages * spmv_bulk_stage_bytes(a.tile, a.rp_cap, sizeof(ValT), sizeof(IdxT)) +
                      2 * (size_t)a.stages * sizeof(unsigned long long);
  const int threads = a.threads + kWarp;  // consumers + the producer warp
  static int occ[2] = {0, 0};
  static size_t occ_smem[2] = {0, 0};
  static PerDeviceOnce once;
  if (once.first_use()) {
    cudaFuncSetAttribute(spmv_bulk_kernel<IdxT, ValT, XT, HALO>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  }
  int* o = &occ[a.threads == 128 ? 0 : 1];
  if (*o == 0 || occ_smem[a.threads == 128 ? 0 : 1] != smem) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(o, spmv_bulk_kernel<IdxT, ValT, XT, HALO>,
                                                      threads, smem) != cudaSuccess || *o < 1)
      *o = 1;
    occ_smem[a.threads == 128 ? 0 : 1] = smem;
  }
  int bps = *o;
  if (a.bps > 0 && a.bps < bps) bps = a.bps;
  int64_t grid = (int64_t)a.num_sms * bps;
  if (grid > a.nblocks) grid = a.nblocks;
  if (grid < 1) grid = 1;
  spmv_bulk_kernel<IdxT, ValT, XT, HALO><<<(int)grid, threads, smem, st>>>(a);
  return cudaGetLastError();
}
template <typename IdxT, typename ValT, typename XT>
static cudaError_t launch_spmv_bulk(const SpmvArgs& a, cudaStream_t st) {
  const bool halo = a.ghost != nullptr || a.direct_halo;
  return halo ? launch_spmv_bulk_h<IdxT, ValT, XT, true>(a, st)
              : launch_spmv_bulk_h<IdxT, ValT, XT, false>(a, st);
}

size_t spmv_ring_smem(int win_cap, int nwarps, int x_bytes) {
  return (size_t)win_cap * x_bytes + (size_t)nwarps * kRingStrip * x_bytes + sizeof(RingHdr) +
         2 * kRingDepth * sizeof(unsigned long long);
}

template <typename IdxT, typename ValT, typename XT>
static cudaError_t launch_spmv_ring(const SpmvArgs& a, cudaStream_t st) {
  constexpr int MAXW = sizeof(XT) == 8 ? 15 : 7;   // + the producer warp = 16 / 8 warps; 16-byte products: half the warps fit
  int nwarps = a.ring_warps;
  if (nwarps > MAXW) nwarps = MAXW;
  if (nwarps < 1) nwarps = 1;
  // the ring takes what the strips of products leave of the 227 KB, up to the capacity asked for
  int wcap = a.win_cap;
  const int fit = (int)((227 * 1024 - spmv_ring_smem(0, nwarps, sizeof(XT))) / sizeof(XT)) / 256 * 256;
  if (wcap > fit) wcap = fit;
  wcap = wcap / 16 * 16;
  if (wcap < 256) wcap = 256;
  SpmvArgs b = a;
  b.win_cap = wcap;
  if (b.win_half > (wcap - 1536) / 2) b.win_half = (wcap - 1536) / 2;   // room for two rounds of rows
  if (b.win_half < 0) b.win_half = 0;
  const size_t smem = spmv_ring_smem(wcap, nwarps, sizeof(XT));
  static PerDeviceOnce once;
  if (once.first_use()) {
    cudaFuncSetAttribute(spmv_ring_kernel<IdxT, ValT, XT, MAXW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         227 * 1024);
  }
  int64_t grid = a.num_sms;   // one block per SM: the ring wants the shared memory
  if (a.bps > 1) grid *= a.bps;   // (A/B only: more, smaller ranges per SM-resident block)
  const int64_t need = (a.nblocks + nwarps - 1) / nwarps;   // no block without a full round of tiles
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  spmv_ring_kernel<IdxT, ValT, XT, MAXW><<<(int)grid, (nwarps + 1) * kWarp, smem, st>>>(b);
  return cudaGetLastError();
}

size_t spmv_ring2_smem(int win_cap, int nwarps) {
  return (size_t)win_cap * 8 + (size_t)nwarps * kRing2Strip * 12 + sizeof(RingHdr) +
         2 * kRingDepth * sizeof(unsigned long long);
}

template <typename IdxT>
static cudaError_t launch_spmv_ring2(const SpmvArgs& a, cudaStream_t st) {
  int nwarps = a.ring_warps;
  if (nwarps > 31) nwarps = 31;
  if (nwarps < 1) nwarps = 1;
  int wcap = a.win_cap;
  const int fit = (int)((227 * 1024 - spmv_ring2_smem(0, nwarps)) / 8) / 256 * 256;
  if (wcap > fit) wcap = fit;
  wcap = wcap / 16 * 16;
  if (wcap < 256) wcap = 256;
  SpmvArgs b = a;
  b.win_cap = wcap;
  if (b.win_half > (wcap - 1536) / 2) b.win_half = (wcap - 1536) / 2;
  if (b.win_half < 0) b.win_half = 0;
  const size_t smem = spmv_ring2_smem(wcap, nwarps);
  static PerDeviceOnce once;
  if (once.first_use()) {
    cudaFuncSetAttribute(spmv_ring2_kernel<IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  int64_t grid = a.num_sms;
  const int64_t need = (a.nblocks + nwarps - 1) / nwarps;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  spmv_ring2_kernel<IdxT><<<(int)grid, (nwarps + 1) * kWarp, smem, st>>>(b);
  return cudaGetLastError();
}

template <typename IdxT, typename ValT, typename XT, int THREADS>
static cudaError_t launch_spmv_tt(const SpmvArgs& a, cudaStream_t st) {
  // the ring kernel serves float64 vectors (real storage); with complex128 vectors half as many
  // entries fit the ring and the strips, and the tile kernel is faster (1.33 vs 2.0 ms, n = 1e7)
  if constexpr (sizeof(XT) == 8) {
    if (a.window == 2) return launch_spmv_ring2<IdxT>(a, st);
    if (a.window) return launch_spmv_ring<IdxT, ValT, XT>(a, st);
  }
  if (!a.long_rows && a.variant == 0) return launch_spmv_bulk<IdxT, ValT, XT>(a, st);
  if (!a.long_rows && a.variant == 2) return launch_spmv_stream<IdxT, ValT, XT, THREADS>(a, st);
  return a.long_rows ? launch_spmv_ttl<IdxT, ValT, XT, THREADS, true>(a, st)
                     : launch_spmv_ttl<IdxT, ValT, XT, THREADS, false>(a, st);
}
template <typename IdxT, typename ValT, typename XT>
static cudaError_t launch_spmv_x(const SpmvArgs& a, cudaStream_t st) {
  return a.threads == 128 ? launch_spmv_tt<IdxT, ValT, XT, 128>(a, st)
                          : launch_spmv_tt<IdxT, ValT, XT, 256>(a, st);
}
template <typename IdxT, typename ValT>
static cudaError_t launch_spmv_t(const SpmvArgs& a, cudaStream_t st) {
  return launch_spmv_x<IdxT, ValT, cplx>(a, st);
}
template <typename IdxT>
static cudaError_t launch_spmv_real(const SpmvArgs& a, cudaStream_t st) {
  return launch_spmv_x<IdxT, double, double>(a, st);
}

cudaError_t launch_spmv(const SpmvArgs& a, int indptr_bits, int value_kind, cudaStream_t st) {
  if (a.real) {
    if (value_kind != 0) return cudaErrorInvalidValue;  // real vectors need a float64 matrix
    return indptr_bits == 32 ? launch_spmv_real<int32_t>(a, st) : launch_spmv_real<int64_t>(a, st);
  }
  if (indptr_bits == 32) {
    return value_kind == 0 ? launch_spmv_t<int32_t, double>(a, st)
                           : launch_spmv_t<int32_t, cplx>(a, st);
  }
  return value_kind == 0 ? launch_spmv_t<int64_t, double>(a, st)
                         : launch_spmv_t<int64_t, cplx>(a, st);
}

}  // namespace ab200
