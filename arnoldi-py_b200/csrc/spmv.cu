// spmv.cu -- CSR sparse matrix x complex128 vector  (the reference's `A @ V[:, j]`,
// decomposition.py:57-58, which lands in scipy's csr_matvec).
//
// Work decomposition ("nnz tiles with row-aligned edges"):
//   * the nnz range is cut into tiles of `tile` entries; tile b starts at the first
//     row whose indptr >= b * tile (rowblk[], built once per matrix on the device),
//     so every block owns whole rows and about the same number of non-zeros no
//     matter how skewed the row lengths are (the merge-path idea, with the split
//     snapped to row boundaries);
//   * a block stages its tile of values + column ids into shared memory with
//     coalesced 128-bit loads, then one thread per row walks its segment in STORED
//     ORDER with separate multiply and add -- the same operation order as scipy's
//     csr_matvec, so for rows handled this way y is bit-identical to scipy;
//   * a row whose part of the tile is longer than kShortRow (16) entries is summed by a whole
//     warp instead (lanes stride the segment, fixed-order butterfly: deterministic, but not
//     scipy's order), and a row longer than a tile is carried across the block's tile
//     iterations through y.  Without this, one thread walking a 100-entry row serialises
//     the block on power-law matrices.
// x is gathered with 128-bit read-only loads; for banded matrices consecutive rows
// gather consecutive x entries, so the gathers coalesce and hit L2.
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

// ------------------------------------------------------------------ skewed rows: x window in shared memory
// For operators whose rows are long and uneven (BASELINE config 4: Pareto row lengths, columns in a
// band around the diagonal plus a sparse far tail) the streaming roofline is not what limits the
// tile kernel: ncu shows 10 sectors per gather request, L1 hit rate 7% -- every 8/16-byte x
// entry costs a 32-byte sector through L2.  Here the block first pulls the WINDOW of x its tile
// can touch around the diagonal into shared memory with one bulk copy (contiguous, so it runs at
// copy speed and neighbouring tiles hit L2), then streams the tile's values / column ids through
// a two-stage bulk-copy ring in sub-tiles; all threads first turn a sub-tile into products
// (coalesced over the entries, x from the window, only out-of-window / ghost entries go to
// global memory), then rows are summed from the products in stored order (one thread per row,
// a warp for segments > 64 entries).  Work per block is balanced in non-zeros (the plan's tiles),
// not in rows.  Rows of <= 16 entries are still summed in scipy's order (bit-identical).
constexpr int kWinSub = 1024;      // entries per sub-tile
constexpr int kWinLongSeg = 64;    // segments longer than this are summed by a warp

struct WinHdr {
  int64_t r0, r1;    // rows of the tile
  int64_t k0, k1;    // its entries
  int64_t wlo;       // first x entry held in the window
  int wlen;          // entries held
  int nrp;           // row pointers staged (from row r0a)
  int rpoff;
  int pad;
};

template <typename IdxT, typename ValT, typename XT>
__global__ void __launch_bounds__(288) spmv_window_kernel(SpmvArgs a) {
  if (a.ctl != nullptr && a.ctl->stop) return;
  extern __shared__ __align__(128) unsigned char win_smem[];
  const int nthr = blockDim.x - kWarp;
  const int ncw = nthr >> 5;
  const int wcap = a.win_cap;
  const int rpc = a.rp_cap;
  // layout: xwin | 2 x (vals | cols) | prod | rptr | long-row list | hdr | barriers
  size_t off = 0;
  XT* xwin = reinterpret_cast<XT*>(win_smem);
  off += ((size_t)wcap * sizeof(XT) + 127) / 128 * 128;
  unsigned char* ring = win_smem + off;
  const size_t ring_stage = (size_t)kWinSub * (sizeof(ValT) + sizeof(int32_t));
  off += 2 * ring_stage;
  XT* prod = reinterpret_cast<XT*>(win_smem + off);
  off += (size_t)kWinSub * sizeof(XT);
  IdxT* srow = reinterpret_cast<IdxT*>(win_smem + off);
  off += ((size_t)rpc * sizeof(IdxT) + 15) / 16 * 16;
  int* s_long = reinterpret_cast<int*>(win_smem + off);   // [kWinSub / kWinLongSeg + 2] local row ids
  off += sizeof(int) * (kWinSub / kWinLongSeg + 4);
  off = (off + 15) / 16 * 16;
  WinHdr* hdr = reinterpret_cast<WinHdr*>(win_smem + off);
  off += sizeof(WinHdr);
  off = (off + 7) / 8 * 8;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(win_smem + off);
  unsigned long long* win_full = bars;        // window + row pointers + header of a tile landed
  unsigned long long* tile_done = bars + 1;   // every consumer warp is done with the tile
  unsigned long long* ring_full = bars + 2;   // [2]
  unsigned long long* ring_empty = bars + 4;  // [2]
  __shared__ int s_nlong;

  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(win_full, 1);
    mbar_init(tile_done, ncw);
    for (int i = 0; i < 2; ++i) {
      mbar_init(ring_full + i, 1);
      mbar_init(ring_empty + i, ncw);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const IdxT* __restrict__ indptr = static_cast<const IdxT*>(a.indptr);
  const ValT* __restrict__ values = static_cast<const ValT*>(a.values);
  const int32_t* __restrict__ indices = a.indices;
  const XT* __restrict__ x = static_cast<const XT*>(a.x);
  const int64_t* __restrict__ rowblk = a.rowblk;
  const int64_t* __restrict__ nnzblk = a.rowblk + a.nblocks + 1;
  // contiguous share of the tiles for this block: neighbouring tiles share most of their window
  const int per = (a.nblocks + (int)gridDim.x - 1) / (int)gridDim.x;
  const int b0 = (int)blockIdx.x * per;
  const int b1 = b0 + per < a.nblocks ? b0 + per : a.nblocks;

  if (tid >= nthr) {
    // ---------------- producer warp (lane 0)
    if (tid != nthr) return;
    constexpr int RA = 16 / (int)sizeof(IdxT);
    constexpr int XA = 16 / (int)sizeof(XT);  // x entries per 16 bytes
    unsigned use0 = 0, use1 = 0;  // times each ring stage has been filled
    const unsigned long long l2pol = l2_policy_evict_first();
    for (int b = b0, t = 0; b < b1; ++b, ++t) {
      if (t > 0) mbar_wait(tile_done, (unsigned)(t - 1) & 1u);
      const int64_t r0 = rowblk[b], r1 = rowblk[b + 1];
      const int64_t k0 = nnzblk[b], k1 = nnzblk[b + 1];
      const int64_t nrows = r1 - r0;
      int64_t half = ((int64_t)wcap - nrows) / 2;
      if (half < 0) half = 0;
      int64_t wlo = r0 - half;
      if (wlo < 0) wlo = 0;
      wlo &= ~(int64_t)(XA - 1);
      int64_t whi = wlo + wcap;
      if (whi > a.n_local_cols) whi = a.n_local_cols;
      int64_t wlen = whi - wlo;
      if (wlen < 0) wlen = 0;
      const int64_t wcopy = (wlen + XA - 1) & ~(int64_t)(XA - 1);  // columns are padded to 16 elements
      const int64_t r0a = r0 & ~(int64_t)(RA - 1);
      int64_t nrp = nrows + 1 + (r0 - r0a);
      if (nrp > rpc) nrp = rpc;
      nrp = (nrp + RA - 1) & ~(int64_t)(RA - 1);
      hdr->r0 = r0;
      hdr->r1 = r1;
      hdr->k0 = k0;
      hdr->k1 = k1;
      hdr->wlo = wlo;
      hdr->wlen = (int)wlen;
      hdr->nrp = (int)nrp;
      hdr->rpoff = (int)(r0 - r0a);
      const unsigned bw = (unsigned)(wcopy * sizeof(XT)), br = (unsigned)(nrp * sizeof(IdxT));
      mbar_expect_tx(win_full, bw + br);
      if (bw) bulk_g2s(xwin, x + wlo, bw, win_full);
      bulk_g2s(srow, indptr + r0a, br, win_full);
      // the tile's entries, sub-tile by sub-tile, from the 4-aligned address at or below k0
      const int64_t ka = k0 & ~(int64_t)3;
      for (int64_t cs = ka; cs < k1; cs += kWinSub) {
        const int st = (int)(((cs - ka) / kWinSub) & 1);
        const unsigned used = st ? use1 : use0;
        if (used > 0) mbar_wait(ring_empty + st, (used - 1) & 1u);
        int64_t cnt = k1 - cs;
        if (cnt > kWinSub) cnt = kWinSub;
        const unsigned nc = (unsigned)((cnt + 3) & ~(int64_t)3);
        unsigned char* base = ring + (size_t)st * ring_stage;
        mbar_expect_tx(ring_full + st, nc * (unsigned)(sizeof(ValT) + 4));
        bulk_g2s_hint(base, values + cs, nc * (unsigned)sizeof(ValT), ring_full + st, l2pol);
        bulk_g2s_hint(base + (size_t)kWinSub * sizeof(ValT), indices + cs, nc * 4u, ring_full + st, l2pol);
        if (st) use1 += 1; else use0 += 1;
      }
    }
    return;
  }

  // ---------------- consumers
  const XT* __restrict__ ghost = static_cast<const XT*>(a.ghost);
  XT* __restrict__ yout = static_cast<XT*>(a.y);
  const int64_t nloc = a.n_local_cols;
  const double xs = a.xscale ? *a.xscale : 1.0;
  const int lane = tid & 31, warp = tid >> 5;
  unsigned use0 = 0, use1 = 0;
  for (int b = b0, t = 0; b < b1; ++b, ++t) {
    mbar_wait(win_full, (unsigned)t & 1u);
    const WinHdr h = *hdr;
    const int64_t ka = h.k0 & ~(int64_t)3;
    const int nrows = (int)(h.r1 - h.r0);
    if (h.k1 == h.k0) {  // only empty rows
      for (int rl = tid; rl < nrows; rl += nthr) yout[h.r0 + rl] = xzero<XT>();
    }
    for (int64_t cs = ka; cs < h.k1; cs += kWinSub) {
      const int st = (int)(((cs - ka) / kWinSub) & 1);
      mbar_wait(ring_full + st, (st ? use1 : use0) & 1u);
      if (st) use1 += 1; else use0 += 1;
      const ValT* sval = reinterpret_cast<const ValT*>(ring + (size_t)st * ring_stage);
      const int32_t* scol =
          reinterpret_cast<const int32_t*>(ring + (size_t)st * ring_stage + (size_t)kWinSub * sizeof(ValT));
      int64_t ce = cs + kWinSub;
      if (ce > h.k1) ce = h.k1;
      const int cnt = (int)(ce - cs);
      // ---- products, coalesced over the entries; the (few) gathers that miss the window are
      //      issued together so a thread waits for global memory once per sub-tile, not per entry
      constexpr int PU = 4;
      for (int k0 = tid; k0 < cnt; k0 += PU * nthr) {
        XT xg[PU];
        int64_t wi[PU];
#pragma unroll
        for (int u = 0; u < PU; ++u) {
          const int k = k0 + u * nthr;
          wi[u] = -1;
          xg[u] = xzero<XT>();
          if (k < cnt) {
            const int64_t col = scol[k];
            wi[u] = col - h.wlo;
            if (wi[u] < 0 || wi[u] >= h.wlen) {
              xg[u] = (col < nloc) ? ld_ro(x + col) : ld_ro(ghost + (col - nloc));
              wi[u] = -1;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < PU; ++u) {
          const int k = k0 + u * nthr;
          if (k < cnt) prod[k] = vmul(sval[k], wi[u] >= 0 ? xwin[wi[u]] : xg[u]);
        }
      }
      if (tid == 0) s_nlong = 0;
      __syncwarp();
      if (lane == 0) mbar_arrive(ring_empty + st);   // values / column ids of this stage are consumed
      asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");   // products visible to all consumers
      // ---- rows: one thread per row over the products, stored order
      const int64_t lo_k = cs > h.k0 ? cs : h.k0;
      for (int rl = tid; rl < nrows; rl += nthr) {
        const int slot = h.rpoff + rl;
        int64_t rs, re;
        if (slot + 1 < h.nrp) {
          rs = (int64_t)srow[slot];
          re = (int64_t)srow[slot + 1];
        } else {
          rs = (int64_t)indptr[h.r0 + rl];
          re = (int64_t)indptr[h.r0 + rl + 1];
        }
        if (re <= lo_k && !(rs == re && cs == ka)) continue;  // finished in an earlier sub-tile
        if (rs >= ce && rs != re) continue;                    // starts in a later one
        const int64_t lo = rs > lo_k ? rs : lo_k;
        const int64_t hi = re < ce ? re : ce;
        if (hi - lo > kWinLongSeg) {
          const int q = atomicAdd(&s_nlong, 1);
          s_long[q] = rl;
          continue;
        }
        XT acc = (rs < lo_k) ? yout[h.r0 + rl] : xzero<XT>();
        for (int k = (int)(lo - cs); k < (int)(hi - cs); ++k) acc = cadd_rn(acc, prod[k]);
        if (re <= ce) acc = cscale(acc, xs);
        yout[h.r0 + rl] = acc;
      }
      asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");   // the long-row list is complete
      const int nlong = s_nlong;
      for (int l = warp; l < nlong; l += ncw) {
        const int rl = s_long[l];
        const int slot = h.rpoff + rl;
        int64_t rs, re;
        if (slot + 1 < h.nrp) {
          rs = (int64_t)srow[slot];
          re = (int64_t)srow[slot + 1];
        } else {
          rs = (int64_t)indptr[h.r0 + rl];
          re = (int64_t)indptr[h.r0 + rl + 1];
        }
        const int64_t lo = rs > lo_k ? rs : lo_k;
        const int64_t hi = re < ce ? re : ce;
        XT acc = xzero<XT>();
        for (int k = (int)(lo - cs) + lane; k < (int)(hi - cs); k += kWarp) acc = cadd_rn(acc, prod[k]);
        acc = warp_sum(acc);
        if (lane == 0) {
          XT tsum = (rs < lo_k) ? cadd_rn(yout[h.r0 + rl], acc) : acc;
          if (re <= ce) tsum = cscale(tsum, xs);
          yout[h.r0 + rl] = tsum;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");   // products / list may be overwritten
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(tile_done);
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
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
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

size_t spmv_window_smem(int win_cap, int rp_cap, int val_bytes, int idx_bytes, int x_bytes) {
  size_t off = ((size_t)win_cap * x_bytes + 127) / 128 * 128;
  off += 2 * (size_t)kWinSub * (val_bytes + 4);
  off += (size_t)kWinSub * x_bytes;
  off += ((size_t)rp_cap * idx_bytes + 15) / 16 * 16;
  off += sizeof(int) * (kWinSub / kWinLongSeg + 4);
  off = (off + 15) / 16 * 16;
  off += sizeof(WinHdr);
  off = (off + 7) / 8 * 8;
  return off + 6 * sizeof(unsigned long long);
}

template <typename IdxT, typename ValT, typename XT>
static cudaError_t launch_spmv_window(const SpmvArgs& a, cudaStream_t st) {
  const size_t smem = spmv_window_smem(a.win_cap, a.rp_cap, sizeof(ValT), sizeof(IdxT), sizeof(XT));
  const int threads = 256 + kWarp;
  static PerDeviceOnce once;
  if (once.first_use()) {
    cudaFuncSetAttribute(spmv_window_kernel<IdxT, ValT, XT>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  }
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spmv_window_kernel<IdxT, ValT, XT>, threads,
                                                    smem) != cudaSuccess || occ < 1)
    occ = 1;
  if (a.bps > 0 && a.bps < occ) occ = a.bps;
  int64_t grid = (int64_t)a.num_sms * occ;
  if (grid > a.nblocks) grid = a.nblocks;
  if (grid < 1) grid = 1;
  spmv_window_kernel<IdxT, ValT, XT><<<(int)grid, threads, smem, st>>>(a);
  return cudaGetLastError();
}

template <typename IdxT, typename ValT, typename XT, int THREADS>
static cudaError_t launch_spmv_tt(const SpmvArgs& a, cudaStream_t st) {
  if (a.window) return launch_spmv_window<IdxT, ValT, XT>(a, st);
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
