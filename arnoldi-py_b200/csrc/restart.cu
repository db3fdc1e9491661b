// restart.cu -- Krylov-Schur truncation of the basis, in place and in one pass
// (krylov_schur.py:78 `V[:, :p] = V_active @ Qp` and :81 `V[:, p] = V[:, m]`).
//
// Row r of the result depends only on row r of the input, so a block reads the
// m inputs of its rows, synchronises, and then overwrites the first p+1 columns of
// the same rows: no second n x p buffer, 16 n (m + p + 2) bytes of traffic.
// Block = 32 rows (lanes, so each column access of a warp is 512 contiguous bytes)
// x PW warps; warp y accumulates outputs [y*PT, y*PT+PT) in registers.  The m x p
// coefficient block (Q with the lazy column scales folded into its rows) sits in
// shared memory and is read as warp-wide broadcasts.
#include "kernels.cuh"

namespace ab200 {

// coherent 128-bit load whose issue point is pinned (volatile): batches of these are
// issued back to back so several columns of a row are in flight per thread
__device__ __forceinline__ cplx ld_pinned(const cplx* p) {
  cplx r;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

template <int PT, bool QS, bool REAL>
__global__ void __launch_bounds__(512) restart_kernel(RestartArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = a.m, p = a.p;
  // coefficients: shared memory when they fit comfortably, else straight from global
  // (uniform addresses: one L1-resident broadcast load per warp)
  cplx* sq_s = reinterpret_cast<cplx*>(smem_raw);  // [m][p]
  if (QS) {
    for (int k = threadIdx.x; k < m * p; k += blockDim.x) sq_s[k] = a.q[k];
    __syncthreads();
  }
  const cplx* __restrict__ sq_g = a.q;

  const int lane = threadIdx.x & 31;
  const int wy = threadIdx.x >> 5;
  const int k0 = wy * PT;
  int nk = p - k0;
  nk = nk < 0 ? 0 : (nk > PT ? PT : nk);
  const int64_t ld = a.ld;
  const int64_t nchunks = (a.n + kWarp - 1) / kWarp;
  constexpr int B = 4;  // columns per load batch; two batches are in flight
  const int mlast = a.copy_tail ? m : m - 1;

  for (int64_t q = blockIdx.x; q < nchunks; q += gridDim.x) {
    const int64_t row = q * kWarp + lane;
    const bool ok = row < a.n;
    const cplx* src = a.U + (ok ? row : 0);  // coherent loads: U is also written here
    cplx acc[PT];
#pragma unroll
    for (int k = 0; k < PT; ++k) acc[k] = make_double2(0.0, 0.0);
    cplx cur[B], nxt[B];
#pragma unroll
    for (int t = 0; t < B; ++t) cur[t] = ld_pinned(src + (int64_t)(t < m ? t : m - 1) * ld);
    for (int i = 0; i < m; i += B) {
#pragma unroll
      for (int t = 0; t < B; ++t) {
        const int col = i + B + t;
        // past the end: column m (read for the tail anyway) or, without a tail, column m - 1 again
        nxt[t] = ld_pinned(src + (int64_t)(col < m ? col : mlast) * ld);
      }
#pragma unroll
      for (int t = 0; t < B; ++t) {
        if (i + t < m) {
          const int qo = (i + t) * p + k0;
#pragma unroll
          for (int k = 0; k < PT; ++k)
            if (k < nk) addax<REAL>(acc[k], cur[t], QS ? sq_s[qo + k] : __ldg(sq_g + qo + k));
        }
      }
#pragma unroll
      for (int t = 0; t < B; ++t) cur[t] = nxt[t];
    }
    cplx tail = make_double2(0.0, 0.0);
    if (wy == 0 && a.copy_tail) tail = cscale(ld_pinned(src + (int64_t)m * ld), a.scale_m);
    if (blockDim.x > kWarp) __syncthreads();  // every warp of the block has read these rows
    if (ok) {
      cplx* dst = a.U + row;
#pragma unroll
      for (int k = 0; k < PT; ++k)
        if (k < nk) st_stream(dst + (int64_t)(k0 + k) * ld, acc[k]);
      if (wy == 0 && a.copy_tail) st_stream(dst + (int64_t)p * ld, tail);
    }
  }
}

template <int PT>
static cudaError_t launch_restart_t(const RestartArgs& a, int num_sms, cudaStream_t st) {
  const int nw = (a.p + PT - 1) / PT;
  const int threads = nw * kWarp;
  size_t smem = sizeof(cplx) * (size_t)a.m * a.p;
  const bool qs = smem <= 100 * 1024;  // stage the coefficients in shared memory
  if (!qs) smem = 0;
  static PerDeviceOnce once;
  if (once.first_use()) {
    cudaFuncSetAttribute(restart_kernel<PT, true, true>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(restart_kernel<PT, true, false>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  }
  const int64_t nchunks = (a.n + kWarp - 1) / kWarp;
  int bps = 2048 / threads;
  if (bps > 16) bps = 16;
  if (bps < 1) bps = 1;
  int64_t grid = (int64_t)num_sms * bps;
  if (grid > nchunks) grid = nchunks;
  if (grid < 1) grid = 1;
  if (qs) {
    if (a.real)
      restart_kernel<PT, true, true><<<(int)grid, threads, smem, st>>>(a);
    else
      restart_kernel<PT, true, false><<<(int)grid, threads, smem, st>>>(a);
  } else {
    if (a.real)
      restart_kernel<PT, false, true><<<(int)grid, threads, smem, st>>>(a);
    else
      restart_kernel<PT, false, false><<<(int)grid, threads, smem, st>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launch_restart(const RestartArgs& a, int num_sms, cudaStream_t st, int variant) {
  // Outputs per warp: the fewest warps per row group (p <= 16 needs no block barrier at all),
  // and the p outputs spread evenly over them (p = 25 -> 13 + 12, not 16 + 9: the block is as
  // slow as its busiest warp and this kernel is FP64-bound)
  int pt;
  if (variant > 0) {
    pt = variant;
  } else {
    const int nw = (a.p + 15) / 16;
    pt = (a.p + nw - 1) / nw;
    if (nw == 1) pt = a.p <= 8 ? 8 : 16;  // single warp: the full-width variants measured faster
  }
  while ((a.p + pt - 1) / pt > 16) pt *= 2;
  switch (pt) {
    case 4: return launch_restart_t<4>(a, num_sms, st);
    case 8: return launch_restart_t<8>(a, num_sms, st);
    case 9: return launch_restart_t<9>(a, num_sms, st);
    case 10: return launch_restart_t<10>(a, num_sms, st);
    case 11: return launch_restart_t<11>(a, num_sms, st);
    case 12: return launch_restart_t<12>(a, num_sms, st);
    case 13: return launch_restart_t<13>(a, num_sms, st);
    case 14: return launch_restart_t<14>(a, num_sms, st);
    case 15: return launch_restart_t<15>(a, num_sms, st);
    case 16: return launch_restart_t<16>(a, num_sms, st);
    default: break;
  }
  if (pt < 8) return launch_restart_t<8>(a, num_sms, st);
  return launch_restart_t<16>(a, num_sms, st);
}

// ------------------------------------------------------------------ materialise scales
__global__ void __launch_bounds__(256) materialize_kernel(cplx* U, int64_t n, int64_t ld, int col0,
                                                          int ncols, double* scale) {
  for (int c = 0; c < ncols; ++c) {
    const double s = scale[col0 + c];
    if (s == 1.0) continue;
    cplx* col = U + (int64_t)(col0 + c) * ld;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x)
      col[r] = cscale(col[r], s);
  }
}
// real storage <-> complex128: column conversions used at the boundary and at a mode switch
__global__ void __launch_bounds__(256) pack_real_kernel(const cplx* __restrict__ src, double* dst,
                                                        int64_t n) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x)
    dst[r] = src[r].x;
}
__global__ void __launch_bounds__(256) unpack_real_kernel(const double* __restrict__ src, cplx* dst,
                                                          int64_t n, double scale) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x)
    dst[r] = make_double2(src[r] * scale, 0.0);
}
cudaError_t launch_pack_real(const cplx* src, double* dst, int64_t n, int num_sms, cudaStream_t st) {
  int64_t grid = (n + 255) / 256;
  if (grid > (int64_t)num_sms * 8) grid = (int64_t)num_sms * 8;
  if (grid < 1) grid = 1;
  pack_real_kernel<<<(int)grid, 256, 0, st>>>(src, dst, n);
  return cudaGetLastError();
}
cudaError_t launch_unpack_real(const double* src, cplx* dst, int64_t n, double scale, int num_sms,
                               cudaStream_t st) {
  int64_t grid = (n + 255) / 256;
  if (grid > (int64_t)num_sms * 8) grid = (int64_t)num_sms * 8;
  if (grid < 1) grid = 1;
  unpack_real_kernel<<<(int)grid, 256, 0, st>>>(src, dst, n, scale);
  return cudaGetLastError();
}

// y = (*scale) * x : the device-operator path hands the operator the true v_j = s_j U_j
__global__ void __launch_bounds__(256) scaled_copy_kernel(const cplx* __restrict__ x, cplx* __restrict__ y,
                                                          int64_t n, const double* scale) {
  const double s = scale ? *scale : 1.0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (int64_t)gridDim.x * blockDim.x)
    st_stream(y + r, cscale(ld_stream(x + r), s));
}
cudaError_t launch_scaled_copy(const cplx* x, cplx* y, int64_t n, const double* scale, int real,
                               int num_sms, cudaStream_t st) {
  (void)real;  // n counts 16-byte elements in both storage modes
  int64_t grid = (n + 255) / 256;
  if (grid > (int64_t)num_sms * 8) grid = (int64_t)num_sms * 8;
  if (grid < 1) grid = 1;
  scaled_copy_kernel<<<(int)grid, 256, 0, st>>>(x, y, n, scale);
  return cudaGetLastError();
}

__global__ void reset_scale_kernel(double* scale, int col0, int ncols) {
  const int i = threadIdx.x;
  if (i < ncols) scale[col0 + i] = 1.0;
}

cudaError_t launch_materialize(cplx* U, int64_t n, int64_t ld, int col0, int ncols, double* scale,
                               int num_sms, cudaStream_t st) {
  if (ncols <= 0) return cudaSuccess;
  int64_t grid = (n + 255) / 256;
  if (grid > (int64_t)num_sms * 8) grid = (int64_t)num_sms * 8;
  if (grid < 1) grid = 1;
  materialize_kernel<<<(int)grid, 256, 0, st>>>(U, n, ld, col0, ncols, scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  for (int c0 = 0; c0 < ncols; c0 += 256) {
    const int nc = ncols - c0 < 256 ? ncols - c0 : 256;
    reset_scale_kernel<<<1, 256, 0, st>>>(scale, col0 + c0, nc);
  }
  return cudaGetLastError();
}

}  // namespace ab200
