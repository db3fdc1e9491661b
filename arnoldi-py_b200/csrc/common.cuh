// common.cuh -- shared device helpers for the sm_100a Krylov-Schur kernels.
//
// Everything n-length in this library is complex128 stored as interleaved
// (re, im) doubles, i.e. one 16-byte element = one 128-bit load.  The Krylov
// basis is column-major with a padded leading dimension so every column starts
// on a 128-byte line.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ab200 {

typedef double2 cplx;  // .x = re, .y = im

constexpr int kWarp = 32;
constexpr int kNumSMsB200 = 148;
constexpr int kMaxDim = 256;        // basis columns a solver may hold (ab200_create)
constexpr int kPass1MaxCols = 128;  // columns one CGS pass-1 / fused launch covers (16 warps x 8)
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------- loads / stores
// Streaming 128-bit load of read-only data that is touched once per kernel
// (columns of the Krylov basis): bypass L1 allocation, keep L1 for w / coefficients.
__device__ __forceinline__ cplx ld_stream(const cplx* p) {
  cplx r;
  asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
// Read-only 128-bit load that may be re-used by neighbouring warps (w, gathered x).
__device__ __forceinline__ cplx ld_ro(const cplx* p) {
  cplx r;
  asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
// Plain (coherent) 128-bit load for data written earlier in the same kernel family.
__device__ __forceinline__ cplx ld_plain(const cplx* p) {
  cplx r;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
  return r;
}
// Coherent 128-bit load the compiler may schedule freely (for data this kernel also writes,
// at addresses it has not written yet).
__device__ __forceinline__ cplx ld_coherent(const cplx* p) {
  cplx r;
  asm("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y)
               : "memory");
}

// ---------------------------------------------------------------- complex arithmetic
// acc += conj(a) * b        (the inner product <a, b> term, as zgemv trans=2 / vdot)
__device__ __forceinline__ void cfma_conj(cplx& acc, const cplx a, const cplx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}
// acc += a * b
__device__ __forceinline__ void cfma(cplx& acc, const cplx a, const cplx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
// acc -= a * b
__device__ __forceinline__ void cfms(cplx& acc, const cplx a, const cplx b) {
  acc.x = fma(-a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(-a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}
// ---- the same three updates when the basis is stored REAL: one 16-byte element then holds
// two consecutive real rows (x = row 2k, y = row 2k+1) and every coefficient is real (.x)
template <bool REAL>
__device__ __forceinline__ void dotacc(cplx& acc, const cplx a, const cplx b) {
  if (REAL) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
  } else {
    cfma_conj(acc, a, b);
  }
}
template <bool REAL>
__device__ __forceinline__ void addax(cplx& acc, const cplx a, const cplx c) {
  if (REAL) {
    acc.x = fma(a.x, c.x, acc.x);
    acc.y = fma(a.y, c.x, acc.y);
  } else {
    cfma(acc, a, c);
  }
}
template <bool REAL>
__device__ __forceinline__ void subax(cplx& acc, const cplx a, const cplx c) {
  if (REAL) {
    acc.x = fma(-a.x, c.x, acc.x);
    acc.y = fma(-a.y, c.x, acc.y);
  } else {
    cfms(acc, a, c);
  }
}
__device__ __forceinline__ cplx cadd(const cplx a, const cplx b) {
  return make_double2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ cplx cscale(const cplx a, const double s) {
  return make_double2(a.x * s, a.y * s);
}

// ---------------------------------------------------------------- reductions
// Fixed-order butterfly: the result is bit-identical on every lane and from run
// to run (no atomics on floating point anywhere in this library).
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ cplx warp_sum(cplx v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}

// "last block done" ticket: returns true in exactly one block, after every other
// block's global writes that preceded its own ticket are visible.
__device__ __forceinline__ bool last_block_ticket(unsigned* ticket, unsigned nblocks,
                                                  int* smem_flag) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    int last = (t == nblocks - 1);
    if (last) {
      *ticket = 0;  // ready for the next launch on the same stream
      __threadfence();
    }
    *smem_flag = last;
  }
  __syncthreads();
  return *smem_flag != 0;
}

// ---------------------------------------------------------------- mbarrier / bulk copy (sm_90+)
// SASS: SYNCS.* for the mbarrier operations, UBLKCP for cp.async.bulk.
__device__ __forceinline__ unsigned smem_addr(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// the same for waits that can last microseconds (a whole round of another warp): sleep between
// polls so the spinning warp does not take issue slots from the working ones
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(64);
}
// global -> shared bulk copy; dst, src and bytes are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes,
                                         unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_addr(dst)),
      "l"(src), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}

// ---------------------------------------------------------------- L2 eviction hints
// Streams that are read exactly once (CSR values / column ids) are marked evict-first so they
// do not push the gathered vector x -- which is re-read all the time -- out of L2.
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ cplx ld_stream_ef(const cplx* p, unsigned long long pol) {
  cplx r;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
      : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ int4 ld_stream_ef(const int4* p, unsigned long long pol) {
  int4 r;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, unsigned bytes,
                                              unsigned long long* bar, unsigned long long pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_addr(dst)),
      "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(pol)
      : "memory");
}

// Function attributes (opt-in shared memory) are per device: remember per device ordinal
// whether a launcher has configured its kernel yet.
constexpr int kMaxDevices = 64;
struct PerDeviceOnce {
  bool done[kMaxDevices] = {false};
  bool first_use() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

}  // namespace ab200
