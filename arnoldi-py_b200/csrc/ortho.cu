// ortho.cu -- fused orthogonalisation kernels (the reference's ortho.py).
//
// dgks_gs  (ortho.py:56-107): one round = pass 1  h = V^H w and ||w||^2 in ONE sweep
//                                          pass 2  w -= V h and ||w||^2 in ONE sweep
// dgks_mgs (ortho.py:9-53):   one sweep = c+1 kernels, kernel i fuses the axpy of
//                             column i-1 with the dot product against column i.
//
// The basis is stored un-normalised: column i holds U_i with V_i = scale[i] * U_i
// (scale[i] = 1/beta_i, decomposition.py:65-66), so the reference's `w /= beta`
// pass over n elements never happens; the scale is folded into the c-length
// coefficient vectors here.
//
// Every reduction is two-stage in a fixed order (lane butterfly -> per-block
// partial -> last block sums the partials in block order), so results are
// bit-reproducible run to run.  No floating-point atomics.
#include <type_traits>

#include "kernels.cuh"

namespace ab200 {

// ------------------------------------------------------------------ finalisers
// Sum per-block partials of column i in block order: one warp per column.
__device__ __forceinline__ cplx sum_partials(const cplx* part, int nblocks, int lane) {
  cplx a = make_double2(0.0, 0.0);
  for (int b = lane; b < nblocks; b += kWarp) a = cadd(a, part[b]);
  return warp_sum(a);
}
__device__ __forceinline__ double sum_partials(const double* part, int nblocks, int lane) {
  double a = 0.0;
  for (int b = lane; b < nblocks; b += kWarp) a += part[b];
  return warp_sum(a);
}

// Last block: red[2i], red[2i+1] = sum over blocks of column i's partial, red[2c] = sum of the
// norm partials.  Same arithmetic per column as sum_partials (lanes stride the blocks, fixed
// butterfly), but a warp takes FOUR columns at a time so that their loads are all in flight
// together: the last block is alone on the device here and every dependent L2 round trip of
// its tail is paid by the whole grid (and, on several GPUs, by every rank waiting for it).
__device__ __forceinline__ void reduce_partials(const cplx* part, int gcap, const double* npart, int nblocks,
                                                int c, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  constexpr int G = 4;
  for (int i0 = warp * G; i0 < c; i0 += nwarps * G) {
    cplx acc[G];
#pragma unroll
    for (int k = 0; k < G; ++k) acc[k] = make_double2(0.0, 0.0);
    for (int b = lane; b < nblocks; b += kWarp) {
      cplx v[G];
#pragma unroll
      for (int k = 0; k < G; ++k)
        v[k] = (i0 + k < c) ? part[(size_t)(i0 + k) * gcap + b] : make_double2(0.0, 0.0);
#pragma unroll
      for (int k = 0; k < G; ++k) acc[k] = cadd(acc[k], v[k]);
    }
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const cplx g = warp_sum(acc[k]);
      if (lane == 0 && i0 + k < c) {
        red[2 * (i0 + k)] = g.x;
        red[2 * (i0 + k) + 1] = g.y;
      }
    }
  }
  if (warp == nwarps - 1) {   // the warp with the fewest columns
    const double s = sum_partials(npart, nblocks, lane);
    if (lane == 0) red[2 * c] = s;
  }
}

// Cross-GPU sum of `count` doubles held in red[] (local result), in rank order.
// Called by every thread of the last block; returns with red[] = global sums.
//
// One NVLink trip, no flag round: every double travels as two 8-byte packets
// {32 data bits, 32-bit sequence number}, each written with ONE 64-bit store into slot[my rank]
// of every rank.  An aligned 64-bit store is single-copy atomic, so a reader that sees the
// sequence number sees the data that came with it: no fence between payload and flag, no wait
// for the remote stores to be acknowledged (the first version pushed the payload, fenced at
// system scope -- one NVLink round trip -- and only then raised a flag).  Slots alternate with
// the parity of the sequence number, so a packet of exchange seq can only be overwritten by
// exchange seq + 2, which no rank starts before every rank has finished reading seq.
// The system-scope fence BEFORE the stores keeps the release semantics the halo reads rely on:
// whoever receives this rank's packets also sees every basis entry this rank wrote before.
__device__ __forceinline__ void st_packet(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_packet(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ void peer_allreduce(const PeerComm& pc, double* red, int count, StepCtl* ctl) {
  if (pc.nranks <= 1) return;
  __shared__ unsigned long long s_seq;
  const int tid = threadIdx.x;
  if (tid == 0) s_seq = *pc.seq + 1ull;
  __syncthreads();
  const unsigned long long seq = s_seq;
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  const int par = (int)(seq & 1ull);
  __threadfence_system();   // release: this rank's basis writes precede its packets
  // 1. my partial into slot[my rank] of every rank (mine included)
  const size_t my_off = 2 * ((size_t)par * kMaxRanks + pc.rank) * pc.slot_doubles;
  for (int k = tid; k < count; k += blockDim.x) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(red[k]);
    const unsigned long long lo = tag | (bits & 0xffffffffull), hi = tag | (bits >> 32);
    for (int r = 0; r < pc.nranks; ++r) {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(pc.slots[r]) + my_off + 2 * (size_t)k;
      st_packet(dst, lo);
      st_packet(dst + 1, hi);
    }
  }
  __syncthreads();   // red[] has been read by everyone before it is overwritten below
  // 2. collect: wait for both packets of every rank's k-th double, sum in rank order
  //    (wall-clock bound: a peer that died or diverged must not hang the box; on timeout every
  //    later kernel becomes a no-op)
  const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(pc.slots[pc.rank]) +
                                   2 * (size_t)par * kMaxRanks * pc.slot_doubles;
  const unsigned long long t0 = global_ns();
  bool dead = false;
  for (int k = tid; k < count; k += blockDim.x) {
    double a = 0.0;
    for (int r = 0; r < pc.nranks; ++r) {
      const unsigned long long* src = mine + 2 * ((size_t)r * pc.slot_doubles + k);
      unsigned long long lo, hi;
      unsigned spins = 0;
      for (;;) {
        lo = ld_packet(src);
        hi = ld_packet(src + 1);
        if (((lo ^ tag) >> 32) == 0 && ((hi ^ tag) >> 32) == 0) break;
        if (dead || ((++spins & 0x3ffu) == 0 && global_ns() - t0 > kPeerTimeoutNs)) {
          dead = true;
          ctl->comm_error = 1;
          ctl->stop = 1;
          break;
        }
      }
      a += __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
    }
    red[k] = a;
  }
  __syncthreads();
  __threadfence_system();   // acquire: later reads of peer memory see what the senders released
  if (tid == 0) *pc.seq = seq;
}

// After pass 1 (or an MGS dot): g_i -> h_i = s_i g_i, H column, pass-2 coefficients.
//   red[2*i], red[2*i+1] = g_i ; red[2*ncols] = ||w||^2 (when want_norm)
__device__ void finish_dots(const OrthoArgs& a, int col0, int ncols, double* red, bool want_norm) {
  const int tid = threadIdx.x;
  for (int i = tid; i < ncols; i += blockDim.x) {
    const double s = a.scale[col0 + i];
    cplx h = make_double2(red[2 * i] * s, red[2 * i + 1] * s);
    cplx old = a.hcol[col0 + i];
    a.hcol[col0 + i] = a.accumulate ? cadd(old, h) : h;
    a.coef[col0 + i] = cscale(h, s);
  }
  if (want_norm && tid == 0) a.ctl->nrm0sq = red[2 * ncols];
}

// End of an Arnoldi step: breakdown test, H[j+1, j], lazy scale of the new column.
__device__ void finalize_step(const OrthoArgs& a, double beta) {
  StepCtl* ctl = a.ctl;
  if (!a.finalize) return;
  if (a.finalize == 1) ctl->steps_total += 1;  // 2 = normalise a column outside an Arnoldi step
  if (beta < a.tol) {  // ortho.py:107, decomposition.py:61-63
    ctl->stop = 1;
    ctl->broke_at = a.j;
    a.scale[a.j + 1] = 1.0;  // the reference leaves w un-normalised
  } else {
    a.hcol[a.j + 1] = make_double2(beta, 0.0);  // decomposition.py:65
    a.scale[a.j + 1] = 1.0 / beta;              // decomposition.py:66, applied lazily
  }
}

// DGKS test after the first round (ortho.py:101, strict `<`); records the decision.
__device__ bool dgks_decide(const OrthoArgs& a, double beta) {
  StepCtl* ctl = a.ctl;
  ctl->beta = beta;
  ctl->rounds_total += 1;
  const bool again = beta < a.eta * sqrt(ctl->nrm0sq);
  ctl->round2 = again ? 1 : 0;
  if (again) ctl->second_total += 1;
  if (a.step_flag) *a.step_flag = again ? 1 : 0;
  return again;
}

// After pass 2 (or the last MGS axpy): beta, DGKS decision, breakdown, H[j+1,j], scale.
__device__ void finish_norm(const OrthoArgs& a, double nrmsq) {
  StepCtl* ctl = a.ctl;
  const double beta = sqrt(nrmsq);
  bool again = false;
  if (a.round == 1) {
    again = dgks_decide(a, beta);
  } else {
    ctl->beta = beta;
    ctl->rounds_total += 1;
    ctl->round2 = 0;
  }
  if (!again) finalize_step(a, beta);
}

// ------------------------------------------------------------------ CGS pass 1
// Block = W warps.  All warps sweep the SAME rows (so w is read from HBM once and
// re-used through L1); warp k owns columns [k*CT, k*CT+CT) and keeps their
// accumulators in registers for the whole kernel.  Lanes run along n, so each
// load instruction of a warp covers 512 contiguous bytes of one column.
template <int CT, int R, bool REAL>
__global__ void __launch_bounds__(CT < 8 ? 256 : 512) cgs_pass1_kernel(OrthoArgs a) {
  StepCtl* ctl = a.ctl;
  if (ctl->stop) return;
  if (a.round == 2 && !ctl->round2) return;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int c = a.p1_ncols;  // columns [p1_col0, p1_col0 + c) of the basis
  const int mycol0 = warp * CT;
  int mycols = c - mycol0;
  mycols = mycols < 0 ? 0 : (mycols > CT ? CT : mycols);
  const int64_t ld = a.ld;
  const cplx* __restrict__ w = a.w;
  const cplx* __restrict__ U = a.U + (int64_t)(a.p1_col0 + mycol0) * ld;

  cplx acc[CT];
#pragma unroll
  for (int k = 0; k < CT; ++k) acc[k] = make_double2(0.0, 0.0);
  double nacc = 0.0;

  constexpr int ROWS = kWarp * R;
  const int64_t nchunks = (a.n + ROWS - 1) / ROWS;
  for (int64_t q = blockIdx.x; q < nchunks; q += gridDim.x) {
    const int64_t base = q * ROWS + lane;
    cplx wv[R];
    cplx v[CT][R];
    if (q * ROWS + ROWS <= a.n) {
#pragma unroll
      for (int r = 0; r < R; ++r) wv[r] = ld_ro(w + base + r * kWarp);
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        if (k < mycols) {
#pragma unroll
          for (int r = 0; r < R; ++r) v[k][r] = ld_stream(U + (int64_t)k * ld + base + r * kWarp);
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool ok = base + r * kWarp < a.n;
        wv[r] = ok ? ld_ro(w + base + r * kWarp) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        if (k < mycols) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const bool ok = base + r * kWarp < a.n;
            v[k][r] = ok ? ld_stream(U + (int64_t)k * ld + base + r * kWarp)
                         : make_double2(0.0, 0.0);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < CT; ++k) {
      if (k < mycols) {
#pragma unroll
        for (int r = 0; r < R; ++r) dotacc<REAL>(acc[k], v[k][r], wv[r]);
      }
    }
    if (warp == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        nacc = fma(wv[r].x, wv[r].x, nacc);
        nacc = fma(wv[r].y, wv[r].y, nacc);
      }
    }
  }

  // per-block partials, column-major by block so the final sums are coalesced
  const int gcap = a.grid_cap;
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    if (k < mycols) {
      cplx s = warp_sum(acc[k]);
      if (lane == 0) a.part[(size_t)(mycol0 + k) * gcap + blockIdx.x] = s;
    }
  }
  if (warp == 0) {
    double s = warp_sum(nacc);
    if (lane == 0) a.npart[blockIdx.x] = s;
  }

  __shared__ int s_last;
  if (!last_block_ticket(a.ticket, gridDim.x, &s_last)) return;

  // ---- last block: reduce across blocks (fixed order), across GPUs, finish
  extern __shared__ double red[];  // 2*c + 1 doubles
  const int nwarps = blockDim.x >> 5;
  reduce_partials(a.part, gcap, a.npart, gridDim.x, c, red);
  __syncthreads();
  peer_allreduce(a.comm, red, 2 * c + 1, ctl);
  __syncthreads();
  finish_dots(a, a.p1_col0, c, red, a.round == 1 && a.p1_col0 == 0);
}

// ------------------------------------------------------------------ CGS pass 2
// w -= U * coef ; ||w||^2.  Warp = R*32 contiguous rows, all c columns; the
// coefficients sit in shared memory and are broadcast.
template <int R, int UC, bool REAL>
__global__ void __launch_bounds__(256) cgs_pass2_kernel(OrthoArgs a) {
  StepCtl* ctl = a.ctl;
  if (ctl->stop) return;
  if (a.round == 2 && !ctl->round2) return;

  extern __shared__ double smem_d[];
  cplx* scoef = reinterpret_cast<cplx*>(smem_d);
  const int c = a.ncols;
  for (int i = threadIdx.x; i < c; i += blockDim.x) scoef[i] = a.coef[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int64_t ld = a.ld;
  const cplx* __restrict__ U = a.U;
  cplx* w = a.w;
  double nacc = 0.0;

  constexpr int ROWS = kWarp * R;
  const int64_t nchunks = (a.n + ROWS - 1) / ROWS;
  for (int64_t q = (int64_t)blockIdx.x * nwarps + warp; q < nchunks;
       q += (int64_t)gridDim.x * nwarps) {
    const int64_t base = q * ROWS + lane;
    const bool full = q * ROWS + ROWS <= a.n;
    cplx acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool ok = full || base + r * kWarp < a.n;
      acc[r] = ok ? ld_plain(w + base + r * kWarp) : make_double2(0.0, 0.0);
    }
    int i = 0;
    if (full) {
      for (; i + UC <= c; i += UC) {
        cplx v[UC][R];
#pragma unroll
        for (int u = 0; u < UC; ++u)
#pragma unroll
          for (int r = 0; r < R; ++r)
            v[u][r] = ld_stream(U + (int64_t)(i + u) * ld + base + r * kWarp);
#pragma unroll
        for (int u = 0; u < UC; ++u) {
          const cplx cf = scoef[i + u];
#pragma unroll
          for (int r = 0; r < R; ++r) subax<REAL>(acc[r], v[u][r], cf);
        }
      }
    }
    for (; i < c; ++i) {
      cplx v[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool ok = full || base + r * kWarp < a.n;
        v[r] = ok ? ld_stream(U + (int64_t)i * ld + base + r * kWarp) : make_double2(0.0, 0.0);
      }
      const cplx cf = scoef[i];
#pragma unroll
      for (int r = 0; r < R; ++r) subax<REAL>(acc[r], v[r], cf);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool ok = full || base + r * kWarp < a.n;
      if (ok) {
        st_stream(w + base + r * kWarp, acc[r]);
        nacc = fma(acc[r].x, acc[r].x, nacc);
        nacc = fma(acc[r].y, acc[r].y, nacc);
      }
    }
  }

  __shared__ double s_warp[32];
  __shared__ int s_last;
  double s = warp_sum(nacc);
  if (lane == 0) s_warp[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < nwarps; ++k) t += s_warp[k];
    a.npart[blockIdx.x] = t;
  }
  if (!last_block_ticket(a.ticket, gridDim.x, &s_last)) return;

  __shared__ double s_red[1];
  if (warp == 0) {
    double t = sum_partials(a.npart, gridDim.x, lane);
    if (lane == 0) s_red[0] = t;
  }
  __syncthreads();
  peer_allreduce(a.comm, s_red, 1, ctl);
  __syncthreads();
  if (threadIdx.x == 0) finish_norm(a, s_red[0]);
}

// ------------------------------------------------------------------ CGS fused pass
// Round-1 pass 2 and round-2 pass 1 in ONE sweep over the basis:
//     w' = w - U coef          (ortho.py:96)
//     g' = U^H w', ||w'||^2    (ortho.py:98 and, should the DGKS test fire, :102)
// Row r of w' depends only on row r of U, so while a chunk of U is in registers it is
// used twice.  If the DGKS test then fires (100% of the steps on the 2-D Laplacian) the
// second round only needs its pass 2; if it does not, g' is simply dropped.  Same block
// shape as pass 1: warp k owns columns [k*CT, k*CT+CT); the per-warp partial sums of
// U coef meet in shared memory (double buffered: one __syncthreads per chunk).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;  // src-size 0 => the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// PF = true: the block's NEXT chunk is staged into shared memory with cp.async (LDGSTS,
// no registers held) while the current one is processed, so requests stay in flight across
// the block barrier.  PF = false: plain register loads (more blocks per SM instead).
template <int CT, int R, bool PF, bool REAL, int S>
__global__ void __launch_bounds__(CT < 5 ? 256 : 512) cgs_fused_kernel(OrthoArgs a) {
  StepCtl* ctl = a.ctl;
  if (ctl->stop) return;

  constexpr int ROWS = kWarp * R;
  extern __shared__ __align__(16) unsigned char fused_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int c = a.ncols;
  cplx* spart = reinterpret_cast<cplx*>(fused_smem);  // [2][nwarps][ROWS]
  cplx* scoef = spart + 2 * nwarps * ROWS;            // [nwarps * CT]
  cplx* swp = scoef + nwarps * CT;                    // [ROWS]                  w' of the chunk
  cplx* sw = swp + ROWS;                              // PF: [S][ROWS]           chunk of w
  cplx* sv = sw + S * ROWS;                           // PF: [S][nwarps][CT][ROWS] chunk of U
  for (int i = threadIdx.x; i < nwarps * CT; i += blockDim.x)
    scoef[i] = i < c ? a.coef[i] : make_double2(0.0, 0.0);
  __syncthreads();

  const int mycol0 = warp * CT;
  int mycols = c - mycol0;
  mycols = mycols < 0 ? 0 : (mycols > CT ? CT : mycols);
  const int64_t ld = a.ld;
  cplx* w = a.w;
  const cplx* __restrict__ U = a.U + (int64_t)mycol0 * ld;
  cplx cf[CT];
#pragma unroll
  for (int k = 0; k < CT; ++k) cf[k] = scoef[mycol0 + k];

  cplx acc[CT];
#pragma unroll
  for (int k = 0; k < CT; ++k) acc[k] = make_double2(0.0, 0.0);
  double nacc = 0.0;

  const int64_t nchunks = (a.n + ROWS - 1) / ROWS;
  auto stage_v = [&](int64_t q, int b) {  // my columns of chunk q -> sv[b][warp]
    const int64_t base = q * ROWS + lane;
    cplx* dst = sv + ((size_t)b * nwarps + warp) * CT * ROWS;
#pragma unroll
    for (int k = 0; k < CT; ++k) {
      if (k < mycols) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int64_t row = base + r * kWarp;
          const bool ok = row < a.n;
          cp_async16(dst + k * ROWS + r * kWarp + lane, U + (int64_t)k * ld + (ok ? row : 0), ok);
        }
      }
    }
  };
  auto stage_w = [&](int64_t q, int b) {  // chunk q of w -> sw[b]   (warp 0 only)
    const int64_t base = q * ROWS + lane;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = base + r * kWarp;
      const bool ok = row < a.n;
      cp_async16(sw + (size_t)b * ROWS + r * kWarp + lane, w + (ok ? row : 0), ok);
    }
  };

  int buf = 0;
  if (PF) {
    // prologue: S - 1 chunks in flight before the first one is consumed
#pragma unroll
    for (int st = 0; st < S - 1; ++st) {
      const int64_t q0 = (int64_t)blockIdx.x + (int64_t)st * gridDim.x;
      if (q0 < nchunks) {
        if (warp == 0) stage_w(q0, st);
        stage_v(q0, st);
      }
      cp_async_commit();
    }
  }
  unsigned iter = 0;
  for (int64_t q = blockIdx.x; q < nchunks; q += gridDim.x, buf ^= 1, ++iter) {
    const int64_t base = q * ROWS + lane;
    const bool full = q * ROWS + ROWS <= a.n;
    const int slot = (int)(iter % S);
    cplx wv[R];
    cplx v[CT][R];
    if (PF) {
      // refill the slot consumed in the previous iteration (its readers are past barrier 2)
      const int64_t qn = q + (int64_t)(S - 1) * gridDim.x;
      const int nslot = (int)((iter + S - 1) % S);
      if (qn < nchunks) {
        if (warp == 0) stage_w(qn, nslot);
        stage_v(qn, nslot);
      }
      cp_async_commit();
      cp_async_wait<S - 1>();  // all but the S - 1 newest groups have landed: this chunk is in
      __syncwarp();
      const cplx* src = sv + ((size_t)slot * nwarps + warp) * CT * ROWS;
#pragma unroll
      for (int k = 0; k < CT; ++k)
        if (k < mycols) {
#pragma unroll
          for (int r = 0; r < R; ++r) v[k][r] = src[k * ROWS + r * kWarp + lane];
        }
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool ok = full || base + r * kWarp < a.n;
        wv[r] = ok ? ld_coherent(w + base + r * kWarp) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        if (k < mycols) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const bool ok = full || base + r * kWarp < a.n;
            v[k][r] =
                ok ? ld_stream(U + (int64_t)k * ld + base + r * kWarp) : make_double2(0.0, 0.0);
          }
        }
      }
    }
    // my columns' share of U coef for these rows
    cplx* mine = spart + ((size_t)buf * nwarps + warp) * ROWS;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      cplx t = make_double2(0.0, 0.0);
#pragma unroll
      for (int k = 0; k < CT; ++k)
        if (k < mycols) addax<REAL>(t, v[k][r], cf[k]);
      mine[r * kWarp + lane] = t;
    }
    __syncthreads();  // partial sums (and, PF, warp 0's chunk of w) visible to the block
    // w' of the chunk is rebuilt by the first ROWS threads of the block, one element each (so
    // two or more warps share what used to be one warp's serial section between the barriers),
    // stored, and published through shared memory; everybody reads it back.  The W partial sums
    // of an element are added in warp order, as before.  (Every warp doing the W-term sum for
    // all ROWS elements itself saturated the shared-memory pipe.)
    {
      const cplx* all = spart + (size_t)buf * nwarps * ROWS;
      for (int e = threadIdx.x; e < ROWS; e += blockDim.x) {
        cplx win;
        if (PF) {
          win = sw[(size_t)slot * ROWS + e];
        } else {
          win = wv[0];   // element e = r * 32 + lane sits in this thread's own wv[r]
#pragma unroll
          for (int r = 1; r < R; ++r)
            if ((e >> 5) == r) win = wv[r];
        }
        cplx t = make_double2(0.0, 0.0);
        for (int k2 = 0; k2 < nwarps; ++k2) t = cadd(t, all[k2 * ROWS + e]);
        win.x -= t.x;
        win.y -= t.y;
        swp[e] = win;
        const int64_t row = q * ROWS + e;
        if (row < a.n) {
          st_stream(w + row, win);
          nacc = fma(win.x, win.x, nacc);
          nacc = fma(win.y, win.y, nacc);
        }
      }
    }
    __syncthreads();  // w' of the chunk is published
#pragma unroll
    for (int r = 0; r < R; ++r) wv[r] = swp[r * kWarp + lane];
#pragma unroll
    for (int k = 0; k < CT; ++k) {
      if (k < mycols) {
#pragma unroll
        for (int r = 0; r < R; ++r) dotacc<REAL>(acc[k], v[k][r], wv[r]);
      }
    }
  }
  if (PF) cp_async_wait<0>();

  const int gcap = a.grid_cap;
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    if (k < mycols) {
      cplx s = warp_sum(acc[k]);
      if (lane == 0) a.part[(size_t)(mycol0 + k) * gcap + blockIdx.x] = s;
    }
  }
  __shared__ double s_nw[16];
  {
    double s = warp_sum(nacc);
    if (lane == 0) s_nw[warp] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < nwarps; ++k) t += s_nw[k];
    a.npart[blockIdx.x] = t;
  }

  __shared__ int s_last;
  __shared__ int s_again;
  if (!last_block_ticket(a.ticket, gridDim.x, &s_last)) return;

  double* red = reinterpret_cast<double*>(fused_smem);  // 2*c + 1 doubles (spart is free now)
  reduce_partials(a.part, gcap, a.npart, gridDim.x, c, red);
  __syncthreads();
  peer_allreduce(a.comm, red, 2 * c + 1, ctl);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double beta = sqrt(red[2 * c]);
    const bool again = dgks_decide(a, beta);
    if (!again) finalize_step(a, beta);
    s_again = again ? 1 : 0;
  }
  __syncthreads();
  // second round wanted: h += s g', pass-2 coefficients of round 2   (ortho.py:102-103)
  if (s_again) finish_dots(a, 0, c, red, false);
}

// ------------------------------------------------------------------ CGS fused pass, pipelined
// Same sweep as cgs_fused_kernel (w' = w - U coef ; g' = U^H w' ; ||w'||^2) with the two block
// barriers per chunk taken out.  ncu on cgs_fused_kernel: top stall = the __syncthreads around
// the serial rebuild of w' (one warp works, the others wait, twice per chunk).  Here the
// chunks are software-pipelined and the hand-offs are mbarriers that only the warp that needs
// the data waits on:
//
//   every warp, iteration i :  partial sums of U coef for chunk i -> spart[i & 1]; arrive full[i & 1]
//   rebuilder(i) = warp i % W:  wait full[i & 1]; w'(i) = w(i) - sum of partials -> swp[i & 1],
//                               global store, norm; arrive ready[i & 1]
//   every warp, iteration i :  wait ready[(i-1) & 1]; dots of chunk i-1 against w'(i-1)
//
// so while one warp rebuilds w'(i) the others already run the dots of chunk i-1 and the partial
// sums of chunk i+1: nobody idles at a block-wide rendezvous.  U is staged by cp.async into a
// 3-slot ring (each warp refills only its own columns' slice, so the ring needs no barrier);
// the dots re-read their chunk from the ring instead of holding it in registers.
template <int CT, int R, bool REAL, int MAXT>
__global__ void __launch_bounds__(MAXT) cgs_fused_pipe_kernel(OrthoArgs a) {
  StepCtl* ctl = a.ctl;
  if (ctl->stop) return;

  constexpr int ROWS = kWarp * R;
  constexpr int S = 2;
  extern __shared__ __align__(16) unsigned char fused_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int c = a.ncols;
  cplx* spart = reinterpret_cast<cplx*>(fused_smem);  // [2][nwarps][ROWS]
  cplx* scoef = spart + 2 * nwarps * ROWS;            // [nwarps * CT]
  cplx* swp = scoef + nwarps * CT;                    // [2][ROWS]            w' of a chunk
  cplx* sw = swp + 2 * ROWS;                          // [4][ROWS]            chunk of w
  cplx* sv = sw + 4 * ROWS;                           // [S][nwarps][CT][ROWS] chunk of U
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sv + (size_t)S * nwarps * CT * ROWS);
  unsigned long long* full = bars;       // [2] partial sums of a chunk are complete
  unsigned long long* ready = bars + 2;  // [2] w' of a chunk is published
  for (int i = threadIdx.x; i < nwarps * CT; i += blockDim.x)
    scoef[i] = i < c ? a.coef[i] : make_double2(0.0, 0.0);
  if (threadIdx.x == 0) {
    // full: one arrival per warp (the rebuilder counts itself); ready: the rebuilder alone
    mbar_init(full, nwarps), mbar_init(full + 1, nwarps);
    mbar_init(ready, 1), mbar_init(ready + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int mycol0 = warp * CT;
  int mycols = c - mycol0;
  mycols = mycols < 0 ? 0 : (mycols > CT ? CT : mycols);
  const int64_t ld = a.ld;
  cplx* w = a.w;
  const cplx* __restrict__ U = a.U + (int64_t)mycol0 * ld;
  cplx cf[CT];
#pragma unroll
  for (int k = 0; k < CT; ++k) cf[k] = scoef[mycol0 + k];
  cplx acc[CT];
#pragma unroll
  for (int k = 0; k < CT; ++k) acc[k] = make_double2(0.0, 0.0);
  double nacc = 0.0;

  const int64_t nchunks = (a.n + ROWS - 1) / ROWS;
  const int64_t nit = nchunks > blockIdx.x ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto chunk_of = [&](int64_t i) { return (int64_t)blockIdx.x + i * gridDim.x; };
  // chunk i: my columns -> sv[i % 2][warp] (a slice only its own warp touches); warp 0 also
  // stages w, in a 4-slot ring: chunk i + 4 is staged in iteration i + 2, i.e. after warp 0 ran
  // the dots of chunk i, which waited for the rebuilder of chunk i -- the reader of w(i)
  auto stage = [&](int64_t i) {
    if (i < nit) {
      const int64_t base = chunk_of(i) * ROWS + lane;
      cplx* dst = sv + ((size_t)(i % S) * nwarps + warp) * CT * ROWS;
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        if (k < mycols) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int64_t row = base + r * kWarp;
            const bool ok = row < a.n;
            cp_async16(dst + k * ROWS + r * kWarp + lane, U + (int64_t)k * ld + (ok ? row : 0), ok);
          }
        }
      }
      if (warp == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int64_t row = base + r * kWarp;
          const bool ok = row < a.n;
          cp_async16(sw + (size_t)(i % 4) * ROWS + r * kWarp + lane, w + (ok ? row : 0), ok);
        }
      }
    }
    cp_async_commit();
  };

  cplx vold[CT][R];  // chunk i - 1 of my columns, kept for its dots
  stage(0);
  stage(1);
  for (int64_t i = 0; i <= nit; ++i) {
    const int b = (int)(i & 1);
    cplx v[CT][R];
    if (i < nit) {
      cp_async_wait<1>();  // chunk i has landed (chunk i + 1 may still be in flight)
      __syncwarp();
      const cplx* src = sv + ((size_t)(i % S) * nwarps + warp) * CT * ROWS;
#pragma unroll
      for (int k = 0; k < CT; ++k)
        if (k < mycols) {
#pragma unroll
          for (int r = 0; r < R; ++r) v[k][r] = src[k * ROWS + r * kWarp + lane];
        }
      __syncwarp();
      stage(i + 2);  // the slice just read into registers is free: refill it
      // ---- my columns' share of U coef for chunk i
      cplx* mine = spart + ((size_t)b * nwarps + warp) * ROWS;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        cplx t = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < CT; ++k)
          if (k < mycols) addax<REAL>(t, v[k][r], cf[k]);
        mine[r * kWarp + lane] = t;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(full + b);  // release: my partial sums (and warp 0's chunk of w)
      // ---- the rebuilder of chunk i
      if (warp == (int)(i % nwarps)) {
        mbar_wait(full + b, (unsigned)((i >> 1) & 1));
        const int64_t base = chunk_of(i) * ROWS + lane;
        const bool fullchunk = chunk_of(i) * ROWS + ROWS <= a.n;
        const cplx* all = spart + (size_t)b * nwarps * ROWS;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          cplx wv = sw[(size_t)(i % 4) * ROWS + r * kWarp + lane];
          cplx t = make_double2(0.0, 0.0);
          for (int k2 = 0; k2 < nwarps; ++k2) t = cadd(t, all[k2 * ROWS + r * kWarp + lane]);
          wv.x -= t.x;
          wv.y -= t.y;
          swp[b * ROWS + r * kWarp + lane] = wv;
          const bool ok = fullchunk || base + r * kWarp < a.n;
          if (ok) {
            st_stream(w + base + r * kWarp, wv);
            nacc = fma(wv.x, wv.x, nacc);
            nacc = fma(wv.y, wv.y, nacc);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ready + b);  // release: w'(i) is in swp[b]
      }
    } else {
      cp_async_commit();
    }
    // ---- dots of the previous chunk (its w' was rebuilt while this warp did the work above)
    if (i > 0) {
      const int64_t j = i - 1;
      const int bj = (int)(j & 1);
      if (warp != (int)(j % nwarps)) mbar_wait(ready + bj, (unsigned)((j >> 1) & 1));
      cplx wv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) wv[r] = swp[bj * ROWS + r * kWarp + lane];
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        if (k < mycols) {
#pragma unroll
          for (int r = 0; r < R; ++r) dotacc<REAL>(acc[k], vold[k][r], wv[r]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < CT; ++k)
#pragma unroll
      for (int r = 0; r < R; ++r) vold[k][r] = v[k][r];
  }
  cp_async_wait<0>();

  const int gcap = a.grid_cap;
#pragma unroll
  for (int k = 0; k < CT; ++k) {
    if (k < mycols) {
      cplx s = warp_sum(acc[k]);
      if (lane == 0) a.part[(size_t)(mycol0 + k) * gcap + blockIdx.x] = s;
    }
  }
  __shared__ double s_nw[16];
  {
    double s = warp_sum(nacc);
    if (lane == 0) s_nw[warp] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < nwarps; ++k) t += s_nw[k];
    a.npart[blockIdx.x] = t;
  }

  __shared__ int s_last;
  __shared__ int s_again;
  if (!last_block_ticket(a.ticket, gridDim.x, &s_last)) return;

  double* red = reinterpret_cast<double*>(fused_smem);  // 2*c + 1 doubles (spart is free now)
  reduce_partials(a.part, gcap, a.npart, gridDim.x, c, red);
  __syncthreads();
  peer_allreduce(a.comm, red, 2 * c + 1, ctl);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double beta = sqrt(red[2 * c]);
    const bool again = dgks_decide(a, beta);
    if (!again) finalize_step(a, beta);
    s_again = again ? 1 : 0;
  }
  __syncthreads();
  if (s_again) finish_dots(a, 0, c, red, false);
}

// ------------------------------------------------------------------ CGS fused pass, warp tiles
// Same mathematics as cgs_fused_kernel (w' = w - U coef ; g' = U^H w' ; ||w'||^2 in one sweep),
// different decomposition: every WARP owns a chunk of 32 elements x ALL c columns, staged into
// a warp-private shared-memory tile with cp.async (double buffered).  The tile is read twice --
// once to build w', once for the c dot products, whose accumulators live in registers -- so
// no block barrier and no exchange of partial sums is needed in the sweep; warps only meet at
// the very end.  Used for c <= 64 (tile size); wider bases use cgs_fused_kernel.
template <int CMAX, bool REAL>
__global__ void __launch_bounds__(256) cgs_fused_warp_kernel(OrthoArgs a) {
  StepCtl* ctl = a.ctl;
  if (ctl->stop) return;
  extern __shared__ __align__(16) unsigned char fw_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int c = a.ncols;
  const int64_t ld = a.ld;
  cplx* w = a.w;
  const cplx* __restrict__ U = a.U;
  // layout: coef[CMAX] | per warp: tile[2][c + 1][32]   (slot c of a tile holds the chunk of w)
  cplx* scoef = reinterpret_cast<cplx*>(fw_smem);
  cplx* tiles = scoef + CMAX + (size_t)warp * 2 * (c + 1) * kWarp;
  for (int i = threadIdx.x; i < CMAX; i += blockDim.x)
    scoef[i] = i < c ? a.coef[i] : make_double2(0.0, 0.0);
  __syncthreads();

  typedef typename std::conditional<REAL, double, cplx>::type AccT;
  AccT acc[CMAX];
#pragma unroll
  for (int i = 0; i < CMAX; ++i) acc[i] = AccT();
  double nacc = 0.0;

  const int64_t nchunks = (a.n + kWarp - 1) / kWarp;
  const int64_t stride = (int64_t)gridDim.x * nwarps;
  auto stage = [&](int64_t q, int b) {
    const int64_t row = q * kWarp + lane;
    const bool ok = row < a.n;
    cplx* dst = tiles + (size_t)b * (c + 1) * kWarp + lane;
    const cplx* src = U + (ok ? row : 0);
    for (int i = 0; i < c; ++i) cp_async16(dst + i * kWarp, src + (int64_t)i * ld, ok);
    cp_async16(dst + c * kWarp, w + (ok ? row : 0), ok);
  };
  int64_t q = (int64_t)blockIdx.x * nwarps + warp;
  if (q < nchunks) stage(q, 0);
  cp_async_commit();
  int buf = 0;
  for (; q < nchunks; q += stride, buf ^= 1) {
    if (q + stride < nchunks) stage(q + stride, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    const cplx* tile = tiles + (size_t)buf * (c + 1) * kWarp + lane;
    cplx t = make_double2(0.0, 0.0), t2 = make_double2(0.0, 0.0);
    int i = 0;
    for (; i + 2 <= c; i += 2) {  // two chains: the sums are short dependent chains otherwise
      addax<REAL>(t, tile[i * kWarp], scoef[i]);
      addax<REAL>(t2, tile[(i + 1) * kWarp], scoef[i + 1]);
    }
    if (i < c) addax<REAL>(t, tile[i * kWarp], scoef[i]);
    cplx wv = tile[c * kWarp];
    wv.x -= (t.x + t2.x);
    wv.y -= (t.y + t2.y);
    const int64_t row = q * kWarp + lane;
    if (row < a.n) {
      st_stream(w + row, wv);
      nacc = fma(wv.x, wv.x, nacc);
      nacc = fma(wv.y, wv.y, nacc);
    }
#pragma unroll
    for (int k = 0; k < CMAX; ++k) {
      if (k < c) {
        const cplx v = tile[k * kWarp];
        if (REAL) {
          double& ar = *reinterpret_cast<double*>(&acc[k]);
          ar = fma(v.x, wv.x, ar);
          ar = fma(v.y, wv.y, ar);
        } else {
          cfma_conj(*reinterpret_cast<cplx*>(&acc[k]), v, wv);
        }
      }
    }
    __syncwarp();  // the tile may be refilled by the next iteration's staging
  }
  cp_async_wait<0>();
  __syncthreads();  // every warp is done with its tiles: the space is re-used below

  // block-level combine: warp sums -> shared -> one partial per block and column
  cplx* bp = reinterpret_cast<cplx*>(fw_smem) + CMAX;  // [nwarps][CMAX]
#pragma unroll
  for (int k = 0; k < CMAX; ++k) {
    if (k < c) {
      cplx v;
      if (REAL) {
        v = make_double2(warp_sum(*reinterpret_cast<double*>(&acc[k])), 0.0);
      } else {
        v = warp_sum(*reinterpret_cast<cplx*>(&acc[k]));
      }
      if (lane == 0) bp[warp * CMAX + k] = v;
    }
  }
  __shared__ double s_nw[8];
  {
    const double sN = warp_sum(nacc);
    if (lane == 0) s_nw[warp] = sN;
  }
  __syncthreads();
  const int gcap = a.grid_cap;
  for (int k = threadIdx.x; k < c; k += blockDim.x) {
    cplx v = make_double2(0.0, 0.0);
    for (int wi = 0; wi < nwarps; ++wi) v = cadd(v, bp[wi * CMAX + k]);
    a.part[(size_t)k * gcap + blockIdx.x] = v;
  }
  if (threadIdx.x == 0) {
    double tN = 0.0;
    for (int wi = 0; wi < nwarps; ++wi) tN += s_nw[wi];
    a.npart[blockIdx.x] = tN;
  }

  __shared__ int s_last;
  __shared__ int s_again;
  if (!last_block_ticket(a.ticket, gridDim.x, &s_last)) return;

  double* red = reinterpret_cast<double*>(fw_smem);  // 2*c + 1 doubles
  reduce_partials(a.part, gcap, a.npart, gridDim.x, c, red);
  __syncthreads();
  peer_allreduce(a.comm, red, 2 * c + 1, ctl);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double beta = sqrt(red[2 * c]);
    const bool again = dgks_decide(a, beta);
    if (!again) finalize_step(a, beta);
    s_again = again ? 1 : 0;
  }
  __syncthreads();
  if (s_again) finish_dots(a, 0, c, red, false);
}

// ------------------------------------------------------------------ MGS step kernel
// Kernel i of an MGS sweep (ortho.py:39-41 / :47-50), i = 0..c:
//   if i > 0:  w -= coef[i-1] * U_{i-1}          (axpy of the previous column)
//   if i < c:  g_i = <U_i, w>                     (dot with the next column)
//   if i == 0 and round 1: also ||w||^2           (ortho.py:36)
//   if i == c: ||w||^2 of the result              (ortho.py:43 / :52)
template <int R, bool REAL>
__global__ void __launch_bounds__(256) mgs_step_kernel(OrthoArgs a, int i) {
  StepCtl* ctl = a.ctl;
  if (ctl->stop) return;
  if (a.round == 2 && !ctl->round2) return;

  const int c = a.ncols;
  const bool do_axpy = i > 0;
  const bool do_dot = i < c;
  const bool norm_in = (i == 0);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int64_t ld = a.ld;
  const cplx* __restrict__ Uprev = a.U + (int64_t)(i - 1) * ld;
  const cplx* __restrict__ Ucur = a.U + (int64_t)i * ld;
  cplx* w = a.w;
  const cplx cf = do_axpy ? a.coef[i - 1] : make_double2(0.0, 0.0);

  cplx dacc = make_double2(0.0, 0.0);
  double nacc = 0.0;
  constexpr int ROWS = kWarp * R;
  const int64_t nchunks = (a.n + ROWS - 1) / ROWS;
  for (int64_t q = (int64_t)blockIdx.x * nwarps + warp; q < nchunks;
       q += (int64_t)gridDim.x * nwarps) {
    const int64_t base = q * ROWS + lane;
    cplx wv[R], up[R], uc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool ok = base + r * kWarp < a.n;
      wv[r] = ok ? ld_plain(w + base + r * kWarp) : make_double2(0.0, 0.0);
      up[r] = (ok && do_axpy) ? ld_stream(Uprev + base + r * kWarp) : make_double2(0.0, 0.0);
      uc[r] = (ok && do_dot) ? ld_stream(Ucur + base + r * kWarp) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool ok = base + r * kWarp < a.n;
      if (do_axpy) {
        subax<REAL>(wv[r], up[r], cf);
        if (ok) st_stream(w + base + r * kWarp, wv[r]);
      }
      if (do_dot) dotacc<REAL>(dacc, uc[r], wv[r]);
      if (norm_in || !do_dot) {
        nacc = fma(wv[r].x, wv[r].x, nacc);
        nacc = fma(wv[r].y, wv[r].y, nacc);
      }
    }
  }

  __shared__ double s_w[32 * 3];
  __shared__ int s_last;
  cplx ds = warp_sum(dacc);
  double ns = warp_sum(nacc);
  if (lane == 0) {
    s_w[warp * 3] = ds.x;
    s_w[warp * 3 + 1] = ds.y;
    s_w[warp * 3 + 2] = ns;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0.0, y = 0.0, t = 0.0;
    for (int k = 0; k < nwarps; ++k) {
      x += s_w[k * 3];
      y += s_w[k * 3 + 1];
      t += s_w[k * 3 + 2];
    }
    a.part[blockIdx.x] = make_double2(x, y);
    a.npart[blockIdx.x] = t;
  }
  if (!last_block_ticket(a.ticket, gridDim.x, &s_last)) return;

  __shared__ double s_red[3];
  if (warp == 0) {
    cplx g = sum_partials(a.part, gridDim.x, lane);
    double t = sum_partials(a.npart, gridDim.x, lane);
    if (lane == 0) {
      s_red[0] = g.x;
      s_red[1] = g.y;
      s_red[2] = t;
    }
  }
  __syncthreads();
  peer_allreduce(a.comm, s_red, 3, ctl);
  __syncthreads();
  if (do_dot) {
    // finish_dots expects red[2*ncols] to hold the norm: here ncols == 1
    finish_dots(a, i, 1, s_red, norm_in && a.round == 1);
  } else if (threadIdx.x == 0) {
    finish_norm(a, s_red[2]);
  }
}

// ------------------------------------------------------------------ peer barrier
// All ranks' streams meet here: everything a rank enqueued before this kernel (restart
// update, column uploads) is complete and visible before any rank runs what follows (the
// first halo gather of an expansion).  One 32-thread block, one NVLink round trip.
// The token carries this rank's storage mode: the ranks must agree on it (halo reads use the
// reader's element size on the owner's memory), so a mixed sum is a communicator error.
__global__ void peer_barrier_kernel(PeerComm pc, StepCtl* ctl, int real_mode) {
  __shared__ double token[1];
  if (threadIdx.x == 0) token[0] = real_mode ? 1.0 : 0.0;
  __syncthreads();
  __threadfence_system();
  peer_allreduce(pc, token, 1, ctl);
  __syncthreads();
  if (threadIdx.x == 0 && token[0] != 0.0 && token[0] != (double)pc.nranks) {
    ctl->comm_error = 2;
    ctl->stop = 1;
  }
}
cudaError_t launch_peer_barrier(const PeerComm& pc, StepCtl* ctl, int real_mode, cudaStream_t st) {
  if (pc.nranks <= 1) return cudaSuccess;
  peer_barrier_kernel<<<1, 32, 0, st>>>(pc, ctl, real_mode);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ launchers
static int pick_grid(int64_t work_items, int blocks_per_sm, int num_sms, int cap) {
  int64_t g = (int64_t)num_sms * blocks_per_sm;
  if (g > work_items) g = work_items;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Column-tile width for pass 1: the narrowest tile that keeps the block at <= 16 warps
// while giving at least 4 warps to hide latency.
static void pass1_shape(int c, int* ct, int* warps) {
  int t = (c + 7) / 8;  // aim for ~8 warps
  if (t < 1) t = 1;
  if (t > 8) t = 8;
  int w = (c + t - 1) / t;
  if (w < 1) w = 1;
  *ct = t;
  *warps = w;
}

template <typename K>
static int resident_blocks(K kernel, int threads, size_t smem, int* cache);

template <int CT, int R, bool REAL>
static cudaError_t launch_pass1_tr(const OrthoArgs& a, int warps, int num_sms, cudaStream_t st,
                                   int grid_mult) {
  OrthoArgs args = a;
  const int threads = warps * kWarp;
  const int64_t nchunks = (a.n + kWarp * R - 1) / (kWarp * R);
  const size_t smem = sizeof(double) * (2 * a.p1_ncols + 2);
  static int occ[17] = {0};
  const int bps = grid_mult > 0 ? grid_mult
                                : resident_blocks(cgs_pass1_kernel<CT, R, REAL>, threads, smem, &occ[warps]);
  const int grid = pick_grid(nchunks, bps, num_sms, a.grid_cap);
  cgs_pass1_kernel<CT, R, REAL><<<grid, threads, smem, st>>>(args);
  return cudaGetLastError();
}
template <int CT, int R>
static cudaError_t launch_pass1_t(const OrthoArgs& a, int warps, int num_sms, cudaStream_t st,
                                  int grid_mult) {
  return a.real ? launch_pass1_tr<CT, R, true>(a, warps, num_sms, st, grid_mult)
                : launch_pass1_tr<CT, R, false>(a, warps, num_sms, st, grid_mult);
}

static cudaError_t launch_pass1_group(const OrthoArgs& a, int num_sms, cudaStream_t st,
                                      int grid_mult) {
  int ct, warps;
  pass1_shape(a.p1_ncols, &ct, &warps);
  if (warps > 16) return cudaErrorInvalidValue;
  switch (ct) {
    case 1: return launch_pass1_t<1, 4>(a, warps, num_sms, st, grid_mult);
    case 2: return launch_pass1_t<2, 4>(a, warps, num_sms, st, grid_mult);
    case 3: return launch_pass1_t<3, 4>(a, warps, num_sms, st, grid_mult);
    case 4: return launch_pass1_t<4, 4>(a, warps, num_sms, st, grid_mult);
    case 5: return launch_pass1_t<5, 2>(a, warps, num_sms, st, grid_mult);
    case 6: return launch_pass1_t<6, 2>(a, warps, num_sms, st, grid_mult);
    case 7: return launch_pass1_t<7, 2>(a, warps, num_sms, st, grid_mult);
    default: return launch_pass1_t<8, 2>(a, warps, num_sms, st, grid_mult);
  }
}

// One launch covers at most 128 columns (16 warps x 8 columns); wider bases are swept in
// balanced column groups, each with its own reduction (the norm rides on the first group).
cudaError_t launch_cgs_pass1(const OrthoArgs& a, int num_sms, cudaStream_t st, int grid_mult) {
  const int c = a.ncols;
  const int groups = (c + kPass1MaxCols - 1) / kPass1MaxCols;
  const int per = (c + groups - 1) / groups;
  for (int col0 = 0; col0 < c; col0 += per) {
    OrthoArgs g = a;
    g.p1_col0 = col0;
    g.p1_ncols = c - col0 < per ? c - col0 : per;
    cudaError_t e = launch_pass1_group(g, num_sms, st, grid_mult);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// resident blocks per SM for a kernel / block shape (cached: the query costs microseconds)
template <typename K>
static int resident_blocks(K kernel, int threads, size_t smem, int* cache) {
  if (*cache == 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess ||
        nb < 1)
      nb = 1;
    *cache = nb;
  }
  return *cache;
}

template <int CT, int R, bool PF, bool REAL, int S>
static cudaError_t launch_fused_trs(const OrthoArgs& a, int warps, int num_sms, cudaStream_t st,
                                    int grid_mult) {
  OrthoArgs args = a;
  args.accumulate = 1;  // the dots it produces belong to round 2
  const int threads = warps * kWarp;
  constexpr int ROWS = kWarp * R;
  const int64_t nchunks = (a.n + ROWS - 1) / ROWS;
  size_t smem = sizeof(cplx) * ((size_t)2 * warps * ROWS + (size_t)warps * CT + ROWS);
  if (PF) smem += sizeof(cplx) * ((size_t)S * ROWS + (size_t)S * warps * CT * ROWS);
  const size_t need = sizeof(double) * (2 * a.ncols + 2);
  if (smem < need) smem = need;
  static int occ[17] = {0};
  static PerDeviceOnce once;
  if (once.first_use()) {
    cudaFuncSetAttribute(cgs_fused_kernel<CT, R, PF, REAL, S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         200 * 1024);
  }
  const int bps = grid_mult > 0 ? grid_mult
                                : resident_blocks(cgs_fused_kernel<CT, R, PF, REAL, S>, threads, smem,
                                                  &occ[warps]);
  const int grid = pick_grid(nchunks, bps, num_sms, a.grid_cap);
  cgs_fused_kernel<CT, R, PF, REAL, S><<<grid, threads, smem, st>>>(args);
  return cudaGetLastError();
}
template <int CT, int R, bool PF, bool REAL>
static cudaError_t launch_fused_tr(const OrthoArgs& a, int warps, int num_sms, cudaStream_t st,
                                   int grid_mult) {
  // staging depth: as many chunks in flight as fit beside a second resident block (<= ~100 KB)
  if (PF) {
    const size_t per_stage = sizeof(cplx) * (size_t)kWarp * R * ((size_t)warps * CT + 1);
    const int want = a.stages > 0 ? a.stages : 2;  // measured: deeper rings cost a resident block
    if (want >= 4 && 4 * per_stage <= 200 * 1024)
      return launch_fused_trs<CT, R, PF, REAL, 4>(a, warps, num_sms, st, grid_mult);
    if (want == 3 && 3 * per_stage <= 200 * 1024)
      return launch_fused_trs<CT, R, PF, REAL, 3>(a, warps, num_sms, st, grid_mult);
  }
  return launch_fused_trs<CT, R, PF, REAL, 2>(a, warps, num_sms, st, grid_mult);
}
template <int CT, int R, bool PF>
static cudaError_t launch_fused_t(const OrthoArgs& a, int warps, int num_sms, cudaStream_t st,
                                  int grid_mult) {
  return a.real ? launch_fused_tr<CT, R, PF, true>(a, warps, num_sms, st, grid_mult)
                : launch_fused_tr<CT, R, PF, false>(a, warps, num_sms, st, grid_mult);
}

template <int CMAX, bool REAL>
static cudaError_t launch_fused_warp_t(const OrthoArgs& a, int num_sms, cudaStream_t st,
                                       int grid_mult) {
  OrthoArgs args = a;
  args.accumulate = 1;
  const size_t per_warp = sizeof(cplx) * (size_t)2 * (a.ncols + 1) * kWarp;
  int warps = (int)((190 * 1024 - sizeof(cplx) * CMAX) / per_warp);
  if (warps > 8) warps = 8;
  if (warps < 1) return cudaErrorInvalidValue;
  size_t smem = sizeof(cplx) * CMAX + per_warp * warps;
  const size_t need = sizeof(cplx) * CMAX * (warps + 1) + sizeof(double) * (2 * a.ncols + 2);
  if (smem < need) smem = need;
  static PerDeviceOnce once;
  if (once.first_use()) {
    cudaFuncSetAttribute(cgs_fused_warp_kernel<CMAX, REAL>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  const int64_t nchunks = (a.n + kWarp - 1) / kWarp;
  const int64_t nb = (nchunks + warps - 1) / warps;
  const int grid = pick_grid(nb, grid_mult > 0 ? grid_mult : 1, num_sms, a.grid_cap);
  cgs_fused_warp_kernel<CMAX, REAL><<<grid, warps * kWarp, smem, st>>>(args);
  return cudaGetLastError();
}
template <int CMAX>
static cudaError_t launch_fused_warp(const OrthoArgs& a, int num_sms, cudaStream_t st, int gm) {
  return a.real ? launch_fused_warp_t<CMAX, true>(a, num_sms, st, gm)
                : launch_fused_warp_t<CMAX, false>(a, num_sms, st, gm);
}

template <int CT, int R, bool REAL, int MAXT>
static cudaError_t launch_fused_pipe_trm(const OrthoArgs& a, int warps, int num_sms, cudaStream_t st,
                                         int grid_mult) {
  OrthoArgs args = a;
  args.accumulate = 1;
  const int threads = warps * kWarp;
  constexpr int ROWS = kWarp * R;
  const int64_t nchunks = (a.n + ROWS - 1) / ROWS;
  size_t smem = sizeof(cplx) * ((size_t)2 * warps * ROWS + (size_t)warps * CT + 2 * ROWS + 4 * ROWS +
                                (size_t)2 * warps * CT * ROWS) + 4 * sizeof(unsigned long long);
  const size_t need = sizeof(double) * (2 * a.ncols + 2);
  if (smem < need) smem = need;
  static int occ[17] = {0};
  static PerDeviceOnce once;
  if (once.first_use()) {
    cudaFuncSetAttribute(cgs_fused_pipe_kernel<CT, R, REAL, MAXT>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  const int bps = grid_mult > 0 ? grid_mult
                                : resident_blocks(cgs_fused_pipe_kernel<CT, R, REAL, MAXT>, threads, smem,
                                                  &occ[warps]);
  const int grid = pick_grid(nchunks, bps, num_sms, a.grid_cap);
  cgs_fused_pipe_kernel<CT, R, REAL, MAXT><<<grid, threads, smem, st>>>(args);
  return cudaGetLastError();
}
template <int CT, int R, bool REAL>
static cudaError_t launch_fused_pipe_tr(const OrthoArgs& a, int warps, int num_sms, cudaStream_t st,
                                        int grid_mult) {
  // up to 8 warps: a 256-thread bound leaves the registers for two chunks of U per lane
  return warps <= 8 ? launch_fused_pipe_trm<CT, R, REAL, 256>(a, warps, num_sms, st, grid_mult)
                    : launch_fused_pipe_trm<CT, R, REAL, 512>(a, warps, num_sms, st, grid_mult);
}
template <int CT, int R>
static cudaError_t launch_fused_pipe_t(const OrthoArgs& a, int warps, int num_sms, cudaStream_t st,
                                       int grid_mult) {
  return a.real ? launch_fused_pipe_tr<CT, R, true>(a, warps, num_sms, st, grid_mult)
                : launch_fused_pipe_tr<CT, R, false>(a, warps, num_sms, st, grid_mult);
}

// variant 0: cp.async-staged (prefetching) kernel; variant 2: register loads only.
// fused_ct > 0 forces the column-tile width (when the block shape allows it).
cudaError_t launch_cgs_fused(const OrthoArgs& a, int num_sms, cudaStream_t st, int grid_mult,
                             int variant, int fused_ct) {
  if (variant == 4 && a.ncols <= 64) {  // warp-tile kernel
    const int c = a.ncols;
    if (c <= 16) return launch_fused_warp<16>(a, num_sms, st, grid_mult);
    if (c <= 24) return launch_fused_warp<24>(a, num_sms, st, grid_mult);
    if (c <= 32) return launch_fused_warp<32>(a, num_sms, st, grid_mult);
    if (c <= 40) return launch_fused_warp<40>(a, num_sms, st, grid_mult);
    if (c <= 48) return launch_fused_warp<48>(a, num_sms, st, grid_mult);
    return launch_fused_warp<64>(a, num_sms, st, grid_mult);
  }
  int ct, warps;
  pass1_shape(a.ncols, &ct, &warps);
  if (a.ncols >= 20) {
    // measured: 5-column tiles are the sweet spot (fewer partial sums to exchange than with
    // narrower tiles, fewer registers than with wider ones); up to 16 warps per block
    ct = a.ncols <= 80 ? 5 : (a.ncols + 15) / 16;
    warps = (a.ncols + ct - 1) / ct;
  }
  if (fused_ct > 0 && fused_ct <= 8) {
    const int wf = (a.ncols + fused_ct - 1) / fused_ct;
    if (wf <= (fused_ct < 5 ? 8 : 16)) ct = fused_ct, warps = wf;
  }
  if (warps > 16) return cudaErrorInvalidValue;
  // cp.async staging needs 2 x warps x CT x R x 512 B of shared memory: keep it to shapes that
  // still leave two blocks per SM
  const int rr = ct <= 5 ? 2 : 1;
  const size_t stage_bytes = (size_t)2 * warps * ct * rr * kWarp * sizeof(cplx);
  // variant 5: the mbarrier-pipelined sweep
  if (variant == 5 && warps >= 3 && stage_bytes <= 96 * 1024) {
    const int r1 = a.fused_r == 1 ? 1 : rr;
    switch (ct) {
      case 1: return r1 == 1 ? launch_fused_pipe_t<1, 1>(a, warps, num_sms, st, grid_mult)
                             : launch_fused_pipe_t<1, 2>(a, warps, num_sms, st, grid_mult);
      case 2: return r1 == 1 ? launch_fused_pipe_t<2, 1>(a, warps, num_sms, st, grid_mult)
                             : launch_fused_pipe_t<2, 2>(a, warps, num_sms, st, grid_mult);
      case 3: return r1 == 1 ? launch_fused_pipe_t<3, 1>(a, warps, num_sms, st, grid_mult)
                             : launch_fused_pipe_t<3, 2>(a, warps, num_sms, st, grid_mult);
      case 4: return r1 == 1 ? launch_fused_pipe_t<4, 1>(a, warps, num_sms, st, grid_mult)
                             : launch_fused_pipe_t<4, 2>(a, warps, num_sms, st, grid_mult);
      case 5: return r1 == 1 ? launch_fused_pipe_t<5, 1>(a, warps, num_sms, st, grid_mult)
                             : launch_fused_pipe_t<5, 2>(a, warps, num_sms, st, grid_mult);
      case 6: return launch_fused_pipe_t<6, 1>(a, warps, num_sms, st, grid_mult);
      case 7: return launch_fused_pipe_t<7, 1>(a, warps, num_sms, st, grid_mult);
      default: return launch_fused_pipe_t<8, 1>(a, warps, num_sms, st, grid_mult);
    }
  }
  if (variant != 2 && a.fused_r == 1) {  // A/B: one row pair per lane and chunk
    switch (ct) {
      case 3: return launch_fused_t<3, 1, true>(a, warps, num_sms, st, grid_mult);
      case 4: return launch_fused_t<4, 1, true>(a, warps, num_sms, st, grid_mult);
      case 5: return launch_fused_t<5, 1, true>(a, warps, num_sms, st, grid_mult);
      default: break;
    }
  }
  if (variant != 2 && stage_bytes <= 96 * 1024) {
    switch (ct) {
      case 1: return launch_fused_t<1, 2, true>(a, warps, num_sms, st, grid_mult);
      case 2: return launch_fused_t<2, 2, true>(a, warps, num_sms, st, grid_mult);
      case 3: return launch_fused_t<3, 2, true>(a, warps, num_sms, st, grid_mult);
      case 4: return launch_fused_t<4, 2, true>(a, warps, num_sms, st, grid_mult);
      case 5: return launch_fused_t<5, 2, true>(a, warps, num_sms, st, grid_mult);
      case 6: return launch_fused_t<6, 1, true>(a, warps, num_sms, st, grid_mult);
      case 7: return launch_fused_t<7, 1, true>(a, warps, num_sms, st, grid_mult);
      default: return launch_fused_t<8, 1, true>(a, warps, num_sms, st, grid_mult);
    }
  }
  switch (ct) {
    case 1: return launch_fused_t<1, 4, false>(a, warps, num_sms, st, grid_mult);
    case 2: return launch_fused_t<2, 4, false>(a, warps, num_sms, st, grid_mult);
    case 3: return launch_fused_t<3, 2, false>(a, warps, num_sms, st, grid_mult);
    case 4: return launch_fused_t<4, 2, false>(a, warps, num_sms, st, grid_mult);
    case 5: return launch_fused_t<5, 2, false>(a, warps, num_sms, st, grid_mult);
    case 6: return launch_fused_t<6, 1, false>(a, warps, num_sms, st, grid_mult);
    case 7: return launch_fused_t<7, 1, false>(a, warps, num_sms, st, grid_mult);
    default: return launch_fused_t<8, 1, false>(a, warps, num_sms, st, grid_mult);
  }
}

cudaError_t launch_cgs_pass2(const OrthoArgs& a, int num_sms, cudaStream_t st, int grid_mult) {
  constexpr int R = 2, UC = 4, THREADS = 256;
  const int64_t nchunks = (a.n + kWarp * R - 1) / (kWarp * R);
  const int64_t nblocks = (nchunks + (THREADS / kWarp) - 1) / (THREADS / kWarp);
  const int grid = pick_grid(nblocks, grid_mult > 0 ? grid_mult : 4, num_sms, a.grid_cap);
  const size_t smem = sizeof(cplx) * (a.ncols + 1);
  if (a.real)
    cgs_pass2_kernel<R, UC, true><<<grid, THREADS, smem, st>>>(a);
  else
    cgs_pass2_kernel<R, UC, false><<<grid, THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_mgs_step(const OrthoArgs& a, int i, int num_sms, cudaStream_t st,
                            int grid_mult) {
  constexpr int R = 4, THREADS = 256;
  const int64_t nchunks = (a.n + kWarp * R - 1) / (kWarp * R);
  const int64_t nblocks = (nchunks + (THREADS / kWarp) - 1) / (THREADS / kWarp);
  const int grid = pick_grid(nblocks, grid_mult > 0 ? grid_mult : 4, num_sms, a.grid_cap);
  if (a.real)
    mgs_step_kernel<R, true><<<grid, THREADS, 0, st>>>(a, i);
  else
    mgs_step_kernel<R, false><<<grid, THREADS, 0, st>>>(a, i);
  return cudaGetLastError();
}

}  // namespace ab200
