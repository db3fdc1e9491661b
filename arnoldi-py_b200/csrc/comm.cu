// comm.cu -- multi-GPU plumbing of one 8 x B200 box: CUDA-IPC peer mappings and the halo
// gather that precedes each SpMV.
//
// Rows of A and of the basis are block-row sharded, one process per GPU.  Before the
// SpMV of step j every rank needs the entries of v_j = s_j U_j owned by other ranks that
// its columns touch (the halo).  Peers' bases are mapped into this process (cudaIpc), so
// the halo is a gather kernel that loads straight from peer HBM over NVLink -- no staging
// buffer on the owner, no host involvement, no collective.  Ordering comes for free from
// the reductions: a rank starts step j only after it received every rank's partial of the
// last reduction of step j-1, and a rank publishes that partial (system-scope fence +
// flag) only after all its writes to its rows of U_j.
#include "kernels.cuh"

namespace ab200 {

__global__ void __launch_bounds__(256) halo_gather_kernel(HaloArgs a) {
  if (a.ctl != nullptr && a.ctl->stop) return;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a.nghost;
       k += (int64_t)gridDim.x * blockDim.x) {
    int q = 0;
#pragma unroll
    for (int r = 1; r < kMaxRanks; ++r)
      if (r < a.nranks && k >= a.seg_start[r]) q = r;
    // peer memory: plain coherent load (it was written by another GPU)
    if (a.real) {
      // real storage: a column is peer_ld[q] complex slots = 2 * peer_ld[q] doubles
      const double* src = reinterpret_cast<const double*>(a.peer_base[q]) +
                          (int64_t)a.col * a.peer_ld[q] + a.src_off[k];
      double v;
      asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(src) : "memory");
      reinterpret_cast<double*>(a.ghost)[k] = v;
    } else {
      const cplx* src = a.peer_base[q] + (int64_t)a.col * a.peer_ld[q] + a.src_off[k];
      a.ghost[k] = ld_plain(src);
    }
  }
}

// ---------------------------------------------------------------- owner-side push
// For scattered halos (power-law operators: millions of isolated remote entries) pulling
// 16 bytes at a time over NVLink is slow.  The OWNER gathers the entries a peer needs from
// its own HBM (random local reads at HBM speed) and streams them into the peer's ghost
// buffer with coalesced remote stores; the last block then raises this rank's "halo k
// delivered" flag on every peer.  The receiver runs halo_wait_kernel before its SpMV.
__global__ void __launch_bounds__(256) halo_push_kernel(HaloPushArgs a) {
  if (a.ctl != nullptr && a.ctl->stop) return;
  const cplx* __restrict__ src = a.U_col;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a.nsend;
       k += (int64_t)gridDim.x * blockDim.x) {
    int r = 0;
#pragma unroll
    for (int q = 1; q < kMaxRanks; ++q)
      if (q < a.nranks && k >= a.send_ptr[q]) r = q;
    if (a.real) {
      const double v = __ldg(reinterpret_cast<const double*>(src) + a.send_idx[k]);
      double* dst = reinterpret_cast<double*>(a.peer_ghost[r]) + a.dst_off[r] + (k - a.send_ptr[r]);
      asm volatile("st.global.f64 [%0], %1;" ::"l"(dst), "d"(v) : "memory");
    } else {
      cplx v = ld_ro(src + a.send_idx[k]);
      cplx* dst = a.peer_ghost[r] + a.dst_off[r] + (k - a.send_ptr[r]);
      asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(dst), "d"(v.x), "d"(v.y) : "memory");
    }
  }
  __shared__ int s_last;
  __threadfence_system();
  if (!last_block_ticket(a.ticket, gridDim.x, &s_last)) return;
  __threadfence_system();
  if (threadIdx.x < a.nranks && threadIdx.x != a.rank) {
    volatile unsigned long long* f = a.peer_hflags[threadIdx.x] + a.rank;
    *f = a.seq;
  }
}

__global__ void halo_wait_kernel(const unsigned long long* hflags, unsigned need_mask,
                                 unsigned long long seq, StepCtl* ctl) {
  if (ctl != nullptr && ctl->stop) return;
  const int q = threadIdx.x;
  if (q < kMaxRanks && ((need_mask >> q) & 1u)) {
    volatile const unsigned long long* f = hflags + q;
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    while (*f < seq) {
      if ((++spins & 0x3ffu) == 0 && global_ns() - t0 > kPeerTimeoutNs) {
        if (ctl) ctl->comm_error = 1, ctl->stop = 1;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
}

cudaError_t launch_halo_push(const HaloPushArgs& a, int num_sms, cudaStream_t st) {
  int64_t grid = (a.nsend + 255) / 256;
  if (grid > (int64_t)num_sms * 8) grid = (int64_t)num_sms * 8;
  if (grid < 1) grid = 1;  // even with nothing to send the flags must be raised
  halo_push_kernel<<<(int)grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_halo_wait(const unsigned long long* hflags, unsigned need_mask,
                             unsigned long long seq, StepCtl* ctl, cudaStream_t st) {
  halo_wait_kernel<<<1, 32, 0, st>>>(hflags, need_mask, seq, ctl);
  return cudaGetLastError();
}

cudaError_t launch_halo_gather(const HaloArgs& a, int num_sms, cudaStream_t st) {
  if (a.nghost <= 0) return cudaSuccess;
  int64_t grid = (a.nghost + 255) / 256;
  if (grid > (int64_t)num_sms * 4) grid = (int64_t)num_sms * 4;
  halo_gather_kernel<<<(int)grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace ab200
