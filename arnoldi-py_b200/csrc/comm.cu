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
    const cplx* src = a.peer_base[q] + (int64_t)a.col * a.peer_ld[q] + a.src_off[k];
    // peer memory: plain coherent load (it was written by another GPU)
    a.ghost[k] = ld_plain(src);
  }
}

cudaError_t launch_halo_gather(const HaloArgs& a, int num_sms, cudaStream_t st) {
  if (a.nghost <= 0) return cudaSuccess;
  int64_t grid = (a.nghost + 255) / 256;
  if (grid > (int64_t)num_sms * 4) grid = (int64_t)num_sms * 4;
  halo_gather_kernel<<<(int)grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace ab200
