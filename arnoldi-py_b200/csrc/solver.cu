// solver.cu -- the C ABI declared in include/arnoldi_b200.h.
//
// One ab200_solver owns, on one GPU: the Krylov basis (column-major complex128,
// un-normalised columns + a lazy scale per column), the CSR block of A, the device
// copy of H, the reduction scratch and a control block.  ab200_expand enqueues a
// whole Arnoldi expansion (decomposition.py:56-66) without a host round trip and
// synchronises once at the end to hand the new columns of H to the host driver.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <thread>
#include <chrono>
#include <vector>

#include "../../include/arnoldi_b200.h"
#include "kernels.cuh"

using namespace ab200;

// ---------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return set_err(e__ == cudaErrorMemoryAllocation ? AB200_ENOMEM : AB200_ECUDA,       \
                     "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,   \
                     __LINE__);                                                           \
  } while (0)
#define REQUIRE(cond, ...) \
  do {                     \
    if (!(cond)) return set_err(AB200_EINVAL, __VA_ARGS__); \
  } while (0)

// ---------------------------------------------------------------------------- solver
enum KClass { K_SPMV = 0, K_PASS1, K_PASS2, K_FUSED, K_MGS, K_RESTART, K_OTHER, K_NCLASS };

struct TimedLaunch {
  int cls;
  int step;   // Arnoldi step j (or -1)
  int round;  // 1, 2 or 0
  double bytes;
  cudaEvent_t a, b;
  bool a_shared;   // `a` is the end event of the launch enqueued just before (not owned)
  int executed;    // -1 until the expansion's flags are known
};

struct ab200_solver {
  int device = 0;
  int num_sms = kNumSMsB200;
  int64_t n_global = 0, row0 = 0, n = 0, ld = 0;
  int max_dim = 0;
  cudaStream_t stream = nullptr;

  cplx* V = nullptr;       // ld x (max_dim + 1)
  cplx* wtmp = nullptr;    // n-length scratch (stand-alone ortho / spmv output)
  cplx* xtmp = nullptr;    // n_global-length scratch (stand-alone spmv input)
  double* scale = nullptr;  // [max_dim + 1]
  cplx* Hdev = nullptr;    // (max_dim + 1) x max_dim column-major
  cplx* hscratch = nullptr;  // [max_dim + 2] h column for the stand-alone ortho
  cplx* coef = nullptr;    // [max_dim + 1]
  cplx* part = nullptr;    // [(max_dim + 1) * grid_cap]
  double* npart = nullptr;  // [grid_cap]
  unsigned* ticket = nullptr;
  StepCtl* ctl = nullptr;
  int* step_round2 = nullptr;  // [max_dim] 1 when step j ran the second round
  cplx* qdev = nullptr;    // [max_dim * max_dim] restart coefficients
  int grid_cap = 0;

  // pinned host mirrors
  cplx* h_H = nullptr;
  double* h_scale = nullptr;
  StepCtl* h_ctl = nullptr;
  int* h_step_round2 = nullptr;
  cplx* h_q = nullptr;

  // CSR block
  void* indptr = nullptr;
  int indptr_bits = 32;
  int32_t* indices = nullptr;
  void* values = nullptr;
  int value_kind = AB200_F64;
  int64_t nnz = -1;
  int64_t* rowblk = nullptr;
  int64_t* rowblk_ring = nullptr;   // second tiling (short tiles, one per warp) for the ring kernels
  int tile_ring = 0, nblk_ring = 0;
  int nblk = 0;
  int tile = 0;
  int spmv_threads = 128;
  int spmv_algo = AB200_SPMV_AUTO;
  int spmv_stages = 3, spmv_rp_cap = 0;
  int spmv_window = 0, spmv_win_cap = 0, spmv_win_half = 0;   // skewed rows: x ring in smem
  int spmv_locality_pm = 0;                // per-mille of sampled entries within the window
  int max_row_len = 0;
  // user-supplied device operator (ab200_set_operator) instead of a CSR block
  ab200_apply_fn op_fn = nullptr;
  void* op_user = nullptr;
  cplx* ghost = nullptr;
  int64_t n_local_cols = 0;

  PeerComm comm;
  // multi-GPU state
  int rank = 0, nranks = 1;
  int64_t row_starts[kMaxRanks + 1] = {0};
  double* slots = nullptr;                 // my receive area: [2][kMaxRanks][slot_doubles] x 2 packets
  unsigned long long* flags = nullptr;     // my flags: [2][kMaxRanks]
  unsigned long long* seq = nullptr;       // exchange counter
  void* peer_V[kMaxRanks] = {nullptr};     // IPC mappings (nullptr for self / unused)
  void* peer_slots[kMaxRanks] = {nullptr};
  void* peer_flags[kMaxRanks] = {nullptr};
  int64_t peer_ld[kMaxRanks] = {0};
  int64_t* ghost_off = nullptr;            // [nghost] offsets inside the owner's block
  int64_t nghost = 0;
  int64_t seg_start[kMaxRanks + 1] = {0};
  // owner-side push of the halo
  bool push = false;
  unsigned long long* hflags = nullptr;        // my delivery flags [kMaxRanks]
  unsigned long long hseq = 0;
  void* peer_ghost[kMaxRanks] = {nullptr};
  void* peer_hflags[kMaxRanks] = {nullptr};
  int64_t* send_idx = nullptr;
  int64_t send_ptr[kMaxRanks + 1] = {0};
  int64_t dst_off[kMaxRanks] = {0};
  unsigned* push_ticket = nullptr;

  // options
  int opt_grid_mult = 0, opt_restart_variant = 0, opt_ortho_variant = 0, opt_spmv_tile = 0,
      opt_fused_ct = 0, opt_spmv_threads = 0, opt_fused_stages = 0, opt_fused_r = 0, opt_spmv_variant = 0,
      opt_spmv_stages = 0, opt_spmv_bps = 0, opt_halo_fold = 1, opt_spmv_window = 0,
      opt_spmv_win_half = 0, opt_spmv_ring_warps = 0;
  bool disconnected = false;
  // download path: two pinned bounce buffers + a copy stream (ab200_get_columns)
  void* bounce[2] = {nullptr, nullptr};
  size_t bounce_bytes = 0;
  cudaStream_t stream2 = nullptr;
  cudaEvent_t bounce_ev[2] = {nullptr, nullptr};

  // stats
  bool timing = false;
  std::vector<TimedLaunch> pending;
  std::vector<cudaEvent_t> pool;
  ab200_stats st;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  // DGKS history: did the second round run on most steps of the last expansion?  Decides
  // between the 3-sweep schedule (pass 1, fused, pass 2) and the 2(+2)-sweep one.
  bool dgks_hot = false;
  // Real storage: while A is float64 and every column uploaded / every Q applied so far had zero
  // imaginary parts, the basis is provably real and is kept as float64 (half the bytes, bit-
  // identical real parts).  The first complex input converts it in place, once.
  bool real_mode = true;
  bool pristine = true;   // no column has been written yet
  cudaEvent_t chain_event = nullptr;   // end event of the launch enqueued last, while nothing followed it
  bool pooled = false;   // big buffers come from the device's stream-ordered pool (single-GPU blocks)
};

static cudaEvent_t get_event(ab200_solver* s) {
  if (!s->pool.empty()) {
    cudaEvent_t e = s->pool.back();
    s->pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct LaunchScope {
  ab200_solver* s;
  TimedLaunch t;
  bool on;
  LaunchScope(ab200_solver* s_, int cls, int step, int round, double bytes) : s(s_) {
    on = s->timing;
    t.cls = cls;
    t.step = step;
    t.round = round;
    t.bytes = bytes;
    t.a_shared = false;
    t.executed = -1;
    if (on) {
      // consecutive launches share their boundary event: one record per launch instead of two
      // (event records sit on the stream between the kernels: ~3 us each at this kernel size)
      if (s->chain_event != nullptr) {
        t.a = s->chain_event;
        t.a_shared = true;
      } else {
        t.a = get_event(s);
        cudaEventRecord(t.a, s->stream);
      }
      t.b = get_event(s);
    } else {
      t.a = t.b = nullptr;
    }
  }
  ~LaunchScope() {
    if (on) {
      cudaEventRecord(t.b, s->stream);
      s->chain_event = t.b;
    }
    s->pending.push_back(t);
    s->st.kernel_launches += 1;
  }
};
// anything enqueued outside a LaunchScope ends the chain of shared boundary events
static inline void break_chain(ab200_solver* s) { s->chain_event = nullptr; }

// The multi-GB buffers (basis, CSR arrays) of a solver that owns the WHOLE operator come from the
// device's stream-ordered memory pool, with the pool told to keep what is freed: cudaFree of an
// 11 GB basis was measured at 0.08-1.2 s (it hands the pages back to the OS), a large share of a
// short solve; the next solver of the process gets the memory back in microseconds.  A block of a
// sharded operator keeps cudaMalloc: its basis is exported over CUDA IPC, which pool memory is not.
static cudaError_t big_alloc(ab200_solver* s, void** p, size_t bytes) {
  if (!s->pooled) return cudaMalloc(p, bytes);
  return cudaMallocAsync(p, bytes, s->stream);
}
template <typename T>
static cudaError_t big_alloc(ab200_solver* s, T** p, size_t bytes) {
  return big_alloc(s, reinterpret_cast<void**>(p), bytes);
}
static void big_free(ab200_solver* s, void* p) {
  if (p == nullptr) return;
  if (s->pooled && s->stream)
    cudaFreeAsync(p, s->stream);
  else
    cudaFree(p);
}

// Mark which pending launches really ran (needs the expansion's flags: cheap, no CUDA call) and,
// when `flush` or the backlog is large, read their event times and fold them into the stats.
// The stream must be idle.
static void resolve_pending(ab200_solver* s, const int* step_round2 /* may be null */, bool flush = false) {
  break_chain(s);
  for (auto& t : s->pending) {
    if (t.executed >= 0) continue;
    bool executed = true;
    if (t.round == 2 && t.step >= 0 && step_round2 != nullptr && !step_round2[t.step])
      executed = false;
    if (t.step >= 0 && s->h_ctl && s->h_ctl->stop && t.step > s->h_ctl->broke_at) executed = false;
    t.executed = executed ? 1 : 0;
  }
  // (bounded backlog: every pending launch holds a CUDA event, and events are slow to create and
  //  destroy -- 4096 of them cost 0.9 s at teardown)
  if (!flush && s->pending.size() < 512) return;
  for (auto& t : s->pending) {
    const bool executed = t.executed != 0;
    float ms = 0.f;
    if (t.a) {
      cudaEventElapsedTime(&ms, t.a, t.b);
      if (!t.a_shared) s->pool.push_back(t.a);
      s->pool.push_back(t.b);
    }
    if (!executed) continue;
    switch (t.cls) {
      case K_SPMV:
        s->st.spmv_ms += ms, s->st.spmv_bytes += t.bytes, s->st.spmv_launches++;
        break;
      case K_PASS1:
        s->st.ortho_pass1_ms += ms, s->st.ortho_pass1_bytes += t.bytes,
            s->st.ortho_pass1_launches++;
        break;
      case K_PASS2:
        s->st.ortho_pass2_ms += ms, s->st.ortho_pass2_bytes += t.bytes,
            s->st.ortho_pass2_launches++;
        break;
      case K_FUSED:
        s->st.ortho_fused_ms += ms, s->st.ortho_fused_bytes += t.bytes,
            s->st.ortho_fused_launches++;
        break;
      case K_MGS:
        s->st.mgs_ms += ms, s->st.mgs_bytes += t.bytes, s->st.mgs_launches++;
        break;
      case K_RESTART:
        s->st.restart_ms += ms, s->st.restart_bytes += t.bytes, s->st.restart_launches++;
        break;
      default:
        break;
    }
  }
  s->pending.clear();
}

__global__ void init_ctl_kernel(StepCtl* ctl, double* scale, int nscale, bool reset_counters) {
  if (threadIdx.x == 0) {
    ctl->stop = 0;
    ctl->broke_at = -1;
    ctl->round2 = 0;
    ctl->comm_error = 0;
    if (reset_counters) {
      ctl->rounds_total = 0;
      ctl->second_total = 0;
      ctl->steps_total = 0;
    }
    ctl->nrm0sq = 0.0;
    ctl->beta = 0.0;
  }
  if (scale != nullptr)
    for (int i = threadIdx.x; i < nscale; i += blockDim.x) scale[i] = 1.0;
}
__global__ void set_scale_kernel(double* scale, int col0, int ncols, double v) {
  for (int i = threadIdx.x; i < ncols; i += blockDim.x) scale[col0 + i] = v;
}

static inline void* col_ptr(ab200_solver* s, int j) {
  return s->real_mode ? static_cast<void*>(reinterpret_cast<double*>(s->V) + (size_t)j * s->ld)
                      : static_cast<void*>(s->V + (size_t)j * s->ld);
}
static inline double elem_bytes(const ab200_solver* s) { return s->real_mode ? 8.0 : 16.0; }

// real -> complex storage, in place, columns from the highest down (the complex image of
// column i >= 1 lies beyond its real image and only covers real columns already converted;
// column 0 goes through the scratch vector)
static int switch_to_complex(ab200_solver* s) {
  if (!s->real_mode) return AB200_OK;
  if (!s->pristine) {
    double* vr = reinterpret_cast<double*>(s->V);
    for (int i = s->max_dim; i >= 1; --i)
      CU(launch_unpack_real(vr + (size_t)i * s->ld, s->V + (size_t)i * s->ld, s->n, 1.0, s->num_sms,
                            s->stream));
    CU(cudaMemcpyAsync(s->wtmp, vr, sizeof(double) * (size_t)s->n, cudaMemcpyDeviceToDevice,
                       s->stream));
    CU(launch_unpack_real(reinterpret_cast<const double*>(s->wtmp), s->V, s->n, 1.0, s->num_sms,
                          s->stream));
    s->st.kernel_launches += s->max_dim + 1;
  }
  s->real_mode = false;
  return AB200_OK;
}

static OrthoArgs make_ortho_args(ab200_solver* s, cplx* w, int ncols, int j, double tol, double eta,
                                 cplx* hcol, int finalize) {
  OrthoArgs a;
  a.U = s->V;
  a.w = w;
  a.real = s->real_mode ? 1 : 0;
  a.n = s->real_mode ? (s->n + 1) / 2 : s->n;
  a.ld = s->real_mode ? s->ld / 2 : s->ld;
  a.ncols = ncols;
  a.p1_col0 = 0;
  a.p1_ncols = ncols;
  a.j = j;
  a.round = 1;
  a.accumulate = 0;
  a.finalize = finalize;
  a.grid_cap = s->grid_cap;
  a.stages = s->opt_fused_stages;
  a.fused_r = s->opt_fused_r;
  a.tol = tol;
  a.eta = eta;
  a.scale = s->scale;
  a.hcol = hcol;
  a.coef = s->coef;
  a.part = s->part;
  a.npart = s->npart;
  a.ticket = s->ticket;
  a.ctl = s->ctl;
  a.step_flag = nullptr;
  a.comm = s->comm;
  return a;
}

// enqueue one orthogonalisation (both possible rounds) of w against U[:, :ncols]
static int enqueue_ortho(ab200_solver* s, OrthoArgs a, int ortho_kind) {
  const double nb = elem_bytes(s) * (double)s->n;
  const int c = a.ncols;
  const bool fuse = (s->opt_ortho_variant == 0 ? s->dgks_hot : s->opt_ortho_variant != 1) &&
                    c <= kPass1MaxCols;  // the fused sweep holds all c columns in one block
  if (ortho_kind == AB200_ORTHO_CGS2 && fuse) {
    // default CGS2/DGKS schedule: 3 sweeps when the DGKS test fires, 2 when it does not
    a.round = 1;
    a.accumulate = 0;
    {
      LaunchScope ls(s, K_PASS1, a.j, 1, nb * (c + 1));
      CU(launch_cgs_pass1(a, s->num_sms, s->stream, s->opt_grid_mult));
    }
    {
      LaunchScope ls(s, K_FUSED, a.j, 1, nb * (c + 2));
      CU(launch_cgs_fused(a, s->num_sms, s->stream, s->opt_grid_mult, s->opt_ortho_variant,
                          s->opt_fused_ct));
    }
    a.round = 2;
    a.accumulate = 1;
    {
      LaunchScope ls(s, K_PASS2, a.j, 2, nb * (c + 2));
      CU(launch_cgs_pass2(a, s->num_sms, s->stream, s->opt_grid_mult));
    }
    return AB200_OK;
  }
  for (int round = 1; round <= 2; ++round) {
    a.round = round;
    a.accumulate = (round == 2);
    if (ortho_kind == AB200_ORTHO_CGS2) {
      {
        LaunchScope ls(s, K_PASS1, a.j, round, nb * (c + 1));
        CU(launch_cgs_pass1(a, s->num_sms, s->stream, s->opt_grid_mult));
      }
      {
        LaunchScope ls(s, K_PASS2, a.j, round, nb * (c + 2));
        CU(launch_cgs_pass2(a, s->num_sms, s->stream, s->opt_grid_mult));
      }
    } else {
      for (int i = 0; i <= c; ++i) {
        // kernel i touches w (read, and written when i > 0), U_{i-1} and U_i
        double vecs = 1.0 + (i > 0 ? 2.0 : 0.0) + (i < c ? 1.0 : 0.0);
        LaunchScope ls(s, K_MGS, a.j, round, nb * vecs);
        CU(launch_mgs_step(a, i, s->num_sms, s->stream, s->opt_grid_mult));
      }
    }
  }
  return AB200_OK;
}

// ---------------------------------------------------------------------------- ABI
extern "C" {

int ab200_abi_version(void) { return AB200_ABI_VERSION; }
const char* ab200_last_error(void) { return g_err; }

int ab200_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return set_err(AB200_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
}

// Close every mapping of peer memory.  Multi-GPU teardown is two-phase: every rank disconnects,
// the ranks meet (host barrier), and only then ab200_destroy frees the buffers the peers had
// mapped -- freeing an exported allocation while an importer still maps it is undefined.
int ab200_comm_disconnect(ab200_solver* s) {
  REQUIRE(s != nullptr, "solver is null");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  if (s->stream) CU(cudaStreamSynchronize(s->stream));
  for (int r = 0; r < kMaxRanks; ++r) {
    if (s->peer_V[r]) cudaIpcCloseMemHandle(s->peer_V[r]);
    if (s->peer_slots[r]) cudaIpcCloseMemHandle(s->peer_slots[r]);
    if (s->peer_flags[r]) cudaIpcCloseMemHandle(s->peer_flags[r]);
    if (s->peer_ghost[r]) cudaIpcCloseMemHandle(s->peer_ghost[r]);
    if (s->peer_hflags[r]) cudaIpcCloseMemHandle(s->peer_hflags[r]);
    s->peer_V[r] = s->peer_slots[r] = s->peer_flags[r] = s->peer_ghost[r] = s->peer_hflags[r] = nullptr;
  }
  if (s->nranks > 1) s->disconnected = true;  // no further multi-GPU work on this handle
  return AB200_OK;
}

int ab200_destroy(ab200_solver* s) {
  if (!s) return AB200_OK;
  const bool trace = getenv("AB200_TRACE_DESTROY") != nullptr;
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tm = now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const double t = now();
    fprintf(stderr, "[ab200_destroy] %-12s %.4f s\n", what, t - tm);
    tm = t;
  };
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  lap("sync");
  resolve_pending(s, nullptr, true);
  lap("flush");
  for (auto e : s->pool) cudaEventDestroy(e);
  lap("events");
  if (s->t0) cudaEventDestroy(s->t0), cudaEventDestroy(s->t1);
  // imports first (a no-op after ab200_comm_disconnect), then the buffers this rank owns
  for (int r = 0; r < kMaxRanks; ++r) {
    if (s->peer_V[r]) cudaIpcCloseMemHandle(s->peer_V[r]);
    if (s->peer_slots[r]) cudaIpcCloseMemHandle(s->peer_slots[r]);
    if (s->peer_flags[r]) cudaIpcCloseMemHandle(s->peer_flags[r]);
    if (s->peer_ghost[r]) cudaIpcCloseMemHandle(s->peer_ghost[r]);
    if (s->peer_hflags[r]) cudaIpcCloseMemHandle(s->peer_hflags[r]);
  }
  lap("ipc close");
  big_free(s, s->V), big_free(s, s->wtmp), big_free(s, s->xtmp), cudaFree(s->scale), cudaFree(s->Hdev);
  lap("free V");
  cudaFree(s->hscratch), cudaFree(s->coef), cudaFree(s->part), cudaFree(s->npart);
  cudaFree(s->ticket), cudaFree(s->ctl), cudaFree(s->step_round2), cudaFree(s->qdev);
  big_free(s, s->indptr), big_free(s, s->indices), big_free(s, s->values), cudaFree(s->rowblk);
  cudaFree(s->rowblk_ring);
  if (s->pooled && s->stream) cudaStreamSynchronize(s->stream);
  cudaFree(s->ghost), cudaFree(s->ghost_off);
  cudaFree(s->slots), cudaFree(s->flags), cudaFree(s->seq);
  cudaFree(s->hflags), cudaFree(s->send_idx), cudaFree(s->push_ticket);
  lap("free rest");
  if (s->bounce[0]) cudaFreeHost(s->bounce[0]);
  if (s->bounce[1]) cudaFreeHost(s->bounce[1]);
  lap("free bounce");
  if (s->stream2) cudaStreamDestroy(s->stream2);
  cudaFreeHost(s->h_H), cudaFreeHost(s->h_scale), cudaFreeHost(s->h_ctl);
  cudaFreeHost(s->h_step_round2), cudaFreeHost(s->h_q);
  if (s->stream) cudaStreamDestroy(s->stream);
  lap("host+stream");
  delete s;
  return AB200_OK;
}

int ab200_create(ab200_solver** out, int device, int64_t n_global, int64_t row0,
                 int64_t nrows_local, int max_dim) {
  REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  REQUIRE(n_global > 0 && nrows_local > 0 && row0 >= 0 && row0 + nrows_local <= n_global,
          "bad row block: n_global=%lld row0=%lld nrows_local=%lld", (long long)n_global,
          (long long)row0, (long long)nrows_local);
  REQUIRE(max_dim >= 1 && max_dim <= kMaxDim, "max_dim must be in [1, %d], got %d", kMaxDim, max_dim);
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  REQUIRE(device >= 0 && device < ndev, "device %d not available (%d visible)", device, ndev);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return set_err(AB200_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                   device, prop.major, prop.minor);

  ab200_solver* s = new ab200_solver();
  memset(&s->st, 0, sizeof(s->st));
  memset(&s->comm, 0, sizeof(s->comm));
  s->comm.nranks = 1;
  s->device = device;
  s->num_sms = prop.multiProcessorCount;
  s->n_global = n_global;
  s->row0 = row0;
  s->n = nrows_local;
  s->n_local_cols = nrows_local;
  s->ld = (nrows_local + 15) / 16 * 16;  // columns start on 128-byte lines in both storage modes
  s->max_dim = max_dim;
  s->grid_cap = s->num_sms * 8;
  const int md1 = max_dim + 1;
#define CUX(call)                                   \
  do {                                              \
    cudaError_t e__ = (call);                       \
    if (e__ != cudaSuccess) {                       \
      int rc__ = set_err(e__ == cudaErrorMemoryAllocation ? AB200_ENOMEM : AB200_ECUDA, \
                         "%s failed: %s", #call, cudaGetErrorString(e__));              \
      ab200_destroy(s);                             \
      return rc__;                                  \
    }                                               \
  } while (0)
  CUX(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  if (nrows_local == n_global) {
    cudaMemPool_t pool = nullptr;
    unsigned long long keep = ~0ull;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess &&
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess)
      s->pooled = true;
    else
      cudaGetLastError();
  }
  CUX(big_alloc(s, &s->V, sizeof(cplx) * (size_t)s->ld * md1));
  CUX(big_alloc(s, &s->wtmp, sizeof(cplx) * (size_t)s->ld));
  CUX(cudaMalloc(&s->scale, sizeof(double) * md1));
  CUX(cudaMalloc(&s->Hdev, sizeof(cplx) * (size_t)md1 * max_dim));
  CUX(cudaMalloc(&s->hscratch, sizeof(cplx) * (md1 + 1)));
  CUX(cudaMalloc(&s->coef, sizeof(cplx) * md1));
  CUX(cudaMalloc(&s->part, sizeof(cplx) * (size_t)md1 * s->grid_cap));
  CUX(cudaMalloc(&s->npart, sizeof(double) * s->grid_cap));
  CUX(cudaMalloc(&s->ticket, sizeof(unsigned)));
  CUX(cudaMalloc(&s->ctl, sizeof(StepCtl)));
  CUX(cudaMalloc(&s->step_round2, sizeof(int) * max_dim));
  CUX(cudaMalloc(&s->qdev, sizeof(cplx) * (size_t)max_dim * max_dim));
  CUX(cudaMallocHost(&s->h_H, sizeof(cplx) * (size_t)md1 * max_dim));
  CUX(cudaMallocHost(&s->h_scale, sizeof(double) * md1));
  CUX(cudaMallocHost(&s->h_ctl, sizeof(StepCtl)));
  CUX(cudaMallocHost(&s->h_step_round2, sizeof(int) * max_dim));
  CUX(cudaMallocHost(&s->h_q, sizeof(cplx) * (size_t)max_dim * max_dim));
  // krylov_schur.py:42-43: V and H start as zeros
  CUX(cudaMemsetAsync(s->V, 0, sizeof(cplx) * (size_t)s->ld * md1, s->stream));
  CUX(cudaMemsetAsync(s->wtmp, 0, sizeof(cplx) * (size_t)s->ld, s->stream));
  CUX(cudaMemsetAsync(s->Hdev, 0, sizeof(cplx) * (size_t)md1 * max_dim, s->stream));
  CUX(cudaMemsetAsync(s->ticket, 0, sizeof(unsigned), s->stream));
  CUX(cudaMemsetAsync(s->step_round2, 0, sizeof(int) * max_dim, s->stream));
  init_ctl_kernel<<<1, 128, 0, s->stream>>>(s->ctl, s->scale, md1, true);
  CUX(cudaGetLastError());
  CUX(cudaStreamSynchronize(s->stream));
  for (int i = 0; i < md1; ++i) s->h_scale[i] = 1.0;
  memset(s->h_ctl, 0, sizeof(StepCtl));
  s->h_ctl->broke_at = -1;
#undef CUX
  *out = s;
  return AB200_OK;
}

int ab200_set_csr(ab200_solver* s, const void* indptr, int indptr_bits, const int32_t* indices,
                  const void* values, int value_kind, int64_t nnz, int spmv_algo) {
  REQUIRE(s != nullptr, "solver is null");
  REQUIRE(indptr_bits == 32 || indptr_bits == 64, "indptr_bits must be 32 or 64");
  REQUIRE(value_kind == AB200_F64 || value_kind == AB200_C128, "bad value_kind %d", value_kind);
  REQUIRE(nnz >= 0, "nnz < 0");
  REQUIRE(indptr != nullptr && (nnz == 0 || (indices != nullptr && values != nullptr)),
          "null CSR array");
  REQUIRE(spmv_algo >= AB200_SPMV_AUTO && spmv_algo <= AB200_SPMV_MERGE, "bad spmv_algo");
  const int64_t first = indptr_bits == 32 ? ((const int32_t*)indptr)[0] : ((const int64_t*)indptr)[0];
  const int64_t last =
      indptr_bits == 32 ? ((const int32_t*)indptr)[s->n] : ((const int64_t*)indptr)[s->n];
  REQUIRE(first == 0 && last == nnz, "indptr[0]=%lld, indptr[n]=%lld, nnz=%lld: not a CSR block",
          (long long)first, (long long)last, (long long)nnz);
  CU(cudaSetDevice(s->device));
  break_chain(s);
  CU(cudaStreamSynchronize(s->stream));
  if (value_kind == AB200_C128) {
    int rc = switch_to_complex(s);
    if (rc != AB200_OK) return rc;
  }
  big_free(s, s->indptr), big_free(s, s->indices), big_free(s, s->values), cudaFree(s->rowblk);
  cudaFree(s->rowblk_ring);
  s->indptr = s->indices = nullptr, s->values = nullptr, s->rowblk = nullptr, s->rowblk_ring = nullptr;
  s->tile_ring = s->nblk_ring = 0;
  s->nnz = -1;
  s->op_fn = nullptr;
  const size_t ipb = (size_t)indptr_bits / 8, vb = value_kind == AB200_F64 ? 8 : 16;
  // spare entries: the SpMV stages whole 16-byte groups (4 column ids / values, 4 or 2 row
  // pointers), so every array may be read a little past its end
  CU(big_alloc(s, &s->indptr, ipb * (s->n + 1 + 8)));
  CU(big_alloc(s, &s->indices, sizeof(int32_t) * (size_t)(nnz + 8)));
  CU(big_alloc(s, &s->values, vb * (size_t)(nnz + 8)));
  CU(cudaMemsetAsync(static_cast<char*>(s->indptr) + ipb * (s->n + 1), 0, ipb * 8, s->stream));
  CU(cudaMemsetAsync(s->indices + nnz, 0, sizeof(int32_t) * 8, s->stream));
  CU(cudaMemsetAsync(static_cast<char*>(s->values) + vb * (size_t)nnz, 0, vb * 8, s->stream));
  CU(cudaMemcpyAsync(s->indptr, indptr, ipb * (s->n + 1), cudaMemcpyHostToDevice, s->stream));
  if (nnz > 0) {
    CU(cudaMemcpyAsync(s->indices, indices, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice,
                       s->stream));
    CU(cudaMemcpyAsync(s->values, values, vb * (size_t)nnz, cudaMemcpyHostToDevice, s->stream));
  }
  // longest row decides which kernels apply
  CU(cudaMemsetAsync(s->ticket, 0, sizeof(unsigned), s->stream));
  CU(launch_spmv_maxrow(s->indptr, indptr_bits, s->n, reinterpret_cast<int*>(s->ticket), s->stream));
  int maxrow = 0;
  CU(cudaMemcpyAsync(&maxrow, s->ticket, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaMemsetAsync(s->ticket, 0, sizeof(unsigned), s->stream));
  s->max_row_len = maxrow;
  // AB200_SPMV_STREAM is the one-thread-per-row pipeline: it needs short rows
  if (spmv_algo == AB200_SPMV_STREAM && maxrow > 16)
    return set_err(AB200_EINVAL,
                   "AB200_SPMV_STREAM needs rows of at most 16 entries (longest row: %d); use "
                   "AB200_SPMV_VECTOR, AB200_SPMV_MERGE or AB200_SPMV_AUTO", maxrow);
  s->spmv_algo = spmv_algo;
  const bool short_rows = maxrow <= 16 && spmv_algo != AB200_SPMV_VECTOR && spmv_algo != AB200_SPMV_MERGE;
  // nnz tile: about one row per thread of the block, within [512, 4096]
  int threads = short_rows ? 256 : 128;
  if (s->opt_spmv_threads == 128 || s->opt_spmv_threads == 256) threads = s->opt_spmv_threads;
  int tile = s->opt_spmv_tile;
  if (tile <= 0) {
    double avg = (double)nnz / (double)s->n;
    tile = (int)(avg * threads);
    tile = (tile + 127) / 128 * 128;
    if (tile < 512) tile = 512;
    if (tile > 4096) tile = 4096;
    if (!short_rows && tile > 1280) tile = 1280;  // skewed rows: measured best on the power-law operator
  }
  // Skewed rows: gather x from a shared-memory ring that slides along the diagonal
  // (spmv_ring_kernel) when most entries sit near the diagonal -- sampled on the device.
  // AB200_SPMV_MERGE asks for it regardless; AB200_SPMV_VECTOR keeps the plain tile kernel.
  s->spmv_window = 0;
  s->spmv_locality_pm = 0;
  if (!short_rows && spmv_algo != AB200_SPMV_VECTOR && nnz > 0) {
    // ring capacity in x entries (float64).  The cp.async kernel (default) takes what its 31 strip
    // buffers leave of the shared memory (~11 000 entries: +-4.7 k around a round's rows); the
    // register-staged kernel (spmv_variant = 4) measured best with 8192, which leaves ~90 KB of
    // L1 to the gathers that miss the ring (profiles/r02_spmv_ring.md)
    int wcap = s->opt_spmv_window > 0 ? s->opt_spmv_window : (s->opt_spmv_variant == 4 ? 8192 : 11008);
    if (wcap < 256) wcap = 256;
    if (wcap > 16384) wcap = 16384;
    wcap = wcap / 256 * 256;
    // entries kept on each side of a round's rows; the rest of the ring lets the warps drift apart
    const int half = s->opt_spmv_win_half > 0 ? s->opt_spmv_win_half : (wcap >= 12288 ? 4608 : (wcap - 1536) / 2);
    unsigned long long* cnt = nullptr;
    CU(cudaMalloc(&cnt, 2 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), s->stream));
    CU(launch_spmv_locality(s->indptr, indptr_bits, s->indices, s->n, s->n_local_cols, half, cnt,
                            s->stream));
    unsigned long long h[2] = {0, 0};
    CU(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    cudaFree(cnt);
    s->spmv_locality_pm = h[1] ? (int)(1000.0 * (double)h[0] / (double)h[1]) : 0;
    if (spmv_algo == AB200_SPMV_MERGE || s->spmv_locality_pm >= 500) {
      s->spmv_window = 1;
      s->spmv_win_cap = wcap;
      s->spmv_win_half = half;
    }
  }
  // the ring kernels walk their own, shorter tiling: one tile = one warp's strip
  int tile_ring = 0;
  if (s->spmv_window) {
    tile_ring = s->opt_spmv_tile > 0 ? s->opt_spmv_tile : (s->opt_spmv_variant == 4 ? 512 : 224);
    if (tile_ring > 8192) tile_ring = 8192;
    tile_ring = (tile_ring + 7) / 8 * 8;
  }
  if (tile > 8192) tile = 8192;   // 8192 complex entries + column ids = 164 KB of shared memory
  tile = (tile + 7) / 8 * 8;
  s->spmv_threads = threads;
  // measured on lap2d(4096) / mark(4000): two stages of one-row-per-thread tiles are enough
  // (6.47 TB/s; 3 stages 6.30, 4 stages 5.83 -- deeper rings only take L1 away from the gathers)
  s->spmv_stages = s->opt_spmv_stages >= 2 && s->opt_spmv_stages <= 8 ? s->opt_spmv_stages : 2;
  s->spmv_rp_cap = 2 * threads + 8;
  int64_t nblk = (nnz + tile - 1) / tile;
  if (nblk < 1) nblk = 1;
  REQUIRE(nblk < (1ll << 30), "too many SpMV tiles");
  CU(cudaMalloc(&s->rowblk, sizeof(int64_t) * 2 * (size_t)(nblk + 1)));
  CU(launch_spmv_plan(s->indptr, indptr_bits, s->n, nnz, tile, (int)nblk, s->rowblk, s->stream));
  if (tile_ring > 0) {
    int64_t nb2 = (nnz + tile_ring - 1) / tile_ring;
    if (nb2 < 1) nb2 = 1;
    REQUIRE(nb2 < (1ll << 30), "too many SpMV tiles");
    CU(cudaMalloc(&s->rowblk_ring, sizeof(int64_t) * 2 * (size_t)(nb2 + 1)));
    CU(launch_spmv_plan(s->indptr, indptr_bits, s->n, nnz, tile_ring, (int)nb2, s->rowblk_ring, s->stream));
    s->tile_ring = tile_ring;
    s->nblk_ring = (int)nb2;
  }
  CU(cudaStreamSynchronize(s->stream));
  s->indptr_bits = indptr_bits;
  s->value_kind = value_kind;
  s->nnz = nnz;
  s->tile = tile;
  s->nblk = (int)nblk;
  return AB200_OK;
}

// A device operator in place of a CSR block: `fn(user, x, y, n, is_real, stream)` must enqueue
// y = A x on `stream` (device pointers; float64 entries when is_real, else interleaved
// complex128) and return 0.  The generic-operator row of SURVEY.md section 8f
// (README.md:119 of the reference: "LinearOperator support").  Single GPU.
int ab200_set_operator(ab200_solver* s, ab200_apply_fn fn, void* user, int value_kind) {
  REQUIRE(s != nullptr && fn != nullptr, "null argument");
  REQUIRE(value_kind == AB200_F64 || value_kind == AB200_C128, "bad value_kind %d", value_kind);
  if (s->nranks > 1) return set_err(AB200_ESTATE, "device operators are single-GPU");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  CU(cudaStreamSynchronize(s->stream));
  if (value_kind == AB200_C128) {
    int rc = switch_to_complex(s);
    if (rc != AB200_OK) return rc;
  }
  if (!s->xtmp) CU(big_alloc(s, &s->xtmp, sizeof(cplx) * (size_t)s->ld));
  s->op_fn = fn;
  s->op_user = user;
  s->value_kind = value_kind;
  s->nnz = 0;  // "an operator is set"
  return AB200_OK;
}

int ab200_set_columns(ab200_solver* s, int col0, int ncols, const double* host, int64_t ld_host) {
  REQUIRE(s != nullptr && host != nullptr, "null argument");
  REQUIRE(col0 >= 0 && ncols >= 1 && col0 + ncols <= s->max_dim + 1, "column range [%d, %d) out of [0, %d]",
          col0, col0 + ncols, s->max_dim + 1);
  REQUIRE(ld_host >= s->n, "ld_host < nrows_local");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  if (s->real_mode) {
    // stay real only if every imaginary part is exactly zero
    bool all_real = true;
    for (int c = 0; c < ncols && all_real; ++c) {
      const double* col = host + 2 * (size_t)c * ld_host;
      for (int64_t r = 0; r < s->n; ++r)
        if (col[2 * r + 1] != 0.0) {
          all_real = false;
          break;
        }
    }
    if (!all_real) {
      int rc = switch_to_complex(s);
      if (rc != AB200_OK) return rc;
    }
  }
  if (s->real_mode) {
    for (int c = 0; c < ncols; ++c) {
      CU(cudaMemcpyAsync(s->wtmp, host + 2 * (size_t)c * ld_host, sizeof(cplx) * (size_t)s->n,
                         cudaMemcpyHostToDevice, s->stream));
      CU(launch_pack_real(s->wtmp, static_cast<double*>(col_ptr(s, col0 + c)), s->n, s->num_sms,
                          s->stream));
      s->st.kernel_launches += 1;
    }
  } else {
    CU(cudaMemcpy2DAsync(s->V + (size_t)col0 * s->ld, sizeof(cplx) * s->ld, host,
                         sizeof(cplx) * ld_host, sizeof(cplx) * s->n, ncols,
                         cudaMemcpyHostToDevice, s->stream));
  }
  s->pristine = false;
  set_scale_kernel<<<1, 128, 0, s->stream>>>(s->scale, col0, ncols, 1.0);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  for (int i = 0; i < ncols; ++i) s->h_scale[col0 + i] = 1.0;
  return AB200_OK;
}

// Device -> pageable host through two pinned bounce buffers: the D2H copy of chunk i+1 runs
// while host threads move chunk i out of its bounce buffer (a multi-GB result is otherwise
// limited by the driver's pageable path and by first-touch page faults on one thread).
static const size_t kBounceBytes = (size_t)32 << 20;
static int ensure_bounce(ab200_solver* s) {
  if (s->bounce[0]) return AB200_OK;
  CU(cudaMallocHost(&s->bounce[0], kBounceBytes));
  CU(cudaMallocHost(&s->bounce[1], kBounceBytes));
  CU(cudaEventCreateWithFlags(&s->bounce_ev[0], cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&s->bounce_ev[1], cudaEventDisableTiming));
  s->bounce_bytes = kBounceBytes;
  return AB200_OK;
}
static void parallel_copy(char* dst, const char* src, size_t bytes) {
  const int nt = bytes >= ((size_t)8 << 20) ? 4 : 1;
  if (nt == 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> th;
  const size_t per = (bytes / nt + 4095) & ~(size_t)4095;
  for (int t = 0; t < nt; ++t) {
    const size_t o = (size_t)t * per;
    if (o >= bytes) break;
    const size_t len = bytes - o < per ? bytes - o : per;
    th.emplace_back([=] { memcpy(dst + o, src + o, len); });
  }
  for (auto& t : th) t.join();
}
// stream-ordered device -> host copy of `bytes` from `dev` to pageable `host`
static int download(ab200_solver* s, char* host, const char* dev, size_t bytes) {
  if (bytes < ((size_t)4 << 20)) {
    CU(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return AB200_OK;
  }
  int rc = ensure_bounce(s);
  if (rc != AB200_OK) return rc;
  const size_t cb = s->bounce_bytes;
  const size_t nchunk = (bytes + cb - 1) / cb;
  auto issue = [&](size_t i) -> cudaError_t {
    const size_t o = i * cb, len = bytes - o < cb ? bytes - o : cb;
    cudaError_t e = cudaMemcpyAsync(s->bounce[i & 1], dev + o, len, cudaMemcpyDeviceToHost, s->stream);
    if (e != cudaSuccess) return e;
    return cudaEventRecord(s->bounce_ev[i & 1], s->stream);
  };
  CU(issue(0));
  for (size_t i = 0; i < nchunk; ++i) {
    if (i + 1 < nchunk) CU(issue(i + 1));
    CU(cudaEventSynchronize(s->bounce_ev[i & 1]));
    const size_t o = i * cb, len = bytes - o < cb ? bytes - o : cb;
    parallel_copy(host + o, static_cast<const char*>(s->bounce[i & 1]), len);
  }
  return AB200_OK;
}

int ab200_get_columns(ab200_solver* s, int col0, int ncols, double* host, int64_t ld_host) {
  REQUIRE(s != nullptr && host != nullptr, "null argument");
  REQUIRE(col0 >= 0 && ncols >= 1 && col0 + ncols <= s->max_dim + 1, "column range [%d, %d) out of [0, %d]",
          col0, col0 + ncols, s->max_dim + 1);
  REQUIRE(ld_host >= s->n, "ld_host < nrows_local");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  char* out = reinterpret_cast<char*>(host);
  if (s->real_mode) {
    // expand column by column through the scratch vector, lazy scale applied on the way out
    for (int c = 0; c < ncols; ++c) {
      CU(launch_unpack_real(static_cast<const double*>(col_ptr(s, col0 + c)), s->wtmp, s->n,
                            s->h_scale[col0 + c], s->num_sms, s->stream));
      s->st.kernel_launches += 1;
      int rc = download(s, out + sizeof(cplx) * (size_t)c * ld_host, reinterpret_cast<const char*>(s->wtmp),
                        sizeof(cplx) * (size_t)s->n);
      if (rc != AB200_OK) return rc;
    }
    CU(cudaStreamSynchronize(s->stream));
    return AB200_OK;
  }
  // apply the lazy scales in place first (a no-op for columns whose scale is 1)
  CU(launch_materialize(s->V, s->n, s->ld, col0, ncols, s->scale, s->num_sms, s->stream));
  s->st.kernel_launches += 2;
  for (int c = 0; c < ncols; ++c) {
    int rc = download(s, out + sizeof(cplx) * (size_t)c * ld_host,
                      reinterpret_cast<const char*>(s->V + (size_t)(col0 + c) * s->ld),
                      sizeof(cplx) * (size_t)s->n);
    if (rc != AB200_OK) return rc;
  }
  CU(cudaStreamSynchronize(s->stream));
  for (int i = 0; i < ncols; ++i) s->h_scale[col0 + i] = 1.0;
  return AB200_OK;
}

static int enqueue_spmv(ab200_solver* s, const void* x, void* y, const double* xscale, int step,
                        bool in_expand, bool real_vectors = false) {
  const bool real = (in_expand && s->real_mode) || real_vectors;
  if (s->op_fn != nullptr) {
    // device operator: hand it the true vector v_j = s_j U_j in a scratch buffer
    const int64_t n16 = real ? (s->n + 1) / 2 : s->n;
    LaunchScope ls(s, K_SPMV, step, 0, 0.0);
    CU(launch_scaled_copy(static_cast<const cplx*>(x), s->xtmp, n16, xscale, real ? 1 : 0, s->num_sms,
                          s->stream));
    s->st.kernel_launches += 1;
    const int rc = s->op_fn(s->op_user, s->xtmp, y, s->n, real ? 1 : 0, (void*)s->stream);
    if (rc != 0) return set_err(AB200_ECUDA, "device operator callback returned %d", rc);
    return AB200_OK;
  }
  SpmvArgs a;
  memset(&a, 0, sizeof(a));
  a.indptr = s->indptr;
  a.indices = s->indices;
  a.values = s->values;
  a.rowblk = s->rowblk;
  a.x = x;
  a.ghost = s->ghost;
  a.y = y;
  a.xscale = xscale;
  a.n = s->n;
  a.n_local_cols = s->n_local_cols;
  a.nblocks = s->nblk;
  a.tile = s->tile;
  a.threads = s->spmv_threads;
  a.real = real ? 1 : 0;
  a.long_rows = (s->max_row_len > 16 || s->spmv_algo == AB200_SPMV_VECTOR ||
                 s->spmv_algo == AB200_SPMV_MERGE) ? 1 : 0;
  a.variant = s->opt_spmv_variant;
  a.num_sms = s->num_sms;
  a.ctl = in_expand ? s->ctl : nullptr;
  a.stages = s->spmv_stages;
  a.rp_cap = s->spmv_rp_cap;
  a.bps = s->opt_spmv_bps;
  a.nranks = s->nranks;
  // ring kernels: float64 vectors only (complex128 vectors: half as many entries fit the ring and
  // the tile kernel is faster); 2 = cp.async strips (default), 1 = register-staged (spmv_variant 4)
  a.window = 0;
  if (s->spmv_window && a.real && s->rowblk_ring != nullptr) {
    if (s->opt_spmv_variant == 0 || s->opt_spmv_variant == 5) a.window = 2;
    if (s->opt_spmv_variant == 4) a.window = 1;
  }
  if (a.window) {
    a.rowblk = s->rowblk_ring;
    a.nblocks = s->nblk_ring;
    a.tile = s->tile_ring;
  }
  a.contig = s->opt_spmv_variant == 3 ? 1 : 0;
  a.win_cap = s->spmv_win_cap;
  a.win_half = s->spmv_win_half;
  a.ring_warps = s->opt_spmv_ring_warps > 0 ? s->opt_spmv_ring_warps : (a.window == 2 ? 31 : 15);
  const double sv = s->value_kind == AB200_F64 ? 8.0 : 16.0;
  const double eb = a.real ? 8.0 : 16.0;
  const double bytes = (double)s->nnz * (sv + 4.0) + (double)s->n * (s->indptr_bits / 8 + 2.0 * eb) +
                       2.0 * eb * (double)s->nghost;
  LaunchScope ls(s, K_SPMV, step, 0, bytes);
  if (s->push) {
    if (!in_expand) return set_err(AB200_ESTATE, "halo SpMV is only available inside ab200_expand");
    s->hseq += 1;
    HaloPushArgs p;
    p.U_col = static_cast<const cplx*>(x);
    p.real = a.real;
    p.send_idx = s->send_idx;
    for (int r = 0; r <= kMaxRanks; ++r) p.send_ptr[r] = s->send_ptr[r];
    unsigned need = 0;
    for (int r = 0; r < kMaxRanks; ++r) {
      p.dst_off[r] = s->dst_off[r];
      p.peer_ghost[r] = static_cast<cplx*>(s->peer_ghost[r]);
      p.peer_hflags[r] = static_cast<unsigned long long*>(s->peer_hflags[r]);
      if (r < s->nranks && r != s->rank) need |= 1u << r;
    }
    p.nsend = s->send_ptr[s->nranks];
    p.seq = s->hseq;
    p.ticket = s->push_ticket;
    p.rank = s->rank;
    p.nranks = s->nranks;
    p.ctl = s->ctl;
    CU(launch_halo_push(p, s->num_sms, s->stream));
    CU(launch_halo_wait(s->hflags, need, s->hseq, s->ctl, s->stream));
    s->st.kernel_launches += 2;
  } else if (s->nghost > 0) {
    if (!in_expand) return set_err(AB200_ESTATE, "halo SpMV is only available inside ab200_expand");
    // "pull": the halo entries of column `step` are read from the owners' HBM.  Banded
    // operators with short rows read them inside the SpMV itself (no extra launch, no ghost
    // buffer round trip); otherwise a gather kernel fills the ghost buffer first.
    const bool fold = s->opt_halo_fold && !a.long_rows && a.variant == 0;
    const size_t col_bytes = (size_t)step * (a.real ? sizeof(double) : sizeof(cplx));
    if (fold) {
      a.direct_halo = 1;
      a.ghost_off = s->ghost_off;
      for (int r = 0; r < kMaxRanks; ++r) {
        const char* base = r == s->rank ? reinterpret_cast<const char*>(s->V)
                                        : static_cast<const char*>(s->peer_V[r]);
        a.peer_col[r] = base ? base + col_bytes * (size_t)s->peer_ld[r] : nullptr;
      }
      for (int r = 0; r <= kMaxRanks; ++r) a.seg_start[r] = s->seg_start[r];
    } else {
      HaloArgs h;
      for (int r = 0; r < kMaxRanks; ++r) {
        h.peer_base[r] = r == s->rank ? s->V : static_cast<const cplx*>(s->peer_V[r]);
        h.peer_ld[r] = s->peer_ld[r];
      }
      for (int r = 0; r <= kMaxRanks; ++r) h.seg_start[r] = s->seg_start[r];
      h.src_off = s->ghost_off;
      h.ghost = s->ghost;
      h.nghost = s->nghost;
      h.col = step;
      h.real = a.real;
      h.nranks = s->nranks;
      h.ctl = s->ctl;
      CU(launch_halo_gather(h, s->num_sms, s->stream));
      s->st.kernel_launches += 1;
    }
  }
  CU(launch_spmv(a, s->indptr_bits, s->value_kind, s->stream));
  return AB200_OK;
}

int ab200_expand(ab200_solver* s, int start_dim, int end_dim, double tol, double eta,
                 int ortho_kind, double* h_cols, int* n_iter, int* breakdown) {
  REQUIRE(s != nullptr && h_cols != nullptr && n_iter != nullptr && breakdown != nullptr,
          "null argument");
  REQUIRE(0 <= start_dim && start_dim <= end_dim && end_dim <= s->max_dim,
          "need 0 <= start_dim (%d) <= end_dim (%d) <= max_dim (%d)", start_dim, end_dim,
          s->max_dim);
  REQUIRE(ortho_kind == AB200_ORTHO_CGS2 || ortho_kind == AB200_ORTHO_MGS, "bad ortho_kind %d",
          ortho_kind);
  if (s->nnz < 0)
    return set_err(AB200_ESTATE, "ab200_expand called before ab200_set_csr / ab200_set_operator");
  if (s->disconnected) return set_err(AB200_ESTATE, "ab200_expand after ab200_comm_disconnect");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  s->pristine = false;
  const int md1 = s->max_dim + 1;
  init_ctl_kernel<<<1, 32, 0, s->stream>>>(s->ctl, nullptr, 0, false);
  CU(cudaGetLastError());
  s->st.kernel_launches += 1;
  if (s->nranks > 1) {
    // every rank's basis (restart update / uploaded columns) is final before any halo read
    CU(launch_peer_barrier(s->comm, s->ctl, s->real_mode ? 1 : 0, s->stream));
    s->st.kernel_launches += 1;
  }
  for (int j = start_dim; j < end_dim; ++j) {
    void* x = col_ptr(s, j);
    void* w = col_ptr(s, j + 1);
    int rc = enqueue_spmv(s, x, w, s->scale + j, j, true);  // decomposition.py:57-58
    if (rc != AB200_OK) return rc;
    OrthoArgs a = make_ortho_args(s, static_cast<cplx*>(w), j + 1, j, tol, eta,
                                  s->Hdev + (size_t)j * md1, 1);
    a.step_flag = s->step_round2 + j;
    rc = enqueue_ortho(s, a, ortho_kind);  // decomposition.py:60
    if (rc != AB200_OK) return rc;
  }
  if (end_dim > start_dim) {
    CU(cudaMemcpyAsync(s->h_H + (size_t)start_dim * md1, s->Hdev + (size_t)start_dim * md1,
                       sizeof(cplx) * (size_t)md1 * (end_dim - start_dim), cudaMemcpyDeviceToHost,
                       s->stream));
    CU(cudaMemcpyAsync(s->h_step_round2, s->step_round2, sizeof(int) * s->max_dim,
                       cudaMemcpyDeviceToHost, s->stream));
  }
  CU(cudaMemcpyAsync(s->h_scale, s->scale, sizeof(double) * md1, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(s->h_ctl, s->ctl, sizeof(StepCtl), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  resolve_pending(s, s->h_step_round2);
  if (s->h_ctl->comm_error == 2)
    return set_err(AB200_ECOMM, "ranks disagree on real / complex storage of the basis");
  if (s->h_ctl->comm_error) return set_err(AB200_ECOMM, "peer reduction timed out");
  int done = end_dim;
  *breakdown = 0;
  if (s->h_ctl->stop) {
    done = s->h_ctl->broke_at + 1;
    *breakdown = 1;
  }
  *n_iter = done;
  cplx* hc = reinterpret_cast<cplx*>(h_cols);
  for (int j = start_dim; j < done; ++j) {
    const bool broke = (*breakdown && j == done - 1);
    const int rows = broke ? j + 1 : j + 2;
    memcpy(hc + (size_t)j * md1, s->h_H + (size_t)j * md1, sizeof(cplx) * rows);
  }
  s->st.arnoldi_steps = s->h_ctl->steps_total;
  s->st.ortho_rounds = s->h_ctl->rounds_total;
  s->st.second_rounds = s->h_ctl->second_total;
  if (done > start_dim) {
    int fired = 0;
    for (int j = start_dim; j < done; ++j) fired += s->h_step_round2[j] ? 1 : 0;
    s->dgks_hot = 2 * fired > (done - start_dim);
  }
  return AB200_OK;
}

// shared by ab200_restart and ab200_combine: U[:, col0:col0+p] = U[:, col0:col0+m] q (+ tail)
static int apply_q(ab200_solver* s, const double* q, int64_t ldq, int col0, int m, int p,
                   bool copy_tail) {
  CU(cudaSetDevice(s->device));
  break_chain(s);
  CU(cudaStreamSynchronize(s->stream));  // the pinned Q staging buffer is free again
  const cplx* qh = reinterpret_cast<const cplx*>(q);
  if (s->real_mode) {
    bool q_real = true;
    for (int k = 0; k < p && q_real; ++k)
      for (int i = 0; i < m; ++i)
        if (qh[(size_t)k * ldq + i].y != 0.0) {
          q_real = false;
          break;
        }
    if (!q_real) {  // complex coefficients: the basis becomes complex from here on
      int rc = switch_to_complex(s);
      if (rc != AB200_OK) return rc;
    }
  }
  // fold the lazy column scales into the rows of Q:  V_i = scale[i] U_i
  for (int i = 0; i < m; ++i)
    for (int k = 0; k < p; ++k) {
      const cplx v = qh[(size_t)k * ldq + i];
      const double sc = s->h_scale[col0 + i];
      s->h_q[(size_t)i * p + k] = make_double2(v.x * sc, v.y * sc);
    }
  CU(cudaMemcpyAsync(s->qdev, s->h_q, sizeof(cplx) * (size_t)m * p, cudaMemcpyHostToDevice,
                     s->stream));
  RestartArgs a;
  a.copy_tail = copy_tail ? 1 : 0;
  a.real = s->real_mode ? 1 : 0;
  a.U = static_cast<cplx*>(col_ptr(s, col0));
  a.n = s->real_mode ? (s->n + 1) / 2 : s->n;
  a.ld = s->real_mode ? s->ld / 2 : s->ld;
  a.m = m;
  a.p = p;
  a.q = s->qdev;
  a.scale_m = copy_tail ? s->h_scale[col0 + m] : 1.0;
  {
    LaunchScope ls(s, K_RESTART, -1, 0, elem_bytes(s) * (double)s->n * (m + p + (copy_tail ? 2 : 0)));
    CU(launch_restart(a, s->num_sms, s->stream, s->opt_restart_variant));
  }
  const int nreset = p + (copy_tail ? 1 : 0);
  set_scale_kernel<<<1, 128, 0, s->stream>>>(s->scale, col0, nreset, 1.0);
  CU(cudaGetLastError());
  s->st.kernel_launches += 1;
  // no synchronisation here: the host goes on to update H and to enqueue the next expansion
  // while the update runs (the next call that reads results synchronises)
  for (int i = 0; i < nreset; ++i) s->h_scale[col0 + i] = 1.0;
  s->pristine = false;
  return AB200_OK;
}

int ab200_restart(ab200_solver* s, const double* q, int64_t ldq, int m, int p) {
  REQUIRE(s != nullptr && q != nullptr, "null argument");
  REQUIRE(m >= 1 && m <= s->max_dim && p >= 1 && p < m, "need 1 <= p (%d) < m (%d) <= max_dim (%d)", p,
          m, s->max_dim);
  REQUIRE(ldq >= m, "ldq < m");
  return apply_q(s, q, ldq, 0, m, p, true);
}

// V[:, col0:col0+p] = V[:, col0:col0+m] q   (p <= m), nothing else touched: Ritz / Schur
// vectors out of a block of basis columns (explicit_restarts.py:139,167; the last rotation of
// a solve that ran in a real basis).
int ab200_combine(ab200_solver* s, const double* q, int64_t ldq, int col0, int m, int p) {
  REQUIRE(s != nullptr && q != nullptr, "null argument");
  REQUIRE(col0 >= 0 && m >= 1 && p >= 1 && p <= m && col0 + m <= s->max_dim + 1,
          "need col0 >= 0, 1 <= p (%d) <= m (%d), col0 + m <= max_dim + 1 (%d)", p, m, s->max_dim + 1);
  REQUIRE(ldq >= m, "ldq < m");
  return apply_q(s, q, ldq, col0, m, p, false);
}

// Orthonormalise basis column `col` against columns [0, ncols) on the device (ncols <= col):
// CGS2/DGKS or MGS/DGKS as inside an expansion, then the column is normalised (lazily).
// *beta is its norm after the projections; *breakdown = 1 when beta < tol (the column lies in
// the span of the others and is left un-normalised).  Used to append a fresh direction after
// a happy breakdown and for the deflation step of the explicit-restart solver
// (explicit_restarts.py:63-77,111,141).
int ab200_orthonormalize_column(ab200_solver* s, int col, int ncols, double tol, double eta,
                                int ortho_kind, double* beta, int* breakdown) {
  REQUIRE(s != nullptr && beta != nullptr && breakdown != nullptr, "null argument");
  REQUIRE(col >= 0 && col <= s->max_dim && ncols >= 0 && ncols <= col && ncols <= s->max_dim,
          "need 0 <= ncols (%d) <= col (%d) <= max_dim (%d)", ncols, col, s->max_dim);
  REQUIRE(ortho_kind == AB200_ORTHO_CGS2 || ortho_kind == AB200_ORTHO_MGS, "bad ortho_kind %d",
          ortho_kind);
  if (s->disconnected)
    return set_err(AB200_ESTATE, "ab200_orthonormalize_column after ab200_comm_disconnect");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  s->pristine = false;
  init_ctl_kernel<<<1, 32, 0, s->stream>>>(s->ctl, nullptr, 0, false);
  CU(cudaGetLastError());
  CU(cudaMemsetAsync(s->hscratch, 0, sizeof(cplx) * (s->max_dim + 2), s->stream));
  s->st.kernel_launches += 1;
  cplx* w = static_cast<cplx*>(col_ptr(s, col));
  if (ncols == 0) {
    // nothing to project out: one norm pass (pass 2 with no columns) normalises the column
    // round 1 with nothing projected: nrm0sq is 0, so the DGKS test cannot fire
    OrthoArgs a = make_ortho_args(s, w, 0, col - 1, tol, eta, s->hscratch, 2);
    a.round = 1;
    CU(launch_cgs_pass2(a, s->num_sms, s->stream, s->opt_grid_mult));
    s->st.kernel_launches += 1;
  } else {
    // hscratch plays column `col - 1` of H: finalize writes beta at row col and scale[col]
    OrthoArgs a = make_ortho_args(s, w, ncols, col - 1, tol, eta, s->hscratch, 2);
    const int saved = s->opt_ortho_variant;
    s->opt_ortho_variant = 1;  // plain two-sweep rounds: the fused sweep assumes j == ncols - 1
    int rc = enqueue_ortho(s, a, ortho_kind);
    s->opt_ortho_variant = saved;
    if (rc != AB200_OK) return rc;
  }
  CU(cudaMemcpyAsync(s->h_scale, s->scale, sizeof(double) * (s->max_dim + 1), cudaMemcpyDeviceToHost,
                     s->stream));
  CU(cudaMemcpyAsync(s->h_ctl, s->ctl, sizeof(StepCtl), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  {
    int r2 = s->h_ctl->second_total > s->st.second_rounds ? 1 : 0;
    std::vector<int> flags(s->max_dim + 1, r2);
    resolve_pending(s, flags.data());
  }
  if (s->h_ctl->comm_error) return set_err(AB200_ECOMM, "peer reduction timed out");
  *beta = s->h_ctl->beta;
  *breakdown = s->h_ctl->stop ? 1 : 0;
  s->st.arnoldi_steps = s->h_ctl->steps_total;
  s->st.ortho_rounds = s->h_ctl->rounds_total;
  s->st.second_rounds = s->h_ctl->second_total;
  return AB200_OK;
}

// h[i] = <V_i, A V_col>, i < nrows: one SpMV into the scratch vector and one pass-1 sweep
// (explicit_restarts.py:150-151).
int ab200_project(ab200_solver* s, int col, int nrows, double* h_host) {
  REQUIRE(s != nullptr && h_host != nullptr, "null argument");
  REQUIRE(col >= 0 && col <= s->max_dim && nrows >= 1 && nrows <= s->max_dim + 1 && nrows <= kMaxDim,
          "need 0 <= col (%d) <= max_dim and 1 <= nrows (%d) <= max_dim + 1", col, nrows);
  if (s->nnz < 0) return set_err(AB200_ESTATE, "ab200_project called before an operator was set");
  if (s->disconnected) return set_err(AB200_ESTATE, "ab200_project after ab200_comm_disconnect");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  init_ctl_kernel<<<1, 32, 0, s->stream>>>(s->ctl, nullptr, 0, false);
  CU(cudaGetLastError());
  s->st.kernel_launches += 1;
  if (s->nranks > 1) {
    CU(launch_peer_barrier(s->comm, s->ctl, s->real_mode ? 1 : 0, s->stream));
    s->st.kernel_launches += 1;
  }
  CU(cudaMemsetAsync(s->hscratch, 0, sizeof(cplx) * (s->max_dim + 2), s->stream));
  int rc = enqueue_spmv(s, col_ptr(s, col), s->wtmp, s->scale + col, col, true);
  if (rc != AB200_OK) return rc;
  // the scale vector has max_dim + 1 entries: pass 1 reads scale[i] for i < nrows <= max_dim + 1
  OrthoArgs a = make_ortho_args(s, s->wtmp, nrows, nrows - 1, 0.0, 0.0, s->hscratch, 0);
  a.round = 1;
  a.accumulate = 0;
  {
    LaunchScope ls(s, K_PASS1, -1, 1, elem_bytes(s) * (double)s->n * (nrows + 1));
    CU(launch_cgs_pass1(a, s->num_sms, s->stream, s->opt_grid_mult));
  }
  CU(cudaMemcpyAsync(s->h_H, s->hscratch, sizeof(cplx) * nrows, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(s->h_ctl, s->ctl, sizeof(StepCtl), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  resolve_pending(s, nullptr);
  if (s->h_ctl->comm_error) return set_err(AB200_ECOMM, "peer reduction timed out");
  memcpy(h_host, s->h_H, sizeof(cplx) * nrows);
  return AB200_OK;
}

int ab200_spmv(ab200_solver* s, const double* x_host, double* y_host) {
  REQUIRE(s != nullptr && x_host != nullptr && y_host != nullptr, "null argument");
  if (s->nnz < 0) return set_err(AB200_ESTATE, "ab200_spmv called before ab200_set_csr");
  if (s->n != s->n_global)
    return set_err(AB200_ESTATE, "ab200_spmv is a single-GPU entry point (row block is partial)");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  if (!s->xtmp) CU(big_alloc(s, &s->xtmp, sizeof(cplx) * (size_t)s->ld));
  // a real x on a real operator takes the float64 kernels, exactly as inside a real-storage
  // expansion (the complex result has zero imaginary parts either way)
  bool real = s->real_mode && s->value_kind == AB200_F64 && s->op_fn == nullptr;
  for (int64_t r = 0; r < s->n && real; ++r)
    if (x_host[2 * r + 1] != 0.0) real = false;
  if (real) {
    // xtmp (ld complex slots = 2 ld doubles) holds x in its first half and y in its second
    double* xr = reinterpret_cast<double*>(s->xtmp);
    double* yr = xr + s->ld;
    CU(cudaMemcpyAsync(s->wtmp, x_host, sizeof(cplx) * (size_t)s->n, cudaMemcpyHostToDevice, s->stream));
    CU(launch_pack_real(s->wtmp, xr, s->n, s->num_sms, s->stream));
    int rc = enqueue_spmv(s, xr, yr, nullptr, -1, false, true);
    if (rc != AB200_OK) return rc;
    CU(launch_unpack_real(yr, s->wtmp, s->n, 1.0, s->num_sms, s->stream));
    s->st.kernel_launches += 2;
  } else {
    CU(cudaMemcpyAsync(s->xtmp, x_host, sizeof(cplx) * (size_t)s->n_global, cudaMemcpyHostToDevice,
                       s->stream));
    int rc = enqueue_spmv(s, s->xtmp, s->wtmp, nullptr, -1, false);
    if (rc != AB200_OK) return rc;
  }
  CU(cudaMemcpyAsync(y_host, s->wtmp, sizeof(cplx) * (size_t)s->n, cudaMemcpyDeviceToHost,
                     s->stream));
  CU(cudaStreamSynchronize(s->stream));
  resolve_pending(s, nullptr);
  return AB200_OK;
}

int ab200_ortho(ab200_solver* s, int ncols, double* w_host, double* h_host, double tol, double eta,
                int ortho_kind, double* beta, int* breakdown) {
  REQUIRE(s != nullptr && w_host != nullptr && h_host != nullptr && beta != nullptr &&
              breakdown != nullptr,
          "null argument");
  REQUIRE(ncols >= 1 && ncols <= s->max_dim, "ncols (%d) must be in [1, max_dim=%d]", ncols,
          s->max_dim);
  REQUIRE(ortho_kind == AB200_ORTHO_CGS2 || ortho_kind == AB200_ORTHO_MGS, "bad ortho_kind %d",
          ortho_kind);
  if (s->disconnected) return set_err(AB200_ESTATE, "ab200_ortho after ab200_comm_disconnect");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  {  // the stand-alone plug takes an arbitrary complex w: work on complex storage
    int rc = switch_to_complex(s);
    if (rc != AB200_OK) return rc;
  }
  CU(cudaMemcpyAsync(s->wtmp, w_host, sizeof(cplx) * (size_t)s->n, cudaMemcpyHostToDevice,
                     s->stream));
  CU(cudaMemsetAsync(s->hscratch, 0, sizeof(cplx) * (s->max_dim + 2), s->stream));
  init_ctl_kernel<<<1, 32, 0, s->stream>>>(s->ctl, nullptr, 0, false);
  CU(cudaGetLastError());
  s->st.kernel_launches += 1;
  OrthoArgs a = make_ortho_args(s, s->wtmp, ncols, ncols - 1, tol, eta, s->hscratch, 0);
  int rc = enqueue_ortho(s, a, ortho_kind);
  if (rc != AB200_OK) return rc;
  CU(cudaMemcpyAsync(w_host, s->wtmp, sizeof(cplx) * (size_t)s->n, cudaMemcpyDeviceToHost,
                     s->stream));
  CU(cudaMemcpyAsync(s->h_H, s->hscratch, sizeof(cplx) * ncols, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaMemcpyAsync(s->h_ctl, s->ctl, sizeof(StepCtl), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  // stand-alone call: every enqueued round-2 launch is attributed by the second_total delta
  int r2 = s->h_ctl->second_total > s->st.second_rounds ? 1 : 0;
  std::vector<int> flags(s->max_dim, r2);
  resolve_pending(s, flags.data());
  if (s->h_ctl->comm_error) return set_err(AB200_ECOMM, "peer reduction timed out");
  memcpy(h_host, s->h_H, sizeof(cplx) * ncols);
  *beta = s->h_ctl->beta;
  *breakdown = (*beta < tol) ? 1 : 0;
  s->st.ortho_rounds = s->h_ctl->rounds_total;
  s->st.second_rounds = s->h_ctl->second_total;
  return AB200_OK;
}

struct CommBlob {
  cudaIpcMemHandle_t v, slots, flags;
  int64_t ld;
  int64_t n;
  int max_dim;
  int magic;
};
static_assert(sizeof(CommBlob) <= AB200_COMM_BLOB_BYTES, "blob too large");
static const int kSlotDoubles = 2 * 129 + 8;

int ab200_comm_export(ab200_solver* s, void* blob) {
  REQUIRE(s != nullptr && blob != nullptr, "null argument");
  if (s->pooled)
    return set_err(AB200_ESTATE, "this solver owns the whole operator (nrows_local == n_global): nothing to share");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  if (!s->slots) {
    // two 8-byte packets per double (peer_allreduce)
    CU(cudaMalloc(&s->slots, 2 * sizeof(double) * 2 * kMaxRanks * kSlotDoubles));
    CU(cudaMalloc(&s->flags, sizeof(unsigned long long) * 2 * kMaxRanks));
    CU(cudaMalloc(&s->seq, sizeof(unsigned long long)));
    CU(cudaMemset(s->slots, 0, 2 * sizeof(double) * 2 * kMaxRanks * kSlotDoubles));
    CU(cudaMemset(s->flags, 0, sizeof(unsigned long long) * 2 * kMaxRanks));
    CU(cudaMemset(s->seq, 0, sizeof(unsigned long long)));
    CU(cudaDeviceSynchronize());
  }
  CommBlob b;
  memset(&b, 0, sizeof(b));
  CU(cudaIpcGetMemHandle(&b.v, s->V));
  CU(cudaIpcGetMemHandle(&b.slots, s->slots));
  CU(cudaIpcGetMemHandle(&b.flags, s->flags));
  b.ld = s->ld;
  b.n = s->n;
  b.max_dim = s->max_dim;
  b.magic = 0xab200;
  memset(blob, 0, AB200_COMM_BLOB_BYTES);
  memcpy(blob, &b, sizeof(b));
  return AB200_OK;
}

int ab200_comm_connect(ab200_solver* s, int rank, int nranks, const void* blobs,
                       const int64_t* row_starts) {
  REQUIRE(s != nullptr && blobs != nullptr && row_starts != nullptr, "null argument");
  REQUIRE(nranks >= 1 && nranks <= kMaxRanks && rank >= 0 && rank < nranks,
          "need 0 <= rank (%d) < nranks (%d) <= %d", rank, nranks, kMaxRanks);
  if (!s->slots) return set_err(AB200_ESTATE, "ab200_comm_connect before ab200_comm_export");
  REQUIRE(row_starts[0] == 0 && row_starts[nranks] == s->n_global, "row_starts must span [0, n_global]");
  REQUIRE(row_starts[rank] == s->row0 && row_starts[rank + 1] == s->row0 + s->n,
          "row_starts[rank] does not match this solver's row block");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  const unsigned char* bl = static_cast<const unsigned char*>(blobs);
  for (int r = 0; r < nranks; ++r) {
    CommBlob b;
    memcpy(&b, bl + (size_t)r * AB200_COMM_BLOB_BYTES, sizeof(b));
    REQUIRE(b.magic == 0xab200 && b.max_dim == s->max_dim, "blob of rank %d is not from a matching solver", r);
    REQUIRE(b.n == row_starts[r + 1] - row_starts[r], "blob of rank %d has %lld rows, partition says %lld", r,
            (long long)b.n, (long long)(row_starts[r + 1] - row_starts[r]));
    s->peer_ld[r] = b.ld;
    s->row_starts[r] = row_starts[r];
    if (r == rank) {
      s->comm.slots[r] = s->slots;
      s->comm.flags[r] = s->flags;
      continue;
    }
    CU(cudaIpcOpenMemHandle(&s->peer_V[r], b.v, cudaIpcMemLazyEnablePeerAccess));
    CU(cudaIpcOpenMemHandle(&s->peer_slots[r], b.slots, cudaIpcMemLazyEnablePeerAccess));
    CU(cudaIpcOpenMemHandle(&s->peer_flags[r], b.flags, cudaIpcMemLazyEnablePeerAccess));
    s->comm.slots[r] = static_cast<double*>(s->peer_slots[r]);
    s->comm.flags[r] = static_cast<unsigned long long*>(s->peer_flags[r]);
  }
  s->row_starts[nranks] = row_starts[nranks];
  s->rank = rank;
  s->nranks = nranks;
  s->comm.nranks = nranks;
  s->comm.rank = rank;
  s->comm.slot_doubles = kSlotDoubles;
  s->comm.seq = s->seq;
  return AB200_OK;
}

// Latency of the in-kernel peer reduction: `iters` back-to-back launches of the 1-block
// exchange kernel (remote stores of the partial into every peer, system fence, flag, wait for
// the peers' flags, rank-order sum), timed with CUDA events on the solver's stream.  Collective:
// every rank must call it with the same `iters`.
int ab200_comm_bench(ab200_solver* s, int iters, double* us_per_exchange) {
  REQUIRE(s != nullptr && us_per_exchange != nullptr && iters >= 1, "bad argument");
  if (s->nranks <= 1) return set_err(AB200_ESTATE, "ab200_comm_bench needs ab200_comm_connect first");
  if (s->disconnected) return set_err(AB200_ESTATE, "ab200_comm_bench after ab200_comm_disconnect");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  init_ctl_kernel<<<1, 32, 0, s->stream>>>(s->ctl, nullptr, 0, false);
  for (int i = 0; i < 8; ++i) CU(launch_peer_barrier(s->comm, s->ctl, s->real_mode ? 1 : 0, s->stream));
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a));
  CU(cudaEventCreate(&b));
  CU(cudaEventRecord(a, s->stream));
  for (int i = 0; i < iters; ++i) CU(launch_peer_barrier(s->comm, s->ctl, s->real_mode ? 1 : 0, s->stream));
  CU(cudaEventRecord(b, s->stream));
  CU(cudaEventSynchronize(b));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a), cudaEventDestroy(b);
  CU(cudaMemcpy(s->h_ctl, s->ctl, sizeof(StepCtl), cudaMemcpyDeviceToHost));
  if (s->h_ctl->comm_error) return set_err(AB200_ECOMM, "peer reduction timed out");
  s->st.kernel_launches += iters + 9;
  *us_per_exchange = 1e3 * (double)ms / iters;
  return AB200_OK;
}

int ab200_set_halo(ab200_solver* s, const int64_t* ghost_cols, int64_t nghost) {
  REQUIRE(s != nullptr && nghost >= 0 && (nghost == 0 || ghost_cols != nullptr), "bad argument");
  if (nghost > 0 && s->nranks <= 1)
    return set_err(AB200_ESTATE, "ab200_set_halo with ghost columns needs ab200_comm_connect first");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  CU(cudaStreamSynchronize(s->stream));
  cudaFree(s->ghost), cudaFree(s->ghost_off);
  s->ghost = nullptr, s->ghost_off = nullptr, s->nghost = 0;
  if (nghost == 0) return AB200_OK;
  std::vector<int64_t> off((size_t)nghost);
  int q = 0;
  for (int r = 0; r <= kMaxRanks; ++r) s->seg_start[r] = nghost;
  s->seg_start[0] = 0;
  int64_t prev = -1;
  for (int64_t k = 0; k < nghost; ++k) {
    const int64_t g = ghost_cols[k];
    REQUIRE(g > prev && g >= 0 && g < s->n_global, "ghost_cols must be strictly increasing in [0, n)");
    REQUIRE(g < s->row0 || g >= s->row0 + s->n, "ghost column %lld is inside the local block", (long long)g);
    prev = g;
    while (g >= s->row_starts[q + 1]) {
      ++q;
      s->seg_start[q] = k;
    }
    off[(size_t)k] = g - s->row_starts[q];
  }
  for (int r = q + 1; r <= kMaxRanks; ++r) s->seg_start[r] = nghost;
  CU(cudaMalloc(&s->ghost, sizeof(cplx) * (size_t)nghost));
  CU(cudaMalloc(&s->ghost_off, sizeof(int64_t) * (size_t)nghost));
  CU(cudaMemcpy(s->ghost_off, off.data(), sizeof(int64_t) * (size_t)nghost, cudaMemcpyHostToDevice));
  s->nghost = nghost;
  return AB200_OK;
}

struct HaloBlob {
  cudaIpcMemHandle_t ghost, hflags;
  int64_t nghost;
  int magic;
};
static_assert(sizeof(HaloBlob) <= AB200_COMM_BLOB_BYTES, "blob too large");

int ab200_halo_export(ab200_solver* s, void* blob) {
  REQUIRE(s != nullptr && blob != nullptr, "null argument");
  if (s->nranks <= 1) return set_err(AB200_ESTATE, "ab200_halo_export needs ab200_comm_connect first");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  if (!s->hflags) {
    CU(cudaMalloc(&s->hflags, sizeof(unsigned long long) * kMaxRanks));
    CU(cudaMemset(s->hflags, 0, sizeof(unsigned long long) * kMaxRanks));
    CU(cudaMalloc(&s->push_ticket, sizeof(unsigned)));
    CU(cudaMemset(s->push_ticket, 0, sizeof(unsigned)));
  }
  if (!s->ghost) {  // a rank that reads nothing remote still owns a (1-entry) buffer to export
    CU(cudaMalloc(&s->ghost, sizeof(cplx)));
  }
  CU(cudaDeviceSynchronize());
  HaloBlob b;
  memset(&b, 0, sizeof(b));
  CU(cudaIpcGetMemHandle(&b.ghost, s->ghost));
  CU(cudaIpcGetMemHandle(&b.hflags, s->hflags));
  b.nghost = s->nghost;
  b.magic = 0xab201;
  memset(blob, 0, AB200_COMM_BLOB_BYTES);
  memcpy(blob, &b, sizeof(b));
  return AB200_OK;
}

int ab200_halo_connect(ab200_solver* s, const void* blobs, const int64_t* send_idx,
                       const int64_t* send_ptr, const int64_t* dst_off) {
  REQUIRE(s != nullptr && blobs != nullptr && send_ptr != nullptr && dst_off != nullptr,
          "null argument");
  if (!s->hflags) return set_err(AB200_ESTATE, "ab200_halo_connect before ab200_halo_export");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  const unsigned char* bl = static_cast<const unsigned char*>(blobs);
  REQUIRE(send_ptr[0] == 0, "send_ptr[0] must be 0");
  for (int r = 0; r < s->nranks; ++r) {
    REQUIRE(send_ptr[r + 1] >= send_ptr[r], "send_ptr must be non-decreasing");
    s->send_ptr[r] = send_ptr[r];
    s->dst_off[r] = dst_off[r];
    if (r == s->rank) {
      REQUIRE(send_ptr[r + 1] == send_ptr[r], "a rank does not send to itself");
      continue;
    }
    HaloBlob b;
    memcpy(&b, bl + (size_t)r * AB200_COMM_BLOB_BYTES, sizeof(b));
    REQUIRE(b.magic == 0xab201, "halo blob of rank %d is invalid", r);
    REQUIRE(dst_off[r] >= 0 && dst_off[r] + (send_ptr[r + 1] - send_ptr[r]) <= (b.nghost > 0 ? b.nghost : 1),
            "entries for rank %d overflow its ghost buffer", r);
    CU(cudaIpcOpenMemHandle(&s->peer_ghost[r], b.ghost, cudaIpcMemLazyEnablePeerAccess));
    CU(cudaIpcOpenMemHandle(&s->peer_hflags[r], b.hflags, cudaIpcMemLazyEnablePeerAccess));
  }
  for (int r = s->nranks; r <= kMaxRanks; ++r) s->send_ptr[r] = send_ptr[s->nranks];
  const int64_t nsend = send_ptr[s->nranks];
  for (int64_t k = 0; k < nsend; ++k)
    REQUIRE(send_idx[k] >= 0 && send_idx[k] < s->n, "send_idx[%lld] out of the local block", (long long)k);
  cudaFree(s->send_idx);
  s->send_idx = nullptr;
  CU(cudaMalloc(&s->send_idx, sizeof(int64_t) * (size_t)(nsend > 0 ? nsend : 1)));
  if (nsend > 0)
    CU(cudaMemcpy(s->send_idx, send_idx, sizeof(int64_t) * (size_t)nsend, cudaMemcpyHostToDevice));
  s->push = true;
  s->hseq = 0;
  return AB200_OK;
}

int ab200_set_timing(ab200_solver* s, int enabled) {
  REQUIRE(s != nullptr, "solver is null");
  s->timing = enabled != 0;
  return AB200_OK;
}

int ab200_reset_stats(ab200_solver* s) {
  REQUIRE(s != nullptr, "solver is null");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  CU(cudaStreamSynchronize(s->stream));
  resolve_pending(s, nullptr, true);
  memset(&s->st, 0, sizeof(s->st));
  init_ctl_kernel<<<1, 32, 0, s->stream>>>(s->ctl, nullptr, 0, true);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  return AB200_OK;
}

int ab200_get_stats(ab200_solver* s, ab200_stats* out) {
  REQUIRE(s != nullptr && out != nullptr, "null argument");
  if (!s->pending.empty()) {  // launches whose event times have not been read yet
    CU(cudaSetDevice(s->device));
    break_chain(s);
    CU(cudaStreamSynchronize(s->stream));
    resolve_pending(s, s->h_step_round2, true);
  }
  *out = s->st;
  out->real_storage = s->real_mode ? 1 : 0;
  return AB200_OK;
}

int ab200_synchronize(ab200_solver* s) {
  REQUIRE(s != nullptr, "solver is null");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  CU(cudaStreamSynchronize(s->stream));
  return AB200_OK;
}

int ab200_timer_start(ab200_solver* s) {
  REQUIRE(s != nullptr, "solver is null");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  if (!s->t0) {
    CU(cudaEventCreate(&s->t0));
    CU(cudaEventCreate(&s->t1));
  }
  CU(cudaEventRecord(s->t0, s->stream));
  return AB200_OK;
}
int ab200_timer_stop(ab200_solver* s, double* elapsed_ms) {
  REQUIRE(s != nullptr && elapsed_ms != nullptr, "null argument");
  if (!s->t0) return set_err(AB200_ESTATE, "ab200_timer_stop without ab200_timer_start");
  CU(cudaSetDevice(s->device));
  break_chain(s);
  CU(cudaEventRecord(s->t1, s->stream));
  CU(cudaEventSynchronize(s->t1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, s->t0, s->t1));
  *elapsed_ms = ms;
  return AB200_OK;
}

int ab200_set_option(ab200_solver* s, const char* key, int64_t value) {
  REQUIRE(s != nullptr && key != nullptr, "null argument");
  if (!strcmp(key, "grid_mult"))
    s->opt_grid_mult = (int)value;
  else if (!strcmp(key, "restart_variant"))
    s->opt_restart_variant = (int)value;
  else if (!strcmp(key, "ortho_variant"))
    s->opt_ortho_variant = (int)value;
  else if (!strcmp(key, "spmv_variant"))
    s->opt_spmv_variant = (int)value;
  else if (!strcmp(key, "spmv_stages"))
    s->opt_spmv_stages = (int)value;
  else if (!strcmp(key, "spmv_bps"))
    s->opt_spmv_bps = (int)value;
  else if (!strcmp(key, "halo_fold"))
    s->opt_halo_fold = (int)value;
  else if (!strcmp(key, "spmv_window"))
    s->opt_spmv_window = (int)value;
  else if (!strcmp(key, "spmv_win_half"))
    s->opt_spmv_win_half = (int)value;
  else if (!strcmp(key, "spmv_ring_warps"))
    s->opt_spmv_ring_warps = (int)value;
  else if (!strcmp(key, "fused_r"))
    s->opt_fused_r = (int)value;
  else if (!strcmp(key, "fused_stages"))
    s->opt_fused_stages = (int)value;
  else if (!strcmp(key, "fused_ct"))
    s->opt_fused_ct = (int)value;
  else if (!strcmp(key, "real_mode")) {
    if (value == 0) {
      CU(cudaSetDevice(s->device));
  break_chain(s);
      int rc = switch_to_complex(s);
      if (rc != AB200_OK) return rc;
    } else if (!s->pristine && !s->real_mode) {
      return set_err(AB200_ESTATE, "real storage can only be chosen before any column is written");
    } else if (s->pristine) {
      s->real_mode = true;
    }
  } else if (!strcmp(key, "spmv_threads"))
    s->opt_spmv_threads = (int)value;
  else if (!strcmp(key, "spmv_tile"))
    s->opt_spmv_tile = (int)value;
  else
    return set_err(AB200_EINVAL, "unknown option '%s'", key);
  return AB200_OK;
}

int ab200_host_alloc(void** out, int64_t bytes) {
  REQUIRE(out != nullptr && bytes > 0, "bad argument");
  CU(cudaMallocHost(out, (size_t)bytes));
  return AB200_OK;
}
int ab200_host_free(void* p) {
  if (p) CU(cudaFreeHost(p));
  return AB200_OK;
}

}  // extern "C"
