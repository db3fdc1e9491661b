/*
 * arnoldi_b200.h -- C ABI of the B200-native Krylov-Schur hot path.
 *
 * Drop-in boundary for cournape/arnoldi-py's partial_schur().  The reference is
 * pure Python and has no FFI of its own; each entry point
 * below replaces one n-length call site of the reference, cited as file:line
 * under /root/reference/src/arnoldi.  Everything n-length runs as sm_100a CUDA
 * and stays in HBM; the (max_dim+1) x max_dim projected matrix H, the Schur
 * reorder, convergence tests and sort_function stay on the host
 * (krylov_schur.py:69-76,83-101).
 *
 * Conventions
 *   - plain C types only; complex128 is passed as interleaved (re, im) doubles.
 *   - every function returns 0 on success, a negative AB200_E* code otherwise;
 *     ab200_last_error() gives the message for the calling thread.
 *   - one handle = one GPU = one block of rows [row0, row0 + nrows_local) of A
 *     and of the Krylov basis V.  A single-GPU solve has row0 = 0,
 *     nrows_local = n_global.
 *   - a handle is used from one host thread at a time (krylov_schur.py is
 *     single-threaded); different handles may be used concurrently.
 *   - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef ARNOLDI_B200_H
#define ARNOLDI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AB200_ABI_VERSION 2

#define AB200_OK 0
#define AB200_EINVAL (-1)  /* bad argument (maps to AssertionError in Python) */
#define AB200_ECUDA (-2)   /* CUDA runtime / kernel failure                   */
#define AB200_ENOMEM (-3)  /* device or pinned allocation failed              */
#define AB200_ESTATE (-4)  /* call order violated (e.g. expand before set_csr) */
#define AB200_ECOMM (-5)   /* multi-GPU communicator failure                  */

/* value_kind of the CSR values array */
#define AB200_F64 0   /* float64 values (mark, Laplacians: matrices.py:73) */
#define AB200_C128 1  /* complex128 values (scripts/benchmark-partial-schur.py:78) */

/* ortho_kind: the orthonormalisation plug (ortho.py) */
#define AB200_ORTHO_CGS2 0 /* dgks_gs,  ortho.py:56-107 (the reference default, decomposition.py:60) */
#define AB200_ORTHO_MGS 1  /* dgks_mgs, ortho.py:9-53 */

/* spmv_algo for ab200_set_csr */
#define AB200_SPMV_AUTO 0   /* STREAM when no row has more than 16 entries; else MERGE when at   \
                               least half of the (sampled) entries lie within the ring's reach   \
                               of the diagonal, else VECTOR                                      */
#define AB200_SPMV_VECTOR 1 /* nnz tiles snapped to rows, one block per tile, x gathered from    \
                               global memory; a warp sums every segment > 16 entries            */
#define AB200_SPMV_STREAM 2 /* one thread per row, tiles moved by bulk copies (TMA) through a   \
                               shared-memory ring; rows of <= 16 entries only (else EINVAL);    \
                               bit-identical to scipy's csr_matvec                              */
#define AB200_SPMV_MERGE 3  /* skewed rows: equal-nnz tiles snapped to rows (the merge-path      \
                               split), one WARP per tile, lanes along the entries for the        \
                               products and along the rows for the sums; x gathered from a      \
                               shared-memory ring that slides along the diagonal (bulk copies).  \
                               float64 vectors (real storage); complex128 vectors use VECTOR    */

typedef struct ab200_solver ab200_solver; /* opaque */

/* Per-kernel-class accounting since the last ab200_reset_stats(): device time
 * (CUDA events on the solver's stream, only when timing is enabled), launches,
 * and ALGORITHMIC bytes (SURVEY.md section 8d formulas, evaluated with the
 * column counts and DGKS rounds actually executed). */
typedef struct ab200_stats {
  /* ortho_fused = the sweep that does round-1 pass 2 and round-2 pass 1 together */
  double spmv_ms, ortho_pass1_ms, ortho_pass2_ms, ortho_fused_ms, mgs_ms, restart_ms;
  double spmv_bytes, ortho_pass1_bytes, ortho_pass2_bytes, ortho_fused_bytes, mgs_bytes,
      restart_bytes;
  int64_t spmv_launches, ortho_pass1_launches, ortho_pass2_launches, ortho_fused_launches,
      mgs_launches, restart_launches;
  int64_t arnoldi_steps;   /* true operator applications (matvecs)          */
  int64_t ortho_rounds;    /* CGS/MGS rounds executed (1 or 2 per step)     */
  int64_t second_rounds;   /* steps where the DGKS test fired               */
  int64_t kernel_launches; /* every kernel this library launched            */
  int64_t real_storage;    /* 1 while the basis is held as float64 (provably real) */
} ab200_stats;

int ab200_abi_version(void);
const char *ab200_last_error(void);
/* number of visible CUDA devices, or a negative error code */
int ab200_device_count(void);

/* Allocate the solver state on `device`: V (nrows_local x (max_dim+1),
 * complex128, column-major, zero-filled -- krylov_schur.py:42), two n-length
 * work vectors, the device copy of H and reduction scratch.
 * Replaces the allocations at krylov_schur.py:42-43. */
int ab200_create(ab200_solver **out, int device, int64_t n_global, int64_t row0,
                 int64_t nrows_local, int max_dim);
int ab200_destroy(ab200_solver *s);

/* Upload this GPU's block of rows of A in CSR form (what scipy holds for
 * `A @ x`, decomposition.py:58).  indptr has nrows_local+1 entries of
 * indptr_bits (32 or 64) and is relative to the block (indptr[0] == 0);
 * indices are GLOBAL column ids (int32, as scipy stores them below 2^31 nnz);
 * values are float64 or complex128 according to value_kind. */
int ab200_set_csr(ab200_solver *s, const void *indptr, int indptr_bits, const int32_t *indices,
                  const void *values, int value_kind, int64_t nnz, int spmv_algo);

/* A device operator instead of a CSR block (the duck-typed `A @ x` of decomposition.py:58 for
 * operators without stored entries; README.md:119 "LinearOperator support").  Inside
 * ab200_expand, step j calls fn(user, x, y, n, is_real, stream): x and y are DEVICE pointers to
 * n entries (float64 when is_real, else interleaved complex128), x holds v_j; fn must ENQUEUE
 * y = A x on the CUDA stream `stream` (a cudaStream_t) without synchronising, and return 0.
 * value_kind says whether A is real (AB200_F64: the basis may stay in float64 storage) or
 * complex.  Single GPU. */
typedef int (*ab200_apply_fn)(void *user, const void *x, void *y, int64_t n, int is_real,
                              void *stream);
int ab200_set_operator(ab200_solver *s, ab200_apply_fn fn, void *user, int value_kind);

/* Copy host complex128 columns into / out of V (column-major, leading dimension
 * ld_host in elements).  set_columns replaces `V[:, 0] = v0` (krylov_schur.py:46);
 * get_columns replaces the final `V[:, :nev]` view (krylov_schur.py:110). */
int ab200_set_columns(ab200_solver *s, int col0, int ncols, const double *host, int64_t ld_host);
int ab200_get_columns(ab200_solver *s, int col0, int ncols, double *host, int64_t ld_host);

/* Arnoldi expansion from column start_dim to end_dim on the device
 * (arnoldi_decomposition, decomposition.py:13-68): for each j the SpMV
 * (decomposition.py:57-58), the orthogonalisation against V[:, :j+1]
 * (decomposition.py:60 -> ortho.py), the breakdown test (decomposition.py:61-63)
 * and the normalisation (decomposition.py:65-66).  No host round trip inside.
 *
 * h_cols (host, complex128, column-major with leading dimension max_dim+1)
 * receives, for every executed column j in [start_dim, *n_iter), rows 0..j
 * (the projections) and, unless step j broke down, row j+1 (beta).
 * *n_iter is end_dim, or j+1 if step j broke down (beta < tol), in which
 * case *breakdown = 1 and V[:, j+1] is left un-normalised as in the reference. */
int ab200_expand(ab200_solver *s, int start_dim, int end_dim, double tol, double eta,
                 int ortho_kind, double *h_cols, int *n_iter, int *breakdown);

/* Krylov-Schur truncation (krylov_schur.py:78 and :81 in one pass):
 * V[:, :p] = V[:, :m] Q[:, :p];  V[:, p] = V[:, m].
 * q is host complex128, column-major m x p with leading dimension ldq. */
int ab200_restart(ab200_solver *s, const double *q, int64_t ldq, int m, int p);

/* V[:, col0:col0+p] = V[:, col0:col0+m] q  (p <= m; q host complex128 column-major m x p):
 * Ritz / Schur vectors out of a block of basis columns, nothing else touched
 * (explicit_restarts.py:139 `ritz.vectors[:, 0]`, :167 `V[:, :nev] @ Y`). */
int ab200_combine(ab200_solver *s, const double *q, int64_t ldq, int col0, int m, int p);

/* Orthonormalise basis column `col` against columns [0, ncols), ncols <= col, with the same
 * kernels as an Arnoldi step; *beta = its norm after the projections, *breakdown = 1 when
 * beta < tol (column left un-normalised).  The deflation `mgs(V[:, :k], v)` of
 * explicit_restarts.py:63-77,111,141, and the fresh direction appended after a happy
 * breakdown (krylov_schur.py:57-59 raises there). */
int ab200_orthonormalize_column(ab200_solver *s, int col, int ncols, double tol, double eta,
                                int ortho_kind, double *beta, int *breakdown);

/* h[i] = <V[:, i], A V[:, col]> for i in [0, nrows)  (host complex128 out): the projection
 * explicit_restarts.py:150-151 computes column by column with np.vdot. */
int ab200_project(ab200_solver *s, int col, int nrows, double *h_host);

/* ---- single-operation entry points (the plugs the reference exposes) ---- */

/* y = A x with host vectors (the duck-typed `A @ x`, decomposition.py:58).
 * x has n_global entries (single GPU) -- complex128. */
int ab200_spmv(ab200_solver *s, const double *x_host, double *y_host);

/* dgks_gs / dgks_mgs (ortho.py:56,9): orthogonalise host vector w (nrows_local
 * complex128, updated in place) against V[:, :ncols] currently on the device;
 * h (ncols complex128) receives the projections. */
int ab200_ortho(ab200_solver *s, int ncols, double *w_host, double *h_host, double tol,
                double eta, int ortho_kind, double *beta, int *breakdown);

/* ---- one box, several GPUs: block-row shards, one process per GPU ----
 * The reference has no distributed path; these calls carry the
 * north star's sharding.  Bootstrap (exchanging the opaque blobs below between the
 * processes) is the caller's job -- the Python driver uses torch.distributed.
 *
 * ab200_comm_export  writes this rank's AB200_COMM_BLOB_BYTES-byte blob (CUDA IPC handles of
 *                    its basis, reduction slots and flags, plus its leading dimension).
 * ab200_comm_connect takes every rank's blob (rank-major) and the row partition
 *                    row_starts[0..nranks]; maps the peers' memory.  After it, every
 *                    reduction inside ab200_expand / ab200_ortho is a global one (partials
 *                    pushed to all peers over NVLink inside the reducing kernel, summed in rank
 *                    order: identical bits on every rank) and all ranks must make the same
 *                    sequence of calls.
 * ab200_set_halo     declares which remote entries of v the local CSR block reads: ghost_cols
 *                    are GLOBAL column ids, strictly increasing, none inside the local block.
 *                    The CSR passed to ab200_set_csr must then use LOCAL column numbering:
 *                    id < nrows_local = local row id, id >= nrows_local = nrows_local + index
 *                    into ghost_cols.  */
#define AB200_COMM_BLOB_BYTES 256
#define AB200_MAX_RANKS 8
int ab200_comm_export(ab200_solver *s, void *blob);
int ab200_comm_connect(ab200_solver *s, int rank, int nranks, const void *blobs,
                       const int64_t *row_starts);
int ab200_set_halo(ab200_solver *s, const int64_t *ghost_cols, int64_t nghost);
/* First half of the multi-GPU teardown: synchronise and unmap every peer buffer.  Every rank
 * calls it, the ranks meet (host barrier), then each calls ab200_destroy -- so no rank frees
 * memory a peer still maps or reads.  The handle accepts no further multi-GPU work. */
int ab200_comm_disconnect(ab200_solver *s);
/* Measure the fused peer reduction alone: `iters` back-to-back 1-block exchanges (launch + remote
 * stores + system fence + flags + rank-order sum), microseconds per exchange.  Collective. */
int ab200_comm_bench(ab200_solver *s, int iters, double *us_per_exchange);
/* Optional owner-side push of the halo (for scattered halos).  After ab200_set_halo on every
 * rank: ab200_halo_export writes a blob with the IPC handles of this rank's ghost buffer and
 * delivery flags; ab200_halo_connect takes all ranks' blobs plus, for every peer r, the LOCAL
 * rows of this rank that r reads (send_idx[send_ptr[r] .. send_ptr[r+1]), ascending = the order
 * of r's ghost_cols) and dst_off[r] = index in r's ghost buffer where this rank's entries start.
 * From then on ab200_expand pushes instead of pulling. */
int ab200_halo_export(ab200_solver *s, void *blob);
int ab200_halo_connect(ab200_solver *s, const void *blobs, const int64_t *send_idx,
                       const int64_t *send_ptr, const int64_t *dst_off);

/* ---- measurement ---- */
int ab200_set_timing(ab200_solver *s, int enabled); /* CUDA-event timing per kernel class */
int ab200_reset_stats(ab200_solver *s);
int ab200_get_stats(ab200_solver *s, ab200_stats *out);
int ab200_synchronize(ab200_solver *s);
/* Device-side stopwatch on the solver's stream (CUDA events): start, then stop returns
 * the elapsed milliseconds of everything enqueued in between (after it has finished). */
int ab200_timer_start(ab200_solver *s);
int ab200_timer_stop(ab200_solver *s, double *elapsed_ms);
/* Select kernel variants (for A/B measurements); value 0 = automatic.
 *   "ortho_variant"   CGS2 schedule: 0 = adaptive (fused sweep while the DGKS test fires on
 *                     most steps, two-sweep rounds otherwise), 1 = always two-sweep rounds,
 *                     2 = always fused (register loads), 3 = always fused (cp.async staging),
 *                     4 = always fused, warp-private tiles (c <= 64)
 *   "fused_ct"        column-tile width of the fused sweep (1..8)
 *   "fused_stages"    cp.async ring depth of the fused sweep (2..4)
 *   "fused_r"         1 = one 16-byte element per lane and chunk in the fused sweep
 *   "restart_variant" outputs per warp of the restart kernel (4, 8, 16)
 *   "real_mode"       0 = keep the basis as complex128 from the start (default: float64
 *                     storage while A, v0 and every Q applied are real -- see DESIGN.md)
 *   "grid_mult"       resident blocks per SM for the orthogonalisation kernels
 *   "spmv_variant"    short rows: 0 = bulk-copy (TMA) pipeline, 1 = plain tile kernel,
 *                     2 = cp.async streaming kernel (round 1)
 *   "spmv_tile"       non-zeros staged per SpMV block    }
 *   "spmv_threads"    SpMV block size, 128 or 256        } take effect at the next
 *   "spmv_stages"     ring depth of the bulk pipeline    } ab200_set_csr
 *   "spmv_bps"        resident SpMV blocks per SM (0 = what fits)
 *   "spmv_window"     MERGE: ring capacity in x entries (multiple of 256, <= 16384)  } at the next
 *   "spmv_win_half"   MERGE: entries kept on each side of a round's rows             } ab200_set_csr
 *   "spmv_ring_warps" MERGE: consumer warps per block (<= 15)
 *   "halo_fold"       0 = separate halo gather kernel before each SpMV (default 1: banded
 *                     operators read their halo inside the SpMV) */
int ab200_set_option(ab200_solver *s, const char *key, int64_t value);

/* ---- host side of a restart: the m x m "rotate" of krylov_schur.py:69-72 ----
 * zgees on H_m, then the reference's ordered_schur: one ?trexc move per target slot in the
 * order `perm` asks for (utils.py:52-63), then Q = Q1 Q2 -- the reference's LAPACK calls in
 * the reference's order, issued from C so the GPUs wait as little as possible.  The library
 * links no LAPACK: the caller passes the routine addresses (the Python driver takes them from
 * scipy.linalg.cython_lapack, i.e. the very OpenBLAS the reference runs on).
 *   phase 1  ab200_host_schur:  T1 (in place in `t`, column-major m x m), Q1 into `q`.
 *   phase 2  ab200_host_reorder: given perm[m] = sort_function(diag(T1)), reorders T in place
 *            and writes Q2 (the accumulated swaps, starting from the identity) into q; the
 *            caller forms Q = Q1 Q2.
 * work: caller-provided scratch of at least 4 m m + 8 m doubles. */
int ab200_host_schur(void *zgees_fn, int m, double *t, double *q, double *work);
/* The same factorisation for an H_m whose imaginary parts are all zero (real operator, real
 * basis), through dgees + one complex Givens rotation per 2 x 2 block (scipy.linalg.rsf2csf's
 * construction): a third of the arithmetic.  A valid complex Schur form, NOT the one zgees
 * returns (order of the diagonal before sorting, phases, rounding): used where the driver
 * leaves the reference's arithmetic anyway (see `fast_real_schur` in INTEGRATION.md). */
int ab200_host_schur_real(void *dgees_fn, int m, double *t, double *q);
/* Real SYMMETRIC H_m (the projected matrix of a symmetric operator): eigendecomposition by dsyevd
 * on (H + H^T) / 2 -- T diagonal (ascending), Q real orthogonal, both written as complex m x m;
 * the ordered form is then a column permutation (no ztrexc).  Returns 1 without writing anything
 * when max |H - H^T| > sym_tol * max |H| (take ab200_host_schur_real then), 0 on success.  Like
 * ab200_host_schur_real: opt-in (fast_real_schur), a valid Schur form, not the reference's bits. */
int ab200_host_eigh_real(void *dsyevd_fn, int m, double *t, double *q, double sym_tol);
int ab200_host_reorder(void *ztrexc_fn, int m, double *t, double *q, const int64_t *perm,
                       double *work);

/* Pinned host memory for callers that want asynchronous, full-speed uploads. */
int ab200_host_alloc(void **out, int64_t bytes);
int ab200_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* ARNOLDI_B200_H */
