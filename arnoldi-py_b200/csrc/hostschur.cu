// hostschur.cu -- host side of a Krylov-Schur restart: the m x m "rotate" step
// (krylov_schur.py:69-72 of the reference):
//
//     T1, Q1 = schur(H_m, output="complex")                      scipy -> LAPACK zgees('V','N')
//     T2, Q2 = ordered_schur(T1, "complex", sort_function)       utils.py:32-67: one ztrexc
//     Q = Q1 @ Q2                                                 move per target slot
//
// No device code here.  The GPUs are idle while this runs (it needs the whole of H_m and
// produces the Q the truncation kernel consumes), so it is issued from C: the same LAPACK
// routines, in the same order, through the addresses the caller hands in (the Python driver
// takes them from scipy.linalg.cython_lapack -- the OpenBLAS the reference itself runs on).
// The library itself links no LAPACK.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../include/arnoldi_b200.h"

namespace {

typedef struct { double re, im; } zc;
typedef void (*zgees_t)(char* jobvs, char* sort, void* select, int* n, zc* a, int* lda, int* sdim,
                        zc* w, zc* vs, int* ldvs, zc* work, int* lwork, double* rwork, int* bwork,
                        int* info);
typedef void (*ztrexc_t)(char* compq, int* n, zc* t, int* ldt, zc* q, int* ldq, int* ifst,
                         int* ilst, int* info);

thread_local std::vector<zc> g_work;
thread_local std::vector<double> g_rwork;
thread_local std::vector<zc> g_w;

}  // namespace

extern "C" {

// t: column-major m x m, H_m on entry, T1 on exit.  q: column-major m x m, Q1 on exit.
// Same call as scipy.linalg.schur(a, output="complex"): jobvs = 'V', sort = 'N', workspace
// size from a query call.  Returns LAPACK's info (0 = success), or AB200_EINVAL.
int ab200_host_schur(void* zgees_fn, int m, double* t, double* q, double* work) {
  (void)work;
  if (zgees_fn == nullptr || t == nullptr || q == nullptr || m < 1) return AB200_EINVAL;
  zgees_t zgees = reinterpret_cast<zgees_t>(zgees_fn);
  char jobvs = 'V', sort = 'N';
  int n = m, lda = m, ldvs = m, sdim = 0, info = 0, lwork = -1, bwork = 0;
  g_w.resize(m);
  g_rwork.resize(m);
  zc query = {0.0, 0.0};
  zgees(&jobvs, &sort, nullptr, &n, reinterpret_cast<zc*>(t), &lda, &sdim, g_w.data(),
        reinterpret_cast<zc*>(q), &ldvs, &query, &lwork, g_rwork.data(), &bwork, &info);
  if (info != 0) return info;
  lwork = (int)query.re;
  if (lwork < 2 * m) lwork = 2 * m;
  if ((int)g_work.size() < lwork) g_work.resize(lwork);
  zgees(&jobvs, &sort, nullptr, &n, reinterpret_cast<zc*>(t), &lda, &sdim, g_w.data(),
        reinterpret_cast<zc*>(q), &ldvs, g_work.data(), &lwork, g_rwork.data(), &bwork, &info);
  return info;
}

// perm[dest] = index (into the UNSORTED diagonal of T1) of the eigenvalue wanted at slot dest.
// utils.py:52-63: a list tracks where every original entry currently sits; the wanted one is
// moved up to its slot with one ztrexc(T, Z, here + 1, dest + 1) (1-based).  As in the
// reference, the swaps are accumulated into Z starting from the identity (q is overwritten with
// Q2); the caller forms Q = Q1 Q2 with the same zgemm the reference's `Q1 @ Q2` uses, so the
// host arithmetic is the reference's bit for bit.
int ab200_host_reorder(void* ztrexc_fn, int m, double* t, double* q, const int64_t* perm,
                       double* work) {
  (void)work;
  if (ztrexc_fn == nullptr || t == nullptr || q == nullptr || perm == nullptr || m < 1)
    return AB200_EINVAL;
  ztrexc_t ztrexc = reinterpret_cast<ztrexc_t>(ztrexc_fn);
  std::vector<int> slots(m);
  std::vector<char> seen(m, 0);
  for (int i = 0; i < m; ++i) slots[i] = i;
  memset(q, 0, sizeof(double) * 2 * (size_t)m * m);
  for (int i = 0; i < m; ++i) q[2 * ((size_t)i * m + i)] = 1.0;
  for (int dest = 0; dest < m; ++dest) {
    const int64_t original = perm[dest];
    if (original < 0 || original >= m || seen[original]) return AB200_EINVAL;  // not a permutation
    seen[original] = 1;
    int here = dest;
    while (here < m && slots[here] != (int)original) ++here;
    if (here == m) return AB200_EINVAL;
    if (here == dest) continue;
    char compq = 'V';
    int n = m, ldt = m, ldq = m, ifst = here + 1, ilst = dest + 1, info = 0;
    ztrexc(&compq, &n, reinterpret_cast<zc*>(t), &ldt, reinterpret_cast<zc*>(q), &ldq, &ifst, &ilst,
           &info);
    if (info != 0) return info;
    const int moved = slots[here];
    for (int i = here; i > dest; --i) slots[i] = slots[i - 1];
    slots[dest] = moved;
  }
  return 0;
}

}  // extern "C"
