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
#include <math.h>
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

typedef void (*dgees_t)(char* jobvs, char* sort, void* select, int* n, double* a, int* lda,
                        int* sdim, double* wr, double* wi, double* vs, int* ldvs, double* work,
                        int* lwork, int* bwork, int* info);
typedef void (*dsyevd_t)(char* jobz, char* uplo, int* n, double* a, int* lda, double* w, double* work,
                         int* lwork, int* iwork, int* liwork, int* info);

thread_local std::vector<double> g_da, g_dvs, g_dwork, g_wr, g_wi;
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

// Complex Schur form of a REAL matrix through the real routine: dgees (a third of zgees's
// arithmetic), then every 2 x 2 block of the quasi-triangular factor -- a complex-conjugate
// pair -- is triangularised by one complex Givens rotation (the construction of
// scipy.linalg.rsf2csf).  With only real eigenvalues the result is real.  h: column-major
// complex m x m whose imaginary parts are all zero (checked; AB200_EINVAL otherwise).
// A valid Schur form of the same matrix as ab200_host_schur's, not the same one: diagonal
// order before sorting, signs and phases of the vectors and the rounding differ.
int ab200_host_schur_real(void* dgees_fn, int m, double* t, double* q) {
  if (dgees_fn == nullptr || t == nullptr || q == nullptr || m < 1) return AB200_EINVAL;
  dgees_t dgees = reinterpret_cast<dgees_t>(dgees_fn);
  const size_t mm = (size_t)m * m;
  g_da.resize(mm);
  g_dvs.resize(mm);
  g_wr.resize(m);
  g_wi.resize(m);
  for (size_t i = 0; i < mm; ++i) {
    if (t[2 * i + 1] != 0.0) return AB200_EINVAL;
    g_da[i] = t[2 * i];
  }
  char jobvs = 'V', sort = 'N';
  int n = m, lda = m, ldvs = m, sdim = 0, info = 0, lwork = -1, bwork = 0;
  double query = 0.0;
  dgees(&jobvs, &sort, nullptr, &n, g_da.data(), &lda, &sdim, g_wr.data(), g_wi.data(),
        g_dvs.data(), &ldvs, &query, &lwork, &bwork, &info);
  if (info != 0) return info;
  lwork = (int)query;
  if (lwork < 3 * m) lwork = 3 * m;
  if ((int)g_dwork.size() < lwork) g_dwork.resize(lwork);
  dgees(&jobvs, &sort, nullptr, &n, g_da.data(), &lda, &sdim, g_wr.data(), g_wi.data(),
        g_dvs.data(), &ldvs, g_dwork.data(), &lwork, &bwork, &info);
  if (info != 0) return info;
  zc* T = reinterpret_cast<zc*>(t);
  zc* Z = reinterpret_cast<zc*>(q);
  for (size_t i = 0; i < mm; ++i) {
    T[i].re = g_da[i], T[i].im = 0.0;
    Z[i].re = g_dvs[i], Z[i].im = 0.0;
  }
  auto at = [m](zc* M, int r, int c) -> zc& { return M[(size_t)c * m + r]; };
  auto mul = [](zc a, zc b) { zc r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; };
  auto add = [](zc a, zc b) { zc r = {a.re + b.re, a.im + b.im}; return r; };
  const double eps = 2.220446049250313e-16;
  for (int k = m - 1; k >= 1; --k) {
    const double sub = at(T, k, k - 1).re;
    if (fabs(sub) <= eps * (fabs(at(T, k - 1, k - 1).re) + fabs(at(T, k, k).re))) {
      at(T, k, k - 1).re = 0.0;
      continue;
    }
    // eigenvalue (positive imaginary part) of the block minus its (k, k) entry
    const double a = at(T, k - 1, k - 1).re, b = at(T, k - 1, k).re, d = at(T, k, k).re;
    const double half = 0.5 * (a - d), disc = half * half + b * sub;
    zc mu;
    if (disc < 0.0) {
      mu.re = half, mu.im = sqrt(-disc);
    } else {  // a real pair left as a block: rotate to its larger root
      mu.re = half + (half >= 0 ? sqrt(disc) : -sqrt(disc)), mu.im = 0.0;
    }
    const double r = sqrt(mu.re * mu.re + mu.im * mu.im + sub * sub);
    const zc c = {mu.re / r, mu.im / r};
    const zc cc = {c.re, -c.im};
    const double sn = sub / r;
    const zc s_p = {sn, 0.0}, s_m = {-sn, 0.0};
    // T[k-1:k+1, k-1:] = G T[k-1:k+1, k-1:]   with G = [[conj(c), s], [-s, c]]
    for (int j = k - 1; j < m; ++j) {
      const zc x = at(T, k - 1, j), y = at(T, k, j);
      at(T, k - 1, j) = add(mul(cc, x), mul(s_p, y));
      at(T, k, j) = add(mul(s_m, x), mul(c, y));
    }
    // T[:k+1, k-1:k+1] = T[:k+1, k-1:k+1] G^H ;  Z[:, k-1:k+1] = Z[:, k-1:k+1] G^H
    // G^H = [[c, -s], [s, conj(c)]]
    for (int i = 0; i <= k; ++i) {
      const zc x = at(T, i, k - 1), y = at(T, i, k);
      at(T, i, k - 1) = add(mul(x, c), mul(y, s_p));
      at(T, i, k) = add(mul(x, s_m), mul(y, cc));
    }
    for (int i = 0; i < m; ++i) {
      const zc x = at(Z, i, k - 1), y = at(Z, i, k);
      at(Z, i, k - 1) = add(mul(x, c), mul(y, s_p));
      at(Z, i, k) = add(mul(x, s_m), mul(y, cc));
    }
    at(T, k, k - 1).re = 0.0, at(T, k, k - 1).im = 0.0;
  }
  return 0;
}

// Schur form of a real SYMMETRIC matrix = its eigendecomposition: dsyevd on (H + H^T) / 2, a
// quarter of dgees's time at m = 40, and the ordered form is then a permutation of the columns
// (no ztrexc moves).  The projected matrix of a symmetric operator is symmetric up to the
// rounding of the orthogonalisation coefficients; the routine measures max |H - H^T| and
// declines (returns 1, nothing written) when it exceeds sym_tol * max |H| -- the caller then
// takes ab200_host_schur_real.  On success T is diagonal (ascending) and Q real orthogonal, both
// written as complex column-major m x m.  Like ab200_host_schur_real this is a valid Schur form
// of H to within sym_tol, not the reference's bit for bit.
int ab200_host_eigh_real(void* dsyevd_fn, int m, double* t, double* q, double sym_tol) {
  if (dsyevd_fn == nullptr || t == nullptr || q == nullptr || m < 1 || !(sym_tol >= 0.0)) return AB200_EINVAL;
  dsyevd_t dsyevd = reinterpret_cast<dsyevd_t>(dsyevd_fn);
  const size_t mm = (size_t)m * m;
  double hmax = 0.0, asym = 0.0;
  for (int c = 0; c < m; ++c)
    for (int r = 0; r < m; ++r) {
      const size_t i = (size_t)c * m + r, j = (size_t)r * m + c;
      if (t[2 * i + 1] != 0.0) return AB200_EINVAL;
      const double a = fabs(t[2 * i]), d = fabs(t[2 * i] - t[2 * j]);
      hmax = a > hmax ? a : hmax;
      asym = d > asym ? d : asym;
    }
  if (asym > sym_tol * hmax) return 1;
  g_da.resize(mm);
  g_wr.resize(m);
  for (int c = 0; c < m; ++c)
    for (int r = 0; r < m; ++r)
      g_da[(size_t)c * m + r] = 0.5 * (t[2 * ((size_t)c * m + r)] + t[2 * ((size_t)r * m + c)]);
  char jobz = 'V', uplo = 'L';
  int n = m, lda = m, info = 0, lwork = -1, liwork = -1, iquery = 0;
  double query = 0.0;
  dsyevd(&jobz, &uplo, &n, g_da.data(), &lda, g_wr.data(), &query, &lwork, &iquery, &liwork, &info);
  if (info != 0) return info < 0 ? AB200_EINVAL : info + 1;
  lwork = (int)query;
  liwork = iquery;
  if (lwork < 1 + 6 * m + 2 * m * m) lwork = 1 + 6 * m + 2 * m * m;
  if (liwork < 3 + 5 * m) liwork = 3 + 5 * m;
  if ((int)g_dwork.size() < lwork) g_dwork.resize(lwork);
  std::vector<int> iwork(liwork);
  dsyevd(&jobz, &uplo, &n, g_da.data(), &lda, g_wr.data(), g_dwork.data(), &lwork, iwork.data(), &liwork,
         &info);
  if (info != 0) return info < 0 ? AB200_EINVAL : info + 1;
  memset(t, 0, sizeof(double) * 2 * mm);
  for (int i = 0; i < m; ++i) t[2 * ((size_t)i * m + i)] = g_wr[i];
  for (size_t i = 0; i < mm; ++i) q[2 * i] = g_da[i], q[2 * i + 1] = 0.0;
  return 0;
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
