// Host-side float64 linear algebra used at set-up time (factor the covariance / Gram matrices once;
// the per-candidate work is all on the GPU).
#include <algorithm>
#include <atomic>
#include <cmath>

#include "bogp_internal.h"

namespace bogp {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

const char* last_error_cstr() { return g_err.c_str(); }

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int pivoted_cholesky(const std::vector<cplx>& A, int n, double tol, std::vector<cplx>& L, int& rank) {
  // Outer-product pivoted Cholesky; L is returned in ORIGINAL row order, column j = j-th pivot step.
  std::vector<double> d(n);
  std::vector<int> perm(n);
  double dmax0 = 0.0;
  for (int i = 0; i < n; ++i) {
    d[i] = A[(size_t)i * n + i].real();
    perm[i] = i;
    dmax0 = std::max(dmax0, d[i]);
  }
  L.assign((size_t)n * n, cplx(0.0, 0.0));
  rank = 0;
  if (!(dmax0 > 0.0)) return 0;
  for (int j = 0; j < n; ++j) {
    int p = j;
    for (int i = j + 1; i < n; ++i)
      if (d[perm[i]] > d[perm[p]]) p = i;
    if (!(d[perm[p]] > tol * dmax0)) break;
    std::swap(perm[j], perm[p]);
    const int pj = perm[j];
    const double ljj = std::sqrt(d[pj]);
    L[(size_t)pj * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
      const int pi = perm[i];
      cplx s = A[(size_t)pi * n + pj];
      for (int k = 0; k < j; ++k) s -= L[(size_t)pi * n + k] * std::conj(L[(size_t)pj * n + k]);
      s /= ljj;
      L[(size_t)pi * n + j] = s;
      d[pi] -= std::norm(s);
    }
    rank = j + 1;
  }
  return 0;
}

bool cholesky_lower(std::vector<double>& A, int n) {
  // Row-oriented (Cholesky-Banachiewicz): inner dot products run along contiguous rows.
  for (int i = 0; i < n; ++i) {
    double* Ai = &A[(size_t)i * n];
    for (int j = 0; j <= i; ++j) {
      const double* Aj = &A[(size_t)j * n];
      double s = Ai[j];
      for (int k = 0; k < j; ++k) s -= Ai[k] * Aj[k];
      if (i == j) {
        if (!(s > 0.0) || !std::isfinite(s)) return false;
        Ai[j] = std::sqrt(s);
      } else {
        Ai[j] = s / Aj[j];
      }
    }
    for (int j = i + 1; j < n; ++j) Ai[j] = 0.0;
  }
  return true;
}

void invert_lower(std::vector<double>& L, int n) {
  // X = L^-1 by forward substitution on the columns of I, computed row by row:
  // X[i, j] = (delta_ij - sum_{k=j}^{i-1} L[i,k] X[k,j]) / L[i,i].  X^T is kept so the k-loop is contiguous.
  std::vector<double> Xt((size_t)n * n, 0.0);   // Xt[j*n + i] = X[i, j]
  for (int j = 0; j < n; ++j) {
    double* xj = &Xt[(size_t)j * n];
    for (int i = j; i < n; ++i) {
      const double* Li = &L[(size_t)i * n];
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = j; k < i; ++k) s -= Li[k] * xj[k];
      xj[i] = s / Li[i];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) L[(size_t)i * n + j] = (j <= i) ? Xt[(size_t)j * n + i] : 0.0;
}

}  // namespace bogp

extern "C" const char* bogp_last_error(void) { return bogp::last_error_cstr(); }
extern "C" int bogp_version(void) { return 100; }
extern "C" long long bogp_kernel_launches(void) { return bogp::launches(); }
