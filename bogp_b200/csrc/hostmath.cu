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
static std::atomic<long long> g_h2d{0};
cudaError_t h2d_async(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  g_h2d.fetch_add((long long)bytes, std::memory_order_relaxed);
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
}
long long h2d_bytes() { return g_h2d.load(std::memory_order_relaxed); }

// ---------------------------------------------------------------- dominant-kernel timing (bench.py roofline)
// When enabled, the launcher of a dominant kernel brackets it with a CUDA event pair recorded on the launching
// stream; bogp_timing_read() synchronises the pairs, sums their elapsed times and recycles them.
static bool g_timing = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_ev_pool, g_ev_live;
bool timing_enabled() { return g_timing; }
cudaEvent_t timing_begin(cudaStream_t s) {
  if (!g_timing) return nullptr;
  std::pair<cudaEvent_t, cudaEvent_t> pr;
  if (!g_ev_pool.empty()) { pr = g_ev_pool.back(); g_ev_pool.pop_back(); }
  else if (cudaEventCreate(&pr.first) != cudaSuccess || cudaEventCreate(&pr.second) != cudaSuccess) return nullptr;
  cudaEventRecord(pr.first, s);
  g_ev_live.push_back(pr);
  return pr.second;
}
void timing_end(cudaEvent_t stop, cudaStream_t s) { if (stop) cudaEventRecord(stop, s); }

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

int hermitian_eig_factor(const std::vector<cplx>& A, int n, double tol, std::vector<cplx>& L, int& rank) {
  // A = V diag(lambda) V^H by cyclic complex Jacobi rotations (n <= a few hundred, once per mode set); returns the
  // optimal low-rank factor L = V_r diag(lambda_r)^1/2 over the eigenvalues above tol * lambda_max, columns in
  // descending order, laid out like pivoted_cholesky's L ([n, n], column j = j-th factor column): A ~= L L^H with
  // the FEWEST columns for a given accuracy (a pivoted Cholesky factor needs more for the same tolerance).
  std::vector<cplx> B(A), V((size_t)n * n, cplx(0.0, 0.0));
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  auto off2 = [&]() { double s = 0.0; for (int p = 0; p < n; ++p) for (int q = p + 1; q < n; ++q) s += std::norm(B[(size_t)p * n + q]); return s; };
  double diag2 = 0.0;
  for (int i = 0; i < n; ++i) diag2 += B[(size_t)i * n + i].real() * B[(size_t)i * n + i].real();
  const double stop = 1e-30 * (diag2 + 2.0 * off2());
  for (int sweep = 0; sweep < 60 && off2() > stop; ++sweep) {
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const cplx bpq = B[(size_t)p * n + q];
        const double ab = std::abs(bpq);
        if (ab == 0.0) continue;
        const double app = B[(size_t)p * n + p].real(), aqq = B[(size_t)q * n + q].real();
        if (ab <= 1e-18 * std::sqrt(std::fabs(app * aqq)) && sweep > 2) continue;
        const cplx ph = bpq / ab;                       // e^{i phi}
        const double theta = (aqq - app) / (2.0 * ab);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        const cplx sp = s * std::conj(ph), cp = c * std::conj(ph);     // J = [[c, s], [-s e^{-i phi}, c e^{-i phi}]]
        for (int i = 0; i < n; ++i) {                   // B <- B J,  V <- V J  (columns p, q)
          const cplx bp = B[(size_t)i * n + p], bq = B[(size_t)i * n + q];
          B[(size_t)i * n + p] = c * bp - sp * bq;
          B[(size_t)i * n + q] = s * bp + cp * bq;
          const cplx vp = V[(size_t)i * n + p], vq = V[(size_t)i * n + q];
          V[(size_t)i * n + p] = c * vp - sp * vq;
          V[(size_t)i * n + q] = s * vp + cp * vq;
        }
        for (int j = 0; j < n; ++j) {                   // B <- J^H B  (rows p, q)
          const cplx bp = B[(size_t)p * n + j], bq = B[(size_t)q * n + j];
          B[(size_t)p * n + j] = c * bp - std::conj(sp) * bq;
          B[(size_t)q * n + j] = s * bp + std::conj(cp) * bq;
        }
        B[(size_t)p * n + q] = B[(size_t)q * n + p] = 0.0;
        B[(size_t)p * n + p] = B[(size_t)p * n + p].real();
        B[(size_t)q * n + q] = B[(size_t)q * n + q].real();
      }
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return B[(size_t)a * n + a].real() > B[(size_t)b * n + b].real(); });
  L.assign((size_t)n * n, cplx(0.0, 0.0));
  rank = 0;
  const double lmax = n ? B[(size_t)idx[0] * n + idx[0]].real() : 0.0;
  if (!(lmax > 0.0)) return 0;
  for (int j = 0; j < n; ++j) {
    const double lam = B[(size_t)idx[j] * n + idx[j]].real();
    if (!(lam > tol * lmax)) break;
    const double sl = std::sqrt(lam);
    for (int i = 0; i < n; ++i) L[(size_t)i * n + j] = V[(size_t)i * n + idx[j]] * sl;
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
extern "C" long long bogp_h2d_bytes(void) { return bogp::h2d_bytes(); }
extern "C" int bogp_timing_enable(int on) { bogp::g_timing = on != 0; return 0; }
extern "C" int bogp_timing_read(double* total_ms, long long* count) {
  double tot = 0.0;
  long long n = 0;
  for (auto& pr : bogp::g_ev_live) {
    float ms = 0.f;
    if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
      tot += ms;
      ++n;
    }
    bogp::g_ev_pool.push_back(pr);
  }
  bogp::g_ev_live.clear();
  if (total_ms) *total_ms = tot;
  if (count) *count = n;
  return 0;
}
