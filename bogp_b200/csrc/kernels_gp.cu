// Boundary B kernels: batched exact-GP posterior with fused analytic acquisition (EI / LogEI / PI / UCB),
// analytic gradient wrt the candidates, joint posterior for q > 1 and MC qEI.
//
// Replaces, for a batch of candidates, what BoTorch/GPyTorch do per `optimize_acqf` evaluation:
//   SingleTaskGP.posterior(X)  (ref/projects/swellex96_inv/data/bo/ei.py:67, ref/optimization/gpmodels.py:27-30)
//   ExpectedImprovement / LogExpectedImprovement / ProbabilityOfImprovement / UpperConfidenceBound .forward(X)
//   (ei.py:80, logei.py:80, pi.py:81, ucb.py:82) and the autograd backward optimize_acqf's L-BFGS-B needs.
//
// All arithmetic is float64: SURVEY 7.3 item 3 measured up to 18 % error in the posterior variance
// (k** - ||L^-1 k*||^2 cancels) when any part of it is float32.
//
// One CTA owns BC candidates.  Phases (block-wide barriers between them):
//   1  Ks[i][c] = s k(r_ic), gk[i][c] = (dk/dx factor)          thread <-> (i, c) pairs
//   2  V = L^-1 Ks   (lower triangular, streamed from L2)        thread <-> rows i (paired i, n-1-i for balance)
//   3  mean_c = m + alpha . Ks_c ; var_c = s - ||V_c||^2         warp <-> candidate c
//      -> acquisition value and d acq/d mean, d acq/d var
//   4  W = L^-T V = K^-1 Ks                                      thread <-> columns j
//   5  grad_c = sum_i (dA/dmu alpha_i - 2 dA/dvar W_ic) gk_ic (x_c - X_i)/l^2    warp <-> candidate c
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "bogp_internal.h"

namespace bogp {

constexpr int kGpThreads = 256;
constexpr int kGpWarps = kGpThreads / 32;
constexpr int kGpMaxD = 16;
constexpr int kGpMaxN = 1024;
constexpr int kGpMaxQ = 8;
constexpr double kMinVariance = 1e-10;      // gpytorch.settings.min_variance for float64 [recalled, SURVEY App. A]
constexpr double kAcqMinVariance = 1e-12;   // botorch analytic acquisitions clamp [recalled]

struct GpDev {
  int n, d, kind;
  double outputscale, mean, noise;
  const double* X;       // [n][d]
  const double* inv_ls;  // [d]
  const double* alpha;   // [n]
  const double* Linv;    // [n][n]  Linv[i][j], j <= i
  const double* LinvT;   // [n][n]  LinvT[j][i] = Linv[i][j]
};

enum GpMode { kModePosterior = 0, kModeAcq = 1, kModeJointFwd = 2, kModeJointBwd = 3 };

struct GpArgs {
  int mode;
  int acq_kind;
  int q;                 // joint modes: candidates per batch element (<= BC); else 1
  int observation_noise;
  double best_f, beta;
  long long b;           // batch elements
  const double* Xs;      // [b][q][d]
  double* out0;          // posterior: mean[b]      | acq: value[b]   | joint: mean[b][q]
  double* out1;          // posterior: var[b]       | acq: unused     | joint: cov[b][q][q]
  double* grad;          // acq: [b][d] or null     | joint bwd: [b][q][d]
  const double* gmean;   // joint bwd: upstream d/d mean [b][q]
  const double* gcov;    // joint bwd: upstream d/d cov  [b][q][q]
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// k(r) and the factor g with  dk/dx_a = g * (x_a - X_a) / l_a^2   (both include the outputscale).
__device__ __forceinline__ void kernel_eval(int kind, double s, double d2, double& k, double& g) {
  if (kind == BOGP_KERNEL_RBF) {
    k = s * exp(-0.5 * d2);
    g = -k;
  } else {
    const double sqrt5 = 2.23606797749978969641;
    const double r = sqrt(fmax(d2, 1e-30));
    const double e = s * exp(-sqrt5 * r);
    k = (1.0 + sqrt5 * r + (5.0 / 3.0) * d2) * e;
    g = -(5.0 / 3.0) * (1.0 + sqrt5 * r) * e;
    if (d2 < 1e-30) g = 0.0;   // the oracle's clamp has zero gradient below the floor
  }
}

__device__ __forceinline__ double norm_pdf(double u) { return 0.39894228040143267794 * exp(-0.5 * u * u); }
__device__ __forceinline__ double norm_cdf(double u) { return 0.5 * erfc(-u * 0.70710678118654752440); }
__device__ __forceinline__ double log1mexp(double x) {
  return (x > -0.69314718055994530942) ? log(-expm1(x)) : log1p(-exp(x));
}

// log(phi(u) + u Phi(u)) and rho = d/du of it (botorch _log_ei_helper [recalled]; oracle/gp.py log_ei_helper).
__device__ __forceinline__ void log_ei_helper(double u, double& val, double& rho) {
  const double inv_sqrt2 = 0.70710678118654752440, inv_sqrt2pi = 0.39894228040143267794;
  if (u > -1.0) {
    const double h = norm_pdf(u) + u * norm_cdf(u);
    val = log(h);
    rho = norm_cdf(u) / h;
    return;
  }
  const double log_phi = -0.5 * u * u - 0.91893853320467274178;
  if (u > -1e6) {
    const double ex = erfcx(-u * inv_sqrt2);
    const double w = log(ex * fabs(u)) + 0.22579135264472743236;   // + 0.5*log(pi/2)
    val = log_phi + log1mexp(w);
    rho = 0.5 * ex / (inv_sqrt2pi + 0.5 * u * ex);
  } else {
    val = log_phi - 2.0 * log(fabs(u));
    rho = -u - 2.0 / u;
  }
}

// Acquisition value and partial derivatives wrt posterior mean and (unclamped) variance.
__device__ __forceinline__ void acquisition(int kind, double mu, double var_raw, double best_f, double beta,
                                            double& val, double& d_mu, double& d_var) {
  const bool clamped = !(var_raw > kMinVariance);
  const double var = clamped ? kMinVariance : var_raw;
  const double sigma = sqrt(fmax(var, kAcqMinVariance));
  double d_sigma;
  if (kind == BOGP_ACQ_UCB) {
    const double sb = sqrt(beta);
    val = mu + sb * sigma;
    d_mu = 1.0;
    d_sigma = sb;
  } else {
    const double u = (mu - best_f) / sigma;
    if (kind == BOGP_ACQ_EI) {
      const double pdf = norm_pdf(u), cdf = norm_cdf(u);
      val = sigma * (pdf + u * cdf);
      d_mu = cdf;
      d_sigma = pdf;
    } else if (kind == BOGP_ACQ_LOGEI) {
      double h, rho;
      log_ei_helper(u, h, rho);
      val = h + log(sigma);
      d_mu = rho / sigma;
      d_sigma = (1.0 - u * rho) / sigma;
    } else {   // PI
      const double pdf = norm_pdf(u);
      val = norm_cdf(u);
      d_mu = pdf / sigma;
      d_sigma = -u * pdf / sigma;
    }
  }
  d_var = clamped ? 0.0 : d_sigma / (2.0 * sigma);
}

// Row owned by thread t in slot k: slots alternate ascending / descending so that every thread sees the
// same triangular work (row t pairs with row 2*T-1-t, ...).
__device__ __forceinline__ int slot_row(int t, int k) {
  return (k & 1) ? (k + 1) * kGpThreads - 1 - t : k * kGpThreads + t;
}

template <int BC>
__global__ void __launch_bounds__(kGpThreads)
gp_eval_kernel(GpDev gp, GpArgs a) {
  extern __shared__ double sm[];
  const int n = gp.n, d = gp.d, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  double* xs = sm;                                   // [BC][kGpMaxD]
  double* coef = xs + BC * kGpMaxD;                  // [BC][4]: dA/dmu, dA/dvar, (spare)
  double* S = coef + BC * 4;                         // [BC][BC] joint bwd: gcov + gcov^T
  double* Ks = S + BC * BC;                          // [n][BC]   (reused for W in phase 4)
  double* V = Ks + (size_t)n * BC;                   // [n][BC]
  double* gk = V + (size_t)n * BC;                   // [n][BC]   (gradient modes only)
  const bool joint = a.mode == kModeJointFwd || a.mode == kModeJointBwd;
  const bool want_grad = (a.mode == kModeAcq && a.grad != nullptr) || a.mode == kModeJointBwd;
  const int q = joint ? a.q : 1;
  const int per_cta = joint ? 1 : BC;                // batch elements per CTA
  const long long total = a.b * (long long)q;        // candidates overall
  const int n_slots = (n + kGpThreads - 1) / kGpThreads;

  for (long long blk = blockIdx.x; blk * per_cta < a.b; blk += gridDim.x) {
    const long long c0 = blk * (joint ? q : BC);     // first candidate (flat index into Xs rows)
    const int nc = joint ? q : (int)min((long long)BC, total - c0);
    __syncthreads();
    for (int idx = t; idx < BC * kGpMaxD; idx += kGpThreads) {
      const int c = idx / kGpMaxD, dim = idx % kGpMaxD;
      const int cc = c < nc ? c : nc - 1;            // padded candidates replicate the last real one
      xs[idx] = dim < d ? a.Xs[(c0 + cc) * d + dim] : 0.0;
    }
    if (a.mode == kModeJointBwd) {
      for (int idx = t; idx < BC * BC; idx += kGpThreads) {
        const int j = idx / BC, l = idx % BC;
        S[idx] = (j < q && l < q) ? a.gcov[(blk * q + j) * q + l] + a.gcov[(blk * q + l) * q + j] : 0.0;
      }
    }
    __syncthreads();
    // ---- phase 1: cross-covariances
    for (int idx = t; idx < n * BC; idx += kGpThreads) {
      const int i = idx / BC, c = idx % BC;
      double d2 = 0.0;
      for (int dim = 0; dim < d; ++dim) {
        const double df = (xs[c * kGpMaxD + dim] - __ldg(gp.X + (size_t)i * d + dim)) * __ldg(gp.inv_ls + dim);
        d2 = fma(df, df, d2);
      }
      double k, g;
      kernel_eval(gp.kind, gp.outputscale, d2, k, g);
      Ks[idx] = k;
      if (want_grad) gk[idx] = g;
    }
    __syncthreads();
    // ---- phase 2: V = L^-1 Ks
    for (int k = 0; k < n_slots; ++k) {
      const int row = slot_row(t, k);
      if (row < n) {
        double acc[BC];
#pragma unroll
        for (int c = 0; c < BC; ++c) acc[c] = 0.0;
        const double* col = gp.LinvT + row;            // LinvT[j][row]
#pragma unroll 4
        for (int j = 0; j <= row; ++j) {
          const double l = __ldg(col + (size_t)j * n);
#pragma unroll
          for (int c = 0; c < BC; ++c) acc[c] = fma(l, Ks[j * BC + c], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < BC; ++c) V[row * BC + c] = acc[c];
      }
    }
    __syncthreads();
    // ---- phase 3: moments (+ acquisition)
    if (!joint) {
      for (int c = warp; c < nc; c += kGpWarps) {
        double m = 0.0, ss = 0.0;
        for (int i = lane; i < n; i += 32) {
          m = fma(__ldg(gp.alpha + i), Ks[i * BC + c], m);
          const double v = V[i * BC + c];
          ss = fma(v, v, ss);
        }
        m = warp_sum_d(m) + gp.mean;
        ss = warp_sum_d(ss);
        double var = gp.outputscale - ss;
        if (lane == 0) {
          if (a.mode == kModePosterior) {
            if (a.observation_noise) var += gp.noise;
            a.out0[c0 + c] = m;
            a.out1[c0 + c] = fmax(var, kMinVariance);
          } else {
            double val, dmu, dvar;
            acquisition(a.acq_kind, m, var, a.best_f, a.beta, val, dmu, dvar);
            a.out0[c0 + c] = val;
            coef[c * 4 + 0] = dmu;
            coef[c * 4 + 1] = dvar;
          }
        }
      }
    } else if (a.mode == kModeJointFwd) {
      // mean[q] and cov[q][q]: one warp per (j, l) pair, j <= l, plus one per mean
      for (int p = warp; p < q * q + q; p += kGpWarps) {
        if (p < q) {
          double m = 0.0;
          for (int i = lane; i < n; i += 32) m = fma(__ldg(gp.alpha + i), Ks[i * BC + p], m);
          m = warp_sum_d(m);
          if (lane == 0) a.out0[blk * q + p] = m + gp.mean;
        } else {
          const int j = (p - q) / q, l = (p - q) % q;
          if (j > l) continue;
          double ss = 0.0;
          for (int i = lane; i < n; i += 32) ss = fma(V[i * BC + j], V[i * BC + l], ss);
          ss = warp_sum_d(ss);
          if (lane == 0) {
            double d2 = 0.0;
            for (int dim = 0; dim < d; ++dim) {
              const double df = (xs[j * kGpMaxD + dim] - xs[l * kGpMaxD + dim]) * gp.inv_ls[dim];
              d2 = fma(df, df, d2);
            }
            double kk, g;
            kernel_eval(gp.kind, gp.outputscale, d2, kk, g);
            if (j == l) kk = gp.outputscale;
            double cv = kk - ss;
            if (j == l && a.observation_noise) cv += gp.noise;
            a.out1[(blk * q + j) * q + l] = cv;
            a.out1[(blk * q + l) * q + j] = cv;
          }
        }
      }
    }
    if (!want_grad) continue;
    __syncthreads();
    // ---- phase 4: W = L^-T V  (into the Ks buffer; Ks itself is no longer needed)
    double* W = Ks;
    for (int k = 0; k < n_slots; ++k) {
      const int colj = slot_row(t, k);
      if (colj < n) {
        double acc[BC];
#pragma unroll
        for (int c = 0; c < BC; ++c) acc[c] = 0.0;
        const double* rowp = gp.Linv + colj;           // Linv[i][colj]
#pragma unroll 4
        for (int i = colj; i < n; ++i) {
          const double l = __ldg(rowp + (size_t)i * n);
#pragma unroll
          for (int c = 0; c < BC; ++c) acc[c] = fma(l, V[i * BC + c], acc[c]);
        }
        // W aliases Ks, which other threads may still read in this phase?  No: phase 4 reads only V.
#pragma unroll
        for (int c = 0; c < BC; ++c) W[colj * BC + c] = acc[c];
      }
    }
    __syncthreads();
    // ---- phase 5: gradients, warp <-> candidate
    for (int c = warp; c < nc; c += kGpWarps) {
      double g_acc[kGpMaxD];
#pragma unroll
      for (int dim = 0; dim < kGpMaxD; ++dim) g_acc[dim] = 0.0;
      double dmu, dvar;
      if (!joint) { dmu = coef[c * 4 + 0]; dvar = coef[c * 4 + 1]; } else { dmu = a.gmean[blk * q + c]; dvar = 0.0; }
      for (int i = lane; i < n; i += 32) {
        double cf;
        if (!joint) {
          cf = dmu * __ldg(gp.alpha + i) - 2.0 * dvar * W[i * BC + c];
        } else {
          cf = dmu * __ldg(gp.alpha + i);
          for (int l = 0; l < q; ++l) cf -= S[c * BC + l] * W[i * BC + l];
        }
        cf *= gk[i * BC + c];
#pragma unroll
        for (int dim = 0; dim < kGpMaxD; ++dim)
          if (dim < d) g_acc[dim] = fma(cf, xs[c * kGpMaxD + dim] - __ldg(gp.X + (size_t)i * d + dim), g_acc[dim]);
      }
#pragma unroll
      for (int dim = 0; dim < kGpMaxD; ++dim)
        if (dim < d) g_acc[dim] = warp_sum_d(g_acc[dim]);
      if (lane == 0) {
        if (joint) {
          // + sum_{l != c} S[c][l] dk(x_c, x_l)/dx_c
          for (int l = 0; l < q; ++l) {
            if (l == c) continue;
            double d2 = 0.0;
            for (int dim = 0; dim < d; ++dim) {
              const double df = (xs[c * kGpMaxD + dim] - xs[l * kGpMaxD + dim]) * gp.inv_ls[dim];
              d2 = fma(df, df, d2);
            }
            double kk, g;
            kernel_eval(gp.kind, gp.outputscale, d2, kk, g);
            const double sg = S[c * BC + l] * g;
#pragma unroll
            for (int dim = 0; dim < kGpMaxD; ++dim)
              if (dim < d) g_acc[dim] = fma(sg, xs[c * kGpMaxD + dim] - xs[l * kGpMaxD + dim], g_acc[dim]);
          }
        }
#pragma unroll
        for (int dim = 0; dim < kGpMaxD; ++dim)
          if (dim < d) a.grad[(c0 + c) * d + dim] = g_acc[dim] * gp.inv_ls[dim] * gp.inv_ls[dim];
      }
    }
  }
}

// MC qEI from the joint posterior: one warp per batch element.
//   L = chol(cov) (jitter retries 1e-8, 1e-7, 1e-6 as gpytorch psd_safe_cholesky [recalled]),
//   value = mean_s max_j relu(mean_j + (L z_s)_j - best_f).
// When gmean/gcov are non-null also emits d value / d mean [b][q] and d value / d cov [b][q][q]
// (reverse-mode through the samples and the Cholesky factor) for the joint backward kernel.
__global__ void __launch_bounds__(128)
qei_kernel(const double* __restrict__ mean, const double* __restrict__ cov, long long b, int q,
           const double* __restrict__ base, int S, double best_f, double* __restrict__ out,
           double* __restrict__ gmean, double* __restrict__ gcov, int* __restrict__ fail) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long e = (long long)blockIdx.x * 4 + warp;
  if (e >= b) return;
  double A[kGpMaxQ][kGpMaxQ], L[kGpMaxQ][kGpMaxQ], mu[kGpMaxQ];
  for (int j = 0; j < q; ++j) {
    mu[j] = mean[e * q + j];
    for (int l = 0; l < q; ++l) A[j][l] = cov[(e * q + j) * q + l];
  }
  // every lane factorises the same q x q matrix (q <= 8: cheaper than broadcasting)
  bool ok = false;
  double jit = 0.0;
  for (int attempt = 0; attempt < 4 && !ok; ++attempt) {
    jit = attempt == 0 ? 0.0 : 1e-8 * pow(10.0, attempt - 1);
    ok = true;
    for (int j = 0; j < q && ok; ++j) {
      for (int l = 0; l <= j; ++l) {
        double s = A[j][l] + (j == l ? jit : 0.0);
        for (int k = 0; k < l; ++k) s -= L[j][k] * L[l][k];
        if (j == l) {
          if (!(s > 0.0) || !isfinite(s)) { ok = false; break; }
          L[j][j] = sqrt(s);
        } else {
          L[j][l] = s / L[l][l];
        }
      }
    }
  }
  if (!ok) {
    if (lane == 0) { out[e] = nan(""); atomicExch(fail, 1); }
    return;
  }
  double acc = 0.0;
  double gm[kGpMaxQ], gL[kGpMaxQ][kGpMaxQ];
  const bool want = gmean != nullptr;
  if (want)
    for (int j = 0; j < q; ++j) { gm[j] = 0.0; for (int l = 0; l < q; ++l) gL[j][l] = 0.0; }
  for (int s = lane; s < S; s += 32) {
    double best = 0.0;
    int arg = -1;
    for (int j = 0; j < q; ++j) {
      double v = mu[j] - best_f;
      for (int l = 0; l <= j; ++l) v = fma(L[j][l], base[(size_t)s * q + l], v);
      if (v > best) { best = v; arg = j; }
    }
    acc += best;
    if (want && arg >= 0) {
      gm[arg] += 1.0;
      for (int l = 0; l <= arg; ++l) gL[arg][l] += base[(size_t)s * q + l];
    }
  }
  acc = warp_sum_d(acc);
  if (lane == 0) out[e] = acc / (double)S;
  if (!want) return;
  const double invS = 1.0 / (double)S;
  for (int j = 0; j < q; ++j) {
    gm[j] = warp_sum_d(gm[j]) * invS;
    for (int l = 0; l <= j; ++l) gL[j][l] = warp_sum_d(gL[j][l]) * invS;
  }
  if (lane != 0) return;
  // Cholesky reverse mode (unblocked, Murray 2016): given gL (lower), produce gA (symmetric).
  for (int j = q - 1; j >= 0; --j) {
    for (int l = j; l >= 0; --l) {
      if (j == l) {
        gL[j][j] = 0.5 * gL[j][j] / L[j][j];
      } else {
        gL[j][l] = gL[j][l] / L[l][l];
        gL[l][l] -= gL[j][l] * L[j][l];
      }
      for (int k = 0; k < l; ++k) {
        gL[j][k] -= gL[j][l] * L[l][k] * (j == l ? 2.0 : 1.0);
        if (j != l) gL[l][k] -= gL[j][l] * L[j][k];
      }
    }
  }
  for (int j = 0; j < q; ++j) {
    gmean[e * q + j] = gm[j];
    for (int l = 0; l < q; ++l) {
      // gL now holds dF/dA for the lower triangle where A[j][l] and A[l][j] are the same variable;
      // the joint backward symmetrises (S = g + g^T), so hand it half of each off-diagonal term.
      const double g = j == l ? gL[j][j] : 0.5 * (j > l ? gL[j][l] : gL[l][j]);
      gcov[(e * q + j) * q + l] = g;
    }
  }
}

static size_t gp_smem_bytes(int n, int BC, bool grad) {
  return sizeof(double) * ((size_t)BC * kGpMaxD + BC * 4 + BC * BC + (size_t)n * BC * (grad ? 3 : 2));
}

template <int BC>
static int launch_gp(const bogp_gp_s* g, const GpArgs& a, bool grad, cudaStream_t s) {
  GpDev dev{g->n, g->d, g->kind, g->outputscale, g->mean, g->noise, g->X, g->inv_ls, g->alpha, g->Linv, g->LinvT};
  const size_t smem = gp_smem_bytes(g->n, BC, grad);
  BOGP_REQUIRE(smem <= 227 * 1024, "GP kernel: n=%d too large for BC=%d (%zu B shared memory)", g->n, BC, smem);
  BOGP_CUDA(cudaFuncSetAttribute(gp_eval_kernel<BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const bool joint = a.mode == kModeJointFwd || a.mode == kModeJointBwd;
  const long long ctas = joint ? a.b : (a.b + BC - 1) / BC;
  int dev_id = 0, sms = 148;
  cudaGetDevice(&dev_id);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / smem));
  const int grid = (int)std::max<long long>(1, std::min<long long>(ctas, (long long)sms * per_sm));
  cudaEvent_t ev = bogp::timing_begin(s);
  gp_eval_kernel<BC><<<grid, kGpThreads, smem, s>>>(dev, a);
  bogp::timing_end(ev, s);
  bogp::count_launch();
  BOGP_CUDA(cudaGetLastError());
  return 0;
}

// Candidates per CTA: small batches spread over many SMs (latency), large batches amortise the L^-1 stream.
static int pick_bc(const bogp_gp_s* g, long long b, bool grad) {
  int dev_id = 0, sms = 148;
  cudaGetDevice(&dev_id);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
  int bc = 8;
  while (bc > 1 && (b + bc - 1) / bc < 2LL * sms) bc >>= 1;
  while (bc > 1 && gp_smem_bytes(g->n, bc, grad) > 227 * 1024) bc >>= 1;
  return bc;
}

// ---------------------------------------------------------------------------------------------- large batches
// Thousands of candidates against one model (the raw-sample pass of optimize_acqf: 4096 points, ei.py:81-92): the two
// triangular solves are GEMMs with the candidates as one dimension, V^T = K*^T L^-T and W^T = V^T L^-1, run as 64 x 64
// tiles on the FP64 tensor cores (ts_gemm_nt) instead of per-thread dot products that re-stream L^-1 from L2 for every
// 8 candidates.  Panels are [b padded to 64][n padded to 64] in global memory (L2-resident: 16 MB at b = 4096, n = 512).
//   1 cross-covariances  2 V^T (DMMA)  3 moments + acquisition (warp per candidate)  [4 W^T (DMMA)  5 gradient]
__global__ void __launch_bounds__(256)
gp_moments_kernel(GpDev gp, GpArgs a, const double* __restrict__ Kt, const double* __restrict__ Vt,
                  const double* __restrict__ alpha_p, int np, double* __restrict__ coef) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long c = (long long)blockIdx.x * 8 + warp;
  if (c >= a.b) return;
  const double* kr = Kt + (size_t)c * np;
  const double* vr = Vt + (size_t)c * np;
  double m = 0.0, ss = 0.0;
  for (int i = lane; i < np; i += 32) {
    m = fma(alpha_p[i], kr[i], m);
    const double v = vr[i];
    ss = fma(v, v, ss);
  }
  m = warp_sum_d(m) + gp.mean;
  ss = warp_sum_d(ss);
  double var = gp.outputscale - ss;
  if (lane != 0) return;
  if (a.mode == kModePosterior) {
    if (a.observation_noise) var += gp.noise;
    a.out0[c] = m;
    a.out1[c] = fmax(var, kMinVariance);
  } else {
    double val, dmu, dvar;
    acquisition(a.acq_kind, m, var, a.best_f, a.beta, val, dmu, dvar);
    a.out0[c] = val;
    if (coef) { coef[2 * c] = dmu; coef[2 * c + 1] = dvar; }
  }
}

__global__ void __launch_bounds__(256)
gp_grad_kernel(GpDev gp, GpArgs a, const double* __restrict__ Wt, const double* __restrict__ alpha_p, int np,
               const double* __restrict__ coef) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long c = (long long)blockIdx.x * 8 + warp;
  if (c >= a.b) return;
  const int d = gp.d, n = gp.n;
  double xc[kGpMaxD], g_acc[kGpMaxD];
#pragma unroll
  for (int dim = 0; dim < kGpMaxD; ++dim) { xc[dim] = dim < d ? a.Xs[c * d + dim] : 0.0; g_acc[dim] = 0.0; }
  const double dmu = coef[2 * c], dvar = coef[2 * c + 1];
  const double* wr = Wt + (size_t)c * np;
  for (int i = lane; i < n; i += 32) {
    double d2 = 0.0;
#pragma unroll
    for (int dim = 0; dim < kGpMaxD; ++dim)
      if (dim < d) {
        const double df = (xc[dim] - __ldg(gp.X + (size_t)i * d + dim)) * __ldg(gp.inv_ls + dim);
        d2 = fma(df, df, d2);
      }
    double k, g;
    kernel_eval(gp.kind, gp.outputscale, d2, k, g);
    const double cf = (dmu * alpha_p[i] - 2.0 * dvar * wr[i]) * g;
#pragma unroll
    for (int dim = 0; dim < kGpMaxD; ++dim)
      if (dim < d) g_acc[dim] = fma(cf, xc[dim] - __ldg(gp.X + (size_t)i * d + dim), g_acc[dim]);
  }
#pragma unroll
  for (int dim = 0; dim < kGpMaxD; ++dim)
    if (dim < d) {
      const double v = warp_sum_d(g_acc[dim]);
      if (lane == 0) a.grad[c * d + dim] = v * gp.inv_ls[dim] * gp.inv_ls[dim];
    }
}

__global__ void gp_transpose_pad_kernel(const double* __restrict__ Linv_p, int np, double* __restrict__ LinvT_p) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)np * np; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / np), j = (int)(e % np);
    LinvT_p[(size_t)j * np + i] = Linv_p[e];
  }
}

static bool use_dmma_path(const bogp_gp_s* g, const GpArgs& a) {
  if (const char* e = std::getenv("BOGP_GP_DMMA")) { if (e[0] == '0') return false; if (e[0] == '2') return a.q == 1; }
  return a.q == 1 && (a.mode == kModePosterior || a.mode == kModeAcq) && a.b >= 1024 && g->n >= 96 && g->d <= kGpMaxD;
}

static int launch_gp_dmma(bogp_gp_s* g, const GpArgs& a, bool grad, cudaStream_t s) {
  const int n = g->n, np = (n + 63) / 64 * 64;
  const int b = (int)a.b, bp = (b + 63) / 64 * 64;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_li = 0, o_lt = al((size_t)np * np * 8), o_al = o_lt + al((size_t)np * np * 8), o_kt = o_al + al((size_t)np * 8);
  const size_t o_vt = o_kt + al((size_t)bp * np * 8), o_cf = o_vt + al((size_t)bp * np * 8), total = o_cf + al((size_t)bp * 16);
  if (g->mm_bytes < total) {
    if (g->mm_ws) cudaFree(g->mm_ws);
    g->mm_ws = nullptr; g->mm_bytes = 0; g->mm_np = 0;
    BOGP_CUDA(cudaMalloc(&g->mm_ws, total + total / 4));
    g->mm_bytes = total + total / 4;
  }
  unsigned char* ws = static_cast<unsigned char*>(g->mm_ws);
  double *Li = (double*)(ws + o_li), *Lt = (double*)(ws + o_lt), *alp = (double*)(ws + o_al);
  double *Kt = (double*)(ws + o_kt), *Vt = (double*)(ws + o_vt), *coef = (double*)(ws + o_cf);
  if (g->mm_np != np) {          // once per model: padded L^-1 (identity on the padded diagonal), its transpose, alpha
    if (int rc = ts_pad(g, np, Li, alp, s)) return rc;
    gp_transpose_pad_kernel<<<296, 256, 0, s>>>(Li, np, Lt);
    count_launch();
    g->mm_np = np;
  }
  GpDev dev{g->n, g->d, g->kind, g->outputscale, g->mean, g->noise, g->X, g->inv_ls, g->alpha, g->Linv, g->LinvT};
  cudaEvent_t ev = bogp::timing_begin(s);
  if (int rc = ts_cross_kernel(g, a.Xs, b, bp, np, Kt, s)) return rc;
  if (int rc = ts_gemm_tri(Kt, Li, Vt, bp, np, 1, s)) return rc;              // V^T = K*^T L^-T: B = L^-1 rows, lower triangular
  gp_moments_kernel<<<(b + 7) / 8, 256, 0, s>>>(dev, a, Kt, Vt, alp, np, grad ? coef : nullptr);
  count_launch();
  if (grad) {
    if (int rc = ts_gemm_tri(Vt, Lt, Kt, bp, np, 2, s)) return rc;            // W^T = V^T L^-1: B = L^-T rows, upper triangular (into Kt)
    gp_grad_kernel<<<(b + 7) / 8, 256, 0, s>>>(dev, a, Kt, alp, np, coef);
    count_launch();
  }
  bogp::timing_end(ev, s);
  BOGP_CUDA(cudaGetLastError());
  return 0;
}

static int dispatch_gp(const bogp_gp_s* g, const GpArgs& a, int bc, bool grad, cudaStream_t s) {
  if (use_dmma_path(g, a)) return launch_gp_dmma(const_cast<bogp_gp_s*>(g), a, grad, s);
  BOGP_REQUIRE(g->d <= kGpMaxD, "posterior / acquisition kernels take d <= %d input dimensions (model has %d); only "
               "bogp_gp_thompson accepts more", kGpMaxD, g->d);
  switch (bc) {
    case 1: return launch_gp<1>(g, a, grad, s);
    case 2: return launch_gp<2>(g, a, grad, s);
    case 4: return launch_gp<4>(g, a, grad, s);
    default: return launch_gp<8>(g, a, grad, s);
  }
}

}  // namespace bogp

using bogp::GpArgs;

extern "C" int bogp_gp_create(bogp_gp_t* out, int n, int d, const double* X_host, const double* y_host, int kernel_kind,
                              const double* lengthscale_host, double outputscale, double mean, double noise) {
  BOGP_REQUIRE(out, "bogp_gp_create: out is NULL");
  *out = nullptr;
  BOGP_REQUIRE(n >= 1 && n <= bogp::kGpMaxN, "bogp_gp_create: n=%d outside [1,%d]", n, bogp::kGpMaxN);
  BOGP_REQUIRE(d >= 1 && d <= bogp::kGpMaxDim, "bogp_gp_create: d=%d outside [1,%d]", d, bogp::kGpMaxDim);
  BOGP_REQUIRE(X_host && y_host && lengthscale_host, "bogp_gp_create: NULL input");
  BOGP_REQUIRE(kernel_kind == BOGP_KERNEL_MATERN52 || kernel_kind == BOGP_KERNEL_RBF, "bogp_gp_create: unknown kernel %d", kernel_kind);
  BOGP_REQUIRE(outputscale > 0.0 && noise >= 0.0, "bogp_gp_create: outputscale must be > 0 and noise >= 0");
  for (int a = 0; a < d; ++a) BOGP_REQUIRE(lengthscale_host[a] > 0.0, "bogp_gp_create: lengthscale[%d] must be > 0", a);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    bogp::set_error("no CUDA device available; libbogp_sm100 has no CPU fallback");
    return BOGP_ERR_CUDA;
  }
  // K_XX + noise I  (float64, host: O(n^3) once per fit, the per-candidate work is on the GPU)
  std::vector<double> K((size_t)n * n), inv_ls(d);
  for (int a = 0; a < d; ++a) inv_ls[a] = 1.0 / lengthscale_host[a];
  const double sqrt5 = std::sqrt(5.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double d2 = 0.0;
      for (int a = 0; a < d; ++a) {
        const double df = X_host[(size_t)i * d + a] / lengthscale_host[a] - X_host[(size_t)j * d + a] / lengthscale_host[a];
        d2 += df * df;
      }
      double k;
      if (kernel_kind == BOGP_KERNEL_RBF) {
        k = outputscale * std::exp(-0.5 * d2);
      } else {
        const double r = std::sqrt(std::max(d2, 1e-30));
        k = outputscale * (1.0 + sqrt5 * r + (5.0 / 3.0) * d2) * std::exp(-sqrt5 * r);
      }
      K[(size_t)i * n + j] = K[(size_t)j * n + i] = k;
    }
  std::vector<double> L;
  double jitter = 0.0;
  bool ok = false;
  for (int attempt = 0; attempt < 4 && !ok; ++attempt) {
    jitter = attempt == 0 ? 0.0 : 1e-8 * std::pow(10.0, attempt - 1);
    L = K;
    for (int i = 0; i < n; ++i) L[(size_t)i * n + i] += noise + jitter;
    ok = bogp::cholesky_lower(L, n);
  }
  if (!ok) {
    bogp::set_error("bogp_gp_create: K + noise*I not positive definite after jitter 1e-6");
    return BOGP_ERR_NUMERIC;
  }
  // alpha = K^-1 (y - m) via two triangular solves
  std::vector<double> alpha(n);
  for (int i = 0; i < n; ++i) {
    double s = y_host[i] - mean;
    for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * alpha[k];
    alpha[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = alpha[i];
    for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * alpha[k];
    alpha[i] = s / L[(size_t)i * n + i];
  }
  bogp::invert_lower(L, n);
  std::vector<double> LT((size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) LT[(size_t)j * n + i] = L[(size_t)i * n + j];

  bogp_gp_s* g = new bogp_gp_s();
  g->n = n; g->d = d; g->kind = kernel_kind; g->outputscale = outputscale; g->mean = mean; g->noise = noise; g->jitter = jitter;
  auto up = [](double** dst, const double* src, size_t cnt) -> bool {
    if (cudaMalloc(dst, cnt * sizeof(double)) != cudaSuccess) return false;
    return cudaMemcpy(*dst, src, cnt * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
  };
  bool up_ok = up(&g->X, X_host, (size_t)n * d) && up(&g->inv_ls, inv_ls.data(), d) && up(&g->alpha, alpha.data(), n) &&
               up(&g->Linv, L.data(), (size_t)n * n) && up(&g->LinvT, LT.data(), (size_t)n * n);
  if (!up_ok) {
    bogp::set_error("bogp_gp_create: device allocation/upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    bogp_gp_destroy(g);
    return BOGP_ERR_CUDA;
  }
  *out = g;
  return 0;
}

extern "C" int bogp_gp_destroy(bogp_gp_t g) {
  if (!g) return 0;
  cudaFree(g->X); cudaFree(g->inv_ls); cudaFree(g->alpha); cudaFree(g->Linv); cudaFree(g->LinvT);
  if (g->io_dev) cudaFree(g->io_dev);
  if (g->io_pin) cudaFreeHost(g->io_pin);
  if (g->ts_ws) cudaFree(g->ts_ws);
  if (g->mm_ws) cudaFree(g->mm_ws);
  delete g;
  return 0;
}

extern "C" int bogp_gp_jitter(bogp_gp_t g, double* jitter_out) {
  BOGP_REQUIRE(g && jitter_out, "bogp_gp_jitter: NULL argument");
  *jitter_out = g->jitter;
  return 0;
}

extern "C" int bogp_gp_posterior(bogp_gp_t g, const double* Xs_dev, long long b, int observation_noise,
                                 double* mean_dev, double* var_dev, void* stream) {
  BOGP_REQUIRE(g && b >= 0 && (b == 0 || (Xs_dev && mean_dev && var_dev)), "bogp_gp_posterior: bad argument");
  if (b == 0) return 0;
  GpArgs a{};
  a.mode = bogp::kModePosterior; a.q = 1; a.observation_noise = observation_noise; a.b = b; a.Xs = Xs_dev;
  a.out0 = mean_dev; a.out1 = var_dev;
  return bogp::dispatch_gp(g, a, bogp::pick_bc(g, b, false), false, static_cast<cudaStream_t>(stream));
}

extern "C" int bogp_gp_acq(bogp_gp_t g, int acq_kind, double best_f, double beta, const double* Xs_dev, long long b,
                           double* out_dev, double* grad_dev, void* stream) {
  BOGP_REQUIRE(g && b >= 0 && (b == 0 || (Xs_dev && out_dev)), "bogp_gp_acq: bad argument");
  BOGP_REQUIRE(acq_kind >= BOGP_ACQ_EI && acq_kind <= BOGP_ACQ_UCB, "bogp_gp_acq: unknown acquisition %d", acq_kind);
  BOGP_REQUIRE(acq_kind != BOGP_ACQ_UCB || beta >= 0.0, "bogp_gp_acq: beta must be >= 0");
  if (b == 0) return 0;
  GpArgs a{};
  a.mode = bogp::kModeAcq; a.acq_kind = acq_kind; a.q = 1; a.best_f = best_f; a.beta = beta; a.b = b; a.Xs = Xs_dev;
  a.out0 = out_dev; a.grad = grad_dev;
  const bool grad = grad_dev != nullptr;
  return bogp::dispatch_gp(g, a, bogp::pick_bc(g, b, grad), grad, static_cast<cudaStream_t>(stream));
}

// ---- host-buffer variants: what optimize_acqf's L-BFGS-B calls hundreds of times per BO iteration with a handful of
// restarts -- one C call (page-locked staging, async copies, one synchronisation) instead of ~10 framework ops.
static int gp_io(bogp_gp_t g, size_t doubles) {
  if (g->io_doubles >= doubles) return 0;
  if (g->io_dev) cudaFree(g->io_dev);
  if (g->io_pin) cudaFreeHost(g->io_pin);
  g->io_dev = nullptr; g->io_pin = nullptr; g->io_doubles = 0;
  const size_t cap = doubles + doubles / 2 + 64;
  BOGP_CUDA(cudaMalloc(&g->io_dev, cap * sizeof(double)));
  BOGP_CUDA(cudaMallocHost(&g->io_pin, cap * sizeof(double)));
  g->io_doubles = cap;
  return 0;
}

extern "C" int bogp_gp_acq_host(bogp_gp_t g, int acq_kind, double best_f, double beta, const double* Xs_host, long long b,
                                double* out_host, double* grad_host, void* stream) {
  BOGP_REQUIRE(g && b >= 0 && (b == 0 || (Xs_host && out_host)), "bogp_gp_acq_host: bad argument");
  if (b == 0) return 0;
  const size_t nx = (size_t)b * g->d, total = nx + (size_t)b + (grad_host ? nx : 0);
  if (int rc = gp_io(g, total)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::memcpy(g->io_pin, Xs_host, nx * sizeof(double));
  BOGP_CUDA(cudaMemcpyAsync(g->io_dev, g->io_pin, nx * sizeof(double), cudaMemcpyHostToDevice, s));
  double* out_dev = g->io_dev + nx;
  double* grad_dev = grad_host ? out_dev + b : nullptr;
  if (int rc = bogp_gp_acq(g, acq_kind, best_f, beta, g->io_dev, b, out_dev, grad_dev, stream)) return rc;
  BOGP_CUDA(cudaMemcpyAsync(g->io_pin + nx, out_dev, (total - nx) * sizeof(double), cudaMemcpyDeviceToHost, s));
  BOGP_CUDA(cudaStreamSynchronize(s));
  std::memcpy(out_host, g->io_pin + nx, (size_t)b * sizeof(double));
  if (grad_host) std::memcpy(grad_host, g->io_pin + nx + b, nx * sizeof(double));
  return 0;
}

extern "C" int bogp_gp_qei_host(bogp_gp_t g, const double* Xs_host, long long b, int q, const double* base_dev, int S,
                                double best_f, double* out_host, double* grad_host, void* stream) {
  BOGP_REQUIRE(g && b >= 0 && q >= 1 && (b == 0 || (Xs_host && out_host && base_dev)), "bogp_gp_qei_host: bad argument");
  if (b == 0) return 0;
  const size_t nx = (size_t)b * q * g->d, total = nx + (size_t)b + (grad_host ? nx : 0);
  if (int rc = gp_io(g, total)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::memcpy(g->io_pin, Xs_host, nx * sizeof(double));
  BOGP_CUDA(cudaMemcpyAsync(g->io_dev, g->io_pin, nx * sizeof(double), cudaMemcpyHostToDevice, s));
  double* out_dev = g->io_dev + nx;
  double* grad_dev = grad_host ? out_dev + b : nullptr;
  if (int rc = bogp_gp_qei(g, g->io_dev, b, q, base_dev, S, best_f, out_dev, grad_dev, stream)) return rc;
  BOGP_CUDA(cudaMemcpyAsync(g->io_pin + nx, out_dev, (total - nx) * sizeof(double), cudaMemcpyDeviceToHost, s));
  BOGP_CUDA(cudaStreamSynchronize(s));
  std::memcpy(out_host, g->io_pin + nx, (size_t)b * sizeof(double));
  if (grad_host) std::memcpy(grad_host, g->io_pin + nx + b, nx * sizeof(double));
  return 0;
}

static int bc_for_q(int q) { return q <= 1 ? 1 : q <= 2 ? 2 : q <= 4 ? 4 : 8; }

extern "C" int bogp_gp_posterior_joint(bogp_gp_t g, const double* Xs_dev, long long b, int q, double* mean_dev,
                                       double* cov_dev, void* stream) {
  BOGP_REQUIRE(g && b >= 0 && (b == 0 || (Xs_dev && mean_dev && cov_dev)), "bogp_gp_posterior_joint: bad argument");
  BOGP_REQUIRE(q >= 1 && q <= bogp::kGpMaxQ, "bogp_gp_posterior_joint: q=%d outside [1,%d]", q, bogp::kGpMaxQ);
  if (b == 0) return 0;
  GpArgs a{};
  a.mode = bogp::kModeJointFwd; a.q = q; a.b = b; a.Xs = Xs_dev; a.out0 = mean_dev; a.out1 = cov_dev;
  return bogp::dispatch_gp(g, a, bc_for_q(q), false, static_cast<cudaStream_t>(stream));
}

extern "C" int bogp_gp_posterior_joint_backward(bogp_gp_t g, const double* Xs_dev, long long b, int q,
                                                const double* gmean_dev, const double* gcov_dev, double* grad_dev,
                                                void* stream) {
  BOGP_REQUIRE(g && b >= 0 && (b == 0 || (Xs_dev && gmean_dev && gcov_dev && grad_dev)), "bogp_gp_posterior_joint_backward: bad argument");
  BOGP_REQUIRE(q >= 1 && q <= bogp::kGpMaxQ, "bogp_gp_posterior_joint_backward: q=%d outside [1,%d]", q, bogp::kGpMaxQ);
  if (b == 0) return 0;
  GpArgs a{};
  a.mode = bogp::kModeJointBwd; a.q = q; a.b = b; a.Xs = Xs_dev; a.gmean = gmean_dev; a.gcov = gcov_dev; a.grad = grad_dev;
  return bogp::dispatch_gp(g, a, bc_for_q(q), true, static_cast<cudaStream_t>(stream));
}

extern "C" int bogp_gp_qei(bogp_gp_t g, const double* Xs_dev, long long b, int q, const double* base_dev, int S,
                           double best_f, double* out_dev, double* grad_dev, void* stream) {
  BOGP_REQUIRE(g && b >= 0 && (b == 0 || (Xs_dev && base_dev && out_dev)), "bogp_gp_qei: bad argument");
  BOGP_REQUIRE(q >= 1 && q <= bogp::kGpMaxQ, "bogp_gp_qei: q=%d outside [1,%d]", q, bogp::kGpMaxQ);
  BOGP_REQUIRE(S >= 1, "bogp_gp_qei: need S >= 1 base samples");
  if (b == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // scratch: mean[b,q], cov[b,q,q] (+ gmean, gcov), fail flag
  const size_t nm = (size_t)b * q, ncv = (size_t)b * q * q;
  double* scratch = nullptr;
  BOGP_CUDA(cudaMalloc(&scratch, (2 * nm + 2 * ncv + 2) * sizeof(double)));
  double *mean = scratch, *cov = mean + nm, *gmean = cov + ncv, *gcov = gmean + nm;
  int* fail = reinterpret_cast<int*>(gcov + ncv);
  int rc = 0;
  cudaError_t e = cudaMemsetAsync(fail, 0, sizeof(int), s);
  if (e != cudaSuccess) { bogp::set_error("bogp_gp_qei: %s", cudaGetErrorString(e)); rc = BOGP_ERR_CUDA; }
  if (!rc) rc = bogp_gp_posterior_joint(g, Xs_dev, b, q, mean, cov, stream);
  if (!rc) {
    bogp::qei_kernel<<<(unsigned)((b + 3) / 4), 128, 0, s>>>(mean, cov, b, q, base_dev, S, best_f, out_dev,
                                                            grad_dev ? gmean : nullptr, grad_dev ? gcov : nullptr, fail);
    bogp::count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { bogp::set_error("bogp_gp_qei: %s", cudaGetErrorString(e)); rc = BOGP_ERR_CUDA; }
  }
  if (!rc && grad_dev) rc = bogp_gp_posterior_joint_backward(g, Xs_dev, b, q, gmean, gcov, grad_dev, stream);
  int failed = 0;
  if (!rc) {
    e = cudaMemcpyAsync(&failed, fail, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { bogp::set_error("bogp_gp_qei: %s", cudaGetErrorString(e)); rc = BOGP_ERR_CUDA; }
  } else {
    cudaStreamSynchronize(s);
  }
  cudaFree(scratch);
  if (!rc && failed) {
    bogp::set_error("bogp_gp_qei: joint posterior covariance not positive definite after jitter 1e-6");
    return BOGP_ERR_NUMERIC;
  }
  return rc;
}
