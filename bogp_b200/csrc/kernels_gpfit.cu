// GP hyper-parameter fit on the GPU: the reference's loop
//     for _ in range(100): optimizer.zero_grad(); loss = -mll(model(X), train_Y); loss.backward(); optimizer.step()
// (ref/projects/swellex96_inv/data/bo/ei.py:69-78, AdamW lr 0.1 over the raw GPyTorch parameters) as ONE kernel
// launch: every step -- kernel matrix, blocked Cholesky, explicit inverse, marginal log-likelihood, analytic
// gradient, AdamW update -- runs inside a persistent CTA with the n x n matrix resident in shared memory (n <= ~150
// at d = 7: 227 KB) or in an L2-resident scratch for larger n; no host round trip, no library call between steps.
// One CTA per independent fit: a batch of B starting points is B concurrent CTAs.
//
//   -mll / n,  mll = log N(y | m, K + s2 I) + sum log-priors                      (BoTorch ExactMarginalLogLikelihood)
//   d mll / d theta = 1/2 tr( (alpha alpha^T - K^-1) dK/dtheta ),   alpha = K^-1 (y - m)
//
// float64 throughout (the reference fits in float64, ei.py:24).  Linear algebra, all on 32 x 32 blocks:
//   * diagonal blocks: Cholesky and triangular inverse fused level by level across the CTA (64 block barriers per
//     32 x 32 block; a single-warp register/shuffle version was 5x slower: 4000 dependent instructions per block);
//   * panel solve L_ik = A_ik D^-T, trailing update A_ij -= L_ik L_jk^T, blocked in-place inverse X = L^-1,
//     K^-1 = X^T X into the (free) upper triangle: warp-per-row loops over shared memory, odd leading dimension.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "bogp_internal.h"
#include "gpfit_device.cuh"

namespace bogp {

// SMEM = true: the matrix and X live in shared memory (the compiler sees shared-space pointers: LDS/STS with
// 32-bit addressing instead of generic loads); SMEM = false: matrix in the global scratch.
// DL: compile-time bound of the loops over input dimensions (4, 8 or 16 >= d): the per-pair loops are predicated, and
// predicated-off iterations still take issue slots.
template <bool SMEM, int DL>
__global__ void __launch_bounds__(kFitThreads, 1) gp_fit_kernel(const FitArgs p) {
  extern __shared__ double fsm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = p.n, d = p.d, ld = p.ld, np = d + 3;
  const int T = (n + 31) / 32;
  double* sp = fsm;
  double* A = SMEM ? sp : p.work + (size_t)blockIdx.x * n * ld;
  if (SMEM) sp += n * ld;
  double* Xs = sp; if (SMEM || p.x_in_smem) sp += n * d;
  const double* Xp = (SMEM || p.x_in_smem) ? Xs : p.X;
  double* rv = sp; sp += n;            // y - mean
  double* alpha = sp; sp += n;
  double* zv = sp; sp += n;
  double* kdiag = sp; sp += n;         // diag(K^-1)
  double* T1 = sp; sp += 32 * 33;      // inverse of the current diagonal block / block product
  double* Db = sp; sp += 32 * 33;      // factor of the current diagonal block
  double* Lc = sp; sp += 64;           // current factor column, double-buffered by level parity
  double* red = sp; sp += kFitWarps * (kFitMaxD + 4);
  double* prm = sp; sp += kFitMaxD + 8;   // ls[d] | os, noise, mean, jitter at kFitMaxD..
  double* raw = sp; sp += kFitNP;
  double* am = sp; sp += kFitNP;       // Adam first moment
  double* av = sp; sp += kFitNP;       // Adam second moment
  double* gsum = sp; sp += kFitMaxD + 4;
  int* flag = reinterpret_cast<int*>(sp);

  if (SMEM || p.x_in_smem)
    for (int i = tid; i < n * d; i += kFitThreads) Xs[i] = p.X[i];
  if (tid < np) { raw[tid] = p.raw[(size_t)blockIdx.x * np + tid]; am[tid] = 0.0; av[tid] = 0.0; }
  if (tid == 0) *flag = 0;
  __syncthreads();

  for (int step = 1; step <= p.steps; ++step) {
    // ---------------------------------------------------------------- constrained parameters
    if (tid < d) prm[tid] = p.ls_hi > p.ls_lo ? p.ls_lo + (p.ls_hi - p.ls_lo) * sigmoid_d(raw[tid]) : softplus_d(raw[tid]);
    if (tid == 32) prm[kFitMaxD + 0] = p.os_hi > p.os_lo ? p.os_lo + (p.os_hi - p.os_lo) * sigmoid_d(raw[d]) : softplus_d(raw[d]);
    if (tid == 33) prm[kFitMaxD + 1] = p.noise_lo + (p.noise_hi - p.noise_lo) * sigmoid_d(raw[d + 1]);
    if (tid == 34) prm[kFitMaxD + 2] = raw[d + 2];
    __syncthreads();
    const double os = prm[kFitMaxD], noise = prm[kFitMaxD + 1], mean = prm[kFitMaxD + 2];
    double inv_ls[DL];
#pragma unroll
    for (int k = 0; k < DL; ++k) inv_ls[k] = k < d ? 1.0 / prm[k] : 0.0;
    for (int i = tid; i < n; i += kFitThreads) rv[i] = p.y[i] - mean;

    // ---------------------------------------------------------------- K (lower) and its Cholesky, jitter retries
    bool pd = false;
    for (int attempt = 0; attempt < 4 && !pd; ++attempt) {
      const double jitter = attempt == 0 ? 0.0 : (attempt == 1 ? 1e-8 : (attempt == 2 ? 1e-7 : 1e-6));
      for (int i = warp; i < n; i += kFitWarps)
        for (int j = lane; j <= i; j += 32) {
          double d2 = 0.0;
#pragma unroll
          for (int k = 0; k < DL; ++k)
            if (k < d) { const double s = (Xp[i * d + k] - Xp[j * d + k]) * inv_ls[k]; d2 = fma(s, s, d2); }
          double kc, g;
          corr_and_grad(p.kind, d2, kc, g);
          A[i * ld + j] = os * kc + (i == j ? noise + jitter : 0.0);
        }
      __syncthreads();
      pd = true;
      for (int kb = 0; kb < T; ++kb) {
        const int r0 = kb * 32, bs = min(32, n - r0), r1 = r0 + bs;
        if (!block_chol_inv(A, ld, r0, bs, Db, T1, Lc, flag)) { pd = false; break; }
        if (r1 < n) {
          // panel: L_i,blk = A_i,blk D^-T, two rows per warp (rows are private to a warp: read everything, then
          // write); D^-1 in T1 has a zero upper triangle and bs = 32 here (a trailing panel exists)
          for (int i = r1 + 2 * warp; i < n; i += 2 * kFitWarps) {
            const double* a0 = A + i * ld + r0;
            const double* a1 = (i + 1 < n) ? a0 + ld : a0;
            const double* t = T1 + lane * 33;
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
            for (int c = 0; c < 32; ++c) { const double tv = t[c]; acc0 = fma(a0[c], tv, acc0); acc1 = fma(a1[c], tv, acc1); }
            __syncwarp();
            A[i * ld + r0 + lane] = acc0;
            if (i + 1 < n) A[(i + 1) * ld + r0 + lane] = acc1;
          }
          __syncthreads();
          // trailing update: A_ij -= sum_c L_ic L_jc,  r1 <= j <= i: rows (i, i+1) per warp, lanes over j
          for (int i = r1 + 2 * warp; i < n; i += 2 * kFitWarps) {
            const bool two = i + 1 < n;
            const double* li0 = A + i * ld + r0;
            const double* li1 = two ? li0 + ld : li0;
            const int jend = two ? i + 1 : i;
            for (int j = r1 + lane; j <= jend; j += 32) {
              const double* lj = A + j * ld + r0;
              double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
              for (int c = 0; c < 32; ++c) { const double v = lj[c]; acc0 = fma(li0[c], v, acc0); acc1 = fma(li1[c], v, acc1); }
              if (j <= i) A[i * ld + j] -= acc0;
              if (two) A[(i + 1) * ld + j] -= acc1;
            }
          }
          __syncthreads();
        }
      }
      if (!pd) {
        __syncthreads();
        if (tid == 0) *flag = 0;
        __syncthreads();
      }
    }
    if (!pd) {
      if (tid == 0) p.status[blockIdx.x] = 1;
      return;
    }

    // ---------------------------------------------------------------- X = L^-1 in place (lower)
    // (the diagonal blocks already hold their inverses: block_chol_inv)
    for (int k = 0; k + 1 < T; ++k) {
      const int c0 = k * 32;                     // column block (always full: k < T - 1)
      for (int ib = k + 1; ib < T; ++ib) {
        const int i0 = ib * 32, bsi = min(32, n - i0);
        // S = sum_{j = k}^{ib - 1} L_(ib, j) X_(j, k): thread (warp a, lane b) owns S[a][b] and S[a + 16][b]
        {
          const int a = warp;
          const double* l0 = A + (i0 + min(a, bsi - 1)) * ld;
          const double* l1 = A + (i0 + min(a + 16, bsi - 1)) * ld;
          double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
          for (int c = 0; c < 32; ++c) {                     // j = k: X_kk is lower triangular
            const double xv = c >= lane ? A[(c0 + c) * ld + c0 + lane] : 0.0;
            acc0 = fma(l0[c0 + c], xv, acc0); acc1 = fma(l1[c0 + c], xv, acc1);
          }
#pragma unroll 4
          for (int jj = c0 + 32; jj < i0; ++jj) {
            const double xv = A[jj * ld + c0 + lane];
            acc0 = fma(l0[jj], xv, acc0); acc1 = fma(l1[jj], xv, acc1);
          }
          T1[a * 33 + lane] = acc0;
          T1[(a + 16) * 33 + lane] = acc1;
        }
        __syncthreads();
        // X_(ib, k) = - X_(ib, ib) S   (the upper triangle of the diagonal block of A is not zero: select)
        {
          const int a = warp;
          const double* x0 = A + (i0 + min(a, bsi - 1)) * ld + i0;
          const double* x1 = A + (i0 + min(a + 16, bsi - 1)) * ld + i0;
          double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
          for (int c = 0; c < 32; ++c) {
            const double sv = T1[c * 33 + lane];
            acc0 = fma(c <= a ? x0[c] : 0.0, sv, acc0);
            acc1 = fma((c <= a + 16 && c < bsi) ? x1[c] : 0.0, sv, acc1);
          }
          __syncwarp();
          if (a < bsi) A[(i0 + a) * ld + c0 + lane] = -acc0;
          if (a + 16 < bsi) A[(i0 + a + 16) * ld + c0 + lane] = -acc1;
        }
        __syncthreads();
      }
    }

    // ---------------------------------------------------------------- alpha = X^T X r, log det, quadratic form
    for (int k = tid; k < n; k += kFitThreads) {
      double acc = 0.0;
      for (int j = 0; j <= k; ++j) acc = fma(A[k * ld + j], rv[j], acc);
      zv[k] = acc;
    }
    __syncthreads();
    double part_quad = 0.0, part_ld = 0.0, part_as = 0.0;
    for (int j = tid; j < n; j += kFitThreads) {
      double acc = 0.0;
      for (int k = j; k < n; ++k) acc = fma(A[k * ld + j], zv[k], acc);
      alpha[j] = acc;
      part_quad = fma(rv[j], acc, part_quad);
      part_ld -= log(A[j * ld + j]);          // log L_jj = -log X_jj
      part_as += acc;
    }
    // ---------------------------------------------------------------- K^-1 = X^T X: strict upper triangle + diagonal
    for (int j = 2 * warp; j < n; j += 2 * kFitWarps) {      // columns (j, j + 1) per warp, lanes over i
      const bool two = j + 1 < n;
      const int iend = two ? j + 1 : j;
      for (int i = lane; i <= iend; i += 32) {
        double acc0 = i <= j ? A[j * ld + i] * A[j * ld + j] : 0.0, acc1 = 0.0;      // k = j term (column j only)
#pragma unroll 4
        for (int k = j + 1; k < n; ++k) {
          const double xi = A[k * ld + i];
          acc0 = fma(xi, A[k * ld + j], acc0);
          acc1 = fma(xi, A[k * ld + j + 1], acc1);
        }
        if (i < j) A[i * ld + j] = acc0; else if (i == j) kdiag[j] = acc0;
        if (two) { if (i <= j) A[i * ld + j + 1] = acc1; else kdiag[j + 1] = acc1; }
      }
    }
    __syncthreads();

    // ---------------------------------------------------------------- gradient traces over the lower triangle
    double g_ls[DL], g_os = 0.0, g_noise = 0.0;
#pragma unroll
    for (int k = 0; k < DL; ++k) g_ls[k] = 0.0;
    for (int i = warp; i < n; i += kFitWarps) {
      const double ai = alpha[i];
      for (int j = lane; j <= i; j += 32) {
        double s2[DL], d2 = 0.0;
#pragma unroll
        for (int k = 0; k < DL; ++k) {
          s2[k] = 0.0;
          if (k < d) { const double s = (Xp[i * d + k] - Xp[j * d + k]) * inv_ls[k]; s2[k] = s * s; d2 += s2[k]; }
        }
        double kc, g;
        corr_and_grad(p.kind, d2, kc, g);
        const double kinv = i == j ? kdiag[i] : A[j * ld + i];
        const double W = (i == j ? 0.5 : 1.0) * (ai * alpha[j] - kinv);      // 1/2 * (1 or 2) * W_ij
        g_os = fma(W, kc, g_os);
        if (i == j) g_noise += W;
        const double wg = W * os * g;
#pragma unroll
        for (int k = 0; k < DL; ++k)
          if (k < d) g_ls[k] = fma(wg, s2[k] * inv_ls[k], g_ls[k]);
      }
    }
    // block reduction of d + 5 sums (fixed order: deterministic)
    // block reduction of d + 4 sums (fixed order: deterministic); slot q < kFitMaxD: length-scale q, then os, noise, ...
    {
      double vals[DL + 4];
#pragma unroll
      for (int k = 0; k < DL; ++k) vals[k] = g_ls[k];
      vals[DL] = g_os; vals[DL + 1] = g_noise; vals[DL + 2] = part_quad; vals[DL + 3] = part_ld;
#pragma unroll
      for (int q = 0; q < DL + 4; ++q) {
        double v = vals[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0) red[warp * (kFitMaxD + 4) + (q < DL ? q : kFitMaxD + q - DL)] = v;
      }
    }
    {
      double v = part_as;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
      part_as = v;
    }
    __syncthreads();
    if (tid < kFitMaxD + 4 && (tid < DL || tid >= kFitMaxD)) {
      double v = 0.0;
      for (int w = 0; w < kFitWarps; ++w) v += red[w * (kFitMaxD + 4) + tid];
      gsum[tid] = v;
    }
    __syncthreads();
    if (lane == 0) red[warp] = part_as;
    __syncthreads();

    // ---------------------------------------------------------------- loss, chain rule, AdamW (one thread)
    if (tid == 0) {
      double sum_alpha = 0.0;
      for (int w = 0; w < kFitWarps; ++w) sum_alpha += red[w];
      const double nn = (double)n;
      double ll = -0.5 * gsum[kFitMaxD + 2] - gsum[kFitMaxD + 3] - 0.5 * nn * 1.8378770664093454836;   // log(2 pi)
      double grad[kFitNP];
      for (int k = 0; k < d; ++k) {
        const double ls = prm[k];
        double gl = gsum[k];
        if (p.has_ls_prior) {
          ll += p.ls_conc * log(p.ls_rate) - lgamma(p.ls_conc) + (p.ls_conc - 1.0) * log(ls) - p.ls_rate * ls;
          gl += (p.ls_conc - 1.0) / ls - p.ls_rate;
        }
        const double sk = sigmoid_d(raw[k]);
        grad[k] = -gl * (p.ls_hi > p.ls_lo ? (p.ls_hi - p.ls_lo) * sk * (1.0 - sk) : sk) / nn;
      }
      {
        double go = gsum[kFitMaxD];
        if (p.has_os_prior) {
          ll += p.os_conc * log(p.os_rate) - lgamma(p.os_conc) + (p.os_conc - 1.0) * log(os) - p.os_rate * os;
          go += (p.os_conc - 1.0) / os - p.os_rate;
        }
        const double so = sigmoid_d(raw[d]);
        grad[d] = -go * (p.os_hi > p.os_lo ? (p.os_hi - p.os_lo) * so * (1.0 - so) : so) / nn;
        const double sg = sigmoid_d(raw[d + 1]);
        grad[d + 1] = -gsum[kFitMaxD + 1] * (p.noise_hi - p.noise_lo) * sg * (1.0 - sg) / nn;
        grad[d + 2] = -sum_alpha / nn;
      }
      if (p.loss) p.loss[(size_t)blockIdx.x * p.steps + step - 1] = -ll / nn;
      if (p.grad_out)
        for (int k = 0; k < np; ++k) p.grad_out[(size_t)blockIdx.x * np + k] = grad[k];
      const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
      const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
      for (int k = 0; k < np && !p.grad_out; ++k) {
        double x = raw[k] * (1.0 - p.lr * p.wd);
        am[k] = b1 * am[k] + (1.0 - b1) * grad[k];
        av[k] = b2 * av[k] + (1.0 - b2) * grad[k] * grad[k];
        const double denom = sqrt(av[k]) / sqrt(bc2) + eps;
        raw[k] = x - (p.lr / bc1) * am[k] / denom;
      }
    }
    __syncthreads();
  }
  if (tid < np) p.raw[(size_t)blockIdx.x * np + tid] = raw[tid];
  if (tid == 0) p.status[blockIdx.x] = 0;
}

}  // namespace bogp

extern "C" int bogp_gp_fit_adamw(int n, int d, const double* X_host, const double* y_host, int kernel_kind, int steps,
                                 double lr, double weight_decay, double noise_lo, double noise_hi,
                                 const double* lengthscale_prior, const double* outputscale_prior, int B,
                                 double* raw_inout_host, double* loss_host, void* stream) {
  return bogp_gp_fit_adamw_bounded(n, d, X_host, y_host, kernel_kind, steps, lr, weight_decay, noise_lo, noise_hi,
                                   lengthscale_prior, outputscale_prior, nullptr, nullptr, B, raw_inout_host, loss_host,
                                   stream);
}

// grad_host != null: one evaluation -- loss[B] and d(loss)/d(raw)[B][d + 3] at raw_inout_host, which is left unchanged
static int fit_impl(int n, int d, const double* X_host, const double* y_host, int kernel_kind,
                    int steps, double lr, double weight_decay, double noise_lo, double noise_hi,
                    const double* lengthscale_prior, const double* outputscale_prior,
                    const double* lengthscale_bounds, const double* outputscale_bounds, int B,
                    double* raw_inout_host, double* loss_host, double* grad_host, void* stream) {
  using namespace bogp;
  BOGP_REQUIRE(!lengthscale_bounds || (lengthscale_bounds[0] > 0.0 && lengthscale_bounds[1] > lengthscale_bounds[0]),
               "bogp_gp_fit_adamw: bad lengthscale bounds");
  BOGP_REQUIRE(!outputscale_bounds || (outputscale_bounds[0] > 0.0 && outputscale_bounds[1] > outputscale_bounds[0]),
               "bogp_gp_fit_adamw: bad outputscale bounds");
  BOGP_REQUIRE(n >= 1 && n <= kFitMaxN, "bogp_gp_fit_adamw: n=%d outside [1,%d]", n, kFitMaxN);
  BOGP_REQUIRE(d >= 1 && d <= kFitMaxD, "bogp_gp_fit_adamw: d=%d outside [1,%d]", d, kFitMaxD);
  BOGP_REQUIRE(X_host && y_host && raw_inout_host, "bogp_gp_fit_adamw: NULL input");
  BOGP_REQUIRE(kernel_kind == BOGP_KERNEL_MATERN52 || kernel_kind == BOGP_KERNEL_RBF, "bogp_gp_fit_adamw: unknown kernel %d", kernel_kind);
  BOGP_REQUIRE(steps >= 1 && steps <= 100000 && B >= 1 && B <= 4096, "bogp_gp_fit_adamw: bad steps=%d or B=%d", steps, B);
  BOGP_REQUIRE(lr > 0.0 && weight_decay >= 0.0 && noise_lo > 0.0 && noise_hi > noise_lo, "bogp_gp_fit_adamw: bad lr / weight decay / noise bounds");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available; libbogp_sm100 has no CPU fallback");
    return BOGP_ERR_CUDA;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FitArgs p{};
  p.n = n; p.d = d; p.kind = kernel_kind; p.steps = steps; p.ld = n | 1;
  p.lr = lr; p.wd = weight_decay; p.noise_lo = noise_lo; p.noise_hi = noise_hi;
  if (lengthscale_bounds) { p.ls_lo = lengthscale_bounds[0]; p.ls_hi = lengthscale_bounds[1]; }
  if (outputscale_bounds) { p.os_lo = outputscale_bounds[0]; p.os_hi = outputscale_bounds[1]; }
  if (lengthscale_prior) { p.has_ls_prior = 1; p.ls_conc = lengthscale_prior[0]; p.ls_rate = lengthscale_prior[1]; }
  if (outputscale_prior) { p.has_os_prior = 1; p.os_conc = outputscale_prior[0]; p.os_rate = outputscale_prior[1]; }
  const size_t fixed = ((size_t)4 * n + 2 * 32 * 33 + 64 + kFitWarps * (kFitMaxD + 4) + (kFitMaxD + 8) + 3 * kFitNP + (kFitMaxD + 4) + 2) * sizeof(double);
  const size_t a_bytes = (size_t)n * p.ld * sizeof(double), x_bytes = (size_t)n * d * sizeof(double);
  const size_t limit = 227 * 1024;
  p.x_in_smem = fixed + x_bytes <= limit ? 1 : 0;
  p.a_in_smem = fixed + (p.x_in_smem ? x_bytes : 0) + a_bytes <= limit ? 1 : 0;
  const size_t smem = fixed + (p.x_in_smem ? x_bytes : 0) + (p.a_in_smem ? a_bytes : 0);
  const int np = d + 3;
  double *dX = nullptr, *dy = nullptr, *draw = nullptr, *dloss = nullptr, *dwork = nullptr, *dgrad = nullptr;
  int* dstatus = nullptr;
  auto cleanup = [&]() { cudaFree(dX); cudaFree(dy); cudaFree(draw); cudaFree(dloss); cudaFree(dwork); cudaFree(dstatus); cudaFree(dgrad); };
#define FIT_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); cleanup(); return BOGP_ERR_CUDA; } } while (0)
  FIT_CUDA(cudaMalloc(&dX, x_bytes));
  FIT_CUDA(cudaMalloc(&dy, n * sizeof(double)));
  FIT_CUDA(cudaMalloc(&draw, (size_t)B * np * sizeof(double)));
  FIT_CUDA(cudaMalloc(&dstatus, B * sizeof(int)));
  if (loss_host) FIT_CUDA(cudaMalloc(&dloss, (size_t)B * steps * sizeof(double)));
  if (grad_host) FIT_CUDA(cudaMalloc(&dgrad, (size_t)B * np * sizeof(double)));
  if (!p.a_in_smem) FIT_CUDA(cudaMalloc(&dwork, (size_t)B * a_bytes));      // (one-CTA kernel only; the cooperative path brings its own)
  FIT_CUDA(cudaMemcpyAsync(dX, X_host, x_bytes, cudaMemcpyHostToDevice, s));
  FIT_CUDA(cudaMemcpyAsync(dy, y_host, n * sizeof(double), cudaMemcpyHostToDevice, s));
  FIT_CUDA(cudaMemcpyAsync(draw, raw_inout_host, (size_t)B * np * sizeof(double), cudaMemcpyHostToDevice, s));
  p.X = dX; p.y = dy; p.raw = draw; p.loss = dloss; p.work = dwork; p.status = dstatus; p.grad_out = dgrad;
  const bool in_smem = p.a_in_smem && p.x_in_smem;
  const int dl = d <= 4 ? 4 : (d <= 8 ? 8 : 16);
  // models whose matrix does not fit one SM's shared memory: one fit over a cooperative grid (kernels_gpfit_coop.cu),
  // starting points one after the other.  BOGP_FIT_COOP=0 keeps them on the one-CTA kernel, =1 forces the grid.
  bool coop = !p.a_in_smem;
  if (const char* ce = std::getenv("BOGP_FIT_COOP")) coop = ce[0] == '1' ? true : (ce[0] == '0' ? false : coop);
  if (coop) {
    int dev_id = 0, sms = 148;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
    for (int b = 0; b < B; ++b) {
      FitArgs pb = p;
      pb.raw = draw + (size_t)b * np;
      pb.loss = dloss ? dloss + (size_t)b * steps : nullptr;
      pb.status = dstatus + b;
      pb.grad_out = dgrad ? dgrad + (size_t)b * np : nullptr;
      FIT_CUDA(launch_gp_fit_coop(pb, sms, s));
    }
  } else {
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaEvent_t ev = timing_begin(s);
    kern<<<B, kFitThreads, smem, s>>>(p);
    timing_end(ev, s);
    return cudaGetLastError();
  };
  if (in_smem) FIT_CUDA(dl == 4 ? launch(gp_fit_kernel<true, 4>) : dl == 8 ? launch(gp_fit_kernel<true, 8>) : launch(gp_fit_kernel<true, 16>));
  else FIT_CUDA(dl == 4 ? launch(gp_fit_kernel<false, 4>) : dl == 8 ? launch(gp_fit_kernel<false, 8>) : launch(gp_fit_kernel<false, 16>));
  count_launch();
  }
  FIT_CUDA(cudaGetLastError());
  std::vector<int> status(B);
  if (grad_host) FIT_CUDA(cudaMemcpyAsync(grad_host, dgrad, (size_t)B * np * sizeof(double), cudaMemcpyDeviceToHost, s));
  else FIT_CUDA(cudaMemcpyAsync(raw_inout_host, draw, (size_t)B * np * sizeof(double), cudaMemcpyDeviceToHost, s));
  FIT_CUDA(cudaMemcpyAsync(status.data(), dstatus, B * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (loss_host) FIT_CUDA(cudaMemcpyAsync(loss_host, dloss, (size_t)B * steps * sizeof(double), cudaMemcpyDeviceToHost, s));
  FIT_CUDA(cudaStreamSynchronize(s));
#undef FIT_CUDA
  cleanup();
  for (int b = 0; b < B; ++b)
    if (status[b] != 0) {
      set_error("bogp_gp_fit_adamw: K + noise*I not positive definite after jitter 1e-6 (fit %d)", b);
      return BOGP_ERR_NUMERIC;
    }
  return 0;
}

extern "C" int bogp_gp_fit_adamw_bounded(int n, int d, const double* X_host, const double* y_host, int kernel_kind,
                                         int steps, double lr, double weight_decay, double noise_lo, double noise_hi,
                                         const double* lengthscale_prior, const double* outputscale_prior,
                                         const double* lengthscale_bounds, const double* outputscale_bounds, int B,
                                         double* raw_inout_host, double* loss_host, void* stream) {
  return fit_impl(n, d, X_host, y_host, kernel_kind, steps, lr, weight_decay, noise_lo, noise_hi, lengthscale_prior,
                  outputscale_prior, lengthscale_bounds, outputscale_bounds, B, raw_inout_host, loss_host, nullptr, stream);
}

extern "C" int bogp_gp_mll_grad(int n, int d, const double* X_host, const double* y_host, int kernel_kind, double noise_lo,
                                double noise_hi, const double* lengthscale_prior, const double* outputscale_prior,
                                const double* lengthscale_bounds, const double* outputscale_bounds, int B,
                                const double* raw_host, double* loss_host, double* grad_host, void* stream) {
  BOGP_REQUIRE(loss_host && grad_host, "bogp_gp_mll_grad: NULL output");
  return fit_impl(n, d, X_host, y_host, kernel_kind, 1, 1.0, 0.0, noise_lo, noise_hi, lengthscale_prior, outputscale_prior,
                  lengthscale_bounds, outputscale_bounds, B, const_cast<double*>(raw_host), loss_host, grad_host, stream);
}
