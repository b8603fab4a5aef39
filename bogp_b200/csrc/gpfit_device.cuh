// Device code shared by the GP hyper-parameter fit kernels (kernels_gpfit.cu: one CTA per fit; kernels_gpfit_coop.cu:
// one fit spread over a cooperative grid).
#pragma once
#include "bogp_internal.h"

namespace bogp {

constexpr int kFitThreads = 512;
constexpr int kFitWarps = kFitThreads / 32;
constexpr int kFitMaxD = 16;
constexpr int kFitMaxN = 1024;
constexpr int kFitNP = kFitMaxD + 3;        // raw parameters: lengthscale[d], outputscale, noise, mean

struct FitArgs {
  int n, d, kind, steps, ld, a_in_smem, x_in_smem;
  double lr, wd, noise_lo, noise_hi;
  double ls_lo, ls_hi, os_lo, os_hi;      // Interval constraints (hi > lo) instead of softplus (baxus.py:310-319)
  int has_ls_prior, has_os_prior;
  double ls_conc, ls_rate, os_conc, os_rate;
  const double* X;       // [n, d]
  const double* y;       // [n]
  double* raw;           // [B][d + 3] in/out
  double* loss;          // [B][steps] or null
  double* grad_out;      // null, or [B][d + 3]: write d(loss)/d(raw) of every step here and leave raw alone (bogp_gp_mll_grad)
  double* work;          // [B][n * ld] when the matrix does not fit in shared memory
  int* status;           // [B]: 0 ok, 1 not positive definite after jitter 1e-6
};

constexpr unsigned kFull = 0xffffffffu;

// Cholesky AND inverse of the bs x bs diagonal block at (r0, r0) by the whole CTA, fused level by level (two block
// barriers per level): on exit the block of A holds D^-1 (lower) -- the factor D itself is never needed again: the
// panel solve, the blocked inverse and log det all use D^-1 -- and Xb holds D^-1 with a zero upper triangle.
// Thread (warp w, lane c) owns elements (w, c) and (w + 16, c) of the 32 x 32 tiles Db (factor) and Xb (inverse);
// rows >= bs are identity.  Level j:  A(j) column j of the factor (into Lc[j & 1]) | C(j-1) inverse rows > j-1 -= ..
// -- barrier --  B(j) trailing update, store column j, finalise inverse row j  -- barrier --.
// Returns false (uniform) on a non-positive pivot.
__device__ __forceinline__ bool block_chol_inv(double* A, int ld, int r0, int bs, double* Db, double* Xb, double* Lc, int* flag,
                                               bool write_back = true) {
  const int w = threadIdx.x >> 5, c = threadIdx.x & 31;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = w + 16 * h;
    Db[i * 33 + c] = (i < bs && c <= i) ? A[(r0 + i) * ld + r0 + c] : (i == c ? 1.0 : 0.0);
    Xb[i * 33 + c] = i == c ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < 32; ++j) {
    double* Lj = Lc + 32 * (j & 1);
    if (w == 0) {                                            // A(j): one warp, lane = row
      double ajj = Db[j * 33 + j];
      if (!(ajj > 0.0)) { if (c == 0) *flag = 1; ajj = 1.0; }
      // 1 / sqrt(pivot) from the float approximation and two Newton steps instead of DSQRT + DDIV (the two longest
      // dependent instruction sequences of the level: this chain is the critical path of the whole factorisation)
      double rs;
      if (ajj > 1e-30 && ajj < 1e30) {
        rs = (double)rsqrtf((float)ajj);
        rs = rs * fma(-0.5 * ajj, rs * rs, 1.5);
        rs = rs * fma(-0.5 * ajj, rs * rs, 1.5);
        rs = rs * fma(-0.5 * ajj, rs * rs, 1.5);      // (third step: the seed is only good to ~2^-22)
      } else {
        rs = 1.0 / sqrt(ajj);
      }
      if (c >= j) Lj[c] = c == j ? ajj * rs : Db[c * 33 + j] * rs;
      if (c == 0) Db[j * 33 + 32] = rs;                      // padding column of the factor tile: 1 / L_jj for the inverse's row j
    }
    if (j > 0 && c <= j - 1) {                               // C(j-1)
      const double* Lp = Lc + 32 * ((j - 1) & 1);
      const double xk = Xb[(j - 1) * 33 + c];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = w + 16 * h;
        if (i > j - 1) Xb[i * 33 + c] = fma(-Lp[i], xk, Xb[i * 33 + c]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {                            // B(j)
      const int i = w + 16 * h;
      if (c > j && i >= c) Db[i * 33 + c] = fma(-Lj[i], Lj[c], Db[i * 33 + c]);
      if (c == j && i >= j) Db[i * 33 + j] = Lj[i];
      if (i == j && c <= j) Xb[j * 33 + c] = Xb[j * 33 + c] * Db[j * 33 + 32];
    }
    __syncthreads();
  }
  if (write_back) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = w + 16 * h;
      if (i < bs && c <= i) A[(r0 + i) * ld + r0 + c] = Xb[i * 33 + c];
    }
  }
  __syncthreads();
  return *flag == 0;
}

__device__ __forceinline__ double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); }
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

// correlation k(r) (no outputscale) and g = -2 dk/d(r^2) so that dk/dl_k = g * s_k^2 / l_k with s_k = delta_k / l_k
__device__ __forceinline__ void corr_and_grad(int kind, double d2, double& kc, double& g) {
  if (kind == BOGP_KERNEL_RBF) {
    kc = exp(-0.5 * d2);
    g = kc;
  } else {
    const double s5 = 2.23606797749978969641;
    const double r = sqrt(fmax(d2, 1e-30));
    const double e = exp(-s5 * r);
    kc = (1.0 + s5 * r + (5.0 / 3.0) * d2) * e;
    g = (5.0 / 3.0) * (1.0 + s5 * r) * e;
  }
}

// kernels_gpfit_coop.cu: one fit over a cooperative grid (device pointers prepared by the caller; synchronises the stream)
cudaError_t launch_gp_fit_coop(FitArgs p, int sms, cudaStream_t s);

}  // namespace bogp
