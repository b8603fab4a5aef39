// GP hyper-parameter fit for the larger models (n above ~150: the n x n matrix no longer fits one SM's shared memory):
// the same loop as kernels_gpfit.cu -- steps x {kernel matrix, blocked Cholesky, explicit inverse, -mll / n, analytic
// gradient, AdamW} in ONE launch, float64 -- with every phase spread over a cooperative grid of G CTAs that meet on a
// grid-wide barrier (one atomic counter, all CTAs co-resident: cudaLaunchCooperativeKernel).  The matrices live in
// global memory (n = 512: 2 MB each, L2-resident).  Replaces the torch / cuSOLVER loop the round-1 build fell back to
// above n = 320 (ref/projects/swellex96_inv/data/bo/ei.py:69-78 with budgets up to 500 trials, ei.py:22-23).
//
// Per step (T = ceil(n / 32) block columns), barriers in brackets:
//   K(theta) lower triangle, rows over all warps of the grid                                          [1]
//   per block column: the 32 x 32 diagonal block is factored and inverted by EVERY CTA redundantly (same input, same
//     arithmetic: no broadcast of D^-1, which all of them need for the panel solve), panel rows over the grid [1],
//     trailing update rows over the grid [1]
//   X = L^-1 out of place, one column block per CTA (columns are independent)                        [1]
//   z = X r [1];  alpha = X^T z, K^-1 = X^T X (upper triangle), columns over the grid                [1]
//   gradient traces over the lower triangle, per-CTA partial sums in fixed order                     [1]
//   every CTA sums the G partials in the same order and takes the same AdamW step (no broadcast)
// 2 T + 5 barriers per step: n = 256 -> 21, n = 512 -> 37.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "bogp_internal.h"
#include "gpfit_device.cuh"

namespace bogp {

constexpr int kCoopPart = kFitMaxD + 8;      // per-CTA partial sums: g_ls[16] | g_os, g_noise, quad, logdet, sum(alpha)

struct CoopArgs {
  FitArgs f;             // work = A [n * ld]
  double* Xg;            // [n * ld]   X = L^-1 (lower), out of place
  double* Dinv;          // [T][32 * 33] inverses of the diagonal blocks
  double* vec;           // zv[n] | alpha[n] | kdiag[n]
  double* part;          // [G][kCoopPart]
  unsigned* bar;         // grid barrier counter (zeroed by the host)
  int* err;
};

// Grid-wide barrier: thread 0 of every CTA bumps the counter and spins until all G have arrived at this generation.
// Release / acquire at gpu scope around it (the fences also drop this SM's L1 lines, so the plain loads that follow see
// what the other CTAs wrote).  Bounded: a CTA that never arrives traps instead of hanging the device.
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned& target, unsigned nblocks, int* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v;
    long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if ((int)(v - target) >= 0) break;
      if (spin == 0) t0 = clock64();
      if ((spin & 4095u) == 4095u && clock64() - t0 > 6000000000LL) {
        atomicExch(err, 1 + (int)blockIdx.x);
        __threadfence_system();
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

#ifndef BOGP_COOP_DBG
#define BOGP_COOP_DBG 0
#endif
#define STAMP(i) do { if (BOGP_COOP_DBG && step == 3 && blockIdx.x == 0 && tid == 0) stamps[i] = clock64(); } while (0)

template <int DL>
__global__ void __launch_bounds__(kFitThreads, 1) gp_fit_coop_kernel(const CoopArgs q) {
  long long stamps[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  extern __shared__ double csm[];
  const FitArgs& p = q.f;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = p.n, d = p.d, ld = p.ld, np = d + 3;
  const int T = (n + 31) / 32;
  const int G = gridDim.x, W = G * kFitWarps, gw = blockIdx.x * kFitWarps + warp;
  double* A = p.work;
  double* Xg = q.Xg;
  double* zv = q.vec;
  double* alpha = q.vec + n;
  double* kdiag = q.vec + 2 * n;
  double* sp = csm;
  double* Xs = sp; if (p.x_in_smem) sp += n * d;
  const double* Xp = p.x_in_smem ? Xs : p.X;
  double* rv = sp; sp += n;            // y - mean (every CTA keeps its own copy)
  double* T1 = sp; sp += 32 * 33;      // inverse of the current diagonal block / block product
  double* Db = sp; sp += 32 * 33;
  double* Lc = sp; sp += 64;
  double* red = sp; sp += kFitWarps * kCoopPart;
  double* prm = sp; sp += kFitMaxD + 8;
  double* raw = sp; sp += kFitNP;
  double* am = sp; sp += kFitNP;
  double* av = sp; sp += kFitNP;
  double* gsum = sp; sp += kCoopPart;
  int* flag = reinterpret_cast<int*>(sp);
  unsigned bar_target = 0;
  auto gsync = [&]() { grid_sync(q.bar, bar_target, (unsigned)G, q.err); };

  if (p.x_in_smem)
    for (int i = tid; i < n * d; i += kFitThreads) Xs[i] = p.X[i];
  if (tid < np) { raw[tid] = p.raw[tid]; am[tid] = 0.0; av[tid] = 0.0; }
  if (tid == 0) *flag = 0;
  __syncthreads();

  for (int step = 1; step <= p.steps; ++step) {
    STAMP(0);
    // ---------------------------------------------------------------- constrained parameters (every CTA, identical)
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
      // rows paired (h, n - 1 - h): every warp gets the same number of elements
      for (int h = gw; h < (n + 1) / 2; h += W) {
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
          const int i = side ? n - 1 - h : h;
          if (side && i == h) break;
          for (int j = lane; j <= i; j += 32) {
            double d2 = 0.0;
#pragma unroll
            for (int k = 0; k < DL; ++k)
              if (k < d) { const double s = (Xp[i * d + k] - Xp[j * d + k]) * inv_ls[k]; d2 = fma(s, s, d2); }
            double kc, g;
            corr_and_grad(p.kind, d2, kc, g);
            A[(size_t)i * ld + j] = os * kc + (i == j ? noise + jitter : 0.0);
          }
        }
      }
      gsync();
      STAMP(1);
      pd = true;
      for (int kb = 0; kb < T; ++kb) {
        const int r0 = kb * 32, bs = min(32, n - r0), r1 = r0 + bs;
        // every CTA factors and inverts the diagonal block (T1 = D^-1 afterwards); CTA 0 files D^-1 for the later phases
        if (!block_chol_inv(A, ld, r0, bs, Db, T1, Lc, flag, false)) { pd = false; break; }
        if (blockIdx.x == 0)
          for (int e = tid; e < 32 * 33; e += kFitThreads) q.Dinv[(size_t)kb * 32 * 33 + e] = T1[e];
        if (r1 < n) {
          // panel: L_i,blk = A_i,blk D^-T, two rows per warp
          for (int i = r1 + 2 * gw; i < n; i += 2 * W) {
            const double* a0 = A + (size_t)i * ld + r0;
            const double* a1 = (i + 1 < n) ? a0 + ld : a0;
            const double* t = T1 + lane * 33;
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
            for (int c = 0; c < 32; ++c) { const double tv = t[c]; acc0 = fma(a0[c], tv, acc0); acc1 = fma(a1[c], tv, acc1); }
            __syncwarp();
            A[(size_t)i * ld + r0 + lane] = acc0;
            if (i + 1 < n) A[(size_t)(i + 1) * ld + r0 + lane] = acc1;
          }
          gsync();
          // trailing update: A_ij -= sum_c L_ic L_jc,  r1 <= j <= i: rows (i, i + 1) per warp, lanes over j
          for (int i = r1 + 2 * gw; i < n; i += 2 * W) {
            const bool two = i + 1 < n;
            const double* li0 = A + (size_t)i * ld + r0;
            const double* li1 = two ? li0 + ld : li0;
            const int jend = two ? i + 1 : i;
            for (int j = r1 + lane; j <= jend; j += 32) {
              const double* lj = A + (size_t)j * ld + r0;
              double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
              for (int c = 0; c < 32; ++c) { const double v = lj[c]; acc0 = fma(li0[c], v, acc0); acc1 = fma(li1[c], v, acc1); }
              if (j <= i) A[(size_t)i * ld + j] -= acc0;
              if (two) A[(size_t)(i + 1) * ld + j] -= acc1;
            }
          }
          gsync();
        }
      }
      if (!pd) {          // uniform over the grid (same data, same arithmetic): retry with more jitter
        __syncthreads();
        if (tid == 0) *flag = 0;
        gsync();
      }
    }
    if (!pd) {
      if (tid == 0 && blockIdx.x == 0) p.status[0] = 1;
      return;
    }
    gsync();          // D^-1 of the LAST diagonal block (filed by CTA 0 after the loop's last barrier) is read by every CTA below
    STAMP(2);

    // ---------------------------------------------------------------- X = L^-1 out of place, one column block per CTA
    for (int k = blockIdx.x; k < T; k += G) {
      const int c0 = k * 32, bsk = min(32, n - c0);
      const double* Dk = q.Dinv + (size_t)k * 32 * 33;
      for (int e = tid; e < 32 * 32; e += kFitThreads) {
        const int a = e >> 5, b = e & 31;
        if (a < bsk && b <= a) Xg[(size_t)(c0 + a) * ld + c0 + b] = Dk[a * 33 + b];
      }
      __syncthreads();
      for (int ib = k + 1; ib < T; ++ib) {
        const int i0 = ib * 32, bsi = min(32, n - i0);
        // S = sum_{j = k}^{ib - 1} L_(ib, j) X_(j, k): thread (warp a, lane b) owns S[a][b] and S[a + 16][b]
        {
          const int a = warp;
          const double* l0 = A + (size_t)(i0 + min(a, bsi - 1)) * ld;
          const double* l1 = A + (size_t)(i0 + min(a + 16, bsi - 1)) * ld;
          double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
          for (int c = 0; c < 32; ++c) {                     // j = k: X_kk is lower triangular
            const double xv = c >= lane ? Dk[c * 33 + lane] : 0.0;
            acc0 = fma(l0[c0 + c], xv, acc0); acc1 = fma(l1[c0 + c], xv, acc1);
          }
#pragma unroll 4
          for (int jj = c0 + 32; jj < i0; ++jj) {
            const double xv = Xg[(size_t)jj * ld + c0 + lane];
            acc0 = fma(l0[jj], xv, acc0); acc1 = fma(l1[jj], xv, acc1);
          }
          T1[a * 33 + lane] = acc0;
          T1[(a + 16) * 33 + lane] = acc1;
        }
        __syncthreads();
        // X_(ib, k) = - X_(ib, ib) S
        {
          const int a = warp;
          const double* Di = q.Dinv + (size_t)ib * 32 * 33;
          double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
          for (int c = 0; c < 32; ++c) {
            const double sv = T1[c * 33 + lane];
            acc0 = fma(c <= a ? Di[a * 33 + c] : 0.0, sv, acc0);
            acc1 = fma(c <= a + 16 ? Di[(a + 16) * 33 + c] : 0.0, sv, acc1);
          }
          if (a < bsi) Xg[(size_t)(i0 + a) * ld + c0 + lane] = -acc0;
          if (a + 16 < bsi) Xg[(size_t)(i0 + a + 16) * ld + c0 + lane] = -acc1;
        }
        __syncthreads();
      }
    }
    gsync();
    STAMP(3);

    // ---------------------------------------------------------------- z = X r (warp per row)
    for (int k = gw; k < n; k += W) {
      double acc = 0.0;
      for (int j = lane; j <= k; j += 32) acc = fma(Xg[(size_t)k * ld + j], rv[j], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
      if (lane == 0) zv[k] = acc;
    }
    gsync();
    STAMP(4);
    // ---------------------------------------------------------------- alpha = X^T z, log det, quadratic form (warp per column)
    double part_quad = 0.0, part_ld = 0.0, part_as = 0.0;        // lane 0 of every warp
    for (int j = gw; j < n; j += W) {
      double acc = 0.0;
      for (int k = j + lane; k < n; k += 32) acc = fma(Xg[(size_t)k * ld + j], zv[k], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
      if (lane == 0) {
        alpha[j] = acc;
        part_quad = fma(rv[j], acc, part_quad);
        part_ld -= log(Xg[(size_t)j * ld + j]);          // log L_jj = -log X_jj
        part_as += acc;
      }
    }
    // ---------------------------------------------------------------- K^-1 = X^T X: strict upper triangle of A + kdiag
    for (int j = 2 * gw; j < n; j += 2 * W) {      // columns (j, j + 1) per warp, lanes over i
      const bool two = j + 1 < n;
      const int iend = two ? j + 1 : j;
      for (int i = lane; i <= iend; i += 32) {
        double acc0 = i <= j ? Xg[(size_t)j * ld + i] * Xg[(size_t)j * ld + j] : 0.0, acc1 = 0.0;      // k = j term (column j only)
#pragma unroll 4
        for (int k = j + 1; k < n; ++k) {
          const double xi = Xg[(size_t)k * ld + i];
          acc0 = fma(xi, Xg[(size_t)k * ld + j], acc0);
          acc1 = fma(xi, Xg[(size_t)k * ld + j + 1], acc1);
        }
        if (i < j) A[(size_t)i * ld + j] = acc0; else if (i == j) kdiag[j] = acc0;
        if (two) { if (i <= j) A[(size_t)i * ld + j + 1] = acc1; else kdiag[j + 1] = acc1; }
      }
    }
    gsync();
    STAMP(5);

    // ---------------------------------------------------------------- gradient traces over the lower triangle
    double g_ls[DL], g_os = 0.0, g_noise = 0.0;
#pragma unroll
    for (int k = 0; k < DL; ++k) g_ls[k] = 0.0;
    for (int h = gw; h < (n + 1) / 2; h += W) {
#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        const int i = side ? n - 1 - h : h;
        if (side && i == h) break;
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
          const double kinv = i == j ? kdiag[i] : A[(size_t)j * ld + i];
          const double Wt = (i == j ? 0.5 : 1.0) * (ai * alpha[j] - kinv);
          g_os = fma(Wt, kc, g_os);
          if (i == j) g_noise += Wt;
          const double wg = Wt * os * g;
#pragma unroll
          for (int k = 0; k < DL; ++k)
            if (k < d) g_ls[k] = fma(wg, s2[k] * inv_ls[k], g_ls[k]);
        }
      }
    }
    // fixed-order reductions: lanes -> warp, warps -> CTA partial, CTA partials -> every CTA sums all G of them
    {
      double vals[DL + 2];
#pragma unroll
      for (int k = 0; k < DL; ++k) vals[k] = g_ls[k];
      vals[DL] = g_os; vals[DL + 1] = g_noise;
#pragma unroll
      for (int qq = 0; qq < DL + 2; ++qq) {
        double v = vals[qq];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0) red[warp * kCoopPart + (qq < DL ? qq : kFitMaxD + qq - DL)] = v;
      }
      if (lane == 0) {
        red[warp * kCoopPart + kFitMaxD + 2] = part_quad;
        red[warp * kCoopPart + kFitMaxD + 3] = part_ld;
        red[warp * kCoopPart + kFitMaxD + 4] = part_as;
        for (int k = DL; k < kFitMaxD; ++k) red[warp * kCoopPart + k] = 0.0;
      }
    }
    __syncthreads();
    if (tid < kFitMaxD + 5) {
      double v = 0.0;
      for (int w = 0; w < kFitWarps; ++w) v += red[w * kCoopPart + tid];
      q.part[(size_t)blockIdx.x * kCoopPart + tid] = v;
    }
    gsync();
    if (tid < kFitMaxD + 5) {
      double v = 0.0;
      for (int b = 0; b < G; ++b) v += q.part[(size_t)b * kCoopPart + tid];
      gsum[tid] = v;
    }
    __syncthreads();

    STAMP(6);
    if (BOGP_COOP_DBG && step == 3 && blockIdx.x == 0 && tid == 0)
      printf("[coop dbg] n=%d G=%d cycles: K %lld | chol %lld | inverse %lld | z %lld | alpha+Kinv %lld | grad %lld\n", n, G, stamps[1] - stamps[0],
             stamps[2] - stamps[1], stamps[3] - stamps[2], stamps[4] - stamps[3], stamps[5] - stamps[4], stamps[6] - stamps[5]);
    // ---------------------------------------------------------------- loss, chain rule, AdamW (thread 0 of every CTA)
    if (tid == 0) {
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
        grad[d + 2] = -gsum[kFitMaxD + 4] / nn;
      }
      if (p.loss && blockIdx.x == 0) p.loss[step - 1] = -ll / nn;
      if (p.grad_out && blockIdx.x == 0)
        for (int k = 0; k < np; ++k) p.grad_out[k] = grad[k];
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
  if (blockIdx.x == 0) {
    if (tid < np) p.raw[tid] = raw[tid];
    if (tid == 0) p.status[0] = 0;
  }
}

// One fit (B = 1 per launch; the caller loops over starting points).  Device pointers X, y, raw, loss, status as
// prepared by bogp_gp_fit_adamw_bounded; returns a CUDA error code.
cudaError_t launch_gp_fit_coop(FitArgs p, int sms, cudaStream_t s) {
  const int n = p.n, d = p.d;
  const int T = (n + 31) / 32;
  int G = std::max(4, std::min(std::min(sms, 32), n / 16));      // measured: no gain beyond ~n / 16 CTAs (the column chain of the diagonal blocks is serial)
  if (const char* e = std::getenv("BOGP_FIT_COOP_CTAS")) G = std::max(1, std::min(sms, std::atoi(e)));
  const size_t mat = (size_t)n * p.ld * sizeof(double);
  const size_t bytes = 2 * mat + (size_t)T * 32 * 33 * 8 + (size_t)3 * n * 8 + (size_t)G * kCoopPart * 8 + 512;
  unsigned char* ws = nullptr;
  cudaError_t e = cudaMalloc(&ws, bytes);
  if (e != cudaSuccess) return e;
  CoopArgs q{};
  q.f = p;
  size_t off = 0;
  q.f.work = reinterpret_cast<double*>(ws + off); off += mat;
  q.Xg = reinterpret_cast<double*>(ws + off); off += mat;
  q.Dinv = reinterpret_cast<double*>(ws + off); off += (size_t)T * 32 * 33 * 8;
  q.vec = reinterpret_cast<double*>(ws + off); off += (size_t)3 * n * 8;
  q.part = reinterpret_cast<double*>(ws + off); off += (size_t)G * kCoopPart * 8;
  q.bar = reinterpret_cast<unsigned*>(ws + off); off += 256;
  q.err = reinterpret_cast<int*>(ws + off);
  const size_t fixed = ((size_t)n + 2 * 32 * 33 + 64 + kFitWarps * kCoopPart + (kFitMaxD + 8) + 3 * kFitNP + kCoopPart + 2) * sizeof(double);
  const size_t x_bytes = (size_t)n * d * sizeof(double);
  q.f.x_in_smem = fixed + x_bytes <= 200 * 1024 ? 1 : 0;
  const size_t smem = fixed + (q.f.x_in_smem ? x_bytes : 0);
  const int dl = d <= 4 ? 4 : (d <= 8 ? 8 : 16);
  void* kern = dl == 4 ? (void*)gp_fit_coop_kernel<4> : dl == 8 ? (void*)gp_fit_coop_kernel<8> : (void*)gp_fit_coop_kernel<16>;
  e = cudaMemsetAsync(ws + 2 * mat, 0, bytes - 2 * mat, s);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    void* args[] = {&q};
    cudaEvent_t ev = timing_begin(s);
    e = cudaLaunchCooperativeKernel(kern, dim3(G), dim3(kFitThreads), args, smem, s);
    timing_end(ev, s);
    count_launch();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(ws);
  return e;
}

}  // namespace bogp
