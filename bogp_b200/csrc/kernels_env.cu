// Geoacoustic-inversion scoring (config 5): every candidate environment has its own mode set, so
// nothing is shared between candidates and the kernel is HBM-bound: per (candidate, tonal) it streams
// Phi_rec^T [Mmax, N] float32 once (4*N*M bytes) plus k and the source amplitudes.
// Replaces the pool of MatchedFieldProcessor calls of ref/projects/swellex96_inv/data/bo/obj.py:28-30, 73-83.
// Also: arg-max / arg-min reduction (ref/projects/swellex96_localization/data/collate.py:50).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "bogp_internal.h"
#include "device_math.cuh"

struct bogp_envbatch_s {
  long long C = 0;
  int F = 0, N = 0, Mmax = 0;
  double *k_re = nullptr, *k_im = nullptr, *range = nullptr;
  float *amp_re = nullptr, *amp_im = nullptr, *phiT = nullptr;
  float *CkT_re = nullptr, *CkT_im = nullptr;   // [F][N][Jmax]
  int* J = nullptr;                             // [F]
  float* trK = nullptr;                         // [F]
  int Jmax = 0;
  bool has_cov = false;
  // streaming layout: one contiguous record per candidate (see envstream_kernel); null when a record cannot be
  // double-buffered in shared memory
  unsigned char* rec = nullptr;
  unsigned rec_bytes = 0, off_kre = 0, off_kim = 0, off_g = 0, off_r = 0;
  // tilted array (bogp_envbatch_set_tilt): per candidate sin(tilt) and per (candidate, receiver) lever z_pivot - z_n
  double *sin_tilt = nullptr, *lever = nullptr;
  bool tilted = false;
};

namespace bogp {

constexpr int kEnvWarps = 8;

// One warp per candidate; lanes run over receivers (coalesced rows of Phi^T), modes are broadcast
// from shared memory.
__global__ void __launch_bounds__(kEnvWarps * 32)
envbatch_kernel(long long C, int F, int N, int Mmax, const double* __restrict__ k_re, const double* __restrict__ k_im,
                const float* __restrict__ amp_re, const float* __restrict__ amp_im, const float* __restrict__ phiT,
                const double* __restrict__ range, const float* __restrict__ CkT_re, const float* __restrict__ CkT_im,
                const int* __restrict__ Jf, const float* __restrict__ trK, int Jmax, int atype, int mf,
                float* __restrict__ out) {
  extern __shared__ float2 smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* a_s = smem + (size_t)warp * (Mmax + N);
  float2* p_s = a_s + Mmax;
  for (long long c = (long long)blockIdx.x * kEnvWarps + warp; c < C; c += (long long)gridDim.x * kEnvWarps) {
    const double r = range[c];
    const float rs = (float)rsqrt(r);
    float acc = 0.f;
    for (int f = 0; f < F; ++f) {
      const size_t cf = (size_t)c * F + f;
      for (int m = lane; m < Mmax; m += 32) {
        const double kre = k_re[cf * Mmax + m], kim = k_im[cf * Mmax + m];
        const float ar = amp_re[cf * Mmax + m], ai = amp_im ? amp_im[cf * Mmax + m] : 0.f;
        float2 a = make_float2(0.f, 0.f);
        if (ar != 0.f || ai != 0.f) {
          // k^{-1/2} in float64 (one per mode), propagation with float64 phase reduction
          const double mag = rsqrt(sqrt(kre * kre + kim * kim)), ang = -0.5 * atan2(kim, kre);
          double s2, c2;
          sincos(ang, &s2, &c2);
          float s, co;
          sincosf(reduce_phase(kre * r), &s, &co);
          const float amp = expf((float)(kim * r)) * rs * (float)mag;
          a = cmul(cmul(make_float2(ar, ai), make_float2((float)c2, (float)s2)), make_float2(amp * co, -amp * s));
        }
        a_s[m] = a;
      }
      __syncwarp();
      const float* P = phiT + cf * (size_t)Mmax * N;
      float norm = 0.f;
      for (int n = lane; n < N; n += 32) {
        float pr = 0.f, pi = 0.f;
#pragma unroll 4
        for (int m = 0; m < Mmax; ++m) {
          const float ph = __ldg(P + (size_t)m * N + n);
          const float2 a = a_s[m];
          pr = fmaf(ph, a.x, pr);
          pi = fmaf(ph, a.y, pi);
        }
        p_s[n] = make_float2(pr, pi);
        norm += pr * pr + pi * pi;
      }
      __syncwarp();
      float num = 0.f;
      const int J = Jf[f];
      const float* cr = CkT_re + (size_t)f * N * Jmax;
      const float* ci = CkT_im + (size_t)f * N * Jmax;
      for (int j = lane; j < J; j += 32) {
        float sr = 0.f, si = 0.f;
        for (int n = 0; n < N; ++n) {
          const float2 p = p_s[n];
          const float a = __ldg(cr + (size_t)n * Jmax + j), b = __ldg(ci + (size_t)n * Jmax + j);
          sr += a * p.x - b * p.y;
          si += a * p.y + b * p.x;
        }
        num += sr * sr + si * si;
      }
      norm = warp_sum(norm);
      num = warp_sum(num);
      __syncwarp();
      const float b = num / norm;
      acc = combine_tonal(acc, atype == BOGP_ATYPE_CBF ? b : trK[f] - b, mf, f == 0);
    }
    if (lane == 0) out[c] = (mf == BOGP_MF_MEAN) ? acc / (float)F : acc;
  }
}

// ---- tilted array, one environment per candidate (the inversion's search space has `tilt` in [-3, 3] degrees,
// ref/projects/swellex96_inv/data/bo/common.py:63, and the simulated truth is tilted): receiver n sits at range
// r_n = r - (z_pivot - z_n) sin(tau), so the phase depends on (receiver, mode) and the replica is evaluated term by term
//     p_n = r_n^-1/2 sum_m Phi[n, m] g_m exp(-i k_m r_n),   g_m = phi_m(z_s) k_m^-1/2
// with the phase product and its reduction in float64 (as everywhere else).  One warp per candidate, lanes over
// receivers; same Bartlett epilogue as envbatch_kernel.
__global__ void __launch_bounds__(kEnvWarps * 32)
envbatch_tilt_kernel(long long C, int F, int N, int Mmax, const double* __restrict__ k_re, const double* __restrict__ k_im,
                     const float* __restrict__ amp_re, const float* __restrict__ amp_im, const float* __restrict__ phiT,
                     const double* __restrict__ range, const double* __restrict__ sin_tilt, const double* __restrict__ lever,
                     const float* __restrict__ CkT_re, const float* __restrict__ CkT_im, const int* __restrict__ Jf,
                     const float* __restrict__ trK, int Jmax, int atype, int mf, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char tsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t per_warp = (size_t)Mmax * (sizeof(double) + sizeof(float) + sizeof(float2)) + (size_t)N * sizeof(float2);
  unsigned char* base = tsm + (size_t)warp * ((per_warp + 15) & ~(size_t)15);
  double* kre_s = reinterpret_cast<double*>(base);
  float2* g_s = reinterpret_cast<float2*>(kre_s + Mmax);
  float2* p_s = g_s + Mmax;
  float* kim_s = reinterpret_cast<float*>(p_s + N);
  for (long long c = (long long)blockIdx.x * kEnvWarps + warp; c < C; c += (long long)gridDim.x * kEnvWarps) {
    const double r = range[c], st = sin_tilt[c];
    float acc = 0.f;
    for (int f = 0; f < F; ++f) {
      const size_t cf = (size_t)c * F + f;
      for (int m = lane; m < Mmax; m += 32) {
        const double kre = k_re[cf * Mmax + m], kim = k_im[cf * Mmax + m];
        const float ar = amp_re[cf * Mmax + m], ai = amp_im ? amp_im[cf * Mmax + m] : 0.f;
        float2 g = make_float2(0.f, 0.f);
        if (ar != 0.f || ai != 0.f) {
          const double mag = rsqrt(sqrt(kre * kre + kim * kim)), ang = -0.5 * atan2(kim, kre);
          double s2, c2;
          sincos(ang, &s2, &c2);
          g = cmul(make_float2(ar, ai), make_float2((float)(mag * c2), (float)(mag * s2)));
        }
        kre_s[m] = kre; kim_s[m] = (float)kim; g_s[m] = g;
      }
      __syncwarp();
      const float* P = phiT + cf * (size_t)Mmax * N;
      float norm = 0.f;
      for (int n = lane; n < N; n += 32) {
        const double rn = r - lever[(size_t)c * N + n] * st;
        const float rnf = (float)rn, rs = (float)rsqrt(rn);
        float pr = 0.f, pi = 0.f;
        for (int m = 0; m < Mmax; ++m) {
          const float2 g = g_s[m];
          if (g.x == 0.f && g.y == 0.f) continue;
          float s, co;
          sincos_big(kre_s[m] * rn, s, co);
          const float a = expf(kim_s[m] * rnf) * rs * __ldg(P + (size_t)m * N + n);
          const float2 t = cmul(g, make_float2(a * co, -a * s));
          pr += t.x; pi += t.y;
        }
        p_s[n] = make_float2(pr, pi);
        norm += pr * pr + pi * pi;
      }
      __syncwarp();
      float num = 0.f;
      const int J = Jf[f];
      const float* cr = CkT_re + (size_t)f * N * Jmax;
      const float* ci = CkT_im + (size_t)f * N * Jmax;
      for (int j = lane; j < J; j += 32) {
        float sr = 0.f, si = 0.f;
        for (int n = 0; n < N; ++n) {
          const float2 pp = p_s[n];
          const float a = __ldg(cr + (size_t)n * Jmax + j), b = __ldg(ci + (size_t)n * Jmax + j);
          sr += a * pp.x - b * pp.y;
          si += a * pp.y + b * pp.x;
        }
        num += sr * sr + si * si;
      }
      norm = warp_sum(norm);
      num = warp_sum(num);
      __syncwarp();
      const float b = num / norm;
      acc = combine_tonal(acc, atype == BOGP_ATYPE_CBF ? b : trK[f] - b, mf, f == 0);
    }
    if (lane == 0) out[c] = (mf == BOGP_MF_MEAN) ? acc / (float)F : acc;
  }
}

// ---- streaming variant (the fast path).  Each candidate is ONE contiguous record in HBM
//   [Phi^T (F x Mmax x N float32) | k_re (F x Mmax float64) | k_im (float32) | g = phi_src * k^-1/2 (float2) | range]
// pulled into shared memory by a single cp.async.bulk per candidate (mbarrier complete_tx), S stages per warp, so
// the HBM stream never waits for the arithmetic: a warp computes candidate c from shared memory while the copies
// of the next candidates (its own next stage and the other warps') are in flight.  Algorithmic bytes per
// (candidate, tonal): 4 N M + 20 M; the record IS those bytes (plus zero padding to Mmax).
__device__ __forceinline__ uint32_t env_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void env_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

__device__ __forceinline__ void env_fetch(uint32_t dst, const unsigned char* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

struct EnvStreamParams {
  long long C;
  int F, N, Mmax, Jmax, atype, mf, warps, stages;
  unsigned rec_bytes, off_kre, off_kim, off_g, off_r, warp_bytes;
  const unsigned char* rec;
  const float* CkT_re;
  const float* CkT_im;
  const int* J;
  const float* trK;
  float* out;
};

constexpr int kEnvJ = 4;      // covariance ranks handled in registers by the streaming kernel

__global__ void __launch_bounds__(512) envstream_kernel(const EnvStreamParams p) {
  extern __shared__ __align__(128) unsigned char esm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* base = esm + (size_t)warp * p.warp_bytes;                 // [stages][rec_bytes] then a_s
  float2* a_s = reinterpret_cast<float2*>(base + (size_t)p.stages * p.rec_bytes);
  const uint32_t bar0 = env_smem_u32(esm + (size_t)p.warps * p.warp_bytes) + 16u * (uint32_t)warp;   // 2 barriers per warp
  if (lane == 0) {
    for (int st = 0; st < p.stages; ++st) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * st));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const long long TW = (long long)gridDim.x * p.warps;
  long long c = (long long)blockIdx.x * p.warps + warp;
  const int FM = p.F * p.Mmax;
  // prologue: fill the ring
  if (lane == 0) {
    long long cc = c;
    for (int st = 0; st < p.stages && cc < p.C; ++st, cc += TW)
      env_fetch(env_smem_u32(base + (size_t)st * p.rec_bytes), p.rec + (size_t)cc * p.rec_bytes, p.rec_bytes, bar0 + 8u * st);
  }
  int st = 0;
  uint32_t phase = 0;      // bit st = parity of stage st
  for (; c < p.C; c += TW) {
    env_wait(bar0 + 8u * st, (phase >> st) & 1u);
    phase ^= 1u << st;
    const unsigned char* buf = base + (size_t)st * p.rec_bytes;
    const float* P = reinterpret_cast<const float*>(buf);
    const double* kre = reinterpret_cast<const double*>(buf + p.off_kre);
    const float* kim = reinterpret_cast<const float*>(buf + p.off_kim);
    const float2* g = reinterpret_cast<const float2*>(buf + p.off_g);
    const double r = *reinterpret_cast<const double*>(buf + p.off_r);
    const float rf = (float)r, rs = (float)rsqrt(r);
    for (int idx = lane; idx < FM; idx += 32) {
      const float2 gg = g[idx];
      float2 a = make_float2(0.f, 0.f);
      if (gg.x != 0.f || gg.y != 0.f) {
        float s, co;
        sincos_big(kre[idx] * r, s, co);
        const float amp = expf(kim[idx] * rf) * rs;
        a = cmul(gg, make_float2(amp * co, -amp * s));
      }
      a_s[idx] = a;
    }
    __syncwarp();
    float acc = 0.f;
    for (int f = 0; f < p.F; ++f) {
      const float* Pf = P + (size_t)f * p.Mmax * p.N;
      const float2* af = a_s + f * p.Mmax;
      const int J = p.J[f];
      const float* cr = p.CkT_re + (size_t)f * p.N * p.Jmax;
      const float* ci = p.CkT_im + (size_t)f * p.N * p.Jmax;
      float norm = 0.f, sr[kEnvJ], si[kEnvJ];
#pragma unroll
      for (int j = 0; j < kEnvJ; ++j) sr[j] = si[j] = 0.f;
      for (int n = lane; n < p.N; n += 32) {
        // four modes per step, eight independent accumulators: the loop is bound by shared-memory load latency
        // (12 warps per SM), so the loads of a whole unrolled body are issued before its FMAs
        float pr0 = 0.f, pi0 = 0.f, pr1 = 0.f, pi1 = 0.f, pr2 = 0.f, pi2 = 0.f, pr3 = 0.f, pi3 = 0.f;
        const float* pn = Pf + n;
        const int N = p.N;
#pragma unroll 4
        for (int m = 0; m < p.Mmax; m += 4) {          // Mmax is a multiple of 4
          const float ph0 = pn[m * N], ph1 = pn[(m + 1) * N], ph2 = pn[(m + 2) * N], ph3 = pn[(m + 3) * N];
          const float4 a01 = *reinterpret_cast<const float4*>(af + m), a23 = *reinterpret_cast<const float4*>(af + m + 2);
          pr0 = fmaf(ph0, a01.x, pr0); pi0 = fmaf(ph0, a01.y, pi0);
          pr1 = fmaf(ph1, a01.z, pr1); pi1 = fmaf(ph1, a01.w, pi1);
          pr2 = fmaf(ph2, a23.x, pr2); pi2 = fmaf(ph2, a23.y, pi2);
          pr3 = fmaf(ph3, a23.z, pr3); pi3 = fmaf(ph3, a23.w, pi3);
        }
        const float pr = (pr0 + pr1) + (pr2 + pr3), pi = (pi0 + pi1) + (pi2 + pi3);
        norm = fmaf(pr, pr, fmaf(pi, pi, norm));
#pragma unroll
        for (int j = 0; j < kEnvJ; ++j)
          if (j < J) {
            const float a = __ldg(cr + (size_t)n * p.Jmax + j), b = __ldg(ci + (size_t)n * p.Jmax + j);
            sr[j] += a * pr - b * pi;
            si[j] += a * pi + b * pr;
          }
      }
      norm = warp_sum(norm);
      float num = 0.f;
#pragma unroll
      for (int j = 0; j < kEnvJ; ++j)
        if (j < J) {
          const float x = warp_sum(sr[j]), y = warp_sum(si[j]);
          num += x * x + y * y;
        }
      const float b = num / norm;
      acc = combine_tonal(acc, p.atype == BOGP_ATYPE_CBF ? b : p.trK[f] - b, p.mf, f == 0);
    }
    if (lane == 0) p.out[c] = (p.mf == BOGP_MF_MEAN) ? acc / (float)p.F : acc;
    __syncwarp();                                     // every lane is done with this stage's buffer
    const long long cn = c + (long long)p.stages * TW;
    if (cn < p.C && lane == 0)
      env_fetch(env_smem_u32(base + (size_t)st * p.rec_bytes), p.rec + (size_t)cn * p.rec_bytes, p.rec_bytes, bar0 + 8u * st);
    st = (st + 1 == p.stages) ? 0 : st + 1;
  }
}

// ------------------------------------------------------------------------------------------------ argmax
struct ValIdx {
  float v;
  long long i;
};

template <bool MAX>
__device__ __forceinline__ ValIdx better(ValIdx a, ValIdx b) {
  // NaNs never win; ties go to the lower index (np.argmax / np.argmin: first occurrence)
  const bool b_wins = (b.i >= 0) && ((a.i < 0) || (MAX ? b.v > a.v : b.v < a.v) || (b.v == a.v && b.i < a.i));
  return b_wins ? b : a;
}

template <bool MAX>
__device__ __forceinline__ ValIdx block_reduce(ValIdx x) {
  __shared__ float sv[32];
  __shared__ long long si[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = better<MAX>(x, y);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = x.v; si[warp] = x.i; }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    x.v = lane < nw ? sv[lane] : 0.f;
    x.i = lane < nw ? si[lane] : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ValIdx y;
      y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
      y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
      x = better<MAX>(x, y);
    }
  }
  return x;
}

template <bool MAX>
__global__ void argext_stage1(const float* __restrict__ x, long long n, float* __restrict__ pv, long long* __restrict__ pi) {
  ValIdx best{0.f, -1};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (v == v) best = better<MAX>(best, ValIdx{v, i});
  }
  best = block_reduce<MAX>(best);
  if (threadIdx.x == 0) { pv[blockIdx.x] = best.v; pi[blockIdx.x] = best.i; }
}

template <bool MAX>
__global__ void argext_stage2(const float* __restrict__ pv, const long long* __restrict__ pi, int nb, float* ov, long long* oi) {
  ValIdx best{0.f, -1};
  for (int i = threadIdx.x; i < nb; i += blockDim.x) best = better<MAX>(best, ValIdx{pv[i], pi[i]});
  best = block_reduce<MAX>(best);
  if (threadIdx.x == 0) { *ov = best.v; *oi = best.i; }
}

template <bool MAX>
static int argext(const float* x_dev, long long n, float* val_host, long long* idx_host, cudaStream_t s) {
  BOGP_REQUIRE(x_dev && val_host && idx_host && n >= 1, "bogp_arg%s_f32: bad argument", MAX ? "max" : "min");
  const int nb = (int)std::max<long long>(1, std::min<long long>((n + 1023) / 1024, 592));
  float* pv = nullptr;
  BOGP_CUDA(cudaMalloc(&pv, (size_t)(nb + 1) * (sizeof(float) + sizeof(long long)) + 16));
  long long* pi = reinterpret_cast<long long*>(pv + ((nb + 1 + 1) & ~1));
  float* ov = pv + nb;
  long long* oi = pi + nb;
  argext_stage1<MAX><<<nb, 256, 0, s>>>(x_dev, n, pv, pi);
  bogp::count_launch();
  argext_stage2<MAX><<<1, 256, 0, s>>>(pv, pi, nb, ov, oi);
  bogp::count_launch();
  cudaError_t e = cudaMemcpyAsync(val_host, ov, sizeof(float), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(idx_host, oi, sizeof(long long), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(pv);
  if (e != cudaSuccess) { set_error("bogp_argext: %s", cudaGetErrorString(e)); return BOGP_ERR_CUDA; }
  return 0;
}

}  // namespace bogp

extern "C" int bogp_argmax_f32(const float* x_dev, long long n, float* val_host, long long* idx_host, void* stream) {
  return bogp::argext<true>(x_dev, n, val_host, idx_host, static_cast<cudaStream_t>(stream));
}
extern "C" int bogp_argmin_f32(const float* x_dev, long long n, float* val_host, long long* idx_host, void* stream) {
  return bogp::argext<false>(x_dev, n, val_host, idx_host, static_cast<cudaStream_t>(stream));
}
extern "C" int bogp_argext_f32_dev(const float* x_dev, long long n, int maximize, float* val_dev, long long* idx_dev,
                                   void* scratch_dev, long long scratch_bytes, void* stream) {
  BOGP_REQUIRE(x_dev && val_dev && idx_dev && scratch_dev && n >= 1, "bogp_argext_f32_dev: bad argument");
  BOGP_REQUIRE(scratch_bytes >= 8192, "bogp_argext_f32_dev: scratch must be >= 8192 bytes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nb = (int)std::max<long long>(1, std::min<long long>((n + 1023) / 1024, 592));
  long long* pi = static_cast<long long*>(scratch_dev);          // [592] int64 then [592] float
  float* pv = reinterpret_cast<float*>(pi + 592);
  if (maximize) {
    bogp::argext_stage1<true><<<nb, 256, 0, s>>>(x_dev, n, pv, pi);
    bogp::argext_stage2<true><<<1, 256, 0, s>>>(pv, pi, nb, val_dev, idx_dev);
  } else {
    bogp::argext_stage1<false><<<nb, 256, 0, s>>>(x_dev, n, pv, pi);
    bogp::argext_stage2<false><<<1, 256, 0, s>>>(pv, pi, nb, val_dev, idx_dev);
  }
  bogp::count_launch(2);
  BOGP_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ env batch API
template <typename T>
static int dev_copy(T** dst, const T* src, size_t n) {
  BOGP_CUDA(cudaMalloc(dst, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) BOGP_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int bogp_envbatch_create(bogp_envbatch_t* out, long long C, int F, int N, int Mmax, const double* k_re_host,
                                    const double* k_im_host, const double* amp_re_host, const double* amp_im_host,
                                    const float* phi_rec_T_host, const double* range_m_host) {
  BOGP_REQUIRE(out, "bogp_envbatch_create: out is NULL");
  *out = nullptr;
  BOGP_REQUIRE(C >= 1 && F >= 1 && F <= BOGP_MAX_TONALS && N >= 1 && N <= 1024 && Mmax >= 1 && Mmax <= 1024,
               "bogp_envbatch_create: bad sizes C=%lld F=%d N=%d Mmax=%d", C, F, N, Mmax);
  BOGP_REQUIRE(k_re_host && k_im_host && amp_re_host && phi_rec_T_host && range_m_host, "bogp_envbatch_create: NULL input");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { bogp::set_error("no CUDA device available; no CPU fallback"); return BOGP_ERR_CUDA; }
  bogp_envbatch_s* h = new bogp_envbatch_s();
  h->C = C; h->F = F; h->N = N; h->Mmax = Mmax;
  const size_t nm = (size_t)C * F * Mmax;
  std::vector<float> ar(nm), ai(amp_im_host ? nm : 0);
  for (size_t i = 0; i < nm; ++i) { ar[i] = (float)amp_re_host[i]; if (amp_im_host) ai[i] = (float)amp_im_host[i]; }
  int rc = 0;
  rc |= dev_copy(&h->k_re, k_re_host, nm);
  rc |= dev_copy(&h->k_im, k_im_host, nm);
  rc |= dev_copy(&h->amp_re, ar.data(), nm);
  if (amp_im_host) rc |= dev_copy(&h->amp_im, ai.data(), nm);
  rc |= dev_copy(&h->phiT, phi_rec_T_host, nm * N);
  rc |= dev_copy(&h->range, range_m_host, (size_t)C);
  if (rc) { bogp_envbatch_destroy(h); return BOGP_ERR_CUDA; }
  // streaming records (fast path): [Phi^T | k_re f64 | k_im f32 | g = phi_src k^-1/2 (float2) | range f64], 16-byte padded
  if (Mmax % 4 == 0) {
    const size_t FM = (size_t)F * Mmax;
    const auto pad16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t o_kre = pad16(FM * N * 4), o_kim = o_kre + FM * 8, o_g = pad16(o_kim + FM * 4), o_r = o_g + FM * 8;
    const size_t R = pad16(o_r + 8);
    if (2 * R + FM * 8 + 64 <= 220 * 1024 && (unsigned long long)C * R < (1ull << 40)) {
      std::vector<unsigned char> rec((size_t)C * R, 0);
      for (long long c = 0; c < C; ++c) {
        unsigned char* b = rec.data() + (size_t)c * R;
        std::memcpy(b, phi_rec_T_host + (size_t)c * FM * N, FM * N * 4);
        double* kr = reinterpret_cast<double*>(b + o_kre);
        float* ki = reinterpret_cast<float*>(b + o_kim);
        float* g = reinterpret_cast<float*>(b + o_g);
        for (size_t i = 0; i < FM; ++i) {
          const size_t s = (size_t)c * FM + i;
          kr[i] = k_re_host[s];
          ki[i] = (float)k_im_host[s];
          const bogp::cplx amp(amp_re_host[s], amp_im_host ? amp_im_host[s] : 0.0);
          const bogp::cplx gg = (amp == bogp::cplx(0.0, 0.0)) ? amp : amp / std::sqrt(bogp::cplx(k_re_host[s], k_im_host[s]));
          g[2 * i] = (float)gg.real();
          g[2 * i + 1] = (float)gg.imag();
        }
        *reinterpret_cast<double*>(b + o_r) = range_m_host[c];
      }
      if (dev_copy(&h->rec, rec.data(), rec.size())) { bogp_envbatch_destroy(h); return BOGP_ERR_CUDA; }
      h->rec_bytes = (unsigned)R; h->off_kre = (unsigned)o_kre; h->off_kim = (unsigned)o_kim;
      h->off_g = (unsigned)o_g; h->off_r = (unsigned)o_r;
    }
  }
  *out = h;
  return 0;
}

extern "C" int bogp_envbatch_set_covariance(bogp_envbatch_t h, const double* K_re_host, const double* K_im_host) {
  BOGP_REQUIRE(h && K_re_host && K_im_host, "bogp_envbatch_set_covariance: NULL argument");
  const int F = h->F, N = h->N;
  std::vector<std::vector<bogp::cplx>> Ls(F);
  std::vector<int> J(F);
  std::vector<float> tr(F);
  int Jmax = 1;
  for (int f = 0; f < F; ++f) {
    std::vector<bogp::cplx> K((size_t)N * N);
    double t = 0.0;
    for (int a = 0; a < N; ++a)
      for (int b = 0; b < N; ++b) {
        size_t ab = (size_t)f * N * N + (size_t)a * N + b, ba = (size_t)f * N * N + (size_t)b * N + a;
        K[(size_t)a * N + b] = 0.5 * (bogp::cplx(K_re_host[ab], K_im_host[ab]) + std::conj(bogp::cplx(K_re_host[ba], K_im_host[ba])));
      }
    for (int a = 0; a < N; ++a) t += K[(size_t)a * N + a].real();
    tr[f] = (float)t;
    bogp::pivoted_cholesky(K, N, 1e-13, Ls[f], J[f]);
    Jmax = std::max(Jmax, J[f]);
  }
  std::vector<float> cr((size_t)F * N * Jmax, 0.f), ci((size_t)F * N * Jmax, 0.f);
  for (int f = 0; f < F; ++f)
    for (int n = 0; n < N; ++n)
      for (int j = 0; j < J[f]; ++j) {
        cr[((size_t)f * N + n) * Jmax + j] = (float)Ls[f][(size_t)n * N + j].real();
        ci[((size_t)f * N + n) * Jmax + j] = (float)(-Ls[f][(size_t)n * N + j].imag());
      }
  cudaFree(h->CkT_re); cudaFree(h->CkT_im); cudaFree(h->J); cudaFree(h->trK);
  h->CkT_re = h->CkT_im = nullptr; h->J = nullptr; h->trK = nullptr;
  int rc = 0;
  rc |= dev_copy(&h->CkT_re, cr.data(), cr.size());
  rc |= dev_copy(&h->CkT_im, ci.data(), ci.size());
  rc |= dev_copy(&h->J, J.data(), (size_t)F);
  rc |= dev_copy(&h->trK, tr.data(), (size_t)F);
  if (rc) return BOGP_ERR_CUDA;
  h->Jmax = Jmax;
  h->has_cov = true;
  return 0;
}

extern "C" int bogp_envbatch_set_tilt(bogp_envbatch_t h, const double* tilt_deg_host, const double* lever_host) {
  BOGP_REQUIRE(h != nullptr, "bogp_envbatch_set_tilt: NULL handle");
  cudaFree(h->sin_tilt); cudaFree(h->lever);
  h->sin_tilt = h->lever = nullptr;
  h->tilted = false;
  if (!tilt_deg_host) return 0;                      // back to a vertical array
  BOGP_REQUIRE(lever_host != nullptr, "bogp_envbatch_set_tilt: lever is NULL");
  std::vector<double> st((size_t)h->C);
  bool any = false;
  for (long long c = 0; c < h->C; ++c) {
    st[(size_t)c] = std::sin(tilt_deg_host[c] * 0.017453292519943295);
    any |= tilt_deg_host[c] != 0.0;
  }
  if (!any) return 0;
  if (dev_copy(&h->sin_tilt, st.data(), st.size()) || dev_copy(&h->lever, lever_host, (size_t)h->C * h->N)) return BOGP_ERR_CUDA;
  h->tilted = true;
  return 0;
}

extern "C" int bogp_envbatch_bartlett(bogp_envbatch_t h, int atype, int mf_method, float* out_dev, void* stream) {
  BOGP_REQUIRE(h && out_dev, "bogp_envbatch_bartlett: NULL argument");
  BOGP_REQUIRE(h->has_cov, "bogp_envbatch_bartlett: call bogp_envbatch_set_covariance first");
  BOGP_REQUIRE((atype == 0 || atype == 1) && (mf_method == 0 || mf_method == 1), "bogp_envbatch_bartlett: bad atype/mf");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (h->tilted) {
    const size_t per_warp = (((size_t)h->Mmax * (sizeof(double) + sizeof(float) + sizeof(float2)) + (size_t)h->N * sizeof(float2)) + 15) & ~(size_t)15;
    const size_t smem_t = (size_t)bogp::kEnvWarps * per_warp;
    BOGP_REQUIRE(smem_t <= 227 * 1024, "bogp_envbatch_bartlett: Mmax+N too large");
    if (smem_t > 48 * 1024) BOGP_CUDA(cudaFuncSetAttribute(bogp::envbatch_tilt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    const int grid_t = (int)std::max<long long>(1, std::min<long long>((h->C + bogp::kEnvWarps - 1) / bogp::kEnvWarps, (long long)sms * 8));
    bogp::envbatch_tilt_kernel<<<grid_t, bogp::kEnvWarps * 32, smem_t, static_cast<cudaStream_t>(stream)>>>(
        h->C, h->F, h->N, h->Mmax, h->k_re, h->k_im, h->amp_re, h->amp_im, h->phiT, h->range, h->sin_tilt, h->lever,
        h->CkT_re, h->CkT_im, h->J, h->trK, h->Jmax, atype, mf_method, out_dev);
    bogp::count_launch();
    BOGP_CUDA(cudaGetLastError());
    return 0;
  }
  const int force_simt = std::getenv("BOGP_ENV_SIMT") ? std::atoi(std::getenv("BOGP_ENV_SIMT")) : 0;
  if (h->rec && h->Jmax <= bogp::kEnvJ && !force_simt) {
    // streaming kernel: as many warps as fit with `stages` record buffers each.  Measured (profiles/paths_r1b.jsonl):
    // one stage per warp wins -- the kernel is then bound by the warps' arithmetic, not by copies in flight, and
    // twice the warps beat a private second stage (BOGP_ENV_STAGES=2 keeps the comparison runnable)
    const int stages_env = std::getenv("BOGP_ENV_STAGES") ? std::atoi(std::getenv("BOGP_ENV_STAGES")) : 1;
    bogp::EnvStreamParams p{};
    p.stages = std::min(2, std::max(1, stages_env));
    p.warp_bytes = (unsigned)(((size_t)p.stages * h->rec_bytes + (size_t)h->F * h->Mmax * 8 + 127) & ~(size_t)127);
    p.warps = (int)std::min<size_t>(16, (226 * 1024) / ((size_t)p.warp_bytes + 16));
    if (p.warps >= 1) {
      p.C = h->C; p.F = h->F; p.N = h->N; p.Mmax = h->Mmax; p.Jmax = h->Jmax; p.atype = atype; p.mf = mf_method;
      p.rec_bytes = h->rec_bytes; p.off_kre = h->off_kre; p.off_kim = h->off_kim; p.off_g = h->off_g; p.off_r = h->off_r;
      p.rec = h->rec; p.CkT_re = h->CkT_re; p.CkT_im = h->CkT_im; p.J = h->J; p.trK = h->trK; p.out = out_dev;
      const size_t smem = (size_t)p.warps * p.warp_bytes + 16 * (size_t)p.warps;
      BOGP_CUDA(cudaFuncSetAttribute(bogp::envstream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int grid = (int)std::max<long long>(1, std::min<long long>((h->C + p.warps - 1) / p.warps, (long long)sms));
      cudaEvent_t ev = bogp::timing_begin(static_cast<cudaStream_t>(stream));
      bogp::envstream_kernel<<<grid, p.warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
      bogp::timing_end(ev, static_cast<cudaStream_t>(stream));
      bogp::count_launch();
      BOGP_CUDA(cudaGetLastError());
      return 0;
    }
  }
  const size_t smem = (size_t)bogp::kEnvWarps * (h->Mmax + h->N) * sizeof(float2);
  BOGP_REQUIRE(smem <= 227 * 1024, "bogp_envbatch_bartlett: Mmax+N too large");
  if (smem > 48 * 1024) BOGP_CUDA(cudaFuncSetAttribute(bogp::envbatch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::max<long long>(1, std::min<long long>((h->C + bogp::kEnvWarps - 1) / bogp::kEnvWarps, (long long)sms * 8));
  bogp::envbatch_kernel<<<grid, bogp::kEnvWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      h->C, h->F, h->N, h->Mmax, h->k_re, h->k_im, h->amp_re, h->amp_im, h->phiT, h->range, h->CkT_re, h->CkT_im, h->J,
      h->trK, h->Jmax, atype, mf_method, out_dev);
  bogp::count_launch();
  BOGP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bogp_envbatch_destroy(bogp_envbatch_t h) {
  if (!h) return 0;
  cudaFree(h->k_re); cudaFree(h->k_im); cudaFree(h->range); cudaFree(h->amp_re); cudaFree(h->amp_im);
  cudaFree(h->phiT); cudaFree(h->rec); cudaFree(h->sin_tilt); cudaFree(h->lever); cudaFree(h->CkT_re); cudaFree(h->CkT_im); cudaFree(h->J); cudaFree(h->trK);
  delete h;
  return 0;
}
