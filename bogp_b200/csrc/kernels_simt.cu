// CUDA-core (SIMT) Bartlett kernels: scattered candidates, tilted array, small grids.
//
// One warp evaluates one candidate over all tonals.  Vertical array: mode-space form
//   a_m   = phi_m(z) k_m^{-1/2} r^{-1/2} e^{k_im r} e^{-i k_re r}
//   T     = W_f a          (W_f = [R_f ; C_f],  G = Phi^H Phi = R^H R,  Phi^H K Phi = C^H C)
//   B_f   = ||C a||^2 / ||R a||^2  ( = w^H K w with w = p/||p|| ),  cbf_ml = tr K - B_f
// Tilted array (per-receiver range r_n = r - (z_pivot - z_n) sin(tilt)): receiver-space form
//   p_n   = r_n^{-1/2} sum_m Phi[n,m] phi_m(z) k_m^{-1/2} e^{-i k_m r_n},  B_f = ||Ck p||^2 / ||p||^2.
// Replaces runner + beamformer + multifreq combine of MatchedFieldProcessor.__call__
// (ref/projects/swellex96_inv/data/bo/obj.py:28-30, ref/projects/swellex96_localization/data/mfp.py:213-214).
#include <algorithm>
#include <cmath>

#include "bogp_internal.h"
#include "device_math.cuh"

namespace bogp {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

// ------------------------------------------------------------------------------------------------
// Mode-space evaluation shared by the points and grid kernels.  `a_s` holds a_m for this warp.
__device__ __forceinline__ float modespace_tonal(const ModesetDev& ms, const TonalMeta& t, const float2* a_s,
                                                 float2* T_s, int lane, int atype) {
  const float* Wt = ms.Wt + t.offWt;   // [Mp][rowsP]
  const int rowsP = t.rowsP;
  for (int rb = 0; rb < t.rows; rb += 32) {
    const int row = rb + lane;
    float tr = 0.f, ti = 0.f;
    for (int m = 0; m < t.M; ++m) {
      const float w = __ldg(Wt + (size_t)m * rowsP + row);
      const float2 a = a_s[m];
      tr = fmaf(w, a.x, tr);
      ti = fmaf(w, a.y, ti);
    }
    T_s[row] = make_float2(tr, ti);
  }
  __syncwarp();
  float norm = 0.f, num = 0.f;
  for (int i = lane; i < t.n_plain; i += 32) {
    const float2 v = T_s[i];
    norm += v.x * v.x + v.y * v.y;
  }
  const int nb = t.n_plain;
  for (int i = lane; i < t.n_npair; i += 32) {
    const float2 u1 = T_s[nb + i], u2 = T_s[nb + t.n_npair + i];
    const float re = u1.x - u2.y, im = u1.y + u2.x;
    norm += re * re + im * im;
  }
  const int qb = nb + 2 * t.n_npair;
  for (int i = lane; i < t.n_qpair; i += 32) {
    const float2 u1 = T_s[qb + i], u2 = T_s[qb + t.n_qpair + i];
    const float re = u1.x - u2.y, im = u1.y + u2.x;
    num += re * re + im * im;
  }
  norm = warp_sum(norm);
  num = warp_sum(num);
  __syncwarp();
  const float b = num / norm;
  return atype == BOGP_ATYPE_CBF ? b : t.trK - b;
}

// q_m = phi_m(z) * k_m^{-1/2}  (complex), phi by linear interpolation of the depth table.
__device__ __forceinline__ float2 source_coeff(const ModesetDev& ms, const TonalMeta& t, int m, int i0, float w) {
  const float* tr = ms.tab_re + t.offTab + (size_t)i0 * t.Mp + m;
  float pr = fmaf(w, __ldg(tr + t.Mp) - __ldg(tr), __ldg(tr));
  float2 phi = make_float2(pr, 0.f);
  if (ms.src_complex) {
    const float* ti = ms.tab_im + t.offTab + (size_t)i0 * t.Mp + m;
    phi.y = fmaf(w, __ldg(ti + t.Mp) - __ldg(ti), __ldg(ti));
  }
  return cmul(phi, make_float2(__ldg(ms.c_re + t.offM + m), __ldg(ms.c_im + t.offM + m)));
}

// e^{-i k r} with float64 phase reduction, times e^{k_im r}.
__device__ __forceinline__ float2 propagate(const ModesetDev& ms, const TonalMeta& t, int m, double r) {
  const double kre = __ldg(ms.k_re + t.offM + m), kim = __ldg(ms.k_im + t.offM + m);
  float s, c;
  sincosf(reduce_phase(kre * r), &s, &c);
  const float amp = expf((float)(kim * r));
  return make_float2(amp * c, -amp * s);
}

__global__ void __launch_bounds__(kThreads)
points_modespace_kernel(ModesetDev ms, const double* __restrict__ cand, long long C, int atype, int mf,
                        float* __restrict__ out, int aStride, int tStride) {
  extern __shared__ float2 smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* a_s = smem + (size_t)warp * (aStride + tStride);
  float2* T_s = a_s + aStride;
  for (long long c = (long long)blockIdx.x * kWarpsPerBlock + warp; c < C; c += (long long)gridDim.x * kWarpsPerBlock) {
    const double r = cand[3 * c + 0], z = cand[3 * c + 1];
    int i0; float w;
    depth_index(z, ms.z0, ms.dz, ms.nz_tab, i0, w);
    const float rs = (float)rsqrt(r);
    float acc = 0.f;
    for (int f = 0; f < ms.F; ++f) {
      const TonalMeta& t = ms.t[f];
      for (int m = lane; m < t.M; m += 32) {
        float2 a = cmul(source_coeff(ms, t, m, i0, w), propagate(ms, t, m, r));
        a_s[m] = make_float2(a.x * rs, a.y * rs);
      }
      __syncwarp();
      const float bf = modespace_tonal(ms, t, a_s, T_s, lane, atype);
      acc = combine_tonal(acc, bf, mf, f == 0);
    }
    if (lane == 0) out[c] = (mf == BOGP_MF_MEAN) ? acc / (float)ms.F : acc;
  }
}

// Grid: a_m = phiz[j][m] * E[i][m] from tables built once per call in float64.
__global__ void __launch_bounds__(kThreads)
grid_modespace_kernel(ModesetDev ms, const float2* __restrict__ E, const float2* __restrict__ PZ, int Nr, int Nz,
                      int atype, int mf, float* __restrict__ out, int aStride, int tStride) {
  extern __shared__ float2 smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* a_s = smem + (size_t)warp * (aStride + tStride);
  float2* T_s = a_s + aStride;
  const long long C = (long long)Nr * Nz;
  for (long long c = (long long)blockIdx.x * kWarpsPerBlock + warp; c < C; c += (long long)gridDim.x * kWarpsPerBlock) {
    const int j = (int)(c / Nr), i = (int)(c % Nr);
    float acc = 0.f;
    for (int f = 0; f < ms.F; ++f) {
      const TonalMeta& t = ms.t[f];
      const float2* Ef = E + (size_t)t.offM * Nr + (size_t)i * t.Mp;     // [Nr][Mp] per tonal
      const float2* Pf = PZ + (size_t)t.offM * Nz + (size_t)j * t.Mp;    // [Nz][Mp] per tonal
      for (int m = lane; m < t.M; m += 32) a_s[m] = cmul(__ldg(Pf + m), __ldg(Ef + m));
      __syncwarp();
      const float bf = modespace_tonal(ms, t, a_s, T_s, lane, atype);
      acc = combine_tonal(acc, bf, mf, f == 0);
    }
    if (lane == 0) out[c] = (mf == BOGP_MF_MEAN) ? acc / (float)ms.F : acc;
  }
}

// E[f][i][m] = exp(-i k_m r_i) / sqrt(k_m r_i) (float64 arithmetic, float32 result);
// PZ[f][j][m] = phi_m(z_j) (complex if the modes are).
__global__ void grid_tables_kernel(ModesetDev ms, const double* __restrict__ r, int Nr, const double* __restrict__ z,
                                   int Nz, float2* __restrict__ E, float2* __restrict__ PZ) {
  const int f = blockIdx.y;
  const TonalMeta& t = ms.t[f];
  const long long nE = (long long)Nr * t.Mp, nP = (long long)Nz * t.Mp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < nE + nP;
       idx += (long long)gridDim.x * blockDim.x) {
    if (idx < nE) {
      const int i = (int)(idx / t.Mp), m = (int)(idx % t.Mp);
      float2 v = make_float2(0.f, 0.f);
      if (m < t.M) {
        const double kre = ms.k_re[t.offM + m], kim = ms.k_im[t.offM + m], rr = r[i];
        // 1/sqrt(k r) = (k r)^{-1/2}: polar form of the complex number k r
        const double ar = kre * rr, ai = kim * rr;
        const double mag = rsqrt(sqrt(ar * ar + ai * ai));
        const double ang = -0.5 * atan2(ai, ar);
        double s, c, s2, c2;
        sincos(-(kre * rr), &s, &c);      // e^{-i k_re r}
        sincos(ang, &s2, &c2);
        const double amp = exp(kim * rr) * mag;
        v.x = (float)(amp * (c * c2 - s * s2));
        v.y = (float)(amp * (s * c2 + c * s2));
      }
      E[(size_t)t.offM * Nr + idx] = v;
    } else {
      const long long q = idx - nE;
      const int j = (int)(q / t.Mp), m = (int)(q % t.Mp);
      float2 v = make_float2(0.f, 0.f);
      if (m < t.M) {
        int i0; float w;
        depth_index(z[j], ms.z0, ms.dz, ms.nz_tab, i0, w);
        const float* tr = ms.tab_re + t.offTab + (size_t)i0 * t.Mp + m;
        v.x = fmaf(w, tr[t.Mp] - tr[0], tr[0]);
        if (ms.src_complex) {
          const float* ti = ms.tab_im + t.offTab + (size_t)i0 * t.Mp + m;
          v.y = fmaf(w, ti[t.Mp] - ti[0], ti[0]);
        }
      }
      PZ[(size_t)t.offM * Nz + q] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tilted array, receiver space.
__global__ void __launch_bounds__(kThreads)
points_tilt_kernel(ModesetDev ms, const double* __restrict__ cand, long long C, int atype, int mf,
                   float* __restrict__ out, int qStride, int pStride) {
  extern __shared__ float2 smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* q_s = smem + (size_t)warp * (qStride + pStride);
  float2* p_s = q_s + qStride;
  const int N = ms.N;
  for (long long c = (long long)blockIdx.x * kWarpsPerBlock + warp; c < C; c += (long long)gridDim.x * kWarpsPerBlock) {
    const double r = cand[3 * c + 0], z = cand[3 * c + 1], tilt = cand[3 * c + 2];
    const double sin_t = sin(tilt * 0.017453292519943295);
    int i0; float w;
    depth_index(z, ms.z0, ms.dz, ms.nz_tab, i0, w);
    float acc = 0.f;
    for (int f = 0; f < ms.F; ++f) {
      const TonalMeta& t = ms.t[f];
      for (int m = lane; m < t.M; m += 32) q_s[m] = source_coeff(ms, t, m, i0, w);
      __syncwarp();
      float norm = 0.f;
      for (int n = lane; n < N; n += 32) {
        const double rn = r - (double)__ldg(ms.rec_lever + n) * sin_t;
        float pr = 0.f, pi = 0.f;
        for (int m = 0; m < t.M; ++m) {
          float2 h = cmul(q_s[m], propagate(ms, t, m, rn));
          const float ph = __ldg(ms.recT_re + t.offRec + (size_t)m * N + n);
          if (ms.rec_complex) {
            h = cmul(h, make_float2(ph, __ldg(ms.recT_im + t.offRec + (size_t)m * N + n)));
            pr += h.x; pi += h.y;
          } else {
            pr = fmaf(ph, h.x, pr);
            pi = fmaf(ph, h.y, pi);
          }
        }
        const float rs = (float)rsqrt(rn);
        pr *= rs; pi *= rs;
        p_s[n] = make_float2(pr, pi);
        norm += pr * pr + pi * pi;
      }
      __syncwarp();
      float num = 0.f;
      const float* cr = ms.CkT_re + t.offC;   // [N][J]
      const float* ci = ms.CkT_im + t.offC;
      for (int j = lane; j < t.J; j += 32) {
        float sr = 0.f, si = 0.f;
        for (int n = 0; n < N; ++n) {
          const float2 p = p_s[n];
          const float a = __ldg(cr + (size_t)n * t.J + j), b = __ldg(ci + (size_t)n * t.J + j);
          sr += a * p.x - b * p.y;
          si += a * p.y + b * p.x;
        }
        num += sr * sr + si * si;
      }
      norm = warp_sum(norm);
      num = warp_sum(num);
      __syncwarp();
      const float b = num / norm;
      acc = combine_tonal(acc, atype == BOGP_ATYPE_CBF ? b : t.trK - b, mf, f == 0);
    }
    if (lane == 0) out[c] = (mf == BOGP_MF_MEAN) ? acc / (float)ms.F : acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Tilted array on a (range x depth x tilt) grid (config 4).  e^{-i k_m r_n} with r_n = r - l_n sin(tau) factors into
// e^{-i k_m r} (the range table E) and e^{+i k_m l_n sin(tau)}, which depends on the tilt only: one block works on one
// tilt value and keeps Theta_tau[m][n] = Phi[n,m] e^{i k_m l_n sin(tau)} of the current tonal in shared memory, so the
// per-candidate work is N x M complex multiply-adds and no transcendental (the scattered-candidate kernel above pays a
// sincos per (receiver, mode)).  p_n = (r / r_n)^{1/2} sum_m Theta[m][n] phi_m(z) E[m, r]; one warp per candidate,
// two receivers per lane (N <= 64).
constexpr int kTiltChunk = 1024;     // candidates of one tilt plane per block
constexpr int kTiltCand = 4;         // candidates a warp evaluates together

__global__ void __launch_bounds__(kThreads)
grid_tilt_kernel(ModesetDev ms, const float2* __restrict__ E, const float2* __restrict__ PZ, const double* __restrict__ r,
                 const double* __restrict__ tilt, int Nr, int Nz, int atype, int mf, float* __restrict__ out, int MpMax) {
  extern __shared__ float2 smem[];
  float2* Th = smem;                                              // [MpMax][64]
  float* acc = reinterpret_cast<float*>(Th + (size_t)MpMax * 64); // [kTiltChunk]
  float2* a_all = reinterpret_cast<float2*>(acc + kTiltChunk);    // [warps][MpMax]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* a_s = a_all + (size_t)warp * kTiltCand * MpMax;
  const int plane = Nr * Nz;
  const int nchunks = (plane + kTiltChunk - 1) / kTiltChunk;
  const int t = blockIdx.x / nchunks, c0 = (blockIdx.x % nchunks) * kTiltChunk;
  const int cn = min(kTiltChunk, plane - c0);
  const double sin_t = sin(tilt[t] * 0.017453292519943295);
  const int N = ms.N;
  const int n0 = lane, n1 = lane + 32;
  const float l0 = n0 < N ? ms.rec_lever[n0] : 0.f, l1 = n1 < N ? ms.rec_lever[n1] : 0.f;
  const float fs = (float)sin_t;
  for (int f = 0; f < ms.F; ++f) {
    const TonalMeta& tm = ms.t[f];
    __syncthreads();       // previous tonal's Theta no longer in use
    for (int idx = threadIdx.x; idx < tm.Mp * 64; idx += blockDim.x) {
      const int m = idx >> 6, n = idx & 63;
      float2 v = make_float2(0.f, 0.f);
      if (n < N && m < tm.M) {
        const double ls = (double)ms.rec_lever[n] * sin_t;
        const double kre = ms.k_re[tm.offM + m], kim = ms.k_im[tm.offM + m];
        float sn, cs;
        sincosf(reduce_phase(kre * ls), &sn, &cs);
        const float amp = expf((float)(-kim * ls));
        float2 ph = make_float2(ms.rec_re[tm.offRec + (size_t)n * tm.Mp + m], 0.f);
        if (ms.rec_complex) ph.y = ms.rec_im[tm.offRec + (size_t)n * tm.Mp + m];
        v = cmul(ph, make_float2(amp * cs, amp * sn));
      }
      Th[idx] = v;
    }
    __syncthreads();
    const float2* Ef = E + (size_t)tm.offM * Nr;
    const float2* Pf = PZ + (size_t)tm.offM * Nz;
    const float* cr = ms.CkT_re + tm.offC;   // [N][J]
    const float* ci = ms.CkT_im + tm.offC;
    // kTiltCand candidates per warp at a time: every Theta element fetched from shared memory feeds kTiltCand complex
    // multiply-adds (the loop is bound by shared-memory loads otherwise)
    for (int cb = warp * kTiltCand; cb < cn; cb += kWarpsPerBlock * kTiltCand) {
      int ii[kTiltCand];
#pragma unroll
      for (int q = 0; q < kTiltCand; ++q) {
        const int c = c0 + min(cb + q, cn - 1), j = c / Nr, i = c - j * Nr;
        ii[q] = i;
        for (int m = lane; m < tm.M; m += 32)
          a_s[q * MpMax + m] = cmul(__ldg(Pf + (size_t)j * tm.Mp + m), __ldg(Ef + (size_t)i * tm.Mp + m));
      }
      __syncwarp();
      float2 p0[kTiltCand], p1[kTiltCand];
#pragma unroll
      for (int q = 0; q < kTiltCand; ++q) p0[q] = p1[q] = make_float2(0.f, 0.f);
#pragma unroll 2
      for (int m = 0; m < tm.M; ++m) {
        const float2 t0 = Th[m * 64 + n0], t1 = Th[m * 64 + n1];
#pragma unroll
        for (int q = 0; q < kTiltCand; ++q) {
          const float2 a = a_s[q * MpMax + m];
          p0[q].x = fmaf(t0.x, a.x, fmaf(-t0.y, a.y, p0[q].x));
          p0[q].y = fmaf(t0.x, a.y, fmaf(t0.y, a.x, p0[q].y));
          p1[q].x = fmaf(t1.x, a.x, fmaf(-t1.y, a.y, p1[q].x));
          p1[q].y = fmaf(t1.x, a.y, fmaf(t1.y, a.x, p1[q].y));
        }
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < kTiltCand; ++q) {
        const int cl = cb + q;
        if (cl >= cn) break;                       // warp-uniform
        const float rinv = 1.f / (float)r[ii[q]];
        const float s0 = rsqrtf(1.f - l0 * fs * rinv), s1 = rsqrtf(1.f - l1 * fs * rinv);     // (r / r_n)^{1/2}
        float2 u0 = make_float2(p0[q].x * s0, p0[q].y * s0), u1 = make_float2(p1[q].x * s1, p1[q].y * s1);
        const float norm = warp_sum(u0.x * u0.x + u0.y * u0.y + u1.x * u1.x + u1.y * u1.y);
        float num = 0.f;
        for (int jj = 0; jj < tm.J; ++jj) {
          float sr = 0.f, si = 0.f;
          if (n0 < N) {
            const float a = __ldg(cr + (size_t)n0 * tm.J + jj), b = __ldg(ci + (size_t)n0 * tm.J + jj);
            sr += a * u0.x - b * u0.y; si += a * u0.y + b * u0.x;
          }
          if (n1 < N) {
            const float a = __ldg(cr + (size_t)n1 * tm.J + jj), b = __ldg(ci + (size_t)n1 * tm.J + jj);
            sr += a * u1.x - b * u1.y; si += a * u1.y + b * u1.x;
          }
          sr = warp_sum(sr); si = warp_sum(si);
          num += sr * sr + si * si;
        }
        const float b = num / norm;
        if (lane == 0) acc[cl] = combine_tonal(acc[cl], atype == BOGP_ATYPE_CBF ? b : tm.trK - b, mf, f == 0);
      }
    }
  }
  __syncthreads();
  const float scale = mf == BOGP_MF_MEAN ? 1.f / (float)ms.F : 1.f;
  for (int cl = threadIdx.x; cl < cn; cl += blockDim.x) out[(size_t)t * plane + c0 + cl] = acc[cl] * scale;
}

// cand[(t*Nz + j)*Nr + i] = (r_i, z_j, tilt_t)
__global__ void make_grid_candidates_kernel(const double* __restrict__ r, int Nr, const double* __restrict__ z, int Nz,
                                            const double* __restrict__ tilt, int Nt, double* __restrict__ cand) {
  const long long C = (long long)Nr * Nz * Nt;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < C; c += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(c % Nr), j = (int)((c / Nr) % Nz), tt = (int)(c / ((long long)Nr * Nz));
    cand[3 * c + 0] = r[i];
    cand[3 * c + 1] = z[j];
    cand[3 * c + 2] = tilt ? tilt[tt] : 0.0;
  }
}

static int sm_count() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

static int grid_for(long long C, size_t smem_bytes) {
  // persistent-style: a multiple of the SM count, capped by the work
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem_bytes, 1024)));
  long long want = (C + kWarpsPerBlock - 1) / kWarpsPerBlock;
  return (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count() * per_sm));
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) BOGP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  BOGP_REQUIRE(bytes <= 227 * 1024, "mode set too large for the SIMT kernels (%zu B shared memory per block)", bytes);
  return 0;
}

int launch_points(const bogp_modeset_s* h, const double* cand_dev, long long C, int atype, int mf, bool any_tilt,
                  float* out_dev, cudaStream_t s) {
  if (!any_tilt) {
    const int aStride = h->max_Mp, tStride = pad_to(std::max(h->max_rows, 1), 32);
    const size_t smem = (size_t)kWarpsPerBlock * (aStride + tStride) * sizeof(float2);
    if (int rc = set_smem(points_modespace_kernel, smem)) return rc;
    cudaEvent_t ev = timing_begin(s);
    points_modespace_kernel<<<grid_for(C, smem), kThreads, smem, s>>>(h->dev, cand_dev, C, atype, mf, out_dev, aStride, tStride);
    timing_end(ev, s);
    bogp::count_launch();
  } else {
    const int qStride = h->max_Mp, pStride = h->N;
    const size_t smem = (size_t)kWarpsPerBlock * (qStride + pStride) * sizeof(float2);
    if (int rc = set_smem(points_tilt_kernel, smem)) return rc;
    cudaEvent_t ev = timing_begin(s);
    points_tilt_kernel<<<grid_for(C, smem), kThreads, smem, s>>>(h->dev, cand_dev, C, atype, mf, out_dev, qStride, pStride);
    timing_end(ev, s);
    bogp::count_launch();
  }
  BOGP_CUDA(cudaGetLastError());
  return 0;
}

int launch_grid_simt(bogp_modeset_s* h, const double* r_host, int Nr, const double* z_host, int Nz,
                     const double* tilt_host, int Nt, int atype, int mf, float* out_dev, cudaStream_t s) {
  const ModesetDev& d = h->dev;
  int sumMp = 0;
  for (int f = 0; f < d.F; ++f) sumMp += d.t[f].Mp;
  const size_t bytes_rz = ((size_t)Nr + Nz + Nt) * sizeof(double);
  const size_t off_tab = pad_to((int)bytes_rz, 256);
  if (!tilt_host) {
    const size_t bytes_E = (size_t)sumMp * Nr * sizeof(float2), bytes_P = (size_t)sumMp * Nz * sizeof(float2);
    if (int rc = ensure_scratch(h, off_tab + bytes_E + bytes_P)) return rc;
    char* base = static_cast<char*>(h->grid_scratch);
    double* r_dev = reinterpret_cast<double*>(base);
    double* z_dev = r_dev + Nr;
    float2* E = reinterpret_cast<float2*>(base + off_tab);
    float2* PZ = reinterpret_cast<float2*>(base + off_tab + bytes_E);
    BOGP_CUDA(h2d_async(r_dev, r_host, Nr * sizeof(double), s));
    BOGP_CUDA(h2d_async(z_dev, z_host, Nz * sizeof(double), s));
    dim3 tg(std::max(1, std::min(64, (int)(((long long)(Nr + Nz) * h->max_Mp + 255) / 256))), d.F);
    grid_tables_kernel<<<tg, 256, 0, s>>>(d, r_dev, Nr, z_dev, Nz, E, PZ);
    bogp::count_launch();
    BOGP_CUDA(cudaGetLastError());
    const int aStride = h->max_Mp, tStride = pad_to(std::max(h->max_rows, 1), 32);
    const size_t smem = (size_t)kWarpsPerBlock * (aStride + tStride) * sizeof(float2);
    if (int rc = set_smem(grid_modespace_kernel, smem)) return rc;
    cudaEvent_t ev = timing_begin(s);
    grid_modespace_kernel<<<grid_for((long long)Nr * Nz, smem), kThreads, smem, s>>>(d, E, PZ, Nr, Nz, atype, mf, out_dev, aStride, tStride);
    timing_end(ev, s);
    bogp::count_launch();
    BOGP_CUDA(cudaGetLastError());
    return 0;
  }
  if (d.N <= 64) {
    // tilt-plane kernel: range / depth tables as in the vertical-array grid, Theta_tau per block in shared memory
    const size_t bytes_E = (size_t)sumMp * Nr * sizeof(float2), bytes_P = (size_t)sumMp * Nz * sizeof(float2);
    if (int rc = ensure_scratch(h, off_tab + bytes_E + bytes_P)) return rc;
    char* base = static_cast<char*>(h->grid_scratch);
    double* r_dev = reinterpret_cast<double*>(base);
    double* z_dev = r_dev + Nr;
    double* t_dev = z_dev + Nz;
    float2* E = reinterpret_cast<float2*>(base + off_tab);
    float2* PZ = reinterpret_cast<float2*>(base + off_tab + bytes_E);
    BOGP_CUDA(h2d_async(r_dev, r_host, Nr * sizeof(double), s));
    BOGP_CUDA(h2d_async(z_dev, z_host, Nz * sizeof(double), s));
    BOGP_CUDA(h2d_async(t_dev, tilt_host, Nt * sizeof(double), s));
    dim3 tg(std::max(1, std::min(64, (int)(((long long)(Nr + Nz) * h->max_Mp + 255) / 256))), d.F);
    grid_tables_kernel<<<tg, 256, 0, s>>>(d, r_dev, Nr, z_dev, Nz, E, PZ);
    bogp::count_launch();
    BOGP_CUDA(cudaGetLastError());
    const size_t smem = ((size_t)h->max_Mp * 64 + (size_t)kWarpsPerBlock * kTiltCand * h->max_Mp) * sizeof(float2) + kTiltChunk * sizeof(float);
    if (int rc = set_smem(grid_tilt_kernel, smem)) return rc;
    const long long plane = (long long)Nr * Nz;
    const long long blocks = (long long)Nt * ((plane + kTiltChunk - 1) / kTiltChunk);
    BOGP_REQUIRE(blocks < (1ll << 31), "bogp_bartlett_grid: tilt grid too large for one launch");
    cudaEvent_t ev = timing_begin(s);
    grid_tilt_kernel<<<(unsigned)blocks, kThreads, smem, s>>>(d, E, PZ, r_dev, t_dev, Nr, Nz, atype, mf, out_dev, h->max_Mp);
    timing_end(ev, s);
    bogp::count_launch();
    BOGP_CUDA(cudaGetLastError());
    return 0;
  }
  const long long C = (long long)Nr * Nz * Nt;
  if (int rc = ensure_scratch(h, off_tab + (size_t)C * 3 * sizeof(double))) return rc;
  char* base = static_cast<char*>(h->grid_scratch);
  double* r_dev = reinterpret_cast<double*>(base);
  double* z_dev = r_dev + Nr;
  double* t_dev = z_dev + Nz;
  double* cand = reinterpret_cast<double*>(base + off_tab);
  BOGP_CUDA(h2d_async(r_dev, r_host, Nr * sizeof(double), s));
  BOGP_CUDA(h2d_async(z_dev, z_host, Nz * sizeof(double), s));
  BOGP_CUDA(h2d_async(t_dev, tilt_host, Nt * sizeof(double), s));
  make_grid_candidates_kernel<<<std::max(1, (int)std::min<long long>((C + 255) / 256, 148 * 8)), 256, 0, s>>>(r_dev, Nr, z_dev, Nz, t_dev, Nt, cand);
  bogp::count_launch();
  BOGP_CUDA(cudaGetLastError());
  return launch_points(h, cand, C, atype, mf, true, out_dev, s);
}

}  // namespace bogp
