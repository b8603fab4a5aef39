// tcgen05 (UMMA) Bartlett grid kernel for sm_100a: the separable (range x depth) ambiguity surface as a 3xTF32
// tensor-core contraction with TMEM accumulators and a fused Bartlett epilogue.
//
// Math (mode space, see kernels_simt.cu): for tonal f, range r_i and source depth z_j
//     T[row, (i,j)] = sum_m W_f[row, m] phi_m(z_j) E_f[m, i],   E_f[m, i] = e^{-i k_m r_i} / sqrt(k_m r_i)
//     B_f(i, j)     = sum_{numerator rows} |T|^2 / sum_{norm rows} |T|^2
// As a GEMM per (depth j, tonal f):   D[i, row] = sum_m  A[i, m] * Bop[row, m]
//     A   = E_f^T, the 128-range tile of the phase table (re and im parts -> two accumulators), resident in TENSOR
//           MEMORY for the whole depth sweep of a tonal (tcgen05.mma with A in TMEM: with K = 8 per TF32 MMA an A
//           operand in shared memory costs 4 KB of shared-memory reads per instruction and bounds the pipe);
//     Bop = W_f diag(phi(z_j)), streamed per depth (several depths stacked along the MMA N dimension).
// 3xTF32: every fp32 operand x is split x = hi + lo (hi = tf32-rounded); D = A_lo B_hi + A_hi B_lo + A_hi B_hi
// accumulates in fp32 in TMEM (1xTF32 fails the 1e-5 bar, 3xTF32 gives ~3e-7).
// D puts range i on TMEM lane i, so each epilogue thread owns one range and reduces over the operator rows of its
// own lane: no cross-lane shuffles; the replica field never exists outside TMEM.
//
// Pipeline per CTA (persistent over (range tile, depth chunk) units), 24 warps, roles aligned to warpgroups so that
// setmaxnreg moves registers from the data-movement roles to the epilogue (kEpiRegs / kAuxRegs):
//     warps 0-15   epilogue: 2 buffer sets x 2 depth parities x 4 lane quadrants; tcgen05.ld -> sum of squares ->
//                  Bartlett -> multi-frequency combine in shared memory
//     warps 16-19  E loaders: global -> registers -> tcgen05.st, one tonal ahead of the MMAs
//     warp 20      bulk-copy producer: the ring of Bop stages (cp.async.bulk + mbarrier complete_tx)
//     warps 21, 22 MMA issuers, one per accumulator buffer (stage parity): tcgen05.mma.kind::tf32 with A from TMEM and
//                  B from shared memory, tcgen05.commit to the ring / TMEM barriers; warp 21 owns the TMEM allocation
//     warp 23      idle (fills the last warpgroup)
// The Bop table is built per call directly in the UMMA canonical K-major no-swizzle layout ([8-row group][16-byte K
// chunk][8 rows][4 floats]; LBO = 128 B, SBO = 32 * Mp B), so the producer issues plain 1-D bulk copies.
// TMEM map (512 columns): two accumulator buffers of (re | im) x NA columns, then the A region.
// The two issuers take turns (BAR_TURN): stages enter the tensor pipe in order, so that a stage's epilogue overlaps the
// next stage's MMAs instead of both buffers sitting in the epilogue together.  The arg-max / arg-min of the surface is
// folded into the write-out (packed atomicMax, last CTA decodes): bogp_bartlett_grid_argext.
//
// Template parameter TILT = true is the tilted-array engine (config 4): same pipeline, B operand = Theta_tau =
// Phi exp(i k l sin tau) per (tilt, receiver half), A operand = E_f phi(z_j) rebuilt in TMEM per (unit, tonal), epilogue
// in receiver space with s_n = (r / r_n)^1/2 -- see the comments at `if (TILT)` in the loader and epilogue roles.
//
// Replaces the loop `for z in src_z: amb[zz,:] = MFP({"src_z": z, "rec_r": rvec})`
// (ref/projects/swellex96_localization/data/mfp.py:213-214) for a vertical array, and the same sweep with a `tilt`
// key (ref/projects/swellex96_loc_tilt/data/mfp.py:196-230, ref/projects/swellex96_loc_tilt/data/env.py:73-74).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include <cuda_fp16.h>

#include "bogp_internal.h"
#include "device_math.cuh"
#include "umma_ptx.cuh"

// Per-role wait counters (tools/umma_dbg_counters.py) are compiled in only with -DBOGP_UMMA_DBG=1
// (BOGP_UMMA_DBG_BUILD=1 python -m bogp_b200.build): their 64-bit accumulators cost registers in every role.
#ifndef BOGP_UMMA_DBG
#define BOGP_UMMA_DBG 0
#endif

namespace bogp {

constexpr int kTileR = 128;            // ranges per tile = MMA M = TMEM lanes
constexpr int kStages = 6;             // Bop ring depth
#ifndef BOGP_EPI_REGS
#define BOGP_EPI_REGS 96
#define BOGP_LOAD_REGS 56
#define BOGP_CTL_REGS 40
#endif
constexpr int kEpiSub = 2;             // epilogue warps per (buffer set, lane quadrant): depth j goes to warp j % kEpiSub
constexpr int kEpiWarps = 8 * kEpiSub;  // 2 accumulator-buffer sets x kEpiSub depth residues x 4 TMEM lane quadrants
constexpr int kLoadWarps = 4;
// Warp roles, laid out by warpgroup (4 aligned warps) so that setmaxnreg can move registers from the data-movement
// roles to the epilogue: warps 0..15 epilogue, 16..19 A-operand loaders, 20 Bop producer, 21 / 22 MMA issuers, 23 idle.
constexpr int kLoadWarp0 = kEpiWarps;
constexpr int kProdWarp = kEpiWarps + kLoadWarps;
constexpr int kMmaWarpA = kProdWarp + 1;
constexpr int kMmaWarpB = kProdWarp + 2;
constexpr int kUmmaThreads = (kProdWarp + 4) * 32;
constexpr int kEpiRegs = BOGP_EPI_REGS, kLoadRegs = BOGP_LOAD_REGS, kCtlRegs = BOGP_CTL_REGS;
static_assert(16 * kEpiRegs + 4 * kLoadRegs + 4 * kCtlRegs <= 24 * 80, "register pool of the CTA (24 warps launched at 80)");
constexpr unsigned kSmemLimit = 227u * 1024u;

struct UTonal {
  int Mp, rowsP;
  int kc;                      // TMEM columns of one split of the A image: Mp (TF32, one mode per column) or Kp / 2 (FP16 pairs, Kp = Mp padded to 16)
  float sA, sB;                // FP16 mode: power-of-two scales of the A and B operands (largest magnitude -> 2^14)
  int nq;                      // numerator (re, im) row pairs = rank of Phi^H K Phi: columns [0, 2 nq)
  int n1;                      // 2 nq + 2 n_npair: pair columns in [2 nq, n1) feed the norm; plain columns from n1 on
  int nj;                      // depths stacked per stage
  int wait_prev;               // this tonal's E tile overlaps the previous tonal's in shared memory
  unsigned e_bytes;            // 2048 * Mp  (re_hi | re_lo | im_hi | im_lo) x 128 ranges
  unsigned e_off;              // offset of the tonal inside one range tile's block of the E table
  unsigned b_zbytes;           // rowsP * Mp * 4: one depth of one split of Bop
  unsigned long long b_off;    // offset of the tonal inside the Bop tables
  float trK;
  int src;                     // index of the tonal in the mode set (the plan may process the tonals in another order)
  int fold;                    // folded operator: numerator = |c1 y0 + c2 y1|^2 from the first two columns, norm = all columns
  float c1r, c1i, c2r, c2i;
  // tilted-array mode only
  int offM;                    // column offset of the tonal in the phi(z) table
  int offC;                    // offset of the receiver-space covariance factor [N][J]
  int J;
};

struct UParams {
  int F, Nr, Nz, nrt, ZC, units, atype, mf;
  unsigned e_rt_stride, stage_bytes;
  unsigned smem_B, smem_acc, smem_bar, smem_aux;     // smem_aux (tilt mode): lever[64] | covariance factor [F][64][4] float2 | partial-sum exchange
  int NA;                      // TMEM columns per accumulator half; buffers at 0 and 2 NA, A region at [4 NA, 4 NA + AR)
  int AR;                      // TMEM columns of the A region
  const unsigned char* E;
  const unsigned char* Bhi;
  const unsigned char* Blo;
  float* out;
  int* err;
  long long* dbg;              // optional [grid][16] cycle counters (BOGP_UMMA_DBG=1): where each role waits
  // fused arg-ext (collate.py:50): the epilogue folds every value it writes into a packed (ordered value, ~index) key
  // with one atomicMax per warp and unit; the last CTA to finish decodes it.  ext_mode: 0 off, 1 arg-max, 2 arg-min.
  int ext_mode;
  int two_slot;                // epilogue: both depth slots of a small tonal loaded at once (A/B knob BOGP_UMMA_TWO_SLOT)
  unsigned long long* ext_key;
  unsigned int* ext_done;
  float* ext_val;
  long long* ext_idx;
  // ---- tilted-array mode (umma_grid_kernel<true>): Nz counts PSEUDO-depths k = (tilt, receiver half) in stage order
  // (k = 4q + {0,1,2,3} -> (tilt 2q, half 0), (2q+1, 0), (2q, 1), (2q+1, 1)), ZC = 2 * tilts per unit, and a unit is
  // (range tile, real depth, tilt chunk)
  int ntc;                     // tilt chunks per real depth
  int Nzr, Nt;                 // real depths, real tilts
  int N, sumMp;
  const float* phiz;           // [Nzr][sumMp]  phi_m(z_j)
  const float* phiz_im;        // imaginary part for complex (KRAKENC) source tables, else null
  const float* sin_t;          // [Nz / 2]      sin(tilt), padded
  const float* rinv;           // [Nr]          1 / r
  const float* lever;          // [64]          z_pivot - z_n (0 beyond N)
  const float* CkT_re;         // receiver-space covariance factor (ModesetDev::CkT_re / CkT_im)
  const float* CkT_im;
  int f16;                     // tilt engine: FP16-split operands (kind::f16) instead of 3xTF32
  int skip;                    // debug build only: 1 = issue no MMAs, 2 = epilogue releases without reading (bottleneck bracketing)
  UTonal t[BOGP_MAX_TONALS];
};

// ------------------------------------------------------------------------------------------------ table kernels
__device__ __forceinline__ float tf32_rna(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// offset (floats) of element (row, m) inside a K-major no-swizzle block of leading dimension Mp
__device__ __forceinline__ size_t canon_off(int row, int m, int Mp) {
  return ((size_t)((row >> 3) * (Mp >> 2) + (m >> 2)) * 8 + (row & 7)) * 4 + (m & 3);
}

// ---- operand tables, one launch: blocks [0, nE) build the E table, blocks [nE, nE + nB) the Bop tables.
// E table: per (range tile, tonal) the TMEM image of E_f^T: columns [re_hi (Mp) | re_lo | im_hi | im_lo] for each of the
// 128 ranges, stored as [16-byte column chunk][range][4 floats] so that a loader warp reads coalesced float4s.
// E[m, i] = exp(-i k_m r_i) / sqrt(k_m r_i) = r^-1/2 e^{k_im r} * k^-1/2 * (cos(k_re r) - i sin(k_re r)), in float64
// (k r reaches 1.6e4 rad), then split x = hi + lo with hi rounded to TF32.
// Bop tables: per (tonal, depth) the block W_f diag(phi(z_j)) [rowsP, Mp] in the UMMA canonical K-major no-swizzle
// layout, hi and lo arrays.  One block = kTabDepths depths of one tonal; the block's share of W_f stays in registers.
constexpr int kTabThreads = 128;
constexpr int kTabDepths = 8;
constexpr int kTabRows = 32;          // ranges per E block
constexpr int kTabMaxQ = 8;           // float4 items of W_f per thread: rowsP * Mp / 4 <= 1024

__global__ void __launch_bounds__(kTabThreads) umma_tables_kernel(ModesetDev ms, UParams p, const double* __restrict__ r,
                                                                  const double* __restrict__ z, int nE) {
  __shared__ double2 kinv[128];       // k_m^-1/2
  __shared__ float phi_s[kTabDepths][64];     // phi_m(z_j) of the block's depths
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && tid == 0) { *p.err = 0; if (p.ext_mode) { *p.ext_key = 0ull; *p.ext_done = 0u; } }
  if ((int)blockIdx.x < nE) {
    // ------------------------------------------------------------------ E table
    const int per_f = p.nrt * (kTileR / kTabRows);
    const int f = blockIdx.x / per_f, rem = blockIdx.x % per_f;
    const int rt = rem / (kTileR / kTabRows), row0 = (rem % (kTileR / kTabRows)) * kTabRows;
    const UTonal& u = p.t[f];
    const TonalMeta& t = ms.t[u.src];
    for (int m = tid; m < u.Mp; m += kTabThreads) {
      double cr = 0.0, ci = 0.0;
      if (m < t.M) {
        const double kre = ms.k_re[t.offM + m], kim = ms.k_im[t.offM + m];
        const double mag = rsqrt(sqrt(kre * kre + kim * kim)), ang = -0.5 * atan2(kim, kre);
        sincos(ang, &ci, &cr);
        cr *= mag; ci *= mag;
      }
      kinv[m] = make_double2(cr, ci);
    }
    __syncthreads();
    float* base = reinterpret_cast<float*>(const_cast<unsigned char*>(p.E) + (size_t)rt * p.e_rt_stride + u.e_off);
    for (int idx = tid; idx < kTabRows * u.Mp; idx += kTabThreads) {
      const int m = idx % u.Mp, row = row0 + idx / u.Mp, i = rt * kTileR + row;
      double vr = 0.0, vi = 0.0;
      if (i < p.Nr && m < t.M) {
        const double kre = ms.k_re[t.offM + m], kim = ms.k_im[t.offM + m], rr = r[i];
        double s, c;
        sincos(kre * rr, &s, &c);
        const double amp = exp(kim * rr) * rsqrt(rr);
        const double2 q = kinv[m];
        vr = amp * (q.x * c + q.y * s);          // (q.x + i q.y) (c - i s)
        vi = amp * (q.y * c - q.x * s);
      }
      const float rh = tf32_rna((float)vr), ih = tf32_rna((float)vi);
      // the low parts are rounded to TF32 here: the tensor core would otherwise TRUNCATE them (biased towards zero,
      // twice the error, and the bias adds up coherently over the modes)
      const float v[4] = {rh, tf32_rna((float)(vr - (double)rh)), ih, tf32_rna((float)(vi - (double)ih))};
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int col = b * u.Mp + m;
        base[((size_t)(col >> 2) * kTileR + row) * 4 + (col & 3)] = v[b];
      }
    }
    return;
  }
  // -------------------------------------------------------------------- Bop tables
  const int nzb = (p.Nz + kTabDepths - 1) / kTabDepths;
  const int bb = blockIdx.x - nE;
  const int f = bb / nzb, j0 = (bb % nzb) * kTabDepths;
  const UTonal& u = p.t[f];
  const TonalMeta& t = ms.t[u.src];
  const int KC = u.Mp >> 2, n_q = u.rowsP * KC;
  // items q = tid, tid + 128, ...: q = (row group * KC + K chunk) * 8 + row8 is the float4 index inside a depth block
  float4 wq[kTabMaxQ];
  int kc4[kTabMaxQ];
#pragma unroll
  for (int i = 0; i < kTabMaxQ; ++i) {
    const int q = min(tid + i * kTabThreads, n_q - 1);
    const int pc = q >> 3, kc = pc % KC, rg = pc / KC;
    kc4[i] = kc * 4;
    wq[i] = __ldg(reinterpret_cast<const float4*>(ms.Wu + t.offWu + (size_t)(rg * 8 + (q & 7)) * t.Mp) + kc);
  }
  float4* Bhi = reinterpret_cast<float4*>(const_cast<unsigned char*>(p.Bhi) + u.b_off);
  float4* Blo = reinterpret_cast<float4*>(const_cast<unsigned char*>(p.Blo) + u.b_off);
  const int nd = min(kTabDepths, p.Nz - j0);
  // all depths' mode shapes first (one barrier), then the stores stream without synchronisation
  for (int e = tid; e < nd * u.Mp; e += kTabThreads) {
    const int dj = e / u.Mp, m = e - dj * u.Mp;
    float v = 0.f;
    if (m < t.M) {
      int i0; float w;
      depth_index(z[j0 + dj], ms.z0, ms.dz, ms.nz_tab, i0, w);
      const float* tr = ms.tab_re + t.offTab + (size_t)i0 * t.Mp + m;
      v = fmaf(w, __ldg(tr + t.Mp) - __ldg(tr), __ldg(tr));
    }
    phi_s[dj][m] = v;
  }
  __syncthreads();
  for (int dj = 0; dj < nd; ++dj) {
    const int j = j0 + dj;
    const float* ph = phi_s[dj];
    float4* hi = Bhi + (size_t)j * n_q;
    float4* lo = Blo + (size_t)j * n_q;
#pragma unroll
    for (int i = 0; i < kTabMaxQ; ++i) {
      const int q = tid + i * kTabThreads;
      if (q < n_q) {
        const float4 pv = *reinterpret_cast<const float4*>(ph + kc4[i]);
        const float x0 = wq[i].x * pv.x, x1 = wq[i].y * pv.y, x2 = wq[i].z * pv.z, x3 = wq[i].w * pv.w;
        const float h0 = tf32_rna(x0), h1 = tf32_rna(x1), h2 = tf32_rna(x2), h3 = tf32_rna(x3);
        hi[q] = make_float4(h0, h1, h2, h3);
        lo[q] = make_float4(tf32_rna(x0 - h0), tf32_rna(x1 - h1), tf32_rna(x2 - h2), tf32_rna(x3 - h3));
      }
    }
  }
}

// ---- tilted-array operand tables
// Theta tables: per (tonal, pseudo-depth k = (tilt, receiver half) in stage order) the block [64 rows, Mp]; receiver
// pair (q, q+1) occupies rows 2q .. 2q+3 = (Re q, Re q+1, Im q, Im q+1) of Theta_tau[n, m] = Phi[n, m] exp(i k_m l_n sin tau), n = 32 half + q, in the UMMA canonical
// K-major layout, hi and lo arrays.  The phase k_m l_n sin(tau) is formed in float64.
__global__ void __launch_bounds__(128) umma_tilt_theta_kernel(ModesetDev ms, UParams p, const double* __restrict__ tilt_deg) {
  const int f = blockIdx.y, k = blockIdx.x;
  const UTonal& u = p.t[f];
  const TonalMeta& t = ms.t[u.src];
  const int tau = min(2 * (k >> 2) + (k & 1), p.Nt - 1), half = (k >> 1) & 1;
  const double st = sin(tilt_deg[tau] * 0.017453292519943295);
  float* Bhi = reinterpret_cast<float*>(const_cast<unsigned char*>(p.Bhi) + u.b_off + (size_t)k * u.b_zbytes);
  float* Blo = reinterpret_cast<float*>(const_cast<unsigned char*>(p.Blo) + u.b_off + (size_t)k * u.b_zbytes);
  const int Mcols = p.f16 ? 2 * u.kc : u.Mp;          // modes per row of the tile (padded)
  for (int idx = threadIdx.x; idx < 32 * Mcols; idx += blockDim.x) {
    const int q = idx / Mcols, m = idx - q * Mcols, n = 32 * half + q;
    double vr = 0.0, vi = 0.0;
    if (n < ms.N && m < t.M) {
      const double ls = (double)ms.rec_lever[n] * st;
      double sn, cs;
      sincos(ms.k_re[t.offM + m] * ls, &sn, &cs);
      const double amp = exp(-ms.k_im[t.offM + m] * ls);
      const double pr = ms.rec_re[t.offRec + (size_t)n * t.Mp + m];
      const double pi = ms.rec_complex ? (double)ms.rec_im[t.offRec + (size_t)n * t.Mp + m] : 0.0;
      vr = amp * (pr * cs - pi * sn);
      vi = amp * (pr * sn + pi * cs);
    }
    // rows of a receiver pair (q even, q + 1): Re q, Re q+1, Im q, Im q+1 -- the epilogue then forms T of both
    // receivers with packed fp32x2 instructions on adjacent accumulator columns
    const int row_re = 4 * (q >> 1) + (q & 1);
    if (p.f16) {
      // FP16-split operands: [64 rows x Kp] halves, 8 halves per 16-byte K chunk, scaled by sB
      const int Kp = 2 * u.kc;
      __half* Hhi = reinterpret_cast<__half*>(Bhi);
      __half* Hlo = reinterpret_cast<__half*>(Blo);
      auto hoff = [&](int row) { return ((size_t)((row >> 3) * (Kp >> 3) + (m >> 3)) * 8 + (row & 7)) * 8 + (m & 7); };
      const float xr = (float)(vr * (double)u.sB), xi = (float)(vi * (double)u.sB);
      const __half hr = __float2half_rn(xr), hi_ = __float2half_rn(xi);
      Hhi[hoff(row_re)] = hr; Hlo[hoff(row_re)] = __float2half_rn(xr - __half2float(hr));
      Hhi[hoff(row_re + 2)] = hi_; Hlo[hoff(row_re + 2)] = __float2half_rn(xi - __half2float(hi_));
      continue;
    }
    const float rh = tf32_rna((float)vr), ih = tf32_rna((float)vi);
    const size_t o0 = canon_off(row_re, m, u.Mp), o1 = canon_off(row_re + 2, m, u.Mp);
    Bhi[o0] = rh; Blo[o0] = tf32_rna((float)(vr - (double)rh));
    Bhi[o1] = ih; Blo[o1] = tf32_rna((float)(vi - (double)ih));
  }
}

// phi_m(z_j) for every real depth (source-depth table interpolation of the data contract), sin(tilt), 1 / r, levers
__global__ void umma_tilt_aux_kernel(ModesetDev ms, UParams p, const double* __restrict__ r, const double* __restrict__ z,
                                     const double* __restrict__ tilt_deg, float* phiz, float* phiz_im, float* sin_t,
                                     float* rinv, float* lever) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int idx = tid; idx < p.Nzr * p.sumMp; idx += nth) {
    const int j = idx / p.sumMp, col = idx - j * p.sumMp;
    int f = 0;
    while (f + 1 < p.F && col >= p.t[f + 1].offM) ++f;
    const int m = col - p.t[f].offM;
    const TonalMeta& t = ms.t[p.t[f].src];
    float v = 0.f;
    if (m < t.M) {
      int i0; float w;
      depth_index(z[j], ms.z0, ms.dz, ms.nz_tab, i0, w);
      const float* tr = ms.tab_re + t.offTab + (size_t)i0 * t.Mp + m;
      v = fmaf(w, __ldg(tr + t.Mp) - __ldg(tr), __ldg(tr));
      if (phiz_im) {
        const float* ti = ms.tab_im + t.offTab + (size_t)i0 * t.Mp + m;
        phiz_im[idx] = fmaf(w, __ldg(ti + t.Mp) - __ldg(ti), __ldg(ti));
      }
    } else if (phiz_im) {
      phiz_im[idx] = 0.f;
    }
    phiz[idx] = v;
  }
  for (int i = tid; i < p.Nz / 2; i += nth) sin_t[i] = (float)sin(tilt_deg[min(i, p.Nt - 1)] * 0.017453292519943295);
  for (int i = tid; i < p.Nr; i += nth) rinv[i] = (float)(1.0 / r[i]);
  for (int i = tid; i < 64; i += nth) lever[i] = i < ms.N ? ms.rec_lever[i] : 0.f;
}

// ------------------------------------------------------------------------------------------------ main kernel
// Debug build only: per-stage time stamps of CTA 0 in three windows of stages (one per tonal class), see launch_grid_umma.
__device__ __forceinline__ void dbg_stamp(const UParams& p, int ev, uint32_t sc) {
#if BOGP_UMMA_DBG
  if (p.dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
    int idx = -1;
    if (sc >= 20 && sc < 40) idx = (int)sc - 20;
    else if (sc >= 100 && sc < 120) idx = 20 + (int)sc - 100;
    else if (sc >= 280 && sc < 300) idx = 40 + (int)sc - 280;
    if (idx >= 0) p.dbg[3360 + ev * 60 + idx] = clock64();
  }
#endif
}
// barrier slots (8 bytes each) at smem_bar
enum { BAR_BFULL = 0, BAR_BEMPTY = kStages, BAR_EFULL = 2 * kStages, BAR_EDONE = 2 * kStages + 2,
       BAR_TFULL = 2 * kStages + 4, BAR_TEMPTY = 2 * kStages + 6, BAR_TURN = 2 * kStages + 8, BAR_COUNT = 2 * kStages + 10 };

// TMEM column of tonal g's E image: even tonals grow from the low end of the A region, odd ones sit at the high end,
// so tonal g+1 can be loaded while tonal g is still being multiplied whenever the two fit side by side.
__device__ __forceinline__ uint32_t a_column(const UParams& p, int g, int Mp) {
  return (uint32_t)(4 * p.NA) + ((g & 1) ? (uint32_t)(p.AR - 4 * Mp) : 0u);
}

// unit -> (range tile, first [pseudo-]depth of the chunk, depths in the chunk, real depth (tilt mode))
template <bool TILT>
__device__ __forceinline__ void unit_decode(const UParams& p, int unit, int& rt, int& z0, int& nzu, int& jreal) {
  rt = unit % p.nrt;
  const int zc = unit / p.nrt;
  if (TILT) { jreal = zc / p.ntc; z0 = (zc % p.ntc) * p.ZC; }
  else { jreal = 0; z0 = zc * p.ZC; }
  nzu = min(p.ZC, p.Nz - z0);
}

struct MmaState {
  const UParams* p;
  uint32_t tmem_base, s_base, bar0;
  uint32_t sc;       // global stage counter
  int g;             // global tonal counter
  int nzu;
  int id;            // issuer 0 / 1: owns accumulator buffer `id` and the stages of that parity
  bool dbg;
  long long w_efull, w_bfull, w_tempty;
};

// The stages of one tonal of one unit that belong to issuer st.id (stage parity = accumulator buffer = issuer: two
// issuer warps alternate stages so that one warp's barrier waits and address set-up overlap the other's MMAs -- a
// single issuer blocks on the MMA queue and then leaves the pipe idle while it prepares the next stage).  Templated on the K steps so that a stage is straight-line code: two barrier
// waits, 6 * KS tcgen05.mma, two commits.  (The per-stage scalar code runs on a single warp at ~5 cycles per dependent
// instruction, while a 128 x 64 x 8 MMA lasts 32 cycles: anything not hoisted out of the stage loop idles the pipe.)
template <int KS, bool F16 = false>
__device__ __forceinline__ void mma_tonal(MmaState& st, const UTonal& u) {
  const UParams& p = *st.p;
  // Mp = TMEM columns per split = 8 per K step in both modes (TF32: 8 modes, FP16: 16 modes as 8 packed pairs), and the
  // B tile's 8-row-group stride is 256 bytes per K step in both (32 Mp bytes of floats = 16 Kp bytes of halves)
  const uint32_t Mp = 8u * KS;
  const uint32_t a_col = st.tmem_base + a_column(p, st.g, (int)Mp);
  const uint32_t desc_hi = (uint32_t)(make_desc(0, 128u, 32u * Mp) >> 32);
  const uint32_t desc_lo0 = (uint32_t)make_desc(0, 128u, 32u * Mp);       // LBO field; start address added per stage
  const uint32_t idesc0 = F16 ? ((1u << 4) | ((kTileR >> 4) << 24)) : ((1u << 4) | (2u << 7) | (2u << 10) | ((kTileR >> 4) << 24));
  auto bar = [&](int i) { return st.bar0 + 8u * (uint32_t)i; };
  mbar_wait_t(bar(BAR_EFULL + (st.g & 1)), (uint32_t)((st.g >> 1) & 1), p.err, 4, (BOGP_UMMA_DBG && st.dbg) ? &st.w_efull : nullptr);
  if (BOGP_UMMA_DBG && st.dbg && st.id == 0 && blockIdx.x == 0 && st.g < 32 && (threadIdx.x & 31) == 0) p.dbg[16 * 200 + st.g] = clock64();
  const int nj = u.nj, nzu = st.nzu;
  uint32_t sc = st.sc;
#pragma unroll 1
  for (int j0 = 0; j0 < nzu; j0 += nj, ++sc) {
    if ((int)(sc & 1u) != st.id) continue;     // the other issuer's stage
    const int njs = min(nj, nzu - j0);
    const uint32_t slot = sc % kStages, ph = (sc / kStages) & 1u;
    const uint32_t buf = sc & 1u, tph = (sc >> 1) & 1u;
    mbar_wait_t(bar(BAR_BFULL + slot), ph, p.err, 5, (BOGP_UMMA_DBG && st.dbg) ? &st.w_bfull : nullptr);
    mbar_wait_t(bar(BAR_TEMPTY + buf), tph ^ 1u, p.err, 6, (BOGP_UMMA_DBG && st.dbg) ? &st.w_tempty : nullptr);
    // Issue in stage order: wait until the other issuer has ISSUED stage sc - 1.  Without the turn the two issuers
    // interleave their MMAs in the tensor pipe's queue, both stages complete together, both buffers sit in the
    // epilogue together and the pipe idles meanwhile (measured lock-step: both issuers ready within ~150 cycles of
    // each other, then nothing for ~2300); in order, stage sc - 1 completes one stage time earlier and its epilogue
    // overlaps this stage's MMAs.
    tc_fence_after();
    dbg_stamp(p, 0, sc);
    const uint32_t N = (uint32_t)(njs * u.rowsP);
    const uint32_t idesc = idesc0 | ((N >> 3) << 17);
    const uint32_t b_hi = st.s_base + p.smem_B + slot * p.stage_bytes;
    const uint32_t b_lo = b_hi + (uint32_t)njs * u.b_zbytes;
    const uint32_t d_re = st.tmem_base + buf * (uint32_t)(2 * p.NA), d_im = d_re + (uint32_t)p.NA;
    // (the turn is taken as late and passed on as early as possible: only the MMA instructions themselves are
    // serialised between the issuers, ~20 cycles each; barrier waits, address set-up and commits overlap)
    if (sc > 0) mbar_wait(bar(BAR_TURN + (st.id ^ 1)), ((sc - 1u) >> 1) & 1u, p.err, 8);
    if (elect_one()) {
      if (!(BOGP_UMMA_DBG && p.skip == 1))
        issue_stage<KS, F16>(d_re, d_im, a_col, Mp, desc_lo0 | ((b_hi >> 4) & 0x3fffu), desc_lo0 | ((b_lo >> 4) & 0x3fffu), desc_hi, idesc);
      mbar_arrive(bar(BAR_TURN + st.id));
      tc_commit(bar(BAR_BEMPTY + slot));
      tc_commit(bar(BAR_TFULL + buf));
    }
    __syncwarp();
    dbg_stamp(p, 1, sc);
  }
  st.sc = sc;
  if (elect_one()) tc_commit(bar(BAR_EDONE + (st.g & 1)));
  __syncwarp();
}

template <bool TILT, bool F16 = false>
__global__ void __launch_bounds__(kUmmaThreads, 1) umma_grid_kernel(const __grid_constant__ UParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + p.smem_bar;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.smem_bar + BAR_COUNT * 8);
  // acc_seq[set][depth parity][quadrant]: tonals (global count) whose multi-frequency updates that epilogue warp has finished
  volatile int* acc_seq = reinterpret_cast<volatile int*>(smem + p.smem_bar + 192);
  static_assert(BAR_COUNT * 8 + 4 <= 192, "barriers | TMEM slot | acc_seq[16] share the 256-byte control block");
  if (threadIdx.x < 16) acc_seq[threadIdx.x] = 0;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar(BAR_BFULL + s), 1); mbar_init(bar(BAR_BEMPTY + s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(BAR_EFULL + b), kLoadWarps);
      mbar_init(bar(BAR_EDONE + b), 2);          // both MMA issuers commit
      mbar_init(bar(BAR_TFULL + b), 1);
      mbar_init(bar(BAR_TURN + b), 1);
      mbar_init(bar(BAR_TEMPTY + b), kEpiWarps / 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarpA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (TILT) {
    // tilt mode: the epilogue's per-receiver constants in shared memory (broadcast LDS instead of dependent __ldg)
    float* lev_s = reinterpret_cast<float*>(smem + p.smem_aux);
    float2* c_s = reinterpret_cast<float2*>(lev_s + 64);              // [F][64][4]
    for (int i = threadIdx.x; i < 64; i += blockDim.x) lev_s[i] = p.lever[i];
    float4* cp_s = reinterpret_cast<float4*>(lev_s + 64 + p.F * 512 + 2 * 2 * 4 * 9 * 32);      // [F][32 receiver pairs]: (cx0, cx1, cy0, cy1), rank 1
    for (int i = threadIdx.x; i < p.F * 32; i += blockDim.x) {
      const int f = i >> 5, n0 = 2 * (i & 31);
      const UTonal& u = p.t[f];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 < p.N) { v.x = p.CkT_re[u.offC + n0 * u.J]; v.z = p.CkT_im[u.offC + n0 * u.J]; }
      if (n0 + 1 < p.N) { v.y = p.CkT_re[u.offC + (n0 + 1) * u.J]; v.w = p.CkT_im[u.offC + (n0 + 1) * u.J]; }
      cp_s[i] = v;
    }
    for (int i = threadIdx.x; i < p.F * 256; i += blockDim.x) {
      const int f = i >> 8, n = (i >> 2) & 63, j = i & 3;
      const UTonal& u = p.t[f];
      c_s[i] = (n < p.N && j < u.J) ? make_float2(p.CkT_re[u.offC + n * u.J + j], p.CkT_im[u.offC + n * u.J + j]) : make_float2(0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // broadcast through a shuffle so that ptxas treats the TMEM base (and everything derived from it: accumulator and
  // A-operand addresses of every MMA) as warp-uniform and keeps it in uniform registers instead of R2UR-ing per stage
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  long long* const dbg = (BOGP_UMMA_DBG && p.dbg) ? p.dbg + 16 * blockIdx.x : nullptr;
  const long long t_start = clock64();

  if (warp >= kProdWarp) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtlRegs));
  if (warp == kProdWarp) {
    // ============================================================== Bop producer (whole warp, one elected lane issues)
    long long w_bempty = 0;
    uint32_t sc = 0;       // global stage counter
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      int rt_, z0, nzu, jr_;
      unit_decode<TILT>(p, unit, rt_, z0, nzu, jr_);
      for (int f = 0; f < p.F; ++f) {
        const UTonal& u = p.t[f];
        for (int j0 = 0; j0 < nzu; j0 += u.nj, ++sc) {
          const int njs = min(u.nj, nzu - j0);
          const uint32_t slot = sc % kStages, ph = (sc / kStages) & 1u;
          mbar_wait_relaxed<TILT ? 512 : 256>(bar(BAR_BEMPTY + slot), ph ^ 1u, p.err, 3);
          const uint32_t bytes = (uint32_t)njs * u.b_zbytes;
          const uint32_t dst = s_base + p.smem_B + slot * p.stage_bytes;
          const size_t src = (size_t)u.b_off + (size_t)(z0 + j0) * u.b_zbytes;
          if (elect_one()) {
            mbar_expect_tx(bar(BAR_BFULL + slot), 2 * bytes);
            bulk_g2s(dst, p.Bhi + src, bytes, bar(BAR_BFULL + slot));
            bulk_g2s(dst + bytes, p.Blo + src, bytes, bar(BAR_BFULL + slot));
          }
          __syncwarp();
        }
      }
    }
    (void)w_bempty;
  } else if (warp == kMmaWarpA || warp == kMmaWarpB) {
    // ============================================================== MMA issuers (whole warp, one elected lane issues)
    MmaState st{};
    st.id = warp == kMmaWarpA ? 0 : 1;
    st.p = &p; st.tmem_base = tmem_base; st.s_base = s_base; st.bar0 = bar0; st.dbg = dbg != nullptr;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      int rt_, z0_, jr_;
      unit_decode<TILT>(p, unit, rt_, z0_, st.nzu, jr_);
      for (int f = 0; f < p.F; ++f, ++st.g) {
        if (F16) {
          switch (p.t[f].kc >> 3) {
            case 1: mma_tonal<1, F16>(st, p.t[f]); break;
            case 2: mma_tonal<2, F16>(st, p.t[f]); break;
            case 3: mma_tonal<3, F16>(st, p.t[f]); break;
            default: mma_tonal<4, F16>(st, p.t[f]); break;
          }
        } else
        switch (p.t[f].kc >> 3) {
          case 1: mma_tonal<1>(st, p.t[f]); break;
          case 2: mma_tonal<2>(st, p.t[f]); break;
          case 3: mma_tonal<3>(st, p.t[f]); break;
          case 4: mma_tonal<4>(st, p.t[f]); break;
          case 5: mma_tonal<5>(st, p.t[f]); break;
          case 6: mma_tonal<6>(st, p.t[f]); break;
          case 7: mma_tonal<7>(st, p.t[f]); break;
          default: mma_tonal<8>(st, p.t[f]); break;
        }
      }
    }
    if (dbg && lane == 0 && st.id == 0) { dbg[3] = clock64() - t_start; dbg[4] = st.w_efull; dbg[5] = st.w_bfull; dbg[6] = st.w_tempty; dbg[7] = st.sc; }
  }
  } else if (warp >= kLoadWarp0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kLoadRegs));
    // ============================================================== E loaders: global -> registers -> TMEM
    const int quad = warp & 3;
    const int tl = quad * 32 + lane;
    long long w_edone = 0;
    int g = 0;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      int rt, z0_, nzu_, jreal;
      unit_decode<TILT>(p, unit, rt, z0_, nzu_, jreal);
      for (int f = 0; f < p.F; ++f, ++g) {
        const UTonal& u = p.t[f];
        // (tilt mode: a tonal lasts ~100k cycles and every poll costs issue slots the epilogue needs)
        if (g >= 2) mbar_wait_relaxed<TILT ? 4096 : 256>(bar(BAR_EDONE + (g & 1)), (uint32_t)(((g - 2) >> 1) & 1), p.err, 1);
        if (g >= 1) {
          const UTonal& up = p.t[f == 0 ? p.F - 1 : f - 1];
          if (4 * (up.kc + u.kc) > p.AR)
            mbar_wait_relaxed<TILT ? 2048 : 256>(bar(BAR_EDONE + ((g - 1) & 1)), (uint32_t)(((g - 1) >> 1) & 1), p.err, 2);
        }
        tc_fence_after();
        const float4* src = reinterpret_cast<const float4*>(p.E + (size_t)rt * p.e_rt_stride + u.e_off) + tl;
        const uint32_t dst = tmem_base + ((uint32_t)(quad * 32) << 16) + a_column(p, g, u.kc);
        const int n16 = u.Mp >> 2;            // 4 Mp columns in groups of 16
        if (TILT) {
          // tilted array: the depth goes into the A operand, A[i, m] = E[i, m] phi_m(z_j), because the B operand
          // (Theta_tau = Phi e^{i k l sin tau}) is per tilt: rebuild the fp32 value from the table's (hi, lo), scale,
          // split again.  Eight modes per step: (re | im) x (hi, lo) chunks in, two 8-column stores out per part.
          const float* phi = p.phiz + (size_t)jreal * p.sumMp + u.offM;
          const int mc = u.Mp >> 2;           // 16-byte chunks per block of Mp columns
          if (F16) {
            // FP16-split operands: 16 modes per step -> the scaled value x sA, hi = fp16(x), lo = fp16(x - hi), two
            // modes per 32-bit TMEM column; image = [re_hi | re_lo | im_hi | im_lo] x kc columns
            const float* phj = p.phiz_im ? p.phiz_im + (size_t)jreal * p.sumMp + u.offM : nullptr;
            const uint32_t kc = (uint32_t)u.kc;
            for (int q16 = 0; q16 < (u.kc >> 3); ++q16) {
              float xr[16], xi[16];
#pragma unroll
              for (int h4 = 0; h4 < 4; ++h4) {
                const int c = 4 * q16 + h4;           // 16-byte chunk (4 modes) of the table's blocks
                if (4 * c < u.Mp) {
                  const float4 rh = __ldg(src + (size_t)(0 * mc + c) * kTileR), rl = __ldg(src + (size_t)(1 * mc + c) * kTileR);
                  const float4 ih = __ldg(src + (size_t)(2 * mc + c) * kTileR), il = __ldg(src + (size_t)(3 * mc + c) * kTileR);
                  const float4 a = __ldg(reinterpret_cast<const float4*>(phi + 4 * c));
                  const float er[4] = {rh.x + rl.x, rh.y + rl.y, rh.z + rl.z, rh.w + rl.w};
                  const float ei[4] = {ih.x + il.x, ih.y + il.y, ih.z + il.z, ih.w + il.w};
                  const float pr[4] = {a.x * u.sA, a.y * u.sA, a.z * u.sA, a.w * u.sA};
                  if (phj) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(phj + 4 * c));
                    const float pi[4] = {b.x * u.sA, b.y * u.sA, b.z * u.sA, b.w * u.sA};
#pragma unroll
                    for (int e = 0; e < 4; ++e) { xr[4 * h4 + e] = fmaf(er[e], pr[e], -ei[e] * pi[e]); xi[4 * h4 + e] = fmaf(er[e], pi[e], ei[e] * pr[e]); }
                  } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) { xr[4 * h4 + e] = er[e] * pr[e]; xi[4 * h4 + e] = ei[e] * pr[e]; }
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e) { xr[4 * h4 + e] = 0.f; xi[4 * h4 + e] = 0.f; }
                }
              }
              float hi[8], lo[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const __half h0 = __float2half_rn(xr[2 * e]), h1 = __float2half_rn(xr[2 * e + 1]);
                const __half l0 = __float2half_rn(xr[2 * e] - __half2float(h0)), l1 = __float2half_rn(xr[2 * e + 1] - __half2float(h1));
                hi[e] = __uint_as_float((uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16));
                lo[e] = __uint_as_float((uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16));
              }
              tc_st8(dst + (uint32_t)(8 * q16), hi);
              tc_st8(dst + kc + (uint32_t)(8 * q16), lo);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const __half h0 = __float2half_rn(xi[2 * e]), h1 = __float2half_rn(xi[2 * e + 1]);
                const __half l0 = __float2half_rn(xi[2 * e] - __half2float(h0)), l1 = __float2half_rn(xi[2 * e + 1] - __half2float(h1));
                hi[e] = __uint_as_float((uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16));
                lo[e] = __uint_as_float((uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16));
              }
              tc_st8(dst + 2u * kc + (uint32_t)(8 * q16), hi);
              tc_st8(dst + 3u * kc + (uint32_t)(8 * q16), lo);
            }
          } else
          if (p.phiz_im) {
            // complex source mode shapes (KRAKENC): a = (E_re + i E_im)(phi_re + i phi_im), both parts per 8 modes
            const float* phj = p.phiz_im + (size_t)jreal * p.sumMp + u.offM;
            for (int q2 = 0; q2 < (u.Mp >> 3); ++q2) {
              float er[8], ei[8], pr[8], pi[8];
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const float4 rh = __ldg(src + (size_t)(0 * mc + 2 * q2 + h2) * kTileR), rl = __ldg(src + (size_t)(1 * mc + 2 * q2 + h2) * kTileR);
                const float4 ih = __ldg(src + (size_t)(2 * mc + 2 * q2 + h2) * kTileR), il = __ldg(src + (size_t)(3 * mc + 2 * q2 + h2) * kTileR);
                const float4 a = __ldg(reinterpret_cast<const float4*>(phi + 8 * q2 + 4 * h2)), b = __ldg(reinterpret_cast<const float4*>(phj + 8 * q2 + 4 * h2));
                er[4 * h2 + 0] = rh.x + rl.x; er[4 * h2 + 1] = rh.y + rl.y; er[4 * h2 + 2] = rh.z + rl.z; er[4 * h2 + 3] = rh.w + rl.w;
                ei[4 * h2 + 0] = ih.x + il.x; ei[4 * h2 + 1] = ih.y + il.y; ei[4 * h2 + 2] = ih.z + il.z; ei[4 * h2 + 3] = ih.w + il.w;
                pr[4 * h2 + 0] = a.x; pr[4 * h2 + 1] = a.y; pr[4 * h2 + 2] = a.z; pr[4 * h2 + 3] = a.w;
                pi[4 * h2 + 0] = b.x; pi[4 * h2 + 1] = b.y; pi[4 * h2 + 2] = b.z; pi[4 * h2 + 3] = b.w;
              }
              float hi[8], lo[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) { const float x = fmaf(er[e], pr[e], -ei[e] * pi[e]); hi[e] = tf32_rna(x); lo[e] = tf32_rna(x - hi[e]); }
              tc_st8(dst + (uint32_t)(8 * q2), hi);
              tc_st8(dst + (uint32_t)(u.Mp + 8 * q2), lo);
#pragma unroll
              for (int e = 0; e < 8; ++e) { const float x = fmaf(er[e], pi[e], ei[e] * pr[e]); hi[e] = tf32_rna(x); lo[e] = tf32_rna(x - hi[e]); }
              tc_st8(dst + (uint32_t)(2 * u.Mp + 8 * q2), hi);
              tc_st8(dst + (uint32_t)(3 * u.Mp + 8 * q2), lo);
            }
          } else
          for (int part = 0; part < 2; ++part) {
            const int hb = 2 * part * mc, lb = hb + mc;
            for (int q2 = 0; q2 < (u.Mp >> 3); ++q2) {
              const float4 h0 = __ldg(src + (size_t)(hb + 2 * q2) * kTileR), h1 = __ldg(src + (size_t)(hb + 2 * q2 + 1) * kTileR);
              const float4 l0 = __ldg(src + (size_t)(lb + 2 * q2) * kTileR), l1 = __ldg(src + (size_t)(lb + 2 * q2 + 1) * kTileR);
              const float4 pa = __ldg(reinterpret_cast<const float4*>(phi + 8 * q2)), pb = __ldg(reinterpret_cast<const float4*>(phi + 8 * q2 + 4));
              const float x[8] = {(h0.x + l0.x) * pa.x, (h0.y + l0.y) * pa.y, (h0.z + l0.z) * pa.z, (h0.w + l0.w) * pa.w,
                                  (h1.x + l1.x) * pb.x, (h1.y + l1.y) * pb.y, (h1.z + l1.z) * pb.z, (h1.w + l1.w) * pb.w};
              float hi[8], lo[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) { hi[e] = tf32_rna(x[e]); lo[e] = tf32_rna(x[e] - hi[e]); }
              tc_st8(dst + (uint32_t)(2 * part * u.Mp + 8 * q2), hi);
              tc_st8(dst + (uint32_t)((2 * part + 1) * u.Mp + 8 * q2), lo);
            }
          }
        } else
        for (int c = 0; c < n16; c += 2) {
          const float4 v0 = __ldg(src + (size_t)(4 * c + 0) * kTileR), v1 = __ldg(src + (size_t)(4 * c + 1) * kTileR);
          const float4 v2 = __ldg(src + (size_t)(4 * c + 2) * kTileR), v3 = __ldg(src + (size_t)(4 * c + 3) * kTileR);
          const float4 v4 = __ldg(src + (size_t)(4 * c + 4) * kTileR), v5 = __ldg(src + (size_t)(4 * c + 5) * kTileR);
          const float4 v6 = __ldg(src + (size_t)(4 * c + 6) * kTileR), v7 = __ldg(src + (size_t)(4 * c + 7) * kTileR);
          tc_st16(dst + (uint32_t)(16 * c), v0, v1, v2, v3);
          tc_st16(dst + (uint32_t)(16 * c + 16), v4, v5, v6, v7);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_EFULL + (g & 1)));
      }
    }
    if (dbg && lane == 0 && quad == 0) { dbg[12] = clock64() - t_start; }
  } else {
    // ============================================================== epilogue
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    if (TILT) {
      // ------------------------------------------------------------ tilted array (receiver space)
      // Stage k of a tonal holds, for one tilt and one half of the array, T_n = sum_m Theta_tau[n, m] a_m as adjacent
      // (re-row, im-row) column pairs; p_n = s_n T_n with s_n = (r / r_n)^1/2 = (1 - l_n sin(tau) / r)^-1/2, a factor
      // that depends on the lane's range and sits between the modes and the covariance, so the reduction over
      // receivers happens here:  B = sum_j |sum_n conj(L_nj) s_n T_n|^2 / sum_n s_n^2 |T_n|^2.
      // Stage order (k = 4q + {0,1,2,3} -> (tilt 2q, half 0), (2q+1, 0), (2q, 1), (2q+1, 1)) gives both halves of a
      // tilt the same parity, i.e. the same accumulator buffer and the same pair of warps, which keep their partial
      // sums in registers between the two stages; the pair splits every stage (16 receivers each) and meets once per
      // tilt on a 64-thread named barrier to add the two partial sums.
      constexpr int kTJ = 4;                  // covariance ranks handled in registers
      const int ew = warp, quad = warp & 3, set = (ew >> 2) & 1, sub = ew >> 3;
      const int tl = quad * 32 + lane;
      const uint32_t NA = (uint32_t)p.NA;
      const uint32_t t_re = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)set * (2 * NA);
      float* acc = reinterpret_cast<float*>(smem + p.smem_acc);     // [ZC / 2][128]
      uint32_t sc = 0;
      long long w_tfull = 0;
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        int rt, z0, nzu, jreal;
        unit_decode<true>(p, unit, rt, z0, nzu, jreal);
        const int i = rt * kTileR + tl;
        const float rinv = i < p.Nr ? __ldg(p.rinv + i) : 0.f;
        const int t0 = z0 >> 1;
        for (int f = 0; f < p.F; ++f) {
          const UTonal& u = p.t[f];
          const int J = u.J;
          const float* lev_s = reinterpret_cast<const float*>(smem + p.smem_aux);
          const float2* c_f = reinterpret_cast<const float2*>(lev_s + 64) + f * 256;     // [64][4]
          float* xch = const_cast<float*>(lev_s) + 64 + p.F * 512;                       // [2][2 sets][4 quads][9][32]
          // the rank-1 (packed) and general (scalar) reductions keep different state: two instantiations of the tonal's
          // stage loop, so that neither carries the other's registers
          auto tonal_loop = [&](auto rank1_tag) {
          constexpr bool R1 = decltype(rank1_tag)::value;
          const f32x2 one2 = pk(1.f, 1.f);
          const f32x2* lev2 = reinterpret_cast<const f32x2*>(lev_s);
          const ulonglong2* cp_f = reinterpret_cast<const ulonglong2*>(lev_s + 64 + p.F * 512 + 2 * 2 * 4 * 9 * 32) + f * 32;
          // sc is even at the start of every tonal (nzu is a multiple of 4), so this set's stages are k = 4q + set
          // (half 0) and 4q + 2 + set (half 1) of tilt 2q + set: iterate over the tilts, both halves unrolled
          for (int k4 = 0; k4 < nzu; k4 += 4) {
            const int tloc = (k4 >> 1) + set;
            const float xs = __ldg(p.sin_t + t0 + tloc) * rinv;
            float norm = 0.f, ur[kTJ], ui[kTJ];
#pragma unroll
            for (int j = 0; j < kTJ; ++j) ur[j] = ui[j] = 0.f;
            f32x2 norm2 = 0ull, u1 = 0ull, w1 = 0ull, v1 = 0ull, v2 = 0ull;      // rank-1 path: packed over receiver pairs
            const f32x2 nxs2 = pk(-xs, -xs);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
            const uint32_t s_own = sc + (uint32_t)(k4 + 2 * half + set);
            mbar_wait_t(bar(BAR_TFULL + set), (s_own >> 1) & 1u, p.err, 7, dbg ? &w_tfull : nullptr);
            tc_fence_after();
            if (BOGP_UMMA_DBG && p.skip == 2) {
              tc_fence_before();
              if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + set));
              continue;
            }
            // The two warps (sub 0 / 1) of this (buffer, quadrant) split the stage: 16 receivers = 32 columns each, all
            // four tcgen05.ld up front, release at the single wait, then the receiver-space reduction from registers.
            auto reduce16 = [&](const float* re, const float* im, int nb) {
              if (R1) {
                // rank-1 covariance (K = d d^H, the reference's per-snapshot case), two receivers per instruction:
                // columns 4p .. 4p+3 of a chunk are (Re n, Re n+1, Im n, Im n+1) of receiver pair p
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                  const f32x2 r01 = pk(re[4 * pp], re[4 * pp + 1]), r23 = pk(re[4 * pp + 2], re[4 * pp + 3]);
                  const f32x2 i01 = pk(im[4 * pp], im[4 * pp + 1]), i23 = pk(im[4 * pp + 2], im[4 * pp + 3]);
                  const f32x2 tr = sub2(r01, i23), ti = add2(i01, r23);
                  const int pr = (nb >> 1) + pp;
                  const f32x2 dd = fma2(lev2[pr], nxs2, one2);                    // 1 - l_n sin(tau) / r
                  float d0, d1;
                  upk(dd, d0, d1);
                  const f32x2 sn = pk(rsqrt_fast(d0), rsqrt_fast(d1));
                  const f32x2 a = mul2(sn, tr), b = mul2(sn, ti);                 // p_n = s_n T_n
                  norm2 = fma2(a, a, norm2);
                  norm2 = fma2(b, b, norm2);
                  const ulonglong2 c = cp_f[pr];                                  // (cx0, cx1), (cy0, cy1)
                  u1 = fma2(c.x, a, u1); w1 = fma2(c.y, b, w1);                   // Re u = sum cx a - cy b
                  v1 = fma2(c.x, b, v1); v2 = fma2(c.y, a, v2);                   // Im u = sum cx b + cy a
                }
                return;
              }
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int n = nb + q, c0 = 4 * (q >> 1) + (q & 1);
                const float tr = re[c0] - im[c0 + 2], ti = im[c0] + re[c0 + 2];
                const float sn = rsqrt_fast(fmaf(-lev_s[n], xs, 1.f));
                const float a = sn * tr, b = sn * ti;
                norm = fmaf(a, a, fmaf(b, b, norm));
#pragma unroll
                for (int j = 0; j < kTJ; ++j)
                  if (j < J) {
                    const float2 c = c_f[n * 4 + j];
                    ur[j] = fmaf(c.x, a, fmaf(-c.y, b, ur[j]));
                    ui[j] = fmaf(c.x, b, fmaf(c.y, a, ui[j]));
                  }
              }
            };
            const uint32_t tc0 = t_re + 32u * (uint32_t)sub;
            const int nb0 = 32 * half + 16 * sub;
            float ar[16], ai[16], br[16], bi[16];
            tc_ld16(tc0, ar); tc_ld16(tc0 + NA, ai);
            tc_ld16(tc0 + 16u, br); tc_ld16(tc0 + NA + 16u, bi);
            tc_wait_ld();
            tc_fence_before();
            if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + set));
            reduce16(ar, ai, nb0);
            reduce16(br, bi, nb0 + 8);
            }
            {
              if (R1) {
                float x0, x1, y0, y1;
                upk(norm2, x0, x1); norm = x0 + x1;
                upk(u1, x0, x1); upk(w1, y0, y1); ur[0] = (x0 + x1) - (y0 + y1);
                upk(v1, x0, x1); upk(v2, y0, y1); ui[0] = (x0 + x1) + (y0 + y1);
              }
              // sub 1 hands its partial sums to sub 0 through shared memory (double-buffered by tilt pair)
              float* x = xch + ((((tloc >> 1) & 1) * 2 + set) * 4 + quad) * (32 * (1 + 2 * kTJ)) + lane;
              if (sub == 1) {
                x[0] = norm;
#pragma unroll
                for (int j = 0; j < kTJ; ++j) { x[32 * (1 + 2 * j)] = ur[j]; x[32 * (2 + 2 * j)] = ui[j]; }
              }
              asm volatile("bar.sync %0, %1;" ::"r"(5 + set * 4 + quad), "r"(64) : "memory");
              if (sub == 0) {
                norm += x[0];
                float num = 0.f;
#pragma unroll
                for (int j = 0; j < kTJ; ++j) {
                  const float vr = ur[j] + x[32 * (1 + 2 * j)], vi = ui[j] + x[32 * (2 + 2 * j)];
                  num = fmaf(vr, vr, fmaf(vi, vi, num));
                }
                const float bq = __fdividef(num, norm);
                const float val = p.atype == BOGP_ATYPE_CBF ? bq : u.trK - bq;
                float* a = acc + (size_t)tloc * kTileR + tl;
                *a = (f == 0) ? val : (p.mf == BOGP_MF_MEAN ? *a + val : *a * val);
              }
            }
          }
          sc += (uint32_t)nzu;
          };
          if (J == 1) tonal_loop(std::true_type{}); else tonal_loop(std::false_type{});
        }
        constexpr int kPerQuad = kEpiWarps / 4;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "r"(kPerQuad * 32) : "memory");
        const float scale = p.mf == BOGP_MF_MEAN ? 1.f / (float)p.F : 1.f;
        const int k4 = ew >> 2, ntu = nzu >> 1;
        unsigned long long best = 0ull;
        if (i < p.Nr)
          for (int tloc = k4; tloc < ntu; tloc += kPerQuad) {
            const int tau = t0 + tloc;
            if (tau >= p.Nt) continue;               // padding tilt (odd Nt)
            const float v = acc[(size_t)tloc * kTileR + tl] * scale;
            const size_t flat = ((size_t)tau * p.Nzr + jreal) * p.Nr + i;
            p.out[flat] = v;
            if (p.ext_mode && v == v) {
              uint32_t uu = __float_as_uint(v);
              uu = (uu & 0x80000000u) ? ~uu : (uu | 0x80000000u);
              if (p.ext_mode == 2) uu = ~uu;
              const unsigned long long key = ((unsigned long long)uu << 32) | (0xffffffffu - (uint32_t)flat);
              best = key > best ? key : best;
            }
          }
        if (p.ext_mode) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
          }
          if (lane == 0 && best) atomicMax(p.ext_key, best);
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "r"(kPerQuad * 32) : "memory");
      }
    } else {
    // 16 warps = 2 buffer sets x 2 depth parities x 4 TMEM lane quadrants.  Set b drains accumulator buffer b only,
    // so the two buffers are processed concurrently by different warps, and a warp releases its buffer as soon as its
    // last tcgen05.ld has landed in registers -- the squares, the division and the combine run after the release.
    const int ew = warp;                     // 0..15
    const int quad = warp & 3;               // TMEM lane quadrant this warp may read
    const int set = (ew >> 2) & 1;           // accumulator buffer (= stage parity) this warp drains
    const int sub = ew >> 3;                 // depth residue (j % kEpiSub) this warp handles
    const int tl = quad * 32 + lane;         // lane in the tile = range index in the tile
    const uint32_t NA = (uint32_t)p.NA;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)set * (2 * NA);
    float* acc = reinterpret_cast<float*>(smem + p.smem_acc);     // [ZC][128]
    uint32_t sc = 0;
    long long w_tfull = 0;
    // The running mean / product of depth j lives in acc[j] and is updated tonal after tonal, by this warp or by its
    // partner (other accumulator set, same depth parity and quadrant) depending on the stage parity.  Before it starts
    // a tonal a warp makes sure the partner has finished the previous tonal (release / acquire on a counter
    // in shared memory, once per tonal): with one-stage tonals on tiny grids two consecutive stages touch the same
    // depth and the order of the two updates would otherwise be a race.
    int gt = 0;                              // tonals finished by this warp (over all units)
    const uint32_t seq_mine = smem_u32(const_cast<int*>(acc_seq) + (set * 2 + sub) * 4 + quad);
    const uint32_t seq_partner = smem_u32(const_cast<int*>(acc_seq) + ((set ^ 1) * 2 + sub) * 4 + quad);
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      const int rt = unit % p.nrt, zc = unit / p.nrt;
      const int z0 = zc * p.ZC, nzu = min(p.ZC, p.Nz - z0);
      for (int f = 0; f < p.F; ++f) {
        const UTonal& u = p.t[f];
        const int rowsP = u.rowsP, nq = u.nq;
        if (gt > 0) {
          int v;
          do { asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(seq_partner) : "memory"); } while (v < gt);
        }
        const bool fast = (u.n1 == 2 * nq) && nq <= 8;
        for (int j0 = 0; j0 < nzu; j0 += u.nj, ++sc) {
          if ((int)(sc & 1u) != set) continue;
          const int njs = min(u.nj, nzu - j0);
          mbar_wait_t(bar(BAR_TFULL + set), (sc >> 1) & 1u, p.err, 7, dbg ? &w_tfull : nullptr);
          const long long tw = (dbg && ew == 0 && blockIdx.x == 0) ? clock64() : 0;
          tc_fence_after();
          if (quad == 0 && sub == 0) dbg_stamp(p, 2, sc);
          // slots of this stage owned by this warp: s = s_first, s_first + kEpiSub, ...
          const int s_first = (sub - j0 % kEpiSub + kEpiSub) % kEpiSub;
          bool released = false;
          if ((p.two_slot & 1) && fast && rowsP == 16 && (nq == 1 || u.fold) && s_first + kEpiSub < njs && s_first + 2 * kEpiSub >= njs &&
              !(BOGP_UMMA_DBG && p.skip == 2)) {
            // small tonals (one 16-column chunk per depth slot, two slots for this warp): both slots' loads in flight
            // at once, one wait, release, then the arithmetic -- the per-stage cost of these tonals is TMEM-load
            // latency, not tensor work
            const int sA = s_first, sB = s_first + kEpiSub;
            float reA[16], imA[16], reB[16], imB[16];
            const uint32_t tA = t_lane + (uint32_t)(sA * 16), tB = t_lane + (uint32_t)(sB * 16);
            tc_ld16(tA, reA); tc_ld16(tA + NA, imA);
            tc_ld16(tB, reB); tc_ld16(tB + NA, imB);
            tc_wait_ld();
            tc_fence_before();
            if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + set));
            released = true;
            if (quad == 0 && sub == 0) dbg_stamp(p, 3, sc);
            auto slot = [&](const float* re, const float* im, int sidx) {
              float num, numsq;
              if (u.fold) {
                const float tr = fmaf(u.c1r, re[0], fmaf(-u.c1i, im[0], fmaf(u.c2r, re[1], -u.c2i * im[1])));
                const float ti = fmaf(u.c1r, im[0], fmaf(u.c1i, re[0], fmaf(u.c2r, im[1], u.c2i * re[1])));
                num = fmaf(tr, tr, ti * ti);
                numsq = 0.f;
              } else {
                const float tr = re[0] - im[1], ti = im[0] + re[1];
                num = fmaf(tr, tr, ti * ti);
                numsq = fmaf(re[0], re[0], fmaf(im[0], im[0], fmaf(re[1], re[1], im[1] * im[1])));
              }
              unsigned long long t0 = 0ull, t1 = 0ull, t2 = 0ull, t3 = 0ull;
#pragma unroll
              for (int q = 0; q < 16; q += 4) {
                sq_acc2(t0, re[q], re[q + 1]);
                sq_acc2(t1, im[q], im[q + 1]);
                sq_acc2(t2, re[q + 2], re[q + 3]);
                sq_acc2(t3, im[q + 2], im[q + 3]);
              }
              const float norm = ((sum2(t0) + sum2(t1)) + (sum2(t2) + sum2(t3))) - numsq;
              const float b = __fdividef(num, norm);
              const float val = p.atype == BOGP_ATYPE_CBF ? b : u.trK - b;
              float* a = acc + (size_t)(j0 + sidx) * kTileR + tl;
              *a = (f == 0) ? val : (p.mf == BOGP_MF_MEAN ? *a + val : *a * val);
            };
            slot(reA, imA, sA);
            slot(reB, imB, sB);
          } else if ((p.two_slot & 2) && fast && rowsP == 32 && (nq == 1 || u.fold) && s_first < njs && s_first + kEpiSub >= njs &&
                     !(BOGP_UMMA_DBG && p.skip == 2)) {
            // one slot of two 16-column chunks: all four loads in flight, one wait
            float reA[16], imA[16], reB[16], imB[16];
            const uint32_t tA = t_lane + (uint32_t)(s_first * 32);
            tc_ld16(tA, reA); tc_ld16(tA + NA, imA);
            tc_ld16(tA + 16u, reB); tc_ld16(tA + NA + 16u, imB);
            tc_wait_ld();
            tc_fence_before();
            if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + set));
            released = true;
            if (quad == 0 && sub == 0) dbg_stamp(p, 3, sc);
            float num, numsq;
            if (u.fold) {
              const float tr = fmaf(u.c1r, reA[0], fmaf(-u.c1i, imA[0], fmaf(u.c2r, reA[1], -u.c2i * imA[1])));
              const float ti = fmaf(u.c1r, imA[0], fmaf(u.c1i, reA[0], fmaf(u.c2r, imA[1], u.c2i * reA[1])));
              num = fmaf(tr, tr, ti * ti);
              numsq = 0.f;
            } else {
              const float tr = reA[0] - imA[1], ti = imA[0] + reA[1];
              num = fmaf(tr, tr, ti * ti);
              numsq = fmaf(reA[0], reA[0], fmaf(imA[0], imA[0], fmaf(reA[1], reA[1], imA[1] * imA[1])));
            }
            unsigned long long t0 = 0ull, t1 = 0ull, t2 = 0ull, t3 = 0ull;
#pragma unroll
            for (int q = 0; q < 16; q += 4) {
              sq_acc2(t0, reA[q], reA[q + 1]);
              sq_acc2(t1, imA[q], imA[q + 1]);
              sq_acc2(t2, reA[q + 2], reA[q + 3]);
              sq_acc2(t3, imA[q + 2], imA[q + 3]);
            }
#pragma unroll
            for (int q = 0; q < 16; q += 4) {
              sq_acc2(t0, reB[q], reB[q + 1]);
              sq_acc2(t1, imB[q], imB[q + 1]);
              sq_acc2(t2, reB[q + 2], reB[q + 3]);
              sq_acc2(t3, imB[q + 2], imB[q + 3]);
            }
            const float norm = ((sum2(t0) + sum2(t1)) + (sum2(t2) + sum2(t3))) - numsq;
            const float b = __fdividef(num, norm);
            const float val = p.atype == BOGP_ATYPE_CBF ? b : u.trK - b;
            float* a = acc + (size_t)(j0 + s_first) * kTileR + tl;
            *a = (f == 0) ? val : (p.mf == BOGP_MF_MEAN ? *a + val : *a * val);
          } else
          for (int s = s_first; s < njs && !(BOGP_UMMA_DBG && p.skip == 2); s += kEpiSub) {
            const int j = j0 + s;
            const bool last_slot = s + kEpiSub >= njs;
            const uint32_t t_re = t_lane + (uint32_t)(s * rowsP);
            float norm = 0.f, num = 0.f;
            if (fast) {
              // Real receiver modes (KRAKEN): every |T|^2 term is a plain square except the nq numerator pairs, and
              // (r0 - i1)^2 + (i0 + r1)^2 = r0^2 + i0^2 + r1^2 + i1^2 + 2 (i0 r1 - r0 i1).  One FFMA per value over
              // ALL columns gives tot = norm + numsq; the numerator pairs sit in the first columns already loaded.
              // The epilogue is ISSUE-bound (all warps of a lane quadrant share one scheduler), so this path is kept
              // to the bare instruction count: 2 loads + 32 FFMA per 16 columns, rank-1 numerator special-cased.
              unsigned long long t0 = 0ull, t1 = 0ull, t2 = 0ull, t3 = 0ull;      // packed (even, odd column) partial sums
              float numsq = 0.f;
              const int c_last = rowsP - 16;
              for (int c0 = 0; c0 <= c_last; c0 += 16) {
                float re[16], im[16];
                tc_ld16(t_re + (uint32_t)c0, re);
                tc_ld16(t_re + NA + (uint32_t)c0, im);
                tc_wait_ld();
                if (last_slot && c0 == c_last) {
                  tc_fence_before();
                  if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + set));
                  released = true;
                  if (quad == 0 && sub == 0) dbg_stamp(p, 3, sc);
                }
                if (c0 == 0) {
                  if (u.fold) {
                    const float tr = fmaf(u.c1r, re[0], fmaf(-u.c1i, im[0], fmaf(u.c2r, re[1], -u.c2i * im[1])));
                    const float ti = fmaf(u.c1r, im[0], fmaf(u.c1i, re[0], fmaf(u.c2r, im[1], u.c2i * re[1])));
                    num = fmaf(tr, tr, ti * ti);
                  } else if (nq == 1) {
                    const float tr = re[0] - im[1], ti = im[0] + re[1];
                    num = fmaf(tr, tr, ti * ti);
                    numsq = fmaf(re[0], re[0], fmaf(im[0], im[0], fmaf(re[1], re[1], im[1] * im[1])));
                  } else {
                    numerator16(re, im, nq, num, numsq);
                  }
                }
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                  sq_acc2(t0, re[q], re[q + 1]);
                  sq_acc2(t1, im[q], im[q + 1]);
                  sq_acc2(t2, re[q + 2], re[q + 3]);
                  sq_acc2(t3, im[q + 2], im[q + 3]);
                }
              }
              norm = ((sum2(t0) + sum2(t1)) + (sum2(t2) + sum2(t3))) - numsq;
            } else {
              const int nq2 = 2 * nq, n1 = u.n1;
              for (int c0 = 0; c0 < rowsP; c0 += 16) {
                float re[16], im[16];
                tc_ld16(t_re + (uint32_t)c0, re);
                tc_ld16(t_re + NA + (uint32_t)c0, im);
                tc_wait_ld();
                if (last_slot && c0 + 16 >= rowsP) {
                  tc_fence_before();
                  if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + set));
                  released = true;
                  if (quad == 0 && sub == 0) dbg_stamp(p, 3, sc);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const int c = c0 + 2 * q;
                  const float r0 = re[2 * q], r1 = re[2 * q + 1], i0 = im[2 * q], i1 = im[2 * q + 1];
                  if (c >= n1) {
                    norm += r0 * r0 + i0 * i0 + r1 * r1 + i1 * i1;
                  } else {
                    const float tr = r0 - i1, ti = i0 + r1;
                    const float v = tr * tr + ti * ti;
                    if (c < nq2) num += v; else norm += v;
                  }
                }
              }
            }
            const float b = __fdividef(num, norm);
            const float val = p.atype == BOGP_ATYPE_CBF ? b : u.trK - b;
            float* a = acc + (size_t)j * kTileR + tl;
            *a = (f == 0) ? val : (p.mf == BOGP_MF_MEAN ? *a + val : *a * val);
          }
          if (quad == 0 && sub == 0) dbg_stamp(p, 4, sc);
          if (!released) {      // no slot of this stage belongs to this warp
            tc_fence_before();
            if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + set));
          }
          if (tw && lane == 0 && unit == (int)blockIdx.x) { p.dbg[16 * 201 + f] += clock64() - tw; p.dbg[16 * 203 + f] += 1; }
        }
        ++gt;
        __syncwarp();
        if (lane == 0) asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(seq_mine), "r"(gt) : "memory");
      }
      // acc[j][tl] is updated by the two sets' warps of this (quadrant, parity) in stage order; the buffer hand-shake
      // orders those updates (release/acquire on the mbarriers), but the write-out below reads both sets' updates, so
      // the warps of this quadrant meet on a named barrier first (4 warps = 128 threads per quadrant).
      constexpr int kPerQuad = kEpiWarps / 4;      // warps sharing a lane quadrant
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "r"(kPerQuad * 32) : "memory");
      const int i = rt * kTileR + tl;
      const float scale = p.mf == BOGP_MF_MEAN ? 1.f / (float)p.F : 1.f;
      const int k4 = ew >> 2;                  // this warp writes depths j = k4, k4 + kPerQuad, ...
      unsigned long long best = 0ull;
      if (i < p.Nr)
        for (int j = k4; j < nzu; j += kPerQuad) {
          const float v = acc[(size_t)j * kTileR + tl] * scale;
          p.out[(size_t)(z0 + j) * p.Nr + i] = v;
          if (p.ext_mode && v == v) {
            uint32_t u = __float_as_uint(v);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);            // order-preserving map float -> uint32
            if (p.ext_mode == 2) u = ~u;                                // arg-min: largest key = smallest value
            const unsigned long long key = ((unsigned long long)u << 32) | (0xffffffffu - ((uint32_t)(z0 + j) * (uint32_t)p.Nr + (uint32_t)i));
            best = key > best ? key : best;                             // ties: lowest flat index (np.argmax rule)
          }
        }
      if (p.ext_mode) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
          best = other > best ? other : best;
        }
        if (lane == 0 && best) atomicMax(p.ext_key, best);
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "r"(kPerQuad * 32) : "memory");     // before the next unit overwrites acc
    }
    if (dbg && lane == 0 && (ew == 0 || ew == 4)) { dbg[8 + (ew >> 2) * 2] = clock64() - t_start; dbg[9 + (ew >> 2) * 2] = w_tfull; }
      }
}
  tc_fence_before();
  __syncthreads();
  if (p.ext_mode && threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.ext_done, 1u) == gridDim.x - 1) {
      __threadfence();
      const unsigned long long key = atomicMax(p.ext_key, 0ull);
      if (key == 0ull) { *p.ext_val = 0.f; *p.ext_idx = -1; }
      else {
        uint32_t u = (uint32_t)(key >> 32);
        if (p.ext_mode == 2) u = ~u;
        u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        *p.ext_val = __uint_as_float(u);
        *p.ext_idx = (long long)(0xffffffffu - (uint32_t)(key & 0xffffffffu));
      }
    }
  }
  if (warp == kMmaWarpA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int device_sms() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

// smallest, largest, second smallest, second largest, ... (by padded mode count); BOGP_UMMA_TONAL_ORDER=0: as given
static void tonal_order(const ModesetDev& d, int* order) {
  int idx[BOGP_MAX_TONALS];
  for (int f = 0; f < d.F; ++f) idx[f] = f;
  std::stable_sort(idx, idx + d.F, [&](int a, int b) { return d.t[a].Mp < d.t[b].Mp; });
  const char* e = std::getenv("BOGP_UMMA_TONAL_ORDER");
  const bool interleave = !(e && e[0] == '0');
  for (int i = 0, lo = 0, hi = d.F - 1; i < d.F; ++i) order[i] = !interleave ? i : ((i & 1) ? idx[hi--] : idx[lo++]);
}

// Fills the size-independent part of the plan; returns false when the mode set does not fit the kernel.
static bool plan_tonals(const bogp_modeset_s* h, UParams& p, int ZC) {
  const ModesetDev& d = h->dev;
  if (d.src_complex) return false;     // complex source tables would need a complex Bop (not built yet)
  unsigned e_off = 0;
  int maxMp = 0, maxRows = 0;
  // Tonal order: small and large tonals alternate (smallest, largest, second smallest, ...) so that the E images of
  // consecutive tonals fit side by side in the TMEM A region and the loaders never wait for the previous tonal's MMAs
  // (ascending order puts the 40- and 48-mode tonals next to each other: 4 (40 + 48) > 256 columns).  The mean /
  // product over tonals is accumulated in this order.
  int order[BOGP_MAX_TONALS];
  tonal_order(d, order);
  for (int f = 0; f < d.F; ++f) {
    const TonalMeta& t = d.t[order[f]];
    UTonal& u = p.t[f];
    u.src = order[f];
    u.Mp = t.Mp; u.kc = t.Mp; u.rowsP = t.u_rowsP; u.n1 = t.u_p0; u.nq = t.u_fold ? 0 : t.n_qpair; u.trK = t.trK;
    u.fold = t.u_fold; u.c1r = t.u_c[0]; u.c1i = t.u_c[1]; u.c2r = t.u_c[2]; u.c2i = t.u_c[3];
    u.e_bytes = 2048u * (unsigned)u.Mp;
    u.e_off = e_off;
    e_off += u.e_bytes;
    u.b_zbytes = (unsigned)(u.rowsP * u.Mp * 4);
    if (u.rowsP * u.Mp / 4 > kTabMaxQ * kTabThreads) return false;
    maxMp = std::max(maxMp, u.Mp);
    maxRows = std::max(maxRows, u.rowsP);
  }
  // TMEM: 4 NA accumulator columns + AR columns for the E image (two tonals side by side when they fit)
  if (4 * maxMp > 256) return false;
  p.AR = maxRows <= 64 ? 256 : pad_to(4 * maxMp, 32);
  p.NA = std::min(128, ((512 - p.AR) / 4) & ~15);
  if (maxRows > p.NA) return false;
  p.F = d.F;
  p.e_rt_stride = e_off;
  const unsigned acc_bytes = (unsigned)ZC * kTileR * 4u;
  const unsigned bar_bytes = 256;
  unsigned stage = (kSmemLimit - acc_bytes - bar_bytes) / kStages;
  stage = std::min(stage & ~127u, 32u * 1024u);
  for (int f = 0; f < d.F; ++f) {
    UTonal& u = p.t[f];
    const unsigned per = 2u * u.b_zbytes;          // hi + lo of one depth
    if (per > stage) return false;
    u.nj = (int)std::min<unsigned>(std::min<unsigned>(stage / per, (unsigned)(p.NA / u.rowsP)), 16u);
    if (u.nj < 1) return false;
  }
  p.stage_bytes = stage;
  p.smem_B = 0;
  p.smem_acc = p.smem_B + kStages * stage;
  p.smem_bar = p.smem_acc + acc_bytes;
  return true;
}

static int choose_zc(int Nr, int Nz, int sms) {
  const int nrt = (Nr + kTileR - 1) / kTileR;
  const int chunks = std::max(1, sms / nrt);
  int zc = (Nz + chunks - 1) / chunks;
  zc = std::max(8, std::min(64, (zc + 7) & ~7));
  return zc;
}

bool umma_supported(const bogp_modeset_s* h) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10 || !h->has_cov) return false;
  if (umma16_supported(h)) return true;
  UParams p{};
  return plan_tonals(h, p, 64);
}

int launch_grid_umma(bogp_modeset_s* h, const double* r_host, int Nr, const double* z_host, int Nz, int atype, int mf,
                     float* out_dev, cudaStream_t s, int ext_mode, float* ext_val_dev, long long* ext_idx_dev,
                     const PeerArgs* peer, bool* peer_done) {
  // FP16-split engine (kernels_umma16.cu) unless BOGP_UMMA_KIND=tf32 or the mode set does not fit it
  if (peer_done) *peer_done = false;
  if (umma16_supported(h)) {
    const bool fuse = peer && ext_mode && (long long)Nr * Nz < (1ll << 32);
    if (peer_done) *peer_done = fuse;
    return launch_grid_umma16(h, r_host, Nr, z_host, Nz, atype, mf, out_dev, s, ext_mode, ext_val_dev, ext_idx_dev, fuse ? peer : nullptr);
  }
  const ModesetDev& d = h->dev;
  const int sms = device_sms();
  UParams p{};
  const int ZC = choose_zc(Nr, Nz, sms);
  if (!plan_tonals(h, p, ZC)) { set_error("UMMA engine: mode set does not fit (operator rows, Mp > 64 or complex source table)"); return BOGP_ERR_UNSUPPORTED; }
  p.Nr = Nr; p.Nz = Nz; p.ZC = ZC; p.atype = atype; p.mf = mf;
  { const char* e = std::getenv("BOGP_UMMA_TWO_SLOT"); p.two_slot = e ? std::atoi(e) : 3; }      // bit 0: 16-row tonals (two slots at once), bit 1: 32-row tonals (both chunks at once)
  p.nrt = (Nr + kTileR - 1) / kTileR;
  const int nzc = (Nz + ZC - 1) / ZC;
  p.units = p.nrt * nzc;
  unsigned long long b_off = 0;
  for (int f = 0; f < d.F; ++f) { p.t[f].b_off = b_off; b_off += (unsigned long long)Nz * p.t[f].b_zbytes; }
  // scratch: [r | z doubles][err int][E table][Bhi][Blo]
  const size_t off_rz = 0, bytes_rz = ((size_t)Nr + Nz) * sizeof(double);
  const size_t off_err = (bytes_rz + 255) & ~(size_t)255;
  const size_t off_E = off_err + 256;
  const size_t bytes_E = (size_t)p.nrt * p.e_rt_stride;
  const size_t off_Bhi = (off_E + bytes_E + 255) & ~(size_t)255;
  const size_t off_Blo = (off_Bhi + b_off + 255) & ~(size_t)255;
  if (int rc = ensure_scratch(h, off_Blo + b_off + 256)) return rc;
  unsigned char* base = static_cast<unsigned char*>(h->grid_scratch);
  double* r_dev = reinterpret_cast<double*>(base + off_rz);
  double* z_dev = r_dev + Nr;
  p.err = reinterpret_cast<int*>(base + off_err);
  if (ext_mode && (long long)Nr * Nz < (1ll << 32)) {
    p.ext_mode = ext_mode;
    p.ext_key = reinterpret_cast<unsigned long long*>(base + off_err + 16);
    p.ext_done = reinterpret_cast<unsigned int*>(base + off_err + 32);
    p.ext_val = ext_val_dev; p.ext_idx = ext_idx_dev;
  }
  p.E = base + off_E; p.Bhi = base + off_Bhi; p.Blo = base + off_Blo;
  p.out = out_dev;
  const char* dbg_env = BOGP_UMMA_DBG ? std::getenv("BOGP_UMMA_DBG") : nullptr;
  static long long* dbg_dev = nullptr;
  if (BOGP_UMMA_DBG) { const char* sk = std::getenv("BOGP_UMMA_SKIP"); p.skip = sk ? std::atoi(sk) : 0; }
  if (dbg_env && dbg_env[0] == '1') {
    if (!dbg_dev) BOGP_CUDA(cudaMalloc(&dbg_dev, 256 * 16 * sizeof(long long)));
    BOGP_CUDA(cudaMemsetAsync(dbg_dev, 0, 256 * 16 * sizeof(long long), s));
    p.dbg = dbg_dev;
  }
  // two small pageable copies: the driver sends them inline with the command stream, which costs less GPU time than
  // one DMA from page-locked staging (measured: 8 us per step)
  BOGP_CUDA(h2d_async(r_dev, r_host, Nr * sizeof(double), s));
  BOGP_CUDA(h2d_async(z_dev, z_host, Nz * sizeof(double), s));
  const int nE = d.F * p.nrt * (kTileR / kTabRows), nB = d.F * ((Nz + kTabDepths - 1) / kTabDepths);
  umma_tables_kernel<<<nE + nB, kTabThreads, 0, s>>>(d, p, r_dev, z_dev, nE);
  count_launch();
  BOGP_CUDA(cudaGetLastError());
  const size_t smem = (size_t)p.smem_bar + 256;
  BOGP_CUDA(cudaFuncSetAttribute(umma_grid_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t ev = timing_begin(s);
  umma_grid_kernel<false><<<std::min(p.units, sms), kUmmaThreads, smem, s>>>(p);
  timing_end(ev, s);
  count_launch();
  BOGP_CUDA(cudaGetLastError());
  if (p.dbg) {
    std::vector<long long> hdbg(256 * 16);
    BOGP_CUDA(cudaMemcpyAsync(hdbg.data(), dbg_dev, hdbg.size() * sizeof(long long), cudaMemcpyDeviceToHost, s));
    BOGP_CUDA(cudaStreamSynchronize(s));
    const int grid = std::min(p.units, sms);
    const char* names[16] = {"-", "-", "-", "mma_total", "mma_wait_efull",
                             "mma_wait_bfull", "mma_wait_tempty", "stages", "epi0_total", "epi0_wait_tfull",
                             "epi4_total", "epi4_wait_tfull", "load_total", "-", "-", "-"};
    for (int k = 0; k < 16; ++k) {
      double sum = 0; long long mx = 0;
      for (int b = 0; b < grid; ++b) { sum += (double)hdbg[b * 16 + k]; mx = std::max(mx, hdbg[b * 16 + k]); }
      fprintf(stderr, "[umma dbg] %-18s avg %12.0f max %12lld\n", names[k], sum / grid, mx);
    }
    fprintf(stderr, "[umma dbg] tonal start deltas (CTA 0, issuer 0):");
    for (int f = 1; f < d.F && f < 32; ++f) fprintf(stderr, " %lld", hdbg[16 * 200 + f] - hdbg[16 * 200 + f - 1]);
    fprintf(stderr, "\n[umma dbg] epilogue cycles per stage by tonal (CTA 0, warp 0):");
    for (int f = 0; f < d.F && f < 32; ++f) fprintf(stderr, " %lld", hdbg[16 * 201 + f] / std::max<long long>(1, hdbg[16 * 203 + f]));
    fprintf(stderr, "\n");
    for (int w = 0; w < 3; ++w) {
      fprintf(stderr, "[umma dbg] window %d: stage | ready->issued | issued->tfull-wake | hold (wake->release) | release->done | release->next ready(same buf)\n", w);
      for (int i = 0; i < 18; ++i) {
        const long long* e = hdbg.data() + 3360;
        const int k = w * 20 + i;
        fprintf(stderr, "   %3d  %6lld %6lld %6lld %6lld %6lld   (t0 delta %lld)\n", i, e[60 + k] - e[k], e[120 + k] - e[60 + k], e[180 + k] - e[120 + k],
                e[240 + k] - e[180 + k], e[k + 2] - e[180 + k], e[k + 1] - e[k]);
      }
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ tilted-array engine
bool umma_tilt_supported(const bogp_modeset_s* h) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10 || !h->has_cov) return false;
  const ModesetDev& d = h->dev;
  if (d.N > 64 || h->max_J > 4 || h->max_Mp > 64) return false;
  return true;
}

// Grid with a tilt axis on the tensor cores: out[Nt, Nz, Nr].  Units = (range tile, depth, chunk of tilts).
int launch_grid_umma_tilt(bogp_modeset_s* h, const double* r_host, int Nr, const double* z_host, int Nz,
                          const double* tilt_host, int Nt, int atype, int mf, float* out_dev, cudaStream_t s,
                          int ext_mode, float* ext_val_dev, long long* ext_idx_dev) {
  const ModesetDev& d = h->dev;
  const int sms = device_sms();
  UParams p{};
  const int Ntp = (Nt + 1) & ~1;                       // tilts padded to even: both halves of a pair share a stage quad
  int TC = std::min(Ntp, 32);                          // tilts per unit (even)
  // balance: enough units for every SM, but as few A-operand rebuilds as possible
  const int nrt = (Nr + kTileR - 1) / kTileR;
  while (TC > 2 && (long long)nrt * Nz * ((Ntp + TC - 1) / TC) < 2ll * sms) TC = std::max(2, (TC / 2) & ~1);
  p.F = d.F; p.Nr = Nr; p.Nzr = Nz; p.Nt = Nt; p.Nz = 2 * Ntp; p.ZC = 2 * TC; p.nrt = nrt;
  p.ntc = (Ntp + TC - 1) / TC;
  p.units = nrt * Nz * p.ntc;
  p.atype = atype; p.mf = mf; p.N = d.N;
  p.AR = 256; p.NA = 64;
  unsigned e_off = 0;
  unsigned long long b_off = 0;
  int sumMp = 0, maxMp = 0;
  int order[BOGP_MAX_TONALS];
  tonal_order(d, order);
  // FP16-split operands (default; BOGP_UMMA_KIND=tf32 keeps 3xTF32): same three products per MAC at the kind::f16 rate,
  // operands scaled by per-tonal powers of two (the Bartlett value is a ratio: the scales cancel)
  bool f16 = true;
  if (const char* e = std::getenv("BOGP_UMMA_KIND")) f16 = e[0] != 't';
  if ((int)h->u16.k_host.size() != d.F || (int)h->u16.phi_max.size() != d.F) f16 = false;
  p.f16 = f16 ? 1 : 0;
  double rmin = r_host[0], rmax = r_host[0];
  const double lmax = h->u16.lever_max;
  for (int i = 1; i < Nr; ++i) { rmin = std::min(rmin, r_host[i]); rmax = std::max(rmax, r_host[i]); }
  auto pow2 = [](double bound) { return (bound > 0.0 && std::isfinite(bound)) ? (float)std::ldexp(1.0, (int)std::floor(std::log2(16384.0 / bound))) : 1.f; };
  for (int f = 0; f < d.F; ++f) {
    const TonalMeta& t = d.t[order[f]];
    UTonal& u = p.t[f];
    u.src = order[f];
    u.Mp = t.Mp; u.rowsP = 64; u.nj = 1; u.nq = 0; u.n1 = 0; u.trK = t.trK;
    u.kc = f16 ? pad_to(t.M, 16) / 2 : t.Mp;
    u.e_bytes = 2048u * (unsigned)u.Mp; u.e_off = e_off; e_off += u.e_bytes;
    u.b_zbytes = f16 ? 64u * (unsigned)(2 * u.kc) * 2u : 64u * (unsigned)u.Mp * 4u;
    u.b_off = b_off; b_off += (unsigned long long)p.Nz * u.b_zbytes;
    u.offM = sumMp; sumMp += t.Mp;
    u.offC = t.offC; u.J = t.J;
    maxMp = std::max(maxMp, t.Mp);
    u.sA = u.sB = 1.f;
    if (f16) {
      double be = 0.0, kim = 0.0, pr = 0.0;
      for (const cplx& k : h->u16.k_host[order[f]]) {
        be = std::max(be, std::exp(std::max(k.imag() * rmin, k.imag() * rmax)) / std::sqrt(std::abs(k) * std::max(rmin * 0.5, 1e-3)));
        kim = std::max(kim, std::fabs(k.imag()));
      }
      for (const cplx& v : h->phi_rec[order[f]]) pr = std::max(pr, std::abs(v));
      u.sA = pow2(be * (double)h->u16.phi_max[order[f]]);      // |E phi|: E bounded at half the smallest range (receivers of a tilted array sit closer)
      u.sB = pow2(pr * std::exp(kim * lmax));                   // |Phi e^{i k l sin(tau)}|
    }
  }
  p.sumMp = sumMp;
  p.e_rt_stride = e_off;
  const unsigned acc_bytes = (unsigned)TC * kTileR * 4u;
  const unsigned aux_bytes = (64u * 4u + (unsigned)d.F * 256u * 8u + 2u * 2u * 4u * 9u * 32u * 4u + (unsigned)d.F * 32u * 16u + 127u) & ~127u;
  unsigned stage = (kSmemLimit - acc_bytes - aux_bytes - 256u) / kStages;
  stage = std::min(stage & ~127u, 32u * 1024u);
  for (int f = 0; f < d.F; ++f)
    if (2u * p.t[f].b_zbytes > stage || 4 * p.t[f].kc > 256) { set_error("UMMA tilt engine: mode set does not fit"); return BOGP_ERR_UNSUPPORTED; }
  p.stage_bytes = stage; p.smem_B = 0; p.smem_acc = kStages * stage; p.smem_aux = p.smem_acc + acc_bytes;
  p.smem_bar = p.smem_aux + aux_bytes;
  // scratch: [r | z | tilt doubles][err][phiz][sin][rinv][lever][E table][Bhi][Blo]
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t off_err = al(((size_t)Nr + Nz + Nt) * sizeof(double));
  const size_t off_phi = off_err + 256;
  const size_t off_phi_im = al(off_phi + (size_t)Nz * sumMp * 4);
  const size_t off_sin = al(off_phi_im + (d.src_complex ? (size_t)Nz * sumMp * 4 : 0));
  const size_t off_rinv = al(off_sin + (size_t)Ntp * 4);
  const size_t off_lev = al(off_rinv + (size_t)Nr * 4);
  const size_t off_E = al(off_lev + 64 * 4);
  const size_t off_Bhi = al(off_E + (size_t)nrt * e_off);
  const size_t off_Blo = al(off_Bhi + b_off);
  if (int rc = ensure_scratch(h, off_Blo + b_off + 256)) return rc;
  unsigned char* base = static_cast<unsigned char*>(h->grid_scratch);
  double* r_dev = reinterpret_cast<double*>(base);
  double* z_dev = r_dev + Nr;
  double* t_dev = z_dev + Nz;
  p.err = reinterpret_cast<int*>(base + off_err);
  if (ext_mode && (long long)Nr * Nz * Nt < (1ll << 32)) {
    p.ext_mode = ext_mode;
    p.ext_key = reinterpret_cast<unsigned long long*>(base + off_err + 16);
    p.ext_done = reinterpret_cast<unsigned int*>(base + off_err + 32);
    p.ext_val = ext_val_dev; p.ext_idx = ext_idx_dev;
  }
  float* phiz = reinterpret_cast<float*>(base + off_phi);
  float* sin_t = reinterpret_cast<float*>(base + off_sin);
  float* rinv = reinterpret_cast<float*>(base + off_rinv);
  float* lever = reinterpret_cast<float*>(base + off_lev);
  float* phiz_im = d.src_complex ? reinterpret_cast<float*>(base + off_phi_im) : nullptr;
  p.phiz = phiz; p.phiz_im = phiz_im; p.sin_t = sin_t; p.rinv = rinv; p.lever = lever;
  p.CkT_re = d.CkT_re; p.CkT_im = d.CkT_im;
  p.E = base + off_E; p.Bhi = base + off_Bhi; p.Blo = base + off_Blo;
  p.out = out_dev;
  BOGP_CUDA(h2d_async(r_dev, r_host, Nr * sizeof(double), s));
  BOGP_CUDA(h2d_async(z_dev, z_host, Nz * sizeof(double), s));
  BOGP_CUDA(h2d_async(t_dev, tilt_host, Nt * sizeof(double), s));
  const int nE = d.F * nrt * (kTileR / kTabRows);
  {
    UParams pe = p;              // the E blocks of the shared table kernel read Nr / nrt / e offsets only
    umma_tables_kernel<<<nE, kTabThreads, 0, s>>>(d, pe, r_dev, z_dev, nE);
  }
  umma_tilt_theta_kernel<<<dim3((unsigned)p.Nz, (unsigned)d.F), 128, 0, s>>>(d, p, t_dev);
  umma_tilt_aux_kernel<<<std::max(1, std::min(296, (Nz * sumMp + 255) / 256)), 256, 0, s>>>(d, p, r_dev, z_dev, t_dev, phiz, phiz_im, sin_t, rinv, lever);
  count_launch(3);
  BOGP_CUDA(cudaGetLastError());
  const size_t smem = (size_t)p.smem_bar + 256;
  auto kern = p.f16 ? umma_grid_kernel<true, true> : umma_grid_kernel<true, false>;
  BOGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t ev = timing_begin(s);
  kern<<<std::min(p.units, sms), kUmmaThreads, smem, s>>>(p);
  timing_end(ev, s);
  count_launch();
  BOGP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace bogp
