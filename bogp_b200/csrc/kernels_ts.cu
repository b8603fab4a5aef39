// Thompson sampling over a candidate set: one joint draw f ~ N(mu, Sigma) of the GP posterior at b candidates and its
// arg-max.  This is what BAxUS's create_candidate does through botorch's MaxPosteriorSampling
// (reference: projects/swellex96_inv/data/bo/baxus.py:121-141, n_candidates = min(5000, max(2000, 200 d)) :45):
//     mu    = m + K_*^T K^-1 (y - m),     Sigma = K_** - K_*^T K^-1 K_*        (noise-free posterior, float64)
//     L L^T = Sigma (+ jitter 1e-8, 1e-7, 1e-6 only if the factorisation fails: gpytorch psd_safe_cholesky [recalled])
//     f     = mu + L z,  z ~ N(0, I) supplied by the caller;  X_next = X_cand[argmax f]
// Everything is float64; the level-3 work runs on the FP64 tensor cores:
//     ts_kernel_matrix   K_*^T [b][n] and K_** [b][b] (lower tiles), 64 x 64 tiles, input dimensions streamed through smem
//     ts_gemm_nt         C (-)= A B^T, 64 x 64 tile / CTA on the FP64 tensor cores (mma.sync m8n8k4): whitening V^T = K_*^T Linv^T (triangular K
//                        range), Sigma -= V^T V (lower tiles), and the two level-3 steps of the blocked Cholesky
//     (diagonal blocks)  64 x 64 Cholesky + inverse in shared memory, by the CTA that has just updated the block (the
//                        factorisation of block k+1 hides behind the rest of trailing update k), failure flag
//     ts_mean / ts_sample / ts_argmax   mu, f = mu + L z (warp per row), arg-max without replacement over samples
// All matrices are padded (b to 64, n to 64) with zeros / an identity diagonal, so no kernel has edge predicates.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>

#include "bogp_internal.h"

namespace bogp {

constexpr int kTsT = 64;          // tile edge = Cholesky panel width
constexpr int kTsKC = 16;         // K chunk of the GEMM
constexpr int kTsThreads = 256;
constexpr int kTsMaxB = 16384;
constexpr int kTsMaxS = 64;

__device__ __forceinline__ double ts_kernel_value(int kind, double os, double d2) {
  if (kind == BOGP_KERNEL_RBF) return os * exp(-0.5 * d2);
  const double sqrt5 = 2.23606797749978969641;
  const double r = sqrt(fmax(d2, 1e-30));
  return os * (1.0 + sqrt5 * r + (5.0 / 3.0) * d2) * exp(-sqrt5 * r);
}

// lower-triangular tile index t -> (ti, tj), tj <= ti
__device__ __forceinline__ void tri_decode(int t, int& ti, int& tj) {
  int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((long long)(i + 1) * (i + 2) / 2 <= t) ++i;
  while ((long long)i * (i + 1) / 2 > t) --i;
  ti = i;
  tj = t - (int)((long long)i * (i + 1) / 2);
}

// out[i][j] = k(P_i, Q_j); rows i >= Pvalid / columns j >= Qvalid are padding: 0, or (lower mode) 1 on the diagonal.
// lower: P == Q, only tiles tj <= ti are written, `diag_add` is added to the diagonal (jitter).
struct KmArgs {
  const double *P, *Q, *inv_ls;
  double* out;
  int ldo, d, Pvalid, Qvalid, kind, lower, tiles_n;
  double os, diag_add;
};

__global__ void __launch_bounds__(kTsThreads) ts_kernel_matrix(KmArgs a) {
  __shared__ double Ps[kTsKC][kTsT], Qs[kTsKC][kTsT];
  int ti, tj;
  if (a.lower) tri_decode(blockIdx.x, ti, tj);
  else { ti = blockIdx.x / a.tiles_n; tj = blockIdx.x % a.tiles_n; }
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lr = tid & 63, lk = (tid >> 6) * 4;          // loader: row lr of the tile, dimensions lk .. lk+3 of the chunk
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  const int pi = ti * kTsT + lr, qj = tj * kTsT + lr;
  for (int a0 = 0; a0 < a.d; a0 += kTsKC) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int dim = a0 + lk + e;
      const double w = dim < a.d ? __ldg(a.inv_ls + dim) : 0.0;
      Ps[lk + e][lr] = (dim < a.d && pi < a.Pvalid) ? __ldg(a.P + (size_t)pi * a.d + dim) * w : 0.0;
      Qs[lk + e][lr] = (dim < a.d && qj < a.Qvalid) ? __ldg(a.Q + (size_t)qj * a.d + dim) * w : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTsKC; ++k) {
      double p[4], q[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { p[i] = Ps[k][ty * 4 + i]; q[i] = Qs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double df = p[i] - q[j]; acc[i][j] = fma(df, df, acc[i][j]); }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = ti * kTsT + ty * 4 + i;
    double v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = tj * kTsT + tx * 4 + j;
      const bool valid = gi < a.Pvalid && gj < a.Qvalid;
      v[j] = valid ? ts_kernel_value(a.kind, a.os, acc[i][j]) : 0.0;
      if (a.lower && gi == gj) v[j] = valid ? v[j] + a.diag_add : 1.0;
    }
    double2* o = reinterpret_cast<double2*>(a.out + (size_t)gi * a.ldo + tj * kTsT + tx * 4);
    o[0] = make_double2(v[0], v[1]);
    o[1] = make_double2(v[2], v[3]);
  }
}

// 1 / sqrt(x) in float64 from the float approximation and two Newton steps (a dependent DDIV + DSQRT per column costs
// more than the column's whole rank-1 update)
__device__ __forceinline__ double rsqrt_d(double x) {
  if (!(x > 1e-30 && x < 1e30)) return 1.0 / sqrt(x);
  double y = (double)rsqrtf((float)x);
  y = y * fma(-0.5 * x, y * y, 1.5);
  y = y * fma(-0.5 * x, y * y, 1.5);
  return y;
}

// Diagonal block of a panel: D = L L^T in place (lower), Tinv = L^-1 (dense 64 x 64, zero upper) for the panel solve.
// Blocked in shared memory (16-wide sub-blocks): the factorisation is a chain of dependent column steps, so the work per
// step is kept to the sub-block (<= 4 predicated slots per thread) and everything else is done as small matrix products
// between barriers -- 64 short column steps + 3 trailing updates, then 15 substitution steps on the four diagonal
// sub-blocks at once + 3 x 2 product phases for the off-diagonal blocks of the inverse.
constexpr int kDiagLd = kTsT + 1;

__device__ __forceinline__ void diag_factor_invert(double (*D)[kDiagLd], double (*X)[kDiagLd], double* invd,
                                                   double (*Tm)[16][17], int tid, int* flag,
                                                   long long* stamps = nullptr) {
#define TS_STAMP(i) do { if (stamps && tid == 0) stamps[i] = clock64(); } while (0)
  TS_STAMP(0);
  // A dependent FP64 operation costs ~40 cycles here (measured: tools/micro/diag_time.cu), so the column chain is kept
  // to: pivot -> float reciprocal seed -> one cubic correction (3 DFMA) -> rank-1 update.  Scaling the columns by
  // 1 / sqrt(pivot) is off that chain (one parallel pass per sub-block), and everything that is a sum runs as several
  // independent accumulators.
  const int r = tid >> 2, cp = tid & 3;
  double* pivs = &Tm[0][0][0];                 // pivots of the current sub-block (Tm is free until the inversion)
  for (int e = tid; e < kTsT * kTsT; e += kTsThreads) X[e >> 6][e & 63] = 0.0;
  for (int sb = 0; sb < 4; ++sb) {
    const int j0 = 16 * sb, j1 = j0 + 16;
    TS_STAMP(4 + 4 * sb);
    for (int j = j0; j < j1; ++j) {
      double piv = D[j][j];
      const bool bad = !(piv > 0.0) || !isfinite(piv);
      if (bad) piv = 1.0;
      const double rj = r > j ? D[r][j] : 0.0;
      double w;
      if (piv > 1e-30 && piv < 1e30) {
        const double y0 = (double)__frcp_rn((float)piv);       // 1 / piv to ~1e-7
        const double w0 = -rj * y0;
        const double err = fma(-piv, y0, 1.0);                 // 1 / piv = y0 (1 + err + err^2 + ...)
        w = fma(w0, fma(err, err, err), w0);
      } else {
        w = -rj / piv;
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int c = j + 1 + cp + 4 * m;
        if (c < j1 && c <= r) D[r][c] = fma(w, D[c][j], D[r][c]);
      }
      if (r == j && cp == 0) { pivs[j - j0] = piv; if (bad) *flag = 1; }
      __syncthreads();
    }
    TS_STAMP(5 + 4 * sb);
    // scale the sub-block's columns: L[r][j] = D[r][j] / sqrt(piv_j)
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int j = j0 + cp + 4 * m;
      const double pv = pivs[j - j0], rs = rsqrt_d(pv);
      if (r > j) D[r][j] *= rs;
      if (r == j) { D[j][j] = pv * rs; invd[j] = rs; }
    }
    __syncthreads();
    TS_STAMP(6 + 4 * sb);
    if (j1 < kTsT) {      // trailing update of the lower triangle right of the sub-block, <= 5 elements per thread at once
      const int nt = kTsT - j1, total = nt * (nt + 1) / 2;
      int rr[5], cc[5];
      double acc[5];
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        // triangle folded into an (nt / 2) x (nt + 1) rectangle: row p of the rectangle = row p of the triangle
        // (p + 1 elements) followed by row nt - 1 - p (nt - p elements)
        const int e = tid + kTsThreads * m;
        int a = 0, bq = 0;
        if (e < total) {
          const int p = nt == 48 ? e / 49 : (nt == 32 ? e / 33 : e / 17), q = e - p * (nt + 1);
          if (q <= p) { a = p; bq = q; } else { a = nt - 1 - p; bq = q - (p + 1); }
        }
        rr[m] = j1 + a; cc[m] = j1 + bq;
        acc[m] = D[rr[m]][cc[m]];
      }
#pragma unroll 4
      for (int j = 0; j < 16; ++j) {
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[m] = fma(-D[rr[m]][j0 + j], D[cc[m]][j0 + j], acc[m]);
      }
#pragma unroll
      for (int m = 0; m < 5; ++m)
        if (tid + kTsThreads * m < total) D[rr[m]][cc[m]] = acc[m];
      __syncthreads();
    }
  }
  TS_STAMP(1);
  // X = L^-1.  Diagonal sub-blocks: forward substitution, 4 blocks x 16 columns x 4 lanes at once
  {
    const int b0 = 16 * (tid >> 6), cc = b0 + ((tid >> 2) & 15), part = tid & 3;
    if (part == 0) X[cc][cc] = invd[cc];
    __syncwarp();
    for (int i = 1; i < 16; ++i) {
      const int gi = b0 + i;
      double s0 = 0.0, s1 = 0.0;
      if (gi > cc) {
        const int kk = cc + part;
        if (kk < gi) s0 = D[gi][kk] * X[kk][cc];
        if (kk + 4 < gi) s1 = D[gi][kk + 4] * X[kk + 4][cc];
        if (kk + 8 < gi) s0 = fma(D[gi][kk + 8], X[kk + 8][cc], s0);
        if (kk + 12 < gi) s1 = fma(D[gi][kk + 12], X[kk + 12][cc], s1);
      }
      double sum = s0 + s1;
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      if (part == 0 && gi > cc) X[gi][cc] = -sum * invd[gi];
      __syncwarp();
    }
  }
  __syncthreads();
  TS_STAMP(2);
  // off-diagonal blocks by distance from the diagonal: X_ij = -X_ii (sum_{k=j}^{i-1} L_ik X_kj); X is zero above its
  // diagonal, so the sums run over whole 16-blocks in four independent accumulators
  for (int dd = 1; dd < 4; ++dd) {
    const int nblk = 4 - dd;
    for (int e = tid; e < nblk * 256; e += kTsThreads) {
      const int bj = e >> 8, bi = bj + dd, pp = (e >> 4) & 15, q = e & 15;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      for (int kk = 16 * bj; kk < 16 * bi; kk += 4) {
        a0 = fma(D[16 * bi + pp][kk], X[kk][16 * bj + q], a0);
        a1 = fma(D[16 * bi + pp][kk + 1], X[kk + 1][16 * bj + q], a1);
        a2 = fma(D[16 * bi + pp][kk + 2], X[kk + 2][16 * bj + q], a2);
        a3 = fma(D[16 * bi + pp][kk + 3], X[kk + 3][16 * bj + q], a3);
      }
      Tm[bj][pp][q] = (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
    for (int e = tid; e < nblk * 256; e += kTsThreads) {
      const int bj = e >> 8, bi = bj + dd, pp = (e >> 4) & 15, q = e & 15;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int m = 0; m < 16; m += 4) {       // X_ii is zero above its diagonal
        a0 = fma(X[16 * bi + pp][16 * bi + m], Tm[bj][m][q], a0);
        a1 = fma(X[16 * bi + pp][16 * bi + m + 1], Tm[bj][m + 1][q], a1);
        a2 = fma(X[16 * bi + pp][16 * bi + m + 2], Tm[bj][m + 2][q], a2);
        a3 = fma(X[16 * bi + pp][16 * bi + m + 3], Tm[bj][m + 3][q], a3);
      }
      X[16 * bi + pp][16 * bj + q] = -((a0 + a1) + (a2 + a3));
    }
    __syncthreads();
  }
  TS_STAMP(3);
#undef TS_STAMP
}

// C[i][j] = A B^T (accumulate 0) or C[i][j] -= A B^T (accumulate 1); A [M][K], B [N][K], all dimensions padded:
// M, N multiples of 64, K of 16.  lower: square problem, tiles tj <= ti only.  ktri: B is lower triangular with
// K == N: the K loop of tile column tj stops after its last row.  A may alias C when N == 64 (a CTA owns whole rows
// and has read all of A before it writes).
struct GemmArgs {
  const double *A, *B;
  double* C;
  int lda, ldb, ldc, K, tiles_n, lower, accumulate, ktri;      // ktri 1: B lower triangular (K stops after tile column tj); 2: B upper triangular (K starts at tile column tj)
  // fuse_diag: the CTA of tile (0, 0) -- the next diagonal block of the blocked Cholesky -- factors and inverts it
  // right after updating it, while the other CTAs are still on the rest of the trailing matrix
  int fuse_diag;
  double* Tinv;
  int* flag;
};

// FP64 tensor-core MMA (DMMA): D[8x8] += A[8x4] B[4x8]; lane (g = lane / 4, t = lane % 4) holds A[g][t], B[t][g] and
// D[g][2t], D[g][2t + 1].  Measured on B200 (tools/micro/fp64_peak.cu): 37.1 TFLOP/s against 33.9 for DFMA, and -- what
// matters here -- operand fragments are spread over the warp instead of broadcast, so a k-step costs 6 LDS.64 per
// 2048 FMA where the 4 x 4 DFMA register tile needed 4 LDS.128 per 512 (shared-memory bound at half the DFMA peak).
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

constexpr int kGemmSmemDoubles = 2 * 2 * kTsKC * 68;      // As | Bs, double-buffered
constexpr size_t kGemmSmem = kGemmSmemDoubles * sizeof(double);
constexpr size_t kGemmDiagSmem = kGemmSmem + (size_t)(kTsT * kDiagLd + kTsT + 3 * 16 * 17) * sizeof(double);
static_assert(kTsT * kDiagLd <= kGemmSmemDoubles, "the diagonal block aliases the operand buffers");
constexpr int kTsLd = 68;      // shared-memory row stride in doubles: the fragment loads (4 k x 8 rows) hit 32 distinct banks

__global__ void __launch_bounds__(kTsThreads) ts_gemm_nt(GemmArgs g) {
  extern __shared__ __align__(16) double ts_smem[];
  double (*As)[kTsKC][kTsLd] = reinterpret_cast<double (*)[kTsKC][kTsLd]>(ts_smem);
  double (*Bs)[kTsKC][kTsLd] = As + 2;
  int ti, tj;
  if (g.lower) tri_decode(blockIdx.x, ti, tj);
  else { ti = blockIdx.x / g.tiles_n; tj = blockIdx.x % g.tiles_n; }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3, wm = (warp >> 2) * 32, wn = (warp & 3) * 16;    // warp tile 32 x 16
  const int lr = tid & 63, lk = (tid >> 6) * 4;
  const double* Ap = g.A + (size_t)(ti * kTsT + lr) * g.lda + lk;
  const double* Bp = g.B + (size_t)(tj * kTsT + lr) * g.ldb + lk;
  const int K = g.ktri == 1 ? min(g.K, (tj + 1) * kTsT) : g.K;
  const int Kfirst = g.ktri == 2 ? tj * kTsT : 0;
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double2 a0 = *reinterpret_cast<const double2*>(Ap + Kfirst), a1 = *reinterpret_cast<const double2*>(Ap + Kfirst + 2);
  double2 b0 = *reinterpret_cast<const double2*>(Bp + Kfirst), b1 = *reinterpret_cast<const double2*>(Bp + Kfirst + 2);
  int buf = 0;
  for (int k0 = Kfirst; k0 < K; k0 += kTsKC) {
    As[buf][lk][lr] = a0.x; As[buf][lk + 1][lr] = a0.y; As[buf][lk + 2][lr] = a1.x; As[buf][lk + 3][lr] = a1.y;
    Bs[buf][lk][lr] = b0.x; Bs[buf][lk + 1][lr] = b0.y; Bs[buf][lk + 2][lr] = b1.x; Bs[buf][lk + 3][lr] = b1.y;
    __syncthreads();
    if (k0 + kTsKC < K) {      // next chunk in flight while this one is multiplied
      a0 = *reinterpret_cast<const double2*>(Ap + k0 + kTsKC); a1 = *reinterpret_cast<const double2*>(Ap + k0 + kTsKC + 2);
      b0 = *reinterpret_cast<const double2*>(Bp + k0 + kTsKC); b1 = *reinterpret_cast<const double2*>(Bp + k0 + kTsKC + 2);
    }
#pragma unroll
    for (int k4 = 0; k4 < kTsKC; k4 += 4) {
      double a[4], b[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[buf][k4 + tq][wm + 8 * i + gq];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = Bs[buf][k4 + tq][wn + 8 * j + gq];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(acc[i][j], a[i], b[j]);
    }
    buf ^= 1;      // the other buffer was last read before the barrier above: safe to fill without a second barrier
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double2* c = reinterpret_cast<double2*>(g.C + (size_t)(ti * kTsT + wm + 8 * i + gq) * g.ldc + tj * kTsT + wn + 8 * j + 2 * tq);
      if (g.accumulate) {
        const double2 c0 = *c;
        *c = make_double2(c0.x - acc[i][j][0], c0.y - acc[i][j][1]);
      } else {
        *c = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
  if (g.fuse_diag && blockIdx.x == 0) {
    __syncthreads();      // the tile is complete in global memory and the operand buffers are free
    double (*D)[kDiagLd] = reinterpret_cast<double (*)[kDiagLd]>(ts_smem);            // aliases As / Bs
    double (*X)[kDiagLd] = reinterpret_cast<double (*)[kDiagLd]>(ts_smem + kGemmSmemDoubles);
    double* invd = reinterpret_cast<double*>(X + kTsT);
    double (*Tm)[16][17] = reinterpret_cast<double (*)[16][17]>(invd + kTsT);
    for (int e = tid; e < kTsT * kTsT; e += kTsThreads) D[e >> 6][e & 63] = g.C[(size_t)(e >> 6) * g.ldc + (e & 63)];
    __syncthreads();
    diag_factor_invert(D, X, invd, Tm, tid, g.flag);
    for (int e = tid; e < kTsT * kTsT; e += kTsThreads) {
      const int r = e >> 6, c = e & 63;
      if (c <= r) g.C[(size_t)r * g.ldc + c] = D[r][c];
      g.Tinv[e] = c <= r ? X[r][c] : 0.0;
    }
  }
}

// Linv / alpha of the model into the padded workspace
__global__ void ts_pad_kernel(const double* Linv, const double* alpha, int n, int np, double* Linvp, double* alphap) {
  const size_t total = (size_t)np * np;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / np), j = (int)(e % np);
    Linvp[e] = (i < n && j < n) ? Linv[(size_t)i * n + j] : 0.0;
    if (j == 0) alphap[i] = i < n ? alpha[i] : 0.0;
  }
}

// mu_c = m + Kt[c] . alpha   (warp per candidate)
__global__ void ts_mean_kernel(const double* Kt, int np, const double* alphap, double mean, int b, double* mu) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= b) return;
  double s = 0.0;
  for (int j = lane; j < np; j += 32) s = fma(Kt[(size_t)w * np + j], alphap[j], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) mu[w] = mean + s;
}

// f[s][c] = mu_c + sum_{j <= c} L[c][j] z[s][j]   (warp per candidate row)
__global__ void ts_sample_kernel(const double* L, int ld, const double* mu, const double* z, int b, int S, double* f) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= b) return;
  for (int s = 0; s < S; ++s) {
    double acc = 0.0;
    for (int j = lane; j <= w; j += 32) acc = fma(L[(size_t)w * ld + j], z[(size_t)s * b + j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) f[(size_t)s * b + w] = mu[w] + acc;
  }
}

// arg-max of each sample in turn; a candidate picked by an earlier sample is excluded (replacement = False)
__global__ void __launch_bounds__(1024) ts_argmax_kernel(const double* f, int b, int S, long long* idx, double* val) {
  __shared__ double sv[32];
  __shared__ int si[32];
  __shared__ int chosen[kTsMaxS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int s = 0; s < S; ++s) {
    double bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = tid; c < b; c += blockDim.x) {
      bool taken = false;
      for (int t = 0; t < s; ++t) taken |= chosen[t] == c;
      const double v = f[(size_t)s * b + c];
      if (!taken && (v > bv || (v == bv && c < bi))) { bv = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = sv[lane]; bi = si[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { chosen[s] = bi; idx[s] = bi; val[s] = bv; }
    }
    __syncthreads();
  }
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// ---- building blocks for the large-batch posterior / acquisition path (kernels_gp.cu): everything padded to 64
int ts_pad(const bogp_gp_s* g, int np, double* Linv_p, double* alpha_p, cudaStream_t s) {
  ts_pad_kernel<<<std::min(1184, (np * np + 255) / 256), 256, 0, s>>>(g->Linv, g->alpha, g->n, np, Linv_p, alpha_p);
  count_launch();
  BOGP_CUDA(cudaGetLastError());
  return 0;
}
// Kt[bp][np] = k(Xc_i, X_j) (zero outside the b x n block)
int ts_cross_kernel(const bogp_gp_s* g, const double* Xc_dev, int b, int bp, int np, double* Kt, cudaStream_t s) {
  KmArgs km{};
  km.P = Xc_dev; km.Q = g->X; km.inv_ls = g->inv_ls; km.out = Kt; km.ldo = np; km.d = g->d; km.Pvalid = b; km.Qvalid = g->n;
  km.kind = g->kind; km.lower = 0; km.tiles_n = np / kTsT; km.os = g->outputscale; km.diag_add = 0.0;
  ts_kernel_matrix<<<(bp / kTsT) * (np / kTsT), kTsThreads, 0, s>>>(km);
  count_launch();
  BOGP_CUDA(cudaGetLastError());
  return 0;
}
// C[bp][np] = A[bp][np] B^T with B[np][np] lower (tri = 1) or upper (tri = 2) triangular, on the FP64 tensor cores
int ts_gemm_tri(const double* A, const double* B, double* C, int bp, int np, int tri, cudaStream_t s) {
  GemmArgs wh{};
  wh.A = A; wh.lda = np; wh.B = B; wh.ldb = np; wh.C = C; wh.ldc = np; wh.K = np; wh.tiles_n = np / kTsT; wh.ktri = tri;
  BOGP_CUDA(cudaFuncSetAttribute(ts_gemm_nt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmDiagSmem));
  ts_gemm_nt<<<(bp / kTsT) * (np / kTsT), kTsThreads, kGemmSmem, s>>>(wh);
  count_launch();
  BOGP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace bogp

using namespace bogp;

extern "C" int bogp_gp_thompson(bogp_gp_t g, const double* Xc_dev, int b, const double* z_dev, int S,
                                long long* idx_dev, double* val_dev, double* samples_dev, double* mean_dev,
                                double* jitter_out, void* stream) {
  BOGP_REQUIRE(g && Xc_dev && z_dev && idx_dev && val_dev, "bogp_gp_thompson: NULL argument");
  BOGP_REQUIRE(b >= 1 && b <= kTsMaxB, "bogp_gp_thompson: b=%d outside [1,%d]", b, kTsMaxB);
  BOGP_REQUIRE(S >= 1 && S <= kTsMaxS && S <= b, "bogp_gp_thompson: num_samples=%d outside [1,min(%d,b)]", S, kTsMaxS);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = g->n, d = g->d;
  const int bp = (b + kTsT - 1) / kTsT * kTsT, np = (n + kTsT - 1) / kTsT * kTsT;
  const int tb = bp / kTsT, tn = np / kTsT;
  // workspace
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off = al256(off + bytes); return o; };
  const size_t o_kt = take((size_t)bp * np * 8), o_vt = take((size_t)bp * np * 8), o_li = take((size_t)np * np * 8);
  const size_t o_al = take((size_t)np * 8), o_sg = take((size_t)bp * bp * 8), o_ti = take((size_t)kTsT * kTsT * 8);
  const size_t o_mu = take((size_t)bp * 8), o_f = take((size_t)S * b * 8), o_flag = take(256);
  if (g->ts_bytes < off) {
    if (g->ts_ws) cudaFree(g->ts_ws);
    g->ts_ws = nullptr; g->ts_bytes = 0;
    BOGP_CUDA(cudaMalloc(&g->ts_ws, off));
    g->ts_bytes = off;
  }
  unsigned char* ws = static_cast<unsigned char*>(g->ts_ws);
  double *Kt = (double*)(ws + o_kt), *Vt = (double*)(ws + o_vt), *Li = (double*)(ws + o_li), *al = (double*)(ws + o_al);
  double *Sg = (double*)(ws + o_sg), *Ti = (double*)(ws + o_ti), *mu = (double*)(ws + o_mu);
  double* f = samples_dev ? samples_dev : (double*)(ws + o_f);
  int* flag = (int*)(ws + o_flag);

  ts_pad_kernel<<<std::min(1184, (np * np + 255) / 256), 256, 0, s>>>(g->Linv, g->alpha, n, np, Li, al);
  KmArgs km{};
  km.P = Xc_dev; km.Q = g->X; km.inv_ls = g->inv_ls; km.out = Kt; km.ldo = np; km.d = d; km.Pvalid = b; km.Qvalid = n;
  km.kind = g->kind; km.lower = 0; km.tiles_n = tn; km.os = g->outputscale; km.diag_add = 0.0;
  ts_kernel_matrix<<<tb * tn, kTsThreads, 0, s>>>(km);
  GemmArgs wh{};
  wh.A = Kt; wh.lda = np; wh.B = Li; wh.ldb = np; wh.C = Vt; wh.ldc = np; wh.K = np; wh.tiles_n = tn; wh.ktri = 1;
  ts_gemm_nt<<<tb * tn, kTsThreads, kGemmSmem, s>>>(wh);
  ts_mean_kernel<<<(b * 32 + 255) / 256, 256, 0, s>>>(Kt, np, al, g->mean, b, mu);
  count_launch(4);

  BOGP_CUDA(cudaFuncSetAttribute(ts_gemm_nt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmDiagSmem));
  const int tri = tb * (tb + 1) / 2;
  double jitter = 0.0;
  bool ok = false;
  cudaEvent_t ev = timing_begin(s);
  for (int attempt = 0; attempt < 4 && !ok; ++attempt) {
    jitter = attempt == 0 ? 0.0 : 1e-8 * std::pow(10.0, attempt - 1);
    BOGP_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
    KmArgs kk = km;
    kk.Q = Xc_dev; kk.Qvalid = b; kk.out = Sg; kk.ldo = bp; kk.lower = 1; kk.diag_add = jitter;
    ts_kernel_matrix<<<tri, kTsThreads, 0, s>>>(kk);
    GemmArgs sy{};
    sy.A = Vt; sy.lda = np; sy.B = Vt; sy.ldb = np; sy.C = Sg; sy.ldc = bp; sy.K = np; sy.lower = 1; sy.accumulate = 1;
    sy.fuse_diag = 1; sy.Tinv = Ti; sy.flag = flag;      // + diagonal block 0
    ts_gemm_nt<<<tri, kTsThreads, kGemmDiagSmem, s>>>(sy);
    count_launch(2);
    for (int k = 0; k + 1 < tb; ++k) {
      const int rem = tb - 1 - k;
      double* panel = Sg + (size_t)(k + 1) * kTsT * bp + (size_t)k * kTsT;
      GemmArgs tr{};
      tr.A = panel; tr.lda = bp; tr.B = Ti; tr.ldb = kTsT; tr.C = panel; tr.ldc = bp; tr.K = kTsT; tr.tiles_n = 1;
      ts_gemm_nt<<<rem, kTsThreads, kGemmSmem, s>>>(tr);
      GemmArgs up{};
      up.A = panel; up.lda = bp; up.B = panel; up.ldb = bp; up.C = Sg + (size_t)(k + 1) * kTsT * bp + (size_t)(k + 1) * kTsT;
      up.ldc = bp; up.K = kTsT; up.lower = 1; up.accumulate = 1;
      up.fuse_diag = 1; up.Tinv = Ti; up.flag = flag;    // + diagonal block k + 1, needed by the next panel solve
      ts_gemm_nt<<<rem * (rem + 1) / 2, kTsThreads, kGemmDiagSmem, s>>>(up);
      count_launch(2);
    }
    int hflag = 0;
    BOGP_CUDA(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    BOGP_CUDA(cudaStreamSynchronize(s));
    ok = hflag == 0;
  }
  timing_end(ev, s);
  if (!ok) {
    set_error("bogp_gp_thompson: posterior covariance not positive definite after jitter 1e-6");
    return BOGP_ERR_NUMERIC;
  }
  ts_sample_kernel<<<(b * 32 + 255) / 256, 256, 0, s>>>(Sg, bp, mu, z_dev, b, S, f);
  ts_argmax_kernel<<<1, 1024, 0, s>>>(f, b, S, idx_dev, val_dev);
  count_launch(2);
  if (mean_dev) BOGP_CUDA(cudaMemcpyAsync(mean_dev, mu, (size_t)b * 8, cudaMemcpyDeviceToDevice, s));
  BOGP_CUDA(cudaGetLastError());
  if (jitter_out) *jitter_out = jitter;
  return 0;
}
