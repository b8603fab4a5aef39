// Internal declarations shared by the translation units of libbogp_sm100.so.
#pragma once
#include <cuda_runtime.h>

#include <complex>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/bogp.h"

namespace bogp {

typedef std::complex<double> cplx;

void set_error(const char* fmt, ...);
// Number of kernels this library has launched (bench.py reports it as gpu_launches).
void count_launch(int n = 1);
// host -> device copy of a call's inputs on its stream, counted (bogp_h2d_bytes: what bench.py reports per step)
cudaError_t h2d_async(void* dst, const void* src, size_t bytes, cudaStream_t s);
// Event bracket around a dominant kernel (no-ops returning nullptr unless bogp_timing_enable(1)).
bool timing_enabled();
cudaEvent_t timing_begin(cudaStream_t s);
void timing_end(cudaEvent_t stop, cudaStream_t s);

#define BOGP_CUDA(call)                                                                      \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      bogp::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call,                  \
                      cudaGetErrorString(e__));                                              \
      return BOGP_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

#define BOGP_REQUIRE(cond, ...)                                                              \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      bogp::set_error(__VA_ARGS__);                                                          \
      return BOGP_ERR_INVALID;                                                               \
    }                                                                                        \
  } while (0)

inline int pad_to(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------- host linear algebra (float64)
// Pivoted Cholesky of a Hermitian PSD matrix A[n,n] (row-major): A ~= L L^H, L[n, rank] (row-major,
// leading dimension n).  Stops when the largest remaining diagonal <= tol * max initial diagonal.
int pivoted_cholesky(const std::vector<cplx>& A, int n, double tol, std::vector<cplx>& L, int& rank);
// Same contract from the eigen-decomposition (optimal rank for the tolerance: eigenvalues above tol * lambda_max).
int hermitian_eig_factor(const std::vector<cplx>& A, int n, double tol, std::vector<cplx>& L, int& rank);
// Dense Cholesky A = L L^T (real SPD, row-major, lower); returns false if not PD.
bool cholesky_lower(std::vector<double>& A, int n);
// In-place inverse of a lower-triangular matrix (row-major).
void invert_lower(std::vector<double>& L, int n);

// ---------------------------------------------------------------- device-side tables of one tonal
// All per-tonal arrays are concatenated; `off*` give the element offsets.
struct TonalMeta {
  int M;         // modes
  int Mp;        // modes padded to a multiple of 8 (zero-filled)
  int rows;      // rows of the mode-space operator W
  int rowsP;     // rows padded to a multiple of 32 (leading dimension of Wt)
  int n_plain;   // rows [0, n_plain): norm rows, real operator           T = W a
  int n_npair;   // next 2*n_npair rows: norm rows of a complex operator  T = (Wre + i Wim) a
  int n_qpair;   // next 2*n_qpair rows: numerator (projection) rows      T = (Cre + i Cim) a
  int J;         // rank of the receiver-space covariance factor
  int offM;      // offset into [sum Mp] arrays
  int offW;      // offset into W (rows * Mp floats)
  int offWt;     // offset into Wt (Mp * rowsP floats)
  int offTab;    // offset into phi_src table (nz_tab * Mp)
  int offRec;    // offset into phi_rec (N * Mp)
  int offC;      // offset into receiver-space factor (J * N complex)
  float trK;     // trace of K_f (cbf_ml)
  // tcgen05 (UMMA) row order of the same operator: the complex rows as ADJACENT (re, im) pairs -- numerator pairs
  // first, norm pairs after -- then the plain rows from the even row u_p0 on, zero-padded to u_rowsP (x16).
  int u_p0;      // first plain row (even) = 2 n_qpair + 2 n_npair
  int u_rowsP;   // rows padded to a multiple of 16 (the MMA N granularity at M = 128)
  int offWu;     // offset into Wu (u_rowsP * Mp floats, row-major [row][Mp])
  // Folded operator (real receiver modes, rank-1 covariance): the norm rows R are rotated by an orthogonal Q whose first
  // two rows span Re/Im of the numerator row written in R's row space (C = g^T R): ||Q R a|| = ||R a|| and
  // C a = u_c1 (Q R a)_0 + u_c2 (Q R a)_1, so the numerator costs no operator rows of its own (u_p0 = 0, rank(G) rows).
  int u_fold;
  float u_c[4];  // c1 (re, im), c2 (re, im)
  int u_rows;    // operator rows the tcgen05 engine multiplies by (before padding): the algorithmic row count
};

#define BOGP_MAX_TONALS 32

struct ModesetDev {
  int F, N, nz_tab;
  double z0, dz, z_pivot;
  bool src_complex, rec_complex;
  TonalMeta t[BOGP_MAX_TONALS];
  // [sum Mp]
  const double* k_re;
  const double* k_im;
  const float* c_re;   // k^{-1/2}
  const float* c_im;
  // [sum nz_tab*Mp]
  const float* tab_re;
  const float* tab_im;   // null unless src_complex
  // [sum N*Mp]
  const float* rec_re;
  const float* rec_im;   // null unless rec_complex
  const float* recT_re;  // transposed copies [Mp][N] (coalesced over receivers)
  const float* recT_im;
  const float* rec_lever;   // [N]  (z_pivot - z_n)
  // set by covariance_set
  const float* W;        // [sum rows*Mp]   row-major (UMMA operand generation)
  const float* Wt;       // [sum Mp*rowsP]  transposed (SIMT kernels)
  const float* Wu;       // [sum u_rowsP*Mp] UMMA row order
  const float* CkT_re;   // [sum N*J]       receiver-space factor, transposed [N][J]
  const float* CkT_im;
};

}  // namespace bogp

struct bogp_modeset_s {
  bogp::ModesetDev dev;           // device pointers + metadata (passed by value to kernels)
  std::vector<void*> allocs;      // every cudaMalloc owned by this handle
  std::vector<void*> cov_allocs;  // the five covariance-dependent device arrays (W, Wt, Wu, Ck re/im), reused while they fit
  std::vector<size_t> cov_capacity;
  // covariance-independent half of bogp_covariance_set, cached: pivoted Cholesky factor of G_f = Phi_f^H Phi_f
  double rank_tol = 0.0;                 // relative eigenvalue below which directions of G = Phi^H Phi are dropped (0: pivoted Cholesky)
  std::vector<std::vector<bogp::cplx>> LG_cache;
  std::vector<int> rG_cache;
  // folded tcgen05 operator: R_f = LG_f^H (rG x M) and U_f = Phi_f R_f^+ (N x rG), both covariance-independent
  std::vector<std::vector<double>> Rm_cache, U_cache;
  // host copies (float64) kept for covariance_set
  int F = 0, N = 0;
  std::vector<int> M;
  std::vector<std::vector<bogp::cplx>> phi_rec;   // per tonal [N*M]
  bool has_cov = false;
  int max_rows = 0, max_Mp = 0, max_J = 0;
  // scratch for the grid paths (grown on demand)
  void* grid_scratch = nullptr;
  size_t grid_scratch_bytes = 0;
  // device staging for the *_host entry points (grown on demand)
  void* io_scratch = nullptr;
  size_t io_scratch_bytes = 0;
  // small page-locked, device-mapped host block: the arg-ext record (index, value) of the *_host_argmax entry point is
  // written into it by the kernel itself and read after the call's one synchronisation (no device -> host copies)
  unsigned char* pin_host = nullptr;
  unsigned char* pin_dev = nullptr;
  size_t pin_bytes = 0;
  // FP16-split tcgen05 engine (kernels_umma16.cu): operand magnitudes for the per-tonal scales, its own device buffers,
  // and the range vector the cached E table was built for
  struct U16 {
    std::vector<float> wu_max, phi_max;             // per tonal: max |Wu|, max |phi_src table|
    double lever_max = 0.0;                         // max |z_pivot - z_n|
    std::vector<std::vector<bogp::cplx>> k_host;    // per tonal wavenumbers
    void* E = nullptr; size_t E_bytes = 0;
    void* B = nullptr; size_t B_bytes = 0;
    void* misc = nullptr; size_t misc_bytes = 0;
    // streamed Bop tables: per tonal position a device counter of finished table blocks (never reset) and the value it
    // reaches when the tonal's tables of the current call are complete
    unsigned int* tab_cnt = nullptr;
    unsigned int tab_target[BOGP_MAX_TONALS] = {};
    // source mode shapes at the depths of z_cached, [tonal position][depth][Kp] floats (covariance-independent)
    void* phi = nullptr; size_t phi_bytes = 0;
    bool phi_valid = false;
    std::vector<double> r_cached, z_cached;   // host copies of the vectors the device copies / the E table were built from
  } u16;
};

namespace bogp {
// api_mfp.cu
int ensure_pinned(bogp_modeset_s* h, size_t bytes);
// kernels_simt.cu
int launch_points(const bogp_modeset_s* h, const double* cand_dev, long long C, int atype, int mf, bool any_tilt,
                  float* out_dev, cudaStream_t s);
int launch_grid_simt(bogp_modeset_s* h, const double* r_host, int Nr, const double* z_host, int Nz,
                     const double* tilt_host, int Nt, int atype, int mf, float* out_dev, cudaStream_t s);
// kernels_umma.cu
// ext_mode 1 / 2: also write arg-max / arg-min (value, flat index) to ext_val_dev / ext_idx_dev from inside the kernel;
// returns with ext_mode ignored (caller falls back to the arg-ext kernels) only when Nr * Nz >= 2^32.
struct PeerArgs;
// peer != nullptr: the last CTA also publishes (index_offset + index, value) to the peers and collects theirs; *peer_done
// tells the caller whether the engine that ran did so (the 3xTF32 kernel does not)
int launch_grid_umma(bogp_modeset_s* h, const double* r_host, int Nr, const double* z_host, int Nz, int atype,
                     int mf, float* out_dev, cudaStream_t s, int ext_mode = 0, float* ext_val_dev = nullptr,
                     long long* ext_idx_dev = nullptr, const PeerArgs* peer = nullptr, bool* peer_done = nullptr);
bool umma_supported(const bogp_modeset_s* h);
// tilted-array grid (tilt axis) on the tensor cores: N <= 64 receivers, covariance rank <= 4, real source tables
bool umma_tilt_supported(const bogp_modeset_s* h);
int launch_grid_umma_tilt(bogp_modeset_s* h, const double* r_host, int Nr, const double* z_host, int Nz,
                          const double* tilt_host, int Nt, int atype, int mf, float* out_dev, cudaStream_t s,
                          int ext_mode = 0, float* ext_val_dev = nullptr, long long* ext_idx_dev = nullptr);
int ensure_scratch(bogp_modeset_s* h, size_t bytes);
// kernels_umma16.cu: the same surface with FP16-split operands (the default tcgen05 engine for a vertical array)
bool umma16_supported(const bogp_modeset_s* h);
struct PeerArgs;
int launch_grid_umma16(bogp_modeset_s* h, const double* r_host, int Nr, const double* z_host, int Nz, int atype, int mf,
                       float* out_dev, cudaStream_t s, int ext_mode, float* ext_val_dev, long long* ext_idx_dev,
                       const PeerArgs* peer = nullptr);
}  // namespace bogp

// Peer-memory record exchange (kernels_peer.cu); the tcgen05 surface kernel can run the exchange from its last CTA
struct bogp_peer_s {
  int rank = 0, world = 1;
  unsigned char* local = nullptr;               // this rank's buffer (device)
  std::vector<unsigned char*> mapped;           // peers' buffers as mapped here (mapped[rank] == local)
  unsigned char** mapped_dev = nullptr;         // the same array on the device
  unsigned long long epoch = 0;
  size_t bytes = 0;
  int pipelined = 0;                            // bogp_peer_set_pipelined: publish step k, hand back the records of step k - 1
  int last_pipelined = 0;                       // mode of the previous step
};
namespace bogp {
// What a surface kernel needs to run the exchange of the (index, value) record itself: see peer_exchange_kernel.
struct PeerArgs {
  unsigned char* const* bufs = nullptr;      // every rank's buffer as mapped here (device array), null = no exchange
  int rank = 0, world = 1;
  unsigned long long epoch = 0;
  long long index_offset = 0;                // added to the local flat index before the record is published
  long long* recs_out = nullptr;             // [world][2] records of every rank (device or mapped host memory)
  int* err = nullptr;
  int pipelined = 0;                         // 1: collect epoch - 1, then publish epoch (no wait for the peers' current step); 2: collective step after a pipelined one
};
// fills `a` for the next exchange of `p` (advances its epoch)
int peer_next(bogp_peer_s* p, long long index_offset, long long* recs_out, PeerArgs& a);
// the stand-alone exchange kernel on a record already on the device (engines without the fused exchange)
int peer_exchange_launch(const PeerArgs& a, const long long* rec_dev, cudaStream_t s);
}  // namespace bogp

// GP posterior handle (kernels_gp.cu creates it; kernels_ts.cu samples from it)
struct bogp_gp_s {
  int n = 0, d = 0, kind = 0;
  double outputscale = 1.0, mean = 0.0, noise = 0.0, jitter = 0.0;
  double *X = nullptr, *inv_ls = nullptr, *alpha = nullptr, *Linv = nullptr, *LinvT = nullptr;
  // staging of the *_host entry points (grown on demand): device [Xs | out | grad], page-locked host mirror
  double* io_dev = nullptr;
  double* io_pin = nullptr;
  size_t io_doubles = 0;
  // workspace of bogp_gp_thompson (grown on demand)
  void* ts_ws = nullptr;
  size_t ts_bytes = 0;
  // large-batch posterior / acquisition path (FP64 tensor cores): padded L^-1, L^-T, alpha (built once per model) and
  // the per-call [b x n] panels
  void* mm_ws = nullptr;
  size_t mm_bytes = 0;
  int mm_np = 0;               // padding the cached L^-1 / L^-T / alpha were built for (0: not built)
};
namespace bogp {
// kernels_ts.cu: FP64 tensor-core building blocks shared with the large-batch acquisition path
int ts_pad(const bogp_gp_s* g, int np, double* Linv_p, double* alpha_p, cudaStream_t s);
int ts_cross_kernel(const bogp_gp_s* g, const double* Xc_dev, int b, int bp, int np, double* Kt, cudaStream_t s);
int ts_gemm_tri(const double* A, const double* B, double* C, int bp, int np, int tri, cudaStream_t s);
constexpr int kGpMaxDim = 256;      // input dimensions a model may have (the acquisition kernels take <= 16 of them)
}
