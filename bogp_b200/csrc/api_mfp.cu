// C-ABI entry points of Boundary A (matched-field objective): set-up and dispatch.
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstring>

#include "bogp_internal.h"

using bogp::cplx;

namespace {

template <typename T>
int upload(bogp_modeset_s* h, std::vector<void*>& owner, const std::vector<T>& v, const T** out) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  BOGP_CUDA(cudaMalloc(&p, bytes));
  owner.push_back(p);
  if (!v.empty()) BOGP_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = static_cast<const T*>(p);
  (void)h;
  return 0;
}

// Covariance-dependent arrays are re-uploaded at every bogp_covariance_set (once per time step of a sweep): keep the
// device buffer while the new contents fit instead of cudaFree + cudaMalloc (each an implicit device synchronisation).
template <typename T>
int upload_reuse(bogp_modeset_s* h, size_t slot, const std::vector<T>& v, const T** out) {
  if (h->cov_allocs.size() <= slot) { h->cov_allocs.resize(slot + 1, nullptr); h->cov_capacity.resize(slot + 1, 0); }
  const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  if (h->cov_capacity[slot] < bytes) {
    if (h->cov_allocs[slot]) cudaFree(h->cov_allocs[slot]);
    h->cov_allocs[slot] = nullptr; h->cov_capacity[slot] = 0;
    const size_t cap = bytes + bytes / 4 + 256;          // head-room: the operator rows change with the covariance rank
    BOGP_CUDA(cudaMalloc(&h->cov_allocs[slot], cap));
    h->cov_capacity[slot] = cap;
  }
  if (!v.empty()) BOGP_CUDA(cudaMemcpy(h->cov_allocs[slot], v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = static_cast<const T*>(h->cov_allocs[slot]);
  return 0;
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    bogp::set_error("no CUDA device available (%s); libbogp_sm100 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return BOGP_ERR_CUDA;
  }
  return 0;
}

}  // namespace

namespace bogp {
int ensure_scratch(bogp_modeset_s* h, size_t bytes) {
  if (h->grid_scratch_bytes >= bytes) return 0;
  if (h->grid_scratch) cudaFree(h->grid_scratch);
  h->grid_scratch = nullptr;
  h->grid_scratch_bytes = 0;
  BOGP_CUDA(cudaMalloc(&h->grid_scratch, bytes));
  h->grid_scratch_bytes = bytes;
  return 0;
}
}  // namespace bogp

namespace bogp {
int ensure_pinned(bogp_modeset_s* h, size_t bytes) {
  if (h->pin_bytes >= bytes) return 0;
  if (h->pin_host) cudaFreeHost(h->pin_host);
  h->pin_host = h->pin_dev = nullptr; h->pin_bytes = 0;
  const size_t cap = std::max<size_t>(bytes, 4096);
  void* hp = nullptr; void* dp = nullptr;
  BOGP_CUDA(cudaHostAlloc(&hp, cap, cudaHostAllocMapped));
  BOGP_CUDA(cudaHostGetDevicePointer(&dp, hp, 0));
  h->pin_host = static_cast<unsigned char*>(hp); h->pin_dev = static_cast<unsigned char*>(dp); h->pin_bytes = cap;
  return 0;
}

}  // namespace bogp

static int io_scratch(bogp_modeset_s* h, size_t bytes) {
  if (h->io_scratch_bytes >= bytes) return 0;
  if (h->io_scratch) cudaFree(h->io_scratch);
  h->io_scratch = nullptr;
  h->io_scratch_bytes = 0;
  BOGP_CUDA(cudaMalloc(&h->io_scratch, bytes));
  h->io_scratch_bytes = bytes;
  return 0;
}

extern "C" int bogp_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  if (int rc = require_device()) return rc;
  int dev = 0;
  BOGP_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  BOGP_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

extern "C" int bogp_modeset_create(bogp_modeset_t* out, int F, int N, const int* M_host,
                                   const double* k_re_host, const double* k_im_host,
                                   const double* phi_rec_re_host, const double* phi_rec_im_host,
                                   int nz_tab, double z0, double dz, const double* phi_src_re_host,
                                   const double* phi_src_im_host, const double* rec_z_host,
                                   double z_pivot) {
  BOGP_REQUIRE(out != nullptr, "bogp_modeset_create: out is NULL");
  *out = nullptr;
  BOGP_REQUIRE(F >= 1 && F <= BOGP_MAX_TONALS, "bogp_modeset_create: F=%d outside [1,%d]", F, BOGP_MAX_TONALS);
  BOGP_REQUIRE(N >= 1 && N <= 1024, "bogp_modeset_create: N=%d outside [1,1024]", N);
  BOGP_REQUIRE(nz_tab >= 2 && dz > 0.0, "bogp_modeset_create: need nz_tab>=2 and dz>0");
  BOGP_REQUIRE(M_host && k_re_host && k_im_host && phi_rec_re_host && phi_src_re_host && rec_z_host,
               "bogp_modeset_create: NULL input");
  for (int f = 0; f < F; ++f)
    BOGP_REQUIRE(M_host[f] >= 1 && M_host[f] <= 512, "bogp_modeset_create: M[%d]=%d outside [1,512]", f, M_host[f]);
  if (int rc = require_device()) return rc;

  bogp_modeset_s* h = new bogp_modeset_s();
  bogp::ModesetDev& d = h->dev;
  std::memset(&d, 0, sizeof(d));
  d.F = F; d.N = N; d.nz_tab = nz_tab; d.z0 = z0; d.dz = dz; d.z_pivot = z_pivot;
  d.src_complex = phi_src_im_host != nullptr;
  d.rec_complex = phi_rec_im_host != nullptr;
  h->F = F; h->N = N;
  h->M.assign(M_host, M_host + F);
  h->phi_rec.resize(F);
  h->u16.k_host.resize(F);
  h->u16.phi_max.assign(F, 0.f);

  int sumMp = 0, sumTab = 0, sumRec = 0;
  for (int f = 0; f < F; ++f) {
    bogp::TonalMeta& t = d.t[f];
    t.M = M_host[f];
    t.Mp = bogp::pad_to(t.M, 8);
    t.offM = sumMp; t.offTab = sumTab; t.offRec = sumRec;
    sumMp += t.Mp; sumTab += nz_tab * t.Mp; sumRec += N * t.Mp;
    h->max_Mp = std::max(h->max_Mp, t.Mp);
  }
  std::vector<double> kre(sumMp, 1.0), kim(sumMp, 0.0);
  std::vector<float> cre(sumMp, 0.f), cim(sumMp, 0.f);
  std::vector<float> tab_re(sumTab, 0.f), tab_im(d.src_complex ? sumTab : 0, 0.f);
  std::vector<float> rec_re(sumRec, 0.f), rec_im(d.rec_complex ? sumRec : 0, 0.f);
  std::vector<float> recT_re(sumRec, 0.f), recT_im(d.rec_complex ? sumRec : 0, 0.f);
  std::vector<float> lever(N);
  for (int n = 0; n < N; ++n) { lever[n] = (float)(z_pivot - rec_z_host[n]); h->u16.lever_max = std::max(h->u16.lever_max, std::fabs(z_pivot - rec_z_host[n])); }

  size_t ik = 0, irec = 0, isrc = 0;
  for (int f = 0; f < F; ++f) {
    const bogp::TonalMeta& t = d.t[f];
    for (int m = 0; m < t.M; ++m) {
      kre[t.offM + m] = k_re_host[ik + m];
      kim[t.offM + m] = k_im_host[ik + m];
      cplx c = 1.0 / std::sqrt(cplx(k_re_host[ik + m], k_im_host[ik + m]));
      cre[t.offM + m] = (float)c.real();
      cim[t.offM + m] = (float)c.imag();
      h->u16.k_host[f].push_back(cplx(k_re_host[ik + m], k_im_host[ik + m]));
    }
    h->phi_rec[f].resize((size_t)N * t.M);
    for (int n = 0; n < N; ++n)
      for (int m = 0; m < t.M; ++m) {
        double re = phi_rec_re_host[irec + (size_t)n * t.M + m];
        double im = d.rec_complex ? phi_rec_im_host[irec + (size_t)n * t.M + m] : 0.0;
        h->phi_rec[f][(size_t)n * t.M + m] = cplx(re, im);
        rec_re[t.offRec + (size_t)n * t.Mp + m] = (float)re;
        if (d.rec_complex) rec_im[t.offRec + (size_t)n * t.Mp + m] = (float)im;
        recT_re[t.offRec + (size_t)m * N + n] = (float)re;
        if (d.rec_complex) recT_im[t.offRec + (size_t)m * N + n] = (float)im;
      }
    for (int i = 0; i < nz_tab; ++i)
      for (int m = 0; m < t.M; ++m) {
        tab_re[t.offTab + (size_t)i * t.Mp + m] = (float)phi_src_re_host[isrc + (size_t)i * t.M + m];
        h->u16.phi_max[f] = std::max(h->u16.phi_max[f], std::fabs((float)phi_src_re_host[isrc + (size_t)i * t.M + m]));
        if (d.src_complex) tab_im[t.offTab + (size_t)i * t.Mp + m] = (float)phi_src_im_host[isrc + (size_t)i * t.M + m];
      }
    ik += t.M; irec += (size_t)N * t.M; isrc += (size_t)nz_tab * t.M;
  }
  int rc = 0;
  rc |= upload(h, h->allocs, kre, &d.k_re);
  rc |= upload(h, h->allocs, kim, &d.k_im);
  rc |= upload(h, h->allocs, cre, &d.c_re);
  rc |= upload(h, h->allocs, cim, &d.c_im);
  rc |= upload(h, h->allocs, tab_re, &d.tab_re);
  if (d.src_complex) rc |= upload(h, h->allocs, tab_im, &d.tab_im);
  rc |= upload(h, h->allocs, rec_re, &d.rec_re);
  if (d.rec_complex) rc |= upload(h, h->allocs, rec_im, &d.rec_im);
  rc |= upload(h, h->allocs, recT_re, &d.recT_re);
  if (d.rec_complex) rc |= upload(h, h->allocs, recT_im, &d.recT_im);
  rc |= upload(h, h->allocs, lever, &d.rec_lever);
  if (rc) { bogp_modeset_destroy(h); return BOGP_ERR_CUDA; }
  *out = h;
  return 0;
}

extern "C" int bogp_modeset_destroy(bogp_modeset_t h) {
  if (!h) return 0;
  for (void* p : h->allocs) cudaFree(p);
  for (void* p : h->cov_allocs) cudaFree(p);
  if (h->grid_scratch) cudaFree(h->grid_scratch);
  if (h->io_scratch) cudaFree(h->io_scratch);
  if (h->pin_host) cudaFreeHost(h->pin_host);
  if (h->u16.E) cudaFree(h->u16.E);
  if (h->u16.B) cudaFree(h->u16.B);
  if (h->u16.misc) cudaFree(h->u16.misc);
  if (h->u16.tab_cnt) cudaFree(h->u16.tab_cnt);
  if (h->u16.phi) cudaFree(h->u16.phi);
  delete h;
  return 0;
}

extern "C" int bogp_covariance_set(bogp_modeset_t h, const double* K_re_host, const double* K_im_host) {
  BOGP_REQUIRE(h && K_re_host && K_im_host, "bogp_covariance_set: NULL argument");
  const int F = h->F, N = h->N;
  bogp::ModesetDev& d = h->dev;
  const double tol = 1e-13;
  std::vector<float> W, Wt, Wu, Ckre, Ckim;
  h->max_rows = 0; h->max_J = 0;
  for (int f = 0; f < F; ++f) {
    bogp::TonalMeta& t = d.t[f];
    const int M = t.M;
    const std::vector<cplx>& Phi = h->phi_rec[f];   // [N, M]
    // Hermitian part of K_f (the processor takes Re(w^H K w), which only sees (K + K^H)/2).
    std::vector<cplx> K((size_t)N * N);
    double trK = 0.0;
    for (int a = 0; a < N; ++a)
      for (int b = 0; b < N; ++b) {
        size_t ab = (size_t)f * N * N + (size_t)a * N + b, ba = (size_t)f * N * N + (size_t)b * N + a;
        K[(size_t)a * N + b] = 0.5 * (cplx(K_re_host[ab], K_im_host[ab]) + std::conj(cplx(K_re_host[ba], K_im_host[ba])));
      }
    for (int a = 0; a < N; ++a) trK += K[(size_t)a * N + a].real();
    t.trK = (float)trK;
    // G = Phi^H Phi does not depend on the covariance: its pivoted Cholesky factor is computed once per handle.
    if ((int)h->LG_cache.size() != F) { h->LG_cache.assign(F, {}); h->rG_cache.assign(F, -1); }
    if (h->rG_cache[f] < 0) {
      std::vector<cplx> G((size_t)M * M);
      for (int a = 0; a < M; ++a)
        for (int b = a; b < M; ++b) {
          cplx s = 0.0;
          for (int n = 0; n < N; ++n) s += std::conj(Phi[(size_t)n * M + a]) * Phi[(size_t)n * M + b];
          G[(size_t)a * M + b] = s;
          G[(size_t)b * M + a] = std::conj(s);
        }
      for (int a = 0; a < M; ++a) G[(size_t)a * M + a] = G[(size_t)a * M + a].real();
      // Norm rows R (G = R^H R) from the eigen-decomposition of G, truncated at lambda_k > rank_tol * lambda_max: the
      // fewest operator rows for a given accuracy.  Dropping the directions below rank_tol = 1e-10 (1e-5 in amplitude)
      // changes ||Phi a||^2 by at most rank_tol * lambda_max ||a||^2 -- measured on config 2: Bartlett power within
      // 8e-10 relative (un-floored) of the full-rank value, against 4e-7 of fp32 arithmetic -- and removes the
      // near-null directions of a partial-aperture array (13-tonal SWellEx shape: 307 -> 288 padded rows).
      // BOGP_RANK_TOL overrides; BOGP_RANK_TOL=chol keeps the pivoted Cholesky factor at 1e-13.
      const char* rt = std::getenv("BOGP_RANK_TOL");
      if (rt && rt[0] == 'c') bogp::pivoted_cholesky(G, M, tol, h->LG_cache[f], h->rG_cache[f]);
      else {
        h->rank_tol = rt ? std::max(1e-15, std::atof(rt)) : 1e-10;
        bogp::hermitian_eig_factor(G, M, h->rank_tol, h->LG_cache[f], h->rG_cache[f]);
      }
    }
    const std::vector<cplx>& LG = h->LG_cache[f];
    const int rG = h->rG_cache[f];
    std::vector<cplx> LH, LK;
    int rH = 0, rK = 0;
    bogp::pivoted_cholesky(K, N, tol, LK, rK);
    if (rK <= M) {
      // K = LK LK^H  =>  H = Phi^H K Phi = C^H C with C = LK^H Phi (rK x M): the numerator rows directly, no N^2 M
      // product and no second factorisation (rank-1 K = d d^H, the per-snapshot case, costs N M)
      rH = rK;
      LH.assign((size_t)M * M, cplx(0.0, 0.0));
      for (int j = 0; j < rK; ++j)
        for (int m = 0; m < M; ++m) {
          cplx s = 0.0;
          for (int n = 0; n < N; ++n) s += std::conj(LK[(size_t)n * N + j]) * Phi[(size_t)n * M + m];
          LH[(size_t)m * M + j] = std::conj(s);          // LH^H = C
        }
    } else {
      // rank K > M: factor the M x M mode-space matrix instead (at most M numerator rows)
      std::vector<cplx> KP((size_t)N * M), H((size_t)M * M);
      for (int n = 0; n < N; ++n)
        for (int m = 0; m < M; ++m) {
          cplx s = 0.0;
          for (int q = 0; q < N; ++q) s += K[(size_t)n * N + q] * Phi[(size_t)q * M + m];
          KP[(size_t)n * M + m] = s;
        }
      for (int a = 0; a < M; ++a)
        for (int b = a; b < M; ++b) {
          cplx s = 0.0;
          for (int n = 0; n < N; ++n) s += std::conj(Phi[(size_t)n * M + a]) * KP[(size_t)n * M + b];
          H[(size_t)a * M + b] = s;
          H[(size_t)b * M + a] = std::conj(s);
        }
      for (int a = 0; a < M; ++a) H[(size_t)a * M + a] = H[(size_t)a * M + a].real();
      bogp::pivoted_cholesky(H, M, tol, LH, rH);
    }
    // Mode-space operator W: rows = R = LG^H (rG x M) then C = LH^H (rH x M); T = W a.
    t.n_plain = d.rec_complex ? 0 : rG;
    t.n_npair = d.rec_complex ? rG : 0;
    t.n_qpair = rH;
    t.rows = t.n_plain + 2 * t.n_npair + 2 * t.n_qpair;
    t.offW = (int)W.size();
    W.resize(W.size() + (size_t)t.rows * t.Mp, 0.f);
    float* Wf = &W[t.offW];
    int row = 0;
    // R[j, m] = conj(LG[m, j])
    if (!d.rec_complex) {
      for (int j = 0; j < rG; ++j, ++row)
        for (int m = 0; m < M; ++m) Wf[(size_t)row * t.Mp + m] = (float)LG[(size_t)m * M + j].real();
    } else {
      for (int j = 0; j < rG; ++j, ++row)
        for (int m = 0; m < M; ++m) Wf[(size_t)row * t.Mp + m] = (float)LG[(size_t)m * M + j].real();
      for (int j = 0; j < rG; ++j, ++row)
        for (int m = 0; m < M; ++m) Wf[(size_t)row * t.Mp + m] = (float)(-LG[(size_t)m * M + j].imag());
    }
    for (int j = 0; j < rH; ++j, ++row)
      for (int m = 0; m < M; ++m) Wf[(size_t)row * t.Mp + m] = (float)LH[(size_t)m * M + j].real();
    for (int j = 0; j < rH; ++j, ++row)
      for (int m = 0; m < M; ++m) Wf[(size_t)row * t.Mp + m] = (float)(-LH[(size_t)m * M + j].imag());
    t.rowsP = bogp::pad_to(std::max(t.rows, 1), 32);
    t.offWt = (int)Wt.size();
    Wt.resize(Wt.size() + (size_t)t.Mp * t.rowsP, 0.f);
    for (int r2 = 0; r2 < t.rows; ++r2)
      for (int m = 0; m < M; ++m) Wt[t.offWt + (size_t)m * t.rowsP + r2] = Wf[(size_t)r2 * t.Mp + m];
    // Folded UMMA operator (bogp_internal.h): rank-1 numerator carried by the first two rows of the rotated norm rows.
    // Phi = U R with U = Phi R^T (R R^T)^-1 (cached per handle: covariance-independent), so the numerator row
    // C = l^H Phi (l = the rank-1 factor of K) is g^T R with g = U^T conj(l); two Householder reflections Q = H2 H1
    // take Re/Im of g to span(e0, e1), and R' = Q R replaces R.
    t.u_fold = 0; t.u_c[0] = t.u_c[1] = t.u_c[2] = t.u_c[3] = 0.f;
    {
      const char* nf = std::getenv("BOGP_NO_FOLD");
      if (!d.rec_complex && rK == 1 && rH == 1 && rG >= 1 && !(nf && nf[0] == '1')) {
        if ((int)h->Rm_cache.size() != F) { h->Rm_cache.assign(F, {}); h->U_cache.assign(F, {}); }
        std::vector<double>& Rm = h->Rm_cache[f];
        std::vector<double>& U = h->U_cache[f];
        bool ok = true;
        if (Rm.empty()) {
          Rm.assign((size_t)rG * M, 0.0);
          for (int j = 0; j < rG; ++j)
            for (int m = 0; m < M; ++m) Rm[(size_t)j * M + m] = LG[(size_t)m * M + j].real();
          std::vector<double> A((size_t)rG * rG);
          for (int a2 = 0; a2 < rG; ++a2)
            for (int b2 = 0; b2 <= a2; ++b2) {
              double sacc = 0.0;
              for (int m = 0; m < M; ++m) sacc += Rm[(size_t)a2 * M + m] * Rm[(size_t)b2 * M + m];
              A[(size_t)a2 * rG + b2] = A[(size_t)b2 * rG + a2] = sacc;
            }
          U.assign((size_t)N * rG, 0.0);
          if (bogp::cholesky_lower(A, rG)) {
            for (int n = 0; n < N; ++n) {           // row n of U solves (R R^T) u = R phi_n
              double* u = &U[(size_t)n * rG];
              for (int a2 = 0; a2 < rG; ++a2) {
                double sacc = 0.0;
                for (int m = 0; m < M; ++m) sacc += Rm[(size_t)a2 * M + m] * Phi[(size_t)n * M + m].real();
                u[a2] = sacc;
              }
              for (int i = 0; i < rG; ++i) {
                double sacc = u[i];
                for (int k = 0; k < i; ++k) sacc -= A[(size_t)i * rG + k] * u[k];
                u[i] = sacc / A[(size_t)i * rG + i];
              }
              for (int i = rG - 1; i >= 0; --i) {
                double sacc = u[i];
                for (int k = i + 1; k < rG; ++k) sacc -= A[(size_t)k * rG + i] * u[k];
                u[i] = sacc / A[(size_t)i * rG + i];
              }
            }
          } else {
            U.clear();      // marks "no fold" for this tonal
          }
        }
        ok = !U.empty();
        if (ok) {
          std::vector<double> gr(rG, 0.0), gi(rG, 0.0);
          for (int n = 0; n < N; ++n) {
            const double lr = LK[(size_t)n * N].real(), li = -LK[(size_t)n * N].imag();     // conj(l)
            const double* u = &U[(size_t)n * rG];
            for (int j = 0; j < rG; ++j) { gr[j] += lr * u[j]; gi[j] += li * u[j]; }
          }
          // check C = g^T R against the numerator row built above
          double res = 0.0, cmax = 0.0;
          for (int m = 0; m < M; ++m) {
            double er = -LH[(size_t)m * M].real(), ei = LH[(size_t)m * M].imag();
            for (int j = 0; j < rG; ++j) { er += gr[j] * Rm[(size_t)j * M + m]; ei += gi[j] * Rm[(size_t)j * M + m]; }
            res = std::max(res, std::max(std::fabs(er), std::fabs(ei)));
            cmax = std::max(cmax, std::max(std::fabs(LH[(size_t)m * M].real()), std::fabs(LH[(size_t)m * M].imag())));
          }
          // (with truncated norm rows the numerator row is reproduced up to the dropped directions: 4 sqrt(rank_tol))
          ok = res <= std::max(1e-9, 4.0 * std::sqrt(h->rank_tol)) * std::max(cmax, 1e-300);
          if (ok) {
            double nr = 0.0, ni = 0.0;
            for (int j = 0; j < rG; ++j) { nr += gr[j] * gr[j]; ni += gi[j] * gi[j]; }
            const bool r_first = nr >= ni;
            std::vector<double> x = r_first ? gr : gi, y = r_first ? gi : gr;
            std::vector<double> Rp(Rm);
            // reflection across v (I - 2 v v^T / v^T v) applied to a vector / to the rows of Rp, rows [from, rG)
            auto reflect_vec = [&](const std::vector<double>& v, double vv, std::vector<double>& w, int from) {
              double dot = 0.0;
              for (int j = from; j < rG; ++j) dot += v[j] * w[j];
              const double sc = 2.0 * dot / vv;
              for (int j = from; j < rG; ++j) w[j] -= sc * v[j];
            };
            auto reflect_rows = [&](const std::vector<double>& v, double vv, int from) {
              for (int m = 0; m < M; ++m) {
                double dot = 0.0;
                for (int j = from; j < rG; ++j) dot += v[j] * Rp[(size_t)j * M + m];
                const double sc = 2.0 * dot / vv;
                for (int j = from; j < rG; ++j) Rp[(size_t)j * M + m] -= sc * v[j];
              }
            };
            double alpha = 0.0, s0 = 0.0, beta = 0.0;
            {
              double nx = 0.0;
              for (int j = 0; j < rG; ++j) nx += x[j] * x[j];
              nx = std::sqrt(nx);
              if (nx > 0.0) {
                alpha = x[0] > 0.0 ? -nx : nx;
                std::vector<double> v(x);
                v[0] -= alpha;
                double vv = 0.0;
                for (int j = 0; j < rG; ++j) vv += v[j] * v[j];
                if (vv > 0.0) { reflect_vec(v, vv, y, 0); reflect_rows(v, vv, 0); }
                else alpha = x[0];
              }
              s0 = y[0];
              if (rG >= 2) {
                double ny = 0.0;
                for (int j = 1; j < rG; ++j) ny += y[j] * y[j];
                ny = std::sqrt(ny);
                if (ny > 0.0) {
                  beta = y[1] > 0.0 ? -ny : ny;
                  std::vector<double> v(y);
                  v[0] = 0.0;
                  v[1] -= beta;
                  double vv = 0.0;
                  for (int j = 1; j < rG; ++j) vv += v[j] * v[j];
                  if (vv > 0.0) reflect_rows(v, vv, 1);
                  else beta = y[1];
                }
              }
            }
            if (ok) {
              // Q g_first = alpha e0, Q g_second = s0 e0 + beta e1
              if (r_first) { t.u_c[0] = (float)alpha; t.u_c[1] = (float)s0; t.u_c[2] = 0.f; t.u_c[3] = (float)beta; }
              else { t.u_c[0] = (float)s0; t.u_c[1] = (float)alpha; t.u_c[2] = (float)beta; t.u_c[3] = 0.f; }
              t.u_fold = 1;
              t.u_p0 = 0;
              t.u_rows = rG;
              t.u_rowsP = bogp::pad_to(std::max(rG, 1), 16);
              t.offWu = (int)Wu.size();
              Wu.resize(Wu.size() + (size_t)t.u_rowsP * t.Mp, 0.f);
              float* Uo = &Wu[t.offWu];
              for (int k = 0; k < rG; ++k)
                for (int m = 0; m < M; ++m) Uo[(size_t)k * t.Mp + m] = (float)Rp[(size_t)k * M + m];
            }
          }
        }
      }
    }
    // UMMA row order (bogp_internal.h): (re, im) numerator pairs | (re, im) norm pairs | plain rows | zero pad
    if (!t.u_fold) {
    t.u_p0 = 2 * t.n_qpair + 2 * t.n_npair;        // first plain row (even)
    t.u_rowsP = bogp::pad_to(std::max(t.u_p0 + t.n_plain, 1), 16);
    t.offWu = (int)Wu.size();
    Wu.resize(Wu.size() + (size_t)t.u_rowsP * t.Mp, 0.f);
    {
      float* U = &Wu[t.offWu];
      const int src_np = t.n_plain, src_q = t.n_plain + 2 * t.n_npair;
      int dst = 0;
      for (int j = 0; j < t.n_qpair; ++j, dst += 2)
        for (int m = 0; m < M; ++m) {
          U[(size_t)dst * t.Mp + m] = Wf[(size_t)(src_q + j) * t.Mp + m];
          U[(size_t)(dst + 1) * t.Mp + m] = Wf[(size_t)(src_q + t.n_qpair + j) * t.Mp + m];
        }
      for (int j = 0; j < t.n_npair; ++j, dst += 2)
        for (int m = 0; m < M; ++m) {
          U[(size_t)dst * t.Mp + m] = Wf[(size_t)(src_np + j) * t.Mp + m];
          U[(size_t)(dst + 1) * t.Mp + m] = Wf[(size_t)(src_np + t.n_npair + j) * t.Mp + m];
        }
      for (int j = 0; j < t.n_plain; ++j, ++dst)
        for (int m = 0; m < M; ++m) U[(size_t)dst * t.Mp + m] = Wf[(size_t)j * t.Mp + m];
    }
    t.u_rows = t.u_p0 + t.n_plain;
    }
    // Receiver-space factor: num = || LK^H p ||^2 ; Ck[j, n] = conj(LK[n, j]); stored transposed [N][J]
    t.J = rK;
    t.offC = (int)Ckre.size();
    for (int n = 0; n < N; ++n)
      for (int j = 0; j < rK; ++j) {
        Ckre.push_back((float)LK[(size_t)n * N + j].real());
        Ckim.push_back((float)(-LK[(size_t)n * N + j].imag()));
      }
    h->max_rows = std::max(h->max_rows, t.rows);
    h->max_J = std::max(h->max_J, t.J);
  }
  h->u16.wu_max.assign(F, 0.f);
  for (int f = 0; f < F; ++f) {
    const bogp::TonalMeta& t = d.t[f];
    for (size_t i = 0; i < (size_t)t.u_rowsP * t.Mp; ++i) h->u16.wu_max[f] = std::max(h->u16.wu_max[f], std::fabs(Wu[t.offWu + i]));
  }
  // kernels of a previous call may still be reading the arrays that are overwritten in place below
  BOGP_CUDA(cudaDeviceSynchronize());
  int rc = 0;
  rc |= upload_reuse(h, 0, W, &d.W);
  rc |= upload_reuse(h, 1, Wt, &d.Wt);
  rc |= upload_reuse(h, 2, Wu, &d.Wu);
  rc |= upload_reuse(h, 3, Ckre, &d.CkT_re);
  rc |= upload_reuse(h, 4, Ckim, &d.CkT_im);
  if (rc) return BOGP_ERR_CUDA;
  h->has_cov = true;
  return 0;
}

extern "C" int bogp_modeset_info(bogp_modeset_t h, int f, int* M, int* rows, int* J) {
  BOGP_REQUIRE(h && f >= 0 && f < h->F, "bogp_modeset_info: bad argument");
  BOGP_REQUIRE(h->has_cov, "bogp_modeset_info: call bogp_covariance_set first");
  if (M) *M = h->dev.t[f].M;
  if (rows) *rows = h->dev.t[f].rows;
  if (J) *J = h->dev.t[f].J;
  return 0;
}

extern "C" int bogp_modeset_info_umma(bogp_modeset_t h, int f, int* rows, int* rows_padded, int* folded) {
  BOGP_REQUIRE(h && f >= 0 && f < h->F, "bogp_modeset_info_umma: bad argument");
  BOGP_REQUIRE(h->has_cov, "bogp_modeset_info_umma: call bogp_covariance_set first");
  if (rows) *rows = h->dev.t[f].u_rows;
  if (rows_padded) *rows_padded = h->dev.t[f].u_rowsP;
  if (folded) *folded = h->dev.t[f].u_fold;
  return 0;
}

extern "C" int bogp_umma_kind(bogp_modeset_t h) {
  if (!h || !h->has_cov) return 0;
  if (bogp::umma16_supported(h)) return 16;
  return bogp::umma_supported(h) ? 32 : 0;
}

static int check_modes(int atype, int mf) {
  BOGP_REQUIRE(atype == BOGP_ATYPE_CBF || atype == BOGP_ATYPE_CBF_ML, "unknown atype %d", atype);
  BOGP_REQUIRE(mf == BOGP_MF_MEAN || mf == BOGP_MF_PRODUCT, "unknown multifreq method %d", mf);
  return 0;
}

// ext_mode 0: surface only.  1 / 2: arg-max / arg-min of the surface into (ext_val_dev, ext_idx_dev) as well -- inside the
// tcgen05 kernel when that engine runs (no extra launch), by the two arg-ext kernels otherwise.
__global__ void pack_record_kernel(const float* val, const long long* idx, long long offset, long long* rec) {
  const long long i = *idx;
  rec[0] = i < 0 ? i : i + offset;
  rec[1] = (long long)__float_as_uint(*val);
}

static int grid_impl(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host, int Nz,
                     const double* tilt_deg_host, int Nt, int atype, int mf_method, int engine, float* out_dev,
                     void* stream, int ext_mode, float* ext_val_dev, long long* ext_idx_dev, void* ext_scratch,
                     const bogp::PeerArgs* peer = nullptr, bool* peer_done = nullptr) {
  if (peer_done) *peer_done = false;
  BOGP_REQUIRE(h != nullptr, "bogp_bartlett_grid: NULL handle");
  BOGP_REQUIRE(h->has_cov, "bogp_bartlett_grid: call bogp_covariance_set first");
  BOGP_REQUIRE(Nr >= 0 && Nz >= 0 && Nt >= 1, "bogp_bartlett_grid: bad sizes Nr=%d Nz=%d Nt=%d", Nr, Nz, Nt);
  if (int rc = check_modes(atype, mf_method)) return rc;
  if (Nr == 0 || Nz == 0) return 0;   // empty surface: nothing to read or write (pointers may be NULL)
  BOGP_REQUIRE(r_m_host && z_host && out_dev, "bogp_bartlett_grid: NULL argument");
  for (int i = 0; i < Nr; ++i) BOGP_REQUIRE(r_m_host[i] > 0.0, "bogp_bartlett_grid: range[%d]=%g must be > 0", i, r_m_host[i]);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bool tilted = false;
  if (tilt_deg_host)
    for (int i = 0; i < Nt; ++i) tilted |= (tilt_deg_host[i] != 0.0);
  else
    BOGP_REQUIRE(Nt == 1, "bogp_bartlett_grid: Nt must be 1 when tilt is NULL");
  // The receiver-space tensor-core engine also covers complex source mode shapes (KRAKENC) on an untilted array: the
  // depth dependence sits in its A operand, where a complex phi(z) costs nothing extra (tilt = 0, one plane).
  static const double zero_tilt = 0.0;
  const bool tilt_axis = tilted || Nt > 1 || h->dev.src_complex;
  const double* tilt_arg = tilt_deg_host ? tilt_deg_host : &zero_tilt;
  if (engine == BOGP_ENGINE_UMMA) {
    const bool ok = tilt_axis ? bogp::umma_tilt_supported(h) : bogp::umma_supported(h);
    if (!ok) { bogp::set_error("UMMA engine unsupported for this mode set / device"); return BOGP_ERR_UNSUPPORTED; }
  }
  const long long total = (long long)std::max(Nt, 1) * Nz * Nr;
  const bool fuse_ext = total < (1ll << 32);
  int rc;
  if (tilt_axis && (engine == BOGP_ENGINE_UMMA || (engine == BOGP_ENGINE_AUTO && bogp::umma_tilt_supported(h) && total >= 65536))) {
    // tilted array on the tensor cores (receiver-space epilogue)
    rc = bogp::launch_grid_umma_tilt(h, r_m_host, Nr, z_host, Nz, tilt_arg, Nt, atype, mf_method, out_dev, s, ext_mode, ext_val_dev, ext_idx_dev);
    if (rc || !ext_mode || fuse_ext) return rc;
  } else if (!tilt_axis && (engine == BOGP_ENGINE_UMMA || (engine == BOGP_ENGINE_AUTO && bogp::umma_supported(h) && total >= 16384))) {
    rc = bogp::launch_grid_umma(h, r_m_host, Nr, z_host, Nz, atype, mf_method, out_dev, s, ext_mode, ext_val_dev, ext_idx_dev, peer, peer_done);
    if (rc || !ext_mode || fuse_ext) return rc;
  } else {
    // (a tilt axis of zeros still has Nt planes to fill)
    rc = bogp::launch_grid_simt(h, r_m_host, Nr, z_host, Nz, (tilted || Nt > 1) ? tilt_arg : nullptr, Nt, atype, mf_method, out_dev, s);
    if (rc || !ext_mode) return rc;
  }
  return bogp_argext_f32_dev(out_dev, (long long)std::max(Nt, 1) * Nz * Nr, ext_mode == 1, ext_val_dev, ext_idx_dev, ext_scratch, 8192, stream);
}

extern "C" int bogp_bartlett_grid(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host, int Nz,
                                  const double* tilt_deg_host, int Nt, int atype, int mf_method, int engine,
                                  float* out_dev, void* stream) {
  return grid_impl(h, r_m_host, Nr, z_host, Nz, tilt_deg_host, Nt, atype, mf_method, engine, out_dev, stream, 0, nullptr, nullptr, nullptr);
}

extern "C" int bogp_bartlett_grid_argext(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host, int Nz,
                                         const double* tilt_deg_host, int Nt, int atype, int mf_method, int engine,
                                         float* out_dev, int maximize, float* val_dev, long long* idx_dev,
                                         void* scratch_dev, long long scratch_bytes, void* stream) {
  BOGP_REQUIRE(val_dev && idx_dev && scratch_dev && scratch_bytes >= 8192, "bogp_bartlett_grid_argext: val/idx/scratch (>= 8192 bytes) required");
  if ((long long)std::max(Nt, 1) * Nz * Nr == 0) return 0;
  return grid_impl(h, r_m_host, Nr, z_host, Nz, tilt_deg_host, Nt, atype, mf_method, engine, out_dev, stream, maximize ? 1 : 2, val_dev, idx_dev, scratch_dev);
}

// Same choice as grid_impl: does this call run on a tcgen05 engine (where the arg-ext is fused into the surface kernel)?
static bool picks_tensor_engine(bogp_modeset_t h, int engine, const double* tilt_deg_host, int Nt, long long total) {
  if (!h || !h->has_cov || engine == BOGP_ENGINE_SIMT) return false;
  bool tilted = Nt > 1 || h->dev.src_complex;
  if (tilt_deg_host)
    for (int i = 0; i < Nt; ++i) tilted |= (tilt_deg_host[i] != 0.0);
  if (tilted) return bogp::umma_tilt_supported(h) && (engine == BOGP_ENGINE_UMMA || total >= 65536);
  return bogp::umma_supported(h) && (engine == BOGP_ENGINE_UMMA || total >= 16384);
}

// Page-locked, device-mapped host memory (cudaHostAlloc / torch pin_memory): the kernels can write the surface straight
// into it -- the posted PCIe writes of the epilogue overlap the computation and the separate device -> host copy of the
// finished surface disappears from the call.  Returns the device alias of `host`, or nullptr for pageable memory.
static float* mapped_alias(float* host) {
  if (const char* off = std::getenv("BOGP_NO_ZEROCOPY")) { if (off[0] == '1') return nullptr; }
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return nullptr;
  return static_cast<float*>(at.devicePointer);
}

extern "C" int bogp_bartlett_grid_host(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host, int Nz,
                                       const double* tilt_deg_host, int Nt, int atype, int mf_method, int engine,
                                       float* out_host, void* stream) {
  BOGP_REQUIRE(out_host != nullptr, "bogp_bartlett_grid_host: out is NULL");
  BOGP_REQUIRE(h != nullptr, "bogp_bartlett_grid_host: NULL handle");
  size_t n = (size_t)std::max(Nt, 1) * Nz * Nr;
  if (n == 0) return 0;
  if (float* alias = mapped_alias(out_host)) {
    int rc = bogp_bartlett_grid(h, r_m_host, Nr, z_host, Nz, tilt_deg_host, Nt, atype, mf_method, engine, alias, stream);
    if (rc == 0 && cudaStreamSynchronize(static_cast<cudaStream_t>(stream)) != cudaSuccess) {
      bogp::set_error("bogp_bartlett_grid_host: %s", cudaGetErrorString(cudaGetLastError()));
      rc = BOGP_ERR_CUDA;
    }
    return rc;
  }
  if (int rc0 = io_scratch(h, n * sizeof(float))) return rc0;
  float* tmp = static_cast<float*>(h->io_scratch);
  int rc = bogp_bartlett_grid(h, r_m_host, Nr, z_host, Nz, tilt_deg_host, Nt, atype, mf_method, engine, tmp, stream);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out_host, tmp, n * sizeof(float), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream));
    if (e == cudaSuccess) e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) { bogp::set_error("bogp_bartlett_grid_host: %s", cudaGetErrorString(e)); rc = BOGP_ERR_CUDA; }
  }
  return rc;
}

extern "C" int bogp_bartlett_grid_host_argmax(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host,
                                              int Nz, const double* tilt_deg_host, int Nt, int atype, int mf_method,
                                              int engine, float* out_host, int maximize, float* val_host,
                                              long long* idx_host, void* stream) {
  BOGP_REQUIRE(out_host && val_host && idx_host, "bogp_bartlett_grid_host_argmax: NULL output");
  BOGP_REQUIRE(h != nullptr, "bogp_bartlett_grid_host_argmax: NULL handle");
  const size_t n = (size_t)std::max(Nt, 1) * Nz * Nr;
  *val_host = 0.f;
  *idx_host = -1;
  if (n == 0) return 0;
  // staging: [surface n floats][pad to 16][idx int64, val float][argext scratch 8192]
  const size_t off_pair = (n * sizeof(float) + 15) & ~(size_t)15;
  if (int rc0 = io_scratch(h, off_pair + 16 + 8192)) return rc0;
  char* base = static_cast<char*>(h->io_scratch);
  // mapped host memory: the kernel writes it directly (only where the reduction is fused: the separate arg-ext
  // kernels would re-read the surface over PCIe)
  float* alias = picks_tensor_engine(h, engine, tilt_deg_host, Nt, (long long)n) ? mapped_alias(out_host) : nullptr;
  float* tmp = alias ? alias : reinterpret_cast<float*>(base);
  // the (index, value) record lands in mapped page-locked memory: written by the kernel, read after the one
  // synchronisation below -- no device -> host copies for it
  if (int rc1 = bogp::ensure_pinned(h, 64)) return rc1;
  long long* idx_dev = reinterpret_cast<long long*>(h->pin_dev);
  float* val_dev = reinterpret_cast<float*>(h->pin_dev + 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = bogp_bartlett_grid_argext(h, r_m_host, Nr, z_host, Nz, tilt_deg_host, Nt, atype, mf_method, engine, tmp,
                                     maximize, val_dev, idx_dev, base + off_pair + 16, 8192, stream);
  if (rc == 0) {
    cudaError_t e = alias ? cudaSuccess : cudaMemcpyAsync(out_host, tmp, n * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { bogp::set_error("bogp_bartlett_grid_host_argmax: %s", cudaGetErrorString(e)); rc = BOGP_ERR_CUDA; }
    else {
      *idx_host = *reinterpret_cast<volatile long long*>(h->pin_host);
      *val_host = *reinterpret_cast<volatile float*>(h->pin_host + 8);
    }
  }
  return rc;
}

// Surface + arg-ext + exchange of the (index_offset + index, value) records with the other GPUs of the box.  On the
// FP16-split tcgen05 engine the exchange runs in the surface kernel's last CTA (no launch of its own); otherwise the
// record is packed by a one-thread kernel and exchanged by peer_exchange_kernel.
static int grid_peer_impl(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host, int Nz,
                          const double* tilt_deg_host, int Nt, int atype, int mf_method, int engine, float* out_dev,
                          int maximize, bogp_peer_t peer, long long index_offset, long long* recs_out_dev, float* val_dev,
                          long long* idx_dev, void* scratch_dev, void* stream) {
  BOGP_REQUIRE(peer && recs_out_dev && val_dev && idx_dev && scratch_dev, "bogp_bartlett_grid_argext_peer: NULL argument");
  BOGP_REQUIRE((long long)std::max(Nt, 1) * Nz * Nr > 0, "bogp_bartlett_grid_argext_peer: empty surface");
  bogp::PeerArgs a;
  if (int rc = bogp::peer_next(peer, index_offset, recs_out_dev, a)) return rc;
  bool done = false;
  int rc = grid_impl(h, r_m_host, Nr, z_host, Nz, tilt_deg_host, Nt, atype, mf_method, engine, out_dev, stream,
                     maximize ? 1 : 2, val_dev, idx_dev, scratch_dev, &a, &done);
  if (rc || done) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long* rec = reinterpret_cast<long long*>(static_cast<char*>(scratch_dev) + 8192 - 16);      // the arg-ext kernels use the first 7104 bytes
  pack_record_kernel<<<1, 1, 0, s>>>(val_dev, idx_dev, index_offset, rec);
  bogp::count_launch();
  return bogp::peer_exchange_launch(a, rec, s);
}

extern "C" int bogp_bartlett_grid_argext_peer(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host, int Nz,
                                              const double* tilt_deg_host, int Nt, int atype, int mf_method, int engine,
                                              float* out_dev, int maximize, bogp_peer_t peer, long long index_offset,
                                              long long* recs_out_dev, float* val_dev, long long* idx_dev,
                                              void* scratch_dev, long long scratch_bytes, void* stream) {
  BOGP_REQUIRE(scratch_bytes >= 8192, "bogp_bartlett_grid_argext_peer: scratch must be >= 8192 bytes");
  return grid_peer_impl(h, r_m_host, Nr, z_host, Nz, tilt_deg_host, Nt, atype, mf_method, engine, out_dev, maximize, peer,
                        index_offset, recs_out_dev, val_dev, idx_dev, scratch_dev, stream);
}

extern "C" int bogp_bartlett_grid_host_argmax_peer(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host,
                                                   int Nz, const double* tilt_deg_host, int Nt, int atype, int mf_method,
                                                   int engine, float* out_host, int maximize, bogp_peer_t peer,
                                                   long long index_offset, long long* recs_host, void* stream) {
  BOGP_REQUIRE(out_host && recs_host && peer, "bogp_bartlett_grid_host_argmax_peer: NULL argument");
  BOGP_REQUIRE(h != nullptr, "bogp_bartlett_grid_host_argmax_peer: NULL handle");
  const size_t n = (size_t)std::max(Nt, 1) * Nz * Nr;
  BOGP_REQUIRE(n > 0, "bogp_bartlett_grid_host_argmax_peer: empty surface");
  const int world = peer->world;
  const size_t off_pair = (n * sizeof(float) + 15) & ~(size_t)15;
  if (int rc0 = io_scratch(h, off_pair + 16 + 8192)) return rc0;
  char* base = static_cast<char*>(h->io_scratch);
  float* alias = picks_tensor_engine(h, engine, tilt_deg_host, Nt, (long long)n) ? mapped_alias(out_host) : nullptr;
  float* tmp = alias ? alias : reinterpret_cast<float*>(base);
  // mapped page-locked block: [idx, val | records of every rank], written by the kernels, read after the one synchronisation
  if (int rc1 = bogp::ensure_pinned(h, 64 + (size_t)world * 16)) return rc1;
  long long* idx_dev = reinterpret_cast<long long*>(h->pin_dev);
  float* val_dev = reinterpret_cast<float*>(h->pin_dev + 8);
  long long* recs_dev = reinterpret_cast<long long*>(h->pin_dev + 64);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = grid_peer_impl(h, r_m_host, Nr, z_host, Nz, tilt_deg_host, Nt, atype, mf_method, engine, tmp, maximize, peer,
                          index_offset, recs_dev, val_dev, idx_dev, base + off_pair + 16, stream);
  if (rc == 0) {
    cudaError_t e = alias ? cudaSuccess : cudaMemcpyAsync(out_host, tmp, n * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { bogp::set_error("bogp_bartlett_grid_host_argmax_peer: %s", cudaGetErrorString(e)); rc = BOGP_ERR_CUDA; }
    else std::memcpy(recs_host, h->pin_host + 64, (size_t)world * 16);
  }
  return rc;
}

// any(cand[:,2] != 0) on the device, so the tilt-free fast path is chosen without a host copy of cand.
__global__ void any_tilt_kernel(const double* __restrict__ cand, long long C, int* flag) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  int any = 0;
  for (; i < C; i += stride) any |= (cand[3 * i + 2] != 0.0);
  if (__syncthreads_or(any) && threadIdx.x == 0) atomicOr(flag, 1);
}

extern "C" int bogp_bartlett_points(bogp_modeset_t h, const double* cand_dev, long long C, int atype, int mf_method,
                                    float* out_dev, void* stream) {
  BOGP_REQUIRE(h && (C == 0 || (cand_dev && out_dev)), "bogp_bartlett_points: NULL argument");
  BOGP_REQUIRE(h->has_cov, "bogp_bartlett_points: call bogp_covariance_set first");
  BOGP_REQUIRE(C >= 0, "bogp_bartlett_points: C < 0");
  if (int rc = check_modes(atype, mf_method)) return rc;
  if (C == 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = bogp::ensure_scratch(h, 256)) return rc;
  int* flag = static_cast<int*>(h->grid_scratch);
  BOGP_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
  int blocks = (int)std::min<long long>((C + 255) / 256, 1184);
  any_tilt_kernel<<<blocks, 256, 0, s>>>(cand_dev, C, flag);
  bogp::count_launch();
  int any = 0;
  BOGP_CUDA(cudaMemcpyAsync(&any, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  BOGP_CUDA(cudaStreamSynchronize(s));
  return bogp::launch_points(h, cand_dev, C, atype, mf_method, any != 0, out_dev, s);
}

extern "C" int bogp_bartlett_points_host(bogp_modeset_t h, const double* cand_host, long long C, int atype,
                                         int mf_method, float* out_host, void* stream) {
  BOGP_REQUIRE(C >= 0 && (C == 0 || (cand_host && out_host)), "bogp_bartlett_points_host: bad argument");
  if (C == 0) return 0;
  BOGP_REQUIRE(h != nullptr, "bogp_bartlett_points_host: NULL handle");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc0 = io_scratch(h, (size_t)C * (3 * sizeof(double) + sizeof(float)))) return rc0;
  double* cd = static_cast<double*>(h->io_scratch);
  float* od = reinterpret_cast<float*>(cd + (size_t)C * 3);
  int rc = 0;
  cudaError_t e = bogp::h2d_async(cd, cand_host, (size_t)C * 3 * sizeof(double), s);
  if (e != cudaSuccess) { bogp::set_error("H2D: %s", cudaGetErrorString(e)); rc = BOGP_ERR_CUDA; }
  if (!rc) rc = bogp_bartlett_points(h, cd, C, atype, mf_method, od, stream);
  if (!rc) {
    e = cudaMemcpyAsync(out_host, od, (size_t)C * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { bogp::set_error("D2H: %s", cudaGetErrorString(e)); rc = BOGP_ERR_CUDA; }
  }
  return rc;
}
