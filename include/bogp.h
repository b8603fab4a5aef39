/* bogp.h -- C ABI of libbogp_sm100.so: the B200-native hot path of NeptuneProjects/BOGP.
 *
 * The reference has no FFI of its own: its hot path is reached through two Python
 * protocols.  Each entry point below names the reference interface it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the
 * reference side.
 *
 *   Boundary A  objective callable  tritonoa.sp.mfp.MatchedFieldProcessor(...).__call__(dict)
 *               built at ref/projects/swellex96_localization/data/mfp.py:198-204,
 *               ref/projects/swellex96_inv/data/bo/obj.py:62-70,
 *               ref/projects/swellex96_localization/conf/optimization/run.py:89-96
 *   Boundary B  botorch Model.posterior / AcquisitionFunction.forward / optimize_acqf
 *               used at ref/projects/swellex96_inv/data/bo/ei.py:65-92, ucb.py:82,
 *               ref/projects/swellex96_localization/data/demo_1d.py:92-112,
 *               ref/optimization/gpmodels.py:16-30
 *
 * Conventions: every function returns 0 on success, <0 on error (bogp_last_error() gives
 * the text, thread-local).  Pointers named *_host are host memory, *_dev are device memory
 * on the current CUDA device, caller-owned.  `stream` is a cudaStream_t passed as void*
 * (NULL = legacy default stream).  Handles are not thread-safe; distinct handles are.
 * Ranges are in METRES at this boundary (the Python objective takes km, mfp.py:214).
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef BOGP_H
#define BOGP_H

#ifdef __cplusplus
extern "C" {
#endif

#define BOGP_OK 0
#define BOGP_ERR_INVALID (-1)
#define BOGP_ERR_CUDA (-2)
#define BOGP_ERR_NUMERIC (-3)
#define BOGP_ERR_UNSUPPORTED (-4)

/* beamformer `atype` (tritonoa.sp.beamforming.beamformer; obj.py:68, loc_tilt/conf/run.py:146) */
#define BOGP_ATYPE_CBF 0
#define BOGP_ATYPE_CBF_ML 1
/* MatchedFieldProcessor `multifreq_method` (loc_tilt/conf/run.py:263-264) */
#define BOGP_MF_MEAN 0
#define BOGP_MF_PRODUCT 1
/* covariance kernels of the surrogate (baxus.py:310-319 Matern-5/2 ARD; gpmodels.py:22-24 RBF ARD) */
#define BOGP_KERNEL_MATERN52 0
#define BOGP_KERNEL_RBF 1
/* analytic acquisitions (ei.py:80, logei.py:80, pi.py:81, ucb.py:82) */
#define BOGP_ACQ_EI 0
#define BOGP_ACQ_LOGEI 1
#define BOGP_ACQ_PI 2
#define BOGP_ACQ_UCB 3
/* engine selection for the grid path */
#define BOGP_ENGINE_AUTO 0
#define BOGP_ENGINE_SIMT 1   /* CUDA-core kernel */
#define BOGP_ENGINE_UMMA 2   /* tcgen05 3xTF32 kernel */

typedef struct bogp_modeset_s* bogp_modeset_t;
typedef struct bogp_envbatch_s* bogp_envbatch_t;
typedef struct bogp_gp_s* bogp_gp_t;

const char* bogp_last_error(void);
int bogp_version(void);
/* Kernels launched by this library since load (monotonic; bench.py reports the per-step difference). */
long long bogp_kernel_launches(void);
/* Bytes of call inputs (range / depth / tilt vectors, candidate lists) copied host -> device by the Bartlett entry points
 * since load (monotonic): vectors that have not changed since the previous call on the same handle are not copied
 * again.  bench.py reports the per-step difference as e2e.h2d_bytes_per_step. */
long long bogp_h2d_bytes(void);
/* Roofline instrumentation: while enabled, each launch of a dominant kernel (the Bartlett grid/points kernels, the
 * GP evaluation kernel) is bracketed by CUDA events on its own stream; bogp_timing_read() waits for them and returns
 * the summed kernel time and the number of launches since the last read. */
int bogp_timing_enable(int on);
int bogp_timing_read(double* total_ms, long long* count);
/* sm_count / compute capability of the current device; fails without a GPU. */
int bogp_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------ Boundary A: matched-field objective
 * Replaces the modal-sum half of tritonoa...run_kraken (the `runner` of MatchedFieldProcessor) for one
 * environment: modes precomputed on the host (KRAKEN stays external), uploaded once.
 *   M_host[F]                     modes per tonal
 *   k_re/k_im_host[sum M]         horizontal wavenumbers
 *   phi_rec_re/_im_host           per tonal row-major [N, M_f], concatenated; _im may be NULL (KRAKEN real modes)
 *   phi_src_re/_im_host           per tonal row-major [nz_tab, M_f] on the mesh z0 + i*dz; _im may be NULL
 *   rec_z_host[N], z_pivot        array geometry for the tilted-array path
 */
int bogp_modeset_create(bogp_modeset_t* out, int F, int N, const int* M_host,
                        const double* k_re_host, const double* k_im_host,
                        const double* phi_rec_re_host, const double* phi_rec_im_host,
                        int nz_tab, double z0, double dz,
                        const double* phi_src_re_host, const double* phi_src_im_host,
                        const double* rec_z_host, double z_pivot);
int bogp_modeset_destroy(bogp_modeset_t h);

/* `covariance_matrix` argument of MatchedFieldProcessor: K[F, N, N] complex, row-major, split planes
 * (layout [F,N,N] per ref/optimization/utils.py:23-32, 42-56).  Factorises K on the host (pivoted
 * Cholesky, float64) into the projection rows the kernels consume. */
int bogp_covariance_set(bogp_modeset_t h, const double* K_re_host, const double* K_im_host);

/* Shape of tonal f after bogp_covariance_set: modes M_f, rows of the mode-space operator W_f = [R_f ; C_f]
 * (G = Phi^H Phi = R^H R, Phi^H K Phi = C^H C; complex rows count twice) and rank J of K_f.  The algorithmic work of
 * one candidate x tonal on the vertical-array path is 4 * rows * M flop. */
int bogp_modeset_info(bogp_modeset_t h, int f, int* M, int* rows, int* J);
/* Operator of tonal f as the tcgen05 engine multiplies by it: rows before / after padding to the MMA granularity, and
 * whether the rank-1 numerator is folded into the norm rows (real receiver modes, rank-1 covariance: an orthogonal
 * rotation of R puts the numerator row into the span of its first two rows, so the operator has rank(G) rows instead
 * of rank(G) + 2).  Valid after bogp_covariance_set. */
int bogp_modeset_info_umma(bogp_modeset_t h, int f, int* rows, int* rows_padded, int* folded);
/* Arithmetic of the tcgen05 engine a vertical-array grid call on this handle runs: 16 = three FP16-split products
 * (kind::f16, csrc/kernels_umma16.cu), 32 = three TF32-split products (kind::tf32, csrc/kernels_umma.cu; also
 * BOGP_UMMA_KIND=tf32), 0 = the tensor-core engines do not take this mode set / device.  Same 22-bit operand precision
 * either way (bench.py picks the roofline denominator from it).  Valid after bogp_covariance_set. */
int bogp_umma_kind(bogp_modeset_t h);

/* Ambiguity surface: replaces the loop `for z in src_z: amb[zz,:] = MFP({"src_z": z, "rec_r": rvec})`
 * (mfp.py:213-214) including runner + beamformer + multifreq combine.
 * out_dev[Nt, Nz, Nr] float32 (Nt = 1 and tilt_deg_host = NULL for a vertical array). */
int bogp_bartlett_grid(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host, int Nz,
                       const double* tilt_deg_host, int Nt, int atype, int mf_method, int engine,
                       float* out_dev, void* stream);
/* Surface + the final reduction np.unravel_index(np.argmax(surface)) (collate.py:50) in one call, all on the device:
 * val_dev / idx_dev receive the extremum and its flat index (-1 if every value is NaN).  With the tcgen05 engine the
 * reduction happens inside the surface kernel's epilogue (packed atomicMax, last CTA decodes): no extra launch.
 * scratch_dev: >= 8192 bytes (used only when the reduction needs the separate arg-ext kernels). */
int bogp_bartlett_grid_argext(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host, int Nz,
                              const double* tilt_deg_host, int Nt, int atype, int mf_method, int engine,
                              float* out_dev, int maximize, float* val_dev, long long* idx_dev,
                              void* scratch_dev, long long scratch_bytes, void* stream);

/* Same, host output buffer (device->host copy inside; synchronises the stream). */
int bogp_bartlett_grid_host(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host, int Nz,
                            const double* tilt_deg_host, int Nt, int atype, int mf_method, int engine,
                            float* out_host, void* stream);

/* Surface + its arg-max (arg-min for the minimised cbf_ml objective) in one call: what mfp.py:213-214 followed by
 * np.unravel_index(np.argmax(surface)) (collate.py:50) produce.  One stream synchronisation for both results. */
int bogp_bartlett_grid_host_argmax(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host, int Nz,
                                   const double* tilt_deg_host, int Nt, int atype, int mf_method, int engine,
                                   float* out_host, int maximize, float* val_host, long long* idx_host, void* stream);

/* Scattered candidates (BO raw samples / trials): cand_dev[C, 3] float64 = (range m, depth m, tilt deg).
 * Replaces one MatchedFieldProcessor.__call__ per candidate (obj.py:28-30). */
int bogp_bartlett_points(bogp_modeset_t h, const double* cand_dev, long long C, int atype, int mf_method,
                         float* out_dev, void* stream);
int bogp_bartlett_points_host(bogp_modeset_t h, const double* cand_host, long long C, int atype,
                              int mf_method, float* out_host, void* stream);


/* Geoacoustic inversion (obj.py:73-83): candidate c has its OWN mode set (host KRAKEN run per candidate).
 * Mode-major packing, padded to Mmax:  k_re/k_im_host[C, F, Mmax], amp_re/amp_im_host[C, F, Mmax] =
 * phi_m(z_s) already evaluated at the candidate's source depth (0 for padded modes),
 * phi_rec_T_host[C, F, Mmax, N] float32 (real modes).  range_m_host[C].  K is shared (bogp_envbatch_set_covariance). */
int bogp_envbatch_create(bogp_envbatch_t* out, long long C, int F, int N, int Mmax,
                         const double* k_re_host, const double* k_im_host,
                         const double* amp_re_host, const double* amp_im_host,
                         const float* phi_rec_T_host, const double* range_m_host);
int bogp_envbatch_set_covariance(bogp_envbatch_t h, const double* K_re_host, const double* K_im_host);
/* Tilted array for the environment batch: tilt_deg_host[C] and lever_host[C][N] = z_pivot - z_n of each candidate's
 * array (receiver n then sits at range r - lever_n sin(tilt): the `tilt` key of the inversion's search space,
 * projects/swellex96_inv/data/bo/common.py:63, evaluated through obj.py:62-70).  tilt_deg_host = NULL or all zeros
 * returns the batch to the vertical-array kernels. */
int bogp_envbatch_set_tilt(bogp_envbatch_t h, const double* tilt_deg_host, const double* lever_host);
int bogp_envbatch_bartlett(bogp_envbatch_t h, int atype, int mf_method, float* out_dev, void* stream);
int bogp_envbatch_destroy(bogp_envbatch_t h);

/* np.unravel_index(np.argmax(surface)) (ref/projects/swellex96_localization/data/collate.py:50):
 * first index of the maximum of x_dev[n]; NaNs are ignored. */
int bogp_argmax_f32(const float* x_dev, long long n, float* val_host, long long* idx_host, void* stream);
int bogp_argmin_f32(const float* x_dev, long long n, float* val_host, long long* idx_host, void* stream);
/* Same without the host round trip (multi-GPU final reduction: the (value, index) pair stays on the device for the
 * NCCL all-gather).  scratch_dev: >= 8192 bytes of device memory.  idx_dev gets -1 if every element is NaN. */
int bogp_argext_f32_dev(const float* x_dev, long long n, int maximize, float* val_dev, long long* idx_dev,
                        void* scratch_dev, long long scratch_bytes, void* stream);

/* ------------------------------------------------------------------ Boundary B: GP posterior + acquisition
 * Replaces SingleTaskGP(X, Y, ...) in eval mode (ei.py:67; gpmodels.py:16-30): exact GP with constant mean,
 * outputscale * {Matern-5/2 | RBF} ARD kernel, homoskedastic noise.  Factorises K_XX + noise*I = L L^T
 * (float64, jitter retries 1e-8,1e-7,1e-6) and uploads L^-1 and alpha.  All GP arithmetic is float64. */
int bogp_gp_create(bogp_gp_t* out, int n, int d, const double* X_host, const double* y_host, int kernel_kind,
                   const double* lengthscale_host, double outputscale, double mean, double noise);
int bogp_gp_destroy(bogp_gp_t g);
/* jitter that had to be added to the diagonal (0 if none). */
int bogp_gp_jitter(bogp_gp_t g, double* jitter_out);

/* Model.posterior(X[b,1,d]).mean / .variance  (variance clamped at 1e-10 as gpytorch does for float64). */
int bogp_gp_posterior(bogp_gp_t g, const double* Xs_dev, long long b, int observation_noise,
                      double* mean_dev, double* var_dev, void* stream);
/* AcquisitionFunction.forward(X[b,1,d]) -> out_dev[b]; grad_dev[b,d] = d out / d X (NULL to skip):
 * fused posterior + EI/LogEI/PI/UCB epilogue (+ analytic backward for optimize_acqf's L-BFGS-B). */
int bogp_gp_acq(bogp_gp_t g, int acq_kind, double best_f, double beta, const double* Xs_dev, long long b,
                double* out_dev, double* grad_dev, void* stream);
/* Joint posterior for q > 1: mean_dev[b,q], cov_dev[b,q,q]. */
int bogp_gp_posterior_joint(bogp_gp_t g, const double* Xs_dev, long long b, int q,
                            double* mean_dev, double* cov_dev, void* stream);
/* Backward of the joint posterior: given d F/d mean [b,q] and d F/d cov [b,q,q] (entries of the full matrix; the
 * kernel symmetrises), grad_dev[b,q,d] = d F/d X.  This is the autograd hook botorch's MC acquisitions use through
 * Model.posterior (optimize_acqf's L-BFGS-B needs it). */
int bogp_gp_posterior_joint_backward(bogp_gp_t g, const double* Xs_dev, long long b, int q,
                                     const double* gmean_dev, const double* gcov_dev, double* grad_dev, void* stream);
/* MC qEI (Ax Models.GPEI path, ref/optimization/strategy.py:61): base_dev[S,q] standard-normal QMC samples.
 * out_dev[b]; grad_dev[b,q,d] or NULL. */
int bogp_gp_qei(bogp_gp_t g, const double* Xs_dev, long long b, int q, const double* base_dev, int S,
                double best_f, double* out_dev, double* grad_dev, void* stream);

/* Thompson sampling over a candidate set (BAxUS create_candidate, projects/swellex96_inv/data/bo/baxus.py:121-141:
 * botorch MaxPosteriorSampling(model, replacement=False)(X_cand, num_samples)).  One joint posterior draw per sample
 * at the b candidates, f = mu + chol(Sigma) z with the noise-free posterior (mu, Sigma) in float64 and the standard
 * normal base samples z_dev[S][b] supplied by the caller; sample s picks the arg-max among the candidates no earlier
 * sample picked.  idx_dev[S] / val_dev[S]: index into Xc and sampled value.  samples_dev[S][b] and mean_dev[b] are
 * optional outputs (NULL to skip).  jitter_out (host, optional): diagonal jitter the factorisation needed (0, or 1e-8 /
 * 1e-7 / 1e-6 after a failed attempt: gpytorch psd_safe_cholesky); BOGP_ERR_NUMERIC if 1e-6 is not enough.
 * Unlike the acquisition entry points this one accepts models with up to 256 input dimensions (BAxUS grows its
 * target space).  b <= 16384, S <= 64.  Synchronises the stream once per factorisation attempt. */
int bogp_gp_thompson(bogp_gp_t g, const double* Xc_dev, int b, const double* z_dev, int S, long long* idx_dev,
                     double* val_dev, double* samples_dev, double* mean_dev, double* jitter_out, void* stream);

/* Host-buffer variants of bogp_gp_acq / bogp_gp_qei for the optimiser loop of optimize_acqf (ei.py:81-92): L-BFGS-B calls
 * value + gradient hundreds of times per BO iteration on a handful of restarts; one C call with page-locked staging
 * replaces the framework round trip.  Xs_host[b,(q,)d], out_host[b], grad_host[b,(q,)d] or NULL; base_dev stays on the
 * device.  Synchronises the stream. */
int bogp_gp_acq_host(bogp_gp_t g, int acq_kind, double best_f, double beta, const double* Xs_host, long long b,
                     double* out_host, double* grad_host, void* stream);
int bogp_gp_qei_host(bogp_gp_t g, const double* Xs_host, long long b, int q, const double* base_dev, int S,
                     double best_f, double* out_host, double* grad_host, void* stream);

/* GP hyper-parameter fit: the loop of ref/projects/swellex96_inv/data/bo/ei.py:69-78 --
 * `steps` x AdamW(lr, weight_decay; betas 0.9/0.999, eps 1e-8 as torch.optim.AdamW) on -mll/n over the raw GPyTorch
 * parameters -- as ONE persistent kernel launch per call (kernel matrix, blocked Cholesky, K^-1, analytic gradient and
 * the AdamW update all on the device, float64).  raw_inout_host[B][d+3] = raw {lengthscale[d], outputscale, noise,
 * mean}: lengthscale = softplus(raw), outputscale = softplus(raw), noise = lo + (hi-lo) sigmoid(raw) (the
 * Interval(lo, hi) constraint of ei.py:66), mean = raw.  B independent starting points run as B concurrent CTAs.
 * lengthscale_prior / outputscale_prior: Gamma {concentration, rate} or NULL.  loss_host[B][steps] (optional):
 * -mll/n before each update.  Returns BOGP_ERR_NUMERIC if K + noise I stays indefinite after jitter 1e-6. */
int bogp_gp_fit_adamw(int n, int d, const double* X_host, const double* y_host, int kernel_kind, int steps,
                      double lr, double weight_decay, double noise_lo, double noise_hi,
                      const double* lengthscale_prior, const double* outputscale_prior, int B,
                      double* raw_inout_host, double* loss_host, void* stream);
/* The same with Interval constraints on the length scales / output scale instead of softplus (the explicit
 * ScaleKernel(MaternKernel(lengthscale_constraint=Interval(0.005, 10)), outputscale_constraint=Interval(0.05, 10)) of
 * baxus.py:310-319): value = lo + (hi - lo) sigmoid(raw).  lengthscale_bounds / outputscale_bounds: {lo, hi} or NULL. */
int bogp_gp_fit_adamw_bounded(int n, int d, const double* X_host, const double* y_host, int kernel_kind, int steps,
                              double lr, double weight_decay, double noise_lo, double noise_hi,
                              const double* lengthscale_prior, const double* outputscale_prior,
                              const double* lengthscale_bounds, const double* outputscale_bounds, int B,
                              double* raw_inout_host, double* loss_host, void* stream);
/* One evaluation of the same objective: loss_host[B] = -(log marginal likelihood + log priors) / n and
 * grad_host[B][d + 3] = d(loss)/d(raw) at raw_host[B][d + 3] (raw lengthscale[d], outputscale, noise, mean), on the
 * same kernels (one step, no update).  What a hand-rolled loop `loss = -mll(model(X), y); loss.backward()`
 * (projects/swellex96_inv/data/bo/ei.py:72-77) and botorch.fit.fit_gpytorch_mll's scipy L-BFGS-B closure
 * (projects/swellex96_localization/data/demo_1d.py:137, projects/swellex96_inv/data/bo/baxus.py:333) evaluate. */
int bogp_gp_mll_grad(int n, int d, const double* X_host, const double* y_host, int kernel_kind, double noise_lo,
                     double noise_hi, const double* lengthscale_prior, const double* outputscale_prior,
                     const double* lengthscale_bounds, const double* outputscale_bounds, int B,
                     const double* raw_host, double* loss_host, double* grad_host, void* stream);

/* ---- Final reduction across the GPUs of one box over peer memory (NVLink P2P stores), no collective library.
 * Replaces the host-side np.argmax over collated surfaces (projects/swellex96_localization/data/collate.py:50) and the
 * NCCL all-gather of one 16-byte (flat index, value bits) record per rank.  Set-up: every rank calls bogp_peer_create
 * (allocates its buffer, returns a BOGP_IPC_HANDLE_BYTES CUDA IPC handle), the ranks exchange the handles by any means
 * (torch.distributed all_gather), every rank calls bogp_peer_open with all handles in rank order.  Per step:
 * bogp_peer_exchange stores rec_dev[2] into every peer and fills recs_out_dev[world][2] once every peer's record of
 * the same step has arrived (one single-block kernel on `stream`; all ranks must call it the same number of times). */
#define BOGP_IPC_HANDLE_BYTES 64
typedef struct bogp_peer_s* bogp_peer_t;
int bogp_peer_create(bogp_peer_t* out, int rank, int world, unsigned char* ipc_handle_out);
int bogp_peer_open(bogp_peer_t p, const unsigned char* all_handles);
int bogp_peer_exchange(bogp_peer_t p, const long long* rec_dev, long long* recs_out_dev, void* stream);
int bogp_peer_destroy(bogp_peer_t p);
/* Pipelined form of the exchange (for a sequence of steps such as the time-step sweep of
 * projects/swellex96_localization/data/mfp.py:166-216, where the reduced record of a step is only consumed when its
 * results are written): with bogp_peer_set_pipelined(p, 1) every exchange -- stand-alone or fused into a surface call
 * below -- publishes this rank's record of step k without waiting for anybody and fills recs_out with the records of
 * step k - 1 (index -1 on the first step), so a step never ends on the slowest rank of the box; bogp_peer_collect
 * (one single-block kernel) waits for and returns the records of the last published step.  All ranks must switch
 * modes at the same step. */
int bogp_peer_set_pipelined(bogp_peer_t p, int on);
int bogp_peer_collect(bogp_peer_t p, long long* recs_out_dev, void* stream);
/* Surface + arg-ext + the exchange above in ONE call: record = (index_offset + flat index of this rank's extremum,
 * value bits); recs_out[world][2] gets every rank's record of the same step.  On the FP16-split tcgen05 engine the
 * exchange runs inside the surface kernel (its last CTA stores the record into the peers' buffers and waits for
 * theirs: no launch of its own); other engines pack the record and run the exchange kernel.  The _host form takes host
 * buffers like bogp_bartlett_grid_host_argmax and returns the records of all ranks after its one synchronisation (what
 * the reference's collate step reduces with np.argmax, projects/swellex96_localization/data/collate.py:50). */
int bogp_bartlett_grid_argext_peer(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host, int Nz,
                                   const double* tilt_deg_host, int Nt, int atype, int mf_method, int engine,
                                   float* out_dev, int maximize, bogp_peer_t peer, long long index_offset,
                                   long long* recs_out_dev, float* val_dev, long long* idx_dev, void* scratch_dev,
                                   long long scratch_bytes, void* stream);
int bogp_bartlett_grid_host_argmax_peer(bogp_modeset_t h, const double* r_m_host, int Nr, const double* z_host, int Nz,
                                        const double* tilt_deg_host, int Nt, int atype, int mf_method, int engine,
                                        float* out_host, int maximize, bogp_peer_t peer, long long index_offset,
                                        long long* recs_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BOGP_H */
