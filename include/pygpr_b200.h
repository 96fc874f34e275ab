/*
 * pygpr_b200 — C ABI of the B200-native (sm_100a) PyGPR hot path.
 *
 * The reference (sarath-srinivas/PyGPR) has NO FFI layer: its hot path is torch-CPU calls made
 * from Python classes.  Each entry point below replaces one group of those call sites; the
 * reference file:line it stands in for is cited on every declaration.  A maintainer binds them
 * with ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major FP64 unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it;
 *   - the library never allocates or frees persistent device memory: workspaces are passed in;
 *   - return value: 0 on success, negative gpr_status on failure (gpr_last_error() has the text);
 *     numerical failure of a factorisation is reported LAPACK-style through info[B]
 *     (0 = ok, k>0 = leading minor of order k is not positive definite), never as a return code; INT_MIN in every entry
 *     means the single-launch factorisation was aborted by its hang guard (a tile task waited > 30 s for its
 *     dependencies on a wedged GPU) — a scheduling stall, not a numerical verdict;
 *   - matrices that are factorised are stored PADDED to a multiple of GPR_TILE (=128) in both
 *     dimensions (`npad`); the padding block is the identity, so factors of the padded matrix
 *     equal the padded factors.  Inputs are stored pre-scaled by the ARD length-scales
 *     (a = x * ls, covar.py:137-140) with the feature dimension zero-padded to `dpad`
 *     (a multiple of 4, the DMMA k-step).
 */
#ifndef PYGPR_B200_H
#define PYGPR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GPR_TILE 128
#define GPR_ABI_VERSION 8

typedef enum {
    GPR_OK = 0,
    GPR_ERR_BAD_ARG = -1,
    GPR_ERR_CUDA = -2,
    GPR_ERR_UNSUPPORTED = -3
} gpr_status;

int gpr_abi_version(void);
const char* gpr_last_error(void);
/* developer A/B knobs of the schedules: "potrf_mode" = auto|columns|columns1|dataflow, "trsv_mode" = auto|single.  Their
 * initial values come from PYGPR_POTRF_MODE / PYGPR_TRSV_MODE, read once when the library is loaded. */
int gpr_set_option(const char* name, const char* value);
/* number of SMs / max opt-in shared memory of the current device (0 on error) */
int gpr_device_sm_count(void);

/* bytes of scratch gpr_predict needs for a given padded order (per launch, independent of B, m) */
long long gpr_predict_scratch_bytes(int npad);
/* bytes of scratch gpr_lml_grad needs */
long long gpr_lml_grad_scratch_bytes(int B, int npad, int d);

/*
 * a[b,i,:] = x[b,i,:] * ls[b,:]   (zero padded to [npad, dpad]),   nrm[b,i] = sum_k a[b,i,k]^2
 * Replaces covar.py:137-140 (x.mul(ls)) and covar.py:106 (x.square().sum) — call sites C1.
 *   x        [B or 1, n, d]   x_bstride = n*d, or 0 when one x is shared by all experts
 *   hp       row b at hp + b*hp_stride (hp_stride = 0 shares one row); ls = hp[1..d]
 */
int gpr_se_scale(const double* x, long long x_bstride, const double* hp, int hp_stride,
                 int B, int n, int d, int npad, int dpad, double* a, double* nrm, void* stream);

/*
 * K[b] = sig_b^2 exp(-(|a_i|^2 + |a_j|^2 - 2 a_i.a_j)) + diag_add_b * I      (GEMM form, DMMA cross term)
 * Replaces Squared_exponential.kernel covar.py:129-167 (+ White_noise covar.py:227-240 and the
 * jitter gpr.py:68 through diag_add) — call sites C1-C3.
 *   a,nrm     from gpr_se_scale ([B,npad,dpad], [B,npad])
 *   hp        sig = hp[b*hp_stride + 0]
 *   diag_add  [B] or NULL
 *   K         [B, ldk, ldk] with ldk >= n.  rows/cols >= n (up to ldk) are written as identity
 *             when pad_identity != 0, left untouched otherwise.
 *   lower_only != 0 writes only tiles on or below the block diagonal.
 */
int gpr_se_build(const double* a, const double* nrm, const double* hp, int hp_stride,
                 const double* diag_add, int B, int n, int npad, int dpad,
                 double* K, int ldk, int pad_identity, int lower_only, void* stream);

/*
 * Ks[b,i,j] = sig_b^2 exp(-|as_i - a_j|^2)   rows = test points, cols = training points
 * Replaces the xp branch of Squared_exponential.kernel covar.py:151-160 — call site C4.
 *   as,nrms  [B, mpad, dpad],[B, mpad] scaled test points;  a,nrm [B, npad, dpad],[B,npad]
 *   Ks       [B, m, n] dense (no padding)
 */
int gpr_se_cross(const double* as, const double* nrms, int mpad, const double* a, const double* nrm, int npad,
                 const double* hp, int hp_stride, int B, int m, int n, int dpad, double* Ks, void* stream);
/* The same with an explicit leading dimension / batch stride of Ks (ld >= n, bstride >= m * ld): the emulated prediction writes
 * K* with rows padded to whole k-blocks of the digit-plane split (the padding columns are the caller's, kept zero). */
int gpr_se_cross_ld(const double* as, const double* nrms, int mpad, const double* a, const double* nrm, int npad,
                    const double* hp, int hp_stride, int B, int m, int n, int dpad, double* Ks, long long ld, long long bstride,
                    void* stream);
/*
 * K* written straight as the digit planes of gpr_ozaki_split(Ks, bound = sig^2) — bit for bit the planes / scales that
 * gpr_se_cross_ld followed by gpr_ozaki_split(bound) produce, without the FP64 round trip through HBM (covar.py:151-160 feeding
 * gpr.py:100 on the emulated path).  Every row tile whole: the test points come padded to mpad rows.
 *   planes [slices][B * mpad][npad] int8 (columns n .. npad zero), scale [B * mpad];  bound >= max |K*| (sig^2 of the experts)
 * gpr_se_cross_planes_ok: 1 when the shared-memory staging fits for this input width / plane count.
 */
int gpr_se_cross_planes_ok(int dpad, int slices);
int gpr_se_cross_planes(const double* as, const double* nrms, int mpad, const double* a, const double* nrm, int npad,
                        const double* hp, int hp_stride, int B, int n, int dpad, int slices, double bound, void* planes,
                        double* scale, void* stream);

/*
 * dK[b,0] = 2 K_se/sig ; dK[b,1+k] = -2 ls_k (x_ik - x_jk)^2 K_se      (covar.py:169-206)
 * Materialises the gradient tensor for API parity of kernel_and_grad only; MLE never calls it.
 *   Kse [B,n,n] dense, x [B,n,d], dK [B,d+1,n,n]
 */
int gpr_se_kernel_grad(const double* Kse, const double* x, const double* hp, int hp_stride,
                       int B, int n, int d, double* dK, void* stream);

/*
 * Batched lower Cholesky, blocked left-looking, trailing products on FP64 tensor cores (DMMA).
 * Replaces tc.cholesky gpr.py:69 / loss.py:39,64,97 (call site C5) and, through y/z, the forward
 * half of tc.cholesky_solve gpr.py:70-72 / loss.py:44 (C6) and the NLML scalar loss.py:49-51 (C14).
 *
 *   L        [B, npad, npad] in: see `gen`; out: lower factor (strict upper triangle of each
 *            diagonal block zeroed; blocks above the block diagonal are used as scratch)
 *   gen == 0 the lower triangle of K (+ identity padding) is already in L (in-place factorisation)
 *   gen != 0 K is generated on the fly from (a, nrm, hp, diag_add) — L is write-only
 *   invdiag  [B, npad/128, 128, 128] out: inverses of the diagonal blocks of L (lower triangular)
 *   y        [B, npad] or NULL (zero padded).  If given:  z = L^-1 y  is written to z [B, npad]
 *   logdet_half [B] out: sum_i log L_ii          (NULL allowed)
 *   zz       [B] out: z.z = y^T K^-1 y            (NULL allowed; needs y)
 *   info     [B] int32 out
 */
int gpr_potrf_batched(double* L, int gen, const double* a, const double* nrm, const double* hp, int hp_stride,
                      const double* diag_add, int B, int n, int npad, int dpad,
                      double* invdiag, const double* y, double* z, double* logdet_half, double* zz,
                      int* info, void* work, void* stream);
/* bytes of `work` for gpr_potrf_batched (split-K slots + counters used when a launch has fewer tiles than SMs);
 * work == NULL is allowed and disables the split (slower for few experts, same results up to summation order) */
long long gpr_potrf_work_bytes(int B, int npad);

/*
 * The same factorisation for experts that all start with the SAME ng rows under the SAME hyper-parameters — GRBCM's local
 * experts, whose training sets are the communication set followed by their own block (gr_bcm.py:19-26), when one
 * hyper-parameter vector is shared.  The leading ng x ng block of every K is then the communication expert's matrix, so its
 * factor (Lg, invdiag_g from gpr_potrf_batched on (xg, yg), gr_bcm.py:28) is read, not recomputed: only block rows >= ng/128
 * are produced (7/8 of the flops at ng = n/2).  The caller pre-loads z[b, 0:ng], logdet_half[b], zz[b], info[b] with the
 * communication expert's values (they are accumulated onto) and copies Lg / invdiag_g into the leading blocks of L / invdiag
 * if a later call (gpr_trsv_backward, gpr_lml_grad, ...) needs the full factor.  ng must be a multiple of 128.
 */
int gpr_potrf_batched_shared(double* L, const double* a, const double* nrm, const double* hp, int hp_stride,
                             const double* diag_add, int B, int n, int npad, int dpad, double* invdiag, const double* y,
                             double* z, double* logdet_half, double* zz, int* info, void* work, void* stream, int ng,
                             const double* Lg, int npad_g, const double* invdiag_g);

/*
 * gpr_potrf_batched (communication expert, B = 1) and gpr_potrf_batched_shared (local experts) as ONE launch: the
 * communication expert's tile tasks lead every shared block column and the local tiles wait on its progress counters, so
 * its serial chain of ng/128 block columns runs under the local experts' work.  *_g arguments as for gpr_potrf_batched with
 * B = 1 (hp_g: one row); outputs of both models as in the two separate calls; z[b, 0:ng] is filled by the kernel.
 * ng need not be a multiple of 128: the floor(ng/128) whole blocks are shared, the ragged rest is each model's own.
 */
int gpr_potrf_grbcm(double* L, const double* a, const double* nrm, const double* hp, int hp_stride, const double* diag_add,
                    int B, int n, int npad, int dpad, double* invdiag, const double* y, double* z, double* logdet_half,
                    double* zz, int* info, double* Lg, const double* a_g, const double* nrm_g, const double* hp_g,
                    const double* diag_add_g, int ng, int npad_g, double* invdiag_g, const double* y_g, double* z_g,
                    double* logdet_half_g, double* zz_g, int* info_g, void* work, void* stream);
long long gpr_potrf_grbcm_work_bytes(int B, int npad, int npad_g);

/* out[0] = sum_{i<rows} log L_ii, out[1] = sum_{i<rows} z_i^2 of a finished B = 1 factorisation: the values gpr_potrf_batched_shared's
 * logdet_half / zz are pre-loaded with when only the first `rows` rows are shared. */
int gpr_factor_partials(const double* L, int npad, const double* z, int rows, double* out, void* stream);

/* L[b, 0:ng, 0:ng] = Lg[0:ng, 0:ng] for every expert b (tile granular; lower_only: 0 all tiles, 1 tiles on / below the block
 * diagonal, 2 tiles above it) — the copy gpr_potrf_batched_shared / gpr_potrf_grbcm ask the caller for. */
int gpr_copy_leading_block(double* L, int B, int npad, const double* Lg, int npad_g, int ng, int lower_only, void* stream);

/*
 * alpha = L^-T z  (backward half of cholesky_solve, gpr.py:70-72).  HBM-bound: reads L once.
 *   z [B,npad] in, alpha [B,npad] out (may alias z only when work == NULL)
 *   work  >= gpr_trsv_work_bytes(B) or NULL.  With a workspace and fewer experts than SMs the solve runs as a
 *         cooperative launch of several column-owner CTAs per expert (one CTA per expert cannot saturate HBM).
 */
int gpr_trsv_backward(const double* L, const double* invdiag, int B, int npad,
                      const double* z, double* alpha, void* work, void* stream);
long long gpr_trsv_work_bytes(int B);

/*
 * Batched prediction: mean and variance of every expert at m test points, in one persistent kernel.
 *   V^T = K* L^-T  by blocked forward substitution (DMMA), K* tiles generated on the fly (never in HBM)
 *   mean[b,c] = sum_j K*[c,j] alpha[b,j]                                   gpr.py:79-85   (C4, C7)
 *   var[b,c]  = prior_b - sum_j V[j,c]^2 ,  prior_b = sig_b^2 + noise_prior_b   gpr.py:96-106 (C8, C9)
 *   xs       [m, d] raw test points shared by all experts;  hp as above (sig, ls)
 *   prior    [B] prior variance at a test point (sig^2 + sig_n^2: K** includes the noise, gpr.py:98)
 *   want_var == 0 skips the triangular solve (SK_WRAP path, scikit_model.py:34)
 *   scratch  >= gpr_predict_scratch_bytes(npad)
 */
int gpr_predict(const double* L, const double* invdiag, const double* a, const double* nrm,
                const double* alpha, const double* hp, int hp_stride, const double* prior,
                const double* xs, int B, int n, int npad, int d, int dpad, int m,
                double* mean, double* var, int want_var, void* scratch, void* stream);

/* gpr_predict that also keeps  V^T = K* L^-T  in vt [B, ceil(m/128), 128, npad] (rows >= m of the last tile are padding). */
int gpr_predict_keep(const double* L, const double* invdiag, const double* a, const double* nrm, const double* alpha,
                     const double* hp, int hp_stride, const double* prior, const double* xs, int B, int n, int npad,
                     int d, int dpad, int m, double* mean, double* var, double* vt, void* scratch, void* stream);

/*
 * Prediction of experts that share their first ng training rows and their hyper-parameters with a communication expert
 * (see gpr_potrf_batched_shared).  vt_g / mean_g / var_g come from gpr_predict_keep on the communication expert for the same
 * test points; z = L^-1 y of the local experts [B, npad].  Each expert is evaluated as an update of the communication
 * expert:  mean = mean_g + V_l^T z_l ,  var = var_g - |V_l|^2  — identical to gpr_predict's K* alpha and prior - |V|^2
 * (gpr.py:79-106) because the first ng rows of V and z are the communication expert's; 3/4 of the flops at ng = n/2.
 */
int gpr_predict_shared(const double* L, const double* invdiag, const double* a, const double* nrm, const double* z,
                       const double* hp, int hp_stride, const double* xs, int B, int n, int npad, int d, int dpad, int m,
                       double* mean, double* var, void* scratch, void* stream, int ng, const double* vt_g, int npad_g,
                       const double* mean_g, const double* var_g);

/*
 * X[b] = L[b]^-1 (lower triangular, row-major, [B, npad, npad]) from a factor of gpr_potrf_batched; overwrites the strict
 * block-upper part of L with L^-T on the way (like gpr_potri_batched).  scratch >= gpr_potri_scratch_bytes(B, npad).
 */
int gpr_tril_inverse(double* L, const double* invdiag, int B, int npad, double* X, void* scratch, void* stream);

/*
 * gpr_predict_keep in GEMM form:  V^T = K* X^T  with X = L^-1 from gpr_tril_inverse — (m/128) x (n/128) independent tiles
 * instead of ceil(m/128) serial substitutions, for the single communication expert whose V^T every local expert waits for.
 *   mean = V^T z (= K* alpha, gpr.py:79-85),  var = prior - |V|^2 (gpr.py:96-106),  vt [B, ceil(m/128), 128, npad]
 *   z = L^-1 y [B, npad];  scratch >= gpr_predict_gemm_scratch_bytes(B, npad, dpad, m)
 *   mean_part / var_part (or NULL): the same sums restricted to the first rows_part training rows (a multiple of 128) — what
 *   gpr_predict_shared continues from when only those rows are shared with the local experts
 */
int gpr_predict_gemm(const double* X, const double* a, const double* nrm, const double* z, const double* hp, int hp_stride,
                     const double* prior, const double* xs, int B, int n, int npad, int d, int dpad, int m, double* mean,
                     double* var, double* vt, void* scratch, void* stream, int rows_part, double* mean_part, double* var_part);
long long gpr_predict_gemm_scratch_bytes(int B, int npad, int dpad, int m);

/*
 * Negative log marginal likelihood gradient without forming dK:
 *   U = L^-T (blocked, DMMA) written into the strict block-upper part of L's buffer, then
 *   K^-1 = U U^T tile by tile with a fused epilogue reducing  W = K^-1 - alpha alpha^T  against
 *   K_se generated on the fly:   loss.py:59-128 (C11-C13)
 *     grad[b,0]     = (1/sig)  sum_ij W_ij Kse_ij
 *     grad[b,1+k]   = -ls_k    sum_ij W_ij Kse_ij (x_ik - x_jk)^2
 *     grad[b,d+1]   = sig_n tr(W)                       (only when has_noise)
 *   L, invdiag, a, nrm, alpha from the calls above; grad [B, d+1+has_noise]
 */
int gpr_lml_grad(double* L, const double* invdiag, const double* a, const double* nrm, const double* alpha,
                 const double* hp, int hp_stride, int has_noise, int B, int n, int npad, int d, int dpad,
                 double* grad, void* scratch, void* stream);

/*
 * gpr_lml_grad for experts factorised by gpr_potrf_batched_shared, when the leading block copied into L also carries the
 * communication expert's L_g^-T in its strict block-upper part (run gpr_potri_batched on the communication factor BEFORE
 * copying it).  Kinv_g [ng, ldg] is that call's K_g^-1.  The shared blocks of L^-T and of K^-1 = U U^T are not recomputed.
 */
int gpr_lml_grad_shared(double* L, const double* invdiag, const double* a, const double* nrm, const double* alpha,
                        const double* hp, int hp_stride, int has_noise, int B, int n, int npad, int d, int dpad,
                        double* grad, void* scratch, void* stream, int ng, const double* Kinv_g, int ldg);

/*
 * GRBCM aggregation (diag), gr_bcm.py:116-149 (C15), split for expert sharding:
 *   partial : S[0,c] = sum_i beta_i, S[1,c] = sum_i beta_i/var_i, S[2,c] = sum_i beta_i mean_i/var_i
 *             over the nc_local experts held here; beta_i = 1 for global expert index 0
 *             (first_global_index + i == 0), else 0.5 (log var_g - log var_i).
 *             Optionally writes beta/prec rows (row stride m) for the reference's self.beta/self.prec.
 *   finalize: var_A = 1/(S1 + (1-S0)/var_g),  mu_A = var_A (S2 + (1-S0) mean_g/var_g)
 */
int grbcm_partial(const double* mean_l, const double* var_l, const double* var_g,
                  int nc_local, int first_global_index, int m, double* S,
                  double* beta_rows, double* prec_rows, void* stream);
int grbcm_finalize(const double* S, const double* mean_g, const double* var_g, int m,
                   double* mu, double* var, double* beta0, double* prec0, void* stream);

/* ---- "next" rows: full predictive covariance and its GRBCM aggregation ------------------- */

/* scratch bytes for gpr_predict_full (holds V^T [B, mpad, npad]) and gpr_potri_batched */
long long gpr_predict_full_scratch_bytes(int B, int npad, int m);
long long gpr_potri_scratch_bytes(int B, int npad);

/*
 * mean, variance AND full predictive covariance  cov[b] = K**(xs,xs) + noise_var_b I - K* K^-1 K*^T   [B, m, m]
 * Replaces Exact_GP.predict_covar gpr.py:108-120 (and predict(var="full"), gpr.py:86-87).
 *   as, nrms  scaled test points from gpr_se_scale(xs, hp) ([B, mpad, dpad], [B, mpad]); mpad = ceil(m/128)*128
 *   noise_var [B] sig_n^2 (K** = cov.kernel(params, xp) includes the white noise, no jitter, gpr.py:110) or NULL
 */
int gpr_predict_full(const double* L, const double* invdiag, const double* a, const double* nrm, const double* alpha,
                     const double* hp, int hp_stride, const double* prior, const double* noise_var, const double* xs,
                     const double* as, const double* nrms, int B, int n, int npad, int d, int dpad, int m, int mpad,
                     double* mean, double* var, double* cov, void* scratch, void* stream);

/*
 * Kinv[b] = (L L^T)^-1 from a factor produced by gpr_potrf_batched (overwrites the block-upper part of L with L^-T).
 * Replaces tc.cholesky_solve(identity, chol) gr_bcm.py:101-103 and tc.cholesky_inverse gr_bcm.py:112.
 *   Kinv [B, ldo, ldo], only the leading n x n block is written (both triangles)
 */
int gpr_potri_batched(double* L, const double* invdiag, int B, int n, int npad, double* Kinv, int ldo, void* scratch,
                      void* stream);

/* prec[i,j] = sum_c Kinv[c,i,j] * (beta[c,i] + beta[c,j]) / 2      GRBCM.aggregate_full_covar gr_bcm.py:105-111 */
int grbcm_full_combine(const double* Kinv, const double* beta, int nc1, int m, double* prec, void* stream);

/* ---- dense path: model-level support for ANY Covar composition (covar.py:28-81), kernel matrices materialised ---------------- */
/*
 * V^T = Ks X^T with X = L^-1 (gpr_tril_inverse) for materialised cross-covariances Ks [B, mpad, npad] (zero padded,
 * mpad = ceil(m/128)*128):  mean = V^T z = Ks alpha (gpr.py:79-85), var = prior[b, i] - |V_i|^2 (gpr.py:96-106), vt [B, mpad, npad].
 * prior [B, m] is the prior variance k(x*_i, x*_i) PER TEST POINT — diag(cov.kernel(params, xp)), gpr.py:98,102 — so that
 * non-stationary compositions of the Covar protocol are handled like the reference handles them.
 */
int gpr_predict_dense(const double* X, const double* Ks, const double* z, const double* prior, int B, int npad, int m,
                      double* mean, double* var, double* vt, void* scratch, void* stream);
long long gpr_predict_dense_scratch_bytes(int B, int npad, int m);
/* cov[b] = Kss[b] - V^T V   (gpr.py:108-120), Kss [B, m, m] materialised, vt from gpr_predict_dense */
int gpr_covar_dense(const double* vt, const double* Kss, int B, int npad, int m, double* cov, void* stream);
/* grad[b,p] = 1/2 sum_ij (Kinv[b] - alpha_b alpha_b^T)_ij dK[b,p,i,j]   (loss.py:77-84,116-121), dK [B, P, n, n] materialised */
int gpr_trace_grad(const double* Kinv, const double* alpha, int alpha_stride, const double* dK, int B, int P, int n,
                   double* grad, void* scratch, void* stream);
long long gpr_trace_grad_scratch_bytes(int B, int P, int n);

/* ---- "next" row N3: data partitioning (sampler.py) -------------------------------------------------------- */
/*
 * D[i,j] = |x_i|^2 + |y_j|^2 - 2 x_i.y_j  (GEMM form, cross term on DMMA) — euclidean_dist sampler.py:94-100, used by
 * cluster_samples / MATERN1.cluster_samples (nearest-centre assignment, sampler.py:62-80,103-119).
 *   ax,nx / ay,ny  from gpr_se_scale with unit length-scales ([1,npad,dpad],[1,npad]);  D [nx_rows, ny_rows]
 */
int gpr_sq_dist(const double* ax, const double* nx, int npadx, const double* ay, const double* ny, int npady,
                int nx_rows, int ny_rows, int dpad, double* D, void* stream);

/* ---- "next" row N4: FP64-accurate products on the INT8 tcgen05 tensor cores (emulated FP64 / reduced-precision opt-in) ----- */
/*
 * C = beta*C + A B^T with A [M, K], B [N, K] given as balanced radix-256 digit planes (Ozaki-style splitting): stands in for
 * the FP64 GEMMs behind tc.cholesky_solve(K*^T, L) gpr.py:100 and tc.bmm gpr.py:80,115 wherever they are expressed as a
 * product with a cached L^-1.  Every plane product runs on tcgen05.mma kind::i8 with exact int32 accumulation in TMEM;
 * the FP64 recombination happens in the epilogue.  `slices` (1..8) is the accuracy knob: the result carries
 * ~2^(2-8*slices) of the product of the row maxima (6: FP64 grade, 3-4: the FP32-grade opt-in mode of north_star).
 *   gpr_ozaki_split      X [rows, k] (leading dimension ld) -> planes [slices][rpad][kpad] int8 + scale [rpad]
 *                        (rpad = gpr_ozaki_rows_padded(rows), kpad = gpr_ozaki_k_padded(k); planes of
 *                        gpr_ozaki_planes_bytes bytes; padding rows/columns are zero).  bound > 0: the caller vouches that
 *                        |x| <= bound everywhere and every row is scaled by that bound (one pass instead of two);
 *                        bound <= 0: each row is scaled by its own maximum.  batch > 1: `batch` matrices x_batch_stride
 *                        doubles apart (e.g. the same window of every expert's factor), planes / scales stacked along the
 *                        padded row dimension ([slices][batch * rpad][kpad], [batch * rpad]) as gpr_ozaki_gemm_nt's batch wants
 *   gpr_ozaki_gemm_nt    both operands split with the SAME slices and k;  b_lower = -1: dense B;  b_lower = o >= 0:
 *                        B[j, k] = 0 for k > j + o (a lower-triangular / lower-trapezoidal B, e.g. rows of L^-1) and the zero
 *                        k-blocks are skipped;  K * slices < 2^17 (int32 headroom).  batch > 1: `batch` independent products
 *                        in one launch — the plane arrays hold the operands of all batch entries stacked along the (padded)
 *                        row dimension (split the stacked matrices; rows must then be multiples of 128), C at c_batch_stride.
 *                        a_upper = o >= 0: A[i, k] = 0 for k < i + o (rows of an upper-triangular factor; the leading zero
 *                        k-blocks are skipped), -1: dense A.  c_lower != 0: only the 128 x 64 tiles that touch the lower
 *                        triangle of C are computed and written (K^-1 = U U^T for the NLML gradient, loss.py:77,116).
 *                        C = beta*C + alpha * A B^T.  b_shared != 0 (with batch > 1): ONE B operand (planesB holds N rows) is
 *                        used by every batch entry (e.g. the communication expert's L_g^-1 against every local expert's block).
 */
long long gpr_ozaki_rows_padded(int rows);
long long gpr_ozaki_k_padded(int k);
long long gpr_ozaki_planes_bytes(int rows, int k, int slices);
int gpr_ozaki_split(const double* X, long long ld, int rows, int k, int slices, double bound, void* planes, double* scale,
                    int batch, long long x_batch_stride, void* stream);
int gpr_ozaki_gemm_nt(const void* planesA, const double* scaleA, int M, const void* planesB, const double* scaleB, int N,
                      int K, int slices, int b_lower, double beta, double* C, long long ldc, int batch, long long c_batch_stride,
                      int a_upper, int c_lower, double alpha, int b_shared, void* stream);
/* The same on a WINDOW of rows of each operand: every batch entry of A occupies a_rows_total (padded) rows of planesA / scaleA and the
 * product uses the M rows starting a_row_off rows in (likewise B) — a block-row range of an operand that was split once, e.g. the parts
 * of a Cholesky panel that update the next panel's columns and the rest of the trailing matrix separately (engine.fit_emulated).
 * Offsets are multiples of 64 rows; a_rows_total / b_rows_total = 0 means the whole operand (gpr_ozaki_gemm_nt). */
int gpr_ozaki_gemm_nt_window(const void* planesA, const double* scaleA, int M, int a_rows_total, int a_row_off, const void* planesB,
                             const double* scaleB, int N, int b_rows_total, int b_row_off, int K, int slices, int b_lower, double beta,
                             double* C, long long ldc, int batch, long long c_batch_stride, int a_upper, int c_lower, double alpha,
                             int b_shared, void* stream);

/* pieces of the emulated-FP64 prediction (opt-in): V_l^T = K* X_l^T with X_l = rows of the explicit inverse L^-1, the product on
 * the INT8 tensor cores (gpr_ozaki_gemm_nt).  Stands in for tc.cholesky_solve(K*^T, L) + the row sums of gpr.py:100-104 and the
 * tc.bmm of gpr.py:80-83 for experts that extend the communication expert.
 *   gpr_trtri_batched      U = L^-T in place in the strict block-upper part of L [B, npad, npad]; block columns < gt must already
 *                          hold the shared L_g^-T (gpr_copy_leading_block mode 2); scratch: gpr_potri_scratch_bytes(B, npad).
 *                          seg > 0: only the diagonal segments of seg block rows are inverted (the blocks between segments are
 *                          products of segment inverses: U_AB = -U_AA L_BA^T U_BB, formed by gpr_ozaki_gemm_nt)
 *   gpr_tril_inverse_rows  X [B, npad - row0, npad] = rows row0.. of L^-1 (dense row-major, zeros right of the diagonal block)
 *   gpr_vt_reduce          mean[e, t] = mean_g[t] + sum_c V[e, t, c] z[e, c],  var[e, t] = var_g[t] - sum_c V[e, t, c]^2
 *                          (V [E, m, ldv] with nl used columns; z rows z_stride apart)
 */
int gpr_trtri_batched(double* L, const double* invdiag, int B, int npad, void* scratch, void* stream, int gt, int seg);
/* window [r0, r0+nr) x [c0, c0+nc) of U = L^-T (kind 0) or X = L^-1 (kind 1) as a dense block out [B, nr, nc] (multiples of 32),
 * read out of the in-place L^-T and the diagonal-block inverses: operands of the products that assemble L^-T from diagonal segments */
int gpr_factor_window(const double* L, const double* invdiag, int B, int npad, int kind, int r0, int nr, int c0, int nc, double* out,
                      void* stream);
int gpr_tril_inverse_rows(const double* L, const double* invdiag, int B, int npad, int row0, double* X, void* stream);
int gpr_vt_reduce(const double* V, long long ldv, const double* z, long long z_stride, const double* mean_g, const double* var_g,
                  int E, int m, int nl, double* mean, double* var, void* stream);

/* pieces of the emulated-FP64 NLML gradient (opt-in): K^-1 = U U^T (U = L^-T, tc.cholesky_solve(dK, L) + traces of loss.py:77-84,
 * 116-121 reformulated as in gpr_lml_grad) with the product on the INT8 tensor cores.
 *   gpr_triu_rows        Ud [B, npad, npad] = the rows of U as a dense upper-triangular matrix (zeros below the diagonal blocks),
 *                        out of the in-place L^-T in the block-upper part of L and the diagonal-block inverses
 *   gpr_grad_from_kinv   grad [B, P] from materialised K^-1 (lower 128 x 128 tiles of Kinv [B, ldk, ldk] read): the epilogue of
 *                        gpr_lml_grad — M = (K^-1 - alpha alpha^T) o K_se with K_se generated on chip, reduced per hyper-parameter;
 *                        scratch: gpr_lml_grad_scratch_bytes(B, npad, d)
 */
int gpr_triu_rows(const double* L, const double* invdiag, int B, int npad, double* Ud, void* stream);
int gpr_grad_from_kinv(const double* Kinv, int ldk, const double* a, const double* nrm, const double* alpha, const double* hp,
                       int hp_stride, int has_noise, int B, int n, int npad, int d, int dpad, double* grad, void* scratch, void* stream);

/* y[b, i] -= sum_k A[b, i, k] z[b, k]   for i < rows, k < kdim: A = a window of every expert's factor (row stride lda, experts
 * a_batch_stride apart), y / z rows y_stride / z_stride apart.  The forward solve's update of the remaining right-hand side
 * behind a finished block column (z = L^-1 y, gpr.py:70-72) in the panel-wise factorisation of engine.fit_emulated.  HBM bound. */
int gpr_gemv_sub(const double* A, long long lda, long long a_batch_stride, int B, int rows, int kdim, const double* z,
                 long long z_stride, double* y, long long y_stride, void* stream);

/*
 * Communicator — for consumers without torch.distributed: the two collectives of the path, all-reduce(sum, f64) of the
 * [3, m] partial sums of GRBCM.aggregate (gr_bcm.py:133,143-144; grbcm_partial -> HERE -> grbcm_finalize) and of the [1+P] NLML /
 * gradient scalars of a shared-vector training step.  Thin wrappers over NCCL, bound at run time (dlopen "libnccl.so.2"):
 * GPR_ERR_UNSUPPORTED where NCCL is absent, every other entry point is unaffected.
 *   gpr_comm_unique_id        rank 0 fills id[gpr_comm_unique_id_bytes()] (128) and hands it to the other ranks by its own means
 *   gpr_comm_init             collective; each rank on its own current CUDA device
 *   gpr_comm_allreduce_f64    in place on a device buffer, ordered on `stream`
 */
int gpr_comm_unique_id_bytes(void);
int gpr_comm_unique_id(void* id);
int gpr_comm_init(void** comm, int world, int rank, const void* id);
int gpr_comm_allreduce_f64(void* comm, double* buf, long long count, void* stream);
int gpr_comm_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif
