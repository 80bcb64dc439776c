/*
 * cosmolike.h — C ABI of libcosmolike_b200.so, the B200-native batched cosmological likelihood engine.
 *
 * This is the drop-in boundary for ONE hot path of franciscotln/cosmology-model-fit: evaluating
 * chi^2(theta) / log L(theta) / log P(theta) for whole batches of parameter vectors.  The reference has no
 * FFI; its de-facto operator interface is a set of module-level Python callables repeated in every fit
 * script.  Each entry point below names the reference interface it replaces.
 *
 * Conventions: plain C, plain pointers and sizes, no Python/torch types.  All floating point is IEEE
 * float64.  Every function returns 0 on success or a negative CL_E_* code; cl_last_error() gives the text.
 * There is NO CPU fallback: cl_create() fails with CL_E_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef COSMOLIKE_H
#define COSMOLIKE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CL_ABI_VERSION 4u
#define CL_NCCL_UID_BYTES 128 /* sizeof(ncclUniqueId) */
#define CL_MAX_DIM 12      /* max length of one parameter vector theta */
#define CL_MAX_VEL 3       /* max peculiar-velocity template amplitudes (step: 1, dipole xyz: 3) */
#define CL_MAX_GAUSS 4     /* max extra Gaussian terms of each kind */
#define CL_MAX_BAO 32      /* max BAO data points */
#define CL_MAX_GL 128     /* max Gauss-Legendre nodes */
#define CL_MAX_CC 64       /* max cosmic-chronometer points */
#define CL_SN_SMALL_MAX 64 /* SN blocks up to this size use the in-kernel quadratic form (no GEMM) */

/* chi-squared engines for large SN blocks (option "chi2_engine"):
 *   DMMA    — FP64 tensor pipe (mma.sync f64), bit-stable FP64 contraction;
 *   TCGEN05 — 5th-generation tensor cores: the residual rows and W = L^-1 are split into "chi2_slices" (5..7, default 6)
 *             int8 digit planes, multiplied exactly with tcgen05.mma kind::i8 (int32 accumulators in TMEM) and recombined in
 *             FP64.  7 planes (default) carry all 53 bits of every row: same accuracy class as the FP64 engine
 *             (|d chi2| ~ 1e-10 at chi2 ~ 6e4); 6 planes carry 46 bits: |d chi2| / chi2 ~ 2e-12, 30 % faster.
 * Default: TCGEN05 with 7 planes. */
enum { CL_CHI2_ENGINE_DMMA = 0, CL_CHI2_ENGINE_TCGEN05 = 1 };

/* error codes */
enum {
  CL_OK = 0,
  CL_E_INVALID = -1,   /* bad argument / inconsistent spec */
  CL_E_NO_DEVICE = -2, /* no CUDA device of compute capability 10.x */
  CL_E_CUDA = -3,      /* CUDA runtime/driver error, see cl_last_error */
  CL_E_NOMEM = -4,
  CL_E_NUMERIC = -5    /* covariance factor not invertible etc. */
};

/* E(z) family.  LATE: H = H0 sqrt(Om (1+z)^3 + (1-Om) f_DE(z))            (sn/pantheon.py:28-31, bao/desi.py:31-35)
 *               FULL: H = H0 sqrt(Or(1+z)^4 + Obc(1+z)^3 + Onu*Omnu_z(z) + Ode f_DE(z)), densities from
 *                     h = H0/100, omega_b, omega_c and the constants of one cmb/data_*_compression.py
 *                     (bao/desi_cmb_union3.py:37-57, bao/desi_des5y_bbn_theta_star.py:31-52, cmb/cmb.py:11-25) */
enum { CL_FAMILY_LATE = 0, CL_FAMILY_FULL = 1 };

/* dark-energy density factor f_DE(z) = rho_DE(z)/rho_DE(0)
 *   LCDM 1 | WCDM (1+z)^(3(1+w0)) | CPL (1+z)^(3(1+w0+wa)) exp(-3 wa z/(1+z))   (bao/desi_fs_lya_cmb.py:18-22)
 *   THAWING (2(1+z)^3 / ((1+w0) + (1-w0)(1+z)^3))^2                                (bao/desi.py:25-28, README.md:29) */
enum { CL_DE_LCDM = 0, CL_DE_WCDM = 1, CL_DE_CPL = 2, CL_DE_THAWING = 3 };

/* SN covariance operand.  CHOLESKY: lower factor L of C (scipy cho_factor(lower=True)[0]); chi2 = |L^-1 d|^2
 *                          (solve_triangular.py:5-14, sn/pantheon.py:14,61).
 *                          INVCOV:   C^-1; chi2 = d^T C^-1 d (sn/union3_1.py:8,57). */
enum { CL_SN_CHOLESKY = 0, CL_SN_INVCOV = 1 };

/* peculiar-velocity redshift shift: DIVIDE   z_cosmo = (1+z_cmb)/(1+s) - 1       (sn/pantheon.py:43-49)
 *                                   MULTIPLY z_cosmo = max((1+z_cmb)(1+s) - 1, 1e-8) (bao/desi_pantheon_cc.py:83-90)
 * with s = vel_scale * sum_k theta[col_vel[k]] * weight[k][i] / c */
enum { CL_VEL_DIVIDE = 0, CL_VEL_MULTIPLY = 1 };

/* BAO quantity codes (bao/desi_cmb_union3.py:72) */
enum { CL_BAO_DV_OVER_RS = 0, CL_BAO_DM_OVER_RS = 1, CL_BAO_DH_OVER_RS = 2, CL_BAO_F_AP = 3 };
/* D_H at BAO redshifts: EXACT c/H(z) (bao/desi_cmb_pantheon.py:61-63) or PCHIP through dh_grid
 * (interpolator.py:111-114, bao/desi_cmb_union3.py:83) */
enum { CL_DH_EXACT = 0, CL_DH_PCHIP = 1 };
/* sound horizon at drag: FIXED constant (bao/desi.py:10), PARAM theta column (bao/desi_des5y_rd.py:68),
 * FIT r_drag(omega_b, omega_m) (cmb/data_planck_act_compression.py:102-124) */
enum { CL_RD_FIXED = 0, CL_RD_PARAM = 1, CL_RD_FIT = 2 };

/* compressed-CMB vector: R_LA_WB (R, l_A, omega_b) (cmb/data_planck_act_compression.py:200-212);
 *                        THETA_WB_WM (theta*, omega_b, omega_m) (cmb/data_early_lcdm_compression.py:200-207) */
enum { CL_CMB_NONE = 0, CL_CMB_R_LA_WB = 1, CL_CMB_THETA_WB_WM = 2 };

/* what cl_eval writes */
enum {
  CL_OUT_CHI2 = 0,    /* chi_squared(theta)                                       (sn/pantheon.py:57-61) */
  CL_OUT_LOGLIKE = 1, /* log_likelihood(theta) = -chi2/2 (+ CC normalisation), guards applied
                         (sn/pantheon.py:64-65, bao/desi_fs_lya_cmb.py:117-121) */
  CL_OUT_LOGPROB = 2  /* log_probability(theta) = log_prior + log_likelihood, -inf outside the box
                         (sn/pantheon.py:80-97) */
};

/* Constants of one cmb/data_*_compression.py module (the five modules differ only in these numbers). */
typedef struct cl_cmb_consts {
  double Or_h2;        /* Omega_r h^2 used in E(z)            (cmb/data_planck_act_compression.py:39-43) */
  double Omnu_h2;      /* massive-neutrino Omega_nu h^2       (:36) */
  double Ogamma_h2;    /* photon density for R_b in r_s       (:29) */
  double nu_m0;        /* m_nu / T_nu0                        (:35) */
  double nu_rho0;      /* compute_rho0(m0)                    (:47, nu_evolution.py:23-28) */
  double nu_q2[5];     /* qs**2                               (:48-49, nu_evolution.py:10-16) */
  double nu_w[5];      /* 5-node weights                      (nu_evolution.py:20) */
  double zstar_s1, zstar_s2, zstar_b, zstar_m; /* z_star fit  (:86-99) */
  double rdrag_b, rdrag_m;                     /* r_drag fit  (:102-124) */
} cl_cmb_consts;

typedef struct cl_spec {
  uint32_t abi_version; /* must be CL_ABI_VERSION */
  int32_t ndim;         /* length of one theta row */

  /* ---- expansion history ---- */
  int32_t family;   /* CL_FAMILY_* */
  int32_t de_model; /* CL_DE_* */
  int32_t col_H0;   /* theta column of H0 (or h); -1: use H0_fixed (sn/union3_1.py:11) */
  double H0_fixed;
  double H0_scale;  /* H0 = H0_scale * theta[col_H0]; 100 when the script samples h (bao/desi.py:33) */
  int32_t col_Om;   /* LATE: column of Omega_m */
  int32_t Om_is_physical; /* LATE: column holds omega_m = Omega_m h^2 (bao/desi_omh2.py) */
  int32_t col_obh2; /* FULL: omega_b column */
  int32_t col_och2; /* FULL: omega_c column */
  int32_t col_w0;   /* -1 when the DE model has no w0 */
  int32_t col_wa;
  cl_cmb_consts cmbc; /* needed for FULL, for CL_RD_FIT and for the CMB block */

  /* ---- redshift grid of the cumulative trapezoid (sn/pantheon.py:16-17) ---- */
  const double* z_grid; /* the caller's np.linspace(0, z_max+0.1, 4000) bits */
  int32_t n_grid;       /* 16..4096 */

  /* ---- supernova block ---- */
  int32_t n_sn;            /* 0 = no SN term */
  const double* sn_zcmb;   /* [n_sn] */
  const double* sn_zhel;   /* [n_sn] */
  const double* sn_obs;    /* [n_sn] m_b or mu */
  int32_t sn_cov_form;     /* CL_SN_* */
  const double* sn_mat;    /* [n_sn*n_sn] row-major L (upper part ignored) or C^-1 */
  int32_t col_offset;      /* M / Delta-M column, -1 = none */
  int32_t n_vel;           /* number of velocity templates (0 = no mu_corr) */
  int32_t col_vel[CL_MAX_VEL];
  const double* sn_vel_weight; /* [n_vel*n_sn] e.g. +1 where z_cmb<=z_turn else -1 (sn/pantheon.py:46) */
  double vel_scale;        /* 100 when v is sampled in units of 100 km/s */
  int32_t vel_mode;        /* CL_VEL_* */
  const double* sn_mu_fixed; /* nullable [n_sn]: where finite, the model distance modulus 25 + 5 log10(d_L) is replaced by this
                                value (SH0ES Cepheid calibrators, sn/pantheon_and_sh0es.py:47,63-69; mu_corr still applies);
                                NaN = use the model */
  int32_t n_lin;           /* linear-in-magnitude templates: delta -= sum_k theta[col_lin[k]] * sn_lin_template[k][i]
                              (bao/desi_cmb_pantheon_H0trgb.py:103-106: 100 (5/ln10) / (c z_i)) */
  int32_t col_lin[CL_MAX_VEL];
  const double* sn_lin_template; /* [n_lin*n_sn] */

  /* ---- BAO block ---- */
  int32_t n_bao;              /* 0 = none */
  const double* bao_z;        /* [n_bao] */
  const double* bao_value;    /* [n_bao] */
  const int32_t* bao_qty;     /* [n_bao] CL_BAO_* */
  const double* bao_inv_cov;  /* [n_bao*n_bao] */
  int32_t bao_dh_mode;        /* CL_DH_* */
  int32_t rd_mode;            /* CL_RD_* */
  double rd_fixed;
  int32_t col_rd;

  /* ---- compressed CMB block ---- */
  int32_t cmb_mode;       /* CL_CMB_* */
  double cmb_prior[3];    /* DISTANCE_PRIORS */
  double cmb_weight[9];   /* 3x3 weight W: chi2 = d^T W d.  Sub-selections of the reference (l_A only
                             d^2/cov[1,1], rows [1:], ...) are expressed by zero rows/columns. */
  const double* gl_x;     /* Gauss-Legendre nodes on [-1,1]: the caller's np.polynomial.legendre.leggauss(100)[0]
                             (cmb/data_planck_act_compression.py:150); NULL = computed by the library */
  const double* gl_w;     /* weights */
  int32_t n_gl;           /* <= CL_MAX_GL; 0 with NULL pointers = 100 nodes */

  /* ---- cosmic chronometers: chi2_cc = f^2 d^T W d, d = H_obs - H(z)   (ohd/cc.py:16-38) ---- */
  int32_t n_cc;             /* 0 = none */
  const double* cc_z;       /* [n_cc] */
  const double* cc_H;       /* [n_cc] */
  const double* cc_inv_cov; /* [n_cc*n_cc] */
  int32_t col_fcc;          /* -1: f = 1 */
  double cc_logdet;         /* log det C_cc */
  double cc_norm_sign;      /* log L gets -0.5*(n ln 2pi + logdet) + cc_norm_sign * n ln f;
                               +1: chi2_cc = f^2 d^T C^-1 d (ohd/cc.py:25,33); -1: f inflates the errors, chi2_cc = f^-2 d^T C^-1 d
                               (ohd/cc_pantheon.py:63,92); 0 disables the normalisation (chi2_cc = f^2 ...) */

  /* ---- extra Gaussian chi2 terms ((theta[col]-mean)/sigma)^2, e.g. H0 TRGB
         (bao/desi_cmb_pantheon_H0trgb.py:124) ---- */
  int32_t n_gauss_chi2;
  int32_t gauss_chi2_col[CL_MAX_GAUSS];
  double gauss_chi2_mean[CL_MAX_GAUSS];
  double gauss_chi2_sigma[CL_MAX_GAUSS];

  /* ---- prior (only used by CL_OUT_LOGPROB) ---- */
  int32_t has_bounds;      /* open box lo < theta < hi else -inf (sn/pantheon.py:81-83) */
  double lo[CL_MAX_DIM];
  double hi[CL_MAX_DIM];
  double log_prior_norm;   /* e.g. -sum log(hi-lo) (sn/pantheon.py:77) */
  int32_t n_gauss_prior;   /* -0.5 ((theta[col]-mean)/sigma)^2 in the log-prior (sn/pantheon.py:85) */
  int32_t gauss_prior_col[CL_MAX_GAUSS];
  double gauss_prior_mean[CL_MAX_GAUSS];
  double gauss_prior_sigma[CL_MAX_GAUSS];

  /* ---- guards ---- */
  int32_t guard_cpl;   /* CPL only: w0+wa >= 0 -> log L = guard_value (bao/desi_fs_lya_cmb.py:119-120) */
  double guard_value;  /* -1e8 */
} cl_spec;

typedef struct cl_ctx cl_ctx;

/* Build an engine for one likelihood (replaces the import-time section of a fit script:
 * sn/pantheon.py:10-19 — load data, cho_factor, grid).  Copies every array of the spec to the device;
 * for CL_SN_CHOLESKY with n_sn > CL_SN_SMALL_MAX it also forms W = L^-1 on the host in extended precision.
 * `device` is the CUDA ordinal (one process per GPU). */
int cl_create(const cl_spec* spec, int device, cl_ctx** out);
int cl_destroy(cl_ctx* ctx);
const char* cl_last_error(const cl_ctx* ctx); /* ctx may be NULL for errors of cl_create */

/* Batched evaluation from HOST memory: theta[B][ld] row-major float64 -> out[B].
 * Replaces chi_squared / log_likelihood / log_probability (sn/pantheon.py:57-97) and the batch wrappers
 * log_probs_vectorized (bao/desi.py:100-106), log_likelihood(batch) (bao/desi_union3_cc_theta_star.py:142-147).
 * H2D and D2H copies go through the context's pinned staging buffers; returns after the result is in out. */
int cl_eval(cl_ctx* ctx, const double* theta, int64_t B, int64_t ld, int what, double* out);

/* Same with DEVICE pointers, asynchronous on `stream` (a cudaStream_t cast to void*; NULL = the context's
 * own non-blocking stream — pass cudaStreamLegacy (0x1) to target the legacy default stream).  The caller
 * synchronises.  One stream at a time per context: the workspace is shared between calls. */
int cl_eval_device(cl_ctx* ctx, const double* d_theta, int64_t B, int64_t ld, int what, double* d_out, void* stream);

/* chi2 components, host memory: out[B][4] = (sn, bao, cmb, cc+gaussian)
 * (bao/desi_cmb_union3.py:97-135 chi2_bao / chi2_sn / chi2_cmb) */
int cl_eval_components(cl_ctx* ctx, const double* theta, int64_t B, int64_t ld, double* out);

/* Profile-likelihood mode with the magnitude offset marginalised / profiled analytically (not in
 * the reference — parity unpinned).  Only valid when col_offset >= 0 and n_sn > CL_SN_SMALL_MAX.
 * out[B][3] = (y.y, y.u, u.u) with y = L^-1 d(offset=0), u = L^-1 1, so that
 * chi2_sn(M) = yy - 2 M yu + M^2 uu,  min_M chi2_sn = yy - yu^2/uu. */
int cl_eval_sn_moments(cl_ctx* ctx, const double* theta, int64_t B, int64_t ld, double* out);

/* Distances at arbitrary redshifts (helper exports DM_z / DH_z used by the scripts' main() and by the parity
 * tests: sn/pantheon.py:34-40, bao/desi_cmb_pantheon.py:61-72).  DM[B][nq] from the trapezoid grid + Hermite
 * interpolation, DH[B][nq] = c/H(z) exactly (either may be NULL). */
int cl_distances(cl_ctx* ctx, const double* theta, int64_t B, int64_t ld, const double* zq, int64_t nq, double* DM, double* DH);

/* BAO theory vector out[B][n_bao] (bao_theory: bao/desi_cmb_union3.py:76-94) */
int cl_bao_theory(cl_ctx* ctx, const double* theta, int64_t B, int64_t ld, double* out);

/* CMB derived quantities out[B][8] = (v0, v1, v2 of the compressed vector, z*, r_s(z*), D_M(z*), r_drag, 100 theta*)
 * (cmb/data_planck_act_compression.py:200-212, cmb/cmb.py:48-63 blobs) */
int cl_cmb(cl_ctx* ctx, const double* theta, int64_t B, int64_t ld, double* out);

/* SN residual vector delta[B][n_sn] = obs - offset - mu_corr - mu_theory (sn/pantheon.py:58-60) */
int cl_sn_residuals(cl_ctx* ctx, const double* theta, int64_t B, int64_t ld, double* out);

/* Timing of the most recent cl_eval / cl_eval_device on this context (CUDA events on the launching stream):
 * ms[0] stage 1+2 (Friedmann distances + residuals), ms[1] stage 3 (chi-squared GEMM), ms[2] finalize,
 * ms[3] total device time incl. copies for cl_eval.  Blocks until the events have completed. */
int cl_last_timing(cl_ctx* ctx, double ms[4]);
/* Split of ms[1] for the most recent n evaluations, oldest first: ms[i][0] = forming the int8 digit planes of the residual
 * rows (0 with the DMMA engine), ms[i][1] = the contraction kernel itself.  Returns the number of entries written. */
int cl_stage3_split(cl_ctx* ctx, int n, double* ms);
/* Same for the most recent n evaluations (the library keeps the last 64), oldest first: ms[i][4].
 * Returns the number of entries written (<= n) or a negative error. */
int cl_timing_history(cl_ctx* ctx, int n, double* ms);
/* Page-locked host memory for theta / result buffers.  cl_eval(), cl_eval_components() and cl_eval_sn_moments() recognise
 * page-locked pointers (from here, cudaHostAlloc, cudaHostRegister or torch pin_memory) and move them by DMA directly;
 * pageable buffers go through the library's own pinned staging area (one extra host copy each way).  ctx may be NULL. */
int cl_host_alloc(cl_ctx* ctx, size_t bytes, void** ptr);
int cl_host_free(cl_ctx* ctx, void* ptr);

/* Number of kernels this library launched on the context since creation. */
int64_t cl_launch_count(const cl_ctx* ctx);

/* Options (integers): "chi2_engine" (CL_CHI2_ENGINE_*), "chi2_slices" (5..7), "max_rows_per_pass", "gemm_ctas",
 * "stage12_ctas", "stage12_lean" (0: always the full stage-1+2 kernel), "fuse_planes" (default 1: the lean stage-2 kernel writes the digit planes itself instead of the FP64 residual rows; 0: separate slicing kernel, same bits), "chi2_guard" (0 switches the accuracy guard of the tcgen05 engine off), "chi2_guard_mode" (see cl_guard_info), "gemm_group_rb", "chi2_slice_tpb", and for the DMMA engine "gemm_dynamic", "gemm_diag_skip".  "dbg" is for
 * profiling builds of the library (nvcc -DOZ_PROF=1; the production build compiles the counters out): bit 2 prints the cycle
 * counters of the tcgen05 contraction to stderr, bit 3 writes the event trace of CTA 0 to $COSMOLIKE_TRACE (default
 * oz_trace.txt); bits 4-7 are timing experiments that INVALIDATE the results (loads / folds switched off).
 * Returns CL_E_INVALID for unknown names. */
int cl_set_option(cl_ctx* ctx, const char* name, int64_t value);

/* CUDA graphs of small evaluations.  A call of cl_eval / cl_eval_components / cl_eval_sn_moments (pageable host buffers)
 * or cl_eval_device (caller's stream) with at most "cuda_graph_max_rows" rows (option, default 4096; "cuda_graphs" = 0
 * switches the mechanism off) is captured into a CUDA graph the SECOND time its shape is seen - entry point, rows, output
 * selector and, for cl_eval_device, the two device pointers, ld and the stream - and replayed from then on: one graph
 * launch (upload, stage 1+2, contraction, finalize, the accuracy-guard fallback, download) instead of ~10 stream operations,
 * which is what a stock emcee / nautilus call of 75-100 rows costs most (sn/pantheon.py:119-125 hands over one row at a
 * time).  Same kernels, same arguments, same bits.  Replays record no per-stage events: cl_last_timing / cl_timing_history /
 * cl_stage3_split keep describing the most recent evaluation that was launched the ordinary way.  Every option change and
 * every regrowth of a workspace drops the captured graphs.  out[0] = graphs held, out[1] = replays since cl_create. */
int cl_graph_info(cl_ctx* ctx, int64_t out[2]);

/* Floating-point options: "chi2_guard_abs" (default 5e-7), "chi2_guard_rel" (default 1e-12) - see cl_guard_info. */
int cl_set_option_f64(cl_ctx* ctx, const char* name, double value);

/* Accuracy guard of the tcgen05 (int8 digit plane) chi-squared engine.  The engine replaces the reference's FP64 forward
 * substitution (solve_triangular.py:5-14) by an exact integer contraction of S digit planes per operand row; its only
 * errors are the fixed-point rounding of the operands (one power-of-two scale per row) and the dropped products below the
 * last kept digit: every term W_nk r_k is off by at most 2^eR_b 2^eW_n eps_S, eps_S = 2^(2-8S) (1 + (S-1) 256/255).  Two
 * a-priori bounds per row b follow, selected by option "chi2_guard_mode":
 *   0 (default) probabilistic - the term errors are bounded, mean zero and independent (roundings of balanced digits), so by
 *     Hoeffding's inequality  |d chi2_b| <= 2 lambda sqrt(chi2_b) 2^eR_b eps_S Omega_pr + rho_b^2,
 *     Omega_pr = max_n sqrt(nnz_n) 2^eW_n,  except with probability < 2 exp(-lambda^2/2) = 2.5e-14 per row (lambda = 8);
 *   1 worst case - every error at its maximum and of one sign:  |d chi2_b| <= 2 sqrt(chi2_b) rho_b + rho_b^2,
 *     rho_b = 2^eR_b eps_S Omega_wc,  Omega_wc = sqrt(sum_n (nnz_n 2^eW_n)^2)   (both Omegas static, from W = L^-1).
 * With option "chi2_guard" = 1 (default) every row whose bound exceeds max(chi2_guard_abs, chi2_guard_rel * chi2_b) is
 * recomputed on the FP64 tensor pipe (DMMA engine) inside the same call; the values the caller sees therefore always meet
 * the tolerance (in the sense of the selected bound) or are FP64 results.  out[0] = rows recomputed since cl_create,
 * out[1] = rows recomputed by the most recent pass, out[2] = Omega and out[3] = the coefficient of the linear term
 * (lambda eps_S Omega_pr or eps_S Omega_wc) for the current mode and plane count.  Synchronises the device. */
int cl_guard_info(cl_ctx* ctx, double out[4]);

/* ---- multi-GPU: one process per GPU, one context per process, NCCL bound at run time (dlopen of libnccl.so.2, or the path
 * in $COSMOLIKE_NCCL_LIB) -------------------------------------------------------------------------------------------------
 * The likelihood of every parameter vector is independent - the reference farms rows out with multiprocessing.Pool /
 * numba prange (sn/pantheon.py:119-125, bao/desi.py:100-106) - so a batch is row-sharded over the ranks, every rank holds the
 * static operands, and the only communication is the all-gather of the per-row results over NVLink.
 * cl_comm_unique_id: rank 0 obtains the 128-byte NCCL id and ships it to the other ranks by any means (MPI, a file, a socket,
 * torch.distributed); cl_comm_init: collective over all ranks, binds the communicator to the context's device. */
int cl_comm_unique_id(void* uid /* CL_NCCL_UID_BYTES */);
int cl_comm_init(cl_ctx* ctx, int rank, int nranks, const void* uid);
int cl_comm_destroy(cl_ctx* ctx);            /* also done by cl_destroy */
int cl_comm_info(const cl_ctx* ctx, int* rank, int* nranks);   /* nranks = 0: no communicator */

/* Sharded evaluation (collective): this rank's B_local rows theta_local[B_local][ld] (HOST memory; the same B_local on every
 * rank - pad the last shard) -> out_all[nranks * B_local] in rank order.  root < 0: every rank receives the gathered vector;
 * root >= 0: only that rank downloads it (out_all may be NULL elsewhere) - the master/worker shape of the reference's
 * Pool.map.  Page-locked buffers are moved by DMA directly, as in cl_eval. */
int cl_eval_allgather(cl_ctx* ctx, const double* theta_local, int64_t B_local, int64_t ld, int what, double* out_all, int root);
/* Same with DEVICE pointers, asynchronous on `stream`: d_out_all[nranks * B_local]; this rank's results are written in place
 * at d_out_all + rank * B_local and gathered with one ncclAllGather.  The caller synchronises. */
int cl_eval_allgather_device(cl_ctx* ctx, const double* d_theta_local, int64_t B_local, int64_t ld, int what, double* d_out_all, void* stream);

/* ---- profile-likelihood grids generated on the device (BASELINE.json config 4; not in the reference) ---------------------
 * A Cartesian grid over n_axes theta columns, axis a = np.linspace(lo[a], hi[a], n[a]) (LAST axis fastest: the order of
 * np.meshgrid(..., indexing="ij").ravel()); the other columns are `fixed`.  cl_eval_grid evaluates the points
 * first .. first + count - 1 of the flattened grid: theta never crosses PCIe, and the values only do when `out` is given. */
typedef struct cl_grid {
  int32_t n_axes;
  int32_t col[CL_MAX_DIM];   /* theta column of axis a */
  int64_t n[CL_MAX_DIM];     /* points of axis a */
  double lo[CL_MAX_DIM], hi[CL_MAX_DIM];
  double fixed[CL_MAX_DIM];  /* value of every theta column that is not an axis (indexed by column) */
} cl_grid;
/* what: CL_OUT_CHI2 / CL_OUT_LOGLIKE / CL_OUT_LOGPROB, or for a large Cholesky SN block the chi2 with the magnitude offset
 * handled in closed form from the two-dot epilogue (cl_eval_sn_moments): */
enum { CL_GRID_PROFILE = 16 /* min_M chi2 = yy - yu^2/uu */, CL_GRID_MARGINAL = 17 /* -2 ln int dM exp(-chi2/2) */ };
typedef struct cl_grid_stats {
  double best;      /* smallest chi2 (largest log L / log P for the CL_OUT_LOG* selectors) over the points evaluated */
  int64_t index;    /* its flattened grid index (the smallest one among ties), -1 if no finite value */
  double log_sum;   /* ln sum exp(-chi2/2)  (ln sum exp(log L) for the CL_OUT_LOG* selectors) */
  int64_t count;    /* points evaluated */
  int32_t larger_is_better;   /* set by cl_eval_grid: 1 for the CL_OUT_LOG* selectors, 0 for chi2-like values */
  int32_t reserved;
} cl_grid_stats;
int cl_eval_grid(cl_ctx* ctx, const cl_grid* grid, int64_t first, int64_t count, int what, double* out /* nullable [count] */, cl_grid_stats* stats);
/* Collective: combines the stats of every rank's slice of the grid (one all-gather of 32 bytes per rank, folded in rank
 * order on every rank: the same bits everywhere). */
int cl_grid_allreduce(cl_ctx* ctx, cl_grid_stats* stats);

/* ---- nested-sampling proposals generated, evaluated and filtered on the device (the reference drives
 * nautilus with ~100 points per likelihood call: bao/desi_cmb_pantheon.py:153-170) -----------------------------------------
 * n points uniform in the ellipsoid {mu + L z : |z| < 1} of the unit cube (counter-based Philox4x32-10: row i of the call
 * sees the numbers of (seed, offset + i) whatever the batch size), mapped through the prior transform (uniform on [lo, hi]
 * per column, or Gaussian mean + sigma ndtri(u) where gauss[col] != 0), evaluated with selector `what`, and the rows that lie
 * inside the unit cube with value > thresh are returned IN DRAW ORDER, at most max_keep of them:
 * u_out[max_keep][ndim], theta_out[max_keep][ndim], val_out[max_keep] (host memory; page-locked buffers move by DMA).
 * counts[0] = rows inside the unit cube, counts[1] = rows accepted, counts[2] = rows returned = min(counts[1], max_keep). */
typedef struct cl_proposal {
  int32_t ndim;                          /* = the spec's ndim */
  int32_t gauss[CL_MAX_DIM];
  double mu[CL_MAX_DIM];
  double L[CL_MAX_DIM * CL_MAX_DIM];     /* row-major [ndim][ndim], lower triangular */
  double lo[CL_MAX_DIM], hi[CL_MAX_DIM], mean[CL_MAX_DIM], sigma[CL_MAX_DIM];
} cl_proposal;
int cl_propose_eval(cl_ctx* ctx, const cl_proposal* prop, int64_t n, uint64_t seed, uint64_t offset, int what, double thresh, int64_t max_keep,
                    double* u_out, double* theta_out, double* val_out, int64_t counts[3]);

/* Library / device description string, e.g. "cosmolike_b200 abi 3, sm_100a, NVIDIA B200 (148 SMs)". */
const char* cl_describe(const cl_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* COSMOLIKE_H */
