/*
 * cosmo_oracle.c — CPU restatement of the reference's likelihood arithmetic.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for libcosmolike_b200.so and the "port" CPU baseline of bench.py.  It is
 * never linked into, imported by or executed from the product path (only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it).
 *
 * It follows the reference's algorithms operation by operation — sequential cumulative trapezoid, bisection
 * searchsorted, Hermite/PCHIP formulas, 100-point Gauss-Legendre sums in node order, forward substitution —
 * in plain scalar C (gcc -O2, no -ffast-math), so that it reproduces the numba results to rounding.
 * Pinned against the golden vectors in tests/golden/ that were produced by running the unmodified reference
 * (tests/golden/make_golden.py); see tests/test_oracle_golden.py.
 *
 * Reference files restated here (paths relative to the reference root):
 *   interpolator.py:5-119, solve_triangular.py:5-14, nu_evolution.py:5-28 (constants arrive via the spec),
 *   cmb/data_planck_act_compression.py:53-212, sn/pantheon.py:22-97, sn/union3_1.py:17-57,
 *   bao/desi.py:24-106, bao/desi_cmb_union3.py:31-140, bao/desi_des5y_bbn_theta_star.py:25-151,
 *   bao/desi_cmb_pantheon.py:25-135, bao/desi_fs_lya_cmb.py:18-121, ohd/cc.py:16-38.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#include "../include/cosmolike.h"

#define C_KMS 299792.458 /* scipy.constants.c / 1000 (sn/pantheon.py:12) */

typedef struct {
  double H0, h;
  double Om;              /* LATE */
  double Or, Obc, Onu, Ode; /* FULL */
  double obh2, och2;
  double w0, wa;
} cosmo_t;

/* ---- parameter unpacking (role -> column map) ---- */
static void unpack(const cl_spec* s, const double* th, cosmo_t* c) {
  memset(c, 0, sizeof *c);
  c->H0 = s->col_H0 >= 0 ? s->H0_scale * th[s->col_H0] : s->H0_fixed;
  c->h = c->H0 / 100;
  c->w0 = s->col_w0 >= 0 ? th[s->col_w0] : -1.0;
  c->wa = s->col_wa >= 0 ? th[s->col_wa] : 0.0;
  if (s->family == CL_FAMILY_LATE) {
    c->Om = th[s->col_Om];
    if (s->Om_is_physical) c->Om = c->Om / (c->h * c->h); /* bao/desi_omh2.py: Om = Omh2 / h**2 */
  } else {
    /* bao/desi_cmb_union3.py:38-42 */
    double h2 = c->h * c->h;
    c->obh2 = th[s->col_obh2];
    c->och2 = th[s->col_och2];
    c->Onu = s->cmbc.Omnu_h2 / h2;
    c->Or = s->cmbc.Or_h2 / h2;
    c->Obc = (c->obh2 + c->och2) / h2;
    c->Ode = 1.0 - c->Obc - c->Or - c->Onu;
  }
}

/* cmb/data_planck_act_compression.py:53-66 */
static double Omnu_z(const cl_cmb_consts* k, double z) {
  double zp1 = 1.0 + z;
  double r = k->nu_m0 / zp1;
  double mz_sq = r * r;
  double f0 = sqrt(k->nu_q2[0] + mz_sq), f1 = sqrt(k->nu_q2[1] + mz_sq), f2 = sqrt(k->nu_q2[2] + mz_sq);
  double f3 = sqrt(k->nu_q2[3] + mz_sq), f4 = sqrt(k->nu_q2[4] + mz_sq);
  double ws = f0 * k->nu_w[0] + f1 * k->nu_w[1] + f2 * k->nu_w[2] + f3 * k->nu_w[3] + f4 * k->nu_w[4];
  double zp1_2 = zp1 * zp1;
  return zp1_2 * zp1_2 * ws / k->nu_rho0;
}

/* dark-energy density factor (bao/desi_cmb_pantheon.py:25-31 menu) */
static double fde(const cl_spec* s, const cosmo_t* c, double z) {
  double zp1 = 1.0 + z;
  switch (s->de_model) {
    case CL_DE_WCDM: return pow(zp1, 3 * (1.0 + c->w0));
    case CL_DE_CPL: return pow(zp1, 3 * (1 + c->w0 + c->wa)) * exp(-3 * c->wa * z / zp1);
    case CL_DE_THAWING: {
      double cubed = zp1 * zp1 * zp1;
      double q = 2 * cubed / ((1.0 + c->w0) + (1.0 - c->w0) * cubed);
      return q * q;
    }
    default: return 1.0;
  }
}

/* H(z): sn/pantheon.py:28-31 (late), bao/desi_cmb_union3.py:37-57 (full) */
static double H_of_z(const cl_spec* s, const cosmo_t* c, double z) {
  double zp1 = 1.0 + z;
  double cubed = zp1 * zp1 * zp1;
  if (s->family == CL_FAMILY_LATE) {
    double de = (s->de_model == CL_DE_LCDM) ? (1.0 - c->Om) : (1.0 - c->Om) * fde(s, c, z);
    return c->H0 * sqrt(c->Om * cubed + de);
  }
  double radiation = c->Or * (cubed * zp1);
  double matter = c->Obc * cubed;
  double neutrino = c->Onu * Omnu_z(&s->cmbc, z);
  double de = (s->de_model == CL_DE_LCDM) ? c->Ode : c->Ode * fde(s, c, z);
  return c->H0 * sqrt(radiation + matter + de + neutrino);
}

/* DM_grid: bao/desi_cmb_des5y.py:60-66 == sn/pantheon.py:36-39 */
static void build_grid(const cl_spec* s, const cosmo_t* c, double* dh_grid, double* cum_dm) {
  int G = s->n_grid;
  for (int i = 0; i < G; i++) dh_grid[i] = C_KMS / H_of_z(s, c, s->z_grid[i]);
  cum_dm[0] = 0.0;
  double acc = 0.0;
  for (int i = 0; i < G - 1; i++) {
    double dh = (dh_grid[i] + dh_grid[i + 1]) / 2;
    double dz = s->z_grid[i + 1] - s->z_grid[i];
    acc += dh * dz; /* np.cumsum: sequential */
    cum_dm[i + 1] = acc;
  }
}

/* np.searchsorted(x, xi) (side='left'): first index with x[idx] >= xi */
static int searchsorted_left(const double* x, int n, double xi) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (x[mid] < xi) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* interpolator.py:71-108 */
static double pchip_eval(double xi, const double* x, const double* y, const double* d, int n, int exact) {
  if (!exact) {
    if (xi <= x[0]) return y[0];
    if (xi >= x[n - 1]) return y[n - 1];
  } else {
    if (xi <= x[0]) return y[0] + d[0] * (xi - x[0]);
    if (xi >= x[n - 1]) return y[n - 1] + d[n - 1] * (xi - x[n - 1]);
  }
  int i = searchsorted_left(x, n, xi) - 1;
  double h_i = x[i + 1] - x[i];
  double t = (xi - x[i]) / h_i;
  double t2 = t * t, t3 = t2 * t;
  double h00 = 2 * t3 - 3 * t2 + 1;
  double h10 = t3 - 2 * t2 + t;
  double h01 = -2 * t3 + 3 * t2;
  double h11 = t3 - t2;
  return h00 * y[i] + h10 * h_i * d[i] + h01 * y[i + 1] + h11 * h_i * d[i + 1];
}

static double sgn(double v) { return (v > 0) - (v < 0); }

/* interpolator.py:5-68 */
static void pchip_slopes(const double* x, const double* y, int n, double* d, double* h, double* delta) {
  for (int i = 0; i < n; i++) d[i] = 0.0;
  if (n < 2) return;
  for (int i = 0; i < n - 1; i++) {
    h[i] = x[i + 1] - x[i];
    delta[i] = (y[i + 1] - y[i]) / h[i];
  }
  if (n == 2) { d[0] = delta[0]; d[1] = delta[0]; return; }
  for (int i = 1; i < n - 1; i++) {
    double dm1 = delta[i - 1], di = delta[i], hm1 = h[i - 1], hi = h[i];
    if (dm1 != 0.0 && di != 0.0 && dm1 * di > 0.0) {
      double w1 = 2.0 * hi + hm1, w2 = hi + 2.0 * hm1;
      d[i] = (w1 + w2) / (w1 / dm1 + w2 / di);
    } else d[i] = 0.0;
  }
  double d0 = ((2 * h[0] + h[1]) * delta[0] - h[0] * delta[1]) / (h[0] + h[1]);
  if (delta[0] == 0.0 || sgn(d0) != sgn(delta[0])) d[0] = 0.0;
  else if (sgn(delta[0]) != sgn(delta[1]) && fabs(d0) > fabs(3 * delta[0])) d[0] = 3 * delta[0];
  else d[0] = d0;
  double dn = ((2 * h[n - 2] + h[n - 3]) * delta[n - 2] - h[n - 2] * delta[n - 3]) / (h[n - 2] + h[n - 3]);
  if (delta[n - 2] == 0.0 || sgn(dn) != sgn(delta[n - 2])) d[n - 1] = 0.0;
  else if (sgn(delta[n - 2]) != sgn(delta[n - 3]) && fabs(dn) > fabs(3 * delta[n - 2])) d[n - 1] = 3 * delta[n - 2];
  else d[n - 1] = dn;
}

/* standalone exports of the interpolator for its own golden test */
void oracle_interp_hermite(const double* xq, int nq, const double* x, const double* y, const double* yp, int n, double* out) {
  for (int k = 0; k < nq; k++) out[k] = pchip_eval(xq[k], x, y, yp, n, 1);
}
void oracle_pchip_slopes(const double* x, const double* y, int n, double* d) {
  double* h = malloc(sizeof(double) * n);
  double* delta = malloc(sizeof(double) * n);
  pchip_slopes(x, y, n, d, h, delta);
  free(h); free(delta);
}
void oracle_interp_pchip(const double* xq, int nq, const double* x, const double* y, int n, double* out) {
  double* d = malloc(sizeof(double) * n);
  oracle_pchip_slopes(x, y, n, d);
  for (int k = 0; k < nq; k++) out[k] = pchip_eval(xq[k], x, y, d, n, 0);
  free(d);
}

/* solve_triangular.py:5-14.  np.dot is BLAS ddot in numba; four partial sums stand in for its SIMD lanes. */
static double dot4(const double* a, const double* b, int n) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int i = 0;
  for (; i + 4 <= n; i += 4) { s0 += a[i] * b[i]; s1 += a[i + 1] * b[i + 1]; s2 += a[i + 2] * b[i + 2]; s3 += a[i + 3] * b[i + 3]; }
  for (; i < n; i++) s0 += a[i] * b[i];
  return (s0 + s1) + (s2 + s3);
}
double oracle_solve_triangular(const double* L, const double* b, int n, double* y) {
  for (int i = 0; i < n; i++) y[i] = (b[i] - dot4(L + (size_t)i * n, y, i)) / L[(size_t)i * n + i];
  return dot4(y, y, n);
}
/* The same forward substitution (solve_triangular.py:10-14) carried in 80-bit extended precision: the arithmetic yardstick of
 * the conditioning tests (tests/test_gpu_conditioning.py).  It is NOT the reference's arithmetic (that is the FP64 version
 * above); it measures how far the FP64 result of anybody - the reference included - is from the exact value for the given
 * (L, delta).  out[b] = |L^-1 R[b]|^2. */
int oracle_solve_triangular_ld(const double* L, const double* R, int64_t B, int n, double* out) {
  long double* y = malloc(sizeof(long double) * (size_t)n);
  if (!y) return -1;
  for (int64_t b = 0; b < B; b++) {
    const double* r = R + (size_t)b * n;
    long double chi2 = 0.0L;
    for (int i = 0; i < n; i++) {
      const double* Li = L + (size_t)i * n;
      long double acc = r[i];
      for (int k = 0; k < i; k++) acc -= (long double)Li[k] * y[k];
      y[i] = acc / (long double)Li[i];
      chi2 += y[i] * y[i];
    }
    out[b] = (double)chi2;
  }
  free(y);
  return 0;
}
/* delta @ M @ delta: (delta @ M) is a vector-matrix product, then a dot (sn/union3_1.py:57) */
static double quad_form(const double* M, const double* d, int n) {
  double acc = 0.0;
  for (int j = 0; j < n; j++) {
    double t = 0.0;
    for (int i = 0; i < n; i++) t += d[i] * M[(size_t)i * n + j];
    acc += t * d[j];
  }
  return acc;
}

/* ---- CMB fits: cmb/data_planck_act_compression.py:86-124 ---- */
static double z_star_fit(const cl_cmb_consts* k, double wb, double wm) {
  wb = pow(wb, k->zstar_b);
  wm = pow(wm, k->zstar_m);
  return pow(wm, -0.7316314841257655) +
         k->zstar_s1 * 391.6723594873167 * pow(wb, 0.9368102670600895) * pow(wm, -0.35300106475765136) +
         k->zstar_s2 * 937.4224935298015 * pow(wm, 0.0192950634264157) * pow(wb, -0.04285000485853785);
}
static double r_drag_fit(const cl_cmb_consts* k, double wb, double wm) {
  wb = pow(wb, k->rdrag_b);
  wm = pow(wm, k->rdrag_m);
  const double a1 = 0.00257366, a2 = 0.05032, a3 = 0.013, a4 = 0.7720642, a5 = 0.24346362, a6 = 0.00641072,
               a7 = 0.5350899, a8 = 32.7525, a9 = 0.315473;
  double den = (a1 * pow(wb, a2)) + (a3 * pow(wb, a4) * pow(wm, a5)) + (a6 * pow(wm, a7));
  return 1.0 / den - a8 / pow(wm, a9);
}

/* Gauss-Legendre nodes when the caller passes none: Newton on P_n (agrees with numpy.leggauss to ~1e-16) */
static void gauss_legendre(int n, double* x, double* w) {
  for (int i = 0; i < n; i++) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 1.0;
    for (int it = 0; it < 100; it++) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; j++) { double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1); }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    x[n - 1 - i] = z;
    w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
  }
}

typedef struct { double v[3], zstar, rs, dm, rdrag, theta100; } cmb_out_t;

/* cmb_distances: cmb/data_planck_act_compression.py:160-212 (and data_early_lcdm_compression.py:200-207) */
static void cmb_eval(const cl_spec* s, const cosmo_t* c, const double* glx, const double* glw, int ngl, cmb_out_t* o) {
  const cl_cmb_consts* k = &s->cmbc;
  double Om_h2 = c->och2 + c->obh2 + k->Omnu_h2;
  double zstar = z_star_fit(k, c->obh2, Om_h2);
  /* rs_z */
  double a_lim = 1.0 / (1.0 + zstar), hw = a_lim / 2.0, mid = a_lim / 2.0, integ = 0.0;
  for (int i = 0; i < ngl; i++) {
    double a = hw * glx[i] + mid;
    double z = (1.0 / a) - 1.0;
    double Rb = (3.0 / 4.0) * (c->obh2 / k->Ogamma_h2) * a;
    integ += glw[i] * (C_KMS / (a * a * H_of_z(s, c, z) * sqrt(3.0 * (1.0 + Rb))));
  }
  double rs = hw * integ;
  /* DM_z */
  hw = zstar / 2.0; mid = zstar / 2.0; integ = 0.0;
  for (int i = 0; i < ngl; i++) integ += glw[i] * (C_KMS / H_of_z(s, c, hw * glx[i] + mid));
  double dm = hw * integ;
  o->zstar = zstar; o->rs = rs; o->dm = dm;
  o->theta100 = 100 * (rs / dm);
  o->rdrag = r_drag_fit(k, c->obh2, Om_h2);
  if (s->cmb_mode == CL_CMB_THETA_WB_WM) {
    o->v[0] = rs / dm; o->v[1] = c->obh2; o->v[2] = Om_h2;
  } else {
    o->v[0] = 100 * sqrt(Om_h2) * dm / C_KMS; o->v[1] = M_PI * dm / rs; o->v[2] = c->obh2;
  }
}

/* ---- per-thread workspace ---- */
typedef struct {
  double *dh_grid, *cum_dm, *slopes, *sh, *sdelta, *delta_sn, *y, *vec;
  double glx[CL_MAX_GL], glw[CL_MAX_GL];
  int ngl;
} work_t;

static int work_init(const cl_spec* s, work_t* w) {
  int G = s->n_grid, n = s->n_sn > 0 ? s->n_sn : 1;
  w->dh_grid = malloc(sizeof(double) * G); w->cum_dm = malloc(sizeof(double) * G);
  w->slopes = malloc(sizeof(double) * G); w->sh = malloc(sizeof(double) * G); w->sdelta = malloc(sizeof(double) * G);
  w->delta_sn = malloc(sizeof(double) * n); w->y = malloc(sizeof(double) * n);
  w->vec = malloc(sizeof(double) * (CL_MAX_BAO + CL_MAX_CC));
  if (s->gl_x && s->gl_w && s->n_gl > 0) {
    w->ngl = s->n_gl; memcpy(w->glx, s->gl_x, sizeof(double) * s->n_gl); memcpy(w->glw, s->gl_w, sizeof(double) * s->n_gl);
  } else { w->ngl = 100; gauss_legendre(100, w->glx, w->glw); }
  return w->dh_grid && w->cum_dm && w->slopes && w->sh && w->sdelta && w->delta_sn && w->y && w->vec ? 0 : -1;
}
static void work_free(work_t* w) {
  free(w->dh_grid); free(w->cum_dm); free(w->slopes); free(w->sh); free(w->sdelta); free(w->delta_sn); free(w->y); free(w->vec);
}

static double rd_value(const cl_spec* s, const cosmo_t* c, const double* th) {
  if (s->rd_mode == CL_RD_FIXED) return s->rd_fixed;
  if (s->rd_mode == CL_RD_PARAM) return th[s->col_rd];
  double obh2 = s->family == CL_FAMILY_FULL ? c->obh2 : th[s->col_obh2];
  double wm = s->family == CL_FAMILY_FULL ? c->obh2 + c->och2 + s->cmbc.Omnu_h2 : c->Om * c->h * c->h;
  return r_drag_fit(&s->cmbc, obh2, wm);
}

/* bao_theory: bao/desi_cmb_union3.py:76-94 (pchip), bao/desi_cmb_pantheon.py:85-99 (exact) */
static void bao_theory(const cl_spec* s, const cosmo_t* c, const double* th, work_t* w, int have_slopes, double* out) {
  int G = s->n_grid;
  double rd = rd_value(s, c, th);
  if (s->bao_dh_mode == CL_DH_PCHIP && !have_slopes) pchip_slopes(s->z_grid, w->dh_grid, G, w->slopes, w->sh, w->sdelta);
  for (int k = 0; k < s->n_bao; k++) {
    double z = s->bao_z[k];
    double DM = pchip_eval(z, s->z_grid, w->cum_dm, w->dh_grid, G, 1);
    double DH = s->bao_dh_mode == CL_DH_PCHIP ? pchip_eval(z, s->z_grid, w->dh_grid, w->slopes, G, 0) : C_KMS / H_of_z(s, c, z);
    switch (s->bao_qty[k]) {
      case CL_BAO_DV_OVER_RS: out[k] = pow(z * DH * (DM * DM), 1.0 / 3) / rd; break;
      case CL_BAO_DM_OVER_RS: out[k] = DM / rd; break;
      case CL_BAO_DH_OVER_RS: out[k] = DH / rd; break;
      default: out[k] = DM / DH; break;
    }
  }
}

/* SN residual: sn/pantheon.py:43-60, bao/desi_cmb_union3.py:103-123 */
static void sn_residuals(const cl_spec* s, const double* th, work_t* w, double* delta) {
  int G = s->n_grid, n = s->n_sn;
  double offset = s->col_offset >= 0 ? th[s->col_offset] : 0.0;
  for (int i = 0; i < n; i++) {
    double zc = s->sn_zcmb[i];
    double DM = pchip_eval(zc, s->z_grid, w->cum_dm, w->dh_grid, G, 1);
    double mu_corr = 0.0;
    if (s->n_vel > 0) {
      double v_km_s = 0.0;
      for (int k = 0; k < s->n_vel; k++) v_km_s += s->vel_scale * th[s->col_vel[k]] * s->sn_vel_weight[(size_t)k * n + i];
      double z_pec = v_km_s / C_KMS, z_cosmo;
      if (s->vel_mode == CL_VEL_DIVIDE) z_cosmo = -1.0 + (1.0 + zc) / (1.0 + z_pec);
      else { z_cosmo = (1.0 + zc) * (1.0 + z_pec) - 1.0; if (z_cosmo < 1e-8) z_cosmo = 1e-8; }
      double DMc = pchip_eval(z_cosmo, s->z_grid, w->cum_dm, w->dh_grid, G, 1);
      mu_corr = 5.0 * log10(DMc / DM);
    }
    /* sn/pantheon_and_sh0es.py:63-69: mu_pred = where(ceph_mask, ceph_dists, mu_theory(DM_cmb)) */
    double mu = (s->sn_mu_fixed && isfinite(s->sn_mu_fixed[i])) ? s->sn_mu_fixed[i]
                                                                : 25.0 + 5 * log10((1.0 + s->sn_zhel[i]) * DM);
    double lin = 0.0; /* bao/desi_cmb_pantheon_H0trgb.py:103-106 */
    for (int k = 0; k < s->n_lin; k++) lin += th[s->col_lin[k]] * s->sn_lin_template[(size_t)k * n + i];
    delta[i] = s->sn_obs[i] - (offset + lin) - mu_corr - mu;
  }
}

typedef struct { double sn, bao, cmb, extra, cc_norm; int guard; } comps_t;

static void eval_one(const cl_spec* s, const double* th, work_t* w, int grid_builds, comps_t* o) {
  cosmo_t c;
  unpack(s, th, &c);
  memset(o, 0, sizeof *o);
  if (s->guard_cpl && s->de_model == CL_DE_CPL && c.w0 + c.wa >= 0.0) o->guard = 1;
  int need_grid = s->n_sn > 0 || s->n_bao > 0;
  if (need_grid) for (int r = 0; r < (grid_builds > 0 ? grid_builds : 1); r++) build_grid(s, &c, w->dh_grid, w->cum_dm);
  if (s->n_sn > 0) {
    sn_residuals(s, th, w, w->delta_sn);
    if (s->sn_cov_form == CL_SN_CHOLESKY) o->sn = oracle_solve_triangular(s->sn_mat, w->delta_sn, s->n_sn, w->y);
    else o->sn = quad_form(s->sn_mat, w->delta_sn, s->n_sn);
  }
  if (s->n_bao > 0) {
    bao_theory(s, &c, th, w, 0, w->vec);
    for (int k = 0; k < s->n_bao; k++) w->vec[k] = s->bao_value[k] - w->vec[k];
    o->bao = quad_form(s->bao_inv_cov, w->vec, s->n_bao);
  }
  if (s->cmb_mode != CL_CMB_NONE) {
    cmb_out_t co;
    cmb_eval(s, &c, w->glx, w->glw, w->ngl, &co);
    double d[3] = {s->cmb_prior[0] - co.v[0], s->cmb_prior[1] - co.v[1], s->cmb_prior[2] - co.v[2]};
    o->cmb = quad_form(s->cmb_weight, d, 3);
  }
  if (s->n_cc > 0) {
    /* ohd/cc.py:22-34 */
    double f = s->col_fcc >= 0 ? th[s->col_fcc] : 1.0;
    double* d = w->vec + CL_MAX_BAO;
    for (int k = 0; k < s->n_cc; k++) d[k] = s->cc_H[k] - H_of_z(s, &c, s->cc_z[k]);
    /* cc_norm_sign < 0: f inflates the errors, chi2 * f ** -2 (ohd/cc_pantheon.py:63); otherwise f^2 * chi2 (ohd/cc.py:25) */
    o->extra += (s->cc_norm_sign < 0.0 ? 1.0 / (f * f) : f * f) * quad_form(s->cc_inv_cov, d, s->n_cc);
    if (s->cc_norm_sign != 0.0)
      o->cc_norm = s->n_cc * log(2 * M_PI) + s->cc_logdet - s->cc_norm_sign * 2 * s->n_cc * log(f);
  }
  for (int g = 0; g < s->n_gauss_chi2; g++) {
    double r = (th[s->gauss_chi2_col[g]] - s->gauss_chi2_mean[g]) / s->gauss_chi2_sigma[g];
    o->extra += r * r;
  }
}

/* log_prior: sn/pantheon.py:80-85 */
static double log_prior(const cl_spec* s, const double* th) {
  if (s->has_bounds)
    for (int j = 0; j < s->ndim; j++)
      if (!(s->lo[j] < th[j] && th[j] < s->hi[j])) return -INFINITY;
  double lp = s->log_prior_norm;
  for (int g = 0; g < s->n_gauss_prior; g++) {
    double r = th[s->gauss_prior_col[g]] - s->gauss_prior_mean[g];
    lp += -0.5 * (r * r) / (s->gauss_prior_sigma[g] * s->gauss_prior_sigma[g]);
  }
  return lp;
}

static int n_threads_for(int nthreads) {
  if (nthreads > 0) return nthreads;
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

int oracle_max_threads(void) { return n_threads_for(0); }

typedef struct {
  const cl_spec* s; const double* theta; int64_t B, ld; int what; double* out; double* comps;
  int grid_builds, tid, nt, err;
} job_t;

static void* eval_worker(void* arg) {
  job_t* j = (job_t*)arg;
  const cl_spec* s = j->s;
  work_t w;
  if (work_init(s, &w) != 0) { j->err = -1; work_free(&w); return NULL; }
  /* rows are dealt round-robin in blocks of 4, like a prange over batch rows (bao/desi.py:104-105) */
  for (int64_t b0 = (int64_t)j->tid * 4; b0 < j->B; b0 += (int64_t)j->nt * 4) {
    for (int64_t b = b0; b < b0 + 4 && b < j->B; b++) {
      const double* th = j->theta + b * j->ld;
      comps_t o;
      double lp = 0.0;
      if (j->what == CL_OUT_LOGPROB) {
        lp = log_prior(s, th);
        if (isinf(lp)) { /* sn/pantheon.py:90-92: the model is not evaluated */
          if (j->out) j->out[b] = -INFINITY;
          if (j->comps) for (int q = 0; q < 4; q++) j->comps[b * 4 + q] = NAN;
          continue;
        }
      }
      eval_one(s, th, &w, j->grid_builds, &o);
      double chi2 = o.sn + o.bao + o.cmb + o.extra;
      if (j->comps) { j->comps[b * 4] = o.sn; j->comps[b * 4 + 1] = o.bao; j->comps[b * 4 + 2] = o.cmb; j->comps[b * 4 + 3] = o.extra; }
      if (!j->out) continue;
      if (j->what == CL_OUT_CHI2) j->out[b] = chi2;
      else {
        double ll = o.guard ? s->guard_value : -0.5 * (chi2 + o.cc_norm);
        j->out[b] = j->what == CL_OUT_LOGLIKE ? ll : lp + ll;
      }
    }
  }
  work_free(&w);
  return NULL;
}

/* what: CL_OUT_*.  comps (nullable) receives [B][4] = sn, bao, cmb, extra.
 * grid_builds mirrors how many times the reference script rebuilds the z-grid per evaluation
 * (sn/pantheon.py:59-60,49 -> 2); it only affects timing, not values. */
int oracle_eval(const cl_spec* s, const double* theta, int64_t B, int64_t ld, int what, double* out, double* comps,
                int nthreads, int grid_builds) {
  int nt = n_threads_for(nthreads);
  if (nt > 256) nt = 256;
  if ((int64_t)nt > (B + 3) / 4) nt = (int)((B + 3) / 4);
  if (nt < 1) nt = 1;
  job_t jobs[256];
  pthread_t tids[256];
  for (int t = 0; t < nt; t++) {
    job_t j = {s, theta, B, ld, what, out, comps, grid_builds, t, nt, 0};
    jobs[t] = j;
  }
  if (nt == 1) { eval_worker(&jobs[0]); return jobs[0].err; }
  int started = 0, err = 0;
  for (int t = 0; t < nt; t++) { if (pthread_create(&tids[t], NULL, eval_worker, &jobs[t]) != 0) { err = -2; break; } started++; }
  for (int t = 0; t < started; t++) pthread_join(tids[t], NULL);
  for (int t = 0; t < started; t++) if (jobs[t].err) err = jobs[t].err;
  return err;
}

int oracle_distances(const cl_spec* s, const double* theta, int64_t B, int64_t ld, const double* zq, int64_t nq, double* DM, double* DH) {
  work_t w;
  if (work_init(s, &w) != 0) return -1;
  for (int64_t b = 0; b < B; b++) {
    cosmo_t c;
    unpack(s, theta + b * ld, &c);
    build_grid(s, &c, w.dh_grid, w.cum_dm);
    for (int64_t k = 0; k < nq; k++) {
      if (DM) DM[b * nq + k] = pchip_eval(zq[k], s->z_grid, w.cum_dm, w.dh_grid, s->n_grid, 1);
      if (DH) DH[b * nq + k] = C_KMS / H_of_z(s, &c, zq[k]);
    }
  }
  work_free(&w);
  return 0;
}

int oracle_bao_theory(const cl_spec* s, const double* theta, int64_t B, int64_t ld, double* out) {
  work_t w;
  if (work_init(s, &w) != 0) return -1;
  for (int64_t b = 0; b < B; b++) {
    cosmo_t c;
    unpack(s, theta + b * ld, &c);
    build_grid(s, &c, w.dh_grid, w.cum_dm);
    bao_theory(s, &c, theta + b * ld, &w, 0, out + b * s->n_bao);
  }
  work_free(&w);
  return 0;
}

int oracle_cmb(const cl_spec* s, const double* theta, int64_t B, int64_t ld, double* out) {
  work_t w;
  if (work_init(s, &w) != 0) return -1;
  for (int64_t b = 0; b < B; b++) {
    cosmo_t c;
    cmb_out_t o;
    unpack(s, theta + b * ld, &c);
    cmb_eval(s, &c, w.glx, w.glw, w.ngl, &o);
    double* r = out + b * 8;
    r[0] = o.v[0]; r[1] = o.v[1]; r[2] = o.v[2]; r[3] = o.zstar; r[4] = o.rs; r[5] = o.dm; r[6] = o.rdrag; r[7] = o.theta100;
  }
  work_free(&w);
  return 0;
}

int oracle_sn_residuals(const cl_spec* s, const double* theta, int64_t B, int64_t ld, double* out) {
  work_t w;
  if (work_init(s, &w) != 0) return -1;
  for (int64_t b = 0; b < B; b++) {
    cosmo_t c;
    unpack(s, theta + b * ld, &c);
    build_grid(s, &c, w.dh_grid, w.cum_dm);
    sn_residuals(s, theta + b * ld, &w, out + b * s->n_sn);
  }
  work_free(&w);
  return 0;
}
