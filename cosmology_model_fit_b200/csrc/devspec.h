// devspec.h — device-side view of one likelihood (a flattened cl_spec whose pointers are device pointers).
// Passed to the kernels by value as a __grid_constant__ parameter.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/cosmolike.h"

namespace cosmolike {

constexpr double kC_KMS = 299792.458;  // scipy.constants.c / 1000 (sn/pantheon.py:12)

// aux planes written by the stage-1/2 kernel, [AUX_COUNT][B] (structure of arrays)
// (raw block sums: the scalar algebra on top of them - compressed-CMB vector and chi2, chronometer normalisation, Gaussian
// terms - is done by the finalize kernel, one thread per row, instead of by one thread of the stage-1/2 CTA while the CTA's
// other warps wait for it at the next row's first barrier)
enum { AUX_BAO = 0, AUX_GL_DM = 1, AUX_GL_RS = 2, AUX_CC = 3, AUX_LOGPRIOR = 4, AUX_FLAGS = 5, AUX_SN_SMALL = 6, AUX_ZSTAR = 7, AUX_COUNT = 8 };
enum { FLAG_GUARD = 1, FLAG_OUTSIDE = 2 };

// what the stage-1/2 kernel produces
enum { MODE_EVAL = 0, MODE_DIST = 1, MODE_BAO = 2, MODE_CMB = 3, MODE_RESID = 4 };

struct DevSpec {
  int ndim, family, de_model;
  int col_H0, col_Om, Om_is_physical, col_obh2, col_och2, col_w0, col_wa;
  double H0_fixed, H0_scale;
  cl_cmb_consts k;
  double nu_inv_rho0;  // 1 / k.nu_rho0
  // z grid
  const double* z_grid;
  const double* grid_omnu_T;   // [17][256] Omnu_z(z) at node 16 t + k (theta-independent), transposed: FULL family grid pass
  const double* grid_ln1pz_T;  // [17][256] ln(1 + z) at the np.linspace nodes, transposed: wCDM / CPL grid pass
  int G, grid_uniform;
  double step;      // z_grid[i] == i*step bit-for-bit when grid_uniform (np.linspace from 0), except the last node
  double z_last;    // z_grid[G-1] (np.linspace stores `stop` there exactly)
  double inv_step;
  // SN block
  int n_sn, sn_small, sn_form, col_offset, n_vel, vel_mode, vel_pm1;
  int col_vel[CL_MAX_VEL];
  const double* sn_pack;   // [n_sn][4] = {z_cmb, first velocity-template weight, 1 + z_hel, obs} (general path)
  const double2* sn_zs;    // [n_sn] {1 + z_cmb, w} with a +-1 step template, {z_cmb, 0} without one (fast path)
  const double* sn_obsp;   // [n_sn] obs - 25 - 5 log10(1 + z_hel)
  const double2* sn_zs4;   // [4][sn_q4] quad-interleaved copy of sn_zs: entry [q][m] = supernova 4 m + q (fused digit planes)
  const double* sn_obsp4;  // [4][sn_q4] the same for sn_obsp; both padded with copies of the last supernova
  int sn_q4;
  const double *sn_vel_w, *sn_mat_small;
  const double* sn_mu_fixed;  // nullable [n_sn], NaN = model
  const double* sn_lin_t;     // [n_lin][n_sn]
  int n_lin, col_lin[CL_MAX_VEL];
  const double2* logtab;   // [128] {1/c_j, log10(c_j)} for fast_5log10 (second entry scaled by 5)
  double vel_scale;
  // BAO block
  int n_bao, dh_mode, rd_mode, col_rd;
  const double *bao_z, *bao_val, *bao_W;
  const int32_t* bao_qty;
  double rd_fixed;
  // CMB block
  int cmb_mode, n_gl;
  double cmb_prior[3], cmb_W[9];
  const double *gl_x, *gl_w;
  // CC block
  int n_cc, col_fcc;
  const double *cc_z, *cc_H, *cc_W;
  double cc_logdet, cc_norm_sign;
  // Gaussian terms, prior, guard
  int n_gc, n_gp, has_bounds, guard_cpl;
  int gc_col[CL_MAX_GAUSS], gp_col[CL_MAX_GAUSS];
  double gc_mean[CL_MAX_GAUSS], gc_sigma[CL_MAX_GAUSS], gp_mean[CL_MAX_GAUSS], gp_sigma[CL_MAX_GAUSS];
  double lo[CL_MAX_DIM], hi[CL_MAX_DIM];
  double lp_norm, guard_value;
};

// arguments of one stage-1/2 launch
struct Stage12Args {
  const double* theta;  // [B][ld]
  int64_t B, ld;
  int mode, what;
  int dbg;              // profiling experiments: bit0 skip the grid pass, bit1 skip the SN loop (results invalid)
  int zero_offset;      // moments mode: residuals with the magnitude offset set to 0
  double* R;            // MODE_EVAL: residual rows [B][ldR]; MODE_RESID: [B][n_sn]
  int64_t ldR;
  double* aux;          // [AUX_COUNT][B]
  const double* zq;     // MODE_DIST query redshifts
  int nq;
  double *outDM, *outDH;  // MODE_DIST
  double* out;            // MODE_BAO [B][n_bao], MODE_CMB [B][8]
  // fused digit planes (lean kernel, fast SN path, planes_ld <= 8 threads-per-CTA): the residual row goes straight to the int8
  // planes of the tcgen05 contraction ([planes_S][B][planes_ld], the layout of k_oz_slice_rows) and R is not written
  signed char* planes;
  int64_t planes_ld;
  double* rowscale;       // [B] 2^e of the row
  int planes_S;
  // accuracy-guard fallback pass: only the rows of flagged 128-row blocks are evaluated (guard[1] = rows flagged in this
  // pass: the kernel returns at once when it is 0; guard[2 + rb] != 0 marks row block rb)
  const int* guard;
};

}  // namespace cosmolike
