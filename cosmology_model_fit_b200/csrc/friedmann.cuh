// friedmann.cuh — stages 1 and 2 of the likelihood path, one CTA per parameter vector.
//
// Stage 1 (Friedmann distances): E(z) on the caller's z-grid, dh = c/H, cumulative trapezoid D_M with a
//   block-wide scan, both kept in shared memory (reference: sn/pantheon.py:28-40, bao/desi_cmb_des5y.py:60-66);
//   Gauss-Legendre D_M(z*) and r_s(z*) for the compressed CMB (cmb/data_planck_act_compression.py:160-197).
// Stage 2 (residuals): Hermite / PCHIP interpolation to every SN and BAO redshift (interpolator.py:71-119),
//   mu with the z_pec step correction (sn/pantheon.py:43-60), BAO ratios (bao/desi_cmb_union3.py:76-94),
//   (R, l_A, omega_b), the small quadratic forms, priors and guards.  The SN residual row goes to HBM for the
//   stage-3 chi-squared GEMM; everything else is reduced to a handful of scalars per theta.
//
// Arithmetic notes (all FP64; the FP64 pipe is the bound of this kernel):
//   * dh = (c/H0) * rsqrt(E^2) instead of c / (H0 * sqrt(E^2)): <= 2 ulp from the reference's three roundings.
//   * on the np.linspace grid the Hermite abscissa is t = z/step - i (one FMA) instead of (z - x_i)/h_i; the two
//     differ by ~1e-13 relative in t, i.e. < 1e-13 relative in D_M (the interpolant is C1 across nodes).
//   * log10 is a 128-entry table + degree-7 polynomial (abs. error < 3e-16 on the D_L range), no special cases.
//   * grid {D_M, dh} pairs are interleaved as double2 with one pad slot per 16 entries: conflict-free 16-byte
//     stores from the per-thread chunks, one 16-byte load per node when interpolating.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "devspec.h"
#include "digits.cuh"

namespace cosmolike {

// CL_S12_DBG=1 (profiling builds only) compiles the timing switches `dbg` bit 0 (skip the grid pass) and bit 1 (skip the SN pass) in;
// both invalidate the results, so the production build does not carry them
#ifndef CL_S12_DBG
#define CL_S12_DBG 0
#endif
#ifndef CL_S12_MINBLOCKS
#define CL_S12_MINBLOCKS 3
#endif
constexpr int kS12Threads = 256;
constexpr int kPPT = 16;  // grid points per thread: 256*16 = 4096 >= n_grid
constexpr int kMaxGrid = kS12Threads * kPPT;

__host__ __device__ __forceinline__ int pad_idx(int i) { return i + (i >> 4); }
constexpr int kPaddedGrid = kMaxGrid + kMaxGrid / 16;

struct Cosmo {
  double H0, h, K /* c/H0 */, Om, Or, Obc, Onu, Ode, obh2, och2, w0, wa;
};

// ---- shared-memory access by 32-bit shared-window offset (keeps address arithmetic out of the generic space) ----
__device__ __forceinline__ uint32_t s12_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 lds_d2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_d2(uint32_t addr, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}

// 1/sqrt(x) for normal positive x: MUFU.RSQ64H seed + one third-order correction (the sequence CUDA's rsqrt() uses,
// without its special-case branch); <= 1 ulp
__device__ __forceinline__ double rsqrt_pos(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = fma(-x, y0 * y0, 1.0);
  return fma(fma(e, 0.375, 0.5), y0 * e, y0);
}
__device__ __forceinline__ double fast_sqrt(double x) { return x * rsqrt_pos(x); }  // x > 0, <= 2 ulp
// 1/x for normal positive x: MUFU.RCP64H seed, then y0 (1 + e + e^2) with e = 1 - x y0; <= 1 ulp
__device__ __forceinline__ double rcp_pos(double x) {
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = fma(-x, y0, 1.0);
  return fma(y0, fma(e, e, e), y0);
}

__device__ __forceinline__ void unpack(const DevSpec& s, const double* __restrict__ th, Cosmo& c) {
  c.H0 = s.col_H0 >= 0 ? s.H0_scale * th[s.col_H0] : s.H0_fixed;
  c.h = c.H0 / 100;
  c.K = kC_KMS * rcp_pos(c.H0);  // c/H0 to 1 ulp
  c.w0 = s.col_w0 >= 0 ? th[s.col_w0] : -1.0;
  c.wa = s.col_wa >= 0 ? th[s.col_wa] : 0.0;
  c.Om = c.Or = c.Obc = c.Onu = c.Ode = c.obh2 = c.och2 = 0.0;
  if (s.family == CL_FAMILY_LATE) {
    c.Om = th[s.col_Om];
    if (s.Om_is_physical) c.Om = c.Om / (c.h * c.h);
    if (s.col_obh2 >= 0) c.obh2 = th[s.col_obh2];
  } else {
    double h2 = c.h * c.h;
    c.obh2 = th[s.col_obh2];
    c.och2 = th[s.col_och2];
    c.Onu = s.k.Omnu_h2 / h2;
    c.Or = s.k.Or_h2 / h2;
    c.Obc = (c.obh2 + c.och2) / h2;
    c.Ode = 1.0 - c.Obc - c.Or - c.Onu;
  }
}

// 5 log10(x) = 5 e log10(2) + 5 log10(c_j) + 5 log1p(r)/ln 10 coefficients, k = 1..7: (-1)^(k+1) 5 / (k ln 10)
__constant__ double kLog5Poly[8] = {0.0, 2.171472409516259, -1.0857362047581296, 0.7238241365054197, -0.5428681023790648,
                                    0.4342944819032518, -0.36191206825270983, 0.3102103442166084};
__constant__ double kLog5Two[2] = {1.5051499791443348 /* 30-bit head of 5 log10(2): e*head is exact */, -8.244288170221258e-10};

// 5-node massive-neutrino density (cmb/data_planck_act_compression.py:53-66):
//   Omnu_z = zp1^4 sum_k w_k sqrt(q_k^2 + (m0/zp1)^2) / rho0 = zp1^3 sum_k w_k sqrt(q_k^2 zp1^2 + m0^2) / rho0
// (the second form needs no division; it differs from the first by rounding only)
__device__ __forceinline__ double omnu_z(const DevSpec& s, double zp1) {
  const cl_cmb_consts& k = s.k;
  const double z2 = zp1 * zp1, m2 = k.nu_m0 * k.nu_m0;
  const double f0 = fast_sqrt(fma(k.nu_q2[0], z2, m2)), f1 = fast_sqrt(fma(k.nu_q2[1], z2, m2));
  const double f2 = fast_sqrt(fma(k.nu_q2[2], z2, m2)), f3 = fast_sqrt(fma(k.nu_q2[3], z2, m2));
  const double f4 = fast_sqrt(fma(k.nu_q2[4], z2, m2));
  const double ws = f0 * k.nu_w[0] + f1 * k.nu_w[1] + f2 * k.nu_w[2] + f3 * k.nu_w[3] + f4 * k.nu_w[4];
  return (z2 * zp1) * ws * s.nu_inv_rho0;
}

template <int DE>
__device__ __forceinline__ double fde(const Cosmo& c, double z, double zp1, double cubed) {
  if (DE == CL_DE_WCDM) return pow(zp1, 3 * (1.0 + c.w0));
  if (DE == CL_DE_CPL) return pow(zp1, 3 * (1 + c.w0 + c.wa)) * exp(-3 * c.wa * z / zp1);
  if (DE == CL_DE_THAWING) {
    double q = 2 * cubed / ((1.0 + c.w0) + (1.0 - c.w0) * cubed);
    return q * q;
  }
  return 1.0;
}

// E(z)^2 = (H/H0)^2 (sn/pantheon.py:28-31 late family, bao/desi_cmb_union3.py:37-57 full family)
template <int FAM, int DE>
__device__ __forceinline__ double E2_of_zp1(const DevSpec& s, const Cosmo& c, double z, double zp1) {
  double cubed = zp1 * zp1 * zp1;
  if (FAM == CL_FAMILY_LATE) {
    double de = (DE == CL_DE_LCDM) ? (1.0 - c.Om) : (1.0 - c.Om) * fde<DE>(c, z, zp1, cubed);
    return c.Om * cubed + de;
  }
  double radiation = c.Or * (cubed * zp1);
  double matter = c.Obc * cubed;
  double neutrino = c.Onu * omnu_z(s, zp1);
  double de = (DE == CL_DE_LCDM) ? c.Ode : c.Ode * fde<DE>(c, z, zp1, cubed);
  return radiation + matter + de + neutrino;
}
// E^2 at grid node i of the np.linspace grid.  Two theta-independent node tables (built once by cl_create) take the
// expensive pieces out of the 4000-node loop: ln(1+z_i) turns the wCDM / CPL power into one exp (instead of pow + exp +
// a division), and Omnu_z(z_i) replaces five square roots per node in the FULL family.
template <int FAM, int DE>
__device__ __forceinline__ double E2_at_node(const Cosmo& c, double zp1, double ln1pz_i, double omnu_i) {
  const double cubed = zp1 * zp1 * zp1;
  double f = 1.0;  // dark-energy density factor
  if (DE == CL_DE_WCDM) f = exp(3 * (1.0 + c.w0) * ln1pz_i);
  if (DE == CL_DE_CPL)  // z/(1+z) = 1 - 1/(1+z)
    f = exp(fma(3 * (1 + c.w0 + c.wa), ln1pz_i, -3 * c.wa * (1.0 - rcp_pos(zp1))));
  if (DE == CL_DE_THAWING) {  // (2 a^-3 / ((1 + w0) + (1 - w0) a^-3))^2 with the reciprocal to 1 ulp instead of an IEEE division per node
    const double q = (2 * cubed) * rcp_pos(fma(1.0 - c.w0, cubed, 1.0 + c.w0));
    f = q * q;
  }
  if (FAM == CL_FAMILY_LATE) return c.Om * cubed + ((DE == CL_DE_LCDM) ? (1.0 - c.Om) : (1.0 - c.Om) * f);
  // massive neutrinos: the theta-independent Omnu_z(z_i) comes from the static node table
  const double de = (DE == CL_DE_LCDM) ? c.Ode : c.Ode * f;
  return c.Or * (cubed * zp1) + c.Obc * cubed + de + c.Onu * omnu_i;
}
template <int FAM, int DE>
__device__ __forceinline__ double E2_of_z(const DevSpec& s, const Cosmo& c, double z) {
  return E2_of_zp1<FAM, DE>(s, c, z, 1.0 + z);
}
template <int FAM, int DE>
__device__ __forceinline__ double DH_of_z(const DevSpec& s, const Cosmo& c, double z) {  // c / H(z)
  return c.K * rsqrt_pos(E2_of_z<FAM, DE>(s, c, z));
}
template <int FAM, int DE>
__device__ __forceinline__ double H_of_z(const DevSpec& s, const Cosmo& c, double z) {
  return c.H0 * fast_sqrt(E2_of_z<FAM, DE>(s, c, z));
}

__device__ __forceinline__ double grid_z(const DevSpec& s, int i) {
  if (s.grid_uniform) return i == s.G - 1 ? s.z_last : (double)i * s.step;
  return __ldg(s.z_grid + i);
}

// 5 log10(x) for normal positive x: x = 2^e m, m in [1,2); table entry j = top 7 mantissa bits holds
// {fl(1/c_j), -5 log10 of that rounded value}; log10(m) = log10(m/c_j) + log10(c_j) with |m/c_j - 1| < 2^-8.
// `tab` is the 32-bit shared-memory address of the table.
__device__ __forceinline__ double fast_5log10(double x, uint32_t tab) {
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const double ed = (double)((hi >> 20) - 1023);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double2 tc = lds_d2(tab + (((uint32_t)hi >> 9) & 0x7f0u));  // 16 bytes per entry, index = (hi >> 13) & 127
  const double r = fma(m, tc.x, -1.0);   // |r| < 2^-8
  // degree 6 (the r^7 term is < 5e-18); the coefficients of r^4 .. r^6 are rounded to their high words (|effect| < 1e-17),
  // which makes them instruction immediates instead of constant-bank loads
  double p = fma(-0.3619120121002197, r, 0.4342944622039795);
  p = fma(p, r, -0.5428681373596191);
  p = fma(p, r, kLog5Poly[3]);
  p = fma(p, r, kLog5Poly[2]);
  p = fma(p, r, kLog5Poly[1]);
  double res = fma(ed, kLog5Two[0], tc.y);
  res = fma(ed, kLog5Two[1], res);
  return fma(r, p, res);
}
// true for a normal positive finite double (the domain of fast_5log10): one integer add and one unsigned compare on the high word
__device__ __forceinline__ bool normal_positive(double x) {
  return (uint32_t)(__double2hiint(x) - 0x00100000) < 0x7fe00000u;
}

// cubic Hermite segment in Horner form: y(t) on [node i, node i+1], t in [0,1], hd = h * slope
__device__ __forceinline__ double hermite_seg(double y0, double hd0, double y1, double hd1, double t) {
  const double D = y1 - y0;
  const double c2 = fma(3.0, D, -fma(2.0, hd0, hd1));
  const double c3 = (hd0 + hd1) - 2.0 * D;
  return fma(t, fma(t, fma(t, c3, c2), hd0), y0);
}

// D_M(xq): cubic Hermite with analytic node derivatives y' = dh (interp_hermite, interpolator.py:71-108,117-119);
// linear extrapolation with the end slope outside the grid.  gd[] holds {D_M, hscale * dh} per node (hscale = step
// on the np.linspace grid, 1 otherwise).  This is the general (any redshift, any grid) form; the SN loop has its own
// in-range fast path.
__device__ __forceinline__ double hermite_dm(const DevSpec& s, const double2* __restrict__ gdraw,
                                             const double* __restrict__ off, double xq) {
  const int G = s.G;
  auto node = [&](int i) { double2 v = gdraw[pad_idx(i)]; v.x += off[i >> 4]; return v; };
  if (s.grid_uniform) {
    if (xq > 0.0 && xq < s.z_last) {
      const int i = min((int)(xq * s.inv_step), G - 2);
      const double t = fma(xq, s.inv_step, -(double)i);
      const double2 a = node(i), b = node(i + 1);
      return hermite_seg(a.x, a.y, b.x, b.y, t);
    }
    if (xq <= 0.0) { const double2 a = node(0); return a.x + (a.y * s.inv_step) * xq; }
    const double2 b = node(G - 1);
    return b.x + (b.y * s.inv_step) * (xq - s.z_last);
  }
  const double x0 = __ldg(s.z_grid), xn = __ldg(s.z_grid + G - 1);
  if (xq <= x0) { const double2 a = node(0); return a.x + a.y * (xq - x0); }
  if (xq >= xn) { const double2 b = node(G - 1); return b.x + b.y * (xq - xn); }
  int lo = 0, hi = G;  // np.searchsorted(x, xq) - 1 (interpolator.py:94)
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(s.z_grid + mid) < xq) lo = mid + 1; else hi = mid;
  }
  const int i = lo - 1;
  const double xi = __ldg(s.z_grid + i), h_i = __ldg(s.z_grid + i + 1) - xi;
  const double2 a = node(i), b = node(i + 1);
  return hermite_seg(a.x, h_i * a.y, b.x, h_i * b.y, (xq - xi) / h_i);
}

__device__ __forceinline__ double sgn(double v) { return (double)((v > 0) - (v < 0)); }

// Fritsch-Carlson slope at node j of (z_grid, dh_grid) (_pchip_slopes, interpolator.py:5-68); local stencil
__device__ double pchip_slope(const DevSpec& s, const double2* __restrict__ gd, int j) {
  const int n = s.G;
  auto H = [&](int i) { return grid_z(s, i + 1) - grid_z(s, i); };
  const double ih = s.grid_uniform ? s.inv_step : 1.0;
  auto D = [&](int i) { return (gd[pad_idx(i + 1)].y * ih - gd[pad_idx(i)].y * ih) / H(i); };
  if (j == 0) {
    double h0 = H(0), h1 = H(1), d0 = D(0), d1 = D(1);
    double v = ((2 * h0 + h1) * d0 - h0 * d1) / (h0 + h1);
    if (d0 == 0.0 || sgn(v) != sgn(d0)) return 0.0;
    if (sgn(d0) != sgn(d1) && fabs(v) > fabs(3 * d0)) return 3 * d0;
    return v;
  }
  if (j == n - 1) {
    double h2 = H(n - 2), h3 = H(n - 3), d2 = D(n - 2), d3 = D(n - 3);
    double v = ((2 * h2 + h3) * d2 - h2 * d3) / (h2 + h3);
    if (d2 == 0.0 || sgn(v) != sgn(d2)) return 0.0;
    if (sgn(d2) != sgn(d3) && fabs(v) > fabs(3 * d2)) return 3 * d2;
    return v;
  }
  double dm1 = D(j - 1), di = D(j), hm1 = H(j - 1), hi = H(j);
  if (dm1 != 0.0 && di != 0.0 && dm1 * di > 0.0) {
    double w1 = 2.0 * hi + hm1, w2 = hi + 2.0 * hm1;
    return (w1 + w2) / (w1 / dm1 + w2 / di);
  }
  return 0.0;
}

// interp_pchip(xq, z_grid, dh_grid) (interpolator.py:111-114): clamps outside the grid.  Only <= 32 BAO points
// per theta use this, so it follows the reference formulas literally.
__device__ double pchip_dh(const DevSpec& s, const double2* __restrict__ gd, double xq) {
  const int G = s.G;
  const double ih = s.grid_uniform ? s.inv_step : 1.0;
  if (xq <= grid_z(s, 0)) return gd[0].y * ih;
  if (xq >= grid_z(s, G - 1)) return gd[pad_idx(G - 1)].y * ih;
  int i;
  if (s.grid_uniform) {
    // np.searchsorted(x, xq) - 1 on the np.linspace grid: x_i < xq <= x_(i+1); start from floor(xq / step) and fix up
    i = min(max((int)(xq * s.inv_step), 0), G - 2);
    while (i > 0 && grid_z(s, i) >= xq) i--;
    while (i < G - 2 && grid_z(s, i + 1) < xq) i++;
    // Equal intervals: the same interpolant in units of the node spacing.  With the differences delta_k = f_(k+1) - f_k the
    // Fritsch-Carlson slope times h is the harmonic mean 2 delta_(k-1) delta_k / (delta_(k-1) + delta_k) (0 at a sign change)
    // and (3 d0 - d1) / 2 with the end rules at the two boundary nodes: TWO divisions per query instead of the eleven of the
    // literal form below (a BAO point is a dependent chain on one thread that a whole CTA waits for; agreement ~1e-15).
    const double t = fma(xq, s.inv_step, -(double)i);
    auto f = [&](int k) { return gd[pad_idx(k)].y; };            // step * dh at node k
    const double f0 = f(i), f1 = f(i + 1), dc = f1 - f0;
    auto end_slope = [&](double d0, double d1) {                 // _pchip_slopes, interpolator.py:40-68, h0 = h1
      const double v = 0.5 * (3.0 * d0 - d1);
      if (d0 == 0.0 || sgn(v) != sgn(d0)) return 0.0;
      if (sgn(d0) != sgn(d1) && fabs(v) > fabs(3.0 * d0)) return 3.0 * d0;
      return v;
    };
    auto mid_slope = [&](double dm, double dp) { return (dm != 0.0 && dp != 0.0 && dm * dp > 0.0) ? 2.0 * dm * dp / (dm + dp) : 0.0; };
    const double m0 = i == 0 ? end_slope(dc, f(2) - f1) : mid_slope(f0 - f(i - 1), dc);
    const double m1 = i == G - 2 ? end_slope(dc, f0 - f(G - 3)) : mid_slope(dc, f(i + 2) - f1);
    const double t2 = t * t, t3 = t2 * t;
    const double h00 = 2 * t3 - 3 * t2 + 1, h10 = t3 - 2 * t2 + t, h01 = -2 * t3 + 3 * t2, h11 = t3 - t2;
    return (h00 * f0 + h10 * m0 + h01 * f1 + h11 * m1) * ih;
  } else {
    int lo = 0, hi = G;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (grid_z(s, mid) < xq) lo = mid + 1; else hi = mid;
    }
    i = lo - 1;
  }
  double xi = grid_z(s, i);
  double h_i = grid_z(s, i + 1) - xi;
  double t = (xq - xi) / h_i;
  double t2 = t * t, t3 = t2 * t;
  double h00 = 2 * t3 - 3 * t2 + 1, h10 = t3 - 2 * t2 + t, h01 = -2 * t3 + 3 * t2, h11 = t3 - t2;
  double d0 = pchip_slope(s, gd, i), d1 = pchip_slope(s, gd, i + 1);
  return h00 * (gd[pad_idx(i)].y * ih) + h10 * h_i * d0 + h01 * (gd[pad_idx(i + 1)].y * ih) + h11 * h_i * d1;
}

// closed-form fits (cmb/data_planck_act_compression.py:86-124).  Every term of the reference is a product of powers of
// wb and wm, e.g. (wb^b)^p (wm^m)^q = exp(p b ln wb + q m ln wm): two logs and one exp per term instead of the 14 pow()
// calls of the literal form (agreement ~4e-16 relative, the fits are needed to 1e-9).
// The fits are spread over threads: term k of the seven exponentials is one thread's work (two logs + one exp: the serial
// form - 2 logs + 7 exps on one thread - kept every other warp of the CTA waiting at the barrier behind the grid pass).
//   z* = T0 + s1 391.67 T1 + s2 937.42 T2,   r_d = 1 / (a1 T3 + a3 T4 + a6 T5) - a8 T6
constexpr int kFitTerms = 7;
// The two logarithms go through the shared-memory table of the SN pass (fast_5log10: |err| < 4e-16 absolute on 5 log10 x, i.e.
// ~2e-16 on ln x - a quarter of libm's instruction count on what is a serial chain at the top of every row); arguments that
// are not normal positive numbers take libm and give the reference's NaN / -inf.
__device__ __forceinline__ double cmb_fit_term(const cl_cmb_consts& k, double wb, double wm, int term, uint32_t tab) {
  constexpr double kLn10Over5 = 0.46051701859880917;   // ln x = (5 log10 x) ln(10) / 5
  const double lb = normal_positive(wb) ? fast_5log10(wb, tab) * kLn10Over5 : log(wb);
  const double lm = normal_positive(wm) ? fast_5log10(wm, tab) * kLn10Over5 : log(wm);
  const double zb = k.zstar_b * lb, zm = k.zstar_m * lm, rb = k.rdrag_b * lb, rm = k.rdrag_m * lm;
  double arg;
  switch (term) {
    case 0: arg = -0.7316314841257655 * zm; break;
    case 1: arg = 0.9368102670600895 * zb - 0.35300106475765136 * zm; break;
    case 2: arg = 0.0192950634264157 * zm - 0.04285000485853785 * zb; break;
    case 3: arg = 0.05032 * rb; break;
    case 4: arg = 0.7720642 * rb + 0.24346362 * rm; break;
    case 5: arg = 0.5350899 * rm; break;
    default: arg = -0.315473 * rm; break;
  }
  return exp(arg);
}
__device__ __forceinline__ double zstar_from_terms(const cl_cmb_consts& k, const double* t) {
  return t[0] + k.zstar_s1 * 391.6723594873167 * t[1] + k.zstar_s2 * 937.4224935298015 * t[2];
}
__device__ __forceinline__ double rdrag_from_terms(const double* t) {
  const double den = 0.00257366 * t[3] + 0.013 * t[4] + 0.00641072 * t[5];
  return 1.0 / den - 32.7525 * t[6];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sums of the values selected by `mask` (fixed order -> deterministic); result valid in every thread.  ONE barrier:
// the scratch is double-buffered by the caller (`s_red` alternates between rows), so a row's reads cannot meet the next
// row's writes - a thread reaches those only behind the next row's barriers.
// `finish` = false stops behind the barrier (the caller adds the eight partials of a value itself: block_partials_sum).
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* s_red /* [NV*8] */, unsigned mask, bool finish = true) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; q++) {
    if (!((mask >> q) & 1u)) continue;
    double w = warp_sum(v[q]);
    if (lane == 0) s_red[q * 8 + warp] = w;
  }
  __syncthreads();
  if (!finish) return;
#pragma unroll
  for (int q = 0; q < NV; q++) {
    if (!((mask >> q) & 1u)) continue;
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < kS12Threads / 32; w++) a += s_red[q * 8 + w];
    v[q] = a;
  }
}

// compressed-CMB vector from the two Gauss-Legendre sums (cmb/data_planck_act_compression.py:160-212)
__device__ __forceinline__ void cmb_vector(const DevSpec& s, const Cosmo& c, double zstar, double gl_dm, double gl_rs, double (&cmbv)[3],
                                           double& rs, double& dm) {
  dm = (zstar / 2.0) * gl_dm;
  rs = ((1.0 / (1.0 + zstar)) / 2.0) * gl_rs;
  const double Om_h2 = c.och2 + c.obh2 + s.k.Omnu_h2;
  if (s.cmb_mode == CL_CMB_THETA_WB_WM) { cmbv[0] = rs / dm; cmbv[1] = c.obh2; cmbv[2] = Om_h2; }
  else { cmbv[0] = 100 * sqrt(Om_h2) * dm / kC_KMS; cmbv[1] = M_PI * dm / rs; cmbv[2] = c.obh2; }
}

// the eight warp partials of value q in warp order (the order block_sum itself uses: same bits)
__device__ __forceinline__ double block_partials_sum(const double* s_red, int q) {
  double a = 0.0;
#pragma unroll
  for (int w = 0; w < kS12Threads / 32; w++) a += s_red[q * 8 + w];
  return a;
}

struct S12Smem {
  double2 gd[kPaddedGrid];  // {D_M relative to the first node of the 16-node chunk, hscale * dh} per grid node, padded
  double off[kS12Threads];  // D_M at the first node of each chunk
  double2 logtab[128];
  double wsum[8];
  double red[2][5 * 8];   // block-sum scratch, alternating between rows
  double vec[CL_MAX_BAO + CL_MAX_CC + CL_SN_SMALL_MAX];
  double scal[8];  // the seven exponential terms of the z* / r_drag fits
  double theta[2][CL_MAX_DIM];  // this row's and the next row's parameter vector
};

template <int FAM, int DE, int LEAN>
__global__ void __launch_bounds__(kS12Threads, CL_S12_MINBLOCKS)
k_friedmann_residuals(const __grid_constant__ DevSpec s, const __grid_constant__ Stage12Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // LEAN = 1: the instantiation for plain evaluations of a large SN block alone (no BAO / CMB / CC terms, no helper modes):
  // the probe switches below become compile-time constants and the dead phases drop out of the code (the full kernel is
  // ~140 KB of SASS, and instruction-fetch stalls showed in its profile)
  // LEAN = 2: plain evaluations of a large SN block on the fast path TOGETHER WITH the small probes (BAO / compressed CMB /
  // cosmic chronometers, switched at run time): stage 2 writes the digit planes itself here too.  The helper modes, the
  // general SN path and the FP64 row stores drop out; the BAO / CC residuals are formed BEFORE the supernova pass by the
  // CTA's LAST threads (which have no second supernova trip), so that the barrier of the row's power-of-two scale is also the
  // one behind which the small residual vectors are complete, and the Gauss-Legendre nodes follow the plane stores.
  // (macros, not locals: a local copy of a kernel parameter occupies a register in the full kernel, which is at its 80-register cap)
#define mode (LEAN ? (int)MODE_EVAL : a.mode)
#define n_bao (LEAN == 1 ? 0 : s.n_bao)
#define n_cc (LEAN == 1 ? 0 : s.n_cc)
#define cmb_mode (LEAN == 1 ? (int)CL_CMB_NONE : s.cmb_mode)
#define sn_small (LEAN ? false : (bool)s.sn_small)
  S12Smem& sm = *reinterpret_cast<S12Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Thread layout of the small probes.  They are long dependent chains on a handful of threads (a BAO point is two PCHIP
  // slopes with their divisions, a Gauss-Legendre node a 5-term neutrino sum and three square roots), so what matters is that
  // they do NOT share a warp: a warp executes its lanes' different branches one after the other, and with everything indexed
  // from thread 0 warp 0 ran the small-SN residuals, then the BAO points, then the chronometers, then its 32 Gauss-Legendre
  // nodes while the other warps waited at the barrier (30 % of the stall samples of bao/desi_cmb_union3.py).  Now: BAO and
  // chronometers from the top down (pt: warp 7), the small SN block below them (st: threads 223, 222, ...: warp 6 and down),
  // the 200 Gauss-Legendre nodes from the bottom up (gt: warps 0-6; from the top down in the small-probe instantiation, whose
  // low warps carry the second supernova trip); the block sums leave through eight lanes of warp 0 (one aux plane each).
  const int pt = kS12Threads - 1 - tid;
  const int st = kS12Threads - 33 - tid;
  const int gt = LEAN == 2 ? pt : tid;
  const int G = s.G;
  if (tid < 128) sm.logtab[tid] = s.logtab[tid];  // visible after the first __syncthreads of the loop body
  const uint32_t gd_addr = s12_smem_u32(sm.gd), tab_addr = s12_smem_u32(sm.logtab);

  if (a.guard != nullptr && a.guard[1] == 0) return;   // fallback pass with nothing flagged
  if (tid < s.ndim && (int64_t)blockIdx.x < a.B) sm.theta[0][tid] = a.theta[(int64_t)blockIdx.x * a.ld + tid];
  __syncthreads();
  int tb = 0;
  for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x, tb ^= 1) {
    const double* __restrict__ th = sm.theta[tb];
    const double* __restrict__ scal = sm.scal;
    // theta of the next row is staged into the other buffer while this row computes (its L2 latency would otherwise
    // stall every warp); the write happens after the first barrier of the iteration, when no thread can still be
    // reading that buffer for the previous row
    const bool stage_next = tid < s.ndim && b + gridDim.x < a.B;
    double th_next = 0.0;
    if (stage_next) th_next = a.theta[(b + gridDim.x) * a.ld + tid];  // consumed only at the end of the iteration
    if (a.guard != nullptr && a.guard[2 + (b >> 7)] == 0) {   // fallback pass: this row's block is not flagged (uniform across the CTA)
      __syncthreads();
      if (stage_next) sm.theta[tb ^ 1][tid] = th_next;
      __syncthreads();
      continue;
    }
    Cosmo c;
    unpack(s, th, c);

    // ---- prior box / guard: rows that the reference never evaluates (sn/pantheon.py:80-92) ----
    double lp = 0.0;
    int flags = 0;
    if (mode == MODE_EVAL) {
      if (a.what == CL_OUT_LOGPROB) {
        if (s.has_bounds)
          for (int j = 0; j < s.ndim; j++)
            if (!(s.lo[j] < th[j] && th[j] < s.hi[j])) flags |= FLAG_OUTSIDE;
        lp = s.lp_norm;
        for (int g = 0; g < s.n_gp; g++) {
          double r = th[s.gp_col[g]] - s.gp_mean[g];
          lp += -0.5 * (r * r) / (s.gp_sigma[g] * s.gp_sigma[g]);
        }
      }
      if (s.guard_cpl && DE == CL_DE_CPL && c.w0 + c.wa >= 0.0 && a.what != CL_OUT_CHI2) flags |= FLAG_GUARD;
      if (flags) {  // uniform across the CTA
        if (tid == 0) {
          a.aux[AUX_FLAGS * a.B + b] = (double)flags;
          a.aux[AUX_LOGPRIOR * a.B + b] = lp;
        }
          __syncthreads();
        if (stage_next) sm.theta[tb ^ 1][tid] = th_next;
        __syncthreads();
        continue;
      }
    }

    const bool need_grid = (mode == MODE_EVAL && (s.n_sn > 0 || n_bao > 0)) || mode == MODE_DIST ||
                           mode == MODE_BAO || mode == MODE_RESID;
    const bool need_cmb = (mode == MODE_EVAL && cmb_mode != CL_CMB_NONE) || mode == MODE_CMB;
    const bool need_rd = s.rd_mode == CL_RD_FIT && ((mode == MODE_EVAL && n_bao > 0) || mode == MODE_BAO || mode == MODE_CMB);

    // ================= stage 1: dh = c/H on the grid, cumulative trapezoid =================
    // thread t owns nodes [16t, 16t+16) and the 16 intervals that start at them.  It stores {D_M relative to its first
    // node, hscale * dh} per node as it goes (hscale = step on the np.linspace grid: the interpolant wants h * slope);
    // the block-wide exclusive scan of the chunk totals goes to sm.off[] and is added when a node is read.
    double run = 0.0;
    const int i0 = tid * kPPT;
    if (need_grid) {
      const uint32_t dst = gd_addr + (uint32_t)pad_idx(i0) * 16u;
      if (CL_S12_DBG && (a.dbg & 1)) {   // timing experiment (profiling build): no grid pass, results invalid
        run = 1.0;
      } else if (s.grid_uniform) {
        // Static node tables (ln(1+z_i), Omnu_z(z_i)).  A thread owns 16 CONSECUTIVE nodes, so the tables are stored transposed
        // ([k][thread] = node 16 thread + k, k = 0..16, built by cl_create): every load is coalesced across the warp, nothing is
        // staged in shared memory, and the loads run a few nodes ahead of their use (the stores to shared memory are
        // volatile asm statements with a memory clobber, so the compiler keeps this placement).
        constexpr bool kTabLn = DE == CL_DE_WCDM || DE == CL_DE_CPL, kTabOm = FAM == CL_FAMILY_FULL;
        constexpr int kAhead = 2;
        double lnv[kPPT + 1], omv[kPPT + 1];   // indexed by compile-time constants after unrolling: only a window is ever live
        auto fetch = [&](int k) {
          lnv[k] = 0.0; omv[k] = 0.0;
          // (volatile: the compiler would otherwise hoist all 17 loads to the top of the unrolled loop and spill)
          if (kTabLn) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(lnv[k]) : "l"(s.grid_ln1pz_T + k * kS12Threads + tid));
          if (kTabOm) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(omv[k]) : "l"(s.grid_omnu_T + k * kS12Threads + tid));
        };
        auto tab = [&](int k, double& ln, double& om) { ln = lnv[k]; om = omv[k]; if (k + kAhead <= kPPT) fetch(k + kAhead); };
#pragma unroll
        for (int k = 0; k < kAhead; k++) fetch(k);
        const double Ks = c.K * s.step, zp1_0 = fma((double)i0, s.step, 1.0);
        double ln_i, om_i;
        tab(0, ln_i, om_i);
        double prev = Ks * rsqrt_pos(E2_at_node<FAM, DE>(c, zp1_0, ln_i, om_i));
        if (i0 < G) {
          // The chunk that holds the grid's last node runs the same straight-line loop as every other one (a separate
          // predicated loop made its warp execute both - and the whole CTA wait for it at the barrier behind the grid pass):
          // the node tables and the shared-memory slots exist for all 16 * 256 positions, what is computed and stored beyond
          // node G - 1 is never read (the interpolants stop at node G - 1, the chunk offsets behind this chunk are unused).
#pragma unroll
          for (int k = 0; k < kPPT; k++) {
            const double zp1 = fma((double)(k + 1), s.step, zp1_0);  // 1 + z_grid[i0+k+1] to 1 ulp
            tab(k + 1, ln_i, om_i);
            const double nxt = Ks * rsqrt_pos(E2_at_node<FAM, DE>(c, zp1, ln_i, om_i));
            sts_d2(dst + 16u * k, run, prev);
            run = fma(prev + nxt, 0.5, run);
            prev = nxt;
          }
          // the pad slot behind the chunk gets hd of the NEXT chunk's first node, so that the SN pass finds hd_{j+1} at a
          // fixed 24 bytes behind node j's pair, also across a chunk boundary
          asm volatile("st.shared.f64 [%0], %1;" ::"r"(dst + 16u * kPPT + 8u), "d"(prev) : "memory");
        }
      } else if (i0 < G) {
        double prev = DH_of_z<FAM, DE>(s, c, __ldg(s.z_grid + i0));
#pragma unroll
        for (int k = 0; k < kPPT; k++) {
          if (i0 + k < G) sts_d2(dst + 16u * k, run, prev);
          if (i0 + k + 1 < G) {
            const double z0 = __ldg(s.z_grid + i0 + k), z1 = __ldg(s.z_grid + i0 + k + 1);
            const double nxt = DH_of_z<FAM, DE>(s, c, z1);
            run += ((prev + nxt) / 2) * (z1 - z0);
            prev = nxt;
          }
        }
      }
      // block-exclusive scan of the per-thread totals
      double inc = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) sm.wsum[warp] = inc;
      run = inc - run;  // exclusive prefix within the warp
    }
    // the closed-form fits: one exponential term per thread on the CTA's last threads (they own no grid nodes: 16 * 249 < G)
    if (tid >= kS12Threads - kFitTerms && (need_cmb || need_rd)) {
      const double wm = (FAM == CL_FAMILY_FULL) ? c.och2 + c.obh2 + s.k.Omnu_h2 : c.Om * c.h * c.h;
      const int term = tid - (kS12Threads - kFitTerms);
      if (term >= 3 ? need_rd : need_cmb) sm.scal[term] = cmb_fit_term(s.k, c.obh2, wm, term, tab_addr);
    }
    __syncthreads();

    if (need_grid) {
      // off = run + (the totals of the warps before this one), added in warp order.  The eight totals are loaded up front and
      // the terms of later warps enter as +0.0 (which changes nothing): the same additions in the same order as the obvious
      // loop, without its dependent load -> add -> compare -> branch per trip - the last warp ran seven of those between the
      // two barriers that bracket this step while the other seven warps waited for it (7 % of the row's stall samples for
      // 4 % of its instructions, profiles/r04c ncu source view).
      const double2* w2 = reinterpret_cast<const double2*>(sm.wsum);
      const double2 wa = w2[0], wb = w2[1], wc = w2[2], wd = w2[3];
      double off = run;
      off += warp > 0 ? wa.x : 0.0;
      off += warp > 1 ? wa.y : 0.0;
      off += warp > 2 ? wb.x : 0.0;
      off += warp > 3 ? wb.y : 0.0;
      off += warp > 4 ? wc.x : 0.0;
      off += warp > 5 ? wc.y : 0.0;
      off += warp > 6 ? wd.x : 0.0;
      sm.off[tid] = off;
      __syncthreads();
    }

    // ================= helper outputs =================
    if (mode == MODE_DIST) {
      for (int q = tid; q < a.nq; q += kS12Threads) {
        double z = a.zq[q];
        if (a.outDM) a.outDM[b * a.nq + q] = hermite_dm(s, sm.gd, sm.off, z);
        if (a.outDH) a.outDH[b * a.nq + q] = DH_of_z<FAM, DE>(s, c, z);
      }
      if (stage_next) sm.theta[tb ^ 1][tid] = th_next;
      __syncthreads();
      continue;
    }

    // BAO theory (bao_theory, bao/desi_cmb_union3.py:76-94 / bao/desi_cmb_pantheon.py:85-99) and cosmic chronometers
    // (ohd/cc.py:22-26): residual vectors into sm.vec
    auto small_probe_residuals = [&]() {
      if ((mode == MODE_EVAL || mode == MODE_BAO) && n_bao > 0 && pt < n_bao) {
        double rd = s.rd_mode == CL_RD_FIXED ? s.rd_fixed : (s.rd_mode == CL_RD_PARAM ? th[s.col_rd] : rdrag_from_terms(scal));
        double z = __ldg(s.bao_z + pt);
        double DM = hermite_dm(s, sm.gd, sm.off, z);
        double DH = s.dh_mode == CL_DH_PCHIP ? pchip_dh(s, sm.gd, z) : DH_of_z<FAM, DE>(s, c, z);
        int q = __ldg(s.bao_qty + pt);
        double v;
        if (q == CL_BAO_DV_OVER_RS) v = cbrt(z * DH * (DM * DM)) / rd;  // reference: x ** (1/3), equal to ~1e-16
        else if (q == CL_BAO_DM_OVER_RS) v = DM / rd;
        else if (q == CL_BAO_DH_OVER_RS) v = DH / rd;
        else v = DM / DH;
        if (mode == MODE_BAO) a.out[b * n_bao + pt] = v;
        else sm.vec[pt] = __ldg(s.bao_val + pt) - v;
      }
      if (mode == MODE_EVAL && pt < n_cc)
        sm.vec[CL_MAX_BAO + pt] = __ldg(s.cc_H + pt) - H_of_z<FAM, DE>(s, c, __ldg(s.cc_z + pt));
    };
    if (LEAN == 2) small_probe_residuals();

    // ================= stage 2: residuals =================
    const int n_sn = s.n_sn;
    bool row_synced = false;   // the fused digit-plane path ends the row's shared-memory traffic with its own barrier
    if ((mode == MODE_EVAL || mode == MODE_RESID) && n_sn > 0 && !(CL_S12_DBG && (a.dbg & 2))) {
      const double offset = (s.col_offset >= 0 && !a.zero_offset) ? th[s.col_offset] : 0.0;
      const int64_t ld = mode == MODE_RESID ? (int64_t)n_sn : a.ldR;
      double* __restrict__ Rrow = a.R + b * ld;
      const bool to_smem = sn_small && mode == MODE_EVAL;
      if (LEAN == 2 || (s.grid_uniform && (s.n_vel == 0 || s.vel_pm1) && s.sn_mu_fixed == nullptr && s.n_lin == 0)) {
        // fast path.  Static per-SN operands: zs = {1 + z_cmb, w} (or {z_cmb, 0} without a velocity template) and
        // obsp = obs - 25 - 5 log10(1 + z_hel), so that delta = obsp - offset - 5 log10 D_M(z_cosmo):
        // mu_theory + mu_corr = 25 + 5 log10((1+z_hel) D_M(z_cosmo)), D_M(z_cmb) cancels.
        // step template (w = +-1): 1/(1 + w z_pec) = ravg + w rdif with ravg = 1/(1 - z_pec^2), rdif = -z_pec ravg
        const bool shift = s.n_vel > 0;
        double ravg = 1.0, rdif = 0.0;
        if (shift) {
          const double z_pec = (s.vel_scale * th[s.col_vel[0]]) * (1.0 / kC_KMS);
          ravg = rcp_pos(fma(-z_pec, z_pec, 1.0));
          rdif = -z_pec * ravg;
        }
        const double inv_step = s.inv_step, obs_off = -offset;
        const uint32_t imax = (uint32_t)(G - 2);
        // the shared-memory window base is laundered through an empty asm: the compiler otherwise rebuilds it (S2R + MOV + LEA)
        // for every supernova instead of keeping one register
        uint32_t gd_base = gd_addr;
        asm volatile("" : "+r"(gd_base));
        const uint32_t off_base = gd_base + (s12_smem_u32(sm.off) - gd_addr), tab_base = gd_base + (tab_addr - gd_addr);
        // floor(z/step) without conversion instructions: adding 1.5 * 2^52 leaves round(x) in the low word
        const double kMagic = 6755399441055744.0;
        auto resid = [&](double2 zs, double ob) -> double {
          const double zq = shift ? fma(zs.x, fma(zs.y, rdif, ravg), -1.0) : zs.x;  // (1+z_cmb)/(1+z_pec) - 1
          const double w = fma(zq, inv_step, -0.5) + kMagic;  // round(z/step - 1/2): the interval index (ties: either side)
          const uint32_t j = (uint32_t)__double2loint(w);
          // inside the grid <=> 0 <= round(z/step - 1/2) <= G - 2: the high word of w is then the magic constant's and the low
          // word the index (two integer compares; NaN, negative and huge redshifts fail the first)
          if (__double2hiint(w) == 0x43380000 && j <= imax) {
            const double t = fma(zq, inv_step, -(w - kMagic));  // in [0, 1] up to rounding
            const uint32_t a0 = gd_base + ((j + (j >> 4)) << 4);
            const double2 n0 = lds_d2(a0);
            double hd1, base;   // hd of node j + 1: 24 bytes on (the pad slot behind a chunk carries the next chunk's first hd)
            asm volatile("ld.shared.f64 %0, [%1+24];" : "=d"(hd1) : "r"(a0));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(base) : "r"(off_base + ((j >> 4) << 3)));
            // On the trapezoid-built grid D_M(i+1) - D_M(i) = (hd_i + hd_{i+1})/2, so the cubic Hermite segment
            // (interpolator.py:96-108) collapses to the quadratic y_i + t hd_i + t^2 (hd_{i+1} - hd_i)/2; the dropped
            // cubic coefficient is pure rounding of the cumulative sum (~1e-16 D_M).
            const double dm = fma(t, fma(0.5 * t, hd1 - n0.y, n0.y), n0.x + base);
            // the table log10 decodes the bits of a NORMAL POSITIVE number; a NaN / Inf / non-positive distance (NaN or Inf
            // parameters, E^2 < 0) takes the libm call below, which gives the reference's NaN (np.log10)
            if (normal_positive(dm)) return (ob + obs_off) - fast_5log10(dm, tab_base);
            return (ob + obs_off) - 5.0 * log10(dm);
          }
          return (ob + obs_off) - 5.0 * log10(hermite_dm(s, sm.gd, sm.off, zq));  // outside the grid / non-positive distance
        };
        // Four supernovae at once, branch-free: the in-range case (every supernova of a sane parameter vector) is straight-line
        // code, so the four dependency chains interleave and their shared-memory latencies overlap; whether ALL four took
        // the fast formulas is one flag, and the rare group with an exception (a redshift outside the grid, a NaN / Inf /
        // non-positive distance) is redone by the scalar function above.  Same arithmetic per supernova: same bits.
        auto resid4 = [&](const double2 (&zs)[4], const double (&ob)[4], double (&d)[4]) {
          double zq[4], w[4], dm[4];
          uint32_t j[4];
          bool ok = true;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            zq[q] = shift ? fma(zs[q].x, fma(zs[q].y, rdif, ravg), -1.0) : zs[q].x;
            w[q] = fma(zq[q], inv_step, -0.5) + kMagic;
            const uint32_t jj = (uint32_t)__double2loint(w[q]);
            ok = ok && __double2hiint(w[q]) == 0x43380000 && jj <= imax;
            j[q] = min(jj, imax);   // a valid address also when the group is going to be redone
          }
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const double t = fma(zq[q], inv_step, -(w[q] - kMagic));
            const uint32_t a0 = gd_base + ((j[q] + (j[q] >> 4)) << 4);
            const double2 n0 = lds_d2(a0);
            double hd1, base;
            asm volatile("ld.shared.f64 %0, [%1+24];" : "=d"(hd1) : "r"(a0));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(base) : "r"(off_base + ((j[q] >> 4) << 3)));
            dm[q] = fma(t, fma(0.5 * t, hd1 - n0.y, n0.y), n0.x + base);
            ok = ok && normal_positive(dm[q]);
          }
#pragma unroll
          for (int q = 0; q < 4; q++) d[q] = (ob[q] + obs_off) - fast_5log10(dm[q], tab_base);
          if (!ok) {   // (unrolled: a run-time index would move the operand arrays to local memory)
#pragma unroll
            for (int q = 0; q < 4; q++) d[q] = resid(zs[q], ob[q]);
          }
        };
        // static operands of supernovae 4 m .. 4 m + 3 from the quad-interleaved copies ([q][m], coalesced for consecutive m)
        const int q4 = s.sn_q4;
        auto load4 = [&](int m, double2 (&zs)[4], double (&ob)[4]) {
#pragma unroll
          for (int q = 0; q < 4; q++) { zs[q] = __ldg(s.sn_zs4 + q * q4 + m); ob[q] = __ldg(s.sn_obsp4 + q * q4 + m); }
        };
        if (LEAN == 2 || (LEAN && a.planes != nullptr)) {
          // Fused digit planes: stage 2 writes the int8 planes of the tcgen05 contraction itself and the FP64 row never goes
          // through HBM.  A thread owns FOUR CONSECUTIVE supernovae per trip (4 m .. 4 m + 3 for m = tid, tid + 256), keeps its
          // <= 8 residuals in registers, the CTA agrees on the row's power-of-two scale with one barrier, and the balanced
          // base-256 digits of the four values are one 64-bit add + XOR each and a 4 x S byte transpose (oz_digits4, the
          // arithmetic of k_oz_slice_rows_reg: the planes are the same bits) - one 32-bit store per plane and trip, 128
          // contiguous bytes per warp.  The static operands come from the quad-interleaved copies sn_zs4 / sn_obsp4
          // ([4][sn_q4] with entry [q][m] = supernova 4 m + q, padded with copies of the last supernova), so every load is coalesced.
          double dv[8];
          const int m4 = (int)(a.planes_ld >> 2);        // groups of four columns, a multiple of 32: whole warps drop out
#pragma unroll
          for (int trip = 0; trip < 2; trip++) {
            const int m = tid + trip * kS12Threads;
            double d[4] = {0.0, 0.0, 0.0, 0.0};
            if (m < m4) {
              double2 zs[4];
              double ob[4];
              load4(m, zs, ob);
              resid4(zs, ob, d);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) dv[4 * trip + q] = d[q];
          }
          // the row's power-of-two scale only needs the largest EXPONENT: an integer maximum over the high words of |delta|
          // (NaN / Inf sort above every finite value), one REDUX per warp and one barrier
          uint32_t mh = 0;
#pragma unroll
          for (int k = 0; k < 8; k++) mh = max(mh, (uint32_t)__double2hiint(dv[k]) & 0x7fffffffu);
          mh = __reduce_max_sync(0xffffffffu, mh);
          // its own slot: the SN-only kernel has no block sums; with small probes the grid scan's warp totals are free here
          // (written before the row's first barrier, read before its second)
          uint32_t* s_mh = reinterpret_cast<uint32_t*>(LEAN == 2 ? sm.wsum : sm.red[0]);
          if (lane == 0) s_mh[warp] = mh;
          // SN-only kernel: this barrier is also the row's LAST one - every thread has finished reading the grid nodes, so the
          // next row may overwrite them, and the next row's parameter vector (parked here, before the barrier) is visible behind
          // it.  With small probes it is the barrier behind which sm.vec is complete; the row ends with the block sum's.
          if (LEAN != 2) {
            if (stage_next) sm.theta[tb ^ 1][tid] = th_next;
            row_synced = true;
          }
          __syncthreads();
          {
            const uint4 m0 = *reinterpret_cast<const uint4*>(s_mh), m1 = *reinterpret_cast<const uint4*>(s_mh + 4);
            mh = max(max(max(m0.x, m0.y), max(m0.z, m0.w)), max(max(m1.x, m1.y), max(m1.z, m1.w)));
          }
          const bool bad = mh >= 0x7ff00000u;
          const int S = a.planes_S, frac_bits = 6 + 8 * (S - 1);
          // 2^e = the next power of two above max |delta| (k_oz_slice_rows: ilogb + 1); a zero / denormal row takes the floor
          const int e = (mh >> 20) ? (int)(mh >> 20) - 1022 : -900;
          const double up = __longlong_as_double((long long)(1023 + frac_bits - e) << 52);   // 2^(frac_bits - e)
          if (tid == 0) a.rowscale[b] = bad ? __longlong_as_double(0x7ff8000000000000LL) : __longlong_as_double((long long)(1023 + e) << 52);
          signed char* __restrict__ prow = a.planes + b * a.planes_ld;
          const int64_t pstride = a.B * a.planes_ld;
#pragma unroll
          for (int trip = 0; trip < 2; trip++) {
            const int m = tid + trip * kS12Threads;
            if (m < m4) {
              long long v[4];
#pragma unroll
              for (int q = 0; q < 4; q++)   // |v| <= 2^frac_bits for a finite row (a bad row's digits are garbage: its scale is NaN)
                v[q] = __double2ll_rn(dv[4 * trip + q] * up);
              if (4 * m + 4 > n_sn) {   // the group that straddles the end of the row, and the padding columns: zero digits
#pragma unroll
                for (int q = 0; q < 4; q++) if (4 * m + q >= n_sn) v[q] = 0LL;
              }
              signed char* dst = prow + 4 * m;
#define CL_STORE_PLANES(SS)                                                                 \
              { uint32_t w[SS]; oz_digits4<SS>(v, w);                                        \
                _Pragma("unroll") for (int p = 0; p < SS; p++) { *reinterpret_cast<uint32_t*>(dst) = w[p]; dst += pstride; } }
              if (S == 7) CL_STORE_PLANES(7) else if (S == 6) CL_STORE_PLANES(6) else CL_STORE_PLANES(5)
#undef CL_STORE_PLANES
            }
          }
        } else {
        // four consecutive supernovae per thread and trip; the residuals leave as two 16-byte stores (the row is 128-byte aligned)
        double* __restrict__ outp = to_smem ? (sm.vec + CL_MAX_BAO + CL_MAX_CC) : Rrow;
        const int m4 = (n_sn + 3) >> 2;
        for (int m = to_smem ? st : tid; (unsigned)m < (unsigned)m4; m += kS12Threads) {   // (small block: threads 223, 222, ...)
          double2 zs[4];
          double ob[4], d[4];
          load4(m, zs, ob);
          resid4(zs, ob, d);
          if (4 * m + 4 <= n_sn && !to_smem && mode != MODE_RESID) {   // MODE_RESID rows are packed (pitch n_sn): scalar stores
            *reinterpret_cast<double2*>(outp + 4 * m) = make_double2(d[0], d[1]);
            *reinterpret_cast<double2*>(outp + 4 * m + 2) = make_double2(d[2], d[3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; q++) if (4 * m + q < n_sn) outp[4 * m + q] = d[q];
          }
        }
        }
      } else {
        const double2* __restrict__ pack = reinterpret_cast<const double2*>(s.sn_pack);
        double vamp[CL_MAX_VEL];
#pragma unroll
        for (int k = 0; k < CL_MAX_VEL; k++) vamp[k] = k < s.n_vel ? s.vel_scale * th[s.col_vel[k]] : 0.0;
        // D_M(zq): inside a np.linspace grid the same collapsed (quadratic) Hermite segment as the fast path, the literal
        // reference formulas otherwise
        const bool uni = s.grid_uniform != 0;
        const double inv_step_g = s.inv_step, z_last_g = s.z_last;
        const int imax_g = G - 2;
        const uint32_t off_addr_g = s12_smem_u32(sm.off);
        auto dm_of = [&](double zq) -> double {
          if (uni && zq > 1e-9 && zq < z_last_g) {
            const double kMagicG = 6755399441055744.0;
            const double w = fma(zq, inv_step_g, -0.5) + kMagicG;
            const int j = min(__double2loint(w), imax_g);
            const double t = fma(zq, inv_step_g, -(w - kMagicG));
            const double2 n0 = lds_d2(gd_addr + ((uint32_t)(j + (j >> 4)) << 4));
            const double2 n1 = lds_d2(gd_addr + ((uint32_t)(j + 1 + ((j + 1) >> 4)) << 4));
            double base;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(base) : "r"(off_addr_g + ((uint32_t)(j >> 4) << 3)));
            return fma(t, fma(0.5 * t, n1.y - n0.y, n0.y), n0.x + base);
          }
          return hermite_dm(s, sm.gd, sm.off, zq);
        };
        // (a small SN block sits on threads 223, 222, ...: see the thread layout at the top)
        for (int i = to_smem ? st : tid; (unsigned)i < (unsigned)n_sn; i += kS12Threads) {
          const double2 p0 = __ldg(pack + 2 * i), p1 = __ldg(pack + 2 * i + 1);  // {z_cmb, w0}, {1+z_hel, obs}
          double zq = p0.x;
          if (s.n_vel > 0) {
            double v_km_s = 0.0;
            for (int k = 0; k < s.n_vel; k++) v_km_s += vamp[k] * __ldg(s.sn_vel_w + (size_t)k * n_sn + i);
            const double z_pec = v_km_s * (1.0 / kC_KMS);
            if (s.vel_mode == CL_VEL_DIVIDE) zq = fma(1.0 + p0.x, rcp_pos(1.0 + z_pec), -1.0);   // |z_pec| << 1: 1 + z_pec > 0
            else zq = fmax((1.0 + p0.x) * (1.0 + z_pec) - 1.0, 1e-8);
          }
          const double DM = dm_of(zq);
          // 5 log10 through the shared-memory table of the fast path (|err| < 4e-16) for normal positive arguments
          auto log5 = [&](double x) { return normal_positive(x) ? fast_5log10(x, tab_addr) : 5.0 * log10(x); };
          double mu;
          if (s.sn_mu_fixed != nullptr && isfinite(__ldg(s.sn_mu_fixed + i))) {
            // SH0ES calibrator: fixed distance modulus, mu_corr = 5 log10(D_M(z_cosmo)/D_M(z_cmb)) still applies
            mu = __ldg(s.sn_mu_fixed + i);
            if (s.n_vel > 0) mu += log5(DM) - log5(dm_of(p0.x));
          } else {
            mu = 25.0 + log5(p1.x * DM);
          }
          double lin = 0.0;
          for (int k = 0; k < s.n_lin; k++) lin += th[s.col_lin[k]] * __ldg(s.sn_lin_t + (size_t)k * n_sn + i);
          const double d = (p1.y - (offset + lin)) - mu;
          if (to_smem) sm.vec[CL_MAX_BAO + CL_MAX_CC + i] = d; else Rrow[i] = d;
        }
      }
    }
    if (mode == MODE_RESID) {
      if (stage_next) sm.theta[tb ^ 1][tid] = th_next;
      __syncthreads();
      continue;
    }

    if (LEAN != 2) small_probe_residuals();
    if (mode == MODE_BAO) {
      if (stage_next) sm.theta[tb ^ 1][tid] = th_next;
      __syncthreads();
      continue;
    }

    // Gauss-Legendre integrands (cmb/data_planck_act_compression.py:160-197): thread q < n_gl -> D_M node,
    // n_gl <= q < 2 n_gl -> r_s node (in scale factor)
    double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    double zstar = 0.0;
    if (need_cmb) {
      zstar = zstar_from_terms(s.k, scal);
      if (gt < s.n_gl) {
        double hw = zstar / 2.0;
        double z = hw * __ldg(s.gl_x + gt) + hw;
        v[0] = __ldg(s.gl_w + gt) * DH_of_z<FAM, DE>(s, c, z);
      } else if (gt < 2 * s.n_gl) {
        int q = gt - s.n_gl;
        double a_lim = 1.0 / (1.0 + zstar);
        double hw = a_lim / 2.0;
        double av = hw * __ldg(s.gl_x + q) + hw;
        double z = (1.0 / av) - 1.0;
        double Rb = (3.0 / 4.0) * (c.obh2 / s.k.Ogamma_h2) * av;
        v[1] = __ldg(s.gl_w + q) * (DH_of_z<FAM, DE>(s, c, z) / (av * av * sqrt(3.0 * (1.0 + Rb))));
      }
    }
    const bool need_red = need_cmb || n_bao > 0 || n_cc > 0 || sn_small;  // uniform
    if (need_red && LEAN != 2) __syncthreads();  // sm.vec complete (LEAN = 2: behind the barrier of the row scale)

    if (mode == MODE_EVAL) {
      // sum_i d[i] M[i * stride_i + ...]: the matrix elements come from L2 (all of L1 is carved out as shared memory), so they
      // are fetched four at a time ahead of the multiply-adds; the additions keep their order (same bits as the plain loop)
      auto dot_strided = [&](const double* __restrict__ d, const double* __restrict__ m, int n, int stride) {
        double t = 0.0;
        int i = 0;
        for (; i + 4 <= n; i += 4) {
          const double m0 = __ldg(m + (i + 0) * stride), m1 = __ldg(m + (i + 1) * stride), m2 = __ldg(m + (i + 2) * stride),
                       m3 = __ldg(m + (i + 3) * stride);
          t += d[i] * m0; t += d[i + 1] * m1; t += d[i + 2] * m2; t += d[i + 3] * m3;
        }
        for (; i < n; i++) t += d[i] * __ldg(m + i * stride);
        return t;
      };
      if (pt < n_bao)   // delta @ inv_cov @ delta (bao/desi_cmb_union3.py:97-100)
        v[2] = dot_strided(sm.vec, s.bao_W + pt, n_bao, n_bao) * sm.vec[pt];
      if (pt < n_cc) {
        const double* d = sm.vec + CL_MAX_BAO;
        v[3] = dot_strided(d, s.cc_W + pt, n_cc, n_cc) * d[pt];
      }
      if (sn_small && (unsigned)st < (unsigned)n_sn) {
        const double* d = sm.vec + CL_MAX_BAO + CL_MAX_CC;
        if (s.sn_form == CL_SN_INVCOV) {  // delta @ inv_cov @ delta (sn/union3_1.py:57)
          v[4] = dot_strided(d, s.sn_mat_small + st, n_sn, n_sn) * d[st];
        } else {  // |L^-1 delta|^2 with W = L^-1 (solve_triangular.py:5-14)
          const double t = dot_strided(d, s.sn_mat_small + st * n_sn, st + 1, 1);
          v[4] = t * t;
        }
      }
    }
    const double rd_out = (need_rd && tid == 0) ? rdrag_from_terms(scal) : 0.0;
    // The theta row of the next iteration (loaded into a register at the top) is parked in the other buffer; no thread
    // reads that buffer in this iteration (the previous row's readers all passed this iteration's first barrier).
    if (stage_next && !row_synced) sm.theta[tb ^ 1][tid] = th_next;
    // The next iteration stores its grid nodes before its first barrier, so every thread must be done reading gd.
    const unsigned mask = (need_cmb ? 3u : 0u) | (n_bao > 0 ? 4u : 0u) | (n_cc > 0 ? 8u : 0u) | (sn_small ? 16u : 0u);
    if (need_red) block_sum<5>(v, sm.red[tb], mask, mode != MODE_EVAL);
    else if (!row_synced) __syncthreads();

    if (mode == MODE_EVAL) {
      // The raw block sums leave as they are, one aux plane per lane of warp 0: lanes 0-4 add the eight warp partials of one
      // value each (warp order: deterministic), lanes 5-7 carry the log-prior, the flags and z*.  k_finalize does the scalar
      // algebra on top (one thread per row there; here it was ~270 instructions with divisions, a square root and a
      // logarithm on ONE thread - and every thread of the CTA adding all partials of all values - while the CTA's other
      // warps waited at the next row's grid-pass barrier).
      if (tid < 8) {
        double val;
        if (tid < 5) val = (need_red && ((mask >> tid) & 1u)) ? block_partials_sum(sm.red[tb], tid) : 0.0;
        else val = tid == 5 ? lp : tid == 6 ? 0.0 : zstar;
        // value order of v[]: GL D_M sum, GL r_s sum, BAO, chronometers, small SN block
        const int slot = tid == 0 ? AUX_GL_DM : tid == 1 ? AUX_GL_RS : tid == 2 ? AUX_BAO : tid == 3 ? AUX_CC : tid == 4 ? AUX_SN_SMALL
                       : tid == 5 ? AUX_LOGPRIOR : tid == 6 ? AUX_FLAGS : AUX_ZSTAR;
        a.aux[slot * a.B + b] = val;
      }
    } else if (tid == 0) {   // MODE_CMB helper (block_sum has left the sums in v)
      double cmbv[3] = {0.0, 0.0, 0.0}, rs = 0.0, dm = 0.0;
      cmb_vector(s, c, zstar, v[0], v[1], cmbv, rs, dm);
      double* r = a.out + b * 8;
      r[0] = cmbv[0]; r[1] = cmbv[1]; r[2] = cmbv[2]; r[3] = zstar; r[4] = rs; r[5] = dm;
      r[6] = rd_out; r[7] = 100 * (rs / dm);
    }
    // (the barrier inside block_sum orders this iteration's shared-memory reads before the next row's writes)
  }
}
#undef mode
#undef n_bao
#undef n_cc
#undef cmb_mode
#undef sn_small

// ---- finalize: combine the SN chi2 partials of stage 3 with the scalar terms ----
// Accuracy guard of the int8 digit-plane engine ("error bound"): with S planes every residual row and
// every row of W carries FRAC = 8 S - 2 fractional bits below its power-of-two scale, and the products with i + j >= S are
// dropped, so every term of y_n = sum_k W_nk r_k is off by at most  c_n 2^eR_b,  c_n = 2^eW_n eps_S,
// eps_S = 2^(2 - 8 S) (1 + (S - 1) 256 / 255).  Two a-priori bounds on |d chi2_b| = |2 y.dy + dy.dy| follow:
//   worst case (every error at its maximum, all of one sign):  2 sqrt(chi2_b) rho_b + rho_b^2,  rho_b = 2^eR_b kappa_wc,
//       kappa_wc = eps_S sqrt(sum_n (nnz_n 2^eW_n)^2);
//   probabilistic (the term errors are roundings of balanced digits: bounded, mean zero, independent; Hoeffding over the
//       N (N + 1) / 2 terms of y.dy):  2 lambda sqrt(chi2_b) 2^eR_b kappa_pr + rho_b^2,  kappa_pr = eps_S max_n sqrt(nnz_n) 2^eW_n,
//       exceeded with probability < 2 exp(-lambda^2 / 2) per row (lambda = 8: 2.5e-14).
// `kappa` below is the coefficient of the linear term of the selected bound (cl_set_option "chi2_guard_mode": 0 =
// probabilistic, the default; 1 = worst case), `kappa_sq` the worst-case one of the quadratic term (both static: formed by
// cl_create).  Rows whose bound exceeds max(tol_abs, tol_rel chi2) are flagged and recomputed on the FP64 tensor pipe by the
// fallback pass (k_chi2_gemm restricted to the flagged 128-row blocks).
struct GuardArgs {
  const double* rowscale;   // nullptr = guard off
  double kappa, kappa_sq, tol_abs, tol_rel;
  int* guard;               // [0] rows flagged since cl_create, [1] rows flagged in this pass, [2 + rb] row-block marks
  unsigned char* rowflag;   // [B] 1 = this row is recomputed by the fallback pass
  int only_flagged;         // fallback pass: rewrite the flagged rows only (and do not flag again)
};
__device__ __forceinline__ bool guard_row(const GuardArgs& q, int64_t b, double chi2_sn) {
  const double rho = q.rowscale[b] * q.kappa, rho_sq = q.rowscale[b] * q.kappa_sq;
  const double bound = fma(2.0 * sqrt(fmax(chi2_sn, 0.0)), rho, rho_sq * rho_sq);
  const bool flag = bound > fmax(q.tol_abs, q.tol_rel * chi2_sn);   // NaN rows (bad residuals) compare false: chi2 stays NaN
  q.rowflag[b] = flag ? 1 : 0;
  if (flag) { q.guard[2 + (b >> 7)] = 1; atomicAdd(q.guard + 1, 1); atomicAdd(q.guard, 1); }
  return flag;
}

struct FinalizeArgs {
  int64_t B;
  int what, n_part, sn_large;
  const double* part;  // [n_part][B] per-column-tile partial sums of |W r|^2
  const double* aux;   // [AUX_COUNT][B]
  const double* theta; // [B][ld] the pass's parameter vectors (the scalar terms need a few of their columns)
  int64_t ld;
  double* out;         // [B]
  double* comps;       // nullable [B][4]
  double guard_value;
  GuardArgs q;
};

__global__ void __launch_bounds__(256) k_finalize(const __grid_constant__ DevSpec s, const __grid_constant__ FinalizeArgs f) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= f.B) return;
  if (f.q.only_flagged && (f.q.guard[1] == 0 || !f.q.rowflag[b])) return;
  int flags = (int)f.aux[AUX_FLAGS * f.B + b];
  double lp = f.aux[AUX_LOGPRIOR * f.B + b];
  if (flags) {
    double r = (flags & FLAG_OUTSIDE) ? -INFINITY : (f.what == CL_OUT_LOGPROB ? lp + f.guard_value : f.guard_value);
    if (f.out) f.out[b] = r;
    if (f.comps) for (int j = 0; j < 4; j++) f.comps[b * 4 + j] = NAN;
    if (f.q.rowscale && !f.q.only_flagged) f.q.rowflag[b] = 0;
    return;
  }
  double sn = 0.0;
  if (f.sn_large) for (int t = 0; t < f.n_part; t++) sn += f.part[(int64_t)t * f.B + b];
  else sn = f.aux[AUX_SN_SMALL * f.B + b];
  if (f.q.rowscale && !f.q.only_flagged) guard_row(f.q, b, sn);
  // scalar algebra on the raw block sums of stage 1+2 (bao/desi_cmb_union3.py:103-135, ohd/cc.py:29-33, ohd/cc_pantheon.py:63)
  const double* __restrict__ th = f.theta + b * f.ld;
  const double bao = f.aux[AUX_BAO * f.B + b];
  double cmb = 0.0, extra = 0.0, ccnorm = 0.0;
  if (s.cmb_mode != CL_CMB_NONE) {
    Cosmo c;
    unpack(s, th, c);
    double cmbv[3], rs, dm;
    cmb_vector(s, c, f.aux[AUX_ZSTAR * f.B + b], f.aux[AUX_GL_DM * f.B + b], f.aux[AUX_GL_RS * f.B + b], cmbv, rs, dm);
    const double d[3] = {s.cmb_prior[0] - cmbv[0], s.cmb_prior[1] - cmbv[1], s.cmb_prior[2] - cmbv[2]};
    for (int j = 0; j < 3; j++) {
      double t = 0.0;
      for (int i = 0; i < 3; i++) t += d[i] * s.cmb_W[i * 3 + j];
      cmb += t * d[j];
    }
  }
  if (s.n_cc > 0) {
    const double fcc = s.col_fcc >= 0 ? th[s.col_fcc] : 1.0;
    extra += (s.cc_norm_sign < 0.0 ? 1.0 / (fcc * fcc) : fcc * fcc) * f.aux[AUX_CC * f.B + b];   // error-inflation form: chi2 * f ** -2 (ohd/cc_pantheon.py:63)
    if (s.cc_norm_sign != 0.0) ccnorm = s.n_cc * log(2 * M_PI) + s.cc_logdet - s.cc_norm_sign * 2 * s.n_cc * log(fcc);
  }
  for (int g = 0; g < s.n_gc; g++) {
    const double r = (th[s.gc_col[g]] - s.gc_mean[g]) / s.gc_sigma[g];
    extra += r * r;
  }
  if (f.comps) { f.comps[b * 4] = sn; f.comps[b * 4 + 1] = bao; f.comps[b * 4 + 2] = cmb; f.comps[b * 4 + 3] = extra; }
  if (!f.out) return;
  double chi2 = sn + bao + cmb + extra;
  if (f.what == CL_OUT_CHI2) f.out[b] = chi2;
  else {
    double ll = -0.5 * (chi2 + ccnorm);
    f.out[b] = f.what == CL_OUT_LOGLIKE ? ll : lp + ll;
  }
}

// moments mode: out[b] = (y.y, y.u, u.u)
__global__ void __launch_bounds__(256) k_sum_parts(const double* __restrict__ part, const double* __restrict__ part_u, int T,
                                                   int64_t B, double uu, double* __restrict__ out, const GuardArgs q) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (q.only_flagged && (q.guard[1] == 0 || !q.rowflag[b])) return;
  double yy = 0.0, yu = 0.0;
  for (int t = 0; t < T; t++) { yy += part[(int64_t)t * B + b]; yu += part_u[(int64_t)t * B + b]; }
  if (q.rowscale && !q.only_flagged) guard_row(q, b, yy);
  out[b * 3] = yy; out[b * 3 + 1] = yu; out[b * 3 + 2] = uu;
}

}  // namespace cosmolike
