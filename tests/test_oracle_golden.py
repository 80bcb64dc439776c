"""Pins the CPU oracle (oracle/cosmo_oracle.c) to golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Tolerances: distances 1e-12 relative (north star asks 1e-9), chi2 1e-7 absolute
(north star asks 1e-6) even where chi2 ~ 1e6."""
import numpy as np
import pytest

import oracle.oracle as O
from cases import CHI2_CASES, golden, rel_err, spec

CHI2_ATOL = 1e-7
DIST_RTOL = 1e-12


@pytest.mark.parametrize("name", CHI2_CASES)
def test_chi2_matches_reference(name):
    g = golden(name)
    got = O.Oracle(spec(name)).chi_squared(g["theta"])
    assert np.max(np.abs(got - g["chi2"])) < CHI2_ATOL


def test_grid_is_the_reference_grid():
    for name in ("sn_pantheon", "sn_union3_1", "bao_desi", "bao_desi_cmb_union3", "bao_desi_des5y_bbn_theta_star"):
        assert np.array_equal(spec(name).z_grid, golden(name)["z_grid"]), name


def test_pantheon_distances_residuals_logprob():
    g = golden("sn_pantheon")
    o = O.Oracle(spec("sn_pantheon"))
    dm, _ = o.distances(g["theta"][:8], g["zq"])
    assert rel_err(dm, g["dm"]) < DIST_RTOL
    assert np.max(np.abs(o.sn_residuals(g["theta"][:4]) - g["delta"])) < 1e-12
    lp = o.log_probability(g["theta_logp"])
    assert np.array_equal(np.isneginf(lp), np.isneginf(g["logp"]))
    fin = np.isfinite(g["logp"])
    assert np.max(np.abs(lp[fin] - g["logp"][fin])) < CHI2_ATOL


def test_union3_distances_and_known_answers():
    g = golden("sn_union3_1")
    o = O.Oracle(spec("sn_union3_1"))
    assert rel_err(o.distances(g["theta"][:8], g["zq"])[0], g["dm"]) < DIST_RTOL
    # SURVEY.md G2 pins
    chi2 = o.chi_squared([[-0.05, 0.3, -3.0], [-0.05, 0.3, 0.0], [0.027, 0.335, 0.0]])
    assert np.allclose(chi2, [160.89519602172209, 178.67819338900432, 28.761102862275692], rtol=0, atol=1e-9)


def test_bao_desi_theory_and_float32_batch():
    g = golden("bao_desi")
    o = O.Oracle(spec("bao_desi"))
    assert rel_err(o.bao_theory(g["theta"]), g["theory"]) < DIST_RTOL
    lp = o.log_probability(g["batch"])
    assert np.array_equal(lp.astype(np.float32), g["logp32"])  # bao/desi.py:103 returns float32
    assert np.array_equal(np.isneginf(lp), np.isneginf(g["logp64"]))


def test_config2_components_and_cmb():
    g = golden("bao_desi_cmb_union3")
    o = O.Oracle(spec("bao_desi_cmb_union3"))
    c = o.components(g["theta"])
    assert np.max(np.abs(c[:, 0] - g["chi2_sn"])) < CHI2_ATOL
    assert np.max(np.abs(c[:, 1] - g["chi2_bao"])) < CHI2_ATOL
    assert np.max(np.abs(c[:, 2] - g["chi2_cmb"])) < CHI2_ATOL
    cm = o.cmb(g["theta"])
    assert rel_err(cm[:, :3], g["cmb_distances"]) < DIST_RTOL
    assert rel_err(cm[:, 3], g["z_star"]) < DIST_RTOL
    assert rel_err(cm[:, 6], g["r_drag"]) < DIST_RTOL
    assert rel_err(o.bao_theory(g["theta"]), g["bao_theory"]) < DIST_RTOL
    # SURVEY.md G4 pin
    assert abs(o.chi_squared([-0.0519, 68.42, 0.02257, 0.11738, -3.0])[0] - 39.686971513276205) < 1e-9


def test_cmb_constants_match_reference_module():
    g = golden("bao_desi_cmb_union3")
    k = spec("bao_desi_cmb_union3").cmb_consts
    assert k.Or_h2 == float(g["Or_h2"]) and k.Omnu_h2 == float(g["Omnu_h2"]) and k.Ogamma_h2 == float(g["O_GAMMA_H2"])
    assert k.nu_m0 == float(g["m0"]) and k.nu_rho0 == float(g["rho0"])
    assert np.array_equal(k.nu_q, g["qs"]) and np.array_equal(np.array(k.nu_w), g["ws"])
    assert np.array_equal(k.priors, g["cmb_priors"]) and np.array_equal(k.covariance, g["cmb_cov"])
    x, w = np.polynomial.legendre.leggauss(100)
    assert np.array_equal(x, g["GL_X"]) and np.array_equal(w, g["GL_W"])


def test_cpl_guard_and_loglike():
    g = golden("bao_desi_fs_lya_cmb")
    o = O.Oracle(spec("bao_desi_fs_lya_cmb"))
    ll = o.log_likelihood(g["theta"])
    guard = g["loglike"] == -1e8
    assert guard.sum() > 0 and np.array_equal(ll == -1e8, guard)
    assert np.max(np.abs(ll - g["loglike"])) < CHI2_ATOL
    ok = ~guard
    c = o.components(g["theta"][ok])
    assert np.max(np.abs(c[:, 1] - g["chi2_bao"][ok])) < CHI2_ATOL
    assert np.max(np.abs(c[:, 2] - g["chi2_cmb"][ok])) < CHI2_ATOL


def test_cmb_only_blobs():
    g = golden("cmb_cmb")
    o = O.Oracle(spec("cmb_cmb"))
    assert np.max(np.abs(o.log_likelihood(g["theta"]) - g["loglike"])) < CHI2_ATOL
    cm = o.cmb(g["theta"])
    blobs = np.c_[cm[:, 7], cm[:, 4], cm[:, 5] / 1000, cm[:, 3]]  # cmb/cmb.py:62-63
    assert rel_err(blobs, g["blobs"]) < DIST_RTOL
    lp = o.log_probability(g["theta_logp"])
    assert np.array_equal(np.isneginf(lp), np.isneginf(g["logp"]))


@pytest.mark.parametrize("name", ["cmb_cmb_act", "cmb_cmb_planck_lens", "cmb_cmb_planck"])
def test_cmb_constant_sets_behind_cmb_cmb(name):
    """cmb/data_act_compression.py, data_planck_lens_compression.py, data_planck_compression.py swapped into cmb/cmb.py's import
    (tests/golden/make_golden.py::_cmb_cmb_with_module): pins spec.cmb_act / cmb_planck_lens / cmb_planck."""
    g = golden(name)
    sp = spec(name)
    assert np.array_equal(np.asarray(sp.cmb_consts.priors), g["priors"])
    o = O.Oracle(sp)
    assert np.max(np.abs(o.log_likelihood(g["theta"]) - g["loglike"])) < CHI2_ATOL
    cm = o.cmb(g["theta"])
    blobs = np.c_[cm[:, 7], cm[:, 4], cm[:, 5] / 1000, cm[:, 3]]
    assert rel_err(blobs, g["blobs"]) < DIST_RTOL


def test_config1_logprob_with_bbn_prior():
    g = golden("bao_desi_des5y_bbn_theta_star")
    o = O.Oracle(spec("bao_desi_des5y_bbn_theta_star"))
    lp = o.log_probability(g["theta_logp"])
    fin = np.isfinite(g["logp"])
    assert np.array_equal(np.isneginf(lp), ~fin)
    assert np.max(np.abs(lp[fin] - g["logp"][fin])) < CHI2_ATOL
    assert rel_err(o.bao_theory(g["theta"][:8]), g["bao_theory"]) < DIST_RTOL


def test_interpolator_known_answers():
    g = golden("interpolator")
    assert np.array_equal(O.interp_hermite(g["u_xq"], g["u_x"], g["u_y"], g["u_yp"]), g["u_herm"])
    assert np.array_equal(O.interp_pchip(g["u_xq"], g["u_x"], g["u_y"]), g["u_pchip"])
    assert np.array_equal(O.pchip_slopes(g["u_x"], g["u_y"]), g["u_slopes"])
    assert np.array_equal(O.pchip_slopes(g["n_x"], g["n_y"]), g["n_slopes"])
    assert np.array_equal(O.interp_pchip(g["n_xq"], g["n_x"], g["n_y"]), g["n_pchip"])
    assert np.array_equal(O.interp_pchip(g["u_xq"], g["u_x"], g["d_y"]), g["d_pchip"])


def test_solve_triangular_known_answers():
    g = golden("solve_triangular")
    for b, v in zip(g["b"], g["value"]):
        assert abs(O.solve_triangular(g["L"], b) - v) < 1e-12 * abs(v)


def test_threads_do_not_change_results():
    g = golden("sn_pantheon")
    o = O.Oracle(spec("sn_pantheon"))
    assert np.array_equal(o.chi_squared(g["theta"], nthreads=1), o.chi_squared(g["theta"], nthreads=4))


from cases import GENERIC_LOGLIKE_CASES, GENERIC_LOGP_CASES  # noqa: E402


@pytest.mark.parametrize("name", GENERIC_LOGLIKE_CASES)
def test_generic_log_likelihood(name):
    """log_likelihood incl. the cosmic-chronometer normalisation term in both sign conventions (ohd/cc.py:33,
    ohd/cc_pantheon.py:92)."""
    import oracle.oracle as O
    g = golden(name)
    assert np.max(np.abs(O.Oracle(spec(name)).log_likelihood(g["theta"]) - g["loglike"])) < CHI2_ATOL


@pytest.mark.parametrize("name", GENERIC_LOGP_CASES)
def test_generic_log_probability(name):
    """The script's own log_probability: box prior (-inf rows are never evaluated) + its normalisation constant."""
    import oracle.oracle as O
    g = golden(name)
    lp = O.Oracle(spec(name)).log_probability(g["theta_logp"])
    assert np.array_equal(np.isneginf(lp), np.isneginf(g["logp"])) and np.isneginf(g["logp"]).sum() == 2
    fin = np.isfinite(g["logp"])
    assert np.max(np.abs(lp[fin] - g["logp"][fin])) < CHI2_ATOL


def test_pchip_on_equal_intervals_in_units_of_the_node_spacing():
    """The CUDA path evaluates interp_pchip (interpolator.py:5-68,111-114) on the np.linspace grid in units of the node
    spacing (csrc/friedmann.cuh: pchip_dh): slope x h = harmonic mean of the neighbouring differences, (3 d0 - d1) / 2 with the
    end rules at the two boundary nodes - two divisions per query instead of eleven.  Restated here in numpy and held to the
    oracle's literal transcription of the reference: monotone data (the dh grid of a real fit is), data with sign changes and
    flat runs, queries in the first and last interval and on nodes."""
    import oracle.oracle as O
    rng = np.random.default_rng(5)

    def fast(xq, x, y):
        n, step = len(x), x[1] - x[0]
        out = np.empty_like(xq)
        sgn = np.sign
        def end_slope(d0, d1):
            v = 0.5 * (3.0 * d0 - d1)
            if d0 == 0.0 or sgn(v) != sgn(d0):
                return 0.0
            if sgn(d0) != sgn(d1) and abs(v) > abs(3.0 * d0):
                return 3.0 * d0
            return v
        def mid_slope(dm, dp):
            return 2.0 * dm * dp / (dm + dp) if (dm != 0.0 and dp != 0.0 and dm * dp > 0.0) else 0.0
        for k, q in enumerate(xq):
            if q <= x[0]:
                out[k] = y[0]; continue
            if q >= x[-1]:
                out[k] = y[-1]; continue
            i = int(np.searchsorted(x, q) - 1)
            t = (q - x[i]) / step
            dc = y[i + 1] - y[i]
            m0 = end_slope(dc, y[2] - y[1]) if i == 0 else mid_slope(y[i] - y[i - 1], dc)
            m1 = end_slope(dc, y[n - 2] - y[n - 3]) if i == n - 2 else mid_slope(dc, y[i + 2] - y[i + 1])
            t2, t3 = t * t, t * t * t
            out[k] = (2 * t3 - 3 * t2 + 1) * y[i] + (t3 - 2 * t2 + t) * m0 + (-2 * t3 + 3 * t2) * y[i + 1] + (t3 - t2) * m1
        return out

    x = np.linspace(0.0, 2.4, 401)
    cases = [1.0 / np.sqrt(0.3 * (1 + x) ** 3 + 0.7),                 # dh-like: smooth, decreasing
             np.cumsum(rng.uniform(0.0, 1.0, x.size)),                # increasing, rough
             rng.standard_normal(x.size),                             # sign changes everywhere
             np.repeat(rng.standard_normal(x.size // 4 + 1), 4)[:x.size]]   # flat runs
    xq = np.concatenate([rng.uniform(-0.1, 2.5, 500), x[:3] + 1e-4, x[-3:] - 1e-4, x[5:8], [x[0], x[-1]]])
    for y in cases:
        want = O.interp_pchip(xq, x, y)
        got = fast(xq, x, np.asarray(y, dtype=np.float64))
        scale = np.max(np.abs(y))
        assert np.max(np.abs(got - want)) <= 1e-13 * scale, np.max(np.abs(got - want)) / scale
