"""Adversarial-conditioning parity of stage 3 (chi2_sn = |L^-1 delta|^2, solve_triangular.py:5-14 / sn/pantheon.py:14,61).

Every large-N golden case uses the benign synthetic covariance of `synthetic.py` (cond ~ 2e2).  The default stage-3 engine
is NOT FP64 arithmetic (int8 digit planes, one power-of-two scale per row, products below the last digit dropped), and both
engines contract with W = fl(L^-1) instead of substituting forward, so robustness is shown here on
inputs built to hurt:

  * covariances with cond 1e6 / 1e8 / 1e10 of two kinds - strongly correlated calibration-like modes (C = D + sum a a^T
    with huge amplitudes) and a diagonal spanning many decades;
  * a W = L^-1 whose rows span > 1e8 in magnitude;
  * residual rows with one 1e3 x outlier supernova;
  * N = 1701 (the headline shape, all Pantheon+ rows) and N = 1820 (DES-Dovekie).

Yardstick: forward substitution in 80-bit extended precision (`oracle.solve_triangular_ld`) on the ENGINE'S OWN residual rows
(`cl_sn_residuals`), which isolates stage 3: at cond 1e10 a 1-ulp difference in a residual moves chi2 by more than 1e-6, for
the reference as for anybody.  Bar for engines 0 (FP64 DMMA), 7 planes (default, guard on): |d chi2| <= max(1e-6, 1e-12 chi2)
or - where FP64 itself cannot do better - 4 x the error of the reference's own FP64 forward substitution against the same
yardstick.  6 planes run with the guard at its 6-plane tolerance.  The guard must fire where its bound says so and the
guarded values must then be the FP64 engine's bits."""
import numpy as np
import pytest

from cases import golden
from cosmology_model_fit_b200 import datasets, fits
from cosmology_model_fit_b200.synthetic import synthetic_sn_covariance, uniform_theta

pytestmark = pytest.mark.gpu


def _pantheon(cut):
    return datasets.pantheon_plus(cut=cut)


def _cov_modes(sigma, cond, seed):
    """diag(sigma^2) + 3 calibration-like modes scaled so that cond(C) ~ cond."""
    n = sigma.size
    rng = np.random.default_rng(seed)
    C = np.diag(sigma**2)
    amp = np.sqrt(cond * np.min(sigma**2) / n)
    for k in range(3):
        a = amp * (1.0 + 0.3 * rng.standard_normal(n)) / (k + 1)
        C = C + np.outer(a, a)
    return C


def _cov_diag_range(sigma, cond, seed):
    """rank-40 systematics + a diagonal whose entries span sqrt(cond) in sigma (a few very precise, a few very poor SNe)."""
    n = sigma.size
    rng = np.random.default_rng(seed)
    s = sigma * np.exp(rng.uniform(-0.25, 0.25, n) * np.log(cond))
    return synthetic_sn_covariance(s, seed=seed, amp=0.02 * float(np.min(s)) / float(np.min(sigma)))


def _case(kind, cond, n_sn):
    if n_sn == 1820:
        z_cmb, z_hel, obs, cov0 = datasets.des_dovekie()
        builder = fits.sn_des5y
    else:
        z_cmb, z_hel, obs, cov0 = _pantheon(cut=(n_sn == 1590))
        builder = fits.sn_pantheon
    assert z_cmb.size == n_sn
    sigma = np.sqrt(np.diag(cov0))
    sigma = np.minimum(sigma, 0.5)   # the synthetic diagonal has a 1.4 mag tail; keep the base matrix tame
    obs = obs.copy()
    if kind == "modes":
        cov = _cov_modes(sigma, cond, seed=n_sn)
    elif kind == "diag":
        cov = _cov_diag_range(sigma, cond, seed=n_sn + 1)
    elif kind == "outlier":
        cov = synthetic_sn_covariance(sigma, seed=7)
        obs[n_sn // 3] += 1000.0 * sigma[n_sn // 3] * 10.0   # one supernova ~1e3 x off
    else:
        raise ValueError(kind)
    sp = builder((z_cmb, z_hel, obs, cov))
    if n_sn == 1820:   # sn/des5y.py is a nautilus script: its prior box lives in the golden file, not in the spec
        sp.theta_box = golden("sn_des5y")["bounds"]
    else:
        sp.theta_box = sp.bounds
    return sp, cov


CASES = [("modes", 1e6, 1701), ("modes", 1e8, 1701), ("modes", 1e10, 1701), ("diag", 1e6, 1701), ("diag", 1e8, 1701),
         ("diag", 1e10, 1701), ("outlier", 0.0, 1701), ("modes", 1e8, 1820), ("diag", 1e10, 1820), ("outlier", 0.0, 1820)]


@pytest.fixture(scope="module")
def prepared():
    """spec, theta, the engine's own residual rows, the extended-precision yardstick and the reference-FP64 error per case."""
    import oracle.oracle as O
    from cosmology_model_fit_b200 import Engine
    cache = {}

    def get(kind, cond, n_sn):
        key = (kind, cond, n_sn)
        if key not in cache:
            sp, cov = _case(kind, cond, n_sn)
            theta = uniform_theta(sp.theta_box, 160, seed=11)
            with Engine(sp) as e:
                R = e.sn_residuals(theta)
            L = np.ascontiguousarray(sp.sn_mat)
            truth = O.solve_triangular_ld(L, R)
            ref64 = O.solve_triangular_batch(L, R)
            cache[key] = dict(spec=sp, theta=theta, R=R, truth=truth, ref_err=np.abs(ref64 - truth), cond=np.linalg.cond(cov))
        return cache[key]
    return get


def _tol(p):
    return np.maximum(np.maximum(1e-6, 1e-12 * np.abs(p["truth"])), 4.0 * p["ref_err"])


@pytest.mark.parametrize("kind,cond,n_sn", CASES)
def test_fp64_engine_vs_extended_precision(prepared, kind, cond, n_sn):
    from cosmology_model_fit_b200 import Engine
    p = prepared(kind, cond, n_sn)
    with Engine(p["spec"]) as e:
        e.set_option("chi2_engine", 0)
        got = e.components(p["theta"])[:, 0]
    err = np.abs(got - p["truth"])
    assert np.all(err <= _tol(p)), (kind, cond, n_sn, p["cond"], err.max(), p["ref_err"].max(), np.abs(p["truth"]).max())


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("kind,cond,n_sn", CASES)
def test_digit_plane_engine_guarded_vs_extended_precision(prepared, kind, cond, n_sn, mode):
    """Default engine (7 planes, guard on): every value meets the bar, and flagged rows carry the FP64 engine's bits - with the
    probabilistic bound (mode 0, the default) and with the worst-case bound (mode 1); either bound really bounds the error."""
    from cosmology_model_fit_b200 import Engine
    p = prepared(kind, cond, n_sn)
    with Engine(p["spec"]) as e:
        e.set_option("chi2_guard_mode", mode)
        got = e.components(p["theta"])[:, 0]
        info = e.guard_info()
        e.set_option("chi2_guard", 0)
        raw = e.components(p["theta"])[:, 0]
        e.set_option("chi2_engine", 0)
        fp64 = e.components(p["theta"])[:, 0]
    err = np.abs(got - p["truth"])
    assert np.all(err <= _tol(p)), (kind, cond, n_sn, err.max(), p["ref_err"].max(), info)
    # the a-priori bound really bounds the unguarded engine's error against the FP64 contraction of the same W
    # (+ the rounding of the FP64 contraction itself, which the bound does not cover)
    scale = np.ldexp(1.0, np.frexp(np.abs(p["R"]).max(axis=1))[1])
    rho = scale * info["kappa"]
    bound = 2 * np.sqrt(np.abs(fp64)) * rho + (rho if mode else 0.0)**2   # (mode 0: the quadratic term is the worst-case one, ~1e-17)
    assert np.all(np.abs(raw - fp64) <= bound + 4.0 * p["ref_err"] + 1e-13 * np.abs(fp64)), (np.abs(raw - fp64).max(), bound.max())
    flagged = bound > np.maximum(5e-7, 1e-12 * np.abs(raw))
    # rows the guard recomputed are bit-identical to the FP64 engine; the others to the unguarded planes
    sure = bound > 2 * np.maximum(5e-7, 1e-12 * np.abs(raw))      # away from the threshold (the kernel uses its own chi2)
    assert np.array_equal(got[sure], fp64[sure])
    calm = bound < 0.5 * np.maximum(5e-7, 1e-12 * np.abs(raw))
    assert np.array_equal(got[calm], raw[calm])
    assert info["rows_last_pass"] >= int(sure.sum()) and info["rows_last_pass"] <= int((~calm).sum())
    if kind == "outlier" and mode == 1:
        assert flagged.all() and info["rows_last_pass"] == p["theta"].shape[0]


@pytest.mark.parametrize("kind,cond,n_sn", [("modes", 1e8, 1701), ("diag", 1e10, 1820), ("outlier", 0.0, 1701)])
def test_six_planes_with_its_own_tolerance(prepared, kind, cond, n_sn):
    """6 planes (46 bits per row): the guard holds whatever tolerance the caller sets; here 1e-4 absolute / 1e-9 relative."""
    from cosmology_model_fit_b200 import Engine
    p = prepared(kind, cond, n_sn)
    with Engine(p["spec"]) as e:
        e.set_option("chi2_slices", 6)
        e.set_option("chi2_guard_abs", 1e-4)
        e.set_option("chi2_guard_rel", 1e-9)
        got = e.components(p["theta"])[:, 0]
    err = np.abs(got - p["truth"])
    assert np.all(err <= np.maximum(np.maximum(1e-4, 1e-9 * np.abs(p["truth"])), 4.0 * p["ref_err"])), err.max()


def test_headline_size_vs_oracle_and_no_fallback():
    """N = 1701 (the size named in the metric): 512 random rows against the CPU oracle at the north-star tolerance, the guard
    silent on the benign covariance, and every output selector."""
    import oracle.oracle as O
    from cosmology_model_fit_b200 import Engine
    sp = fits.sn_pantheon(_pantheon(cut=False))
    theta = uniform_theta(sp.bounds, 512, seed=3)
    orc = O.Oracle(sp)
    want = orc.chi_squared(theta, nthreads=0)
    with Engine(sp) as e:
        got = e.chi_squared(theta)
        info = e.guard_info()
        lp = e.log_probability(theta)
    assert np.all(np.abs(got - want) <= np.maximum(1e-6, 1e-12 * np.abs(want))), np.abs(got - want).max()
    assert info["rows_total"] == 0, info
    assert np.all(np.abs(lp - orc.log_probability(theta, nthreads=0)) <= np.maximum(1e-6, 1e-12 * np.abs(want)))


def test_guard_is_batch_independent():
    """A row's value does not depend on which other rows share its batch, also when some of them are flagged."""
    from cosmology_model_fit_b200 import Engine
    sp, _ = _case("modes", 1e8, 1701)
    theta = uniform_theta(sp.theta_box, 700, seed=5)
    with Engine(sp) as e:
        e.set_option("chi2_guard_mode", 1)     # worst-case bound
        e.set_option("chi2_guard_abs", 2e-8)   # a threshold inside this batch's range of bounds (1e-8 .. 3e-8): it splits the batch
        full = e.chi_squared(theta)
        n_flag = e.guard_info()["rows_last_pass"]
        part = np.concatenate([e.chi_squared(theta[:300]), e.chi_squared(theta[300:301]), e.chi_squared(theta[301:])])
    assert np.array_equal(full, part)
    assert 0 < n_flag < theta.shape[0], n_flag


@pytest.mark.parametrize("name", ["bao_desi_cmb_pantheon", "bao_desi_des5y_bbn_theta_star"])
def test_guard_fallback_of_the_small_probe_kernel(name):
    """Configurations whose stage 2 writes the digit planes from the small-probe instantiation (SN block + BAO / CMB terms):
    when the guard fires, the FP64 residual rows of the flagged blocks are regenerated by the FULL kernel and contracted on the
    FP64 tensor pipe - the values are then the FP64 engine's, bit for bit, for flagged rows, and the small terms are untouched."""
    from cases import golden, spec
    from cosmology_model_fit_b200 import Engine
    from cosmology_model_fit_b200.synthetic import uniform_theta
    theta = uniform_theta(golden(name)["bounds"], 1000, seed=3)
    with Engine(spec(name)) as e:
        base = e.chi_squared(theta)
        comp = e.components(theta)
        assert e.guard_info()["rows_total"] == 0
        e.set_option("chi2_guard_abs", 1e-30)
        e.set_option("chi2_guard_rel", 0.0)          # every finite row exceeds this tolerance: all rows take the fallback
        guarded = e.chi_squared(theta)
        comp_g = e.components(theta)
        n_flag = e.guard_info()["rows_last_pass"]
        e.set_option("chi2_engine", 0)
        fp64 = e.chi_squared(theta)
        comp_f = e.components(theta)
    fin = np.isfinite(fp64)
    assert n_flag == int(fin.sum()) and n_flag > 900
    assert np.array_equal(comp_g[:, 0], comp_f[:, 0], equal_nan=True)      # SN chi2: the FP64 engine's bits
    assert np.array_equal(guarded, fp64, equal_nan=True)
    assert np.allclose(comp_g[fin, 1:], comp[fin, 1:], rtol=1e-9, atol=1e-9)
    assert np.all(np.abs(base[fin] - fp64[fin]) <= np.maximum(1e-6, 1e-12 * np.abs(fp64[fin])))
