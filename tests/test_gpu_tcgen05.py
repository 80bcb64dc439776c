"""GPU parity tests of the tcgen05 chi-squared engine (csrc/chi2_ozaki.cuh): the SN contraction |L^-1 delta|^2 of
solve_triangular.py:5-14 evaluated with int8 digit planes on the 5th-generation tensor cores, against the golden
vectors of the unmodified reference, the CPU oracle and the FP64 DMMA engine.

Tolerances: 7 digit planes carry every bit of an FP64 row, so they must meet the same bar as the FP64 engine
(|d chi2| < 1e-6 absolute, relaxed to 1e-12 relative where chi2 itself exceeds 1e6).  6 planes carry 46 bits per
row: the bar is 1e-6 absolute below chi2 = 1e5 and 1e-11 relative above."""
import numpy as np
import pytest

from cases import golden, spec

pytestmark = pytest.mark.gpu

#: golden cases with a large SN block (stage 3 runs); the others never reach the contraction
LARGE_SN_CASES = ["sn_pantheon", "sn_des5y", "bao_desi_des5y_bbn_theta_star", "bao_desi_cmb_pantheon", "bao_desi_cmb_des5y",
                  "bao_desi_pantheon_cc", "sn_pantheon_dipole_xyz", "sn_pantheon_and_sh0es", "bao_desi_cmb_pantheon_H0trgb",
                  "sn_pantheon_dipole", "ohd_cc_des5y"]


def close(got, want, slices):
    got, want = np.asarray(got), np.asarray(want)
    tol = np.maximum(1e-6, (1e-12 if slices >= 7 else 1e-11) * np.abs(want))
    bad = np.abs(got - want) > tol
    assert not bad.any(), (got[bad][:5], want[bad][:5], np.abs(got - want)[bad][:5])


@pytest.fixture(scope="module")
def engines():
    from cosmology_model_fit_b200 import Engine
    from cosmology_model_fit_b200.engine import CHI2_ENGINE_TCGEN05
    cache = {}

    def get(name, slices):
        key = (name, slices)
        if key not in cache:
            e = Engine(spec(name))
            if slices:
                e.set_option("chi2_engine", CHI2_ENGINE_TCGEN05)
                e.set_option("chi2_slices", slices)
            else:
                e.set_option("chi2_engine", 0)
            cache[key] = e
        return cache[key]
    yield get
    for e in cache.values():
        e.close()


@pytest.mark.parametrize("slices", [6, 7])
@pytest.mark.parametrize("name", LARGE_SN_CASES)
def test_chi2_vs_reference_golden(engines, name, slices):
    g = golden(name)
    close(engines(name, slices).chi_squared(g["theta"]), g["chi2"], slices)


@pytest.mark.parametrize("slices", [6, 7])
@pytest.mark.parametrize("name", ["sn_pantheon", "sn_des5y", "bao_desi_cmb_pantheon"])
def test_chi2_vs_oracle_random_batch(engines, name, slices):
    import oracle.oracle as O
    from cosmology_model_fit_b200.synthetic import uniform_theta
    theta = uniform_theta(golden(name)["bounds"], 300, seed=7)
    close(engines(name, slices).chi_squared(theta), O.Oracle(spec(name)).chi_squared(theta, nthreads=0), slices)


def test_five_planes_documented_accuracy(engines):
    """38 bits per row: 1e-9 relative (not a parity mode; kept for coarse grid scans)."""
    g = golden("sn_pantheon")
    got = engines("sn_pantheon", 5).chi_squared(g["theta"])
    assert np.max(np.abs(got - g["chi2"]) / np.abs(g["chi2"])) < 5e-9


@pytest.mark.parametrize("slices", [6, 7])
def test_rows_are_independent_of_the_batch(engines, slices):
    """One exponent per residual row and exact integer accumulation: a row's value does not depend on the batch
    size, the row-block edges of the 128-row tile, the order of the rows or the number of passes."""
    from cosmology_model_fit_b200 import Engine
    from cosmology_model_fit_b200.synthetic import uniform_theta
    e = engines("sn_pantheon", slices)
    theta = uniform_theta(golden("sn_pantheon")["bounds"], 1000, seed=3)
    full = e.chi_squared(theta)
    for b in (1, 127, 128, 129):
        assert np.array_equal(e.chi_squared(theta[:b]), full[:b])
    perm = np.random.default_rng(0).permutation(1000)
    assert np.array_equal(e.chi_squared(theta[perm]), full[perm])
    with Engine(spec("sn_pantheon")) as e2:
        e2.set_option("chi2_engine", 1)
        e2.set_option("chi2_slices", slices)
        e2.set_option("max_rows_per_pass", 256)
        assert np.array_equal(e2.chi_squared(theta), full)
        e2.set_option("gemm_group_rb", 8)       # another L2 group order of the items: same integers
        assert np.array_equal(e2.chi_squared(theta), full)
        e2.set_option("gemm_ctas", 3)           # three persistent CTAs walk all items
        assert np.array_equal(e2.chi_squared(theta), full)


def test_prior_and_guard_rows(engines):
    """-inf rows (outside the prior box) are never evaluated: their stale residual rows still go through the slicing
    kernel and must not disturb anything (sn/pantheon.py:80-97)."""
    g = golden("sn_pantheon")
    lp = engines("sn_pantheon", 6).log_probability(g["theta_logp"])
    want = g["logp"]
    fin = np.isfinite(want)
    assert np.array_equal(np.isneginf(lp), np.isneginf(want)) and not np.isnan(lp).any()
    close(-2 * lp[fin], -2 * want[fin], 6)


@pytest.mark.parametrize("slices", [6, 7])
def test_full_size_batch_against_fp64_engine(engines, slices):
    """BASELINE.json's batch (65536 rows, Pantheon+ N = 1590 as fitted): every row against the FP64 DMMA engine,
    plus the size-independent property that chi2 is exactly quadratic in the magnitude offset."""
    from cosmology_model_fit_b200.synthetic import uniform_theta
    B = 65536
    theta = uniform_theta(golden("sn_pantheon")["bounds"], B, seed=42)
    ref = engines("sn_pantheon", 0).chi_squared(theta)
    e = engines("sn_pantheon", slices)
    c0 = e.chi_squared(theta)
    close(c0, ref, slices)
    d = 0.01
    tp, tm = theta.copy(), theta.copy()
    tp[:, 0] += d; tm[:, 0] -= d
    second = (e.chi_squared(tp) - 2 * c0 + e.chi_squared(tm)) / d**2
    uu = e.sn_moments(theta[:1])[0, 2]
    assert np.max(np.abs(second - 2 * uu) / (2 * uu)) < 1e-6


def test_components_and_moments_with_the_tcgen05_engine(engines):
    """components() and sn_moments() (y.y, y.u, u.u: the two-dot epilogue of SURVEY N3) through the sliced contraction."""
    g = golden("bao_desi_cmb_pantheon")
    e, ref = engines("bao_desi_cmb_pantheon", 7), engines("bao_desi_cmb_pantheon", 0)
    a, b = e.components(g["theta"]), ref.components(g["theta"])
    assert np.max(np.abs(a - b)) < 1e-6
    th = golden("sn_pantheon")["theta"]
    m, m0 = engines("sn_pantheon", 7).sn_moments(th), engines("sn_pantheon", 0).sn_moments(th)
    assert np.max(np.abs(m - m0) / np.abs(m0)) < 1e-12
    m6 = engines("sn_pantheon", 6).sn_moments(th)
    assert np.max(np.abs(m6 - m0) / np.abs(m0)) < 1e-10
    # chi2(M) = yy - 2 M yu + M^2 uu is what chi_squared() returns (cancellation: M ~ -19.5 enters squared)
    chi2 = engines("sn_pantheon", 7).chi_squared(th)
    M = th[:, 0]
    direct = m[:, 0] - 2 * M * m[:, 1] + M * M * m[:, 2]
    assert np.max(np.abs(chi2 - direct) / np.abs(direct)) < 1e-9


def test_nan_residuals_propagate(engines):
    """A peculiar-velocity amplitude far outside the prior makes z_cosmo negative for the nearest SNe: log10 of a negative
    distance is NaN in the reference (sn/pantheon.py:43-54) and chi2 must be NaN with either engine, not a finite number
    assembled from clamped digits; the other rows of the batch are unaffected."""
    import oracle.oracle as O
    g = golden("sn_pantheon")
    theta = g["theta"][:8].copy()
    theta[3, 3] = 60.0      # 6000 km/s: z_pec = 0.02 > min(z_cmb) = 0.01
    want = O.Oracle(spec("sn_pantheon")).chi_squared(theta)
    assert np.isnan(want[3]) and np.isfinite(np.delete(want, 3)).all()
    for slices in (0, 6, 7):
        got = engines("sn_pantheon", slices).chi_squared(theta)
        assert np.isnan(got[3]), slices
        close(np.delete(got, 3), np.delete(want, 3), slices or 7)


def test_stage3_split_timing(engines):
    """cl_stage3_split: planes + contraction add up to the stage-3 time of cl_last_timing; zero planes time with the DMMA
    engine."""
    from cosmology_model_fit_b200.synthetic import uniform_theta
    theta = uniform_theta(golden("sn_pantheon")["bounds"], 4096, seed=1)
    for slices in (7, 0):
        e = engines("sn_pantheon", slices)
        for fuse in ((1, 0) if slices else (1,)):
            e.set_option("fuse_planes", fuse)
            e.chi_squared(theta)
            planes, contraction = e.stage3_split()
            t = e.last_timing()
            assert contraction > 0 and abs(planes + contraction - t["stage3_ms"]) < 1e-3
            # a separate slicing kernel only runs for the tcgen05 engine when stage 2 does not write the planes itself
            assert (planes > 0.005) == (bool(slices) and not fuse)
        e.set_option("fuse_planes", 0)
    hist = engines("sn_pantheon", 7).stage3_split(3)
    assert hist.shape[1] == 2 and len(hist) >= 1


@pytest.mark.parametrize("n", [65, 100, 129, 191, 640, 1000])
def test_ragged_sn_counts_vs_oracle(n):
    """SN blocks whose size is no multiple of anything in the kernel (column tile NT = 64 / 80, k block 64, TMA box 128
    rows): the first column tile starts at a negative column (TMA zero fill), the plane pitch is padded, the last k block is
    partial.  Pantheon+ rows 0..n-1 with the matching corner of the covariance, all three engines against the CPU oracle."""
    import oracle.oracle as O
    from cosmology_model_fit_b200 import Engine, datasets, fits
    from cosmology_model_fit_b200.synthetic import uniform_theta
    z, zh, mb, cov = datasets.pantheon_plus(cut=False)
    sub = fits.sn_pantheon((z[:n].copy(), zh[:n].copy(), mb[:n].copy(), np.ascontiguousarray(cov[:n, :n])))
    theta = uniform_theta(sub.bounds, 257, seed=n)
    want = O.Oracle(sub).chi_squared(theta, nthreads=0)
    with Engine(sub) as e:
        for slices in (0, 6, 7):
            e.set_option("chi2_engine", 1 if slices else 0)
            if slices:
                e.set_option("chi2_slices", slices)
            close(e.chi_squared(theta), want, slices or 7)


@pytest.mark.parametrize("slices", [5, 6, 7])
@pytest.mark.parametrize("name", ["sn_pantheon", "sn_des5y"])
def test_fused_digit_planes_are_the_slicing_kernels_bits(name, slices):
    """Stage 2 of the lean kernel can write the digit planes itself (`fuse_planes`, opt-in): same scale, same digits as the
    separate slicing kernel, hence the same chi2 bits - random rows, prior rows, a NaN row, the SN moments, ragged batches."""
    from cosmology_model_fit_b200 import Engine
    from cosmology_model_fit_b200.synthetic import uniform_theta
    g = golden(name)
    theta = np.concatenate([g["theta"], uniform_theta(g["bounds"], 517, seed=23)])
    if name == "sn_pantheon":
        theta[5, 3] = 60.0   # z_cosmo < 0 for the nearest SNe: NaN residuals (test_nan_residuals_propagate)
    res = []
    for fuse in (1, 0):
        with Engine(spec(name)) as e:
            e.set_option("chi2_slices", slices)
            e.set_option("fuse_planes", fuse)
            res.append((e.chi_squared(theta), e.log_probability(theta * 1.01), e.sn_moments(theta[:100]), e.chi_squared(theta[:129])))
    for a, b in zip(*res):
        assert np.array_equal(a, b, equal_nan=True)
    if name == "sn_pantheon":
        assert np.isnan(res[0][0][5]) and np.isfinite(np.delete(res[0][0], 5)).all()
