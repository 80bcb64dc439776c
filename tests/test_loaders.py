"""Readers of the reference's on-disk formats (cosmology_model_fit_b200/loaders.py): synthetic files in the layouts of
y2022pantheonSHOES / y2025DESdovekie / y2026union3_1 / y2025BAO, and — when the reference checkout is present — its own
data files against the committed column fixtures."""
import os

import numpy as np
import pytest

from cosmology_model_fit_b200 import datasets, loaders

REF = os.environ.get("COSMO_REFERENCE", "/root/reference")


def _write_pantheon(tmp, n=40, layout="header"):
    rng = np.random.default_rng(0)
    z = np.sort(rng.uniform(0.002, 1.5, n))
    cols = {"CID": [f"sn{i}" for i in range(n)], "IDSURVEY": rng.integers(1, 150, n), "zHD": z, "zHEL": z * 0.999,
            "m_b_corr": 20 + 5 * np.log10(z + 0.01), "RA": rng.uniform(0, 360, n), "DEC": rng.uniform(-60, 60, n),
            "CEPH_DIST": np.where(np.arange(n) < 3, 30.0, -9.0), "IS_CALIBRATOR": (np.arange(n) < 3).astype(int)}
    with open(tmp / "distances.txt", "w") as f:
        f.write(" ".join(cols) + "\n")
        for i in range(n):
            f.write(" ".join(repr(cols[k][i].item()) if hasattr(cols[k][i], "item") else str(cols[k][i]) for k in cols) + "\n")
    A = rng.standard_normal((n, n))
    cov = A @ A.T / n + np.eye(n)
    with open(tmp / "cov.txt", "w") as f:
        f.write("cov_mu_shoes\n" if layout == "header" else f"{n}\n")
        f.write("\n".join(repr(float(x)) for x in cov.ravel()) + "\n")
    return cols, cov


@pytest.mark.parametrize("layout", ["header", "count"])
def test_pantheon_files_both_covariance_layouts(tmp_path, layout):
    cols, cov = _write_pantheon(tmp_path, layout=layout)
    z, zh, mb, c = loaders.pantheon_plus_files(tmp_path / "distances.txt", tmp_path / "cov.txt")
    keep = np.where(cols["zHD"] > 0.01)[0]                      # y2022pantheonSHOES/data.py:25
    assert np.array_equal(z, cols["zHD"][keep]) and np.array_equal(mb, cols["m_b_corr"][keep])
    assert np.array_equal(c, cov[np.ix_(keep, keep)])
    full = loaders.pantheon_plus_files(tmp_path / "distances.txt", tmp_path / "cov.txt", cut=False, with_positions=True)
    assert full[0].size == 40 and full[6].dtype == np.int32 and np.array_equal(full[4], cols["RA"])
    # the second read comes from the .npy cache written next to the text file
    cached = [f for f in os.listdir(tmp_path) if f.endswith(".npy")]
    assert len(cached) == 1
    os.remove(tmp_path / cached[0])
    np.save(tmp_path / cached[0], cov * 0 + 7.0)
    assert np.all(loaders.read_flat_covariance(tmp_path / "cov.txt") == 7.0)
    assert np.array_equal(loaders.read_flat_covariance(tmp_path / "cov.txt", cache=False), cov)
    # SH0ES selection keeps the calibrators at any redshift (y2022pantheonSHOES/data_shoes.py:24-39)
    sh = loaders.pantheon_plus_shoes_files(tmp_path / "distances.txt", tmp_path / "cov.txt", cache=False)
    sel = np.where((cols["IS_CALIBRATOR"] == 1) | (cols["zHD"] > 0.01))[0]
    assert np.array_equal(sh[0], cols["zHD"][sel]) and np.array_equal(sh[3], cols["CEPH_DIST"][sel])


def test_bad_covariance_size_is_an_error(tmp_path):
    _write_pantheon(tmp_path)
    with open(tmp_path / "cov.txt", "a") as f:
        f.write("1.0\n2.0\n")
    with pytest.raises(ValueError):
        loaders.read_flat_covariance(tmp_path / "cov.txt", cache=False)


def test_des_union3_bao_files(tmp_path):
    rng = np.random.default_rng(1)
    n = 12
    z = rng.uniform(0.02, 1.1, n)
    with open(tmp_path / "d.csv", "w") as f:
        f.write("CID IDSURVEY zHD zHEL MU MUERR MUERR_VPEC MUERR_SYS PROBIA_BEAMS\n")
        for i in range(n):
            f.write(f"sn{i}  10 {float(z[i])!r} {float(z[i] * 1.001)!r} {float(35 + z[i])!r}   0.1  0.07  0.05  1.0\n")
    cov = np.diag(np.arange(1.0, n + 1))
    np.save(tmp_path / "c.npy", cov)
    zz, zh, mu, c = loaders.des_dovekie_files(tmp_path / "d.csv", tmp_path / "c.npy")
    o = np.argsort(z)                                            # y2025DESdovekie/data.py: rows sorted by zHD
    assert np.array_equal(zz, z[o]) and np.array_equal(np.diag(c), np.diag(cov)[o])
    with open(tmp_path / "u.csv", "w") as f:
        f.write("zcmb,zhel,mb\n0.05,0.051,36.1\n0.5,0.5,42.2\n")
    with open(tmp_path / "ucov.txt", "w") as f:
        f.write("1.0 0.1\n0.1 2.0\n")
    u = loaders.union3_1_files(tmp_path / "u.csv", tmp_path / "ucov.txt")
    assert np.array_equal(u[2], [36.1, 42.2]) and u[3][1, 1] == 2.0
    with open(tmp_path / "b.csv", "w") as f:
        f.write("z,value,quantity\n0.295,7.94,DV_over_rs\n0.51,13.59,DM_over_rs\n0.51,21.86,DH_over_rs\n")
    with open(tmp_path / "bcov.txt", "w") as f:
        f.write("1 0 0\n0 2 0.5\n0 0.5 3\n")
    b = loaders.bao_files(tmp_path / "b.csv", tmp_path / "bcov.txt")
    assert list(b[2]) == ["DV_over_rs", "DM_over_rs", "DH_over_rs"] and b[3][1, 2] == 0.5


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_reference_files_match_the_committed_fixtures():
    """The reference's own raw files through loaders.py against cosmology_model_fit_b200/data/data_*.npz (the covariance blobs of
    Pantheon+ and DES are missing from the checkout: columns only)."""
    t = loaders._read_table(f"{REF}/y2022pantheonSHOES/raw-data/distances.txt")
    d = np.load(os.path.join(datasets._DATA, "data_pantheon_plus.npz"))
    assert np.array_equal(t["zHD"], d["zHD"]) and np.array_equal(t["m_b_corr"], d["m_b_corr"]) and t["zHD"].size == 1701
    t = loaders._read_table(f"{REF}/y2025DESdovekie/raw-data/distances.csv")
    d = np.load(os.path.join(datasets._DATA, "data_des_dovekie.npz"))
    assert np.array_equal(t["MU"], d["MU"]) and t["MU"].size == 1820
    u = loaders.union3_1_files(f"{REF}/y2026union3_1/raw-data/bins_union_3_1.csv", f"{REF}/y2026union3_1/raw-data/covariance.txt")
    for a, b in zip(u, datasets.union3_1()):
        assert np.array_equal(a, b)
    b = loaders.bao_files(f"{REF}/y2025BAO/raw-data/data.csv", f"{REF}/y2025BAO/raw-data/covariance.txt")
    for a, c in zip(b, datasets.desi_dr2()):
        assert np.array_equal(a, c)
