"""Readers for the reference's on-disk data formats.

The reference parses its inputs at import time of every script: the Pantheon+ stat+sys covariance is a 2.9-million-line
text file read with pandas (`y2022pantheonSHOES/data.py:9,21-23`, seconds per import), DES-SN5YR Dovekie ships a `.npy`
(`y2025DESdovekie/data.py:18`), Union3.1 and DESI small csv / whitespace tables (`y2026union3_1/data.py`,
`y2025BAO/data.py`).  These functions read the same files and return the tuples `fits.py` takes — the same row selection
and ordering as the reference loaders — and keep a binary cache of the large covariance next to the text file (or under
`cache_dir`), so the second start of a fit costs milliseconds.

Two layouts of the Pantheon+ covariance are accepted: the reference's `covariance_stat_sys.txt` (one header line, then N^2
values) and the public release `Pantheon+SH0ES_STAT+SYS.cov` (first line N, then N^2 values).
"""
from __future__ import annotations

import hashlib
import os

import numpy as np


def _read_table(path, sep=None):
    """Whitespace / comma separated table with a header line -> dict of column arrays (strings stay strings)."""
    with open(path, "r") as f:
        header = f.readline()
        sep = sep or ("," if "," in header else None)
        names = [h.strip() for h in (header.split(sep) if sep else header.split())]
        cols = [[] for _ in names]
        for line in f:
            parts = line.split(sep) if sep else line.split()
            if len(parts) < len(names):
                continue
            for c, p in zip(cols, parts):
                c.append(p.strip())
    out = {}
    for n, c in zip(names, cols):
        try:
            out[n] = np.array(c, dtype=np.float64)
        except ValueError:
            out[n] = np.array(c)
    return out


def _cache_path(src, cache_dir):
    st = os.stat(src)
    tag = hashlib.sha1(f"{os.path.abspath(src)}:{st.st_size}:{st.st_mtime_ns}".encode()).hexdigest()[:16]
    base = os.path.basename(src) + f".{tag}.npy"
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        return os.path.join(cache_dir, base)
    return os.path.join(os.path.dirname(os.path.abspath(src)), base)


def read_flat_covariance(path, n=None, cache=True, cache_dir=None):
    """One value per line (optionally preceded by a header word or by N): the N x N matrix, row-major.  The parsed matrix
    is cached as .npy keyed by the file's path, size and mtime."""
    cpath = _cache_path(path, cache_dir) if cache else None
    if cpath and os.path.exists(cpath):
        return np.load(cpath)
    with open(path, "rb") as f:
        tokens = f.read().split()
    if tokens:
        try:
            float(tokens[0])
        except ValueError:
            tokens = tokens[1:]                          # a column name (the reference's `cov_mu_shoes`)
    vals = np.array(tokens, dtype=np.float64)            # C loop of float(): ~1 s for the 2.9 M values of Pantheon+
    m = int(round(np.sqrt(vals.size)))
    if m * m != vals.size:
        m = int(round(np.sqrt(max(vals.size - 1, 0))))
        if m * m + 1 == vals.size and vals[0] == m:
            vals = vals[1:]                              # public .cov layout: the first line is N
        else:
            raise ValueError(f"{path}: {vals.size} values do not form a square matrix")
    if n is not None and m != n:
        raise ValueError(f"{path}: covariance of order {m}, expected {n}")
    cov = vals.reshape(m, m)
    if cpath:
        try:
            np.save(cpath, cov)
        except OSError:
            pass                                        # read-only data directory: no cache
    return cov


def pantheon_plus_files(distances, covariance, cut=True, with_positions=False, cache=True, cache_dir=None):
    """Pantheon+ (y2022pantheonSHOES/data.py): (z_cmb, z_hel, m_b, cov) with the zHD > 0.01 selection of the reference
    (`cut=False` keeps all 1701 rows); `with_positions=True` appends (RA, DEC, IDSURVEY) as `get_data_with_position`."""
    t = _read_table(distances)
    cov = read_flat_covariance(covariance, n=t["zHD"].size, cache=cache, cache_dir=cache_dir)
    keep = np.where(t["zHD"] > 0.01)[0] if cut else np.arange(t["zHD"].size)
    out = (t["zHD"][keep], t["zHEL"][keep], t["m_b_corr"][keep], cov[np.ix_(keep, keep)])
    if with_positions:
        out += (t["RA"][keep], t["DEC"][keep], t["IDSURVEY"][keep].astype(np.int32))
    return out


def pantheon_plus_shoes_files(distances, covariance, cache=True, cache_dir=None):
    """Pantheon+SH0ES (y2022pantheonSHOES/data_shoes.py:24-39): calibrators kept at any redshift;
    (z_cmb, z_hel, m_b, ceph_dist, cov)."""
    t = _read_table(distances)
    cov = read_flat_covariance(covariance, n=t["zHD"].size, cache=cache, cache_dir=cache_dir)
    sel = np.where((t["IS_CALIBRATOR"] == 1) | (t["zHD"] > 0.01))[0]
    return t["zHD"][sel], t["zHEL"][sel], t["m_b_corr"][sel], t["CEPH_DIST"][sel], cov[np.ix_(sel, sel)]


def des_dovekie_files(distances, covariance_npy):
    """DES-SN5YR Dovekie (y2025DESdovekie/data.py): rows sorted by zHD; (z_cmb, z_hel, mu, cov)."""
    t = _read_table(distances)
    cov = np.load(covariance_npy)
    o = np.argsort(t["zHD"])
    return t["zHD"][o], t["zHEL"][o], t["MU"][o], cov[o, :][:, o]


def union3_1_files(bins_csv, covariance_txt):
    """Union3.1 binned distances (y2026union3_1/data.py): (z_cmb, z_hel, mu, cov)."""
    t = _read_table(bins_csv, sep=",")
    n = t["zcmb"].size
    cov = np.array(open(covariance_txt).read().split(), dtype=np.float64).reshape(n, n)
    return t["zcmb"], t["zhel"], t["mb"], cov


def bao_files(data_csv, covariance_txt):
    """DESI-style BAO table (y2025BAO/data.py, y2025BAO/data_fs_lya.py): columns z, value, quantity; (z, value, quantity, cov)."""
    t = _read_table(data_csv, sep=",")
    n = t["z"].size
    cov = np.array(open(covariance_txt).read().split(), dtype=np.float64).reshape(n, n)
    return t["z"], t["value"], t["quantity"], cov
