"""Batched ensemble sampler sized for the GPU likelihood.

emcee / nautilus are what the reference's `main()`s drive (`sn/pantheon.py:108-125`); with stock settings they hand the
likelihood 50-100 rows per call, far too few for a GPU.  This is a minimal affine-invariant ensemble
sampler (Goodman & Weare stretch move + differential-evolution move, the red/blue half-ensemble update emcee uses)
whose only contact with the model is ONE vectorised call per half step: `log_prob_fn(theta[n, d]) -> [n]`.  With
131072 walkers each call is a B = 65536 batch — the shape `bench.py` measures.

It is deliberately sampler-library-agnostic test/driver infrastructure: the same seed with the CPU oracle or the CUDA
engine as `log_prob_fn` produces the same chain up to accept/reject flips at the 1e-9 level.
"""
from __future__ import annotations

import numpy as np


class EnsembleSampler:
    def __init__(self, nwalkers, ndim, log_prob_fn, a=2.0, de_fraction=0.0, seed=42):
        if nwalkers % 2 or nwalkers < 2 * ndim:
            raise ValueError("nwalkers must be even and >= 2 * ndim")
        self.nwalkers, self.ndim, self.log_prob_fn = nwalkers, ndim, log_prob_fn
        self.a, self.de_fraction = float(a), float(de_fraction)
        self.rng = np.random.default_rng(seed)
        self.chain = None
        self.log_prob = None
        self.n_accepted = 0
        self.n_proposed = 0
        self.n_calls = 0

    def _lp(self, theta):
        self.n_calls += 1
        lp = np.asarray(self.log_prob_fn(np.ascontiguousarray(theta)), dtype=np.float64)
        if np.isnan(lp).any():
            raise ValueError("log_prob_fn returned NaN")  # emcee treats NaN as fatal too
        return lp

    def run_mcmc(self, p0, nsteps):
        p = np.array(p0, dtype=np.float64)
        if p.shape != (self.nwalkers, self.ndim):
            raise ValueError("p0 must be [nwalkers, ndim]")
        lp = self._lp(p)
        half = self.nwalkers // 2
        chain = np.empty((nsteps, self.nwalkers, self.ndim))
        lps = np.empty((nsteps, self.nwalkers))
        for step in range(nsteps):
            for first in (0, 1):
                s = slice(0, half) if first == 0 else slice(half, None)      # walkers being updated
                c = slice(half, None) if first == 0 else slice(0, half)      # complementary ensemble
                cur, comp = p[s], p[c]
                n = cur.shape[0]
                use_de = self.rng.random() < self.de_fraction
                if use_de:  # differential evolution: x + gamma (c_j - c_k), symmetric -> plain Metropolis ratio
                    j = self.rng.integers(0, comp.shape[0], n)
                    k = (j + self.rng.integers(1, comp.shape[0], n)) % comp.shape[0]
                    gamma = 2.38 / np.sqrt(2 * self.ndim) * (1 + 1e-5 * self.rng.standard_normal(n))
                    prop = cur + gamma[:, None] * (comp[j] - comp[k])
                    log_factor = np.zeros(n)
                else:       # stretch move: z ~ g(z) on [1/a, a], y = c_j + z (x - c_j), ratio z^(d-1)
                    z = ((self.a - 1.0) * self.rng.random(n) + 1.0) ** 2 / self.a
                    j = self.rng.integers(0, comp.shape[0], n)
                    prop = comp[j] + z[:, None] * (cur - comp[j])
                    log_factor = (self.ndim - 1) * np.log(z)
                lp_prop = self._lp(prop)
                accept = np.log(self.rng.random(n)) < log_factor + lp_prop - lp[s]
                cur[accept] = prop[accept]
                lp_s = lp[s]
                lp_s[accept] = lp_prop[accept]
                self.n_accepted += int(accept.sum())
                self.n_proposed += n
            chain[step], lps[step] = p, lp
        self.chain, self.log_prob = chain, lps
        return p, lp

    @property
    def acceptance_fraction(self):
        return self.n_accepted / max(1, self.n_proposed)

    def get_chain(self, discard=0, flat=False):
        c = self.chain[discard:]
        return c.reshape(-1, self.ndim) if flat else c


def laplace_log_evidence(log_prob_fn, theta_map, step=1e-4, scales=None):
    """Laplace approximation of ln Z around a mode (log_evidence.py:7-70 uses numdifftools for the Hessian); the
    2 d^2 + 1 finite-difference points are evaluated as ONE batch."""
    theta_map = np.asarray(theta_map, dtype=np.float64)
    d = theta_map.size
    h = step * (np.abs(theta_map) + 1e-3 if scales is None else np.asarray(scales, dtype=np.float64))
    pts = [theta_map.copy()]
    for i in range(d):
        for j in range(i, d):
            for si, sj in ((1, 1), (1, -1), (-1, 1), (-1, -1)):
                t = theta_map.copy()
                t[i] += si * h[i]
                t[j] += sj * h[j]
                pts.append(t)
    lp = np.asarray(log_prob_fn(np.array(pts)), dtype=np.float64)
    H = np.empty((d, d))
    k = 1
    for i in range(d):
        for j in range(i, d):
            pp, pm, mp, mm = lp[k:k + 4]
            k += 4
            if i == j:  # points are theta +- 2 h e_i (twice) and theta (twice)
                H[i, i] = (pp - 2 * lp[0] + mm) / (4 * h[i] * h[i])
            else:
                H[i, j] = H[j, i] = (pp - pm - mp + mm) / (4 * h[i] * h[j])
    sign, logdet = np.linalg.slogdet(-H)
    if sign <= 0:
        raise ValueError("Hessian is not negative definite at theta_map")
    return float(lp[0] + 0.5 * d * np.log(2 * np.pi) - 0.5 * logdet), H
