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


class BoxPrior:
    """Independent priors as the reference's nautilus scripts declare them (`prior.add_parameter(name, dist=(lo, hi))` or
    `dist=norm(loc, scale)`, e.g. bao/desi_bbn_theta_star.py:117-120): uniform on [lo, hi] per column, optionally Gaussian
    on some columns.  `transform` maps the unit cube to theta."""

    def __init__(self, bounds, gauss=None):
        self.bounds = np.asarray(bounds, dtype=np.float64)
        self.ndim = self.bounds.shape[0]
        self.gauss = dict(gauss or {})

    def transform(self, u):
        from scipy.special import ndtri
        lo, hi = self.bounds[:, 0], self.bounds[:, 1]
        theta = lo + np.asarray(u, dtype=np.float64) * (hi - lo)
        for col, (mean, sigma) in self.gauss.items():
            theta[:, col] = mean + sigma * ndtri(np.clip(u[:, col], 1e-300, 1.0 - 1e-16))
        return theta


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on numpy uint32 arrays (the generator of csrc/multi.cuh, bit for bit)."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & mask, p1 & mask, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & mask, p0 & mask
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & mask, (k1 + np.uint64(0xBB67AE85)) & mask
    return c0, c1, c2, c3


def _u53(hi, lo):
    return ((((hi >> np.uint64(5)) << np.uint64(26)) | (lo >> np.uint64(6))).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def ellipsoid_points(mu, L, n, seed, offset):
    """The points k_propose (csrc/multi.cuh) draws for rows offset .. offset + n - 1: uniform in {mu + L z, |z| < 1}.  Same
    integers; the floating-point steps (log, sincos, pow) agree with the device to a few ulp."""
    d = len(mu)
    row = np.arange(offset, offset + n, dtype=np.uint64)
    lo32, hi32 = row & np.uint64(0xFFFFFFFF), row >> np.uint64(32)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    pairs = (d + 1) // 2
    z = np.empty((n, 2 * pairs))
    for j in range(pairs + 1):
        c = philox4x32_10(lo32, hi32, np.full(n, j, dtype=np.uint64), np.zeros(n, dtype=np.uint64), k0, k1)
        a, b = _u53(c[0], c[1]), _u53(c[2], c[3])
        if j == pairs:
            ur = a
            break
        r = np.sqrt(-2.0 * np.log(a))
        z[:, 2 * j], z[:, 2 * j + 1] = r * np.cos(2.0 * np.pi * b), r * np.sin(2.0 * np.pi * b)
    z = z[:, :d]
    z *= (ur ** (1.0 / d) / np.linalg.norm(z, axis=1))[:, None]
    return np.asarray(mu) + z @ np.tril(np.asarray(L)).T


class HostProposer:
    """Proposals drawn with the counter-based generator on the host and evaluated by any `log_like_fn` (the CPU twin of
    DeviceProposer: same points, same order)."""

    def __init__(self, prior, log_like_fn, seed=0):
        self.prior, self.log_like_fn, self.seed, self.offset = prior, log_like_fn, int(seed), 0

    def draw(self, mu, L, n, thresh, need):
        u = ellipsoid_points(mu, L, n, self.seed, self.offset)
        self.offset += n
        inside = np.all((u > 0.0) & (u < 1.0), axis=1)
        u = u[inside]
        theta = np.ascontiguousarray(self.prior.transform(u)) if len(u) else np.empty((0, self.prior.ndim))
        ll = np.asarray(self.log_like_fn(theta), dtype=np.float64) if len(u) else np.empty(0)
        good = np.flatnonzero(ll > thresh)
        ok = good[:need]
        return u[ok], theta[ok], ll[ok], (int(inside.sum()), len(good), len(ok)), n


class DeviceProposer:
    """Proposals generated, prior-transformed, evaluated and filtered on the GPU (cl_propose_eval): per call only the
    accepted rows the sampler still needs come back over PCIe."""

    def __init__(self, prior, engine, what=1, seed=0):
        self.prior, self.engine, self.what, self.seed, self.offset = prior, engine, int(what), int(seed), 0

    def draw(self, mu, L, n, thresh, need):
        u, theta, ll, cnt = self.engine.propose_eval(mu, L, self.prior.bounds, n, self.seed, self.offset, self.what, thresh, need,
                                                     gauss=self.prior.gauss)
        self.offset += n
        return u, theta, ll, cnt, n


class NestedSampler:
    """Batched nested sampling sized for the GPU likelihood (the reference drives nautilus,
    `bao/desi_cmb_pantheon.py:154-170`, whose stock batches are ~100 points).

    Static nested sampling with `n_replace` deletions per iteration: the worst K live points are removed one at a time
    (live count n, n-1, ..., n-K+1, so ln X shrinks by 1/n, 1/(n-1), ...), then K replacements are drawn uniformly from
    the prior region above the K-th threshold.  Replacements come from rejection sampling inside one bounding ellipsoid
    of the live points in the unit cube (enlarged; enough for the unimodal posteriors of these fits) and every likelihood
    call evaluates `batch` (>= 65536 by default) proposals at once; accepted proposals are used in the order drawn.
    The same seed with the CPU oracle or the CUDA engine as `log_like_fn` follows the same path up to accept/reject flips
    at the 1e-9 level."""

    def __init__(self, prior, log_like_fn, n_live=2000, n_replace=None, batch=65536, enlarge=1.3, seed=0, min_batch=1024, proposer=None):
        self.prior, self.log_like_fn = prior, log_like_fn
        self.proposer = proposer   # HostProposer / DeviceProposer: counter-based proposals (the default draws with numpy's generator)
        self.ndim = prior.ndim
        self.n_live = int(n_live)
        self.n_replace = int(n_replace or max(1, n_live // 5))
        if not 1 <= self.n_replace < self.n_live:
            raise ValueError("n_replace must be in [1, n_live)")
        self.batch, self.enlarge, self.min_batch = int(batch), float(enlarge), int(min_batch)
        self._acc = 0.5       # running estimate of the acceptance of ellipsoid proposals
        self.rng = np.random.default_rng(seed)
        self.n_calls = 0      # likelihood calls (batches)
        self.n_evals = 0      # likelihood evaluations (rows)

    def _ll(self, u):
        theta = np.ascontiguousarray(self.prior.transform(u))
        self.n_calls += 1
        self.n_evals += len(theta)
        ll = np.asarray(self.log_like_fn(theta), dtype=np.float64)
        if np.isnan(ll).any():
            raise ValueError("log_like_fn returned NaN")
        return theta, ll

    def _ellipsoid(self, u_live):
        """Bounding ellipsoid of the live points: mean, Cholesky factor of the covariance scaled to contain all of them."""
        mu = u_live.mean(0)
        cov = np.cov(u_live, rowvar=False).reshape(self.ndim, self.ndim) + 1e-18 * np.eye(self.ndim)
        L = np.linalg.cholesky(cov)
        d = np.linalg.solve(L, (u_live - mu).T)
        r2 = float(np.max(np.sum(d * d, axis=0)))
        return mu, L * np.sqrt(r2) * self.enlarge ** (1.0 / self.ndim)

    def _draw(self, mu, L, n):
        z = self.rng.standard_normal((n, self.ndim))
        z *= (self.rng.random(n) ** (1.0 / self.ndim) / np.linalg.norm(z, axis=1))[:, None]   # uniform in the unit ball
        u = mu + z @ L.T
        return u[np.all((u > 0.0) & (u < 1.0), axis=1)]

    def run(self, dlogz=0.01, max_iter=100000):
        n, K, d = self.n_live, self.n_replace, self.ndim
        u = self.rng.random((n, d))
        theta, ll = self._ll(u)
        log_x = 0.0                       # ln of the enclosed prior volume
        logz = -np.inf
        h_num = 0.0                       # sum of w_i L_i ln L_i for the information H
        dead_theta, dead_logw, dead_ll = [], [], []
        # ln(X_{i-1} - X_i) for one deletion at live count m: ln X_{i-1} + ln(1 - exp(-1/m))
        for it in range(max_iter):
            order = np.argsort(ll, kind="stable")
            worst = order[:K]
            # the K deletions at live counts n, n - 1, ..., n - K + 1, vectorised (a Python loop over 8192 deletions per
            # iteration cost more than the likelihood batch)
            m = n - np.arange(K, dtype=np.float64)
            shrink = np.concatenate([[0.0], np.cumsum(1.0 / m)])
            logw = (log_x - shrink[:-1]) + np.log1p(-np.exp(-1.0 / m)) + ll[worst]
            dead_theta.append(theta[worst]); dead_logw.append(logw); dead_ll.append(ll[worst])
            logz = np.logaddexp(logz, np.logaddexp.reduce(logw))
            log_x -= shrink[-1]
            thresh = ll[worst[-1]]
            keep = order[K:]
            mu, L = self._ellipsoid(u[keep])
            new_u, new_theta, new_ll = [], [], []
            need = K
            while need > 0:
                # enough proposals for the points still needed at the acceptance seen so far, never more than `batch`
                n_prop = int(min(self.batch, max(self.min_batch, 1.2 * need / max(self._acc, 1e-4))))
                if self.proposer is not None:
                    u_ok, th_ok, ll_ok, (n_in, n_good, _), n_eval = self.proposer.draw(mu, L, n_prop, thresh, need)
                    if np.isnan(ll_ok).any():
                        raise ValueError("log_like_fn returned NaN")
                    self.n_calls += 1
                    self.n_evals += n_eval
                    self._acc = 0.5 * self._acc + 0.5 * max(n_good, 1) / n_prop
                    new_u.append(u_ok); new_theta.append(th_ok); new_ll.append(ll_ok)
                    need -= len(ll_ok)
                    continue
                cand = self._draw(mu, L, n_prop)
                if len(cand) == 0:
                    continue
                th_c, ll_c = self._ll(cand)
                good = np.flatnonzero(ll_c > thresh)
                self._acc = 0.5 * self._acc + 0.5 * max(len(good), 1) / n_prop
                ok = good[:need]
                new_u.append(cand[ok]); new_theta.append(th_c[ok]); new_ll.append(ll_c[ok])
                need -= len(ok)
            u = np.vstack([u[keep]] + new_u)
            theta = np.vstack([theta[keep]] + new_theta)
            ll = np.concatenate([ll[keep]] + new_ll)
            # remaining evidence bounded by max(L_live) X
            if np.max(ll) + log_x < logz + np.log(dlogz):
                break
        # the live points share the remaining volume
        logw_live = log_x - np.log(n) + ll
        logz = np.logaddexp(logz, np.logaddexp.reduce(logw_live))
        all_theta = np.vstack(dead_theta + [theta])
        all_logw = np.concatenate(dead_logw + [logw_live])
        all_ll = np.concatenate(dead_ll + [ll])
        w = np.exp(all_logw - logz)
        info = float(np.sum(w * all_ll) - logz)            # H = int P ln(L/Z)
        return {"logz": float(logz), "logz_err": float(np.sqrt(max(info, 0.0) / n)), "information": info, "n_iter": it + 1,
                "n_calls": self.n_calls, "n_evals": self.n_evals, "samples": all_theta, "weights": w, "log_like": all_ll}
