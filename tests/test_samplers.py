"""The batched ensemble sampler: (CPU) it recovers a known Gaussian posterior; (GPU) driven by the CUDA engine and by the
CPU oracle with the same seed it gives posterior means that agree within Monte Carlo noise (north star: config 0)."""
import numpy as np
import pytest

from cosmology_model_fit_b200.samplers import EnsembleSampler, laplace_log_evidence


def test_recovers_gaussian_posterior():
    mean = np.array([1.0, -2.0, 0.5])
    cov = np.array([[1.0, 0.6, 0.0], [0.6, 2.0, -0.3], [0.0, -0.3, 0.5]])
    icov = np.linalg.inv(cov)
    lp = lambda th: -0.5 * np.einsum("ni,ij,nj->n", th - mean, icov, th - mean)
    s = EnsembleSampler(256, 3, lp, de_fraction=0.3, seed=1)
    p0 = np.random.default_rng(0).normal(0, 1, (256, 3))
    s.run_mcmc(p0, 400)
    flat = s.get_chain(discard=150, flat=True)
    assert np.all(np.abs(flat.mean(0) - mean) < 0.08)
    assert np.all(np.abs(np.cov(flat.T) - cov) < 0.2)
    assert 0.2 < s.acceptance_fraction < 0.8
    assert s.n_calls == 1 + 2 * 400  # one vectorised call per half step


def test_laplace_evidence_of_a_gaussian():
    cov = np.array([[0.5, 0.1], [0.1, 0.2]])
    icov = np.linalg.inv(cov)
    lp = lambda th: -0.5 * np.einsum("ni,ij,nj->n", th, icov, th)
    lnz, H = laplace_log_evidence(lp, np.zeros(2), scales=np.ones(2))
    assert abs(lnz - 0.5 * np.log((2 * np.pi) ** 2 * np.linalg.det(cov))) < 1e-6
    assert np.allclose(H, -icov, atol=1e-5)


@pytest.mark.gpu
def test_config0_posterior_means_gpu_vs_oracle():
    import oracle.oracle as O
    from cases import spec
    from cosmology_model_fit_b200 import Engine
    sp = spec("sn_pantheon")
    orc = O.Oracle(sp)
    nwalk, nsteps, burn = 64, 260, 100
    rng = np.random.default_rng(3)
    p0 = np.array([-19.35, 70.4, 0.33, 0.0]) + rng.normal(0, 1, (nwalk, 4)) * np.array([0.02, 1.0, 0.02, 0.3])
    with Engine(sp) as eng:
        s_gpu = EnsembleSampler(nwalk, 4, lambda th: eng.log_probs_vectorized(th, dtype=np.float64), seed=7)
        s_gpu.run_mcmc(p0, nsteps)
    s_cpu = EnsembleSampler(nwalk, 4, lambda th: orc.log_probability(th, nthreads=0), seed=7)
    s_cpu.run_mcmc(p0, nsteps)
    a, b = s_gpu.get_chain(discard=burn, flat=True), s_cpu.get_chain(discard=burn, flat=True)
    # Monte Carlo noise of the mean: posterior std / sqrt(effective samples) with a generous autocorrelation time of 40
    noise = b.std(0) / np.sqrt(len(b) / 40.0)
    assert np.all(np.abs(a.mean(0) - b.mean(0)) < 3 * noise), (a.mean(0), b.mean(0), noise)
    # the chains are in fact identical until an accept/reject decision flips at the 1e-9 level
    same = np.all(s_gpu.chain == s_cpu.chain, axis=(1, 2))
    assert same[:20].all()
