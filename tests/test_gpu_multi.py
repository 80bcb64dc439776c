"""The multi-GPU entry points of the C ABI on ONE GPU: a single-rank NCCL communicator (cl_comm_init with nranks = 1) makes
cl_eval_allgather / cl_grid_allreduce runnable here; the device-generated grid (cl_eval_grid) is held to the host statement
of the same grid.  Two ranks need two GPUs: tools/multi_gpu_check.py under torchrun (profiles/)."""
import numpy as np
import pytest

from cases import golden, spec
from cosmology_model_fit_b200 import Engine
from cosmology_model_fit_b200.engine import GRID_MARGINAL, GRID_PROFILE
from cosmology_model_fit_b200.parallel import ShardedEngine, grid_points, grid_stats_of
from cosmology_model_fit_b200.profile import offset_profile
from cosmology_model_fit_b200.spec import OUT_CHI2, OUT_LOGLIKE, OUT_LOGPROB
from cosmology_model_fit_b200.synthetic import uniform_theta

pytestmark = pytest.mark.gpu


def test_single_rank_communicator_allgather_host_and_device():
    import torch
    g = golden("sn_pantheon")
    theta = uniform_theta(g["bounds"], 700, seed=9)
    with Engine(spec("sn_pantheon")) as e:
        want = e.log_probability(theta)
        assert e.comm_info() == (0, 0)
        with pytest.raises(Exception, match="communicator"):
            e.eval_allgather(theta, OUT_LOGPROB)
        e.comm_init(0, 1, Engine.nccl_unique_id())
        assert e.comm_info() == (0, 1)
        assert np.array_equal(e.eval_allgather(theta, OUT_LOGPROB), want)
        pinned_t = e.pinned_empty(theta.shape); pinned_t[...] = theta
        pinned_o = e.pinned_empty((700,))
        assert e.eval_allgather(pinned_t, OUT_LOGPROB, out_all=pinned_o, root=0) is pinned_o and np.array_equal(pinned_o, want)
        e.set_option("max_rows_per_pass", 256)       # several passes per call
        assert np.array_equal(e.eval_allgather(theta, OUT_LOGPROB), want)
        d_t = torch.from_numpy(theta).cuda()
        d_o = torch.empty(700, dtype=torch.float64, device="cuda")
        st = torch.cuda.Stream()
        e.eval_allgather_device(d_t.data_ptr(), 700, 4, OUT_LOGPROB, d_o.data_ptr(), st.cuda_stream)
        st.synchronize()
        assert np.array_equal(d_o.cpu().numpy(), want)
        # ShardedEngine on an explicit (rank, world) pair: world 1 needs no communicator
        sh = ShardedEngine(spec("sn_pantheon"), engine=e, rank=0, world=1)
        assert np.array_equal(sh.log_probability(theta), want)
        assert np.array_equal(sh.evaluate_local(theta[:33], OUT_LOGPROB), want[:33])


@pytest.mark.parametrize("what", [GRID_PROFILE, GRID_MARGINAL])
def test_device_grid_with_profiled_offset_matches_the_host_grid(what):
    """theta = (M, H0, Om, v): grid over (Om, v) with M profiled / marginalised from the two-dot epilogue; the device-generated
    points are np.linspace's bits, the values those of cl_eval_sn_moments on the same rows, the stats those of the values."""
    axes = {2: (0.15, 0.55, 41), 3: (-2.5, 2.5, 23)}
    fixed = {0: 0.0, 1: 70.0}
    n = 41 * 23
    with Engine(spec("sn_pantheon")) as e:
        pts = grid_points(axes, fixed, 4, 0, n)
        want, _ = offset_profile(e.sn_moments(pts), "profile" if what == GRID_PROFILE else "marginal")
        grid = e.make_grid(axes, fixed)
        stats, vals = e.eval_grid(grid, 0, n, what, want_values=True)
        assert np.array_equal(vals, want)
        ref = grid_stats_of(want, 0)
        assert stats["index"] == ref["index"] and stats["best"] == ref["best"] and stats["count"] == n
        assert abs(stats["log_sum"] - ref["log_sum"]) < 1e-10
        # slices: values are the same bits, stats fold to the whole
        e.set_option("max_rows_per_pass", 256)
        s1, v1 = e.eval_grid(grid, 0, 500, what, want_values=True)
        s2, v2 = e.eval_grid(grid, 500, n - 500, what, want_values=True)
        assert np.array_equal(np.concatenate([v1, v2]), want)
        from cosmology_model_fit_b200.parallel import fold_grid_stats
        tot = fold_grid_stats([s1, s2])
        assert tot["index"] == ref["index"] and tot["best"] == ref["best"] and abs(tot["log_sum"] - ref["log_sum"]) < 1e-10
        s0, none = e.eval_grid(grid, 0, n, what)
        assert none is None and s0["index"] == ref["index"]
        # single-rank all-reduce is the identity
        e.comm_init(0, 1, Engine.nccl_unique_id())
        sa, _ = e.eval_grid(grid, 0, n, what, allreduce=True)
        assert sa == s0
        with pytest.raises(Exception, match="outside the grid"):
            e.eval_grid(grid, n - 3, 10, what)


def test_device_grid_plain_selectors_and_prior_rows():
    """chi2 / log L / log P on a grid that leaves the prior box (Om from -0.1): -inf rows carry no weight and never win."""
    axes = {0: (-19.6, -19.1, 6), 2: (-0.1, 0.6, 15), 1: (60.0, 80.0, 5)}
    fixed = {3: 0.5}
    n = 6 * 15 * 5
    with Engine(spec("sn_pantheon")) as e:
        pts = grid_points(axes, fixed, 4, 0, n)
        grid = e.make_grid(axes, fixed)
        for what, larger in ((OUT_CHI2, False), (OUT_LOGLIKE, True), (OUT_LOGPROB, True)):
            want = e._eval(pts, what)
            stats, vals = e.eval_grid(grid, 0, n, what, want_values=True)
            assert np.array_equal(vals, want, equal_nan=True), what
            ref = grid_stats_of(want, 0, larger)
            assert stats["index"] == ref["index"] and stats["best"] == ref["best"], (what, stats, ref)
            assert abs(stats["log_sum"] - ref["log_sum"]) < 1e-10
        assert np.isneginf(e._eval(pts, OUT_LOGPROB)).any()
