"""world_size-2 (and 3) gloo tests of the row-sharding / all-gather host logic, with the CPU oracle as the injected
evaluator (the CUDA engine itself is covered by the -m gpu tests)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from cosmology_model_fit_b200.parallel import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_all_rows():
    for n in (0, 1, 5, 64, 65536, 100003):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    from cases import golden, spec
    import oracle.oracle as O
    from cosmology_model_fit_b200.parallel import ShardedEngine
    from cosmology_model_fit_b200.spec import OUT_CHI2
    dist.init_process_group("gloo")
    sp = spec("sn_union3_1")
    orc = O.Oracle(sp)
    eng = ShardedEngine(sp, evaluator=lambda th, what: orc._eval(th, what))
    g = golden("sn_union3_1")
    for B in (44, 7, 1):
        theta = g["theta"][:B]
        got = eng.chi_squared(theta)
        assert got.shape == (B,)
        assert np.max(np.abs(got - g["chi2"][:B])) < 1e-7, (dist.get_rank(), B)
        buf = np.empty(B)
        assert eng.chi_squared(theta, out=buf) is buf and np.array_equal(buf, got)
        only0 = eng.chi_squared(theta, root=0)          # master / worker shape: only rank 0 receives
        assert (only0 is None) == (dist.get_rank() != 0)
        if only0 is not None:
            assert np.array_equal(only0, got)
    # a Cartesian grid sharded by contiguous slices of the flattened index, reduced to (best, index, log-sum-exp)
    from cosmology_model_fit_b200.parallel import grid_points, grid_stats_of
    axes = {{1: (0.2, 0.4, 7), 0: (-0.2, 0.2, 5)}}       # axis 0 = column 1 (slowest), axis 1 = column 0 (fastest)
    fixed = {{2: -3.0}}
    stats, vals, (first, count) = eng.grid(axes, fixed, OUT_CHI2, want_values=True)
    full = orc._eval(grid_points(axes, fixed, sp.ndim, 0, 35), OUT_CHI2)
    assert np.array_equal(vals, full[first:first + count])
    want = grid_stats_of(full, 0)
    assert stats["index"] == want["index"] == int(np.argmin(full)) and stats["best"] == want["best"] == full.min()
    assert abs(stats["log_sum"] - want["log_sum"]) < 1e-12 and stats["count"] == 35
    om, dm = np.meshgrid(np.linspace(0.2, 0.4, 7), np.linspace(-0.2, 0.2, 5), indexing="ij")
    pts = grid_points(axes, fixed, sp.ndim, 0, 35)
    assert np.array_equal(pts[:, 1], om.ravel()) and np.array_equal(pts[:, 0], dm.ravel()) and np.all(pts[:, 2] == -3.0)
    from cosmology_model_fit_b200.spec import OUT_LOGLIKE
    ll_stats, _, _ = eng.grid(axes, fixed, OUT_LOGLIKE)
    assert ll_stats["index"] == want["index"] and abs(ll_stats["best"] + 0.5 * want["best"]) < 1e-12
    assert abs(ll_stats["log_sum"] - want["log_sum"]) < 1e-12
    dist.barrier()
    dist.destroy_process_group()
    print("rank" + os.environ["RANK"] + "ok", flush=True)
""")


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_engine_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = 29600 + world
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == world
