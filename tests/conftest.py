import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def _gpu_ready():
    """True when the in-tree CUDA library exists and a compute-capability-10 device answers."""
    from cosmology_model_fit_b200.engine import library_path
    if not os.path.exists(library_path()):
        return False, "libcosmolike_b200.so is not built"
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device visible"
        if torch.cuda.get_device_capability(0)[0] != 10:
            return False, "device 0 is not sm_100"
    except Exception as e:  # pragma: no cover
        return False, f"torch.cuda probe failed: {e}"
    return True, ""


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) on a box without a B200 or without the built library.  An explicit
    `-m gpu` run on such a box would then skip everything silently, so it fails loudly instead."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    ok, why = _gpu_ready()
    if ok:
        return
    if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or ""):
        raise pytest.UsageError(f"-m gpu requested but {why}: the GPU tests have no CPU fallback")
    skip = pytest.mark.skip(reason=why)
    for it in gpu_items:
        it.add_marker(skip)
