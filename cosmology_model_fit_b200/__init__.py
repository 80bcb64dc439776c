"""cosmology_model_fit_b200 — B200-native batched cosmological likelihood engine.

Drop-in for the chi_squared / log_likelihood / log_probability / log_probs_vectorized path of
franciscotln/cosmology-model-fit.  The compute lives in csrc/ (hand-written CUDA for sm_100a behind the C ABI
of include/cosmolike.h); this package is the thin host side: model specs and a ctypes binding.
"""
from . import spec, fits, datasets, synthetic, samplers, profile  # noqa: F401
from .spec import LikelihoodSpec  # noqa: F401
from .engine import Engine, EngineError, library_path  # noqa: F401

__all__ = ["LikelihoodSpec", "Engine", "EngineError", "library_path", "spec", "fits", "datasets", "synthetic"]
