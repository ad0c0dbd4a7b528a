"""vihmc -- B200-native batched-chain HMC engine for the VI-HMC hot path.

Host side (this package) mirrors the reference's call surface; all arithmetic runs in
hand-written sm_100a CUDA behind the C ABI declared in ``include/vihmc.h`` (``libvihmc.so``,
loaded with ctypes).  There is no CPU fallback: anything that computes raises if the library or a
GPU is missing.
"""
from . import engine, samplers, spec, synth, util, validate  # noqa: F401
from .samplers import Integrator, Sampler, predict_model, sample, sample_model  # noqa: F401
from .spec import DeepONetArch, LogProbSpec, MLPArch  # noqa: F401

__version__ = "0.1.0"
