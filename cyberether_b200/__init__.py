"""cyberether_b200 — Blackwell (sm_100a) compute backend for the Jetstream DSP module compute() path.

Layout:
  csrc/          CUDA kernels + the C ABI (include/b200dsp.h) -> libb200dsp.so
  _native.py     ctypes loader for that library (no fallback)
  jetstream.py   host-side mirror of the reference Module / Runtime / Scheduler interface
  blocks.py      spectrum_engine / filter / fm block wiring on this provider
"""
from . import _native  # noqa: F401
from .jetstream import (  # noqa: F401
    Result, Taint, Tensor, TensorLink, Module, TestContext, NativeCudaRuntime, SynchronousScheduler,
    build_module, list_available_modules, last_error, amplitude_scaling_coeff, range_coefficients,
)
