"""mitsuba-alvrl_b200: B200-native hot path of the Adaptive LightSlice VRL integrator.

The product is csrc/ (hand-written sm_100a CUDA kernels + host control in C++) behind the C ABI of
include/alvrl.h, built in-tree as libalvrl.so.  This package is only the Python face of that ABI (ctypes) for
tests, bench.py and torch.distributed plumbing.  There is no CPU fallback: `integrator()` raises if the CUDA
library is missing."""
import os

from . import binding, rms, scenes, sharding, volfile  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ALVRL_LIB", os.path.join(_HERE, "libalvrl.so"))   # ALVRL_LIB: kernel-variant experiments (tools/build_variant.sh)
_api = None


def api():
    global _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA extension with `python -c 'import __graft_entry__ as g; "
                "g.build()'` -- this package has no CPU fallback")
        _api = binding.Api(LIB_PATH, "alvrl_")
    return _api


def integrator(device=0, **params):
    """Create a vrl integrator handle on a CUDA device (parameter names of vrlIntegrator.cpp:128-208)."""
    return binding.Integrator(api(), device, **params)
