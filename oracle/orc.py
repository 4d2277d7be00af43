"""TEST INFRASTRUCTURE ONLY: Python loader of the CPU oracle (oracle/liborc.so).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
The product package never imports this module."""
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import alvrl_loader  # noqa: E402

_pkg = alvrl_loader.load()
binding = _pkg.binding
C = binding.C


def build(force=False):
    lib = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "oracle_core.hpp", "oracle_prep.hpp")]
    stale = force or not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return lib


_api = {}


def api(fast=False):
    """liborc.so (IEEE, the parity definition) or liborc_fast.so (speed baseline only)."""
    if fast not in _api:
        build()
        _api[fast] = binding.Api(os.path.join(_HERE, "liborc_fast.so" if fast else "liborc.so"), "orc_")
    return _api[fast]


def film(frames, filter, param=0.0):
    """frames: [passes, H, W, 3] (or one [H, W, 3] frame) -> developed film [H, W, 3] (orc_film: rfilter table, ImageBlock::put
    in raster order, weight division)"""
    fr = np.ascontiguousarray(frames, dtype=np.float32)
    if fr.ndim == 3:
        fr = fr[None]
    n, H, W, _ = fr.shape
    out = np.zeros((H, W, 3), np.float32)
    rc = api().lib.orc_film(C.c_uint32(W), C.c_uint32(H), C.c_int(filter), C.c_float(param), binding._p(fr), C.c_uint32(n), binding._p(out))
    if rc != 0:
        raise RuntimeError("orc_film failed")
    return out


class Oracle(binding.Integrator):
    def __init__(self, fast=False, threads=None, **params):
        super().__init__(api(fast), 0, **params)
        self.set_threads(threads or os.cpu_count() or 1)

    def set_threads(self, n):
        self._call("set_threads", C.c_int(n))

    def set_no_visibility(self, on):
        self._call("set_no_visibility", C.c_int(int(on)))

    def set_graze_tolerance(self, tol):
        self._call("set_graze_tolerance", C.c_float(tol))

    def R_graze(self, r0=0, r1=None):
        """uint8 [rows, N]: 1 where some shadow ray of the entry had an occlusion decision within the tolerance of flipping"""
        S, G = self.num_slices()
        r1 = G if r1 is None else r1
        out = np.zeros((r1 - r0, self.N), np.uint8)
        self._call("get_R_graze", C.c_uint32(r0), C.c_uint32(r1), binding._p(out))
        return out

    def gather_points(self):
        P = self.W * self.H
        pos, d = np.zeros((P, 3), np.float32), np.zeros((P, 3), np.float32)
        self._call("get_gather_points", binding._p(pos), binding._p(d))
        return pos, d

    def primary_ties(self):
        t = np.zeros(self.W * self.H, np.uint8)
        self._call("get_primary_ties", binding._p(t))
        return t

    def build_R_record_tape(self):
        S, G = self.num_slices()
        K = (2 * self.params.volVolSamples + self.params.volSurfSamples) * self.params.Rsamples
        tape = np.zeros(G * self.N * K, np.float32)
        self._call("build_R_record_tape", binding._p(tape), C.c_uint64(tape.size))
        return tape.reshape(G, self.N, K)

    def cluster_diag(self):
        a, b, c = C.c_uint32(), C.c_uint64(), C.c_uint64()
        self._call("get_cluster_diag", C.byref(a), C.byref(b), C.byref(c))
        return dict(near_tie_splits=a.value, splits=b.value, variance_steps=c.value)

    def render_pixels(self, pixels, clustered=True):
        px = binding._u32(pixels)
        out = np.zeros((len(px), 3), np.float32)
        self._call("render_pixels", binding._p(px), C.c_uint32(len(px)), binding._p(out), C.c_int(int(clustered)))
        return out

    def integrate_pair(self, pixel, vrl, uniforms):
        u = binding._f32(uniforms)
        out = np.zeros(5, np.float32)
        self._call("integrate_pair", C.c_uint32(pixel), C.c_uint32(vrl), binding._p(u), C.c_uint32(len(u)),
                   binding._p(out))
        return out


def sfmt_ulongs(seed, n):
    out = np.zeros(n, np.uint64)
    api().lib.orc_sfmt_ulongs(C.c_uint64(seed), binding._p(out), C.c_uint32(n))
    return out


def sfmt_floats(seed, n):
    out = np.zeros(n, np.float32)
    api().lib.orc_sfmt_floats(C.c_uint64(seed), binding._p(out), C.c_uint32(n))
    return out


def sfmt_clone_ulongs(seed, skip_parent, n):
    out = np.zeros(n, np.uint64)
    api().lib.orc_sfmt_clone_ulongs(C.c_uint64(seed), C.c_uint32(skip_parent), binding._p(out), C.c_uint32(n))
    return out
