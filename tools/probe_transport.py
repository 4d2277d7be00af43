"""GPU probe: time the R-build transport kernel on a C2-shaped workload (rows x VRLs), both flavours."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import alvrl_loader  # noqa: E402

pkg = alvrl_loader.load()
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--width", type=int, default=1024)
ap.add_argument("--height", type=int, default=1024)
ap.add_argument("--vrls", type=int, default=100_000)
ap.add_argument("--undersampling", type=float, default=64.0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--strict", type=int, default=0)
ap.add_argument("--grid", type=int, default=None)
ap.add_argument("--occluders", type=int, default=None)
a = ap.parse_args()

scene, vrls, params = pkg.scenes.make_config(a.config, width=a.width, height=a.height, n_vrls=a.vrls, grid=a.grid, occluders=a.occluders)
params["targetPixelUndersampling"] = a.undersampling
g = pkg.integrator(0, **params)
g._call("set_math_mode", pkg.binding.C.c_int(a.strict))
t = time.time(); g.set_scene(scene); g.set_vrls(*vrls); print("upload %.2fs" % (time.time() - t))
t = time.time(); g.build_slices(); print("build_slices %.3fs" % (time.time() - t), g.num_slices())
t = time.time(); g.sample_slice_mapping(); print("slice mapping %.3fs" % (time.time() - t), g.num_slices())
for r in range(a.reps):
    t = time.time(); g.build_R(); wall = time.time() - t
    st = g.stats()
    S, G = g.num_slices()
    pairs = G * g.N
    F = 160 + (136 if params.get("hg") is None else 158) * params["volVolSamples"] + 91 * params["volSurfSamples"]
    print(f"build_R rep{r}: kernel {st.msTransportKernelR:.2f} ms wall {wall*1e3:.1f} ms pairs {pairs:.3e} "
          f"-> {pairs / st.msTransportKernelR / 1e6:.3f} Gpairs/s, {pairs * F / st.msTransportKernelR / 1e9:.2f} algorithmic TFLOP/s")
