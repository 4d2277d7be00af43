"""GPU probe: wall-clock time of every C-ABI call of one C2 frame next to the library's own phase timers."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import alvrl_loader  # noqa: E402
pkg = alvrl_loader.load()
import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2"); ap.add_argument("--width", type=int, default=None); ap.add_argument("--height", type=int, default=None)
ap.add_argument("--vrls", type=int, default=None); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--slice-range", type=int, nargs=2, default=None, help="what one rank of a multi-GPU job owns")
a = ap.parse_args()
scene, vrls, params = pkg.scenes.make_config(a.config, width=a.width, height=a.height, n_vrls=a.vrls)
g = pkg.integrator(0, **params)
g.set_scene(scene); g.set_vrls(*vrls)
for r in range(a.reps):
    t = {}
    t0 = time.time()
    def lap(name):
        global t0
        now = time.time(); t[name] = (now - t0) * 1e3; t0 = now
    g.build_slices(); lap("build_slices")
    if a.slice_range: g.set_slice_range(*a.slice_range)
    g.sample_slice_mapping(); lap("slice_mapping")
    g.build_R(); lap("build_R")
    g.build_clusters(); lap("build_clusters")
    img = g.render(); lap("render")
    st = g.stats()
    print(f"rep{r}: " + " ".join(f"{k}={v:.0f}" for k, v in t.items()) + f" total={sum(t.values()):.0f} ms | lib: slices={st.msSlices:.0f} R={st.msBuildR:.0f} "
          f"(kernel {st.msTransportKernelR:.0f}) clusters={st.msClusters:.0f} render={st.msRender:.0f} (kernel {st.msTransportKernelRender:.0f}) launches={st.kernelLaunches}")
