"""GPU probe: time buildClusters on a config (ALVRL_PROFILE=1 prints the per-phase split)."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import alvrl_loader
pkg = alvrl_loader.load()
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2"); ap.add_argument("--width", type=int, default=None); ap.add_argument("--height", type=int, default=None)
ap.add_argument("--vrls", type=int, default=None); ap.add_argument("--slice-range", type=int, nargs=2, default=None, help="only these slices (what one rank of a multi-GPU job owns)"); ap.add_argument("--lists", action="store_true", help="also fetch the cluster lists (computes the lazy global / fallback clustering)")
a = ap.parse_args()
scene, vrls, params = pkg.scenes.make_config(a.config, width=a.width, height=a.height, n_vrls=a.vrls)
g = pkg.integrator(0, **params)
g.set_scene(scene); g.set_vrls(*vrls); g.build_slices()
if a.slice_range: g.set_slice_range(*a.slice_range)
g.sample_slice_mapping(); g.build_R()
t = time.time(); g.build_clusters(); print("build_clusters %.3f s" % (time.time() - t))
print("launches", g.stats().kernelLaunches)
if not a.lists: sys.exit(0)
cl_off = g.clusters()["offset"]
import numpy as np
print("clusters per slice: mean %.1f max %d" % (np.diff(cl_off).mean(), np.diff(cl_off).max()), "launches", g.stats().kernelLaunches)
