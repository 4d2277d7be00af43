"""Generates tests/golden/c1_tiny.npz with the CPU oracle (the reference itself cannot be built in this image, so
these fixtures pin the oracle against regressions, not the oracle against the reference).  Run from the repo root."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402

pkg = orc._pkg
W = H = 40
N = 80
seed, S = 21, 16
scene, vrls, params = pkg.scenes.make_config("C1", width=W, height=H, n_vrls=N)
params.update(seed=seed, targetNumSlices=S)
o = orc.Oracle(**params)
o.set_scene(scene)
o.set_vrls(*vrls)
o.build_slices()
o.prepass()
cl = o.clusters()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c1_tiny.npz"), width=W, height=H, n_vrls=N, seed=seed,
                    targetNumSlices=S, prim=o.primary_hits()[0], pixel_to_slice=o.pixel_to_slice(),
                    row_pixel=o.rep_pixels()[1], R=o.get_R(), cluster_offset=cl["offset"], cluster_vrls=cl["vrls"],
                    cluster_weights=cl["weights"], image=o.render())
print("wrote c1_tiny.npz")
