"""Generates tests/golden/walks_tiny.npz with the CPU oracle: a traced VRL set (vrlTracer.h) and a ground-truth image (volpath with
onlyVRLpaths) of the glass + conductor scene.  Like c1_tiny.npz these fixtures pin the oracle against regressions (the
reference itself cannot be built in this image).  Run from the repo root."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc  # noqa: E402

pkg = orc._pkg
W, H, seed, target = 16, 12, 31, 120
scene, em, rad = pkg.scenes.tracer_scene(W, H, glass=True)
o = orc.Oracle(volVolSamples=2, volSurfSamples=2, targetNumSlices=4, seed=seed, vrlTargetNum=target)
o.set_scene(scene)
o.set_area_emitter(em, rad)
o.trace_vrls()
s, e, p, pc = o.get_vrls()
img = o.volpath_render(spp=3, internal_samples=2)
img_all = o.volpath_render(spp=2, internal_samples=1, flags=0, max_depth=5)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "walks_tiny.npz"), width=W, height=H, seed=seed, target=target,
                    vrl_start=s, vrl_end=e, vrl_power=p, particles=pc, volpath=img, volpath_all_depth5=img_all)
print("wrote walks_tiny.npz:", len(s), "VRLs from", pc, "particles; image mean", float(img.mean()))
