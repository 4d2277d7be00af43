"""GPU probe: the fast flavour of k_build_R against the strict oracle on C2's slice 0 x a VRL prefix; prints the error
distribution and dumps the worst entries (row, vrl, gpu, oracle) to gpurun_out/ for analysis.  The oracle's R is cached in
/tmp so that several kernel variants (ALVRL_LIB=build/libalvrl_<name>.so) can be compared in one session."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import alvrl_loader  # noqa: E402

pkg = alvrl_loader.load()
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--vrls", type=int, default=30000)
ap.add_argument("--tag", default="base")
ap.add_argument("--dump", type=int, default=0)
a = ap.parse_args()

scene, vrls, params = pkg.scenes.make_config(a.config, n_vrls=a.vrls)
cache = f"/tmp/parity_{a.config}_{a.vrls}.npz"
if os.path.exists(cache):
    z = np.load(cache); Ro, graze, rows = z["Ro"], z["graze"], int(z["rows"])
else:
    import orc
    o = orc.Oracle(**params); o.set_scene(scene); o.set_vrls(*vrls); o.set_graze_tolerance(1e-5)
    o.build_slices(); o.set_slice_range(0, 1); o.sample_slice_mapping()
    rows = int(o.rep_pixels()[0][1])
    o.build_R(); Ro, graze = o.get_R(0, rows), o.R_graze(0, rows).astype(bool)
    np.savez(cache, Ro=Ro, graze=graze, rows=rows)
g = pkg.integrator(0, **params)
g.set_scene(scene); g.set_vrls(*vrls); g.build_slices(); g.set_slice_range(0, 1); g.sample_slice_mapping(); g.build_R()
Rg = g.get_R(0, rows)
mo, mg = Ro[..., 0], Rg[..., 0]
floor = 1e-12 * np.abs(mo).max()
em = np.abs(mg - mo) / (np.abs(mo) + floor)
ev = np.abs(Rg[..., 1] - Ro[..., 1]) / (Ro[..., 1] + mo * mo + floor * floor)
clean = ~graze
for name, e in (("mean", em), ("var", ev)):
    print(f"[{a.tag}] {name}: median {np.median(e):.2e} p99.9 {np.quantile(e, 0.999):.2e} p99.99 {np.quantile(e, 0.9999):.2e} "
          f"frac>1e-4 {(e > 1e-4).mean():.3e} unflagged {((e > 1e-4) & clean).sum() / clean.sum():.3e} "
          f"frac>3e-4 unflagged {((e > 3e-4) & clean).sum() / clean.sum():.3e} frac>1e-2 unflagged {((e > 1e-2) & clean).sum() / clean.sum():.3e}")
print(f"[{a.tag}] kernel {g.stats().msTransportKernelR:.2f} ms for {rows} rows x {a.vrls} VRLs")
if a.dump:
    bad = np.argwhere((em > 1e-4) & clean)
    order = np.argsort(-em[bad[:, 0], bad[:, 1]])[: a.dump]
    sel = bad[order]
    off, px = g.rep_pixels()
    np.savez(os.path.join(ROOT, "gpurun_out", f"parity_bad_{a.tag}.npz"), row=sel[:, 0], vrl=sel[:, 1], pixel=px[sel[:, 0]],
             mg=mg[sel[:, 0], sel[:, 1]], mo=mo[sel[:, 0], sel[:, 1]], vg=Rg[sel[:, 0], sel[:, 1], 1], vo=Ro[sel[:, 0], sel[:, 1], 1])
