import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import alvrl_loader, orc
pkg = alvrl_loader.load()
for grid in (32, 128, 384):
    scene, vrls, params = pkg.scenes.make_config("C3", width=48, height=48, n_vrls=128, grid=grid)
    params.update(targetNumSlices=8)
    g = pkg.integrator(0, **params); o = orc.Oracle(**params)
    for it in (g, o):
        it.set_scene(scene); it.set_vrls(*vrls); it.build_slices(); it.sample_slice_mapping(); it.build_R()
    Rg, Ro = g.get_R()[..., 0], o.get_R()[..., 0]
    g._call("set_math_mode", pkg.binding.C.c_int(1)); g.build_R(); Rs = g.get_R()[..., 0]
    z_o, z_g = Ro == 0, Rg == 0
    both = ~z_o & ~z_g
    rel = np.abs(Rg - Ro)[both] / np.abs(Ro[both])
    print(f"grid {grid}: entries {Ro.size} oracle zero {z_o.mean():.3f} gpu zero {z_g.mean():.3f} o0&g!0 {(z_o & ~z_g).sum()} o!0&g0 {(~z_o & z_g).sum()} "
          f"both!=0: median rel {np.median(rel):.2e} p99 {np.quantile(rel, 0.99):.2e} frac>1e-4 {(rel > 1e-4).mean():.4f} max {rel.max():.2e}; strict==oracle {np.array_equal(Rs, Ro)}")
    bad = both.copy(); bad[both] = rel > 1e-4
    idx = np.argwhere(bad)[:6]
    for r, v in idx: print("   row", r, "vrl", v, "oracle", Ro[r, v], "gpu", Rg[r, v], "ratio", Rg[r, v] / Ro[r, v])
