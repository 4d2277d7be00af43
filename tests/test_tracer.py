"""VRL tracer (src/integrators/vrl/vrlTracer.h:14-58, 91-230): light particles from an area emitter through the
medium and off the surfaces, every in-medium path segment a VRL, until vrlTargetNum VRLs exist.

CPU part: the oracle's restatement against what the construction implies.  GPU part: the device tracer against the oracle
bit for bit, and the traced set through the whole path."""
import numpy as np
import pytest

from conftest import setup


def _make(pkg, cls, glass, target, seed=5, **extra):
    scene, em, rad = pkg.scenes.tracer_scene(40, 40, glass=glass)
    params = dict(volVolSamples=2, volSurfSamples=2, targetNumSlices=6, seed=seed, vrlTargetNum=target)
    params.update(extra)
    it = cls(**params)
    it.set_scene(scene)
    it.set_area_emitter(em, rad)
    return it, scene


@pytest.mark.parametrize("glass", [False, True], ids=["cornell", "glass+conductor"])
def test_oracle_tracer_properties(pkg, orc, glass):
    o, scene = _make(pkg, orc.Oracle, glass, 400)
    o.trace_vrls()
    s, e, p, pc = o.get_vrls()
    assert len(s) >= 400 and 0 < pc <= len(s)                     # the loop stops after the particle that reaches the target
    assert np.isfinite(s).all() and np.isfinite(e).all() and (p >= 0).all() and (p.max(1) > 0).all()
    assert (np.linalg.norm(e - s, axis=1) > 0).all()               # vrlVector::put drops zero-length VRLs
    lo, hi = -1e-4, 1 + 1e-4
    inside = ((s > lo) & (s < hi)).all(1)
    assert inside.all()                                            # the box is closed: the walk stays inside
    # the first VRL of the first particle starts on the light, with the emitter's power (radiance * pi * area, area.cpp:198)
    assert abs(s[0, 1] - 0.998) < 1e-6 and 0.35 <= s[0, 0] <= 0.65
    assert np.allclose(p[0], np.array([18.0, 15.0, 12.0]) * np.float32(np.pi) * 0.09, rtol=1e-5)
    # short VRLs end at the scattering point: consecutive VRLs of a particle chain up
    chained = (np.abs(s[1:] - e[:-1]).max(1) == 0).mean()
    assert chained > 0.5
    # deterministic, and a prefix property: a smaller target gives a prefix of the same particle sequence
    o2, _ = _make(pkg, orc.Oracle, glass, 150)
    o2.trace_vrls()
    s2, e2, p2, pc2 = o2.get_vrls()
    assert pc2 <= pc and np.array_equal(s2, s[:len(s2)]) and np.array_equal(p2, p[:len(s2)])


def test_oracle_tracer_long_vrls_reach_the_surface(pkg, orc):
    o, scene = _make(pkg, orc.Oracle, False, 200, shortVrls=0)
    o.trace_vrls()
    s, e, p, pc = o.get_vrls()
    on_wall = (np.abs(e - np.round(e)).min(1) < 2e-3) | (np.abs(e[:, 1] - 0.998) < 1e-3)
    assert on_wall.mean() > 0.6                                    # long VRLs end on the next surface (vrlTracer.h:150-160); boxes aside


@pytest.mark.gpu
@pytest.mark.parametrize("glass,extra", [(False, {}), (True, {}), (False, dict(shortVrls=0)), (True, dict(rrDepth=2, maxParticleDepth=12))],
                         ids=["cornell", "glass+conductor", "long-vrls", "rr2-depth12"])
def test_tracer_matches_oracle_bit_exact(pkg, orc, glass, extra):
    g, scene = _make(pkg, lambda **kw: pkg.integrator(0, **kw), glass, 3000, **extra)
    o, _ = _make(pkg, orc.Oracle, glass, 3000, **extra)
    g.trace_vrls(); o.trace_vrls()
    sg, eg, pg, pcg = g.get_vrls()
    so, eo, po, pco = o.get_vrls()
    assert pcg == pco and len(sg) == len(so) >= 3000
    assert np.array_equal(sg, so) and np.array_equal(eg, eo)
    assert np.array_equal(pg, po), float(np.abs(pg - po).max())


@pytest.mark.gpu
def test_traced_vrls_through_the_path(pkg, orc):
    """tracer -> slices -> R -> clusters -> render on the device; the same VRL set in the oracle gives the same clusters"""
    g, scene = _make(pkg, lambda **kw: pkg.integrator(0, **kw), False, 600)
    g._call("set_math_mode", pkg.binding.C.c_int(1))
    g.trace_vrls()
    g.build_slices(); g.prepass()
    img = g.render()
    assert np.isfinite(img).all() and img.max() > 0
    o, _ = _make(pkg, orc.Oracle, False, 600)
    o.trace_vrls()
    o.build_slices(); o.prepass()
    co, cg = o.clusters(), g.clusters()
    assert np.array_equal(co["offset"], cg["offset"]) and np.array_equal(co["vrls"], cg["vrls"])
