"""Specular chains (vrlIntegrator.cpp:445-511): camera segments that end on a smooth dielectric or conductor
continue along the delta components, with Russian roulette, medium transitions and per-segment weights.

CPU part: the oracle's restatement against what the construction implies.  GPU part: the device chains against the oracle
bit for bit, the rows of R and the clustered render with chains against the oracle."""
import numpy as np
import pytest

from conftest import setup


def _scene(pkg, W=40, H=40, n_vrls=48):
    scene = pkg.scenes.chain_scene(W, H)
    vrls = pkg.scenes.synthetic_vrls(n_vrls, sigma_t=1.05, seed=77)
    params = dict(volVolSamples=2, volSurfSamples=2, targetNumSlices=6, seed=11)
    return scene, vrls, params


def test_oracle_chain_structure(pkg, orc):
    scene, vrls, params = _scene(pkg)
    o = setup(orc.Oracle(**params), scene, vrls)
    off, segs = o.chain_segments()
    n = np.diff(off)
    prim = o.primary_hits()[0]
    mats = scene["tri_material"]
    hit = prim != 0xFFFFFFFF
    delta = np.zeros(len(prim), bool)
    delta[hit] = mats[prim[hit]] >= 3
    assert (n[~delta] == 0).all()                      # only pixels that see the glass sphere or the mirror grow a chain
    assert (n[delta] >= 1).mean() > 0.9 and n.max() >= 3
    # every child ray starts at the hit point of its parent; the first one at the camera segment's hit point
    codes = segs[:, 14].astype(np.int64)
    assert (codes >= 2).all()
    for pix in np.flatnonzero(n)[:50]:
        s = segs[off[pix]:off[pix + 1]]
        byCode = {int(c): r for c, r in zip(s[:, 14], s)}
        for c, r in byCode.items():
            par = c // 2
            if par in byCode:
                assert np.array_equal(r[0:3], byCode[par][6:9])      # origin == the parent's hit point
        assert np.isfinite(s[:, 10:13]).all() and (s[:, 10:13] >= 0).all()
    # inside the glass there is no medium (interior = vacuum): the transmitted segment below a glass hit is flagged out of it
    glass_first = [segs[off[p]:off[p + 1]] for p in np.flatnonzero(delta & (n >= 2))[:200]]
    assert any(((s[:, 14] == 3) & (s[:, 13] == 0)).any() for s in glass_first)
    # the mirror keeps the medium and has one component
    assert ((segs[:, 13] == 1).sum() > 0)


def test_oracle_chain_adds_radiance_behind_glass(pkg, orc):
    """with the chains the pixels that see glass / mirror receive the medium's radiance from behind the surface"""
    scene, vrls, params = _scene(pkg, 32, 32, 32)
    o = setup(orc.Oracle(**params), scene, vrls)
    o.build_slices(); o.prepass()
    img = o.render()
    plain = dict(scene); plain["mat_bits"] = np.where(scene["mat_bits"] & 6, np.uint32(0), scene["mat_bits"]).astype(np.uint32)
    o2 = setup(orc.Oracle(**params), plain, vrls)
    o2.build_slices(); o2.prepass()
    img2 = o2.render()
    off, _ = o.chain_segments()
    has = (np.diff(off) > 0).reshape(32, 32).T          # pixel index = y + H * x
    assert img[has].sum() > 1.05 * img2[has].sum()
    assert np.array_equal(img[~has] > 0, img2[~has] > 0)


@pytest.mark.gpu
def test_chain_segments_bit_exact(pkg, orc):
    scene, vrls, params = _scene(pkg, 64, 64)
    g = setup(pkg.integrator(0, **params), scene, vrls)
    o = setup(orc.Oracle(**params), scene, vrls)
    og, sg = g.chain_segments()
    oo, so = o.chain_segments()
    assert np.array_equal(og, oo)
    assert len(sg) > 500
    assert np.array_equal(sg, so), np.abs(sg - so).max(0)


@pytest.mark.gpu
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_R_rows_with_chains_vs_oracle(pkg, orc, strict):
    scene, vrls, params = _scene(pkg, 48, 48, 64)
    g = setup(pkg.integrator(0, **params), scene, vrls)
    g._call("set_math_mode", pkg.binding.C.c_int(1 if strict else 0))
    o = setup(orc.Oracle(**params), scene, vrls)
    for it in (g, o):
        it.build_slices(); it.sample_slice_mapping(); it.build_R()
    assert np.array_equal(g.rep_pixels()[1], o.rep_pixels()[1])
    Rg, Ro = g.get_R(), o.get_R()
    off, _ = o.chain_segments()
    rows_with_chain = np.diff(off)[o.rep_pixels()[1]] > 0
    assert rows_with_chain.sum() >= 3
    floor = 1e-12 * np.abs(Ro[..., 0]).max()
    em = np.abs(Rg[..., 0] - Ro[..., 0]) / (np.abs(Ro[..., 0]) + floor)
    ev = np.abs(Rg[..., 1] - Ro[..., 1]) / (Ro[..., 1] + Ro[..., 0] ** 2 + floor * floor)
    bad = (em > 1e-4) | (ev > 1e-4)
    print("rows with chains", int(rows_with_chain.sum()), "of", len(rows_with_chain), "bad entries", int(bad.sum()), "of", bad.size,
          "median", float(np.median(em)))
    assert bad.mean() <= (2e-5 if strict else 1e-3), (int(bad.sum()), bad.size)
    # the chain really contributes: rows with a chain differ from the same rows of a chain-less scene
    assert (Ro[rows_with_chain, :, 0].sum() > 0)


@pytest.mark.gpu
def test_render_with_chains_vs_oracle(pkg, orc):
    """clustered render with the oracle's clusters: per pixel 1e-3 (strict flavour), pixels with chains included"""
    scene, vrls, params = _scene(pkg, 40, 40, 48)
    o = setup(orc.Oracle(**params), scene, vrls)
    o.build_slices(); o.prepass()
    g = setup(pkg.integrator(0, **params), scene, vrls)
    g._call("set_math_mode", pkg.binding.C.c_int(1))
    g.build_slices(); g.sample_slice_mapping(); g.build_R()
    g.set_clusters(o.clusters())
    ig, io = g.render(), o.render()
    off, _ = o.chain_segments()
    has = (np.diff(off) > 0).reshape(40, 40).T
    assert has.sum() > 50
    scale = io.max()
    err = np.abs(ig - io) / (np.abs(io) + 1e-4 * scale)
    print("max rel err", float(err.max()), "with chain", float(err[has].max()), "pixels with chains", int(has.sum()))
    assert err.max() < 1e-3
