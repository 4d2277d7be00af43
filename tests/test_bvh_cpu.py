"""CPU suite: the host builder of the device BVH (csrc/bvh.h) -- the threaded binary tree that primary rays and the strict
flavour walk, and the 4-wide tree collapsed from it for the fast flavour's any-hit query on large scenes (the C4 path).  The
reference's accelerator is a SAH kd-tree (include/mitsuba/render/sahkdtree3.h); semantics, not structure, have to match: no
triangle a ray hits may be culled.  Checked through libalvrl_host.so: the structure (preorder, escape indices, leaf ranges that
partition the triangle order, boxes that hold their triangles and their children) and, ray by ray, that every triangle a
brute-force test hits sits in a leaf the traversal reaches."""
import ctypes as C

import numpy as np
import pytest


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def bvh_check(host_lib, verts, tris, n_rays, seed, lo=-0.3, hi=1.3):
    verts = np.ascontiguousarray(verts, np.float32); tris = np.ascontiguousarray(tris, np.uint32)
    rng = np.random.default_rng(seed)
    a, b = rng.uniform(lo, hi, (n_rays, 3)), rng.uniform(lo, hi, (n_rays, 3))
    # a share of axis-parallel rays (zero direction components) and of rays that start on a vertex
    k = n_rays // 8
    b[:k, 0] = a[:k, 0]; b[k:2 * k, 1] = a[k:2 * k, 1]; b[k:2 * k, 2] = a[k:2 * k, 2]
    a[2 * k:3 * k] = verts[rng.integers(0, len(verts), k)]
    d = b - a
    L = np.linalg.norm(d, axis=1)
    o, d, L = a.astype(np.float32), (d / L[:, None]).astype(np.float32), L.astype(np.float32)
    stats = np.zeros(8, np.uint64)
    assert host_lib.alvrl_host_bvh_check(_p(verts), _p(tris), C.c_uint32(len(tris)), _p(o), _p(d), _p(L), C.c_uint32(n_rays), _p(stats)) == 0
    return dict(zip(("error", "hits", "missed2", "missed4", "visits2", "visits4", "nodes2", "nodes4"), (int(x) for x in stats)))


@pytest.mark.parametrize("name,kw", [("C1", {}), ("C4", dict(occluders=40))], ids=["cornell", "icospheres"])
def test_bvh_of_the_config_scenes(pkg, host_lib, name, kw):
    scene, _, _ = pkg.scenes.make_config(name, width=8, height=8, n_vrls=4, **kw)
    n = 3000 if name == "C1" else 400
    r = bvh_check(host_lib, scene["verts"], scene["tris"], n, 1, lo=0.0 if name == "C4" else -0.3, hi=1.0 if name == "C4" else 1.3)
    print(name, len(scene["tris"]), "triangles:", r)
    assert r["error"] == 0 and r["missed2"] == 0 and r["missed4"] == 0 and r["hits"] > n // 4
    assert r["visits4"] < r["visits2"]                                       # the collapsed tree visits fewer nodes
    assert r["nodes2"] <= 2 * len(scene["tris"]) and r["nodes4"] < r["nodes2"]


def test_bvh_property_based(host_lib):
    """hypothesis over triangle soups: clustered and scattered triangles, slivers, zero-area triangles, duplicates, coincident
    centroids, axis-aligned walls much larger than the rest"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=80, deadline=None, derandomize=True, database=None)
    @given(st.integers(0, 2**31 - 1), st.integers(1, 400), st.sampled_from(["scattered", "clustered", "degenerate", "walls"]))
    def check(seed, nt, kind):
        rng = np.random.default_rng(seed)
        c = rng.uniform(0, 1, (nt, 1, 3))
        if kind == "clustered":
            c = 0.5 + 0.02 * rng.normal(size=(nt, 1, 3))
        size = rng.choice([0.002, 0.02, 0.2], (nt, 1, 1))
        v = c + size * rng.normal(size=(nt, 3, 3))
        if kind == "degenerate":
            v[: nt // 3, 2] = v[: nt // 3, 1]                                  # zero-area triangles
            v[nt // 3: nt // 2] = v[0]                                       # duplicates of the first triangle (coincident centroids)
            v[nt // 2: 2 * nt // 3, :, 0] = 0.5                              # coplanar slivers in the plane x = 0.5
        if kind == "walls" and nt >= 4:
            v[0] = [(0, 0, 0), (1, 0, 0), (1, 0, 1)]; v[1] = [(0, 0, 0), (1, 0, 1), (0, 0, 1)]          # a floor under everything
            v[2] = [(0, 0, 1), (1, 0, 1), (1, 1, 1)]; v[3] = [(0, 0, 1), (1, 1, 1), (0, 1, 1)]          # a back wall
        verts = v.reshape(-1, 3).astype(np.float32)
        tris = np.arange(3 * nt, dtype=np.uint32).reshape(nt, 3)
        r = bvh_check(host_lib, verts, tris, 300, seed ^ 0x5bd1)
        assert r["error"] == 0 and r["missed2"] == 0 and r["missed4"] == 0, (seed, nt, kind, r)
    check()
