"""CPU suite: the occluder compiler (csrc/occluders.h) and its query (csrc/occ_query.h, the function the transport
kernels inline) decide what brute-force triangle tests decide.  The reference's definition of a shadow query is "any
triangle hit inside [mint, maxt]" (Scene::evalTransmittance, src/librender/scene.cpp:619-679, through
ShapeKDTree::rayIntersect, skdtree.cpp:144-204); here it is restated with Moeller-Trumbore in float64."""
import ctypes as C
import math

import numpy as np
import pytest


MAX_SLABS, MAX_PLANES = 24, 24        # csrc/occ_query.h


class OccDev(C.Structure):            # mirrors struct OccDev (csrc/occ_query.h)
    _fields_ = [("slabA", C.c_float * (4 * MAX_SLABS)), ("slabB", C.c_float * (2 * MAX_SLABS)),
                ("planes", C.c_float * (4 * MAX_PLANES)), ("planeInfo", C.c_uint32 * MAX_PLANES),
                ("numBoxes", C.c_uint32), ("numSlabs", C.c_uint32), ("numPlanes", C.c_uint32), ("numTris", C.c_uint32),
                ("cullMargin", C.c_float), ("pad", C.c_float * 3)]     # float4 members: the C++ struct is 16-byte aligned


def compile_occ(host_lib, verts, tris, num_leaves=18):
    verts = np.ascontiguousarray(verts, dtype=np.float32)
    tris = np.ascontiguousarray(tris, dtype=np.uint32)
    counts = np.zeros(6, dtype=np.uint32)
    dev = OccDev()
    recs = np.zeros((3 * 128, 4), dtype=np.float32)
    n = host_lib.alvrl_host_compile_occluders(verts.ctypes.data_as(C.c_void_p), tris.ctypes.data_as(C.c_void_p), C.c_uint32(len(tris)),
                                              C.c_uint32(num_leaves), counts.ctypes.data_as(C.c_void_p), C.byref(dev),
                                              recs.ctypes.data_as(C.c_void_p), C.c_uint32(len(recs)))
    assert n == C.sizeof(OccDev)
    return counts, (dev, recs)


def query(host_lib, counts, occ, o, d, tmin, tmax):
    dev, recs = occ
    o, d = (np.ascontiguousarray(a, dtype=np.float32) for a in (o, d))
    tmin, tmax = (np.ascontiguousarray(a, dtype=np.float32) for a in (tmin, tmax))
    out = np.zeros(len(o), dtype=np.uint8)
    host_lib.alvrl_host_occ_query(C.byref(dev), recs.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p),
                                  d.ctypes.data_as(C.c_void_p), tmin.ctypes.data_as(C.c_void_p), tmax.ctypes.data_as(C.c_void_p),
                                  C.c_uint32(len(o)), out.ctypes.data_as(C.c_void_p))
    return out.astype(bool)


def brute(verts, tris, o, d, tmin, tmax, margin=0.0):
    """(hit, graze): any-hit by Moeller-Trumbore in float64; graze marks segments within `margin` of an edge / range end"""
    v = verts.astype(np.float64)
    o, d = o.astype(np.float64), d.astype(np.float64)
    hit = np.zeros(len(o), dtype=bool)
    graze = np.zeros(len(o), dtype=bool)
    for a, b, c in tris:
        A, e1, e2 = v[a], v[b] - v[a], v[c] - v[a]
        p = np.cross(d, e2)
        det = p @ e1
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            s = o - A
            u = (s * p).sum(1) * inv
            q = np.cross(s, e1)
            w = (d * q).sum(1) * inv
            t = (q @ e2) * inv
        m = margin
        u, w, t = (np.nan_to_num(a, nan=np.inf, posinf=np.inf, neginf=-np.inf) for a in (u, w, t))
        inside = (u >= 0) & (w >= 0) & (u + w <= 1) & (t >= tmin) & (t <= tmax) & (np.abs(det) > 1e-12)
        near = (u >= -m) & (w >= -m) & (u + w <= 1 + m) & (t >= tmin - m) & (t <= tmax + m)
        far = (u >= m) & (w >= m) & (u + w <= 1 - m) & (t >= tmin + m) & (t <= tmax - m) & (np.abs(det) > 1e-6)
        hit |= inside
        graze |= near & ~far
    return hit, graze


def random_segments(rng, n, lo=-0.2, hi=1.2):
    a = rng.uniform(lo, hi, (n, 3))
    b = rng.uniform(lo, hi, (n, 3))
    d = b - a
    L = np.linalg.norm(d, axis=1)
    return a.astype(np.float32), (d / L[:, None]).astype(np.float32), L.astype(np.float32)


def test_cornell_compiles_to_two_boxes_and_five_walls(pkg, host_lib):
    scene, _, _ = pkg.scenes.make_config("C1", width=8, height=8, n_vrls=4)
    counts, (dev, _) = compile_occ(host_lib, scene["verts"], scene["tris"])
    use, slabs, planes, ntris, polys, boxes = (int(x) for x in counts)
    assert (use, slabs, planes, ntris, polys, boxes) == (1, 6, 5, 10, 2, 2)
    assert (dev.numBoxes, dev.numSlabs, dev.numPlanes, dev.numTris) == (2, 6, 5, 10)
    # each box: three slabs with finite lower bounds, unit normals, the last one flagged
    A = np.array(dev.slabA[:4 * slabs]).reshape(slabs, 4)
    B = np.array(dev.slabB[:2 * slabs], dtype=np.float32).reshape(slabs, 2)
    assert np.isfinite(A[:, 3]).all() and (A[:, 3] < B[:, 0]).all()
    assert np.allclose(np.linalg.norm(A[:, :3], axis=1), 1, atol=1e-6)
    assert list(B[:, 1].view(np.uint32)) == [0, 0, 1, 0, 0, 1]


@pytest.mark.parametrize("closed", [False, True])
def test_cornell_queries_match_brute_force(pkg, host_lib, closed):
    scene = pkg.scenes.cornell_scene(8, 8, closed=closed)
    verts, tris = scene["verts"], scene["tris"]
    counts, stream = compile_occ(host_lib, verts, tris)
    assert counts[0] == 1
    rng = np.random.default_rng(5)
    o, d, L = random_segments(rng, 40000)
    tmin = np.zeros_like(L)
    got = query(host_lib, counts, stream, o, d, tmin, L)
    want, graze = brute(verts, tris, o, d, tmin, L, margin=1e-4)
    assert 0.2 < want.mean() < 0.95
    bad = (got != want) & ~graze
    assert not bad.any(), (int(bad.sum()), o[bad][:3], d[bad][:3], L[bad][:3])
    assert graze.mean() < 0.01
    # segments that start ON a surface with the adaptive epsilon of the shadow-ray overload (skdtree.cpp:154-157)
    v = verts.astype(np.float64)
    t = tris[rng.integers(0, len(tris), 20000)]
    bary = rng.dirichlet((1, 1, 1), len(t))
    p = (v[t[:, 0]] * bary[:, :1] + v[t[:, 1]] * bary[:, 1:2] + v[t[:, 2]] * bary[:, 2:]).astype(np.float32)
    q = rng.uniform(0.02, 0.98, (len(t), 3)).astype(np.float32)
    dd = q - p
    L = np.linalg.norm(dd, axis=1).astype(np.float32)
    dd = (dd / L[:, None]).astype(np.float32)
    tmin = (1e-4 * np.abs(p).max(axis=1)).astype(np.float32)
    got = query(host_lib, counts, stream, p, dd, tmin, L)
    want, graze = brute(verts, tris, p, dd, tmin, L, margin=2e-4)
    bad = (got != want) & ~graze
    assert not bad.any(), int(bad.sum())


def _tetra(offset, s=0.3):
    v = np.array([[0, 0, 0], [s, 0, 0], [0, s, 0], [0, 0, s]], dtype=np.float32) + np.asarray(offset, dtype=np.float32)
    t = np.array([[0, 2, 1], [0, 1, 3], [0, 3, 2], [1, 2, 3]], dtype=np.uint32)
    return v, t


def test_general_polytopes_open_and_nonconvex_parts(pkg, host_lib):
    """a tetrahedron (4 lone half-spaces), an L-shaped non-convex closed prism (falls back to planar groups), a lone quad"""
    vt, tt = _tetra((0.1, 0.1, 0.1))
    # L-shaped prism: closed but not convex
    poly = [(0, 0), (0.4, 0), (0.4, 0.15), (0.15, 0.15), (0.15, 0.4), (0, 0.4)]
    base, top = 0.55, 0.8
    vl = np.array([(x + 0.5, base, y + 0.5) for x, y in poly] + [(x + 0.5, top, y + 0.5) for x, y in poly], dtype=np.float32)
    n = len(poly)
    tl = []
    for i in range(n):
        j = (i + 1) % n
        tl += [(i, j, n + j), (i, n + j, n + i)]
    fan = [(0, 1, 2), (0, 2, 3), (0, 3, 4), (0, 4, 5)]
    tl += [(a, c, b) for a, b, c in fan] + [(n + a, n + b, n + c) for a, b, c in fan]
    tl = np.array(tl, dtype=np.uint32)
    vq = np.array([[0.2, 0.9, 0.2], [0.8, 0.9, 0.2], [0.8, 0.9, 0.8], [0.2, 0.9, 0.8]], dtype=np.float32)
    tq = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    verts = np.concatenate([vt, vl, vq])
    tris = np.concatenate([tt, tl + len(vt), tq + len(vt) + len(vl)])
    counts, stream = compile_occ(host_lib, verts, tris, num_leaves=32)
    use, slabs, planes, ntris, polys, boxes = (int(x) for x in counts)
    assert use == 1 and polys == 1 and slabs == 4 and boxes == 0       # only the tetrahedron is a convex solid
    assert np.isinf(np.array(stream[0].slabA[:16]).reshape(4, 4)[:, 3]).all()   # lone half-spaces: c_lo = -inf
    assert ntris == len(tl) + len(tq) and planes == 6 + 2 + 1          # 6 side planes + 2 caps + the quad
    rng = np.random.default_rng(11)
    o, d, L = random_segments(rng, 60000, 0.0, 1.0)
    tmin = np.zeros_like(L)
    got = query(host_lib, counts, stream, o, d, tmin, L)
    want, graze = brute(verts, tris, o, d, tmin, L, margin=1e-4)
    assert 0.05 < want.mean() < 0.9
    bad = (got != want) & ~graze
    assert not bad.any(), int(bad.sum())


def test_compiler_declines_what_it_cannot_hold(pkg, host_lib):
    # more triangles than the shared-memory budget
    v, t = _tetra((0, 0, 0))
    verts = np.concatenate([v + i for i in range(40)])
    tris = np.concatenate([t + 4 * i for i in range(40)])
    counts, stream = compile_occ(host_lib, verts, tris)
    assert counts[0] == 0
    # cheaper as a flat sweep: 30 separate triangles in 2 leaves
    rng = np.random.default_rng(3)
    verts = rng.uniform(0, 1, (90, 3)).astype(np.float32)
    tris = np.arange(90, dtype=np.uint32).reshape(30, 3)
    counts, _ = compile_occ(host_lib, verts, tris, num_leaves=2)
    assert counts[0] == 0
    # degenerate triangles are dropped, a segment inside a solid does not touch its boundary
    v, t = _tetra((0, 0, 0), s=1.0)
    t = np.concatenate([t, [[0, 0, 1]]]).astype(np.uint32)
    counts, stream = compile_occ(host_lib, v, t, num_leaves=8)
    assert counts[0] == 1 and counts[4] == 1 and counts[3] == 0
    o = np.array([[0.1, 0.1, 0.1], [0.1, 0.1, 0.1], [-1, 0.2, 0.2]], dtype=np.float32)
    d = np.array([[1, 0, 0], [1, 0, 0], [1, 0, 0]], dtype=np.float32)
    got = query(host_lib, counts, stream, o, d, np.zeros(3, np.float32), np.array([0.3, 5.0, 0.5], np.float32))
    assert list(got) == [False, True, False]


def test_mixed_solids_boxes_first_and_axis_parallel_segments(pkg, host_lib):
    """a tetrahedron, a rotated box and an axis-aligned box (boxes are ordered first), a slanted and an axis-aligned quad"""
    m = pkg.scenes._Mesh()
    m.box((0.3, 0.3, 0.7), (0.1, 0.15, 0.1), 30.0, 0)
    m.box((0.7, 0.2, 0.3), (0.1, 0.2, 0.1), 0.0, 0)
    m.quad((0.1, 0.8, 0.1), (0.9, 0.9, 0.1), (0.9, 0.9, 0.9), (0.1, 0.8, 0.9), 0)     # slanted
    m.quad((0.0, 0.0, 1.0), (0.0, 1.0, 1.0), (1.0, 1.0, 1.0), (1.0, 0.0, 1.0), 0)     # z = 1
    vb, tb, _ = m.arrays()
    vt, tt = _tetra((0.1, 0.4, 0.1), s=0.25)
    verts = np.concatenate([vt, vb])
    tris = np.concatenate([tt, tb + len(vt)])
    counts, occ = compile_occ(host_lib, verts, tris, num_leaves=32)
    use, slabs, planes, ntris, polys, boxes = (int(x) for x in counts)
    assert (use, slabs, planes, ntris, polys, boxes) == (1, 10, 2, 4, 3, 2)
    dev = occ[0]
    assert np.isfinite(np.array(dev.slabA[:4 * 6]).reshape(6, 4)[:, 3]).all()          # the two boxes come first
    assert np.isinf(np.array(dev.slabA[4 * 6:4 * 10]).reshape(4, 4)[:, 3]).all()       # then the tetrahedron's half-spaces
    rng = np.random.default_rng(21)
    o, d, L = random_segments(rng, 60000, 0.0, 1.0)
    tmin = np.zeros_like(L)
    got = query(host_lib, counts, occ, o, d, tmin, L)
    want, graze = brute(verts, tris, o, d, tmin, L, margin=1e-4)
    assert 0.05 < want.mean() < 0.9
    bad = (got != want) & ~graze
    assert not bad.any(), int(bad.sum())
    # axis-parallel segments (a zero direction component: reciprocal = inf) behave
    o = rng.uniform(0, 1, (20000, 3)).astype(np.float32)
    d = np.zeros((20000, 3), dtype=np.float32)
    d[np.arange(20000), rng.integers(0, 3, 20000)] = rng.choice([-1.0, 1.0], 20000)
    L = rng.uniform(0.05, 1.0, 20000).astype(np.float32)
    got = query(host_lib, counts, occ, o, d, np.zeros_like(L), L)
    want, graze = brute(verts, tris, o, d, np.zeros_like(L), L, margin=1e-4)
    bad = (got != want) & ~graze
    assert not bad.any(), int(bad.sum())


def test_pair_level_culling_never_changes_a_decision(host_lib, pkg, orc):
    """occ_query.h, pair-level culling: the side bits of a camera segment [E, Usurf] and of a VRL [S, End] decide which boxes
    and planes the pair's shadow segments still have to test.  On real camera segments of the Cornell scene (hit points from
    the oracle's primary rays) x synthetic VRLs, 1.5 M random shadow segments (vol->vol from a point of the camera segment,
    vol->surf from the hit point with the adaptive epsilon) give the same answer with and without the culled tests, and the
    culling removes most of the plane tests and a good part of the box tests."""
    from conftest import small_case, setup
    scene, vrls, params = small_case(pkg, "C1", 48, 48, 400)
    counts, occ = compile_occ(host_lib, scene["verts"], scene["tris"])
    assert counts[0] == 1
    o = setup(orc.Oracle(**params), scene, vrls)
    prim, t, pos, nrm = o.primary_hits()
    hit = prim != 0xFFFFFFFF
    rng = np.random.default_rng(11)
    n = 150_000
    pix = rng.choice(np.flatnonzero(hit), n)
    vi = rng.integers(0, len(vrls[0]), n)
    E = np.broadcast_to(scene["camera"]["origin"], (n, 3)).astype(np.float32).copy()
    U = np.ascontiguousarray(pos[pix], dtype=np.float32)
    S = np.ascontiguousarray(vrls[0][vi], dtype=np.float32)
    En = np.ascontiguousarray(vrls[1][vi], dtype=np.float32)
    # VRLs that start or end exactly on a wall (what a tracer produces), and VRLs inside a box
    S[: n // 20, 1] = 0.0
    En[n // 20: n // 10, 0] = 1.0
    stats = np.zeros(6, np.uint64)
    dev, recs = occ
    ext = float(np.ptp(scene["verts"], axis=0).max())
    host_lib.alvrl_host_pair_cull_check(C.byref(dev), recs.ctypes.data_as(C.c_void_p), E.ctypes.data_as(C.c_void_p), U.ctypes.data_as(C.c_void_p),
                                        S.ctypes.data_as(C.c_void_p), En.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(10),
                                        C.c_uint64(3), C.c_float(1e-5 * ext), stats.ctypes.data_as(C.c_void_p))
    mism, boxC, boxT, plC, plT, occl = (int(x) for x in stats)
    print(f"pair culling: {mism} mismatches in {n * 10} segments ({occl} occluded); boxes culled {boxC / boxT:.2%}, planes culled {plC / plT:.2%}")
    assert mism == 0
    assert occl > 0.05 * n * 10
    assert plC / plT > 0.7 and boxC / boxT > 0.2


def _rotation(rng):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _random_box(rng):
    """an arbitrarily rotated box: 8 corners, 12 outward triangles"""
    half = rng.uniform(0.04, 0.22, 3)
    c = rng.uniform(0.2, 0.8, 3)
    R = _rotation(rng) if rng.random() < 0.7 else np.eye(3)
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float64) * half
    v = (corners @ R.T + c).astype(np.float32)
    # corner index = 4 * (x > 0) + 2 * (y > 0) + (z > 0); faces wound counter-clockwise seen from outside
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    t = np.array([f for a, b, cc, d in quads for f in ((a, b, cc), (a, cc, d))], np.uint32)
    return v, t


def _random_quad(rng):
    o = rng.uniform(0.05, 0.95, 3)
    R = _rotation(rng)
    e1, e2 = R[:, 0] * rng.uniform(0.1, 0.5), R[:, 1] * rng.uniform(0.1, 0.5)
    v = np.array([o, o + e1, o + e1 + e2, o + e2], np.float32)
    return v, np.array([(0, 1, 2), (0, 2, 3)], np.uint32)


def test_random_boxes_and_quads_match_brute_force(host_lib):
    """hypothesis over scene compositions: up to four arbitrarily rotated boxes (which may overlap each other) and up to four free
    quads.  Whatever the compiler makes of them -- boxes, general polytopes, planar groups, or nothing (it may decline, then the
    kernels traverse the tree instead) -- the compiled set answers 6 000 random segments like brute-force triangle tests, outside
    a 1e-4 grazing margin."""
    from hypothesis import given, settings, strategies as st
    stats = {"compiled": 0, "declined": 0}

    @settings(max_examples=120, deadline=None, derandomize=True, database=None)          # (600 examples: also clean, 32 s)
    @given(st.integers(0, 2**31 - 1), st.integers(0, 4), st.integers(0, 4))
    def check(seed, n_boxes, n_quads):
        if n_boxes + n_quads == 0:
            return
        rng = np.random.default_rng(seed)
        vs, ts, base = [], [], 0
        for k in range(n_boxes + n_quads):
            v, t = _random_box(rng) if k < n_boxes else _random_quad(rng)
            vs.append(v); ts.append(t + base); base += len(v)
        verts, tris = np.concatenate(vs), np.concatenate(ts).astype(np.uint32)
        counts, stream = compile_occ(host_lib, verts, tris, num_leaves=max(1, len(tris) // 2))
        if counts[0] != 1:
            stats["declined"] += 1
            return
        stats["compiled"] += 1
        o, d, L = random_segments(rng, 6000)
        tmin = np.zeros_like(L)
        got = query(host_lib, counts, stream, o, d, tmin, L)
        want, graze = brute(verts, tris, o, d, tmin, L, margin=1e-4)
        bad = (got != want) & ~graze
        assert not bad.any(), (seed, n_boxes, n_quads, int(bad.sum()), o[bad][:2], d[bad][:2], L[bad][:2])
        assert graze.mean() < 0.02
    check()
    assert stats["compiled"] >= 50, stats


def test_pair_level_culling_on_random_scenes(host_lib):
    """the pair-level culling (side bits of camera segment and VRL decide which boxes and planes a pair's shadow segments still
    test) on random compositions of rotated boxes and quads: camera segments from random eyes to random points ON the surfaces,
    random VRLs, some starting on a surface -- culled and unculled queries agree on every shadow segment"""
    from hypothesis import given, settings, strategies as st
    tot = {"scenes": 0, "segments": 0, "occluded": 0, "box_culled": 0, "box_tests": 0, "plane_culled": 0, "plane_tests": 0}

    @settings(max_examples=50, deadline=None, derandomize=True, database=None)
    @given(st.integers(0, 2**31 - 1), st.integers(0, 4), st.integers(0, 4))
    def check(seed, n_boxes, n_quads):
        if n_boxes + n_quads == 0:
            return
        rng = np.random.default_rng(seed)
        vs, ts, base = [], [], 0
        for k in range(n_boxes + n_quads):
            v, t = _random_box(rng) if k < n_boxes else _random_quad(rng)
            vs.append(v); ts.append(t + base); base += len(v)
        verts, tris = np.concatenate(vs), np.concatenate(ts).astype(np.uint32)
        counts, (dev, recs) = compile_occ(host_lib, verts, tris, num_leaves=max(1, len(tris) // 2))
        if counts[0] != 1:
            return
        n = 8000
        v64 = verts.astype(np.float64)
        t = tris[rng.integers(0, len(tris), n)]
        bary = rng.dirichlet((1, 1, 1), n)
        U = (v64[t[:, 0]] * bary[:, :1] + v64[t[:, 1]] * bary[:, 1:2] + v64[t[:, 2]] * bary[:, 2:]).astype(np.float32)
        E = rng.uniform(-0.3, 1.3, (n, 3)).astype(np.float32)
        S = rng.uniform(-0.1, 1.1, (n, 3)).astype(np.float32)
        En = rng.uniform(-0.1, 1.1, (n, 3)).astype(np.float32)
        S[: n // 10] = U[rng.permutation(n)[: n // 10]]                      # VRLs that start on a surface, as a tracer's do
        stats = np.zeros(6, np.uint64)
        ext = float(np.ptp(verts, axis=0).max())
        host_lib.alvrl_host_pair_cull_check(C.byref(dev), recs.ctypes.data_as(C.c_void_p), E.ctypes.data_as(C.c_void_p), U.ctypes.data_as(C.c_void_p),
                                            S.ctypes.data_as(C.c_void_p), En.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(6),
                                            C.c_uint64(seed & 0xFFFF), C.c_float(1e-5 * ext), stats.ctypes.data_as(C.c_void_p))
        mism, boxC, boxT, plC, plT, occl = (int(x) for x in stats)
        assert mism == 0, (seed, n_boxes, n_quads, mism)
        tot["scenes"] += 1; tot["segments"] += 6 * n; tot["occluded"] += occl
        tot["box_culled"] += boxC; tot["box_tests"] += boxT; tot["plane_culled"] += plC; tot["plane_tests"] += plT
    check()
    print(tot)
    assert tot["scenes"] >= 20 and tot["occluded"] > 0.02 * tot["segments"]
    assert tot["box_culled"] > 0 and tot["plane_culled"] > 0                 # the culling is exercised, not vacuous
