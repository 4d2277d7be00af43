"""Analytic shapes at the boundary: `rectangle` and `sphere` handed to the path as triangles (csrc/shapes.h behind
alvrl_add_rectangle / alvrl_add_sphere), on the CPU through libalvrl_host.so -- the geometry against the shapes' definitions
(src/shapes/rectangle.cpp:76-118,170-196; src/shapes/sphere.cpp:108-131,160-251) and, through the oracle's ray caster, against
the analytic intersections."""
import ctypes as C

import numpy as np
import pytest


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def tessellate(host_lib, shape, arg, radius=1.0, flip=False, theta_steps=16):
    a = np.ascontiguousarray(arg, np.float32).reshape(-1)
    nv, nt = C.c_uint32(), C.c_uint32()
    rc = host_lib.alvrl_host_tessellate(shape, _p(a), C.c_float(radius), int(flip), C.c_uint32(theta_steps), None, C.byref(nv), None, C.byref(nt))
    if rc != 0:
        return None, None
    v, t = np.zeros((nv.value, 3), np.float32), np.zeros((nt.value, 3), np.uint32)
    assert host_lib.alvrl_host_tessellate(shape, _p(a), C.c_float(radius), int(flip), C.c_uint32(theta_steps), _p(v), C.byref(nv), _p(t), C.byref(nt)) == 0
    return v, t


def _normals(v, t):
    v = v.astype(np.float64)
    return np.cross(v[t[:, 1]] - v[t[:, 0]], v[t[:, 2]] - v[t[:, 0]])


def _to_world(rot_axis, angle, scale, translate, mirror=False):
    ax = np.asarray(rot_axis, np.float64); ax /= np.linalg.norm(ax)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K
    S = np.diag(scale).astype(np.float64)
    if mirror:
        S = S @ np.diag([-1.0, 1.0, 1.0])
    M = np.eye(4); M[:3, :3] = R @ S; M[:3, 3] = translate
    return M


@pytest.mark.parametrize("mirror", [False, True])
@pytest.mark.parametrize("flip", [False, True])
def test_rectangle_is_the_transformed_square_with_the_shapes_normal(host_lib, flip, mirror):
    M = _to_world((1, 2, 0.5), 0.7, (0.3, 0.8, 1.0), (0.5, 0.25, -1.0), mirror)
    v, t = tessellate(host_lib, 0, M, flip=flip)
    corners = np.array([(-1, -1, 0, 1), (1, -1, 0, 1), (1, 1, 0, 1), (-1, 1, 0, 1)], np.float64)       # rectangle.cpp:179-182
    assert np.allclose(v, (corners @ M.T)[:, :3], atol=1e-6) and len(t) == 2
    assert {tuple(sorted(x)) for x in t.tolist()} == {(0, 1, 2), (0, 2, 3)}                           # 190-196: the two halves
    # the shape's normal: toWorld' (Normal(0, 0, 1)) with toWorld' = toWorld * scale(1, 1, -1) when flipped (82-83, 104):
    # normals transform with the inverse transpose
    Mf = M[:3, :3] @ np.diag([1, 1, -1.0 if flip else 1.0])
    n = np.linalg.inv(Mf).T @ np.array([0, 0, 1.0])
    gn = _normals(v, t)
    assert (gn @ n > 0).all()
    assert np.allclose(gn / np.linalg.norm(gn, axis=1, keepdims=True), n / np.linalg.norm(n), atol=1e-5)
    # the two triangles cover the square exactly: areas add up to |dpdu x dpdv| (102-103)
    area = 0.5 * np.linalg.norm(gn, axis=1).sum()
    assert np.isclose(area, np.linalg.norm(np.cross(M[:3, :3] @ [2, 0, 0], M[:3, :3] @ [0, 2, 0])), rtol=1e-5)


@pytest.mark.parametrize("flip", [False, True])
@pytest.mark.parametrize("T", [3, 4, 16, 64])
def test_sphere_is_a_closed_outward_mesh_with_its_vertices_on_the_sphere(host_lib, T, flip):
    c, r = np.array([0.3, -0.2, 1.5], np.float32), 0.37
    v, t = tessellate(host_lib, 1, c, r, flip, T)
    P = 2 * T
    assert len(v) == 2 + P * (T - 2) and len(t) == 2 * P * (T - 2)
    assert np.allclose(np.linalg.norm(v.astype(np.float64) - c, axis=1), r, rtol=2e-6)
    # closed 2-manifold: every undirected edge in exactly two triangles, every directed edge once; Euler characteristic 2
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]).astype(np.int64)
    directed = e[:, 0] * len(v) + e[:, 1]
    assert len(np.unique(directed)) == len(directed)
    und = np.sort(e, axis=1)
    _, counts = np.unique(und[:, 0] * len(v) + und[:, 1], return_counts=True)
    assert (counts == 2).all()
    assert len(v) - len(counts) + len(t) == 2
    # no degenerate triangle; the geometric normals point away from the centre (towards it with flipNormals, sphere.cpp:248-249)
    gn = _normals(v, t)
    assert (np.linalg.norm(gn, axis=1) > 1e-9 * r * r).all()
    cen = v.astype(np.float64)[t].mean(1) - c
    s = np.einsum("ij,ij->i", gn, cen)
    assert (s < 0).all() if flip else (s > 0).all()
    # the stated deviation of the surface from the sphere
    dev = 1 - np.linalg.norm(cen, axis=1).min() / r
    assert dev <= 1 - np.cos(np.pi / (T - 1)) + 1e-6          # (the chord of one polar AND one azimuthal step)
    if T == 64:
        assert dev < 1.5e-3


def test_bad_shapes_are_refused(host_lib):
    assert tessellate(host_lib, 1, np.zeros(3), 0.0)[0] is None           # "Cannot create spheres of radius <= 0", sphere.cpp:130-131
    assert tessellate(host_lib, 1, np.zeros(3), 1.0, theta_steps=2)[0] is None
    M = np.eye(4); M[1, 1] = 0.0
    assert tessellate(host_lib, 0, M)[0] is None                          # singular toWorld


def test_tessellated_shapes_against_the_analytic_intersections(pkg, orc, host_lib):
    """rays through the oracle's ray caster (Scene::rayIntersect restated) on the tessellated sphere and rectangle against the
    analytic intersections of sphere.cpp:160-186 and rectangle.cpp:120-140: same hit / miss decisions away from the silhouette,
    distances within the stated deviation; the rectangle is exact"""
    c, r = np.array([0.5, 0.45, 0.5], np.float32), 0.2
    sv, st = tessellate(host_lib, 1, c, r, False, 64)
    M = _to_world((0, 1, 0), 0.4, (0.25, 0.15, 1.0), (0.5, 0.1, 0.5))
    rv, rt = tessellate(host_lib, 0, M)
    scene, vrls, params = pkg.scenes.make_config("C1", width=8, height=8, n_vrls=4)
    scene = dict(scene, verts=np.concatenate([sv, rv]), tris=np.concatenate([st, rt + len(sv)]).astype(np.uint32),
                 tri_material=np.zeros(len(st) + len(rt), np.uint32))
    o = orc.Oracle(**params); o.set_scene(scene)
    rng = np.random.default_rng(9)
    n = 20000
    org = rng.uniform(-0.5, 1.5, (n, 3)); org[:, 1] = rng.uniform(0.9, 1.6, n)               # from above
    tgt = np.concatenate([c + rng.normal(0, 0.15, (n // 2, 3)), (M @ np.concatenate([rng.uniform(-1.3, 1.3, (n - n // 2, 2)), np.zeros((n - n // 2, 1)), np.ones((n - n // 2, 1))], 1).T).T[:, :3]])
    d = tgt - org; d /= np.linalg.norm(d, axis=1, keepdims=True)
    org32, d32 = org.astype(np.float32), d.astype(np.float32)
    prim, tt = o.trace_rays(org32, d32, np.zeros(n, np.float32), np.full(n, np.inf, np.float32))[:2]
    org, d = org32.astype(np.float64), d32.astype(np.float64)
    # analytic sphere (sphere.cpp:160-186)
    oc = org - c
    B = 2 * np.einsum("ij,ij->i", oc, d); Cq = np.einsum("ij,ij->i", oc, oc) - r * r
    disc = B * B - 4 * Cq
    ts = np.where(disc > 0, (-B - np.sqrt(np.maximum(disc, 0))) / 2, np.inf)
    ts[ts <= 0] = np.inf
    # analytic rectangle (rectangle.cpp:120-140): the ray in object space meets z = 0 inside [-1, 1]^2
    Mi = np.linalg.inv(M)
    oo = (Mi @ np.concatenate([org, np.ones((n, 1))], 1).T).T[:, :3]; od = (Mi[:3, :3] @ d.T).T
    hit_t = -oo[:, 2] / od[:, 2]
    hp = oo + hit_t[:, None] * od
    inside = (hit_t > 0) & (np.abs(hp[:, 0]) <= 1) & (np.abs(hp[:, 1]) <= 1)
    tr = np.where(inside, hit_t, np.inf)
    want_t = np.minimum(ts, tr)
    hit = prim != pkg.binding.NO_HIT
    # away from the sphere's silhouette (impact parameter below 0.98 r) and the rectangle's edge the decisions agree
    impact = np.sqrt(np.maximum(0, np.einsum("ij,ij->i", oc, oc) - (B / 2) ** 2)) / r
    edge = np.minimum(np.abs(np.abs(hp[:, 0]) - 1), np.abs(np.abs(hp[:, 1]) - 1)) < 1e-4
    clear = ((impact < 0.98) | (impact > 1.0)) & ~edge
    assert clear.mean() > 0.9 and hit[clear].sum() > 5000
    assert np.array_equal(hit[clear], np.isfinite(want_t[clear]))
    on_rect = clear & hit & (tr < ts)
    on_sph = clear & hit & (ts <= tr)
    assert on_rect.sum() > 500 and on_sph.sum() > 2000
    rel = np.abs(tt[on_rect] - want_t[on_rect]) / want_t[on_rect]                              # the rectangle is exact: fp32 rounding only
    assert rel.max() < 2e-4 and np.median(rel) < 1e-6
    assert (prim[on_rect] >= len(st)).all() and (prim[on_sph] < len(st)).all()
    err = np.abs(tt[on_sph] - want_t[on_sph]) / r
    assert err.max() < 1.5e-3 / np.sqrt(1 - 0.98 ** 2) and np.median(err) < 7e-4             # radial deviation / cos(incidence)


# ---- a scene with analytic shapes: the ceiling light as a `rectangle`, a diffuse `sphere` in the fog -------------------------

LIGHT_TO_WORLD = np.array([[0.15, 0, 0, 0.5], [0, 0, -1, 0.998], [0, 0.15, 0, 0.5], [0, 0, 0, 1]], np.float32)   # faces down (-y)
BALL_CENTER, BALL_RADIUS, BALL_STEPS = np.array([0.7, 0.72, 0.3], np.float32), 0.1, 12


def shapes_scene(pkg, width=40, height=32):
    """tracer_scene without its light triangles; two more materials (the light's, the ball's) for the shapes to come.
    Returns (scene, material id of the light, material id of the ball, radiance)"""
    scene, em, rad = pkg.scenes.tracer_scene(width, height, glass=False)
    scene = dict(scene)
    keep = np.ones(len(scene["tris"]), bool); keep[em] = False
    scene["tris"] = np.ascontiguousarray(scene["tris"][keep]); scene["tri_material"] = np.ascontiguousarray(scene["tri_material"][keep])
    n = len(scene["albedo"])
    scene["albedo"] = np.concatenate([scene["albedo"], [[0.73, 0.73, 0.73], [0.2, 0.4, 0.8]]]).astype(np.float32)
    scene["mat_bits"] = np.ones(n + 2, np.uint32)
    return scene, n, n + 1, rad


def with_tessellated_shapes(host_lib, scene, light_mat, ball_mat):
    """the same scene with the shapes' triangles appended by hand (what alvrl_add_rectangle / alvrl_add_sphere do inside)"""
    rv, rt = tessellate(host_lib, 0, LIGHT_TO_WORLD)
    sv, st = tessellate(host_lib, 1, BALL_CENTER, BALL_RADIUS, False, BALL_STEPS)
    nv, nt = len(scene["verts"]), len(scene["tris"])
    out = dict(scene)
    out["verts"] = np.concatenate([scene["verts"], rv, sv]).astype(np.float32)
    out["tris"] = np.concatenate([scene["tris"], rt + nv, st + nv + len(rv)]).astype(np.uint32)
    out["tri_material"] = np.concatenate([scene["tri_material"], np.full(len(rt), light_mat), np.full(len(st), ball_mat)]).astype(np.uint32)
    return out, np.arange(nt, nt + len(rt), dtype=np.uint32)


def test_oracle_walks_a_scene_with_a_rectangle_light_and_a_sphere(pkg, orc, host_lib):
    """the tessellated shapes in the whole path on the oracle: the light (a `rectangle`, facing down) emits into the box, the
    VRL tracer produces its set, the frame renders, and the ball shows up in the primary hits"""
    scene, light_mat, ball_mat, rad = shapes_scene(pkg)
    full, em = with_tessellated_shapes(host_lib, scene, light_mat, ball_mat)
    o = orc.Oracle(volVolSamples=2, volSurfSamples=2, targetNumSlices=6, seed=5, vrlTargetNum=200)
    o.set_scene(full)
    o.set_area_emitter(em, rad)
    o.trace_vrls()
    s, e, p, pc = o.get_vrls()
    on_light = s[:, 1] == np.float32(0.998)                           # the walks' first segments
    assert len(s) >= 200 and on_light.sum() > 20 and (e[on_light, 1] < s[on_light, 1]).all()      # leave downwards: it faces down
    o.build_slices(); o.prepass()
    img = o.render()
    assert np.isfinite(img).all() and img.max() > 0
    prim = o.primary_hits()[0].reshape(-1)
    assert (full["tri_material"][prim[prim != pkg.binding.NO_HIT]] == ball_mat).sum() > 3
