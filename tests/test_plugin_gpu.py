"""GPU suite, part 3: the drop-in boundary end to end.  vrl.so -- the `integrator type="vrl"` plugin shim -- is driven the
way Mitsuba drives the reference plugin: CreateInstance(props with vrlFile) -> preprocess(scene) -> prepass -> render
(src/integrators/vrl/vrlIntegrator.cpp:237-356, src/librender/scene.cpp:416-449, integrator.cpp:380-440), and its image
must equal the image of the same calls made directly on the C ABI.  The VRL set arrives through the reference's ASCII
file format (VRL.h:43-54, 120-128), which is the plugin's only ingestion path (alvrl_load_vrl_file)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _plugin():
    lib = C.CDLL(os.path.join(ROOT, "mitsuba-alvrl_b200", "vrl.so"))
    lib.alvrl_plugin_props_new.restype = C.c_void_p
    lib.alvrl_plugin_scene_new.restype = C.c_void_p
    return lib


def _by_material(scene):
    """one TriMesh per material, as Mitsuba shapes carry one BSDF each; returns the regrouped flat scene too"""
    meshes, verts, tris, mats = [], scene["verts"], scene["tris"], scene["tri_material"]
    for m in range(len(scene["albedo"])):
        t = np.ascontiguousarray(tris[mats == m], dtype=np.uint32)
        if len(t):
            meshes.append((np.ascontiguousarray(verts, dtype=np.float32), t, np.ascontiguousarray(scene["albedo"][m], dtype=np.float32), m))
    flat = dict(scene)
    nv = len(verts)
    flat["verts"] = np.concatenate([mv for mv, _, _, _ in meshes])
    flat["tris"] = np.concatenate([t + i * nv for i, (_, t, _, _) in enumerate(meshes)])
    flat["tri_material"] = np.concatenate([np.full(len(t), i, np.uint32) for i, (_, t, _, _) in enumerate(meshes)])
    flat["albedo"] = np.stack([a for _, _, a, _ in meshes])
    flat["mat_bits"] = np.ones(len(meshes), np.uint32)
    return meshes, flat


@pytest.mark.parametrize("clustered", [True, False], ids=["clustered", "unclustered"])
def test_plugin_frame_equals_direct_abi(pkg, tmp_path, clustered):
    lib = _plugin()
    scene, vrls, params = pkg.scenes.make_config("C1", width=64, height=48, n_vrls=180)
    start, end, power, pc = vrls
    # a few lines the reader must filter (VRL.h:148-158): zero power, zero length
    start = np.concatenate([start, [[0.5, 0.5, 0.5], [0.2, 0.2, 0.2]]]).astype(np.float32)
    end = np.concatenate([end, [[0.6, 0.5, 0.5], [0.2, 0.2, 0.2]]]).astype(np.float32)
    power = np.concatenate([power, [[0, 0, 0], [1, 1, 1]]]).astype(np.float32)
    path = str(tmp_path / "cornell.vrl")
    pkg.scenes.write_vrl_file(path, start, end, power)
    meshes, flat = _by_material(scene)
    xml = dict(params, targetNumSlices=12, seed=5, vrlFile=path)
    if not clustered:
        xml.update(globalCluster=False, localRefinement=False)

    # --- through the plugin ---
    p = C.c_void_p(lib.alvrl_plugin_props_new())
    for k, v in xml.items():
        if isinstance(v, bool):
            lib.alvrl_plugin_props_set_bool(p, k.encode(), int(v))
        elif isinstance(v, int):
            lib.alvrl_plugin_props_set_int(p, k.encode(), v)
        elif isinstance(v, float):
            lib.alvrl_plugin_props_set_float(p, k.encode(), C.c_float(v))
        else:
            lib.alvrl_plugin_props_set_string(p, k.encode(), str(v).encode())
    inst = C.c_void_p()
    err = C.create_string_buffer(1024)
    assert lib.alvrl_plugin_create(p, C.byref(inst), err, 1024) == 0, err.value
    assert lib.alvrl_plugin_unqueried(p) == 0
    sc = C.c_void_p(lib.alvrl_plugin_scene_new())
    fp, up = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    for v, t, a, _ in meshes:
        lib.alvrl_plugin_scene_add_mesh(sc, v.ctypes.data_as(fp), C.c_uint32(len(v)), t.ctypes.data_as(up), C.c_uint32(len(t)), a.ctypes.data_as(fp), 1)
    med = scene["medium"]
    sa, ss = np.ascontiguousarray(med["sigmaA"], np.float32), np.ascontiguousarray(med["sigmaS"], np.float32)
    lib.alvrl_plugin_scene_add_medium_homogeneous(sc, sa.ctypes.data_as(fp), ss.ctypes.data_as(fp), C.c_float(-1.0), 0, C.c_float(0.0))
    cam = scene["camera"]
    s2c = np.ascontiguousarray(cam["sampleToCamera"], np.float32).reshape(16)
    c2w = np.ascontiguousarray(cam["cameraToWorld"], np.float32).reshape(16)
    pos = np.ascontiguousarray(cam["origin"], np.float32)
    W, H = cam["width"], cam["height"]
    lib.alvrl_plugin_scene_set_sensor(sc, s2c.ctypes.data_as(fp), c2w.ctypes.data_as(fp), C.c_uint32(W), C.c_uint32(H),
                                      C.c_float(cam["near"]), C.c_float(cam["far"]), pos.ctypes.data_as(fp))
    img_plugin = np.zeros((H, W, 3), np.float32)
    rc = lib.alvrl_plugin_render_frame(inst, sc, img_plugin.ctypes.data_as(fp), err, 1024)
    assert rc == 0, err.value
    lib.alvrl_plugin_destroy(inst); lib.alvrl_plugin_scene_free(sc); lib.alvrl_plugin_props_free(p)

    # --- the same calls on the C ABI ---
    direct = {k: v for k, v in xml.items() if k != "vrlFile"}
    g = pkg.integrator(0, **direct)
    g.set_scene(flat)
    g.load_vrl_file(path)
    assert g.N == 180                                   # the two bad lines were dropped by the put() filter
    if clustered:
        g.build_slices(); g.prepass()
    img_direct = g.render(clustered)
    assert img_plugin.max() > 0
    assert np.array_equal(img_plugin, img_direct)

    # and the file path gives the VRL set that the array path gives
    g2 = pkg.integrator(0, **direct)
    g2.set_scene(flat)
    g2.set_vrls(start, end, power, 0)
    if clustered:
        g2.build_slices(); g2.prepass()
    np.testing.assert_allclose(g2.render(clustered), img_direct, rtol=1e-5, atol=1e-9)


def test_plugin_reports_missing_vrl_file(pkg, tmp_path):
    lib = _plugin()
    p = C.c_void_p(lib.alvrl_plugin_props_new())
    lib.alvrl_plugin_props_set_string(p, b"vrlFile", str(tmp_path / "nope.vrl").encode())
    inst = C.c_void_p()
    err = C.create_string_buffer(1024)
    assert lib.alvrl_plugin_create(p, C.byref(inst), err, 1024) == 0
    sc = C.c_void_p(lib.alvrl_plugin_scene_new())
    img = np.zeros(3, np.float32)
    rc = lib.alvrl_plugin_render_frame(inst, sc, img.ctypes.data_as(C.POINTER(C.c_float)), err, 1024)
    assert rc != 0 and b"medium" in err.value           # vrlIntegrator.cpp:244-248: exactly one medium
    lib.alvrl_plugin_destroy(inst); lib.alvrl_plugin_scene_free(sc); lib.alvrl_plugin_props_free(p)


def test_plugin_frame_with_specular_chains(pkg, tmp_path):
    """a glass sphere and a conductor in the scene: the shim marshals BSDF types, optics and the shapes' interior / exterior media
    (bsdf.h:230-284, shape.h:427-433), and the frame with the specular chains of LiInternal (vrlIntegrator.cpp:445-511) equals
    the frame of the same calls on the C ABI"""
    lib = _plugin()
    scene = pkg.scenes.chain_scene(48, 40)
    start, end, power, pc = pkg.scenes.synthetic_vrls(96, sigma_t=1.05, seed=3)
    path = str(tmp_path / "chain.vrl")
    pkg.scenes.write_vrl_file(path, start, end, power)
    meshes, flat = _by_material(scene)
    keep = [m for _, _, _, m in meshes]
    flat["mat_bits"] = scene["mat_bits"][keep]
    flat["optics"] = scene["optics"][keep]
    xml = dict(volVolSamples=2, volSurfSamples=2, targetNumSlices=8, seed=9, vrlFile=path)
    p = C.c_void_p(lib.alvrl_plugin_props_new())
    for k, v in xml.items():
        if isinstance(v, int):
            lib.alvrl_plugin_props_set_int(p, k.encode(), v)
        else:
            lib.alvrl_plugin_props_set_string(p, k.encode(), str(v).encode())
    inst = C.c_void_p()
    err = C.create_string_buffer(1024)
    assert lib.alvrl_plugin_create(p, C.byref(inst), err, 1024) == 0, err.value
    sc = C.c_void_p(lib.alvrl_plugin_scene_new())
    fp, up = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    S = pkg.scenes
    for v, t, a, m in meshes:
        bits = int(scene["mat_bits"][m])
        lib.alvrl_plugin_scene_add_mesh(sc, v.ctypes.data_as(fp), C.c_uint32(len(v)), t.ctypes.data_as(up), C.c_uint32(len(t)), a.ctypes.data_as(fp), int(bits & 1))
        if bits & (S.BSDF_DIELECTRIC | S.BSDF_CONDUCTOR):
            eta = np.ascontiguousarray(scene["optics"][m, 0:3], np.float32); kk = np.ascontiguousarray(scene["optics"][m, 3:6], np.float32)
            lib.alvrl_plugin_scene_set_mesh_bsdf(sc, 1 if bits & S.BSDF_DIELECTRIC else 2, eta.ctypes.data_as(fp), kk.ctypes.data_as(fp),
                                                 int(bool(bits & S.MAT_TRANSITION)), int(bool(bits & S.MAT_INTERIOR_MEDIUM)), int(bool(bits & S.MAT_EXTERIOR_MEDIUM)))
    med = scene["medium"]
    sa, ss = np.ascontiguousarray(med["sigmaA"], np.float32), np.ascontiguousarray(med["sigmaS"], np.float32)
    lib.alvrl_plugin_scene_add_medium_homogeneous(sc, sa.ctypes.data_as(fp), ss.ctypes.data_as(fp), C.c_float(-1.0), 0, C.c_float(0.0))
    cam = scene["camera"]
    s2c = np.ascontiguousarray(cam["sampleToCamera"], np.float32).reshape(16)
    c2w = np.ascontiguousarray(cam["cameraToWorld"], np.float32).reshape(16)
    pos = np.ascontiguousarray(cam["origin"], np.float32)
    W, H = cam["width"], cam["height"]
    lib.alvrl_plugin_scene_set_sensor(sc, s2c.ctypes.data_as(fp), c2w.ctypes.data_as(fp), C.c_uint32(W), C.c_uint32(H),
                                      C.c_float(cam["near"]), C.c_float(cam["far"]), pos.ctypes.data_as(fp))
    img_plugin = np.zeros((H, W, 3), np.float32)
    assert lib.alvrl_plugin_render_frame(inst, sc, img_plugin.ctypes.data_as(fp), err, 1024) == 0, err.value
    lib.alvrl_plugin_destroy(inst); lib.alvrl_plugin_scene_free(sc); lib.alvrl_plugin_props_free(p)

    g = pkg.integrator(0, **{k: v for k, v in xml.items() if k != "vrlFile"})
    g.set_scene(flat)
    g.load_vrl_file(path)
    off, _ = g.chain_segments()
    assert off[-1] > 100                                  # the chains exist
    g.build_slices(); g.prepass()
    assert np.array_equal(img_plugin, g.render())


def _scene_to_plugin(lib, sc, scene, meshes):
    """the mock Scene from a flat scene dict: meshes, the one medium, the sensor; returns the arrays that must stay alive"""
    fp, up = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    for v, t, a, _ in meshes:
        lib.alvrl_plugin_scene_add_mesh(sc, v.ctypes.data_as(fp), C.c_uint32(len(v)), t.ctypes.data_as(up), C.c_uint32(len(t)), a.ctypes.data_as(fp), 1)
    med = scene["medium"]
    sa, ss = np.ascontiguousarray(med["sigmaA"], np.float32), np.ascontiguousarray(med["sigmaS"], np.float32)
    lib.alvrl_plugin_scene_add_medium_homogeneous(sc, sa.ctypes.data_as(fp), ss.ctypes.data_as(fp), C.c_float(-1.0), 0, C.c_float(0.0))
    cam = scene["camera"]
    s2c = np.ascontiguousarray(cam["sampleToCamera"], np.float32).reshape(16)
    c2w = np.ascontiguousarray(cam["cameraToWorld"], np.float32).reshape(16)
    pos = np.ascontiguousarray(cam["origin"], np.float32)
    lib.alvrl_plugin_scene_set_sensor(sc, s2c.ctypes.data_as(fp), c2w.ctypes.data_as(fp), C.c_uint32(cam["width"]), C.c_uint32(cam["height"]),
                                      C.c_float(cam["near"]), C.c_float(cam["far"]), pos.ctypes.data_as(fp))
    return sa, ss, s2c, c2w, pos


@pytest.mark.parametrize("passes,rfilter", [(1, 0), (3, 0), (2, 2)], ids=["one-pass", "three-passes-box", "two-passes-gaussian"])
def test_plugin_traces_its_vrls_and_renders_progressive_passes(pkg, passes, rfilter):
    """vrlFile == "": every prepass traces a fresh VRL set (vrlIntegrator.cpp:276-280) and ProgressiveMonteCarloIntegrator::render
    runs prepass + render pass maxPasses times into one film (integrator.cpp:380-440).  Through vrl.so == the same calls on the ABI."""
    lib = _plugin()
    scene, em, rad = pkg.scenes.tracer_scene(40, 32, glass=False)
    scene = dict(scene)
    # the emitter's quad as a shape of its own (an area emitter is attached to one shape): a material id of its own
    scene["albedo"] = np.concatenate([scene["albedo"], scene["albedo"][scene["tri_material"][em[0]]][None]]).astype(np.float32)
    tm = scene["tri_material"].copy(); tm[em] = len(scene["albedo"]) - 1
    scene["tri_material"] = tm
    scene["mat_bits"] = np.ones(len(scene["albedo"]), np.uint32)
    meshes, flat = _by_material(scene)
    em_mesh = [i for i, (_, _, _, m) in enumerate(meshes) if m == len(scene["albedo"]) - 1][0]
    em_flat = np.nonzero(flat["tri_material"] == em_mesh)[0].astype(np.uint32)
    assert len(em_flat) == len(em)
    xml = dict(volVolSamples=2, volSurfSamples=2, targetNumSlices=8, seed=5, vrlTargetNum=300, maxPasses=passes)

    p = C.c_void_p(lib.alvrl_plugin_props_new())
    for k, v in xml.items():
        lib.alvrl_plugin_props_set_int(p, k.encode(), v)
    inst = C.c_void_p()
    err = C.create_string_buffer(1024)
    assert lib.alvrl_plugin_create(p, C.byref(inst), err, 1024) == 0, err.value
    sc = C.c_void_p(lib.alvrl_plugin_scene_new())
    keep = _scene_to_plugin(lib, sc, scene, meshes)
    fp = C.POINTER(C.c_float)
    H, W = scene["camera"]["height"], scene["camera"]["width"]
    img_plugin = np.zeros((H, W, 3), np.float32)
    # no emitter in the scene: the tracer has nothing to start from
    assert lib.alvrl_plugin_render_frame(inst, sc, img_plugin.ctypes.data_as(fp), err, 1024) != 0 and b"emitter" in err.value
    lib.alvrl_plugin_destroy(inst)
    assert lib.alvrl_plugin_create(p, C.byref(inst), err, 1024) == 0, err.value
    r = np.ascontiguousarray(rad, np.float32)
    lib.alvrl_plugin_scene_add_area_emitter(sc, C.c_uint32(em_mesh), r.ctypes.data_as(fp))
    rc = lib.alvrl_plugin_render_frame_filtered(inst, sc, rfilter, C.c_float(0.0), img_plugin.ctypes.data_as(fp), err, 1024)
    assert rc == 0, err.value
    lib.alvrl_plugin_destroy(inst); lib.alvrl_plugin_scene_free(sc); lib.alvrl_plugin_props_free(p)
    del keep

    g = pkg.integrator(0, **{k: v for k, v in xml.items() if k != "maxPasses"})
    g.set_scene(flat)
    g.set_area_emitter(em_flat, rad)
    g.build_slices()
    use_film = passes > 1 or rfilter != 0
    if use_film:
        g.film_configure(rfilter, 0.0)
    sets = []
    for k in range(passes):
        if k:
            g.set_seed(5 + k)
        g.trace_vrls()
        sets.append(g.get_vrls()[0])
        g.prepass()
        img_direct = g.render()
        if use_film:
            g.film_put(None)
    if use_film:
        img_direct = g.film_develop()
    assert img_plugin.max() > 0 and np.array_equal(img_plugin, img_direct)
    if passes > 1:                                       # every pass has its own VRL set
        assert not np.array_equal(sets[0][:50], sets[1][:50])


def test_set_seed_equals_a_fresh_handle(pkg):
    """alvrl_set_seed(s) on a used handle == a handle created with seed s: slice mapping, R, clusters and the image"""
    scene, vrls, params = pkg.scenes.make_config("C1", width=48, height=40, n_vrls=150)
    params.update(targetNumSlices=8)
    a = pkg.integrator(0, **dict(params, seed=3))
    a.set_scene(scene); a.set_vrls(*vrls); a.build_slices(); a.prepass()
    img3 = a.render()
    a.set_seed(11)
    a.prepass()
    img11 = a.render()
    b = pkg.integrator(0, **dict(params, seed=11))
    b.set_scene(scene); b.set_vrls(*vrls); b.build_slices(); b.prepass()
    assert np.array_equal(img11, b.render()) and not np.array_equal(img3, img11)
    ca, cb = a.clusters(), b.clusters()
    assert np.array_equal(ca["offset"], cb["offset"]) and np.array_equal(ca["vrls"], cb["vrls"])

