"""CPU suite: the plugin shim's host logic end to end, with the CPU oracle standing in for the device library.

tests/support/plugin_on_oracle.cpp compiles csrc/plugin/vrl_plugin.cpp (unchanged) against the oracle's orc_* mirror of the C
ABI.  Frames driven the way Mitsuba drives the reference plugin -- CreateInstance(props) -> preprocess(scene) -> render(scene)
= maxPasses x (prepass + render pass) -- must equal the frames of the same calls made directly on the oracle.  What this
pins without a GPU: the marshalling out of the scene (one material per shape, BSDF bits and optics, media of the shapes, the
emitter's triangles, analytic shapes handed over as triangles), the VRL file path, the call order, the per-pass seeds and
the film.  tests/test_plugin_gpu.py runs the same frames through the real vrl.so on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from test_plugin_gpu import _by_material, _scene_to_plugin


@pytest.fixture(scope="module")
def lib(orc, tmp_path_factory):
    orc.build()
    out = str(tmp_path_factory.mktemp("plugin_on_oracle") / "vrl_on_oracle.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", os.path.join(ROOT, "tests", "support", "plugin_on_oracle.cpp"),
                           "-L" + os.path.join(ROOT, "oracle"), "-l:liborc.so", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", out])
    so = C.CDLL(out)
    so.alvrl_plugin_props_new.restype = C.c_void_p
    so.alvrl_plugin_scene_new.restype = C.c_void_p
    return so


def _instance(lib, **xml):
    p = C.c_void_p(lib.alvrl_plugin_props_new())
    for k, v in xml.items():
        if isinstance(v, bool):
            lib.alvrl_plugin_props_set_bool(p, k.encode(), int(v))
        elif isinstance(v, (int, np.integer)):
            lib.alvrl_plugin_props_set_int(p, k.encode(), int(v))
        elif isinstance(v, float):
            lib.alvrl_plugin_props_set_float(p, k.encode(), C.c_float(v))
        else:
            lib.alvrl_plugin_props_set_string(p, k.encode(), str(v).encode())
    inst = C.c_void_p()
    err = C.create_string_buffer(1024)
    assert lib.alvrl_plugin_create(p, C.byref(inst), err, 1024) == 0, err.value
    assert lib.alvrl_plugin_unqueried(p) == 0
    return p, inst


def _frame(lib, inst, sc, scene, rfilter=0):
    fp = C.POINTER(C.c_float)
    H, W = scene["camera"]["height"], scene["camera"]["width"]
    img = np.zeros((H, W, 3), np.float32)
    err = C.create_string_buffer(1024)
    rc = lib.alvrl_plugin_render_frame_filtered(inst, sc, rfilter, C.c_float(0.0), img.ctypes.data_as(fp), err, 1024)
    assert rc == 0, err.value
    return img


@pytest.mark.parametrize("clustered", [True, False], ids=["clustered", "unclustered"])
def test_frame_from_a_vrl_file(pkg, orc, lib, tmp_path, clustered):
    scene, vrls, params = pkg.scenes.make_config("C1", width=40, height=30, n_vrls=90)
    start, end, power, pc = vrls
    path = str(tmp_path / "cornell.vrl")
    pkg.scenes.write_vrl_file(path, start, end, power)
    meshes, flat = _by_material(scene)
    xml = dict(params, targetNumSlices=8, seed=5, vrlFile=path)
    if not clustered:
        xml.update(globalCluster=False, localRefinement=False)
    p, inst = _instance(lib, **xml)
    sc = C.c_void_p(lib.alvrl_plugin_scene_new())
    keep = _scene_to_plugin(lib, sc, scene, meshes)
    img_plugin = _frame(lib, inst, sc, scene)
    lib.alvrl_plugin_destroy(inst); lib.alvrl_plugin_scene_free(sc); lib.alvrl_plugin_props_free(p)
    del keep
    o = orc.Oracle(**{k: v for k, v in xml.items() if k != "vrlFile"})
    o.set_scene(flat)
    o.set_vrls(start, end, power, 0)                     # (the file's lines parse back to exactly these floats)
    if clustered:
        o.build_slices(); o.prepass()
    img = o.render(clustered)
    assert img_plugin.max() > 0 and np.array_equal(img_plugin, img)


def test_frame_with_specular_chains(pkg, orc, lib, tmp_path):
    """BSDF types, optics and the shapes' interior / exterior media reach the path (bsdf.h:230-284, shape.h:427-433)"""
    scene = pkg.scenes.chain_scene(32, 28)
    start, end, power, pc = pkg.scenes.synthetic_vrls(48, sigma_t=1.05, seed=3)
    path = str(tmp_path / "chain.vrl")
    pkg.scenes.write_vrl_file(path, start, end, power)
    meshes, flat = _by_material(scene)
    keepm = [m for _, _, _, m in meshes]
    flat["mat_bits"] = scene["mat_bits"][keepm]
    flat["optics"] = scene["optics"][keepm]
    xml = dict(volVolSamples=2, volSurfSamples=2, targetNumSlices=6, seed=9, vrlFile=path)
    p, inst = _instance(lib, **xml)
    sc = C.c_void_p(lib.alvrl_plugin_scene_new())
    fp, up = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    S = pkg.scenes
    alive = []
    for v, t, a, m in meshes:
        bits = int(scene["mat_bits"][m])
        lib.alvrl_plugin_scene_add_mesh(sc, v.ctypes.data_as(fp), C.c_uint32(len(v)), t.ctypes.data_as(up), C.c_uint32(len(t)), a.ctypes.data_as(fp), int(bits & 1))
        if bits & (S.BSDF_DIELECTRIC | S.BSDF_CONDUCTOR):
            eta = np.ascontiguousarray(scene["optics"][m, 0:3], np.float32); kk = np.ascontiguousarray(scene["optics"][m, 3:6], np.float32)
            alive += [eta, kk]
            lib.alvrl_plugin_scene_set_mesh_bsdf(sc, 1 if bits & S.BSDF_DIELECTRIC else 2, eta.ctypes.data_as(fp), kk.ctypes.data_as(fp),
                                                 int(bool(bits & S.MAT_TRANSITION)), int(bool(bits & S.MAT_INTERIOR_MEDIUM)), int(bool(bits & S.MAT_EXTERIOR_MEDIUM)))
    keep = _scene_to_plugin(lib, sc, scene, [])          # medium + sensor only: the meshes are in already
    img_plugin = _frame(lib, inst, sc, scene)
    lib.alvrl_plugin_destroy(inst); lib.alvrl_plugin_scene_free(sc); lib.alvrl_plugin_props_free(p)
    del keep, alive
    o = orc.Oracle(**{k: v for k, v in xml.items() if k != "vrlFile"})
    o.set_scene(flat)
    o.set_vrls(start, end, power, 0)
    assert o.chain_segments()[0][-1] > 50
    o.build_slices(); o.prepass()
    assert np.array_equal(img_plugin, o.render())


@pytest.mark.parametrize("passes,rfilter", [(1, 0), (3, 0), (2, 2)], ids=["one-pass", "three-passes-box", "two-passes-gaussian"])
def test_traced_vrls_progressive_passes_and_analytic_shapes(pkg, orc, lib, host_lib, passes, rfilter):
    """vrlFile == "": every prepass traces its VRLs (vrlIntegrator.cpp:276-280) from the emitter -- here a `rectangle` shape --,
    ProgressiveMonteCarloIntegrator::render accumulates the passes in the film (integrator.cpp:380-440); a `sphere` shape sits in
    the fog.  The shim hands both shapes over as triangles (alvrl_add_rectangle / alvrl_add_sphere)."""
    from test_shapes_cpu import BALL_CENTER, BALL_RADIUS, BALL_STEPS, LIGHT_TO_WORLD, shapes_scene, with_tessellated_shapes
    scene, light_mat, ball_mat, rad = shapes_scene(pkg, 32, 24)
    meshes, flat = _by_material(scene)
    n_mesh = len(meshes)
    flat["albedo"] = np.concatenate([flat["albedo"], scene["albedo"][[light_mat, ball_mat]]]).astype(np.float32)
    flat["mat_bits"] = np.ones(n_mesh + 2, np.uint32)
    xml = dict(volVolSamples=2, volSurfSamples=2, targetNumSlices=6, seed=5, vrlTargetNum=150, maxPasses=passes, sphereTessellation=BALL_STEPS)
    p, inst = _instance(lib, **xml)
    sc = C.c_void_p(lib.alvrl_plugin_scene_new())
    keep = _scene_to_plugin(lib, sc, scene, meshes)
    fp = C.POINTER(C.c_float)
    m16 = np.ascontiguousarray(LIGHT_TO_WORLD, np.float32).reshape(16)
    la, ba = np.ascontiguousarray(scene["albedo"][light_mat], np.float32), np.ascontiguousarray(scene["albedo"][ball_mat], np.float32)
    cen, r = np.ascontiguousarray(BALL_CENTER, np.float32), np.ascontiguousarray(rad, np.float32)
    lib.alvrl_plugin_scene_add_rectangle(sc, m16.ctypes.data_as(fp), 0, la.ctypes.data_as(fp))
    lib.alvrl_plugin_scene_add_sphere(sc, cen.ctypes.data_as(fp), C.c_float(BALL_RADIUS), 0, ba.ctypes.data_as(fp))
    lib.alvrl_plugin_scene_add_area_emitter_on_shape(sc, C.c_uint32(0), r.ctypes.data_as(fp))
    img_plugin = _frame(lib, inst, sc, scene, rfilter)
    lib.alvrl_plugin_destroy(inst); lib.alvrl_plugin_scene_free(sc); lib.alvrl_plugin_props_free(p)
    del keep

    full, em = with_tessellated_shapes(host_lib, flat, n_mesh, n_mesh + 1)          # the shapes' triangles appended by hand
    o = orc.Oracle(**{k: v for k, v in xml.items() if k not in ("maxPasses", "sphereTessellation")})
    o.set_scene(full)
    o.set_area_emitter(em, rad)
    o.build_slices()
    frames = []
    for k in range(passes):
        if k:
            o.set_seed(5 + k)
        o.trace_vrls()
        o.prepass()
        frames.append(o.render())
    use_film = passes > 1 or rfilter != 0
    img = orc.film(np.stack(frames), rfilter, 0.0) if use_film else frames[0]
    assert img_plugin.max() > 0 and np.array_equal(img_plugin, img)
    if passes > 1:
        assert not np.array_equal(frames[0], frames[1])


def test_dump_passes_names_the_pass_file_like_the_reference(pkg, orc, lib, tmp_path):
    """dumpPasses (src/librender/integrator.cpp:361-378, 436-438; passFileSuffix, vrlIntegrator.cpp:357-364): with a pass count the
    last pass is dumped under <destination>_passNNN_precpu.._prewall.._rencpu.._renwall.._prevrl.._renvrl...blahExtensionTODO --
    cumulative times in %.4e, the two StatsCounters ("Number of integrated VRLs during preprocessing / rendering") as floats"""
    import re
    scene, vrls, params = pkg.scenes.make_config("C1", width=24, height=20, n_vrls=40)
    start, end, power, pc = vrls
    path = str(tmp_path / "set.vrl")
    pkg.scenes.write_vrl_file(path, start, end, power)
    meshes, flat = _by_material(scene)
    xml = dict(params, targetNumSlices=5, seed=2, vrlFile=path, maxPasses=2)
    fp = C.POINTER(C.c_float)
    names = {}
    for dump in (True, False):
        p, inst = _instance(lib, dumpPasses=dump, **xml)
        sc = C.c_void_p(lib.alvrl_plugin_scene_new())
        keep = _scene_to_plugin(lib, sc, scene, meshes)
        img = np.zeros((20, 24, 3), np.float32)
        name, err = C.create_string_buffer(600), C.create_string_buffer(600)
        rc = lib.alvrl_plugin_render_frame_dump(inst, sc, b"/renders/cornell", img.ctypes.data_as(fp), name, 600, err, 600)
        assert rc == 0, err.value
        names[dump] = name.value.decode()
        lib.alvrl_plugin_destroy(inst); lib.alvrl_plugin_scene_free(sc); lib.alvrl_plugin_props_free(p)
        del keep
    assert names[False] == ""
    num = r"(\d\.\d{4}e[+-]\d{2})"
    m = re.fullmatch(r"/renders/cornell_pass(\d{3})_precpu" + num + "_prewall" + num + "_rencpu" + num + "_renwall" + num + "_prevrl" + num +
                     "_renvrl" + num + r"\.blahExtensionTODO", names[True])
    assert m, names[True]
    assert m.group(1) == "002" and float(m.group(3)) > 0 and float(m.group(5)) > 0
    # the counters: the same two passes on the oracle itself
    o = orc.Oracle(**{k: v for k, v in xml.items() if k not in ("vrlFile", "maxPasses")})
    o.set_scene(flat); o.set_vrls(start, end, power, 0); o.build_slices()
    for k in range(2):
        if k:
            o.set_seed(2 + k)
        o.prepass(); o.render()
    st = o.stats()
    assert m.group(6) == "%.4e" % np.float32(st.pairsPreprocess) and m.group(7) == "%.4e" % np.float32(st.pairsRender)


@pytest.mark.parametrize("mode", ["slicesFalseColor", "numVrlFalseColor", "numVrlFalseColor-unclustered", "convergenceFalseColor"])
def test_false_colour_debug_outputs(pkg, orc, lib, tmp_path, mode):
    """numVrlFalseColor / slicesFalseColor / convergenceFalseColor (vrlIntegrator.cpp:199-201, 514-520, 574-584, 806-807) for scenes
    without delta surfaces, made by the shim from the library's hit mask, slice map and cluster counts.  Expected images are
    written out here from the reference's formulas; the pixel indexing (m_slices[y + sizeY * x]) is cross-checked against the
    radiance image: the same pixels are black."""
    scene, vrls, params = pkg.scenes.make_config("C1", width=36, height=28, n_vrls=60)
    cam = pkg.scenes.perspective_camera(36, 28, origin=(0.45, 0.55, -2.4), target=(0.5, 0.5, 0.0), fov=40.0)     # further back: rays pass the box
    scene = dict(scene, camera=cam, extra_bounds=cam["origin"].reshape(1, 3))
    start, end, power, pc = vrls
    path = str(tmp_path / "set.vrl")
    pkg.scenes.write_vrl_file(path, start, end, power)
    meshes, flat = _by_material(scene)
    xml = dict(params, targetNumSlices=7, seed=4, vrlFile=path)
    key = mode.split("-")[0]
    unclustered = mode.endswith("unclustered")
    if unclustered:
        xml.update(globalCluster=False, localRefinement=False)

    def frame(**extra):
        p, inst = _instance(lib, **dict(xml, **extra))
        sc = C.c_void_p(lib.alvrl_plugin_scene_new())
        keep = _scene_to_plugin(lib, sc, scene, meshes)
        img = _frame(lib, inst, sc, scene)
        lib.alvrl_plugin_destroy(inst); lib.alvrl_plugin_scene_free(sc); lib.alvrl_plugin_props_free(p)
        del keep
        return img

    img = frame(**{key: True})
    radiance = frame()
    o = orc.Oracle(**{k: v for k, v in xml.items() if k != "vrlFile"})
    o.set_scene(flat); o.set_vrls(start, end, power, 0)
    H, W = 28, 36
    hit = (o.primary_hits()[0] != pkg.binding.NO_HIT).reshape(W, H).T                 # the library's pixel index is y + H * x
    assert 0 < hit.sum() < hit.size                                                  # some camera rays leave the open box
    lit = radiance.sum(-1) > 0                                                       # ... and those pixels are black in the radiance image
    assert not lit[~hit].any() and lit[hit].mean() > 0.99
    if key == "convergenceFalseColor":                                               # quirk B4: no delta surface, no effect
        assert np.array_equal(img, radiance)
        return
    if unclustered:
        want = np.where(hit[..., None], np.float32(1.0), np.float32(0.0)) * np.ones(3, np.float32)
    else:
        o.build_slices(); o.prepass()
        sl = o.pixel_to_slice().reshape(W, H).T.astype(np.uint64)
        if key == "slicesFalseColor":
            s = sl % (1 << 32)
            rgbv = np.stack([((s + s * s) % (1 << 32) % 43) / 43.0, ((7 * s + 2 * s * s + 7) % (1 << 32) % 41) / 41.0,
                             ((23 * s + 5 * s * s + s * s * s + 17) % (1 << 32) % 53) / 53.0], -1).astype(np.float32)
        else:
            cl = o.clusters()
            counts = np.diff(cl["offset"]).astype(np.float32)
            assert (counts > 0).all()
            rgbv = (counts[np.minimum(sl, len(counts) - 1).astype(np.int64)] / np.float32(len(start)))[..., None] * np.ones(3, np.float32)
        want = np.where(hit[..., None], rgbv, np.float32(0.0)).astype(np.float32)
        assert len(np.unique(sl[hit])) == o.num_slices()[0] > 1
    assert np.array_equal(img, want)


def test_false_colour_refuses_specular_scenes_and_unclustered_slices(pkg, lib, tmp_path):
    scene = pkg.scenes.chain_scene(16, 16)
    meshes, flat = _by_material(scene)
    fp, up = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    for xml, needle, specular in ((dict(slicesFalseColor=True, vrlFile="/x.vrl"), "specular surfaces", True),
                                  (dict(slicesFalseColor=True, localRefinement=False, globalCluster=False, vrlFile="/x.vrl"), "without clustering", False)):
        p, inst = _instance(lib, **xml)
        sc = C.c_void_p(lib.alvrl_plugin_scene_new())
        for v, t, a, m in meshes:
            lib.alvrl_plugin_scene_add_mesh(sc, v.ctypes.data_as(fp), C.c_uint32(len(v)), t.ctypes.data_as(up), C.c_uint32(len(t)), a.ctypes.data_as(fp), 1)
            if specular and int(scene["mat_bits"][m]) & pkg.scenes.BSDF_DIELECTRIC:
                lib.alvrl_plugin_scene_set_mesh_bsdf(sc, 1, None, None, 1, 0, 1)
        keep = _scene_to_plugin(lib, sc, scene, [])
        img = np.zeros((16, 16, 3), np.float32)
        err = C.create_string_buffer(600)
        rc = lib.alvrl_plugin_render_frame(inst, sc, img.ctypes.data_as(fp), err, 600)
        assert rc != 0 and needle.encode() in err.value, err.value
        lib.alvrl_plugin_destroy(inst); lib.alvrl_plugin_scene_free(sc); lib.alvrl_plugin_props_free(p)
        del keep


def test_render_until_cancelled_dumps_every_pass(pkg, orc, lib, tmp_path):
    """maxPasses < 0 (integrator.cpp:398, 428-430): passes run until cancel(); with dumpPasses every pass is dumped.  The film of the
    test hook cancels from inside the third dump, as a GUI thread would between passes: three passes, three pass files numbered
    001..003, and the final film is the film of the third dump = three passes accumulated"""
    import re
    scene, vrls, params = pkg.scenes.make_config("C1", width=20, height=16, n_vrls=30)
    start, end, power, pc = vrls
    path = str(tmp_path / "set.vrl")
    pkg.scenes.write_vrl_file(path, start, end, power)
    meshes, flat = _by_material(scene)
    xml = dict(params, targetNumSlices=4, seed=7, vrlFile=path, maxPasses=-1, dumpPasses=True)
    p, inst = _instance(lib, **xml)
    sc = C.c_void_p(lib.alvrl_plugin_scene_new())
    keep = _scene_to_plugin(lib, sc, scene, meshes)
    img = np.zeros((16, 20, 3), np.float32)
    names, err = C.create_string_buffer(4000), C.create_string_buffer(600)
    rc = lib.alvrl_plugin_render_until_cancelled(inst, sc, b"/out/img", 3, img.ctypes.data_as(C.POINTER(C.c_float)), names, 4000, err, 600)
    assert rc == 3, err.value
    lib.alvrl_plugin_destroy(inst); lib.alvrl_plugin_scene_free(sc); lib.alvrl_plugin_props_free(p)
    del keep
    files = names.value.decode().split()
    assert [re.match(r"/out/img_pass(\d{3})_", f).group(1) for f in files] == ["001", "002", "003"]
    o = orc.Oracle(**{k: v for k, v in xml.items() if k not in ("vrlFile", "maxPasses", "dumpPasses")})
    o.set_scene(flat); o.set_vrls(start, end, power, 0); o.build_slices()
    frames = []
    for k in range(3):
        if k:
            o.set_seed(7 + k)
        o.prepass(); frames.append(o.render())
    assert np.array_equal(img, orc.film(np.stack(frames), 0, 0.0))
