"""GPU suite, last file on purpose: the tests of the boundary additions of round 2's final session, which had no GPU time left.
They were written and dry-run with the CPU oracle standing in for the device handle (profiles/r2_session5_cpu_only.log), so
their first execution on a B200 is the round-end run -- they sort last so that nothing they might trip over can stop the
verified suites in front of them.

  - a grid medium taken from a "VOL" file (src/volume/gridvolume.cpp:217-287) is the medium taken from the array; the film's
    NumPy output (src/films/mfilm.cpp:337-348) is the developed film  (the parsing / writing itself: tests/test_hostio_cpu.py)
  - `rectangle` / `sphere` shapes handed over as triangles, through the ABI and through vrl.so  (tests/test_shapes_cpu.py)
  - slicesFalseColor through vrl.so  (tests/test_plugin_oracle_cpu.py)"""
import ctypes as C

import numpy as np
import pytest

from conftest import small_case
from test_plugin_gpu import _by_material, _plugin, _scene_to_plugin

pytestmark = pytest.mark.gpu


def _points(n, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(0.02, 0.98, (n, 3)).astype(np.float32), rng.uniform(0.02, 0.98, (n, 3)).astype(np.float32)


@pytest.mark.parametrize("uint8", [False, True], ids=["float32", "uint8"])
def test_grid_medium_from_a_volume_file_equals_the_array(pkg, orc, tmp_path, uint8):
    scene, vrls, params = small_case(pkg, "C3", 16, 16, 16, grid=24)
    m = dict(scene["medium"])
    path = str(tmp_path / "density.vol")
    if uint8:
        q = np.clip(np.rint(m["density"].astype(np.float64) * 255), 0, 255).astype(np.uint8)
        pkg.volfile.write_vol(path, q, m["bbox_min"], m["bbox_max"], pkg.volfile.VOL_UINT8)
        m["density"] = q.astype(np.float32) / np.float32(255.0)             # the reference's density map (gridvolume.cpp:212-215)
    else:
        pkg.volfile.write_vol(path, m["density"], m["bbox_min"], m["bbox_max"])
    scene = dict(scene, medium=m)
    a = pkg.integrator(0, **params); a.set_scene(scene)                       # the array
    b = pkg.integrator(0, **params); b.set_scene(scene)
    b.set_medium_grid_file(path, m["scale"], m["albedo"], m["sigmaS_base"], phase=m["phase"], g=m["g"])     # the file, its own AABB
    o = orc.Oracle(**params); o.set_scene(scene)
    p1, p2 = _points(3000, 5)
    s = np.zeros(len(p1), np.int32)
    ta, tb, to = a.eval_transmittance(p1, s, p2), b.eval_transmittance(p1, s, p2), o.eval_transmittance(p1, s, p2)
    assert np.array_equal(ta, tb)
    np.testing.assert_allclose(tb, to, rtol=1e-5, atol=1e-30)            # (and the oracle agrees, as in test_eval_transmittance_grid_medium)
    # the `min` / `max` override of gridvolume.cpp:112-117: half the box, the same density stretched over it
    half = np.array([0.5, 1.0, 1.0], np.float32)
    b.set_medium_grid_file(path, m["scale"], m["albedo"], m["sigmaS_base"], bmin=m["bbox_min"], bmax=half, phase=m["phase"], g=m["g"])
    a.set_medium_grid(m["density"], m["bbox_min"], half, m["scale"], m["albedo"], m["sigmaS_base"], m["phase"], m["g"])
    assert np.array_equal(a.eval_transmittance(p1, s, p2), b.eval_transmittance(p1, s, p2))


def test_volume_file_errors_reach_the_caller(pkg, tmp_path):
    scene, vrls, params = small_case(pkg, "C3", 16, 16, 16, grid=8)
    g = pkg.integrator(0, **params); g.set_scene(scene)
    m = scene["medium"]
    with pytest.raises(pkg.binding.AlvrlError) as e:
        g.set_medium_grid_file(str(tmp_path / "missing.vol"), m["scale"], m["albedo"], m["sigmaS_base"])
    assert e.value.code == -4
    (tmp_path / "bad.vol").write_bytes(b"VOX\x03" + b"\0" * 60)
    with pytest.raises(pkg.binding.AlvrlError) as e:
        g.set_medium_grid_file(str(tmp_path / "bad.vol"), m["scale"], m["albedo"], m["sigmaS_base"])
    assert e.value.code == -1 and "incorrect header identifier" in str(e.value)
    p1, p2 = _points(64, 1)                                                  # the medium set before is still in place
    assert np.isfinite(g.eval_transmittance(p1, np.zeros(64, np.int32), p2)).all()


def test_film_numpy_output_is_the_developed_film(pkg, tmp_path):
    scene, vrls, params = pkg.scenes.make_config("C1", width=40, height=24, n_vrls=8)
    g = pkg.integrator(0, **params); g.set_scene(scene)
    g.film_configure(2, 0.0)                                                 # gaussian, the scene default
    fr = np.random.default_rng(3).random((24, 40, 3), dtype=np.float32)
    g.film_put(fr)
    path = str(tmp_path / "pass0.npy")
    g.film_write_npy(path)
    back = np.load(path)
    assert back.dtype == np.float32 and back.shape == (24, 40, 3)
    assert np.array_equal(back, g.film_develop())


def test_analytic_shapes_through_the_abi_and_through_the_plugin(pkg, host_lib):
    """a `rectangle` (the ceiling light) and a `sphere` handed over as triangles.  alvrl_add_rectangle /
    alvrl_add_sphere == the same triangles appended by hand (hits, traced VRLs, frame bit-identical), and the frame through
    vrl.so with mts::AnalyticShapeView shapes and the emitter on the rectangle == the frame of the same calls on the ABI."""
    from test_shapes_cpu import BALL_CENTER, BALL_RADIUS, BALL_STEPS, LIGHT_TO_WORLD, shapes_scene, with_tessellated_shapes
    lib = _plugin()
    scene, light_mat, ball_mat, rad = shapes_scene(pkg)
    meshes, flat = _by_material(scene)                   # materials without triangles (the two shapes') drop out: renumber
    n_mesh = len(meshes)
    flat["albedo"] = np.concatenate([flat["albedo"], scene["albedo"][[light_mat, ball_mat]]]).astype(np.float32)
    flat["mat_bits"] = np.ones(n_mesh + 2, np.uint32)
    xml = dict(volVolSamples=2, volSurfSamples=2, targetNumSlices=8, seed=5, vrlTargetNum=300, sphereTessellation=BALL_STEPS)
    direct = {k: v for k, v in xml.items() if k != "sphereTessellation"}

    def frame(g, em):
        g.set_area_emitter(em, rad)
        g.build_slices(); g.trace_vrls(); g.prepass()
        return g.primary_hits()[0], g.get_vrls()[0], g.render()

    # --- the ABI: shapes added by the library ---
    a = pkg.integrator(0, **direct)
    a.set_scene(flat)
    first = a.add_rectangle(LIGHT_TO_WORLD, n_mesh)
    sfirst, scount = a.add_sphere(BALL_CENTER, BALL_RADIUS, n_mesh + 1, theta_steps=BALL_STEPS)
    assert first == len(flat["tris"]) and sfirst == first + 2 and scount == 4 * BALL_STEPS * (BALL_STEPS - 2)
    prim_a, vrl_a, img_a = frame(a, np.arange(first, first + 2, dtype=np.uint32))
    # --- the same triangles appended by hand ---
    full, em = with_tessellated_shapes(host_lib, flat, n_mesh, n_mesh + 1)
    b = pkg.integrator(0, **direct)
    b.set_scene(full)
    prim_b, vrl_b, img_b = frame(b, em)
    assert np.array_equal(prim_a, prim_b) and np.array_equal(vrl_a, vrl_b) and np.array_equal(img_a, img_b)
    assert img_a.max() > 0 and (prim_a[prim_a != pkg.binding.NO_HIT] >= sfirst).sum() > 3          # the ball is in the picture

    # --- through vrl.so ---
    p = C.c_void_p(lib.alvrl_plugin_props_new())
    for k, v in xml.items():
        lib.alvrl_plugin_props_set_int(p, k.encode(), v)
    inst = C.c_void_p()
    err = C.create_string_buffer(1024)
    assert lib.alvrl_plugin_create(p, C.byref(inst), err, 1024) == 0, err.value
    assert lib.alvrl_plugin_unqueried(p) == 0
    sc = C.c_void_p(lib.alvrl_plugin_scene_new())
    keep = _scene_to_plugin(lib, sc, scene, meshes)
    fp = C.POINTER(C.c_float)
    m16 = np.ascontiguousarray(LIGHT_TO_WORLD, np.float32).reshape(16)
    la, ba = np.ascontiguousarray(scene["albedo"][light_mat], np.float32), np.ascontiguousarray(scene["albedo"][ball_mat], np.float32)
    cen, r = np.ascontiguousarray(BALL_CENTER, np.float32), np.ascontiguousarray(rad, np.float32)
    lib.alvrl_plugin_scene_add_rectangle(sc, m16.ctypes.data_as(fp), 0, la.ctypes.data_as(fp))
    lib.alvrl_plugin_scene_add_sphere(sc, cen.ctypes.data_as(fp), C.c_float(BALL_RADIUS), 0, ba.ctypes.data_as(fp))
    lib.alvrl_plugin_scene_add_area_emitter_on_shape(sc, C.c_uint32(0), r.ctypes.data_as(fp))
    H, W = scene["camera"]["height"], scene["camera"]["width"]
    img_plugin = np.zeros((H, W, 3), np.float32)
    rc = lib.alvrl_plugin_render_frame(inst, sc, img_plugin.ctypes.data_as(fp), err, 1024)
    assert rc == 0, err.value
    lib.alvrl_plugin_destroy(inst); lib.alvrl_plugin_scene_free(sc); lib.alvrl_plugin_props_free(p)
    del keep
    assert np.array_equal(img_plugin, img_a)


def test_slices_false_colour_through_the_plugin(pkg, orc, tmp_path):
    """slicesFalseColor through vrl.so on the device: the shim colours every hit pixel by its slice id with the reference's formula
    (vrlIntegrator.cpp:577-584).  Same body as the CPU test (where the shim runs on the oracle); the slice map is bit-identical
    between device and oracle, so the expected image is the same."""
    from test_plugin_oracle_cpu import test_false_colour_debug_outputs as body
    body(pkg, orc, _plugin(), tmp_path, "slicesFalseColor")
