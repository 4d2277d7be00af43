"""CPU suite, part 3: the Mitsuba plugin shim (vrl.so) keeps the reference plugin's XML parameters, defaults and
constructor errors (src/integrators/vrl/vrlIntegrator.cpp:128-208) and exports the plugin ABI symbols."""
import ctypes as C
import os

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def plugin(pkg):
    path = os.path.join(ROOT, "mitsuba-alvrl_b200", "vrl.so")
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(path)
    lib.alvrl_plugin_props_new.restype = C.c_void_p
    lib.GetDescription.restype = C.c_char_p
    return lib


def _create(plugin, pkg, **props):
    p = C.c_void_p(plugin.alvrl_plugin_props_new())
    for k, v in props.items():
        if isinstance(v, bool):
            plugin.alvrl_plugin_props_set_bool(p, k.encode(), int(v))
        elif isinstance(v, int):
            plugin.alvrl_plugin_props_set_int(p, k.encode(), v)
        elif isinstance(v, float):
            plugin.alvrl_plugin_props_set_float(p, k.encode(), C.c_float(v))
        else:
            plugin.alvrl_plugin_props_set_string(p, k.encode(), str(v).encode())
    inst = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = plugin.alvrl_plugin_create(p, C.byref(inst), err, 512)
    unq = plugin.alvrl_plugin_unqueried(p)
    params = None
    if rc == 0:
        params = pkg.binding.Params()
        plugin.alvrl_plugin_get_params(inst, C.byref(params))
        plugin.alvrl_plugin_destroy(inst)
    plugin.alvrl_plugin_props_free(p)
    return rc, err.value.decode(), params, unq


def test_plugin_abi_symbols(plugin):
    assert hasattr(plugin, "CreateInstance") and hasattr(plugin, "GetDescription")     # cobject.h:99-107
    assert b"Adaptive Lightslice" in plugin.GetDescription()


def test_defaults_and_overrides(plugin, pkg):
    rc, err, p, unq = _create(plugin, pkg)
    assert rc == 0, err
    assert (p.volVolSamples, p.volSurfSamples, p.targetNumSlices, p.Rsamples, p.maxPasses) == (2, 2, 100, 1, 1)
    assert p.targetPixelUndersampling == 64.0 and p.fallBackUndersampling == 5.0 and p.localUndersampling == -1.0
    assert p.shortVrls == 1 and p.globalCluster == 0 and p.localRefinement == 1
    rc, err, p, unq = _create(plugin, pkg, volVolSamples=4, volSurfSamples=4, targetNumSlices=512, sliceCurvatureFactor=0.25,
                              globalCluster=True, vrlFile="/tmp/x.vrl", maxPasses=3, dumpPasses=False, rrDepth=7)
    assert rc == 0 and unq == 0                      # every known attribute is queried: the loader would not warn
    assert (p.volVolSamples, p.volSurfSamples, p.targetNumSlices, p.globalCluster, p.maxPasses) == (4, 4, 512, 1, 3)
    assert abs(p.sliceCurvatureFactor - 0.25) < 1e-7


def test_reference_constructor_errors(plugin, pkg):
    rc, err, _, _ = _create(plugin, pkg, nc=3)
    assert rc != 0 and "neighbourCount" in err       # vrlIntegrator.cpp:129-131
    rc, err, _, _ = _create(plugin, pkg, volVolSamples=1)
    assert rc != 0 and "volVolSamples" in err        # 149-151
    rc, err, _, _ = _create(plugin, pkg, volSurfSamples=1)
    assert rc != 0 and "volSurfSamples" in err       # 154-156
    rc, err, _, _ = _create(plugin, pkg, volVolSamples=0, volSurfSamples=0)
    assert rc == 0


def test_unknown_attribute_is_left_unqueried(plugin, pkg):
    rc, err, p, unq = _create(plugin, pkg, notAParameter=1)
    assert rc == 0 and unq == 1                      # scenehandler.cpp:792-795 would warn about it


def test_preprocess_checks_the_scene_before_it_touches_the_device(plugin):
    """what preprocess() refuses is decided on the host (no CUDA device needed): a VRL file wants exactly one medium
    (vrlIntegrator.cpp:244-248); without a file the VRLs are traced in every prepass (276-280), which on the device needs one
    medium and one area emitter on a mesh"""
    import numpy as np
    plugin.alvrl_plugin_scene_new.restype = C.c_void_p
    fp = C.POINTER(C.c_float)
    img = np.zeros(3, np.float32)
    one = np.ones(3, np.float32)

    def frame(scene_setup, **props):
        p = C.c_void_p(plugin.alvrl_plugin_props_new())
        for k, v in props.items():
            if isinstance(v, int):
                plugin.alvrl_plugin_props_set_int(p, k.encode(), v)
            else:
                plugin.alvrl_plugin_props_set_string(p, k.encode(), str(v).encode())
        inst = C.c_void_p()
        err = C.create_string_buffer(512)
        assert plugin.alvrl_plugin_create(p, C.byref(inst), err, 512) == 0, err.value
        sc = C.c_void_p(plugin.alvrl_plugin_scene_new())
        scene_setup(sc)
        rc = plugin.alvrl_plugin_render_frame(inst, sc, img.ctypes.data_as(fp), err, 512)
        plugin.alvrl_plugin_destroy(inst); plugin.alvrl_plugin_scene_free(sc); plugin.alvrl_plugin_props_free(p)
        return rc, err.value.decode()

    def medium(sc):
        plugin.alvrl_plugin_scene_add_medium_homogeneous(sc, one.ctypes.data_as(fp), one.ctypes.data_as(fp), C.c_float(-1.0), 0, C.c_float(0.0))

    rc, err = frame(lambda sc: None, vrlFile="/tmp/some.vrl")
    assert rc != 0 and "exactly one medium" in err
    rc, err = frame(lambda sc: None)
    assert rc != 0 and "one medium" in err
    rc, err = frame(medium)
    assert rc != 0 and "area emitter" in err

    def medium_and_dangling_emitter(sc):
        medium(sc)
        plugin.alvrl_plugin_scene_add_area_emitter(sc, C.c_uint32(4), one.ctypes.data_as(fp))
    rc, err = frame(medium_and_dangling_emitter)
    assert rc != 0 and "not a mesh" in err

    def medium_and_emitter_on_a_missing_shape(sc):       # analytic shapes (mts::AnalyticShapeView) are counted separately
        medium(sc)
        plugin.alvrl_plugin_scene_add_area_emitter_on_shape(sc, C.c_uint32(0), one.ctypes.data_as(fp))
    rc, err = frame(medium_and_emitter_on_a_missing_shape)
    assert rc != 0 and "not a mesh" in err


def test_sphere_tessellation_parameter(plugin, pkg):
    """`sphereTessellation` (no equivalent in the reference, like cudaDevice): polar steps of the triangles a `sphere` shape is
    handed to the path as; 0 = the library's default"""
    rc, err, params, unq = _create(plugin, pkg, sphereTessellation=32)
    assert rc == 0 and unq == 0
    rc, err, _, _ = _create(plugin, pkg, sphereTessellation=2)
    assert rc != 0 and "sphereTessellation" in err


def test_inherited_parameters_and_their_errors(plugin, pkg):
    """MonteCarloIntegrator's constructor checks (src/librender/integrator.cpp:300-305); rrDepth reaches the parameter block
    (it is the roulette depth of the VRL tracer, vrlIntegrator.cpp:279)"""
    rc, err, params, _ = _create(plugin, pkg, rrDepth=9)
    assert rc == 0 and params.rrDepth == 9
    rc, err, _, _ = _create(plugin, pkg, rrDepth=0)
    assert rc != 0 and "rrDepth" in err
    rc, err, _, _ = _create(plugin, pkg, maxDepth=0)
    assert rc != 0 and "maxDepth" in err
    rc, err, _, _ = _create(plugin, pkg, maxDepth=-1)
    assert rc == 0


def test_serialize_round_trip_in_the_reference_field_order(plugin, pkg):
    """what travels to a network node: numPasses | rrDepth maxDepth strictNormals hideEmitters | maxPasses dumpPasses |
    volVolSamples volSurfSamples globalCluster localRefinement specRRdepth initialSpecularThroughput shortVrls
    (integrator.cpp:61-63, 315-320, 356-359; vrlIntegrator.cpp:226-235) -- ints and floats 4 bytes, bools 1 byte"""
    import struct
    p = C.c_void_p(plugin.alvrl_plugin_props_new())
    for k, v in dict(numPasses=2, rrDepth=7, maxDepth=12, maxPasses=3, volVolSamples=4, volSurfSamples=6, specularForcedRRdepth=50).items():
        plugin.alvrl_plugin_props_set_int(p, k.encode(), v)
    for k, v in dict(strictNormals=True, hideEmitters=False, dumpPasses=True, globalCluster=True, localRefinement=False, shortVrls=False).items():
        plugin.alvrl_plugin_props_set_bool(p, k.encode(), int(v))
    plugin.alvrl_plugin_props_set_float(p, b"initialSpecularThroughput", C.c_float(12.5))
    inst = C.c_void_p()
    err = C.create_string_buffer(512)
    assert plugin.alvrl_plugin_create(p, C.byref(inst), err, 512) == 0, err.value
    n = plugin.alvrl_plugin_serialize(inst, None, 0)
    buf = (C.c_uint8 * n)()
    assert plugin.alvrl_plugin_serialize(inst, buf, n) == n
    want = struct.pack("=iii??i?ii??if?", 2, 7, 12, True, False, 3, True, 4, 6, True, False, 50, 12.5, False)
    assert bytes(buf) == want
    inst2 = C.c_void_p()
    assert plugin.alvrl_plugin_unserialize(buf, n, C.byref(inst2), err, 512) == 0, err.value
    a, b = pkg.binding.Params(), pkg.binding.Params()
    plugin.alvrl_plugin_get_params(inst, C.byref(a)); plugin.alvrl_plugin_get_params(inst2, C.byref(b))
    for f in ("rrDepth", "maxPasses", "volVolSamples", "volSurfSamples", "globalCluster", "localRefinement", "specularForcedRRdepth",
              "initialSpecularThroughput", "shortVrls"):
        assert getattr(a, f) == getattr(b, f), f
    n2 = plugin.alvrl_plugin_serialize(inst2, None, 0)
    buf2 = (C.c_uint8 * n2)()
    plugin.alvrl_plugin_serialize(inst2, buf2, n2)
    assert bytes(buf2) == want                           # what was not transmitted (the clustering parameters) is at its default
    assert plugin.alvrl_plugin_unserialize(buf, n - 3, C.byref(inst2), err, 512) != 0 and b"beyond the end" in err.value
    plugin.alvrl_plugin_destroy(inst); plugin.alvrl_plugin_props_free(p)
