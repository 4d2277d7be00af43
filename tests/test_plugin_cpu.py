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
    medium and one area emitter on a mesh; maxPasses < 0 ("until cancelled", integrator.cpp:398) is refused"""
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
    rc, err = frame(medium, vrlFile="/tmp/some.vrl", maxPasses=-1)
    assert rc != 0 and "maxPasses" in err
