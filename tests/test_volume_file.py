"""GPU side of the two file formats (the parsing and writing itself is tested on the CPU, tests/test_hostio_cpu.py):
a grid medium taken from a "VOL" file (src/volume/gridvolume.cpp:217-287) is the medium taken from the array, and the
film's NumPy output (src/films/mfilm.cpp:337-348) is the developed film."""
import numpy as np
import pytest

from conftest import small_case

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
