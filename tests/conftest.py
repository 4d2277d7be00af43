import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import alvrl_loader
    return alvrl_loader.load()


@pytest.fixture(scope="session")
def orc():
    import orc as _orc   # the CPU oracle: test infrastructure only
    _orc.build()
    return _orc


@pytest.fixture(scope="session")
def host_lib(pkg):
    """libalvrl_host.so: the plugin's host-side logic without any device code (built by __graft_entry__.build())."""
    import ctypes
    path = os.path.join(ROOT, "mitsuba-alvrl_b200", "libalvrl_host.so")
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(path)


def small_case(pkg, name="C1", width=64, height=64, n_vrls=200, **extra):
    scene, vrls, params = pkg.scenes.make_config(name, width=width, height=height, n_vrls=n_vrls,
                                                 grid=extra.pop("grid", None), occluders=extra.pop("occluders", None))
    params.update(extra)
    return scene, vrls, params


def setup(it, scene, vrls):
    it.set_scene(scene)
    it.set_vrls(*vrls)
    return it


def rel_err(a, b, floor):
    return np.abs(a - b) / (np.abs(b) + floor)
