"""Import helper: the package directory is `mitsuba-alvrl_b200/` (the name the project brief fixes), which is
not a valid Python identifier, so it is registered in sys.modules as `mitsuba_alvrl_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "mitsuba-alvrl_b200")


def load():
    name = "mitsuba_alvrl_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
