"""Import helper: the package directory is named `ray-tracer-engine_b200` (not a valid Python
identifier), so it is registered under the module name `ray_tracer_engine_b200`."""
import importlib.util
import os
import sys

_NAME = "ray_tracer_engine_b200"
_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ray-tracer-engine_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_ROOT, "__init__.py"), submodule_search_locations=[_ROOT])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = load()
