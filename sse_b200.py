"""Import shim: `import sse_b200` loads the package in `stochasticseriesexpansion.jl_b200/`.

The package directory carries the reference's name (with a dot), which Python cannot import by
name; this module registers it under the importable name `sse_b200`.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stochasticseriesexpansion.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "sse_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sse_b200"] = _mod
_spec.loader.exec_module(_mod)
