"""Imports the package directory `hierarchicalmatrices.jl_b200/` (its name, fixed by
the project layout, contains a dot and so is not importable by a plain `import`)
under the module name `hierarchicalmatrices_jl_b200`."""
import importlib.util
import os
import sys

NAME = "hierarchicalmatrices_jl_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hierarchicalmatrices.jl_b200")


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(
        NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
