"""Import helper: the package directory is named `tortoise.cpp_b200` (with a dot), which a
plain `import` statement cannot express; load it by path under the module name
`tortoise_cpp_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "tortoise.cpp_b200")


def import_pkg():
    name = "tortoise_cpp_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def import_sub(sub: str):
    import_pkg()
    return importlib.import_module(f"tortoise_cpp_b200.{sub}")
