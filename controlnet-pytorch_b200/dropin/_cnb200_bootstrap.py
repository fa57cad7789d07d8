"""Loads the hyphen-named package `controlnet-pytorch_b200` by path (without putting the repo root on sys.path) and
re-exports one of its modules under the reference's import name.  Used by the shim packages next to this file."""
import importlib
import importlib.util
import os
import sys

PKG = "controlnet-pytorch_b200"
PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def package():
    if PKG in sys.modules:
        return sys.modules[PKG]
    spec = importlib.util.spec_from_file_location(PKG, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG] = mod
    spec.loader.exec_module(mod)
    return mod


def reexport(shim_globals, sub):
    """Make the shim module `models.x` / `scheduler.x` expose every public and module-level name of PKG.<sub>."""
    package()
    real = importlib.import_module(PKG + "." + sub)
    for k, v in vars(real).items():
        if not (k.startswith("__") and k.endswith("__")):
            shim_globals[k] = v
    shim_globals["__doc__"] = real.__doc__
    shim_globals["__cnb200_real__"] = real
    return real
