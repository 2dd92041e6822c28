"""Loader for the UNMODIFIED reference kernels (test infrastructure only).

Imports the reference's own Numba kernels from /root/reference without PCRaster/netCDF4 by
stubbing the few non-numerical imports (recipe: SURVEY.md appendix A.8).  It only works in the
build container (where /root/reference exists); it is used by tests/golden/make_golden.py to
produce the committed golden vectors and by tests that pin the C oracle when the tree is present.
Nothing in the product path imports this file.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("LISFLOOD_REFERENCE", "/root/reference")
_R = os.path.join(REF_ROOT, "src", "lisflood")


def available():
    return os.path.isdir(os.path.join(_R, "hydrological_modules"))


_loaded = {}


def _load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load():
    """Returns (kwpt, kwp, soilloop) reference modules."""
    if _loaded:
        return _loaded["kwpt"], _loaded["kwp"], _loaded["sl"]
    import pandas  # noqa: F401  (must precede the numexpr stub: pandas probes numexpr's version)

    nx = types.ModuleType("numexpr")
    nx.__version__ = "stub"

    def _evaluate(expr, local_dict=None, global_dict=None):
        # numexpr resolves names in the caller's frame unless the dicts are given
        fr = sys._getframe(1)
        env = {}
        env.update(fr.f_globals if global_dict is None else global_dict)
        env.update(fr.f_locals if local_dict is None else local_dict)
        return eval(expr, {"__builtins__": {}}, env)

    nx.evaluate = _evaluate
    sys.modules["numexpr"] = nx
    nine = types.ModuleType("nine")
    nine.range = range
    sys.modules["nine"] = nine
    for name, sub in (("lisflood", ""), ("lisflood.global_modules", "/global_modules"),
                      ("lisflood.hydrological_modules", "/hydrological_modules")):
        m = types.ModuleType(name)
        m.__path__ = [_R + sub]
        sys.modules[name] = m
    sys.modules["lisflood.hydrological_modules"].HydroModule = object
    st = types.ModuleType("lisflood.global_modules.settings")
    for n in ("LisSettings", "MaskInfo", "EPICSettings"):
        setattr(st, n, type(n, (), {}))
    sys.modules["lisflood.global_modules.settings"] = st

    _load = _load_module
    _load("lisflood.global_modules.errors", _R + "/global_modules/errors.py")
    kwpt = _load("lisflood.hydrological_modules.kinematic_wave_parallel_tools",
                 _R + "/hydrological_modules/kinematic_wave_parallel_tools.py")
    kwp = _load("lisflood.hydrological_modules.kinematic_wave_parallel",
                _R + "/hydrological_modules/kinematic_wave_parallel.py")
    sl = _load("lisflood.hydrological_modules.soilloop", _R + "/hydrological_modules/soilloop.py")
    _loaded.update(kwpt=kwpt, kwp=kwp, sl=sl)
    return kwpt, kwp, sl
