"""Loads the UNMODIFIED settings module of the reference (global_modules/settings.py) under stand-ins for the packages it
imports but the XML parsing does not use (future, nine, pcraster, netCDF4, cftime).  TEST INFRASTRUCTURE ONLY (build
container; needs /root/reference): tests/test_settings_live_reference.py compares the product's settings-XML parser with it
on the reference's own shipped settings files."""
import importlib.util
import os
import sys
import types
from collections import OrderedDict

from . import ref_loader

_R = ref_loader._R
_mod = None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Unavailable(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return None


def load():
    """The reference's settings module (its LisSettings class is NOT instantiated: __init__ needs the calendar machinery;
    `parse` below calls the XML methods on a bare object)."""
    global _mod
    if _mod is not None:
        return _mod
    def ensure(name, **attrs):        # other harness modules may have installed barer stand-ins already: complete them
        m = sys.modules.get(name)
        if m is None:
            return _stub(name, **attrs)
        for k, v in attrs.items():
            if not hasattr(m, k):
                setattr(m, k, v)
        return m
    fb = ensure("future.backports", OrderedDict=OrderedDict)
    fu = ensure("future.utils", with_metaclass=lambda meta, *bases: meta("NewBase", bases or (object,), {}))
    ensure("future", backports=fb, utils=fu)
    ensure("nine", iteritems=lambda d: d.items(), str=str, range=range, map=map, nine=lambda c: c)
    for n in ("pcraster", "netCDF4", "cftime"):
        if n not in sys.modules:
            sys.modules[n] = _Unavailable(n)
    pk = types.ModuleType("lisflood_ref_settings")
    pk.__path__ = [_R]
    sys.modules["lisflood_ref_settings"] = pk
    gm = types.ModuleType("lisflood_ref_settings.global_modules")
    gm.__path__ = [os.path.join(_R, "global_modules")]
    sys.modules["lisflood_ref_settings.global_modules"] = gm

    def _load(sub):
        name = "lisflood_ref_settings.global_modules." + sub
        spec = importlib.util.spec_from_file_location(name, os.path.join(_R, "global_modules", sub + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m
    for sub in ("errors", "decorators", "default_options"):
        _load(sub)
    _mod = _load("settings")
    return _mod


def parse(settings_file, sys_args=()):
    """(user variables, bindings, options, flags) of a settings file as the reference's own methods produce them
    (settings.py:502-607)."""
    import xml.dom.minidom
    st = load()
    dom = xml.dom.minidom.parse(settings_file)
    obj = object.__new__(st.LisSettings)
    obj.settings_dir = os.path.normpath(os.path.dirname(os.path.abspath(settings_file)))
    user, binding = obj._bindings(dom)
    return user, binding, st.LisSettings._options(dom), dict(st.LisSettings._flags(list(sys_args)))
