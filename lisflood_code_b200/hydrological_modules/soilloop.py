"""soilloop -- HydroModule mirror (reference: src/lisflood/hydrological_modules/soilloop.py:437-704).

`self.var` is a lisflood_code_b200.hotpath.HotPathModel (the device-resident model object).  The reference runs
canopy, soil columns, open/sealed, per-pixel sums and groundwater as five module calls
(Lisflood_dynamic.py:114-149); on the device they are ONE fused stage (lf_model_soil), executed by the first of
the five calls of a step -- the others check the call order and return, because nothing of the hot path can
observe the intermediate states (the optional modules that could -- rice, water abstraction -- are out of scope).
"""
import ctypes as C

import numpy as np

from . import HydroModule
from .. import _capi


def _f64(a, name):
    a = np.asarray(a)
    if a.dtype != np.float64 or not a.flags.c_contiguous:
        raise TypeError("%s must be a C-contiguous float64 array (it is updated in place / read without a copy)" % name)
    return a


def interception_water_balance(Interception, TaInterception, LeafDrainage, CumInterception, LAI, Rain, TaInterceptionMax,
                               drainageK):
    """The reference's Numba kernel (hydrological_modules/soilloop.py:27-70) on the device: same eight arguments,
    (vegetation, pixel) float64 arrays updated in place, returns None."""
    V, N = np.shape(Interception)
    out = [_f64(a, n) for a, n in ((Interception, "Interception"), (TaInterception, "TaInterception"),
                                   (LeafDrainage, "LeafDrainage"), (CumInterception, "CumInterception"))]
    lai = np.ascontiguousarray(LAI, np.float64)
    rain = np.ascontiguousarray(Rain, np.float64)
    tmax = np.ascontiguousarray(TaInterceptionMax, np.float64)
    if lai.shape != (V, N) or tmax.shape != (V, N) or rain.shape != (N,):
        raise ValueError("interception_water_balance: shapes do not match (vegetation, pixel) = (%d, %d)" % (V, N))
    _capi.check(_capi.lib().lf_interception_water_balance(*[_capi.ptr(a) for a in out], _capi.ptr(lai), _capi.ptr(rain),
                                                          _capi.ptr(tmax), float(drainageK), V, N))


_SOIL_ARG_ORDER = (   # positional order of soilColumnsWaterBalance (soilloop.py:79-99)
    "index_landuse_all", "is_irrigated", "is_paddy_irrig", "paddy_inactive", "DtDay", "AvailableWaterForInfiltration", "Rain",
    "SnowMelt", "LeafDrainage", "Interception", "DSLR", "AvWaterThreshold", "ESAct", "ESMax", "isFrozenSoil", "b_Xinanjiang",
    "StoreMaxPervious", "PowerInfPot", "PrefFlow", "PowerPrefFlow", "Infiltration", "CourantCrit", "PoreSpaceNotZero1a",
    "PoreSpaceNotZero1b", "PoreSpaceNotZero2", "KSat1a", "KSat1b", "KSat2", "GenuInvM1a", "GenuInvM1b", "GenuInvM2", "GenuM1a",
    "GenuM1b", "GenuM2", "W1a", "W1b", "W1", "W2", "Theta1a", "Theta1b", "Theta2", "Sat1a", "Sat1b", "Sat1", "Sat2",
    "SeepTopToSubA", "SeepTopToSubB", "SeepSubToGW", "WRes1a", "WRes1b", "WRes1", "WRes2", "WWP1a", "WWP1b", "WWP1", "WWP2",
    "WFC1a", "WFC1b", "WFC1", "WFC2", "SoilDepth1a", "SoilDepth1b", "SoilDepth2", "WS1a", "WS1b", "WS1", "WS2", "UpperZoneK",
    "DrainedFraction", "GwPercStep", "UZOutflow", "UZ", "GwPercUZLZ")
_SOIL_SCALARS = ("DtDay", "AvWaterThreshold", "CourantCrit", "DrainedFraction")
_SOIL_IN_PLACE = ("AvailableWaterForInfiltration", "DSLR", "ESAct", "PrefFlow", "Infiltration", "W1a", "W1b", "W1", "W2",
                  "Theta1a", "Theta1b", "Theta2", "Sat1a", "Sat1b", "Sat1", "Sat2", "SeepTopToSubA", "SeepTopToSubB",
                  "SeepSubToGW", "UZOutflow", "UZ", "GwPercUZLZ")


def soilColumnsWaterBalance(*args, **kwargs):
    """The reference's Numba kernel (hydrological_modules/soilloop.py:78-355) on the device: the same 73 arguments in
    the same order (keywords accepted), in-place on the (vegetation, pixel) float64 arrays, returns None.  Paddy-rice
    fractions (EPIC) are out of scope: `is_paddy_irrig` must be all False."""
    if len(args) > len(_SOIL_ARG_ORDER):
        raise TypeError("soilColumnsWaterBalance takes %d arguments" % len(_SOIL_ARG_ORDER))
    a = dict(zip(_SOIL_ARG_ORDER, args))
    for k, v in kwargs.items():
        if k not in _SOIL_ARG_ORDER or k in a:
            raise TypeError("soilColumnsWaterBalance: unexpected or repeated argument %r" % k)
        a[k] = v
    missing = [k for k in _SOIL_ARG_ORDER if k not in a]
    if missing:
        raise TypeError("soilColumnsWaterBalance: missing arguments %s" % ", ".join(missing))
    V, N = np.shape(a["Interception"])
    S = _capi.SoilColumnsArgs()
    keep = []
    idx = np.ascontiguousarray(a["index_landuse_all"], np.int64)
    irr = np.ascontiguousarray(a["is_irrigated"]).astype(np.uint8)
    paddy = np.ascontiguousarray(a["is_paddy_irrig"]).astype(np.uint8)
    S.num_vegs, S.num_pixs, S.num_landuses = V, N, int(np.shape(a["KSat1a"])[0])
    S.index_landuse_all = idx.ctypes.data_as(_capi._I)
    S.is_irrigated, S.is_paddy_irrig = irr.ctypes.data_as(_capi._U), paddy.ctypes.data_as(_capi._U)
    keep += [idx, irr, paddy]
    for k in _SOIL_SCALARS:
        setattr(S, k, float(a[k]))
    for name, typ in _capi.SoilColumnsArgs._fields_:
        if name in _SOIL_SCALARS or name in ("num_vegs", "num_pixs", "num_landuses", "index_landuse_all", "is_irrigated",
                                             "is_paddy_irrig", "NoSubS_out"):
            continue
        if typ is _capi._U:
            arr = np.ascontiguousarray(a[name]).astype(np.uint8)
        elif name in _SOIL_IN_PLACE:
            arr = _f64(a[name], name)
        else:
            arr = np.ascontiguousarray(a[name], np.float64)
        keep.append(arr)
        setattr(S, name, arr.ctypes.data_as(typ))
    _capi.check(_capi.lib().lf_soil_columns_water_balance(C.byref(S)))


def suctionUnsaturatedSoilPF(index_landuse_all, pF0, pF1, pF2, W1a, W1b, W2, WRes1a, WRes1b, WRes2, WS1a, WS1b, WS2,
                             PoreSpaceNotZero1a, PoreSpaceNotZero1b, PoreSpaceNotZero2, GenuInvAlpha1a, GenuInvAlpha1b,
                             GenuInvAlpha2, GenuInvM1a, GenuInvM1b, GenuInvM2, GenuInvN1a, GenuInvN1b, GenuInvN2, HeadMax):
    """The reference's Numba kernel of option simulatePF (hydrological_modules/soilloop.py:402-424) on the device: the same
    26 arguments; pF0, pF1, pF2 -- (vegetation, pixel) float64 -- are written in place, returns None."""
    V, N = np.shape(pF0)
    S = _capi.SoilPfArgs()
    idx = np.ascontiguousarray(index_landuse_all, np.int64)
    keep = [idx]
    S.num_vegs, S.num_pixs, S.num_landuses = V, N, int(np.shape(WRes1a)[0])
    S.index_landuse_all = idx.ctypes.data_as(_capi._I)
    S.HeadMax = float(HeadMax)
    groups = {"pF": (pF0, pF1, pF2), "W": (W1a, W1b, W2), "WRes": (WRes1a, WRes1b, WRes2), "WS": (WS1a, WS1b, WS2),
              "PoreSpaceNotZero": (PoreSpaceNotZero1a, PoreSpaceNotZero1b, PoreSpaceNotZero2),
              "GenuInvAlpha": (GenuInvAlpha1a, GenuInvAlpha1b, GenuInvAlpha2), "GenuInvM": (GenuInvM1a, GenuInvM1b, GenuInvM2),
              "GenuInvN": (GenuInvN1a, GenuInvN1b, GenuInvN2)}
    for name, arrays in groups.items():
        for layer, a in enumerate(arrays):
            if name == "pF":
                arr = _f64(a, "pF%d" % layer)
            elif name == "PoreSpaceNotZero":
                arr = np.ascontiguousarray(a).astype(np.uint8)
            else:
                arr = np.ascontiguousarray(a, np.float64)
            want = (V, N) if name in ("pF", "W") else (S.num_landuses, N)
            if arr.shape != want:
                raise ValueError("suctionUnsaturatedSoilPF: %s of layer %d has shape %s, expected %s" % (name, layer, arr.shape, want))
            keep.append(arr)
            getattr(S, name)[layer] = arr.ctypes.data_as(_capi._U if name == "PoreSpaceNotZero" else _capi._D)
    _capi.check(_capi.lib().lf_suction_unsaturated_soil_pf(C.byref(S)))


class soilloop(HydroModule):
    input_files_keys = {'wateruse': []}
    module_name = 'SoilLoop'

    def __init__(self, soilloop_variable):
        self.var = soilloop_variable

    def initial(self):
        pass

    def dynamic_canopy(self):
        self.var._soil_stage_call("dynamic_canopy")

    def dynamic_soil(self):
        self.var._soil_stage_call("dynamic_soil")
