"""Runs the UNMODIFIED hot-path module classes of the reference (routing, surface_routing, soilloop,
soil, opensealed, groundwater: their dynamic*() methods) on a synthetic model object.
TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).

The classes import PCRaster / netCDF4 / xarray at module level but their dynamic() parts are NumPy +
Numba only, so the unavailable packages are replaced by empty stubs and `LisSettings` / `MaskInfo` by
minimal stand-ins.  The reference tree is never modified.  The model object (`RefVar`) carries the
reference's attribute names and restates the small index helpers of
Lisflood_initial.py:266-396 that the modules call.
"""
import sys
import types
from collections import OrderedDict

import numpy as np

from . import ref_loader

_R = ref_loader._R
_mods = {}


class _Options(dict):
    def __missing__(self, key):
        return False


class _FakeSettings(object):
    _inst = None

    def __init__(self):
        self.options = _Options()
        self.binding = {}
        self.flags = {"nancheck": False}

    @classmethod
    def instance(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst


class _FakeMaskInfo(object):
    _inst = None

    class _Info(object):
        pass

    @classmethod
    def instance(cls):
        return cls._inst

    @classmethod
    def set(cls, land_mask):
        m = cls()
        m.info = cls._Info()
        m.info.mask = ~np.asarray(land_mask, bool)
        m.info.mapC = (int(np.asarray(land_mask).sum()),)
        cls._inst = m
        return m

    def in_zero(self):
        return np.zeros(self.info.mapC[0])


class _EPIC(object):
    _inst = None

    def __init__(self):
        self.soil_uses = ["Rainfed", "Forest", "Irrigated"]
        self.prescribed_vegetation = [f + "_prescribed" for f in self.soil_uses]
        self.vegetation_landuse = OrderedDict(zip(self.prescribed_vegetation, self.soil_uses))
        self.landuse_vegetation = OrderedDict([(v, [k]) for k, v in self.vegetation_landuse.items()])

    @classmethod
    def instance(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst


class _AnyModule(types.ModuleType):
    """Module whose every attribute is a do-nothing callable (stands in for pcraster etc.)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _unavailable(*a, **k):
            raise RuntimeError("%s.%s is not available in the oracle harness" % (self.__name__, name))
        return _unavailable


class _NoOpSubModule(object):
    def __init__(self, var):
        self.var = var

    def __getattr__(self, name):
        return lambda *a, **k: None


def load():
    """Returns dict name -> reference class (routing, surface_routing, soilloop, soil, opensealed, groundwater)."""
    if _mods:
        return _mods
    kwpt, kwp, sl_kernels = ref_loader.load()  # also installs the lisflood package stubs
    for name in ("pcraster", "pcraster.operations", "pcraster.framework", "xarray", "netCDF4", "cftime", "future",
                 "pyproj"):
        if name not in sys.modules:
            sys.modules[name] = _AnyModule(name)
    st = sys.modules["lisflood.global_modules.settings"]
    st.LisSettings, st.MaskInfo, st.EPICSettings = _FakeSettings, _FakeMaskInfo, _EPIC
    add1 = types.ModuleType("lisflood.global_modules.add1")
    for n in ("loadmap", "loadmap_base", "compressArray", "decompress", "makenumpy", "defsoil", "readnetcdf"):
        setattr(add1, n, None)
    sys.modules["lisflood.global_modules.add1"] = add1
    hm = sys.modules["lisflood.hydrological_modules"]
    for sub in ("polder", "inflow", "transmission"):
        m = types.ModuleType("lisflood.hydrological_modules." + sub)
        setattr(m, sub, _NoOpSubModule)
        sys.modules[m.__name__] = m
        setattr(hm, sub, m)
    # lakes / reservoir: the real classes (their dynamic_inloop() return at once unless the option is switched on)
    for sub in ("lakes", "reservoir"):
        m = ref_loader._load_module("lisflood.hydrological_modules." + sub, _R + "/hydrological_modules/%s.py" % sub)
        setattr(hm, sub, m)
        _mods[sub] = getattr(m, sub)
    # soilloop was imported by ref_loader before the settings stand-ins existed: rebind its globals
    sl_kernels.LisSettings, sl_kernels.MaskInfo, sl_kernels.EPICSettings = _FakeSettings, _FakeMaskInfo, _EPIC
    _mods["soilloop"] = sl_kernels.soilloop
    _mods["soilloop_module"] = sl_kernels
    for name in ("routing", "surface_routing", "soil", "opensealed", "groundwater"):
        m = ref_loader._load_module("lisflood.hydrological_modules." + name, _R + "/hydrological_modules/%s.py" % name)
        _mods[name] = getattr(m, name)
    return _mods


def load_feeders():
    """The reference's snow and frost classes (feeder modules): hydrological_modules/snow.py,
    frost.py.  Their dynamic() parts are NumPy only."""
    load()
    out = {}
    for name in ("snow", "frost"):
        key = "lisflood.hydrological_modules." + name
        m = sys.modules.get(key) or ref_loader._load_module(key, _R + "/hydrological_modules/%s.py" % name)
        out[name] = getattr(m, name)
    return out


class _DA(np.ndarray):
    """(vegetation, pixel) array with the tiny xarray surface dynamic_canopy uses on LAITerm."""

    def __new__(cls, a):
        return np.asarray(a).view(cls)

    def sel(self, **kw):
        return self

    @property
    def values(self):
        return np.asarray(self)


def numpy_modified(a, dims):
    from lisflood_code_b200.global_modules.add1 import NumpyModified
    return NumpyModified(np.ascontiguousarray(a), dims)


class RefVar(object):
    """Synthetic stand-in for the reference's model object, built from synthetic.full_stack()."""
    VEG_DIMS = ["vegetation", "pixel"]
    LU_DIMS = ["landuse", "pixel"]

    def __init__(self, S, options=None):
        load()
        sett = _FakeSettings.instance()
        sett.options.clear()
        sett.options.update(options or {})
        sett.options.setdefault("SplitRouting", bool(S.get("SplitRouting")))
        self.maskinfo = _FakeMaskInfo.set(S["mask"])
        self.settings = sett
        self.epic_settings = _EPIC.instance()
        self.SOIL_USES = list(self.epic_settings.soil_uses)
        self.PRESCRIBED_VEGETATION = list(self.epic_settings.prescribed_vegetation)
        self.VEGETATION_LANDUSE = OrderedDict(zip(self.PRESCRIBED_VEGETATION, self.SOIL_USES))
        self.LANDUSE_VEGETATION = OrderedDict([(v, [k]) for k, v in self.VEGETATION_LANDUSE.items()])
        self.prescribed_vegetation = list(self.PRESCRIBED_VEGETATION)
        self.vegetation = list(self.PRESCRIBED_VEGETATION)
        n = S["N"]
        vn = lambda: numpy_modified(np.zeros((3, n)), self.VEG_DIMS)
        for k, v in S.items():
            if isinstance(v, np.ndarray) and v.ndim == 2 and v.shape == (3, n) and k not in ("mask",):
                dims = self.VEG_DIMS if k in ("SoilFraction", "W1a", "W1b", "W1", "W2", "UZ", "DSLR", "CumInterception") \
                    else (["runoff", "pixel"] if k in ("OFAlpha", "InvOFAlpha") else self.LU_DIMS)
                setattr(self, k, numpy_modified(v.copy(), dims))
            elif isinstance(v, np.ndarray):
                setattr(self, k, v.copy())
            else:
                setattr(self, k, v)
        for k in ("Interception", "TaInterception", "LeafDrainage", "potential_transpiration", "Ta", "ESAct", "PrefFlow",
                  "Infiltration", "SeepTopToSubA", "SeepTopToSubB", "SeepSubToGW", "Theta", "Theta1a", "Theta1b", "Theta2",
                  "Sat1a", "Sat1b", "Sat1", "Sat2", "AvailableWaterForInfiltration", "SoilMoistureStressDays", "RWS",
                  "GwPercUZLZ", "UZOutflow", "LAI"):
            setattr(self, k, vn())
        for k in ("TaCUM", "TaInterceptionCUM", "ESActCUM", "GwLossCUM", "sumDis", "CumQ", "DischargeM3Out", "WaterDepth"):
            setattr(self, k, np.zeros(n))
        self.TimeSinceStart = 0
        self.dim_runoff = ("runoff", ["Other", "Forest", "Direct"])
        self.dim_landuse = ("landuse", self.SOIL_USES[:])
        self.dim_pixel = ("pixel", np.arange(n))
        self.num_pixel = n

    # ---- helpers restated from Lisflood_initial.py:266-396 ----
    def allocateDataArray(self, dimensions, dtype=float):
        coords = OrderedDict(dimensions)
        return numpy_modified(np.zeros([len(v) for v in coords.values()], dtype), coords.keys())

    def allocateVariableAllVegetation(self, dtype=float):
        return self.allocateDataArray([("vegetation", self.vegetation[:]), self.dim_pixel], dtype)

    def get_landuse_and_indexes_from_vegetation_epic(self, veg):
        iveg = self.vegetation.index(veg)
        landuse = self.epic_settings.vegetation_landuse[veg]
        return iveg, self.epic_settings.soil_uses.index(landuse), landuse

    def get_indexes_from_landuse_and_veg_list_GLOBAL(self, landuse, veg_list):
        ilanduse = self.SOIL_USES.index(landuse)
        return ([self.vegetation.index(v) for v in veg_list], [self.PRESCRIBED_VEGETATION.index(v) for v in veg_list],
                ilanduse)

    def deffraction(self, variable):
        ax = variable.dims.index("vegetation")
        return (self.SoilFraction.values * variable.values).sum(ax)

    def set_forcing(self, F):
        self.Rain, self.SnowMelt = F["Rain"].copy(), F["SnowMelt"].copy()
        self.ETRef, self.EWRef, self.ESRef = F["ETRef"].copy(), F["EWRef"].copy(), F["ESRef"].copy()
        self.LAI.values[:] = F["LAI"]
        self.LAITerm = _DA(F["LAITerm"].copy())
        self.isFrozenSoil = F["isFrozenSoil"].copy()


class RefModel(object):
    """The reference's per-step call order for the hot path (Lisflood_dynamic.py:114-229)."""

    def __init__(self, S, options=None):
        M = load()
        self.var = RefVar(S, options)
        v = self.var
        self.soilloop = M["soilloop"](v)
        self.soilloop.initial()
        self.soil = M["soil"](v)
        self.opensealed = M["opensealed"](v)
        self.groundwater = M["groundwater"](v)
        self.surface_routing = M["surface_routing"](v)
        self.routing = M["routing"](v)
        kwp = sys.modules["lisflood.hydrological_modules.kinematic_wave_parallel"]
        land = S["mask"]
        # surface_routing.initialSecond / routing.initialSecond (need compressArray): restated calls
        sr = self.surface_routing
        idx = v.dim_runoff[1].index
        sr.direct_surface_router = kwp.kinematicWave(v.LddToChan, land, v.OFAlpha.values[idx("Direct")], v.Beta, v.PixelLength, v.DtSec)
        sr.other_surface_router = kwp.kinematicWave(v.LddToChan, land, v.OFAlpha.values[idx("Other")], v.Beta, v.PixelLength, v.DtSec)
        sr.forest_surface_router = kwp.kinematicWave(v.LddToChan, land, v.OFAlpha.values[idx("Forest")], v.Beta, v.PixelLength, v.DtSec)
        self.routing.river_router = kwp.kinematicWave(v.LddKinematic, land, v.ChannelAlpha, v.Beta, v.ChanLength, v.DtRouting,
                                                      alpha_floodplains=getattr(v, "ChannelAlpha2", None))

    def step(self, F):
        v = self.var
        v.TimeSinceStart += 1
        v.set_forcing(F)
        self.soilloop.dynamic_canopy()
        self.soilloop.dynamic_soil()
        self.opensealed.dynamic()
        self.soil.dynamic_perpixel()
        self.groundwater.dynamic()
        self.surface_routing.dynamic()
        # Lisflood_dynamic.py:176-229
        v.sumDisDay = v.maskinfo.in_zero()
        for s in range(v.NoRoutSteps):
            self.routing.dynamic(s)
        opt = v.settings.options
        if opt["InitLisflood"] or not opt["SplitRouting"]:
            v.ChanM3 = v.ChanM3Kin.copy()
        else:
            v.ChanM3 = v.ChanM3Kin + v.Chan2M3Kin - v.Chan2M3Start
        v.TotalCrossSectionArea = v.ChanM3 * v.InvChanLength
        v.sumDis += v.sumDisDay
        v.ChanQAvg = v.sumDisDay / v.NoRoutSteps
        v.DischargeM3Out += np.where(v.AtLastPointC, v.ChanQ * v.DtSec, 0)
