"""Runs the UNMODIFIED `initial()` / `initialSecond()` methods of the reference's soil and routing modules on
synthetic inputs, to pin the init-time parameter derivation.
TEST INFRASTRUCTURE ONLY (build container; needs /root/reference) -- used by tests/golden/make_golden.py.

What is the reference's own arithmetic here: every NumPy expression of soil.initial (hydrological_modules/soil.py:76-470)
and of routing.initial / initialSecond (routing.py:61-397): van Genuchten storages, channel geometry, kinematic-wave
alpha, initial volume / discharge, split-routing limits.  What is NOT: PCRaster.  The ldd operators the reference
calls (lddmask, lddrepair, accuflux, downstream, upstream, pit, catchment ...) are stubbed by the NumPy restatements of
lisflood_code_b200/global_modules/ldd_ops.py working on compressed 1-D arrays, so the drainage-network outputs
(LddKinematic, LddToChan, UpArea, Catchments, downstruct) are pinned only to the PCRaster manual's semantics.
"""
import sys
from collections import OrderedDict

import numpy as np

from . import ref_modules


class _Pcr(np.ndarray):
    """A 'PCRaster map' of the harness: compressed 1-D array."""

    def __new__(cls, a):
        return np.asarray(a).view(cls)


def _install(raw, land_mask):
    """Binds loadmap / compressArray / decompress / the pcraster operators in the loaded reference modules."""
    from lisflood_code_b200.global_modules import ldd_ops
    M = ref_modules.load()
    land = np.asarray(land_mask, bool)
    n = int(land.sum())

    def loadmap(name, pcr=False, lddflag=False, **kw):
        v = raw[name]
        if np.ndim(v) == 0:
            return float(v)
        return _Pcr(np.array(v, np.float64)) if pcr else np.array(v, np.float64)

    def compress(m, *a, **k):
        return np.asarray(m).copy()

    def decompress(a, *aa, **k):
        return _Pcr(np.asarray(a))

    def lddmask(ldd, keep):
        return _Pcr(ldd_ops.lddmask_codes(np.asarray(ldd, np.float64), np.asarray(keep) != 0, land))

    def lddrepair(ldd):
        return _Pcr(ldd_ops.lddrepair_codes(np.asarray(ldd, np.float64), land))

    def ds_of(ldd):
        return ldd_ops.downstream_index(np.asarray(ldd, np.float64), land)

    def accuflux(ldd, x):
        return _Pcr(ldd_ops.accuflux(ds_of(ldd), np.broadcast_to(np.asarray(x, np.float64), (n,))))

    def boolean(x):
        return _Pcr(np.asarray(x) != 0)

    def pit(ldd):
        d = ds_of(ldd)
        out = np.zeros(n, np.int64)
        p = np.flatnonzero(d < 0)
        out[p] = np.arange(1, p.size + 1)
        return _Pcr(out)

    def ifthenelse(c, a, b):
        return _Pcr(np.where(np.asarray(c) != 0, a, b))

    def downstream(ldd, x):
        d = ds_of(ldd)
        x = np.asarray(x)
        return _Pcr(np.where(d >= 0, x[np.maximum(d, 0)], x))

    def upstream(ldd, x):
        return _Pcr(ldd_ops.upstream_sum(ds_of(ldd), np.asarray(x, np.float64)))

    def uniqueid(b):
        b = np.asarray(b) != 0
        out = np.zeros(n, np.int64)
        out[b] = np.arange(1, int(b.sum()) + 1)
        return _Pcr(out)

    def nominal(x):
        return _Pcr(np.asarray(x).astype(np.int64))

    def catchment(ldd, points):
        d = ds_of(ldd)
        root = np.where(d >= 0, d, np.arange(n))
        while True:
            nxt = root[root]
            if np.array_equal(nxt, root):
                break
            root = nxt
        return _Pcr(np.asarray(points)[root])

    soil_mod = sys.modules["lisflood.hydrological_modules.soil"]
    rout_mod = sys.modules["lisflood.hydrological_modules.routing"]
    soil_mod.loadmap = loadmap

    def makenumpy(m):   # global_modules/add1.py: scalar -> array over the mask
        return np.zeros(n) + m if np.ndim(m) == 0 else np.asarray(m, np.float64)
    for name in ("surface_routing", "groundwater"):
        mod = sys.modules["lisflood.hydrological_modules." + name]
        mod.loadmap, mod.loadmap_base, mod.makenumpy, mod.compressArray, mod.decompress = (loadmap, loadmap, makenumpy,
                                                                                            compress, decompress)
    for k, f in dict(loadmap=loadmap, loadmap_base=loadmap, compressArray=compress, decompress=decompress, lddmask=lddmask,
                     lddrepair=lddrepair, accuflux=accuflux, boolean=boolean, pit=pit, ifthenelse=ifthenelse,
                     downstream=downstream, upstream=upstream, uniqueid=uniqueid, nominal=nominal,
                     catchment=catchment).items():
        setattr(rout_mod, k, f)
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_   # the reference predates NumPy 2
    return M, loadmap


class _InitVar(object):
    """Model object of the reference during initial(): helper API restated from Lisflood_initial.py:266-396."""

    def __init__(self, land_mask, raw, loadmap, options, DtSec, DtSecChannel):
        from lisflood_code_b200.global_modules.add1 import NumpyModified
        self._NM = NumpyModified
        self._loadmap, self._raw = loadmap, raw
        sett = ref_modules._FakeSettings.instance()
        sett.options.clear()
        sett.options.update(options)
        self.settings = sett
        self.maskinfo = ref_modules._FakeMaskInfo.set(land_mask)
        n = int(np.asarray(land_mask).sum())
        self.SOIL_USES = ["Rainfed", "Forest", "Irrigated"]
        self.PRESCRIBED_VEGETATION = [u + "_prescribed" for u in self.SOIL_USES]
        self.prescribed_vegetation = list(self.PRESCRIBED_VEGETATION)
        self.vegetation = list(self.PRESCRIBED_VEGETATION)
        self.VEGETATION_LANDUSE = OrderedDict(zip(self.PRESCRIBED_VEGETATION, self.SOIL_USES))
        self.dim_pixel = ("pixel", np.arange(n))
        self.dim_landuse = ("landuse", self.SOIL_USES[:])
        self.dim_runoff = ("runoff", ["Other", "Forest", "Direct"])
        self.coord_landuse = OrderedDict([self.dim_landuse, self.dim_pixel])
        self.DtSec, self.DtSecChannel = float(DtSec), float(DtSecChannel)
        self.DtDay = self.DtSec / 86400.0

    def allocateDataArray(self, dimensions, dtype=float):
        coords = OrderedDict(dimensions)
        return self._NM(np.zeros([len(v) for v in coords.values()], dtype), coords.keys())

    def allocateVariableAllVegetation(self, dtype=float):
        return self.allocateDataArray([("vegetation", self.vegetation[:]), self.dim_pixel], dtype)

    def defsoil(self, name_1, name_2=None, name_3=None, coords=None):   # Lisflood_initial.py:371-391 (non-EPIC branch)
        if coords is None:
            coords = self.coord_landuse
        data = self.allocateDataArray(coords)

        def read(name, backup=None):   # readInputWithBackup
            if name is None:
                return backup
            if isinstance(name, str):
                return self._loadmap(name) if name in self._raw else backup
            return float(name)
        v1 = read(name_1)
        data.values[0][:] = v1
        data.values[1][:] = read(name_2, v1)
        data.values[2][:] = read(name_3, v1)
        return data


def _collect(var):
    out = {}
    for k, v in var.__dict__.items():
        if k.startswith("_") or k in ("settings", "maskinfo"):
            continue
        if isinstance(v, (float, int, np.floating, np.integer)) and not isinstance(v, bool):
            out[k] = np.float64(v)
        elif isinstance(v, np.ndarray) and v.dtype.kind in "fiub":
            out[k] = np.asarray(v)
    return out


def soil_initial(land_mask, raw, state, options=None, DtSec=86400.0):
    """Outputs of the reference's soil.initial() as {attribute: array}.  `state`: attributes other modules set before
    (SoilFraction (3,N), RiceFraction, WaterFraction, OtherFraction, IrrigationFraction, ForestFraction,
    DirectRunoffFraction)."""
    M, loadmap = _install(raw, land_mask)
    var = _InitVar(land_mask, raw, loadmap, dict(options or {}), DtSec, 3600.0)
    for k, v in state.items():
        setattr(var, k, var._NM(np.array(v, np.float64), ["vegetation", "pixel"]) if np.ndim(v) == 2 else np.array(v, np.float64))
    before = set(var.__dict__)
    M["soil"](var).initial()
    return {k: v for k, v in _collect(var).items() if k not in before or k == "SoilFraction"}


def routing_initial(land_mask, raw, state, options=None, DtSec=86400.0, DtSecChannel=3600.0):
    """Outputs of the reference's routing.initial() + initialSecond() (the kinematicWave object it builds is dropped)."""
    M, loadmap = _install(raw, land_mask)
    opts = dict(options or {})
    var = _InitVar(land_mask, raw, loadmap, opts, DtSec, DtSecChannel)
    n = int(np.asarray(land_mask).sum())
    var.MaskMap = _Pcr(np.ones(n, bool))
    for k, v in state.items():
        setattr(var, k, np.array(v, np.float64))
    var.PixelAreaPcr = _Pcr(var.PixelArea)
    before = set(var.__dict__)
    r = M["routing"](var)
    r.initial()
    r.initialSecond()
    out = {k: v for k, v in _collect(var).items() if k not in before}
    return out


def surface_and_groundwater_initial(land_mask, raw, state, options=None, DtSec=86400.0):
    """Outputs of the reference's surface_routing.initial() and groundwater.initial().  `state`: what miscInitial,
    soil.initial and routing.initial leave on the model object (PixelLength, InvPixelLength, MMtoM, NManning (3,N), Beta,
    InvBeta, AlpPow, GwPerc, GwLoss)."""
    M, loadmap = _install(raw, land_mask)
    var = _InitVar(land_mask, raw, loadmap, dict(options or {}), DtSec, 3600.0)
    for k, v in state.items():
        if np.ndim(v) == 2:
            setattr(var, k, var._NM(np.array(v, np.float64), ["runoff", "pixel"]))
        elif np.ndim(v) == 1:
            setattr(var, k, np.array(v, np.float64))
        else:
            setattr(var, k, float(v))
    before = set(var.__dict__)
    M["surface_routing"](var).initial()
    M["groundwater"](var).initial()
    return {k: v for k, v in _collect(var).items() if k not in before}


def structures_initial(land_mask, raw, tables, state, options=None, DtSec=86400.0, DtSecChannel=3600.0):
    """Outputs of the reference's reservoir.initial() and lakes.initial() (reservoir.py:52-170, lakes.py:52-196).
    `tables`: binding name -> two-column array [site id, value] (PCRaster lookup tables); `state`: IsChannel,
    IsStructureKinematic, LddKinematic, downstruct, ChanQ, DtRouting as routing.initial leaves them.  The PCRaster calls
    (ifthen, defined, boolean, cover, downstream, lookupscalar) are stubbed on compressed arrays like in _install()."""
    from lisflood_code_b200.global_modules import ldd_ops
    from lisflood_code_b200.hydrological_modules.reservoir import lookupscalar as lookup_restated
    M, loadmap = _install(raw, land_mask)
    land = np.asarray(land_mask, bool)
    n = int(land.sum())
    opts = dict(options or {})
    var = _InitVar(land_mask, raw, loadmap, opts, DtSec, DtSecChannel)
    var.settings.binding = {k: k for k in tables}
    for k, v in state.items():
        setattr(var, k, np.array(v) if np.ndim(v) else v)

    def makenumpy(m):
        return np.zeros(n) + m if np.ndim(m) == 0 else np.asarray(m, np.float64)

    def compress(m, *a, **k):
        return np.asarray(m, np.float64).copy()

    def decompress(a, *aa, **k):
        return _Pcr(np.asarray(a))

    fns = dict(
        ifthen=lambda c, x: _Pcr(np.where(np.asarray(c) != 0, np.asarray(x, np.float64), np.nan)),
        defined=lambda x: _Pcr(~np.isnan(np.asarray(x, np.float64))),
        boolean=lambda x: _Pcr(np.nan_to_num(np.asarray(x, np.float64)) != 0),
        cover=lambda x, y: _Pcr(np.where(np.isnan(np.asarray(x, np.float64)), y, x)),
        lookupscalar=lambda name, sites: _Pcr(lookup_restated(tables[str(name)], np.asarray(sites, np.float64))),
        downstream=lambda ldd, x: _Pcr(np.where(ldd_ops.downstream_index(np.asarray(ldd, np.float64), land) >= 0,
                                                np.asarray(x)[np.maximum(ldd_ops.downstream_index(np.asarray(ldd, np.float64), land), 0)],
                                                np.asarray(x))))
    pcr = sys.modules["pcraster"]
    for name in ("reservoir", "lakes"):
        mod = sys.modules["lisflood.hydrological_modules." + name]
        mod.loadmap, mod.compressArray, mod.decompress, mod.makenumpy = loadmap, compress, decompress, makenumpy
        for k, f in fns.items():
            setattr(mod, k, f)
    saved = {k: pcr.__dict__.get(k) for k in fns}
    for k, f in fns.items():
        setattr(pcr, k, f)
    try:
        before = set(var.__dict__)
        M["lakes"](var).initial()
        M["reservoir"](var).initial()
    finally:
        for k, f in saved.items():
            if f is None:
                pcr.__dict__.pop(k, None)
            else:
                setattr(pcr, k, f)
    out = {}
    for k, v in var.__dict__.items():
        if k in before and k != "IsStructureKinematic":
            continue
        if isinstance(v, tuple):                       # reservoir.py:155: a stray comma makes the cold-start fill a 1-tuple
            v = np.asarray(v[0])
        if isinstance(v, np.ndarray) and v.dtype.kind in "fiub":
            out[k] = np.squeeze(np.asarray(v)) if v.ndim == 2 and v.shape[0] == 1 else np.asarray(v)
    return out


def misc_and_landuse_initial(land_mask, raw, options=None, cell=None):
    """Outputs of the reference's miscInitial.initial() (hydrological_modules/miscInitial.py:44-133) followed by
    landusechange.initial() (landusechange.py:53-93, static maps).  options['gridSizeUserDefined'] chooses between the
    'PixelLengthUser' / 'PixelAreaUser' inputs and the cell size of the mask map (`cell`, what MaskAttrs holds)."""
    import types
    M, loadmap = _install(raw, land_mask)
    opts = dict(options or {})
    st = sys.modules["lisflood.global_modules.settings"]

    class _MaskAttrs(object):
        @classmethod
        def instance(cls):
            return {"cell": cell}
    st.MaskAttrs, st.calendar = _MaskAttrs, (lambda date, calendar_type: date)
    if "lisflood.global_modules.netcdf" not in sys.modules:
        sys.modules["lisflood.global_modules.netcdf"] = types.ModuleType("lisflood.global_modules.netcdf")
        setattr(sys.modules["lisflood.global_modules"], "netcdf", sys.modules["lisflood.global_modules.netcdf"])
    n_pix = int(np.asarray(land_mask).sum())
    sys.modules["lisflood.global_modules.netcdf"].read_lat_from_template = lambda binding: np.asarray(
        raw.get("lat_deg", np.zeros(n_pix)), np.float64)
    ref_modules._FakeSettings.instance().binding.update(CalendarDayStart="01/01/1990", calendar_type="proleptic_gregorian")
    pcr = sys.modules["pcraster"]
    pcr.celllength, pcr.scalar = (lambda: cell), (lambda x: x)
    add1 = sys.modules["lisflood.global_modules.add1"]
    from lisflood_code_b200.global_modules.add1 import NumpyModified
    # (a writable copy: with today's pandas, DataFrame.T.values is read-only and the reference assigns into it)
    add1.loadmap, add1.compressArray, add1.NumpyModified, add1.readnetcdf = loadmap, (lambda m, *a, **k: np.asarray(m).copy()), \
        (lambda a, dims=None: NumpyModified(np.array(a), dims)), None
    epic = ref_modules._EPIC.instance()
    epic.landuse_inputmap = OrderedDict(zip(epic.landuse_vegetation.keys(), ["OtherFraction", "ForestFraction",
                                                                               "IrrigationFraction"]))   # settings.py:343-345
    mods = {}
    for name in ("miscInitial", "landusechange"):
        mods[name] = ref_modules.ref_loader._load_module("lisflood.hydrological_modules." + name,
                                                         ref_modules._R + "/hydrological_modules/%s.py" % name)
    var = _InitVar(land_mask, raw, loadmap, opts, 0.0, 0.0)
    del var.DtSec, var.DtSecChannel, var.DtDay           # miscInitial sets them from the bindings
    var.coord_prescribed_vegetation = OrderedDict([("vegetation", var.prescribed_vegetation[:]), var.dim_pixel])
    before = set(var.__dict__)
    mods["miscInitial"].miscInitial(var).initial()
    mods["landusechange"].landusechange(var).initial()
    return {k: v for k, v in _collect(var).items() if k not in before}


def feeders_initial(land_mask, raw, options=None):
    """Outputs of the reference's snow.initial() (hydrological_modules/snow.py:53-93), frost.initial() (frost.py:44-58) and
    leafarea.initial() (leafarea.py:44-72; the 36 x 3 prescribed LAI maps are not read: loadLAI is stubbed)."""
    import types
    M, loadmap = _install(raw, land_mask)
    F = ref_modules.load_feeders()
    for name in ("snow", "frost"):
        sys.modules["lisflood.hydrological_modules." + name].loadmap = loadmap
    n = int(np.asarray(land_mask).sum())

    class _Stack(object):                       # the surface of xarray.DataArray leafarea.initial touches
        def __init__(self, data, coords=None, dims=None):
            self.interval = types.SimpleNamespace(values=np.arange(data.shape[0]))
            self.loc = self

        def __setitem__(self, key, value):
            pass
    add1 = sys.modules["lisflood.global_modules.add1"]
    add1.loadmap = loadmap
    zus = sys.modules.get("lisflood.global_modules.zusatz")
    if zus is None:
        zus = types.ModuleType("lisflood.global_modules.zusatz")
        sys.modules["lisflood.global_modules.zusatz"] = zus
        setattr(sys.modules["lisflood.global_modules"], "zusatz", zus)
    for k in ("generateName", "loadLAI"):
        if not hasattr(zus, k):
            setattr(zus, k, lambda *a, **kw: None)
        if not hasattr(add1, k):
            setattr(add1, k, lambda *a, **kw: None)
    sys.modules["xarray"].DataArray = _Stack
    key = "lisflood.hydrological_modules.leafarea"
    lmod = sys.modules.get(key) or ref_modules.ref_loader._load_module(key, ref_modules._R + "/hydrological_modules/leafarea.py")
    lmod.loadmap, lmod.loadLAI, lmod.generateName = loadmap, (lambda *a, **kw: np.zeros(n)), (lambda *a, **kw: "")
    var = _InitVar(land_mask, raw, loadmap, dict(options or {}), 86400.0, 3600.0)
    var.coord_prescribed_vegetation = OrderedDict([("vegetation", var.prescribed_vegetation[:]), var.dim_pixel])
    var.PRESCRIBED_LAI = OrderedDict(zip(var.prescribed_vegetation, ["LAIOtherMaps", "LAIForestMaps", "LAIIrrigationMaps"]))
    ref_modules._FakeSettings.instance().binding.update(LAIOtherMaps="", LAIForestMaps="", LAIIrrigationMaps="")
    before = set(var.__dict__)
    F["snow"](var).initial()
    F["frost"](var).initial()
    lmod.leafarea(var).initial()
    out = {k: v for k, v in _collect(var).items() if k not in before}
    out["SnowCoverS"] = np.stack([np.zeros(n) + x for x in var.SnowCoverS])
    out["L1"] = np.asarray(var.L1)
    return out


def structures_module_initial(land_mask, ldd_kinematic, is_structure, options=None):
    """Outputs of the reference's structures.initial() (hydrological_modules/structures.py:43-61) for a channel network and
    the structure pixels reservoir.initial() / lakes.initial() marked: LddStructuresKinematic, IsUpsOfStructureKinematicC and
    the cut LddKinematic.  downstream / lddrepair / ifthenelse / cover / boolean are the stand-ins of _install()."""
    from lisflood_code_b200.global_modules import ldd_ops
    M, loadmap = _install({}, land_mask)
    land = np.asarray(land_mask, bool)
    key = "lisflood.hydrological_modules.structures"
    mod = sys.modules.get(key) or ref_modules.ref_loader._load_module(key, ref_modules._R + "/hydrological_modules/structures.py")

    def ds_of(ldd):
        return ldd_ops.downstream_index(np.asarray(ldd, np.float64), land)

    def downstream(ldd, x):
        d = ds_of(ldd)
        x = np.asarray(x)
        return _Pcr(np.where(d >= 0, x[np.maximum(d, 0)], x))
    mod.downstream = downstream
    mod.boolean = lambda x: _Pcr(np.asarray(x) != 0)
    mod.cover = lambda x, y: _Pcr(np.asarray(x))
    mod.lddrepair = lambda ldd: _Pcr(ldd_ops.lddrepair_codes(np.asarray(ldd, np.float64), land))
    mod.ifthenelse = lambda c, a, b: _Pcr(np.where(np.asarray(c) != 0, a, b))
    mod.decompress = lambda a, *aa, **k: _Pcr(np.asarray(a))
    mod.compressArray = lambda m, *a, **k: np.asarray(m).copy()
    mod.LisSettings = ref_modules._FakeSettings
    var = _InitVar(land_mask, {}, loadmap, dict(options or {}), 86400.0, 3600.0)
    var.LddKinematic = _Pcr(np.array(ldd_kinematic, np.float64))
    var.IsStructureKinematic = np.asarray(is_structure) != 0
    mod.structures(var).initial()
    return {"LddStructuresKinematic": np.asarray(var.LddStructuresKinematic, np.float64),
            "IsUpsOfStructureKinematicC": np.asarray(var.IsUpsOfStructureKinematicC) != 0,
            "LddKinematic": np.asarray(var.LddKinematic, np.float64)}
