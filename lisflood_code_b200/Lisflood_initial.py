"""Host-side model object of the initialisation phase and the helper API the hot-path modules call on it
(reference: src/lisflood/Lisflood_initial.py:266-396, global_modules/add1.py:48-61).

The reference's `LisfloodModel_ini` builds ~40 modules; only the helpers the hot-path modules use during `initial()`
are kept here: `allocateDataArray`, `allocateVariableAllVegetation`, `defsoil`, `deffraction`, the
vegetation / land-use index helpers and a `loadmap` bound to a dictionary of inputs (float or float64[N]) falling
back to the settings binding.  `initial()` of soil / routing then derives the parameter maps exactly as the reference
does (tests/test_init_golden.py pins them to the reference's own `initial()` run on the same inputs), and
`HotPathModel(InitialVariables.state())` puts them on the device.
"""
from collections import OrderedDict

import numpy as np

from .global_modules.add1 import MaskInfo, NumpyModified

SOIL_USES = ["Rainfed", "Forest", "Irrigated"]                         # Lisflood_initial.py:108-112
PRESCRIBED_VEGETATION = [u + "_prescribed" for u in SOIL_USES]
RUNOFF = ["Other", "Forest", "Direct"]


class InitialVariables(object):
    def __init__(self, land_mask, maps, options=None, DtSec=86400.0, DtSecChannel=3600.0):
        self.maskinfo = MaskInfo(~np.asarray(land_mask, bool))
        self.maps = dict(maps)
        self.options = dict(options or {})
        self.num_pixel = self.maskinfo.num_pixels
        self.SOIL_USES = list(SOIL_USES)
        self.PRESCRIBED_VEGETATION = list(PRESCRIBED_VEGETATION)
        self.prescribed_vegetation = list(PRESCRIBED_VEGETATION)
        self.vegetation = list(PRESCRIBED_VEGETATION)
        self.VEGETATION_LANDUSE = OrderedDict(zip(self.PRESCRIBED_VEGETATION, self.SOIL_USES))
        self.LANDUSE_VEGETATION = OrderedDict((u, [v]) for v, u in self.VEGETATION_LANDUSE.items())
        self.dim_pixel = ("pixel", np.arange(self.num_pixel))
        self.dim_landuse = ("landuse", list(SOIL_USES))
        self.dim_runoff = ("runoff", list(RUNOFF))
        self.coord_landuse = OrderedDict([self.dim_landuse, self.dim_pixel])
        self.coord_vegetation = OrderedDict([("vegetation", self.vegetation[:]), self.dim_pixel])
        # miscInitial.py:44-60
        self.DtSec, self.DtSecChannel = float(DtSec), float(DtSecChannel)
        self.DtDay = self.DtSec / 86400.0
        self.InvDtSec, self.InvDtDay = 1 / self.DtSec, 1 / self.DtDay

    def misc_initial(self):
        """The numeric part of miscInitial.initial() the hot path needs (reference: hydrological_modules/miscInitial.py:
        52-133): grid size from the 'PixelLengthUser' / 'PixelAreaUser' inputs (option gridSizeUserDefined; scalars or
        maps), unit multipliers, groundwater percolation / loss per step."""
        zeros = self.maskinfo.in_zero
        self.PixelLength = self.loadmap('PixelLengthUser')
        area = self.loadmap('PixelAreaUser') if self._has('PixelAreaUser') else self.PixelLength ** 2
        self.PixelArea = zeros() + area if isinstance(area, float) else area
        self.InvPixelLength = 1.0 / self.PixelLength
        self.MMtoM, self.MtoMM = 0.001, 1000
        self.MMtoM3 = 0.001 * self.PixelArea
        self.M3toMM = 1 / self.MMtoM3
        loss = self.loadmap('GwLoss')
        self.GwLoss = zeros() + loss if isinstance(loss, float) else loss
        self.GwPerc = np.maximum(self.loadmap('GwPercValue'), self.GwLoss)
        self.GwPercStep = self.GwPerc * self.DtDay
        self.GwLossStep = self.GwLoss * self.DtDay
        if self._has('PrScaling'):                       # miscInitial.py:142-143 (read by the feeder kernel)
            self.PrScaling = self.loadmap('PrScaling')
        if self._has('CalEvaporation'):
            self.CalEvaporation = self.loadmap('CalEvaporation')

    def landuse_initial(self):
        """Land-use fractions (reference: hydrological_modules/landusechange.py:53-93, static maps): SoilFraction rows =
        Other (rainfed), Forest, Irrigation."""
        for nm in ("Forest", "DirectRunoff", "Water", "Irrigation", "Rice", "Other"):
            x = self.loadmap(nm + "Fraction")
            setattr(self, nm + "Fraction", self.maskinfo.in_zero() + x if isinstance(x, float) else x.copy())
        self.SoilFraction = self.allocateVariableAllVegetation()
        self.SoilFraction[0] = self.OtherFraction
        self.SoilFraction[1] = self.ForestFraction
        self.SoilFraction[2] = self.IrrigationFraction

    # ---- inputs ----
    def option(self, name):
        return bool(self.options.get(name, False))

    def loadmap(self, name):
        """float when the input is a number, else float64[N] (global_modules/add1.py:318-541 return convention)."""
        if name in self.maps:
            v = self.maps[name]
            if np.ndim(v) == 0:
                return float(v)
            v = np.asarray(v)
            if v.ndim == 2:
                from .global_modules.add1 import compressArray
                return compressArray(v, self.maskinfo).astype(np.float64)
            return np.ascontiguousarray(v, np.float64)
        from .global_modules.add1 import loadmap
        return loadmap(name, maskinfo=self.maskinfo)

    def loadtable(self, name):
        """A PCRaster lookup table (`TabTotStorage` ...) as a two-column array [site id, value]: from `maps` or, through
        the settings binding, from a .npy file or from the reference's own text tables ("id value" per line, .txt)."""
        if name in self.maps:
            return np.asarray(self.maps[name], np.float64).reshape(-1, 2)
        import os
        from .global_modules.settings import LisSettings
        path = LisSettings.instance().binding[name]
        for cand in (path, path + ".txt"):
            if os.path.isfile(cand) and not cand.endswith(".npy"):
                return np.loadtxt(cand, ndmin=2).astype(np.float64).reshape(-1, 2)
        return np.load(path).astype(np.float64).reshape(-1, 2)

    def _has(self, name):
        if name is None:
            return False
        if name in self.maps:
            return True
        try:
            from .global_modules.settings import LisSettings
            return name in LisSettings.instance().binding
        except RuntimeError:
            return False

    # ---- helper API (Lisflood_initial.py:266-396) ----
    def allocateDataArray(self, dimensions, dtype=float):
        coords = OrderedDict(dimensions)
        return NumpyModified(np.zeros([len(v) for v in coords.values()], dtype), coords.keys())

    def allocateVariableAllVegetation(self, dtype=float):
        return self.allocateDataArray([("vegetation", self.vegetation[:]), self.dim_pixel], dtype)

    def defsoil(self, name_1, name_2=None, name_3=None, coords=None):
        """(landuse|runoff, pixel) array from up to three inputs; a missing second / third input repeats the first
        (Lisflood_initial.py:371-391).  An input may be a binding name or a number."""
        if coords is None:
            coords = self.coord_landuse
        if list(coords.keys())[0] not in ("landuse", "runoff"):
            raise Exception("Coords key not found!")
        data = self.allocateDataArray(coords)

        def read(name, backup=None):
            if name is None:
                return backup
            if isinstance(name, str):
                return self.loadmap(name) if self._has(name) else backup
            return float(name)
        first = read(name_1)
        data.values[0][:] = first
        data.values[1][:] = read(name_2, first)
        data.values[2][:] = read(name_3, first)
        return data

    def deffraction(self, variable):
        ax = variable.dims.index("vegetation")
        return (self.SoilFraction.values * variable.values).sum(ax)

    def get_landuse_and_indexes_from_vegetation_epic(self, veg):
        landuse = self.VEGETATION_LANDUSE[veg]
        return self.vegetation.index(veg), self.SOIL_USES.index(landuse), landuse

    def get_indexes_from_landuse_and_veg_list_GLOBAL(self, landuse, veg_list):
        return ([self.vegetation.index(v) for v in veg_list], [self.PRESCRIBED_VEGETATION.index(v) for v in veg_list],
                self.SOIL_USES.index(landuse))

    # ---- hand-over to the device ----
    def state(self):
        """Every NumPy / scalar attribute set by the modules' initial(), as the dict HotPathModel takes."""
        out = {}
        for k, v in self.__dict__.items():
            if k in ("maps", "options", "maskinfo") or k.startswith("_"):
                continue
            if isinstance(v, (np.ndarray, float, int, bool, np.floating, np.integer)):
                out[k] = np.asarray(v) if isinstance(v, NumpyModified) else v
        out["mask"] = self.maskinfo.land_mask
        out["N"] = self.num_pixel
        out["rows"], out["cols"] = self.maskinfo.shape
        return out


def initialise(land_mask, maps=None, options=None, DtSec=86400.0, DtSecChannel=3600.0):
    """Runs the hot-path modules' initial() in the reference's order (Lisflood_initial.py:174-262: misc, land use,
    soil, routing, groundwater, surface routing, [reservoirs, lakes, structures with options simulateReservoirs /
    simulateLakes,] routing second part) on raw inputs by binding name and returns the InitialVariables object; `.state()`
    is what HotPathModel takes."""
    from .hydrological_modules.groundwater import groundwater
    from .hydrological_modules.routing import routing
    from .hydrological_modules.soil import soil
    from .hydrological_modules.surface_routing import surface_routing
    var = InitialVariables(land_mask, maps or {}, options, DtSec=DtSec, DtSecChannel=DtSecChannel)
    var.misc_initial()
    var.landuse_initial()
    soil(var).initial()
    r = routing(var)
    r.initial()
    groundwater(var).initial()
    surface_routing(var).initial()
    if var.option('simulateReservoirs') or var.option('simulateLakes'):
        from .hydrological_modules.lakes import lakes
        from .hydrological_modules.reservoir import reservoir
        from .hydrological_modules.structures import structures
        reservoir(var).initial()
        lakes(var).initial()
        structures(var).initial()
        var.simulateReservoirs, var.simulateLakes = var.option('simulateReservoirs'), var.option('simulateLakes')
    r.initialSecond()
    return var
