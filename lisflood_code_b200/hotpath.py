"""HotPathModel -- the device-resident model object of the hot path.

It plays the role of the reference's shared model object `self.var` for the modules on the hot path
(reference: src/lisflood/Lisflood_dynamic.py:114-229; attribute names: SURVEY.md §A.3): every map lives in
HBM inside one `lf_model` (C ABI, include/lisflood_b200.h) and is translated to / from the reference's
compressed float64[N] (or (3, N)) NumPy arrays only when Python reads or writes the attribute.  The HydroModule
mirrors in lisflood_code_b200/hydrological_modules/ drive it with the reference's call protocol.
"""
import ctypes as C

import numpy as np

from . import _capi
from .global_modules.add1 import NumpyModified

VEG_DIMS = ["vegetation", "pixel"]
LU_DIMS = ["landuse", "pixel"]

# maps the model needs before the first step (name -> rows)
PARAMETERS = {
    "b_Xinanjiang": 1, "PowerPrefFlow": 1, "UpperZoneK": 1, "GwPercStep": 1, "LowerZoneK": 1, "LZThreshold": 1,
    "GwLossStep": 1, "SoilFraction": 3, "DirectRunoffFraction": 1, "WaterFraction": 1, "MMtoM3": 1, "PixelArea": 1,
    "KSat1a": 3, "KSat1b": 3, "KSat2": 3, "GenuInvM1a": 3, "GenuInvM1b": 3, "GenuInvM2": 3, "WRes1a": 3, "WRes1b": 3,
    "WRes2": 3, "WS1a": 3, "WS1b": 3, "WS2": 3, "WWP1a": 3, "WWP1b": 3, "WWP2": 3, "WFC1a": 3, "WFC1b": 3, "WFC2": 3,
    "SoilDepth1a": 3, "SoilDepth1b": 3, "SoilDepth2": 3, "CropCoef": 3, "CropGroupNumber": 3, "OFAlpha": 3,
    "ChanLength": 1, "ChannelAlpha": 1,
}
SPLIT_PARAMETERS = {"ChannelAlpha2": 1, "QLimit": 1, "M3Limit": 1, "Chan2M3Start": 1, "Chan2QStart": 1}
STATE = {
    "CumInterception": 3, "W1a": 3, "W1b": 3, "W2": 3, "UZ": 3, "DSLR": 3, "LZ": 1, "CumInterSealed": 1, "LZInflowCUM": 1,
    "OFQOther": 1, "OFQForest": 1, "OFQDirect": 1, "ChanQKin": 1, "ChanM3Kin": 1, "ChanQ": 1,
}
SPLIT_STATE = {"Chan2QKin": 1, "Chan2M3Kin": 1, "CrossSection2Area": 1, "Sideflow1Chan": 1}
FORCING = {"Rain": 1, "SnowMelt": 1, "ETRef": 1, "EWRef": 1, "ESRef": 1, "LAI": 3, "LAITerm": 3}
FLAGS = ("isFrozenSoil", "IsChannel", "IsChannelKinematic", "AtLastPointC")
# structures in the routing sub-step loop: per-structure arrays, the reference's names
RESERVOIR_ARRAYS = ("TotalReservoirStorageM3CC", "ConservativeStorageLimitCC", "NormalStorageLimitCC",
                    "Normal_FloodStorageLimitCC", "FloodStorageLimitCC", "MinReservoirOutflowCC", "NormalReservoirOutflowCC",
                    "NonDamagingReservoirOutflowCC", "DeltaO", "DeltaLN", "DeltaNFL", "ReservoirStorageM3CC")
LAKE_ARRAYS = ("LakeAreaCC", "LakeFactor", "LakeFactorSqr", "LakeStorageM3CC", "LakeOutflowCC", "LakeInflowOldCC",
               "LakeStorageM3BalanceCC")
STRUCTURE_OUTPUTS = ("ReservoirFillCC", "QResOutM3DtCC", "LakeLevelCC", "QLakeOutM3DtCC")
# full maps the reference fills with np.put at the end of the sub-step loop (reservoir.py:300-322, lakes.py:280-297)
STRUCTURE_MAPS = {"ReservoirStorageM3": ("ReservoirStorageM3CC", "res"), "ReservoirFill": ("ReservoirFillCC", "res"),
                  "QResOutM3Dt": ("QResOutM3DtCC", "res"), "LakeStorageM3": ("LakeStorageM3CC", "lake"),
                  "LakeLevel": ("LakeLevelCC", "lake"), "LakeOutflow": ("LakeOutflowCC", "lake"),
                  "LakeInflowOld": ("LakeInflowOldCC", "lake"), "LakeStorageM3Balance": ("LakeStorageM3BalanceCC", "lake"),
                  "QLakeOutM3Dt": ("QLakeOutM3DtCC", "lake")}
DIAGNOSTIC_ONLY = ("WWP2", "WFC2", "SoilDepth1a", "SoilDepth1b", "SoilDepth2", "PixelArea")


THREE_ROWS = frozenset(k for d in (PARAMETERS, STATE, FORCING) for k, r in d.items() if r == 3) | {
    "SnowCoverS",
    "Interception", "TaInterception", "LeafDrainage", "potential_transpiration", "Ta", "ESAct", "PrefFlow",
    "Infiltration", "AvailableWaterForInfiltration", "SeepTopToSubA", "SeepTopToSubB", "SeepSubToGW", "Theta1a",
    "Theta1b", "Theta2", "Sat1a", "Sat1b", "Sat1", "Sat2", "UZOutflow", "GwPercUZLZ", "RWS", "Theta", "SurfaceRunSoil",
    "W1"}


class HotPathModel(object):
    def __init__(self, S, diagnostics=False, graphs=None, n_active=None, pick=None):
        """S: dict with the reference's attribute names (see synthetic.full_stack for the full list).
        graphs=(overland, channel): two device graphs of the same pixel subset (lf_graph_restrict: the local part of a
        raster cut over several GPUs, lisflood_code_b200/parallel.py) with n_active pixels; `pick` then maps a global
        map of S to its local pixels."""
        L = _capi.lib()
        if graphs is not None:
            mask = None
            rows_cols = (int(S["rows"]), int(S["cols"]))
        elif "mask_device" in S:
            mask = None
            rows_cols, n_active = (int(S["rows"]), int(S["cols"])), int(S["N"])
        else:
            mask = np.ascontiguousarray(S["mask"]).astype(np.uint8)
            rows_cols, n_active = mask.shape, int(mask.sum())
        cfg = _capi.ModelConfig()
        cfg.rows, cfg.cols = rows_cols
        plen = np.asarray(S["PixelLength"], np.float64)     # a map under option gridSizeUserDefined (miscInitial.py:52-68)
        if plen.ndim and plen.size and not (plen == plen.flat[0]).all():
            raise NotImplementedError("the device model takes ONE pixel length (overland routing, surface_routing.py:143-149): "
                                      "the PixelLength map is not uniform")
        cfg.DtSec, cfg.Beta, cfg.PixelLength = float(S["DtSec"]), float(S["Beta"]), float(plen.flat[0])
        cfg.NoRoutSteps, cfg.SplitRouting = int(S["NoRoutSteps"]), 1 if S.get("SplitRouting") else 0
        cfg.CourantCrit, cfg.AvWaterThreshold = float(S["CourantCrit"]), float(S["AvWaterThreshold"])
        cfg.LeafDrainageK, cfg.DrainedFraction = float(S["LeafDrainageK"]), float(S["DrainedFraction"])
        cfg.SMaxSealed = float(S["SMaxSealed"])
        cfg.diagnostics = 1 if diagnostics else 0
        self.__dict__["_h"] = C.c_void_p()
        self.__dict__["diagnostics"] = bool(diagnostics)
        self.__dict__["split"] = bool(S.get("SplitRouting"))
        self.__dict__["DtDay"] = float(S["DtSec"]) / 86400.0
        # option simulatePF: the parameters only the pF kernel needs (soil.py:184-190,466), kept on the host
        self.__dict__["_pf_params"] = {k: S[k] for k in ("HeadMax",) + tuple(p + x for p in ("GenuInvAlpha", "GenuInvN")
                                                                             for x in ("1a", "1b", "2")) if k in S}
        self.__dict__["N"] = n_active
        h = C.c_void_p()
        if graphs is not None:
            n_active = int(n_active)
            self.__dict__["N"] = n_active
            _capi.check(L.lf_model_create_from_graphs(C.byref(cfg), graphs[0], graphs[1], C.byref(h)))
        else:
            ldd_oc, ldd_kin = S["LddToChan"], S["LddKinematic"]
            if S.get("simulateReservoirs") or S.get("simulateLakes"):
                # the device graph is levelled on the network BEFORE structures.initial cuts it (structures.py:43-61); the
                # cut itself is applied by the channel kernel (lf_model_set_structures)
                ldd_kin = S["LddStructuresKinematic"]
            if not hasattr(ldd_oc, "data_ptr"):   # NumPy input; torch CUDA tensors are passed through as they are
                ldd_oc = np.ascontiguousarray(ldd_oc, np.float64)
                ldd_kin = np.ascontiguousarray(ldd_kin, np.float64)
            mask_arg = S["mask_device"] if "mask_device" in S else mask
            _capi.check(L.lf_model_create(C.byref(cfg), _capi.ptr(mask_arg), _capi.ptr(ldd_oc), _capi.ptr(ldd_kin),
                                          C.byref(h)))
        self.__dict__["_h"] = h
        self.__dict__["_rows"] = {}
        self.__dict__["NoRoutSteps"] = int(S["NoRoutSteps"])
        self.__dict__["_soil_calls"] = []
        self.__dict__["_nancheck"] = False
        self.__dict__["_nan_warned"] = False
        todo = dict(PARAMETERS)
        todo.update(STATE)
        if self.split:
            todo.update(SPLIT_PARAMETERS)
            todo.update(SPLIT_STATE)
        for name, rows in todo.items():
            if name in DIAGNOSTIC_ONLY and not diagnostics:
                continue
            if name in S:
                self.set(name, pick(S[name]) if pick else S[name], rows)
        for name in FLAGS:
            if name in S:
                self.set_flags(name, pick(S[name]) if pick else S[name])
        self.__dict__["_res_index"] = np.zeros(0, np.int64)
        self.__dict__["_lake_index"] = np.zeros(0, np.int64)
        if (S.get("simulateReservoirs") or S.get("simulateLakes")) and graphs is None:
            self.set_structures(S)     # on a cut raster the caller hands over this rank's structures (parallel.py)

    # ---- raw access --------------------------------------------------------------------------------
    def set(self, name, values, rows=None):
        n = self.N
        if hasattr(values, "data_ptr"):   # torch tensor (host or CUDA), float64, contiguous, compressed order
            assert values.is_contiguous() and values.element_size() == 8
            rows = rows or (3 if values.dim() == 2 else 1)
            assert values.numel() == rows * n, (name, values.shape)
            self._rows[name] = rows
            _capi.check(_capi.lib().lf_model_set(self._h, name.encode(), _capi.ptr(values), values.numel()))
            return
        a = np.asarray(values, np.float64)
        if rows is None:
            rows = 3 if (a.ndim == 2) else 1
        a = np.ascontiguousarray(np.broadcast_to(a, (rows, n) if rows > 1 else (n,)))
        self._rows[name] = rows
        _capi.check(_capi.lib().lf_model_set(self._h, name.encode(), _capi.ptr(a), a.size))

    def set_structures(self, S):
        """Reservoirs / lakes of the routing sub-step loop from the attributes reservoir.initial / lakes.initial leave on
        the model object (reference: reservoir.py:52-170, lakes.py:57-197)."""
        res = np.ascontiguousarray(S["ReservoirIndex"], np.int64) if S.get("simulateReservoirs") else np.zeros(0, np.int64)
        lak = np.ascontiguousarray(S["LakeIndex"], np.int64) if S.get("simulateLakes") else np.zeros(0, np.int64)
        self.__dict__["_res_index"], self.__dict__["_lake_index"] = res, lak
        _capi.check(_capi.lib().lf_model_set_structures(self._h, res.size, _capi.ptr(res) if res.size else None, lak.size,
                                                        _capi.ptr(lak) if lak.size else None))
        for names, count in ((RESERVOIR_ARRAYS, res.size), (LAKE_ARRAYS, lak.size)):
            if count:
                for name in names:
                    self.set_structure_array(name, S[name])

    def set_structure_array(self, name, values):
        a = np.ascontiguousarray(values, np.float64)
        _capi.check(_capi.lib().lf_model_structure_array(self._h, name.encode(), _capi.ptr(a), a.size, 1))

    def get_structure_array(self, name):
        n = self._lake_index.size if (name in LAKE_ARRAYS or name in ("LakeLevelCC", "QLakeOutM3DtCC")) else self._res_index.size
        a = np.zeros(n, np.float64)
        _capi.check(_capi.lib().lf_model_structure_array(self._h, name.encode(), _capi.ptr(a), a.size, 0))
        return a

    # ---- feeder modules (readmeteo scaling + snow + frost fused; leafarea) --------------------------------
    def set_feeder(self, P, state=None):
        """Parameters of the feeder kernel by the reference's names (hydrological_modules/snow.py FEEDER_PARAMETERS): maps
        or Python floats, like the float-or-array results of the reference's loadmap; state: SnowCoverS (3, N), FrostIndex."""
        for name, v in P.items():
            if np.ndim(v) == 0 and not hasattr(v, "data_ptr"):
                _capi.check(_capi.lib().lf_model_set_scalar(self._h, name.encode(), float(v)))
            else:
                self.set(name, v, 1)
        for name, v in (state or {}).items():
            self.set(name, v, 3 if name == "SnowCoverS" else 1)

    def feed(self, raw, calendar_day, asynchronous=False, packing=None, decode="float32"):
        """Raw meteo maps of the step -> forcing of the soil stage (lf_model_feed).  raw: Precipitation, Tavg, ET0, E0;
        float32 or float64; NumPy arrays or torch tensors (host or CUDA), compressed order.
        packing: {name: (scale_factor, add_offset)} for int16 maps still packed the CF way (the attributes of the NetCDF
        variable; netcdf.py:231-232): unpacked on the device (lf_model_feed_packed) in `decode` arithmetic."""
        from .hydrological_modules.snow import season_coefficients
        keys = ("Precipitation", "Tavg", "ET0", "E0")
        maps = [raw[k] for k in keys]
        size = lambda a: a.element_size() if hasattr(a, "element_size") else a.dtype.itemsize
        count = lambda a: a.numel() if hasattr(a, "numel") else a.size
        es = size(maps[0])
        sizes = (2,) if packing is not None else (4, 8)
        if es not in sizes or any(size(a) != es or count(a) != self.N for a in maps):
            raise ValueError("feed: four %s maps of %d pixels" % ("int16" if packing is not None else "float32 or four float64",
                                                                  self.N))
        if not hasattr(maps[0], "data_ptr"):
            maps = [np.ascontiguousarray(a) for a in maps]
            if asynchronous:
                self.__dict__["_raw_keepalive"] = maps     # borrowed until the next synchronising call
        c, ice_n, ice_s = season_coefficients(int(calendar_day))
        if packing is not None:
            if decode not in ("float32", "float64"):
                raise ValueError("feed: decode is 'float32' or 'float64'")
            for a in maps:
                dt = str(a.dtype)
                if not dt.endswith("int16") or dt.endswith("uint16"):
                    raise ValueError("feed: packed maps are int16, got %s" % dt)
            scale = np.array([float(packing[k][0]) for k in keys])
            offset = np.array([float(packing[k][1]) for k in keys])
            _capi.check(_capi.lib().lf_model_feed_packed(self._h, *[_capi.ptr(a) for a in maps], _capi.ptr(scale),
                                                         _capi.ptr(offset), 1 if decode == "float32" else 0, c, ice_n, ice_s,
                                                         1 if asynchronous else 0))
            return
        _capi.check(_capi.lib().lf_model_feed(self._h, *[_capi.ptr(a) for a in maps], 1 if es == 4 else 0, c, ice_n, ice_s,
                                              1 if asynchronous else 0))

    def set_lai(self, lai):
        """LAI maps (3, N) of the current 10-day interval; LAITerm = exp(-kgb LAI) is derived on the device."""
        if hasattr(lai, "data_ptr"):
            assert lai.is_contiguous() and lai.element_size() == 8 and lai.numel() == 3 * self.N
            a = lai
        else:
            a = np.ascontiguousarray(lai, np.float64)
        self._rows["LAI"], self._rows["LAITerm"] = 3, 3
        _capi.check(_capi.lib().lf_model_set_lai(self._h, _capi.ptr(a), 3 * self.N))

    def set_async(self, name, values):
        """Queue a new value for a map without waiting (see lf_model_set_async); `values`: contiguous float64
        NumPy array or torch tensor that stays untouched until the next get()."""
        size = values.numel() if hasattr(values, "numel") else values.size
        _capi.check(_capi.lib().lf_model_set_async(self._h, name.encode(), _capi.ptr(values), size))

    def get(self, name, rows=None):
        if name in RESERVOIR_ARRAYS or name in LAKE_ARRAYS or name in STRUCTURE_OUTPUTS:
            return self.get_structure_array(name)
        if name in STRUCTURE_MAPS:       # np.put(full map, index, per-structure values), reservoir.py:300-322
            cc, kind = STRUCTURE_MAPS[name]
            full = np.zeros(self.N)
            np.put(full, self._res_index if kind == "res" else self._lake_index, self.get_structure_array(cc))
            return full
        if name in ("LZOutflowToChannelPixel", "LZOutflowToChannel"):   # groundwater.py:142,180
            name = "LZOutflow"
        elif name == "M3all":                                           # surface_routing.py:196
            return self.get("OFM3Direct") + self.get("OFM3Other") + self.get("OFM3Forest")
        elif name == "WaterDepth":                                      # surface_routing.py:203
            return self.get("M3all") * (1 / self.get("MMtoM3"))
        elif name == "SoilMoistureStressDays":                          # option repStressDays, soilloop.py:597-598
            return np.where(self.get("RWS", 3) < 1, self.DtDay, 0.0)    # (RWS is a map of the diagnostics build)
        if rows is None:
            rows = self._rows.get(name, 1)
        return self.get_into(name, np.empty((rows, self.N) if rows > 1 else (self.N,), np.float64))

    def suction_pf(self, GenuInvAlpha=None, GenuInvN=None, HeadMax=None):
        """Option simulatePF (soilloop.py:673-705): pF0, pF1, pF2 -- log10 of the capillary head of the layers 1a, 1b, 2 --
        from the soil moisture of the model, by the device operator suctionUnsaturatedSoilPF.  GenuInvAlpha, GenuInvN:
        per layer ("1a", "1b", "2") the (landuse, pixel) maps soil.initial derives (soil.py:184-190), HeadMax (:466);
        by default the ones the model was created with.  The other parameters are the model's own."""
        from .hydrological_modules.soilloop import suctionUnsaturatedSoilPF
        lay = ("1a", "1b", "2")
        try:
            if GenuInvAlpha is None:
                GenuInvAlpha = {x: self._pf_params["GenuInvAlpha" + x] for x in lay}
            if GenuInvN is None:
                GenuInvN = {x: self._pf_params["GenuInvN" + x] for x in lay}
            if HeadMax is None:
                HeadMax = self._pf_params["HeadMax"]
        except KeyError as e:
            raise KeyError("suction_pf: %s was neither passed nor part of the model's parameters" % e.args[0])
        pf = [np.empty((3, self.N)) for _ in lay]
        maps = lambda k: [self.get(k + x, 3) for x in lay]
        pore = [self.get("WS" + x, 3) != 0 for x in lay]   # PoreSpaceNotZero, soil.py:226 (WS = ThetaS * depth: 0 with the depth)
        suctionUnsaturatedSoilPF(np.arange(3), pf[0], pf[1], pf[2], *maps("W"), *maps("WRes"), *maps("WS"), *pore,
                                 *[GenuInvAlpha[x] for x in lay], *maps("GenuInvM"), *[GenuInvN[x] for x in lay], HeadMax)
        return pf

    def get_into(self, name, out):
        """Copies a map into `out` (NumPy array or torch tensor, host or CUDA, float64, contiguous)."""
        size = out.numel() if hasattr(out, "numel") else out.size
        _capi.check(_capi.lib().lf_model_get(self._h, name.encode(), _capi.ptr(out), size))
        return out

    def get_async(self, name, out):
        """Queues the copy of a map into `out` (page-locked NumPy array / torch tensor) without waiting; valid after
        wait_outputs().  A float32 `out` gets the map narrowed on the device (OutputMapsDataType = float32,
        netcdf.py:478): half the bytes over the host link."""
        size = out.numel() if hasattr(out, "numel") else out.size
        dt = str(out.dtype)
        if dt.endswith("float64"):
            _capi.check(_capi.lib().lf_model_get_async(self._h, name.encode(), _capi.ptr(out), size))
        elif dt.endswith("float32"):
            _capi.check(_capi.lib().lf_model_get_async_f32(self._h, name.encode(), _capi.ptr(out), size))
        else:
            raise TypeError("get_async: `out` is float64 or float32, got %s" % dt)

    def wait_outputs(self):
        _capi.check(_capi.lib().lf_model_wait_outputs(self._h))

    def set_flags(self, name, values):
        if hasattr(values, "data_ptr"):
            assert values.element_size() == 1 and values.is_contiguous()
            _capi.check(_capi.lib().lf_model_set_flags(self._h, name.encode(), _capi.ptr(values), values.numel()))
            return
        a = np.ascontiguousarray(values).astype(np.uint8)
        _capi.check(_capi.lib().lf_model_set_flags(self._h, name.encode(), _capi.ptr(a), a.size))

    def set_forcing(self, F):
        """Meteorological input of the coming step (what readmeteo/snow/frost/leafarea leave on self.var)."""
        for name, rows in FORCING.items():
            if name in F:          # maps that did not change (the 10-day LAI maps, leafarea.py:76-90) may be left out
                self.set(name, F[name], rows)
        if "isFrozenSoil" in F:
            self.set_flags("isFrozenSoil", F["isFrozenSoil"])

    # ---- HydroModule call protocol (hydrological_modules/*.py mirrors) ----------------------------------
    _SOIL_SEQUENCE = ("dynamic_canopy", "dynamic_soil", "opensealed", "dynamic_perpixel", "groundwater")

    def _soil_stage_call(self, who):
        """The five soil-stage module calls of a step (Lisflood_dynamic.py:114-149) in the reference's order; the
        first one runs the fused device stage."""
        calls = self._soil_calls
        expected = self._SOIL_SEQUENCE[len(calls) % 5]
        if who != expected:
            raise RuntimeError("%s called out of order (expected %s): the hot-path modules follow "
                               "Lisflood_dynamic.py:114-149" % (who, expected))
        if len(calls) % 5 == 0:
            del calls[:]
            self.soil()
        calls.append(who)

    def _require_soil_stage_done(self):
        if len(self._soil_calls) != 5:
            raise RuntimeError("surface_routing.dynamic before the soil-stage modules of this step have all run")
        del self._soil_calls[:]

    # ---- stages ---------------------------------------------------------------------------------------
    def soil(self):
        _capi.check(_capi.lib().lf_model_soil(self._h))

    def surface_routing(self):
        _capi.check(_capi.lib().lf_model_surface_routing(self._h))

    def channel(self):
        _capi.check(_capi.lib().lf_model_channel(self._h))
        if self._nancheck:
            self.check_finite()

    def step(self, F=None):
        if F is not None:
            self.set_forcing(F)
        _capi.check(_capi.lib().lf_model_step(self._h))
        if self._nancheck:
            self.check_finite()

    def set_option(self, name, value):
        """Execution options of lf_model_set_option ("overlap_isolated", "early_blocks_per_sm", "isolated_blocks_per_sm",
        "narrow_runs", "cuda_graphs", "accumulate_discharge", "flagnancheck")."""
        _capi.check(_capi.lib().lf_model_set_option(self._h, name.encode(), float(value)))
        if name == "flagnancheck":
            self.__dict__["_nancheck"] = bool(value)

    def check_finite(self):
        """The reference's `-n` check (kinematic_wave_parallel.py:180-184) on the channel discharge: warns ONCE."""
        import warnings
        from .global_modules.errors import LisfloodWarning
        flag = C.c_int()
        _capi.check(_capi.lib().lf_model_nonfinite(self._h, C.byref(flag)))
        if flag.value and not self._nan_warned:
            self.__dict__["_nan_warned"] = True
            warnings.warn(LisfloodWarning("Non-finite discharge values found in the channel routing output"))
        return not flag.value

    def stage_times(self, reset=False):
        """Device milliseconds spent in (soil, overland, channel) stages of step() since the last reset."""
        a, b, c, k = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        _capi.check(_capi.lib().lf_model_stage_times(self._h, 1 if reset else 0, C.byref(a), C.byref(b), C.byref(c),
                                                     C.byref(k)))
        return {"soil_ms": a.value, "overland_ms": b.value, "channel_ms": c.value, "steps": k.value}

    def soil_stats(self, enable_timing=True):
        """Deferred-column counts per sub-step bucket and per-kernel device time of the last soil stage."""
        cnt = np.zeros(6, np.int64)
        ms = np.zeros(8, np.float64)
        _capi.check(_capi.lib().lf_model_soil_stats(self._h, 1 if enable_timing else 0, _capi.ptr(cnt), _capi.ptr(ms)))
        return {"deferred_columns": cnt.tolist(), "deferred_fraction": float(cnt.sum()) / (3.0 * self.N),
                "kernel_ms": {"first_pass": round(float(ms[0]), 3), "k_soil_veg_deferred": round(float(ms[1:7].sum()), 3),
                              "k_soil_pixel_flagged": round(float(ms[7]), 3)}}

    def info(self):
        v = [C.c_int64() for _ in range(5)]
        _capi.check(_capi.lib().lf_model_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("n_pixels", "levels_overland", "levels_channel", "isolated_channel_pixels", "device_bytes"),
                        [x.value for x in v]))

    # ---- attribute protocol of the reference's model object ----------------------------------------------
    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        rows = 3 if name in THREE_ROWS else 1
        try:
            a = self.get(name, rows)
        except _capi.LisfloodB200Error as e:
            if e.code == _capi.LF_ERR_INVALID:
                raise AttributeError(name)
            raise
        return NumpyModified(a, VEG_DIMS) if rows == 3 else a

    def __setattr__(self, name, value):
        if name.startswith("_"):
            self.__dict__[name] = value
            return
        if name in FLAGS:
            self.set_flags(name, value)
        else:
            self.set(name, value, 3 if name in THREE_ROWS else 1)

    def close(self):
        L = _capi._lib
        if L is not None and self.__dict__.get("_h"):
            L.lf_model_destroy(self._h)
            self.__dict__["_h"] = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
