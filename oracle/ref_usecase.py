"""The inputs of the reference's own test catchment (tests/data/LF_ETRS89_UseCase) by binding name, read through the settings
XML the reference ships (bindings resolved by the product's settings parser, itself checked against the reference's parser)
and the harness readers of oracle/ref_maps.py.  TEST INFRASTRUCTURE ONLY (build container; needs /root/reference)."""
import os

import numpy as np

from . import ref_loader, ref_maps

ROOT = os.path.normpath(os.path.join(ref_loader._R, "..", "..", "tests", "data", "LF_ETRS89_UseCase"))
EXTRA_KEYS = ("PixelLengthUser", "PixelAreaUser", "GwLoss", "GwPercValue", "DtSec", "DtSecChannel", "ForestFraction",
              "DirectRunoffFraction", "WaterFraction", "IrrigationFraction", "RiceFraction", "OtherFraction",
              "LZAvInflowMap", "AvgDis", "QSplitMult")     # LZAvInflowMap, AvgDis: pre-run products the reference ships


def load_inputs(settings="base.xml", keys=()):
    """(land mask, {binding name: float or float64[N]}, bindings) for the init-time inputs of the hot-path modules."""
    from lisflood_code_b200.global_modules.settings import LisSettings
    from lisflood_code_b200.hydrological_modules import groundwater, routing, soil, surface_routing
    import contextlib
    import io
    before = LisSettings._instance
    with contextlib.redirect_stdout(io.StringIO()):       # (the lat / lon settings name a user variable they do not define)
        b = LisSettings(settings if os.path.isabs(settings) else os.path.join(ROOT, "settings", settings)).binding
    LisSettings._instance = before

    def path_of(v):       # loadmap: the name as given, then NetCDF / PCRaster by extension (add1.py:318-541)
        for c in (v, v + ".nc", v + ".map", os.path.splitext(v)[0] + ".nc"):
            if os.path.isfile(c):
                return c
        raise FileNotFoundError(v)
    mask = ref_maps.read_pcraster(path_of(b["MaskMap"])) == 1
    wanted = set(EXTRA_KEYS) | set(keys)
    for m in (soil.soil, routing.routing, groundwater.groundwater, surface_routing.surface_routing):
        for ks in m.input_files_keys.values():
            wanted |= set(ks)
    raw = {}
    for k in sorted(wanted):
        v = b[k]
        try:
            raw[k] = float(v)
            continue
        except ValueError:
            pass
        p = path_of(v)
        a = ref_maps.read_pcraster(p) if p.endswith(".map") else ref_maps.read_netcdf4_2d(p)
        raw[k] = a[mask].astype(np.float64)
        if not np.isfinite(raw[k]).all():
            raise ValueError("%s (%s) has missing values inside the mask" % (k, p))
    return mask, raw, b


def _stack(path):
    """(3-D array, time values, (seconds per unit, reference date)) of the only 3-D variable of a NetCDF-4 stack."""
    import datetime
    import re
    f = ref_maps.H5File(path if path.endswith(".nc") else path + ".nc")
    links = f.links()
    name = [k for k in links if len(f.dataset(links[k]).get("dims", ())) == 3][0]
    data = f.read(f.dataset(links[name]))
    tinfo = f.dataset(links["time"]) if "time" in links else None
    units = str(tinfo["attrs"].get("units", "")) if tinfo else ""
    if " since " not in units:                       # the 36 prescribed LAI maps: addressed by interval, not by date
        return data, None, None
    kind, stamp = units.split(" since ")
    m = re.match(r"(\d+)-(\d+)-(\d+)[ T](\d+):(\d+):(\d+)", stamp.strip())
    ref = datetime.datetime(*[int(x) for x in m.groups()])
    return data, f.read(tinfo), ({"hours": 3600.0, "days": 86400.0}[kind], ref)


class OracleRun(object):
    """The CPU restatement of the hot path (oracle/lisf_oracle_model.py + lisf_oracle_feeders.py) on the reference's test
    catchment: static state from the host init mirrors on the real inputs, forcing read by date from the shipped meteo
    stacks as add1.readnetcdf does (:700-720), LAI by calendar day (leafarea.py:82), CalendarDay as Lisflood_dynamic.py:46-48.
    Only the modules of the hot path run: no water use, rice, open-water evaporation, lakes or reservoirs."""

    def __init__(self, dt_sec=86400.0, split=False, settings="base.xml", init_lisflood=False):
        from lisflood_code_b200.Lisflood_initial import initialise
        from lisflood_code_b200.hydrological_modules.snow import feeder_arguments, frost, leafarea, snow
        from . import lisf_oracle_model as om
        from .lisf_oracle_feeders import FeederOracle
        keys = set(snow.input_files_keys["all"]) | set(frost.input_files_keys["all"]) | {"kdf", "PrScaling", "CalEvaporation"}
        self.mask, raw, self.binding = load_inputs(settings, keys)
        raw["DtSec"] = float(dt_sec)
        n = int(self.mask.sum())
        var = initialise(self.mask, raw, {"SplitRouting": split, "drainedIrrigation": split, "gridSizeUserDefined": True,
                                          "InitLisflood": bool(init_lisflood)},
                         DtSec=raw["DtSec"], DtSecChannel=raw["DtSecChannel"])
        snow(var).initial()
        frost(var).initial()
        leafarea(var).initial()
        S = var.state()
        S["mask"], S["SplitRouting"] = self.mask, bool(split)
        for k in ("N", "rows", "cols", "NoRoutSteps"):
            S[k] = int(S[k])
        P, state = feeder_arguments(var, np.full(n, 0.8))       # northern hemisphere: only the sign of the latitude is used
        self.kgb = P["kgb"]
        self.feeder = FeederOracle({k: (np.full(n, v) if np.ndim(v) == 0 else np.asarray(v, np.float64))
                                    for k, v in P.items() if k != "kgb"}, state, S["DtSec"])
        self.model = om.OracleModel(S)
        self.S = S
        b = self.binding
        self.forcing = {name: _stack(b[bind]) for name, bind in (("Precipitation", "PrecipitationMaps"), ("Tavg", "TavgMaps"),
                                                                  ("ET0", "ET0Maps"), ("E0", "E0Maps"))}
        self.lai = [_stack(b[bind])[0] for bind in ("LAIOtherMaps", "LAIForestMaps", "LAIIrrigationMaps")]

    def step(self, date):
        """One model step whose time stamp is `date` (datetime)."""
        from lisflood_code_b200.hydrological_modules.snow import lai_interval
        from .lisf_oracle_feeders import lai_term
        day = int(date.strftime("%j"))
        raw = {}
        for name, (data, tv, (unit_s, ref)) in self.forcing.items():
            idx = np.flatnonzero(tv == (date - ref).total_seconds() / unit_s)
            if idx.size != 1:
                raise KeyError("%s: no map stamped %s" % (name, date))
            raw[name] = data[idx[0]][self.mask]
        o = self.feeder.step(raw, day)
        j = lai_interval(day)
        lai = np.stack([self.lai[i][j][self.mask].astype(np.float64) for i in range(3)])
        self.model.step({"Rain": o["Rain"], "SnowMelt": o["SnowMelt"], "ETRef": o["ETRef"], "EWRef": o["EWRef"],
                         "ESRef": o["ESRef"], "isFrozenSoil": o["isFrozenSoil"], "LAI": lai, "LAITerm": lai_term(self.kgb, lai)})
        return self.model.var


def shipped_output(run, name):
    """Array of an output file the reference ships (tests/data/LF_ETRS89_UseCase/reference/<run>/<name>.nc): (time, y, x)
    for a stack, (y, x) for a map."""
    f = ref_maps.H5File(os.path.join(ROOT, "reference", run, name + ".nc"))
    return f.read(f.dataset(f.links()[name]))
