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
    before = LisSettings._instance
    b = LisSettings(os.path.join(ROOT, "settings", settings)).binding
    LisSettings._instance = before

    def path_of(v):
        for c in (v, v + ".nc", v + ".map"):
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
