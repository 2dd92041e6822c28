"""End-state export, warm-start import and the products of the InitLisflood pre-run.

The reference checkpoints a run through its "end maps" (global_modules/default_options.py: the ReportedMap entries with
end=['repEndMaps']) and restarts from them through the `...InitValue` bindings its modules' initial() read (-9999 = cold
start): soil moisture travels as THETA, not as W (soil.py:234-277), overland storage as volume (surface_routing.py:49-66),
the channel as cross-section area + discharge (routing.py:203-218,326-327).  The pre-run (option InitLisflood, single
routing with one routing sub-step, routing.py:73-82) leaves `AvgDis` (-> QLimit of the split routing, routing.py:364) and
`LZAvInflowMap` (-> steady-state lower zone, groundwater.py:76-95).

Everything here is host code around a HotPathModel: `export_end_state` reads the device state and forms the end maps with
the reference's expressions, `init_bindings` turns them into the inputs of the init chain
(lisflood_code_b200/Lisflood_initial.py::initialise), `prerun_products` forms the two pre-run maps."""
from collections import OrderedDict

import numpy as np

# end map -> (init binding, model attribute, row)           default_options.py / the modules' input_files_keys
END_MAPS = OrderedDict([
    ("Theta1End", ("ThetaInit1Value", "Theta1a", 0)), ("Theta1ForestEnd", ("ThetaForestInit1Value", "Theta1a", 1)),
    ("Theta1IrrigationEnd", ("ThetaIrrigationInit1Value", "Theta1a", 2)),
    ("Theta2End", ("ThetaInit2Value", "Theta1b", 0)), ("Theta2ForestEnd", ("ThetaForestInit2Value", "Theta1b", 1)),
    ("Theta2IrrigationEnd", ("ThetaIrrigationInit2Value", "Theta1b", 2)),
    ("Theta3End", ("ThetaInit3Value", "Theta2", 0)), ("Theta3ForestEnd", ("ThetaForestInit3Value", "Theta2", 1)),
    ("Theta3IrrigationEnd", ("ThetaIrrigationInit3Value", "Theta2", 2)),
    ("UZEnd", ("UZInitValue", "UZ", 0)), ("UZForestEnd", ("UZForestInitValue", "UZ", 1)),
    ("UZIrrigationEnd", ("UZIrrigationInitValue", "UZ", 2)), ("LZEnd", ("LZInitValue", "LZ", None)),
    ("DSLREnd", ("DSLRInitValue", "DSLR", 0)), ("DSLRForestEnd", ("DSLRForestInitValue", "DSLR", 1)),
    ("DSLRIrrigationEnd", ("DSLRIrrigationInitValue", "DSLR", 2)),
    ("CumInterceptionEnd", ("CumIntInitValue", "CumInterception", 0)),
    ("CumInterceptionForestEnd", ("CumIntForestInitValue", "CumInterception", 1)),
    ("CumInterceptionIrrigationEnd", ("CumIntIrrigationInitValue", "CumInterception", 2)),
    ("CumIntSealedEnd", ("CumIntSealedInitValue", "CumInterSealed", None)),
    ("OFDirectEnd", ("OFDirectInitValue", "OFM3Direct", None)), ("OFOtherEnd", ("OFOtherInitValue", "OFM3Other", None)),
    ("OFForestEnd", ("OFForestInitValue", "OFM3Forest", None)),
    ("ChanCrossSectionEnd", ("TotalCrossSectionAreaInitValue", "TotalCrossSectionArea", None)),
    ("ChanQEnd", ("PrevDischarge", "ChanQ", None)),
    ("CrossSection2End", ("CrossSection2AreaInitValue", "CrossSection2Area", None)),
    ("ChSideEnd", ("PrevSideflowInitValue", "Sideflow1Chan", None)),
    ("SnowCoverAEnd", ("SnowCoverAInitValue", "SnowCoverS", 0)), ("SnowCoverBEnd", ("SnowCoverBInitValue", "SnowCoverS", 1)),
    ("SnowCoverCEnd", ("SnowCoverCInitValue", "SnowCoverS", 2)), ("FrostIndexEnd", ("FrostIndexInitValue", "FrostIndex", None)),
])
SPLIT_ONLY = ("CrossSection2End", "ChSideEnd")
FEEDER_ONLY = ("SnowCoverAEnd", "SnowCoverBEnd", "SnowCoverCEnd", "FrostIndexEnd")


def _theta(W, depth, pore):
    """thetaFun, soilloop.py:386-387."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(pore, W / depth, 0.0)


def export_end_state(M, S, feeders=False):
    """{end-map name: float64[N]} from the device-resident model M (HotPathModel) and the static maps S it was built from
    (SoilDepth*, PoreSpaceNotZero* or WS*, OFAlpha, PixelLength, Beta)."""
    split = bool(S.get("SplitRouting"))
    beta, plen = float(S["Beta"]), float(S["PixelLength"])
    W = {"Theta1a": M.get("W1a", 3), "Theta1b": M.get("W1b", 3), "Theta2": M.get("W2", 3)}
    lay = {"Theta1a": "1a", "Theta1b": "1b", "Theta2": "2"}
    cache = {}

    def attr(name):
        if name not in cache:
            if name in W:          # Theta = W / SoilDepth where there is pore space (soilloop.py:330-332)
                d = np.asarray(S["SoilDepth" + lay[name]])
                pore = np.asarray(S["PoreSpaceNotZero" + lay[name]]) if ("PoreSpaceNotZero" + lay[name]) in S else \
                    np.logical_and(d != 0, np.asarray(S["WS" + lay[name]]) != 0)
                cache[name] = _theta(W[name], d, pore)
            elif name.startswith("OFM3"):   # surface_routing.py:191-193
                k = {"OFM3Other": 0, "OFM3Forest": 1, "OFM3Direct": 2}[name]
                cache[name] = plen * np.asarray(S["OFAlpha"])[k] * M.get("OFQ" + name[4:]) ** beta
            else:
                rows = 3 if name in ("UZ", "DSLR", "CumInterception", "SnowCoverS") else 1
                cache[name] = M.get(name, rows)
        return cache[name]

    out = OrderedDict()
    for end, (binding, name, row) in END_MAPS.items():
        if (end in SPLIT_ONLY and not split) or (end in FEEDER_ONLY and not feeders):
            continue
        a = attr(name)
        out[end] = np.array(a[row] if row is not None else a, np.float64)
    return out


def init_bindings(end_state):
    """{init binding: map} for the modules' initial() (a warm start); bindings not covered keep their -9999 default."""
    return {END_MAPS[end][0]: v for end, v in end_state.items()}


def prerun_products(M, steps, DtDay):
    """The two maps an InitLisflood pre-run leaves behind: AvgDis = CumQ / TimeSinceStart (Lisflood_dynamic.py:224-227; the
    model must run with option "accumulate_discharge") and LZAvInflowMap = LZInflowCUM / DtDay / TimeSinceStart
    (groundwater.py:177)."""
    return {"AvgDis": M.get("CumQ") / steps, "LZAvInflowMap": (M.get("LZInflowCUM") * (1 / DtDay)) / steps}


def tss_line(values):
    """One line of a PCRaster .tss time series as the reference writes it: 6 significant digits of the float32-rounded
    values (the gauges are sampled from a REAL4 PCRaster map; global_modules/output.py::TssWriter)."""
    return " ".join("%.6g" % np.float32(v) for v in values)
