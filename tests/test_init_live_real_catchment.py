"""Init-time parameter derivation on the REAL inputs of the reference's own test catchment (tests/data/LF_ETRS89_UseCase:
65 maps + 39 scalars by binding name, through the settings XML the reference ships): the reference's OWN soil.initial(),
routing.initial() / initialSecond(), surface_routing.initial() and groundwater.initial() (oracle/ref_init.py) against the
host mirrors, bit for bit, single and split routing.  Only where /root/reference exists (the build container)."""
import numpy as np
import pytest

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


LAT_LON = "LF_lat_lon_UseCase/run_lat_lon.xml"      # the reference's second use case: a lat / lon grid, 14 x 31, PCRaster maps


@pytest.mark.parametrize("settings,split", [("base.xml", False), ("base.xml", True), (LAT_LON, False), (LAT_LON, True)])
def test_initial_on_the_reference_use_case(settings, split):
    import os
    from lisflood_code_b200.Lisflood_initial import InitialVariables
    from oracle import ref_init, ref_usecase
    from test_init_golden import _build, compare_with_reference
    if settings == LAT_LON:
        settings = os.path.join(os.path.dirname(ref_usecase.ROOT), *LAT_LON.split("/"))
    mask, raw, binding = ref_usecase.load_inputs(settings)
    n = int(mask.sum())
    assert n in (2847, 177) and sum(np.ndim(v) > 0 for v in raw.values()) >= 55
    dt_sec = raw["DtSec"]
    opts = {"SplitRouting": split, "drainedIrrigation": split, "gridSizeUserDefined": True}
    # what miscInitial / landusechange leave behind (mirrors pinned in tests/test_oracle_live_reference.py)
    v0 = InitialVariables(mask, raw, opts, DtSec=dt_sec, DtSecChannel=raw["DtSecChannel"])
    v0.misc_initial()
    v0.landuse_initial()
    state = {k: np.asarray(getattr(v0, k)) for k in ("SoilFraction", "RiceFraction", "WaterFraction", "OtherFraction",
                                                     "IrrigationFraction", "ForestFraction", "DirectRunoffFraction", "PixelArea")}
    g = {"mask": mask, "DtSec": np.float64(dt_sec), "SplitRouting": np.bool_(split)}
    g.update({"raw__" + k: np.asarray(v) for k, v in raw.items()})
    g.update({"state__" + k: v for k, v in state.items()})
    soil_state = {k: v for k, v in state.items() if k != "PixelArea"}
    g.update({"soil__" + k: v for k, v in ref_init.soil_initial(mask, raw, soil_state, opts, DtSec=dt_sec).items()})
    g.update({"routing__" + k: v for k, v in ref_init.routing_initial(mask, raw, {"PixelArea": state["PixelArea"]}, opts,
                                                                      DtSec=dt_sec, DtSecChannel=raw["DtSecChannel"]).items()})
    gwloss = np.zeros(n) + raw["GwLoss"]
    st2 = {"PixelLength": raw["PixelLengthUser"], "InvPixelLength": 1.0 / raw["PixelLengthUser"], "MMtoM": 0.001,
           "NManning": g["soil__NManning"], "Beta": g["routing__Beta"], "InvBeta": g["routing__InvBeta"],
           "AlpPow": g["routing__AlpPow"], "GwLoss": gwloss, "GwPerc": np.maximum(raw["GwPercValue"], gwloss)}
    g.update({"surfgw__" + k: v for k, v in ref_init.surface_and_groundwater_initial(mask, raw, st2, opts, DtSec=dt_sec).items()})
    checked = compare_with_reference(*_build(None, g))
    assert checked >= 140


def test_feeder_initial_on_the_reference_use_case():
    """snow.initial(), frost.initial(), leafarea.initial() of the live reference on the use case's real inputs against the
    feeder mirrors (hydrological_modules/snow.py), bit for bit; feeder_arguments() hands them to HotPathModel.set_feeder."""
    from lisflood_code_b200.Lisflood_initial import InitialVariables
    from lisflood_code_b200.hydrological_modules.snow import (FEEDER_PARAMETERS, feeder_arguments, frost, leafarea, snow)
    from oracle import ref_init, ref_usecase
    keys = set(snow.input_files_keys["all"]) | set(frost.input_files_keys["all"]) | {"kdf", "PrScaling", "CalEvaporation"}
    mask, raw, binding = ref_usecase.load_inputs("base.xml", keys)
    n = int(mask.sum())
    want = ref_init.feeders_initial(mask, raw)
    var = InitialVariables(mask, raw, {}, DtSec=raw["DtSec"], DtSecChannel=raw["DtSecChannel"])
    var.misc_initial()
    snow(var).initial()
    frost(var).initial()
    leafarea(var).initial()
    checked = 0
    for k, w in want.items():
        got = np.stack([np.zeros(n) + x for x in var.SnowCoverS]) if k == "SnowCoverS" else np.asarray(getattr(var, k))
        assert np.array_equal(np.broadcast_to(got, np.shape(w)), w), k
        checked += 1
    assert checked >= 18 and np.ndim(want["DeltaTSnow"]) == 1 and np.ndim(want["SnowMeltCoef"]) == 1      # real maps
    P, state = feeder_arguments(var, np.zeros(n))
    assert set(P) == set(FEEDER_PARAMETERS) and state["SnowCoverS"].shape == (3, n) and state["FrostIndex"].shape == (n,)


def test_structures_initial_on_the_reference_use_case():
    """reservoir.initial() / lakes.initial() of the live reference on the use case's real site maps and lookup tables
    (maps/ec_res.nc, ec_lakes.nc, tables/*.txt through the shipped settings) against the host mirrors, bit for bit."""
    from lisflood_code_b200.Lisflood_initial import InitialVariables, initialise
    from lisflood_code_b200.hydrological_modules.lakes import lakes
    from lisflood_code_b200.hydrological_modules.reservoir import reservoir
    from oracle import ref_init, ref_usecase
    from test_structures_init_golden import SKIP
    keys = set(reservoir.input_files_keys["simulateReservoirs"]) | set(lakes.input_files_keys["simulateLakes"])
    tab_keys = sorted(k for k in keys if k.startswith("Tab"))
    mask, raw, binding = ref_usecase.load_inputs("base.xml", keys - set(tab_keys))
    tables = {k: np.loadtxt(binding[k] if binding[k].endswith(".txt") else binding[k] + ".txt", ndmin=2) for k in tab_keys}
    assert len(tables) == 10 and all(t.shape[1] == 2 for t in tables.values())
    opts = {"simulateLakes": True, "simulateReservoirs": True, "gridSizeUserDefined": True}
    base = initialise(mask, raw, {"gridSizeUserDefined": True}, DtSec=raw["DtSec"], DtSecChannel=raw["DtSecChannel"])
    n = int(mask.sum())
    state = {"IsChannel": np.asarray(base.IsChannel), "IsStructureKinematic": np.zeros(n, bool),
             "LddKinematic": np.asarray(base.LddKinematic), "downstruct": np.asarray(base.downstruct),
             "ChanQ": np.asarray(base.ChanQ), "DtRouting": base.DtRouting}
    want = ref_init.structures_initial(mask, raw, tables, state, opts, DtSec=raw["DtSec"], DtSecChannel=raw["DtSecChannel"])
    maps = dict(raw)
    maps.update(tables)
    var = InitialVariables(mask, maps, opts, DtSec=raw["DtSec"], DtSecChannel=raw["DtSecChannel"])
    for k, v in state.items():
        setattr(var, k, v.copy() if np.ndim(v) else v)
    lakes(var).initial()
    reservoir(var).initial()
    checked = 0
    for name, w in want.items():
        if name in SKIP:
            continue
        assert hasattr(var, name), name
        got = np.asarray(getattr(var, name))
        assert got.shape == np.shape(w), (name, got.shape, np.shape(w))
        assert np.array_equal(got.astype(np.asarray(w).dtype), w, equal_nan=True), name
        checked += 1
    assert checked >= 38 and var.ReservoirIndex.size > 0 and var.LakeIndex.size > 0


def test_init_chain_with_structures_on_the_reference_use_case():
    """initialise() with simulateReservoirs / simulateLakes on the real inputs: the structures mirror (LDD cut just upstream
    of every structure, structures.py:43-61) against the live reference's structures.initial(), and the resulting state is
    what the device model and the CPU restatement take (a short run of the latter stays finite)."""
    from lisflood_code_b200.Lisflood_initial import initialise
    from lisflood_code_b200.hydrological_modules.lakes import lakes
    from lisflood_code_b200.hydrological_modules.reservoir import reservoir
    from oracle import lisf_oracle_model as om, ref_init, ref_usecase
    keys = set(reservoir.input_files_keys["simulateReservoirs"]) | set(lakes.input_files_keys["simulateLakes"])
    tab_keys = sorted(k for k in keys if k.startswith("Tab"))
    mask, raw, binding = ref_usecase.load_inputs("base.xml", keys - set(tab_keys))
    maps = dict(raw)
    maps.update({k: np.loadtxt(binding[k] if binding[k].endswith(".txt") else binding[k] + ".txt", ndmin=2) for k in tab_keys})
    opts = {"simulateLakes": True, "simulateReservoirs": True, "gridSizeUserDefined": True, "SplitRouting": True}
    plain = initialise(mask, maps, dict(opts, simulateLakes=False, simulateReservoirs=False), DtSec=raw["DtSec"],
                       DtSecChannel=raw["DtSecChannel"])
    var = initialise(mask, maps, opts, DtSec=raw["DtSec"], DtSecChannel=raw["DtSecChannel"])
    struct = np.asarray(var.IsStructureKinematic) != 0
    assert struct.sum() == var.ReservoirIndex.size + var.LakeIndex.size == 36
    want = ref_init.structures_module_initial(mask, plain.LddKinematic, struct)
    for k, w in want.items():
        assert np.array_equal(np.asarray(getattr(var, k)).astype(w.dtype), w), k
    assert np.array_equal(var.LddStructuresKinematic, plain.LddKinematic)
    cut = np.asarray(var.LddKinematic) != np.asarray(plain.LddKinematic)
    assert cut.sum() >= 36 and (np.asarray(var.LddKinematic)[cut] == 5).all()
    S = var.state()
    assert S["simulateReservoirs"] is True and S["simulateLakes"] is True
    for k in ("LddStructuresKinematic", "ReservoirIndex", "LakeIndex", "TotalReservoirStorageM3CC", "LakeAreaCC", "downstruct"):
        assert k in S, k
