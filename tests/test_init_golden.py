"""Init-time parameter derivation against golden vectors made by the reference's OWN
soil.initial() / routing.initial() / routing.initialSecond() (tests/golden/make_golden.py init, oracle/ref_init.py).

The derivations are NumPy expressions in the same order as the reference, so the comparison is bit-exact for every
map both sides define (tolerance written here: 0).  The drainage-network maps come from the PCRaster operators, which
the golden run stubs with the same ldd_ops restatement (see oracle/ref_init.py): they check consistency, not PCRaster."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden

NETWORK = {"Ldd", "LddToChan", "LddKinematic", "UpArea", "InvUpArea", "Catchments", "InvCatchArea", "downstruct",
           "AtLastPointC", "IsChannel", "IsChannelKinematic", "IsStructureKinematic"}
# attributes the reference's initial() sets that are outside the hot path (reporting / other modules) and not mirrored
NOT_MIRRORED = {"AtLastPoint", "MaskMap", "IsChannelPcr", "avgdis", "Theta1a", "Theta1b", "Theta2", "TaInterception",
                "Interception", "LeafDrainage", "potential_transpiration", "Ta", "ESAct", "PrefFlow", "Infiltration",
                "SeepTopToSubA", "SeepTopToSubB", "SeepSubToGW", "Theta", "AvailableWaterForInfiltration", "RWS",
                "SoilMoistureStressDays"}


def _build(case, g=None):
    from lisflood_code_b200.Lisflood_initial import InitialVariables
    from lisflood_code_b200.global_modules.add1 import NumpyModified
    from lisflood_code_b200.hydrological_modules.groundwater import groundwater
    from lisflood_code_b200.hydrological_modules.routing import routing
    from lisflood_code_b200.hydrological_modules.soil import soil
    from lisflood_code_b200.hydrological_modules.surface_routing import surface_routing
    g = load_golden(case) if g is None else g
    raw = {k[5:]: (float(v) if v.ndim == 0 else v) for k, v in g.items() if k.startswith("raw__")}
    split = bool(g["SplitRouting"])
    var = InitialVariables(g["mask"], raw, {"SplitRouting": split, "drainedIrrigation": split}, DtSec=float(g["DtSec"]),
                           DtSecChannel=float(raw.get("DtSecChannel", 3600.0)))
    for k, v in g.items():
        if k.startswith("state__"):
            setattr(var, k[7:], NumpyModified(v.copy(), ["vegetation", "pixel"]) if v.ndim == 2 else v.copy())
    # the reference's order (Lisflood_initial.py:174-262): misc, ..., groundwater, soil, routing, surface routing
    var.misc_initial()
    soil(var).initial()
    r = routing(var)
    r.initial()
    groundwater(var).initial()
    surface_routing(var).initial()
    r.initialSecond()
    return g, var


@pytest.mark.parametrize("case", golden_cases("init_"))
def test_initial_matches_reference(case):
    compare_with_reference(*_build(case))


def compare_with_reference(g, var, at_least=100):
    checked, missing = 0, []
    for key, want in g.items():
        if not key.startswith(("soil__", "routing__", "surfgw__")):
            continue
        name = key.split("__", 1)[1]
        if name in NOT_MIRRORED:
            continue
        if not hasattr(var, name) or getattr(var, name) is None:
            missing.append(name)
            continue
        got = np.asarray(getattr(var, name))
        assert got.shape == want.shape, (name, got.shape, want.shape)
        if name == "downstruct":   # off-channel pixels are missing values in the reference's LddKinematic: undefined there
            on = np.asarray(var.LddKinematic) != 0
            got, want = got[on], want[on]
        if want.dtype.kind == "f":
            assert np.array_equal(got.astype(np.float64), want, equal_nan=True), (name, float(np.nanmax(np.abs(got - want))))
        else:
            assert np.array_equal(got, want), name
        checked += 1
    assert not missing, missing
    assert checked >= at_least
    return checked


@pytest.mark.parametrize("case", golden_cases("init_"))
def test_state_feeds_the_device_model_layout(case):
    """InitialVariables.state() after the modules' initial() carries every scalar, parameter, state and flag map
    HotPathModel uploads (the dictionary synthetic.full_stack builds by hand)."""
    from lisflood_code_b200 import hotpath
    g, var = _build(case)
    S = var.state()
    split = bool(g["SplitRouting"])
    need = list(hotpath.PARAMETERS) + list(hotpath.STATE) + ["IsChannel", "IsChannelKinematic", "AtLastPointC", "LddToChan",
                                                             "LddKinematic", "DtSec", "Beta", "PixelLength", "NoRoutSteps",
                                                             "CourantCrit", "AvWaterThreshold", "LeafDrainageK",
                                                             "DrainedFraction", "SMaxSealed", "mask"]
    if split:
        need += list(hotpath.SPLIT_PARAMETERS) + list(hotpath.SPLIT_STATE)
    assert not [k for k in need if k not in S], [k for k in need if k not in S]
    n = int(g["mask"].sum())
    for k in list(hotpath.PARAMETERS) + list(hotpath.STATE):
        a = np.asarray(S[k], np.float64)
        assert a.ndim == 0 or a.shape[-1] == n, k
        assert np.all(np.isfinite(a)), k
        if hotpath.PARAMETERS.get(k, hotpath.STATE.get(k)) == 3:
            assert a.shape == (3, n), k


@pytest.mark.parametrize("case", golden_cases("init_"))
def test_initialised_state_runs_through_the_oracle_model(case, oracle):
    """The state derived by the initial() mirrors is a complete model state: the CPU oracle steps it and conserves
    finiteness (no GPU needed: HotPathModel takes the same dictionary)."""
    from lisflood_code_b200 import synthetic
    from oracle import lisf_oracle_model as om
    g, var = _build(case)
    S = var.state()
    S["SplitRouting"] = bool(g["SplitRouting"])
    S["kgb"] = 0.75 * 0.72
    for k in ("W1a", "W1b", "W2"):          # soil-less pixels of the raw inputs divide by zero in the reference's satFun
        assert np.all(np.isfinite(S[k]))
    pore = np.asarray(S["PoreSpaceNotZero1a"]).all(axis=0)
    if not pore.all():
        pytest.skip("synthetic inputs with soil-less pixels: the reference's own kernel raises ZeroDivisionError there")
    O = om.OracleModel(S)
    O.step(synthetic.forcing(S, 0, 7))
    assert np.all(np.isfinite(O.var.ChanQAvg)) and np.all(np.isfinite(O.var.W1a))
