"""Shared driver of the pre-run / warm-start chain tests (reference tests/test_warmstart.py:51-141):
pre-run (InitLisflood) -> AvgDis + LZAvInflowMap -> long cold run; cold run of 1 step -> end maps -> warm start of 1 step
-> end maps -> ... ; every warm-started step must reproduce the long run."""
import numpy as np


def build_state(mask, raw, fractions, options, DtSec=86400.0):
    from lisflood_code_b200.Lisflood_initial import InitialVariables
    from lisflood_code_b200.global_modules.add1 import NumpyModified
    from lisflood_code_b200.hydrological_modules.groundwater import groundwater
    from lisflood_code_b200.hydrological_modules.routing import routing
    from lisflood_code_b200.hydrological_modules.soil import soil
    from lisflood_code_b200.hydrological_modules.surface_routing import surface_routing
    var = InitialVariables(mask, raw, options, DtSec=DtSec)
    for k, v in fractions.items():
        setattr(var, k, NumpyModified(v.copy(), ["vegetation", "pixel"]) if v.ndim == 2 else v.copy())
    var.misc_initial()
    soil(var).initial()
    r = routing(var)
    r.initial()
    groundwater(var).initial()
    surface_routing(var).initial()
    r.initialSecond()
    S = var.state()
    S["SplitRouting"] = bool(options.get("SplitRouting")) and not options.get("InitLisflood")
    S["kgb"] = 0.75 * 0.72
    for k in ("N", "rows", "cols", "NoRoutSteps"):
        S[k] = int(S[k])
    return S


def run_chain(make_model, nsteps=6, seed=33):
    """make_model(S) -> object with step(F), get(name, rows), set_option(name, value).  Returns (long-run discharge per
    step, warm-chain discharge per step, AvgDis, LZAvInflowMap)."""
    from lisflood_code_b200 import state_io, synthetic
    mask, raw, fractions = synthetic.raw_inputs(44, 50, seed=seed, soilless_fraction=0.0, channel_threshold=12)
    raw = dict(raw)
    for k in ("AvgDis", "LZAvInflowMap"):
        raw.pop(k)
    # ---- pre-run: InitLisflood (single routing, one routing sub-step; routing.py:73-82, groundwater.py:75-76)
    S0 = build_state(mask, raw, fractions, {"InitLisflood": True, "SplitRouting": True})
    assert S0["NoRoutSteps"] == 1 and not S0["SplitRouting"]
    P = make_model(S0)
    P.set_option("accumulate_discharge", 1)
    npre = 8
    for t in range(npre):
        P.step(synthetic.forcing(S0, t, seed))
    pre = state_io.prerun_products(P, npre, S0["DtDay"])
    assert np.all(np.isfinite(pre["AvgDis"])) and pre["AvgDis"].max() > 0
    raw.update(pre)
    # ---- long cold run with split routing
    opts = {"SplitRouting": True}
    S = build_state(mask, raw, fractions, opts)
    assert S["SplitRouting"] and S["NoRoutSteps"] == 24
    L = make_model(S)
    long_dis = []
    for t in range(nsteps):
        L.step(synthetic.forcing(S, 100 + t, seed))
        long_dis.append(L.get("ChanQAvg"))
    # ---- chain: cold start for one step, then warm start / stop step by step
    warm_dis = []
    M = make_model(S)
    Sk = S
    for t in range(nsteps):
        M.step(synthetic.forcing(S, 100 + t, seed))
        warm_dis.append(M.get("ChanQAvg"))
        end = state_io.export_end_state(M, Sk)
        rawk = dict(raw)
        rawk.update(state_io.init_bindings(end))
        Sk = build_state(mask, rawk, fractions, opts)
        # cumulative maps are not part of the end state (the reference restarts them at zero as well)
        M = make_model(Sk)
    return long_dis, warm_dis, pre["AvgDis"], pre["LZAvInflowMap"]


def check_chain(long_dis, warm_dis, tol=1e-9):
    from lisflood_code_b200.state_io import tss_line
    worst = 0.0
    for a, b in zip(warm_dis, long_dis):
        worst = max(worst, float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12))))
    assert worst < tol, worst
    # the reference's comparator for time series is exact on the 6 significant digits a .tss file holds
    gauges = np.argsort(-long_dis[-1])[:25]
    for a, b in zip(warm_dis, long_dis):
        assert tss_line(a[gauges]) == tss_line(b[gauges])
    return worst
