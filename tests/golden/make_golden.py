"""Generates the committed golden vectors by running the UNMODIFIED reference kernels
(/root/reference, imported through oracle/ref_loader.py).  Run in the build container only:

    python tests/golden/make_golden.py

The .npz files next to this script are what the CPU tests (oracle pin) and the GPU parity tests
compare against; /root/reference is never needed at test time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from lisflood_code_b200 import synthetic  # noqa: E402
from oracle import ref_loader  # noqa: E402


def routing_case(kwp, name, rows, cols, seed, noise, mask_fraction, dx_array, steps, split, beta=0.6, dt=3600.0,
                 negative_q=False, ldd=None, mask=None, alpha=None, q0=None, q=None, dx=None):
    if ldd is None:
        ldd, mask = synthetic.random_ldd(rows, cols, seed=seed, noise=noise, mask_fraction=mask_fraction)
    n = int(mask.sum())
    a_, q0_, q_ = synthetic.routing_fields(n, seed)
    alpha = a_ if alpha is None else alpha
    q0 = q0_ if q0 is None else q0
    q = q_ if q is None else q
    rng = np.random.default_rng(seed + 1)
    if dx is None:
        dx = rng.uniform(3000.0, 7000.0, n) if dx_array else 5000.0
    if negative_q:  # negative side-flow (water abstraction) exercises the C <= 1e-12 branch
        q = q - 1.5e-4 * (rng.random(n) < 0.3)
        q0 = q0 * (rng.random(n) < 0.7)
    alpha2 = alpha * rng.uniform(1.5, 3.0, n) if split else None
    kw = kwp.kinematicWave(ldd[mask].copy(), mask, alpha, beta, dx, dt, alpha_floodplains=alpha2)
    out = dict(ldd=ldd[mask].astype(np.float64), mask=mask, alpha=alpha, beta=np.float64(beta), dx=np.asarray(dx),
               dt=np.float64(dt), q0=q0, q=q,
               downstream_lookup=kw.downstream_lookup, upstream_lookup=kw.upstream_lookup,
               num_upstream_pixels=kw.num_upstream_pixels, pixels_ordered=kw.pixels_ordered,
               order_start_stop=kw.order_start_stop)
    Q = q0.copy()
    snaps = []
    for s in range(steps):
        kw.kinematicWaveRouting(Q, q, "main_channel")
        snaps.append(Q.copy())
    out["Q_main"] = np.stack(snaps)
    if split:
        out["alpha2"] = alpha2
        Q2 = q0.copy() * 0.5
        snaps = []
        for s in range(steps):
            kw.kinematicWaveRouting(Q2, q * 0.3, "floodplains")
            snaps.append(Q2.copy())
        out["Q_fp"] = np.stack(snaps)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "n=%d levels=%d K=%d" % (n, kw.order_start_stop.shape[0], kw.upstream_lookup.shape[1]))


MODEL_KEYS_V = ["CumInterception", "W1a", "W1b", "W1", "W2", "UZ", "DSLR", "Interception", "TaInterception", "LeafDrainage",
                "potential_transpiration", "Ta", "ESAct", "PrefFlow", "Infiltration", "AvailableWaterForInfiltration",
                "SeepTopToSubA", "SeepTopToSubB", "SeepSubToGW", "Theta1a", "Theta1b", "Theta2", "Sat1a", "Sat1b", "Sat1",
                "Sat2", "UZOutflow", "GwPercUZLZ", "RWS", "Theta", "SurfaceRunSoil"]
MODEL_KEYS_N = ["CumInterSealed", "LZ", "LZInflowCUM", "DirectRunoff", "TASealed", "EWaterAct", "InterSealed", "RainSnowmelt",
                "TaInterceptionAll", "TaPixel", "ESActPixel", "PrefFlowPixel", "InfiltrationPixel", "ThetaAll",
                "SeepTopToSubPixelA", "SeepTopToSubPixelB", "SeepSubToGWPixel", "Theta1aPixel", "Theta1bPixel", "Theta2Pixel",
                "UZOutflowPixel", "GwPercUZLZPixel", "GwLossLZ", "LZOutflow", "LZOutflowToChannelPixel", "LZAvInflow",
                "SurfaceRunoff", "TotalRunoff", "OFQDirect", "OFQOther", "OFQForest", "OFM3Direct", "OFM3Other", "OFM3Forest",
                "Qall", "M3all", "OFToChanM3", "WaterDepth", "ToChanM3Runoff", "ToChanM3RunoffDt", "ChanQKin", "ChanM3Kin",
                "ChanQ", "sumDisDay", "FlowVelocity", "TravelDistance", "ChanM3", "TotalCrossSectionArea", "ChanQAvg",
                "sumDis", "DischargeM3Out", "TaCUM", "TaInterceptionCUM", "ESActCUM", "GwLossCUM"]
MODEL_KEYS_SPLIT = ["Chan2QKin", "Chan2M3Kin", "CrossSection2Area", "Sideflow1Chan", "sumDisDay_NOTlast"]


def model_case(name, rows, cols, seed, split, steps, **kw):
    """Full hot-path step executed by the reference's OWN module classes (oracle/ref_modules.py)."""
    from oracle import ref_modules
    S = synthetic.full_stack(rows, cols, seed=seed, split_routing=split, **kw)
    M = ref_modules.RefModel(S)
    out = {}
    for k, v in S.items():
        out["S__" + k] = np.asarray(v)
    keys = MODEL_KEYS_V + MODEL_KEYS_N + (MODEL_KEYS_SPLIT if split else [])
    for t in range(steps):
        F = synthetic.forcing(S, t, seed)
        for k, v in F.items():
            out["F%d__%s" % (t, k)] = v
        M.step(F)
        for k in keys:
            out["O%d__%s" % (t, k)] = np.asarray(getattr(M.var, k)).copy()
    out["steps"] = np.int64(steps)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "n=%d steps=%d split=%s channel fraction %.3f" % (S["N"], steps, split, S["IsChannel"].mean()))


STRUCTURE_KEYS = ["ReservoirStorageM3", "ReservoirFill", "QResOutM3Dt", "LakeStorageM3", "LakeLevel", "LakeOutflow",
                  "LakeInflowOld", "LakeStorageM3Balance", "QLakeOutM3Dt"]
STRUCTURE_KEYS_CC = ["ReservoirStorageM3CC", "ReservoirFillCC", "LakeStorageM3CC", "LakeOutflowCC", "LakeLevelCC"]


def structures_case(name, rows, cols, seed, split, steps, n_res, n_lakes, **kw):
    """Model steps with reservoirs and lakes in the routing sub-step loop, executed by the
    reference's OWN classes: routing.dynamic -> lakes.dynamic_inloop / reservoir.dynamic_inloop (oracle/ref_modules.py).
    Named structures_* (not model_*): the device path does not take structures yet, the CPU oracle does."""
    from oracle import ref_modules
    S = synthetic.full_stack(rows, cols, seed=seed, split_routing=split, **kw)
    synthetic.add_structures(S, n_res, n_lakes, seed=seed)
    M = ref_modules.RefModel(S, options={"simulateLakes": n_lakes > 0, "simulateReservoirs": n_res > 0})
    out = {"S__" + k: np.asarray(v) for k, v in S.items()}
    keys = MODEL_KEYS_V + MODEL_KEYS_N + (MODEL_KEYS_SPLIT if split else []) + STRUCTURE_KEYS + STRUCTURE_KEYS_CC
    for t in range(steps):
        F = synthetic.forcing(S, t, seed)
        for k, v in F.items():
            out["F%d__%s" % (t, k)] = v
        M.step(F)
        for k in keys:
            out["O%d__%s" % (t, k)] = np.asarray(getattr(M.var, k)).copy()
    out["steps"] = np.int64(steps)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "n=%d steps=%d reservoirs=%s lakes=%s fill %s" % (S["N"], steps, S["ReservoirIndex"], S["LakeIndex"],
                                                                  np.round(M.var.ReservoirFillCC, 3)))


def structures_init_case(name, rows, cols, seed, warm):
    """reservoir.initial() and lakes.initial() executed by the reference's OWN classes (oracle/ref_init.py) on site maps
    and lookup tables; cold start (initial values -9999) or warm start (state maps)."""
    from oracle import ref_init
    S = synthetic.full_stack(rows, cols, seed=seed, split_routing=False, channel_threshold=10)
    ldd0 = S["LddKinematic"].copy()
    synthetic.add_structures(S, 3, 3, seed=seed)
    n, mask = S["N"], S["mask"]
    rng = np.random.default_rng(seed)
    res_sites, lake_sites = np.zeros(n), np.zeros(n)
    res_sites[S["ReservoirIndex"]] = [11, 12, 13]
    lake_sites[S["LakeIndex"]] = [5, 6, 7]
    off = np.flatnonzero(~S["IsChannel"])[:2]
    res_sites[off[0]], lake_sites[off[1]] = 14, 8          # sites off the channel network are dropped
    ids_r, ids_l = np.array([11., 12., 13., 14.]), np.array([5., 6., 7., 8.])
    T = lambda ids, lo, hi: np.stack([ids, rng.uniform(lo, hi, ids.size)], 1)
    tables = {"TabTotStorage": T(ids_r, 1e6, 5e7), "TabConservativeStorageLimit": T(ids_r, .05, .15),
              "TabNormalStorageLimit": T(ids_r, .4, .6), "TabFloodStorageLimit": T(ids_r, .85, .97),
              "TabNonDamagingOutflowQ": T(ids_r, 20., 60.), "TabNormalOutflowQ": T(ids_r, 2., 70.),
              "TabMinOutflowQ": T(ids_r, .5, 3.), "TabLakeArea": T(ids_l, 2e6, 5e7), "TabLakeA": T(ids_l, 5., 60.),
              "TabLakeAvNetInflowEstimate": T(ids_l, .5, 20.)}
    U = lambda lo, hi: rng.uniform(lo, hi, n)
    raw = {"ReservoirSites": res_sites, "LakeSites": lake_sites, "adjust_Normal_Flood": 0.5, "ReservoirRnormqMult": U(.8, 1.2),
           "LakeMultiplier": 1.1, "PrevDischarge": U(.1, 9.),
           "ReservoirInitialFillValue": U(.2, .9) if warm else -9999.0, "LakeInitialLevelValue": U(.2, 3.) if warm else -9999.0,
           "LakePrevInflowValue": U(.1, 9.) if warm else -9999.0, "LakePrevOutflowValue": U(.1, 9.) if warm else -9999.0}
    state = {"IsChannel": S["IsChannel"], "IsStructureKinematic": np.zeros(n, bool), "LddKinematic": ldd0,
             "downstruct": S["downstruct"], "ChanQ": S["ChanQ"], "DtRouting": S["DtRouting"]}
    opts = {"simulateLakes": True, "simulateReservoirs": True}
    out = {"mask": mask, "DtSec": np.float64(S["DtSec"])}
    out.update({"raw__" + k: np.asarray(v) for k, v in raw.items()})
    out.update({"table__" + k: v for k, v in tables.items()})
    out.update({"state__" + k: np.asarray(v) for k, v in state.items()})
    out.update({"out__" + k: v for k, v in ref_init.structures_initial(mask, raw, tables, state, opts, DtSec=S["DtSec"]).items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "n=%d outputs=%d" % (n, sum(k.startswith("out__") for k in out)))


def init_case(name, rows, cols, seed, split, scalar_maps=False, dt_sec=86400.0, soilless_fraction=0.03):
    """soil.initial() and routing.initial()/initialSecond() executed by the reference's OWN classes (oracle/ref_init.py)
    on raw inputs by binding name; stores inputs and every attribute they set."""
    from oracle import ref_init
    mask, raw, state = synthetic.raw_inputs(rows, cols, seed=seed, scalar_maps=scalar_maps, soilless_fraction=soilless_fraction)
    opts = {"SplitRouting": split, "drainedIrrigation": split}
    out = {"mask": mask, "DtSec": np.float64(dt_sec), "SplitRouting": np.bool_(split)}
    out.update({"raw__" + k: np.asarray(v) for k, v in raw.items()})
    out.update({"state__" + k: np.asarray(v) for k, v in state.items()})
    soil_state = {k: v for k, v in state.items() if k != "PixelArea"}
    out.update({"soil__" + k: v for k, v in ref_init.soil_initial(mask, raw, soil_state, opts, DtSec=dt_sec).items()})
    out.update({"routing__" + k: v for k, v in ref_init.routing_initial(mask, raw, {"PixelArea": state["PixelArea"]}, opts,
                                                                        DtSec=dt_sec).items()})
    # surface_routing.initial / groundwater.initial need what miscInitial, soil.initial and routing.initial left behind
    n = int(mask.sum())
    gwloss = np.zeros(n) + raw["GwLoss"]
    st2 = {"PixelLength": raw["PixelLengthUser"], "InvPixelLength": 1.0 / raw["PixelLengthUser"], "MMtoM": 0.001,
           "NManning": out["soil__NManning"], "Beta": out["routing__Beta"], "InvBeta": out["routing__InvBeta"],
           "AlpPow": out["routing__AlpPow"], "GwLoss": gwloss, "GwPerc": np.maximum(raw["GwPercValue"], gwloss)}
    out.update({"surfgw__" + k: v for k, v in ref_init.surface_and_groundwater_initial(mask, raw, st2, opts, DtSec=dt_sec).items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "surface/groundwater maps=%d" % sum(k.startswith("surfgw__") for k in out))
    print(name, "n=%d soil maps=%d routing maps=%d" % (int(mask.sum()), sum(k.startswith("soil__") for k in out),
                                                      sum(k.startswith("routing__") for k in out)))


def adversarial_case(kwp, name, beta):
    """Inputs chosen against the solver's stopping rule (kinematic_wave_parallel_tools.py:73-80): discharges up to 1e9
    (the 1e-12 absolute tolerance is far below one ulp there, so the reference leaves its loop through `Q == previous`
    or wanders between two neighbouring values for 3000 iterations), down to 1e-13 (the tolerance is met by the initial
    guess), alpha / dx / dt over six decades, negative and zero side flow, long chains and wide confluences."""
    rng = np.random.default_rng(977)
    rows, cols = 24, 160
    ldd = np.full((rows, cols), 6.0)            # every row is a chain to the east ...
    ldd[:, -1] = 2.0                            # ... collected by the last column
    ldd[-1, -1] = 5.0
    ldd[::3, ::7] = 5.0                         # some pits: short chains, isolated pixels
    mask = np.ones((rows, cols), bool)
    n = rows * cols
    logu = lambda lo, hi: np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    alpha = logu(1e-3, 1e3)
    dx = logu(10.0, 1e5)
    q0 = logu(1e-13, 1e9)
    q0[rng.random(n) < 0.1] = 0.0
    q = np.where(rng.random(n) < 0.5, logu(1e-12, 1e2), -logu(1e-9, 1e-2))
    q[rng.random(n) < 0.1] = 0.0
    routing_case(kwp, name, rows, cols, 977, 0, 0, True, 5, False, beta=beta, dt=600.0, ldd=ldd, mask=mask, alpha=alpha, q0=q0,
                 q=q, dx=dx)


def feeders_case(name, n, seed, steps, dt_sec, day0):
    """Feeder modules of a step: snow.dynamic() and frost.dynamic() executed by the reference's OWN
    classes (hydrological_modules/snow.py:95-187, frost.py:61-78) on raw meteo maps scaled as readmeteo.dynamic does
    (readmeteo.py:61-81, four NumPy expressions restated here: the class needs xarray); both hemispheres, calendar days
    that cross the summer ice-melt season limits, float32 forcing as read from NetCDF (widened to float64)."""
    from types import SimpleNamespace
    from oracle import ref_modules
    M = ref_modules.load_feeders()
    rng = np.random.default_rng(seed)
    U = lambda lo, hi: rng.uniform(lo, hi, n)
    P = {"PrScaling": U(0.8, 1.2), "CalEvaporation": U(0.8, 1.3), "DeltaTSnow": 0.9674 * U(0, 400.0) * 0.0065,
         "SnowSeason": U(0.5, 2.0) * 0.5, "TempSnow": U(0.5, 1.5), "SnowFactor": U(0.9, 1.3), "SnowMeltCoef": U(2.0, 5.0),
         "TempMelt": U(-0.5, 0.5), "lat_rad": np.radians(U(-60.0, 70.0)), "Kfrost": U(0.4, 0.7), "Afrost": U(0.95, 0.99),
         "FrostIndexThreshold": U(40.0, 60.0), "SnowWaterEquivalent": U(0.05, 0.45)}
    state0 = {"SnowCoverS": np.stack([U(0, 80.0) * (rng.random(n) < 0.6) for _ in range(3)]), "FrostIndex": U(0, 70.0) * (rng.random(n) < 0.7)}
    v = SimpleNamespace(**{k: a.copy() for k, a in P.items()})
    ref_modules._FakeMaskInfo.set(np.ones((1, n), bool))
    v.DtDay = dt_sec / 86400.0
    v.SnowCoverS = [state0["SnowCoverS"][i].copy() for i in range(3)]
    v.FrostIndex = state0["FrostIndex"].copy()
    v.TotalPrecipitation = np.zeros(n)
    sn, fr = M["snow"](v), M["frost"](v)
    v.SnowDayDegrees = 360 / 365.25                       # snow.initial, :66-75
    sn.icemelt_start_N, sn.icemelt_end_N, sn.icemelt_start_S, sn.icemelt_end_S = 165, 257, 347, 74
    v.IceDayDegrees = 2 * v.SnowDayDegrees
    out = {"n": np.int64(n), "steps": np.int64(steps), "DtSec": np.float64(dt_sec)}
    out.update({"P__" + k: a for k, a in P.items()})
    out.update({"S__" + k: a for k, a in state0.items()})
    for t in range(steps):
        day = (day0 + int(t * v.DtDay * 37)) % 366 + 1    # strides through the year: both ice-melt seasons and winter
        raw = {"Precipitation": (rng.gamma(0.8, 8.0, n) * (rng.random(n) < 0.5)).astype(np.float32),
               "Tavg": rng.uniform(-25.0, 25.0, n).astype(np.float32), "ET0": rng.uniform(0, 6.0, n).astype(np.float32),
               "E0": rng.uniform(0, 7.0, n).astype(np.float32)}
        out["CalendarDay%d" % t] = np.int64(day)
        for k, a in raw.items():
            out["R%d__%s" % (t, k)] = a
        v.CalendarDay = day
        v.Precipitation = raw["Precipitation"].astype(np.float64) * v.DtDay * v.PrScaling   # readmeteo.py:66-69
        v.Tavg = raw["Tavg"].astype(np.float64)
        v.ETRef = raw["ET0"].astype(np.float64) * v.DtDay * v.CalEvaporation
        v.EWRef = raw["E0"].astype(np.float64) * v.DtDay * v.CalEvaporation
        v.ESRef = (v.EWRef + v.ETRef) / 2                                                    # :78
        sn.dynamic()
        fr.dynamic()
        for k in ("Precipitation", "ETRef", "EWRef", "ESRef", "Rain", "Snow", "SnowMelt", "SnowCover", "FrostIndex",
                  "isFrozenSoil", "TotalPrecipitation"):
            out["O%d__%s" % (t, k)] = np.asarray(getattr(v, k)).copy()
        out["O%d__SnowCoverS" % t] = np.stack(v.SnowCoverS)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "n=%d steps=%d frozen %.2f snow-covered %.2f" % (n, steps, out["O%d__isFrozenSoil" % (steps - 1)].mean(),
                                                                 (out["O%d__SnowCover" % (steps - 1)] > 0).mean()))


def real_catchment_case(kwp, name, steps=6):
    """The reference's OWN test catchment (tests/data/LF_ETRS89_UseCase/maps: 57 x 80 cells of 5 km, 2847 in the mask, one
    outlet), read with oracle/ref_maps.py:
      * ldd, mask, pixarea and the PCRaster-made upstream-area map ec_upArea -- the pin of global_modules/ldd_ops.py
        (downstream_index / lddrepair at the outlet, whose link leaves the mask; accuflux);
      * the reference's kinematicWave on that network (codes after lddmask / lddrepair, as routing.initial hands them over,
        routing.py:90-101), channel alpha from the catchment's own geometry maps with the formulas of routing.py:184-235,
        space_delta = its chanlength map: graph arrays and discharge snapshots."""
    from oracle import ref_maps
    from lisflood_code_b200.global_modules import ldd_ops
    M = os.path.join(ref_loader._R, "..", "..", "tests", "data", "LF_ETRS89_UseCase", "maps")
    rd = lambda f: ref_maps.read_netcdf4_2d(os.path.join(M, f + ".nc"))
    mask = ref_maps.read_pcraster(os.path.join(M, "mask.map")) == 1
    ldd_raw, up_area, pixarea = rd("ec_ldd"), rd("ec_upArea"), rd("pixarea")
    codes = ldd_ops.lddrepair_codes(ldd_raw[mask].astype(np.float64), mask)
    C = lambda f: rd(f)[mask].astype(np.float64)
    beta = 0.6
    grad = np.maximum(C("changrad"), 0.0001)                                   # routing.py:184 (ChanGradMin of the use case)
    man, bw, depth, sdxdy, length = C("ec_chanman"), C("ec_chanbw"), C("ec_chanbnkf"), C("chans"), C("chanlength")
    wetted = bw + 2 * np.sqrt(depth ** 2 + (depth * sdxdy) ** 2)                # :229-230
    alpha = ((man / np.sqrt(grad)) ** beta * wetted ** (2.0 / 3.0 * beta)).astype(float)   # :232-235
    n = int(mask.sum())
    rng = np.random.default_rng(57)
    q0 = rng.uniform(0.05, 30.0, n)
    q = rng.uniform(0.0, 2.0e-4, n)
    ldd2d = np.zeros(mask.shape)
    ldd2d[mask] = codes
    routing_case(kwp, name, mask.shape[0], mask.shape[1], 57, 0, 0, True, steps, False, beta=beta, dt=3600.0, ldd=ldd2d, mask=mask,
                 alpha=alpha, q0=q0, q=q, dx=length)
    path = os.path.join(HERE, name + ".npz")
    with np.load(path) as z:
        out = {k: z[k] for k in z.files}
    out.update(ldd_raw=ldd_raw[mask], pixarea=pixarea[mask], upArea=up_area[mask])
    np.savez_compressed(path, **out)
    print(name, "outlet links leaving the mask: %d, upstream area of the outlet %.4g m2" % (
        int(((ldd_ops.downstream_index(ldd_raw[mask], mask) < 0) & (ldd_raw[mask] != 5)).sum()), up_area[mask].max()))


def real_usecase_case(name, steps=6, dt_sec=86400.0):
    """The reference's test catchment END TO END as a fixture for the device: static state from the host init mirrors on the
    real input maps (bit-exact against the reference's own initial() there: tests/test_init_live_real_catchment.py), the raw
    meteo maps of the first `steps` steps from 02/01/2016 06:00 (float32, as stored), the LAI maps of their intervals, what
    the CPU restatement gives per step (oracle/ref_usecase.py::OracleRun), and the soil-moisture maps of the output stacks
    the reference SHIPS for that run (reference/output_reference_daily: tha, thfa, thia, thc, thfc, thic)."""
    import datetime
    from oracle import ref_usecase
    from lisflood_code_b200.hydrological_modules.snow import lai_interval
    R = ref_usecase.OracleRun(dt_sec=dt_sec, split=True)
    mask = R.mask
    out = {"steps": np.int64(steps)}
    for k, v in R.S.items():
        if isinstance(v, (np.ndarray, float, int, bool, np.floating, np.integer)):
            out["S__" + k] = np.asarray(v)
    for k, v in R.feeder.P.items():
        out["P__" + k] = np.asarray(v)
    out["P__kgb"] = np.asarray(R.kgb)
    out["Z__SnowCoverS"] = np.stack(R.feeder.SnowCoverS)
    out["Z__FrostIndex"] = R.feeder.FrostIndex.copy()
    shipped = {k: ref_usecase.shipped_output("output_reference_daily", k) for k in ("tha", "thfa", "thia", "thc", "thfc", "thic")}
    start = datetime.datetime(2016, 1, 2, 6, 0)
    keys = ("W1a", "W1b", "W2", "UZ", "LZ", "Theta1a", "Theta2", "ChanQAvg", "ChanQ", "ChanM3", "OFQDirect", "OFQOther",
            "OFQForest", "TotalRunoff", "CumInterception", "DSLR")
    for t in range(steps):
        date = start + datetime.timedelta(seconds=t * dt_sec)
        day = int(date.strftime("%j"))
        out["day%d" % t] = np.int64(day)
        for var_name, (data, tv, (unit_s, ref)) in R.forcing.items():
            idx = np.flatnonzero(tv == (date - ref).total_seconds() / unit_s)[0]
            out["R%d__%s" % (t, var_name)] = data[idx][mask]
        out["L%d" % t] = np.stack([R.lai[i][lai_interval(day)][mask].astype(np.float64) for i in range(3)])
        v = R.step(date)
        for k in keys:
            out["O%d__%s" % (t, k)] = np.asarray(getattr(v, k)).copy()
        for k in ("FrostIndex",):
            out["O%d__%s" % (t, k)] = np.asarray(getattr(R.feeder, k)).copy()
        out["O%d__SnowCoverS" % t] = np.stack(R.feeder.SnowCoverS)
        for k, stack in shipped.items():
            out["X%d__%s" % (t, k)] = stack[t][mask]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "n=%d steps=%d: %d arrays" % (int(mask.sum()), steps, len(out)))


def soil_options_case(name, rows, cols, seed, steps=2):
    """The option-gated extras of soilloop.dynamic_soil, executed by the reference's OWN class with the options switched on:
    repStressDays (SoilMoistureStressDays, soilloop.py:597-598) and simulatePF (the nested Numba kernel
    suctionUnsaturatedSoilPF, soilloop.py:673-705: pF0, pF1, pF2 from the end-of-step soil moisture)."""
    from oracle import ref_modules
    S = synthetic.full_stack(rows, cols, seed=seed, split_routing=False, mask_fraction=0.08, channel_threshold=12)
    n = S["N"]
    rng = np.random.default_rng(seed + 1000)
    M = ref_modules.RefModel(S, options={"simulatePF": True, "repStressDays": True})
    v = M.var
    extra = {"HeadMax": np.float64(1.0e7)}
    for lay in ("1a", "1b", "2"):
        extra["GenuInvAlpha" + lay] = 1 / rng.uniform(0.004, 0.2, (3, n))
        extra["GenuInvN" + lay] = 1 - 1 / S["GenuInvM" + lay]           # n = 1 / (1 - m), soil.py:176-186
        for k in ("GenuInvAlpha" + lay, "GenuInvN" + lay):
            setattr(v, k, ref_modules.numpy_modified(extra[k].copy(), v.LU_DIMS))
    v.HeadMax = float(extra["HeadMax"])
    for k in ("pF0", "pF1", "pF2"):
        setattr(v, k, v.allocateVariableAllVegetation())
    out = {"steps": np.int64(steps)}
    out.update({"S__" + k: np.asarray(a) for k, a in S.items()})
    out.update({"X__" + k: a for k, a in extra.items()})
    for t in range(steps):
        F = synthetic.forcing(S, t, seed)
        if t == 1:
            F = dict(F, Rain=F["Rain"] * 0.0)        # a dry step: water stress on more columns
        for k, a in F.items():
            out["F%d__%s" % (t, k)] = a
        M.step(F)
        for k in ("pF0", "pF1", "pF2", "SoilMoistureStressDays", "RWS", "W1a", "W1b", "W2"):
            out["O%d__%s" % (t, k)] = np.asarray(getattr(v, k).values).copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    last = steps - 1
    print(name, "n=%d: stressed columns %.3f, pF range %.2f..%.2f, pF = -1 on %.4f" % (
        n, (out["O%d__SoilMoistureStressDays" % last] > 0).mean(), out["O%d__pF0" % last].min(), out["O%d__pF2" % last].max(),
        np.mean([np.mean(out["O%d__pF%d" % (last, i)] == -1) for i in range(3)])))


def main():
    import warnings
    warnings.simplefilter("ignore")
    if len(sys.argv) > 1 and sys.argv[1] == "realcatchment":
        kwpt, kwp, sl = ref_loader.load()
        real_catchment_case(kwp, "kwreal_etrs89_57x80")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "realusecase":
        real_usecase_case("realcase_etrs89_daily")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "soiloptions":
        soil_options_case("soilopt_28x33", 28, 33, 91)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "feeders":
        feeders_case("feeders_daily", 900, 51, 10, 86400.0, 150)
        feeders_case("feeders_6h", 600, 52, 8, 21600.0, 340)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "adversarial":
        kwpt, kwp, sl = ref_loader.load()
        adversarial_case(kwp, "kwadv_24x160_beta06", 0.6)
        adversarial_case(kwp, "kwadv_24x160_beta07", 0.7)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "structinit":
        structures_init_case("structinit_40x46_cold", 40, 46, 81, False)
        structures_init_case("structinit_38x42_warm", 38, 42, 82, True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "structures":
        structures_case("structures_40x46_single", 40, 46, 71, False, 4, 3, 2, channel_threshold=10)
        structures_case("structures_36x44_split_6h", 36, 44, 72, True, 4, 2, 2, channel_threshold=10, dt_sec=21600.0)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "init":
        init_case("init_24x31_split", 24, 31, 61, True)
        init_case("init_19x23_single_6h", 19, 23, 62, False, scalar_maps=True, dt_sec=21600.0, soilless_fraction=0.0)
        init_case("init_21x26_split_sound", 21, 26, 63, True, soilless_fraction=0.0)
        return
    model_case("model_26x34_single", 26, 34, 41, False, 3, mask_fraction=0.08, channel_threshold=12)
    model_case("model_30x28_split", 30, 28, 42, True, 3, mask_fraction=0.05, channel_threshold=10)
    model_case("model_20x22_6h", 20, 22, 43, False, 2, channel_threshold=8, dt_sec=21600.0)
    kwpt, kwp, sl = ref_loader.load()
    # known-answer vector of SURVEY.md §8c: 4x4, all south, bottom row pits
    ldd = np.full((4, 4), 2.0)
    ldd[3, :] = 5.0
    mask = np.ones((4, 4), bool)
    routing_case(kwp, "kw_4x4_south", 4, 4, 0, 0, 0, False, 3, False, ldd=ldd, mask=mask, alpha=np.full(16, 1.5),
                 q0=np.ones(16), q=np.full(16, 1e-4), dx=1000.0)
    routing_case(kwp, "kw_40x50_masked", 40, 50, 11, 1.0, 0.25, True, 6, True)
    routing_case(kwp, "kw_64x48_deep", 64, 48, 12, 0.2, 0.0, False, 6, False)
    routing_case(kwp, "kw_48x64_negq", 48, 64, 13, 2.0, 0.1, True, 6, True, negative_q=True)
    routing_case(kwp, "kw_33x29_beta08", 33, 29, 14, 0.5, 0.05, True, 4, False, beta=0.8, dt=21600.0)
    # degenerate shapes (a 1-pixel domain crashes the reference itself in _setRoutingOrders:
    # Series.squeeze() returns a scalar, kinematic_wave_parallel.py:151-155 -- no golden possible)
    routing_case(kwp, "kw_2x1_col", 2, 1, 15, 1.0, 0.0, False, 2, False)
    routing_case(kwp, "kw_1x17_row", 1, 17, 16, 1.0, 0.0, True, 3, False)


if __name__ == "__main__":
    main()
