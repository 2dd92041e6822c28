"""Seeded synthetic catchments for the parity tests and bench.py.

Everything here is host-side NumPy input generation; it is not part of the timed path.
LDD codes are the PCRaster keypad codes the reference consumes
(reference: src/lisflood/hydrological_modules/kinematic_wave_parallel.py:47-51):

      7 8 9        row 0 = north, so 8 = (row-1, col), 2 = (row+1, col), 6 = (row, col+1) ...
      4 5 6        5 = pit
      1 2 3
"""
import numpy as np

# (drow, dcol) -> keypad code
_NEIGH = [(-1, -1, 7), (-1, 0, 8), (-1, 1, 9), (0, -1, 4), (0, 1, 6), (1, -1, 1), (1, 0, 2), (1, 1, 3)]


def random_ldd(rows, cols, seed=0, noise=3.0, tilt=1.0, mask_fraction=0.0, single_outlet=False):
    """Random D8 drainage forest: elevation = tilted plane + Gaussian noise, every cell drains to
    its steepest-descent neighbour, a cell with no strictly lower neighbour is a pit (code 5).

    noise/tilt ~ 3   -> shallow forest (tens of levels, many pits)
    noise/tilt ~ 0.3 -> deep trees (about `rows` levels)
    single_outlet: the south edge collects everything into one outlet (one basin that must be CUT to be shared
    between GPUs).
    Returns (ldd_codes float64[rows, cols], land_mask bool[rows, cols]); cells outside the mask
    carry code 0.
    """
    rng = np.random.default_rng(seed)
    elev = rng.standard_normal((rows, cols)) * noise
    elev += tilt * np.arange(rows, dtype=np.float64)[::-1, None]          # drains to the south edge
    elev += 0.05 * tilt * np.abs(np.arange(cols, dtype=np.float64) - cols / 2)[None, :]  # towards a trunk valley
    if mask_fraction > 0:
        land = rng.random((rows, cols)) >= mask_fraction
    else:
        land = np.ones((rows, cols), bool)
    big = np.float64(1e300)
    e = np.where(land, elev, big)
    pad = np.full((rows + 2, cols + 2), big)
    pad[1:-1, 1:-1] = e
    best = np.zeros((rows, cols))
    code = np.full((rows, cols), 5.0)
    for dr, dc, k in _NEIGH:
        nb = pad[1 + dr:1 + dr + rows, 1 + dc:1 + dc + cols]
        dist = np.sqrt(float(dr * dr + dc * dc))
        drop = (e - nb) / dist
        drop[nb >= big] = -1.0
        better = drop > best
        best = np.where(better, drop, best)
        code = np.where(better, float(k), code)
    if single_outlet:
        # one big basin: the south edge becomes a collector that drains to its centre cell (the only outlet of
        # everything that reaches the edge); interior sinks remain separate small catchments
        c0 = cols // 2
        code[rows - 1, :c0] = 6.0
        code[rows - 1, c0 + 1:] = 4.0
        code[rows - 1, c0] = 5.0
        land = land.copy()
        land[rows - 1, :] = True
    code[~land] = 0.0
    return code, land


def routing_fields(n, seed=0):
    """alpha, Q0, q for config C2: alpha~U(0.5,3), Q0~U(0.1,10), q~U(0,1e-4)."""
    rng = np.random.default_rng(seed + 7919)
    alpha = rng.uniform(0.5, 3.0, n)
    q0 = rng.uniform(0.1, 10.0, n)
    q = rng.uniform(0.0, 1e-4, n)
    return alpha, q0, q


def level_stats(order_start_stop):
    sizes = order_start_stop[:, 1] - order_start_stop[:, 0]
    return {"levels": int(sizes.size), "median_level": float(np.median(sizes)), "max_level": int(sizes.max()),
            "min_level": int(sizes.min())}


# ------------------------------------------------------------------------------------------------
# Full soil + overland + channel stack (config C3 of BASELINE.json)
# ------------------------------------------------------------------------------------------------
VEG = ["Rainfed_prescribed", "Forest_prescribed", "Irrigated_prescribed"]
LANDUSE = ["Rainfed", "Forest", "Irrigated"]


def mualem(w_res, w_sat, genu_alpha, genu_n, genu_m, head_cm):
    """Soil moisture at pressure head [cm] (reference: hydrological_modules/soil.py:30-35)."""
    return w_res + (w_sat - w_res) / ((1 + (genu_alpha * head_cm) ** genu_n) ** genu_m)


def full_stack(rows, cols, seed=0, ldd_noise=0.5, mask_fraction=0.0, channel_threshold=60, split_routing=False,
               dt_sec=86400.0, dt_sec_channel=3600.0, frozen_fraction=0.05, beta=0.6):
    """Static parameters + initial state of the hot path on a seeded synthetic catchment.

    Keys follow the reference's `self.var.<name>` attributes.  (V, N) / (L, N) arrays are
    C-contiguous float64; land use 2 (Irrigated) shares the parameter maps of land use 0 exactly like the
    reference's `defsoil` without a third map (Lisflood_initial.py:371-391).
    """
    from .global_modules import ldd_ops
    rng = np.random.default_rng(seed + 4242)
    ldd, mask = random_ldd(rows, cols, seed=seed, noise=ldd_noise, mask_fraction=mask_fraction)
    n = int(mask.sum())
    S = {"rows": rows, "cols": cols, "mask": mask, "N": n, "Ldd": ldd[mask].astype(np.float64)}
    S["Ldd"] = ldd_ops.lddrepair_codes(S["Ldd"], mask)

    def U(lo, hi, shape=None):
        return rng.uniform(lo, hi, (n,) if shape is None else shape)

    def landuse3(a_other, a_forest):
        return np.ascontiguousarray(np.stack([a_other, a_forest, a_other]))

    # ---- time / geometry constants (miscInitial.py:44-181, routing.py:61-82)
    S["DtSec"], S["DtSecChannel"] = float(dt_sec), float(dt_sec_channel)
    S["DtDay"] = S["DtSec"] / 86400.0
    S["InvDtSec"], S["InvDtDay"] = 1 / S["DtSec"], 1 / S["DtDay"]
    S["NoRoutSteps"] = int(max(1, round(S["DtSec"] / S["DtSecChannel"], 0)))
    S["DtRouting"] = S["DtSec"] / S["NoRoutSteps"]
    S["InvDtRouting"] = 1 / S["DtRouting"]
    S["InvNoRoutSteps"] = 1 / float(S["NoRoutSteps"])
    S["PixelLength"] = 5000.0
    S["InvPixelLength"] = 1.0 / S["PixelLength"]
    S["PixelArea"] = np.full(n, S["PixelLength"] ** 2)
    S["MMtoM"] = 0.001
    S["MMtoM3"] = 0.001 * S["PixelArea"]
    S["M3toMM"] = 1 / S["MMtoM3"]
    S["Beta"] = float(beta)
    S["InvBeta"] = 1 / S["Beta"]
    S["AlpPow"] = 2.0 / 3.0 * S["Beta"]

    # ---- fractions (Dirichlet over other / forest / irrigated / sealed / water)
    fr = rng.dirichlet([4.0, 3.0, 1.0, 0.6, 0.3], n).T
    S["SoilFraction"] = np.ascontiguousarray(fr[:3])
    S["DirectRunoffFraction"] = np.ascontiguousarray(fr[3])
    S["WaterFraction"] = np.ascontiguousarray(fr[4])

    # ---- soil hydraulic parameters per land use and layer (soil.py:109-228)
    for lay, (d_lo, d_hi) in (("1a", (40, 60)), ("1b", (200, 300)), ("2", (500, 900))):
        if lay == "2":
            depth = U(d_lo, d_hi)
            depth = landuse3(depth, depth)
            ths, thr = U(.4, .5), U(.02, .08)
            lam, gal = U(.15, .45), U(.005, .05)
            ks = np.exp(U(np.log(1.0), np.log(500.0)))
            ths, thr, lam, gal, ks = (landuse3(x, x) for x in (ths, thr, lam, gal, ks))
        else:
            depth = landuse3(U(d_lo, d_hi), U(d_lo, d_hi))
            # (zero-depth soils are not generated: the reference's satFun divides by WFC-WWP = 0 and Numba
            #  raises ZeroDivisionError, hydrological_modules/soilloop.py:393-396)
            ths, thr = landuse3(U(.4, .5), U(.4, .5)), landuse3(U(.02, .08), U(.02, .08))
            lam, gal = landuse3(U(.15, .45), U(.15, .45)), landuse3(U(.005, .05), U(.005, .05))
            ks = landuse3(np.exp(U(np.log(1.0), np.log(500.0))), np.exp(U(np.log(1.0), np.log(500.0))))
        gn = 1 + lam
        gm = lam / gn
        S["SoilDepth" + lay], S["KSat" + lay] = depth, ks
        S["GenuM" + lay], S["GenuInvM" + lay] = gm, 1 / gm
        ws, wres = ths * depth, thr * depth
        S["WS" + lay], S["WRes" + lay] = ws, wres
        S["WFC" + lay] = mualem(wres, ws, gal, gn, gm, 100)
        S["WPF3" + {"1a": "a", "1b": "b", "2": "2"}[lay]] = mualem(wres, ws, gal, gn, gm, 1000)
        S["WWP" + lay] = mualem(wres, ws, gal, gn, gm, 15000)
        S["PoreSpaceNotZero" + lay] = np.logical_and(depth != 0, ws != 0)
    for nm in ("WS", "WRes", "WFC", "WWP"):
        S[nm + "1"] = S[nm + "1a"] + S[nm + "1b"]
    S["WPF3"] = S["WPF3a"] + S["WPF3b"]
    S["SoilDepthTotal"] = S["SoilDepth1a"] + S["SoilDepth1b"] + S["SoilDepth2"]
    S["b_Xinanjiang"] = U(.1, .7)
    S["PowerInfPot"] = (S["b_Xinanjiang"] + 1) / S["b_Xinanjiang"]
    S["StoreMaxPervious"] = S["WS1"] / (S["b_Xinanjiang"] + 1)
    S["PowerPrefFlow"] = U(1.0, 5.0)
    S["CropCoef"] = np.ascontiguousarray(np.stack([U(.9, 1.1), U(.9, 1.3), U(.9, 1.2)]))
    S["CropGroupNumber"] = np.ascontiguousarray(np.stack([U(1.0, 5.0), U(2.0, 5.0), U(1.0, 5.0)]))
    S["CourantCrit"] = 0.4
    S["LeafDrainageK"] = float(min(S["DtDay"] * (1 / 1.0), 1))
    S["AvWaterThreshold"] = 5.0 * S["DtDay"]
    S["DrainedFraction"] = 0.0
    S["SMaxSealed"] = 1.0
    S["kgb"] = 0.75 * 0.72

    # ---- groundwater (groundwater.py:44-132, miscInitial.py:127-133)
    S["UpperZoneK"] = np.minimum(S["DtDay"] * (1 / U(5.0, 20.0)), 1)
    S["LowerZoneK"] = np.minimum(S["DtDay"] * (1 / U(50.0, 500.0)), 1)
    gwloss = np.zeros(n)
    S["GwPercStep"] = np.maximum(U(0.2, 1.5), gwloss) * S["DtDay"]
    S["GwLossStep"] = gwloss * S["DtDay"]
    S["LZThreshold"] = U(0.0, 20.0)

    # ---- initial state (soil.py:268-277, 380-410; groundwater.py:97-118)
    for lay in ("1a", "1b", "2"):
        w = np.empty((3, n))
        for v in range(3):
            w[v] = np.where(S["PoreSpaceNotZero" + lay][v], S["WFC" + lay][v] * U(0.7, 1.1), 0)
            w[v] = np.minimum(w[v], S["WS" + lay][v])
        S["W" + lay] = w
    S["W1"] = S["W1a"] + S["W1b"]
    S["UZ"] = np.ascontiguousarray(U(0.0, 10.0, (3, n)))
    S["LZ"] = U(20.0, 200.0)
    S["DSLR"] = np.ascontiguousarray(np.floor(U(1.0, 6.0, (3, n))))
    S["CumInterception"] = np.ascontiguousarray(U(0.0, 0.5, (3, n)))
    S["CumInterSealed"] = U(0.0, 0.5)
    S["LZInflowCUM"] = np.zeros(n)

    # ---- drainage networks (routing.py:90-170)
    ds = ldd_ops.downstream_index(S["Ldd"], mask)
    uparea = ldd_ops.accuflux(ds, np.ones(n))
    S["IsChannel"] = uparea >= channel_threshold
    S["IsChannelKinematic"] = S["IsChannel"].copy()
    S["LddKinematic"] = ldd_ops.lddrepair_codes(ldd_ops.lddmask_codes(S["Ldd"], S["IsChannel"]), mask)
    # cells draining into a non-channel cell cannot exist (channels are downstream-closed by construction)
    S["LddToChan"] = np.where(S["IsChannel"], 5.0, S["Ldd"])
    S["AtLastPointC"] = ds < 0

    # ---- channel geometry (routing.py:184-253)
    S["ChanLength"] = S["PixelLength"] * U(1.0, 1.4)
    S["InvChanLength"] = 1 / S["ChanLength"]
    chan_grad = np.maximum(U(1e-4, 5e-3), 1e-5)
    chan_man = U(0.02, 0.06)
    width = 2.0 + 0.5 * np.sqrt(uparea)
    depth_thr = 0.5 + 0.05 * np.sqrt(uparea)
    sdxdy = 1.0
    S["ChanBottomWidth"] = width
    upper = width + 2 * sdxdy * depth_thr
    bankfull = 0.5 * depth_thr * (upper + width)
    S["TotalCrossSectionAreaBankFull"] = bankfull
    S["TotalCrossSectionArea"] = 0.5 * bankfull
    wd_alpha = np.where(S["IsChannel"], 0.5 * depth_thr, 0.0)
    S["ChanWettedPerimeterAlpha"] = width + 2 * np.sqrt(np.square(wd_alpha) + np.square(wd_alpha * sdxdy))
    alp_term = (chan_man / np.sqrt(chan_grad)) ** S["Beta"]
    S["ChannelAlpha"] = (alp_term * (S["ChanWettedPerimeterAlpha"] ** S["AlpPow"])).astype(float)
    S["InvChannelAlpha"] = 1 / S["ChannelAlpha"]
    S["ChanM3"] = S["TotalCrossSectionArea"] * S["ChanLength"]
    S["ChanM3Kin"] = S["ChanM3"].copy()
    S["ChanQKin"] = np.where(S["ChannelAlpha"] > 0, (S["TotalCrossSectionArea"] / S["ChannelAlpha"]) ** S["InvBeta"], 0)
    S["ChanQ"] = S["ChanQKin"].copy()
    S["SplitRouting"] = bool(split_routing)
    if split_routing:  # routing.py:353-397
        man2 = chan_man * 3.0
        S["ChannelAlpha2"] = (((man2 / np.sqrt(chan_grad)) ** S["Beta"]) * (S["ChanWettedPerimeterAlpha"] ** S["AlpPow"]))
        S["InvChannelAlpha2"] = 1 / S["ChannelAlpha2"]
        dsk = ldd_ops.downstream_index(S["LddKinematic"], mask)
        avgdis = np.where(S["IsChannel"], 0.002 * ldd_ops.accuflux(dsk, np.where(S["IsChannel"], 1.0, 0.0)) + 0.01, 0.0)
        S["QLimit"] = avgdis * 2.0
        S["M3Limit"] = S["ChannelAlpha"] * S["ChanLength"] * (S["QLimit"] ** S["Beta"])
        S["Chan2M3Start"] = S["ChannelAlpha2"] * S["ChanLength"] * (S["QLimit"] ** S["Beta"])
        S["Chan2QStart"] = S["QLimit"] - ldd_ops.upstream_sum(dsk, S["QLimit"])
        S["CrossSection2Area"] = np.zeros(n)
        S["Sideflow1Chan"] = np.zeros(n)
        S["Chan2M3Kin"] = S["CrossSection2Area"] * S["ChanLength"] + S["Chan2M3Start"]
        S["ChanM3Kin"] = S["ChanM3"] - S["Chan2M3Kin"] + S["Chan2M3Start"]
        S["ChanM3Kin"] = np.where((S["ChanM3Kin"] < 0.0) & (S["ChanM3Kin"] > -0.0000001), 0.0, S["ChanM3Kin"])
        S["Chan2QKin"] = (S["Chan2M3Kin"] * S["InvChanLength"] * S["InvChannelAlpha2"]) ** S["InvBeta"]
        S["ChanQKin"] = (S["ChanM3Kin"] * S["InvChanLength"] * S["InvChannelAlpha"]) ** S["InvBeta"]

    # ---- overland flow (surface_routing.py:44-95; runoff order = Other, Forest, Direct)
    grad = np.maximum(U(1e-3, 0.1), 1e-4)
    nman = np.stack([U(0.05, 0.2), U(0.1, 0.4), np.full(n, 0.02)])
    of_wp = S["PixelLength"] + 2 * S["MMtoM"] * 5.0
    S["OFAlpha"] = np.ascontiguousarray(((nman / np.sqrt(grad)) ** S["Beta"]) * (of_wp ** S["AlpPow"]))
    S["InvOFAlpha"] = 1 / S["OFAlpha"]
    for k, nm in enumerate(("Other", "Forest", "Direct")):
        S["OFM3" + nm] = np.zeros(n)
        S["OFQ" + nm] = ((S["OFM3" + nm] * S["InvPixelLength"] * S["InvOFAlpha"][k]) ** S["InvBeta"]).astype(float)
    return S


def complete_stack(S):
    """Adds the derived maps / scalars the reference keeps on `self.var` next to the base maps (sums over the two
    top-soil layers, inverses, unit factors; soil.py:212-228,361-363, routing.py:184-253) to a dictionary that holds
    only what the device model is given.  Existing keys are left untouched."""
    n = int(S["N"])
    D = dict(S)

    def need(k, f):
        if k not in D:
            D[k] = f()

    need("DtDay", lambda: D["DtSec"] / 86400.0)
    need("InvDtSec", lambda: 1 / D["DtSec"])
    need("InvDtDay", lambda: 1 / D["DtDay"])
    need("DtRouting", lambda: D["DtSec"] / D["NoRoutSteps"])
    need("InvDtRouting", lambda: 1 / D["DtRouting"])
    need("InvNoRoutSteps", lambda: 1 / float(D["NoRoutSteps"]))
    need("InvPixelLength", lambda: 1.0 / D["PixelLength"])
    need("PixelArea", lambda: np.full(n, D["PixelLength"] ** 2))
    need("MMtoM", lambda: 0.001)
    need("MMtoM3", lambda: 0.001 * D["PixelArea"])
    need("M3toMM", lambda: 1 / D["MMtoM3"])
    need("InvBeta", lambda: 1 / D["Beta"])
    need("kgb", lambda: 0.75 * 0.72)
    for lay in ("1a", "1b", "2"):
        need("GenuM" + lay, lambda: 1 / D["GenuInvM" + lay])
        need("PoreSpaceNotZero" + lay, lambda: np.logical_and(D["SoilDepth" + lay] != 0, D["WS" + lay] != 0))
    for nm in ("WS", "WRes", "WFC", "WWP", "W"):
        need(nm + "1", lambda: D[nm + "1a"] + D[nm + "1b"])
    need("SoilDepthTotal", lambda: D["SoilDepth1a"] + D["SoilDepth1b"] + D["SoilDepth2"])
    need("PowerInfPot", lambda: (D["b_Xinanjiang"] + 1) / D["b_Xinanjiang"])
    need("StoreMaxPervious", lambda: D["WS1"] / (D["b_Xinanjiang"] + 1))
    need("LZInflowCUM", lambda: np.zeros(n))
    need("InvChanLength", lambda: 1 / D["ChanLength"])
    need("InvChannelAlpha", lambda: 1 / D["ChannelAlpha"])
    if D.get("SplitRouting"):
        need("InvChannelAlpha2", lambda: 1 / D["ChannelAlpha2"])
    need("ChanQ", lambda: D["ChanQKin"].copy())
    for nm in ("Direct", "Other", "Forest"):
        need("OFQ" + nm, lambda: np.zeros(n))
    return D


def forcing(S, step, seed=0):
    """Seeded meteorological forcing of model step `step`: intermittent Gamma rain,
    ETRef/EWRef ~ U(0,6) mm/day, LAI ~ U(0,6), a few frozen pixels."""
    n = S["N"]
    rng = np.random.default_rng((seed + 1) * 100003 + step)
    wet = rng.random(n) < 0.45
    F = {"Rain": np.where(wet, rng.gamma(0.8, 8.0, n), 0.0) * S["DtDay"],
         "SnowMelt": np.where(rng.random(n) < 0.1, rng.uniform(0, 3.0, n), 0.0) * S["DtDay"],
         "ETRef": rng.uniform(0, 6.0, n) * S["DtDay"], "EWRef": rng.uniform(0, 6.0, n) * S["DtDay"],
         "LAI": np.ascontiguousarray(rng.uniform(0, 6.0, (3, n))),
         "isFrozenSoil": rng.random(n) < 0.05}
    F["ESRef"] = (F["EWRef"] + F["ETRef"]) / 2
    F["LAITerm"] = np.exp(-S["kgb"] * F["LAI"])
    return F


def raw_inputs(rows, cols, seed=0, ldd_noise=0.5, mask_fraction=0.1, channel_threshold=20, scalar_maps=False,
               soilless_fraction=0.03):
    """Raw static inputs by BINDING name (what `loadmap` returns: float or float64[N]) for the soil and routing
    modules' initial(), plus the attributes earlier modules leave on the model object.  Returns (mask, raw, state).
    Used by tests/golden/make_golden.py (fed to the reference's own initial()) and by the init tests."""
    from .global_modules import ldd_ops
    rng = np.random.default_rng(seed + 977)
    ldd, mask = random_ldd(rows, cols, seed=seed, noise=ldd_noise, mask_fraction=mask_fraction)
    n = int(mask.sum())
    U = lambda lo, hi: rng.uniform(lo, hi, n)
    codes = ldd_ops.lddrepair_codes(ldd[mask].astype(np.float64), mask)
    uparea = ldd_ops.accuflux(ldd_ops.downstream_index(codes, mask), np.ones(n))
    raw = {"Ldd": codes, "Channels": (uparea >= channel_threshold).astype(np.float64), "beta": 0.6,
           "ChanLength": 5000.0 * U(1.0, 1.4), "ChanGrad": U(0.0, 5e-3), "ChanGradMin": 1e-4, "CalChanMan": U(0.5, 2.0),
           "ChanMan": U(0.02, 0.06), "ChanBottomWidth": 2.0 + 0.5 * np.sqrt(uparea), "ChanDepthThreshold": 0.5 + 0.05 * np.sqrt(uparea),
           "ChanSdXdY": 1.0, "TotalCrossSectionAreaInitValue": np.where(rng.random(n) < 0.5, -9999.0, U(0.5, 3.0)),
           "PrevDischarge": -9999.0, "CrossSection2AreaInitValue": np.where(rng.random(n) < 0.6, -9999.0, U(0.0, 0.05)),
           "PrevSideflowInitValue": -9999.0, "CalChanMan2": U(1.0, 4.0), "QSplitMult": 2.0,
           "AvgDis": np.where(uparea >= channel_threshold, 0.002 * uparea + 0.01, 0.0)}
    for i, (lo, hi) in zip("123", ((40, 60), (200, 300), (500, 900))):
        raw["SoilDepth" + i] = U(lo, hi)
        raw["SoilDepth%sForest" % i] = U(lo, hi)
        for stem, (a, b) in (("MapThetaSat", (.4, .5)), ("MapThetaRes", (.02, .08)), ("MapLambda", (.15, .45)),
                             ("MapGenuAlpha", (.005, .05))):
            raw[stem + i] = U(a, b)
            if i != "3":
                raw[stem + i + "Forest"] = U(a, b)
        raw["MapKSat" + i] = np.exp(U(np.log(1.0), np.log(500.0)))
        if i != "3":
            raw["MapKSat%sForest" % i] = np.exp(U(np.log(1.0), np.log(500.0)))
    raw["SoilDepth1"][rng.random(n) < soilless_fraction] = 0.0   # soil-less pixels: PoreSpaceNotZero False
    raw.update({"CourantCrit": 0.4, "LeafDrainageTimeConstant": 1.0, "AvWaterRateThreshold": 5.0, "MapCropCoef": U(.9, 1.1),
                "MapForestCropCoef": U(.9, 1.3), "MapIrrigationCropCoef": U(.9, 1.2), "MapCropGroupNumber": U(1, 5),
                "MapForestCropGroupNumber": U(2, 5), "MapIrrigationCropGroupNumber": U(1, 5), "MapN": U(.05, .2),
                "MapForestN": U(.1, .4), "b_Xinanjiang": 0.3 if scalar_maps else U(.1, .7),
                "PowerPrefFlow": 3.5 if scalar_maps else U(1, 5), "CumIntSealedInitValue": 0.0, "SMaxSealed": 1.0,
                "DrainedFraction": 0.25})
    for i in "123":
        raw["ThetaInit%sValue" % i] = -9999.0
        raw["ThetaForestInit%sValue" % i] = np.where(rng.random(n) < 0.5, -9999.0, U(.1, .4))
        raw["ThetaIrrigationInit%sValue" % i] = 0.25
    for stem, (lo, hi) in (("DSLR", (0.0, 6.0)), ("CumInt", (0.0, 0.5))):
        raw[stem + "InitValue"], raw[stem + "ForestInitValue"], raw[stem + "IrrigationInitValue"] = U(lo, hi), U(lo, hi), U(lo, hi)
    fr = rng.dirichlet([4.0, 3.0, 1.0, 0.6, 0.3, 0.2], n).T
    state = {"SoilFraction": np.ascontiguousarray(fr[:3]), "OtherFraction": fr[0].copy(), "ForestFraction": fr[1].copy(),
             "IrrigationFraction": fr[2].copy(), "DirectRunoffFraction": fr[3].copy(), "WaterFraction": fr[4].copy(),
             "RiceFraction": fr[5].copy(), "PixelArea": np.full(n, 25.0e6)}
    # overland flow, groundwater, grid size (drawn last: the earlier inputs keep their values)
    raw.update({"OFOtherInitValue": 0.0, "OFForestInitValue": U(0.0, 50.0), "OFDirectInitValue": 0.0, "Grad": U(0.0, 0.1),
                "GradMin": 1e-3, "OFDepRef": 5.0, "UpperZoneTimeConstant": 10.0 if scalar_maps else U(5.0, 20.0),
                "LowerZoneTimeConstant": U(50.0, 500.0), "LZAvInflowMap": U(0.0, 2.0),
                "LZInitValue": np.where(rng.random(n) < 0.5, -9999.0, U(20.0, 200.0)), "LZThreshold": U(0.0, 20.0),
                "UZInitValue": 0.0, "UZForestInitValue": U(0.0, 10.0), "UZIrrigationInitValue": 2.5,
                "PixelLengthUser": 5000.0, "GwLoss": 0.0 if scalar_maps else U(0.0, 0.3), "GwPercValue": U(0.2, 1.5)})
    return mask, raw, state


def add_structures(S, n_reservoirs=3, n_lakes=2, seed=0):
    """Adds reservoirs and lakes to a full_stack() dictionary: the per-structure ("CC") parameter and
    state arrays with the reference's attribute names (hydrological_modules/reservoir.py:60-170, lakes.py:60-196), the
    structures' LDD surgery (structures.py:43-61: cells just upstream of a structure become pits of LddKinematic, inflow
    is gathered through `downstruct` built on the unmodified network, routing.py:151-157).  Sites are channel pixels
    with upstream channel cells, away from outlets and from each other."""
    from .global_modules import ldd_ops
    rng = np.random.default_rng(seed + 555)
    n, mask = S["N"], S["mask"]
    ldd_kin = np.asarray(S["LddKinematic"], np.float64)
    dsk = ldd_ops.downstream_index(ldd_kin, mask)
    nups = np.bincount(dsk[dsk >= 0], minlength=n)
    cand = np.flatnonzero(S["IsChannel"] & (dsk >= 0) & (nups > 0))
    rng.shuffle(cand)
    sites, taken = [], np.zeros(n, bool)
    for p in cand:
        if taken[p] or taken[dsk[p]] or taken[np.flatnonzero(dsk == p)].any():
            continue
        sites.append(int(p))
        taken[p] = True
        taken[dsk[p]] = True
        taken[np.flatnonzero(dsk == p)] = True
        if len(sites) == n_reservoirs + n_lakes:
            break
    if len(sites) < n_reservoirs + n_lakes:
        raise ValueError("catchment too small for %d structures" % (n_reservoirs + n_lakes))
    res = np.sort(np.array(sites[:n_reservoirs], np.int64))
    lak = np.sort(np.array(sites[n_reservoirs:], np.int64))
    S["downstruct"] = np.where(dsk >= 0, dsk, n).astype(np.int32)       # routing.py:151-157 (pits get N)
    struct = np.zeros(n, bool)
    struct[res] = True
    struct[lak] = True
    S["IsStructureKinematic"] = struct
    ups = (dsk >= 0) & struct[np.maximum(dsk, 0)]                       # structures.py:51-55
    S["IsUpsOfStructureKinematicC"] = ups
    S["LddStructuresKinematic"] = ldd_kin.copy()
    S["LddKinematic"] = np.where(ups, 5.0, ldd_kin)                     # structures.py:59
    q_site = lambda idx: np.maximum(np.bincount(S["downstruct"], weights=S["ChanQ"], minlength=n + 1)[idx], 0.05)
    R = res.size
    if R:
        S["simulateReservoirs"] = True
        S["ReservoirIndex"] = res
        sc = np.zeros(n)
        sc[res] = np.arange(1, R + 1)
        S["ReservoirSitesC"] = sc
        qn = q_site(res)
        S["TotalReservoirStorageM3CC"] = qn * 86400.0 * rng.uniform(3.0, 30.0, R)
        S["ConservativeStorageLimitCC"] = rng.uniform(0.05, 0.15, R)
        S["NormalStorageLimitCC"] = rng.uniform(0.4, 0.6, R)
        S["FloodStorageLimitCC"] = rng.uniform(0.85, 0.97, R)
        S["MinReservoirOutflowCC"] = qn * rng.uniform(0.05, 0.2, R)
        S["NonDamagingReservoirOutflowCC"] = qn * rng.uniform(3.0, 6.0, R)
        norm = qn * rng.uniform(0.6, 1.2, R)                              # reservoir.py:141-145
        norm = np.where(norm > S["MinReservoirOutflowCC"], norm, S["MinReservoirOutflowCC"] + 0.01)
        S["NormalReservoirOutflowCC"] = np.where(norm < S["NonDamagingReservoirOutflowCC"], norm,
                                                 S["NonDamagingReservoirOutflowCC"] - 0.01)
        S["Normal_FloodStorageLimitCC"] = S["NormalStorageLimitCC"] + 0.5 * (S["FloodStorageLimitCC"] - S["NormalStorageLimitCC"])
        S["DeltaO"] = S["NormalReservoirOutflowCC"] - S["MinReservoirOutflowCC"]
        S["DeltaLN"] = S["NormalStorageLimitCC"] - 2 * S["ConservativeStorageLimitCC"]
        S["DeltaLF"] = S["FloodStorageLimitCC"] - S["NormalStorageLimitCC"]
        S["DeltaNFL"] = S["FloodStorageLimitCC"] - S["Normal_FloodStorageLimitCC"]
        fill = rng.uniform(0.1, 1.0, R)                                   # spans all four outflow regimes
        S["ReservoirFillCC"] = fill
        S["ReservoirStorageM3CC"] = fill * S["TotalReservoirStorageM3CC"]
        S["ReservoirStorageM3"] = np.zeros(n)
        S["ReservoirStorageM3"][res] = S["ReservoirStorageM3CC"]
    Lk = lak.size
    if Lk:
        S["simulateLakes"] = True
        S["LakeIndex"] = lak
        sc = np.zeros(n)
        sc[lak] = np.arange(1, Lk + 1)
        S["LakeSitesC2"] = sc
        ql = q_site(lak)
        S["LakeAreaCC"] = rng.uniform(2.0e6, 5.0e7, Lk)
        S["LakeACC"] = rng.uniform(5.0, 60.0, Lk)
        S["LakeAvNetCC"] = ql.copy()
        storage = S["LakeAreaCC"] * np.sqrt(S["LakeAvNetCC"] / S["LakeACC"])       # lakes.py:113-117
        S["LakeLevelCC"] = storage / S["LakeAreaCC"]
        S["LakeInflowOldCC"] = np.bincount(S["downstruct"], weights=S["ChanQ"], minlength=n + 1)[lak]
        S["LakeFactor"] = S["LakeAreaCC"] / (S["DtRouting"] * np.sqrt(S["LakeACC"]))
        S["LakeFactorSqr"] = np.square(S["LakeFactor"])
        indicator = storage / S["DtRouting"] + S["LakeAvNetCC"] / 2
        S["LakeOutflowCC"] = np.square(-S["LakeFactor"] + np.sqrt(S["LakeFactorSqr"] + 2 * indicator))
        S["LakeStorageM3CC"] = storage.copy()
        S["LakeStorageM3BalanceCC"] = storage.copy()
        S["LakeStorageM3"] = np.zeros(n)
        S["LakeStorageM3"][lak] = storage
    return S
