"""CPU restatement (NumPy float64 + the C kernels of lisf_oracle*.c) of one model time step of the
LISFLOOD hot path.  TEST INFRASTRUCTURE ONLY -- see the header of lisf_oracle.c.

Call order and arithmetic follow the reference (paths relative to src/lisflood/):
  canopy            hydrological_modules/soilloop.py:519-627    (+ kernel :27-70)
  soil columns      hydrological_modules/soilloop.py:630-665    (+ kernel :78-355)
  open/sealed       hydrological_modules/opensealed.py:41-71
  per-pixel sums    hydrological_modules/soil.py:471-514, Lisflood_initial.py:393-396
  groundwater       hydrological_modules/groundwater.py:134-180
  surface routing   hydrological_modules/surface_routing.py:115-212
  channel sub-steps hydrological_modules/routing.py:435-706, Lisflood_dynamic.py:176-229
It is pinned by tests/test_oracle_golden.py against golden vectors produced by the reference's own
module classes (oracle/ref_modules.py -> tests/golden/make_golden.py).
"""
import ctypes as C
import types

import numpy as np

from . import lisf_oracle as lo

_D = C.POINTER(C.c_double)
_U8 = C.POINTER(C.c_uint8)
_I64 = C.POINTER(C.c_int64)

_SOIL_FIELDS = [
    ("num_vegs", C.c_int64), ("num_pixs", C.c_int64), ("index_landuse_all", _I64), ("is_irrigated", _U8),
    ("DtDay", C.c_double), ("AvWaterThreshold", C.c_double), ("CourantCrit", C.c_double), ("DrainedFraction", C.c_double),
    ("AvailableWaterForInfiltration", _D), ("Rain", _D), ("SnowMelt", _D), ("LeafDrainage", _D), ("Interception", _D),
    ("DSLR", _D), ("ESAct", _D), ("ESMax", _D), ("isFrozenSoil", _U8), ("b_Xinanjiang", _D), ("StoreMaxPervious", _D),
    ("PowerInfPot", _D), ("PrefFlow", _D), ("PowerPrefFlow", _D), ("Infiltration", _D),
    ("PoreSpaceNotZero1a", _U8), ("PoreSpaceNotZero1b", _U8), ("PoreSpaceNotZero2", _U8),
    ("KSat1a", _D), ("KSat1b", _D), ("KSat2", _D), ("GenuInvM1a", _D), ("GenuInvM1b", _D), ("GenuInvM2", _D),
    ("GenuM1a", _D), ("GenuM1b", _D), ("GenuM2", _D),
    ("W1a", _D), ("W1b", _D), ("W1", _D), ("W2", _D), ("Theta1a", _D), ("Theta1b", _D), ("Theta2", _D),
    ("Sat1a", _D), ("Sat1b", _D), ("Sat1", _D), ("Sat2", _D),
    ("SeepTopToSubA", _D), ("SeepTopToSubB", _D), ("SeepSubToGW", _D),
    ("WRes1a", _D), ("WRes1b", _D), ("WRes1", _D), ("WRes2", _D), ("WWP1a", _D), ("WWP1b", _D), ("WWP1", _D), ("WWP2", _D),
    ("WFC1a", _D), ("WFC1b", _D), ("WFC1", _D), ("WFC2", _D),
    ("SoilDepth1a", _D), ("SoilDepth1b", _D), ("SoilDepth2", _D), ("WS1a", _D), ("WS1b", _D), ("WS1", _D), ("WS2", _D),
    ("UpperZoneK", _D), ("GwPercStep", _D), ("UZOutflow", _D), ("UZ", _D), ("GwPercUZLZ", _D), ("NoSubS_out", _I64),
]


class SoilArgs(C.Structure):
    _fields_ = _SOIL_FIELDS


def _lib():
    L = lo.lib()
    if not hasattr(L, "_soil_ready"):
        L.lfo_interception.argtypes = [_D, _D, _D, _D, _D, _D, _D, C.c_double, C.c_int64, C.c_int64]
        L.lfo_interception.restype = None
        L.lfo_soil_columns.argtypes = [C.POINTER(SoilArgs)]
        L.lfo_soil_columns.restype = None
        L._soil_ready = True
    return L


def _p(a, t=_D):
    return a.ctypes.data_as(t)


def interception_water_balance(Interception, TaInterception, LeafDrainage, CumInterception, LAI, Rain,
                               TaInterceptionMax, drainageK):
    """Same 8-argument in-place signature as the reference kernel (soilloop.py:27-28)."""
    V, N = Interception.shape
    _lib().lfo_interception(_p(Interception), _p(TaInterception), _p(LeafDrainage), _p(CumInterception),
                            _p(np.ascontiguousarray(LAI)), _p(Rain), _p(np.ascontiguousarray(TaInterceptionMax)),
                            float(drainageK), V, N)


def suction_unsaturated_soil_pf(index_landuse_all, pF0, pF1, pF2, W1a, W1b, W2, WRes1a, WRes1b, WRes2, WS1a, WS1b, WS2,
                                PoreSpaceNotZero1a, PoreSpaceNotZero1b, PoreSpaceNotZero2, GenuInvAlpha1a, GenuInvAlpha1b,
                                GenuInvAlpha2, GenuInvM1a, GenuInvM1b, GenuInvM2, GenuInvN1a, GenuInvN1b, GenuInvN2, HeadMax):
    """suctionUnsaturatedSoilPF (soilloop.py:402-424 / :673-697) with saturationDegree (:379-383) and pressureHead
    (:428-432) restated in NumPy: same 26 arguments, pF0..2 written in place."""
    idx = np.asarray(index_landuse_all)
    layers = ((pF0, W1a, WRes1a, WS1a, PoreSpaceNotZero1a, GenuInvAlpha1a, GenuInvM1a, GenuInvN1a),
              (pF1, W1b, WRes1b, WS1b, PoreSpaceNotZero1b, GenuInvAlpha1b, GenuInvM1b, GenuInvN1b),
              (pF2, W2, WRes2, WS2, PoreSpaceNotZero2, GenuInvAlpha2, GenuInvM2, GenuInvN2))
    with np.errstate(all="ignore"):
        for pf, w, wres, ws, pore, inva, invm, invn in layers:
            wres, ws, pore = np.asarray(wres)[idx], np.asarray(ws)[idx], np.asarray(pore, bool)[idx]
            sat = np.where(pore, np.maximum(np.minimum((w - wres) / (ws - wres), 1.0), 0.0), 0.0)          # :379-383
            safe = np.where(sat == 0, 1.0, sat)
            head = np.minimum(HeadMax, np.asarray(inva)[idx] * ((1.0 / safe) ** np.asarray(invm)[idx] - 1.0) ** np.asarray(invn)[idx])
            head = np.where(sat == 0, HeadMax, head)                                                       # :429-432
            pf[...] = np.where(head > 0, np.log10(np.where(head > 0, head, 1.0)), -1.0)                    # :422-424


def soil_columns(v, ESMax, nosubs=None):
    """soilColumnsWaterBalance on the attributes of `v` (argument list of soilloop.py:645-665)."""
    a = SoilArgs()
    keep = []
    V, N = v.Interception.shape
    a.num_vegs, a.num_pixs = V, N
    idx = np.arange(V, dtype=np.int64)
    irr = np.array([0, 0, 1], np.uint8)
    frozen = np.ascontiguousarray(v.isFrozenSoil).astype(np.uint8)
    keep += [idx, irr, frozen, ESMax]
    a.index_landuse_all, a.is_irrigated, a.isFrozenSoil, a.ESMax = _p(idx, _I64), _p(irr, _U8), _p(frozen, _U8), _p(ESMax)
    a.DtDay, a.AvWaterThreshold, a.CourantCrit, a.DrainedFraction = v.DtDay, v.AvWaterThreshold, v.CourantCrit, v.DrainedFraction
    for name, typ in _SOIL_FIELDS:
        if name in ("num_vegs", "num_pixs", "index_landuse_all", "is_irrigated", "DtDay", "AvWaterThreshold", "CourantCrit",
                    "DrainedFraction", "isFrozenSoil", "ESMax", "NoSubS_out"):
            continue
        arr = getattr(v, name)
        if typ is _U8:
            arr = np.ascontiguousarray(arr).astype(np.uint8)
            keep.append(arr)
        assert arr.flags.c_contiguous and (typ is _U8 or arr.dtype == np.float64), name
        setattr(a, name, _p(arr, typ))
    a.NoSubS_out = _p(nosubs, _I64) if nosubs is not None else None
    _lib().lfo_soil_columns(C.byref(a))


class OracleModel(object):
    """One object = the reference's `self.var` restricted to the hot path; step(F) advances one model step."""

    def __init__(self, S):
        v = self.var = types.SimpleNamespace()
        n = S["N"]
        for k, val in S.items():
            setattr(v, k, val.copy() if isinstance(val, np.ndarray) else val)
        z3 = lambda: np.zeros((3, n))
        for k in ("Interception", "TaInterception", "LeafDrainage", "potential_transpiration", "Ta", "ESAct", "PrefFlow",
                  "Infiltration", "SeepTopToSubA", "SeepTopToSubB", "SeepSubToGW", "Theta", "Theta1a", "Theta1b", "Theta2",
                  "Sat1a", "Sat1b", "Sat1", "Sat2", "AvailableWaterForInfiltration", "RWS", "GwPercUZLZ", "UZOutflow"):
            setattr(v, k, z3())
        for k in ("TaCUM", "TaInterceptionCUM", "ESActCUM", "GwLossCUM", "sumDis", "DischargeM3Out"):
            setattr(v, k, np.zeros(n))
        v.TimeSinceStart = 0
        self.nosubs = np.zeros((3, n), np.int64)
        mask = S["mask"]
        mk = lambda ldd, alpha, dx, dt, a2=None: lo.KinematicWaveOracle(ldd, mask, alpha, v.Beta, dx, dt, alpha_floodplains=a2)
        # runoff order Other, Forest, Direct (surface_routing.py:108-113)
        self.of_other = mk(v.LddToChan, v.OFAlpha[0], v.PixelLength, v.DtSec)
        self.of_forest = mk(v.LddToChan, v.OFAlpha[1], v.PixelLength, v.DtSec)
        self.of_direct = mk(v.LddToChan, v.OFAlpha[2], v.PixelLength, v.DtSec)
        self.river = mk(v.LddKinematic, v.ChannelAlpha, v.ChanLength, v.DtRouting, getattr(v, "ChannelAlpha2", None))

    # -- soilloop.dynamic_canopy, soilloop.py:519-627 ------------------------------------------------
    def canopy(self):
        v = self.var
        one_minus = 1. - v.LAITerm
        ta_int_max = v.EWRef[None] * one_minus
        interception_water_balance(v.Interception, v.TaInterception, v.LeafDrainage, v.CumInterception, v.LAI, v.Rain,
                                   ta_int_max, v.LeafDrainageK)
        transpir_max = v.CropCoef * v.ETRef[None] * one_minus
        v.potential_transpiration[:] = np.maximum(transpir_max - v.TaInterception, 0)
        e = np.minimum(0.1 * v.ETRef * v.InvDtDay, 1.0)
        for k in range(3):  # vegetation index == land-use index for the three prescribed fractions (:592-627 quirk)
            cgn = v.CropGroupNumber[k]
            p = 1 / (0.76 + 1.5 * e) - 0.10 * (5 - cgn)
            p = np.where(cgn <= 2.5, p + (e - 0.6) / (cgn * (cgn + 3)), p)
            p = np.maximum(np.minimum(p, 1.0), 0)
            wc1 = ((1 - p) * (v.WFC1[k] - v.WWP1[k])) + v.WWP1[k]
            wc1a = ((1 - p) * (v.WFC1a[k] - v.WWP1a[k])) + v.WWP1a[k]
            wc1b = ((1 - p) * (v.WFC1b[k] - v.WWP1b[k])) + v.WWP1b[k]
            with np.errstate(divide="ignore", invalid="ignore"):
                rws = np.where((wc1 - v.WWP1[k]) > 0, (v.W1[k] - v.WWP1[k]) / (wc1 - v.WWP1[k]), 1)
            v.RWS[k] = np.maximum(np.minimum(rws, 1), 0)
            ta = np.minimum(v.RWS[k] * v.potential_transpiration[k], np.maximum(v.W1[k] - v.WWP1[k], 0))
            v.Ta[k] = np.where(v.isFrozenSoil, 0, ta)
            a_free = np.maximum(v.W1a[k] - wc1a, 0)
            b_free = np.maximum(v.W1b[k] - wc1b, 0)
            ta1a = np.minimum(v.Ta[k], a_free)
            rest = np.maximum(v.Ta[k] - ta1a, 0)
            ta1b = np.minimum(rest, b_free)
            rest = np.maximum(rest - ta1b, 0)
            sa = np.maximum(v.W1a[k] - ta1a - v.WWP1a[k], 0)
            sb = np.maximum(v.W1b[k] - ta1b - v.WWP1b[k], 0)
            tot = sa + sb
            ok = tot > 0
            with np.errstate(divide="ignore", invalid="ignore"):
                fa = np.where(ok, sa / tot, 0)
                fb = np.where(ok, sb / tot, 0)
            ta1a = ta1a + fa * rest
            ta1b = ta1b + fb * rest
            v.W1a[k] -= ta1a
            v.W1b[k] -= ta1b
            v.W1[k] = v.W1a[k] + v.W1b[k]

    # -- soilloop.dynamic_soil, soilloop.py:630-665 --------------------------------------------------
    def soil_columns(self):
        v = self.var
        esmax = np.ascontiguousarray(v.ESRef[None] * v.LAITerm)
        soil_columns(v, esmax, self.nosubs)

    # -- opensealed.dynamic, opensealed.py:41-71 -----------------------------------------------------
    def opensealed(self):
        v = self.var
        v.RainSnowmelt = np.maximum(v.Rain + v.SnowMelt, 0.0)
        v.EWaterAct = np.maximum(np.minimum(v.EWRef, v.RainSnowmelt) * 1.0, 0.0)
        v.InterSealed = np.minimum(np.maximum(v.SMaxSealed - v.CumInterSealed, 0.0), v.RainSnowmelt)
        v.CumInterSealed = v.CumInterSealed + v.InterSealed
        v.TASealed = np.maximum(np.minimum(v.CumInterSealed, v.EWRef), 0.0)
        v.CumInterSealed = np.maximum(v.CumInterSealed - v.TASealed, 0.0)
        v.DirectRunoff = v.DirectRunoffFraction * (v.RainSnowmelt - v.InterSealed) + v.WaterFraction * (v.RainSnowmelt - v.EWaterAct)

    def _frac(self, x):  # deffraction, Lisflood_initial.py:393-396
        return (self.var.SoilFraction * x).sum(0)

    # -- soil.dynamic_perpixel, soil.py:471-514 ------------------------------------------------------
    def perpixel(self):
        v = self.var
        v.TaInterceptionAll = self._frac(v.TaInterception) + v.DirectRunoffFraction * v.TASealed
        v.TaInterceptionCUM = v.TaInterceptionCUM + v.TaInterceptionAll
        v.TaPixel = self._frac(v.Ta)
        v.TaCUM = v.TaCUM + v.TaPixel
        v.ESActPixel = self._frac(v.ESAct) + v.WaterFraction * v.EWaterAct
        v.ESActCUM = v.ESActCUM + v.ESActPixel
        v.PrefFlowPixel = self._frac(v.PrefFlow)
        v.InfiltrationPixel = self._frac(v.Infiltration)
        tot_sm = v.W1a + v.W1b + v.W2
        v.Theta = v.SoilFraction * tot_sm / v.SoilDepthTotal
        fsum = np.sum(v.SoilFraction, 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            v.ThetaAll = np.where(fsum > 0, np.sum(v.Theta, 0) / fsum, 0)
        v.SeepTopToSubPixelA = self._frac(v.SeepTopToSubA)
        v.SeepTopToSubPixelB = self._frac(v.SeepTopToSubB)
        v.SeepSubToGWPixel = self._frac(v.SeepSubToGW)
        v.Theta1aPixel = self._frac(v.Theta1a)
        v.Theta1bPixel = self._frac(v.Theta1b)
        v.Theta2Pixel = self._frac(v.Theta2)

    # -- groundwater.dynamic, groundwater.py:134-180 ------------------------------------------------
    def groundwater(self):
        v = self.var
        v.LZOutflow = np.maximum(np.minimum(v.LowerZoneK * v.LZ, v.LZ - v.LZThreshold), 0)
        v.LZOutflowToChannel = v.LZOutflow
        v.LZ = v.LZ - v.LZOutflow
        v.UZOutflowPixel = self._frac(v.UZOutflow)
        v.GwPercUZLZPixel = self._frac(v.GwPercUZLZ)
        v.LZ = v.LZ + v.GwPercUZLZPixel
        v.GwLossLZ = np.maximum(np.minimum(v.GwLossStep, v.LZ), 0.0)
        v.LZ = v.LZ - v.GwLossLZ
        v.LZInflowCUM = np.maximum(v.LZInflowCUM + (v.GwPercUZLZPixel - v.GwLossLZ), 0.0)
        v.GwLossCUM = v.GwLossCUM + v.GwLossLZ
        v.LZAvInflow = (v.LZInflowCUM * v.InvDtDay) / v.TimeSinceStart
        v.LZOutflowToChannelPixel = v.LZOutflowToChannel

    # -- surface_routing.dynamic, surface_routing.py:115-212 ----------------------------------------
    def surface_routing(self):
        v = self.var
        v.SurfaceRunSoil = v.SoilFraction * np.maximum(v.AvailableWaterForInfiltration - v.Infiltration, 0)
        v.SurfaceRunoff = v.DirectRunoff + np.sum(v.SurfaceRunSoil, 0)
        v.TotalRunoff = v.SurfaceRunoff + v.UZOutflowPixel + v.LZOutflowToChannelPixel
        to_q = lambda mm: mm * v.MMtoM3 * v.InvPixelLength * v.InvDtSec
        side_direct = to_q(v.DirectRunoff)
        side_other = to_q(np.sum(v.SurfaceRunSoil[[0, 2]], 0))   # Rainfed + Irrigated
        side_forest = to_q(v.SurfaceRunSoil[1])
        self.of_direct.kinematicWaveRouting(v.OFQDirect, side_direct)
        self.of_other.kinematicWaveRouting(v.OFQOther, side_other)
        self.of_forest.kinematicWaveRouting(v.OFQForest, side_forest)
        v.OFM3Direct = v.PixelLength * v.OFAlpha[2] * v.OFQDirect ** v.Beta
        v.OFM3Other = v.PixelLength * v.OFAlpha[0] * v.OFQOther ** v.Beta
        v.OFM3Forest = v.PixelLength * v.OFAlpha[1] * v.OFQForest ** v.Beta
        v.Qall = v.OFQDirect + v.OFQOther + v.OFQForest
        v.M3all = v.OFM3Direct + v.OFM3Other + v.OFM3Forest
        v.OFToChanM3 = np.where(v.IsChannel, v.Qall * v.DtSec, 0)
        v.WaterDepth = v.M3all * v.M3toMM
        v.ToChanM3Runoff = (v.UZOutflowPixel + v.LZOutflowToChannelPixel) * v.MMtoM3 + v.OFToChanM3
        v.ToChanM3RunoffDt = v.ToChanM3Runoff * v.InvNoRoutSteps

    # -- lakes.dynamic_inloop, lakes.py:199-297 (Modified Puls; per-lake "CC" arrays) ---------------------
    def lakes_inloop(self, s):
        v = self.var
        if s == 0:
            v.LakeStorageM3CC = np.compress(v.LakeSitesC2 > 0, v.LakeStorageM3)
        v.LakeInflowCC = np.bincount(v.downstruct, weights=v.ChanQ)[v.LakeIndex]
        lake_in = (v.LakeInflowCC + v.LakeInflowOldCC) * 0.5
        v.LakeInflowOldCC = v.LakeInflowCC.copy()
        indicator = v.LakeStorageM3CC / v.DtRouting - 0.5 * v.LakeOutflowCC + lake_in
        v.LakeOutflowCC = np.square(-v.LakeFactor + np.sqrt(v.LakeFactorSqr + 2 * indicator))
        out_m3 = v.LakeOutflowCC * v.DtRouting
        v.LakeStorageM3CC = (indicator - v.LakeOutflowCC * 0.5) * v.DtRouting
        bad = np.isnan(v.LakeStorageM3CC) | (v.LakeStorageM3CC < 0)
        v.LakeStorageM3CC[bad] = 0
        v.LakeStorageM3BalanceCC = v.LakeStorageM3BalanceCC + (lake_in * v.DtRouting - out_m3)
        v.LakeLevelCC = v.LakeStorageM3CC / v.LakeAreaCC
        v.QLakeOutM3Dt = np.zeros(v.N)
        np.put(v.QLakeOutM3Dt, v.LakeIndex, out_m3)
        if s == v.NoRoutSteps - 1:
            for name in ("LakeStorageM3Balance", "LakeStorageM3", "LakeLevel", "LakeInflowOld", "LakeOutflow"):
                full = np.zeros(v.N)
                np.put(full, v.LakeIndex, getattr(v, name + "CC"))
                setattr(v, name, full)

    # -- reservoir.dynamic_inloop, reservoir.py:173-322 (four-regime outflow rule) -------------------------
    def reservoir_inloop(self, s):
        v = self.var
        inv_day = 1 / float(86400)
        inflow = np.bincount(v.downstruct, weights=v.ChanQ)[v.ReservoirIndex]
        in_m3 = inflow * v.DtRouting
        if s == 0:
            v.ReservoirStorageM3CC = np.compress(v.ReservoirSitesC > 0, v.ReservoirStorageM3)
        v.ReservoirStorageM3CC = v.ReservoirStorageM3CC + in_m3
        fill = v.ReservoirStorageM3CC / v.TotalReservoirStorageM3CC
        o1 = np.minimum(v.MinReservoirOutflowCC, v.ReservoirStorageM3CC * inv_day)
        o2 = v.MinReservoirOutflowCC + v.DeltaO * (fill - 2 * v.ConservativeStorageLimitCC) / v.DeltaLN
        o3a = v.NormalReservoirOutflowCC
        o3b = v.NormalReservoirOutflowCC + ((fill - v.Normal_FloodStorageLimitCC) / v.DeltaNFL) * (
            v.NonDamagingReservoirOutflowCC - v.NormalReservoirOutflowCC)
        temp = np.minimum(v.NonDamagingReservoirOutflowCC, np.maximum(inflow * 1.2, v.NormalReservoirOutflowCC))
        o4 = np.maximum((fill - v.FloodStorageLimitCC - 0.01) * v.TotalReservoirStorageM3CC * inv_day, temp)
        out = o1.copy()
        out = np.where(fill > 2 * v.ConservativeStorageLimitCC, o2, out)
        out = np.where(fill > v.NormalStorageLimitCC, o3a, out)
        out = np.where(fill > v.Normal_FloodStorageLimitCC, o3b, out)
        out = np.where(fill > v.FloodStorageLimitCC, o4, out)
        temp = np.minimum(out, np.maximum(inflow, v.NormalReservoirOutflowCC))
        out = np.where((out > 1.2 * inflow) & (out > v.NormalReservoirOutflowCC) & (fill < v.FloodStorageLimitCC), temp, out)
        out_m3 = out * v.DtRouting
        out_m3 = np.minimum(out_m3, v.ReservoirStorageM3CC)
        out_m3 = np.maximum(out_m3, v.ReservoirStorageM3CC - v.TotalReservoirStorageM3CC)
        v.ReservoirStorageM3CC = v.ReservoirStorageM3CC - out_m3
        v.ReservoirFillCC = v.ReservoirStorageM3CC / v.TotalReservoirStorageM3CC
        v.ReservoirFillCC[np.isnan(v.ReservoirFillCC)] = 0
        v.ReservoirFillCC[v.ReservoirFillCC < 0] = 0
        v.QResOutM3Dt = np.zeros(v.N)
        np.put(v.QResOutM3Dt, v.ReservoirIndex, out_m3)
        if s == v.NoRoutSteps - 1:
            v.ReservoirStorageM3 = np.zeros(v.N)
            v.ReservoirFill = np.zeros(v.N)
            np.put(v.ReservoirStorageM3, v.ReservoirIndex, v.ReservoirStorageM3CC)
            np.put(v.ReservoirFill, v.ReservoirIndex, v.ReservoirFillCC)

    # -- routing.dynamic, routing.py:435-706 (kinematic wave; structures = lakes + reservoirs when present) ----
    def routing_substep(self, s=0):
        v = self.var
        lakes_on, res_on = bool(getattr(v, "simulateLakes", False)), bool(getattr(v, "simulateReservoirs", False))
        if lakes_on:
            self.lakes_inloop(s)          # :441-443: lakes first, then reservoirs, both on the previous ChanQ
        if res_on:
            self.reservoir_inloop(s)
        side_m3 = v.ToChanM3RunoffDt.copy()
        if lakes_on:
            side_m3 += v.QLakeOutM3Dt     # :472-475
        if res_on:
            side_m3 += v.QResOutM3Dt
        side = np.where(v.IsChannelKinematic, side_m3 * v.InvChanLength * v.InvDtRouting, 0)
        if not v.SplitRouting:
            side[np.isnan(side)] = 0
            self.river.kinematicWaveRouting(v.ChanQKin, side, "main_channel")
            v.ChanM3Kin = np.maximum(v.ChanLength * v.ChannelAlpha * v.ChanQKin ** v.Beta, 0.0)
            v.ChanQKin = (v.ChanM3Kin * v.InvChanLength * v.InvChannelAlpha) ** v.InvBeta
            v.ChanQ = v.ChanQKin.copy()
            v.sumDisDay = v.sumDisDay + v.ChanQ
        else:
            tot = v.ChanM3Kin + v.Chan2M3Kin
            with np.errstate(divide="ignore", invalid="ignore"):
                ratio = np.where(tot > 0, v.ChanM3Kin / tot, 0.0)
            s1 = np.where((tot - v.Chan2M3Start) > v.M3Limit, ratio * side, side)
            v.Sideflow1Chan = np.where(np.abs(side) < 1e-7, side, s1)
            s2 = side - v.Sideflow1Chan
            s2 = s2 + v.Chan2QStart * v.InvChanLength
            self.river.kinematicWaveRouting(v.ChanQKin, v.Sideflow1Chan, "main_channel")
            v.ChanM3Kin = np.maximum(v.ChanLength * v.ChannelAlpha * v.ChanQKin ** v.Beta, 0.0)
            v.ChanQKin = (v.ChanM3Kin * v.InvChanLength * v.InvChannelAlpha) ** v.InvBeta
            self.river.kinematicWaveRouting(v.Chan2QKin, s2, "floodplains")
            m3 = v.ChanLength * v.ChannelAlpha2 * v.Chan2QKin ** v.Beta
            v.Chan2M3Kin = np.where(m3 - v.Chan2M3Start < 0.0, v.Chan2M3Start, m3)
            v.CrossSection2Area = (v.Chan2M3Kin - v.Chan2M3Start) * v.InvChanLength
            v.Chan2QKin = (v.Chan2M3Kin * v.InvChanLength * v.InvChannelAlpha2) ** v.InvBeta
            v.ChanQ = np.maximum(v.ChanQKin + v.Chan2QKin - v.QLimit, 0.0)
            v.sumDisDay_NOTlast = v.sumDisDay.copy()
            v.sumDisDay = v.sumDisDay + v.ChanQ
        area = np.maximum(v.ChanM3Kin * v.InvChanLength, 0.01)
        v.FlowVelocity = np.minimum(v.ChanQKin / area, 0.36 * v.ChanQKin ** 0.24)
        v.FlowVelocity = v.FlowVelocity * np.minimum(np.sqrt(v.PixelArea) * v.InvChanLength, 1)
        v.TravelDistance = v.FlowVelocity * v.DtSec

    # -- Lisflood_dynamic.py:114-229 ---------------------------------------------------------------------
    def set_forcing(self, F):
        v = self.var
        for k in ("Rain", "SnowMelt", "ETRef", "EWRef", "ESRef", "LAI", "LAITerm", "isFrozenSoil"):
            setattr(v, k, F[k].copy())

    def soil_step(self, F):
        self.var.TimeSinceStart += 1
        self.set_forcing(F)
        self.canopy()
        self.soil_columns()
        self.opensealed()
        self.perpixel()
        self.groundwater()

    def routing_step(self):
        v = self.var
        self.surface_routing()
        v.sumDisDay = np.zeros(v.N)
        for s in range(v.NoRoutSteps):
            self.routing_substep(s)
        v.ChanM3 = v.ChanM3Kin.copy() if not v.SplitRouting else v.ChanM3Kin + v.Chan2M3Kin - v.Chan2M3Start
        v.TotalCrossSectionArea = v.ChanM3 * v.InvChanLength
        v.sumDis = v.sumDis + v.sumDisDay
        v.ChanQAvg = v.sumDisDay / v.NoRoutSteps
        v.DischargeM3Out = v.DischargeM3Out + np.where(v.AtLastPointC, v.ChanQ * v.DtSec, 0)

    def step(self, F):
        self.soil_step(F)
        self.routing_step()
