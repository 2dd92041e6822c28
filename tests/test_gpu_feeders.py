"""Feeder modules on the device (readmeteo scaling + snow + frost fused, LAITerm) against goldens made
by the reference's OWN snow / frost classes (tests/golden/make_golden.py::feeders_case), through the C ABI; float32 raw
forcing (as in the NetCDF files), map and scalar parameters, synchronous and asynchronous upload."""
import numpy as np
import pytest

from conftest import golden_cases, rel_err
from test_oracle_feeders_golden import feeder_case

pytestmark = pytest.mark.gpu


def _model(n, dt_sec, diagnostics=True):
    """A model whose drainage network is irrelevant here: n pixels in one row, all pits."""
    from lisflood_code_b200.hotpath import HotPathModel
    S = {"mask": np.ones((1, n), bool), "LddToChan": np.full(n, 5.0), "LddKinematic": np.full(n, 5.0), "DtSec": dt_sec,
         "Beta": 0.6, "PixelLength": 5000.0, "NoRoutSteps": 1, "SplitRouting": False, "CourantCrit": 0.4,
         "AvWaterThreshold": 5.0, "LeafDrainageK": 1.0, "DrainedFraction": 0.0, "SMaxSealed": 1.0}
    return HotPathModel(S, diagnostics=diagnostics)


@pytest.mark.parametrize("case", golden_cases("feeders_"))
@pytest.mark.parametrize("asynchronous", [False, True])
def test_feeders_golden(gpu_lib, case, asynchronous):
    P, S0, dt, R, days, O = feeder_case(case)
    n = S0["FrostIndex"].size
    M = _model(n, dt)
    M.set_feeder(P, S0)
    for t in range(len(R)):
        M.feed(R[t], days[t], asynchronous=asynchronous)
        gpu_lib.synchronize()
        for k, want in O[t].items():
            if k == "isFrozenSoil":
                continue
            got = M.get(k, 3 if want.ndim == 2 else 1)
            assert rel_err(got, want) < 1e-12, (case, t, k, rel_err(got, want))
        # the frozen-soil flag as the soil stage sees it: FrostIndex > threshold (bit-equal decisions away from ties)
        fi = M.get("FrostIndex")
        assert np.array_equal(fi > P["FrostIndexThreshold"], O[t]["isFrozenSoil"])


def test_feeders_scalar_parameters_and_lai(gpu_lib):
    from oracle.lisf_oracle_feeders import FeederOracle, lai_term
    rng = np.random.default_rng(5)
    n = 777
    P = {"PrScaling": 1.0, "CalEvaporation": 1.1, "DeltaTSnow": 0.9674 * 120.0 * 0.0065, "SnowSeason": 0.5, "TempSnow": 1.0,
         "SnowFactor": 1.0, "SnowMeltCoef": 4.0, "TempMelt": 0.0, "lat_rad": np.radians(rng.uniform(-40, 60, n)), "Kfrost": 0.57,
         "Afrost": 0.97, "FrostIndexThreshold": 56.0, "SnowWaterEquivalent": 0.45, "kgb": 0.75 * 0.72}
    S0 = {"SnowCoverS": rng.uniform(0, 50, (3, n)), "FrostIndex": rng.uniform(0, 60, n)}
    M = _model(n, 86400.0)
    M.set_feeder(P, S0)
    Pm = {k: (np.full(n, v) if np.ndim(v) == 0 else v) for k, v in P.items() if k != "kgb"}
    F = FeederOracle(Pm, S0, 86400.0)
    for t, day in enumerate((10, 200, 300, 360)):
        raw = {"Precipitation": rng.gamma(0.8, 8.0, n), "Tavg": rng.uniform(-20, 20, n), "ET0": rng.uniform(0, 6, n),
               "E0": rng.uniform(0, 7, n)}                       # float64 raw maps this time
        want = F.step(raw, day)
        M.feed(raw, day)
        for k in ("Rain", "SnowMelt", "ETRef", "EWRef", "ESRef", "SnowCoverS", "FrostIndex", "Snow", "SnowCover"):
            assert rel_err(M.get(k, 3 if k == "SnowCoverS" else 1), want[k]) < 1e-12, (t, k)
    lai = rng.uniform(0, 6, (3, n))
    M.set_lai(lai)
    assert rel_err(M.get("LAITerm", 3), lai_term(P["kgb"], lai)) < 1e-14
    assert np.array_equal(M.get("LAI", 3), lai)


def test_dynamic_with_raw_forcing_equals_feed_and_step(gpu_lib):
    """LisfloodModel_dyn.dynamic(raw=...) drives readmeteo / leafarea / snow / frost mirrors in the reference's order
    (Lisflood_dynamic.py:79-105) and gives what feed() + step() give."""
    from lisflood_code_b200 import synthetic
    from lisflood_code_b200.hotpath import HotPathModel
    from lisflood_code_b200.Lisflood_dynamic import LisfloodModel_dyn
    S = synthetic.full_stack(50, 60, seed=31, mask_fraction=0.1)
    n = S["N"]
    rng = np.random.default_rng(8)
    P = {"PrScaling": 1.0, "CalEvaporation": 1.0, "DeltaTSnow": rng.uniform(0, 2, n), "SnowSeason": 0.5, "TempSnow": 1.0,
         "SnowFactor": 1.0, "SnowMeltCoef": 4.0, "TempMelt": 0.0, "lat_rad": np.full(n, 0.8), "Kfrost": 0.57, "Afrost": 0.97,
         "FrostIndexThreshold": 56.0, "SnowWaterEquivalent": 0.45, "kgb": 0.75 * 0.72}
    lai = {j: rng.uniform(0, 6, (3, n)) for j in range(36)}
    A, B = HotPathModel(S), HotPathModel(S)
    for M in (A, B):
        M.set_feeder(P, {"SnowCoverS": np.zeros((3, n)), "FrostIndex": np.zeros(n)})
    dyn = LisfloodModel_dyn(B)
    from lisflood_code_b200.hydrological_modules.snow import lai_interval
    last = None
    for t, day in enumerate((9, 10, 11, 12)):       # day 11 starts a new LAI interval
        raw = {"Precipitation": rng.gamma(0.8, 8.0, n).astype(np.float32), "Tavg": rng.uniform(-8, 20, n).astype(np.float32),
               "ET0": rng.uniform(0, 6, n).astype(np.float32), "E0": rng.uniform(0, 6, n).astype(np.float32)}
        j = lai_interval(day)
        if j != last:
            A.set_lai(lai[j])
            last = j
        A.feed(raw, day)
        A.step()
        dyn.dynamic(raw=raw, calendar_day=day, lai_of_interval=lambda k: lai[k])
        for k in ("ChanQAvg", "W1a", "FrostIndex", "LAITerm"):
            rows = 3 if k in ("W1a", "LAITerm") else 1
            assert np.array_equal(A.get(k, rows), B.get(k, rows)), (t, k)
    with pytest.raises(RuntimeError):
        dyn.frost_module.dynamic()              # out of order: no snow.dynamic before it


@pytest.mark.parametrize("decode", ["float32", "float64"])
@pytest.mark.parametrize("asynchronous", [False, True])
def test_packed_forcing_is_unpacked_on_the_device(gpu_lib, decode, asynchronous):
    """int16 forcing with scale_factor / add_offset (CF packing, decoded on the host by the reference's reader:
    netcdf.py:231-232) gives bit for bit what feeding the unpacked maps gives, and that agrees with the feeder oracle."""
    from oracle.lisf_oracle_feeders import FeederOracle, cf_pack, cf_unpack
    rng = np.random.default_rng(17)
    n = 4099
    P = {"PrScaling": 1.0, "CalEvaporation": 1.05, "DeltaTSnow": rng.uniform(0, 2, n), "SnowSeason": 0.5, "TempSnow": 1.0,
         "SnowFactor": 1.0, "SnowMeltCoef": 4.0, "TempMelt": 0.0, "lat_rad": np.radians(rng.uniform(-40, 60, n)), "Kfrost": 0.57,
         "Afrost": 0.97, "FrostIndexThreshold": 56.0, "SnowWaterEquivalent": 0.45, "kgb": 0.75 * 0.72}
    S0 = {"SnowCoverS": rng.uniform(0, 50, (3, n)), "FrostIndex": rng.uniform(0, 60, n)}
    A, B = _model(n, 86400.0), _model(n, 86400.0)
    Pm = {k: (np.full(n, v) if np.ndim(v) == 0 else v) for k, v in P.items() if k != "kgb"}
    F = FeederOracle(Pm, S0, 86400.0)
    for M in (A, B):
        M.set_feeder(P, S0)
    keys = ("Rain", "SnowMelt", "ETRef", "EWRef", "ESRef", "SnowCoverS", "FrostIndex", "Snow", "SnowCover", "Precipitation", "Tavg")
    for t, day in enumerate((15, 120, 260, 350)):
        fields = {"Precipitation": rng.gamma(0.8, 8.0, n), "Tavg": rng.uniform(-25, 25, n), "ET0": rng.uniform(0, 6, n),
                  "E0": rng.uniform(0, 7, n)}
        packed, packing, unpacked = {}, {}, {}
        for k, v in fields.items():
            packed[k], s, o = cf_pack(v)
            packing[k] = (s, o)
            unpacked[k] = cf_unpack(packed[k], s, o, decode)
            assert unpacked[k].dtype == (np.float32 if decode == "float32" else np.float64)
            assert np.abs(unpacked[k] - v).max() <= 0.51 * s * (1 + 1e-3) + 1e-6      # it is the packing of these maps
        A.feed(packed, day, asynchronous=asynchronous, packing=packing, decode=decode)
        B.feed(unpacked, day)
        gpu_lib.synchronize()
        want = F.step({k: v.astype(np.float64) for k, v in unpacked.items()}, day)
        for k in keys:
            rows = 3 if k == "SnowCoverS" else 1
            a = A.get(k, rows)
            assert np.array_equal(a, B.get(k, rows)), (t, k)
            if k in want:
                assert rel_err(a, want[k]) < 1e-12, (t, k)
    with pytest.raises(ValueError):
        A.feed(packed, 1, packing=packing, decode="float16")
    with pytest.raises(ValueError):
        A.feed(unpacked, 1, packing=packing)          # not int16
    with pytest.raises(ValueError):
        A.feed(packed, 1)                             # int16 without its packing attributes


def test_get_async_float32_is_the_narrowed_map(gpu_lib):
    """OutputMapsDataType = float32 (netcdf.py:478): the map is narrowed on the device, equal to get().astype(float32)."""
    from lisflood_code_b200 import _capi, synthetic
    from lisflood_code_b200.hotpath import HotPathModel
    S = synthetic.full_stack(40, 50, seed=3, mask_fraction=0.1)
    M = HotPathModel(S)
    out32 = _capi.pinned_empty(S["N"], np.float32)
    kin32 = _capi.pinned_empty(S["N"], np.float32)
    w32 = _capi.pinned_empty(3 * S["N"], np.float32)
    for t in range(3):
        M.set_forcing(synthetic.forcing(S, t, seed=4))
        M.step()
        M.get_async("ChanQAvg", out32)                # a map in the channel order
        M.get_async("ChanQKin", kin32)                # stored as Q^(1/5) on the device (Beta = 0.6)
        M.get_async("W1a", w32)                       # a (3, N) map in the soil order
        M.wait_outputs()
        assert np.array_equal(out32, M.get("ChanQAvg").astype(np.float32))
        assert np.array_equal(kin32, M.get("ChanQKin").astype(np.float32)) and float(kin32.max()) > 0
        assert np.array_equal(np.asarray(w32).reshape(3, -1), M.get("W1a", 3).astype(np.float32))
    with pytest.raises(TypeError):
        M.get_async("ChanQAvg", np.empty(S["N"], np.int32))
