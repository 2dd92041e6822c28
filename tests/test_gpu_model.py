"""GPU parity of the full hot-path step (fused soil kernel, overland routers, channel sub-step wavefront)
through the C ABI, against golden vectors made by the reference's own module classes and against the CPU
oracle on larger seeded catchments.  Tolerance: 1e-6 relative contractually (abs floor 1e-12); asserted 1e-8."""
import numpy as np
import pytest

from conftest import golden_cases, golden_model, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _model(gpu_lib, S, diagnostics=True):
    from lisflood_code_b200.hotpath import HotPathModel
    return HotPathModel(S, diagnostics=diagnostics)


@pytest.mark.parametrize("case", golden_cases("model_"))
def test_golden_model_step(gpu_lib, case):
    S, F, O = golden_model(case)
    M = _model(gpu_lib, S)
    for t in range(len(F)):
        M.step(F[t])
        worst = {}
        for k, want in O[t].items():
            got = np.asarray(M.get(k, 3 if want.ndim == 2 else 1))
            worst[k] = rel_err(got, want)
        bad = {k: v for k, v in worst.items() if not v < TOL}
        assert not bad, (case, t, bad)


@pytest.mark.parametrize("rows,cols,split,seed,beta", [(150, 120, False, 7, 0.6), (130, 160, True, 8, 0.6),
                                                       (90, 110, True, 9, 0.7)])   # beta != 3/5: general-power path
def test_model_vs_oracle(gpu_lib, oracle, rows, cols, split, seed, beta):
    from lisflood_code_b200 import synthetic
    from oracle import lisf_oracle_model as om
    S = synthetic.full_stack(rows, cols, seed=seed, split_routing=split, ldd_noise=0.4, mask_fraction=0.1, beta=beta)
    O = om.OracleModel(S)
    M = _model(gpu_lib, S)
    keys = ["W1a", "W1b", "W2", "UZ", "LZ", "DSLR", "CumInterception", "Infiltration", "PrefFlow", "ESAct", "Ta",
            "TotalRunoff", "OFQOther", "OFQForest", "OFQDirect", "ToChanM3RunoffDt", "ChanQKin", "ChanM3Kin", "ChanQ",
            "ChanQAvg", "sumDis", "DischargeM3Out", "TotalCrossSectionArea", "ThetaAll", "FlowVelocity"]
    if split:
        keys += ["Chan2QKin", "Chan2M3Kin", "CrossSection2Area", "Sideflow1Chan"]
    for t in range(4):
        F = synthetic.forcing(S, t, seed)
        O.step(F)
        M.step(F)
        bad = {}
        for k in keys:
            want = np.asarray(getattr(O.var, k))
            e = rel_err(M.get(k, 3 if want.ndim == 2 else 1), want)
            if not e < TOL:
                bad[k] = e
        assert not bad, (t, bad)
    # sub-stepping actually happened somewhere (the adaptive Courant loop is exercised)
    assert O.nosubs.max() > 1


def test_lean_mode_matches_diagnostic_mode(gpu_lib):
    """diagnostics=False (the production configuration) computes the same state and discharge."""
    from lisflood_code_b200 import synthetic
    S = synthetic.full_stack(90, 70, seed=11, split_routing=True)
    A, B = _model(gpu_lib, S, True), _model(gpu_lib, S, False)
    for t in range(2):
        F = synthetic.forcing(S, t, 11)
        A.step(F)
        B.step(F)
    for k in ("W1a", "W1b", "W2", "UZ", "LZ", "ChanQAvg", "ChanQKin", "Chan2QKin", "OFQOther"):
        rows = 3 if k in ("W1a", "W1b", "W2", "UZ") else 1
        assert np.array_equal(A.get(k, rows), B.get(k, rows)), k


def test_stage_calls_equal_fused_step(gpu_lib):
    from lisflood_code_b200 import synthetic
    S = synthetic.full_stack(60, 80, seed=12)
    A, B = _model(gpu_lib, S, False), _model(gpu_lib, S, False)
    F = synthetic.forcing(S, 0, 12)
    A.step(F)
    B.set_forcing(F)
    B.soil()
    B.surface_routing()
    B.channel()
    assert np.array_equal(A.get("ChanQAvg"), B.get("ChanQAvg"))
    with pytest.raises(AttributeError):
        A.NoSuchMap


def test_async_input_path_equals_sync(gpu_lib):
    """lf_model_set_async (copy stream + deferred layout translation) feeds the same values as lf_model_set."""
    from lisflood_code_b200 import synthetic
    S = synthetic.full_stack(70, 90, seed=13)
    A, B = _model(gpu_lib, S, False), _model(gpu_lib, S, False)
    for t in range(3):
        F = synthetic.forcing(S, t, 13)
        A.step(F)
        keep = {k: np.ascontiguousarray(F[k], np.float64) for k in ("Rain", "SnowMelt", "ETRef", "EWRef", "ESRef", "LAI",
                                                                    "LAITerm")}
        for k, a in keep.items():
            B.set_async(k, a)
        B.set_flags("isFrozenSoil", F["isFrozenSoil"])
        B.step()
        assert np.array_equal(A.get("ChanQAvg"), B.get("ChanQAvg")), t
    st = A.soil_stats(enable_timing=False)
    assert sum(st["deferred_columns"]) > 0 and 0 < st["deferred_fraction"] < 0.5


LEAN_STATE = ("W1a", "W1b", "W2", "UZ", "DSLR", "CumInterception")
LEAN_PIXEL = ("LZ", "CumInterSealed", "LZInflowCUM", "TaCUM", "TaInterceptionCUM", "ESActCUM", "GwLossCUM", "ChanQAvg",
              "ChanQKin", "OFQOther", "OFQForest", "OFQDirect")


def _assert_same_state(A, B, tag):
    for k in LEAN_STATE:
        assert np.array_equal(A.get(k, 3), B.get(k, 3)), (tag, k)
    for k in LEAN_PIXEL:
        assert np.array_equal(A.get(k), B.get(k)), (tag, k)


@pytest.mark.parametrize("rows,cols,maskf", [(96, 64, 0.0), (61, 71, 0.0), (80, 90, 0.15)])
@pytest.mark.parametrize("variant,plain", [(13, 0), (13, 1), (10, 0), (11, 0), (12, 0)])
def test_first_pass_launch_shapes_agree(gpu_lib, monkeypatch, rows, cols, maskf, variant, plain):
    """The lean first pass of the soil stage -- inputs staged in shared memory by bulk async copies (variants 10-13),
    or staged by ordinary loads (LF_SOIL_PLAIN, and automatically for ragged tiles / odd N) -- produces the bits of the
    diagnostics build (k_soil_fused, direct loads), which the golden tests pin."""
    from lisflood_code_b200 import synthetic
    S = synthetic.full_stack(rows, cols, seed=21, split_routing=False, mask_fraction=maskf)
    A = _model(gpu_lib, S, True)
    monkeypatch.setenv("LF_SOIL_VARIANT", str(variant))
    monkeypatch.setenv("LF_SOIL_PLAIN", str(plain))
    B = _model(gpu_lib, S, False)
    for t in range(3):
        F = synthetic.forcing(S, t, 21)
        A.step(F)
        B.step(F)
        _assert_same_state(A, B, (variant, plain, t))
    st = B.soil_stats(enable_timing=False)
    assert sum(st["deferred_columns"]) > 0


def test_full_bucket_lists_fall_back_to_the_pixel_kernel(gpu_lib, monkeypatch):
    """A column that finds its bucket list full is integrated by k_soil_pixel_flagged: same bits."""
    from lisflood_code_b200 import synthetic
    S = synthetic.full_stack(64, 64, seed=22, split_routing=True)
    A = _model(gpu_lib, S, False)
    monkeypatch.setenv("LF_SOIL_LIST_CAP", "3")
    B = _model(gpu_lib, S, False)
    for t in range(3):
        F = synthetic.forcing(S, t, 22)
        A.step(F)
        B.step(F)
        _assert_same_state(A, B, t)
    assert sum(A.soil_stats(enable_timing=False)["deferred_columns"]) > 6 * 3


@pytest.mark.parametrize("rows,cols,maskf", [(128, 96, 0.0), (75, 83, 0.2)])
def test_lean_model_vs_oracle(gpu_lib, oracle, rows, cols, maskf):
    """The production (lean) configuration against the CPU oracle, bulk-staged (even N, full tiles) and plain-staged."""
    from lisflood_code_b200 import synthetic
    from oracle import lisf_oracle_model as om
    S = synthetic.full_stack(rows, cols, seed=23, split_routing=True, ldd_noise=0.4, mask_fraction=maskf)
    O = om.OracleModel(S)
    M = _model(gpu_lib, S, False)
    for t in range(3):
        F = synthetic.forcing(S, t, 23)
        O.step(F)
        M.step(F)
        bad = {}
        for k in LEAN_STATE + ("LZ", "TaCUM", "ESActCUM", "ChanQAvg", "ChanQKin", "Chan2QKin", "OFQOther", "OFQForest"):
            want = np.asarray(getattr(O.var, k))
            e = rel_err(M.get(k, 3 if want.ndim == 2 else 1), want)
            if not e < TOL:
                bad[k] = e
        assert not bad, (t, bad)


@pytest.mark.parametrize("split", [False, True])
def test_narrow_runs_equal_one_launch_per_diagonal(gpu_lib, split):
    """Channel wavefront: runs of narrow diagonals in one single-block launch (k_chan_narrow_run) against one launch per
    diagonal, bit for bit, on a network deep enough to have both kinds."""
    from lisflood_code_b200 import synthetic
    S = synthetic.full_stack(420, 380, seed=12, split_routing=split, ldd_noise=0.3, mask_fraction=0.05)
    got = {}
    for narrow in (1, 0):
        M = _model(gpu_lib, S)
        M.set_option("narrow_runs", narrow)
        for t in range(3):
            M.step(synthetic.forcing(S, t, 12))
        got[narrow] = {k: M.get(k) for k in ("ChanQAvg", "ChanQ", "ChanM3", "ChanQKin", "sumDis")}
    for k in got[1]:
        assert np.array_equal(got[1][k], got[0][k]), k
    assert float(got[1]["ChanQAvg"].max()) > 0


def test_early_isolated_launch_equals_serial(gpu_lib):
    """"overlap_isolated": the non-channel isolated pixels started beside the soil stage give the same maps, bit for bit."""
    from lisflood_code_b200 import synthetic
    S = synthetic.full_stack(300, 260, seed=14, ldd_noise=0.5, mask_fraction=0.05)
    got = {}
    for overlap in (0, 1):
        M = _model(gpu_lib, S)
        M.set_option("overlap_isolated", overlap)
        for t in range(3):
            M.step(synthetic.forcing(S, t, 14))
        got[overlap] = {k: M.get(k) for k in ("ChanQAvg", "ChanQ", "ChanM3", "ChanQKin", "sumDis", "LZ")}
    for k in got[1]:
        assert np.array_equal(got[1][k], got[0][k]), k
