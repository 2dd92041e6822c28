"""The reference's two Numba kernels as stand-alone device operators: same argument lists, in place,
compared with the CPU oracle (which the reference-made goldens pin).  Tolerance: 1e-6 relative contractually (abs floor
1e-12); asserted 1e-9."""
import copy

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _prepared(seed, rows=70, cols=90, steps=2):
    """Model object after `steps` full steps plus the canopy part of the next one (the state soilColumnsWaterBalance sees)."""
    from lisflood_code_b200 import synthetic
    from oracle import lisf_oracle_model as om
    S = synthetic.full_stack(rows, cols, seed=seed, split_routing=False, mask_fraction=0.1)
    O = om.OracleModel(S)
    for t in range(steps):
        O.step(synthetic.forcing(S, t, seed))
    O.set_forcing(synthetic.forcing(S, steps, seed))
    return S, O


def test_interception_operator(gpu_lib, oracle):
    from lisflood_code_b200.hydrological_modules.soilloop import interception_water_balance
    from oracle import lisf_oracle_model as om
    rng = np.random.default_rng(5)
    V, N = 3, 4099
    lai = np.ascontiguousarray(rng.uniform(0, 7, (V, N)))
    lai[0, :50] = 0.05          # SMax = 0 branch
    lai[1, :20] = 50.0          # LAI > 43.3 branch
    rain = np.where(rng.random(N) < 0.5, rng.gamma(0.8, 8.0, N), 0.0)
    tmax = np.ascontiguousarray(rng.uniform(0, 4, (V, N)))
    cum0 = np.ascontiguousarray(rng.uniform(0, 1.5, (V, N)) * (rng.random((V, N)) < 0.8))
    A = [np.zeros((V, N)), np.zeros((V, N)), np.zeros((V, N)), cum0.copy()]
    B = [np.zeros((V, N)), np.zeros((V, N)), np.zeros((V, N)), cum0.copy()]
    om.interception_water_balance(A[0], A[1], A[2], A[3], lai, rain, tmax, 0.7)
    assert interception_water_balance(B[0], B[1], B[2], B[3], lai, rain, tmax, 0.7) is None
    for a, b, nm in zip(A, B, ("Interception", "TaInterception", "LeafDrainage", "CumInterception")):
        assert rel_err(b, a) < TOL, nm
    with pytest.raises(TypeError):
        interception_water_balance(B[0].astype(np.float32), B[1], B[2], B[3], lai, rain, tmax, 0.7)


@pytest.mark.parametrize("seed,drained", [(31, 0.0), (32, 0.3)])
def test_soil_columns_operator(gpu_lib, oracle, seed, drained):
    from lisflood_code_b200.hydrological_modules import soilloop
    from oracle import lisf_oracle_model as om
    S, O = _prepared(seed)
    O.var.DrainedFraction = drained
    O.canopy()
    v = O.var
    esmax = np.ascontiguousarray(v.ESRef[None] * v.LAITerm)
    g = copy.deepcopy(v)                       # the operator's copy of every array
    om.soil_columns(v, esmax, O.nosubs)        # oracle, in place on v
    args = {}
    for name in soilloop._SOIL_ARG_ORDER:
        if name == "index_landuse_all":
            args[name] = np.arange(3)
        elif name == "is_irrigated":
            args[name] = np.array([False, False, True])
        elif name == "is_paddy_irrig":
            args[name] = np.zeros(3, bool)
        elif name == "paddy_inactive":
            args[name] = []
        elif name == "ESMax":
            args[name] = esmax
        else:
            args[name] = getattr(g, name)
    assert soilloop.soilColumnsWaterBalance(*[args[k] for k in soilloop._SOIL_ARG_ORDER]) is None
    bad = {}
    for name in soilloop._SOIL_IN_PLACE:
        e = rel_err(getattr(g, name), getattr(v, name))
        if not e < TOL:
            bad[name] = e
    assert not bad, bad
    assert O.nosubs.max() > 1      # the adaptive sub-stepping was exercised
    with pytest.raises(Exception):
        a2 = dict(args, is_paddy_irrig=np.array([False, False, True]))
        soilloop.soilColumnsWaterBalance(**a2)


def test_suction_pf_operator(gpu_lib):
    """suctionUnsaturatedSoilPF (option simulatePF) on the device: same 26 arguments as the reference's kernel; against the
    golden made by the reference's own soilloop class, and the edge cases of saturationDegree / pressureHead."""
    from conftest import golden_cases, load_golden
    from lisflood_code_b200.hydrological_modules.soilloop import suctionUnsaturatedSoilPF
    from test_oracle_soil_options_golden import LAYERS, edge_case_arguments, pf_arguments, split_golden
    cases = golden_cases("soilopt_")
    assert cases
    for case in cases:
        g = load_golden(case)
        S, X = split_golden(g)
        for t in range(int(g["steps"])):
            W = [np.ascontiguousarray(g["O%d__W%s" % (t, lay)]) for lay in LAYERS]
            a = pf_arguments(S, X, W)
            suctionUnsaturatedSoilPF(*a)
            for i in range(3):
                assert rel_err(a[1 + i], g["O%d__pF%d" % (t, i)]) < 1e-12, (case, t, i)
    a, want = edge_case_arguments()
    suctionUnsaturatedSoilPF(*a)
    for i in range(3):
        assert np.array_equal(a[1 + i], want), (i, a[1 + i])
    with pytest.raises(ValueError):
        a[4] = a[4][:, :1]
        suctionUnsaturatedSoilPF(*a)


def test_model_stress_days_and_pf(gpu_lib):
    """The option-gated extras at model level: SoilMoistureStressDays (repStressDays) and pF0..2 (simulatePF) of a
    diagnostics model after the steps of the golden, against the reference's own soilloop class."""
    from conftest import golden_cases, golden_model, load_golden
    from lisflood_code_b200.hotpath import HotPathModel
    from test_oracle_soil_options_golden import LAYERS
    for case in golden_cases("soilopt_"):
        S, F, O = golden_model(case)
        X = {k[3:]: v for k, v in load_golden(case).items() if k.startswith("X__")}
        M = HotPathModel(S, diagnostics=True)
        for t in range(len(F)):
            M.step(F[t])
            rws, want = O[t]["RWS"], O[t]["SoilMoistureStressDays"]
            assert rel_err(M.get("RWS", 3), rws) < 1e-9
            got = M.get("SoilMoistureStressDays")
            clear = np.abs(rws - 1) > 1e-9           # the decision RWS < 1 is not a rounding matter there
            assert np.array_equal(got[clear], want[clear]) and (got != want).mean() < 1e-3, t
            pf = M.suction_pf({x: X["GenuInvAlpha" + x] for x in LAYERS}, {x: X["GenuInvN" + x] for x in LAYERS}, float(X["HeadMax"]))
            for i in range(3):
                assert rel_err(pf[i], O[t]["pF%d" % i]) < 1e-9, (case, t, i)
