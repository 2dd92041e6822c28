"""GPU parity tests of the kinematic-wave path, through the C ABI (ctypes), against the committed
golden vectors of the reference and against the C oracle on seeded synthetic catchments.

Bar: graph/ordering arrays bit-exact; discharge <= 1e-6 relative (abs floor 1e-12).
The observed error is ~1e-13; the tests assert 1e-9 so a regression in the solver shows up long before
the contractual tolerance is reached.
"""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9
GRAPH_KEYS = ("downstream_lookup", "upstream_lookup", "num_upstream_pixels", "pixels_ordered", "order_start_stop")


def _dx(g):
    return g["dx"] if g["dx"].ndim else float(g["dx"])


def _kw(gpu_lib):
    from lisflood_code_b200.hydrological_modules.kinematic_wave_parallel import kinematicWave
    return kinematicWave


# kwreal_: the reference's own test catchment (57 x 80, 2847 px).  Added after the round's GPU budget was spent and not run
# on a GPU yet, hence a non-strict expected failure (XPASS = fine); the CPU restatement passes it (tests/test_oracle_golden.py)
KW_CASES = golden_cases("kw_") + [
    pytest.param(c, marks=pytest.mark.xfail(strict=False, reason="never run on a GPU yet (added after the round's GPU budget "
                                                                  "was spent)")) for c in golden_cases("kwreal_")]


@pytest.mark.parametrize("case", KW_CASES)
def test_golden_graph_bit_exact(gpu_lib, case):
    g = load_golden(case)
    kw = _kw(gpu_lib)(g["ldd"], g["mask"], g["alpha"], float(g["beta"]), _dx(g), float(g["dt"]))
    for k in GRAPH_KEYS:
        got = getattr(kw, k)
        assert got.dtype == g[k].dtype and got.shape == g[k].shape, k
        assert np.array_equal(got, g[k]), k


@pytest.mark.parametrize("case", KW_CASES)
def test_golden_routing(gpu_lib, case):
    g = load_golden(case)
    a2 = g.get("alpha2")
    kw = _kw(gpu_lib)(g["ldd"], g["mask"], g["alpha"], float(g["beta"]), _dx(g), float(g["dt"]), alpha_floodplains=a2)
    Q = g["q0"].copy()
    for s in range(g["Q_main"].shape[0]):
        assert kw.kinematicWaveRouting(Q, g["q"], "main_channel") is None
        assert rel_err(Q, g["Q_main"][s]) < TOL, (case, s)
    if a2 is not None:
        Q2 = g["q0"] * 0.5
        for s in range(g["Q_fp"].shape[0]):
            kw.kinematicWaveRouting(Q2, g["q"] * 0.3, "floodplains")
            assert rel_err(Q2, g["Q_fp"][s]) < TOL, (case, s)
    # device-resident wavefront == repeated calls
    steps = g["Q_main"].shape[0]
    kw.set_discharge(g["q0"])
    kw.set_lateral_inflow(g["q"])
    kw.run(steps)
    assert rel_err(kw.get_discharge(), g["Q_main"][-1]) < TOL


@pytest.mark.parametrize("case", golden_cases("kwadv_"))
def test_solver_stopping_rule_adversarial(gpu_lib, case):
    """Inputs chosen against the stopping rule (tests/golden/make_golden.py::adversarial_case): discharges from 1e-13
    to 1e9, where the reference leaves its Newton loop through `Q == previous` or wanders between two neighbouring
    values until its 3000-iteration cap (kinematic_wave_parallel_tools.py:73-80), while the device solver also stops
    when the relative Newton step is <= 1e-8 (lf_kw_solve.cuh).  The values the UNMODIFIED reference returned are the
    golden; the device result may differ from them only at the rounding level.  Observed: <= 3e-14 relative everywhere
    except where a negative side flow cancels against a*Qold^beta (C = a*Qold^beta + q*dx loses 3-4 digits, so the last-bit
    difference between the device's z^3 and the reference's pow(Q, 0.6) is amplified accordingly: 4.4e-12 on one pixel whose
    discharge is 1e-10) -- the same amplification the reference shows between numexpr's and NumPy's pow."""
    g = load_golden(case)
    kw = _kw(gpu_lib)(g["ldd"], g["mask"], g["alpha"], float(g["beta"]), _dx(g), float(g["dt"]))
    Q = g["q0"].copy()
    worst = 0.0
    for s in range(g["Q_main"].shape[0]):
        kw.kinematicWaveRouting(Q, g["q"])
        assert np.array_equal(Q == 0, g["Q_main"][s] == 0), (case, s)     # the dry / wet decision is the reference's
        worst = max(worst, rel_err(Q, g["Q_main"][s]))
    print("%s: worst rel. deviation from the reference %.2e" % (case, worst))
    assert worst < 1e-10, (case, worst)
    kw.set_discharge(g["q0"])
    kw.set_lateral_inflow(g["q"])
    kw.run(g["Q_main"].shape[0])
    assert rel_err(kw.get_discharge(), g["Q_main"][-1]) < 1e-10


@pytest.mark.parametrize("rows,cols,noise,maskf,seed,dxmap", [
    (257, 190, 3.0, 0.0, 21, True),     # shallow forest
    (300, 211, 0.3, 0.15, 22, False),   # deep trees with holes
    (1000, 1000, 0.3, 0.0, 23, True),   # ~1000 levels
    (64, 3000, 1.0, 0.3, 24, True),     # wide and short
])
def test_synthetic_vs_oracle(gpu_lib, oracle, rows, cols, noise, maskf, seed, dxmap):
    from lisflood_code_b200 import synthetic
    ldd, mask = synthetic.random_ldd(rows, cols, seed=seed, noise=noise, mask_fraction=maskf)
    n = int(mask.sum())
    alpha, q0, q = synthetic.routing_fields(n, seed)
    dx = np.random.default_rng(seed).uniform(3000, 7000, n) if dxmap else 5000.0
    ora = oracle.KinematicWaveOracle(ldd[mask], mask, alpha, 0.6, dx, 3600.0)
    kw = _kw(gpu_lib)(ldd[mask], mask, alpha, 0.6, dx, 3600.0)
    for k in GRAPH_KEYS:
        assert np.array_equal(getattr(kw, k), getattr(ora, k)), k
    # per-call API (host arrays, in place)
    Qg, Qo = q0.copy(), q0.copy()
    for s in range(3):
        kw.kinematicWaveRouting(Qg, q)
        ora.kinematicWaveRouting(Qo, q)
        assert rel_err(Qg, Qo) < TOL
    # device-resident wavefront with a time-varying inflow multiplier, compared every 10 steps
    rng = np.random.default_rng(seed + 5)
    kw.set_discharge(q0)
    kw.set_lateral_inflow(q)
    Qo = q0.copy()
    for chunk in range(3):
        scale = rng.uniform(0.2, 3.0, 10)
        kw.run(10, inflow_scale=scale)
        for s in range(10):
            ora.kinematicWaveRouting(Qo, q * scale[s])
        assert rel_err(kw.get_discharge(), Qo) < TOL, chunk


def test_layout_properties(gpu_lib):
    """Size-independent invariants of the breadth-first device layout on a 2000x2000 catchment:
    a permutation; level spans match order_start_stop; children contiguous and one level below."""
    from lisflood_code_b200 import synthetic
    ldd, mask = synthetic.random_ldd(2000, 2000, seed=31, noise=0.5)
    n = int(mask.sum())
    kw = _kw(gpu_lib)(ldd[mask], mask, np.ones(n), 0.6, 5000.0, 3600.0)
    pop, ls = kw.storage_layout()
    assert np.array_equal(np.sort(pop), np.arange(n))
    oss = kw.order_start_stop
    assert np.array_equal(ls[:-1], oss[:, 0]) and np.array_equal(ls[1:], oss[:, 1])
    order = kw.pixels_ordered
    for l in (0, len(ls) // 2, len(ls) - 2):
        assert np.array_equal(np.sort(pop[ls[l]:ls[l + 1]]), order[ls[l]:ls[l + 1]])
    # every non-outlet pixel's downstream sits exactly one level later
    pos = np.empty(n, np.int64)
    pos[pop] = np.arange(n)
    ds = kw.downstream_lookup.astype(np.int64)
    lev = np.searchsorted(ls, pos, side="right") - 1
    has = ds >= 0
    assert np.all(lev[ds[has]] == lev[has] + 1)
    assert np.all(lev[~has] == len(ls) - 2)


def test_edge_cases(gpu_lib, oracle):
    KW = _kw(gpu_lib)
    # single pixel (the reference itself crashes here; we define it as a single pit)
    kw = KW(np.array([5.0]), np.ones((1, 1), bool), np.array([1.5]), 0.6, 1000.0, 3600.0)
    Q = np.array([2.0])
    kw.kinematicWaveRouting(Q, np.array([1e-4]))
    ora = oracle.KinematicWaveOracle(np.array([5.0]), np.ones((1, 1), bool), np.array([1.5]), 0.6, 1000.0, 3600.0)
    Qo = np.array([2.0])
    ora.kinematicWaveRouting(Qo, np.array([1e-4]))
    assert rel_err(Q, Qo) < TOL
    # all-zero state with zero inflow stays exactly zero; negative inflow clamps to zero
    g = load_golden("kw_40x50_masked")
    kw = KW(g["ldd"], g["mask"], g["alpha"], 0.6, _dx(g), 3600.0)
    Q = np.zeros(kw.num_pixels)
    kw.kinematicWaveRouting(Q, np.zeros(kw.num_pixels))
    assert np.all(Q == 0.0)
    kw.kinematicWaveRouting(Q, np.full(kw.num_pixels, -1.0))
    assert np.all(Q == 0.0)
    # bad section -> Exception, like the reference
    with pytest.raises(Exception):
        kw.kinematicWaveRouting(Q, Q, "floodplain")
    with pytest.raises(Exception):  # floodplains without alpha_floodplains
        kw.kinematicWaveRouting(Q, Q, "floodplains")
    # bad LDD codes / cycles are reported, not looped on
    bad = g["ldd"].copy()
    bad[7] = 12.0
    with pytest.raises(gpu_lib.LisfloodB200Error) as e:
        KW(bad, g["mask"], g["alpha"], 0.6, 1000.0, 3600.0)
    assert e.value.code == gpu_lib.LF_ERR_BAD_LDD
    with pytest.raises(gpu_lib.LisfloodB200Error) as e:
        KW(np.array([6.0, 4.0]), np.ones((1, 2), bool), np.ones(2), 0.6, 1000.0, 3600.0)
    assert e.value.code == gpu_lib.LF_ERR_LDD_CYCLE


def test_nancheck_warns_once(gpu_lib):
    import warnings
    g = load_golden("kw_4x4_south")
    kw = _kw(gpu_lib)(g["ldd"], g["mask"], g["alpha"], 0.6, 1000.0, 3600.0, flagnancheck=True)
    Q = np.ones(16)
    Q[5] = np.nan
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        kw.kinematicWaveRouting(Q, np.zeros(16))
        kw.kinematicWaveRouting(Q, np.zeros(16))
    assert len(w) == 1 and kw.kinematic_wave_warning_printed


def test_routing_accepts_any_array_like_the_reference(gpu_lib):
    """The reference mutates whatever array it is given (float32, strided views, NumpyModified): so does the mirror."""
    from lisflood_code_b200.global_modules.add1 import NumpyModified
    g = load_golden("kw_40x50_masked")
    kw = _kw(gpu_lib)(g["ldd"], g["mask"], g["alpha"], float(g["beta"]), _dx(g), float(g["dt"]))
    want = g["q0"].copy()
    kw.kinematicWaveRouting(want, g["q"])
    n = want.size
    strided = np.zeros(2 * n)
    strided[::2] = g["q0"]
    view = strided[::2]                                   # non-contiguous float64
    assert kw.kinematicWaveRouting(view, g["q"]) is None
    assert np.array_equal(strided[::2], want) and np.all(strided[1::2] == 0)
    q32 = g["q0"].astype(np.float32)                      # float32: routed in float64, stored back as float32
    ref32 = q32.astype(np.float64)
    kw.kinematicWaveRouting(ref32, g["q"])
    kw.kinematicWaveRouting(q32, g["q"])
    assert q32.dtype == np.float32 and np.array_equal(q32, ref32.astype(np.float32))
    wrapped = NumpyModified(g["q0"].copy(), ["pixel"])
    kw.kinematicWaveRouting(wrapped, g["q"])
    assert np.array_equal(np.asarray(wrapped), want)


@pytest.mark.parametrize("rows,cols,noise,steps", [(600, 500, 0.3, 30), (200, 3000, 1.0, 7)])
def test_narrow_runs_equal_one_launch_per_diagonal(gpu_lib, rows, cols, noise, steps):
    """Option "narrow_runs": the diagonals of at most 512 work items (here the far ends of the longest paths of one basin)
    run back to back in one block (k_kw_narrow_run); the discharge is bit-identical to one launch per diagonal, with and
    without graph replay."""
    from lisflood_code_b200 import synthetic
    ldd, mask = synthetic.random_ldd(rows, cols, seed=41, noise=noise, single_outlet=True)
    n = int(mask.sum())
    alpha, q0, q = synthetic.routing_fields(n, 41)
    scale = np.random.default_rng(42).uniform(0.2, 3.0, steps)
    got = {}
    for narrow, graphs in ((1, 1), (0, 1), (1, 0), (0, 0)):
        kw = _kw(gpu_lib)(ldd[mask], mask, alpha, 0.6, 5000.0, 3600.0)
        kw.set_option("narrow_runs", narrow)
        kw.set_option("cuda_graphs", graphs)
        kw.set_discharge(q0)
        kw.set_lateral_inflow(q)
        kw.run(steps, inflow_scale=scale)
        kw.run(steps, inflow_scale=scale)          # the second run replays the captured graph
        got[(narrow, graphs)] = kw.get_discharge()
        assert np.isfinite(got[(narrow, graphs)]).all()
    for k, v in got.items():
        assert np.array_equal(v, got[(0, 0)]), k
