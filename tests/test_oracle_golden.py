"""Pins the C oracle (oracle/lisf_oracle.c) against the golden vectors produced by the UNMODIFIED
reference kernels (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, rel_err

GRAPH_KEYS = ("downstream_lookup", "upstream_lookup", "num_upstream_pixels", "pixels_ordered", "order_start_stop")


def _dx(g):
    return g["dx"] if g["dx"].ndim else float(g["dx"])


KW_CASES = golden_cases("kw_") + golden_cases("kwreal_")    # kwreal_: the reference's own test catchment (57 x 80, 2847 px)


@pytest.mark.parametrize("case", KW_CASES)
def test_graph_bit_exact(oracle, case):
    g = load_golden(case)
    kw = oracle.KinematicWaveOracle(g["ldd"], g["mask"], g["alpha"], float(g["beta"]), _dx(g), float(g["dt"]))
    for k in GRAPH_KEYS:
        got = getattr(kw, k)
        assert got.dtype == g[k].dtype, k
        assert np.array_equal(got, g[k]), k


@pytest.mark.parametrize("case", KW_CASES)
def test_routing_matches_reference(oracle, case):
    g = load_golden(case)
    a2 = g.get("alpha2")
    kw = oracle.KinematicWaveOracle(g["ldd"], g["mask"], g["alpha"], float(g["beta"]), _dx(g), float(g["dt"]),
                                    alpha_floodplains=a2)
    Q = g["q0"].copy()
    for s in range(g["Q_main"].shape[0]):
        kw.kinematicWaveRouting(Q, g["q"], "main_channel")
        # numpy's vectorised pow (reference wrapper) vs libm pow differ in the last ulp only
        assert rel_err(Q, g["Q_main"][s]) < 1e-11, (case, s)
    if a2 is not None:
        Q2 = g["q0"] * 0.5
        for s in range(g["Q_fp"].shape[0]):
            kw.kinematicWaveRouting(Q2, g["q"] * 0.3, "floodplains")
            assert rel_err(Q2, g["Q_fp"][s]) < 1e-11, (case, s)


@pytest.mark.parametrize("case", golden_cases("kwadv_"))
def test_routing_adversarial_inputs(oracle, case):
    """Stopping-rule stress (tests/golden/make_golden.py::adversarial_case): Q from 1e-13 to 1e9, alpha / dx over six decades."""
    g = load_golden(case)
    kw = oracle.KinematicWaveOracle(g["ldd"], g["mask"], g["alpha"], float(g["beta"]), _dx(g), float(g["dt"]))
    Q = g["q0"].copy()
    for s in range(g["Q_main"].shape[0]):
        kw.kinematicWaveRouting(Q, g["q"], "main_channel")
        assert np.array_equal(Q == 0, g["Q_main"][s] == 0)
        assert rel_err(Q, g["Q_main"][s]) < 1e-13, (case, s)


def test_known_answer_vector(oracle):
    """4x4, all south; rows after one call = 0.31022586, 0.5392451, 0.71516453, 0.85306031."""
    g = load_golden("kw_4x4_south")
    kw = oracle.KinematicWaveOracle(g["ldd"], g["mask"], g["alpha"], 0.6, 1000.0, 3600.0)
    Q = np.ones(16)
    kw.kinematicWaveRouting(Q, np.full(16, 1e-4))
    assert np.allclose(Q.reshape(4, 4)[:, 0], [0.31022586, 0.5392451, 0.71516453, 0.85306031], rtol=0, atol=5e-9)
    assert np.array_equal(kw.order_start_stop, [[0, 4], [4, 8], [8, 12], [12, 16]])
    assert np.array_equal(kw.pixels_ordered, np.arange(16))
    assert kw.upstream_lookup.shape == (16, 1)


def test_bad_section_and_bad_codes(oracle):
    g = load_golden("kw_4x4_south")
    kw = oracle.KinematicWaveOracle(g["ldd"], g["mask"], g["alpha"], 0.6, 1000.0, 3600.0)
    with pytest.raises(Exception):
        kw.kinematicWaveRouting(np.ones(16), np.zeros(16), "floodplain")
    bad = g["ldd"].copy()
    bad[3] = 11.0
    with pytest.raises(ValueError):
        oracle.KinematicWaveOracle(bad, g["mask"], g["alpha"], 0.6, 1000.0, 3600.0)
    # two pixels draining into each other: the reference would never return (kinematic_wave_parallel.py:99)
    cyc = np.array([[6.0, 4.0]])
    with pytest.raises(ValueError):
        oracle.KinematicWaveOracle(cyc.ravel(), np.ones((1, 2), bool), np.ones(2), 0.6, 1000.0, 3600.0)


@pytest.mark.parametrize("case", golden_cases("kwreal_"))
def test_ldd_operators_on_the_reference_catchment(case):
    """global_modules/ldd_ops.py against PCRaster's own products for the reference's test catchment: the upstream-area
    map ec_upArea = accuflux(ldd, pixarea), bit for bit; the one link that leaves the mask (the outlet) becomes a pit
    (lddmask / lddrepair), every other code is kept."""
    from lisflood_code_b200.global_modules import ldd_ops
    g = load_golden(case)
    mask, raw = g["mask"], g["ldd_raw"].astype(np.float64)
    ds = ldd_ops.downstream_index(raw, mask)
    leaving = (ds < 0) & (raw != 5)
    assert leaving.sum() == 1 and g["upArea"][leaving][0] == g["upArea"].max()
    assert np.array_equal(ldd_ops.accuflux(ds, g["pixarea"].astype(np.float64)), g["upArea"])
    repaired = ldd_ops.lddrepair_codes(raw, mask)
    assert np.array_equal(repaired, g["ldd"]) and np.array_equal(repaired[~leaving], raw[~leaving]) and repaired[leaving][0] == 5
    order, hops = ldd_ops.topological_order(ds)
    assert hops[leaving][0] == 0 and hops.max() + 1 == g["order_start_stop"].shape[0]     # as deep as the reference's ordering
    rank = np.empty(order.size, np.int64)
    rank[order] = np.arange(order.size)
    assert (rank[ds >= 0] < rank[ds[ds >= 0]]).all()                    # every pixel before its downstream pixel
    assert np.array_equal(ldd_ops.catchment_of_pits(ds), np.ones(ds.size, np.int64))    # one basin
    assert np.array_equal(ldd_ops.upstream_sum(ds, g["upArea"]) + g["pixarea"], g["upArea"])   # upArea = own area + upstream(upArea)
