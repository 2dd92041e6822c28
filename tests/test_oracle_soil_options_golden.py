"""Option-gated soil extras (repStressDays, simulatePF) on the CPU: the NumPy restatement of the reference's pF kernel
against goldens made by the reference's OWN soilloop class with the options on (tests/golden/make_golden.py::
soil_options_case), plus the hand-checked edge cases of saturationDegree / pressureHead (soilloop.py:379-383, :428-432)."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, rel_err

LAYERS = ("1a", "1b", "2")


def pf_arguments(S, X, W):
    """The 26 arguments of suctionUnsaturatedSoilPF after the three output arrays: from the static stack S, the extra
    parameters X and the soil moisture W = (W1a, W1b, W2)."""
    a = [np.array([0, 1, 2], np.int64), np.empty_like(W[0]), np.empty_like(W[0]), np.empty_like(W[0]), W[0], W[1], W[2]]
    for k in ("WRes", "WS", "PoreSpaceNotZero"):
        a += [S[k + lay] for lay in LAYERS]
    a += [X["GenuInvAlpha" + lay] for lay in LAYERS]
    a += [S["GenuInvM" + lay] for lay in LAYERS]
    a += [X["GenuInvN" + lay] for lay in LAYERS]
    a.append(float(X["HeadMax"]))
    return a


def split_golden(g):
    S = {k[3:]: v for k, v in g.items() if k.startswith("S__")}
    X = {k[3:]: v for k, v in g.items() if k.startswith("X__")}
    return S, X


@pytest.mark.parametrize("case", golden_cases("soilopt_"))
def test_pf_restatement_matches_the_reference_kernel(case):
    from oracle import lisf_oracle_model as om
    g = load_golden(case)
    S, X = split_golden(g)
    for t in range(int(g["steps"])):
        W = [np.ascontiguousarray(g["O%d__W%s" % (t, lay)]) for lay in LAYERS]
        a = pf_arguments(S, X, W)
        om.suction_unsaturated_soil_pf(*a)
        for i in range(3):
            assert rel_err(a[1 + i], g["O%d__pF%d" % (t, i)]) < 1e-13, (t, i)
        # repStressDays, soilloop.py:597-598
        want = g["O%d__SoilMoistureStressDays" % t]
        assert np.array_equal(np.where(g["O%d__RWS" % t] < 1, float(S["DtSec"]) / 86400.0, 0.0), want)
        assert 0 < (want > 0).mean() < 1


def edge_case_arguments():
    """Two pixels x three fractions: residual / saturated / no pore space / nearly dry columns."""
    V, N = 3, 2
    wres, ws = np.full((3, N), 10.0), np.full((3, N), 110.0)
    pore = np.ones((3, N), bool)
    pore[2, 1] = False
    W = np.array([[10.0, 110.0], [5.0, 500.0], [10.0 + 1e-7, 60.0]])       # sat: 0, 1 | 0 (below), 1 (above) | 1e-9, no pores
    inva, invm, invn = np.full((3, N), 50.0), np.full((3, N), 4.0), np.full((3, N), 0.75)
    a = [np.array([0, 1, 2], np.int64)] + [np.empty((V, N)) for _ in range(3)] + [W.copy(), W.copy(), W.copy()]
    a += [wres] * 3 + [ws] * 3 + [pore] * 3 + [inva] * 3 + [invm] * 3 + [invn] * 3 + [1.0e7]
    want = np.array([[7.0, -1.0], [7.0, -1.0], [7.0, 7.0]])               # head = HeadMax -> 7; head = 0 -> -1
    return a, want


def test_pf_edge_cases():
    from oracle import lisf_oracle_model as om
    a, want = edge_case_arguments()
    om.suction_unsaturated_soil_pf(*a)
    for i in range(3):
        assert np.array_equal(a[1 + i], want), (i, a[1 + i])
