"""Pins the CPU restatement of the feeder modules (oracle/lisf_oracle_feeders.py: readmeteo scaling, snow, frost) to
goldens produced by the reference's OWN snow / frost classes (tests/golden/make_golden.py::feeders_case).  CPU only."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, rel_err


def feeder_case(name):
    g = load_golden(name)
    P = {k[3:]: v for k, v in g.items() if k.startswith("P__")}
    S = {k[3:]: v for k, v in g.items() if k.startswith("S__")}
    steps = int(g["steps"])
    R = [{k.split("__", 1)[1]: v for k, v in g.items() if k.startswith("R%d__" % t)} for t in range(steps)]
    O = [{k.split("__", 1)[1]: v for k, v in g.items() if k.startswith("O%d__" % t)} for t in range(steps)]
    days = [int(g["CalendarDay%d" % t]) for t in range(steps)]
    return P, S, float(g["DtSec"]), R, days, O


@pytest.mark.parametrize("case", golden_cases("feeders_"))
def test_feeders_match_reference(case):
    from oracle.lisf_oracle_feeders import FeederOracle
    P, S, dt, R, days, O = feeder_case(case)
    F = FeederOracle(P, S, dt)
    for t in range(len(R)):
        got = F.step(R[t], days[t])
        for k, want in O[t].items():
            if want.dtype == bool:
                assert np.array_equal(got[k], want), (case, t, k)
            else:
                assert rel_err(got[k], want) < 1e-13, (case, t, k, rel_err(got[k], want))
    assert O[-1]["isFrozenSoil"].any() and (O[-1]["SnowCover"] > 0).any() and (O[-1]["SnowMelt"] > 0).any()


def test_lai_interval_lookup():
    from lisflood_code_b200.hydrological_modules.snow import LAI_INTERVAL_START, lai_interval
    L1, j = [], 0
    for i in range(367):                    # leafarea.py:63-69
        if i >= LAI_INTERVAL_START[j + 1]:
            j += 1
        L1.append(j)
    assert [lai_interval(d) for d in range(367)] == L1
