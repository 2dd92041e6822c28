"""The reference's own test catchment end to end, from a committed fixture (no /root/reference needed): real static maps
through the init mirrors, real meteo maps, and the soil-moisture maps of the output stacks the reference SHIPS for that run
(tests/data/LF_ETRS89_UseCase/reference/output_reference_daily).  CPU: the restatement (feeder oracle + model oracle)
reproduces its own recorded run bit for bit and the shipped maps to 1e-6 (the reference's comparator works at 1e-4)."""
import numpy as np
import pytest

from conftest import golden_cases
from realcase_common import THETA, load


@pytest.mark.parametrize("case", golden_cases("realcase_"))
def test_restatement_reproduces_the_shipped_soil_moisture(oracle, case):
    from oracle import lisf_oracle_model as om
    from oracle.lisf_oracle_feeders import FeederOracle, lai_term
    S, P, state, raw, days, lai, want, shipped = load(case)
    n = S["N"]
    assert n == 2847 and S["SplitRouting"]
    feeder = FeederOracle({k: (np.full(n, v) if np.ndim(v) == 0 else v) for k, v in P.items() if k != "kgb"}, state, S["DtSec"])
    O = om.OracleModel(S)
    for t in range(len(raw)):
        o = feeder.step(raw[t], days[t])
        O.step({"Rain": o["Rain"], "SnowMelt": o["SnowMelt"], "ETRef": o["ETRef"], "EWRef": o["EWRef"], "ESRef": o["ESRef"],
                "isFrozenSoil": o["isFrozenSoil"], "LAI": lai[t], "LAITerm": lai_term(P["kgb"], lai[t])})
        for k, w in want[t].items():
            got = np.stack(feeder.SnowCoverS) if k == "SnowCoverS" else (feeder.FrostIndex if k == "FrostIndex" else
                                                                         np.asarray(getattr(O.var, k)))
            assert np.array_equal(got, w), (t, k)
        for name, (attr, row) in THETA.items():
            assert np.abs(np.asarray(getattr(O.var, attr))[row] - shipped[t][name]).max() < 1e-6, (t, name)
