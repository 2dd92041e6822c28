"""Reservoirs and lakes inside the device sub-step loop against golden vectors made by the reference's
OWN classes (routing.dynamic -> lakes.dynamic_inloop / reservoir.dynamic_inloop; reservoir.py:173-322, lakes.py:199-297,
structures.py:43-61), through the C ABI.  Every map of the golden, single and split routing, daily and 6-hourly."""
import numpy as np
import pytest

from conftest import golden_cases, golden_model, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-8


@pytest.mark.parametrize("case", golden_cases("structures_"))
def test_structures_golden(gpu_lib, case):
    from lisflood_code_b200.hotpath import HotPathModel
    S, F, O = golden_model(case)
    for k in ("ReservoirIndex", "LakeIndex"):
        S[k] = np.asarray(S[k], np.int64)
    M = HotPathModel(S, diagnostics=True)
    for t in range(len(F)):
        M.step(F[t])
        bad = {}
        for k, want in O[t].items():
            e = rel_err(np.asarray(M.get(k, 3 if want.ndim == 2 else 1)), want)
            if not e < TOL:
                bad[k] = e
        assert not bad, (case, t, bad)
    assert float(np.max(M.get("QResOutM3Dt"))) > 0 and float(np.max(M.get("QLakeOutM3Dt"))) > 0


@pytest.mark.parametrize("split", [False, True])
def test_structures_vs_oracle_larger(gpu_lib, oracle, split):
    """A larger seeded catchment with more structures, lean (production) build, against the CPU restatement."""
    from lisflood_code_b200 import synthetic
    from lisflood_code_b200.hotpath import HotPathModel
    from oracle import lisf_oracle_model as om
    S = synthetic.full_stack(140, 150, seed=91, split_routing=split, channel_threshold=12, dt_sec=21600.0 if split else 86400.0)
    synthetic.add_structures(S, 6, 5, seed=91)
    O, M = om.OracleModel(S), HotPathModel(S, diagnostics=False)
    keys = ["ChanQAvg", "ChanQ", "ChanQKin", "ChanM3Kin", "ReservoirStorageM3", "ReservoirFill", "QResOutM3Dt", "LakeStorageM3",
            "LakeLevel", "LakeOutflow", "LakeInflowOld", "LakeStorageM3Balance", "QLakeOutM3Dt"] + (["Chan2QKin"] if split else [])
    for t in range(5):
        F = synthetic.forcing(S, t, 91)
        O.step(F)
        M.step(F)
        bad = {k: rel_err(M.get(k), np.asarray(getattr(O.var, k))) for k in keys}
        bad = {k: v for k, v in bad.items() if not v < TOL}
        assert not bad, (t, bad)
