"""Reservoirs and lakes inside the routing sub-step loop (the next row after the hot path): the CPU
oracle's restatement (oracle/lisf_oracle_model.py: lakes_inloop, reservoir_inloop) against golden vectors produced by
the reference's OWN classes (routing.dynamic -> lakes.dynamic_inloop / reservoir.dynamic_inloop, reservoir.py:173-322,
lakes.py:199-297).  CPU only: the device path does not take structures yet (round 2)."""
import numpy as np
import pytest

from conftest import golden_cases, golden_model, rel_err


@pytest.mark.parametrize("case", golden_cases("structures_"))
def test_structures_step_matches_reference(oracle, case):
    from oracle import lisf_oracle_model as om
    S, F, O = golden_model(case)
    for k in ("ReservoirIndex", "LakeIndex"):
        S[k] = np.asarray(S[k], np.int64)
    M = om.OracleModel(S)
    assert M.var.simulateReservoirs and M.var.simulateLakes
    for t in range(len(F)):
        M.step(F[t])
        for k, want in O[t].items():
            got = np.asarray(getattr(M.var, k))
            assert got.shape == want.shape, (k, got.shape, want.shape)
            assert rel_err(got, want) < 1e-10, (case, t, k, rel_err(got, want))
    # the reservoirs went through more than one outflow regime and the structures changed the discharge
    fills = np.stack([O[t]["ReservoirFillCC"] for t in range(len(F))])
    assert fills.max() - fills.min() > 0.05
    assert float(np.max(O[-1]["QResOutM3Dt"])) > 0 and float(np.max(O[-1]["QLakeOutM3Dt"])) > 0


def test_structure_ldd_surgery():
    """structures.py:43-61: cells just upstream of a structure are pits of LddKinematic; inflow is gathered through
    downstruct built on the unmodified network."""
    from lisflood_code_b200 import synthetic
    from lisflood_code_b200.global_modules import ldd_ops
    S = synthetic.full_stack(40, 46, seed=71, split_routing=False, channel_threshold=10)
    ldd0 = S["LddKinematic"].copy()
    synthetic.add_structures(S, 3, 2, seed=71)
    sites = np.concatenate([S["ReservoirIndex"], S["LakeIndex"]])
    ds0 = ldd_ops.downstream_index(ldd0, S["mask"])
    ups = np.flatnonzero(np.isin(ds0, sites))
    assert ups.size >= sites.size and np.all(S["LddKinematic"][ups] == 5)
    untouched = np.setdiff1d(np.arange(S["N"]), ups)
    assert np.array_equal(S["LddKinematic"][untouched], ldd0[untouched])
    assert np.array_equal(np.flatnonzero(S["IsUpsOfStructureKinematicC"]), ups)
    assert np.all(S["downstruct"][ups] == ds0[ups]) and S["downstruct"].max() == S["N"]
