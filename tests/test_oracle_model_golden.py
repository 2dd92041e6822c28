"""Pins the CPU restatement of the full hot-path step (oracle/lisf_oracle_model.py + lisf_oracle_soil.c)
against golden vectors produced by the reference's OWN module classes
(routing/surface_routing/soilloop/soil/opensealed/groundwater .dynamic(), see oracle/ref_modules.py)."""
import numpy as np
import pytest

from conftest import golden_cases, golden_model, rel_err


@pytest.mark.parametrize("case", golden_cases("model_"))
def test_model_step_matches_reference(oracle, case):
    from oracle import lisf_oracle_model as om
    S, F, O = golden_model(case)
    M = om.OracleModel(S)
    for t in range(len(F)):
        M.step(F[t])
        for k, want in O[t].items():
            got = np.asarray(getattr(M.var, k))
            assert got.shape == want.shape, (k, got.shape, want.shape)
            assert rel_err(got, want) < 1e-10, (case, t, k, rel_err(got, want))


def test_synthetic_generator_is_the_one_that_made_the_goldens():
    """tests that re-generate S from the seed rely on this."""
    from lisflood_code_b200 import synthetic
    S, F, O = golden_model("model_20x22_6h")
    S2 = synthetic.full_stack(20, 22, seed=43, split_routing=False, channel_threshold=8, dt_sec=21600.0)
    for k in ("W1a", "KSat1b", "ChannelAlpha", "LddKinematic", "SoilFraction", "OFAlpha"):
        assert np.array_equal(S[k], S2[k]), k
    F2 = synthetic.forcing(S2, 1, 43)
    assert np.array_equal(F2["Rain"], F[1]["Rain"])
