"""Size-independent properties of the full step on a large catchment (no oracle needed, so they scale to the
bench sizes): per-column and per-pixel water balance of the soil stage, mass balance of the channel network over
the sub-step loop, sign / bound invariants."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols", [(1500, 1400)])
def test_water_and_mass_balance_large(gpu_lib, rows, cols):
    from lisflood_code_b200 import synthetic
    from lisflood_code_b200.hotpath import HotPathModel
    from lisflood_code_b200.hydrological_modules.kinematic_wave_parallel import kinematicWave
    S = synthetic.full_stack(rows, cols, seed=51, ldd_noise=0.4)
    n = S["N"]
    M = HotPathModel(S, diagnostics=True)
    kw = kinematicWave(S["LddKinematic"], S["mask"], S["ChannelAlpha"], 0.6, S["ChanLength"], S["DtRouting"])
    pits = kw.downstream_lookup < 0
    kw.close()
    g3 = lambda k: M.get(k, 3)
    for t in range(2):
        F = synthetic.forcing(S, t, 51)
        before = {k: g3(k) for k in ("CumInterception", "W1a", "W1b", "W2", "UZ")}
        lz0, cs0, m30 = M.get("LZ"), M.get("CumInterSealed"), M.get("ChanM3Kin")
        M.step(F)
        # ---- soil column: rain + snowmelt = d(interception store) + evaporated interception + d(soil water) +
        #      transpiration + soil evaporation + d(upper zone) + upper-zone outflow + percolation + surface runoff
        W = {k: g3(k) for k in ("CumInterception", "W1a", "W1b", "W2", "UZ")}
        dW = (W["W1a"] - before["W1a"]) + (W["W1b"] - before["W1b"]) + (W["W2"] - before["W2"])
        surf = np.maximum(g3("AvailableWaterForInfiltration") - g3("Infiltration"), 0)
        rhs = (W["CumInterception"] - before["CumInterception"]) + g3("TaInterception") + dW + g3("Ta") + g3("ESAct") + \
              (W["UZ"] - before["UZ"]) + g3("UZOutflow") + g3("GwPercUZLZ") + surf
        res = (F["Rain"] + F["SnowMelt"])[None] - rhs
        assert np.abs(res).max() < 1e-9, np.abs(res).max()
        # ---- lower zone and sealed store
        lz = M.get("LZ")
        assert np.abs((lz - lz0) - (M.get("GwPercUZLZPixel") - M.get("LZOutflow") - M.get("GwLossLZ"))).max() < 1e-9
        assert np.abs((M.get("CumInterSealed") - cs0) - (M.get("InterSealed") - M.get("TASealed"))).max() < 1e-12
        # ---- bounds
        assert (W["W1a"] <= S["WS1a"] + 1e-9).all() and (W["W2"] <= S["WS2"] + 1e-9).all()
        for k in ("UZ", "CumInterception"):
            assert (W[k] >= 0).all(), k
        for k in ("ChanQKin", "OFQOther", "OFQForest", "OFQDirect", "ChanQAvg"):
            a = M.get(k)
            assert np.isfinite(a).all() and (a >= 0).all(), k
        # ---- channel network over the NoRoutSteps sub-steps: d(storage) = side inflow - outflow at the outlets
        m3 = M.get("ChanM3Kin")
        inflow = np.where(S["IsChannelKinematic"], M.get("ToChanM3RunoffDt"), 0).sum() * S["NoRoutSteps"]
        out = (M.get("sumDisDay")[pits] * S["DtRouting"]).sum()
        resid = (m3 - m30).sum() - (inflow - out)
        assert abs(resid) < 1e-9 * max(abs(inflow), abs(out)), (resid, inflow, out)
        # dis = ChanQAvg is the sub-step mean (Lisflood_dynamic.py:218)
        assert np.allclose(M.get("ChanQAvg"), M.get("sumDisDay") / S["NoRoutSteps"], rtol=1e-15, atol=0)
