"""reservoir.initial() / lakes.initial() mirrors against goldens made by the reference's OWN methods
(oracle/ref_init.py::structures_initial; reservoir.py:52-170, lakes.py:52-196): site selection, lookup tables, calibration,
cold and warm start.  NumPy expressions in the reference's order: compared bit for bit (tolerance 0).  CPU only."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden

SKIP = {"ReservoirSites", "LakeSitesCC"}     # PCRaster map object / unused duplicate


@pytest.mark.parametrize("case", golden_cases("structinit_"))
def test_structures_initial_matches_reference(case):
    from lisflood_code_b200.Lisflood_initial import InitialVariables
    from lisflood_code_b200.hydrological_modules.lakes import lakes
    from lisflood_code_b200.hydrological_modules.reservoir import reservoir
    g = load_golden(case)
    maps = {k[5:]: (float(v) if v.ndim == 0 else v) for k, v in g.items() if k.startswith("raw__")}
    maps.update({k[7:]: v for k, v in g.items() if k.startswith("table__")})
    var = InitialVariables(g["mask"], maps, {"simulateLakes": True, "simulateReservoirs": True}, DtSec=float(g["DtSec"]))
    for k, v in g.items():
        if k.startswith("state__"):
            setattr(var, k[7:], v.copy() if v.ndim else v.item())
    lakes(var).initial()
    reservoir(var).initial()
    checked = 0
    for key, want in g.items():
        if not key.startswith("out__") or key[5:] in SKIP:
            continue
        name = key[5:]
        assert hasattr(var, name), name
        got = np.asarray(getattr(var, name))
        assert got.shape == want.shape, (name, got.shape, want.shape)
        assert np.array_equal(got.astype(want.dtype), want, equal_nan=True), name
        checked += 1
    assert checked >= 38
    assert var.ReservoirIndex.size == 3 and var.LakeIndex.size == 3      # the off-channel sites were dropped


def test_lookupscalar():
    from lisflood_code_b200.hydrological_modules.reservoir import lookupscalar
    out = lookupscalar([[3, 1.5], [7, 2.5]], np.array([0., 3., 7., 9.]))
    assert np.isnan(out[0]) and out[1] == 1.5 and out[2] == 2.5 and np.isnan(out[3])
