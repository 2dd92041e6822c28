"""Oracle vs the live reference (only where /root/reference exists, i.e. the build container)."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@pytest.mark.parametrize("rows,cols,noise,maskf,seed", [(120, 90, 3.0, 0.0, 1), (90, 130, 0.3, 0.2, 2), (200, 200, 0.1, 0.0, 3)])
def test_oracle_equals_reference(oracle, rows, cols, noise, maskf, seed):
    from lisflood_code_b200 import synthetic
    kwpt, kwp, sl = ref_loader.load()
    ldd, mask = synthetic.random_ldd(rows, cols, seed=seed, noise=noise, mask_fraction=maskf)
    n = int(mask.sum())
    alpha, q0, q = synthetic.routing_fields(n, seed)
    dx = np.random.default_rng(seed).uniform(3000, 7000, n)
    ref = kwp.kinematicWave(ldd[mask].copy(), mask, alpha, 0.6, dx, 3600.0)
    ora = oracle.KinematicWaveOracle(ldd[mask].copy(), mask, alpha, 0.6, dx, 3600.0)
    for k in ("downstream_lookup", "upstream_lookup", "num_upstream_pixels", "pixels_ordered", "order_start_stop"):
        assert np.array_equal(getattr(ref, k), getattr(ora, k)), k
    Qr, Qo = q0.copy(), q0.copy()
    for s in range(10):
        ref.kinematicWaveRouting(Qr, q)
        ora.kinematicWaveRouting(Qo, q)
    assert rel_err(Qo, Qr) < 1e-11
