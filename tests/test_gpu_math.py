"""On-device accuracy of the hand-written float64 math used by the hot kernels (lf_math.cuh, lf_kw_solve.cuh):
Newton division / square root on the hardware reciprocal seeds, table-driven x^y and e^x, fifth / cube roots.
Compared on the GPU against the CUDA math library (<= 2 ulp) over 2^22 pseudo-random arguments."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_math_accuracy(gpu_lib):
    from lisflood_code_b200 import _capi
    err = np.zeros(8, np.float64)
    _capi.check(_capi.lib().lf_math_selftest(1 << 22, 12345, err))
    names = ["div_nr", "sqrt_nr", "pw_tab", "exp_neg_tab", "root5", "root3", "pw", "van_genuchten_term"]
    got = dict(zip(names, err.tolist()))
    eps = 2.0 ** -52
    limits = {"div_nr": 2 * eps, "sqrt_nr": 2 * eps, "pw_tab": 4 * eps, "exp_neg_tab": 4 * eps, "root5": 8 * eps,
              "root3": 8 * eps, "pw": 4 * eps, "van_genuchten_term": 1e-9}
    bad = {k: v for k, v in got.items() if not v <= limits[k]}
    assert not bad, (bad, got)
