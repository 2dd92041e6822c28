"""Pre-run (InitLisflood) through the device path producing AvgDis / LZAvInflowMap, then the warm-start chain through the
theta-based end maps: every warm-started step reproduces the long run (reference: tests/test_warmstart.py:114-141; its
comparators: 6 significant digits for time series, a tolerance for maps -- asserted here: 1e-9 relative, observed ~1e-15:
W -> theta -> W and Q -> volume -> Q round trips are exact only to the last bits)."""
import numpy as np
import pytest

from warmstart_common import check_chain, run_chain

pytestmark = pytest.mark.gpu


def test_prerun_and_warmstart_chain_gpu(gpu_lib, oracle):
    from lisflood_code_b200.hotpath import HotPathModel
    long_dis, warm_dis, avgdis, lzavin = run_chain(lambda S: HotPathModel(S, diagnostics=False))
    worst = check_chain(long_dis, warm_dis)
    print("warm-start chain (device): worst rel. deviation from the long run %.2e" % worst)
    # the device pre-run products agree with the CPU restatement's
    from test_warmstart import _OracleAsModel
    _, _, avgdis_cpu, lzavin_cpu = run_chain(_OracleAsModel, nsteps=1)
    assert np.max(np.abs(avgdis - avgdis_cpu) / np.maximum(np.abs(avgdis_cpu), 1e-12)) < 1e-9
    assert np.max(np.abs(lzavin - lzavin_cpu) / np.maximum(np.abs(lzavin_cpu), 1e-12)) < 1e-9
