"""Pre-run products and the warm-start chain through the end maps with the CPU restatement of the step
as the model -- the host logic (lisflood_code_b200/state_io.py + the init chain) without a GPU.  The same chain runs on
the device in tests/test_gpu_warmstart.py."""
import numpy as np

from warmstart_common import check_chain, run_chain


class _OracleAsModel(object):
    def __init__(self, S):
        from oracle import lisf_oracle_model as om
        self.O = om.OracleModel(S)
        self.acc = False
        self.O.var.CumQ = np.zeros(S["N"])

    def set_option(self, name, value):
        if name == "accumulate_discharge":
            self.acc = bool(value)

    def step(self, F):
        self.O.step(F)
        if self.acc:
            self.O.var.CumQ = self.O.var.CumQ + self.O.var.ChanQ     # Lisflood_dynamic.py:225

    def get(self, name, rows=None):
        return np.asarray(getattr(self.O.var, name), np.float64)


def test_prerun_and_warmstart_chain_cpu(oracle):
    long_dis, warm_dis, avgdis, lzavin = run_chain(_OracleAsModel)
    worst = check_chain(long_dis, warm_dis)
    assert avgdis.shape == long_dis[0].shape and lzavin.min() >= 0
    print("warm-start chain (CPU restatement): worst rel. deviation from the long run %.2e" % worst)
