"""Prints the max relative error of every golden map of the model cases (GPU vs reference goldens)."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_cases, golden_model, rel_err  # noqa: E402
from lisflood_code_b200 import _capi  # noqa: E402
from lisflood_code_b200.hotpath import HotPathModel  # noqa: E402

_capi.check(_capi.lib().lf_device_init(0))
for case in golden_cases("model_"):
    S, F, O = golden_model(case)
    M = HotPathModel(S, diagnostics=True)
    print("==", case, M.info())
    for t in range(len(F)):
        M.step(F[t])
        errs = {}
        for k, want in O[t].items():
            try:
                errs[k] = rel_err(M.get(k, 3 if want.ndim == 2 else 1), want)
            except Exception as e:  # noqa: BLE001
                errs[k] = str(e)[:60]
        bad = {k: v for k, v in errs.items() if not (isinstance(v, float) and v < 1e-8)}
        worst = max((v for v in errs.values() if isinstance(v, float)), default=0.0)
        print(" step", t, "maps", len(errs), "worst %.2e" % worst, "BAD:" if bad else "ok", bad if bad else "")
