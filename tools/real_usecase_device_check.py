"""python tools/real_usecase_device_check.py [fixture name]

The device path on the reference's own test catchment (fixture tests/golden/realcase_*.npz: real static maps through the
init mirrors, real float32 meteo maps, LAI by interval): feeder kernel + soil + overland + split channel routing per step,
compared with the CPU restatement's recorded run (1e-8 relative) and with the soil-moisture maps of the output stacks the
reference SHIPS for that run (1e-6 absolute).  Prints one line per step and "REAL USECASE PASSED" / "FAILED".
Run as its own process by tests/test_gpu_real_usecase.py."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_cases, rel_err  # noqa: E402
from realcase_common import THETA, load  # noqa: E402
from lisflood_code_b200.hotpath import HotPathModel  # noqa: E402


def main():
    cases = sys.argv[1:] or golden_cases("realcase_")
    ok = bool(cases)
    for case in cases:
        S, P, state, raw, days, lai, want, shipped = load(case)
        M = HotPathModel(S, diagnostics=True)
        M.set_feeder(P, state)
        for t in range(len(raw)):
            M.set_lai(lai[t])
            M.feed(raw[t], days[t])
            M.step()
            worst = {k: rel_err(M.get(k, 3 if w.ndim == 2 else 1), w) for k, w in want[t].items()}
            theta = {name: float(np.abs(M.get(attr, 3)[row] - shipped[t][name]).max()) for name, (attr, row) in THETA.items()}
            bad = {k: v for k, v in worst.items() if not v < 1e-8}
            bad.update({k: v for k, v in theta.items() if not v < 1e-6})
            print("%s step %d: vs restatement max rel %.2e (%s); vs shipped theta max abs %.2e%s" % (
                case, t, max(worst.values()), max(worst, key=worst.get), max(theta.values()),
                "  MISMATCH %s" % bad if bad else ""), flush=True)
            ok = ok and not bad
    print("REAL USECASE PASSED" if ok else "REAL USECASE FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
