"""Times one model step of the UNMODIFIED reference modules (Numba kernels + NumPy glue, oracle/ref_modules.py; build
container only) next to the C/OpenMP + NumPy port that bench.py uses as the CPU baseline on the GPU box, on the same
synthetic catchment.  Round 1, 8 host threads, 600x600: reference 1.95-2.07e5, port 2.01e5 cell-updates/s."""
import sys, time, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
from lisflood_code_b200 import synthetic
from oracle import ref_modules, lisf_oracle, lisf_oracle_model as om
import numba
R = int(sys.argv[1]) if len(sys.argv) > 1 else 600
S = synthetic.full_stack(R, R, seed=300, ldd_noise=0.5)
print("N", S["N"], "numba threads", numba.get_num_threads(), "cpus", os.cpu_count(), flush=True)
t0 = time.perf_counter(); M = ref_modules.RefModel(S); print("ref init %.1f s" % (time.perf_counter() - t0), flush=True)
for t in range(3):
    F = synthetic.forcing(S, t, 300)
    t0 = time.perf_counter(); M.step(F); dt = time.perf_counter() - t0
    print("reference step %d: %.2f s -> %.3e cell-updates/s" % (t, dt, S["N"] / dt), flush=True)
O = om.OracleModel(S)
for thr in (1, 8):
    lisf_oracle.set_threads(thr)
    for t in range(2):
        F = synthetic.forcing(S, t, 300)
        t0 = time.perf_counter(); O.step(F); dt = time.perf_counter() - t0
    print("port (%d threads) step: %.2f s -> %.3e cell-updates/s" % (thr, dt, S["N"] / dt), flush=True)
