"""torchrun --nproc-per-node N tools/run_dist_check.py : LDD-cut routing over NCCL must reproduce the single-GPU
router bit for bit (and the CPU oracle to 1e-9)."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
from lisflood_code_b200 import _capi, synthetic  # noqa: E402
from lisflood_code_b200.hydrological_modules.kinematic_wave_parallel import kinematicWave  # noqa: E402
from lisflood_code_b200.parallel import DistributedKinematicWave  # noqa: E402

_capi.check(_capi.lib().lf_device_init(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for (rows, cols, noise, maskf, beta, single) in ((300, 260, 0.3, 0.1, 0.6, True), (500, 400, 2.0, 0.0, 0.6, False),
                                                (240, 300, 0.4, 0.05, 0.8, True), (1200, 900, 0.3, 0.0, 0.6, True)):
    ldd2, mask = synthetic.random_ldd(rows, cols, seed=77, noise=noise, mask_fraction=maskf, single_outlet=single)
    ldd = ldd2[mask]
    n = int(mask.sum())
    alpha, q0, q = synthetic.routing_fields(n, 77)
    dx = np.random.default_rng(7).uniform(3000, 7000, n)
    D = DistributedKinematicWave(ldd, mask, alpha, beta, dx, 3600.0, max_steps=16)
    D.set_discharge(q0)
    D.set_lateral_inflow(q)
    rng = np.random.default_rng(5)
    scales = [rng.uniform(0.5, 1.5, 12) for _ in range(3)]
    for sc in scales:
        D.run(12, inflow_scale=sc)
    out = D.gather_discharge()
    if rank == 0:
        kw = kinematicWave(ldd, mask, alpha, beta, dx, 3600.0)
        kw.set_discharge(q0)
        kw.set_lateral_inflow(q)
        for sc in scales:
            kw.run(12, inflow_scale=sc)
        ref = kw.get_discharge()
        same = np.array_equal(out, ref)
        from oracle import lisf_oracle
        ora = lisf_oracle.KinematicWaveOracle(ldd, mask, alpha, beta, dx, 3600.0)
        Q = q0.copy()
        for sc in scales:
            for s in range(12):
                ora.kinematicWaveRouting(Q, q * sc[s])
        err = float(np.max(np.abs(out - Q) / np.maximum(np.abs(Q), 1e-12)))
        print("case %dx%d beta %.1f: loads %s cut edges %s trunk %d | bit-identical to 1 GPU: %s | vs oracle %.2e" % (
            rows, cols, beta, D.part.loads, D.part.n_cut, int(D.part.trunk.sum()), same, err), flush=True)
        ok = ok and same and err < 1e-9 and (not single or world == 1 or sum(D.part.n_cut) > 0)
dist.barrier()
if rank == 0:
    print("DIST CHECK", "PASSED" if ok else "FAILED", flush=True)
dist.destroy_process_group()
