"""python -m torch.distributed.run --nproc-per-node N tools/run_dist_check.py [--stress]

ONE raster cut along its drainage graph over N processes (one per GPU; with fewer GPUs than processes the ranks share a
device and torch.distributed falls back to gloo for the plumbing) must reproduce the single-GPU result BIT FOR BIT:
  * routing only (DistributedKinematicWave vs kinematicWave), several networks, beta 0.6 and 0.8;
  * the full model (DistributedHotPathModel vs HotPathModel): soil -> overland -> channel sub-steps, single and split;
  * the device partitioner against its NumPy restatement.
--stress: many short runs with rank 0 artificially slowed down (flow control of the exchange regions).
Prints one line per case and "DIST CHECK PASSED" / "FAILED" (rank 0)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
ndev = torch.cuda.device_count()
device = local % ndev
torch.cuda.set_device(device)
from lisflood_code_b200 import _capi, synthetic  # noqa: E402
from lisflood_code_b200.hotpath import HotPathModel  # noqa: E402
from lisflood_code_b200.hydrological_modules.kinematic_wave_parallel import kinematicWave  # noqa: E402
from lisflood_code_b200.parallel import DistributedHotPathModel, DistributedKinematicWave, partition_numpy  # noqa: E402

_capi.check(_capi.lib().lf_device_init(device))
backend = "nccl" if ndev >= world else "gloo"
if backend == "nccl":
    dist.init_process_group("nccl", device_id=torch.device("cuda", device))
else:
    dist.init_process_group("gloo")
stress = "--stress" in sys.argv
ok = True
say = lambda *a: print(*a, flush=True) if rank == 0 else None
say("dist check: world %d on %d device(s), plumbing over %s" % (world, ndev, backend))

# ---- routing only --------------------------------------------------------------------------------------------
cases = ((300, 260, 0.3, 0.1, 0.6, True), (500, 400, 2.0, 0.0, 0.6, False), (240, 300, 0.4, 0.05, 0.8, True),
         (1200, 900, 0.3, 0.0, 0.6, True))
for (rows, cols, noise, maskf, beta, single) in (cases[:1] if stress else cases):
    ldd2, mask = synthetic.random_ldd(rows, cols, seed=77, noise=noise, mask_fraction=maskf, single_outlet=single)
    ldd = ldd2[mask]
    n = int(mask.sum())
    alpha, q0, q = synthetic.routing_fields(n, 77)
    dx = np.random.default_rng(7).uniform(3000, 7000, n)
    D = DistributedKinematicWave(ldd, mask, alpha, beta, dx, 3600.0, max_steps=16)
    D.set_discharge(q0)
    D.set_lateral_inflow(q)
    rng = np.random.default_rng(5)
    nruns, nsteps = (60, 3) if stress else (3, 12)
    scales = [rng.uniform(0.5, 1.5, nsteps) for _ in range(nruns)]
    for k, sc in enumerate(scales):
        if stress and rank == 0 and k % 7 == 3:
            time.sleep(0.05)                     # a slow rank: the others must not overwrite what it has not read yet
        D.run(nsteps, inflow_scale=sc)
    out = D.gather_discharge()
    aborted, epochs = D.status()
    if rank == 0:
        owner_np = partition_numpy(ldd, mask, world) if n <= 120000 else None
        kw = kinematicWave(ldd, mask, alpha, beta, dx, 3600.0)
        kw.set_discharge(q0)
        kw.set_lateral_inflow(q)
        for sc in scales:
            kw.run(nsteps, inflow_scale=sc)
        ref = kw.get_discharge()
        same = bool(np.array_equal(out, ref))
        from oracle import lisf_oracle
        ora = lisf_oracle.KinematicWaveOracle(ldd, mask, alpha, beta, dx, 3600.0)
        Q = q0.copy()
        for sc in scales:
            for s in range(nsteps):
                ora.kinematicWaveRouting(Q, q * sc[s])
        err = float(np.max(np.abs(out - Q) / np.maximum(np.abs(Q), 1e-12)))
        part_ok = True
        if owner_np is not None:
            loads_np = np.bincount(owner_np, minlength=world).tolist()
            part_ok = loads_np == D.loads
        say("routing %dx%d beta %.1f: loads %s cut edges %d (exports of rank 0: %d, imports: %d) trunk %d roots %d | "
            "bit-identical to 1 GPU: %s | vs oracle %.2e | partition == NumPy: %s | aborted %s, runs %d" % (
                rows, cols, beta, D.loads, D.cut_edges, D.exports, D.imports, D.n_trunk, D.n_roots, same, err, part_ok,
                aborted, epochs))
        ok = ok and same and err < 1e-9 and part_ok and not aborted and (not single or world == 1 or D.cut_edges > 0)
    D.close()
    dist.barrier()

# ---- full model ------------------------------------------------------------------------------------------------
model_cases = ((150, 120, False, 7, 0), (130, 160, True, 8, 0), (140, 150, True, 91, 11)) if not stress else ((90, 80, True, 9, 0),)
for (rows, cols, split, seed, nstruct) in model_cases:
    S = synthetic.full_stack(rows, cols, seed=seed, split_routing=split, ldd_noise=0.4, mask_fraction=0.1,
                             **({"channel_threshold": 12, "dt_sec": 21600.0} if nstruct else {}))
    if nstruct:                      # reservoirs and lakes in the sub-step loop, on a cut raster
        synthetic.add_structures(S, nstruct // 2 + 1, nstruct // 2, seed=seed)
    M = DistributedHotPathModel(S, diagnostics=False, subtree_fraction=0.02)   # small sub-trees: the basins of these small rasters get cut
    nsteps = 12 if stress else 3
    for t in range(nsteps):
        if stress and rank == 0 and t % 3 == 1:
            time.sleep(0.05)
        M.step(synthetic.forcing(S, t, seed))
    keys = [("ChanQAvg", 1), ("ChanQKin", 1), ("ChanM3Kin", 1), ("OFQOther", 1), ("OFQDirect", 1), ("W1a", 3), ("UZ", 3),
            ("LZ", 1), ("sumDis", 1)] + ([("Chan2QKin", 1), ("Chan2M3Kin", 1)] if split else []) + \
        ([("ReservoirStorageM3", 1), ("ReservoirFill", 1), ("QResOutM3Dt", 1), ("LakeStorageM3", 1), ("LakeOutflow", 1)] if nstruct else [])
    got = {k: M.gather(k, r) for k, r in keys}
    aborted, epochs = M.status()
    summ = M.plan.summary()
    if rank == 0:
        R = HotPathModel(S, diagnostics=False)
        for t in range(nsteps):
            R.step(synthetic.forcing(S, t, seed))
        diff = [k for k, r in keys if not np.array_equal(got[k], R.get(k, r))]
        say("model %dx%d split=%s structures=%d: loads %s, cut edges overland %d channel %d | maps that differ from 1 GPU: %s | aborted %s" % (
            rows, cols, split, nstruct, M.loads, summ["overland"]["cut_edges"], summ["channel"]["cut_edges"], diff or "none", aborted))
        ok = ok and not diff and not aborted and (world == 1 or summ["channel"]["cut_edges"] + summ["overland"]["cut_edges"] > 0)
        R.close()
    M.close()
    dist.barrier()

if rank == 0:
    print("DIST CHECK", "PASSED" if ok else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok or rank != 0 else 1)
