"""Host logic of the LDD-cut multi-GPU path on CPU: partition invariants, and a world_size-2 (and 3) gloo run
with the CPU oracle as compute stand-in that must reproduce the single-process result BIT FOR BIT."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT


def _case(rows=60, cols=48, seed=5, noise=0.3, maskf=0.1):
    from lisflood_code_b200 import synthetic
    ldd, mask = synthetic.random_ldd(rows, cols, seed=seed, noise=noise, mask_fraction=maskf, single_outlet=True)
    n = int(mask.sum())
    alpha, q0, q = synthetic.routing_fields(n, seed)
    dx = np.random.default_rng(seed).uniform(3000, 7000, n)
    return ldd[mask], mask, alpha, q0, q, dx


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_invariants(world):
    from lisflood_code_b200.parallel import Partition
    ldd, mask, *_ = _case(90, 70, 9, 0.3, 0.05)
    P = Partition(ldd, mask, world)
    n = P.n
    assert sum(P.loads) == n and P.owner.min() >= 0 and P.owner.max() < world
    ds = P.downstream
    has = ds >= 0
    crossing = has & (P.owner != P.owner[np.maximum(ds, 0)])
    # every link that crosses ranks goes INTO rank 0 (the trunk) and is a registered cut edge
    assert np.all(P.owner[ds[crossing]] == 0)
    cut_all = np.concatenate([c for c in P.cut_pixels]) if world > 1 else np.array([], int)
    assert set(np.flatnonzero(crossing)) == set(cut_all.tolist())
    if world > 1:
        assert max(P.loads) <= 1.6 * n / world + 64          # balance
        assert P.trunk.sum() < 0.2 * n
    # xslot bookkeeping
    for r in range(world):
        x = P.local_xslot(r)
        assert (x >= 0).sum() == (P.n_cut[r] if r else 0)
        assert (x <= -2).sum() == (P.n_import if r == 0 else 0)


class CpuStandInBackend(object):
    """Router protocol of lisflood_code_b200.parallel.GpuRouterBackend on top of the CPU oracle (test stand-in:
    it lets the partition / exchange host logic run under gloo without a GPU)."""
    device = "cpu"

    def __init__(self, ldd_local, sub_mask, alpha, beta, dx, dt, xslot, n_exp, n_imp, export, imported, max_steps, world):
        from oracle import lisf_oracle
        self.kw = lisf_oracle.KinematicWaveOracle(ldd_local, sub_mask, alpha, beta, dx, dt)
        self.Q = np.zeros(xslot.size)
        self.q = np.zeros(xslot.size)
        self.cap, self.n_exp, self.n_imp = max_steps, n_exp, n_imp
        self.exp = export.numpy().reshape(-1, max_steps)
        self.imp = imported.numpy().reshape(-1, max_steps)
        self.fixed = (xslot <= -2).astype(np.uint8)
        self.ghost_slot = np.where(xslot <= -2, -2 - xslot, 0)
        self.export_idx = np.flatnonzero(xslot >= 0)
        self.export_slot = xslot[self.export_idx]

    def set_discharge(self, q):
        self.Q = q.copy()

    def set_lateral_inflow(self, q):
        self.q = q.copy()

    def run(self, nsteps, inflow_scale):
        for s in range(nsteps):
            q = self.q if inflow_scale is None else self.q * inflow_scale[s]
            fv = self.imp[self.ghost_slot, s] if self.n_imp else None
            self.kw.kinematicWaveRouting(self.Q, q, fixed=self.fixed if self.n_imp else None, fixed_values=fv)
            if self.n_exp:
                self.exp[self.export_slot, s] = self.Q[self.export_idx]

    def get_discharge(self):
        return self.Q.copy()

    def before_send(self):
        pass

    def after_recv(self):
        pass


def _worker(rank, world, port, tmp):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lisflood_code_b200.parallel import DistributedKinematicWave
    ldd, mask, alpha, q0, q, dx = _case()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_parallel_cpu import CpuStandInBackend
    D = DistributedKinematicWave(ldd, mask, alpha, 0.6, dx, 3600.0, max_steps=8, backend=CpuStandInBackend)
    D.set_discharge(q0)
    D.set_lateral_inflow(q)
    rng = np.random.default_rng(3)
    for chunk in range(3):
        D.run(5, inflow_scale=rng.uniform(0.5, 1.5, 5))
    out = D.gather_discharge()
    if rank == 0:
        np.save(os.path.join(tmp, "dist_w%d.npy" % world), out)
        np.save(os.path.join(tmp, "cuts_w%d.npy" % world), np.array(D.part.n_cut))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_cut_network_is_bit_identical(tmp_path, world, oracle):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / ("dist_w%d.npy" % world))
    cuts = np.load(tmp_path / ("cuts_w%d.npy" % world))
    assert cuts[1:].sum() > 0, "the test catchment must actually be cut"
    ldd, mask, alpha, q0, q, dx = _case()
    ora = oracle.KinematicWaveOracle(ldd, mask, alpha, 0.6, dx, 3600.0)
    Q = q0.copy()
    rng = np.random.default_rng(3)
    for chunk in range(3):
        sc = rng.uniform(0.5, 1.5, 5)
        for s_ in range(5):
            ora.kinematicWaveRouting(Q, q * sc[s_])
    assert np.array_equal(got, Q)
