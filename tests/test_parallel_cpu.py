"""Host logic of the LDD-cut multi-GPU path on CPU: partition invariants, the exchange plan (ghost sets, import blocks,
export targets), and world_size-2 / 3 gloo runs with the CPU oracle as compute stand-in that must reproduce the
single-process result BIT FOR BIT (owner -> owner cut edges in any direction, no hub rank)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT


def _case(rows=60, cols=48, seed=5, noise=0.3, maskf=0.1, kind="synthetic"):
    if kind == "real":      # the reference's own test catchment (tests/golden/kwreal_*.npz: 57 x 80, 2847 pixels, one outlet)
        with np.load(os.path.join(ROOT, "tests", "golden", "kwreal_etrs89_57x80.npz")) as g:
            return g["ldd"], g["mask"], g["alpha"], g["q0"], g["q"], g["dx"]
    from lisflood_code_b200 import synthetic
    ldd, mask = synthetic.random_ldd(rows, cols, seed=seed, noise=noise, mask_fraction=maskf, single_outlet=True)
    n = int(mask.sum())
    alpha, q0, q = synthetic.routing_fields(n, seed)
    dx = np.random.default_rng(seed).uniform(3000, 7000, n)
    return ldd[mask], mask, alpha, q0, q, dx


def _plan(ldd, mask, world, nsec=1, cap=8):
    from lisflood_code_b200.parallel import CutPlan, cut_edges_numpy, partition_numpy
    owner = partition_numpy(ldd, mask, world)
    eu, ed = cut_edges_numpy(ldd, mask, owner)
    return owner, CutPlan({"kw": (eu, ed, owner[eu], owner[ed], nsec, cap)}, world)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_invariants(world):
    from lisflood_code_b200.global_modules import ldd_ops
    ldd, mask, *_ = _case(90, 70, 9, 0.3, 0.05)
    owner, plan = _plan(ldd, mask, world)
    n = owner.size
    loads = np.bincount(owner, minlength=world)
    assert loads.sum() == n and owner.min() >= 0 and owner.max() < world
    ds = ldd_ops.downstream_index(ldd, mask)
    crossing = (ds >= 0) & (owner != owner[np.maximum(ds, 0)])
    eu = plan.g["kw"][0]
    assert set(np.flatnonzero(crossing)) == set(eu.tolist())
    if world > 1:
        assert loads.max() <= 1.6 * n / world + 64          # balance
        assert crossing.sum() > 0
        size = ldd_ops.accuflux(ds, np.ones(n))
        trunk = size > 0.25 * n / world
        assert trunk.sum() < 0.2 * n
        # the trunk is spread over the ranks: a trunk pixel sits with its largest tributary
        for p in np.flatnonzero(trunk)[:50]:
            ups = np.flatnonzero(ds == p)
            assert owner[p] == owner[ups[np.argmax(size[ups])]]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_exchange_plan_is_consistent(world):
    ldd, mask, *_ = _case(90, 70, 9, 0.3, 0.05)
    owner, plan = _plan(ldd, mask, world, nsec=2, cap=6)
    eu, ed, ou, od, nsec, cap = plan.g["kw"]
    seen = {}
    for r in range(world):
        P = plan.rank_plan("kw", r)
        loc = np.flatnonzero((owner == r) | np.isin(np.arange(owner.size), plan.ghosts[r]))
        x = plan.xslot("kw", r, loc)
        assert (x >= 0).sum() == P.n_export and ((x <= -2) & (x != -2 ** 31)).sum() == P.n_import
        assert np.array_equal(loc[x >= 0][np.argsort(x[x >= 0])], P.export_pixels)
        assert np.array_equal(owner[P.export_pixels], np.full(P.n_export, r))
        assert np.all(owner[P.import_pixels] != r)
        # the import block of this rank lies inside its region, blocks of 2 parities
        assert P.import_offset + 2 * P.n_import * nsec * cap <= plan.region_doubles[r]
        for k in range(P.n_export):
            c = int(P.export_peer[k])
            Pc = plan.rank_plan("kw", c)
            slot, rem = divmod(int(P.export_offset[k]) - Pc.import_offset, nsec * cap)
            assert rem == 0 and 0 <= slot < Pc.n_import
            assert Pc.import_pixels[slot] == P.export_pixels[k]          # producer and consumer agree on the slot
            assert P.export_parity_stride[k] == Pc.n_import * nsec * cap
            seen[(c, slot)] = seen.get((c, slot), 0) + 1
    assert len(seen) == eu.size and set(seen.values()) == {1}              # every import slot has exactly one producer


def _worker(rank, world, port, tmp, kind="synthetic"):
    """Stand-in for the device path: every rank routes its sub-mask (own pixels + ghosts) with the CPU oracle, ghost
    values prescribed; the ranks exchange the discharges of their cut edges through the plan's tables and repeat the
    step until nothing changes (a value is final once its upstream path has been exchanged across all its cuts)."""
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import lisf_oracle
    ldd, mask, alpha, q0, q, dx = _case(kind=kind)
    owner, plan = _plan(ldd, mask, world)
    n = owner.size
    keep = owner == rank
    keep[plan.ghosts[rank]] = True
    loc = np.flatnonzero(keep)
    sub = np.zeros(mask.shape, bool)
    sub[mask] = keep
    kw = lisf_oracle.KinematicWaveOracle(ldd[loc], sub, alpha[loc], 0.6, dx[loc], 3600.0)
    P = plan.rank_plan("kw", rank)
    x = plan.xslot("kw", rank, loc)
    ghost = (x <= -2).astype(np.uint8)
    ghost_pos = np.flatnonzero(x <= -2)
    ghost_slot = -2 - x[ghost_pos]
    exp_pos = np.flatnonzero(x >= 0)[np.argsort(x[x >= 0])]
    remote_slot = (P.export_offset - plan.import_offset["kw"][P.export_peer]) // plan.g["kw"][5]
    Q = np.ascontiguousarray(q0[loc])
    rng = np.random.default_rng(3)
    rounds_max = 0
    for step in range(15):
        scale = rng.uniform(0.5, 1.5)
        Qold, imported, rounds = Q.copy(), np.zeros(max(P.n_import, 1)), 0
        while True:
            rounds += 1
            Q = Qold.copy()
            fv = np.zeros(loc.size)
            fv[ghost_pos] = imported[ghost_slot]
            kw.kinematicWaveRouting(Q, q[loc] * scale, fixed=ghost if P.n_import else None, fixed_values=fv if P.n_import else None)
            out = [None] * world
            dist.all_gather_object(out, [(int(P.export_peer[k]), int(remote_slot[k]), float(Q[exp_pos[k]])) for k in range(P.n_export)])
            new = imported.copy()
            for msgs in out:
                for c, slot, val in msgs:
                    if c == rank:
                        new[slot] = val
            changed = [None] * world
            dist.all_gather_object(changed, bool(np.any(new != imported)))
            imported = new
            if not any(changed):
                break
        rounds_max = max(rounds_max, rounds)
    parts = [None] * world
    dist.all_gather_object(parts, (loc[owner[loc] == rank], Q[owner[loc] == rank]))
    if rank == 0:
        full = np.empty(n)
        for idx, val in parts:
            full[idx] = val
        np.save(os.path.join(tmp, "dist_w%d.npy" % world), full)
        np.save(os.path.join(tmp, "meta_w%d.npy" % world), np.array([plan.g["kw"][0].size, rounds_max]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "synthetic"), (3, "synthetic"), (2, "real"), (3, "real")])
def test_gloo_cut_network_is_bit_identical(tmp_path, world, kind, oracle):
    """kind "real": the reference's own test catchment cut over the ranks -- its tests/test_subcatchments.py:111-112 asks
    sub-domain results to be bit-identical to the full run."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path), kind), nprocs=world, join=True)
    got = np.load(tmp_path / ("dist_w%d.npy" % world))
    ncut, rounds = np.load(tmp_path / ("meta_w%d.npy" % world))
    assert ncut > 0, "the test catchment must actually be cut"
    ldd, mask, alpha, q0, q, dx = _case(kind=kind)
    ora = oracle.KinematicWaveOracle(ldd, mask, alpha, 0.6, dx, 3600.0)
    Q = q0.copy()
    rng = np.random.default_rng(3)
    for step in range(15):
        ora.kinematicWaveRouting(Q, q * rng.uniform(0.5, 1.5))
    assert np.array_equal(got, Q)


def test_plan_with_two_graphs_marks_inert_ghosts():
    """The full model cuts two graphs (LddToChan, LddKinematic) with ONE pixel set per rank: a ghost that has a link only in
    one of them is inert in the other; import blocks of the two graphs follow each other inside a rank's region."""
    from lisflood_code_b200 import synthetic
    from lisflood_code_b200.parallel import INERT, CutPlan, cut_edges_numpy, partition_numpy
    S = synthetic.full_stack(80, 70, seed=3, ldd_noise=0.4, mask_fraction=0.1)
    world = 3
    owner = partition_numpy(S["Ldd"], S["mask"], world, subtree_fraction=0.02)
    graphs = {}
    for name, ldd, nsec, cap in (("overland", S["LddToChan"], 3, 1), ("channel", S["LddKinematic"], 1, 24)):
        eu, ed = cut_edges_numpy(ldd, S["mask"], owner)
        graphs[name] = (eu, ed, owner[eu], owner[ed], nsec, cap)
    plan = CutPlan(graphs, world, order=("overland", "channel"))
    assert graphs["overland"][0].size > 0 and graphs["channel"][0].size > 0
    for r in range(world):
        keep = owner == r
        keep[plan.ghosts[r]] = True
        loc = np.flatnonzero(keep)
        xo, xc = plan.xslot("overland", r, loc), plan.xslot("channel", r, loc)
        ghost = ~(owner[loc] == r)
        assert np.all((xo[ghost] <= -2)) and np.all(xc[ghost] <= -2)            # ghosts are never solved, in either graph
        assert np.all(xo[~ghost] >= -1) and np.all(xc[~ghost] >= -1)
        live_o, live_c = (xo <= -2) & (xo != INERT), (xc <= -2) & (xc != INERT)
        assert np.all(live_o | live_c == ghost)                                  # every ghost has a link in at least one graph
        assert (xo == INERT).sum() == ghost.sum() - live_o.sum()
        Po, Pc = plan.rank_plan("overland", r), plan.rank_plan("channel", r)
        assert Po.import_offset == 0 and Pc.import_offset == 2 * Po.n_import * 3 * 1
        assert plan.region_doubles[r] == Pc.import_offset + 2 * Pc.n_import * 1 * 24
