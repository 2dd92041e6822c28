"""LDD-cut domain decomposition of ONE catchment raster across the GPUs of a node (BASELINE.json
configs C3 cut / C4), one process per GPU.

Information in a drainage network flows strictly downstream (a pixel needs only the NEW discharge of its upstream
pixels, reference: hydrological_modules/kinematic_wave_parallel_tools.py:57-58), whole catchments are independent
(the reference proves sub-mask runs bit-identical, tests/test_subcatchments.py:111-112) and the soil / stencil stages
have no neighbour access at all.  So the raster is cut along the drainage GRAPH, not along raster tiles:

  * pixels whose upstream area exceeds `subtree_fraction * N / world` form the trunk (the main stems, a tiny fraction
    of the pixels); what hangs off the trunk -- or drains straight to an outlet -- is a sub-tree;
  * sub-trees are bin-packed over the ranks, largest first; a trunk pixel joins the rank of its largest tributary, so
    the trunk is spread over the ranks too and cut edges go from owner to owner in any direction (no hub rank);
  * a pixel's soil state lives with its routing partition: the per-cell stages never communicate;
  * for every cut link u -> d the rank of d keeps a GHOST of u.  The routing kernels exchange its discharge
    themselves, value by value, over NVLink (csrc/lf_xchg.cuh): the thread that solves (u, step) stores the result
    into the consumer's exchange region, the ghost work item polls its slot.  Every rank lays its pixels out by the
    GLOBAL routing level (lf_graph_restrict), so the wavefront diagonals of all ranks line up, the consumer of a value
    always sits one diagonal after its producer, and the cut network reproduces the uncut one BIT FOR BIT (same
    upstream slots, same summation order).  torch.distributed (NCCL) is the plumbing around it: CUDA-IPC handle
    exchange, barriers, gathers of results and the max-over-ranks timing.

This file holds the host logic: the (NumPy) reference partitioner and the exchange plan (pure host code, covered by
world-size-2/3 gloo tests in tests/test_parallel_cpu.py with a stand-in router), and the device objects
DistributedKinematicWave (routing only, C4) and DistributedHotPathModel (the full stack, C3 cut).
"""
import ctypes as C
import heapq

import numpy as np

from .global_modules import ldd_ops

INERT = -2 ** 31          # xslot of a ghost that has no link in that graph (csrc/lf_xchg.cuh)
HEADER_DOUBLES = 512      # header of an exchange region (lfx::HEADER_BYTES / 8)


# ------------------------------------------------------------------------------------------------------------------
# reference partitioner (NumPy): what lf_graph_partition computes on the device, for the CPU tests and as its check
# ------------------------------------------------------------------------------------------------------------------
def partition_numpy(ldd_codes, land_mask, world, subtree_fraction=0.25):
    """owner int32[N] of every pixel; deterministic.  Same rule as lf_graph_partition (csrc/lf_graph.cu)."""
    mask = np.asarray(land_mask, bool)
    n = int(mask.sum())
    ds = ldd_ops.downstream_index(np.asarray(ldd_codes, np.float64), mask)
    size = ldd_ops.accuflux(ds, np.ones(n))
    order, hops = ldd_ops.topological_order(ds)          # farthest from the outlets first
    thr = max(1.0, subtree_fraction * n / world) if world > 1 else np.inf
    trunk = size > thr
    root = (~trunk) & ((ds < 0) | trunk[np.maximum(ds, 0)])
    label = np.where(root, np.arange(n), -1)
    for p in order[::-1]:                                # outlets first
        if not trunk[p] and not root[p]:
            label[p] = label[ds[p]]
    roots = np.flatnonzero(root)
    idx = sorted(range(roots.size), key=lambda k: (-size[roots[k]], roots[k]))
    heap = [(0.0, r) for r in range(world)]
    heapq.heapify(heap)
    own_root = np.zeros(roots.size, np.int32)
    for k in idx:
        load, r = heapq.heappop(heap)
        own_root[k] = r
        heapq.heappush(heap, (load + size[roots[k]], r))
    owner = np.zeros(n, np.int32)
    tmp = np.zeros(n, np.int32)
    tmp[roots] = own_root
    owner[~trunk] = tmp[label[~trunk]]
    if trunk.any():
        ups = [[] for _ in range(n)]
        for p in range(n):                               # ascending pixel index = the reference's slot order
            if ds[p] >= 0:
                ups[ds[p]].append(p)
        for p in order:                                  # headwaters first: tributaries carry their owner already
            if trunk[p]:
                best, own = -1.0, 0
                for u in ups[p]:
                    if size[u] > best:
                        best, own = size[u], owner[u]
                owner[p] = own
    return owner


def cut_edges_numpy(ldd_codes, land_mask, owner):
    """(edge_u, edge_d): links whose ends have different owners, sorted by u (what lf_graph_cut_edges returns)."""
    ds = ldd_ops.downstream_index(np.asarray(ldd_codes, np.float64), np.asarray(land_mask, bool))
    u = np.flatnonzero((ds >= 0) & (owner != owner[np.maximum(ds, 0)]))
    return u.astype(np.int32), ds[u].astype(np.int32)


# ------------------------------------------------------------------------------------------------------------------
# exchange plan: pure host logic on the (small) cut-edge lists
# ------------------------------------------------------------------------------------------------------------------
class GraphPlan(object):
    """Exchange tables of one graph for one rank (arguments of lf_router_set_exchange / lf_model_set_exchange)."""

    def __init__(self):
        self.export_pixels = self.import_pixels = None   # global compressed indices, ascending
        self.export_peer = self.export_offset = self.export_parity_stride = None
        self.n_export = self.n_import = 0
        self.import_offset = 0


class CutPlan(object):
    """From the cut edges of every graph (sorted by u) and the owners of their two ends: per-rank ghost sets, import
    blocks inside the exchange regions and export targets.  Identical on every rank.

    graphs: {name: (edge_u, edge_d, owner_u, owner_d, nsec, cap)} -- nsec values per edge and step (main channel +
    floodplain; the three overland routers), cap steps per run."""

    def __init__(self, graphs, world, order=None):
        self.world = int(world)
        self.names = list(order or sorted(graphs))
        self.g = {}
        for name in self.names:
            eu, ed, ou, od, nsec, cap = graphs[name]
            eu, ed, ou, od = (np.asarray(a, np.int64) for a in (eu, ed, ou, od))
            assert np.all(np.diff(eu) > 0), "cut edges must be sorted by their upstream end (one link per pixel)"
            assert np.all(ou != od)
            self.g[name] = (eu, ed, ou, od, int(nsec), int(cap))
        # ghosts of rank r: upstream ends of its incoming edges, over all graphs
        self.ghosts = []
        for r in range(self.world):
            gs = [self.g[nm][0][self.g[nm][3] == r] for nm in self.names]
            self.ghosts.append(np.unique(np.concatenate(gs)) if gs else np.zeros(0, np.int64))
        # layout of the regions: per rank, the import blocks of the graphs one after the other
        self.import_offset = {nm: np.zeros(self.world, np.int64) for nm in self.names}
        self.n_import = {nm: np.zeros(self.world, np.int64) for nm in self.names}
        self.region_doubles = np.zeros(self.world, np.int64)
        for r in range(self.world):
            off = 0
            for nm in self.names:
                eu, ed, ou, od, nsec, cap = self.g[nm]
                k = int((od == r).sum())
                self.import_offset[nm][r] = off
                self.n_import[nm][r] = k
                off += 2 * k * nsec * cap
            self.region_doubles[r] = off

    def rank_plan(self, name, rank):
        eu, ed, ou, od, nsec, cap = self.g[name]
        P = GraphPlan()
        imp = od == rank
        exp = ou == rank
        P.import_pixels = eu[imp]
        P.export_pixels = eu[exp]
        P.n_import, P.n_export = int(imp.sum()), int(exp.sum())
        P.import_offset = int(self.import_offset[name][rank])
        peers = od[exp]
        # index of every exported edge among its consumer's imports (both sorted by u)
        remote = np.zeros(P.n_export, np.int64)
        for c in np.unique(peers):
            theirs = eu[od == c]
            mine = peers == c
            remote[mine] = np.searchsorted(theirs, P.export_pixels[mine])
        P.export_peer = peers.astype(np.int32)
        P.export_offset = (self.import_offset[name][peers] + remote * nsec * cap).astype(np.int64)
        P.export_parity_stride = (self.n_import[name][peers] * nsec * cap).astype(np.int64)
        return P

    def xslot(self, name, rank, local_pixels):
        """int32[n_local] for the local pixel list (global indices, ascending)."""
        P = self.rank_plan(name, rank)
        loc = np.asarray(local_pixels, np.int64)
        x = -np.ones(loc.size, np.int32)
        x[np.searchsorted(loc, self.ghosts[rank])] = INERT
        x[np.searchsorted(loc, P.import_pixels)] = (-2 - np.arange(P.n_import)).astype(np.int32)
        x[np.searchsorted(loc, P.export_pixels)] = np.arange(P.n_export, dtype=np.int32)
        return x

    def summary(self):
        return {nm: {"cut_edges": int(self.g[nm][0].size), "imports_per_rank": self.n_import[nm].tolist(),
                     "exports_per_rank": [int((self.g[nm][2] == r).sum()) for r in range(self.world)]} for nm in self.names}


# ------------------------------------------------------------------------------------------------------------------
# device side
# ------------------------------------------------------------------------------------------------------------------
def _dist():
    import torch.distributed as dist
    return dist


class _Device(object):
    """Partition, cut edges, local graphs and the exchange region of this rank, through the C ABI."""

    def __init__(self, rank, world):
        import torch
        from . import _capi
        self.torch, self.capi, self.L = torch, _capi, _capi.lib()
        self.rank, self.world = int(rank), int(world)

    def build_graph(self, ldd, mask, rows, cols):
        h = C.c_void_p()
        self.capi.check(self.L.lf_ldd_build(self.capi.ptr(ldd), self.capi.ptr(mask), rows, cols, C.byref(h)))
        return h

    def partition(self, graph, n, subtree_fraction):
        torch = self.torch
        owner = torch.empty(n, dtype=torch.int32, device="cuda")
        loads = np.zeros(self.world, np.int64)
        ntr, nro = C.c_int64(), C.c_int64()
        self.capi.check(self.L.lf_graph_partition(graph, self.world, float(subtree_fraction), self.capi.ptr(owner),
                                                  self.capi.ptr(loads), C.byref(ntr), C.byref(nro)))
        return owner, loads, ntr.value, nro.value

    def cut_edges(self, graph, owner):
        cap = 1 << 16
        while True:
            eu, ed, cnt = np.zeros(cap, np.int32), np.zeros(cap, np.int32), C.c_int64()
            self.capi.check(self.L.lf_graph_cut_edges(graph, self.capi.ptr(owner), cap, self.capi.ptr(eu), self.capi.ptr(ed),
                                                      C.byref(cnt)))
            if cnt.value <= cap:
                return eu[:cnt.value].copy(), ed[:cnt.value].copy()
            cap = int(cnt.value)

    def owners_of(self, owner, pixels):
        idx = self.torch.as_tensor(np.asarray(pixels, np.int64), device="cuda")
        return owner[idx].cpu().numpy().astype(np.int64)

    def keep_mask(self, owner, ghosts):
        keep = (owner == self.rank).to(self.torch.uint8)
        if len(ghosts):
            keep[self.torch.as_tensor(np.asarray(ghosts, np.int64), device="cuda")] = 1
        return keep

    def restrict(self, graph, keep):
        h = C.c_void_p()
        self.capi.check(self.L.lf_graph_restrict(graph, self.capi.ptr(keep), C.byref(h)))
        return h

    def open_region(self, doubles):
        """Allocates this rank's exchange region and maps every peer's (CUDA IPC handles over torch.distributed)."""
        dist = _dist()
        x = C.c_void_p()
        self.capi.check(self.L.lf_xchg_create(self.rank, self.world, int(doubles), C.byref(x)))
        mine = C.create_string_buffer(64)
        self.capi.check(self.L.lf_xchg_ipc_handle(x, mine))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine.raw))
        for r in range(self.world):
            if r != self.rank:
                self.capi.check(self.L.lf_xchg_open_peer(x, r, handles[r]))
        dist.barrier()
        return x

    def exchange_args(self, P):
        z32, z64 = np.zeros(1, np.int32), np.zeros(1, np.int64)
        peer = np.ascontiguousarray(P.export_peer) if P.n_export else z32
        off = np.ascontiguousarray(P.export_offset) if P.n_export else z64
        stride = np.ascontiguousarray(P.export_parity_stride) if P.n_export else z64
        return P.n_export, self.capi.ptr(peer), self.capi.ptr(off), self.capi.ptr(stride), P.n_import, P.import_offset, \
            (peer, off, stride)

    def status(self, x):
        ab, ep = C.c_int32(), C.c_int64()
        self.capi.check(self.L.lf_xchg_status(x, C.byref(ab), C.byref(ep)))
        return bool(ab.value), int(ep.value)


def _gather_owned(values_local, owned_local, loc, n_global, rank, world):
    """Global map on rank 0 (None elsewhere) from the owned parts of the ranks -- output / tests, not the hot path."""
    dist = _dist()
    part = (np.ascontiguousarray(loc[owned_local]), np.ascontiguousarray(np.asarray(values_local)[..., owned_local]))
    parts = [None] * world if rank == 0 else None
    dist.gather_object(part, parts, dst=0)
    if rank != 0:
        return None
    first = parts[0][1]
    out = np.empty(first.shape[:-1] + (n_global,), first.dtype)
    for idx, val in parts:
        out[..., idx] = val
    return out


class DistributedKinematicWave(object):
    """kinematicWave over an LDD-cut partition of ONE network: same constructor arguments as the reference class (global
    arrays -- NumPy or CUDA tensors -- on every rank), device-resident protocol set_discharge / set_lateral_inflow /
    run / gather_discharge."""

    def __init__(self, compressed_encoded_ldd, land_mask, alpha_channel, beta, space_delta, time_delta, max_steps=64,
                 subtree_fraction=0.25, rows=None, cols=None):
        import torch
        from . import _capi
        from .hydrological_modules.kinematic_wave_parallel import kinematicWave
        dist = _dist()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        D = self.dev = _Device(self.rank, self.world)
        if rows is None:
            rows, cols = np.asarray(land_mask).shape
        as_dev = lambda a, dt: a if hasattr(a, "data_ptr") else torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
        ldd = as_dev(compressed_encoded_ldd, torch.float64)
        mask = as_dev(np.asarray(land_mask).reshape(-1).astype(np.uint8) if not hasattr(land_mask, "data_ptr") else land_mask,
                      torch.uint8)
        n = int(ldd.numel())
        self.n_global, self.max_steps = n, int(max_steps)
        g = D.build_graph(ldd, mask, rows, cols)
        try:
            owner, loads, ntrunk, nroots = D.partition(g, n, subtree_fraction)
            eu, ed = D.cut_edges(g, owner)
            self.plan = CutPlan({"kw": (eu, ed, D.owners_of(owner, eu), D.owners_of(owner, ed), 1, self.max_steps)},
                                self.world)
            keep = D.keep_mask(owner, self.plan.ghosts[self.rank])
            gl = D.restrict(g, keep)
        finally:
            D.L.lf_graph_destroy(g)
        self.loads, self.n_trunk, self.n_roots = loads.tolist(), ntrunk, nroots
        self.loc_dev = torch.nonzero(keep, as_tuple=False).reshape(-1)
        self.loc = self.loc_dev.cpu().numpy()
        self.owned_local = (owner[self.loc_dev] == self.rank).cpu().numpy()
        self.n_local = int(self.loc.size)
        del owner, keep
        pick = lambda v: float(v) if np.ndim(v) == 0 else as_dev(v, torch.float64)[self.loc_dev].contiguous()
        self.kw = kinematicWave.from_graph(gl, self.n_local, pick(alpha_channel), beta, pick(space_delta), time_delta)
        self.xchg = D.open_region(self.plan.region_doubles[self.rank])
        P = self.plan.rank_plan("kw", self.rank)
        xslot = self.plan.xslot("kw", self.rank, self.loc)
        ne, pp, po, ps, ni, io, self._keep = D.exchange_args(P)
        _capi.check(D.L.lf_router_set_exchange(self.kw._router, self.xchg, _capi.ptr(xslot), ne, pp, po, ps, ni, io,
                                               self.max_steps))
        self.cut_edges = int(eu.size)
        self.exports, self.imports = P.n_export, P.n_import
        dist.barrier()

    def _local(self, a):
        if hasattr(a, "data_ptr"):
            return a[self.loc_dev].contiguous()
        return np.ascontiguousarray(np.asarray(a, np.float64)[self.loc])

    def set_discharge(self, discharge_global):
        self.kw.set_discharge(self._local(discharge_global))

    def set_lateral_inflow(self, q_global):
        self.kw.set_lateral_inflow(self._local(q_global))

    def run(self, nsteps, inflow_scale=None):
        """nsteps routing steps of the whole (cut) network; step s uses lateral inflow q * inflow_scale[s].  Returns as
        soon as the work is queued: the ranks exchange boundary discharges from inside the kernels."""
        assert nsteps <= self.max_steps
        self.kw.run(nsteps, inflow_scale=inflow_scale)

    def local_discharge(self):
        return self.kw.get_discharge()

    def gather_discharge(self):
        return _gather_owned(self.local_discharge(), self.owned_local, self.loc, self.n_global, self.rank, self.world)

    def status(self):
        return self.dev.status(self.xchg)

    def close(self):
        if getattr(self, "kw", None) is not None:
            self.kw.close()
            self.kw = None
        if getattr(self, "xchg", None):
            self.dev.L.lf_xchg_destroy(self.xchg)
            self.xchg = None


class DistributedHotPathModel(object):
    """The full hot-path model (soil -> overland -> channel sub-steps) of ONE raster cut over the ranks.

    S: the dictionary HotPathModel takes, holding GLOBAL arrays (NumPy or CUDA tensors) of the whole raster on every
    rank; `Ldd` (the complete drainage network, before the channel mask is applied) drives the partition, LddToChan and
    LddKinematic the two routing graphs.  Maps are then set with GLOBAL arrays (set / set_flags / set_forcing pick the
    local pixels) or with local ones (set_local); get() returns the local part, gather() the global map on rank 0."""

    def __init__(self, S, diagnostics=False, subtree_fraction=0.25):
        import torch
        from . import _capi
        from .hotpath import HotPathModel
        dist = _dist()
        self.torch = torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        D = self.dev = _Device(self.rank, self.world)
        rows, cols, n = int(S["rows"]), int(S["cols"]), int(S["N"])
        as_dev = lambda a, dt: a if hasattr(a, "data_ptr") else torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
        mask = S["mask_device"] if "mask_device" in S else as_dev(np.asarray(S["mask"]).reshape(-1).astype(np.uint8), torch.uint8)
        split = bool(S.get("SplitRouting"))
        nrs = int(S["NoRoutSteps"])
        g = D.build_graph(as_dev(S["Ldd"], torch.float64), mask, rows, cols)
        try:
            owner, loads, ntrunk, nroots = D.partition(g, n, subtree_fraction)
        finally:
            D.L.lf_graph_destroy(g)
        structures = bool(S.get("simulateReservoirs") or S.get("simulateLakes"))
        ldd_kin = S["LddKinematic"]
        if structures:
            # The channel graph is levelled on the network before structures.initial cuts it (hotpath.py), and a structure
            # stays on one rank with the pixels that drain into it (its inflow is their ChanQ of the previous sub-step,
            # np.bincount(downstruct, ChanQ)): those pixels follow the structure's owner, so no cut edge enters a structure.
            ldd_kin = S["LddStructuresKinematic"]
            sites = np.concatenate([np.asarray(S["ReservoirIndex"], np.int64) if S.get("simulateReservoirs") else np.zeros(0, np.int64),
                                    np.asarray(S["LakeIndex"], np.int64) if S.get("simulateLakes") else np.zeros(0, np.int64)])
            down = np.asarray(S["downstruct"], np.int64)
            feeders = np.flatnonzero(np.isin(down, sites))
            if feeders.size:
                f_dev = torch.as_tensor(feeders, device="cuda")
                d_dev = torch.as_tensor(down[feeders], device="cuda")
                for _ in range(64):                      # chains of structures feeding structures settle in a few rounds
                    new = owner[d_dev]
                    if bool((owner[f_dev] == new).all()):
                        break
                    owner[f_dev] = new
                loads = torch.bincount(owner.to(torch.int64), minlength=self.world).cpu().numpy()
        g_of = D.build_graph(as_dev(S["LddToChan"], torch.float64), mask, rows, cols)
        g_ch = D.build_graph(as_dev(ldd_kin, torch.float64), mask, rows, cols)
        try:
            e_of, e_ch = D.cut_edges(g_of, owner), D.cut_edges(g_ch, owner)
            graphs = {"overland": (e_of[0], e_of[1], D.owners_of(owner, e_of[0]), D.owners_of(owner, e_of[1]), 3, 1),
                      "channel": (e_ch[0], e_ch[1], D.owners_of(owner, e_ch[0]), D.owners_of(owner, e_ch[1]),
                                  2 if split else 1, nrs)}
            self.plan = CutPlan(graphs, self.world, order=("overland", "channel"))
            keep = D.keep_mask(owner, self.plan.ghosts[self.rank])
            l_of, l_ch = D.restrict(g_of, keep), D.restrict(g_ch, keep)
        finally:
            D.L.lf_graph_destroy(g_of)
            D.L.lf_graph_destroy(g_ch)
        self.loads, self.n_trunk, self.n_roots, self.n_global = loads.tolist(), ntrunk, nroots, n
        self.loc_dev = torch.nonzero(keep, as_tuple=False).reshape(-1)
        self.loc = self.loc_dev.cpu().numpy()
        self.owned_local = (owner[self.loc_dev] == self.rank).cpu().numpy()
        self.n_local = int(self.loc.size)
        owner_of_sites = D.owners_of(owner, sites) if structures else None
        del owner, keep
        self.model = HotPathModel(S, diagnostics=diagnostics, graphs=(l_of, l_ch), n_active=self.n_local, pick=self._local)
        if structures:
            self._set_local_structures(S, owner_of_sites)
        self.xchg = D.open_region(self.plan.region_doubles[self.rank])
        self._keep = []
        for which, name in enumerate(("overland", "channel")):
            P = self.plan.rank_plan(name, self.rank)
            xslot = self.plan.xslot(name, self.rank, self.loc)
            ne, pp, po, ps, ni, io, keepalive = D.exchange_args(P)
            self._keep.append(keepalive)
            _capi.check(D.L.lf_model_set_exchange(self.model._h, self.xchg, which, _capi.ptr(xslot), ne, pp, po, ps, ni, io))
        dist.barrier()

    def _set_local_structures(self, S, owner_of_sites):
        """Hands this rank's reservoirs / lakes (local pixel indices, the matching slices of the per-structure arrays) to its
        model; maps such as ReservoirStorageM3 then gather like any other map."""
        from .hotpath import LAKE_ARRAYS, RESERVOIR_ARRAYS
        nres = len(S["ReservoirIndex"]) if S.get("simulateReservoirs") else 0
        sub = {"simulateReservoirs": False, "simulateLakes": False}
        for flag, key, names, own in (("simulateReservoirs", "ReservoirIndex", RESERVOIR_ARRAYS, owner_of_sites[:nres]),
                                      ("simulateLakes", "LakeIndex", LAKE_ARRAYS, owner_of_sites[nres:])):
            if not S.get(flag):
                continue
            mine = np.flatnonzero(own == self.rank)
            if mine.size == 0:
                continue
            sub[flag] = True
            sub[key] = np.searchsorted(self.loc, np.asarray(S[key], np.int64)[mine]).astype(np.int64)
            for name in names:
                sub[name] = np.asarray(S[name], np.float64)[mine]
        if sub["simulateReservoirs"] or sub["simulateLakes"]:
            self.model.set_structures(sub)

    def _local(self, a):
        """Local pixels of a global map (last axis = pixels); scalars and local-sized arrays pass through."""
        if hasattr(a, "data_ptr"):
            if a.shape[-1] != self.n_global:
                return a
            return a[..., self.loc_dev].contiguous() if a.is_cuda else a[..., self.torch.as_tensor(self.loc)].contiguous()
        a = np.asarray(a)
        return np.ascontiguousarray(a[..., self.loc]) if a.ndim and a.shape[-1] == self.n_global else a

    def set(self, name, values, rows=None):
        self.model.set(name, self._local(values), rows)

    def set_local(self, name, values, rows=None):
        self.model.set(name, values, rows)

    def set_flags(self, name, values):
        self.model.set_flags(name, self._local(values))

    def set_forcing(self, F):
        self.model.set_forcing({k: self._local(v) for k, v in F.items()})

    def set_feeder(self, P, state=None):
        self.model.set_feeder({k: self._local(v) for k, v in P.items()}, {k: self._local(v) for k, v in (state or {}).items()})

    def feed(self, raw, calendar_day, asynchronous=False, local=False, packing=None, decode="float32"):
        """Raw meteo maps of the step (global maps, or this rank's pixels with local=True); packing / decode as in
        HotPathModel.feed."""
        self.model.feed(raw if local else {k: self._local(v) for k, v in raw.items()}, calendar_day, asynchronous=asynchronous,
                        packing=packing, decode=decode)

    def set_lai(self, lai, local=False):
        self.model.set_lai(lai if local else self._local(lai))

    def set_option(self, name, value):
        self.model.set_option(name, value)

    def stage_times(self, reset=False):
        return self.model.stage_times(reset)

    def soil_stats(self, enable_timing=True):
        return self.model.soil_stats(enable_timing)

    def info(self):
        return self.model.info()

    def get_into(self, name, out):
        return self.model.get_into(name, out)

    def local_forcing(self, F):
        return {k: self._local(v) for k, v in F.items()}

    def step(self, F=None, local=False):
        if F is not None:
            self.model.set_forcing(F if local else {k: self._local(v) for k, v in F.items()})
        self.model.step()

    def get(self, name, rows=None):
        return self.model.get(name, rows)

    def gather(self, name, rows=None):
        return _gather_owned(self.model.get(name, rows), self.owned_local, self.loc, self.n_global, self.rank, self.world)

    def status(self):
        return self.dev.status(self.xchg)

    def close(self):
        if getattr(self, "model", None) is not None:
            self.model.close()
            self.model = None
        if getattr(self, "xchg", None):
            self.dev.L.lf_xchg_destroy(self.xchg)
            self.xchg = None
