"""LDD-cut domain decomposition of the kinematic-wave routing across the GPUs of one node
(BASELINE.json config C4).

Information in a drainage network flows strictly downstream (a pixel needs only the NEW discharge of its
upstream pixels, reference: hydrological_modules/kinematic_wave_parallel_tools.py:57-58) and whole catchments
are independent (the reference proves sub-mask runs are bit-identical, tests/test_subcatchments.py:111-112).  So
the raster is cut along the drainage graph, not along raster tiles:

  * every pixel whose upstream area exceeds a threshold forms the "trunk" (the main stems, a tiny fraction of
    the pixels); everything that hangs off the trunk -- or drains straight to an outlet -- is a sub-tree;
  * sub-trees are bin-packed over the ranks (largest first); the trunk lives on rank 0;
  * the only communication: the root of a sub-tree owned by rank r != 0 feeds its trunk parent on rank 0.
    Its discharge of every routing step of a run is written by the routing kernel into an export buffer and
    sent to rank 0 (one message per run and rank: n_cut_edges x n_steps float64), where the sub-tree root
    exists as a GHOST pixel whose value is read instead of solved.  Coupling is one-way, so ranks != 0 start
    their next run while rank 0 consumes the previous one (pipelining across runs, no per-level barrier).

Every rank builds its router on its own sub-mask of the global raster (links that leave the sub-mask vanish in
lf_ldd_build exactly like in the reference), so the cut network reproduces the uncut one BIT FOR BIT: same
upstream slots, same summation order, same values (checked by tests/test_parallel_cpu.py on gloo with a CPU
stand-in router injected by the test, and by tools/run_dist_check.py on NCCL with the real routers).
"""
import heapq

import numpy as np

from .global_modules import ldd_ops


class Partition(object):
    """Deterministic partition of a drainage network over `world` ranks (identical on every rank)."""

    def __init__(self, ldd_codes, land_mask, world, subtree_fraction=0.25, graph=None):
        """graph: optional kinematicWave built on the GLOBAL network; its device graph then supplies the
        downstream index, the upstream areas (lf_graph_accuflux) and the level ordering -- O(N) instead of the
        O(N * depth) NumPy fallback used by the CPU tests."""
        mask = np.asarray(land_mask, bool)
        ldd = np.asarray(ldd_codes, np.float64)
        n = int(mask.sum())
        self.world, self.n = int(world), n
        if graph is not None:
            ds = graph.downstream_lookup.astype(np.int64)
            size = graph.accuflux(np.ones(n))
            oss = graph.order_start_stop
            segs = np.split(graph.pixels_ordered, oss[1:, 0])   # routing order 0 = farthest from the outlets
        else:
            ds = ldd_ops.downstream_index(ldd, mask)
            size = ldd_ops.accuflux(ds, np.ones(n))
            order, hops = ldd_ops.topological_order(ds)
            bounds = np.flatnonzero(np.diff(hops[order])) + 1
            segs = np.split(order, bounds)      # decreasing hops: farthest first
        self.downstream = ds
        threshold = max(1.0, subtree_fraction * n / max(world, 1))
        trunk = size > threshold if world > 1 else np.zeros(n, bool)
        root = (~trunk) & ((ds < 0) | trunk[np.maximum(ds, 0)])
        # label every non-trunk pixel with the root of its sub-tree: walk the pixels from downstream to upstream
        label = np.where(root, np.arange(n), -1)
        for seg in reversed(segs):              # outlets first
            p = seg[(~trunk[seg]) & (~root[seg])]
            label[p] = label[ds[p]]
        assert (label[~trunk] >= 0).all()
        roots = np.flatnonzero(root)
        sizes = size[roots]
        owner_of_root = np.zeros(roots.size, np.int64)
        load = [(int(trunk.sum()) if r == 0 else 0, r) for r in range(world)]
        heapq.heapify(load)
        for j in np.argsort(-sizes, kind="stable"):
            l, r = heapq.heappop(load)
            owner_of_root[j] = r
            heapq.heappush(load, (l + int(sizes[j]), r))
        owner = np.zeros(n, np.int64)
        tmp = np.zeros(n, np.int64)
        tmp[roots] = owner_of_root
        owner[~trunk] = tmp[label[~trunk]]
        self.owner = owner
        self.trunk = trunk
        # cut edges: sub-tree roots owned by r != 0 whose parent is a trunk pixel (rank 0)
        cut = root & (ds >= 0) & (owner != 0)
        self.cut_pixels = [np.flatnonzero(cut & (owner == r)) for r in range(world)]   # ascending global index
        self.n_cut = [int(c.size) for c in self.cut_pixels]
        self.import_offset = np.concatenate([[0], np.cumsum(self.n_cut)])[:world]
        self.n_import = int(sum(self.n_cut))
        self.loads = [int((owner == r).sum()) for r in range(world)]

    def local_pixels(self, rank):
        """Global compressed indices of the pixels in rank's sub-mask (owned + ghosts), ascending."""
        own = self.owner == rank
        if rank == 0 and self.n_import:
            own = own.copy()
            for c in self.cut_pixels:
                own[c] = True
        return np.flatnonzero(own)

    def local_xslot(self, rank):
        """int32[N_local]: -1 plain, >= 0 export slot, <= -2 ghost slot (see lf_router_set_exchange)."""
        loc = self.local_pixels(rank)
        x = -np.ones(loc.size, np.int32)
        pos = -np.ones(self.n, np.int64)
        pos[loc] = np.arange(loc.size)
        if rank == 0:
            for r in range(1, self.world):
                c = self.cut_pixels[r]
                x[pos[c]] = -2 - (self.import_offset[r] + np.arange(c.size))
        else:
            c = self.cut_pixels[rank]
            x[pos[c]] = np.arange(c.size)
        return x


class _Comm(object):
    """torch.distributed point-to-point plumbing (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()


class GpuRouterBackend(object):
    """Routers of this package + NCCL.  A backend builds the local router of a rank and runs its steps; the CPU
    tests inject a stand-in with the same protocol (tests/test_parallel_cpu.py) to exercise the host logic."""
    device = "cuda"

    def __init__(self, ldd_local, sub_mask, alpha, beta, dx, dt, xslot, n_exp, n_imp, export, imported, max_steps, world):
        from . import _capi
        from .hydrological_modules.kinematic_wave_parallel import kinematicWave
        self._capi = _capi
        self.kw = kinematicWave(ldd_local, sub_mask, alpha, beta, dx, dt)
        if world > 1:
            _capi.check(_capi.lib().lf_router_set_exchange(self.kw._router, _capi.ptr(xslot), n_exp, n_imp,
                                                           _capi.ptr(export), _capi.ptr(imported), max_steps))

    def set_discharge(self, q):
        self.kw.set_discharge(q)

    def set_lateral_inflow(self, q):
        self.kw.set_lateral_inflow(q)

    def run(self, nsteps, inflow_scale):
        self.kw.run(nsteps, inflow_scale=inflow_scale)

    def get_discharge(self):
        return self.kw.get_discharge()

    def before_send(self):
        self._capi.synchronize()       # the export buffer is written on the library's stream

    def after_recv(self):
        import torch
        torch.cuda.current_stream().synchronize()

    @staticmethod
    def global_graph(ldd, mask, beta):
        from .hydrological_modules.kinematic_wave_parallel import kinematicWave
        return kinematicWave(ldd, mask, np.ones(int(np.asarray(mask).sum())), beta, 1.0, 1.0)


class DistributedKinematicWave(object):
    """kinematicWave over an LDD-cut partition: same constructor arguments as the reference class
    (global arrays on every rank), device-resident protocol set_discharge / set_lateral_inflow / run /
    gather_discharge.  `backend`: class with the protocol of GpuRouterBackend (default)."""

    def __init__(self, compressed_encoded_ldd, land_mask, alpha_channel, beta, space_delta, time_delta, max_steps=64,
                 backend=None, subtree_fraction=0.25):
        import torch
        self.torch = torch
        self.comm = _Comm()
        rank, world = self.comm.rank, self.comm.world
        backend = backend or GpuRouterBackend
        mask = np.asarray(land_mask, bool)
        ldd = np.asarray(compressed_encoded_ldd, np.float64)
        graph = None
        if world > 1 and hasattr(backend, "global_graph"):
            graph = backend.global_graph(ldd, mask, beta)   # global graph, used for the partition only
        self.part = P = Partition(ldd, mask, world, subtree_fraction, graph=graph)
        if graph is not None:
            graph.close()
        self.loc = loc = P.local_pixels(rank)
        self.n_local, self.n_global = loc.size, P.n
        self.owned_local = P.owner[loc] == rank          # ghosts are False
        gmask = np.zeros(P.n, bool)
        gmask[loc] = True
        sub = np.zeros(mask.shape, bool)
        sub[mask] = gmask
        self.max_steps = int(max_steps)
        pick = lambda v: v if np.ndim(v) == 0 else np.ascontiguousarray(np.asarray(v, np.float64)[loc])
        self.xslot = P.local_xslot(rank)
        n_exp = P.n_cut[rank] if rank != 0 else 0
        n_imp = P.n_import if rank == 0 else 0
        dev = backend.device
        self.export = torch.zeros(max(n_exp, 1) * self.max_steps, dtype=torch.float64, device=dev)
        self.imported = torch.zeros(max(n_imp, 1) * self.max_steps, dtype=torch.float64, device=dev)
        self.n_exp, self.n_imp = n_exp, n_imp
        self.router = backend(ldd[loc], sub, pick(alpha_channel), beta, pick(space_delta), time_delta, self.xslot, n_exp,
                              n_imp, self.export, self.imported, self.max_steps, world)
        self.beta = beta

    def set_discharge(self, discharge_global):
        self.router.set_discharge(np.ascontiguousarray(np.asarray(discharge_global, np.float64)[self.loc]))

    def set_lateral_inflow(self, q_global):
        self.router.set_lateral_inflow(np.ascontiguousarray(np.asarray(q_global, np.float64)[self.loc]))

    def _exchange_in(self):
        """rank 0: receive every other rank's export block of this run."""
        P, cap = self.part, self.max_steps
        for r in range(1, self.comm.world):
            if P.n_cut[r]:
                o = int(P.import_offset[r]) * cap
                self.comm.dist.recv(self.imported[o:o + P.n_cut[r] * cap], src=r)
        self.router.after_recv()

    def _exchange_out(self):
        if self.n_exp:
            self.router.before_send()
            self.comm.dist.send(self.export[:self.n_exp * self.max_steps], dst=0)

    def run(self, nsteps, inflow_scale=None):
        """nsteps routing steps of the whole (cut) network; step s uses lateral inflow q * inflow_scale[s]."""
        assert nsteps <= self.max_steps
        rank, world = self.comm.rank, self.comm.world
        if world > 1 and rank == 0 and self.n_imp:
            self._exchange_in()
        self.router.run(nsteps, inflow_scale)
        if world > 1 and rank != 0:
            self._exchange_out()

    def local_discharge(self):
        return self.router.get_discharge()

    def gather_discharge(self):
        """Global discharge map on rank 0 (None elsewhere) -- for output / tests, not on the hot path."""
        dist, torch = self.comm.dist, self.torch
        q = self.local_discharge()
        own = self.owned_local
        mine = torch.from_numpy(np.ascontiguousarray(q[own]))
        idx = self.loc[own]
        dev = self.export.device
        if self.comm.rank == 0:
            out = np.empty(self.n_global)
            out[idx] = mine.numpy()
            for r in range(1, self.comm.world):
                cnt = self.part.loads[r]
                buf = torch.empty(cnt, dtype=torch.float64, device=dev)
                dist.recv(buf, src=r)
                out[np.flatnonzero(self.part.owner == r)] = buf.cpu().numpy()
            return out
        dist.send(mine.to(dev), dst=0)
        return None
