"""kinematicWave -- drop-in for the reference's class of the same name
(reference: src/lisflood/hydrological_modules/kinematic_wave_parallel.py:114-184), executing on a B200
through liblisf_b200.so.

Same constructor signature, same `kinematicWaveRouting(discharge, specific_lateral_inflow, section)`
contract (returns None, mutates `discharge` in place, raises Exception on a bad section, warns once on
non-finite output when flagnancheck is set; `discharge` may be any array that supports in-place assignment -- the
fast path is a C-contiguous float64 array) and the same graph attributes (`downstream_lookup`,
`upstream_lookup`, `num_upstream_pixels`, `pixels_ordered`, `order_start_stop`; bit-identical, fetched
lazily from the device).  Additional device-resident methods (`set_discharge`, `set_lateral_inflow`,
`run`, `get_discharge`) keep the state in HBM across many routing steps and execute them as one
space-time wavefront.

Numerics: the solver reproduces the reference's initial guess and Newton iterates (csrc/lf_kw_solve.cuh) and adds one exit
(relative Newton step <= 1e-8); results agree with the reference's to ~1e-13 relative (pinned at < 1e-10 on adversarial
inputs, where cancellation amplifies last-bit differences, by tests/test_gpu_kinwave.py), the contract being 1e-6.  Only the graph arrays are bit-identical.
"""
import ctypes as C
import warnings

import numpy as np

from .. import _capi
from ..global_modules.errors import LisfloodWarning


class kinematicWave:
    """"""

    def __init__(self, compressed_encoded_ldd, land_mask, alpha_channel, beta, space_delta, time_delta,
                 alpha_floodplains=None, flagnancheck=False):
        L = _capi.lib()
        self.kinematic_wave_warning_printed = False
        self.flagnancheck = flagnancheck
        self.space_delta = space_delta
        self.beta = beta
        self.inv_beta = 1 / beta
        self.b_minus_1 = beta - 1
        mask = np.ascontiguousarray(np.asarray(land_mask, bool)).astype(np.uint8)
        if mask.ndim != 2:
            raise ValueError("land_mask must be 2-D")
        n = int(mask.sum())
        ldd = np.ascontiguousarray(compressed_encoded_ldd, np.float64).ravel()
        if ldd.size != n:
            raise ValueError("compressed_encoded_ldd has %d values, land_mask has %d active pixels" % (ldd.size, n))
        self.num_pixels = n
        g = C.c_void_p()
        _capi.check(L.lf_ldd_build(_capi.ptr(ldd), _capi.ptr(mask), mask.shape[0], mask.shape[1], C.byref(g)))
        self._graph = g
        self._router = C.c_void_p()
        alpha = self._as_map(alpha_channel)
        alpha_fp = None if alpha_floodplains is None else self._as_map(alpha_floodplains)
        if np.ndim(space_delta) == 0:
            dx, dxs = None, float(space_delta)
        else:
            dx, dxs = self._as_map(space_delta), 0.0
        r = C.c_void_p()
        _capi.check(L.lf_router_create(g, _capi.ptr(alpha), float(beta), _capi.ptr(dx), dxs, float(time_delta),
                                       _capi.ptr(alpha_fp), 1 if flagnancheck else 0, C.byref(r)))
        self._router = r
        no, k, npix, pits = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _capi.check(L.lf_graph_info(g, C.byref(npix), C.byref(no), C.byref(k), C.byref(pits)))
        self.num_orders, self.max_upstream, self.num_pits = no.value, k.value, pits.value
        self._cache = {}
        self._env_options()

    def _env_options(self):
        """Tuning runs: LF_ROUTER_COOP / LF_ROUTER_GRAPHS / LF_ROUTER_NARROW override the execution options of the router."""
        import os
        if os.environ.get("LF_ROUTER_NARROW") is not None:
            self.set_option("narrow_runs", float(os.environ["LF_ROUTER_NARROW"]))
        if os.environ.get("LF_ROUTER_COOP") is not None:
            self.set_option("cooperative", float(os.environ["LF_ROUTER_COOP"]))
        if os.environ.get("LF_ROUTER_GRAPHS") is not None:
            self.set_option("cuda_graphs", float(os.environ["LF_ROUTER_GRAPHS"]))

    @classmethod
    def from_graph(cls, graph, num_pixels, alpha_channel, beta, space_delta, time_delta, alpha_floodplains=None,
                   flagnancheck=False):
        """A router on an existing device graph (ctypes handle, e.g. from lf_graph_restrict: the local part of a cut
        network, lisflood_code_b200/parallel.py); the router owns the graph.  Maps: NumPy arrays or CUDA tensors in the
        graph's compressed order."""
        L = _capi.lib()
        self = cls.__new__(cls)
        self.kinematic_wave_warning_printed = False
        self.flagnancheck = flagnancheck
        self.space_delta = space_delta
        self.beta, self.inv_beta, self.b_minus_1 = beta, 1 / beta, beta - 1
        self.num_pixels = int(num_pixels)
        self._graph = graph
        self._router = C.c_void_p()
        alpha = self._as_map(alpha_channel)
        alpha_fp = None if alpha_floodplains is None else self._as_map(alpha_floodplains)
        if np.ndim(space_delta) == 0 and not hasattr(space_delta, "data_ptr"):
            dx, dxs = None, float(space_delta)
        else:
            dx, dxs = self._as_map(space_delta), 0.0
        r = C.c_void_p()
        _capi.check(L.lf_router_create(graph, _capi.ptr(alpha), float(beta), _capi.ptr(dx), dxs, float(time_delta),
                                       _capi.ptr(alpha_fp), 1 if flagnancheck else 0, C.byref(r)))
        self._router = r
        no, k, npix, pits = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _capi.check(L.lf_graph_info(graph, C.byref(npix), C.byref(no), C.byref(k), C.byref(pits)))
        self.num_orders, self.max_upstream, self.num_pits = no.value, k.value, pits.value
        self._cache = {}
        self._env_options()
        return self

    def _as_map(self, v):
        if hasattr(v, "data_ptr"):       # torch tensor (host or CUDA): float64, contiguous, one value per pixel
            assert v.is_contiguous() and v.element_size() == 8 and v.numel() == self.num_pixels
            return v
        return np.ascontiguousarray(np.broadcast_to(np.asarray(v, np.float64), (self.num_pixels,)))

    # ---- graph attributes of the reference object (lazy; bit-identical) -------------------------
    def _export(self, name):
        if name not in self._cache:
            n, k, no = self.num_pixels, self.max_upstream, self.num_orders
            bufs = {"pixels_ordered": np.empty(n, np.int64), "order_start_stop": np.empty((no, 2), np.int64),
                    "upstream_lookup": np.empty((n, k), np.int64), "num_upstream_pixels": np.empty(n, np.int64),
                    "downstream_lookup": np.empty(n, np.float64)}
            order = ["pixels_ordered", "order_start_stop", "upstream_lookup", "num_upstream_pixels", "downstream_lookup"]
            args = [_capi.ptr(bufs[o]) if o == name else None for o in order]
            _capi.check(_capi.lib().lf_graph_export(self._graph, *args))
            self._cache[name] = bufs[name]
        return self._cache[name]

    pixels_ordered = property(lambda self: self._export("pixels_ordered"))
    order_start_stop = property(lambda self: self._export("order_start_stop"))
    upstream_lookup = property(lambda self: self._export("upstream_lookup"))
    num_upstream_pixels = property(lambda self: self._export("num_upstream_pixels"))
    downstream_lookup = property(lambda self: self._export("downstream_lookup"))

    def storage_layout(self):
        """(pixel_of_position int32[N], level_start int32[num_orders+1]) of the device layout."""
        pop = np.empty(self.num_pixels, np.int32)
        ls = np.empty(self.num_orders + 1, np.int32)
        _capi.check(_capi.lib().lf_graph_layout(self._graph, _capi.ptr(pop), _capi.ptr(ls)))
        return pop, ls

    # ---- reference operator ----------------------------------------------------------------------
    @staticmethod
    def _section(section):
        if section not in _capi.SECTION:
            raise Exception("The section parameter must be either 'main_channel' or 'floodplain'!")
        return _capi.SECTION[section]

    def _warn(self, nonfinite):
        if self.flagnancheck and not self.kinematic_wave_warning_printed and nonfinite:
            warnings.warn(LisfloodWarning("Warning: NaN or Inf values after kinematicRouting module. Suggestion: "
                                          "please check the input maps (e.g. channel geometry and ldd)"))
            self.kinematic_wave_warning_printed = True

    def kinematicWaveRouting(self, discharge, specific_lateral_inflow, section="main_channel"):
        """Kinematic wave routing of one step; `discharge` (float64[N]) is updated in place."""
        sec = self._section(section)
        q = self._as_map(specific_lateral_inflow)
        bad = C.c_int(0)
        fast = (isinstance(discharge, np.ndarray) and discharge.dtype == np.float64 and discharge.flags.c_contiguous
                and discharge.size == self.num_pixels)
        if fast:
            _capi.check(_capi.lib().lf_router_route(self._router, discharge, q, sec, C.byref(bad)))
        else:
            # like the reference, any array-like that supports item assignment is accepted (float32, views, strided,
            # NumpyModified): routed on a float64 copy, written back in place
            if np.size(discharge) != self.num_pixels:
                raise ValueError("discharge must hold %d pixels" % self.num_pixels)
            tmp = np.ascontiguousarray(np.asarray(discharge, np.float64)).reshape(-1).copy()
            _capi.check(_capi.lib().lf_router_route(self._router, tmp, q, sec, C.byref(bad)))
            discharge[...] = tmp.reshape(np.shape(discharge))
        self._warn(bad.value)

    # ---- device-resident extension ---------------------------------------------------------------
    def set_discharge(self, discharge, section="main_channel"):
        _capi.check(_capi.lib().lf_router_set_discharge(self._router, self._section(section),
                                                        _capi.ptr(self._as_map(discharge))))

    def set_lateral_inflow(self, specific_lateral_inflow, section="main_channel"):
        _capi.check(_capi.lib().lf_router_set_inflow(self._router, self._section(section),
                                                     _capi.ptr(self._as_map(specific_lateral_inflow))))

    def run(self, nsteps, inflow_scale=None, section="main_channel"):
        """nsteps consecutive kinematicWaveRouting calls on the resident state, executed as one
        space-time wavefront; step s uses lateral inflow q * inflow_scale[s]."""
        sc = None
        if inflow_scale is not None:
            sc = np.ascontiguousarray(inflow_scale, np.float64)
            if sc.size != nsteps:
                raise ValueError("inflow_scale must have nsteps entries")
        bad = C.c_int(0)
        _capi.check(_capi.lib().lf_router_run(self._router, self._section(section), int(nsteps), _capi.ptr(sc),
                                              C.byref(bad)))
        self._warn(bad.value)

    def set_option(self, name, value):
        """Execution options of lf_router_set_option ("cooperative", "cuda_graphs", "narrow_runs")."""
        _capi.check(_capi.lib().lf_router_set_option(self._router, name.encode(), float(value)))

    def get_discharge(self, section="main_channel", out=None):
        if out is None:
            out = np.empty(self.num_pixels, np.float64)
        _capi.check(_capi.lib().lf_router_get_discharge(self._router, self._section(section), _capi.ptr(out)))
        return out

    def accuflux(self, x):
        """PCRaster accuflux on this router's graph (downstream-accumulated sum, including the cell)."""
        x = self._as_map(x)
        out = np.empty_like(x)
        _capi.check(_capi.lib().lf_graph_accuflux(self._graph, _capi.ptr(x), _capi.ptr(out)))
        return out

    def close(self):
        L = _capi._lib
        if L is None:
            return
        if getattr(self, "_router", None):
            L.lf_router_destroy(self._router)
            self._router = C.c_void_p()
        if getattr(self, "_graph", None):
            L.lf_graph_destroy(self._graph)
            self._graph = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
