"""ctypes wrapper around the C oracle (oracle/lisf_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
`KinematicWaveOracle` mirrors the reference's kinematicWave class
(reference: src/lisflood/hydrological_modules/kinematic_wave_parallel.py:114-184).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "liblisf_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("lisf_oracle.c", "lisf_oracle_soil.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.lfo_set_threads.argtypes = [C.c_int]
        L.lfo_set_threads.restype = C.c_int
        L.lfo_ldd_graph.argtypes = [_f64p, _u8p, C.c_int64, C.c_int64, _f64p, _i64p, _i64p, _i64p, _i64p,
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.lfo_ldd_graph.restype = C.c_int
        L.lfo_kinematic_wave_routing.argtypes = [_f64p, _f64p, C.c_int64, C.c_void_p, C.c_double, _f64p, _f64p,
                                                 C.c_double, _i64p, C.c_int64, _i64p, _i64p, _i64p, C.c_int64,
                                                 _f64p, C.c_void_p, C.c_void_p]
        L.lfo_kinematic_wave_routing.restype = C.c_int64
        _LIB = L
    return _LIB


def set_threads(n):
    return lib().lfo_set_threads(int(n))


class KinematicWaveOracle:
    """Same constructor / method / attributes as the reference's kinematicWave
    (kinematic_wave_parallel.py:117-184)."""

    def __init__(self, compressed_encoded_ldd, land_mask, alpha_channel, beta, space_delta, time_delta,
                 alpha_floodplains=None, flagnancheck=False):
        L = lib()
        ldd = np.ascontiguousarray(compressed_encoded_ldd, np.float64)
        mask = np.ascontiguousarray(land_mask).astype(np.uint8)
        n = int(mask.sum())
        assert ldd.size == n
        self.num_pixels = n
        self.space_delta = space_delta
        self.beta = float(beta)
        self.inv_beta = 1 / beta
        self.b_minus_1 = beta - 1
        self.a_dx_div_dt_channel = np.ascontiguousarray(
            np.broadcast_to(np.asarray(alpha_channel * space_delta / time_delta, np.float64), (n,)))
        self.b_a_dx_div_dt_channel = beta * self.a_dx_div_dt_channel
        if alpha_floodplains is not None:
            self.a_dx_div_dt_floodplains = np.ascontiguousarray(
                np.broadcast_to(np.asarray(alpha_floodplains * space_delta / time_delta, np.float64), (n,)))
            self.b_a_dx_div_dt_floodplains = beta * self.a_dx_div_dt_floodplains
        self.downstream_lookup = np.empty(n, np.float64)
        ups = np.empty((n, 8), np.int64)
        self.num_upstream_pixels = np.empty(n, np.int64)
        self.pixels_ordered = np.empty(n, np.int64)
        oss = np.empty((max(n, 1), 2), np.int64)
        no, k = C.c_int64(0), C.c_int64(0)
        rc = L.lfo_ldd_graph(ldd, mask.ravel(), mask.shape[0], mask.shape[1], self.downstream_lookup, ups,
                             self.num_upstream_pixels, self.pixels_ordered, oss, C.byref(no), C.byref(k))
        if rc == -1:
            raise ValueError("LDD codes must be integers in 0..9")
        if rc == -2:
            raise ValueError("LDD contains a cycle")
        self.order_start_stop = np.ascontiguousarray(oss[:no.value])
        self.upstream_lookup = np.ascontiguousarray(ups[:, :k.value])
        self._work = np.empty(n, np.float64)
        self.flagnancheck = flagnancheck
        self.kinematic_wave_warning_printed = False
        self.last_newton_iterations = 0

    def kinematicWaveRouting(self, discharge, specific_lateral_inflow, section="main_channel", fixed=None,
                             fixed_values=None):
        """`fixed` / `fixed_values` (optional): pixels whose discharge is prescribed instead of solved -- the
        ghost pixels of an LDD-cut partition; not part of the reference API."""
        if section == "main_channel":
            a, ba = self.a_dx_div_dt_channel, self.b_a_dx_div_dt_channel
        elif section == "floodplains":
            a, ba = self.a_dx_div_dt_floodplains, self.b_a_dx_div_dt_floodplains
        else:
            raise Exception("The section parameter must be either 'main_channel' or 'floodplain'!")
        n = self.num_pixels
        if isinstance(self.space_delta, np.ndarray):
            dx = np.ascontiguousarray(self.space_delta, np.float64)
            dxp, dxs = dx.ctypes.data, 0.0
        else:
            dxp, dxs = None, float(self.space_delta)
        q = np.ascontiguousarray(specific_lateral_inflow, np.float64)
        fx = None if fixed is None else np.ascontiguousarray(fixed, np.uint8)
        fv = None if fixed_values is None else np.ascontiguousarray(fixed_values, np.float64)
        assert discharge.dtype == np.float64 and discharge.flags.c_contiguous
        self.last_newton_iterations = lib().lfo_kinematic_wave_routing(
            discharge, q, n, dxp, dxs, a, ba, self.beta, self.upstream_lookup, self.upstream_lookup.shape[1],
            self.num_upstream_pixels, self.pixels_ordered, self.order_start_stop.ravel(),
            self.order_start_stop.shape[0], self._work,
            None if fx is None else fx.ctypes.data, None if fv is None else fv.ctypes.data)
        if self.flagnancheck and not self.kinematic_wave_warning_printed:
            if not np.all(np.isfinite(discharge)):
                import warnings
                warnings.warn("Warning: NaN or Inf values after kinematicRouting module.")
                self.kinematic_wave_warning_printed = True
