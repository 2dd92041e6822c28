"""ctypes binding of liblisf_b200.so (C ABI: include/lisflood_b200.h).

The library is loaded on first use.  A missing library is a hard error -- the product path never
falls back to a CPU implementation.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "liblisf_b200.so")

LF_OK = 0
LF_ERR_INVALID, LF_ERR_BAD_LDD, LF_ERR_LDD_CYCLE, LF_ERR_CUDA, LF_ERR_NO_DEVICE, LF_ERR_STATE = -1, -2, -3, -4, -5, -6
SECTION = {"main_channel": 0, "floodplains": 1}


class LisfloodB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("liblisf_b200 error %d: %s" % (code, message))
        self.code = code


_f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i64 = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_i64s = C.c_int64

# name -> (restype, argtypes); must list every function declared in include/lisflood_b200.h
SIGNATURES = {
    "lf_last_error": (C.c_char_p, []),
    "lf_version": (C.c_int, []),
    "lf_device_init": (C.c_int, [C.c_int]),
    "lf_device_info": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int64)]),
    "lf_synchronize": (C.c_int, []),
    "lf_timer_start": (C.c_int, []),
    "lf_timer_stop": (C.c_int, [C.POINTER(C.c_double)]),
    "lf_launch_count": (C.c_int64, [C.c_int]),
    "lf_host_launch_count": (C.c_int64, []),
    "lf_host_register": (C.c_int, [_vp, _i64s]),
    "lf_host_unregister": (C.c_int, [_vp]),
    "lf_host_alloc": (C.c_int, [_i64s, C.POINTER(_vp)]),
    "lf_host_free": (C.c_int, [_vp]),
    "lf_math_selftest": (C.c_int, [_i64s, C.c_uint64, _f64]),
    "lf_ldd_build": (C.c_int, [_vp, _vp, _i64s, _i64s, C.POINTER(_vp)]),
    "lf_graph_info": (C.c_int, [_vp, C.POINTER(_i64s), C.POINTER(_i64s), C.POINTER(_i64s), C.POINTER(_i64s)]),
    "lf_graph_export": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "lf_graph_layout": (C.c_int, [_vp, _vp, _vp]),
    "lf_graph_accuflux": (C.c_int, [_vp, _vp, _vp]),
    "lf_graph_destroy": (None, [_vp]),
    "lf_router_create": (C.c_int, [_vp, _vp, C.c_double, _vp, C.c_double, C.c_double, _vp, C.c_int,
                                   C.POINTER(_vp)]),
    "lf_router_route": (C.c_int, [_vp, _f64, _f64, C.c_int, C.POINTER(C.c_int)]),
    "lf_router_set_discharge": (C.c_int, [_vp, C.c_int, _vp]),
    "lf_router_get_discharge": (C.c_int, [_vp, C.c_int, _vp]),
    "lf_router_set_inflow": (C.c_int, [_vp, C.c_int, _vp]),
    "lf_router_run": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.POINTER(C.c_int)]),
    "lf_router_set_exchange": (C.c_int, [_vp, _vp, _vp, C.c_int32, _vp, _vp, _vp, C.c_int32, _i64s, C.c_int32]),
    "lf_graph_partition": (C.c_int, [_vp, C.c_int32, C.c_double, _vp, _vp, C.POINTER(_i64s), C.POINTER(_i64s)]),
    "lf_graph_cut_edges": (C.c_int, [_vp, _vp, _i64s, _vp, _vp, C.POINTER(_i64s)]),
    "lf_graph_restrict": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "lf_xchg_create": (C.c_int, [C.c_int32, C.c_int32, _i64s, C.POINTER(_vp)]),
    "lf_xchg_ipc_handle": (C.c_int, [_vp, C.c_char_p]),
    "lf_xchg_open_peer": (C.c_int, [_vp, C.c_int32, C.c_char_p]),
    "lf_xchg_set_peer_local": (C.c_int, [_vp, C.c_int32, _vp]),
    "lf_xchg_peer_base": (C.c_int, [_vp, C.c_int32, C.POINTER(C.c_uint64)]),
    "lf_xchg_begin": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "lf_xchg_end": (C.c_int, [_vp]),
    "lf_xchg_status": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(_i64s)]),
    "lf_xchg_destroy": (None, [_vp]),
    "lf_model_create_from_graphs": (C.c_int, [_vp, _vp, _vp, C.POINTER(_vp)]),
    "lf_model_set_exchange": (C.c_int, [_vp, _vp, C.c_int32, _vp, C.c_int32, _vp, _vp, _vp, C.c_int32, _i64s]),
    "lf_router_set_option": (C.c_int, [_vp, C.c_char_p, C.c_double]),
    "lf_router_destroy": (None, [_vp]),
    "lf_model_create": (C.c_int, [_vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "lf_model_info": (C.c_int, [_vp, C.POINTER(_i64s), C.POINTER(_i64s), C.POINTER(_i64s), C.POINTER(_i64s),
                                C.POINTER(_i64s)]),
    "lf_model_set": (C.c_int, [_vp, C.c_char_p, _vp, _i64s]),
    "lf_model_set_async": (C.c_int, [_vp, C.c_char_p, _vp, _i64s]),
    "lf_model_get": (C.c_int, [_vp, C.c_char_p, _vp, _i64s]),
    "lf_model_get_async": (C.c_int, [_vp, C.c_char_p, _vp, _i64s]),
    "lf_model_get_async_f32": (C.c_int, [_vp, C.c_char_p, _vp, _i64s]),
    "lf_model_wait_outputs": (C.c_int, [_vp]),
    "lf_model_set_flags": (C.c_int, [_vp, C.c_char_p, _vp, _i64s]),
    "lf_model_soil": (C.c_int, [_vp]),
    "lf_model_surface_routing": (C.c_int, [_vp]),
    "lf_model_channel": (C.c_int, [_vp]),
    "lf_model_step": (C.c_int, [_vp]),
    "lf_model_stage_times": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(_i64s)]),
    "lf_model_soil_stats": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "lf_model_set_scalar": (C.c_int, [_vp, C.c_char_p, C.c_double]),
    "lf_model_feed": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32]),
    "lf_model_feed_packed": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_double, C.c_double, C.c_double,
                                       C.c_int32]),
    "lf_model_set_lai": (C.c_int, [_vp, _vp, _i64s]),
    "lf_model_set_structures": (C.c_int, [_vp, C.c_int32, _vp, C.c_int32, _vp]),
    "lf_model_structure_array": (C.c_int, [_vp, C.c_char_p, _vp, _i64s, C.c_int32]),
    "lf_model_set_option": (C.c_int, [_vp, C.c_char_p, C.c_double]),
    "lf_model_nonfinite": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "lf_model_destroy": (None, [_vp]),
    "lf_interception_water_balance": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double, _i64s, _i64s]),
    "lf_soil_columns_water_balance": (C.c_int, [_vp]),
    "lf_suction_unsaturated_soil_pf": (C.c_int, [_vp]),
}


class ModelConfig(C.Structure):
    """struct lf_model_config (include/lisflood_b200.h)."""
    _fields_ = [("rows", C.c_int64), ("cols", C.c_int64), ("DtSec", C.c_double), ("Beta", C.c_double),
                ("PixelLength", C.c_double), ("NoRoutSteps", C.c_int32), ("SplitRouting", C.c_int32),
                ("CourantCrit", C.c_double), ("AvWaterThreshold", C.c_double), ("LeafDrainageK", C.c_double),
                ("DrainedFraction", C.c_double), ("SMaxSealed", C.c_double), ("diagnostics", C.c_int32),
                ("reserved", C.c_int32)]

_D, _U, _I = C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int64)


class SoilColumnsArgs(C.Structure):
    """struct lf_soil_columns_args (include/lisflood_b200.h): the 73 arguments of soilColumnsWaterBalance."""
    _fields_ = ([("num_vegs", C.c_int64), ("num_pixs", C.c_int64), ("num_landuses", C.c_int64), ("index_landuse_all", _I),
                 ("is_irrigated", _U), ("is_paddy_irrig", _U), ("DtDay", C.c_double),
                 ("AvailableWaterForInfiltration", _D), ("Rain", _D), ("SnowMelt", _D), ("LeafDrainage", _D),
                 ("Interception", _D), ("DSLR", _D), ("AvWaterThreshold", C.c_double), ("ESAct", _D), ("ESMax", _D),
                 ("isFrozenSoil", _U), ("b_Xinanjiang", _D), ("StoreMaxPervious", _D), ("PowerInfPot", _D),
                 ("PrefFlow", _D), ("PowerPrefFlow", _D), ("Infiltration", _D), ("CourantCrit", C.c_double),
                 ("PoreSpaceNotZero1a", _U), ("PoreSpaceNotZero1b", _U), ("PoreSpaceNotZero2", _U)]
                + [(k, _D) for k in ("KSat1a", "KSat1b", "KSat2", "GenuInvM1a", "GenuInvM1b", "GenuInvM2", "GenuM1a", "GenuM1b",
                                     "GenuM2", "W1a", "W1b", "W1", "W2", "Theta1a", "Theta1b", "Theta2", "Sat1a", "Sat1b",
                                     "Sat1", "Sat2", "SeepTopToSubA", "SeepTopToSubB", "SeepSubToGW", "WRes1a", "WRes1b",
                                     "WRes1", "WRes2", "WWP1a", "WWP1b", "WWP1", "WWP2", "WFC1a", "WFC1b", "WFC1", "WFC2",
                                     "SoilDepth1a", "SoilDepth1b", "SoilDepth2", "WS1a", "WS1b", "WS1", "WS2",
                                     "UpperZoneK")]
                + [("DrainedFraction", C.c_double), ("GwPercStep", _D), ("UZOutflow", _D), ("UZ", _D), ("GwPercUZLZ", _D),
                   ("NoSubS_out", _I)])


class SoilPfArgs(C.Structure):
    """struct lf_soil_pf_args (include/lisflood_b200.h): the arguments of suctionUnsaturatedSoilPF, layers 1a, 1b, 2."""
    _fields_ = [("num_vegs", C.c_int64), ("num_pixs", C.c_int64), ("num_landuses", C.c_int64), ("index_landuse_all", _I),
                ("pF", _D * 3), ("W", _D * 3), ("WRes", _D * 3), ("WS", _D * 3), ("PoreSpaceNotZero", _U * 3),
                ("GenuInvAlpha", _D * 3), ("GenuInvM", _D * 3), ("GenuInvN", _D * 3), ("HeadMax", C.c_double)]


_lib = None


def build(force=False):
    """Compiles liblisf_b200.so in-tree with nvcc for sm_100a (no GPU needed)."""
    args = ["make", "-s", "-C", CSRC, "-j4"]
    if force:
        args.append("-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LisfloodB200Error(LF_ERR_NO_DEVICE,
                                    "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                                    " (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != LF_OK:
        raise LisfloodB200Error(rc, lib().lf_last_error().decode("utf-8", "replace"))


def ptr(a):
    """ctypes pointer of a NumPy array, a torch tensor (host or CUDA) or a raw address."""
    if a is None:
        return None
    if isinstance(a, int):
        return _vp(a)
    if hasattr(a, "data_ptr"):
        if getattr(a, "is_cuda", False):
            # the library works on its own stream: whatever torch kernels produce / still read this tensor
            # must have finished before the pointer is used
            import torch
            torch.cuda.current_stream(a.device).synchronize()
        return _vp(a.data_ptr())
    return a.ctypes.data_as(_vp)


class _PinnedBlock(object):
    """cudaHostAlloc'ed memory, freed when the last array that views it is gone."""

    def __init__(self, nbytes):
        self.ptr = _vp()
        check(lib().lf_host_alloc(int(nbytes), C.byref(self.ptr)))
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr and _lib is not None:
                _lib.lf_host_free(self.ptr)
        except Exception:
            pass


class PinnedArray(np.ndarray):
    """NumPy array in page-locked host memory owned by the library (views keep the block alive through `.base`)."""
    _block = None


def pinned_empty(n, dtype=np.float64):
    dt = np.dtype(dtype)
    blk = _PinnedBlock(max(int(n) * dt.itemsize, 1))
    buf = (C.c_char * blk.nbytes).from_address(blk.ptr.value)
    a = np.frombuffer(buf, dt, count=int(n)).view(PinnedArray)
    a._block = blk
    return a


def device_info():
    L = lib()
    name = C.create_string_buffer(128)
    sms, maj, mnr, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
    check(L.lf_device_info(name, C.byref(sms), C.byref(maj), C.byref(mnr), C.byref(mem)))
    return {"name": name.value.decode(), "sm_count": sms.value, "cc": (maj.value, mnr.value), "hbm_bytes": mem.value}


def synchronize():
    check(lib().lf_synchronize())


def timer_start():
    check(lib().lf_timer_start())


def timer_stop():
    ms = C.c_double()
    check(lib().lf_timer_stop(C.byref(ms)))
    return ms.value


def launch_count(reset=False):
    return int(lib().lf_launch_count(1 if reset else 0))


def host_launch_count():
    """Launch API calls since the last launch_count(reset=True) (a replayed CUDA graph counts once)."""
    return int(lib().lf_host_launch_count())
