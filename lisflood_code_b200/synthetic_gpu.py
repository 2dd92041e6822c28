"""Device-side generation of the C3 synthetic catchment (bench.py only).

Same distributions as synthetic.full_stack / synthetic.forcing, but produced with torch on
the GPU so that a 10000x10000 raster (1e8 cells, ~110 maps) is ready in seconds and never exists on the
host.  torch is plumbing here (device memory + RNG); every map is handed to the library as a device pointer
in the reference's compressed order.  The drainage network (steepest descent on tilted noise), the upstream
area (lf_graph_accuflux) and the derived soil parameters follow the same formulas as the host generator.
"""
import ctypes as C
import math

import numpy as np


def _ldd_gpu(torch, rows, cols, seed, noise, tilt=1.0, device="cuda", single_outlet=False):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    elev = torch.randn((rows, cols), generator=g, device=device, dtype=torch.float32) * noise
    elev += tilt * torch.arange(rows - 1, -1, -1, device=device, dtype=torch.float32)[:, None]
    elev += 0.05 * tilt * (torch.arange(cols, device=device, dtype=torch.float32) - cols / 2).abs()[None, :]
    big = 3.0e38
    pad = torch.full((rows + 2, cols + 2), big, device=device, dtype=torch.float32)
    pad[1:-1, 1:-1] = elev
    best = torch.zeros((rows, cols), device=device, dtype=torch.float32)
    code = torch.full((rows, cols), 5.0, device=device, dtype=torch.float64)
    for dr, dc, k in [(-1, -1, 7), (-1, 0, 8), (-1, 1, 9), (0, -1, 4), (0, 1, 6), (1, -1, 1), (1, 0, 2), (1, 1, 3)]:
        nb = pad[1 + dr:1 + dr + rows, 1 + dc:1 + dc + cols]
        drop = (elev - nb) / math.sqrt(dr * dr + dc * dc)
        drop = torch.where(nb >= big, torch.full_like(drop, -1.0), drop)
        better = drop > best
        best = torch.where(better, drop, best)
        code = torch.where(better, torch.full_like(code, float(k)), code)
        del nb, drop, better
    del pad, elev, best
    if single_outlet:
        # one big basin (synthetic.random_ldd): the south edge becomes a collector draining to its centre cell, the only
        # outlet of everything that reaches the edge; interior sinks remain separate small catchments
        c0 = cols // 2
        code[rows - 1, :c0] = 6.0
        code[rows - 1, c0 + 1:] = 4.0
        code[rows - 1, c0] = 5.0
    return code.reshape(-1)


def _host_accuflux_ones(ldd, rows, cols):
    """Upstream area (cells) on the host: the library-free twin of lf_graph_accuflux(ones) for the CPU arm."""
    from .global_modules import ldd_ops
    mask = np.ones((rows, cols), bool)
    ds = ldd_ops.downstream_index(ldd, mask)
    return ldd_ops.accuflux(ds, np.ones(rows * cols))


def c3_generate(torch, rows, cols, seed, emit, ldd_noise=0.5, channel_threshold=60, no_rout_steps=24, dt_sec=86400.0,
                diagnostics=False, accuflux=None, device="cuda", single_outlet=False):
    """The C3 generator: every map of the catchment, one after the other, handed to `emit(kind, name, value)`.

    kind "config": value = dict of the scalars + the mask / LddToChan / LddKinematic tensors (first call);
    kind "map": float64 tensor, (n,) or (3, n), compressed order; kind "flags": uint8 (n,).
    The same code feeds the device model of bench.py (C3Device) and the host dictionary of the CPU arm / the
    parity test (c3_host_stack): same distributions, same seeds, same random streams on the same torch device.
    accuflux(ldd_tensor) -> upstream area tensor; default: the host operator."""
    n = rows * cols
    g = torch.Generator(device=device)
    g.manual_seed(seed + 4242)
    U = lambda lo, hi, shape=(n,): torch.rand(shape, generator=g, device=device, dtype=torch.float64) * (hi - lo) + lo
    ldd = _ldd_gpu(torch, rows, cols, seed, ldd_noise, device=device, single_outlet=single_outlet)
    mask = torch.ones(n, dtype=torch.uint8, device=device)
    if accuflux is None:
        uparea = torch.from_numpy(_host_accuflux_ones(ldd.cpu().numpy(), rows, cols)).to(device)
    else:
        uparea = accuflux(ldd, mask)
    is_chan = uparea >= channel_threshold
    ldd_kin = torch.where(is_chan, ldd, torch.zeros_like(ldd))
    ldd_toc = torch.where(is_chan, torch.full_like(ldd, 5.0), ldd)
    S = {"rows": rows, "cols": cols, "N": n, "mask_device": mask, "Ldd": ldd, "LddToChan": ldd_toc, "LddKinematic": ldd_kin,
         "DtSec": dt_sec, "Beta": 0.6, "PixelLength": 5000.0, "NoRoutSteps": no_rout_steps, "SplitRouting": False,
         "CourantCrit": 0.4, "AvWaterThreshold": 5.0 * dt_sec / 86400.0, "LeafDrainageK": min(dt_sec / 86400.0, 1.0),
         "DrainedFraction": 0.0, "SMaxSealed": 1.0}
    emit("config", "S", S)
    del ldd_kin, ldd_toc
    put = lambda name, t: emit("map", name, t.contiguous())
    dtday = dt_sec / 86400.0
    beta, alppow = 0.6, 2.0 / 3.0 * 0.6
    # ---- fractions (normalised gammas == Dirichlet) ----
    conc = torch.tensor([4.0, 3.0, 1.0, 0.6, 0.3], device=device, dtype=torch.float64)
    fr = torch._standard_gamma(conc[:, None].expand(5, n).contiguous(), generator=g)
    fr = fr / fr.sum(0, keepdim=True)
    put("SoilFraction", fr[:3])
    put("DirectRunoffFraction", fr[3])
    put("WaterFraction", fr[4])
    del fr

    def lu3(a, b):
        return torch.stack([a, b, a]).contiguous()

    # ---- soil hydraulic parameters (soil.py:109-228); Irrigated shares the Rainfed maps ----
    for lay, (d_lo, d_hi) in (("1a", (40, 60)), ("1b", (200, 300)), ("2", (500, 900))):
        if lay == "2":
            x = U(d_lo, d_hi)
            depth = lu3(x, x)
            mk = lambda lo, hi, log=False: (lambda t: lu3(t, t))(torch.exp(U(math.log(lo), math.log(hi))) if log else U(lo, hi))
        else:
            depth = lu3(U(d_lo, d_hi), U(d_lo, d_hi))
            mk = lambda lo, hi, log=False: lu3(torch.exp(U(math.log(lo), math.log(hi))) if log else U(lo, hi),
                                               torch.exp(U(math.log(lo), math.log(hi))) if log else U(lo, hi))
        ths, thr, lam, gal, ks = mk(.4, .5), mk(.02, .08), mk(.15, .45), mk(.005, .05), mk(1.0, 500.0, True)
        gn = 1 + lam
        gm = lam / gn
        ws, wres = ths * depth, thr * depth
        mual = lambda h: wres + (ws - wres) / ((1 + (gal * h) ** gn) ** gm)
        wfc_l, wwp_l = mual(100), mual(15000)
        put("KSat" + lay, ks)
        put("GenuInvM" + lay, 1 / gm)
        put("WRes" + lay, wres)
        put("WS" + lay, ws)
        if lay != "2" or diagnostics:
            put("WWP" + lay, wwp_l)
            put("WFC" + lay, wfc_l)
        if diagnostics:
            put("SoilDepth" + lay, depth)
        # initial soil moisture ~ field capacity (soil.py:268-277)
        w = torch.minimum(wfc_l * U(0.7, 1.1, (3, n)), ws)
        put("W" + lay, w)
        del depth, ths, thr, lam, gal, ks, gn, gm, ws, wres, wfc_l, wwp_l, w
    put("b_Xinanjiang", U(.1, .7))
    put("PowerPrefFlow", U(1.0, 5.0))
    put("CropCoef", torch.stack([U(.9, 1.1), U(.9, 1.3), U(.9, 1.2)]))
    put("CropGroupNumber", torch.stack([U(1.0, 5.0), U(2.0, 5.0), U(1.0, 5.0)]))
    # ---- groundwater ----
    put("UpperZoneK", torch.clamp(dtday * (1 / U(5.0, 20.0)), max=1.0))
    put("LowerZoneK", torch.clamp(dtday * (1 / U(50.0, 500.0)), max=1.0))
    put("GwPercStep", U(0.2, 1.5) * dtday)
    put("GwLossStep", torch.zeros(n, dtype=torch.float64, device=device))
    put("LZThreshold", U(0.0, 20.0))
    put("UZ", U(0.0, 10.0, (3, n)))
    put("LZ", U(20.0, 200.0))
    put("DSLR", torch.floor(U(1.0, 6.0, (3, n))))
    put("CumInterception", U(0.0, 0.5, (3, n)))
    put("CumInterSealed", U(0.0, 0.5))
    pa = torch.full((n,), 5000.0 ** 2, dtype=torch.float64, device=device)
    put("MMtoM3", 0.001 * pa)
    if diagnostics:
        put("PixelArea", pa)
    del pa
    # ---- channel geometry (routing.py:184-253) ----
    chan_len = 5000.0 * U(1.0, 1.4)
    grad = torch.clamp(U(1e-4, 5e-3), min=1e-5)
    man = U(0.02, 0.06)
    width = 2.0 + 0.5 * torch.sqrt(uparea)
    dthr = 0.5 + 0.05 * torch.sqrt(uparea)
    upper = width + 2 * 1.0 * dthr
    half_bank = 0.5 * (0.5 * dthr * (upper + width))
    wd = torch.where(is_chan, 0.5 * dthr, torch.zeros_like(dthr))
    wp = width + 2 * torch.sqrt(wd * wd + (wd * 1.0) ** 2)
    alpha = ((man / torch.sqrt(grad)) ** beta) * (wp ** alppow)
    put("ChanLength", chan_len)
    put("ChannelAlpha", alpha)
    put("ChanM3Kin", half_bank * chan_len)
    qk = (half_bank / alpha) ** (1 / beta)
    put("ChanQKin", qk)
    put("ChanQ", qk)
    del chan_len, grad, man, width, dthr, upper, half_bank, wd, wp, alpha, qk, uparea
    # ---- overland flow (surface_routing.py:69-83) ----
    ograd = torch.clamp(U(1e-3, 0.1), min=1e-4)
    nman = torch.stack([U(0.05, 0.2), U(0.1, 0.4), torch.full((n,), 0.02, dtype=torch.float64, device=device)])
    put("OFAlpha", ((nman / torch.sqrt(ograd)) ** beta) * ((5000.0 + 2 * 0.001 * 5.0) ** alppow))
    del ograd, nman
    is_chan_u8 = is_chan.to(torch.uint8)
    emit("flags", "IsChannel", is_chan_u8)
    emit("flags", "IsChannelKinematic", is_chan_u8)
    emit("flags", "AtLastPointC", (ldd == 5.0).to(torch.uint8))
    emit("stat", "channel_fraction", float(is_chan.double().mean().item()))
    # ---- feeder modules (snow.py:53-93, frost.py:44-58, miscInitial.py:142-143,181, leafarea.py:48): LISFLOOD's usual
    #      scalars, maps for the sub-pixel elevation spread and the latitude; no snow pack / frost at the start ----
    emit("feeder", "parameters", {
        "PrScaling": 1.0, "CalEvaporation": 1.0, "DeltaTSnow": 0.9674 * U(0.0, 300.0) * 0.0065, "SnowSeason": 1.0 * 0.5,
        "TempSnow": 1.0, "SnowFactor": 1.0, "SnowMeltCoef": 4.0, "TempMelt": 0.0, "lat_rad": U(35.0, 70.0) * (math.pi / 180.0),
        "Kfrost": 0.57, "Afrost": 0.97, "FrostIndexThreshold": 56.0, "SnowWaterEquivalent": 0.45, "kgb": 0.75 * 0.72})
    emit("feeder", "state", {"SnowCoverS": torch.zeros((3, n), dtype=torch.float64, device=device),
                             "FrostIndex": torch.zeros(n, dtype=torch.float64, device=device)})
    return g


def c3_forcing(torch, g, n, device="cuda"):
    """RAW meteo maps of one model step from the generator's random stream, float32 like the NetCDF forcing of the
    reference (tensors on `device`): intermittent Gamma precipitation [mm/day], a cool-season temperature field (snow falls
    and melts somewhere on most steps, the frost index builds up slowly), ET0 / E0 [mm/day]."""
    R = lambda shape=(n,): torch.rand(shape, generator=g, device=device, dtype=torch.float64)
    rain = torch.where(R() < 0.45, torch._standard_gamma(torch.full((n,), 0.8, device=device, dtype=torch.float64),
                                                         generator=g) * 8.0,
                       torch.zeros(n, device=device, dtype=torch.float64))
    return {"Precipitation": rain.to(torch.float32), "Tavg": (R() * 24.0 - 6.0).to(torch.float32),
            "ET0": (R() * 6.0).to(torch.float32), "E0": (R() * 6.0).to(torch.float32)}


def c3_lai(torch, g, n, device="cuda"):
    """LAI maps (3, n) of a 10-day interval."""
    return torch.rand((3, n), generator=g, device=device, dtype=torch.float64) * 6.0


def c3_host_stack(rows, cols, seed=0, nforcing=0, device=None, **kw):
    """The same generator collected into host NumPy arrays: (S, feeder, lai, [raw0, raw1, ...]) with S complete for the
    CPU restatement of the step (synthetic.complete_stack adds the derived maps), feeder = (parameters, state) of the feeder
    modules, lai the (3, N) LAI maps, raw_k the float32 meteo maps of step k.  No library call: usable by the CPU arm of
    bench.py.  device: torch device of the random streams (default: cuda when available -- the streams of the bench's own
    raster)."""
    import torch
    from . import synthetic
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    S, feeder = {}, {}
    host = lambda v: v.cpu().numpy() if hasattr(v, "cpu") else v

    def emit(kind, name, value):
        if kind == "config":
            for k, v in value.items():
                if k == "mask_device":
                    S["mask"] = v.cpu().numpy().astype(bool).reshape(rows, cols)
                else:
                    S[k] = host(v)
        elif kind == "map":
            S[name] = value.cpu().numpy()
        elif kind == "flags":
            S[name] = value.cpu().numpy().astype(bool)
        elif kind == "feeder":
            feeder[name] = {k: host(v) for k, v in value.items()}

    kw.setdefault("diagnostics", True)    # the CPU restatement wants every parameter map
    g = c3_generate(torch, rows, cols, seed, emit, device=device, **kw)
    S = synthetic.complete_stack(S)
    lai = c3_lai(torch, g, rows * cols, device).cpu().numpy()
    raw = [{k: v.cpu().numpy() for k, v in c3_forcing(torch, g, rows * cols, device).items()} for _ in range(nforcing)]
    S["rng_device"] = device
    return S, (feeder["parameters"], feeder["state"]), lai, raw


class C3Device(object):
    """Builds the model of a rows x cols catchment entirely on the device: a HotPathModel, or -- under torch.distributed
    with more than one rank and distributed=True -- this rank's part of the SAME raster cut along its drainage graph
    (parallel.DistributedHotPathModel; every rank generates the same global maps from the same seeds and keeps its own
    pixels).  keep_host=True also keeps a host copy of everything handed to the model (parity test of the bench's own data
    on a crop-sized raster)."""

    def __init__(self, rows, cols, seed=0, ldd_noise=0.5, channel_threshold=60, no_rout_steps=24, dt_sec=86400.0,
                 diagnostics=False, keep_host=False, distributed=False, single_outlet=False):
        import torch
        from . import _capi
        from .hotpath import HotPathModel
        self.torch = torch
        L = _capi.lib()
        n = rows * cols
        self.n, self.rows, self.cols = n, rows, cols
        self.host = {} if keep_host else None
        self.feeder_host = {}
        self.distributed = bool(distributed)
        host = lambda v: v.cpu().numpy() if hasattr(v, "cpu") else v

        def accuflux(ldd, mask):
            # upstream area on the full LDD -> channel mask (routing.py:98, 110-118)
            gh = C.c_void_p()
            _capi.check(L.lf_ldd_build(_capi.ptr(ldd), _capi.ptr(mask), rows, cols, C.byref(gh)))
            ones = torch.ones(n, dtype=torch.float64, device="cuda")
            uparea = torch.empty(n, dtype=torch.float64, device="cuda")
            _capi.check(L.lf_graph_accuflux(gh, _capi.ptr(ones), _capi.ptr(uparea)))
            no, k, npx, pits = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
            _capi.check(L.lf_graph_info(gh, C.byref(npx), C.byref(no), C.byref(k), C.byref(pits)))
            self.ldd_levels, self.ldd_pits = no.value, pits.value
            L.lf_graph_destroy(gh)
            return uparea

        def emit(kind, name, value):
            if kind == "config":
                self.S = {k: v for k, v in value.items() if k != "Ldd"}
                if self.distributed:
                    from .parallel import DistributedHotPathModel
                    self.model = DistributedHotPathModel(value, diagnostics=diagnostics)
                else:
                    self.model = HotPathModel(self.S, diagnostics=diagnostics)
                if keep_host:
                    for k, v in value.items():
                        if k == "mask_device":
                            self.host["mask"] = v.cpu().numpy().astype(bool).reshape(rows, cols)
                        else:
                            self.host[k] = host(v)
            elif kind == "map":
                self.model.set(name, value, 3 if value.dim() == 2 else 1)
                if keep_host:
                    self.host[name] = value.cpu().numpy()
            elif kind == "flags":
                self.model.set_flags(name, value)
                if keep_host:
                    self.host[name] = value.cpu().numpy().astype(bool)
            elif kind == "feeder":
                if name == "parameters":
                    self.model.set_feeder(value)
                else:
                    self.model.set_feeder({}, value)
                if keep_host:
                    self.feeder_host[name] = {k: host(v) for k, v in value.items()}
            elif kind == "stat":
                setattr(self, name, value)

        self.gen = c3_generate(torch, rows, cols, seed, emit, ldd_noise=ldd_noise, channel_threshold=channel_threshold,
                               no_rout_steps=no_rout_steps, dt_sec=dt_sec, diagnostics=diagnostics or keep_host,
                               accuflux=accuflux, single_outlet=single_outlet)
        self.n_local = self.model.n_local if self.distributed else n
        torch.cuda.empty_cache()

    def _pick(self, t):
        return self.model._local(t) if self.distributed else t

    def forcing_device(self, step=None):
        """RAW meteo maps of one step (float32 CUDA tensors: Precipitation, Tavg, ET0, E0) -- this rank's pixels."""
        return {k: self._pick(v) for k, v in c3_forcing(self.torch, self.gen, self.n).items()}

    def lai_device(self):
        return self._pick(c3_lai(self.torch, self.gen, self.n))

    def host_stack(self):
        """Host copy of the model's inputs completed with the derived maps (keep_host=True)."""
        from . import synthetic
        return synthetic.complete_stack(dict(self.host))
