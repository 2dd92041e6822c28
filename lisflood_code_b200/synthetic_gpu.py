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


def _ldd_gpu(torch, rows, cols, seed, noise, tilt=1.0):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    elev = torch.randn((rows, cols), generator=g, device="cuda", dtype=torch.float32) * noise
    elev += tilt * torch.arange(rows - 1, -1, -1, device="cuda", dtype=torch.float32)[:, None]
    elev += 0.05 * tilt * (torch.arange(cols, device="cuda", dtype=torch.float32) - cols / 2).abs()[None, :]
    big = 3.0e38
    pad = torch.full((rows + 2, cols + 2), big, device="cuda", dtype=torch.float32)
    pad[1:-1, 1:-1] = elev
    best = torch.zeros((rows, cols), device="cuda", dtype=torch.float32)
    code = torch.full((rows, cols), 5.0, device="cuda", dtype=torch.float64)
    for dr, dc, k in [(-1, -1, 7), (-1, 0, 8), (-1, 1, 9), (0, -1, 4), (0, 1, 6), (1, -1, 1), (1, 0, 2), (1, 1, 3)]:
        nb = pad[1 + dr:1 + dr + rows, 1 + dc:1 + dc + cols]
        drop = (elev - nb) / math.sqrt(dr * dr + dc * dc)
        drop = torch.where(nb >= big, torch.full_like(drop, -1.0), drop)
        better = drop > best
        best = torch.where(better, drop, best)
        code = torch.where(better, torch.full_like(code, float(k)), code)
        del nb, drop, better
    del pad, elev, best
    return code.reshape(-1)


class C3Device(object):
    """Builds a HotPathModel for a rows x cols catchment entirely on the device."""

    def __init__(self, rows, cols, seed=0, ldd_noise=0.5, channel_threshold=60, no_rout_steps=24, dt_sec=86400.0,
                 diagnostics=False):
        import torch
        from . import _capi
        from .hotpath import HotPathModel
        self.torch = torch
        L = _capi.lib()
        n = rows * cols
        self.n, self.rows, self.cols = n, rows, cols
        g = torch.Generator(device="cuda")
        g.manual_seed(seed + 4242)
        self.gen = g
        U = lambda lo, hi, shape=(n,): torch.rand(shape, generator=g, device="cuda", dtype=torch.float64) * (hi - lo) + lo
        ldd = _ldd_gpu(torch, rows, cols, seed, ldd_noise)
        mask = torch.ones(n, dtype=torch.uint8, device="cuda")
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
        del ones
        is_chan = uparea >= channel_threshold
        ldd_kin = torch.where(is_chan, ldd, torch.zeros_like(ldd))
        ldd_toc = torch.where(is_chan, torch.full_like(ldd, 5.0), ldd)
        S = {"rows": rows, "cols": cols, "N": n, "mask_device": mask, "LddToChan": ldd_toc, "LddKinematic": ldd_kin,
             "DtSec": dt_sec, "Beta": 0.6, "PixelLength": 5000.0, "NoRoutSteps": no_rout_steps, "SplitRouting": False,
             "CourantCrit": 0.4, "AvWaterThreshold": 5.0 * dt_sec / 86400.0, "LeafDrainageK": min(dt_sec / 86400.0, 1.0),
             "DrainedFraction": 0.0, "SMaxSealed": 1.0}
        self.S = S
        M = HotPathModel(S, diagnostics=diagnostics)
        self.model = M
        del ldd_kin, ldd_toc
        dtday = dt_sec / 86400.0
        beta, alppow = 0.6, 2.0 / 3.0 * 0.6
        # ---- fractions (normalised gammas == Dirichlet) ----
        conc = torch.tensor([4.0, 3.0, 1.0, 0.6, 0.3], device="cuda", dtype=torch.float64)
        fr = torch._standard_gamma(conc[:, None].expand(5, n).contiguous())
        fr = fr / fr.sum(0, keepdim=True)
        M.set("SoilFraction", fr[:3].contiguous(), 3)
        M.set("DirectRunoffFraction", fr[3].contiguous())
        M.set("WaterFraction", fr[4].contiguous())
        del fr

        def lu3(a, b):
            return torch.stack([a, b, a]).contiguous()

        # ---- soil hydraulic parameters (soil.py:109-228); Irrigated shares the Rainfed maps ----
        ws1 = wfc = None
        store = {}
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
            M.set("KSat" + lay, ks, 3)
            M.set("GenuInvM" + lay, (1 / gm).contiguous(), 3)
            M.set("WRes" + lay, wres, 3)
            M.set("WS" + lay, ws, 3)
            if lay != "2" or diagnostics:
                M.set("WWP" + lay, wwp_l, 3)
                M.set("WFC" + lay, wfc_l, 3)
            if diagnostics:
                M.set("SoilDepth" + lay, depth, 3)
            # initial soil moisture ~ field capacity (soil.py:268-277)
            w = torch.minimum(wfc_l * U(0.7, 1.1, (3, n)), ws)
            M.set("W" + lay, w.contiguous(), 3)
            del depth, ths, thr, lam, gal, ks, gn, gm, ws, wres, wfc_l, wwp_l, w
        M.set("b_Xinanjiang", U(.1, .7))
        M.set("PowerPrefFlow", U(1.0, 5.0))
        M.set("CropCoef", torch.stack([U(.9, 1.1), U(.9, 1.3), U(.9, 1.2)]).contiguous(), 3)
        M.set("CropGroupNumber", torch.stack([U(1.0, 5.0), U(2.0, 5.0), U(1.0, 5.0)]).contiguous(), 3)
        # ---- groundwater ----
        M.set("UpperZoneK", torch.clamp(dtday * (1 / U(5.0, 20.0)), max=1.0))
        M.set("LowerZoneK", torch.clamp(dtday * (1 / U(50.0, 500.0)), max=1.0))
        M.set("GwPercStep", U(0.2, 1.5) * dtday)
        M.set("GwLossStep", torch.zeros(n, dtype=torch.float64, device="cuda"))
        M.set("LZThreshold", U(0.0, 20.0))
        M.set("UZ", U(0.0, 10.0, (3, n)), 3)
        M.set("LZ", U(20.0, 200.0))
        M.set("DSLR", torch.floor(U(1.0, 6.0, (3, n))), 3)
        M.set("CumInterception", U(0.0, 0.5, (3, n)), 3)
        M.set("CumInterSealed", U(0.0, 0.5))
        pa = torch.full((n,), 5000.0 ** 2, dtype=torch.float64, device="cuda")
        M.set("MMtoM3", 0.001 * pa)
        if diagnostics:
            M.set("PixelArea", pa)
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
        M.set("ChanLength", chan_len)
        M.set("ChannelAlpha", alpha)
        M.set("ChanM3Kin", half_bank * chan_len)
        qk = (half_bank / alpha) ** (1 / beta)
        M.set("ChanQKin", qk)
        M.set("ChanQ", qk)
        del chan_len, grad, man, width, dthr, upper, half_bank, wd, wp, alpha, qk, uparea
        # ---- overland flow (surface_routing.py:69-83) ----
        ograd = torch.clamp(U(1e-3, 0.1), min=1e-4)
        nman = torch.stack([U(0.05, 0.2), U(0.1, 0.4), torch.full((n,), 0.02, dtype=torch.float64, device="cuda")])
        M.set("OFAlpha", (((nman / torch.sqrt(ograd)) ** beta) * ((5000.0 + 2 * 0.001 * 5.0) ** alppow)).contiguous(), 3)
        del ograd, nman
        is_chan_u8 = is_chan.to(torch.uint8)
        M.set_flags("IsChannel", is_chan_u8)
        M.set_flags("IsChannelKinematic", is_chan_u8)
        M.set_flags("AtLastPointC", (ldd == 5.0).to(torch.uint8))
        self.channel_fraction = float(is_chan.double().mean().item())
        del is_chan, is_chan_u8, ldd
        torch.cuda.empty_cache()

    def forcing_device(self, step):
        """Forcing of one step as CUDA tensors (device-resident leg)."""
        torch, n, g = self.torch, self.n, self.gen
        R = lambda shape=(n,): torch.rand(shape, generator=g, device="cuda", dtype=torch.float64)
        dtday = self.S["DtSec"] / 86400.0
        rain = torch.where(R() < 0.45, torch._standard_gamma(torch.full((n,), 0.8, device="cuda", dtype=torch.float64)) * 8.0,
                           torch.zeros(n, device="cuda", dtype=torch.float64)) * dtday
        F = {"Rain": rain, "SnowMelt": torch.where(R() < 0.1, R() * 3.0, torch.zeros_like(rain)) * dtday,
             "ETRef": R() * 6.0 * dtday, "EWRef": R() * 6.0 * dtday, "LAI": R((3, n)) * 6.0,
             "isFrozenSoil": (R() < 0.05).to(torch.uint8)}
        F["ESRef"] = (F["EWRef"] + F["ETRef"]) / 2
        F["LAITerm"] = torch.exp(-(0.75 * 0.72) * F["LAI"])
        return F
