#!/usr/bin/env python
"""bench.py -- cell-updates/s of the LISFLOOD raster hot path on B200 (contract: see DESIGN.md §7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2]

Default workload C3 = BASELINE.json configs[2] (the configuration the metric and the >=50x target are quoted on):
synthetic 10000x10000 raster, full soil + infiltration + overland + channel stack, 24 channel sub-steps per
model step; a bench step = one model time step; cell-updates = cells x model steps.
Workload C2 = configs[1]: 2000x2000 raster, kinematic routing only.

One JSON line on stdout (rank 0).  `value` = device-resident throughput (inputs already in HBM),
`e2e` = the same work through the public plugin call with HOST buffers (H2D of the step's inflow map and
D2H of the resulting discharge map inside the timed region), `roofline` for the dominant kernel,
`cpu_baseline` = the CPU oracle (port of the reference algorithm) on this box's host cores.
"""
import argparse
import json
import os
os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's banner out of stdout: the JSON line must be the last line
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES_ROUTING = 44.0  # bytes per (cell, routing solve)
# Fused soil kernel, bytes per cell and model step as laid out in HBM: forcing 89 +
# per-pixel parameters 96 + de-duplicated land-use parameters 272 + state RW 288 + per-pixel state RW 112 +
# runoff outputs 32.  (The reference's layout moves 3*(500+64+160)+360 = 2532 B for the same work)
ALG_BYTES_SOIL = 889.0
ALG_BYTES_CHANNEL_SUBSTEP = 84.0  # fused channel sub-step, single routing


# ------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons of GPU `index` while the timed region runs: NVML in-process every 20 ms (so
    that a timed region of a few tenths of a second still gets samples), `nvidia-smi -lms 100` when NVML cannot be loaded."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    PERIOD_S = 0.02

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.thr, self.nvml, self.h = index, [], None, None, None, None
        self.stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical(index))
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _poll(self):
        nv = self.nvml
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = int(get(self.h))
                self.rows.append([mhz, self.mx] + ["Active" if r & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            self.stop.wait(self.PERIOD_S)

    def __enter__(self):
        self.rows = []
        self.stop.clear()
        try:
            if self.nvml is not None:
                self.thr = threading.Thread(target=self._poll, daemon=True)
            else:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                              "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                             stderr=subprocess.DEVNULL, text=True)
                self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = self.thr = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        elif self.thr:
            self.thr.join(timeout=1)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if str(v).lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def bind_to_gpu_numa(index):
    """Pins this process to the CPU cores NVML reports as local to GPU `index`, so that page-locked host buffers are
    allocated (first touch) on the NUMA node the GPU's PCIe root hangs off: with 8 ranks streaming forcing in and discharge
    out, cross-socket traffic otherwise caps the aggregate host bandwidth.  Best effort; returns the core count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cores = [64 * w + b for w, bits in enumerate(words) for b in range(64) if (bits >> b) & 1]
        if cores:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception:
        pass
    return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# Workload C2 (BASELINE.json configs[1]): 2000x2000 raster, random D8 tree, kinematic routing only
# ------------------------------------------------------------------------------------------------
class C2:
    name = "c2"
    rows = cols = 2000
    timesteps_per_step = 100  # one bench step = 100 kinematicWaveRouting timesteps with a fresh inflow map

    def __init__(self, rank, ldd_kind):
        from lisflood_code_b200 import synthetic
        self.noise = 0.3 if ldd_kind == "deep" else 3.0
        self.ldd_kind = ldd_kind
        self.ldd, self.mask = synthetic.random_ldd(self.rows, self.cols, seed=100 + rank, noise=self.noise)
        self.n = int(self.mask.sum())
        self.alpha, self.q0, self.q = synthetic.routing_fields(self.n, 100 + rank)
        self.dx, self.dt, self.beta = 5000.0, 3600.0, 0.6
        rng = np.random.default_rng(977 + rank)
        self.scales = rng.uniform(0.5, 1.5, (64, self.timesteps_per_step))  # seeded time-varying multiplier

    def describe(self, levels=None):
        return {"workload": "C2 synthetic %dx%d raster, random D8 %s tree (noise/tilt %.1f), kinematic routing only, "
                            "beta 0.6, dx 5000 m, dt 3600 s" % (self.rows, self.cols, self.ldd_kind, self.noise),
                "cells": self.n, "timesteps_per_step": self.timesteps_per_step, "levels": levels,
                "l2_policy": "state+parameters are 6 f64 maps of 32 MB = 192 MB > 126 MB L2, plus a 256 MB L2 flush "
                             "buffer is rewritten between timed steps in the e2e leg"}


def run_ours(args):
    rank, world, local = dist_env()
    from lisflood_code_b200 import _capi
    from lisflood_code_b200.hydrological_modules.kinematic_wave_parallel import kinematicWave
    L = _capi.lib()
    _capi.check(L.lf_device_init(local))
    use_dist = world > 1
    if use_dist:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = C2(rank, args.ldd)
    t0 = time.time()
    kw = kinematicWave(wl.ldd[wl.mask], wl.mask, wl.alpha, wl.beta, wl.dx, wl.dt)
    t_init = time.time() - t0
    tps = wl.timesteps_per_step
    K, W = args.steps, args.warmup

    def barrier():
        if use_dist:
            dist.barrier()
        _capi.synchronize()

    def max_over_ranks(x):
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg -------------------------------------------------------------------
    kw.set_discharge(wl.q0)
    kw.set_lateral_inflow(wl.q)
    for w in range(W):
        kw.run(tps, inflow_scale=wl.scales[w % 64])
    barrier()
    _capi.launch_count(reset=True)
    with ClockSampler(local) as clk:
        barrier()
        _capi.timer_start()
        for k in range(K):
            kw.run(tps, inflow_scale=wl.scales[(W + k) % 64])
        ms = _capi.timer_stop()
        barrier()
    launches = _capi.launch_count()
    ms = max_over_ranks(ms)
    total_cells = wl.n * world
    value = total_cells * tps * K / (ms * 1e-3)

    # ---- end-to-end leg: host buffers through the public API ---------------------------------
    if args.no_e2e:  # profiler runs only
        if rank == 0:
            print(json.dumps({"profile_only": True, "value": value, "ms_per_step": ms / K, "gpu_launches": launches}))
        return
    q_host = [np.ascontiguousarray(wl.q * f) for f in (1.0, 0.9, 1.1)]
    out_host = np.empty(wl.n)
    for a in q_host + [out_host]:
        _capi.check(L.lf_host_register(a.ctypes.data, a.nbytes))
    Ke = max(2, min(K, 5))
    for w in range(2):
        kw.set_lateral_inflow(q_host[w % 3])
        kw.run(tps, inflow_scale=wl.scales[w])
        kw.get_discharge(out=out_host)
    barrier()
    t0 = time.perf_counter()
    for k in range(Ke):
        kw.set_lateral_inflow(q_host[k % 3])            # H2D: this step's inflow map
        kw.run(tps, inflow_scale=wl.scales[k % 64])      # H2D: this step's multipliers
        kw.get_discharge(out=out_host)                   # D2H: resulting discharge map
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_cells * tps * Ke / e2e_s

    # ---- roofline of the dominant kernel (k_kw_diagonal) ----------------------------------------
    peak, peak_kind = measured_peaks()
    achieved = ALG_BYTES_ROUTING * wl.n * tps * K / (ms * 1e-3) / 1e9  # per rank
    roof = {"bound": "hbm", "kernel": "k_kw_diagonal", "achieved": round(achieved, 2), "peak": peak,
            "peak_kind": peak_kind, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
            "alg_bytes_per_cell_solve": ALG_BYTES_ROUTING, "launches": launches,
            "avg_launch_us": round(ms * 1e3 / max(launches, 1), 3),
            "note": "FP64 pow-bound Newton solve; HBM fraction is informational for this kernel (DESIGN.md §5)"}

    # ---- CPU baseline (oracle port) on a bounded sample -----------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_baseline(wl, timesteps=args.cpu_timesteps)
    if rank == 0:
        line = {"metric": "cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": wl.describe(kw.num_orders),
                "e2e": {"value": e2e_value, "unit": "cell-updates/s", "steps": Ke,
                        "h2d_bytes_per_step": int(wl.n * 8 + tps * 8), "d2h_bytes_per_step": int(wl.n * 8)},
                "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof, "cpu_baseline": cpu,
                "init_s": round(t_init, 3)}
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


def run_c3(args):
    """C3 = BASELINE.json configs[2]: 10000x10000, full soil + infiltration + overland + channel stack.  With N > 1 ranks
    the SAME raster (same seeds) is cut along its drainage graph over the N GPUs (strong scaling; `--workload c3-replicas`
    keeps the round-1 behaviour: one independent raster per rank, weak scaling)."""
    rank, world, local = dist_env()
    import torch
    from lisflood_code_b200 import _capi
    from lisflood_code_b200.synthetic_gpu import C3Device
    L = _capi.lib()
    numa_cores = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    _capi.check(L.lf_device_init(local))
    use_dist = world > 1
    cut = use_dist and args.workload == "c3"
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows, cols = args.rows, args.cols
    t0 = time.time()
    dev = C3Device(rows, cols, seed=300 + (0 if cut or not use_dist else rank), ldd_noise=args.ldd_noise, no_rout_steps=24,
                   distributed=cut, single_outlet=args.basin == "single")
    M = dev.model
    if os.environ.get("LF_EARLY_BPS"):      # tuning runs: resident blocks per SM of the early isolated-pixel launch
        M.set_option("early_blocks_per_sm", int(os.environ["LF_EARLY_BPS"]))
    M.set_option("overlap_isolated", 1 if args.overlap else 0)
    if os.environ.get("LF_NARROW") is not None:
        M.set_option("narrow_runs", int(os.environ["LF_NARROW"]))
    if os.environ.get("LF_ISO_BPS") is not None:     # tuning runs: footprint of the isolated-pixel kernel
        M.set_option("isolated_blocks_per_sm", int(os.environ["LF_ISO_BPS"]))
    _capi.synchronize()
    t_init = time.time() - t0
    info = M.info()
    n, nl, K, W = dev.n, dev.n_local, args.steps, args.warmup
    total_cells = n if (cut or not use_dist) else n * world

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        _capi.synchronize()

    def reduce_ranks(x, op="max"):
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident leg: the RAW meteo maps of the step (float32, as the NetCDF forcing holds them) are already in
    #      HBM, two alternating sets; a step = feeder modules (readmeteo scaling + snow + frost) + soil + overland + 24
    #      channel sub-steps.  LAI maps are 10-day maps in LISFLOOD (leafarea.py:76-90): set once, as in the e2e leg.  The
    #      synthetic initial state is cold, so the model is spun up for `--spinup` untimed steps first, like the
    #      reference's pre-run (the W warm-up steps follow).
    M.set_lai(dev.lai_device(), **({"local": True} if cut else {}))
    Fdev = [dev.forcing_device() for _ in range(2)]
    torch.cuda.synchronize()
    feed = (lambda F, day, asynchronous=False, **kw: M.feed(F, day, asynchronous=asynchronous, local=True, **kw)) if cut else \
        (lambda F, day, asynchronous=False, **kw: M.feed(F, day, asynchronous=asynchronous, **kw))
    day = [20]

    def one_step(F):
        day[0] = day[0] % 365 + 1
        feed(F, day[0])
        M.step()

    for w in range(args.spinup):
        one_step(Fdev[w % 2])
    for w in range(W):
        one_step(Fdev[w % 2])
    barrier()
    M.stage_times(reset=True)
    _capi.launch_count(reset=True)
    with ClockSampler(local) as clk:
        barrier()
        _capi.timer_start()
        for k in range(K):
            one_step(Fdev[k % 2])
        ms = _capi.timer_stop()
        barrier()
    host_calls = _capi.host_launch_count()
    launches = _capi.launch_count()
    st = M.stage_times(reset=True)
    M.soil_stats(enable_timing=True)      # one extra (untimed) step with per-kernel events in the soil stage
    one_step(Fdev[0])
    soil_stats = M.soil_stats(enable_timing=False)
    ms = reduce_ranks(ms)
    value = total_cells * K / (ms * 1e-3)
    nst = max(st["steps"], 1)
    soil_ms, of_ms, ch_ms = st["soil_ms"] / nst, st["overland_ms"] / nst, st["channel_ms"] / nst
    xstat = M.status() if cut else None
    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"profile_only": True, "value": value, "ms_per_step": ms / K, "gpu_launches": launches,
                              "stage_ms_per_step": {"soil_ms": soil_ms, "overland_ms": of_ms, "channel_ms": ch_ms},
                              "soil_stats": soil_stats, "exchange_aborted": xstat[0] if xstat else None}))
        if use_dist:
            dist.destroy_process_group()
        return

    # ---- end-to-end leg: per step the four RAW meteo maps (float32) come from pinned HOST memory -- copied on the copy
    #      stream while the previous step computes -- and the discharge map (ChanQAvg = dis) goes back to the host; the
    #      10-day LAI maps stay resident ----
    names = ("Precipitation", "Tavg", "ET0", "E0")
    host_sets = []
    for i in range(2):
        hs = {k: torch.empty(nl, dtype=torch.float32, pin_memory=True) for k in names}
        for k in hs:
            hs[k].copy_(Fdev[i][k])
        host_sets.append(hs)
    dis_host = [torch.empty(nl, dtype=torch.float64, pin_memory=True) for _ in range(2)]
    torch.cuda.synchronize()
    Ke = max(2, min(K, 10))
    Mdev = M.model if cut else M

    def e2e_loop(sets, outs, packings=None):
        """Ke steps fed from the host sets `sets`, discharge into the host buffers `outs`; seconds, max over ranks."""
        kw = (lambda i: {"packing": packings[i % 2]}) if packings else (lambda i: {})

        def e2e_step(i):
            M.step()                                  # forcing of step i was queued by the feed() of the previous call
            day[0] = day[0] % 365 + 1
            # next step's raw maps cross the host link while step i computes
            feed(sets[(i + 1) % 2], day[0], asynchronous=True, **kw(i + 1))
            Mdev.wait_outputs()                       # the discharge map of step i-1 has landed in its host buffer (the host
            #                                           would hand it to the writer thread here: global_modules/output.py)
            Mdev.get_async("ChanQAvg", outs[i % 2])   # D2H of this step's discharge map on the output stream

        feed(sets[0], day[0], asynchronous=True, **kw(0))
        e2e_step(0)
        barrier()
        t0 = time.perf_counter()
        for k in range(Ke):
            e2e_step(k + 1)
        Mdev.wait_outputs()                           # the last discharge map is on the host
        barrier()
        return reduce_ranks(time.perf_counter() - t0)

    e2e_s = e2e_loop(host_sets, dis_host)

    # the same loop with the reference's own narrower formats on the host link: forcing still packed the CF way (int16 +
    # scale_factor / add_offset, unpacked by the feeder kernel instead of the host reader, netcdf.py:231-232) and the
    # discharge map written as float32 (OutputMapsDataType = float32, netcdf.py:478): half the bytes each way
    packed_sets, packings = [], []
    for i in range(2):
        hs, pk = {}, {}
        for k in names:
            x = Fdev[i][k]
            lo, hi = float(x.min()), float(x.max())
            scale = (hi - lo) / 65534.0 if hi > lo else 1.0
            offset = (hi + lo) / 2
            h = torch.empty(nl, dtype=torch.int16, pin_memory=True)
            h.copy_(torch.round((x.double() - offset) / scale).to(torch.int16))
            hs[k], pk[k] = h, (scale, offset)
        packed_sets.append(hs)
        packings.append(pk)
    dis_host32 = [torch.empty(nl, dtype=torch.float32, pin_memory=True) for _ in range(2)]
    torch.cuda.synchronize()
    e2e_packed_s = e2e_loop(packed_sets, dis_host32, packings)
    torch.cuda.synchronize()                          # the last queued upload still borrows the host buffers
    del packed_sets, dis_host32
    e2e_value = total_cells * Ke / e2e_s
    h2d = nl * 4 * len(names)
    d2h = nl * 8

    # what the host link alone allows: the same bytes of one step copied with nothing else running, all ranks at once (GPUs
    # of one node share PCIe switches and host memory), first each direction alone, then both together as in the e2e loop
    scratch = {k: torch.empty(nl, dtype=torch.float32, device="cuda") for k in names}
    dsrc = torch.zeros(nl, dtype=torch.float64, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def copy_ms(do_in, do_out, reps=3):
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for r in range(reps):
            if do_in:
                with torch.cuda.stream(s_in):
                    for k in names:
                        scratch[k].copy_(host_sets[r % 2][k], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    dis_host[r % 2].copy_(dsrc, non_blocking=True)
        torch.cuda.synchronize()
        return reduce_ranks(time.perf_counter() - t0) * 1e3 / reps

    link = {"h2d_alone_ms": round(copy_ms(True, False), 3), "d2h_alone_ms": round(copy_ms(False, True), 3),
            "both_ms": round(copy_ms(True, True), 3)}
    link["h2d_gbs_per_gpu"] = round(h2d / link["h2d_alone_ms"] / 1e6, 1)
    link["d2h_gbs_per_gpu"] = round(d2h / link["d2h_alone_ms"] / 1e6, 1)
    link["note"] = ("one step's host traffic per GPU copied with the GPU otherwise idle, all %d ranks at the same time (max over "
                    "ranks): the floor the host link puts under an end-to-end step" % world)
    del scratch, dsrc

    # ---- rooflines ---------------------------------------------------------------------------------------
    peak, peak_kind = measured_peaks()
    soil_gbs = ALG_BYTES_SOIL * nl / (soil_ms * 1e-3) / 1e9
    chan_bytes = nl * 24 * ALG_BYTES_CHANNEL_SUBSTEP
    chan_gbs = chan_bytes / (ch_ms * 1e-3) / 1e9
    stage = {"soil_ms": round(soil_ms, 3), "overland_ms": round(of_ms, 3), "channel_ms": round(ch_ms, 3),
             "feeder_and_rest_ms": round(ms / K - soil_ms - of_ms - ch_ms, 3)}
    dominant = "soil_ms" if soil_ms >= ch_ms else "channel_ms"
    traffic = None   # measured DRAM bytes of the soil stage per step from the committed ncu capture (same raster size only)
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_soil_stage_c3_traffic.json")) as f:
            tj = json.load(f)
        if int(tj["cells"]) == int(nl):
            traffic = int(tj["soil_stage_bytes"])
    except (OSError, ValueError, KeyError):
        pass
    roof_soil = {"bound": "hbm", "kernel": "soil stage = k_soil_staged (TMA-staged first pass) + k_soil_veg_deferred + k_soil_pixel_flagged "
                 "(per-cell stencil: canopy+soil column+open/sealed+per-pixel sums+groundwater)", "achieved": round(soil_gbs, 1), "peak": peak, "peak_kind": peak_kind,
                 "unit": "GB/s", "frac": round(soil_gbs / peak, 4), "traffic": traffic, "alg_bytes_per_cell": ALG_BYTES_SOIL,
                 "avg_launch_ms": round(soil_ms, 3)}
    fp64 = fp64_roofline(nl, info, ch_ms, clk.summary())
    fp64_step = None      # FP64 warp instructions of the whole step (committed ncu counts, C3 size) against the pipe's peak
    try:
        with open(os.path.join(ROOT, "profiles", "r02_soil_stage_c3_traffic.json")) as f:
            tj = json.load(f)
        if int(tj["cells"]) == int(nl):
            inst = float(tj["fp64_warp_inst_per_step"]["total"])
            mhz = clk.summary().get("sm_mhz") or 1965.0
            pk = 2.0 * 148 * mhz * 1e6
            fp64_step = {"bound": "fp64", "achieved": round(inst / (ms / K * 1e-3) / 1e9, 1), "peak": round(pk / 1e9, 1),
                         "unit": "G warp-inst/s (FP64 pipe)", "frac": round(inst / (ms / K * 1e-3) / pk, 4),
                         "note": "whole step: the hot path is float64 Newton / van Genuchten arithmetic; at 100 %% of the FP64 "
                                 "pipe a step would take %.1f ms" % (inst / pk * 1e3)}
    except (OSError, ValueError, KeyError):
        pass
    roof_chan = {"bound": "hbm", "kernel": "k_chan_diagonal + k_chan_isolated_ws (24 fused channel sub-steps)",
                 "achieved": round(chan_gbs, 1), "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                 "frac": round(chan_gbs / peak, 4), "traffic": None,
                 "alg_bytes_per_cell_substep": ALG_BYTES_CHANNEL_SUBSTEP, "stage_ms": round(ch_ms, 3),
                 "note": "this stage is bound by the FP64 pipe, not by HBM: see roofline_fp64 (DESIGN.md section 4.5)",
                 "roofline_fp64": fp64}
    # headline roofline: the per-cell stencil (the kernel BASELINE.json's HBM target is stated on), HBM-bound.  The stage
    # that takes the most time (channel sub-steps) is bound by the FP64 pipe and reported against that pipe in
    # roofline_routing.roofline_fp64 / roofline_fp64_step -- an HBM fraction would be fictitious for it.
    roofline = dict(roof_soil)
    roofline["note"] = ("the time-dominant stage is %s (%.1f ms), FP64-pipe bound: see roofline_routing.roofline_fp64 and "
                        "roofline_fp64_step" % ("the channel sub-steps" if dominant == "channel_ms" else "this one", max(soil_ms, ch_ms)))
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_baseline_c3(args)
    if rank == 0:
        config = {"workload": "C3 synthetic %dx%d raster, full stack per step: feeder modules (meteo scaling + snow + frost) + "
                              "soil + infiltration + overland + 24 channel sub-steps, single kinematic routing, random D8 LDD "
                              "(noise/tilt %.2f)%s" % (rows, cols, args.ldd_noise,
                                                       "; ONE raster cut along its drainage graph over %d GPUs, boundary "
                                                       "discharges exchanged in-kernel over NVLink peer memory" % world
                                                       if cut else ("; one independent raster per GPU" if use_dist else "")),
                  "cells": total_cells, "cells_per_gpu": nl, "no_rout_steps": 24, "levels_overland": info["levels_overland"],
                  "levels_channel": info["levels_channel"], "channel_fraction": round(dev.channel_fraction, 4),
                  "isolated_channel_pixels": info["isolated_channel_pixels"], "device_bytes_maps": info["device_bytes"],
                  "l2_policy": "every map is %.0f MB per GPU (> 126 MB L2 for more than ~1.6e7 cells per GPU); two raw forcing "
                               "sets alternate between steps, the 10-day LAI maps stay resident" % (nl * 8 / 1e6),
                  "spinup_steps": args.spinup, "co_scheduled_isolated_pixels": bool(args.overlap),
                  "host_cores_bound_to_gpu_numa_node": numa_cores,
                  "drainage": "one basin (south-edge collector)" if args.basin == "single" else
                              "many catchments (steepest descent on tilted noise, every local sink is an outlet)"}
        if cut:
            summ = M.plan.summary()
            config.update({"cells_per_rank": M.loads, "trunk_pixels": M.n_trunk, "subtrees": M.n_roots,
                           "cut_edges": {g: summ[g]["cut_edges"] for g in summ},
                           "cut_edges_in_per_rank": {g: summ[g]["imports_per_rank"] for g in summ},
                           "bytes_exchanged_per_step": int(summ["overland"]["cut_edges"] * 3 * 8 +
                                                           summ["channel"]["cut_edges"] * 24 * 8),
                           "exchange_aborted": xstat[0]})
        line = {"metric": "cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
                "scaling": "strong" if cut else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "e2e": {"value": e2e_value, "unit": "cell-updates/s", "steps": Ke, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": round(e2e_s * 1e3 / Ke, 3), "host_link": link,
                        "note": "per step and GPU: raw Precipitation, Tavg, ET0, E0 (float32, as the reference's NetCDF forcing) from "
                                "pinned host memory through HotPathModel.feed (copy stream, overlapping the previous step), "
                                "discharge map ChanQAvg (float64) back to the host on the output stream (HotPathModel.get_async, "
                                "two host buffers; the timed region ends when the last map has landed); the 10-day LAI maps "
                                "stay resident"},
                "e2e_packed": {"value": total_cells * Ke / e2e_packed_s, "unit": "cell-updates/s", "steps": Ke,
                               "ms_per_step": round(e2e_packed_s * 1e3 / Ke, 3), "h2d_bytes_per_step": int(nl * 2 * len(names)),
                               "d2h_bytes_per_step": int(nl * 4),
                               "note": "the e2e loop with the narrower formats the reference itself reads and writes: forcing "
                                       "as CF-packed int16 (scale_factor / add_offset applied by the feeder kernel, "
                                       "HotPathModel.feed(packing=...)), discharge as float32 (OutputMapsDataType = float32, "
                                       "narrowed on the device); not the headline: `e2e` keeps float32 in / float64 out"},
                "gpu_launches": int(launches), "host_launch_calls": int(host_calls),
                "launch_note": "gpu_launches = kernels of this library executed in the timed region; host_launch_calls = launch "
                               "API calls the host made for them (the level sweep of the overland routers and the wavefront "
                               "diagonals are replayed as CUDA graphs: %.0f calls per step for %.0f kernels)"
                               % (host_calls / max(K, 1), launches / max(K, 1)),
                "clocks": clk.summary(), "roofline": roofline,
                "roofline_stencil": roof_soil, "roofline_routing": roof_chan, "roofline_fp64_step": fp64_step,
                "stage_ms_per_step": stage,
                "soil_stats": soil_stats,
                "cpu_baseline": cpu, "init_s": round(t_init, 2)}
    # multi-GPU default run: the same invocation also measures C4 (BASELINE.json configs[3]: 20000x20000 routing only, ONE
    # basin that has to be cut) and attaches its line, so the scaling record carries a workload with real cut edges too
    c4 = None
    if cut and not args.no_c4:
        del Fdev, host_sets, dis_host
        M.close()
        dev.model = None
        torch.cuda.empty_cache()
        c4 = run_c4(args, attached=True)
    if rank == 0:
        if c4 is not None:
            line["c4_cut"] = {k: c4[k] for k in ("value", "unit", "n_gpus", "ms_per_step", "scaling", "config", "gpu_launches")}
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


# FP64 instructions per pixel and channel sub-step of the isolated-pixel kernel, from the committed ncu capture
# (profiles/r02_chan_isolated_fp64.json: sm__inst_executed_pipe_fp64 / (pixels x sub-steps)); the FP64 pipe of sm_100
# issues 2 warp instructions per clock and SM (64 lanes: 37 TFLOP/s at 1.965 GHz).
def fp64_roofline(n_local, info, ch_ms, clocks):
    try:
        with open(os.path.join(ROOT, "profiles", "r02_chan_isolated_fp64.json")) as f:
            j = json.load(f)
        per = float(j["fp64_warp_inst_per_pixel_substep"])
    except (OSError, ValueError, KeyError):
        return None
    mhz = clocks.get("sm_mhz") or 1965.0
    sms = 148
    inst = per * info["isolated_channel_pixels"] * 24
    peak = 2.0 * sms * mhz * 1e6
    ach = inst / (ch_ms * 1e-3)
    return {"bound": "fp64", "kernel": "k_chan_isolated_ws", "achieved": round(ach / 1e9, 2), "peak": round(peak / 1e9, 2),
            "unit": "G warp-inst/s (FP64 pipe)", "frac": round(ach / peak, 4), "sm_mhz": mhz,
            "fp64_warp_inst_per_pixel_substep": per,
            "note": "lower bound of the pipe's utilisation: the wavefront kernels of the connected network run beside it"}


class _C3Crop(object):
    """The bench's own generator (synthetic_gpu.c3_generate: same code, seed and torch random streams as the device
    raster) at the crop size, collected on the host -- no library call -- with a cyclic list of raw forcing sets and the
    CPU restatement of the feeder modules in front of the model step."""

    def __init__(self, args, rows=None, nforcing=6):
        from lisflood_code_b200 import synthetic_gpu
        from oracle.lisf_oracle_feeders import FeederOracle, lai_term
        r = rows or args.cpu_rows
        self.S, (P, state), lai, self.raw = synthetic_gpu.c3_host_stack(r, r, seed=300, nforcing=nforcing,
                                                                       ldd_noise=args.ldd_noise, no_rout_steps=24)
        n = self.S["N"]
        Pm = {k: (np.full(n, v) if np.ndim(v) == 0 else v) for k, v in P.items() if k != "kgb"}
        self.feeder = FeederOracle(Pm, state, self.S["DtSec"])
        self.lai, self.laiterm = lai, lai_term(P["kgb"], lai)
        self.day = 20

    def forcing(self, S, i, seed=None):
        self.day = self.day % 365 + 1
        o = self.feeder.step(self.raw[i % len(self.raw)], self.day)
        return {"Rain": o["Rain"], "SnowMelt": o["SnowMelt"], "ETRef": o["ETRef"], "EWRef": o["EWRef"], "ESRef": o["ESRef"],
                "isFrozenSoil": o["isFrozenSoil"], "LAI": self.lai, "LAITerm": self.laiterm}


def _c3_oracle(args, rows=None):
    """CPU restatement of the same stack on a crop-sized raster of the SAME generator (C3 is CPU-timed on
    a crop and scaled per cell; tests/test_gpu_bench_configs.py checks the device path against it on these data)."""
    from oracle import lisf_oracle, lisf_oracle_model as om
    crop = _C3Crop(args, rows)
    lisf_oracle.set_threads(os.cpu_count() or 1)
    return crop.S, om.OracleModel(crop.S), crop, lisf_oracle


def cpu_baseline_c3(args):
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))     # the CPU leg may use every host core again
    except Exception:
        pass
    S, O, synthetic, lisf_oracle = _c3_oracle(args)
    cores = os.cpu_count() or 1
    best = None
    for thr in sorted({c for c in (16, 32, 64, cores) if c <= cores}):
        lisf_oracle.set_threads(thr)
        O.step(synthetic.forcing(S, 0, 300))
        t0 = time.perf_counter()
        O.step(synthetic.forcing(S, 1, 300))
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, thr)
    return {"value": S["N"] / best[0], "unit": "cell-updates/s", "cores": best[1], "kind": "port",
            "sample": "1 model step (24 sub-steps) on a %dx%d raster of the bench's own generator (synthetic_gpu.c3_generate, "
                      "random streams on %s); C/OpenMP kernels + NumPy glue exactly as the reference executes them; best of "
                      "the thread counts tried" % (args.cpu_rows, args.cpu_rows, S["rng_device"])}


def run_reference_c3(args):
    """Reference arm of the default workload: the reference's CPU algorithm (C/OpenMP kernels + NumPy glue, executed as
    the reference executes them; the Python/Numba reference itself cannot travel to the GPU box) on the host cores, at
    the best of the thread counts tried, each step a bounded sample (a crop of the same generator) of the C3 step."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    S, O, synthetic, lisf_oracle = _c3_oracle(args)
    cores = os.cpu_count() or 1
    best = None
    for i, thr in enumerate(sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})):   # untimed: also the warm-up
        lisf_oracle.set_threads(thr)
        t0 = time.perf_counter()
        O.step(synthetic.forcing(S, i, 300))
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, thr)
    nthr = lisf_oracle.set_threads(best[1])
    for w in range(args.warmup):
        O.step(synthetic.forcing(S, 5 + w, 300))
    t0 = time.perf_counter()
    for k in range(args.steps):
        O.step(synthetic.forcing(S, 10 + k, 300))
    dt = time.perf_counter() - t0
    value = S["N"] * args.steps / dt
    cpu = {"value": value, "unit": "cell-updates/s", "cores": nthr, "kind": "port",
           "sample": "%d model steps (24 sub-steps each) on a %dx%d raster of the bench's own generator "
                     "(synthetic_gpu.c3_generate, random streams on %s), per-cell throughput; best of the thread counts tried"
                     % (args.steps, args.cpu_rows, args.cpu_rows, S["rng_device"])}
    print(json.dumps({"impl": "reference", "metric": "cell-updates/s", "value": value, "unit": "cell-updates/s",
                      "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "C3 synthetic %dx%d raster, full stack per step: feeder modules (meteo scaling + snow + "
                                             "frost) + soil + infiltration + overland + 24 channel sub-steps, single kinematic "
                                             "routing, random D8 LDD (noise/tilt %.2f)" % (args.rows, args.cols, args.ldd_noise),
                                 "sample": "%dx%d crop of the same generator per step" % (args.cpu_rows, args.cpu_rows),
                                 "cells": S["N"], "no_rout_steps": 24},
                      "cpu_baseline": cpu,
                      "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_c4(args, attached=False):
    """BASELINE.json configs[3]: kinematic routing on ONE raster (default 20000x20000, 4e8 cells) cut along its drainage
    graph across the ranks (lisflood_code_b200/parallel.py): every rank generates the same network on its GPU, the device
    partitioner assigns the pixels, the routing kernels exchange boundary discharges over NVLink peer memory.  Strong
    scaling; with one rank it is the plain single-GPU router on the same raster."""
    rank, world, local = dist_env()
    import torch
    import torch.distributed as dist
    from lisflood_code_b200 import _capi
    from lisflood_code_b200.parallel import DistributedKinematicWave
    from lisflood_code_b200.synthetic_gpu import _ldd_gpu
    torch.cuda.set_device(local)
    _capi.check(_capi.lib().lf_device_init(local))
    own_group = not dist.is_initialized()
    if own_group:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    R = args.c4_rows
    n = R * R
    t0 = time.time()
    ldd = _ldd_gpu(torch, R, R, 400, args.c4_noise, single_outlet=True)   # same seed on every rank: ONE basin
    mask = torch.ones(n, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda")
    g.manual_seed(400 + 7919)
    U = lambda lo, hi: torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * (hi - lo) + lo
    alpha, q0, q = U(0.5, 3.0), U(0.1, 10.0), U(0.0, 1e-4)               # SURVEY.md 8d, C2 / C4 fields
    tps = args.c4_steps
    D = DistributedKinematicWave(ldd, mask, alpha, 0.6, 5000.0, 3600.0, max_steps=tps, rows=R, cols=R)
    del ldd, mask
    D.set_discharge(q0)
    D.set_lateral_inflow(q)
    del alpha, q0, q
    torch.cuda.empty_cache()
    t_init = time.time() - t0
    scales = np.random.default_rng(9).uniform(0.5, 1.5, (8, tps))
    K, W = (3, 2) if attached else (args.steps, max(args.warmup, 2))

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        _capi.synchronize()

    for w in range(W):
        D.run(tps, inflow_scale=scales[w % 8])
    barrier()
    _capi.launch_count(reset=True)
    with ClockSampler(local) as clk:
        barrier()
        t0 = time.perf_counter()
        _capi.timer_start()
        for k in range(K):
            D.run(tps, inflow_scale=scales[(W + k) % 8])
        ms = _capi.timer_stop()
        barrier()
        wall = time.perf_counter() - t0
    t = torch.tensor([ms, wall], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, wall = float(t[0].item()), float(t[1].item())
    launches = _capi.launch_count()
    aborted, runs = D.status()
    line = None
    if rank == 0:
        peak, peak_kind = measured_peaks()
        value = n * tps * K / (ms * 1e-3)
        ach = ALG_BYTES_ROUTING * max(D.loads) * tps * K / (ms * 1e-3) / 1e9
        line = {"metric": "cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C4 synthetic %dx%d raster, kinematic routing only, ONE basin cut along its drainage "
                                       "graph over %d GPU(s) (sub-trees bin-packed, trunk spread over the ranks, owner-to-owner cut "
                                       "edges), boundary discharges exchanged in-kernel over NVLink peer memory" % (R, R, world),
                           "cells": n, "timesteps_per_step": tps, "levels": D.kw.num_orders, "cells_per_rank": D.loads,
                           "cut_edges": D.cut_edges, "trunk_pixels": D.n_trunk, "subtrees": D.n_roots,
                           "bytes_exchanged_per_step": int(D.cut_edges * tps * 8), "exchange_aborted": aborted,
                           "l2_policy": "6 float64 maps of %.0f MB per rank, far above the 126 MB L2" % (max(D.loads) * 8 / 1e6)},
                "e2e": {"value": n * tps * K / wall, "unit": "cell-updates/s", "h2d_bytes_per_step": tps * 8,
                        "d2h_bytes_per_step": 0, "note": "host wall clock around the same K steps (per step the host sends "
                        "the inflow multipliers of its time steps; the state stays resident)"},
                "gpu_launches": int(launches), "clocks": clk.summary(),
                "roofline": {"bound": "hbm", "kernel": "k_kw_diagonal (busiest rank)", "achieved": round(ach, 1), "peak": peak,
                             "peak_kind": peak_kind, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": None,
                             "note": "FP64-bound Newton solve: the HBM fraction is informational (DESIGN.md)"},
                "cpu_baseline": None, "init_s": round(t_init, 2)}
        if not attached:
            print(json.dumps(line), flush=True)
    D.close()
    if own_group:
        dist.destroy_process_group()
    return line


def run_c5(args):
    """BASELINE.json configs[4] (EFAS-like): ~1000x950 raster with sea, 6-hourly steps, 6 routing sub-steps per step, split
    routing, reservoirs and lakes inside the sub-step loop.  A domain of ~5e5 cells is launch-latency bound: the step is
    ~100 small kernels (replayed as CUDA graphs).  N > 1: the same raster cut over the ranks (a structure stays with its
    feeders on one rank) -- reported for completeness; more GPUs cannot speed up a latency-bound step."""
    rank, world, local = dist_env()
    import torch
    from lisflood_code_b200 import _capi, synthetic
    from lisflood_code_b200.hotpath import HotPathModel
    torch.cuda.set_device(local)
    _capi.check(_capi.lib().lf_device_init(local))
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        from lisflood_code_b200.parallel import DistributedHotPathModel
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.time()
    S = synthetic.full_stack(1000, 950, seed=500, split_routing=True, ldd_noise=0.5, mask_fraction=0.45, channel_threshold=25,
                             dt_sec=21600.0)
    synthetic.add_structures(S, 40, 20, seed=500)
    n = S["N"]
    M = DistributedHotPathModel(S, subtree_fraction=0.05) if use_dist else HotPathModel(S)
    Mdev = M.model if use_dist else M
    nl = Mdev.N
    rng = np.random.default_rng(500)
    P = {"PrScaling": 1.0, "CalEvaporation": 1.0, "DeltaTSnow": 0.9674 * rng.uniform(0, 300.0, n) * 0.0065, "SnowSeason": 0.5,
         "TempSnow": 1.0, "SnowFactor": 1.0, "SnowMeltCoef": 4.0, "TempMelt": 0.0, "lat_rad": np.radians(rng.uniform(35, 70, n)),
         "Kfrost": 0.57, "Afrost": 0.97, "FrostIndexThreshold": 56.0, "SnowWaterEquivalent": 0.45, "kgb": 0.75 * 0.72}
    M.set_feeder(P, {"SnowCoverS": np.zeros((3, n)), "FrostIndex": np.zeros(n)})
    M.set_lai(rng.uniform(0, 6.0, (3, n)))
    info = Mdev.info()
    t_init = time.time() - t0
    pick = M._local if use_dist else (lambda a: a)
    raw_host = []
    for i in range(4):
        maps = {"Precipitation": (rng.gamma(0.8, 8.0, n) * (rng.random(n) < 0.45)).astype(np.float32),
                "Tavg": rng.uniform(-6, 18, n).astype(np.float32), "ET0": rng.uniform(0, 6, n).astype(np.float32),
                "E0": rng.uniform(0, 6, n).astype(np.float32)}
        hs = {}
        for k, a in maps.items():
            hs[k] = _capi.pinned_empty(nl, np.float32)
            hs[k][:] = pick(a)
        raw_host.append(hs)
    raw_dev = [{k: torch.as_tensor(np.asarray(a), device="cuda") for k, a in hs.items()} for hs in raw_host]
    dis = [_capi.pinned_empty(nl, np.float64) for _ in range(2)]
    K, W = args.steps, max(args.warmup, 3)

    def sync():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        _capi.synchronize()

    results = {}
    for graphs in (0, 1):
        Mdev.set_option("cuda_graphs", graphs)
        day = 40
        for w in range(W):
            Mdev.feed(raw_dev[w % 4], day + w)
            Mdev.step()
        sync()
        _capi.launch_count(reset=True)
        with ClockSampler(local) as clk:
            _capi.timer_start()
            for k in range(K):
                Mdev.feed(raw_dev[k % 4], day + k)
                Mdev.step()
            ms = _capi.timer_stop()
        sync()
        results[graphs] = (ms / K, _capi.launch_count() / K, clk.summary())
    # end to end: raw float32 forcing from page-locked host memory (asynchronous), discharge map back every step
    Mdev.feed(raw_host[0], 40, asynchronous=True)
    sync()
    t1 = time.perf_counter()
    for k in range(K):
        Mdev.step()
        Mdev.feed(raw_host[(k + 1) % 4], 41 + k, asynchronous=True)
        Mdev.wait_outputs()
        Mdev.get_async("ChanQAvg", dis[k % 2])
    Mdev.wait_outputs()
    sync()
    wall_ms = (time.perf_counter() - t1) / K * 1e3
    ms, launches, clocks = results[1]
    if use_dist:
        t = torch.tensor([ms, wall_ms, results[0][0]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, wall_ms, ms_nograph = (float(x) for x in t.tolist())
        aborted = M.status()[0]
    else:
        ms_nograph, aborted = results[0][0], None
    if rank == 0:
        cfg = {"workload": "C5 EFAS-like synthetic 1000x950 raster with 45 %% sea, 6-hourly steps, feeder modules + soil + overland + "
                           "6 routing sub-steps, split routing, 40 reservoirs + 20 lakes in the sub-step loop%s"
                           % ("; ONE raster cut over %d GPUs" % world if use_dist else ""),
               "cells": n, "cells_per_gpu": nl, "levels_overland": info["levels_overland"], "levels_channel": info["levels_channel"]}
        if use_dist:
            summ = M.plan.summary()
            cfg.update({"cells_per_rank": M.loads, "cut_edges": {g: summ[g]["cut_edges"] for g in summ}, "exchange_aborted": aborted})
        print(json.dumps({"metric": "cell-updates/s", "value": n / (ms * 1e-3), "unit": "cell-updates/s", "n_gpus": world, "steps": K,
                          "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if use_dist else "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                          "e2e": {"value": n / (wall_ms * 1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": int(nl * 16),
                                  "d2h_bytes_per_step": int(nl * 8), "ms_per_step": wall_ms,
                                  "note": "raw float32 forcing from page-locked host memory (asynchronous), discharge map back "
                                          "every step; host wall clock"},
                          "gpu_launches": int(launches * K), "kernels_per_step": launches, "clocks": clocks,
                          "without_cuda_graphs": {"ms_per_step": ms_nograph},
                          "one_year_6_hourly_s": round(1460 * wall_ms / 1e3, 1),
                          "roofline": None, "cpu_baseline": None, "init_s": round(t_init, 2)}), flush=True)
    if use_dist:
        M.close()
        dist.destroy_process_group()


def best_threads(ora, wl, lisf_oracle):
    """Thread count that maximises the oracle's throughput on this host (level-synchronous OpenMP
    does not always scale to every hardware thread); 2 timesteps per candidate."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, 128, cores // 2, cores) if 1 <= c <= cores})
    best, best_t = cores, None
    Q = wl.q0.copy()
    for c in cands:
        lisf_oracle.set_threads(c)
        ora.kinematicWaveRouting(Q, wl.q)
        t0 = time.perf_counter()
        ora.kinematicWaveRouting(Q, wl.q)
        ora.kinematicWaveRouting(Q, wl.q)
        t = time.perf_counter() - t0
        if best_t is None or t < best_t:
            best, best_t = c, t
    return lisf_oracle.set_threads(best)


def cpu_baseline(wl, timesteps=10, threads=None):
    from oracle import lisf_oracle
    ora = lisf_oracle.KinematicWaveOracle(wl.ldd[wl.mask], wl.mask, wl.alpha, wl.beta, wl.dx, wl.dt)
    nthr = lisf_oracle.set_threads(threads) if threads else best_threads(ora, wl, lisf_oracle)
    Q = wl.q0.copy()
    ora.kinematicWaveRouting(Q, wl.q)  # warm-up
    t0 = time.perf_counter()
    for s in range(timesteps):
        ora.kinematicWaveRouting(Q, wl.q * wl.scales[0, s % wl.timesteps_per_step])
    dt = time.perf_counter() - t0
    return {"value": wl.n * timesteps / dt, "unit": "cell-updates/s", "cores": nthr, "kind": "port",
            "sample": "%d kinematicWaveRouting timesteps of the same %dx%d raster (C oracle, OpenMP, %d threads)"
                      % (timesteps, wl.rows, wl.cols, nthr)}


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (C/OpenMP oracle port; the Python/Numba reference
    itself cannot travel to the GPU box) on all host threads, same workload/metric."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    wl = C2(0, args.ldd)
    from oracle import lisf_oracle
    ora = lisf_oracle.KinematicWaveOracle(wl.ldd[wl.mask], wl.mask, wl.alpha, wl.beta, wl.dx, wl.dt)
    nthr = best_threads(ora, wl, lisf_oracle)
    Q = wl.q0.copy()
    sample = args.cpu_timesteps  # timesteps per bench step (bounded sample of the 100-timestep step)
    for w in range(args.warmup):
        ora.kinematicWaveRouting(Q, wl.q)
    t0 = time.perf_counter()
    for k in range(args.steps):
        for s in range(sample):
            ora.kinematicWaveRouting(Q, wl.q * wl.scales[k % 64, s])
    dt = time.perf_counter() - t0
    value = wl.n * sample * args.steps / dt
    cpu = {"value": value, "unit": "cell-updates/s", "cores": nthr, "kind": "port",
           "sample": "%d of the %d timesteps of each step, full %dx%d raster" % (sample, wl.timesteps_per_step, wl.rows,
                                                                                 wl.cols)}
    print(json.dumps({"impl": "reference", "metric": "cell-updates/s", "value": value, "unit": "cell-updates/s",
                      "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": wl.describe(int(ora.order_start_stop.shape[0])),
                      "cpu_baseline": cpu,
                      "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0,
                              "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c3-replicas", "c2", "c4", "c5"])
    ap.add_argument("--c4-steps", type=int, default=96, help="routing time steps per bench step (one wavefront run)")
    ap.add_argument("--basin", default="many", choices=["many", "single"], help="C3 drainage network: many catchments "
                    "(default, the round-1 raster) or ONE basin that has to be cut between the GPUs")
    ap.add_argument("--no-c4", action="store_true", help="multi-GPU default run: skip the attached C4 leg")
    ap.add_argument("--overlap", type=int, default=0, help="1: co-schedule the isolated non-channel pixels with the soil stage")
    ap.add_argument("--c4-rows", type=int, default=20000)
    ap.add_argument("--c4-noise", type=float, default=0.2, help="noise/tilt of the C4 basin (0.2: a single catchment)")
    ap.add_argument("--rows", type=int, default=10000)
    ap.add_argument("--cols", type=int, default=10000)
    ap.add_argument("--ldd-noise", type=float, default=0.5)
    ap.add_argument("--spinup", type=int, default=10, help="untimed model steps before the warm-up (C3: relaxes the cold "
                    "synthetic initial state)")
    ap.add_argument("--cpu-rows", type=int, default=1000, help="edge of the crop the CPU baseline is timed on")
    ap.add_argument("--ldd", default="deep", choices=["deep", "shallow"])
    ap.add_argument("--cpu-timesteps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the e2e and CPU legs (for runs under ncu)")
    args = ap.parse_args()
    if args.workload == "c4" and args.impl != "reference":
        run_c4(args)
    elif args.workload == "c5" and args.impl != "reference":
        run_c5(args)
    elif args.workload in ("c3", "c3-replicas"):
        if args.impl == "reference":
            run_reference_c3(args)
        else:
            run_c3(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
