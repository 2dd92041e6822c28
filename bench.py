#!/usr/bin/env python
"""bench.py -- cell-updates/s of the LISFLOOD raster hot path on B200 (contract: see DESIGN.md §7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

One JSON line on stdout (rank 0).  `value` = device-resident throughput (inputs already in HBM),
`e2e` = the same work through the public plugin call with HOST buffers (H2D of the step's inflow map and
D2H of the resulting discharge map inside the timed region), `roofline` for the dominant kernel,
`cpu_baseline` = the CPU oracle (port of the reference algorithm) on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES_ROUTING = 44.0  # bytes per (cell, routing solve)


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
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


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
        print(json.dumps(line))
    if use_dist:
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
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--ldd", default="deep", choices=["deep", "shallow"])
    ap.add_argument("--cpu-timesteps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the e2e and CPU legs (for runs under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
