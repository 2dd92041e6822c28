"""Times the soil stage of the C3 workload for each tile / occupancy variant of k_soil_staged (LF_SOIL_VARIANT).

  python tools/soil_variants.py [--rows 10000 --cols 10000 --steps 5]

The model is built once on the device (synthetic_gpu.C3Device); each variant runs `steps` model steps after two
warm-up steps and prints the stage timers and the per-kernel soil events.  A checksum of the state after a fixed
number of steps is printed per variant: all variants must agree (same arithmetic, different launch shape).
"""
import argparse
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10000)
    ap.add_argument("--cols", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--variants", default="13,11,10,12,13")
    ap.add_argument("--spinup", type=int, default=10, help="model steps before the first variant (the deferred fraction "
                    "settles from 3.3 % to ~1.2 % over the first ~10 steps; repeat a variant at both ends to bracket drift)")
    args = ap.parse_args()
    import torch
    from lisflood_code_b200 import _capi
    from lisflood_code_b200.synthetic_gpu import C3Device
    torch.cuda.set_device(0)
    _capi.check(_capi.lib().lf_device_init(0))
    dev = C3Device(args.rows, args.cols, seed=300, ldd_noise=0.5, no_rout_steps=24)
    M = dev.model
    F = [dev.forcing_device(i) for i in range(2)]
    torch.cuda.synchronize()
    for w in range(args.spinup):
        M.step(F[w % 2])
    for v in [int(x) for x in args.variants.split(",")]:
        os.environ["LF_SOIL_VARIANT"] = str(v)
        for w in range(2):
            M.step(F[w % 2])
        _capi.synchronize()
        M.stage_times(reset=True)
        for k in range(args.steps):
            M.step(F[k % 2])
        _capi.synchronize()
        st = M.stage_times(reset=True)
        M.soil_stats(enable_timing=True)
        M.step(F[0])
        ss = M.soil_stats(enable_timing=False)
        n = max(st["steps"], 1)
        print(json.dumps({"variant": v, "soil_ms": round(st["soil_ms"] / n, 3), "overland_ms": round(st["overland_ms"] / n, 3),
                          "channel_ms": round(st["channel_ms"] / n, 3), "kernel_ms": ss["kernel_ms"],
                          "deferred_fraction": round(ss["deferred_fraction"], 5)}), flush=True)


if __name__ == "__main__":
    main()
