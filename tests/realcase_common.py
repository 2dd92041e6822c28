"""The reference's test catchment as a fixture (tests/golden/realcase_*.npz, made by make_golden.py::real_usecase_case from
the reference's real maps, meteo stacks and SHIPPED output files): helpers shared by the CPU and the GPU test."""
import numpy as np

from conftest import load_golden

THETA = {"tha": ("Theta1a", 0), "thfa": ("Theta1a", 1), "thia": ("Theta1a", 2), "thc": ("Theta2", 0), "thfc": ("Theta2", 1),
         "thic": ("Theta2", 2)}
RAW = ("Precipitation", "Tavg", "ET0", "E0")


def load(case):
    g = load_golden(case)
    S = {k[3:]: (v.item() if v.ndim == 0 else v) for k, v in g.items() if k.startswith("S__")}
    for k in ("N", "rows", "cols", "NoRoutSteps"):
        S[k] = int(S[k])
    S["SplitRouting"] = bool(S["SplitRouting"])
    P = {k[3:]: (float(v) if v.ndim == 0 else v) for k, v in g.items() if k.startswith("P__")}
    state = {"SnowCoverS": g["Z__SnowCoverS"], "FrostIndex": g["Z__FrostIndex"]}
    steps = int(g["steps"])
    raw = [{k: g["R%d__%s" % (t, k)] for k in RAW} for t in range(steps)]
    days = [int(g["day%d" % t]) for t in range(steps)]
    lai = [g["L%d" % t] for t in range(steps)]
    want = [{k.split("__", 1)[1]: v for k, v in g.items() if k.startswith("O%d__" % t)} for t in range(steps)]
    shipped = [{k.split("__", 1)[1]: v for k, v in g.items() if k.startswith("X%d__" % t)} for t in range(steps)]
    return S, P, state, raw, days, lai, want, shipped
