"""prints the pixels of the adversarial golden where the device solver deviates most from the reference"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from lisflood_code_b200 import _capi
from lisflood_code_b200.hydrological_modules.kinematic_wave_parallel import kinematicWave
_capi.check(_capi.lib().lf_device_init(0))
for case in ("kwadv_24x160_beta06", "kwadv_24x160_beta07"):
    g = dict(np.load(os.path.join("tests", "golden", case + ".npz")))
    kw = kinematicWave(g["ldd"], g["mask"], g["alpha"], float(g["beta"]), g["dx"], float(g["dt"]))
    a = g["alpha"] * g["dx"] / float(g["dt"])
    Q = g["q0"].copy()
    ups = kw.upstream_lookup
    for s in range(g["Q_main"].shape[0]):
        Qold = Q.copy()
        kw.kinematicWaveRouting(Q, g["q"])
        ref = g["Q_main"][s]
        e = np.abs(Q - ref) / np.maximum(np.abs(ref), 1e-12)
        for p in np.argsort(-e)[:4]:
            U = sum(ref[u] for u in ups[p] if u >= 0)
            print(case, "step", s, "pix", p, "err %.2e" % e[p], "ref %.17g got %.17g" % (ref[p], Q[p]), "a %.6g Qold %.6g U %.6g lat %.6g" % (
                a[p], Qold[p], U, g["q"][p] * g["dx"][p]), flush=True)
