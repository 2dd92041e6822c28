#!/usr/bin/env python
"""lisf1.py settings.xml [-q -v -l -c -h -t -d -n -i -s]

Entry surface of LISFLOOD (reference: src/lisf1.py, src/lisflood/main.py:164-226) for the B200 hot path.  The
settings file has the reference's structure (<lfoptions>, <lfuser>, <lfbinding>, $(var) substitution).  Because
PCRaster / NetCDF IO is out of scope here, map bindings name NumPy files (.npy: a 2-D raster or an already compressed
1-D array) or numbers, as `loadmap` accepts them:

    MaskMap        .npy  bool[rows, cols]
    every static input of the hot-path modules by its reference binding name (Ldd, Channels, ChanGrad, MapKSat1,
    SoilDepth1, b_Xinanjiang, ForestFraction ... -- the `input_files_keys` of the module mirrors), DtSec, DtSecChannel:
                   the modules' initial() derive the parameter maps as the reference does (Lisflood_initial.py);
    or StateFile   .npz  the derived maps / initial state by reference attribute name (see synthetic.full_stack)
    ForcingFile    .npz  Rain, SnowMelt, ETRef ... stacked as (steps, ...)     DisOut  .npy  written: ChanQAvg (= dis)
    StepStart, StepEnd   1-based step numbers                                          per step
"""
import sys

import numpy as np


def model_state(settings):
    """The dictionary HotPathModel takes: from the StateFile binding if there is one, else derived from the raw
    static inputs by the modules' initial() (lisflood_code_b200/Lisflood_initial.py::initialise)."""
    b = settings.binding
    mask = np.load(b["MaskMap"]).astype(bool)
    split = bool(settings.options["SplitRouting"]) and not settings.options["InitLisflood"]
    if "StateFile" in b:
        S = {k: (v.item() if v.ndim == 0 else v) for k, v in np.load(b["StateFile"], allow_pickle=False).items()}
    else:
        from lisflood_code_b200.Lisflood_initial import initialise
        var = initialise(mask, {}, settings.options, DtSec=float(b["DtSec"]), DtSecChannel=float(b["DtSecChannel"]))
        S = var.state()
    S["mask"] = mask
    S["SplitRouting"] = split
    for k in ("N", "rows", "cols", "NoRoutSteps"):
        if k in S:
            S[k] = int(S[k])
    return S


def main(*args):
    argv = list(args) if args else sys.argv[1:]
    if not argv:
        print(__doc__)
        return 1
    from lisflood_code_b200.global_modules.settings import LisSettings
    from lisflood_code_b200.hotpath import HotPathModel
    from lisflood_code_b200.Lisflood_dynamic import LisfloodModel_dyn
    settings = LisSettings(argv[0], argv[1:])
    settings.check_supported()
    b, flags = settings.binding, settings.flags
    S = model_state(settings)
    if flags["initonly"]:
        return 0
    var = HotPathModel(S)
    if flags["nancheck"]:
        var.set_option("flagnancheck", 1)      # device-side watch of the channel discharge, warns once (-n)
    model = LisfloodModel_dyn(var)
    with np.load(b["ForcingFile"]) as z:       # every array is read (and decompressed) once, not once per step
        forcing = {k: z[k] for k in z.files}
    first, last = int(b.get("StepStart", 1)), int(b.get("StepEnd", forcing["Rain"].shape[0]))
    dis = []
    for step in range(first, last + 1):
        F = {k: a[step - 1] for k, a in forcing.items()}
        model.dynamic(F)
        dis.append(var.get("ChanQAvg"))
        if flags["loud"]:
            print("%-6i %10.2f" % (step, float(dis[-1].max())))
        elif not (flags["quiet"] or flags["veryquiet"]):
            sys.stdout.write("\r%d" % step)
            sys.stdout.flush()
    np.save(b["DisOut"], np.stack(dis))
    return 0


if __name__ == "__main__":
    sys.exit(main())
