"""Overlapped output / forcing IO on the device: asynchronous D2H of the discharge map + writer thread
and prefetched float32 forcing through HotPathModel.feed give the files a synchronous run gives."""
import datetime

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_output_pipeline_equals_synchronous_run(gpu_lib, tmp_path):
    from scipy.io import netcdf_file
    from lisflood_code_b200 import synthetic
    from lisflood_code_b200.global_modules.output import (ForcingPrefetcher, ForcingStack, MapStackWriter, OutputPipeline,
                                                          TssWriter, write_forcing_stack)
    from lisflood_code_b200.hotpath import HotPathModel
    S = synthetic.full_stack(70, 90, seed=21, mask_fraction=0.1)
    n, mask = S["N"], S["mask"]
    rng = np.random.default_rng(3)
    steps = 6
    raw = {"Precipitation": [(rng.gamma(0.8, 8.0, n) * (rng.random(n) < 0.5)).astype(np.float32) for _ in range(steps)],
           "Tavg": [rng.uniform(-8, 20, n).astype(np.float32) for _ in range(steps)],
           "ET0": [rng.uniform(0, 6, n).astype(np.float32) for _ in range(steps)],
           "E0": [rng.uniform(0, 6, n).astype(np.float32) for _ in range(steps)]}
    P = {"PrScaling": 1.0, "CalEvaporation": 1.0, "DeltaTSnow": rng.uniform(0, 2, n), "SnowSeason": 0.5, "TempSnow": 1.0,
         "SnowFactor": 1.0, "SnowMeltCoef": 4.0, "TempMelt": 0.0, "lat_rad": np.full(n, 0.8), "Kfrost": 0.57, "Afrost": 0.97,
         "FrostIndexThreshold": 56.0, "SnowWaterEquivalent": 0.45, "kgb": 0.75 * 0.72}
    lai = rng.uniform(0, 6, (3, n))

    def model():
        M = HotPathModel(S)
        M.set_feeder(P, {"SnowCoverS": np.zeros((3, n)), "FrostIndex": np.zeros(n)})
        M.set_lai(lai)
        return M
    # synchronous reference run
    A = model()
    want = []
    for k in range(steps):
        A.feed({name: raw[name][k] for name in raw}, 40 + k)
        A.step()
        want.append(A.get("ChanQAvg"))
    # overlapped run: forcing from NetCDF stacks through the prefetcher, discharge through the output pipeline
    stacks = {}
    for name, maps in raw.items():
        write_forcing_stack(str(tmp_path / (name + ".nc")), name, mask, maps)
        stacks[name] = ForcingStack(str(tmp_path / (name + ".nc")), name, mask)
    B = model()
    gauges = np.argsort(-want[-1])[:5]
    out = OutputPipeline(B, "ChanQAvg", [MapStackWriter(str(tmp_path / "dis.nc"), "dis", mask, S["DtSec"],
                                                        datetime.datetime(2016, 1, 1), "discharge", "discharge", "m3/s"),
                                         TssWriter(str(tmp_path / "dis.tss"), gauges)])
    pf = ForcingPrefetcher(stacks, n)
    for k in range(steps):
        i, maps = pf.next()
        assert i == k
        B.feed(maps, 40 + k, asynchronous=True)
        B.step()
        out.report(k + 1)
    out.close()
    nc = netcdf_file(str(tmp_path / "dis.nc"), "r", mmap=False)
    data = nc.variables["dis"][:]
    nc.close()
    assert data.shape[0] == steps
    for k in range(steps):
        assert np.array_equal(data[k][mask], want[k]), k
    rows = open(tmp_path / "dis.tss").read().splitlines()[4 + gauges.size - 1:]
    assert rows[-1] == " %8g" % steps + "".join(" %14g" % v for v in want[-1][gauges])
