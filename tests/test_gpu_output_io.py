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
    assert rows[-1] == " %8g" % steps + "".join(" %14g" % np.float32(v) for v in want[-1][gauges])   # REAL4 sampling


def test_packed_forcing_and_float32_output(gpu_lib, tmp_path):
    """The narrow formats end to end: CF-packed int16 forcing stacks through the prefetcher (unpacked by the feeder
    kernel) and a float32 map stack (OutputMapsDataType = float32) equal a synchronous run fed with the unpacked maps."""
    from scipy.io import netcdf_file
    from lisflood_code_b200 import synthetic
    from lisflood_code_b200.global_modules.output import (ForcingPrefetcher, ForcingStack, MapStackWriter, OutputPipeline,
                                                          write_forcing_stack)
    from lisflood_code_b200.hotpath import HotPathModel
    from oracle.lisf_oracle_feeders import cf_pack, cf_unpack
    S = synthetic.full_stack(60, 80, seed=23, mask_fraction=0.1)
    n, mask = S["N"], S["mask"]
    rng = np.random.default_rng(6)
    steps = 5
    fields = {"Precipitation": rng.gamma(0.8, 8.0, (steps, n)) * (rng.random((steps, n)) < 0.5),
              "Tavg": rng.uniform(-8, 20, (steps, n)), "ET0": rng.uniform(0, 6, (steps, n)), "E0": rng.uniform(0, 6, (steps, n))}
    P = {"PrScaling": 1.0, "CalEvaporation": 1.0, "DeltaTSnow": rng.uniform(0, 2, n), "SnowSeason": 0.5, "TempSnow": 1.0,
         "SnowFactor": 1.0, "SnowMeltCoef": 4.0, "TempMelt": 0.0, "lat_rad": np.full(n, 0.8), "Kfrost": 0.57, "Afrost": 0.97,
         "FrostIndexThreshold": 56.0, "SnowWaterEquivalent": 0.45, "kgb": 0.75 * 0.72}
    lai = rng.uniform(0, 6, (3, n))

    def model():
        M = HotPathModel(S)
        M.set_feeder(P, {"SnowCoverS": np.zeros((3, n)), "FrostIndex": np.zeros(n)})
        M.set_lai(lai)
        return M
    stacks, unpacked = {}, {}
    for name, v in fields.items():
        raw, s, o = cf_pack(v)
        path = str(tmp_path / (name + ".nc"))
        write_forcing_stack(path, name, mask, list(raw), packing=(s, o))
        stacks[name] = ForcingStack(path, name, mask, packed=True)
        unpacked[name] = cf_unpack(raw, s, o, "float32")
    A = model()
    want = []
    for k in range(steps):
        A.feed({name: unpacked[name][k] for name in unpacked}, 40 + k)
        A.step()
        want.append(A.get("ChanQAvg"))
    B = model()
    out = OutputPipeline(B, "ChanQAvg", [MapStackWriter(str(tmp_path / "dis32.nc"), "dis", mask, S["DtSec"],
                                                        datetime.datetime(2016, 1, 1), "discharge", "discharge", "m3/s",
                                                        dtype="f4")], dtype=np.float32)
    pf = ForcingPrefetcher(stacks, n)
    for k in range(steps):
        i, maps = pf.next()
        B.feed(maps, 40 + k, asynchronous=True, packing=pf.packing, decode="float32")
        B.step()
        out.report(k + 1)
    out.close()
    nc = netcdf_file(str(tmp_path / "dis32.nc"), "r", mmap=False)
    data = nc.variables["dis"][:]
    nc.close()
    assert data.dtype.newbyteorder("=") == np.dtype(np.float32) and data.shape[0] == steps
    for k in range(steps):
        assert np.array_equal(data[k][mask], want[k].astype(np.float32)), k
    assert float(want[-1].max()) > 0
