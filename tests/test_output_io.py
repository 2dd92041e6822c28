"""Output / forcing IO on the CPU: the map stack is a CF NetCDF file with the reference's layout
(netcdf.py:432-583), the time series has the reference's text layout (zusatz.py:201-297), the forcing prefetcher delivers
the maps of the stack in order.  The overlap with the device is exercised in tests/test_gpu_output_io.py."""
import datetime

import numpy as np
import pytest


def test_map_stack_and_tss(tmp_path):
    from scipy.io import netcdf_file
    from lisflood_code_b200.global_modules.output import FILL, MapStackWriter, TssWriter
    rng = np.random.default_rng(1)
    mask = rng.random((7, 9)) > 0.3
    n = int(mask.sum())
    w = MapStackWriter(str(tmp_path / "dis.nc"), "dis", mask, 21600.0, datetime.datetime(2016, 1, 1, 6), "discharge",
                       "discharge", "m3/s")
    t = TssWriter(str(tmp_path / "dis.tss"), [0, 5, n - 1], gauge_ids=[11, 12, 13], settings_path="settings.xml")
    maps = [rng.uniform(0, 1e4, n) for _ in range(4)]
    for k, m in enumerate(maps):
        w.append(3 + k, m)
        t.append(3 + k, m)
    w.close()
    t.close()
    nc = netcdf_file(str(tmp_path / "dis.nc"), "r", mmap=False)
    assert nc.Conventions == b"CF-1.6" and nc.variables["dis"].dimensions == ("time", "y", "x")
    assert nc.variables["time"].units == b"hours since 2016-01-01 06:00:00.0"
    assert np.array_equal(nc.variables["time"][:], [12.0, 18.0, 24.0, 30.0])           # (step - 1) * 6 h
    data = nc.variables["dis"][:]
    assert data.shape == (4, 7, 9) and float(nc.variables["dis"]._FillValue) == FILL
    for k, m in enumerate(maps):
        assert np.array_equal(data[k][mask], m) and np.all(data[k][~mask] == FILL)
    nc.close()
    lines = open(tmp_path / "dis.tss").read().splitlines()
    assert lines[0].startswith("timeseries valuescale.scalar settingsfile: settings.xml date: ")
    assert lines[1:6] == ["4", "timestep", "11", "12", "13"]
    want = " %8g" % 3 + "".join(" %14g" % np.float32(v) for v in maps[0][[0, 5, n - 1]])   # REAL4 sampling of the reference
    assert lines[6] == want and len(lines) == 6 + 4


def test_forcing_stack_prefetch(tmp_path):
    from lisflood_code_b200.global_modules.output import ForcingPrefetcher, ForcingStack, write_forcing_stack
    rng = np.random.default_rng(2)
    mask = rng.random((6, 8)) > 0.25
    n = int(mask.sum())
    data = {k: [rng.uniform(-5, 30, n).astype(np.float32) for _ in range(5)] for k in ForcingPrefetcher.NAMES}
    stacks = {}
    for k, maps in data.items():
        write_forcing_stack(str(tmp_path / (k + ".nc")), k, mask, maps)
        stacks[k] = ForcingStack(str(tmp_path / (k + ".nc")), k, mask)
    pf = ForcingPrefetcher(stacks, n, pin=False)
    seen = []
    while True:
        try:
            k, s = pf.next()
        except StopIteration:
            break
        seen.append(k)
        for name in ForcingPrefetcher.NAMES:
            assert s[name].dtype == np.float32 and np.array_equal(s[name], data[name][k])
    assert seen == [0, 1, 2, 3, 4]
    for st in stacks.values():
        st.close()


def test_packed_forcing_stack(tmp_path):
    """An int16 stack with scale_factor / add_offset: unpacked on the host like the reference's reader (default), or handed
    on as stored together with its attributes (packed=True) for the feeder kernel to unpack."""
    from lisflood_code_b200.global_modules.output import ForcingPrefetcher, ForcingStack, write_forcing_stack
    from oracle.lisf_oracle_feeders import cf_pack, cf_unpack
    rng = np.random.default_rng(4)
    mask = rng.random((7, 9)) > 0.25
    n = int(mask.sum())
    packed, attrs, stacks, hosted = {}, {}, {}, {}
    for name in ForcingPrefetcher.NAMES:
        allmaps = rng.uniform(-5, 30, (4, n))
        raw, s, o = cf_pack(allmaps)
        packed[name], attrs[name] = list(raw), (s, o)
        path = str(tmp_path / (name + ".nc"))
        write_forcing_stack(path, name, mask, packed[name], packing=(s, o))
        stacks[name] = ForcingStack(path, name, mask, packed=True)
        hosted[name] = ForcingStack(path, name, mask)
        assert stacks[name].packing == (s, o)
    pf = ForcingPrefetcher(stacks, n, pin=False)
    assert pf.packing == attrs
    for k in range(4):
        i, maps = pf.next()
        assert i == k
        for name in ForcingPrefetcher.NAMES:
            assert maps[name].dtype == np.int16 and np.array_equal(maps[name], packed[name][k])
            got = hosted[name].read_into(k, np.empty(n, np.float32))
            assert np.array_equal(got, cf_unpack(packed[name][k], *attrs[name], decode="float64").astype(np.float32))
    with pytest.raises(StopIteration):
        pf.next()
    with pytest.raises(ValueError):
        ForcingPrefetcher(dict(stacks, Tavg=hosted["Tavg"]), n, pin=False)
    for st in list(stacks.values()) + list(hosted.values()):
        st.close()
    data = [rng.uniform(0, 1, n).astype(np.float32)]
    write_forcing_stack(str(tmp_path / "f.nc"), "x", mask, data)
    with pytest.raises(TypeError):
        ForcingStack(str(tmp_path / "f.nc"), "x", mask, packed=True)
