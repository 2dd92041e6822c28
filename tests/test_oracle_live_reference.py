"""Oracle vs the live reference (only where /root/reference exists, i.e. the build container)."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@pytest.mark.parametrize("rows,cols,noise,maskf,seed", [(120, 90, 3.0, 0.0, 1), (90, 130, 0.3, 0.2, 2), (200, 200, 0.1, 0.0, 3)])
def test_oracle_equals_reference(oracle, rows, cols, noise, maskf, seed):
    from lisflood_code_b200 import synthetic
    kwpt, kwp, sl = ref_loader.load()
    ldd, mask = synthetic.random_ldd(rows, cols, seed=seed, noise=noise, mask_fraction=maskf)
    n = int(mask.sum())
    alpha, q0, q = synthetic.routing_fields(n, seed)
    dx = np.random.default_rng(seed).uniform(3000, 7000, n)
    ref = kwp.kinematicWave(ldd[mask].copy(), mask, alpha, 0.6, dx, 3600.0)
    ora = oracle.KinematicWaveOracle(ldd[mask].copy(), mask, alpha, 0.6, dx, 3600.0)
    for k in ("downstream_lookup", "upstream_lookup", "num_upstream_pixels", "pixels_ordered", "order_start_stop"):
        assert np.array_equal(getattr(ref, k), getattr(ora, k)), k
    Qr, Qo = q0.copy(), q0.copy()
    for s in range(10):
        ref.kinematicWaveRouting(Qr, q)
        ora.kinematicWaveRouting(Qo, q)
    assert rel_err(Qo, Qr) < 1e-11


def test_pf_restatement_equals_the_reference_helpers():
    """saturationDegree / pressureHead of the live reference (soilloop.py:379-383, :428-432), pixel by pixel, against the
    NumPy restatement of the pF kernel: random columns plus dry, saturated, pore-less and beyond-HeadMax ones."""
    from oracle import lisf_oracle_model as om
    from test_oracle_soil_options_golden import edge_case_arguments
    kwpt, kwp, sl = ref_loader.load()
    rng = np.random.default_rng(3)
    V, N = 3, 400
    idx = np.array([2, 0, 1], np.int64)          # a permuted vegetation -> land use map
    a, _ = edge_case_arguments()
    cases = [a]
    wres, ws = rng.uniform(2, 30, (3, N)), rng.uniform(60, 400, (3, N))
    pore = rng.random((3, N)) > 0.05
    W = [np.ascontiguousarray(wres[idx] + rng.uniform(-0.05, 1.05, (V, N)) * (ws[idx] - wres[idx])) for _ in range(3)]
    gm = rng.uniform(0.08, 0.6, (3, N))
    cases.append([idx] + [np.empty((V, N)) for _ in range(3)] + W + [wres] * 3 + [ws] * 3 + [pore] * 3
                 + [1 / rng.uniform(0.004, 0.2, (3, N))] * 3 + [1 / gm] * 3 + [1 - gm] * 3 + [1.0e7])
    for a in cases:
        om.suction_unsaturated_soil_pf(*a)
        index, pf, W = a[0], a[1:4], a[4:7]
        wres, ws, pore, inva, invm, invn, headmax = a[7:10], a[10:13], a[13:16], a[16:19], a[19:22], a[22:25], a[25]
        for layer in range(3):
            want = np.empty_like(pf[layer])
            for v in range(want.shape[0]):
                lu = index[v]
                for p in range(want.shape[1]):
                    sat = sl.saturationDegree(W[layer][v, p], pore[layer][lu, p], wres[layer][lu, p], ws[layer][lu, p])
                    head = sl.pressureHead(sat, inva[layer][lu, p], invm[layer][lu, p], invn[layer][lu, p], headmax)
                    want[v, p] = np.log10(head) if head > 0 else -1.0
            assert rel_err(pf[layer], want) < 1e-13, layer


@pytest.mark.parametrize("user_grid", [True, False])
def test_misc_and_landuse_initial_equal_the_reference(user_grid):
    """miscInitial.initial() (grid size, unit multipliers, groundwater percolation / loss per step) and
    landusechange.initial() (fraction maps -> SoilFraction) of the live reference against the host mirrors
    (Lisflood_initial.InitialVariables.misc_initial / landuse_initial), attribute by attribute, bit for bit."""
    from lisflood_code_b200.Lisflood_initial import InitialVariables
    from oracle import ref_init
    rng = np.random.default_rng(12)
    mask = rng.random((9, 11)) > 0.2
    n = int(mask.sum())
    fr = rng.dirichlet([4, 3, 1, 0.6, 0.3, 0.2], n).T
    raw = {"DtSec": 21600.0, "DtSecChannel": 3600.0, "GwLoss": 0.05, "GwPercValue": rng.uniform(0.0, 1.5, n),
           "PrScaling": 1.0, "CalEvaporation": 1.0}
    raw.update({k + "Fraction": fr[i] for i, k in enumerate(("Other", "Forest", "Irrigation", "DirectRunoff", "Water", "Rice"))})
    if user_grid:
        raw.update(PixelLengthUser=rng.uniform(900.0, 1100.0, n), PixelAreaUser=rng.uniform(0.9e6, 1.1e6, n))
    else:
        raw.update(PixelLengthUser=5000.0)       # the mirror takes the cell size from this input in both modes
    want = ref_init.misc_and_landuse_initial(mask, raw, {"gridSizeUserDefined": user_grid}, cell=5000.0)
    var = InitialVariables(mask, raw, {"gridSizeUserDefined": user_grid}, DtSec=raw["DtSec"], DtSecChannel=raw["DtSecChannel"])
    var.misc_initial()
    var.landuse_initial()
    checked = 0
    for k in ("PixelLength", "PixelArea", "InvPixelLength", "DtSec", "DtDay", "InvDtSec", "InvDtDay", "DtSecChannel", "MMtoM",
              "MtoMM", "MMtoM3", "M3toMM", "GwLoss", "GwPerc", "GwPercStep", "GwLossStep", "ForestFraction",
              "DirectRunoffFraction", "WaterFraction", "IrrigationFraction", "RiceFraction", "OtherFraction", "SoilFraction"):
        got = np.asarray(getattr(var, k), np.float64)
        assert np.array_equal(np.broadcast_to(got, np.shape(want[k])), want[k]), k
        checked += 1
    assert checked == 23


def _shipped_tss():
    import glob
    import os
    if not ref_loader.available():
        return []
    root = os.path.normpath(os.path.join(ref_loader._R, "..", ".."))
    return sorted(glob.glob(os.path.join(root, "tests", "data", "*", "reference", "*", "*.tss")))


@pytest.mark.parametrize("path", _shipped_tss(), ids=lambda p: "/".join(p.split("/")[-3:]))
def test_tss_writer_reproduces_the_shipped_time_series(path, tmp_path):
    """TssWriter (global_modules/output.py) fed with the numbers of a .tss file the reference ships writes that file back
    byte for byte (header layout, column ids, ' %8g' / ' %14g' rows; zusatz.py:201-290) -- except the date stamp."""
    from lisflood_code_b200.global_modules.output import TssWriter
    text = open(path).read()
    lines = text.split("\n")
    head = lines[0]
    assert head.startswith("timeseries ") and " settingsfile: " in head and " date: " in head
    datatype = head.split(" ")[1]
    settings_path = head.split(" settingsfile: ")[1].split(" date: ")[0]
    ncols = int(lines[1])
    assert lines[2] == "timestep"
    ids = lines[3:3 + ncols - 1]
    rows = [ln for ln in lines[3 + ncols - 1:] if ln]
    out = tmp_path / "copy.tss"
    w = TssWriter(str(out), np.arange(ncols - 1), gauge_ids=ids, settings_path=settings_path, datatype=datatype)
    for ln in rows:
        f = ln.split()
        w.append(int(f[0]), np.array([float(x) for x in f[1:]]))
    w.close()
    mine = open(out).read().split("\n")
    assert mine[0].split(" date: ")[0] == head.split(" date: ")[0]
    assert mine[1:] == lines[1:]
    assert datatype == "valuescale.scalar"          # the writer's default


def _shipped_runs():
    import glob
    import os
    if not ref_loader.available():
        return []
    root = os.path.normpath(os.path.join(ref_loader._R, "..", ".."))
    return [d for d in sorted(glob.glob(os.path.join(root, "tests", "data", "LF_ETRS89_UseCase", "reference", "*")))
            if os.path.exists(os.path.join(d, "dis.nc")) and os.path.exists(os.path.join(d, "dis.tss"))]


@pytest.mark.parametrize("run", _shipped_runs(), ids=lambda p: p.split("/")[-1])
def test_shipped_map_stack_against_time_series_and_writer(run, tmp_path):
    """(1) The NetCDF-4 reader of the test harness (oracle/ref_maps.py) is exact: every number of the run's dis.tss is
    '%g' of the float32-rounded value of dis.nc at the gauge pixel (the reference samples gauges from a REAL4 PCRaster map).
    (2) MapStackWriter writes the reference's map stack: same dimensions, coordinate order, fill value, variable
    attributes, calendar and time axis (netcdf.py:432-583); fed with the maps of the shipped dis.nc it gives them back."""
    import datetime
    import os
    from scipy.io import netcdf_file
    from lisflood_code_b200.global_modules.output import MapStackWriter
    from oracle import ref_maps
    f = ref_maps.H5File(os.path.join(run, "dis.nc"))
    links = f.links()
    dis_info, time_info = f.dataset(links["dis"]), f.dataset(links["time"])
    dis, time = f.read(dis_info), f.read(time_info)
    maps = os.path.normpath(os.path.join(run, "..", "..", "maps"))
    outlets = ref_maps.read_netcdf4_2d(os.path.join(maps, "ec_outlets.nc"))
    mask = ref_maps.read_pcraster(os.path.join(maps, "mask.map")) == 1
    lines = open(os.path.join(run, "dis.tss")).read().split("\n")
    ncols = int(lines[1])
    ids = [int(x) for x in lines[3:3 + ncols - 1]]
    rows = [ln.split() for ln in lines[3 + ncols - 1:] if ln]
    assert len(rows) == dis.shape[0]
    table = np.array([[float(x) for x in r[1:]] for r in rows])
    for j, gauge in enumerate(ids):
        (r, c), = np.argwhere(outlets == gauge)
        assert np.array_equal(np.array([float("%g" % np.float32(v)) for v in dis[:, r, c]]), table[:, j]), gauge
    # ---- the writer against this file
    a = dis_info["attrs"]
    assert float(a["_FillValue"]) == -9999.0 and (dis[:, ~mask] == -9999.0).all() and (dis[:, mask] != -9999.0).all()
    units = time_info["attrs"]["units"]
    kind, stamp = units.split(" since ")
    start = datetime.datetime.strptime(stamp, "%Y-%m-%d %H:%M:%S.0")
    dt_sec = {"days": 86400.0, "hours": 3600.0}[kind] * float(time[1] - time[0])
    x, y = f.read(f.dataset(links["x"])), f.read(f.dataset(links["y"]))
    assert (np.diff(y) < 0).all() and (np.diff(x) > 0).all()            # y descending, x ascending
    path = str(tmp_path / "dis_copy.nc")
    w = MapStackWriter(path, "dis", mask, dt_sec, start, a["standard_name"], a["long_name"], a["units"], x=x, y=y,
                       calendar=time_info["attrs"]["calendar"])
    nsteps = min(dis.shape[0], 12)
    for k in range(nsteps):           # step k + 1 after the start date of the time axis carries time[k]
        w.append(int(round(time[k] * {"days": 86400.0, "hours": 3600.0}[kind] / dt_sec)) + 1, dis[k][mask])
    w.close()
    nc = netcdf_file(path, "r", mmap=False)
    v, t = nc.variables["dis"], nc.variables["time"]
    assert t.units.decode() == units and t.calendar.decode() == time_info["attrs"]["calendar"]
    assert t.standard_name.decode() == time_info["attrs"]["standard_name"]
    assert np.array_equal(t[:], time[:nsteps])
    assert v.dimensions == ("time", "y", "x") and v.shape[1:] == dis.shape[1:]
    assert np.array_equal(v[:], dis[:nsteps]) and float(v._FillValue) == -9999.0
    for k in ("standard_name", "long_name", "units"):
        assert getattr(v, k).decode() == a[k], k
    assert np.array_equal(nc.variables["x"][:], x) and np.array_equal(nc.variables["y"][:], y)
    nc.close()


def test_input_files_keys_equal_the_reference_classes():
    """Every HydroModule mirror declares the inputs (binding names per option) its reference class declares; the dynamic
    wave (out of scope) is the only option left out."""
    from lisflood_code_b200.hydrological_modules import (groundwater, lakes, opensealed, reservoir, routing, soil, soilloop,
                                                         surface_routing)
    from oracle import ref_modules
    M = ref_modules.load()
    for name, mirror in (("soil", soil.soil), ("routing", routing.routing), ("groundwater", groundwater.groundwater),
                         ("surface_routing", surface_routing.surface_routing), ("soilloop", soilloop.soilloop),
                         ("opensealed", opensealed.opensealed), ("reservoir", reservoir.reservoir), ("lakes", lakes.lakes)):
        want = {k: sorted(v) for k, v in M[name].input_files_keys.items() if k != "dynamicWave"}
        assert {k: sorted(v) for k, v in mirror.input_files_keys.items()} == want, name
        assert mirror.module_name == M[name].module_name, name


def test_end_maps_table_equals_the_reference_report_table():
    """state_io.END_MAPS against the reference's own table of reported maps (global_modules/default_options.py): every end
    map of the hot-path state is an entry reported with option repEndMaps, from the same model attribute (and row), with the
    same option gating; and no end map of a hot-path state variable is missing."""
    from lisflood_code_b200 import state_io
    from oracle import ref_settings
    ref_settings.load()
    import sys
    table = sys.modules["lisflood_ref_settings.global_modules.default_options"].default_options["reportedmaps"]
    for name, (binding, attr, row) in state_io.END_MAPS.items():
        entry = table[name]
        assert entry.end == ["repEndMaps"], name
        assert entry.output_var == (attr if row is None else "%s[%d]" % (attr, row)), (name, entry.output_var)
        assert ("SplitRouting" in entry.restrictoption) == (name in state_io.SPLIT_ONLY), name
    mine = {(attr if row is None else "%s[%d]" % (attr, row)) for _, attr, row in state_io.END_MAPS.values()}
    hot_path_state = ("Theta", "UZ", "LZ", "DSLR", "CumInter", "OFM3", "TotalCrossSectionArea", "ChanQ", "CrossSection2Area",
                      "Sideflow1Chan", "SnowCoverS", "FrostIndex")
    others = {e.output_var for e in table.values() if e.end == ["repEndMaps"] and e.output_var.startswith(hot_path_state)}
    assert others <= mine, others - mine
