"""Output / forcing IO around the device-resident step, overlapped with it (reference:
global_modules/output.py:68-167, netcdf.py:170-341,432-583, zusatz.py:196-405).

  * MapStackWriter   a CF-1.6 map stack (time, y, x) like the reference's dis.nc: same dimensions, attributes, fill value
                     and time axis (netcdf.py:432-583).  The reference writes NETCDF4/HDF5 through the netCDF4 package;
                     neither it nor libhdf5 exists in this image, so the container is NetCDF-3 (64-bit offset, the
                     record dimension `time` grows step by step), written with scipy.io.netcdf_file -- readable by
                     netCDF4 / xarray like any other NetCDF file.
  * TssWriter        a PCRaster timeoutput time series (dis.tss) in the reference's exact text layout (zusatz.py:201-297).
  * OutputPipeline   per step: lf_model_get_async into one of two page-locked host buffers, then a writer thread
                     decompresses the map to the raster and appends it to the files while the device computes the next
                     step (the D2H copy runs on its own stream).
  * ForcingStack     a float32 or CF-packed int16 (time, y, x) NetCDF-3 forcing stack read through mmap;
                     ForcingPrefetcher compresses the maps of the coming step into page-locked buffers on a thread and
                     hands them to HotPathModel.feed (asynchronous H2D on the copy stream; packed maps cross the host link
                     as int16 and are unpacked by the feeder kernel).
"""
import datetime
import os
import queue
import threading
import time as xtime

import numpy as np

FILL = -9999.0


def time_units(dt_sec, start_date):
    """netcdf.py:556-566."""
    stamp = start_date.strftime("%Y-%m-%d %H:%M:%S.0")
    if dt_sec >= 86400:
        return "days since " + stamp, 86400.0
    if dt_sec >= 3600:
        return "hours since " + stamp, 3600.0
    return "minutes since " + stamp, 60.0


class MapStackWriter(object):
    def __init__(self, path, var_name, land_mask, dt_sec, start_date, standard_name="", long_name="", units="",
                 x=None, y=None, dtype="f8", settings_path="", calendar="proleptic_gregorian"):
        from scipy.io import netcdf_file
        self.mask = np.asarray(land_mask, bool)
        rows, cols = self.mask.shape
        self.nc = nc = netcdf_file(path, "w", version=2)
        nc.settingsfile = os.path.realpath(settings_path) if settings_path else ""
        nc.date_created = xtime.ctime(xtime.time())
        nc.Source_Software = "lisflood_code_b200 (B200 hot path of Lisflood OS)"
        nc.source = "Lisflood output maps"
        nc.keywords = "Lisflood, EFAS, GLOFAS"
        nc.Conventions = "CF-1.6"
        nc.createDimension("time", None)
        nc.createDimension("y", rows)
        nc.createDimension("x", cols)
        t = nc.createVariable("time", "f8", ("time",))
        t.standard_name = "time"
        t.calendar = calendar
        t.units, self.unit_sec = time_units(dt_sec, start_date)
        yv, xv = nc.createVariable("y", "f8", ("y",)), nc.createVariable("x", "f8", ("x",))
        yv[:] = np.arange(rows, 0, -1, dtype=np.float64) if y is None else y      # y descending, x ascending
        xv[:] = np.arange(cols, dtype=np.float64) if x is None else x
        v = nc.createVariable(var_name, dtype, ("time", "y", "x"))
        v._FillValue = np.array(FILL, dtype)
        v.standard_name, v.long_name, v.units = standard_name, long_name, units
        self.var, self.time, self.dt_sec, self.k = v, t, float(dt_sec), 0
        self.raster = np.full((rows, cols), FILL, v.data.dtype if hasattr(v, "data") else np.float64)

    def append(self, step, values):
        """values: float64[N] compressed; step: 1-based model step (time stamp = start + (step - 1) * dt, netcdf.py:536)."""
        self.raster[self.mask] = values                    # decompress (add1.py:287-305): -9999 outside the mask
        self.var[self.k] = self.raster
        self.time[self.k] = (step - 1) * self.dt_sec / self.unit_sec
        self.k += 1

    def close(self):
        self.nc.close()


class TssWriter(object):
    """zusatz.py:201-297: header (spatial datatype, settings file, date; number of columns; 'timestep'; the gauge ids),
    then one row per step: ' %8g' for the step and ' %14g' per gauge.  The reference samples the gauges from a PCRaster
    scalar map (REAL4: pcraster.cellvalue, zusatz.py:340-378), so the printed numbers are the float32-rounded values --
    all 11 190 numbers of the shipped init_daily/dis.tss equal '%g' of the float32-rounded dis.nc values, and 29 of 30
    gauges differ in a last digit somewhere without the rounding; it is applied here too."""

    def __init__(self, path, gauge_pixels, gauge_ids=None, settings_path="", datatype="valuescale.scalar", header=True):
        # datatype: str(pcraster data type).lower() of the gauge map's values; "valuescale.scalar" is what the .tss files
        # shipped with the reference carry (tests/data/*/reference/*/dis.tss)
        self.pix = np.asarray(gauge_pixels, np.int64)
        ids = np.arange(1, self.pix.size + 1) if gauge_ids is None else np.asarray(gauge_ids)
        self.f = open(path, "w")
        if header:
            self.f.write("timeseries {} settingsfile: {} date: {}\n".format(datatype, settings_path, xtime.ctime(xtime.time())))
            self.f.write(str(self.pix.size + 1) + "\n")
            self.f.write("timestep\n")
            for i in ids:
                self.f.write(str(i) + "\n")

    def append(self, step, values):
        row = " %8g" % step
        for v in np.asarray(values)[self.pix]:
            row += " %14g" % np.float32(v)
        self.f.write(row + "\n")

    def close(self):
        self.f.close()


def pinned(n, dtype=np.float64):
    """A page-locked host array (cudaHostAlloc through the library; freed with its last view)."""
    from .. import _capi
    return _capi.pinned_empty(n, dtype)


class OutputPipeline(object):
    """report(step) after every model step: the map leaves the device asynchronously and is written by a thread."""

    def __init__(self, model, name, writers, pin=True, dtype=np.float64):
        """dtype: np.float64, or np.float32 for OutputMapsDataType = float32 (netcdf.py:478) -- the map is then narrowed
        on the device and half the bytes cross the host link."""
        self.M, self.name, self.writers = model, name, list(writers)
        self.bufs = [pinned(model.N, dtype) if pin else np.empty(model.N, dtype) for _ in range(2)]
        self.turn, self.pending = 0, None
        self.q = queue.Queue(maxsize=2)
        self.err = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            step, buf, done = item
            try:
                for w in self.writers:
                    w.append(step, buf)
            except Exception as e:       # surfaced by the next report() / close()
                self.err = e
            done.set()

    def _flush_pending(self):
        if self.pending is not None:
            step, buf, done = self.pending
            self.M.wait_outputs()                      # the D2H copy of that step has landed
            self.q.put((step, buf, done))
            self.pending = None

    def report(self, step):
        if self.err:
            raise self.err
        self._flush_pending()                          # hand the previous step's map to the writer thread
        buf = self.bufs[self.turn]
        done = getattr(self, "_done%d" % self.turn, None)
        if done is not None:
            done.wait()                                # the writer is finished with this buffer (two steps ago)
        done = threading.Event()
        setattr(self, "_done%d" % self.turn, done)
        self.M.get_async(self.name, buf)               # queued behind the step's kernels; returns at once
        self.pending = (step, buf, done)
        self.turn ^= 1

    def close(self):
        self._flush_pending()
        self.q.put(None)
        self.thread.join()
        for w in self.writers:
            w.close()
        if self.err:
            raise self.err


class ForcingStack(object):
    """A (time, y, x) float32 stack in a NetCDF-3 file (what readmeteo's xarray reader delivers per step,
    readmeteo.py:61-69, netcdf.py:170-341), memory-mapped; [k] gives the compressed float32 map of step k."""

    def __init__(self, path, var_name, land_mask, packed=False):
        """packed=True: an int16 variable with scale_factor / add_offset is handed on as stored (read_into fills an int16
        buffer) and unpacked by the feeder kernel (HotPathModel.feed(packing=...)); otherwise it is unpacked here, on the
        host, as the reference's reader does (netcdf.py:231-232)."""
        from scipy.io import netcdf_file
        self.nc = netcdf_file(path, "r", mmap=True, maskandscale=False)
        self.var = self.nc.variables[var_name]
        self.mask = np.asarray(land_mask, bool)
        v = self.var
        self.scale = float(np.asarray(getattr(v, "scale_factor", 1.0)).reshape(-1)[0])
        self.offset = float(np.asarray(getattr(v, "add_offset", 0.0)).reshape(-1)[0])
        self.packed = bool(packed)
        if self.packed and v.data.dtype.newbyteorder("=") != np.dtype(np.int16):
            kind = str(v.data.dtype)
            del v
            self.close()
            raise TypeError("ForcingStack(packed=True): variable '%s' is %s, not int16" % (var_name, kind))

    def __len__(self):
        return self.var.shape[0]

    @property
    def packing(self):
        """(scale_factor, add_offset) of the variable."""
        return self.scale, self.offset

    def read_into(self, k, out):
        a = self.var[k][self.mask]
        if not self.packed and (self.scale != 1.0 or self.offset != 0.0):
            a = a * self.scale + self.offset
        out[:] = a
        return out

    def close(self):
        self.var = None
        self.nc.close()


def write_forcing_stack(path, var_name, land_mask, maps, dt_sec=86400.0, start_date=None, packing=None):
    """Writes compressed float32 maps [(N,), ...] as a (time, y, x) NetCDF-3 stack (test / example data); with
    packing = (scale_factor, add_offset) the maps are int16 already packed with these attributes."""
    from scipy.io import netcdf_file
    mask = np.asarray(land_mask, bool)
    rows, cols = mask.shape
    nc = netcdf_file(path, "w", version=2)
    nc.Conventions = "CF-1.6"
    nc.createDimension("time", None)
    nc.createDimension("y", rows)
    nc.createDimension("x", cols)
    t = nc.createVariable("time", "f8", ("time",))
    t.units = time_units(dt_sec, start_date or datetime.datetime(2000, 1, 1))[0]
    if packing is None:
        v = nc.createVariable(var_name, "f4", ("time", "y", "x"))
        v._FillValue = np.float32(FILL)
        ras = np.full((rows, cols), FILL, np.float32)
    else:
        v = nc.createVariable(var_name, "i2", ("time", "y", "x"))
        v._FillValue = np.int16(-32768)
        v.scale_factor = np.array([packing[0]], np.float64)      # arrays: a Python float would be stored as NC_FLOAT
        v.add_offset = np.array([packing[1]], np.float64)
        ras = np.full((rows, cols), -32768, np.int16)
    for k, m in enumerate(maps):
        ras[mask] = m
        v[k] = ras
        t[k] = k
    nc.close()


class ForcingPrefetcher(object):
    """Reads the four raw meteo maps of the coming steps from their stacks into page-locked float32 buffers on a thread
    while the current step runs; next() returns (step index, maps).  Four buffer sets rotate and at most two finished
    sets wait in the queue, so the set handed out by next() is not refilled before next() has been called twice more:
    by then the asynchronous upload HotPathModel.feed(..., asynchronous=True) started from it has long finished (the
    caller synchronises with the device once per step, e.g. through OutputPipeline.report / wait_outputs)."""
    NAMES = ("Precipitation", "Tavg", "ET0", "E0")

    def __init__(self, stacks, n, first=0, last=None, pin=True):
        self.stacks = stacks
        self.last = min(len(stacks[k]) for k in self.NAMES) if last is None else last
        packed = [bool(getattr(stacks[k], "packed", False)) for k in self.NAMES]
        if any(packed) and not all(packed):
            raise ValueError("ForcingPrefetcher: either all four stacks are handed on packed or none")
        # packed stacks: int16 buffers and the attributes feed() needs -- M.feed(maps, day, packing=prefetcher.packing)
        self.packing = {k: stacks[k].packing for k in self.NAMES} if all(packed) else None
        dt = np.int16 if self.packing else np.float32
        self.sets = [{k: (pinned(n, dt) if pin else np.empty(n, dt)) for k in self.NAMES} for _ in range(4)]
        self.q = queue.Queue(maxsize=1)
        self.thread = threading.Thread(target=self._run, args=(first,), daemon=True)
        self.thread.start()

    def _run(self, first):
        for i, k in enumerate(range(first, self.last)):
            s = self.sets[i % 4]
            for name in self.NAMES:
                self.stacks[name].read_into(k, s[name])
            self.q.put((k, s))               # blocks while a finished set is waiting
        self.q.put(None)

    def next(self):
        item = self.q.get()
        if item is None:
            raise StopIteration
        return item
