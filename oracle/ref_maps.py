"""Minimal readers for the maps of the reference's own test catchment (tests/data/LF_ETRS89_UseCase/maps):
NetCDF-4 files as the reference ships them (HDF5 with version-2 object headers, one 2-D variable stored CONTIGUOUSLY and
uncompressed) and PCRaster CSF maps.  TEST INFRASTRUCTURE ONLY (build container; neither netCDF4 / h5py nor PCRaster
exist in this image): used by tests/golden/make_golden.py to turn the catchment into committed golden vectors.
Format references: HDF5 File Format Specification 3.0 (object header v2: IV.A.1.b; messages 0x01 dataspace, 0x03 datatype,
0x08 layout, 0x10 continuation), PCRaster CSF version 2 header (256 bytes, cell representation at byte 66, rows / columns
at 100 / 104)."""
import re
import struct

import numpy as np


def _messages(b, start, end, creation_order):
    p = start
    while p + 4 <= end:
        kind = b[p]
        size = struct.unpack_from("<H", b, p + 1)[0]
        p += 4 + (2 if creation_order else 0)
        if p + size > end:
            return
        data = b[p:p + size]
        p += size
        if kind == 0x10:                                   # continuation: more messages in an OCHK block
            off, length = struct.unpack_from("<QQ", data, 0)
            if b[off:off + 4] == b"OCHK":
                yield from _messages(b, off + 4, off + length - 4, creation_order)
        else:
            yield kind, data


def read_netcdf4_2d(path):
    """The (only) 2-D variable of a NetCDF-4 map file, as stored (int8 / float32 / float64 ...)."""
    b = open(path, "rb").read()
    if b[:8] != b"\x89HDF\r\n\x1a\n":
        raise ValueError("%s is not an HDF5 file" % path)
    for m in re.finditer(rb"OHDR", b):
        o = m.start()
        if b[o + 4] != 2:
            continue
        flags = b[o + 5]
        p = o + 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
        width = 1 << (flags & 3)
        size = int.from_bytes(b[p:p + width], "little")
        p += width
        info = {}
        for kind, d in _messages(b, p, p + size, bool(flags & 0x04)):
            if kind == 0x01:                                # dataspace: version, rank, flags, dimensions
                info["dims"] = struct.unpack_from("<%dQ" % d[1], d, 4 if d[0] == 2 else 8)
            elif kind == 0x03:                              # datatype: class, bit field, size
                info["class"], info["bits"], info["size"] = d[0] & 15, d[1], struct.unpack_from("<I", d, 4)[0]
            elif kind == 0x08 and d[0] == 3 and d[1] == 1:  # layout version 3, class 1 = contiguous: address, size
                info["addr"], info["nbytes"] = struct.unpack_from("<QQ", d, 2)
        if len(info.get("dims", ())) == 2 and "addr" in info and "class" in info:
            if info["class"] == 1:
                dt = {4: "<f4", 8: "<f8"}[info["size"]]
            elif info["class"] == 0:
                dt = ("<i" if info["bits"] & 8 else "<u") + str(info["size"])
            else:
                continue
            rows, cols = info["dims"]
            if info["nbytes"] != rows * cols * np.dtype(dt).itemsize:
                raise ValueError("%s: unexpected storage size" % path)
            return np.frombuffer(b, dt, count=rows * cols, offset=info["addr"]).reshape(rows, cols).copy()
    raise ValueError("%s: no contiguous 2-D variable found" % path)


def read_pcraster(path):
    """The cells of a PCRaster CSF map (row-major), as stored."""
    b = open(path, "rb").read()
    if b[:27] != b"RUU CROSS SYSTEM MAP FORMAT":
        raise ValueError("%s is not a PCRaster map" % path)
    cell_repr = struct.unpack_from("<H", b, 66)[0]
    rows, cols = struct.unpack_from("<II", b, 100)
    dt = {0x00: "u1", 0x04: "i1", 0x11: "<u2", 0x15: "<i2", 0x22: "<u4", 0x26: "<i4", 0x5a: "<f4", 0xdb: "<f8"}[cell_repr]
    return np.frombuffer(b, dt, count=rows * cols, offset=256).reshape(rows, cols).copy()
