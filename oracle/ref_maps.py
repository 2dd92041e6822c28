"""Minimal readers for the files of the reference's own test catchments (tests/data/*/maps, tests/data/*/reference):
NetCDF-4 files as the reference ships and writes them (HDF5: superblock version 0, version-2 object headers with compact
links, contiguous or chunked (version-1 B-tree) storage, shuffle + deflate filters, fixed and variable-length string
attributes) and PCRaster CSF maps.  TEST INFRASTRUCTURE ONLY (build container; neither netCDF4 / h5py nor PCRaster exist
in this image): used by tests/golden/make_golden.py to turn the catchment into committed golden vectors and by the live
tests that compare the product's writers with the reference's shipped outputs.  Checked against the reference's own
outputs: every number of the shipped dis.tss equals the float32-rounded value read here from dis.nc at its gauge.
Format references: HDF5 File Format Specification 3.0 (object header v2: IV.A.1.b; messages 0x01 dataspace, 0x03 datatype,
0x06 link, 0x08 layout, 0x0B filter pipeline, 0x0C attribute, 0x10 continuation; B-tree v1 chunk index III.A.1; global
heap III.E), PCRaster CSF version 2 header (256 bytes, cell representation at byte 66, rows / columns at 100 / 104)."""
import struct
import zlib

import numpy as np


class H5File(object):
    """links() -> {variable name: address}; dataset(address) -> dims, dtype, attrs, ...; read(info) -> array."""
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n"
        ver = b[8]
        if ver in (0, 1):
            # v0: sizes at 13,14; root symbol table entry at end of superblock
            self.so, self.sl = b[13], b[14]
            p = 24 + (4 if ver == 1 else 0)
            base, free, eof, drv = struct.unpack_from("<QQQQ", b, p)
            p += 32
            # root group symbol table entry: link name offset(8), object header address(8)
            self.root = struct.unpack_from("<Q", b, p + 8)[0]
        else:
            self.so, self.sl = b[9], b[10]
            self.root = struct.unpack_from("<Q", b, 12 + 24)[0]
    def messages(self, addr):
        b = self.b
        if b[addr:addr+4] == b"OHDR":
            flags = b[addr+5]; p = addr + 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
            w = 1 << (flags & 3); size = int.from_bytes(b[p:p+w], "little"); p += w
            yield from self._msgs2(p, p + size, bool(flags & 4))
        else:
            # version 1 object header
            ver = b[addr]; nmsg = struct.unpack_from("<H", b, addr + 2)[0]; size = struct.unpack_from("<I", b, addr + 8)[0]
            yield from self._msgs1(addr + 16, addr + 16 + size)
    def _msgs2(self, p, end, corder):
        b = self.b
        while p + 4 <= end:
            t = b[p]; sz = struct.unpack_from("<H", b, p+1)[0]; p += 4 + (2 if corder else 0)
            if p + sz > end: return
            d = b[p:p+sz]; p += sz
            if t == 0x10:
                off, ln = struct.unpack_from("<QQ", d, 0)
                if b[off:off+4] == b"OCHK":
                    yield from self._msgs2(off + 4, off + ln - 4, corder)
            else:
                yield t, d
    def _msgs1(self, p, end):
        b = self.b
        while p + 8 <= end:
            t, sz = struct.unpack_from("<HH", b, p); p += 8
            d = b[p:p+sz]; p += sz
            if t == 0x10:
                off, ln = struct.unpack_from("<QQ", d, 0)
                yield from self._msgs1(off, off + ln)
            else:
                yield t, d
    def links(self, addr=None):
        out = {}
        for t, d in self.messages(self.root if addr is None else addr):
            if t == 6:
                flags = d[1]; p = 2
                ltype = 0
                if flags & 8: ltype = d[p]; p += 1
                if flags & 4: p += 8
                if flags & 16: p += 1
                w = 1 << (flags & 3); n = int.from_bytes(d[p:p+w], "little"); p += w
                name = d[p:p+n].decode(); p += n
                if ltype == 0:
                    out[name] = struct.unpack_from("<Q", d, p)[0]
        return out
    def _dtype(self, d):
        cls = d[0] & 15; size = struct.unpack_from("<I", d, 4)[0]
        if cls == 0: return np.dtype(("<i" if d[1] & 8 else "<u") + str(size)), 8 + 4
        if cls == 1: return np.dtype("<f%d" % size), 8 + 12
        if cls == 3: return ("str", size), 8
        if cls == 9:
            base, n = self._dtype(d[8:])
            return ("vlen", base, d[1] & 15), 8 + n
        return ("other", cls, size), 8
    def _dataspace(self, d):
        ver, rank, fl = d[0], d[1], d[2]
        off = 4 if ver == 2 else 8
        return struct.unpack_from("<%dQ" % rank, d, off), off + 8 * rank * (2 if fl & 1 else 1)
    def _gheap(self, addr, index):
        b = self.b
        assert b[addr:addr+4] == b"GCOL"
        size = struct.unpack_from("<Q", b, addr + 8)[0]
        p = addr + 16
        while p < addr + size:
            idx, ref = struct.unpack_from("<HH", b, p); osz = struct.unpack_from("<Q", b, p + 8)[0]
            if idx == index: return b[p+16:p+16+osz]
            if idx == 0: break
            p += 16 + (osz + 7) // 8 * 8
        raise KeyError(index)
    def attribute(self, d):
        ver = d[0]
        if ver == 1:
            ns, ts, ss = struct.unpack_from("<HHH", d, 2); p = 8
            name = d[p:p+ns].split(b"\0")[0].decode(); p += (ns + 7) // 8 * 8
            dt, _ = self._dtype(d[p:p+ts]); p += (ts + 7) // 8 * 8
            dims, _ = self._dataspace(d[p:p+ss]) if ss else ((), 0); p += (ss + 7) // 8 * 8
        else:
            ns, ts, ss = struct.unpack_from("<HHH", d, 2); p = 8 + (1 if ver == 3 else 0)
            name = d[p:p+ns].split(b"\0")[0].decode(); p += ns
            dt, _ = self._dtype(d[p:p+ts]); p += ts
            dims, _ = self._dataspace(d[p:p+ss]) if ss else ((), 0); p += ss
        n = int(np.prod(dims)) if dims else 1
        raw = d[p:]
        if isinstance(dt, np.dtype):
            v = np.frombuffer(raw, dt, count=n)
            return name, (v[0] if n == 1 else v.copy())
        if dt[0] == "str":
            return name, raw[:dt[1]].split(b"\0")[0].decode(errors="replace")
        if dt[0] == "vlen":
            vals = []
            for i in range(n):
                ln, ga, gi = struct.unpack_from("<IQI", raw, 16 * i)
                data = self._gheap(ga, gi)[:ln * (dt[1].itemsize if isinstance(dt[1], np.dtype) else 1)]
                vals.append(data.decode(errors="replace") if not isinstance(dt[1], np.dtype) or dt[2] == 1 else np.frombuffer(data, dt[1]))
            return name, (vals[0] if n == 1 else vals)
        return name, None
    def dataset(self, addr):
        info = {"attrs": {}}
        for t, d in self.messages(addr):
            if t == 1: info["dims"] = self._dataspace(d)[0]
            elif t == 3: info["dtype"] = self._dtype(d)[0]
            elif t == 8: info["layout"] = d
            elif t == 0xb: info["filters"] = d
            elif t == 0xc:
                try:
                    k, v = self.attribute(d); info["attrs"][k] = v
                except Exception as e:
                    pass
        return info
    def read(self, info):
        b = self.b; d = info["layout"]; dt = info["dtype"]; dims = info["dims"]
        assert d[0] == 3
        if d[1] == 1:
            addr, nbytes = struct.unpack_from("<QQ", d, 2)
            return np.frombuffer(b, dt, count=int(np.prod(dims)), offset=addr).reshape(dims).copy()
        assert d[1] == 2
        nd = d[2]; btree = struct.unpack_from("<Q", d, 3)[0]
        cdims = struct.unpack_from("<%dI" % nd, d, 11)          # last = element size
        chunk = cdims[:-1]
        filt = self._filters(info.get("filters"))
        out = np.zeros(dims, dt)
        for offs, addr, size, mask in self._chunks(btree, nd):
            raw = b[addr:addr+size]
            for fid in reversed(filt):
                if fid == 1: raw = zlib.decompress(raw)
                elif fid == 2:
                    n = len(raw) // dt.itemsize
                    raw = np.frombuffer(raw, np.uint8).reshape(dt.itemsize, n).T.tobytes()
            a = np.frombuffer(raw, dt, count=int(np.prod(chunk))).reshape(chunk)
            sl = tuple(slice(o, min(o + c, m)) for o, c, m in zip(offs, chunk, dims))
            out[sl] = a[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out
    def _filters(self, d):
        if d is None: return []
        ver, n = d[0], d[1]; p = 8 if ver == 1 else 2; ids = []
        for i in range(n):
            fid, = struct.unpack_from("<H", d, p); p += 2
            if ver == 1 or fid >= 256:
                nl, = struct.unpack_from("<H", d, p); p += 2
            else: nl = 0
            fl, ncv = struct.unpack_from("<HH", d, p); p += 4
            if ver == 1: nl = (nl + 7) // 8 * 8
            p += nl + 4 * ncv
            if ver == 1 and ncv % 2: p += 4
            ids.append(fid)
        return ids
    def _chunks(self, addr, nd):
        b = self.b
        assert b[addr:addr+4] == b"TREE", b[addr:addr+4]
        ntype, level, used = b[addr+4], b[addr+5], struct.unpack_from("<H", b, addr + 6)[0]
        p = addr + 8 + 16
        ksz = 8 + 8 * nd
        for i in range(used):
            size, mask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from("<%dQ" % nd, b, p + 8)[:-1]
            child = struct.unpack_from("<Q", b, p + ksz)[0]
            if level == 0: yield offs, child, size, mask
            else: yield from self._chunks(child, nd)
            p += ksz + 8


def read_netcdf4_2d(path):
    """The (only) 2-D variable of a NetCDF-4 map file, as stored (int8 / float32 / float64 ...)."""
    f = H5File(path)
    for name, addr in f.links().items():
        info = f.dataset(addr)
        if len(info.get("dims", ())) == 2 and isinstance(info.get("dtype"), np.dtype):
            return f.read(info)
    raise ValueError("%s: no 2-D variable found" % path)


def read_pcraster(path):
    """The cells of a PCRaster CSF map (row-major), as stored."""
    b = open(path, "rb").read()
    if b[:27] != b"RUU CROSS SYSTEM MAP FORMAT":
        raise ValueError("%s is not a PCRaster map" % path)
    cell_repr = struct.unpack_from("<H", b, 66)[0]
    rows, cols = struct.unpack_from("<II", b, 100)
    dt = {0x00: "u1", 0x04: "i1", 0x11: "<u2", 0x15: "<i2", 0x22: "<u4", 0x26: "<i4", 0x5a: "<f4", 0xdb: "<f8"}[cell_repr]
    return np.frombuffer(b, dt, count=rows * cols, offset=256).reshape(rows, cols).copy()
