"""Minimal pure-Python reader for the NetCDF-4 (= HDF5) traces the reference writes with `az.to_netcdf`
(age_ens_runs_mcmc/run_age_mcmc_utils.py:425, ng_interp/noble_gas_mcmc.py:288) and reads back with `az.from_netcdf`
(run_age_mcmc_utils.py:434, age_modeling_mcmc.post_plots.py:119-148).

The image has neither h5py nor netCDF4, so this restates just enough of the published HDF5 file format (HDF5 File
Format Specification v3) to walk groups and read numeric datasets:
  superblock v2/v3; version-2 object headers (OHDR / OCHK continuation blocks); link messages stored compactly in the
  header or densely in a fractal heap (FRHP / FHIB / FHDB); dataspace v1/v2; fixed-point, floating-point and fixed-length
  string datatypes; contiguous, compact and chunked (layout v3, version-1 B-tree) storage; deflate + shuffle filters;
  scalar / 1-D attributes of those types.
It is host-side I/O, not part of the GPU hot path.

    tr = read_trace(path)         # {"posterior": {"tau1": array[chain, draw], ...}, "sample_stats": {...},
                                  #  "observed_data": {...}, "attrs": {group: {name: value}}}
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.f = fh.read()
        f = self.f
        if f[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("%s is not an HDF5 file" % path)
        ver = f[8]
        if ver not in (2, 3):
            raise ValueError("superblock version %d not supported (the reference's traces use version 2)" % ver)
        if f[9] != 8 or f[10] != 8:
            raise ValueError("only 8-byte offsets / lengths are supported")
        self.root = self._u(12 + 24, 8)

    def _u(self, off, n):
        return int.from_bytes(self.f[off:off + n], "little")

    # ---- object headers (version 2) ----
    def messages(self, addr):
        f = self.f
        if f[addr:addr + 4] != b"OHDR" or f[addr + 4] != 2:
            raise ValueError("object header at %d is not version 2" % addr)
        flags = f[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        n = 1 << (flags & 3)
        csize = self._u(p, n)
        p += n
        out, blocks = [], [(p, p + csize)]
        while blocks:
            p, end = blocks.pop(0)
            while p + 4 <= end:
                t, sz = f[p], self._u(p + 1, 2)
                p += 4
                if flags & 0x04:
                    p += 2
                data = f[p:p + sz]
                if t == 0x10:                                   # continuation
                    coff, clen = struct.unpack("<QQ", data[:16])
                    if f[coff:coff + 4] != b"OCHK":
                        raise ValueError("bad continuation block")
                    blocks.append((coff + 4, coff + clen - 4))
                elif t != 0:
                    out.append((t, data))
                p += sz
        return out

    # ---- links ----
    @staticmethod
    def _parse_link(d, p=0):
        """Link message at d[p:] -> (name, address or None, next offset)."""
        flags = d[p + 1]
        q = p + 2
        ltype = 0
        if flags & 0x08:
            ltype = d[q]
            q += 1
        if flags & 0x04:
            q += 8
        if flags & 0x10:
            q += 1
        n = 1 << (flags & 3)
        ln = int.from_bytes(d[q:q + n], "little")
        q += n
        name = d[q:q + ln].decode("utf-8", "replace")
        q += ln
        addr = None
        if ltype == 0:
            addr = int.from_bytes(d[q:q + 8], "little")
            q += 8
        elif ltype == 1:                                       # soft link: length + path
            sl = int.from_bytes(d[q:q + 2], "little")
            q += 2 + sl
        else:
            sl = int.from_bytes(d[q:q + 2], "little")
            q += 2 + sl
        return name, addr, q

    def _heap_links(self, heap_addr):
        """Links stored densely: walk the fractal heap's direct blocks and parse the packed link messages."""
        f = self.f
        if f[heap_addr:heap_addr + 4] != b"FRHP":
            raise ValueError("bad fractal heap header")
        p = heap_addr + 5
        heap_id_len, io_filter_len, flags = self._u(p, 2), self._u(p + 2, 2), f[p + 4]
        p += 5
        p += 4                                                  # max size of managed objects
        p += 8 * 12                                             # next huge id .. managed object counts (12 x 8 bytes)
        table_width = self._u(p, 2)
        start_block = self._u(p + 2, 8)
        max_direct = self._u(p + 10, 8)
        max_heap_bits = self._u(p + 18, 2)
        root_addr = self._u(p + 22, 8)
        cur_rows = self._u(p + 30, 2)
        if io_filter_len:
            raise ValueError("filtered fractal heaps are not supported")
        off_bytes = (max_heap_bits + 7) // 8
        has_cksum = bool(flags & 0x02)
        links = []

        def direct(addr, size):
            if addr == UNDEF:
                return
            if f[addr:addr + 4] != b"FHDB":
                raise ValueError("bad fractal heap direct block")
            q = addr + 5 + 8 + off_bytes + (4 if has_cksum else 0)
            end = addr + size
            while q + 4 < end and f[q] == 1:                    # link message version 1
                name, a, q2 = self._parse_link(f, q)
                if not name:
                    break
                links.append((name, a))
                q = q2

        def indirect(addr, nrows):
            if f[addr:addr + 4] != b"FHIB":
                raise ValueError("bad fractal heap indirect block")
            q = addr + 5 + 8 + off_bytes
            max_direct_rows = 2
            s = start_block
            while s < max_direct:
                s *= 2
                max_direct_rows += 1
            for r in range(nrows):
                size = start_block * (1 if r < 2 else 2 ** (r - 1))
                for _ in range(table_width):
                    child = self._u(q, 8)
                    q += 8
                    if r < max_direct_rows:
                        direct(child, size)
                    elif child != UNDEF:
                        raise ValueError("nested indirect blocks are not supported")

        if cur_rows == 0:
            direct(root_addr, start_block)
        else:
            indirect(root_addr, cur_rows)
        return links

    def links(self, addr):
        out = []
        for t, d in self.messages(addr):
            if t == 0x06:
                name, a, _ = self._parse_link(d)
                if a is not None:
                    out.append((name, a))
            elif t == 0x02:                                     # link info: dense storage in a fractal heap
                fl = d[1]
                q = 2 + (8 if fl & 1 else 0)
                heap = int.from_bytes(d[q:q + 8], "little")
                if heap != UNDEF:
                    out.extend((n, a) for n, a in self._heap_links(heap) if a is not None)
        return out

    # ---- datasets ----
    @staticmethod
    def _dtype(d):
        cls, size = d[0] & 0x0F, int.from_bytes(d[4:8], "little")
        bits0 = d[1]
        endian = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype("%s%s%d" % (endian, "i" if bits0 & 0x08 else "u", size))
        if cls == 1:
            return np.dtype("%sf%d" % (endian, size))
        if cls == 3:
            return np.dtype("S%d" % size)
        return None                                              # references, compounds, vlen strings: not needed here

    @staticmethod
    def _dataspace(d):
        ver, rank, flags = d[0], d[1], d[2]
        p = 8 if ver == 1 else 4
        return tuple(int.from_bytes(d[p + 8 * i:p + 8 * i + 8], "little") for i in range(rank))

    @staticmethod
    def _filters(d):
        ver, nf = d[0], d[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(nf):
            fid = int.from_bytes(d[p:p + 2], "little")
            p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = int.from_bytes(d[p:p + 2], "little")
                p += 2
            p += 2
            ncd = int.from_bytes(d[p:p + 2], "little")
            p += 2
            if nlen:
                p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cd = [int.from_bytes(d[p + 4 * i:p + 4 * i + 4], "little") for i in range(ncd)]
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _chunks(self, btree, rank):
        """(offsets, address, stored size, filter mask) of every chunk below a version-1 chunk B-tree node."""
        f = self.f
        if f[btree:btree + 4] != b"TREE" or f[btree + 4] != 1:
            raise ValueError("bad chunk B-tree node")
        level, used = f[btree + 5], self._u(btree + 6, 2)
        p = btree + 24
        keysz = 8 + 8 * (rank + 1)
        out = []
        for _ in range(used):
            size, mask = self._u(p, 4), self._u(p + 4, 4)
            offs = tuple(self._u(p + 8 + 8 * i, 8) for i in range(rank))
            child = self._u(p + keysz, 8)
            p += keysz + 8
            if level == 0:
                out.append((offs, child, size, mask))
            else:
                out.extend(self._chunks(child, rank))
        return out

    def dataset(self, addr):
        """numpy array of the dataset whose object header is at addr, or None for an unsupported datatype."""
        dt = shape = layout = None
        filters = []
        for t, d in self.messages(addr):
            if t == 0x01:
                shape = self._dataspace(d)
            elif t == 0x03:
                dt = self._dtype(d)
            elif t == 0x08:
                layout = d
            elif t == 0x0B:
                filters = self._filters(d)
        if dt is None or shape is None or layout is None:
            return None
        f = self.f
        n = int(np.prod(shape)) if shape else 1
        ver, cls = layout[0], layout[1]
        if ver != 3:
            raise ValueError("data layout version %d not supported" % ver)
        if cls == 0:                                             # compact
            size = int.from_bytes(layout[2:4], "little")
            return np.frombuffer(layout[4:4 + size], dtype=dt, count=n).reshape(shape).copy()
        if cls == 1:                                             # contiguous
            a = int.from_bytes(layout[2:10], "little")
            if a == UNDEF:
                return np.zeros(shape, dtype=dt)
            return np.frombuffer(f, dtype=dt, count=n, offset=a).reshape(shape).copy()
        if cls != 2:
            raise ValueError("unknown layout class %d" % cls)
        rank = layout[2] - 1
        btree = int.from_bytes(layout[3:11], "little")
        cdims = tuple(int.from_bytes(layout[11 + 4 * i:15 + 4 * i], "little") for i in range(rank))
        out = np.zeros(shape, dtype=dt)
        if btree == UNDEF:
            return out
        for offs, a, size, mask in self._chunks(btree, rank):
            raw = f[a:a + size]
            for k, (fid, cd) in reversed(list(enumerate(filters))):
                if mask & (1 << k):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:                                   # shuffle: byte planes -> elements
                    es = cd[0] if cd else dt.itemsize
                    m = len(raw) // es
                    raw = np.frombuffer(raw[:m * es], dtype=np.uint8).reshape(es, m).T.tobytes() + raw[m * es:]
                elif fid == 3:                                   # fletcher32: checksum appended
                    raw = raw[:-4]
                else:
                    raise ValueError("filter %d not supported" % fid)
            chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cdims))).reshape(cdims)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out

    def attributes(self, addr):
        """{name: scalar / array / bytes} of the simple numeric and fixed-string attributes of an object."""
        out = {}
        for t, d in self.messages(addr):
            if t != 0x0C:
                continue
            ver = d[0]
            nlen, dtlen, dslen = (int.from_bytes(d[2 + 2 * i:4 + 2 * i], "little") for i in range(3))
            p = 8 if ver == 1 else (9 if ver == 3 else 8)

            def pad(x):
                return (x + 7) // 8 * 8 if ver == 1 else x
            name = d[p:p + nlen].split(b"\0")[0].decode("utf-8", "replace")
            p += pad(nlen)
            dt = self._dtype(d[p:p + dtlen])
            p += pad(dtlen)
            shape = self._dataspace(d[p:p + dslen]) if d[p + 1] else ()
            p += pad(dslen)
            if dt is None:
                continue
            n = int(np.prod(shape)) if shape else 1
            if len(d) - p < n * dt.itemsize:
                continue
            v = np.frombuffer(d, dtype=dt, count=n, offset=p).reshape(shape)
            if dt.kind == "S":
                v = v.reshape(-1)[0].split(b"\0")[0].decode("utf-8", "replace") if n == 1 else [x.decode() for x in v.reshape(-1)]
            elif not shape:
                v = v.reshape(()).item()
            out[name] = v
        return out


def read_trace(path):
    """All numeric variables of every group of an ArviZ InferenceData file: {group: {var: ndarray}} + {"attrs": {...}}."""
    h = H5File(path)
    res = {"attrs": {"/": h.attributes(h.root)}}
    for gname, gaddr in h.links(h.root):
        try:
            members = h.links(gaddr)
        except ValueError:
            continue
        grp = {}
        for vname, vaddr in members:
            try:
                a = h.dataset(vaddr)
            except ValueError:
                a = None
            if a is not None:
                grp[vname] = a
        res[gname] = grp
        res["attrs"][gname] = h.attributes(gaddr)
    return res
