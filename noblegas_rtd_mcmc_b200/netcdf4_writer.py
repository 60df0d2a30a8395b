"""Minimal pure-Python WRITER of the NetCDF-4 (= HDF5) trace files the reference produces with `az.to_netcdf`
(age_ens_runs_mcmc/run_age_mcmc_utils.py:419-425, ng_interp/noble_gas_mcmc.py:419-447) and its plotting scripts read back
with `az.from_netcdf` (age_modeling_mcmc.post_plots.py:119-148, ng_interp/noble_gas_mcmc.compplots.py:222-233), so that
GPU runs can be consumed by those scripts.  The image has neither h5py / netCDF4 nor arviz; when arviz IS importable,
`diagnostics.to_inference_data` hands the same arrays to `az.from_dict` and the library writes the file itself.

Layout of the file (HDF5 File Format Specification v3; every structure below also occurs, byte for byte in its fixed
parts, in the reference's own traces -- netCDF 4.8.1 / HDF5 1.12.1 -- which noblegas_rtd_mcmc_b200/netcdf4_reader.py parses):
  superblock v2; version-2 object headers with Jenkins lookup3 checksums; groups with compact link storage;
  contiguous little-endian int64 / float64 datasets; version-3 attribute messages (fixed-length strings, int32 / int64 /
  float64 scalars and vectors); the netCDF-4 dimension conventions: every dimension is a coordinate dataset carrying
  CLASS = "DIMENSION_SCALE", NAME, _Netcdf4Dimid and a REFERENCE_LIST (compound {object reference, int32}); every variable
  carries _Netcdf4Coordinates (its dimension ids) and a DIMENSION_LIST (variable-length object references stored in a
  global heap collection) -- the two routes by which netCDF-C / h5netcdf attach dimensions to variables.
Groups follow ArviZ's InferenceData schema: `posterior/<var>[chain, draw(, <var>_dim_0)]`, `sample_stats/...`,
`observed_data/<var>[<var>_dim_0]`, group attributes created_at / arviz_version / inference_library / sampling_time.
Host-side I/O, not part of the GPU hot path.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


# ---- Jenkins lookup3 (hashlittle), the checksum of HDF5 metadata (H5_checksum_lookup3) ----
def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data, initval=0):
    length = len(data)
    a = b = c = (0xdeadbeef + length + initval) & 0xFFFFFFFF
    p = 0
    while length > 12:
        a = (a + int.from_bytes(data[p:p + 4], "little")) & 0xFFFFFFFF
        b = (b + int.from_bytes(data[p + 4:p + 8], "little")) & 0xFFFFFFFF
        c = (c + int.from_bytes(data[p + 8:p + 12], "little")) & 0xFFFFFFFF
        a = (a - c) & 0xFFFFFFFF; a ^= _rot(c, 4); c = (c + b) & 0xFFFFFFFF
        b = (b - a) & 0xFFFFFFFF; b ^= _rot(a, 6); a = (a + c) & 0xFFFFFFFF
        c = (c - b) & 0xFFFFFFFF; c ^= _rot(b, 8); b = (b + a) & 0xFFFFFFFF
        a = (a - c) & 0xFFFFFFFF; a ^= _rot(c, 16); c = (c + b) & 0xFFFFFFFF
        b = (b - a) & 0xFFFFFFFF; b ^= _rot(a, 19); a = (a + c) & 0xFFFFFFFF
        c = (c - b) & 0xFFFFFFFF; c ^= _rot(b, 4); b = (b + a) & 0xFFFFFFFF
        p += 12
        length -= 12
    if length == 0:
        return c
    tail = data[p:p + length] + b"\0" * (12 - length)
    a = (a + int.from_bytes(tail[0:4], "little")) & 0xFFFFFFFF
    b = (b + int.from_bytes(tail[4:8], "little")) & 0xFFFFFFFF
    c = (c + int.from_bytes(tail[8:12], "little")) & 0xFFFFFFFF
    c ^= b; c = (c - _rot(b, 14)) & 0xFFFFFFFF
    a ^= c; a = (a - _rot(c, 11)) & 0xFFFFFFFF
    b ^= a; b = (b - _rot(a, 25)) & 0xFFFFFFFF
    c ^= b; c = (c - _rot(b, 16)) & 0xFFFFFFFF
    a ^= c; a = (a - _rot(c, 4)) & 0xFFFFFFFF
    b ^= a; b = (b - _rot(a, 14)) & 0xFFFFFFFF
    c ^= b; c = (c - _rot(b, 24)) & 0xFFFFFFFF
    return c


# ---- datatype / dataspace messages ----
DT_I64 = bytes.fromhex("100800000800000000004000")
DT_I32 = bytes.fromhex("100800000400000000002000")
DT_F64 = bytes.fromhex("11203f000800000000004000340b0034ff030000")
DT_REF = bytes.fromhex("1700000008000000")
DT_VLEN_REF = bytes.fromhex("1900000010000000") + DT_REF
DT_REFLIST = (bytes.fromhex("3602000010000000") + b"dataset\0" + b"\x00" + DT_REF + b"dimension\0" + b"\x08" + DT_I32)


def dt_string(n):
    return bytes.fromhex("13000000") + struct.pack("<I", n)


def dataspace(shape):
    if shape == ():
        return bytes([2, 0, 0, 0])
    return bytes([2, len(shape), 1, 1]) + b"".join(struct.pack("<Q", s) for s in shape) * 2


def _np_dt(a):
    if a.dtype.kind == "f":
        return DT_F64, a.astype("<f8")
    if a.dtype.kind in "iub":
        return (DT_I32, a.astype("<i4")) if a.dtype.itemsize <= 4 and a.dtype.kind != "b" else (DT_I64, a.astype("<i8"))
    raise TypeError("unsupported dtype %s" % a.dtype)


def attribute(name, value):
    """Attribute message (version 3).  value: str | int | float | 1-D numpy array | ("raw", datatype, dataspace, bytes)."""
    nm = name.encode() + b"\0"
    if isinstance(value, tuple) and value and value[0] == "raw":
        _, dt, ds, data = value
    elif isinstance(value, str):
        raw = value.encode() + b"\0"
        dt, ds, data = dt_string(len(raw)), dataspace(()), raw
    else:
        a = np.asarray(value)
        if a.dtype.kind in "US":
            return attribute(name, str(a.reshape(-1)[0]) if a.size == 1 else ",".join(str(x) for x in a.reshape(-1)))
        dt, arr = _np_dt(np.atleast_1d(a) if a.ndim else a)
        ds = dataspace(tuple(a.shape))
        data = arr.tobytes()
    return struct.pack("<BBHHHB", 3, 0, len(nm), len(dt), len(ds), 0) + nm + dt + ds + data


class _Obj:
    """An object header under construction: a list of (message type, payload) built by a function of the address table."""

    def __init__(self, build):
        self.build = build
        self.addr = None

    def header(self, tab):
        msgs = self.build(tab)
        body = b"".join(struct.pack("<BHB", t, len(d), 0) + d for t, d in msgs)
        head = b"OHDR" + bytes([2, 0x02]) + struct.pack("<I", len(body)) + body
        return head + struct.pack("<I", lookup3(head))


def write_netcdf(path, groups, root_attrs=None):
    """groups: {group name: {"dims": {dim: length}, "vars": {var: (dims tuple, ndarray)}, "attrs": {...}}}.
    Every dimension becomes a coordinate variable 0..n-1 (int64), as ArviZ writes chain / draw / <var>_dim_0."""
    objs = {}                               # key -> _Obj; keys: ("root",), ("g", g), ("v", g, name)
    raw = {}                                # ("v", g, name) -> bytes of the dataset
    heap_objs = []                          # (key of the referenced dim dataset) in heap-object order, index = position + 1
    heap_index = {}                         # (g, var, axis) -> heap object index

    dimid = {}
    nid = 0
    for g, spec in groups.items():
        for d in spec["dims"]:
            dimid[(g, d)] = nid
            nid += 1
    for g, spec in groups.items():
        for v, (dims, arr) in spec["vars"].items():
            if tuple(np.shape(arr)) != tuple(spec["dims"][d] for d in dims):
                raise ValueError("%s/%s: shape %s does not match dims %s" % (g, v, np.shape(arr), dims))
            if v in spec["dims"]:
                raise ValueError("%s/%s: a variable may not carry the name of a dimension" % (g, v))
            for ax, d in enumerate(dims):
                heap_objs.append(("v", g, d))
                heap_index[(g, v, ax)] = len(heap_objs)

    def dataset_builder(key, shape, dt, attrs_fn):
        def build(tab):
            size = len(raw[key])
            msgs = [(0x01, dataspace(shape)), (0x03, dt), (0x05, bytes([3, 0x0A])),
                    (0x08, bytes([3, 1]) + struct.pack("<QQ", tab.get(("raw",) + key, 0) if size else UNDEF, size))]
            msgs += [(0x0C, a) for a in attrs_fn(tab)]
            return msgs
        return build

    for g, spec in groups.items():
        # dimension scales (coordinate variables)
        for d, n in spec["dims"].items():
            key = ("v", g, d)
            raw[key] = np.arange(n, dtype="<i8").tobytes()
            users = [(v, ax) for v, (dims, _) in spec["vars"].items() for ax, dd in enumerate(dims) if dd == d]

            def attrs_fn(tab, g=g, d=d, users=users):
                out = [attribute("CLASS", "DIMENSION_SCALE"), attribute("NAME", d),
                       attribute("_Netcdf4Dimid", np.int32(dimid[(g, d)])),
                       attribute("_Netcdf4Coordinates", np.array([dimid[(g, d)]], dtype=np.int32))]
                if users:
                    data = b"".join(struct.pack("<QiI", tab.get(("v", g, v), 0), ax, 0) for v, ax in users)
                    out.append(attribute("REFERENCE_LIST", ("raw", DT_REFLIST, dataspace((len(users),)), data)))
                return out
            objs[key] = _Obj(dataset_builder(key, (n,), DT_I64, attrs_fn))
        for v, (dims, arr) in spec["vars"].items():
            key = ("v", g, v)
            dt, a = _np_dt(np.asarray(arr))
            raw[key] = np.ascontiguousarray(a).tobytes()

            def attrs_fn(tab, g=g, v=v, dims=dims, is_f=(dt == DT_F64)):
                out = [attribute("_Netcdf4Coordinates", np.array([dimid[(g, d)] for d in dims], dtype=np.int32)),
                       attribute("_Netcdf4Dimid", np.int32(dimid[(g, dims[0])] if dims else 0))]
                if dims:
                    data = b"".join(struct.pack("<IQI", 1, tab.get(("gcol",), 0), heap_index[(g, v, ax)]) for ax in range(len(dims)))
                    out.append(attribute("DIMENSION_LIST", ("raw", DT_VLEN_REF, dataspace((len(dims),)), data)))
                if is_f:
                    out.append(attribute("_FillValue", np.array([np.nan])))
                return out
            objs[key] = _Obj(dataset_builder(key, tuple(np.shape(arr)), dt, attrs_fn))

        def gbuild(tab, g=g, spec=spec):
            msgs = [(0x02, bytes([0, 0]) + struct.pack("<QQ", UNDEF, UNDEF)), (0x0A, bytes([0, 0]))]
            for name in list(spec["dims"]) + list(spec["vars"]):
                nm = name.encode()
                msgs.append((0x06, bytes([1, 0, len(nm)]) + nm + struct.pack("<Q", tab.get(("v", g, name), 0))))
            msgs += [(0x0C, attribute(k, val)) for k, val in (spec.get("attrs") or {}).items()]
            return msgs
        objs[("g", g)] = _Obj(gbuild)

    def rbuild(tab):
        msgs = [(0x02, bytes([0, 0]) + struct.pack("<QQ", UNDEF, UNDEF)), (0x0A, bytes([0, 0]))]
        for g in groups:
            nm = g.encode()
            msgs.append((0x06, bytes([1, 0, len(nm)]) + nm + struct.pack("<Q", tab.get(("g", g), 0))))
        ra = {"_NCProperties": "version=2,ngrtd-b200=1,hdf5-spec=3"}
        ra.update(root_attrs or {})
        msgs += [(0x0C, attribute(k, val)) for k, val in ra.items()]
        return msgs
    objs[("root",)] = _Obj(rbuild)

    # ---- addresses: header sizes do not depend on the address values, so one sizing pass suffices ----
    order = [("root",)] + [("g", g) for g in groups] + [k for k in objs if k[0] == "v"]
    tab = {}
    pos = 48
    for k in order:
        pos = (pos + 7) & ~7
        tab[k] = pos
        pos += len(objs[k].header({}))
    pos = (pos + 7) & ~7
    tab[("gcol",)] = pos
    gcol_size = max(4096, (16 + 24 * len(heap_objs) + 16 + 4095) // 4096 * 4096)
    pos += gcol_size
    for k in order:
        if k[0] == "v":
            pos = (pos + 7) & ~7
            tab[("raw",) + k] = pos
            pos += len(raw[k])
    eof = pos

    out = bytearray(eof)
    sb = b"\x89HDF\r\n\x1a\n" + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", 0, UNDEF, eof, tab[("root",)])
    out[0:48] = sb + struct.pack("<I", lookup3(sb))
    for k in order:
        h = objs[k].header(tab)
        out[tab[k]:tab[k] + len(h)] = h
    # global heap collection: one 8-byte object reference per (variable, axis)
    g0 = tab[("gcol",)]
    body = bytearray()
    for i, key in enumerate(heap_objs):
        body += struct.pack("<HHIQ", i + 1, 1, 0, 8) + struct.pack("<Q", tab[key])
    free = gcol_size - 16 - len(body)
    body += struct.pack("<HHIQ", 0, 0, 0, free)
    out[g0:g0 + 16 + len(body)] = b"GCOL" + bytes([1, 0, 0, 0]) + struct.pack("<Q", gcol_size) + bytes(body)
    for k in order:
        if k[0] == "v":
            a = tab[("raw",) + k]
            out[a:a + len(raw[k])] = raw[k]
    with open(path, "wb") as fh:
        fh.write(bytes(out))
    return eof


def write_trace(path, posterior, sample_stats=None, observed_data=None, attrs=None):
    """ArviZ-style InferenceData file: posterior / sample_stats arrays are [chain, draw] (or [chain, draw, k]), observed_data
    arrays are 1-D.  attrs: group attributes (created_at, arviz_version, inference_library, sampling_time, tuning_steps ...)
    written on every group, as ArviZ does."""
    import datetime
    ga = {"created_at": datetime.datetime.now(datetime.timezone.utc).replace(tzinfo=None).isoformat(), "arviz_version": "0.11.4-compatible",
          "inference_library": "ngrtd-b200", "inference_library_version": "2"}
    ga.update(attrs or {})
    groups = {}

    def sample_group(d):
        dims, vars_ = {}, {}
        for name, a in d.items():
            a = np.asarray(a)
            if a.ndim < 2:
                raise ValueError("%s must be [chain, draw(, ...)]" % name)
            dims.setdefault("chain", a.shape[0])
            dims.setdefault("draw", a.shape[1])
            if (dims["chain"], dims["draw"]) != a.shape[:2]:
                raise ValueError("%s: all variables of a group share chain / draw" % name)
            vd = ["chain", "draw"]
            for ax, n in enumerate(a.shape[2:]):
                dn = "%s_dim_%d" % (name, ax)
                dims[dn] = n
                vd.append(dn)
            vars_[name] = (tuple(vd), a)
        return {"dims": dims, "vars": vars_, "attrs": dict(ga)}
    groups["posterior"] = sample_group(posterior)
    if sample_stats:
        groups["sample_stats"] = sample_group(sample_stats)
    if observed_data:
        dims, vars_ = {}, {}
        for name, a in observed_data.items():
            a = np.atleast_1d(np.asarray(a, dtype=np.float64))
            dn = "%s_dim_0" % name
            dims[dn] = a.shape[0]
            vars_[name] = ((dn,), a)
        groups["observed_data"] = {"dims": dims, "vars": vars_, "attrs": dict(ga)}
    return write_netcdf(path, groups)
