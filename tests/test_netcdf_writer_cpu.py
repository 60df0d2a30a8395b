"""CPU: the NetCDF-4 / HDF5 trace writer (noblegas_rtd_mcmc_b200/netcdf4_writer.py) -- round trip through the repo's own
reader (which parses the reference's traces), checksums, and the netCDF-4 dimension conventions resolved by reference."""
import os
import struct

import numpy as np

from noblegas_rtd_mcmc_b200 import netcdf4_reader as R
from noblegas_rtd_mcmc_b200 import netcdf4_writer as W


def test_lookup3_known_answers():
    # Bob Jenkins' lookup3.c self-test vectors for hashlittle()
    assert W.lookup3(b"Four score and seven years ago", 0) == 0x17770551
    assert W.lookup3(b"Four score and seven years ago", 1) == 0xcd628161
    assert W.lookup3(b"", 0) == 0xdeadbeef


def _write(tmp_path):
    rng = np.random.default_rng(0)
    post = {"tau1": rng.uniform(1, 1000, (3, 50)), "eta1": rng.uniform(1, 5, (3, 50)), "vec": rng.normal(size=(3, 50, 4))}
    stats = {"accepted": (rng.uniform(size=(3, 50)) < 0.3).astype(np.int64), "lambda": rng.uniform(size=(3, 50))}
    obs = {"CFC12": np.array([301.5]), "H3": np.array([4.87, 5.1])}
    path = str(tmp_path / "trace.netcdf")
    W.write_trace(path, post, stats, obs, attrs={"sampling_time": np.array([0.49]), "tuning_steps": np.array([10000])})
    return path, post, stats, obs


def test_round_trip_through_the_reader(tmp_path):
    path, post, stats, obs = _write(tmp_path)
    tr = R.read_trace(path)
    for k, v in post.items():
        assert np.array_equal(tr["posterior"][k], v), k
    assert np.array_equal(tr["posterior"]["chain"], np.arange(3)) and np.array_equal(tr["posterior"]["draw"], np.arange(50))
    assert np.array_equal(tr["posterior"]["vec_dim_0"], np.arange(4))
    for k, v in stats.items():
        assert np.array_equal(tr["sample_stats"][k], v), k
    assert np.array_equal(tr["observed_data"]["H3"], obs["H3"])
    a = tr["attrs"]["posterior"]
    assert a["inference_library"] == "ngrtd-b200" and abs(float(np.ravel(a["sampling_time"])[0]) - 0.49) < 1e-15
    assert int(np.ravel(a["tuning_steps"])[0]) == 10000 and "created_at" in a
    # diagnostics.load_trace recognises the HDF5 signature and returns the same layout as for the reference's files
    from noblegas_rtd_mcmc_b200 import diagnostics
    lt = diagnostics.load_trace(path)
    assert np.array_equal(lt["posterior"]["tau1"], post["tau1"]) and "chain" not in lt["posterior"]


def test_checksums_and_dimension_conventions(tmp_path):
    path, post, stats, obs = _write(tmp_path)
    h = R.H5File(path)
    f = h.f
    assert W.lookup3(f[:44]) == struct.unpack("<I", f[44:48])[0]               # superblock v2 checksum
    assert struct.unpack("<Q", f[28:36])[0] == len(f)                            # end-of-file address

    def check_header(addr):
        assert f[addr:addr + 4] == b"OHDR" and f[addr + 4] == 2
        n = struct.unpack("<I", f[addr + 6:addr + 10])[0]
        end = addr + 10 + n
        assert W.lookup3(f[addr:end]) == struct.unpack("<I", f[end:end + 4])[0]
    check_header(h.root)
    groups = dict(h.links(h.root))
    assert set(groups) == {"posterior", "sample_stats", "observed_data"}
    for g, ga in groups.items():
        check_header(ga)
        members = dict(h.links(ga))
        dimids = {}
        for v, va in members.items():
            check_header(va)
            at = h.attributes(va)
            if at.get("CLASS") == "DIMENSION_SCALE":
                assert at["NAME"] == v
                dimids[int(at["_Netcdf4Dimid"])] = v
        # every variable names its dimensions twice: by id (_Netcdf4Coordinates) and by reference (DIMENSION_LIST -> global heap)
        for v, va in members.items():
            at = h.attributes(va)
            if at.get("CLASS") == "DIMENSION_SCALE":
                continue
            ids = [int(x) for x in np.atleast_1d(at["_Netcdf4Coordinates"])]
            names = [dimids[i] for i in ids]
            arr = h.dataset(va)
            assert arr.ndim == len(names)
            for t, d in h.messages(va):
                if t == 0x0C and b"DIMENSION_LIST\0" in d[:32]:
                    nlen, dtl, dsl = struct.unpack("<HHH", d[2:8])
                    p = 9 + nlen + dtl + dsl
                    assert d[9 + nlen:9 + nlen + dtl] == W.DT_VLEN_REF
                    for ax in range(arr.ndim):
                        ln, gaddr, idx = struct.unpack("<IQI", d[p + 16 * ax:p + 16 * ax + 16])
                        assert ln == 1 and f[gaddr:gaddr + 4] == b"GCOL"
                        q = gaddr + 16
                        while True:                                        # walk the heap objects to index idx
                            oi, rc, _, sz = struct.unpack("<HHIQ", f[q:q + 16])
                            if oi == idx:
                                ref = struct.unpack("<Q", f[q + 16:q + 24])[0]
                                break
                            assert oi != 0
                            q += 16 + (sz + 7) // 8 * 8
                        assert ref == members[names[ax]], (g, v, ax)
                    break
            else:
                raise AssertionError("no DIMENSION_LIST on %s/%s" % (g, v))


def test_fixed_messages_equal_the_reference_files():
    """The datatype messages the writer emits are the ones found in the reference's own traces (netCDF 4.8.1 / HDF5 1.12.1);
    checked against the live files when the reference tree is mounted."""
    import glob
    files = sorted(glob.glob("/root/reference/age_ens_runs_mcmc/conv_traces/*.netcdf"))
    if not files:
        import pytest
        pytest.skip("reference tree not mounted")
    h = R.H5File(files[0])
    seen = set()
    for g, ga in h.links(h.root):
        for v, va in h.links(ga):
            for t, d in h.messages(va):
                if t == 0x03:
                    seen.add(bytes(d))
                if t == 0x0C:
                    nlen, dtl, dsl = struct.unpack("<HHH", d[2:8])
                    seen.add(bytes(d[9 + nlen:9 + nlen + dtl]))
    for dt in (W.DT_I64, W.DT_I32, W.DT_F64, W.DT_VLEN_REF, W.DT_REFLIST, W.dt_string(16)):
        assert dt in seen, dt.hex()
