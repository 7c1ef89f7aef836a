"""The HDF5 file format subset: two independent implementations -- C++
(k-wave-fluid-cuda_b200/csrc/minih5, what the host binary and the reference build link) and Python (tools/h5lite.py) -- read
each other's files, covering what libhdf5 / MATLAB write with default settings for k-Wave files: superblock 0, version-1 object
headers (with a continuation block), symbol-table groups (several symbol-table nodes), contiguous / compact / chunked
datasets (multi-level chunk B-tree, clipped edge chunks), deflate, float32 / uint64 / narrower integer and float64 data,
string / float / long long attributes.  No libhdf5 or h5py exists in this image: what is pinned here is the specification's
byte layout (hand-assembled structures below) and the agreement of the two implementations."""
import os
import struct
import subprocess
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import h5lite  # noqa: E402

PKG = os.path.join(ROOT, "k-wave-fluid-cuda_b200")


@pytest.fixture(scope="module")
def recode(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("h5") / "h5_recode")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O1", "-I", os.path.join(PKG, "csrc", "minih5"), os.path.join(ROOT, "tests", "cpp", "h5_recode.cpp"),
                    os.path.join(PKG, "csrc", "minih5", "minih5.cpp"), "-o", exe, "-lz"], check=True)  # fmt: skip
    return exe


def sample_objects():
    rng = np.random.default_rng(5)
    objs = {"/": {"kind": "group", "attrs": {"file_type": "input", "major_version": "1", "created_by": "test"}},
            "/Nx": {"kind": "u64", "data": np.array(48, np.uint64).reshape(1, 1, 1), "attrs": {"data_type": "long", "domain_type": "real"}},
            "/dt": {"kind": "f32", "data": np.array(2e-8, np.float32).reshape(1, 1, 1), "attrs": {"data_type": "float", "domain_type": "real"}},
            # 5 x 3 x 2 = 30 chunks with clipped edges, deflate
            "/c0": {"kind": "f32", "data": rng.standard_normal((20, 33, 47)).astype(np.float32), "chunk": (4, 11, 32), "deflate": 3,
                    "attrs": {"data_type": "float", "domain_type": "real"}},
            # 300 chunks: a two-level chunk B-tree (64 children per node); deflate level 0 (the reference registers the filter at level 0 too)
            "/p": {"kind": "f32", "data": rng.standard_normal((1, 300, 17)).astype(np.float32), "chunk": (1, 1, 17), "deflate": 0,
                   "attrs": {"data_type": "float", "domain_type": "real", "c_harmonics": 3, "c_period": 12.5}},
            "/sensor_mask_index": {"kind": "u64", "data": rng.integers(1, 10**12, (1, 1, 1000)).astype(np.uint64),
                                   "attrs": {"data_type": "long", "domain_type": "real"}},
            "/p_max": {"kind": "group", "attrs": {}},
            "/p_max/1": {"kind": "f32", "data": rng.standard_normal((4, 5, 6)).astype(np.float32), "chunk": (1, 5, 6), "attrs": {"data_type": "float"}},
            "/p_max/2": {"kind": "f32", "data": rng.standard_normal((2, 3, 9, 2)).astype(np.float32), "attrs": {"data_type": "float"}}}  # fmt: skip
    for i in range(150):  # more members than one symbol-table node holds (2 * 32)
        objs[f"/scalar_{i:03d}"] = {"kind": "f32", "data": np.full((1, 1, 1), i, np.float32), "attrs": {}}
    return objs


def test_python_round_trip(tmp_path):
    objs = sample_objects()
    path = str(tmp_path / "a.h5")
    h5lite.write_hdf5(path, objs)
    back = h5lite.read_hdf5(path)
    assert set(back) == set(objs)
    for p, o in objs.items():
        assert back[p]["kind"] == o["kind"]
        if o["kind"] != "group":
            assert np.array_equal(back[p]["data"], o["data"]), p
            assert tuple(back[p]["chunk"]) == tuple(o.get("chunk", ()))
            assert back[p]["deflate"] == o.get("deflate")
        for k, v in o.get("attrs", {}).items():
            assert back[p]["attrs"][k] == (pytest.approx(v) if isinstance(v, float) else v)
    assert h5lite.read_root_attrs(path)["file_type"] == "input"


def test_cpp_reads_python_files_and_python_reads_cpp_files(tmp_path, recode):
    objs = sample_objects()
    a, b = str(tmp_path / "a.h5"), str(tmp_path / "b.h5")
    h5lite.write_hdf5(a, objs)
    names = [p for p, o in objs.items() if o["kind"] != "group" and not p.startswith("/scalar_")] + ["/scalar_007", "/scalar_149"]
    r = subprocess.run([recode, a, b] + names, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    back = h5lite.read_hdf5(b)
    assert set(back) == set(names) | {"/", "/p_max"}
    assert back["/"]["attrs"] == objs["/"]["attrs"]
    for p in names:
        assert np.array_equal(back[p]["data"], objs[p]["data"]), p
        assert back[p]["deflate"] == 4 and len(back[p]["chunk"]) == objs[p]["data"].ndim
        for k in ("data_type", "domain_type"):
            assert back[p]["attrs"].get(k) == objs[p].get("attrs", {}).get(k), (p, k)
    assert back["/p"]["attrs"]["c_harmonics"] == 3 and back["/p"]["attrs"]["c_period"] == 12.5
    assert back["/p"]["chunk"] == (1, 3, 17)  # 100 clipped chunks written by the C++ side


def _hand_made_file():
    """A file assembled by hand from the specification's tables, using features OTHER writers produce and ours never do: superblock
    version 1, an object header continuation block holding the attribute, a NIL message, a version-2 dataspace, a compact
    dataset, int32 and float64 element types, a version-2 filter pipeline and a space-padded string attribute."""
    U = 0xFFFFFFFFFFFFFFFF
    blob = bytearray(4096)

    def msg(t, body, flags=0):
        body = body + b"\0" * (-len(body) % 8)
        return struct.pack("<HHB3x", t, len(body), flags) + body

    def header(msgs, nmsgs):
        body = b"".join(msgs)
        return struct.pack("<BBHII4x", 1, 0, nmsgs, 1, len(body)) + body

    f32 = struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    f64 = struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    i32 = struct.pack("<BBBBI", 0x10, 0x08, 0, 0, 4) + struct.pack("<HH", 0, 32)
    space_v2 = lambda dims: struct.pack("<BBBB", 2, len(dims), 0, 1) + b"".join(struct.pack("<Q", d) for d in dims)  # noqa: E731
    # attribute (version 1) with a space-padded string, placed in a continuation block at 2048
    name = b"data_type\0"
    sdt = struct.pack("<BBBBI", 0x13, 0x02, 0, 0, 8)
    ssp = struct.pack("<BBB5x", 1, 0, 0)
    pad = lambda b: b + b"\0" * (-len(b) % 8)  # noqa: E731
    attr = msg(0x0C, struct.pack("<BBHHH", 1, 0, len(name), len(sdt), len(ssp)) + pad(name) + pad(sdt) + pad(ssp) + b"long    ")
    blob[2048 : 2048 + len(attr)] = attr
    # dataset A: compact int32 [2][3] with the continuation + a NIL message
    a_data = np.arange(6, dtype="<i4") - 2
    a_hdr = header([msg(0x01, space_v2((2, 3))), msg(0x03, i32, 1), msg(0x00, b"\0" * 8),
                    msg(0x08, struct.pack("<BBH", 3, 0, a_data.nbytes) + a_data.tobytes()), msg(0x10, struct.pack("<QQ", 2048, len(attr)))], 6)
    blob[512 : 512 + len(a_hdr)] = a_hdr
    # dataset B: float64 [4][2], one deflated chunk, version-2 filter pipeline
    b_data = (np.arange(8, dtype="<f8") * 0.5).reshape(4, 2)
    comp = zlib.compress(b_data.tobytes(), 6)
    blob[3000 : 3000 + len(comp)] = comp
    key = lambda size, off: struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in off) + struct.pack("<Q", 0)  # noqa: E731
    node = b"TREE" + struct.pack("<BBHQQ", 1, 0, 1, U, U) + key(len(comp), (0, 0)) + struct.pack("<Q", 3000) + key(0, (4, 0))
    blob[3200 : 3200 + len(node)] = node
    b_hdr = header([msg(0x01, struct.pack("<BBB5x", 1, 2, 0) + struct.pack("<QQ", 4, 2)), msg(0x03, f64, 1),
                    msg(0x0B, struct.pack("<BB", 2, 1) + struct.pack("<HHH", 1, 0, 1) + struct.pack("<I", 6)),
                    msg(0x08, struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", 3200) + struct.pack("<III", 4, 2, 8))], 4)
    blob[1024 : 1024 + len(b_hdr)] = b_hdr
    # root group: heap at 256 (data 288), one symbol-table node at 3500, B-tree at 3900
    heap = b"\0" * 8 + b"A\0" + b"\0" * 6 + b"B\0" + b"\0" * 6
    blob[288 : 288 + len(heap)] = heap
    blob[256:288] = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, 288)
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, 2) + struct.pack("<QQII16x", 8, 512, 0, 0) + struct.pack("<QQII16x", 16, 1024, 0, 0)
    blob[3500 : 3500 + len(snod)] = snod
    bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, U, U) + struct.pack("<QQQ", 0, 3500, 16)
    blob[3900 : 3900 + len(bt)] = bt
    root = header([msg(0x11, struct.pack("<QQ", 3900, 256))], 1)
    blob[128 : 128 + len(root)] = root
    sb = h5lite.SIG + struct.pack("<BBBBBBBBHHI", 1, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + struct.pack("<HH", 32, 0)
    sb += struct.pack("<QQQQ", 0, U, len(blob), U) + struct.pack("<QQII", 0, 128, 1, 0) + struct.pack("<QQ", 3900, 256)
    blob[0 : len(sb)] = sb
    return bytes(blob), a_data.reshape(2, 3), b_data


def test_foreign_structures_are_read_by_both(tmp_path, recode):
    raw, a_data, b_data = _hand_made_file()
    a, b = str(tmp_path / "hand.h5"), str(tmp_path / "hand_out.h5")
    open(a, "wb").write(raw)
    got = h5lite.read_hdf5(a)
    assert np.array_equal(got["/A"]["data"], a_data) and got["/A"]["attrs"]["data_type"].strip() == "long"
    assert np.array_equal(got["/B"]["data"], b_data) and got["/B"]["deflate"] == 6
    r = subprocess.run([recode, a, b, "/A", "/B"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    back = h5lite.read_hdf5(b)
    assert np.array_equal(back["/A"]["data"].astype(np.int64), a_data.astype(np.int64).astype(np.uint64).astype(np.int64))  # widened to 64 bit
    assert np.array_equal(back["/B"]["data"], b_data.astype(np.float32))
    assert back["/A"]["attrs"]["data_type"].strip() == "long"


def test_unsupported_files_fail_loudly(tmp_path, recode):
    bad = str(tmp_path / "bad.h5")
    open(bad, "wb").write(h5lite.SIG + bytes([2]) + b"\0" * 200)  # superblock version 2 (libver 'latest')
    with pytest.raises(ValueError, match="superblock version 2"):
        h5lite.read_hdf5(bad)
    r = subprocess.run([recode, bad, str(tmp_path / "o.h5"), "/x"], capture_output=True, text=True)
    assert r.returncode != 0
