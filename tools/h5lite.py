"""A reader / writer of the HDF5 subset the k-Wave file layout uses (Hdf5/Hdf5File.cpp:97-1086 of the reference; file layout
main.cpp:350-803), written against the HDF5 File Format Specification 1.8 (no libhdf5 / h5py exists in this image):

  superblock version 0 (and 1) . version-1 object headers with continuation blocks . old-style groups (symbol-table
  message, version-1 B-tree of group nodes + local heap) . datasets with contiguous, compact or chunked layout (version-1
  chunk B-tree, any depth) . deflate filter (zlib) . IEEE float32, 64-bit integer (and 32-bit integer / float64 on read)
  . attributes: fixed-length strings, float32 / int64 scalars.

This is the independent Python counterpart of k-wave-fluid-cuda_b200/csrc/minih5 (C++): files written by one are read by
the other in the tests, which is the only cross-check available here.  What it was verified against: the structure tables
of the specification (signatures, field order and sizes) and this round trip -- NOT a real libhdf5.  Test tooling.

API (shared with kwh5.py):  objects = {abs_path: dict(kind='group'|'f32'|'u64', attrs={...}, data=ndarray, chunk=(), deflate=0)}
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K, CHUNK_K = 32, 16, 32  # symbol-table node holds 2*LEAF_K entries, group / chunk B-tree nodes 2*K children


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


# ---- datatype / dataspace messages ---------------------------------------------------------------------------------
def _dt_f32():
    return struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)


def _dt_int(size, signed):
    return struct.pack("<BBBBI", 0x10, 0x08 if signed else 0x00, 0, 0, size) + struct.pack("<HH", 0, 8 * size)


def _dt_str(size):
    return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, size)  # null-terminated ASCII


def _space(dims):
    return struct.pack("<BBB5x", 1, len(dims), 0) + b"".join(struct.pack("<Q", d) for d in dims)


def _msg(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attr_msg(name, value):
    if isinstance(value, str):
        raw = value.encode() + b"\0"
        dt, sp = _dt_str(len(raw)), _space(())
    elif isinstance(value, (int, np.integer)):
        raw, dt, sp = struct.pack("<q", int(value)), _dt_int(8, True), _space((1,))
    else:
        raw, dt, sp = struct.pack("<f", float(value)), _dt_f32(), _space((1,))
    nm = name.encode() + b"\0"
    return _msg(0x0C, struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + raw)


def _object_header(msgs):
    body = b"".join(msgs)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body


# ---- writer ---------------------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self, f):
        self.f = f
        self.pos = 0

    def alloc(self, n, align=8):
        self.pos += -self.pos % align
        a = self.pos
        self.pos += n
        return a

    def put(self, addr, data):
        self.f.seek(addr)
        self.f.write(data)


def _btree_node(node_type, level, entries, key_size, capacity):
    """entries: [(key_bytes, child_addr)], plus the final key in entries[-1][2]."""
    out = b"TREE" + struct.pack("<BBHQQ", node_type, level, len(entries), UNDEF, UNDEF)
    for k, c, _ in entries:
        out += k + struct.pack("<Q", c)
    out += entries[-1][2]
    full = 24 + capacity * (key_size + 8) + key_size
    return out + b"\0" * (full - len(out))


def _write_chunked(w, data, chunk, deflate):
    """Writes the chunks and their B-tree; returns the address of the root node."""
    rank = data.ndim
    esize = data.dtype.itemsize
    counts = [-(-d // c) for d, c in zip(data.shape, chunk)]
    key_size = 8 + 8 * (rank + 1)
    leaves = []  # (key, addr, next_key) per chunk in row-major (= lexicographic) order
    for idx in np.ndindex(*counts):
        off = [i * c for i, c in zip(idx, chunk)]
        block = np.zeros(chunk, dtype=data.dtype)
        sl = tuple(slice(o, min(o + c, d)) for o, c, d in zip(off, chunk, data.shape))
        block[tuple(slice(0, s.stop - s.start) for s in sl)] = data[sl]
        raw = block.tobytes()
        if deflate is not None:
            raw = zlib.compress(raw, deflate)
        addr = w.alloc(len(raw))
        w.put(addr, raw)
        leaves.append((struct.pack("<II", len(raw), 0) + b"".join(struct.pack("<Q", o) for o in off) + struct.pack("<Q", 0), addr, off))
    # the key that closes a node: the offset of the next chunk, or one chunk past the end of the first dimension
    end = [counts[0] * chunk[0]] + [0] * (rank - 1)
    level_entries = []
    for i, (k, a, off) in enumerate(leaves):
        nxt = leaves[i + 1][2] if i + 1 < len(leaves) else end
        level_entries.append((k, a, struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", o) for o in nxt) + struct.pack("<Q", 0)))
    level = 0
    while True:
        nodes = []
        for i in range(0, len(level_entries), 2 * CHUNK_K):
            grp = level_entries[i : i + 2 * CHUNK_K]
            addr = w.alloc(24 + 2 * CHUNK_K * (key_size + 8) + key_size)
            w.put(addr, _btree_node(1, level, grp, key_size, 2 * CHUNK_K))
            nodes.append((grp[0][0], addr, grp[-1][2]))
        if len(nodes) == 1:
            return nodes[0][1]
        level_entries, level = nodes, level + 1


def _write_dataset(w, o):
    kind = o["kind"]
    data = np.ascontiguousarray(o["data"], dtype=np.float32 if kind == "f32" else np.uint64)
    if data.ndim == 0:
        data = data.reshape(1)
    dims = data.shape
    msgs = [_msg(0x01, _space(dims)), _msg(0x03, _dt_f32() if kind == "f32" else _dt_int(8, False), flags=1)]
    chunk = tuple(int(c) for c in o.get("chunk", ()) or ())
    deflate = o.get("deflate", None)
    if chunk:
        msgs.append(_msg(0x05, struct.pack("<BBBB", 2, 3, 2, 0)))  # fill value v2: incremental allocation, fill if set, undefined
        if deflate is not None:
            msgs.append(_msg(0x0B, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I", int(deflate)) + b"\0" * 4))
        root = _write_chunked(w, data, chunk, int(deflate) if deflate is not None else None)
        msgs.append(_msg(0x08, struct.pack("<BBB", 3, 2, len(dims) + 1) + struct.pack("<Q", root) + b"".join(struct.pack("<I", c) for c in chunk) +
                         struct.pack("<I", data.dtype.itemsize)))  # fmt: skip
    else:
        msgs.append(_msg(0x05, struct.pack("<BBBB", 2, 2, 2, 0)))  # late allocation
        raw = data.tobytes()
        addr = w.alloc(len(raw)) if raw else UNDEF
        if raw:
            w.put(addr, raw)
        msgs.append(_msg(0x08, struct.pack("<BB", 3, 1) + struct.pack("<QQ", addr, len(raw))))
    for k, v in o.get("attrs", {}).items():
        msgs.append(_attr_msg(k, v))
    hdr = _object_header(msgs)
    addr = w.alloc(len(hdr))
    w.put(addr, hdr)
    return addr


def _write_group(w, path, tree, objects):
    """Writes the children first, then heap, symbol-table nodes, B-tree and the group's object header.
    Returns (header address, btree address, heap address)."""
    children = sorted(tree.get(path, []))  # symbol-table entries are ordered by name (strcmp)
    entries = []
    for name in children:
        cp = (path.rstrip("/") + "/" + name) if path != "/" else "/" + name
        if objects[cp]["kind"] == "group":
            h, b, hp = _write_group(w, cp, tree, objects)
            entries.append((name, h, 1, struct.pack("<QQ", b, hp)))
        else:
            entries.append((name, _write_dataset(w, objects[cp]), 0, b"\0" * 16))
    # local heap: offset 0 holds the empty string (the B-tree's first key)
    heap = bytearray(b"\0" * 8)
    offs = {}
    for name, *_ in entries:
        offs[name] = len(heap)
        heap += _pad8(name.encode() + b"\0")
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, 16)  # one free block of 16 bytes closing the segment (next = 1: end of the free list)
    heap_data = w.alloc(len(heap))
    w.put(heap_data, bytes(heap))
    heap_addr = w.alloc(32)
    w.put(heap_addr, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
    snods = []
    for i in range(0, max(len(entries), 1), 2 * LEAF_K):
        grp = entries[i : i + 2 * LEAF_K]
        body = b"SNOD" + struct.pack("<BBH", 1, 0, len(grp))
        for name, haddr, ctype, scratch in grp:
            body += struct.pack("<QQII", offs[name], haddr, ctype, 0) + scratch
        body += b"\0" * (8 + 2 * LEAF_K * 40 - len(body))
        addr = w.alloc(len(body))
        w.put(addr, body)
        snods.append((struct.pack("<Q", 0), addr, struct.pack("<Q", offs[grp[-1][0]] if grp else 0)))
    if len(snods) > 2 * INTERNAL_K:
        raise ValueError(f"group {path}: more than {4 * LEAF_K * INTERNAL_K} members are not supported by this writer")
    for i in range(1, len(snods)):  # key i = the largest name of child i - 1
        snods[i] = (snods[i - 1][2], snods[i][1], snods[i][2])
    btree = w.alloc(24 + 2 * INTERNAL_K * 16 + 8)
    w.put(btree, _btree_node(0, 0, snods, 8, 2 * INTERNAL_K))
    msgs = [_msg(0x11, struct.pack("<QQ", btree, heap_addr))]
    for k, v in objects[path].get("attrs", {}).items():
        msgs.append(_attr_msg(k, v))
    hdr = _object_header(msgs)
    haddr = w.alloc(len(hdr))
    w.put(haddr, hdr)
    return haddr, btree, heap_addr


def write_hdf5(path, objects):
    objects = dict(objects)
    objects.setdefault("/", {"kind": "group", "attrs": {}})
    tree = {}
    for p in objects:
        if p == "/":
            continue
        parent, name = p.rsplit("/", 1)
        parent = parent or "/"
        if parent not in objects:
            raise ValueError(f"{p}: parent group {parent} missing")
        tree.setdefault(parent, []).append(name)
    with open(path, "wb") as f:
        w = _Writer(f)
        w.alloc(96)
        haddr, btree, heap = _write_group(w, "/", tree, objects)
        eof = w.alloc(0)
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0) + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, haddr, 1, 0) + struct.pack("<QQ", btree, heap)
        w.put(0, sb)
        f.truncate(eof)


# ---- reader ---------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf):
        self.b = buf

    def u(self, off, n):
        return int.from_bytes(self.b[off : off + n], "little")


def _messages(r, addr):
    """(type, flags, body) of every message of a version-1 object header, continuation blocks included."""
    ver = r.b[addr]
    if ver != 1:
        raise ValueError(f"object header version {ver} at {addr:#x} is not supported (only version 1: libver 'earliest' files)")
    nmsg, size = r.u(addr + 2, 2), r.u(addr + 8, 4)
    blocks, out = [(addr + 16, size)], []
    while blocks and len(out) < nmsg:
        p, left = blocks.pop(0)
        while left >= 8 and len(out) < nmsg:
            t, s, fl = r.u(p, 2), r.u(p + 2, 2), r.b[p + 4]
            body = r.b[p + 8 : p + 8 + s]
            if t == 0x10:
                blocks.append((r.u(p + 8, 8), r.u(p + 16, 8)))
            out.append((t, fl, body))
            p, left = p + 8 + s, left - 8 - s
    return out


def _parse_dtype(b):
    cls, ver = b[0] & 0x0F, b[0] >> 4
    size = int.from_bytes(b[4:8], "little")
    if cls == 1:
        return ("f", size)
    if cls == 0:
        return ("i" if b[1] & 0x08 else "u", size)
    if cls == 3:
        return ("s", size)
    raise ValueError(f"datatype class {cls} (version {ver}) is not supported")


def _parse_space(b):
    ver, rank = b[0], b[1]
    if ver == 1:
        off = 8
    elif ver == 2:
        off = 4
    else:
        raise ValueError(f"dataspace version {ver}")
    return tuple(int.from_bytes(b[off + 8 * i : off + 8 * i + 8], "little") for i in range(rank))


def _np_dtype(dt):
    k, s = dt
    return np.dtype({"f": "<f", "i": "<i", "u": "<u"}[k] + str(s))


def _read_attr(body):
    ver = body[0]
    ns, ds, ss = (int.from_bytes(body[2 + 2 * i : 4 + 2 * i], "little") for i in range(3))
    if ver == 1:
        p = 8
        pad = lambda n: n + (-n % 8)  # noqa: E731
    elif ver in (2, 3):
        p = 8 if ver == 2 else 9
        pad = lambda n: n  # noqa: E731
    else:
        raise ValueError(f"attribute message version {ver}")
    name = body[p : p + ns].split(b"\0")[0].decode()
    p += pad(ns)
    dt = _parse_dtype(body[p : p + ds])
    p += pad(ds)
    dims = _parse_space(body[p : p + ss])
    p += pad(ss)
    n = int(np.prod(dims)) if dims else 1
    if dt[0] == "s":
        return name, body[p : p + dt[1]].split(b"\0")[0].decode()
    v = np.frombuffer(body[p : p + n * dt[1]], dtype=_np_dtype(dt))
    return name, (float(v[0]) if dt[0] == "f" else int(v[0]))


def _chunk_leaves(r, addr, rank, out):
    assert r.b[addr : addr + 4] == b"TREE" and r.b[addr + 4] == 1, "chunk B-tree node expected"
    level, n = r.b[addr + 5], r.u(addr + 6, 2)
    ks = 8 + 8 * (rank + 1)
    p = addr + 24
    for _ in range(n):
        size, mask = r.u(p, 4), r.u(p + 4, 4)
        off = [r.u(p + 8 + 8 * i, 8) for i in range(rank)]
        child = r.u(p + ks, 8)
        if level:
            _chunk_leaves(r, child, rank, out)
        else:
            out.append((off, size, mask, child))
        p += ks + 8


def _read_dataset(r, msgs):
    dims = dt = layout = None
    filters, attrs = [], {}
    for t, fl, body in msgs:
        if t == 0x01:
            dims = _parse_space(body)
        elif t == 0x03:
            dt = _parse_dtype(body)
        elif t == 0x08:
            layout = body
        elif t == 0x0B:
            ver, nf = body[0], body[1]
            p = 8 if ver == 1 else 2
            for _ in range(nf):
                fid = int.from_bytes(body[p : p + 2], "little")
                if ver == 1 or fid >= 256:
                    nl = int.from_bytes(body[p + 2 : p + 4], "little")
                    ncd = int.from_bytes(body[p + 6 : p + 8], "little")
                    p += 8 + (nl + (-nl % 8) if ver == 1 else nl)
                else:
                    ncd = int.from_bytes(body[p + 4 : p + 6], "little")
                    p += 6
                cd = [int.from_bytes(body[p + 4 * i : p + 4 * i + 4], "little") for i in range(ncd)]
                p += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
                filters.append((fid, cd))
        elif t == 0x0C:
            k, v = _read_attr(body)
            attrs[k] = v
    if dims is None or dt is None or layout is None:
        raise ValueError("dataset without dataspace / datatype / layout message")
    npdt = _np_dtype(dt)
    n = int(np.prod(dims)) if dims else 1
    ver, cls = layout[0], layout[1]
    if ver != 3:
        raise ValueError(f"data layout message version {ver} is not supported")
    chunk, deflate = (), None
    if cls == 1:
        addr, size = int.from_bytes(layout[2:10], "little"), int.from_bytes(layout[10:18], "little")
        data = np.frombuffer(r.b[addr : addr + n * npdt.itemsize], dtype=npdt).reshape(dims) if addr != UNDEF else np.zeros(dims, npdt)
    elif cls == 0:
        size = int.from_bytes(layout[2:4], "little")
        data = np.frombuffer(layout[4 : 4 + size], dtype=npdt).reshape(dims)
    elif cls == 2:
        nd = layout[2]
        root = int.from_bytes(layout[3:11], "little")
        cdims = [int.from_bytes(layout[11 + 4 * i : 15 + 4 * i], "little") for i in range(nd)]
        rank = nd - 1
        chunk = tuple(cdims[:rank])
        for fid, cd in filters:
            if fid != 1:
                raise ValueError(f"filter {fid} is not supported (only deflate)")
            deflate = cd[0] if cd else 0
        data = np.zeros(dims, npdt)
        leaves = []
        if root != UNDEF:
            _chunk_leaves(r, root, rank, leaves)
        for off, size, mask, addr in leaves:
            raw = r.b[addr : addr + size]
            if filters and not (mask & 1):
                raw = zlib.decompress(raw)
            block = np.frombuffer(raw, dtype=npdt, count=int(np.prod(chunk))).reshape(chunk)
            sl = tuple(slice(o, min(o + c, d)) for o, c, d in zip(off, chunk, dims))
            data[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
    else:
        raise ValueError(f"layout class {cls}")
    kind = "f32" if dt == ("f", 4) else "u64" if dt[0] in "ui" and dt[1] == 8 else f"{dt[0]}{dt[1]}"
    return {"kind": kind, "attrs": attrs, "data": data, "chunk": chunk, "deflate": deflate}


def _group_entries(r, btree, heap):
    assert r.b[heap : heap + 4] == b"HEAP", "local heap expected"
    hdata = r.u(heap + 24, 8)
    out = []

    def walk(addr):
        assert r.b[addr : addr + 4] == b"TREE" and r.b[addr + 4] == 0, "group B-tree node expected"
        level, n = r.b[addr + 5], r.u(addr + 6, 2)
        p = addr + 24 + 8
        for _ in range(n):
            child = r.u(p, 8)
            p += 16
            if level:
                walk(child)
                continue
            assert r.b[child : child + 4] == b"SNOD", "symbol table node expected"
            ns = r.u(child + 6, 2)
            for i in range(ns):
                e = child + 8 + 40 * i
                noff, haddr = r.u(e, 8), r.u(e + 8, 8)
                end = r.b.index(b"\0", hdata + noff)
                out.append((r.b[hdata + noff : end].decode(), haddr))

    walk(btree)
    return out


def _read_object(r, addr, path, out):
    msgs = _messages(r, addr)
    st = [b for t, _, b in msgs if t == 0x11]
    if st:
        attrs = dict(_read_attr(b) for t, _, b in msgs if t == 0x0C)
        out[path] = {"kind": "group", "attrs": attrs}
        btree, heap = int.from_bytes(st[0][0:8], "little"), int.from_bytes(st[0][8:16], "little")
        for name, haddr in _group_entries(r, btree, heap):
            _read_object(r, haddr, (path.rstrip("/") + "/" + name), out)
    elif any(t in (0x02, 0x06) for t, _, _ in msgs):
        raise ValueError(f"{path}: new-style groups (link messages) are not supported; write the file with libver 'earliest' (the default)")
    else:
        out[path] = _read_dataset(r, msgs)


def read_hdf5(path):
    with open(path, "rb") as f:
        buf = f.read()
    r = _Reader(buf)
    if buf[:8] != SIG:
        raise ValueError("not an HDF5 file (signature missing at offset 0)")
    ver = buf[8]
    if ver not in (0, 1):
        raise ValueError(f"superblock version {ver} is not supported (versions 0 and 1: files written with libver 'earliest')")
    if buf[13] != 8 or buf[14] != 8:
        raise ValueError("only 8-byte offsets and lengths are supported")
    p = 24 + (4 if ver == 1 else 0)
    base = r.u(p, 8)
    if base != 0:
        raise ValueError("non-zero base address")
    root_ste = p + 32
    haddr = r.u(root_ste + 8, 8)
    out = {}
    _read_object(r, haddr, "/", out)
    return out


def read_root_attrs(path):
    """Attributes of the root group only (the datasets are not touched)."""
    with open(path, "rb") as f:
        buf = f.read()
    r = _Reader(buf)
    if buf[:8] != SIG:
        raise ValueError("not an HDF5 file")
    p = 24 + (4 if buf[8] == 1 else 0)
    return dict(_read_attr(b) for t, _, b in _messages(r, r.u(p + 32 + 8, 8)) if t == 0x0C)


def is_hdf5(path):
    try:
        with open(path, "rb") as f:
            return f.read(8) == SIG
    except OSError:
        return False
