"""Reader / writer of the KWH5 container used by minih5 (k-wave-fluid-cuda_b200/csrc/minih5/minih5.cpp) and helpers
that lay a synthetic case out with the k-Wave input-file schema (main.cpp:446-563 of the reference) and read an output
file back.  Test / oracle tooling."""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"KWH5\x00\x01\x00\x00"


def _wstr(f, s):
    b = s.encode()
    f.write(struct.pack("<I", len(b)))
    f.write(b)


def _rstr(f):
    (n,) = struct.unpack("<I", f.read(4))
    return f.read(n).decode()


def write_file(path, objects):
    """objects: ordered dict  abs_path -> dict(kind='group'|'f32'|'u64', attrs={name: str|int|float}, data=ndarray)"""
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(objects)))
        for p, o in objects.items():
            _wstr(f, p)
            kind = {"group": 0, "f32": 1, "u64": 2}[o["kind"]]
            f.write(struct.pack("<B", kind))
            attrs = o.get("attrs", {})
            f.write(struct.pack("<I", len(attrs)))
            for k, v in attrs.items():
                _wstr(f, k)
                if isinstance(v, str):
                    f.write(struct.pack("<B", 0))
                    _wstr(f, v)
                elif isinstance(v, (int, np.integer)):
                    f.write(struct.pack("<Bq", 1, int(v)))
                else:
                    f.write(struct.pack("<Bf", 2, float(v)))
            if kind:
                a = np.ascontiguousarray(o["data"], dtype=np.float32 if kind == 1 else np.uint64)
                f.write(struct.pack("<I", a.ndim))
                f.write(struct.pack("<%dQ" % a.ndim, *a.shape))
                chunk = o.get("chunk", ())
                f.write(struct.pack("<I", len(chunk)))
                if chunk:
                    f.write(struct.pack("<%dQ" % len(chunk), *chunk))
                f.write(struct.pack("<I", int(o.get("deflate", 0))))
                f.write(a.tobytes())


def read_file(path):
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == MAGIC, "not a KWH5 file"
        (n,) = struct.unpack("<Q", f.read(8))
        for _ in range(n):
            p = _rstr(f)
            (kind,) = struct.unpack("<B", f.read(1))
            (na,) = struct.unpack("<I", f.read(4))
            attrs = {}
            for _ in range(na):
                k = _rstr(f)
                (t,) = struct.unpack("<B", f.read(1))
                attrs[k] = _rstr(f) if t == 0 else struct.unpack("<q", f.read(8))[0] if t == 1 else struct.unpack("<f", f.read(4))[0]
            o = {"kind": ["group", "f32", "u64"][kind], "attrs": attrs}
            if kind:
                (rank,) = struct.unpack("<I", f.read(4))
                dims = struct.unpack("<%dQ" % rank, f.read(8 * rank))
                (cr,) = struct.unpack("<I", f.read(4))
                o["chunk"] = struct.unpack("<%dQ" % cr, f.read(8 * cr)) if cr else ()
                (o["deflate"],) = struct.unpack("<I", f.read(4))
                dt = np.float32 if kind == 1 else np.uint64
                cnt = int(np.prod(dims)) if rank else 1
                o["data"] = np.frombuffer(f.read(cnt * np.dtype(dt).itemsize), dtype=dt).reshape(dims)
            out[p] = o
    return out


def read_root_attrs(path):
    """Attributes of the root group only (first record), without touching the datasets."""
    with open(path, "rb") as f:
        assert f.read(8) == MAGIC, "not a KWH5 file"
        f.read(8)
        assert _rstr(f) == "/"
        f.read(1)
        (na,) = struct.unpack("<I", f.read(4))
        attrs = {}
        for _ in range(na):
            k = _rstr(f)
            (t,) = struct.unpack("<B", f.read(1))
            attrs[k] = _rstr(f) if t == 0 else struct.unpack("<q", f.read(8))[0] if t == 1 else struct.unpack("<f", f.read(4))[0]
        return attrs


# ---- k-Wave input file ------------------------------------------------------------------------------------------------
_U64_SCALARS = ("Nx Ny Nz Nt pml_x_size pml_y_size pml_z_size sensor_mask_type p_source_flag p0_source_flag transducer_source_flag "
                "ux_source_flag uy_source_flag uz_source_flag nonuniform_grid_flag absorbing_flag nonlinear_flag u_source_many "
                "u_source_mode p_source_many p_source_mode").split()  # fmt: skip
_F32_SCALARS = "dt dx dy dz c_ref pml_x_alpha pml_y_alpha pml_z_alpha alpha_power".split()


def input_objects(cfg, arrays):
    """Objects of a k-Wave input file (file format 1.1) for a synth.make_case() result."""
    nx, ny, nz = cfg["Nx"], cfg["Ny"], cfg["Nz"]
    objs = {"/": {"kind": "group", "attrs": {
        "created_by": "k-wave-fluid-cuda_b200 synth", "creation_date": "2026-10-18", "file_description": "synthetic input",
        "file_type": "input", "major_version": "1", "minor_version": "1"}}}  # fmt: skip

    def ds(name, data, kind, domain="real"):
        objs["/" + name] = {"kind": kind, "data": data,
                            "attrs": {"data_type": "float" if kind == "f32" else "long", "domain_type": domain}}  # fmt: skip

    for k in _U64_SCALARS:
        if k in cfg:
            if k in ("u_source_many", "u_source_mode") and not (cfg.get("ux_source_flag") or cfg.get("uy_source_flag") or cfg.get("uz_source_flag")):
                continue
            if k in ("p_source_many", "p_source_mode") and not cfg.get("p_source_flag"):
                continue
            ds(k, np.array(cfg[k], np.uint64).reshape(1, 1, 1), "u64")
    for k in _F32_SCALARS:
        if k in cfg and (k != "alpha_power" or cfg.get("absorbing_flag")):
            ds(k, np.array(cfg[k], np.float32).reshape(1, 1, 1), "f32")
    n = nx * ny * nz
    for k, v in arrays.items():
        a = np.asarray(v)
        if a.dtype.kind == "c":
            a = np.ascontiguousarray(a.astype(np.complex64)).view(np.float32)
            shape = (1, 1, a.size)
            if k.startswith(("ddy", "y_shift")):
                shape = (1, a.size // 2, 2)
            if k.startswith(("ddz", "z_shift")):
                shape = (a.size // 2, 1, 2)
            ds(k, a.reshape(shape), "f32", "complex")
        elif a.dtype.kind in "ui":
            if k == "sensor_mask_corners":
                ds(k, a.reshape(1, -1, 6), "u64")
            else:
                ds(k, a.reshape(1, 1, -1), "u64")
        else:
            if a.size == n:
                ds(k, a.reshape(nz, ny, nx), "f32")
            elif k in ("pml_y", "pml_y_sgy"):
                ds(k, a.reshape(1, -1, 1), "f32")
            elif k in ("pml_z", "pml_z_sgz"):
                ds(k, a.reshape(-1, 1, 1), "f32")
            elif k in ("p_source_input", "ux_source_input", "uy_source_input", "uz_source_input") :
                many = cfg.get("p_source_many" if k.startswith("p_") else "u_source_many", 0)
                nsrc = arrays["p_source_index" if k.startswith("p_") else "u_source_index"].size
                ds(k, a.reshape(1, -1, nsrc) if many else a.reshape(1, -1, 1), "f32")
            else:
                ds(k, a.reshape(1, 1, -1), "f32")
    return objs


def write_input(path, cfg, arrays):
    write_file(path, input_objects(cfg, arrays))


def read_output(path):
    """name -> ndarray (datasets) for every dataset of an output file, groups flattened to 'group/child'."""
    return {p.lstrip("/"): o["data"] for p, o in read_file(path).items() if o["kind"] != "group"}
