"""Reader / writer of the HDF5 files the host and the reference build exchange (through tools/h5lite.py, the Python counterpart of
k-wave-fluid-cuda_b200/csrc/minih5/minih5.cpp; legacy KWH5 containers of round 1 are still read) and helpers
that lay a synthetic case out with the k-Wave input-file schema (main.cpp:446-563 of the reference) and read an output
file back.  Test / oracle tooling."""
from __future__ import annotations

import os
import struct
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

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
    """objects: ordered dict  abs_path -> dict(kind='group'|'f32'|'u64', attrs={name: str|int|float}, data=ndarray[, chunk=(), deflate=level])
    Written as a real HDF5 file (tools/h5lite.py); datasets without `chunk` are contiguous, `deflate` adds the zlib filter."""
    import h5lite

    h5lite.write_hdf5(path, objects)


def _read_kwh5(path):
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


def read_file(path):
    """Every object of an HDF5 file (or of a legacy KWH5 container of round 1): abs_path -> dict(kind, attrs, data, chunk, deflate)."""
    import h5lite

    if h5lite.is_hdf5(path):
        return h5lite.read_hdf5(path)
    return _read_kwh5(path)


def read_root_attrs(path):
    """Attributes of the root group only, without touching the datasets."""
    import h5lite

    return h5lite.read_root_attrs(path)


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
