"""ctypes binding of include/kwave_b200.h.  Mirrors the call order of KSpaceFirstOrderSolver
(allocateMemory -> loadInputData -> compute, KSpaceSolver/KSpaceFirstOrderSolver.cpp:124-439)."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "kwave_b200.h")


class KwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"kwave_b200 error {code}: {msg}")
        self.code = code


class KwConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("struct_size", C.c_uint32),
        ("nx", C.c_uint64), ("ny", C.c_uint64), ("nz", C.c_uint64), ("nt", C.c_uint64),
        ("dt", C.c_float), ("dx", C.c_float), ("dy", C.c_float), ("dz", C.c_float), ("c_ref", C.c_float),
        ("alpha_power", C.c_float),
        ("nonlinear_flag", C.c_int32), ("absorbing_flag", C.c_int32), ("nonuniform_grid_flag", C.c_int32),
        ("p_source_flag", C.c_uint64), ("ux_source_flag", C.c_uint64), ("uy_source_flag", C.c_uint64),
        ("uz_source_flag", C.c_uint64), ("transducer_source_flag", C.c_uint64),
        ("p0_source_flag", C.c_int32),
        ("p_source_mode", C.c_int32), ("p_source_many", C.c_int32), ("u_source_mode", C.c_int32), ("u_source_many", C.c_int32),
        ("sensor_mask_type", C.c_int32),
        ("sampling_start_index", C.c_uint64),
        ("c_period", C.c_float), ("c_mos", C.c_uint32), ("c_harmonics", C.c_uint32),
        ("c_no_overlap", C.c_int32), ("c_40bit", C.c_int32),
        ("device", C.c_int32),
        ("raw_rows_capacity", C.c_uint64),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("nccl_unique_id", C.c_void_p),
    ]  # fmt: skip


def _parse_enum(name):
    """Read an enum's identifiers from the header so the ids can never drift from the C side."""
    src = open(_HEADER).read()
    body = re.search(r"enum\s+" + name + r"\s*\{(.*?)\};", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out, val = {}, 0
    for tok in body.split(","):
        tok = tok.strip()
        if not tok:
            continue
        if "=" in tok:
            k, v = tok.split("=")
            val = int(v.strip(), 0)
            tok = k.strip()
        out[tok] = val
        val += 1
    return out


ARRAY_IDS = _parse_enum("kw_array")
STREAM_IDS = _parse_enum("kw_stream")
KW_ABI_VERSION = 1

# dataset name in the k-Wave input file (Utils/MatrixNames.h) -> array id
INPUT_NAMES = {
    "c0": "KW_C0", "rho0": "KW_RHO0", "rho0_sgx": "KW_RHO0_SGX", "rho0_sgy": "KW_RHO0_SGY", "rho0_sgz": "KW_RHO0_SGZ",
    "BonA": "KW_BONA", "alpha_coeff": "KW_ALPHA_COEFF",
    "ddx_k_shift_pos_r": "KW_DDX_K_SHIFT_POS_R", "ddy_k_shift_pos": "KW_DDY_K_SHIFT_POS", "ddz_k_shift_pos": "KW_DDZ_K_SHIFT_POS",
    "ddx_k_shift_neg_r": "KW_DDX_K_SHIFT_NEG_R", "ddy_k_shift_neg": "KW_DDY_K_SHIFT_NEG", "ddz_k_shift_neg": "KW_DDZ_K_SHIFT_NEG",
    "pml_x_sgx": "KW_PML_X_SGX", "pml_y_sgy": "KW_PML_Y_SGY", "pml_z_sgz": "KW_PML_Z_SGZ",
    "pml_x": "KW_PML_X", "pml_y": "KW_PML_Y", "pml_z": "KW_PML_Z",
    "sensor_mask_index": "KW_SENSOR_MASK_INDEX", "sensor_mask_corners": "KW_SENSOR_MASK_CORNERS",
    "p0_source_input": "KW_P0_SOURCE_INPUT", "p_source_input": "KW_P_SOURCE_INPUT", "p_source_index": "KW_P_SOURCE_INDEX",
    "u_source_index": "KW_U_SOURCE_INDEX", "ux_source_input": "KW_UX_SOURCE_INPUT", "uy_source_input": "KW_UY_SOURCE_INPUT",
    "uz_source_input": "KW_UZ_SOURCE_INPUT", "transducer_source_input": "KW_TRANSDUCER_SOURCE_INPUT",
    "delay_mask": "KW_DELAY_MASK",
    "x_shift_neg_r": "KW_X_SHIFT_NEG_R", "y_shift_neg_r": "KW_Y_SHIFT_NEG_R", "z_shift_neg_r": "KW_Z_SHIFT_NEG_R",
}  # fmt: skip

_lib = None


def library_path():
    # KWAVE_B200_LIB selects an alternative build of the same library (kernel-variant experiments)
    return os.environ.get("KWAVE_B200_LIB") or os.path.join(_HERE, "libkwave_b200.so")


def load_library():
    """Load the CUDA library.  Fails loudly when it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise KwError(-2, f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(path)
    u64, vp, i32 = C.c_uint64, C.c_void_p, C.c_int
    lib.kw_last_error.restype = C.c_char_p
    sigs = {
        "kw_abi_version": [],
        "kw_cuda_code_version": [C.POINTER(C.c_int)],
        "kw_ctx_create": [C.POINTER(KwConfig), C.POINTER(vp)],
        "kw_ctx_destroy": [vp],
        "kw_set_array": [vp, i32, vp, u64],
        "kw_get_array": [vp, i32, vp, u64],
        "kw_stream_enable": [vp, i32],
        "kw_preprocess": [vp],
        "kw_run": [vp, u64, C.POINTER(u64), i32],
        "kw_time_index": [vp, C.POINTER(u64)],
        "kw_synchronize": [vp],
        "kw_stream_info": [vp, i32, C.POINTER(u64), C.POINTER(u64)],
        "kw_stream_fetch": [vp, i32, vp, u64, C.POINTER(u64)],
        "kw_stream_peek": [vp, i32, u64, vp, u64],
        "kw_stream_async": [vp, i32],
        "kw_stream_pending": [vp, i32, C.POINTER(u64)],
        "kw_stream_buffer_get": [vp, i32, i32, vp, u64, C.POINTER(u64)],
        "kw_stream_buffer_set": [vp, i32, i32, vp, u64],
        "kw_finish": [vp],
        "kw_set_source_row": [vp, i32, u64, vp, u64],
        "kw_length_supported": [u64],
        "kw_fft_r2c_3d": [u64, u64, u64, vp, vp],
        "kw_fft_c2r_3d": [u64, u64, u64, vp, vp],
        "kw_fft_zmid": [u64, u64, u64, i32, vp, vp, C.c_float, vp, vp, vp, vp, vp, vp],
        "kw_last_run_ms": [vp, C.POINTER(C.c_float)],
        "kw_launch_count": [vp, C.POINTER(u64)],
        "kw_profile": [vp, i32, i32],
        "kw_profile_report": [vp, C.c_char_p, u64],
        "kw_nccl_unique_id": [vp, u64],
        "kw_local_slab": [vp, C.POINTER(u64), C.POINTER(u64)],
        "kw_sensor_layout": [vp, C.POINTER(u64), C.POINTER(u64), vp, u64],
        "kw_comm_bytes": [vp, C.POINTER(C.c_double)],
        "kw_comm_mode": [vp, C.POINTER(C.c_int)],
        "kw_set_time_index": [vp, u64],
        "kw_intensity_avg_block": [vp, vp, i32, u64, u64, vp],
        "kw_q_term": [vp, vp, i32, vp, u64],
        "kw_stream_state_size": [vp, i32, C.POINTER(u64)],
        "kw_stream_state_get": [vp, i32, vp, u64],
        "kw_stream_state_set": [vp, i32, vp, u64],
        "kw_device_memory": [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)],
        "kw_c40_encode": [vp, u64, i32, vp],
        "kw_c40_decode": [vp, u64, i32, vp],
        "kw_compression_bases": [vp, i32, vp, vp, u64, C.POINTER(u64), C.POINTER(u64)],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise KwError(rc, load_library().kw_last_error().decode())


def length_supported(n):
    """2: tuned kernels, 1: run-time-length kernels, 0: unsupported transform length (kw_length_supported)."""
    return int(load_library().kw_length_supported(int(n)))


def fft_r2c_3d(x):
    """CufftComplexMatrix::computeR2CFftND on a host array of shape (nz, ny, nx)."""
    lib = load_library()
    x = np.ascontiguousarray(x, dtype=np.float32)
    nz, ny, nx = x.shape
    out = np.empty((nz, ny, nx // 2 + 1), dtype=np.complex64)
    _check(lib.kw_fft_r2c_3d(nx, ny, nz, x.ctypes.data, out.ctypes.data))
    return out


def fft_c2r_3d(xk, nx):
    lib = load_library()
    xk = np.ascontiguousarray(xk, dtype=np.complex64)
    nz, ny, nxr = xk.shape
    assert nxr == nx // 2 + 1
    out = np.empty((nz, ny, nx), dtype=np.float32)
    _check(lib.kw_fft_c2r_3d(nx, ny, nz, xk.ctypes.data, out.ctypes.data))
    return out


def fft_zmid(x, axis, mul=None, scal=1.0, vec_x=None, vec_y=None, vec_z=None):
    """The fused z pass (k_zmid) on a host half spectrum of shape (nz, ny, nx/2+1); returns one array, or three for axis 3."""
    lib = load_library()
    x = np.ascontiguousarray(x, dtype=np.complex64)
    nz, ny, nxr = x.shape
    nx = 2 * (nxr - 1)
    keep = [x]

    def ptr(a, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    outs = [np.empty_like(x) for _ in range(3 if axis == 3 else 1)]
    op = [o.ctypes.data for o in outs] + [None] * (3 - len(outs))
    _check(lib.kw_fft_zmid(nx, ny, nz, axis, x.ctypes.data, ptr(mul, np.float32), float(scal), ptr(vec_x, np.complex64),
                           ptr(vec_y, np.complex64), ptr(vec_z, np.complex64), op[0], op[1], op[2]))
    return outs if axis == 3 else outs[0]


def c40_encode(values, max_exp):
    """CompressHelper::convertFloatCTo40b on the device: complex64 array -> uint8 array (..., 5)."""
    v = np.ascontiguousarray(values, dtype=np.complex64)
    out = np.empty(v.shape + (5,), dtype=np.uint8)
    _check(load_library().kw_c40_encode(v.ctypes.data, v.size, max_exp, out.ctypes.data))
    return out


def c40_decode(packed, max_exp):
    """CompressHelper::convert40bToFloatC on the device: uint8 array (..., 5) -> complex64 array."""
    b = np.ascontiguousarray(packed, dtype=np.uint8)
    out = np.empty(b.shape[:-1], dtype=np.complex64)
    _check(load_library().kw_c40_decode(b.ctypes.data, out.size, max_exp, out.ctypes.data))
    return out


def intensity_avg_block(p, u_list):
    """computeAverageIntensities on one block: p and every u of shape (steps, n) -> list of intensities of shape (n,)."""
    p = np.ascontiguousarray(p, dtype=np.float32)
    us = [np.ascontiguousarray(u, dtype=np.float32) for u in u_list]
    steps, n = p.shape
    outs = [np.zeros(n, np.float32) for _ in us]
    up = (C.c_void_p * len(us))(*[u.ctypes.data for u in us])
    op = (C.c_void_p * len(us))(*[o.ctypes.data for o in outs])
    _check(load_library().kw_intensity_avg_block(p.ctypes.data, up, len(us), n, steps, op))
    return outs


def nccl_unique_id():
    """128-byte ncclUniqueId for a slab-decomposed run: create on ONE rank, hand to all (e.g. dist.broadcast)."""
    buf = C.create_string_buffer(128)
    _check(load_library().kw_nccl_unique_id(buf, 128))
    return bytes(buf.raw)


class Simulation:
    """One simulation == one kw_ctx.  ``cfg``/``arrays`` use the input-file names (see synth.make_case).

    Slab-decomposed runs (``nranks > 1``, one process per GPU): ``arrays`` may hold either the complete grids (they are
    cut to this rank's z-slab here, see slab.slice_arrays) or slabs already; ``nccl_id`` is the shared ncclUniqueId."""

    def __init__(self, cfg, arrays, streams=(), start_index=0, raw_rows_capacity=0, device=-1, compression=None,
                 rank=0, nranks=1, nccl_id=None, async_output=False):
        self.lib = load_library()
        self.cfg = dict(cfg)
        kc = KwConfig()
        kc.abi_version, kc.struct_size = KW_ABI_VERSION, C.sizeof(KwConfig)
        kc.nx, kc.ny, kc.nz, kc.nt = cfg["Nx"], cfg["Ny"], cfg["Nz"], cfg["Nt"]
        for k in ("dt", "dx", "dy", "dz", "c_ref"):
            setattr(kc, k, float(cfg[k]))
        kc.alpha_power = float(cfg.get("alpha_power", 0.0))
        for k in ("nonlinear_flag", "absorbing_flag", "nonuniform_grid_flag", "p_source_flag", "ux_source_flag",
                  "uy_source_flag", "uz_source_flag", "transducer_source_flag", "p0_source_flag", "p_source_mode",
                  "p_source_many", "u_source_mode", "u_source_many", "sensor_mask_type"):  # fmt: skip
            setattr(kc, k, int(cfg.get(k, 0)))
        kc.sampling_start_index = start_index
        kc.device = device
        kc.raw_rows_capacity = raw_rows_capacity
        kc.rank, kc.nranks = rank, nranks
        self._nccl_id = C.create_string_buffer(bytes(nccl_id), 128) if nranks > 1 else None
        kc.nccl_unique_id = C.cast(self._nccl_id, C.c_void_p) if nranks > 1 else None
        self.rank, self.nranks = rank, nranks
        self._harmonics = (compression or {}).get("harmonics", 1)
        if compression:
            kc.c_period, kc.c_mos, kc.c_harmonics = compression["period"], compression.get("mos", 1), compression.get("harmonics", 1)
            kc.c_no_overlap, kc.c_40bit = int(compression.get("no_overlap", 0)), int(compression.get("c40", 0))
        self.ctx = C.c_void_p()
        _check(self.lib.kw_ctx_create(C.byref(kc), C.byref(self.ctx)))
        self.n = cfg["Nx"] * cfg["Ny"] * cfg["Nz"]
        self.shape = (cfg["Nz"] // nranks, cfg["Ny"], cfg["Nx"])  # this rank's slab of every full-grid array
        if nranks > 1:
            from . import slab

            arrays = slab.slice_arrays(cfg, arrays, rank, nranks)
        for name, arr in arrays.items():
            self.set_array(name, arr)
        self.streams = []
        for s in streams:
            self.enable(s)
        if async_output:  # double-buffered series: full row buffers travel to pinned host memory while the loop goes on
            _check(self.lib.kw_stream_async(self.ctx, 1))
        _check(self.lib.kw_preprocess(self.ctx))

    # -- arrays ----------------------------------------------------------------------------------------------------
    def set_array(self, name, arr):
        aid = ARRAY_IDS[INPUT_NAMES.get(name, name)]
        a = np.asarray(arr)
        if a.dtype.kind in "ui":
            a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
            count = a.size
        elif a.dtype.kind == "c":
            a = np.ascontiguousarray(a, dtype=np.complex64).reshape(-1)
            count = a.size
        else:
            a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
            count = a.size
        _check(self.lib.kw_set_array(self.ctx, aid, a.ctypes.data, count))

    def get_array(self, name, shape=None):
        aid = ARRAY_IDS[INPUT_NAMES.get(name, name)]
        shape = self.shape if shape is None else shape
        out = np.empty(shape, dtype=np.float32)
        _check(self.lib.kw_get_array(self.ctx, aid, out.ctypes.data, out.size))
        return out

    # -- streams ---------------------------------------------------------------------------------------------------
    def enable(self, stream):
        sid = STREAM_IDS[stream]
        _check(self.lib.kw_stream_enable(self.ctx, sid))
        self.streams.append(stream)

    def fetch(self, stream):
        sid = STREAM_IDS[stream]
        row, rows = C.c_uint64(), C.c_uint64()
        _check(self.lib.kw_stream_info(self.ctx, sid, C.byref(row), C.byref(rows)))
        out = np.empty((rows.value, row.value), dtype=np.float32)
        got = C.c_uint64()
        _check(self.lib.kw_stream_fetch(self.ctx, sid, out.ctypes.data, out.size, C.byref(got)))
        return out[: got.value]

    def pending(self, stream):
        n = C.c_uint64()
        _check(self.lib.kw_stream_pending(self.ctx, STREAM_IDS[stream], C.byref(n)))
        return n.value

    # -- time loop -------------------------------------------------------------------------------------------------
    def run(self, nsteps, sync=True):
        done = C.c_uint64()
        _check(self.lib.kw_run(self.ctx, nsteps, C.byref(done), int(sync)))
        return done.value

    def synchronize(self):
        _check(self.lib.kw_synchronize(self.ctx))

    def finish(self):
        _check(self.lib.kw_finish(self.ctx))

    def last_run_ms(self):
        ms = C.c_float()
        _check(self.lib.kw_last_run_ms(self.ctx, C.byref(ms)))
        return ms.value

    def profile(self, enable=True, reset=True):
        _check(self.lib.kw_profile(self.ctx, int(enable), int(reset)))

    def profile_report(self):
        import json

        buf = C.create_string_buffer(1 << 16)
        _check(self.lib.kw_profile_report(self.ctx, buf, len(buf)))
        return json.loads(buf.value.decode())

    def launch_count(self):
        n = C.c_uint64()
        _check(self.lib.kw_launch_count(self.ctx, C.byref(n)))
        return n.value

    def local_slab(self):
        z0, nz = C.c_uint64(), C.c_uint64()
        _check(self.lib.kw_local_slab(self.ctx, C.byref(z0), C.byref(nz)))
        return z0.value, nz.value

    def sensor_layout(self):
        """(total points of the undecomposed row, positions of this rank's points in it)."""
        total, local = C.c_uint64(), C.c_uint64()
        _check(self.lib.kw_sensor_layout(self.ctx, C.byref(total), C.byref(local), None, 0))
        pos = np.empty(local.value, dtype=np.uint64)
        if local.value:
            _check(self.lib.kw_sensor_layout(self.ctx, C.byref(total), C.byref(local), pos.ctypes.data, pos.size))
        return total.value, pos

    # -- checkpoint / restart -----------------------------------------------------------------------------------------
    def save_state(self):
        """Everything a restart needs (cpp:1176-1224): t_index, the seven state arrays, the state of every stream."""
        st = {"t_index": self.t_index, "arrays": {n: self.get_array(n) for n in ("KW_P", "KW_RHOX", "KW_RHOY", "KW_RHOZ", "KW_UX_SGX", "KW_UY_SGY", "KW_UZ_SGZ")},
              "streams": {}}
        for name, sid in STREAM_IDS.items():
            if name == "KW_STREAM_COUNT":
                continue
            n = C.c_uint64()
            _check(self.lib.kw_stream_state_size(self.ctx, sid, C.byref(n)))
            if n.value:
                buf = np.empty(n.value, dtype=np.uint8)
                _check(self.lib.kw_stream_state_get(self.ctx, sid, buf.ctypes.data, buf.size))
                st["streams"][sid] = buf
        return st

    def load_state(self, st):
        for n, a in st["arrays"].items():
            self.set_array(n, a)
        _check(self.lib.kw_set_time_index(self.ctx, st["t_index"]))
        for sid, buf in st["streams"].items():
            _check(self.lib.kw_stream_state_set(self.ctx, sid, buf.ctypes.data, buf.size))

    def q_term(self, intensities):
        """computeQTerm on per-sensor intensities (mask order)."""
        ins = [np.ascontiguousarray(i, dtype=np.float32) for i in intensities]
        out = np.zeros(ins[0].size, np.float32)
        ip = (C.c_void_p * len(ins))(*[i.ctypes.data for i in ins])
        _check(self.lib.kw_q_term(self.ctx, ip, len(ins), out.ctypes.data, out.size))
        return out

    def compression_bases(self, shifted=False):
        """(oSize, bSize, bE, bE_1) as generated by the context (complex64, shape (harmonics, bSize))."""
        osz, bsz = C.c_uint64(), C.c_uint64()
        _check(self.lib.kw_compression_bases(self.ctx, int(shifted), None, None, 0, C.byref(osz), C.byref(bsz)))
        h = int(self._harmonics)
        be = np.empty((h, bsz.value), np.complex64)
        be1 = np.empty((h, bsz.value), np.complex64)
        _check(self.lib.kw_compression_bases(self.ctx, int(shifted), be.ctypes.data, be1.ctypes.data, be.size, C.byref(osz), C.byref(bsz)))
        return osz.value, bsz.value, be, be1

    def comm_mode(self):
        m = C.c_int()
        _check(self.lib.kw_comm_mode(self.ctx, C.byref(m)))
        return {0: "none", 1: "nccl", 2: "peer"}[m.value]

    def comm_bytes(self):
        b = C.c_double()
        _check(self.lib.kw_comm_bytes(self.ctx, C.byref(b)))
        return b.value

    @property
    def t_index(self):
        t = C.c_uint64()
        _check(self.lib.kw_time_index(self.ctx, C.byref(t)))
        return t.value

    def close(self):
        if self.ctx:
            self.lib.kw_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
