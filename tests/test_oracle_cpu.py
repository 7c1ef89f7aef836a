"""CPU-side checks (run with -m "not gpu"): the oracle against the reference's own golden vectors, host logic,
and the C ABI surface of the CUDA library (symbols only; no compute without a GPU)."""
import ctypes
import json
import os
import re
import sys

import numpy as np
import pytest

import fixtures
from oracle import compress_oracle as co
from oracle import kspace_oracle as ko

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tests"))


def f32(u):
    return np.array(u, dtype=np.uint32).view(np.float32)


# ---- compression: pinned against the reference's CompressHelper.cpp (oracle/_ref/compress_ref) ---------------------
@pytest.fixture(scope="module")
def compress_gold():
    return json.load(open(os.path.join(GOLD, "compress_ref.json")))


def test_compression_bases_match_reference(compress_gold):
    for g in compress_gold["bases"]:
        period = float(f32(g["period_bits"]))
        for shifted, sfx in ((False, ""), (True, "_shifted")):
            osize, bsize, be, be1 = co.generate_bases(period, g["mos"], g["harmonics"], True, shifted)
            assert (osize, bsize) == (g["oSize"], g["bSize"])
            for got, key in ((be, "bE" + sfx), (be1, "bE_1" + sfx)):
                ref = np.array(g[key], dtype=np.uint32).view(np.float32).reshape(-1, 2)
                ref = (ref[:, 0] + 1j * ref[:, 1]).reshape(got.shape)
                # libm cosf/sinf vs NumPy's: 1-2 ulp on values of magnitude <= 2/oSize
                assert np.abs(got - ref).max() <= 4e-7 * (2.0 / osize) + 1e-12, key


def test_40bit_codec_bit_exact(compress_gold):
    for c in compress_gold["codec"]:
        re_, im_ = f32(c["re"]), f32(c["im"])
        enc = co.encode40(re_, im_, c["e"])
        assert list(enc) == c["bytes"], (c, list(enc))
        dre, dim = co.decode40(bytes(c["bytes"]), c["e"])
        assert int(np.array(dre).view(np.uint32)) == c["dre"] and int(np.array(dim).view(np.uint32)) == c["dim"], c


def test_known_answers_from_survey():
    """values obtained from the reference's CompressHelper in the survey session."""
    osize, bsize, be, be1 = co.generate_bases(50.0, 1, 2, True, False)
    assert (osize, bsize) == (50, 101)
    assert abs(be[0, 1] - (0.000793691725 - 0.000100266589j)) < 1e-10
    assert abs(be1[0, 1] - (0.0388908945 - 0.00491307164j)) < 1e-8
    assert abs(be[1, 3] - (0.00174952438 - 0.00164291321j)) < 1e-9
    enc = co.encode40(np.float32(12345.678), np.float32(-0.5), 138)
    assert enc.hex() == "62cd810400"
    assert co.decode40(enc, 138) == (np.float32(12345.625), np.float32(-0.5))


def test_compressed_stream_state_machine():
    """Frames come out every oSize sampled steps, alternating accumulators; a pure tone at the basis frequency is
    recovered as amplitude * e^{i phase} (normalised bases)."""
    period, nsens, nt = 20.0, 3, 200
    s = co.CompressedStream(nsens, period, mos=1, harmonics=2, dtype=np.complex128)
    amp = np.array([1.0, 2.5, -0.75])
    frames = []
    for t in range(nt):
        out = s.feed(amp * np.cos(2 * np.pi * t / period))
        if out is not None:
            frames.append((t, out))
    assert [t for t, _ in frames] == list(range(19, nt, 20))
    for t, f in frames[2:]:
        np.testing.assert_allclose(np.abs(f[:, 0]), np.abs(amp), rtol=2e-2)
        assert np.abs(f[:, 1]).max() < 5e-2 * np.abs(amp).max()


# ---- solver oracle --------------------------------------------------------------------------------------------------
def test_oracle_fp32_vs_fp64_noise_floor(synth):
    """The FP32 evaluation of the same restatement stays within 2e-6 of FP64 over 60 steps: the floor against which the
    1e-5 parity bar is read."""
    cfg, arrays = synth.make_case(32, nt=60, source="p_plane")
    a = ko.run(cfg, arrays, dtype=np.float64, record=("p_raw",))["p"]
    b = ko.run(cfg, arrays, dtype=np.float32, record=("p_raw",))["p"]
    assert np.isfinite(a).all()
    assert np.linalg.norm(a - b) / np.linalg.norm(a) < 2e-6


def test_oracle_k_operators(synth):
    cfg, _ = synth.make_case(16, nt=1)
    kappa, n1, n2 = ko.generate_kappa_and_nablas(cfg)
    assert kappa.shape == (16, 16, 9) and kappa[0, 0, 0] == 1.0 and n1[0, 0, 0] == 0.0 and n2[0, 0, 0] == 0.0
    k2 = ko.generate_kappa(cfg)
    assert np.abs(k2 - kappa).max() < 1e-6  # two code paths of the reference (cpp:2440 vs :2556) agree to rounding
    sk = ko.generate_source_kappa(cfg)
    assert sk[0, 0, 0] == 1.0 and (np.abs(sk) <= 1).all()


def test_cuboid_ordering():
    """x fastest inside a cuboid, cuboids concatenated (OutputStreamsCudaKernels.cu:164-188)."""
    o = ko.KSpaceOracle.__new__(ko.KSpaceOracle)
    o.nx, o.ny, o.nz = 8, 4, 4
    o.sensor_corners = np.array([[1, 0, 0, 2, 1, 0], [7, 3, 3, 7, 3, 3]])
    idx = o.cuboid_indices()
    assert idx[0].tolist() == [1, 2, 9, 10] and idx[1].tolist() == [(3 * 4 + 3) * 8 + 7]


@pytest.mark.parametrize("name", fixtures.fixture_names())
def test_oracle_matches_reference_run(name, synth):
    """PIN: outputs of the reference's own solver (its unmodified sources + cuFFT, run on a B200 by
    oracle/make_ref_goldens.py) against the FP64 oracle on the same synthetic input.  Bar: rel-L2 <= 1e-5."""
    shape, kwargs, nt, flags, data = fixtures.load_fixture(name)
    if "--p_c" in flags:
        pytest.skip("compressed streams are checked in test_compression_matches_reference_run")
    cfg, arrays = synth.make_case(*shape, nt=nt, **kwargs)
    out = ko.run(cfg, arrays, nt=nt, dtype=np.float64,
                 record=("p_raw", "p_max", "p_rms", "u_raw", "p_final", "p_max_all", "p_min_all", "u_max"))
    checked = 0
    for key, (got, ref) in fixtures.time_series_views(out["p"], data, nt).items():
        err = fixtures.rel_l2(got, ref)
        print(f"{name}: {key}: rel-L2 {err:.3e}, max-abs {np.abs(got - ref).max():.3e} (scale {np.abs(ref).max():.3e})")
        assert err <= 1e-5, (name, key, err)
        checked += 1
    for key in ("p_max", "p_rms", "p_final", "p_max_all", "p_min_all", "ux", "ux_max"):
        if key in data and key in out:
            err = fixtures.rel_l2(out[key], data[key])
            assert err <= 1e-5, (name, key, err)
            checked += 1
    assert checked


def test_compression_matches_reference_run(synth):
    """PIN of the compression state machine: p_c frames and Ix_avg_c written by the reference (host OpenMP loop,
    IndexOutputStream.cpp:373-470, :299-342) vs CompressedStream fed with the oracle's sampled series."""
    shape, kwargs, nt, flags, data = fixtures.load_fixture("compressed_p_and_intensity")
    cfg, arrays = synth.make_case(*shape, nt=nt, **kwargs)
    out = ko.run(cfg, arrays, nt=nt, dtype=np.float64, record=("p_raw", "u_non_staggered_raw"))
    period, harm = 20.0, 2
    nsens = out["p"].shape[1]
    # the raw non-staggered velocity first (pins computeShiftedVelocity, cpp:2714-2735)
    for a in "xyz":
        ref = data[f"u{a}_non_staggered"].reshape(nt, nsens)
        if a == "x":
            assert fixtures.rel_l2(out[f"u{a}_non_staggered"], ref) <= 1e-5
    sp = co.CompressedStream(nsens, period, 1, harm, shifted=False, nsteps_total=nt, dtype=np.complex128)
    su = co.CompressedStream(nsens, period, 1, harm, shifted=True, nsteps_total=nt, dtype=np.complex128)
    pf, uf, inten = [], [], np.zeros(nsens)
    for t in range(nt):
        a, b = sp.feed(out["p"][t]), su.feed(out["ux_non_staggered"][t])
        if a is not None:
            pf.append(a), uf.append(b)
            inten += co.intensity_frame(a, b)
    ref_pc = data["p_c"].reshape(-1, nsens, harm, 2)
    ref_pc = ref_pc[..., 0] + 1j * ref_pc[..., 1]
    assert ref_pc.shape[0] == len(pf)
    assert fixtures.rel_l2(np.abs(np.stack(pf)), np.abs(ref_pc)) <= 1e-5
    err = np.linalg.norm(np.stack(pf) - ref_pc) / np.linalg.norm(ref_pc)
    assert err <= 1e-5, err
    ref_uc = data["ux_non_staggered_c"].reshape(-1, nsens, harm, 2)
    ref_uc = ref_uc[..., 0] + 1j * ref_uc[..., 1]
    assert np.linalg.norm(np.stack(uf) - ref_uc) / np.linalg.norm(ref_uc) <= 1e-5
    ix = inten / len(pf)  # postProcess: divided by the number of frames (IndexOutputStream.cpp:477-520)
    assert fixtures.rel_l2(ix, data["Ix_avg_c"].reshape(-1)) <= 1e-5


def test_half_step_shift_kernel_equals_the_spectral_shift():
    """kw_intensity_avg_block applies the half-step temporal shift of computeAverageIntensities (cpp:1257-1265, :1433-1470) as a
    circular convolution; the kernel it builds must reproduce the R2C * exp(i pi shift / n) * C2R definition for odd and even n."""
    rng = np.random.default_rng(3)
    for n in (5, 8, 77, 120):
        p, u = rng.standard_normal((n, 6)), rng.standard_normal((n, 6))
        h = co.half_step_kernel(n)
        idx = (np.arange(n)[:, None] - np.arange(n)[None, :]) % n
        assert np.abs((p * (h[idx] @ u)).sum(0) / n - co.intensity_avg(p, u)).max() < 1e-12


def test_q_term_of_a_plane_wave_intensity():
    """computeQTerm restatement: for I_x = sin(2 pi m x / Nx) on the whole grid, Q = -dIx/dx analytically."""
    nx = ny = nz = 16
    cfg = dict(Nx=nx, Ny=ny, Nz=nz, dx=1e-3, dy=1e-3, dz=1e-3)
    x = np.arange(nx)
    ix = np.broadcast_to(np.sin(2 * np.pi * 3 * x / nx), (nz, ny, nx)).reshape(-1)
    idx = np.arange(nx * ny * nz)
    q = co.q_term(cfg, [ix, np.zeros_like(ix), np.zeros_like(ix)], idx)
    want = -np.broadcast_to(2 * np.pi * 3 / (nx * 1e-3) * np.cos(2 * np.pi * 3 * x / nx), (nz, ny, nx)).reshape(-1)
    assert np.abs(q - want).max() <= 1e-6 * np.abs(want).max()


# ---- C ABI surface --------------------------------------------------------------------------------------------------
def declared_symbols():
    src = open(os.path.join(ROOT, "include", "kwave_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kw_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(kw):
    lib = ctypes.CDLL(kw.library_path())
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/kwave_b200.h but not exported"
    assert kw.load_library().kw_abi_version() == 1


def test_supported_transform_lengths(kw):
    """host-side rule of the two kernel families (csrc/fft_inst.cu, csrc/fft_generic.cu); no device needed"""
    lib = kw.load_library()
    tuned = [16, 32, 64, 128, 256, 512, 1024]
    for n in tuned:
        assert lib.kw_length_supported(n) == 2, n
    for n in (24, 40, 48, 56, 72, 80, 96, 112, 120, 160, 192, 240, 320, 384, 480, 640, 768, 960, 1000, 1080, 1296, 1536, 2048):
        assert lib.kw_length_supported(n) == 1, n
    for n in (0, 8, 12, 20, 36, 88, 100, 104, 136, 150, 2056, 4096):  # too short, not 8 m, a prime factor above 7, too long
        assert lib.kw_length_supported(n) == 0, n


def test_enum_ids_follow_the_reference_order(kw):
    assert kw.STREAM_IDS["KW_S_P_RAW"] == 0 and kw.STREAM_IDS["KW_S_P_MAX_ALL"] == 5
    assert kw.STREAM_IDS["KW_S_UX_RAW"] == 7 and kw.STREAM_IDS["KW_S_Q_TERM_C"] == kw.STREAM_IDS["KW_STREAM_COUNT"] - 1
    assert kw.ARRAY_IDS["KW_KAPPA"] == 0 and kw.ARRAY_IDS["KW_P"] == 3


def test_no_cpu_fallback(kw):
    """Without a CUDA device every compute entry point fails loudly (KW_ERR_CUDA), it never computes on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(kw.KwError) as e:
        kw.fft_r2c_3d(np.zeros((16, 16, 16), np.float32))
    assert e.value.code == -2
    cfg, arrays = kw.synth.make_case(16, nt=2)
    with pytest.raises(kw.KwError) as e:
        kw.Simulation(cfg, arrays)
    assert e.value.code == -2


def test_host_file_layer_round_trip(tmp_path):
    """The C++ host's file layer (host/Hdf5Io.h over minih5) on the CPU: create / write rows / close / reopen read-write /
    continue / read hyperslabs -- what checkpoint-restart and the raw-series post-processing rely on; the result is also read
    with the Python reader of the same container."""
    import subprocess
    import sys

    pkg = os.path.join(ROOT, "k-wave-fluid-cuda_b200")
    exe, h5 = str(tmp_path / "hdf5io_roundtrip"), str(tmp_path / "f.h5")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O1", "-I", os.path.join(pkg, "host"), "-I", os.path.join(pkg, "csrc", "minih5"),
                    os.path.join(ROOT, "tests", "cpp", "hdf5io_roundtrip.cpp"), os.path.join(pkg, "csrc", "minih5", "minih5.cpp"), "-o", exe, "-lz"],
                   check=True)  # fmt: skip
    r = subprocess.run([exe, h5], capture_output=True, text=True)
    assert r.returncode == 0 and "round trip ok" in r.stdout, r.stderr
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import kwh5

    f = kwh5.read_file(h5)
    assert f["/p"]["data"].shape == (1, 6, 4) and f["/p"]["data"][0, 5, 3] == 53.0
    assert f["/p_max/1"]["data"].shape == (2, 3, 2) and f["/"]["attrs"]["file_type"] == "output"


def test_configuration_errors_are_reported_before_any_device_work(kw):
    """kw_ctx_create validates the configuration first (the reference's checks: Parameters.cpp:421-424 alpha_power, main.cpp:460
    non-uniform grids) and reports KW_ERR_INVALID with a message -- on any machine, GPU or not."""
    import ctypes as C

    lib = kw.load_library()

    def create(**over):
        kc = kw.capi.KwConfig()
        kc.abi_version, kc.struct_size = kw.capi.KW_ABI_VERSION, C.sizeof(kw.capi.KwConfig)
        kc.nx = kc.ny = kc.nz = 32
        kc.nt, kc.dt, kc.dx, kc.dy, kc.dz, kc.c_ref = 10, 1e-8, 1e-4, 1e-4, 1e-4, 1500.0
        kc.device, kc.rank, kc.nranks = -1, 0, 1
        for k, v in over.items():
            setattr(kc, k, v)
        ctx = C.c_void_p()
        rc = lib.kw_ctx_create(C.byref(kc), C.byref(ctx))
        msg = lib.kw_last_error().decode()
        if rc == 0:
            lib.kw_ctx_destroy(ctx)
        return rc, msg

    for over, needle in ((dict(abi_version=99), "ABI mismatch"), (dict(nonuniform_grid_flag=1), "nonuniform_grid_flag"),
                         (dict(absorbing_flag=1, alpha_power=1.0), "alpha_power == 1"), (dict(nz=0), "Nz must be at least 1"),
                         (dict(nranks=2), "ncclUniqueId"), (dict(nz=1, nranks=2, nccl_unique_id=1), "one GPU")):  # fmt: skip
        rc, msg = create(**over)
        assert rc == -1 and needle in msg, (over, rc, msg)
    rc, msg = create()  # a valid configuration: fails only for lack of a device here, succeeds on a GPU box
    assert rc in (0, -2), (rc, msg)
