"""Grids whose sizes are not powers of two (or exceed 1024): the run-time-length kernels of csrc/fft_generic.cu behind the same launch
table as the tuned ones.  The reference plans whatever Nx, Ny, Nz the input file holds with cuFFT
(MatrixClasses/CufftComplexMatrix.cpp:87-91, :508-534); here every length 8 m <= 2048 with prime factors 2, 3, 5, 7 is accepted, alone
or mixed with tuned power-of-two axes.  Checked like the tuned path: transforms and the fused z pass against the FP64 DFT (NumPy), time
loops against the FP64 oracle (rel-L2 <= 1e-5), whole runs against the reference binary in tests/test_baseline_configs_gpu.py."""
import numpy as np
import pytest

from oracle import kspace_oracle as ko

pytestmark = pytest.mark.gpu
TOL = 1e-5

FFT_SHAPES = [  # (nz, ny, nx): every radix (2, 3, 4, 5, 7) on every axis, generic next to tuned axes, lengths above 1024
    (24, 40, 48), (16, 16, 96), (16, 120, 16), (160, 16, 16), (16, 16, 480), (56, 16, 16), (16, 56, 32), (16, 16, 112),
    (16, 16, 1536), (16, 2048, 16), (1280, 16, 16), (16, 16, 2048), (72, 80, 96), (64, 96, 128), (16, 1000, 16), (192, 240, 64),
]  # fmt: skip


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.complex128 if np.iscomplexobj(a) else np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300))


@pytest.mark.parametrize("shape", FFT_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_r2c_and_c2r_match_dft(kw, shape):
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    want = np.fft.rfftn(x.astype(np.float64))
    got = kw.fft_r2c_3d(x)
    assert got.shape == want.shape
    e1 = rel_l2(got, want)
    xk = want.astype(np.complex64)
    back = kw.fft_c2r_3d(xk, shape[2])
    e2 = rel_l2(back, np.fft.irfftn(xk.astype(np.complex128), s=shape) * x.size)  # unnormalised, as cufftExecC2R
    print(f"{shape}: R2C rel-L2 {e1:.3e}, C2R rel-L2 {e2:.3e}")
    assert e1 < 2e-6 and e2 < 2e-6, (shape, e1, e2)


@pytest.mark.parametrize("axis", [-1, 0, 1, 2, 3])
@pytest.mark.parametrize("n", [24, 96, 120, 224, 480, 1536, 2048])
def test_fused_z_pass_matches_dft(kw, n, axis):
    ny, nx = 16, 40  # nx = 40: the row length is a generic one too (NXP = 32, two 16-wide or four 8-wide tiles per row)
    nxr = nx // 2 + 1
    rng = np.random.default_rng(1000 * n + axis + 11)
    x = (rng.standard_normal((n, ny, nxr)) + 1j * rng.standard_normal((n, ny, nxr))).astype(np.complex64)
    mul = rng.uniform(0.5, 1.5, size=(n, ny, nxr)).astype(np.float32)
    scal = np.float32(1.0 / n)
    cv = lambda m: (rng.standard_normal(m) + 1j * rng.standard_normal(m)).astype(np.complex64)  # noqa: E731
    vx, vy, vz = cv(nxr), cv(ny), cv(n)
    e = np.fft.fft(x.astype(np.complex128), axis=0) * (mul.astype(np.float64) * float(scal))
    ifz = lambda a: np.fft.ifft(a, axis=0) * n  # noqa: E731
    bx, by, bz = vx[None, None, :].astype(np.complex128), vy[None, :, None].astype(np.complex128), vz[:, None, None].astype(np.complex128)
    if axis == 3:
        got = kw.fft_zmid(x, 3, mul=mul, scal=scal, vec_x=vx, vec_y=vy, vec_z=vz)
        for g, b, name in zip(got, (bx, by, bz), "xyz"):
            err = rel_l2(g, ifz(e * b))
            assert err < 2e-6, (n, name, err)
        return
    want = ifz(e if axis < 0 else e * (bx, by, bz)[axis])
    got = kw.fft_zmid(x, axis, mul=mul, scal=scal, vec_x=vx if axis == 0 else None, vec_y=vy if axis == 1 else None,
                      vec_z=vz if axis == 2 else None)
    err = rel_l2(got, want)
    print(f"generic z pass n={n} axis={axis}: rel-L2 {err:.3e}")
    assert err < 2e-6, (n, axis, err)


LOOP_SHAPES = [(96, 40, 24), (40, 120, 24), (24, 40, 96), (48, 48, 48), (64, 96, 32), (160, 32, 32), (32, 32, 112), (80, 64, 1)]


@pytest.mark.parametrize("shape", LOOP_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_time_loop_matches_oracle(kw, synth, shape):
    """nonlinear + absorbing + heterogeneous step with a plane pressure source: every kernel of the step on generic lengths
    (the last shape is a 2-D grid)."""
    nx, ny, nz = shape
    nt = 24
    cfg, arrays = synth.make_case(nx, ny, nz, nt=nt, nonlinear=True, absorbing=True, source="p_plane", n_sensor=nx * ny, shuffle_sensor=True)
    ref = ko.run(cfg, arrays, nt=nt, dtype=np.float64, record=("p_raw", "p_max", "u_raw", "p_final"))
    sim = kw.Simulation(cfg, arrays, streams=["KW_S_P_RAW", "KW_S_P_MAX", "KW_S_UX_RAW"], raw_rows_capacity=nt)
    assert sim.run(nt) == nt
    sim.finish()
    got = {s: sim.fetch(s) for s in ("KW_S_P_RAW", "KW_S_P_MAX", "KW_S_UX_RAW")}
    p_final = sim.get_array("KW_P")
    sim.close()
    for a, b, what in ((got["KW_S_P_RAW"], ref["p"], "p raw"), (got["KW_S_UX_RAW"], ref["ux"], "ux raw"),
                       (got["KW_S_P_MAX"][0], ref["p_max"], "p max"), (p_final, ref["p_final"], "p final")):
        err = rel_l2(a, b)
        print(f"{nx}x{ny}x{nz}: {what}: rel-L2 {err:.3e}, max-abs {np.abs(a - b).max():.3e} (scale {np.abs(b).max():.3e})")
        assert err <= TOL, (shape, what, err)


@pytest.mark.parametrize("shape", [(120, 48, 40), (40, 48, 120)], ids=lambda s: "x".join(map(str, s)))
def test_p0_lossless_with_aggregates_matches_oracle(kw, synth, shape):
    """linear lossless p0 problem: gradient z pass + initial-velocity epilogue, lossless density epilogue with fused whole-domain sampling"""
    nx, ny, nz = shape
    nt = 20
    cfg, arrays = synth.make_case(nx, ny, nz, nt=nt, nonlinear=False, absorbing=False, source="p0", sensor="full_cuboid")
    ref = ko.run(cfg, arrays, nt=nt, dtype=np.float64, record=("p_max_all", "p_rms", "p_final"))
    sim = kw.Simulation(cfg, arrays, streams=["KW_S_P_MAX_ALL", "KW_S_P_RMS"])
    assert sim.run(nt) == nt
    sim.finish()
    for sid, key in (("KW_S_P_MAX_ALL", "p_max_all"), ("KW_S_P_RMS", "p_rms")):
        err = rel_l2(sim.fetch(sid)[0], ref[key])
        assert err <= TOL, (shape, key, err)
    assert rel_l2(sim.get_array("KW_P"), ref["p_final"]) <= TOL
    sim.close()


@pytest.mark.parametrize("shape", [(16, 16, 20), (16, 16, 88), (16, 16, 2056), (16, 104, 16)], ids=lambda s: "x".join(map(str, s)))
def test_unsupported_length_fails_loudly(kw, shape):
    """not a multiple of 8, a prime factor above 7 (88 = 8 * 11, 104 = 8 * 13), or longer than 2048"""
    with pytest.raises(kw.KwError):
        kw.fft_r2c_3d(np.zeros(shape, np.float32))
