"""Parity of the hand-written 3-D R2C / C2R transforms (replacement of CufftComplexMatrix::compute{R2C,C2R}FftND,
MatrixClasses/CufftComplexMatrix.cpp:508-534) against the DFT definition evaluated in FP64 (NumPy pocketfft).
Tolerance: relative L2 <= 2e-6 per transform (FP32 butterflies; cuFFT itself sits at ~3e-7)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [  # (nz, ny, nx)
    (16, 16, 16), (16, 16, 32), (32, 16, 64), (64, 64, 64), (16, 32, 128), (32, 16, 256), (16, 16, 512),
    (16, 16, 1024), (16, 128, 16), (16, 256, 32), (16, 512, 16), (16, 1024, 16), (128, 16, 16), (256, 16, 32),
    (512, 16, 16), (1024, 16, 16), (128, 128, 128), (64, 256, 128),
]  # fmt: skip


def rel_l2(a, b):
    return np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())


@pytest.mark.parametrize("shape", SHAPES)
def test_r2c_matches_dft(kw, shape):
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    got = kw.fft_r2c_3d(x)
    want = np.fft.rfftn(x.astype(np.float64))
    assert got.shape == want.shape
    assert rel_l2(got, want) < 2e-6


@pytest.mark.parametrize("shape", SHAPES)
def test_c2r_matches_dft(kw, shape):
    rng = np.random.default_rng(1 + sum(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    xk = np.fft.rfftn(x.astype(np.float64)).astype(np.complex64)
    got = kw.fft_c2r_3d(xk, shape[2])
    want = np.fft.irfftn(xk.astype(np.complex128), s=shape) * x.size  # unnormalised, as cufftExecC2R
    assert rel_l2(got, want) < 2e-6


def test_linearity_and_roundtrip_at_size(kw):
    """Size-independent properties on a grid too large for the FP64 check to be cheap: C2R(R2C(x)) = N x."""
    shape = (256, 256, 256)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(shape).astype(np.float32)
    back = kw.fft_c2r_3d(kw.fft_r2c_3d(x), shape[2]) / x.size
    assert rel_l2(back, x) < 2e-6


def test_unsupported_length_fails_loudly(kw):
    """lengths outside both kernel families (tuned powers of two; 8 m with factors 2, 3, 5, 7 -- tests/test_generic_lengths_gpu.py)"""
    with pytest.raises(kw.KwError):
        kw.fft_r2c_3d(np.zeros((16, 16, 20), np.float32))
