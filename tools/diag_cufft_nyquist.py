"""Does cuFFT's batched 1-D C2R ignore the imaginary part of the Nyquist (and DC) bin?  The reference's half-cell shift of ux
(computeVelocityShiftInX + C2R, KSpaceSolver/SolverCudaKernels.cu:2617, MatrixClasses/CufftComplexMatrix.cpp:619) feeds it a Nyquist bin
multiplied by x_shift_neg_r[Nx/2] = i, i.e. purely imaginary.  torch.fft.irfft on CUDA calls the same cuFFT C2R."""
import torch
torch.manual_seed(0)
for n in (32, 64, 96, 128, 256, 512, 1024):
    for batch in (64, 4096, 16384):
        x = torch.randn(batch, n, device="cuda")
        X = torch.fft.rfft(x)
        Y = X.clone()
        Y[:, n // 2] = Y[:, n // 2] * 1j        # purely imaginary Nyquist bin
        y = torch.fft.irfft(Y, n=n)
        Z = X.clone(); Z[:, n // 2] = 0          # what "imaginary part ignored" means here: the Nyquist mode vanishes
        z = torch.fft.irfft(Z, n=n)
        print(f"N={n:5d} batch={batch:6d}: |irfft(i*Nyq) - irfft(Nyq dropped)| max {float((y - z).abs().max()):.3e}   (Nyquist mode amplitude {float((X[:, n//2].abs()/n).max()):.3e})")
