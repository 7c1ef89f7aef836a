// Per-length launch table: every FFT-bearing kernel is instantiated once per transform length N in its own translation
// unit (fft_inst.cu compiled with -DKW_N=<N>) so the build parallelises; the solver picks the tables of Nx, Ny, Nz.
#pragma once
#include <cuda_runtime.h>
#include "fft_kernels.cuh"
#include "solver_kernels.cuh"

namespace kw {

struct FftOps {
  int n;
  int col_w;  // kx values per column tile of this length (ColCfg<N>::W)
  int col_wk;  // workers per column transform (ColCfg<N>::WK); the y-blocked layout needs Ny / nranks >= col_wk
  int zmid_w;  // kx values per tile of the fused z pass (ZCfg<N>::W)
  void (*xfwd)(const XFwdArgs&, int nfields, cudaStream_t);
  void (*xinv_store)(const XInvArgs<1>&, const EpiStore&, int nfields, cudaStream_t);
  void (*xinv_add)(const XInvArgs<1>&, const EpiAdd&, cudaStream_t);
  void (*xinv_velocity)(const XInvArgs<1>&, const EpiVelocity&, int nfields, cudaStream_t);
  void (*xinv_density)(const XInvArgs<3>&, const EpiDensity&, cudaStream_t);
  void (*xinv_psum)(const XInvArgs<2>&, const EpiPressureSum&, cudaStream_t);
  void (*col)(const ColArgs&, int dir, int nfields, cudaStream_t);
  void (*zmid)(const ZMidArgs&, cudaStream_t);  // one field per launch
};

const FftOps* get_fft_ops(int n);  // tuned table (powers of two in [16, 1024]), else the run-time-length kernels; nullptr: unsupported
const FftOps* get_generic_fft_ops(int n);  // fft_generic.cu: N = 8 m <= 2048 with prime factors 2, 3, 5, 7
bool generic_length_supported(int n);
int sm_count();

#define KW_DECLARE_OPS(N) extern const FftOps fft_ops_##N;
KW_DECLARE_OPS(16)
KW_DECLARE_OPS(32)
KW_DECLARE_OPS(64)
KW_DECLARE_OPS(128)
KW_DECLARE_OPS(256)
KW_DECLARE_OPS(512)
KW_DECLARE_OPS(1024)

}  // namespace kw
