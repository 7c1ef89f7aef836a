/*
 * kwave_b200.h -- C ABI of the B200-native time-step engine behind kspaceFirstOrder-CUDA's solver.
 *
 * This is the drop-in boundary for the per-timestep hot path of klepo/k-Wave-Fluid-CUDA
 * (KSpaceSolver/KSpaceFirstOrderSolver.cpp:864-943, computeMainLoop).  The reference has no FFI; its host class
 * calls three C++ seams, all replaced by the entry points below (plain pointers and sizes, no C++/torch types):
 *
 *   SolverCudaKernels::*          (KSpaceSolver/SolverCudaKernels.cuh:77-499)      -> kw_run (fused per-step kernels)
 *   CufftComplexMatrix::*         (MatrixClasses/CufftComplexMatrix.h:73-238)      -> kw_run, kw_fft_r2c_3d, kw_fft_c2r_3d
 *   OutputStreamsCudaKernels::*   (OutputStreams/OutputStreamsCudaKernels.cuh:58-105) -> kw_stream_* (in-step sampling)
 *
 * Array ids follow MatrixContainer::MatrixIdx (Containers/MatrixContainer.h:63-207); stream ids follow
 * OutputStreamContainer::OutputStreamIdx (Containers/OutputStreamContainer.h:59-150), so a host built on the
 * reference's containers stays recognisable.  The library owns device memory; the caller owns host buffers.
 *
 * Every function returns KW_OK (0) or a negative error code; kw_last_error() returns the message of the last failure
 * on the calling thread (the reference's convention is: any error -> message -> exit(EXIT_FAILURE),
 * Logger/Logger.cpp:82-89; the C++ host maps non-zero codes to the same exceptions).
 * There is no CPU fallback: without a CUDA device every compute entry point fails with KW_ERR_CUDA.
 *
 * Call order:  kw_ctx_create -> kw_set_array (inputs as stored in the k-Wave input file: 1-based indices, c0 unsquared,
 * rho0_sg* undivided) -> kw_stream_enable -> kw_preprocess -> { kw_run, kw_stream_fetch }* -> kw_finish ->
 * kw_stream_fetch / kw_get_array -> kw_ctx_destroy.
 */
#ifndef KWAVE_B200_H
#define KWAVE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KW_ABI_VERSION 1

enum kw_status {
  KW_OK = 0,
  KW_ERR_INVALID = -1,     /* bad argument / unsupported configuration (std::invalid_argument in the reference) */
  KW_ERR_CUDA = -2,        /* CUDA runtime failure, or no device (std::runtime_error via cudaCheckErrors)      */
  KW_ERR_ALLOC = -3,       /* out of device or host memory (std::bad_alloc)                                    */
  KW_ERR_STATE = -4,       /* call order violated                                                              */
  KW_ERR_STREAM_FULL = -5, /* a raw time-series buffer must be fetched before more steps can run               */
  KW_ERR_COMM = -6         /* NCCL failure in a sharded run                                                    */
};

/* Source modes: Parameters::SourceMode (Parameters/Parameters.h) */
enum kw_source_mode { KW_SRC_DIRICHLET = 0, KW_SRC_ADDITIVE_NO_CORRECTION = 1, KW_SRC_ADDITIVE = 2 };

/* Scalars of the input file that the time loop reads (Parameters/Parameters.cpp:194-553; Appendix B of SURVEY.md). */
typedef struct kw_config {
  uint32_t abi_version; /* KW_ABI_VERSION */
  uint32_t struct_size; /* sizeof(kw_config) */
  uint64_t nx, ny, nz, nt;
  float dt, dx, dy, dz, c_ref;
  float alpha_power;
  int32_t nonlinear_flag, absorbing_flag, nonuniform_grid_flag;
  /* *_source_flag = number of time steps the source is defined for (active while flag > t), cpp:2258,2314 */
  uint64_t p_source_flag, ux_source_flag, uy_source_flag, uz_source_flag, transducer_source_flag;
  int32_t p0_source_flag;
  int32_t p_source_mode, p_source_many, u_source_mode, u_source_many;
  int32_t sensor_mask_type;        /* 0: index, 1: corners (Parameters.cpp:282-295) */
  uint64_t sampling_start_index;   /* 0-based: "-s" minus one (CommandLineParameters.cpp:424) */
  /* compression (Compression/CompressHelper.cpp:48-65); only read when a *_c stream is enabled */
  float c_period;
  uint32_t c_mos, c_harmonics;
  int32_t c_no_overlap, c_40bit;
  int32_t device;                  /* CUDA device ordinal; -1 = current device */
  uint64_t raw_rows_capacity;      /* rows a raw/compressed stream buffers on the device before KW_ERR_STREAM_FULL; 0 = auto */
  /* slab decomposition over GPUs (no reference counterpart) */
  int32_t rank, nranks;
  const void* nccl_unique_id;      /* 128-byte ncclUniqueId shared by all ranks, or NULL for nranks == 1 */
} kw_config;

/* Input / state arrays.  Names = dataset names of Utils/MatrixNames.h; order follows MatrixContainer::MatrixIdx. */
enum kw_array {
  KW_KAPPA = 0, KW_SOURCE_KAPPA, KW_C0 /* c0 in, c2 after kw_preprocess */, KW_P, KW_RHOX, KW_RHOY, KW_RHOZ,
  KW_UX_SGX, KW_UY_SGY, KW_UZ_SGZ, KW_DUXDX, KW_DUYDY, KW_DUZDZ, KW_RHO0,
  KW_RHO0_SGX /* rho0_sgx in, dt/rho0_sgx after kw_preprocess */, KW_RHO0_SGY, KW_RHO0_SGZ,
  KW_DDX_K_SHIFT_POS_R, KW_DDY_K_SHIFT_POS, KW_DDZ_K_SHIFT_POS,
  KW_DDX_K_SHIFT_NEG_R, KW_DDY_K_SHIFT_NEG, KW_DDZ_K_SHIFT_NEG,
  KW_PML_X_SGX, KW_PML_Y_SGY, KW_PML_Z_SGZ, KW_PML_X, KW_PML_Y, KW_PML_Z,
  KW_BONA, KW_ABSORB_TAU, KW_ABSORB_ETA, KW_ABSORB_NABLA1, KW_ABSORB_NABLA2,
  KW_SENSOR_MASK_INDEX, KW_SENSOR_MASK_CORNERS, KW_P0_SOURCE_INPUT, KW_P_SOURCE_INPUT, KW_TRANSDUCER_SOURCE_INPUT,
  KW_UX_SOURCE_INPUT, KW_UY_SOURCE_INPUT, KW_UZ_SOURCE_INPUT, KW_P_SOURCE_INDEX, KW_U_SOURCE_INDEX, KW_DELAY_MASK,
  KW_X_SHIFT_NEG_R, KW_Y_SHIFT_NEG_R, KW_Z_SHIFT_NEG_R,
  KW_ALPHA_COEFF, /* loaded into Temp1 by the reference (MatrixContainer.cpp:389-392) */
  KW_UX_SHIFTED, KW_UY_SHIFTED, KW_UZ_SHIFTED,
  KW_ARRAY_COUNT
};

/* Output streams, in sampling/flush order = OutputStreamContainer::OutputStreamIdx. */
enum kw_stream {
  KW_S_P_RAW = 0, KW_S_P_C, KW_S_P_RMS, KW_S_P_MAX, KW_S_P_MIN, KW_S_P_MAX_ALL, KW_S_P_MIN_ALL,
  KW_S_UX_RAW, KW_S_UY_RAW, KW_S_UZ_RAW, KW_S_UX_C, KW_S_UY_C, KW_S_UZ_C,
  KW_S_UX_NS_RAW, KW_S_UY_NS_RAW, KW_S_UZ_NS_RAW, KW_S_UX_NS_C, KW_S_UY_NS_C, KW_S_UZ_NS_C,
  KW_S_UX_RMS, KW_S_UY_RMS, KW_S_UZ_RMS, KW_S_UX_MAX, KW_S_UY_MAX, KW_S_UZ_MAX, KW_S_UX_MIN, KW_S_UY_MIN, KW_S_UZ_MIN,
  KW_S_UX_MAX_ALL, KW_S_UY_MAX_ALL, KW_S_UZ_MAX_ALL, KW_S_UX_MIN_ALL, KW_S_UY_MIN_ALL, KW_S_UZ_MIN_ALL,
  KW_S_IX_AVG, KW_S_IY_AVG, KW_S_IZ_AVG, KW_S_IX_AVG_C, KW_S_IY_AVG_C, KW_S_IZ_AVG_C, KW_S_Q_TERM, KW_S_Q_TERM_C,
  KW_STREAM_COUNT
};

typedef struct kw_ctx kw_ctx;

/* Library / device ------------------------------------------------------------------------------------------------ */
int kw_abi_version(void);
const char* kw_last_error(void);
/* SolverCudaKernels::getCudaCodeVersion (SolverCudaKernels.cuh:77): __CUDA_ARCH__/10 of the loaded kernels (100). */
int kw_cuda_code_version(int* version);
/* KSpaceFirstOrderSolver::getDeviceMemoryUsage / getAvailableDeviceMemory (cpp:470-492): cudaMemGetInfo of the current device. */
int kw_device_memory(size_t* free_bytes, size_t* total_bytes);

/* Context lifetime (KSpaceFirstOrderSolver ctor / allocateMemory / freeMemory, KSpaceFirstOrderSolver.cpp:91-151). */
int kw_ctx_create(const kw_config* cfg, kw_ctx** out);
int kw_ctx_destroy(kw_ctx* ctx);

/* MatrixContainer::loadDataFromInputFile + copyMatricesToDevice (Containers/MatrixContainer.cpp:486,544).
 * `count` = number of elements (float, float complex pairs, or uint64 for index arrays); count == 1 for a float array
 * selects the homogeneous (scalar) variant of that medium property (Parameters.cpp:426-459). */
int kw_set_array(kw_ctx* ctx, int array_id, const void* host, uint64_t count);
/* D2H of a state/operator array in the reference's layout (RealMatrix::copyFromDevice). count = capacity in elements. */
int kw_get_array(kw_ctx* ctx, int array_id, void* host, uint64_t count);

/* OutputStreamContainer::init / createStreams (Containers/OutputStreamContainer.cpp:70-325): enable before kw_preprocess. */
int kw_stream_enable(kw_ctx* ctx, int stream_id);

/* KSpaceFirstOrderSolver::preProcessing (cpp:784-857): index shift, dt/rho0_sg, kappa/nablas/tau/eta/source kappa,
 * c0 -> c0^2, constants upload. */
int kw_preprocess(kw_ctx* ctx);

/* computeMainLoop body (cpp:885-935) for up to `nsteps` steps starting at the context's t_index, including in-step
 * sampling of every enabled stream.  *steps_done receives the number executed (fewer than nsteps only when Nt is reached
 * or with KW_ERR_STREAM_FULL).  Asynchronous with respect to the host unless `sync` is non-zero. */
int kw_run(kw_ctx* ctx, uint64_t nsteps, uint64_t* steps_done, int sync);
int kw_time_index(kw_ctx* ctx, uint64_t* t_index);
int kw_synchronize(kw_ctx* ctx);

/* Streams: raw / compressed series return the rows buffered since the last fetch (row = one time step or one
 * compressed frame, Nsens (x harmonics x 2) floats, sample order = mask order / cuboids concatenated x-fastest);
 * aggregated streams return their accumulator (after kw_finish: post-processed, e.g. RMS).
 * Available: every id of kw_stream except KW_S_I{X,Y,Z}_AVG and KW_S_Q_TERM, which the reference computes after the run from
 * the STORED raw series (cpp:1231-1534): use kw_intensity_avg_block / kw_q_term below on the series the host has stored.
 * KW_S_I?_AVG_C and KW_S_Q_TERM_C create the do-not-save streams they read (p_c, u?_non_staggered_c, I?_avg_c) themselves
 * (OutputStreamContainer.cpp:273-323); such internal streams cannot be fetched. */
int kw_stream_info(kw_ctx* ctx, int stream_id, uint64_t* row_floats, uint64_t* rows_buffered);
int kw_stream_fetch(kw_ctx* ctx, int stream_id, float* host, uint64_t capacity_floats, uint64_t* rows_fetched);
/* Asynchronous output (OutputStreams/IndexOutputStream.cpp:583-591 writes every sampled step from a mapped host buffer; here rows are
 * buffered on the device): with kw_stream_async(ctx, 1), called before kw_preprocess, every raw / compressed series owns TWO device row
 * buffers and a pinned host buffer -- a full buffer is copied to the host on a separate stream while the time loop goes on sampling into
 * the other, and kw_run only returns KW_ERR_STREAM_FULL when both are occupied.  kw_stream_pending tells how many rows of a stream are on
 * their way to (or already in) pinned memory; kw_stream_info / kw_stream_fetch serve that older chunk first, then the rows still on the
 * device. */
int kw_stream_async(kw_ctx* ctx, int enable);
int kw_stream_pending(kw_ctx* ctx, int stream_id, uint64_t* rows);
/* Part of the running accumulator of an aggregate stream (rms / max / min [_all], I_avg_c) while the loop is in flight: `count`
 * floats from `offset`, device -> host on the solver stream (the reference exposes these only at the end of the run;
 * used for per-step monitoring and by the end-to-end benchmark). */
int kw_stream_peek(kw_ctx* ctx, int stream_id, uint64_t offset, float* host, uint64_t count);
/* OutputStreamContainer::postProcessStreams (cpp:950-973): RMS scaling, I_avg_c division. */
int kw_finish(kw_ctx* ctx);

/* Per-step host->device refresh of one row of a "many" source signal (time-major row t = Nsrc floats). Optional: the
 * whole signal can be given once through kw_set_array. */
int kw_set_source_row(kw_ctx* ctx, int array_id, uint64_t t_index, const float* host_row, uint64_t count);

/* CufftComplexMatrix::computeR2CFftND / computeC2RFftND (MatrixClasses/CufftComplexMatrix.cpp:508-534) on host
 * buffers, cuFFT layout: real [nz][ny][nx], complex [nz][ny][nx/2+1] interleaved, both unnormalised.
 * Grid sizes (here and in kw_ctx_create): each of nx, ny, nz (nz may be 1) a power of two in [16, 1024] (tuned kernels) or any
 * other multiple of 8 up to 2048 whose prime factors are 2, 3, 5, 7 (run-time-length kernels); anything else: KW_ERR_INVALID. */
/* Which kernel family transforms an axis of length n: 2 = tuned (powers of two in [16, 1024]), 1 = run-time-length kernels (other
 * multiples of 8 up to 2048 with prime factors 2, 3, 5, 7), 0 = not supported.  Host-side only (no device work): lets a front end
 * reject a grid before it loads gigabytes of input.  The reference accepts whatever cuFFT plans (CufftComplexMatrix.cpp:87-91). */
int kw_length_supported(uint64_t n);
int kw_fft_r2c_3d(uint64_t nx, uint64_t ny, uint64_t nz, const float* host_real, float* host_complex);
int kw_fft_c2r_3d(uint64_t nx, uint64_t ny, uint64_t nz, const float* host_complex, float* host_real);
/* The fused z pass on host buffers: the kernel that replaces the z stages of cuFFT together with cudaComputePressureGradient,
 * cudaComputeVelocityGradient, cudaComputeAbsorbtionTerm, cudaComputeSourceGradient and cudaComputeVelocityShiftIn{X,Y,Z}
 * (KSpaceSolver/SolverCudaKernels.cu:1139,1210,1812,740,2617-2689).  in, mul: [nz][ny][nx/2+1] (complex / real, mul may be
 * NULL); e = FFT_z(in)*(mul*scal); axis -1: out0 = IFFT_z(e); axis 0|1|2: out0 = IFFT_z(e (x) vec_{x|y|z}[k along that axis]);
 * axis 3: out0/1/2 = the three products with vec_x, vec_y, vec_z (one forward transform feeds three inverse ones).
 * vec_x: nx/2+1, vec_y: ny, vec_z: nz complex values.  Unnormalised transforms.  Unit-test entry; no context needed. */
int kw_fft_zmid(uint64_t nx, uint64_t ny, uint64_t nz, int axis, const float* in, const float* mul, float scal, const float* vec_x,
                const float* vec_y, const float* vec_z, float* out0, float* out1, float* out2);

/* Checkpoint / restart (KSpaceFirstOrderSolver::saveCheckpointData cpp:1176-1224, recovery in loadInputData :186-228).
 * The state of a run is t_index, the seven state arrays (kw_get_array / kw_set_array of KW_P, KW_RHO{X,Y,Z}, KW_U?_SG?, valid
 * after kw_preprocess too) and, per enabled stream, what BaseOutputStream::checkpoint stores (OutputStreams/
 * BaseOutputStream.cpp:528-606): the aggregate buffer, or the two compression accumulators (the reference's
 * Temp_<name>_1/_2 datasets) with the sampled / compressed step counters.  The state blob is opaque; size 0 = the stream is
 * not part of this run (internal do-not-save streams of I_avg_c DO have a state).  Buffered raw rows must be fetched first. */
int kw_set_time_index(kw_ctx* ctx, uint64_t t_index);
int kw_stream_state_size(kw_ctx* ctx, int stream_id, uint64_t* bytes);
int kw_stream_state_get(kw_ctx* ctx, int stream_id, void* buffer, uint64_t bytes);
int kw_stream_state_set(kw_ctx* ctx, int stream_id, const void* buffer, uint64_t bytes);
/* The same state in the pieces the REFERENCE's checkpoint files hold (so that a run interrupted by one code can be resumed by the
 * other): kw_set_time_index also sets the sampled / compressed step counters of every stream from t_index, as IndexOutputStream::reopen
 * does (IndexOutputStream.cpp:203-213); buffer 0 = the accumulator of an aggregate stream, which the reference flushes into the OUTPUT
 * file at a checkpoint and reloads from it (IndexOutputStream.cpp:536-557, :215-230); buffers 1 / 2 = the two compression accumulators
 * of a *_c stream, the reference's Temp_<name>_1 / _2 datasets (BaseOutputStream.cpp:528-606).  host == NULL queries the size. */
int kw_stream_buffer_get(kw_ctx* ctx, int stream_id, int which, float* host, uint64_t capacity_floats, uint64_t* floats);
int kw_stream_buffer_set(kw_ctx* ctx, int stream_id, int which, const float* host, uint64_t floats);

/* Post-processing of stored raw series (--I_avg, --Q_term; the reference re-reads them block-wise from the output file).
 * kw_intensity_avg_block = the block body of KSpaceFirstOrderSolver::computeAverageIntensities (cpp:1231-1534): p and the
 * ncomp non-staggered velocity components of n sensor points over `steps` stored samples, laid out [step][point] as in
 * the file; the velocity is shifted by half a time step spectrally, intensity[f][i] = sum_t p * u_shifted / steps.
 * kw_q_term = computeQTerm (cpp:1783-2080) on per-sensor intensities in mask order (local points of this rank);
 * kw_stream KW_S_Q_TERM_C does the same for the compressed intensities inside kw_finish. */
int kw_intensity_avg_block(const float* p, const float* const* u, int ncomp, uint64_t n, uint64_t steps, float* const* intensity);
int kw_q_term(kw_ctx* ctx, const float* const* intensity, int ncomp, float* q_out, uint64_t capacity);

/* Compression helpers.  kw_c40_encode / kw_c40_decode = CompressHelper::convertFloatCTo40b / convert40bToFloatC
 * (Compression/CompressHelper.cpp:292-389 / :224-290) on n complex values (interleaved re, im <-> 5 bytes each), run by the
 * device code the compressed streams use; max_exp = 138 (pressure) or 114 (velocity), CompressHelper.h:91-92.
 * kw_compression_bases returns the windowed bases bE / bE_1 (harmonics x bSize complex values, CompressHelper::getBE /
 * getBE_1 or their time-shifted variants) the context generated, with oSize and bSize (CompressHelper.cpp:48-65). */
int kw_c40_encode(const float* complex_pairs, uint64_t n, int max_exp, uint8_t* bytes);
int kw_c40_decode(const uint8_t* bytes, uint64_t n, int max_exp, float* complex_pairs);
int kw_compression_bases(kw_ctx* ctx, int shifted, float* be, float* be1, uint64_t capacity_complex, uint64_t* osize, uint64_t* bsize);

/* Slab decomposition over GPUs (no reference counterpart; the reference is single-GPU: MatrixContainer holds whole
 * arrays, Containers/MatrixContainer.cpp:418-476).  One context per rank (one process per GPU); kw_config.rank/nranks and
 * a ncclUniqueId created by ONE rank with kw_nccl_unique_id and distributed by the host (torch.distributed, MPI, a file).
 * Rank r owns planes z in [z_begin, z_begin + z_count) of every Nx*Ny*Nz array: kw_set_array / kw_get_array take and
 * return that slab (count = Nx*Ny*z_count) for full-grid arrays, and the complete data for 1-D vectors, signals and
 * index lists (global 1-based indices; each rank keeps the points of its slab).  Every 3-D transform exchanges the
 * half spectrum once with an all-to-all over NVLink (NCCL).  Stream rows hold the local sensor points only, in list
 * order; kw_sensor_layout returns, for each of them, its position in the row of the undecomposed run. */
int kw_nccl_unique_id(void* out128, uint64_t capacity);
int kw_local_slab(kw_ctx* ctx, uint64_t* z_begin, uint64_t* z_count);
int kw_sensor_layout(kw_ctx* ctx, uint64_t* total_points, uint64_t* local_points, uint64_t* positions, uint64_t capacity);
/* Bytes this rank has sent through the all-to-all since the context was created (NVLink GB/s = bytes / exchange time). */
int kw_comm_bytes(kw_ctx* ctx, double* bytes_sent);
/* How the all-to-all runs: 0 = single GPU (none), 1 = NCCL send/recv groups, 2 = copy-engine pushes into IPC-mapped peer
 * buffers ordered by stream memory operations (default on one node; KW_PEER=0 forces NCCL). */
int kw_comm_mode(kw_ctx* ctx, int* mode);

/* Timing of the device work of the last kw_run (CUDA events on the solver stream), milliseconds. */
int kw_last_run_ms(kw_ctx* ctx, float* ms);
/* Per-kernel device timing for roofline reports: when enabled, every launch of the time loop is bracketed by CUDA
 * events on the solver stream.  kw_profile_report writes a JSON object {kernel: {launches, ms, bytes}} where bytes is the
 * algorithmic traffic of those launches (DESIGN.md section "Kernels"). */
int kw_profile(kw_ctx* ctx, int enable, int reset);
int kw_profile_report(kw_ctx* ctx, char* buf, uint64_t capacity);
/* Number of kernels this library launched since the context was created. */
int kw_launch_count(kw_ctx* ctx, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* KWAVE_B200_H */
