// The reference includes <KSpaceSolver/SolverCUDAKernels.cuh> (upper-case "CUDA") while the file in its tree is
// SolverCudaKernels.cuh (the Windows projects use the upper-case name); on a case-sensitive file system this one-line
// forwarder resolves it.  No reference code is copied.
#include <KSpaceSolver/SolverCudaKernels.cuh>
