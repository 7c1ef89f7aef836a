// Real-space updates of the time step, written as epilogues of the last inverse FFT pass (k_xinv) so that no
// full-grid element-wise pass remains, plus the O(Nsrc)/O(Nsens) scatter, gather and fix-up kernels.
//
// Replaces (KSpaceSolver/SolverCudaKernels.cu): cudaComputeVelocity* :184/:278, cudaComputeDensity{Nonlinear,Linear}
// :1358/:1470, cudaComputePressureTerms* :1577/:1724, cudaSumPressureTerms* :1865/:1966, cudaSumPressure*Lossless
// :2067/:2224, cudaAdd*Source :463/:504/:570/:679/:765/:795, cudaAddInitialPressureSource :864,
// cudaComputeInitialVelocity :949; and (OutputStreams/OutputStreamsCudaKernels.cu) cudaSampleIndex :83,
// cudaSampleCuboid :202, cudaSampleAll :297, cudaPostProcessingRms :359.
#pragma once
#include <cfloat>
#include <cstdint>
#include "fft_core.cuh"

namespace kw {

// a medium property that is either a full-grid array or a scalar (homogeneous variants of the reference kernels)
constexpr int kXThreadsStage = 256;  // = kXThreads: stride between the shared-memory slots of one thread (k_xinv)
struct Fld {
  const float* p;
  float s;
  __device__ __forceinline__ float at(size_t i) const { return p ? __ldg(p + i) : s; }
};

// Whole-domain and cuboid aggregates of a field, accumulated by the kernel that produces the field (in-step sampling:
// the reference runs cudaSampleAll / cudaSampleCuboid as separate full passes, OutputStreamsCudaKernels.cu:202-316).
struct FusedSample {
  float* max_all;
  float* min_all;
  float* rms;  // aggregates over ONE sensor cuboid (x fastest inside the cuboid); more cuboids use the stand-alone kernel
  float* mx;
  float* mn;
  int cub;  // 0: no cuboid aggregate, 1: cuboid == whole domain (buffer index == voxel index), 2: general cuboid
  int x0, x1, y0, y1, z0, z1;
  __device__ __forceinline__ void operator()(size_t i, int x, int y, int z, float p) const {
    if (max_all) max_all[i] = fmaxf(max_all[i], p);
    if (min_all) min_all[i] = fminf(min_all[i], p);
    if (cub == 0) return;
    size_t j = i;
    if (cub == 2) {
      if (x < x0 || x > x1 || y < y0 || y > y1 || z < z0 || z > z1) return;
      j = ((size_t)(z - z0) * (y1 - y0 + 1) + (y - y0)) * (x1 - x0 + 1) + (x - x0);
    }
    if (rms) rms[j] += p * p;
    if (mx) mx[j] = fmaxf(mx[j], p);
    if (mn) mn[j] = fminf(mn[j], p);
  }
};

// ---- epilogues of k_xinv -----------------------------------------------------------------------------------------
// res[f][m] = (value in row a, value in row b) at x = t + m*T; row b = row a + 1 (same z, y+1).

struct EpiStore {
  static constexpr int kStage = 0;
  static constexpr int kMinBlocks = 3;  // CTAs per SM the register budget is sized for  // plain C2R:  out = scale * ifft
  float* out[3];
  float scale;
  template <int NT> __device__ __forceinline__ void apply(float2 (&res)[1][8], int field, int t, size_t row0, int, int, const float* = nullptr, int n_rt = 0) const {
    const int N = NT ? NT : n_rt, T = N / 8;  // NT == 0: length at run time (fft_generic.cu)
    float* o = out[field] + row0 * N;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      o[t + m * T] = res[0][m].x * scale;
      o[N + t + m * T] = res[0][m].y * scale;
    }
  }
};

struct EpiAdd {
  static constexpr int kStage = 0;
  static constexpr int kMinBlocks = 3;  // CTAs per SM the register budget is sized for  // additive (k-space corrected) source: target_j += ifft   (SolverCudaKernels.cu:765-807)
  float* out[3];
  int ntargets;
  template <int NT> __device__ __forceinline__ void apply(float2 (&res)[1][8], int, int t, size_t row0, int, int, const float* = nullptr, int n_rt = 0) const {
    const int N = NT ? NT : n_rt, T = N / 8;  // NT == 0: length at run time (fft_generic.cu)
    for (int j = 0; j < ntargets; ++j) {
      float* o = out[j] + row0 * N;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        o[t + m * T] += res[0][m].x;
        o[N + t + m * T] += res[0][m].y;
      }
    }
  }
};

// u_i = (u_i*pml_i - (fd*g_i)*dtrho_i)*pml_i      (SolverCudaKernels.cu:199-212; homogeneous :287-305)
// init: u_i = g_i * (dtrho_i * (fd*0.5))            (SolverCudaKernels.cu:971-980)
struct EpiVelocity {
  static constexpr int kStage = 0;
  static constexpr int kMinBlocks = 3;  // CTAs per SM the register budget is sized for
  float* u[3];
  Fld dtrho[3];
  const float* pml_sg[3];
  float fd;
  int init;
  template <int NT> __device__ __forceinline__ void apply(float2 (&res)[1][8], int field, int t, size_t row0, int y, int z, const float* = nullptr, int n_rt = 0) const {
    const int N = NT ? NT : n_rt, T = N / 8;  // NT == 0: length at run time (fft_generic.cu)
    const int f = field;
    float* ua = u[f] + row0 * N;
    const Fld d = dtrho[f];
    if (init) {
      const float div = fd * 0.5f;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int x = t + m * T;
        const size_t i = row0 * N + x;
        ua[x] = res[0][m].x * (d.at(i) * div);
        ua[N + x] = res[0][m].y * (d.at(i + N) * div);
      }
      return;
    }
    const float py_a = (f == 1) ? __ldg(pml_sg[1] + y) : (f == 2) ? __ldg(pml_sg[2] + z) : 0.f;
    const float py_b = (f == 1) ? __ldg(pml_sg[1] + y + 1) : py_a;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int x = t + m * T;
      const size_t i = row0 * N + x;
      const float pa = (f == 0) ? __ldg(pml_sg[0] + x) : py_a;
      const float pb = (f == 0) ? pa : py_b;
      ua[x] = (ua[x] * pa - (fd * res[0][m].x) * d.at(i)) * pa;
      ua[N + x] = (ua[N + x] * pb - (fd * res[0][m].y) * d.at(i + N)) * pb;
    }
  }
};

// density update (+ pressure terms or lossless pressure), three gradients at the same voxel.
//  nonlinear: s = (2(rx+ry+rz) + rho0)*dt ; r_i = pml_i*((pml_i*r_i) - s*du_i)        (SolverCudaKernels.cu:1377-1390)
//  linear:    r_i = pml_i*(pml_i*r_i - (dt*rho0)*du_i)                                 (:1486-1494)
//  absorbing: A = rho0*(dux+duy+duz) ; B = sum r ; NL = (BonA*B*B)/(2 rho0) + B        (:1588-1601, :1733-1741)
//  lossless:  p = c2*(B + BonA*(B*B)/(2 rho0))  |  p = c2*B                            (:2079-2082, :2229-2235)
struct EpiDensity {
  static constexpr int kMinBlocks = 2;  // CTAs per SM the register budget is sized for
#ifndef KW_STAGE_DENSITY
#define KW_STAGE_DENSITY 1
#endif
  // shared-memory slots per thread: rho_x, rho_y, rho_z, rho0, BonA at the thread's 16 voxels (slot = k*16 + v)
  static constexpr int kStage = KW_STAGE_DENSITY ? 5 * 16 : 0;
  template <int N> __device__ __forceinline__ void stage(float* stg, int t, size_t row0) const {
    constexpr int T = N / 8;
    const bool need_bona = nonlinear && !defer_terms && bona.p;
#pragma unroll
    for (int v = 0; v < 16; ++v) {
      const size_t i = row0 * N + (v & 1) * N + t + (v >> 1) * T;
      cp_async4(stg + (0 * 16 + v) * kXThreadsStage, rho[0] + i);
      cp_async4(stg + (1 * 16 + v) * kXThreadsStage, rho[1] + i);
      cp_async4(stg + (2 * 16 + v) * kXThreadsStage, rho[2] + i);
      if (rho0.p) cp_async4(stg + (3 * 16 + v) * kXThreadsStage, rho0.p + i);
      if (need_bona) cp_async4(stg + (4 * 16 + v) * kXThreadsStage, bona.p + i);
    }
  }
  float* rho[3];
  Fld rho0, bona, c2;
  const float* pml[3];
  float dt;
  int nonlinear, absorbing;
  int defer_terms;  // a pressure source follows: B / NL / p are (re)computed by k_pressure_terms afterwards
  float* outA;
  float* outB;
  float* outNL;
  float* p;
  FusedSample fs;  // sampling of p when this epilogue produces the final pressure of the step (lossless)
  int sample;
  template <int NT> __device__ __forceinline__ void apply(float2 (&res)[3][8], int, int t, size_t row0, int y, int z, const float* stg = nullptr, int n_rt = 0) const {
    const int N = NT ? NT : n_rt, T = N / 8;  // NT == 0: length at run time (fft_generic.cu)
    // restrict-qualified locals: the arrays never alias, which lets the loads of all voxels be issued ahead of the stores
    float* __restrict__ rxp = rho[0];
    float* __restrict__ ryp = rho[1];
    float* __restrict__ rzp = rho[2];
    float* __restrict__ oA = outA;
    float* __restrict__ oB = outB;
    float* __restrict__ oNL = outNL;
    float* __restrict__ pp = p;
    const float pya = __ldg(pml[1] + y), pyb = __ldg(pml[1] + y + 1), pz = __ldg(pml[2] + z);
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // two batches of 4 x-positions x 2 rows: loads first, then arithmetic, then stores
      float rx[8], ry[8], rz[8], r0[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const size_t i = row0 * N + (q & 1) * N + t + (h * 4 + (q >> 1)) * T;
        if (kStage > 0 && stg) {
          const int v = h * 8 + q;  // = 2*m + row, the slot order of stage()
          rx[q] = stg[(0 * 16 + v) * kXThreadsStage], ry[q] = stg[(1 * 16 + v) * kXThreadsStage], rz[q] = stg[(2 * 16 + v) * kXThreadsStage];
          r0[q] = rho0.p ? stg[(3 * 16 + v) * kXThreadsStage] : rho0.s;
        } else {
          rx[q] = rxp[i], ry[q] = ryp[i], rz[q] = rzp[i], r0[q] = rho0.at(i);
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int m = h * 4 + (q >> 1), x = t + m * T;
        const size_t i = row0 * N + (q & 1) * N + x;
        const float px = __ldg(pml[0] + x), py = (q & 1) ? pyb : pya;
        const float dux = (q & 1) ? res[0][m].y : res[0][m].x;
        const float duy = (q & 1) ? res[1][m].y : res[1][m].x;
        const float duz = (q & 1) ? res[2][m].y : res[2][m].x;
        float ax = rx[q], ay = ry[q], az = rz[q];
        if (nonlinear) {
          const float s = (2.0f * (ax + ay + az) + r0[q]) * dt;
          ax = px * ((px * ax) - s * dux);
          ay = py * ((py * ay) - s * duy);
          az = pz * ((pz * az) - s * duz);
        } else {
          const float d = dt * r0[q];
          ax = px * (px * ax - d * dux);
          ay = py * (py * ay - d * duy);
          az = pz * (pz * az - d * duz);
        }
        rxp[i] = ax, ryp[i] = ay, rzp[i] = az;
        if (absorbing) oA[i] = r0[q] * (dux + duy + duz);
        if (!defer_terms) {
          const float sum = ax + ay + az;
          const float bq = !nonlinear ? 0.f : (kStage > 0 && stg && bona.p) ? stg[(4 * 16 + h * 8 + q) * kXThreadsStage] : bona.at(i);
          if (absorbing) {
            oB[i] = sum;
            if (nonlinear) oNL[i] = ((bq * sum * sum) / (2.0f * r0[q])) + sum;
          } else {
            const float pv = nonlinear ? c2.at(i) * (sum + (bq * (sum * sum) / (2.0f * r0[q]))) : c2.at(i) * sum;
            pp[i] = pv;
            if (sample) fs(i, x, y + (q & 1), z, pv);
          }
        }
      }
    }
  }
};

// p = c2*(NL + fd*((ta*tau) - (tb*eta)))   nonlinear  (SolverCudaKernels.cu:1872-1878)
// p = c2*(B  + fd*(ta*tau - tb*eta))       linear     (:1973-1979)
struct EpiPressureSum {
  static constexpr int kMinBlocks = 2;  // CTAs per SM the register budget is sized for
#ifndef KW_STAGE_PSUM
#define KW_STAGE_PSUM 1
#endif
  // shared-memory slots per thread: base (NL or B), c2, tau, eta at the thread's 16 voxels
  static constexpr int kStage = KW_STAGE_PSUM ? 4 * 16 : 0;
  template <int N> __device__ __forceinline__ void stage(float* stg, int t, size_t row0) const {
    constexpr int T = N / 8;
#pragma unroll
    for (int v = 0; v < 16; ++v) {
      const size_t i = row0 * N + (v & 1) * N + t + (v >> 1) * T;
      cp_async4(stg + (0 * 16 + v) * kXThreadsStage, base + i);
      if (c2.p) cp_async4(stg + (1 * 16 + v) * kXThreadsStage, c2.p + i);
      if (tau.p) cp_async4(stg + (2 * 16 + v) * kXThreadsStage, tau.p + i);
      if (eta.p) cp_async4(stg + (3 * 16 + v) * kXThreadsStage, eta.p + i);
    }
  }
  float* p;
  const float* base;  // NL (nonlinear) or B (linear)
  Fld c2, tau, eta;
  float fd;
  FusedSample fs;
  int sample;
  template <int NT> __device__ __forceinline__ void apply(float2 (&res)[2][8], int, int t, size_t row0, int y, int z, const float* stg = nullptr, int n_rt = 0) const {
    const int N = NT ? NT : n_rt, T = N / 8;  // NT == 0: length at run time (fft_generic.cu)
    float* __restrict__ pp = p;
    float pv[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int m = q >> 1;
      const size_t i = row0 * N + (q & 1) * N + t + m * T;
      const float ta = (q & 1) ? res[0][m].y : res[0][m].x, tb = (q & 1) ? res[1][m].y : res[1][m].x;
      if (kStage > 0 && stg) {
        const float bv = stg[(0 * 16 + q) * kXThreadsStage];
        const float cv = c2.p ? stg[(1 * 16 + q) * kXThreadsStage] : c2.s;
        const float tv = tau.p ? stg[(2 * 16 + q) * kXThreadsStage] : tau.s;
        const float ev = eta.p ? stg[(3 * 16 + q) * kXThreadsStage] : eta.s;
        pv[q] = cv * (bv + fd * ((ta * tv) - (tb * ev)));
      } else {
        pv[q] = c2.at(i) * (__ldg(base + i) + fd * ((ta * tau.at(i)) - (tb * eta.at(i))));
      }
    }
    if (sample) {
      float* __restrict__ mxa = fs.max_all;
      float* __restrict__ mna = fs.min_all;
      float* __restrict__ rms = fs.rms;
      float* __restrict__ cmx = fs.mx;
      float* __restrict__ cmn = fs.mn;
      if (fs.cub != 2) {  // whole-domain buffers: index == voxel index; batches of loads, then stores
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float a0[4], a1[4], a2[4], a3[4], a4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int q = h * 4 + k;
            const size_t i = row0 * N + (q & 1) * N + t + (q >> 1) * T;
            if (mxa) a0[k] = mxa[i];
            if (mna) a1[k] = mna[i];
            if (fs.cub == 1) {
              if (rms) a2[k] = rms[i];
              if (cmx) a3[k] = cmx[i];
              if (cmn) a4[k] = cmn[i];
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int q = h * 4 + k;
            const size_t i = row0 * N + (q & 1) * N + t + (q >> 1) * T;
            if (mxa) mxa[i] = fmaxf(a0[k], pv[q]);
            if (mna) mna[i] = fminf(a1[k], pv[q]);
            if (fs.cub == 1) {
              if (rms) rms[i] = a2[k] + pv[q] * pv[q];
              if (cmx) cmx[i] = fmaxf(a3[k], pv[q]);
              if (cmn) cmn[i] = fminf(a4[k], pv[q]);
            }
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int x = t + (q >> 1) * T;
          fs(row0 * N + (q & 1) * N + x, x, y + (q & 1), z, pv[q]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) pp[row0 * N + (q & 1) * N + t + (q >> 1) * T] = pv[q];
  }
};

// ---- pressure terms from the densities (used after a pressure source touched rho, and on the unfused path) --------
struct TermsArgs {
  const float* rho[3];
  Fld rho0, bona, c2;
  int nonlinear, absorbing;
  float* outB;
  float* outNL;
  float* p;
  const uint64_t* index;  // nullptr: whole grid; else only these voxels (source points)
  size_t n;
};
__device__ __forceinline__ void pressure_terms_voxel(const TermsArgs& a, size_t i) {
  const float sum = a.rho[0][i] + a.rho[1][i] + a.rho[2][i];
  const float r0 = a.rho0.at(i);
  if (a.absorbing) {
    a.outB[i] = sum;
    if (a.nonlinear) a.outNL[i] = ((a.bona.at(i) * sum * sum) / (2.0f * r0)) + sum;
  } else {
    a.p[i] = a.nonlinear ? a.c2.at(i) * (sum + (a.bona.at(i) * (sum * sum) / (2.0f * r0))) : a.c2.at(i) * sum;
  }
}
static __global__ void k_pressure_terms(TermsArgs a) {
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < a.n; j += (size_t)gridDim.x * blockDim.x)
    pressure_terms_voxel(a, a.index ? (size_t)a.index[j] : j);
}

// ---- sources -----------------------------------------------------------------------------------------------------
// mode 0: '=' (Dirichlet), 1: '+=' (additive, no correction); many: signal[t*Nsrc + j] else signal[t]
// (SolverCudaKernels.cu:504-527, :570-629).  ntargets = 1 (velocity component) or 3 (rhox, rhoy, rhoz).
// Slab-decomposed runs keep only the source points of the local slab: index[] holds LOCAL voxel indices and pos[] the
// position of each kept point in the original list (nullptr: identity), nsrc_total = length of the original list.
struct SourceArgs {
  float* target[3];
  int ntargets;
  const float* signal;
  const uint64_t* index;
  const uint64_t* pos;
  size_t nsrc, nsrc_total;
  size_t t;
  int many, mode;
};
static __global__ void k_add_source(SourceArgs a) {
  const size_t base = a.many ? a.t * a.nsrc_total : a.t;
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < a.nsrc; j += (size_t)gridDim.x * blockDim.x) {
    const float s = a.many ? a.signal[base + (a.pos ? a.pos[j] : j)] : a.signal[base];
    const size_t i = a.index[j];
    for (int k = 0; k < a.ntargets; ++k) {
      if (a.mode == 0) a.target[k][i] = s;
      else a.target[k][i] += s;
    }
  }
}
// ux[idx[j]] += signal[delay[j] + t]   (SolverCudaKernels.cu:463-471)
static __global__ void k_add_transducer(float* ux, const uint64_t* index, const float* signal, const uint64_t* delay, size_t nsrc, size_t t) {
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < nsrc; j += (size_t)gridDim.x * blockDim.x)
    ux[index[j]] += signal[delay[j] + t];
}
// scaled[idx[j]] = s_j   into a zeroed grid (SolverCudaKernels.cu:679-697)
static __global__ void k_insert_source(float* grid, const float* signal, const uint64_t* index, const uint64_t* pos, size_t nsrc,
                                       size_t nsrc_total, size_t t, int many) {
  const size_t base = many ? t * nsrc_total : t;
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < nsrc; j += (size_t)gridDim.x * blockDim.x)
    grid[index[j]] = many ? signal[base + (pos ? pos[j] : j)] : signal[base];
}
// p = p0 ; rho_i = p0 / (3*c2)   (SolverCudaKernels.cu:870-883)
// dims = 3, or 2 for Nz == 1 where rho = p0 / (2*c2) and there is no z component (SolverCudaKernels.cu:870-883)
static __global__ void k_initial_pressure(float* p, float* rx, float* ry, float* rz, const float* p0, Fld c2, size_t n, int dims) {
  const float d = (float)dims;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float tmp = p[i] = p0[i];
    tmp = tmp / (d * c2.at(i));
    rx[i] = tmp;
    ry[i] = tmp;
    rz[i] = dims == 3 ? tmp : 0.f;
  }
}

// ---- sampling ----------------------------------------------------------------------------------------------------
enum SampleOp { kOpNone = 0, kOpRms = 1, kOpMax = 2, kOpMin = 3, kOpC = 4, kOpIAvgC = 5, kOpQTermC = 6 };  // BaseOutputStream::ReduceOperator
template <int OP> __device__ __forceinline__ void reduce_into(float* buf, size_t i, float x) {
  if (OP == kOpNone) buf[i] = x;
  else if (OP == kOpRms) buf[i] += x * x;
  else if (OP == kOpMax) buf[i] = fmaxf(buf[i], x);
  else buf[i] = fminf(buf[i], x);
}
template <int OP> static __global__ void k_sample_index(float* buf, const float* __restrict__ src, const uint64_t* __restrict__ mask, size_t n) {
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
    reduce_into<OP>(buf, j, __ldg(src + mask[j]));
}
// all cuboids in one launch: corners are 0-based (x0,y0,z0,x1,y1,z1) per cuboid, offsets = running start of each
// cuboid in the concatenated buffer, x fastest inside a cuboid (OutputStreamsCudaKernels.cu:164-230).
struct CuboidArgs {
  const uint64_t* corners;
  const uint64_t* offsets;  // ncuboids + 1
  int ncuboids;
  int nx, ny;
};
template <int OP> static __global__ void k_sample_cuboid(float* buf, const float* __restrict__ src, CuboidArgs a, size_t total) {
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < total; j += (size_t)gridDim.x * blockDim.x) {
    int c = 0;
    while (c + 1 < a.ncuboids && j >= a.offsets[c + 1]) ++c;
    const uint64_t* k = a.corners + 6 * c;
    const size_t l = j - a.offsets[c];
    const size_t cx = k[3] - k[0] + 1, cy = k[4] - k[1] + 1;
    const size_t x = l % cx, y = (l / cx) % cy, z = l / (cx * cy);
    const size_t i = ((z + k[2]) * a.ny + (y + k[1])) * a.nx + (x + k[0]);
    reduce_into<OP>(buf, j, __ldg(src + i));
  }
}
template <int OP> static __global__ void k_sample_all(float* buf, const float* __restrict__ src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    reduce_into<OP>(buf, i, __ldg(src + i));
}
// inverse of the gathers: grid[voxel of sensor point j] = buf[j]  (computeQTerm, KSpaceFirstOrderSolver.cpp:1799-1863)
static __global__ void k_scatter_index(float* grid, const float* __restrict__ buf, const uint64_t* __restrict__ mask, size_t n) {
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) grid[mask[j]] = buf[j];
}
static __global__ void k_scatter_cuboid(float* grid, const float* __restrict__ buf, CuboidArgs a, size_t total) {
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < total; j += (size_t)gridDim.x * blockDim.x) {
    int c = 0;
    while (c + 1 < a.ncuboids && j >= a.offsets[c + 1]) ++c;
    const uint64_t* k = a.corners + 6 * c;
    const size_t l = j - a.offsets[c];
    const size_t cx = k[3] - k[0] + 1, cy = k[4] - k[1] + 1;
    const size_t x = l % cx, y = (l / cx) % cy, z = l / (cx * cy);
    grid[((z + k[2]) * a.ny + (y + k[1])) * a.nx + (x + k[0])] = buf[j];
  }
}
static __global__ void k_negate(float* buf, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = -buf[i];
}
static __global__ void k_fill(float* buf, float v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = v;
}
// sqrt(buf * 1/(Nt - s))   (OutputStreamsCudaKernels.cu:359-363)
static __global__ void k_post_rms(float* buf, float scaling, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    buf[i] = sqrtf(buf[i] * scaling);
}

// ---- on-the-fly harmonic compression (OutputStreams/IndexOutputStream.cpp:373-470, CuboidOutputStream.cpp:431-532) ---
// Per sampled step and (sensor i, harmonic ih):  acc1 += bE[ih][stepLocal]*x_i ; acc2 += bE_1[ih][stepLocal]*x_i ;
// first saving step (overlap mode): acc2 += acc1.  The products and sums are rounded separately (the reference's host
// loop is SSE2 code without FMA contraction).  no_overlap: acc2 aliases acc1 (BaseOutputStream.cpp:246-257).
__device__ __forceinline__ float2 c_axpy(float2 acc, float2 e, float x) {
  return make_float2(__fadd_rn(acc.x, __fmul_rn(e.x, x)), __fadd_rn(acc.y, __fmul_rn(e.y, x)));
}
struct CompressArgs {
  const float* x;  // the raw samples of this step, sensor order
  float2 *acc1, *acc2;
  uint8_t *q1, *q2;  // 40-bit mode: the accumulators are kept packed and re-quantised every step, as in the reference
  const float2 *be, *be1;
  size_t n;  // Nsens * H
  int H, bsize, step_local, mirror, e;
};
static __global__ void k_compress(CompressArgs a) {
  for (size_t ph = blockIdx.x * (size_t)blockDim.x + threadIdx.x; ph < a.n; ph += (size_t)gridDim.x * blockDim.x) {
    const size_t i = ph / a.H;
    const int ih = (int)(ph % a.H);
    const float x = a.x[i];
    const float2 e = __ldg(a.be + (size_t)ih * a.bsize + a.step_local), e1 = __ldg(a.be1 + (size_t)ih * a.bsize + a.step_local);
    const float2 c1 = c_axpy(a.acc1[ph], e, x);
    a.acc1[ph] = c1;
    float2 c2 = c_axpy(a.acc2 == a.acc1 ? c1 : a.acc2[ph], e1, x);
    if (a.mirror) c2 = make_float2(__fadd_rn(c2.x, c1.x), __fadd_rn(c2.y, c1.y));
    a.acc2[ph] = c2;
  }
}

// 40-bit complex: byte0 = sR|sI|mR[16]|mI[16]|e[3:0], then two little-endian 16-bit mantissas; the shared 4-bit exponent
// is offset by `e` (138 pressure, 114 velocity).  Integer arithmetic of CompressHelper::convert40bToFloatC /
// convertFloatCTo40b (Compression/CompressHelper.cpp:224-389), bit-exact.
__device__ __forceinline__ float2 c40_decode(const uint8_t* b, int e) {
  const uint32_t b0 = b[0];
  uint32_t mr = ((b0 & 0x20u) << 11) | (b[1] | ((uint32_t)b[2] << 8));
  uint32_t mi = ((b0 & 0x10u) << 12) | (b[3] | ((uint32_t)b[4] << 8));
  const uint32_t sr = b0 >> 7, si = (b0 & 0x40u) >> 6;
  int er = (int)(b0 & 0xFu) + e, ei = er;
  mr <<= 6, mi <<= 6;
  if (mr) {
    const int idx = 31 - __clz(mr);
    mr <<= 23 - idx, er -= 22 - idx;
  } else er = 0;
  if (mi) {
    const int idx = 31 - __clz(mi);
    mi <<= 23 - idx, ei -= 22 - idx;
  } else ei = 0;
  return make_float2(__uint_as_float((sr << 31) | ((uint32_t)er << 23) | (mr & 0x007FFFFFu)),
                     __uint_as_float((si << 31) | ((uint32_t)ei << 23) | (mi & 0x007FFFFFu)));
}
__device__ __forceinline__ void c40_encode(float2 c, uint8_t* b, int e) {
  uint32_t mr = __float_as_uint(c.x), mi = __float_as_uint(c.y);
  const uint32_t sr = mr >> 31, si = mi >> 31;
  const int ers = (int)((mr & 0x7F800000u) >> 23) - e, eis = (int)((mi & 0x7F800000u) >> 23) - e;
  int es = ers;
  mr &= 0x007FFFFFu, mi &= 0x007FFFFFu;
  int rsr = 6, rsi = 6;
  if (ers > eis) rsi += ers - eis, es = ers;
  else if (eis > ers) rsr += eis - ers, es = eis;
  if (es < 0) rsr += -es, rsi += -es, es = 0;
  rsr &= 0xFF, rsi &= 0xFF;  // the reference keeps the shifts in uint8_t
  rsr = rsr > 23 ? 23 : rsr, rsi = rsi > 23 ? 23 : rsi;
  mr >>= rsr, mi >>= rsi;
  if (mr > 0 && mr != (0x7FFFFFu >> rsr)) mr += 1;
  if (mi > 0 && mi != (0x7FFFFFu >> rsi)) mi += 1;
  mr |= 1u << (23 - rsr), mr >>= 1;
  mi |= 1u << (23 - rsi), mi >>= 1;
  if (es > 0xF) mr = mi = 0xFFFFu, es = 0xF;
  b[0] = (uint8_t)((sr << 7) | (si << 6) | ((mr & 0x10000u) >> 11) | ((mi & 0x10000u) >> 12) | ((uint32_t)es & 0xFu));
  b[1] = (uint8_t)(mr & 0xFF), b[2] = (uint8_t)((mr >> 8) & 0xFF), b[3] = (uint8_t)(mi & 0xFF), b[4] = (uint8_t)((mi >> 8) & 0xFF);
}
static __global__ void k_compress40(CompressArgs a) {
  for (size_t ph = blockIdx.x * (size_t)blockDim.x + threadIdx.x; ph < a.n; ph += (size_t)gridDim.x * blockDim.x) {
    const size_t i = ph / a.H;
    const int ih = (int)(ph % a.H);
    const float x = a.x[i];
    const float2 e = __ldg(a.be + (size_t)ih * a.bsize + a.step_local), e1 = __ldg(a.be1 + (size_t)ih * a.bsize + a.step_local);
    uint8_t *p1 = a.q1 + ph * 5, *p2 = a.q2 + ph * 5;
    float2 c1 = c40_decode(p1, a.e);
    if (a.q1 == a.q2) {  // no overlap: cc1 += bE*x + bE_1*x  (IndexOutputStream.cpp:419-423)
      c1 = make_float2(__fadd_rn(c1.x, __fadd_rn(__fmul_rn(e.x, x), __fmul_rn(e1.x, x))), __fadd_rn(c1.y, __fadd_rn(__fmul_rn(e.y, x), __fmul_rn(e1.y, x))));
      c40_encode(c1, p1, a.e);
      continue;
    }
    float2 c2 = c40_decode(p2, a.e);
    c1 = c_axpy(c1, e, x), c2 = c_axpy(c2, e1, x);
    c40_encode(c1, p1, a.e);
    if (a.mirror) c2 = make_float2(__fadd_rn(c2.x, c1.x), __fadd_rn(c2.y, c1.y));  // the unquantised cc1, as in the reference (:433-437)
    c40_encode(c2, p2, a.e);
  }
}
// I_avg_c: I[i] += sum_h Re(P_h * conj(U_h)) / 2 over the frames just completed (IndexOutputStream.cpp:299-342)
static __global__ void k_intensity_c(float* I, const float2* pc, const float2* uc, const uint8_t* pq, const uint8_t* uq, size_t nsens, int H, int ep, int eu) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nsens; i += (size_t)gridDim.x * blockDim.x) {
    float acc = I[i];
    for (int ih = 0; ih < H; ++ih) {
      const size_t ph = i * H + ih;
      const float2 P = pq ? c40_decode(pq + ph * 5, ep) : pc[ph], U = uq ? c40_decode(uq + ph * 5, eu) : uc[ph];
      acc = __fadd_rn(acc, __fadd_rn(__fmul_rn(P.x, U.x), __fmul_rn(P.y, U.y)) / 2.0f);
    }
    I[i] = acc;
  }
}
static __global__ void k_divide(float* buf, float d, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = buf[i] / d;
}

// ---- layout helpers ------------------------------------------------------------------------------------------------
// reduced-grid real operator [nz][ny][nxr] (reference layout) -> padded [nz][ny][nxp]
static __global__ void k_pad_real(float* dst, const float* src, int nxr, int nxp, size_t rows) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < rows * nxp; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / nxp;
    const int x = (int)(i % nxp);
    dst[i] = x < nxr ? src[r * nxr + x] : 0.f;
  }
}
static __global__ void k_pad_complex(float2* dst, const float2* src, int nxr, int nxp, size_t rows, int to_padded) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < rows * nxp; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / nxp;
    const int x = (int)(i % nxp);
    if (to_padded) dst[i] = x < nxr ? src[r * nxr + x] : make_float2(0.f, 0.f);
    else if (x < nxr) dst[r * nxr + x] = src[i];
  }
}

}  // namespace kw
