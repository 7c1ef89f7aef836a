// Plane-fused x/y passes: the x pass and the y pass of a 3-D transform in ONE kernel, with the intermediate half
// spectrum of a z plane handed from one pass to the other through a small ring of plane buffers that stays resident in
// the 126 MB L2 (a 512 x 272 complex plane is 1.06 MiB) instead of making a round trip through HBM.
//
//   k_xy_fwd   real planes -> x transform (rows)     -> ring -> y transform (column tiles) -> spectrum [z][y][NXP]
//   k_yx_inv   spectrum    -> y transform (col tiles) -> ring -> x transform (rows) + real-space epilogue
//
// Work is cut into items (a few row pairs of one plane, or one column tile of one plane) that the persistent CTAs
// claim in a fixed global order from an atomic queue.  The order interleaves first-pass items of plane k with
// second-pass items of plane k-L, so the ring only ever holds R > L planes.  An item waits (one thread polls a counter)
// until the items it depends on have completed; those always sit EARLIER in the order, i.e. they have already been
// claimed by CTAs that are running -- progress never depends on a CTA that has not started, so no co-residency
// assumption is needed.  Completion is published with the release pattern  stores -> __syncthreads -> fence -> atomic,
// and consumed with  poll -> fence -> __syncthreads -> ld.global.cg  (ring reads bypass the non-coherent L1).
//
// Replaces, together with k_zmid, cufftExecR2C / cufftExecC2R (MatrixClasses/CufftComplexMatrix.cpp:511,527); HBM
// traffic per transform drops from 4N + 8Nc + 16Nc to 4N + 8Nc (forward) and from 16Nc + 8Nc to 8Nc (inverse).
// Only instantiated for Nx == Ny == KW_N (all benchmark grids); other shapes use the separate k_xfwd / k_col / k_xinv.
#pragma once
#include "fft_kernels.cuh"

namespace kw {

struct PipeArgs {
  unsigned* ctr;        // counter set of this launch: [0] queue head, [1 .. P] first-pass done, [1+P .. 2P] second-pass done
  unsigned* ctr_other;  // the other set: zeroed by this launch for the next one
  int nctr;             // counters per set
  int P;                // planes (or plane groups)
  int I1, I2;           // items per plane in the first / second pass
  int L, R;             // lag (planes) between the passes, ring slots (R > L)
  int* err;             // set to 1 when a dependency wait times out (a bug, never a normal condition)
};

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Claims the next item and waits for its dependencies.  Returns false when the queue is exhausted.
// pass = 1 / 2, plane, sub = item inside the plane.
struct PipeItem {
  int pass, plane, sub;
};
__device__ __forceinline__ bool pipe_next(const PipeArgs& q, int tid, int* s_slot, PipeItem* it) {
  if (tid == 0) {
    const int total = q.P * (q.I1 + q.I2);
    const int idx = (int)atomicAdd(q.ctr, 1u);
    int pass = 0, plane = 0, sub = 0;
    if (idx < total) {
      const int A = q.L * q.I1, per = q.I1 + q.I2, nB = (q.P - q.L) * per;
      if (idx < A) pass = 1, plane = idx / q.I1, sub = idx % q.I1;
      else if (idx - A < nB) {
        const int tick = (idx - A) / per, r = (idx - A) % per;
        if (r < q.I1) pass = 1, plane = q.L + tick, sub = r;
        else pass = 2, plane = tick, sub = r - q.I1;
      } else {
        const int j = idx - A - nB;
        pass = 2, plane = (q.P - q.L) + j / q.I2, sub = j % q.I2;
      }
      // dependencies: second pass of plane p needs all first-pass items of p; first pass of plane p reuses the ring slot
      // of plane p - R and needs all second-pass items of that plane
      const unsigned* c = nullptr;
      unsigned need = 0;
      if (pass == 2) c = q.ctr + 1 + plane, need = (unsigned)q.I1;
      else if (plane >= q.R) c = q.ctr + 1 + q.P + (plane - q.R), need = (unsigned)q.I2;
      if (c) {
        unsigned spins = 0;
        while (ld_relaxed(c) < need) {
          __nanosleep(40);
          if (++spins > (1u << 26)) {  // seconds: something is broken; do not hang the GPU
            *q.err = 1;
            break;
          }
        }
        __threadfence();
      }
    }
    s_slot[0] = pass, s_slot[1] = plane, s_slot[2] = sub;
  }
  __syncthreads();
  it->pass = s_slot[0], it->plane = s_slot[1], it->sub = s_slot[2];
  __syncthreads();  // s_slot may be rewritten by the next claim
  return it->pass != 0;
}
__device__ __forceinline__ void pipe_done(const PipeArgs& q, int tid, const PipeItem& it) {
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    atomicAdd(q.ctr + 1 + (it.pass == 1 ? 0 : q.P) + it.plane, 1u);
  }
}
__device__ __forceinline__ void pipe_prologue(const PipeArgs& q, int tid, int nthreads) {
  for (int i = blockIdx.x * nthreads + tid; i < q.nctr; i += gridDim.x * nthreads) q.ctr_other[i] = 0u;
}

struct XYFwdArgs {
  const float* in[kMaxFields];
  float2* out[kMaxFields];
  float2* ring;
  const float2* tab;
  int nz, nxp;
  PipeArgs pipe;
};

template <int NF> struct YXInvArgs {
  const float2* in[kMaxFields];  // spectra [z][y][NXP]; NF == 1: `nfields` independent fields, NF > 1: the NF fields of a voxel
  float2* ring;                  // R slots of NF planes
  const float2* tab;
  int nz, nxp, nfields;
  PipeArgs pipe;
};

#ifdef KW_N
// shared-memory layout of the fused kernels: [column exchange | row exchange | row twiddles]
template <int N> struct XYCfg {
  using C = ColCfg<N>;
  using PX = Plan<N>;
  static constexpr int THREADS = C::THREADS;
  static constexpr int T = PX::T;              // threads per row pair
  static constexpr int RP = THREADS / T;       // row pairs per x item
  static constexpr int NTW = PX::NTW > 0 ? PX::NTW : 1;
  static constexpr size_t SM_COL = C::SMEM;
  static constexpr size_t SM_ROW = (size_t)RP * N * sizeof(float2);
  static constexpr size_t SM_TW = (size_t)NTW * T * sizeof(float2);
  static constexpr size_t SMEM = SM_COL + SM_ROW + SM_TW;
  static constexpr int XI = (N / 2 + RP - 1) / RP;  // x items per plane
};

template <int N> __device__ __forceinline__ void xy_build_twiddles(float2* stw, const float2* tab, int tid) {
  using X = XYCfg<N>;
  if (tid < X::T) {
    float2 tw[X::NTW];
    load_twiddles<N>(tw, tid, [tab](int m) { return __ldg(tab + m); });
#pragma unroll
    for (int n = 0; n < X::PX::NTW; ++n) stw[n * X::T + tid] = tw[n];
  }
}

// one column tile: src/dst already point at this thread's first point; points are estride apart
template <int N, int DIR, bool COHERENT, class EX>
__device__ __forceinline__ void y_tile(const float2* __restrict__ src, float2* __restrict__ dst, size_t estride, bool valid, int w, EX& ex) {
  constexpr int E = Plan2<N>::E;
  float2 v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) v[e] = COHERENT ? __ldcg(src + e * estride) : __ldg(src + e * estride);
  fft2_worker<N, DIR>(v, w, ex, ConstTab());
  if (valid) {
#pragma unroll
    for (int e = 0; e < E; ++e) dst[e * estride] = v[e];
  }
}

template <int N> __global__ void __launch_bounds__(XYCfg<N>::THREADS, ColCfg<N>::MINB) k_xy_fwd(XYFwdArgs a) {
  using X = XYCfg<N>;
  using C = ColCfg<N>;
  constexpr int W = C::W, WK = C::WK;
  extern __shared__ float2 smem[];
  __shared__ int s_slot[3];
  float2* const srow = smem + X::SM_COL / sizeof(float2);
  float2* const stw = srow + X::SM_ROW / sizeof(float2);
  const int lane = threadIdx.x, w = threadIdx.y, tz = threadIdx.z;
  const int tid = lane + W * (w + WK * tz);
  const int t = tid % X::T, rp = tid / X::T;
  pipe_prologue(a.pipe, tid, X::THREADS);
  xy_build_twiddles<N>(stw, a.tab, tid);
  __syncthreads();
  ColExchange2<W, C::BAR_THREADS> cex{smem + (size_t)tz * N * W + lane, 1 + tz};
  RowExchange<X::T> rex{srow, rp * N, 1 + rp};
  const size_t plane_c = (size_t)N * a.nxp;  // complex elements of a spectrum plane
  const int ngroups = a.nxp / W;
  PipeItem it;
  while (pipe_next(a.pipe, tid, s_slot, &it)) {
    const int f = it.plane / a.nz, z = it.plane % a.nz;
    float2* const slot = a.ring + (size_t)(it.plane % a.pipe.R) * plane_c;
    if (it.pass == 1) {  // rows of the plane: real -> half spectrum into the ring slot
      float2 twr[X::NTW];
#pragma unroll
      for (int n = 0; n < X::PX::NTW; ++n) twr[n] = stw[n * X::T + t];
      RegTw twp{twr};
      const int pair = it.sub * X::RP + rp;
      const bool valid = pair < N / 2;
      const size_t rowl = 2 * (size_t)(valid ? pair : 0);
      xfwd_rows<N>(a.in[f] + (size_t)z * N * N, slot, a.nxp, rowl, rowl * a.nxp, valid, t, twp, rex);
    } else {  // column tiles of the plane: ring slot -> final spectrum
      const int tile = it.sub * C::TPC + tz;
      const bool valid = tile < ngroups;
      const size_t off = (size_t)w * a.nxp + (size_t)(valid ? tile : 0) * W + lane;
      y_tile<N, -1, true>(slot + off, a.out[f] + (size_t)z * plane_c + off, (size_t)WK * a.nxp, valid, w, cex);
    }
    pipe_done(a.pipe, tid, it);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Row pair of NF fields read from ring planes (coherent loads), inverse transform, epilogue.  rowl = row inside the plane.
template <int N, int NF, class Epi, class EX>
__device__ __forceinline__ void yx_rows(const float2* slot, size_t plane_c, int nxp, const Epi& epi, int field, int z, int rowl, bool valid,
                                        int t, const RegTw& twp, EX& ex) {
  constexpr int T = N / 8;
  float2 res[NF][8];
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    const float2* ia = slot + f * plane_c + (size_t)rowl * nxp;
    const float2* ib = ia + nxp;
    float2 v[1][8];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = t + m * T;
      float2 A = __ldcg(ia + k), B = __ldcg(ib + k);
      if (m == 0 && t == 0) A.y = 0.f, B.y = 0.f;
      v[0][m] = make_float2(A.x - B.y, A.y + B.x);
      if (!(m == 0 && t == 0)) ex.put(0, N - k, make_float2(A.x + B.y, B.x - A.y));
    }
    if (t == 0) {
      const float2 A = __ldcg(ia + N / 2), B = __ldcg(ib + N / 2);
      ex.put(0, N / 2, make_float2(A.x, B.x));
    }
    ex.sync();
#pragma unroll
    for (int m = 4; m < 8; ++m) v[0][m] = ex.get(0, t + m * T);
    ex.sync();
    fft_worker<N, +1, 1>(v, t, 0, twp, ex);
#pragma unroll
    for (int m = 0; m < 8; ++m) res[f][m] = v[0][m];
  }
  if (valid) epi.template apply<N>(res, field, t, (size_t)z * N + rowl, rowl, z);
}

template <int N, int NF, class Epi>
__global__ void __launch_bounds__(XYCfg<N>::THREADS, (ColCfg<N>::MINB < Epi::kMinBlocks ? ColCfg<N>::MINB : Epi::kMinBlocks)) k_yx_inv(YXInvArgs<NF> a, Epi epi) {
  using X = XYCfg<N>;
  using C = ColCfg<N>;
  constexpr int W = C::W, WK = C::WK;
  extern __shared__ float2 smem[];
  __shared__ int s_slot[3];
  float2* const srow = smem + X::SM_COL / sizeof(float2);
  float2* const stw = srow + X::SM_ROW / sizeof(float2);
  const int lane = threadIdx.x, w = threadIdx.y, tz = threadIdx.z;
  const int tid = lane + W * (w + WK * tz);
  const int t = tid % X::T, rp = tid / X::T;
  pipe_prologue(a.pipe, tid, X::THREADS);
  xy_build_twiddles<N>(stw, a.tab, tid);
  __syncthreads();
  ColExchange2<W, C::BAR_THREADS> cex{smem + (size_t)tz * N * W + lane, 1 + tz};
  RowExchange<X::T> rex{srow, rp * N, 1 + rp};
  const size_t plane_c = (size_t)N * a.nxp;
  const int ngroups = a.nxp / W;
  const int tiles_per_field = (ngroups + C::TPC - 1) / C::TPC;  // y items per field plane
  PipeItem it;
  while (pipe_next(a.pipe, tid, s_slot, &it)) {
    // NF == 1: plane index runs over (field, z); NF > 1: over z, a slot holding the NF planes of that z
    const int fsel = (NF == 1) ? it.plane / a.nz : 0, z = (NF == 1) ? it.plane % a.nz : it.plane;
    float2* const slot = a.ring + (size_t)(it.plane % a.pipe.R) * plane_c * NF;
    if (it.pass == 1) {  // column tiles: spectrum plane(s) -> ring slot
      const int f = (NF == 1) ? fsel : it.sub / tiles_per_field;
      const int tile = ((NF == 1) ? it.sub : it.sub % tiles_per_field) * C::TPC + tz;
      const bool valid = tile < ngroups;
      const size_t off = (size_t)w * a.nxp + (size_t)(valid ? tile : 0) * W + lane;
      y_tile<N, +1, false>(a.in[f] + (size_t)z * plane_c + off, slot + (NF == 1 ? 0 : f) * plane_c + off, (size_t)WK * a.nxp, valid, w, cex);
    } else {  // rows: ring slot -> real rows + epilogue
      float2 twr[X::NTW];
#pragma unroll
      for (int n = 0; n < X::PX::NTW; ++n) twr[n] = stw[n * X::T + t];
      RegTw twp{twr};
      const int pair = it.sub * X::RP + rp;
      const bool valid = pair < N / 2;
      yx_rows<N, NF>(slot, plane_c, a.nxp, epi, fsel, z, 2 * (valid ? pair : 0), valid, t, twp, rex);
    }
    pipe_done(a.pipe, tid, it);
  }
}
#endif  // KW_N

}  // namespace kw
