// Long column transforms (N = 512, 1024) with 16 points per thread: N = 16 x (16 G), G = 2 (N = 512) or 4 (N = 1024).
//
// The two-stage plan of fft_core2.cuh gives a worker 32 points (N = 512: 16 x 32, N = 1024: 32 x 32); a thread that holds
// 32 complex values needs ~250 registers, so the fused z pass ran with 8 warps per SM and was bound by dependent-issue
// latency (ncu, profiles/r01_k_k_zmid.summary.txt: 12.5 % warps active, 43 % issue slots).  Here a thread holds 16 points
// (<= 128 registers, 16 warps per SM): stage 1 is a register-resident radix-16 butterfly, and the 16G-point butterflies
// of stage 2 are shared by G threads OF ONE WARP -- each runs a radix-16 butterfly on every G-th input, then the G partial
// results are combined with a radix-G butterfly across the threads through warp shuffles.  One shared-memory exchange
// per transform, as before.
//
// Thread layout: blockDim = (W, WK, 1) with W = 32 / G neighbouring kx per warp row and WK = 16 G workers, so that the G
// workers w = G j2 + h (h = 0..G-1) of butterfly j2 sit in one warp at lanes  lane_kx + W h.
//
// Data ownership (what makes the fused z pass cheap): the forward transform takes the natural worker order
// (v[e] = x[w + WK e]) and LEAVES its outputs in the order the butterflies produce them,
//     v[i + B g]  =  conj(W_G)^(h g) X[k],   k = j2 + 16 ((i + B h) mod 16 + 16 g),   B = 16 / G, i < B, g < G,
// and the inverse transform accepts exactly that order (phase included) and returns the natural worker order.  A pointwise
// operator between the two (k_zmid) commutes with the per-thread phase, so neither the reordering nor the phase costs
// anything.  paired_k() gives k for the operator lookups.
//
// Replaces the z stages of cufftExecR2C / cufftExecC2R (MatrixClasses/CufftComplexMatrix.cpp:511,527) for Nz = 512, 1024
// inside k_zmid; index algebra checked against the DFT definition in tools/check_paired_plan.py.
#pragma once
#include "fft_core2.cuh"

namespace kw {

template <int N> struct Plan3 {
  static_assert(N == 512 || N == 1024, "paired plan: N = 512 or 1024");
  static constexpr int E = 16;            // points per thread
  static constexpr int WK = N / 16;       // workers per transform (32, 64)
  static constexpr int G = WK / 16;       // threads sharing a stage-2 butterfly (2, 4)
  static constexpr int B = 16 / G;        // outputs of one frequency group a thread ends up with (8, 4)
  static constexpr int W = 32 / G;        // kx values per tile (16, 8): G workers x W lanes = one warp
  static constexpr int THREADS = W * WK;  // 512
};

// frequency index of register slot `s` (= i + B g) of worker w
template <int N> __device__ __forceinline__ int paired_k(int w, int s) {
  using P = Plan3<N>;
  const int h = w & (P::G - 1), j2 = w / P::G;
  const int i = s % P::B, g = s / P::B;
  return j2 + 16 * (((i + P::B * h) & 15) + 16 * g);
}
// paired_k(w, s) = paired_k0(w) + paired_ks(s) as long as i + B h < 16, which always holds (i < B, h < G)
template <int N> __device__ __forceinline__ int paired_k0(int w) {
  using P = Plan3<N>;
  return w / P::G + 16 * P::B * (w & (P::G - 1));
}
template <int N> __host__ __device__ constexpr int paired_ks(int s) { return 16 * (s % Plan3<N>::B) + 256 * (s / Plan3<N>::B); }

__device__ __forceinline__ float2 shfl2(float2 v, int src_lane) {
  return make_float2(__shfl_sync(0xffffffffu, v.x, src_lane), __shfl_sync(0xffffffffu, v.y, src_lane));
}

// W_{16G}^{m}: compile-time for G = 2 (W32), from the length-N constant table for G = 4 (W64^m = W_N^{16 m})
template <int N, int DIR, class TAB> __device__ __forceinline__ float2 group_twiddle(int h, int kp, TAB tab) {
  constexpr int G = Plan3<N>::G;
  const float2 t = tab(((h * kp) * (N / (16 * G))) & (N - 1));
  return DIR < 0 ? t : cconj(t);
}

// radix-G butterfly over the thread index of a stage-2 group, in rotated order: u[d] belongs to thread (h + d) mod G.
// DIR < 0: p[g] = sum_d W_G^{d g} u[d];  DIR > 0: conjugated.
template <int G, int DIR> __device__ __forceinline__ void dft_group(float2 (&u)[G]) {
  if constexpr (G == 2) {
    const float2 a = u[0];
    u[0] = cadd(a, u[1]);
    u[1] = csub(a, u[1]);
  } else {
    dft4r<DIR>(u[0], u[1], u[2], u[3]);
  }
}

// Forward transform, see the header for the output order.  `hook` runs when the exchange buffer is free again.
template <int N, class EX, class TAB, class HOOK = NoHook>
__device__ __forceinline__ void paired_fwd(float2 (&v)[16], int w, EX& ex, TAB tab, HOOK hook = HOOK()) {
  using P = Plan3<N>;
  constexpr int G = P::G, B = P::B, W = P::W;
  const int h = w & (G - 1), j2 = w / G;
  const int lane = threadIdx.x + W * h;  // lane of this thread inside its warp
  dftR<16, -1, 1, 0, 16>(v);             // stage 1 over e -> k1
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) ex.put(w * 16 + k1, v[k1]);
  ex.sync();
#pragma unroll
  for (int rp = 0; rp < 16; ++rp) v[rp] = ex.get((G * rp + h) * 16 + j2);
  ex.sync();
  hook();
  // twiddle W_N^{j2 r}, r = G r' + h, times the rotation W_N^{(N/G) h r'} that moves this thread's own output block to slots 0..B-1
#pragma unroll
  for (int rp = 0; rp < 16; ++rp) v[rp] = cmul(v[rp], tab((j2 * (G * rp + h) + (N / G) * h * rp) & (N - 1)));
  dftR<16, -1, 1, 0, 16>(v);             // over r' -> slot s, k' = (s + B h) mod 16
  if (h) {
    if constexpr (G == 2) {
      static_for<16>([&](auto S) {
        constexpr int s = decltype(S)::value;
        v[s] = cmulc<-1>(v[s], w32((s + 8) & 15));
      });
    } else {
#pragma unroll
      for (int s = 0; s < 16; ++s) v[s] = cmul(v[s], group_twiddle<N, -1>(h, (s + B * h) & 15, tab));
    }
  }
  // all-to-all inside the group: slot block (G - d) mod G goes to thread (h - d) mod G, i.e. every thread receives, from
  // thread (h + d) mod G, that thread's partial sums for the k' block this thread owns
  float2 u[B][G];
#pragma unroll
  for (int i = 0; i < B; ++i) u[i][0] = v[i];
#pragma unroll
  for (int d = 1; d < G; ++d)
#pragma unroll
    for (int i = 0; i < B; ++i) u[i][d] = shfl2(v[B * (G - d) + i], (lane + W * d) & 31);
#pragma unroll
  for (int i = 0; i < B; ++i) {
    dft_group<G, -1>(u[i]);
#pragma unroll
    for (int g = 0; g < G; ++g) v[i + B * g] = u[i][g];
  }
}

// Inverse transform: input in the order paired_fwd leaves (phase included), output v[e] = x[w + WK e], unnormalised.
template <int N, class EX, class TAB, class HOOK = NoHook>
__device__ __forceinline__ void paired_inv(float2 (&v)[16], int w, EX& ex, TAB tab, HOOK hook = HOOK()) {
  using P = Plan3<N>;
  constexpr int G = P::G, B = P::B, W = P::W;
  const int h = w & (G - 1), j2 = w / G;
  const int lane = threadIdx.x + W * h;
  float2 u[B][G];
#pragma unroll
  for (int i = 0; i < B; ++i) {
#pragma unroll
    for (int g = 0; g < G; ++g) u[i][g] = v[i + B * g];
    dft_group<G, +1>(u[i]);  // u[i][d] = partial result for thread (h + d) mod G
    v[i] = u[i][0];
  }
#pragma unroll
  for (int d = 1; d < G; ++d)
#pragma unroll
    for (int i = 0; i < B; ++i) v[B * (G - d) + i] = shfl2(u[i][d], (lane - W * d) & 31);
  if (h) {
    if constexpr (G == 2) {
      static_for<16>([&](auto S) {
        constexpr int s = decltype(S)::value;
        v[s] = cmulc<+1>(v[s], w32((s + 8) & 15));
      });
    } else {
#pragma unroll
      for (int s = 0; s < 16; ++s) v[s] = cmul(v[s], group_twiddle<N, +1>(h, (s + B * h) & 15, tab));
    }
  }
  dftR<16, +1, 1, 0, 16>(v);  // slots -> r'
#pragma unroll
  for (int rp = 0; rp < 16; ++rp) v[rp] = cmul(v[rp], cconj(tab((j2 * (G * rp + h) + (N / G) * h * rp) & (N - 1))));
#pragma unroll
  for (int rp = 0; rp < 16; ++rp) ex.put((G * rp + h) * 16 + j2, v[rp]);
  ex.sync();
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) v[k1] = ex.get(w * 16 + k1);
  ex.sync();
  hook();
  dftR<16, +1, 1, 0, 16>(v);  // k1 -> e
}

}  // namespace kw
