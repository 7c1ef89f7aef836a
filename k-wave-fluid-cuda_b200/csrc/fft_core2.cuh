// Register-resident radix-16 and radix-32 butterflies (compile-time twiddles) and the two-stage column transform
// built on them: N = R1 x R2 with ONE shared-memory exchange per 1-D transform (N = 64..1024), or no exchange at all
// (N = 16, 32).  Compared with the radix-8 Stockham chain of fft_core.cuh this halves the shared-memory instruction
// count and the barriers per transform, which is what bounds the column passes on sm_100a (ncu: mio_throttle).
#pragma once
#include <utility>

#include "fft_core.cuh"

namespace kw {

#define KW_HD __host__ __device__ __forceinline__

struct cf2 { float x, y; };
// forward roots of unity e^{-2 pi i m/32}; W16^m = W32^{2m}
KW_HD constexpr cf2 w32(int m) {
  constexpr cf2 t[32] = {{1.0f, 0.0f}, {0.9807852804032304f, -0.19509032201612825f}, {0.9238795325112867f, -0.3826834323650898f}, {0.8314696123025452f, -0.5555702330196022f}, {0.7071067811865476f, -0.7071067811865475f}, {0.5555702330196023f, -0.8314696123025452f}, {0.38268343236508984f, -0.9238795325112867f}, {0.19509032201612833f, -0.9807852804032304f}, {0.0f, -1.0f}, {-0.1950903220161282f, -0.9807852804032304f}, {-0.3826834323650897f, -0.9238795325112867f}, {-0.555570233019602f, -0.8314696123025455f}, {-0.7071067811865475f, -0.7071067811865476f}, {-0.8314696123025453f, -0.5555702330196022f}, {-0.9238795325112867f, -0.3826834323650899f}, {-0.9807852804032304f, -0.1950903220161286f}, {-1.0f, 0.0f}, {-0.9807852804032304f, 0.19509032201612836f}, {-0.9238795325112868f, 0.38268343236508967f}, {-0.8314696123025455f, 0.555570233019602f}, {-0.7071067811865477f, 0.7071067811865475f}, {-0.5555702330196022f, 0.8314696123025452f}, {-0.38268343236509034f, 0.9238795325112865f}, {-0.19509032201612866f, 0.9807852804032303f}, {0.0f, 1.0f}, {0.1950903220161283f, 0.9807852804032304f}, {0.38268343236509f, 0.9238795325112866f}, {0.5555702330196018f, 0.8314696123025455f}, {0.7071067811865474f, 0.7071067811865477f}, {0.8314696123025452f, 0.5555702330196022f}, {0.9238795325112865f, 0.3826834323650904f}, {0.9807852804032303f, 0.19509032201612872f}};
  return t[m & 31];
}

// compile-time loop: f(std::integral_constant<int, i>) for i in [0, N) -- the index is a constant expression for the
// front end, so twiddles picked with it become immediates (a plain unrolled loop leaves the table in local memory)
template <int... Is, class F> KW_HD void static_for_impl(std::integer_sequence<int, Is...>, F&& f) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F> KW_HD void static_for(F&& f) { static_for_impl(std::make_integer_sequence<int, N>{}, f); }

// The butterflies below run on the device only; they are written with the (packed) complex helpers of fft_core.cuh.
template <int DIR> __device__ __forceinline__ float2 cmulc(float2 a, cf2 w) {  // a * w (forward) or a * conj(w) (inverse), w compile-time
  return cmul(a, make_float2(w.x, DIR < 0 ? w.y : -w.y));
}
template <int DIR> __device__ __forceinline__ void dft4r(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
  const float2 s13 = cadd(a1, a3), d13 = mul_di<DIR>(csub(a1, a3));
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1 = cadd(d02, d13);
  a3 = csub(d02, d13);
}
// 8-point DFT on references, natural order in and out
template <int DIR> __device__ __forceinline__ void dft8r(float2& v0, float2& v1, float2& v2, float2& v3, float2& v4, float2& v5, float2& v6, float2& v7) {
  constexpr float R = 0.70710678118654752440f;
  dft4r<DIR>(v0, v2, v4, v6);
  dft4r<DIR>(v1, v3, v5, v7);
  const float2 o1 = cmul(v3, make_float2(R, DIR < 0 ? -R : R));
  const float2 o3 = cmul(v7, make_float2(-R, DIR < 0 ? -R : R));
  const float2 o0 = v1, o2 = mul_di<DIR>(v5);
  const float2 e0 = v0, e1 = v2, e2 = v4, e3 = v6;
  v0 = cadd(e0, o0);
  v4 = csub(e0, o0);
  v1 = cadd(e1, o1);
  v5 = csub(e1, o1);
  v2 = cadd(e2, o2);
  v6 = csub(e2, o2);
  v3 = cadd(e3, o3);
  v7 = csub(e3, o3);
}

// R-point DFT of v[off + str*j], j < R, natural order in and out (R = 1, 2, 4, 8, 16, 32); all indices compile time.
template <int R, int DIR, int STR, int OFF, int LEN> __device__ __forceinline__ void dftR(float2 (&v)[LEN]) {
#define KW_V(j) v[OFF + STR * (j)]
  if constexpr (R == 2) {
    const float2 t = KW_V(0);
    KW_V(0) = cadd(t, KW_V(1));
    KW_V(1) = csub(t, KW_V(1));
  } else if constexpr (R == 4) {
    dft4r<DIR>(KW_V(0), KW_V(1), KW_V(2), KW_V(3));
  } else if constexpr (R == 8) {
    dft8r<DIR>(KW_V(0), KW_V(1), KW_V(2), KW_V(3), KW_V(4), KW_V(5), KW_V(6), KW_V(7));
  } else if constexpr (R == 16) {
    // 4 (n2) x 4 (n1):  slot n2 + 4 k1 <- DFT4 over n1 of x[4 n1 + n2]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4r<DIR>(KW_V(n2), KW_V(n2 + 4), KW_V(n2 + 8), KW_V(n2 + 12));
    static_for<4>([&](auto K1) {
      static_for<4>([&](auto N2) {
        constexpr int k1 = decltype(K1)::value, n2 = decltype(N2)::value;
        if constexpr (k1 > 0 && n2 > 0) {
          constexpr cf2 tw = w32(2 * n2 * k1);
          KW_V(n2 + 4 * k1) = cmulc<DIR>(KW_V(n2 + 4 * k1), tw);
        }
      });
    });
    // slot k2 + 4 k1 <- X[k1 + 4 k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4r<DIR>(KW_V(4 * k1), KW_V(4 * k1 + 1), KW_V(4 * k1 + 2), KW_V(4 * k1 + 3));
    float2 t[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) t[s] = KW_V(s);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) KW_V(k1 + 4 * k2) = t[k2 + 4 * k1];
  } else if constexpr (R == 32) {
    // 4 (n2) x 8 (n1):  slot n2 + 4 k1 <- DFT8 over n1 of x[4 n1 + n2]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2)
      dft8r<DIR>(KW_V(n2), KW_V(n2 + 4), KW_V(n2 + 8), KW_V(n2 + 12), KW_V(n2 + 16), KW_V(n2 + 20), KW_V(n2 + 24), KW_V(n2 + 28));
    static_for<8>([&](auto K1) {
      static_for<4>([&](auto N2) {
        constexpr int k1 = decltype(K1)::value, n2 = decltype(N2)::value;
        if constexpr (k1 > 0 && n2 > 0) {
          constexpr cf2 tw = w32(n2 * k1);
          KW_V(n2 + 4 * k1) = cmulc<DIR>(KW_V(n2 + 4 * k1), tw);
        }
      });
    });
    // slot k2 + 4 k1 <- X[k1 + 8 k2]
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) dft4r<DIR>(KW_V(4 * k1), KW_V(4 * k1 + 1), KW_V(4 * k1 + 2), KW_V(4 * k1 + 3));
    float2 t[32];
#pragma unroll
    for (int s = 0; s < 32; ++s) t[s] = KW_V(s);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1)
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) KW_V(k1 + 8 * k2) = t[k2 + 4 * k1];
  }
#undef KW_V
}

// ---- two-stage plan ------------------------------------------------------------------------------------------------
template <int N> struct Plan2 {
  static_assert(N >= 16 && N <= 1024 && (N & (N - 1)) == 0, "N must be a power of two in [16,1024]");
  static constexpr int R1 = (N == 16) ? 16 : (N == 32) ? 32 : (N == 64) ? 8 : (N == 128) ? 8 : (N == 256) ? 16 : (N == 512) ? 16 : 32;
  static constexpr int R2 = N / R1;                       // 1 (single stage), 8, 16, 16, 32, 32
  static constexpr int E = (R1 > R2) ? R1 : R2;           // points a worker owns
  static constexpr int WK = N / E;                        // workers per transform
  static constexpr int B1 = E / R1, B2 = (R2 > 1) ? E / R2 : 0;  // butterflies per worker in each stage
};

// One transform's worth of work of worker w: v[e] holds point  w + WK*e  on entry and on exit (natural order).
// EX: put(idx, x) / get(idx) / sync() over the N-point exchange buffer; TAB(m) = forward twiddle e^{-2 pi i m/N}.
struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};
// `hook` runs once every worker has read its stage-2 inputs back, i.e. when the exchange buffer is free again: the
// kernels use it to start the asynchronous copy of the NEXT tile into that same buffer.
template <int N, int DIR, class EX, class TAB, class HOOK = NoHook>
__device__ __forceinline__ void fft2_worker(float2 (&v)[Plan2<N>::E], int w, EX& ex, TAB tab, HOOK hook = HOOK()) {
  using P = Plan2<N>;
  constexpr int R1 = P::R1, R2 = P::R2, E = P::E, WK = P::WK, B1 = P::B1, B2 = P::B2;
  // stage 1: butterfly jb = w + q*WK takes points jb + r*R2 = w + WK*(q + B1*r)  ->  registers q + B1*r
  if constexpr (B1 == 1) dftR<R1, DIR, 1, 0, E>(v);
  else {
    dftR<R1, DIR, B1, 0, E>(v);
    dftR<R1, DIR, B1, 1, E>(v);
    static_assert(B1 <= 2, "at most two butterflies per worker");
  }
  if constexpr (R2 > 1) {
    // outputs of butterfly jb land at jb*R1 + r
#pragma unroll
    for (int q = 0; q < B1; ++q)
#pragma unroll
      for (int r = 0; r < R1; ++r) ex.put((w + q * WK) * R1 + r, v[q + B1 * r]);
    ex.sync();
    // stage 2: butterfly jb = w + q*WK (< R1) takes points jb + r*R1 = w + WK*(q + B2*r), twiddle W_N^{jb*r}
#pragma unroll
    for (int q = 0; q < B2; ++q)
#pragma unroll
      for (int r = 0; r < R2; ++r) v[q + B2 * r] = ex.get(w + q * WK + r * R1);
    ex.sync();  // everyone holds its stage-2 inputs: the exchange buffer is free
    hook();
#pragma unroll
    for (int q = 0; q < B2; ++q) {
      const int jb = w + q * WK;
#pragma unroll
      for (int r = 1; r < R2; ++r) v[q + B2 * r] = cmul(v[q + B2 * r], twd<DIR>(tab((jb * r) & (N - 1))));
    }
    if constexpr (B2 == 1) dftR<R2, DIR, 1, 0, E>(v);
    else {
      dftR<R2, DIR, B2, 0, E>(v);
      dftR<R2, DIR, B2, 1, E>(v);
      static_assert(B2 <= 2, "at most two butterflies per worker");
    }
  } else {
    hook();
  }
}

}  // namespace kw
