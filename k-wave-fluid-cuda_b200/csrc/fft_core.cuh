// FP32 Stockham FFT building blocks for sm_100a (no cuFFT).
//
// Every 1-D transform of length N = 8^a * b (b in {1,2,4}; N = 16..1024) is computed by N/8 "workers"; a worker owns
// 8 complex values per stage in registers and exchanges them with the other workers of the same transform through
// shared memory between stages.  The first stage reads its inputs straight from registers (filled from global memory
// by the caller) and the last stage leaves its outputs in registers, element index  t + m*N/8  for worker t,
// register m -- the same pattern on input and on output, so a forward transform, a pointwise multiply and an inverse
// transform chain in registers (used by the fused z pass), and a C2R output feeds an R2C input directly.
//
// Replaces MatrixClasses/CufftComplexMatrix.cpp (cufftExecR2C / cufftExecC2R call sites :511,:527) in the reference.
#pragma once
#include <cuda_runtime.h>

namespace kw {

// asynchronous global -> shared copies (no registers, no exposed latency; completion per issuing thread)
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int K> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(K) : "memory"); }



// Complex arithmetic on the packed FP32x2 pipe of sm_100 (FADD2 / FMUL2 / FFMA2: one instruction per complex add, two
// per complex multiply; half swaps, per-half negation and scalar broadcast are operand modifiers, so the make_float2
// shuffles below cost nothing).  The butterflies are issue-bound, not FLOP-bound: halving their instruction count is what
// shortens them.  Measured on B200 (profiles/r01_l_*): the fused z pass of N = 512 gains 4-5 % (1.97 -> 1.88 ms per step),
// but the radix-32 x 32 plan of N = 1024 LOSES 30 % (2.29 -> 1.60 TB/s) and N = 256 loses 3 %: with two warps per scheduler
// the passes are bound by dependent-issue latency, and the packed instructions have the longer one.  Hence packed forms
// only in the N = 512 translation unit (fft_inst.cu is compiled once per length); KW_PACKED=0/1 overrides.
#ifndef KW_PACKED
#ifdef KW_N
#define KW_PACKED (KW_N == 512)
#else
#define KW_PACKED 0
#endif
#endif
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
#if KW_PACKED
  return __ffma2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
#else
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
#endif
}
#if KW_PACKED
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
#endif
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by DIR * i  (DIR = -1: forward e^{-i..}, multiply by -i;  DIR = +1: inverse, multiply by +i)
template <int DIR> __device__ __forceinline__ float2 mul_di(float2 a) {
  return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
// twiddle from a forward table (e^{-2 pi i m / N}); inverse direction conjugates
template <int DIR> __device__ __forceinline__ float2 twd(float2 w) { return DIR < 0 ? w : cconj(w); }

template <int DIR> __device__ __forceinline__ void dft2(float2& a, float2& b) {
  float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

template <int DIR> __device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
  float2 s13 = cadd(a1, a3), d13 = mul_di<DIR>(csub(a1, a3));
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  a1 = cadd(d02, d13);
  a3 = csub(d02, d13);
}

// 8-point DFT, natural order in / natural order out
template <int DIR> __device__ __forceinline__ void dft8(float2 (&v)[8]) {
  constexpr float R = 0.70710678118654752440f;
  // even / odd 4-point transforms
  dft4<DIR>(v[0], v[2], v[4], v[6]);
  dft4<DIR>(v[1], v[3], v[5], v[7]);
  // odd outputs times W8^k
  // odd outputs times W8^k:  W8 = (1 - i)/sqrt2, W8^3 = (-1 - i)/sqrt2 (conjugated for the inverse)
  const float2 o1 = cmul(v[3], make_float2(R, DIR < 0 ? -R : R));
  const float2 o3 = cmul(v[7], make_float2(-R, DIR < 0 ? -R : R));
  float2 o0 = v[1], o2 = mul_di<DIR>(v[5]);
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  v[0] = cadd(e0, o0);
  v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1);
  v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2);
  v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3);
  v[7] = csub(e3, o3);
}

// ---- plan: N = 8^NS8 * TAIL ----------------------------------------------------------------------------------------
template <int N> struct Plan {
  static_assert(N >= 16 && N <= 1024 && (N & (N - 1)) == 0, "N must be a power of two in [16,1024]");
  static constexpr int LOG2 = (N == 16) ? 4 : (N == 32) ? 5 : (N == 64) ? 6 : (N == 128) ? 7 : (N == 256) ? 8
                              : (N == 512) ? 9 : 10;
  static constexpr int NS8 = LOG2 / 3;                 // number of radix-8 stages
  static constexpr int TAIL = 1 << (LOG2 - 3 * NS8);   // 1, 2 or 4
  static constexpr int T = N / 8;                      // workers per transform
  static constexpr int NTW = (NS8 - 1) * 7 + (TAIL == 2 ? 4 : TAIL == 4 ? 6 : 0);  // twiddles a worker needs
};

// Twiddle providers.  A provider answers  get(n, m): n = running index of the twiddle inside the plan (for register
// files filled by load_twiddles), m = index into the forward table e^{-2 pi i m/N}.
struct RegTw {  // loop-invariant twiddles kept in registers (row transforms: every lane has its own set)
  const float2* tw;
  __device__ __forceinline__ float2 get(int n, int) const { return tw[n]; }
};

// Fill a register file in plan order for worker t.  tab(m) returns the forward table entry m.
template <int N, class Tab> __device__ __forceinline__ void load_twiddles(float2* tw, int t, Tab tab) {
  using P = Plan<N>;
  int n = 0;
  int ns = 8;
#pragma unroll
  for (int s = 1; s < P::NS8; ++s) {
    const int k = t & (ns - 1);
    const int stride = N / (8 * ns);
#pragma unroll
    for (int r = 1; r < 8; ++r) tw[n++] = tab(k * r * stride);
    ns *= 8;
  }
  if (P::TAIL == 2) {
#pragma unroll
    for (int q = 0; q < 4; ++q) tw[n++] = tab(t + q * (N / 8));
  } else if (P::TAIL == 4) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int jb = t + q * (N / 8);
#pragma unroll
      for (int r = 1; r < 4; ++r) tw[n++] = tab((jb * r) & (N - 1));
    }
  }
}

// B transforms' worth of work for one thread: worker ids t0 + b*TS, registers v[b][0..7].  EX provides
// put(b, idx, x) / get(b, idx) / sync() over the exchange buffers; TW provides get(n, m) (see above).
template <int N, int DIR, int B, class EX, class TW>
__device__ __forceinline__ void fft_worker(float2 (&v)[B][8], int t0, int ts, const TW& twp, EX& ex) {
  using P = Plan<N>;
  constexpr int T = P::T;
  int n = 0;
  int ns = 1;
#pragma unroll
  for (int s = 0; s < P::NS8; ++s) {
    if (s > 0) {
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const int t = t0 + b * ts;
#pragma unroll
        for (int r = 0; r < 8; ++r) v[b][r] = ex.get(b, t + r * T);
      }
    }
    const bool last = (s == P::NS8 - 1) && (P::TAIL == 1);
    if (!last && s > 0) ex.sync();  // everyone has read its inputs of this stage
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int t = t0 + b * ts;
      const int k = t & (ns - 1);
      if (s > 0) {
        const int stride = N / (8 * ns);
#pragma unroll
        for (int r = 1; r < 8; ++r) v[b][r] = cmul(v[b][r], twd<DIR>(twp.get(n + b * P::NTW + r - 1, k * r * stride)));
      }
      dft8<DIR>(v[b]);
      if (!last) {
        const int base = ((t - k) << 3) + k;  // (t/ns)*ns*8 + k
#pragma unroll
        for (int r = 0; r < 8; ++r) ex.put(b, base + r * ns, v[b][r]);
      }
    }
    if (s > 0) n += 7;
    if (!last) ex.sync();
    ns *= 8;
  }
  if (P::TAIL == 2) {
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int t = t0 + b * ts;
#pragma unroll
      for (int r = 0; r < 8; ++r) v[b][r] = ex.get(b, t + r * T);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[b][q + 4] = cmul(v[b][q + 4], twd<DIR>(twp.get(n + b * P::NTW + q, t + q * T)));
        dft2<DIR>(v[b][q], v[b][q + 4]);
      }
    }
    ex.sync();  // the buffer may be reused by the caller
  } else if (P::TAIL == 4) {
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int t = t0 + b * ts;
#pragma unroll
      for (int r = 0; r < 8; ++r) v[b][r] = ex.get(b, t + r * T);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int jb = t + q * T;
#pragma unroll
        for (int r = 1; r < 4; ++r)
          v[b][q + 2 * r] = cmul(v[b][q + 2 * r], twd<DIR>(twp.get(n + b * P::NTW + q * 3 + r - 1, (jb * r) & (N - 1))));
        dft4<DIR>(v[b][q], v[b][q + 2], v[b][q + 4], v[b][q + 6]);
      }
    }
    ex.sync();
  } else {
    // all-radix-8 plan: the last stage wrote nothing, but its gets must be complete before the buffer is reused
    if (P::NS8 > 1) ex.sync();
  }
}

// ---- exchange policies -------------------------------------------------------------------------------------------
// Row transforms (x axis): consecutive lanes are consecutive workers of one transform.  Flat float2 buffer of all
// transforms of the CTA with an XOR swizzle that makes the three Stockham access patterns (contiguous, stride 8,
// 8-contiguous/stride 64) conflict free per half warp (verified by tools/check_swizzle.py).
// T workers (threads) cooperate on one transform and synchronise among themselves only: a named barrier per transform
// when it spans whole warps, __syncwarp when several transforms share a warp -- the transforms of a CTA drift apart so
// that the load, butterfly and store phases of different rows overlap.
template <int T> struct RowExchange {
  float2* buf;  // CTA buffer
  int base;     // transform index * N
  int bar;      // named barrier id of this transform (1..15)
  __device__ __forceinline__ static int swz(int g) { return g ^ ((g >> 3) & 7) ^ (((g >> 6) & 1) << 3); }
  __device__ __forceinline__ void put(int, int i, float2 x) { buf[swz(base + i)] = x; }
  __device__ __forceinline__ float2 get(int, int i) const { return buf[swz(base + i)]; }
  __device__ __forceinline__ void sync() {
    if (T >= 32) asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(T) : "memory");
    else __syncwarp();
  }
};

// Column transforms (y / z axes): W consecutive lanes hold W neighbouring kx of the same worker, so element i of
// lane l lives at buf[i*W + l]: every access is W contiguous float2 -- conflict free without any swizzle.
// All B transforms of a thread share one N*W tile (they are workers of the same transforms).
template <int W, int NTHREADS = 0> struct ColExchange {
  float2* buf;  // tile buffer + lane
  int bar;      // named barrier of this tile slot (1..15); slots of a CTA synchronise independently
  __device__ __forceinline__ void put(int, int i, float2 x) { buf[i * W] = x; }
  __device__ __forceinline__ float2 get(int, int i) const { return buf[i * W]; }
  __device__ __forceinline__ void sync() {
    if (NTHREADS == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(NTHREADS) : "memory");
  }
};

}  // namespace kw
