// (device code and launchers of fft_generic.cu / fft_generic_ct.cu)
// Transform lengths outside the tuned table (fft_inst.cu: powers of two in [16, 1024]): the same four kernel shapes -- and the same
// launch table (ops.h) -- for any length N = 8 * m whose prime factors are 2, 3, 5 and 7 (96, 120, 160, 192, 240, 480, 768, ...), so
// that the grids cuFFT accepts in the reference (MatrixClasses/CufftComplexMatrix.cpp:87-91 plans whatever Nx, Ny, Nz the input file
// holds) run here too.  One set of kernels with the length and its radix list as RUN-TIME arguments:
//
//   * every transform is a Stockham autosort FFT in shared memory (decimation in frequency, radix 4 / 2 / 3 / 5 / 7 passes ping-ponging
//     between two buffers); twiddles come from a per-CTA table e^{-2 pi i m / N} computed once per CTA in double precision;
//   * rows (x axis): a CTA transforms RP row pairs at a time (two real rows = one complex transform, as in k_xfwd / k_xinv); the
//     inverse hands thread t the points x = t + m * N/8 of both rows, i.e. exactly the layout the fused real-space epilogues
//     (solver_kernels.cuh) are written for -- they are called with N = 0 = "length at run time";
//   * columns (y, z axes): tiles of W neighbouring kx times all N points; the fused z pass applies the k-space operator between the
//     forward and the inverse transform like k_zmid.
//
// These kernels move the same bytes as the tuned ones but make one round trip through shared memory per radix pass instead of one
// register exchange per transform.  Every kernel exists in two forms: CN = 0, length and radix list at run time (any supported length),
// and CN = N for the common lengths listed in fft_generic_ct.cu, where radices, strides and trip counts are constants and the butterfly
// loops unroll (384^3: 12.9 -> 9.1 ms per step, the speed of the reference's cuFFT build).
#pragma once
#include <map>
#include <mutex>

#include "ops.h"

namespace kw {
namespace generic {

constexpr int kThreads = 256;
constexpr int kMaxFactors = 12;

struct GenPlan {
  int n;
  int nf;
  unsigned char r[kMaxFactors];
};

// Radices of a pass: 8 (= 4 x 2 in registers) as long as three factors of two remain, then 4, or 6 (= 3 x 2) / 2 for a single leftover
// two; 3, 5, 7 for the rest.  480 = 8 4 3 5, 384 = 8 8 6, 240 = 8 6 5, 1536 = 8 8 8 3: every pass is one round trip through shared memory.
static bool make_plan(int n, GenPlan* pl) {
  pl->n = n, pl->nf = 0;
  if (n < 16 || n > 2048 || n % 8) return false;
  int m = n, e[8] = {};
  for (int r : {2, 3, 5, 7})
    while (m % r == 0) m /= r, ++e[r];
  if (m != 1) return false;
  auto push = [&](int r) { pl->r[pl->nf++] = (unsigned char)r; };
  for (; e[2] >= 3; e[2] -= 3) push(8);
  if (e[2] == 2) push(4);
  if (e[2] == 1) {
    if (e[3] > 0) push(6), --e[3];
    else push(2);
  }
  for (int r : {3, 5, 7})
    for (; e[r] > 0; --e[r]) push(r);
  return pl->nf <= kMaxFactors;
}
// kx values per column tile: two tile buffers of at most 64 KB, so that three CTAs (24 warps) share an SM -- with one 123 KB CTA per SM
// the first version of these kernels issued on 26 % of the cycles and waited on memory for the rest (ncu, N = 480, 16-wide tiles)
static int tile_w(int n) { return n <= 256 ? 16 : n <= 512 ? 8 : 4; }

// ---- device side ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 gmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <int DIR> __device__ __forceinline__ float2 tw_dir(float2 w) { return DIR < 0 ? w : make_float2(w.x, -w.y); }
// i * a (DIR > 0) or -i * a (DIR < 0)
template <int DIR> __device__ __forceinline__ float2 rot(float2 a) { return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 fma2(float c, float2 a, float2 b) { return make_float2(fmaf(c, a.x, b.x), fmaf(c, a.y, b.y)); }

// forward table e^{-2 pi i m / n}, m < n, one copy per CTA
__device__ __forceinline__ void fill_table(float2* tab, int n) {
  for (int m = threadIdx.x; m < n; m += blockDim.x) {
    double s, c;
    sincospi(-2.0 * (double)m / (double)n, &s, &c);
    tab[m] = make_float2((float)c, (float)s);
  }
}

// DFT of R points in registers (DIR < 0: forward, e^{-2 pi i jk/R}); nr_r = N / R indexes w_R in the table (R = 7 only)
template <int R, int DIR> __device__ __forceinline__ void small_dft(float2 (&v)[R], const float2* tab, int nr_r) {
  if constexpr (R == 2) {
    const float2 a = v[0], b = v[1];
    v[0] = add2(a, b), v[1] = sub2(a, b);
  } else if constexpr (R == 3) {
    const float2 t1 = add2(v[1], v[2]);
    const float2 t2 = fma2(-0.5f, t1, v[0]);
    const float2 d = sub2(v[1], v[2]);
    const float2 t3 = rot<DIR>(make_float2(0.86602540378443865f * d.x, 0.86602540378443865f * d.y));  // -+ i sqrt(3)/2 (a1 - a2)
    v[0] = add2(v[0], t1), v[1] = add2(t2, t3), v[2] = sub2(t2, t3);
  } else if constexpr (R == 4) {
    const float2 s02 = add2(v[0], v[2]), d02 = sub2(v[0], v[2]), s13 = add2(v[1], v[3]), jd = rot<DIR>(sub2(v[1], v[3]));
    v[0] = add2(s02, s13), v[1] = add2(d02, jd), v[2] = sub2(s02, s13), v[3] = sub2(d02, jd);
  } else if constexpr (R == 5) {
    constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f, s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
    const float2 t1 = add2(v[1], v[4]), t2 = add2(v[2], v[3]), t3 = sub2(v[1], v[4]), t4 = sub2(v[2], v[3]);
    const float2 m1 = fma2(c2, t2, fma2(c1, t1, v[0])), m2 = fma2(c1, t2, fma2(c2, t1, v[0]));
    const float2 n1 = rot<DIR>(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
    const float2 n2 = rot<DIR>(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
    v[0] = add2(v[0], add2(t1, t2));
    v[1] = add2(m1, n1), v[4] = sub2(m1, n1), v[2] = add2(m2, n2), v[3] = sub2(m2, n2);
  } else if constexpr (R > 1) {  // 7: the R x R sum with w_R^{jk} = tab[((j k) mod R) nr_r]
    float2 a[R], w[R];
#pragma unroll
    for (int k = 0; k < R; ++k) a[k] = v[k];
#pragma unroll
    for (int k = 1; k < R; ++k) w[k] = tw_dir<DIR>(tab[k * nr_r]);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      float2 acc = a[0];
#pragma unroll
      for (int k = 1; k < R; ++k) acc = add2(acc, (j * k) % R == 0 ? a[k] : gmul(a[k], w[(j * k) % R ? (j * k) % R : 1]));
      v[j] = acc;
    }
  }
}

// One butterfly of radix R = RA * RB of a Stockham pass (decimation in frequency) over a transform of n points stored with element stride ES.
// s = product of the radices of the earlier passes; butterfly t in [0, n / R): p = t / s, q = t % s, ob = q + s R p, nr = n / R;
//   inputs x[t + nr k], k < R;   outputs y[ob + s j] = (sum_k x_k w_R^{jk}) W_n^{p j s}, j < R.
// RB > 1 (radix 8 = 4 x 2, 6 = 3 x 2): the R-point DFT runs in registers as RB transforms of RA points (k = RB ka + kb, over ka), the
// twiddles w_R^{kb ja}, and RA transforms of RB points (over kb) giving j = ja + RA jb -- one round trip through shared memory where two
// passes would make two.
template <int RA, int RB, int ES, int DIR>
__device__ __forceinline__ void butterfly(const float2* __restrict__ x, float2* __restrict__ y, int nr, int s, int t, int p, int ob, const float2* __restrict__ tab) {
  constexpr int R = RA * RB;
  float2 a[R];
#pragma unroll
  for (int k = 0; k < R; ++k) a[k] = x[(t + k * nr) * ES];
#pragma unroll
  for (int kb = 0; kb < RB; ++kb) {
    float2 v[RA];
#pragma unroll
    for (int ka = 0; ka < RA; ++ka) v[ka] = a[RB * ka + kb];
    small_dft<RA, DIR>(v, tab, nr * RB);
#pragma unroll
    for (int ja = 0; ja < RA; ++ja) a[RB * ja + kb] = (RB > 1 && kb * ja) ? gmul(v[ja], tw_dir<DIR>(tab[kb * ja * nr])) : v[ja];
  }
  const int ps = p * s;
#pragma unroll
  for (int ja = 0; ja < RA; ++ja) {
    float2 v[RB];
#pragma unroll
    for (int kb = 0; kb < RB; ++kb) v[kb] = a[RB * ja + kb];
    if constexpr (RB > 1) small_dft<RB, DIR>(v, tab, nr * RA);
#pragma unroll
    for (int jb = 0; jb < RB; ++jb) {
      const int j = ja + RA * jb;
      y[(ob + j * s) * ES] = j ? gmul(v[jb], tw_dir<DIR>(tab[ps * j])) : v[jb];  // W_n^{p j s}
    }
  }
}

// ---- the same passes with the length known at compile time ---------------------------------------------------------------------
// A handful of common lengths (kCommon below) get their own instantiation of every kernel: radices, strides, trip counts and the
// divisions of the index arithmetic become constants and the butterfly loop of a thread unrolls, so that the shared-memory loads of
// its butterflies are in flight together.  Same radix order as make_plan (the two forms agree to rounding; both are tested against the DFT).
template <int M> struct NextRadix {  // M = product of the radices still to do
  static constexpr int value = (M % 8 == 0) ? 8 : (M % 4 == 0) ? 4 : (M % 2 == 0) ? (M % 3 == 0 ? 6 : 2) : (M % 3 == 0) ? 3 : (M % 5 == 0) ? 5 : 7;
};
template <int CN, int M, int S, int W, int DIR, int BD, int CB> struct CtPasses {
  static __device__ __forceinline__ float2* run(float2* a, float2* b, const float2* tab) {
    constexpr int R = NextRadix<M>::value, NB = CN / R, ES = W ? W : 1;
    constexpr int RA = R == 8 ? 4 : R == 6 ? 3 : R, RB = (R == 8 || R == 6) ? 2 : 1;
    constexpr int TOTAL = (W ? W : CB) * NB, TRIPS = (TOTAL + BD - 1) / BD;
#pragma unroll
    for (int k = 0; k < TRIPS; ++k) {
      const int idx = (int)threadIdx.x + k * BD;
      if (TOTAL % BD == 0 || idx < TOTAL) {
        int t, off;
        if constexpr (W == 0) {
          const int bi = idx / NB;
          t = idx - bi * NB, off = bi * CN;
        } else {
          t = idx / W, off = idx % W;
        }
        const int p = t / S, q = t - p * S;
        butterfly<RA, RB, ES, DIR>(a + off, b + off, NB, S, t, p, q + S * R * p, tab);
      }
    }
    __syncthreads();
    return CtPasses<CN, M / R, S * R, W, DIR, BD, CB>::run(b, a, tab);
  }
};
template <int CN, int S, int W, int DIR, int BD, int CB> struct CtPasses<CN, 1, S, W, DIR, BD, CB> {
  static __device__ __forceinline__ float2* run(float2* a, float2*, const float2*) { return a; }
};
// compile-time counterparts of rows_per_cta / the x-inverse block shape / tile_w (host side below)
template <int CN> struct CtCfg {
  static constexpr int RP_FWD = CN == 0 ? 0 : (2048 / (CN ? CN : 1) < 1 ? 1 : 2048 / (CN ? CN : 1) > 16 ? 16 : 2048 / (CN ? CN : 1));
  static constexpr int T = CN / 8;
  static constexpr int RP_INV = CN == 0 ? 0 : (kThreads / (T ? T : 1) < 1 ? 1 : kThreads / (T ? T : 1));
  static constexpr int BD_INV = T * RP_INV;
  static constexpr int W = CN <= 256 ? 16 : CN <= 512 ? 8 : 4;
};

// `batch` transforms at once.  Returns the buffer that holds the result (natural order).  Every thread of the CTA must call it; ends with
// a barrier.  W > 0: column tile, transform c at buf[c + i * W]; W == 0: rows, transform b at buf[b * n + i].
template <int W, int DIR, int CN = 0, int BD = 0, int CB = 0>
__device__ __forceinline__ float2* fft_batch(float2* a, float2* b, int batch, const GenPlan& pl, const float2* tab) {
  if constexpr (CN > 0) return CtPasses<CN, CN, 1, W, DIR, BD, CB>::run(a, b, tab);
  constexpr int ES = W ? W : 1;
  int s = 1;
  for (int f = 0; f < pl.nf; ++f) {
    const int r = pl.r[f], nb = pl.n / r;
    const float inv_s = 1.0f / (float)s, inv_nb = 1.0f / (float)nb;
    const int total = (W ? W : batch) * nb;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      int t, off;
      if constexpr (W == 0) {  // rows: consecutive threads take consecutive butterflies of one transform
        const int bi = __float2int_rz(((float)idx + 0.5f) * inv_nb);  // idx / nb, exact below 2^20
        t = idx - bi * nb, off = bi * pl.n;
      } else {  // column tiles: consecutive threads take the same butterfly of consecutive kx
        t = idx / W, off = idx % W;
      }
      const int p = __float2int_rz(((float)t + 0.5f) * inv_s), q = t - p * s;
      const int ob = q + s * r * p;
      const float2* x = a + off;
      float2* y = b + off;
      switch (r) {
        case 8: butterfly<4, 2, ES, DIR>(x, y, nb, s, t, p, ob, tab); break;
        case 6: butterfly<3, 2, ES, DIR>(x, y, nb, s, t, p, ob, tab); break;
        case 4: butterfly<4, 1, ES, DIR>(x, y, nb, s, t, p, ob, tab); break;
        case 2: butterfly<2, 1, ES, DIR>(x, y, nb, s, t, p, ob, tab); break;
        case 3: butterfly<3, 1, ES, DIR>(x, y, nb, s, t, p, ob, tab); break;
        case 5: butterfly<5, 1, ES, DIR>(x, y, nb, s, t, p, ob, tab); break;
        default: butterfly<7, 1, ES, DIR>(x, y, nb, s, t, p, ob, tab); break;
      }
    }
    __syncthreads();
    float2* tmp = a;
    a = b, b = tmp;
    s *= r;
  }
  return a;
}

// idx / d for idx < 2^20 through a float reciprocal (a 32-bit integer division costs ~20 instructions)
__device__ __forceinline__ int fdiv(int idx, float inv_d) { return __float2int_rz(((float)idx + 0.5f) * inv_d); }

__device__ __forceinline__ size_t row_off(const RowMap& map, size_t row, int nxp) { return map.off(row, nxp); }

// real rows -> half spectra (k_xfwd)
template <int CN> static __global__ void __launch_bounds__(kThreads) g_xfwd(XFwdArgs a, GenPlan pl, int rp_rt) {
  extern __shared__ float2 gsm[];
  const int n = CN ? CN : pl.n, rp = CN ? CtCfg<CN>::RP_FWD : rp_rt;
  float2* tab = gsm;
  float2* A = gsm + n;
  float2* B = A + (size_t)rp * n;
  float2* C = B + (size_t)rp * n;  // the rows of the CTA's next group land here (cp.async) while this group is transformed
  fill_table(tab, n);
  const float* __restrict__ in = a.in[blockIdx.y];
  float2* __restrict__ out = a.out[blockIdx.y];
  const int npairs = a.pair_end - a.pair_begin;
  const int nxr = n / 2 + 1;
  const float inv_n = 1.0f / (float)n, inv_nxp = 1.0f / (float)a.nxp;
  auto fetch = [&](int g, float2* dst) {  // two real rows -> real / imaginary parts of one complex row
    if (g * rp < npairs) {
      const int pair0 = a.pair_begin + g * rp;
      for (int idx = threadIdx.x; idx < rp * n; idx += blockDim.x) {
        const int b = fdiv(idx, inv_n), x = idx - b * n;
        const int pair = pair0 + b;
        if (pair < a.pair_end) {
          const float* r0 = in + 2 * (size_t)pair * n;
          cp_async4(&dst[idx].x, r0 + x);
          cp_async4(&dst[idx].y, r0 + n + x);
        } else {
          dst[idx] = make_float2(0.f, 0.f);
        }
      }
    }
    cp_async_commit();
  };
  fetch(blockIdx.x, A);
  __syncthreads();
  for (int g = blockIdx.x; g * rp < npairs; g += gridDim.x) {
    const int pair0 = a.pair_begin + g * rp;
    fetch(g + gridDim.x, C);
    cp_async_wait<1>();
    __syncthreads();
    const float2* Z = fft_batch<0, -1, CN, kThreads, CtCfg<CN>::RP_FWD>(A, B, rp, pl, tab);
    for (int idx = threadIdx.x; idx < rp * a.nxp; idx += blockDim.x) {
      const int b = fdiv(idx, inv_nxp), k = idx - b * a.nxp;
      const int pair = pair0 + b;
      if (pair >= a.pair_end) continue;
      const size_t off = row_off(a.map, 2 * (size_t)pair, a.nxp);
      float2 va = make_float2(0.f, 0.f), vb = va;  // padding columns: zero
      if (k < nxr) {
        const float2 z = Z[b * n + k], q = Z[b * n + (k ? n - k : 0)];
        va = make_float2(0.5f * (z.x + q.x), 0.5f * (z.y - q.y));
        vb = make_float2(0.5f * (z.y + q.y), -0.5f * (z.x - q.x));
      }
      out[off + k] = va;
      out[off + a.nxp + k] = vb;
    }
    __syncthreads();
    float2* t = A;
    A = C, C = B, B = t;
  }
  cp_async_wait<0>();
}

// half spectra -> real rows + epilogue (k_xinv).  blockDim = (T = n / 8, rp): thread (t, b) ends up with the points x = t + m T of
// row pair b, the layout of the epilogue contract (fft_kernels.cuh).
template <int CN, int NF, class Epi> static __global__ void __launch_bounds__(kThreads, Epi::kMinBlocks) g_xinv(XInvArgs<NF> a, Epi epi, GenPlan pl) {
  extern __shared__ float2 gsm[];
  const int n = CN ? CN : pl.n, T = n / 8, rp = CN ? CtCfg<CN>::RP_INV : (int)blockDim.x / T;
  float2* tab = gsm;
  float2* A = gsm + n;
  float2* B = A + (size_t)rp * n;
  // the half-spectrum rows of the NEXT transform (next field of the group, or the first field of the CTA's next group) are copied into a
  // staging buffer (cp.async) while the current one runs: [pair][row a | row b][n/2 + 1]
  float2* st_cur = B + (size_t)rp * n;
  float2* st_nxt = st_cur + (size_t)rp * (n + 2);
  fill_table(tab, n);
  const int npairs = a.pair_end - a.pair_begin;
  const int field = blockIdx.y + a.field0;
  const int b_own = threadIdx.x / T, t_own = threadIdx.x - b_own * T;
  const int half = n / 2;
  const float inv_h1 = 1.0f / (float)(half + 1);
  auto fetch = [&](int g, int f, float2* st) {
    if (g * rp < npairs) {
      const float2* __restrict__ in = NF == 1 ? a.in[field] : a.in[f];
      const int pair0 = a.pair_begin + g * rp;
      for (int idx = threadIdx.x; idx < rp * (half + 1); idx += blockDim.x) {
        const int b = fdiv(idx, inv_h1), k = idx - b * (half + 1);
        const int pair = pair0 + b;
        float2* sa = st + (size_t)b * (n + 2) + k;
        if (pair < a.pair_end) {
          const size_t off = row_off(a.map, 2 * (size_t)pair, a.nxp);
          cp_async8(sa, in + off + k);
          cp_async8(sa + half + 1, in + off + a.nxp + k);
        } else {
          sa[0] = sa[half + 1] = make_float2(0.f, 0.f);
        }
      }
    }
    cp_async_commit();
  };
  fetch(blockIdx.x, 0, st_cur);
  __syncthreads();
  for (int g = blockIdx.x; g * rp < npairs; g += gridDim.x) {
    const int pair0 = a.pair_begin + g * rp;
    float2 res[NF][8];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      if (f + 1 < NF) fetch(g, f + 1, st_nxt);
      else fetch(g + gridDim.x, 0, st_nxt);
      cp_async_wait<1>();
      __syncthreads();
      for (int idx = threadIdx.x; idx < rp * (half + 1); idx += blockDim.x) {
        const int b = fdiv(idx, inv_h1), k = idx - b * (half + 1);
        const float2* sa = st_cur + (size_t)b * (n + 2) + k;
        const float2 va = sa[0], vb = sa[half + 1];
        float2* Z = A + (size_t)b * n;
        if (k == 0) {
          Z[0] = make_float2(va.x, vb.x);  // C2R ignores the imaginary part of DC ...
        } else if (k == half) {
          Z[half] = make_float2(va.x, vb.x);  // ... and of the Nyquist bin
        } else {
          Z[k] = make_float2(va.x - vb.y, va.y + vb.x);      // A + iB
          Z[n - k] = make_float2(va.x + vb.y, vb.x - va.y);  // conj(A) + i conj(B)
        }
      }
      __syncthreads();
      const float2* R = fft_batch<0, +1, CN, CtCfg<CN>::BD_INV, CtCfg<CN>::RP_INV>(A, B, rp, pl, tab);
      if (b_own < rp) {
#pragma unroll
        for (int m = 0; m < 8; ++m) res[f][m] = R[(size_t)b_own * n + t_own + m * T];
      }
      __syncthreads();
      float2* t = st_cur;
      st_cur = st_nxt, st_nxt = t;
    }
    const int pair = pair0 + b_own;
    if (b_own < rp && pair < a.pair_end) {
      const size_t row0 = 2 * (size_t)pair;
      const int y = (int)(row0 % a.ny), z = (int)(row0 / a.ny);
      epi.template apply<0>(res, field, t_own, row0, y, z, nullptr, n);
    }
  }
  cp_async_wait<0>();
}

// in-place complex transform along y or z of [..][..][NXP] (k_col)
// Three tile buffers: while a tile is transformed between two of them, the next tile of the CTA is on its way into the third one
// (cp.async) -- with synchronous loads the kernel waited 7.8 cycles per issued instruction on the tile load (ncu, N = 480).
template <int CN, int W> static __global__ void __launch_bounds__(kThreads) g_col(ColArgs a, GenPlan pl, int dir) {
  extern __shared__ float2 gsm[];
  const int n = CN ? CN : pl.n;
  float2* tab = gsm;
  float2* cur = gsm + n;
  float2* oth = cur + (size_t)n * W;
  float2* nxt = oth + (size_t)n * W;
  fill_table(tab, n);
  float2* __restrict__ data = a.data[blockIdx.y];
  auto tile_base = [&](int tile) { return data + (size_t)(tile / a.ngroups) * a.outer_stride + (size_t)(tile % a.ngroups) * W; };
  // slab-decomposed runs: the y axis is split into blocks of nyl points, one per rank that owns them after the exchange (RowMap)
  auto point = [&](int i) -> size_t {
    if (a.nyl == 0) return (size_t)i * a.stride;
    const int q = i / a.nyl;
    return (size_t)q * a.blk + (size_t)(i - q * a.nyl) * a.stride;
  };
  auto fetch = [&](int tile, float2* dst) {
    if (tile < a.tile_end) {
      const float2* base = tile_base(tile);
#pragma unroll 4
      for (int idx = threadIdx.x; idx < n * W; idx += kThreads) {
        const int i = idx / W, c = idx - i * W;
        cp_async8(dst + idx, base + point(i) + c);
      }
    }
    cp_async_commit();  // (an empty group when there is no further tile: the wait below counts groups)
  };
  int tile = a.tile_begin + blockIdx.x;
  fetch(tile, cur);
  for (; tile < a.tile_end; tile += gridDim.x) {
    fetch(tile + gridDim.x, nxt);
    cp_async_wait<1>();  // this tile has landed; the next one may still be in flight
    __syncthreads();
    const float2* R = dir < 0 ? fft_batch<W, -1, CN, kThreads, W>(cur, oth, W, pl, tab) : fft_batch<W, +1, CN, kThreads, W>(cur, oth, W, pl, tab);
    float2* base = tile_base(tile);
#pragma unroll 4
    for (int idx = threadIdx.x; idx < n * W; idx += kThreads) {
      const int i = idx / W, c = idx - i * W;
      base[point(i) + c] = R[idx];
    }
    __syncthreads();
    float2* t = cur;
    cur = nxt, nxt = oth, oth = t;
  }
  cp_async_wait<0>();
}

// forward z -> k-space operator -> inverse z (k_zmid; same operator semantics, axis as a run-time argument)
// one field, one operator (axis -1 .. 2): the CTA's next tile travels into a third buffer while this one is processed (a fourth buffer for
// the multiplier would leave one CTA per SM at N = 480; it is read at the operator step instead)
template <int CN, int W> static __global__ void __launch_bounds__(kThreads) g_zmid(ZMidArgs a, GenPlan pl) {
  extern __shared__ float2 gsm[];
  const int n = CN ? CN : pl.n, axis = a.axis;
  float2* tab = gsm;
  float2* cur = gsm + n;
  float2* oth = cur + (size_t)n * W;
  float2* nxt = oth + (size_t)n * W;
  fill_table(tab, n);
  const float2* __restrict__ in = a.f.in;
  const float* __restrict__ mul = a.f.mul;
  const float scal = a.f.scal;
  auto fetch = [&](int tile, float2* dst) {
    if (tile < a.ntiles) {
      const size_t base = (size_t)(tile / a.ngroups) * a.nxp + (size_t)(tile % a.ngroups) * W;
#pragma unroll 4
      for (int idx = threadIdx.x; idx < n * W; idx += kThreads) {
        const int i = idx / W, c = idx - i * W;
        cp_async8(dst + idx, in + base + (size_t)i * a.plane + c);
      }
    }
    cp_async_commit();
  };
  int tile = blockIdx.x;
  fetch(tile, cur);
  for (; tile < a.ntiles; tile += gridDim.x) {
    fetch(tile + gridDim.x, nxt);
    cp_async_wait<1>();
    __syncthreads();
    const int y = tile / a.ngroups, kx0 = (tile % a.ngroups) * W;
    const size_t base = (size_t)y * a.nxp + kx0;
    float2* S = fft_batch<W, -1, CN, kThreads, W>(cur, oth, W, pl, tab);
    float2* O = S == cur ? oth : cur;
#pragma unroll 4  // several multiplier loads in flight per thread
    for (int idx = threadIdx.x; idx < n * W; idx += kThreads) {
      const int i = idx / W, c = idx - i * W;
      const float m = mul ? __ldg(mul + base + (size_t)i * a.plane + c) * scal : scal;
      float2 v = S[idx];
      v = make_float2(v.x * m, v.y * m);
      if (axis == 0) v = gmul(v, __ldg(a.f.vec + kx0 + c));
      else if (axis == 1) v = gmul(v, __ldg(a.f.vec + y));
      else if (axis == 2) v = gmul(v, __ldg(a.f.vec + i));
      S[idx] = v;
    }
    __syncthreads();
    const float2* R = fft_batch<W, +1, CN, kThreads, W>(S, O, W, pl, tab);
#pragma unroll 4
    for (int idx = threadIdx.x; idx < n * W; idx += kThreads) {
      const int i = idx / W, c = idx - i * W;
      a.f.out[base + (size_t)i * a.plane + c] = R[idx];
    }
    __syncthreads();
    float2* t = cur;
    cur = nxt, nxt = oth, oth = t;
  }
  cp_async_wait<0>();
}

// gradient form (axis 3): one forward transform feeds three inverse ones
template <int CN, int W> static __global__ void __launch_bounds__(kThreads) g_zmid_grad(ZMidArgs a, GenPlan pl) {
  extern __shared__ float2 gsm[];
  const int n = CN ? CN : pl.n, axis = a.axis;
  float2* tab = gsm;
  float2* A = gsm + n;
  float2* B = A + (size_t)n * W;
  float2* E = B + (size_t)n * W;  // gradient only: the spectrum times the multiplier, kept for the three operators
  fill_table(tab, n);
  const float2* __restrict__ in = a.f.in;
  const float* __restrict__ mul = a.f.mul;
  const float scal = a.f.scal;
  __syncthreads();
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int y = tile / a.ngroups, kx0 = (tile % a.ngroups) * W;
    const size_t base = (size_t)y * a.nxp + kx0;
#pragma unroll 4
    for (int idx = threadIdx.x; idx < n * W; idx += kThreads) {
      const int i = idx / W, c = idx - i * W;
      A[idx] = __ldg(in + base + (size_t)i * a.plane + c);
    }
    __syncthreads();
    float2* S = fft_batch<W, -1, CN, kThreads, W>(A, B, W, pl, tab);
    float2* O = S == A ? B : A;  // the other buffer
    if (axis == 3) {
#pragma unroll 4
      for (int idx = threadIdx.x; idx < n * W; idx += kThreads) {
        const int i = idx / W, c = idx - i * W;
        const float m = mul ? __ldg(mul + base + (size_t)i * a.plane + c) * scal : scal;
        const float2 v = S[idx];
        E[idx] = make_float2(v.x * m, v.y * m);
      }
      __syncthreads();
      for (int f = 0; f < 3; ++f) {
#pragma unroll 4
        for (int idx = threadIdx.x; idx < n * W; idx += kThreads) {
          const int i = idx / W, c = idx - i * W;
          const float2 w = f == 0 ? __ldg(a.f.vec + kx0 + c) : f == 1 ? __ldg(a.f.vec_y + y) : __ldg(a.f.vec_z + i);
          S[idx] = gmul(E[idx], w);
        }
        __syncthreads();
        const float2* R = fft_batch<W, +1, CN, kThreads, W>(S, O, W, pl, tab);
        float2* __restrict__ out = f == 0 ? a.f.out : f == 1 ? a.f.out_y : a.f.out_z;
#pragma unroll 4
        for (int idx = threadIdx.x; idx < n * W; idx += kThreads) {
          const int i = idx / W, c = idx - i * W;
          out[base + (size_t)i * a.plane + c] = R[idx];
        }
        __syncthreads();
      }
    } else {
#pragma unroll 4
      for (int idx = threadIdx.x; idx < n * W; idx += kThreads) {
        const int i = idx / W, c = idx - i * W;
        const float m = mul ? __ldg(mul + base + (size_t)i * a.plane + c) * scal : scal;
        float2 v = S[idx];
        v = make_float2(v.x * m, v.y * m);
        if (axis == 0) v = gmul(v, __ldg(a.f.vec + kx0 + c));
        else if (axis == 1) v = gmul(v, __ldg(a.f.vec + y));
        else if (axis == 2) v = gmul(v, __ldg(a.f.vec + i));
        S[idx] = v;
      }
      __syncthreads();
      const float2* R = fft_batch<W, +1, CN, kThreads, W>(S, O, W, pl, tab);
#pragma unroll 4
      for (int idx = threadIdx.x; idx < n * W; idx += kThreads) {
        const int i = idx / W, c = idx - i * W;
        a.f.out[base + (size_t)i * a.plane + c] = R[idx];
      }
      __syncthreads();
    }
  }
}

// ---- host side: the launch table ----------------------------------------------------------------------------------------
template <class K> static void opt_in(K kernel, size_t smem) {
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
static int grid_for(int work, int per_sm) {
  const int cap = sm_count() * per_sm;
  return work < cap ? (work > 0 ? work : 1) : cap;
}
static int ctas_per_sm(size_t smem) {
  const int k = (int)((size_t)220 * 1024 / (smem + 1024));
  return k < 1 ? 1 : k > 4 ? 4 : k;
}
static GenPlan plan_of(int n) {
  GenPlan pl{};
  make_plan(n, &pl);
  return pl;
}
static int rows_per_cta(int n) {
  const int rp = 2048 / n;
  return rp < 1 ? 1 : rp > 16 ? 16 : rp;
}

template <int CN> static void xfwd(const XFwdArgs& a, int nfields, cudaStream_t st) {
  const GenPlan pl = plan_of(a.n);
  const int rp = rows_per_cta(a.n);
  const size_t smem = ((size_t)a.n + 3 * (size_t)rp * a.n) * sizeof(float2);
  opt_in(g_xfwd<CN>, smem);
  const int groups = (a.pair_end - a.pair_begin + rp - 1) / rp;
  g_xfwd<CN><<<dim3(grid_for(groups, 4), nfields), kThreads, smem, st>>>(a, pl, rp);
}
template <int CN, int NF, class Epi> static void xinv(const XInvArgs<NF>& a, const Epi& e, int gy, cudaStream_t st) {
  const GenPlan pl = plan_of(a.n);
  const int T = a.n / 8;
  const int rp = kThreads / T < 1 ? 1 : kThreads / T;
  const size_t smem = ((size_t)a.n + 2 * (size_t)rp * a.n + 2 * (size_t)rp * (a.n + 2)) * sizeof(float2);
  opt_in(g_xinv<CN, NF, Epi>, smem);
  const int groups = (a.pair_end - a.pair_begin + rp - 1) / rp;
  g_xinv<CN, NF, Epi><<<dim3(grid_for(groups, 4), gy), T * rp, smem, st>>>(a, e, pl);
}
template <int CN> static void xinv_store(const XInvArgs<1>& a, const EpiStore& e, int nfields, cudaStream_t st) { xinv<CN, 1>(a, e, nfields, st); }
template <int CN> static void xinv_add(const XInvArgs<1>& a, const EpiAdd& e, cudaStream_t st) { xinv<CN, 1>(a, e, 1, st); }
template <int CN> static void xinv_velocity(const XInvArgs<1>& a, const EpiVelocity& e, int nfields, cudaStream_t st) { xinv<CN, 1>(a, e, nfields, st); }
template <int CN> static void xinv_density(const XInvArgs<3>& a, const EpiDensity& e, cudaStream_t st) { xinv<CN, 3>(a, e, 1, st); }
template <int CN> static void xinv_psum(const XInvArgs<2>& a, const EpiPressureSum& e, cudaStream_t st) { xinv<CN, 2>(a, e, 1, st); }

template <int CN> static void col(const ColArgs& a, int dir, int nfields, cudaStream_t st) {
  const GenPlan pl = plan_of(a.n);
  const int W = tile_w(a.n);
  const size_t smem = ((size_t)a.n + 3 * (size_t)a.n * W) * sizeof(float2);
  const dim3 grid(grid_for(a.tile_end - a.tile_begin, ctas_per_sm(smem)), nfields);
  auto go = [&](auto kernel) {
    opt_in(kernel, smem);
    kernel<<<grid, kThreads, smem, st>>>(a, pl, dir);
  };
  if constexpr (CN > 0) go(g_col<CN, CtCfg<CN>::W>);
  else W == 16 ? go(g_col<0, 16>) : W == 8 ? go(g_col<0, 8>) : go(g_col<0, 4>);
}
template <int CN> static void zmid(const ZMidArgs& a, cudaStream_t st) {
  const GenPlan pl = plan_of(a.n);
  const int W = tile_w(a.n);
  // gradient: tile, ping-pong partner, kept spectrum; other forms: tile, partner, next tile
  const size_t smem = ((size_t)a.n + 3 * (size_t)a.n * W) * sizeof(float2);
  const int grid = grid_for(a.ntiles, ctas_per_sm(smem));
  auto go = [&](auto kernel) {
    opt_in(kernel, smem);
    kernel<<<grid, kThreads, smem, st>>>(a, pl);
  };
  if constexpr (CN > 0) {
    if (a.axis == 3) go(g_zmid_grad<CN, CtCfg<CN>::W>);
    else go(g_zmid<CN, CtCfg<CN>::W>);
  } else {
    if (a.axis == 3) W == 16 ? go(g_zmid_grad<0, 16>) : W == 8 ? go(g_zmid_grad<0, 8>) : go(g_zmid_grad<0, 4>);
    else W == 16 ? go(g_zmid<0, 16>) : W == 8 ? go(g_zmid<0, 8>) : go(g_zmid<0, 4>);
  }
}
template <int CN> static FftOps make_ops(int n) {
  const int w = tile_w(n);
  return FftOps{n, w, 1, w, xfwd<CN>, xinv_store<CN>, xinv_add<CN>, xinv_velocity<CN>, xinv_density<CN>, xinv_psum<CN>, col<CN>, zmid<CN>};
}
}  // namespace generic
}  // namespace kw
