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
// These kernels move the same bytes as the tuned ones but make log_r(N) round trips through shared memory per transform instead of one
// register exchange: a correct fallback at a fraction of the tuned speed, not a second fast path.
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

static bool make_plan(int n, GenPlan* pl) {
  pl->n = n, pl->nf = 0;
  if (n < 16 || n > 2048 || n % 8) return false;
  int m = n;
  auto take = [&](int r) {
    while (m % r == 0 && pl->nf < kMaxFactors) pl->r[pl->nf++] = (unsigned char)r, m /= r;
  };
  take(4), take(2), take(3), take(5), take(7);
  return m == 1;
}
// kx values per column tile: two tile buffers of at most 64 KB, so that three CTAs (24 warps) share an SM -- with one 123 KB CTA per SM
// the first version of these kernels issued on 26 % of the cycles and waited on memory for the rest (ncu, N = 480, 16-wide tiles)
static int tile_w(int n) { return n <= 256 ? 16 : n <= 512 ? 8 : 4; }

// ---- device side ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 gmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 tw_dir(float2 w, int dir) { return dir < 0 ? w : make_float2(w.x, -w.y); }

// forward table e^{-2 pi i m / n}, m < n, one copy per CTA
__device__ __forceinline__ void fill_table(float2* tab, int n) {
  for (int m = threadIdx.x; m < n; m += blockDim.x) {
    double s, c;
    sincospi(-2.0 * (double)m / (double)n, &s, &c);
    tab[m] = make_float2((float)c, (float)s);
  }
}

// radix 3 / 5 / 7: the R x R DFT written out with the table entries w_R^{jk} = tab[((j k) mod R) n / R]
template <int R>
__device__ __forceinline__ void odd_butterfly(const float2* __restrict__ x, float2* __restrict__ y, int es, int nr, int s, int t, int p, int ob,
                                              const float2* __restrict__ tab, int dir) {
  float2 a[R], w[R];
#pragma unroll
  for (int k = 0; k < R; ++k) a[k] = x[(size_t)(t + k * nr) * es];
#pragma unroll
  for (int k = 1; k < R; ++k) w[k] = tw_dir(tab[k * nr], dir);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float2 acc = a[0];
#pragma unroll
    for (int k = 1; k < R; ++k) {
      const float2 wk = w[(j * k) % R ? (j * k) % R : 1];
      if ((j * k) % R == 0) acc.x += a[k].x, acc.y += a[k].y;
      else acc.x += a[k].x * wk.x - a[k].y * wk.y, acc.y += a[k].x * wk.y + a[k].y * wk.x;
    }
    y[(size_t)(ob + j * s) * es] = j ? gmul(acc, tw_dir(tab[p * j * s], dir)) : acc;
  }
}

// One radix-r butterfly of a Stockham pass over a transform of n points stored with element stride es.
// s = product of the radices of the earlier passes; butterfly t in [0, n / r): p = t / s, q = t % s;
//   inputs  x[t + (n / r) k],  k < r          outputs  y[q + s (r p + j)] = (sum_k x_k w_r^{jk}) W_n^{p j s},  j < r
__device__ __forceinline__ void butterfly(const float2* __restrict__ x, float2* __restrict__ y, int es, int n, int s, float inv_s, int r, int t,
                                          const float2* __restrict__ tab, int dir) {
  const int p = __float2int_rz(((float)t + 0.5f) * inv_s), q = t - p * s;  // t / s, exact for t < 2^20
  const int nr = n / r;
  const int ob = q + s * r * p;
  if (r == 4) {
    const float2 a0 = x[(size_t)t * es], a1 = x[(size_t)(t + nr) * es], a2 = x[(size_t)(t + 2 * nr) * es], a3 = x[(size_t)(t + 3 * nr) * es];
    const float2 s02 = make_float2(a0.x + a2.x, a0.y + a2.y), d02 = make_float2(a0.x - a2.x, a0.y - a2.y);
    const float2 s13 = make_float2(a1.x + a3.x, a1.y + a3.y), d13 = make_float2(a1.x - a3.x, a1.y - a3.y);
    // forward: -i d13 = (d13.y, -d13.x); inverse: +i d13 = (-d13.y, d13.x)
    const float2 jd = dir < 0 ? make_float2(d13.y, -d13.x) : make_float2(-d13.y, d13.x);
    const int ti = p * s;
    y[(size_t)ob * es] = make_float2(s02.x + s13.x, s02.y + s13.y);
    y[(size_t)(ob + s) * es] = gmul(make_float2(d02.x + jd.x, d02.y + jd.y), tw_dir(tab[ti], dir));
    y[(size_t)(ob + 2 * s) * es] = gmul(make_float2(s02.x - s13.x, s02.y - s13.y), tw_dir(tab[2 * ti], dir));
    y[(size_t)(ob + 3 * s) * es] = gmul(make_float2(d02.x - jd.x, d02.y - jd.y), tw_dir(tab[3 * ti], dir));
    return;
  }
  if (r == 2) {
    const float2 a0 = x[(size_t)t * es], a1 = x[(size_t)(t + nr) * es];
    y[(size_t)ob * es] = make_float2(a0.x + a1.x, a0.y + a1.y);
    y[(size_t)(ob + s) * es] = gmul(make_float2(a0.x - a1.x, a0.y - a1.y), tw_dir(tab[p * s], dir));
    return;
  }
  if (r == 3) odd_butterfly<3>(x, y, es, nr, s, t, p, ob, tab, dir);
  else if (r == 5) odd_butterfly<5>(x, y, es, nr, s, t, p, ob, tab, dir);
  else odd_butterfly<7>(x, y, es, nr, s, t, p, ob, tab, dir);
}

// `batch` transforms at once.  Transform b, point i at buf[b * bs + i * es].  Returns the buffer that holds the result
// (natural order).  Every thread of the CTA must call it; ends with a barrier.
// W > 0: column tile, batch = es = W (compile time), bs = 1; W == 0: rows, es = 1, bs = n.
template <int W>
__device__ __forceinline__ float2* fft_batch(float2* a, float2* b, int batch, const GenPlan& pl, const float2* tab, int dir) {
  int s = 1;
  for (int f = 0; f < pl.nf; ++f) {
    const int r = pl.r[f], nb = pl.n / r;
    const float inv_s = 1.0f / (float)s;
    if constexpr (W == 0) {  // rows: consecutive threads take consecutive butterflies of one transform
      const float inv_nb = 1.0f / (float)nb;
      for (int idx = threadIdx.x; idx < batch * nb; idx += blockDim.x) {
        const int bi = __float2int_rz(((float)idx + 0.5f) * inv_nb), t = idx - bi * nb;
        butterfly(a + (size_t)bi * pl.n, b + (size_t)bi * pl.n, 1, pl.n, s, inv_s, r, t, tab, dir);
      }
    } else {  // column tiles: consecutive threads take the same butterfly of consecutive kx
      for (int idx = threadIdx.x; idx < W * nb; idx += blockDim.x) {
        const int t = idx / W, bi = idx % W;
        butterfly(a + bi, b + bi, W, pl.n, s, inv_s, r, t, tab, dir);
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
static __global__ void __launch_bounds__(kThreads) g_xfwd(XFwdArgs a, GenPlan pl, int rp) {
  extern __shared__ float2 gsm[];
  const int n = pl.n;
  float2* tab = gsm;
  float2* A = gsm + n;
  float2* B = A + (size_t)rp * n;
  fill_table(tab, n);
  const float* __restrict__ in = a.in[blockIdx.y];
  float2* __restrict__ out = a.out[blockIdx.y];
  const int npairs = a.pair_end - a.pair_begin;
  const int nxr = n / 2 + 1;
  const float inv_n = 1.0f / (float)n, inv_nxp = 1.0f / (float)a.nxp;
  __syncthreads();
  for (int g = blockIdx.x; g * rp < npairs; g += gridDim.x) {
    const int pair0 = a.pair_begin + g * rp;
    for (int idx = threadIdx.x; idx < rp * n; idx += blockDim.x) {
      const int b = fdiv(idx, inv_n), x = idx - b * n;
      const int pair = pair0 + b;
      float2 v = make_float2(0.f, 0.f);
      if (pair < a.pair_end) {
        const float* r0 = in + 2 * (size_t)pair * n;
        v = make_float2(__ldg(r0 + x), __ldg(r0 + n + x));
      }
      A[idx] = v;
    }
    __syncthreads();
    const float2* Z = fft_batch<0>(A, B, rp, pl, tab, -1);
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
  }
}

// half spectra -> real rows + epilogue (k_xinv).  blockDim = (T = n / 8, rp): thread (t, b) ends up with the points x = t + m T of
// row pair b, the layout of the epilogue contract (fft_kernels.cuh).
template <int NF, class Epi> static __global__ void __launch_bounds__(kThreads) g_xinv(XInvArgs<NF> a, Epi epi, GenPlan pl) {
  extern __shared__ float2 gsm[];
  const int n = pl.n, T = n / 8, rp = blockDim.x / T;  // (threads beyond T * rp only help with the butterflies)
  float2* tab = gsm;
  float2* A = gsm + n;
  float2* B = A + (size_t)rp * n;
  fill_table(tab, n);
  const int npairs = a.pair_end - a.pair_begin;
  const int field = blockIdx.y + a.field0;
  const int b_own = threadIdx.x / T, t_own = threadIdx.x - b_own * T;
  const int half = n / 2;
  const float inv_h1 = 1.0f / (float)(half + 1);
  __syncthreads();
  for (int g = blockIdx.x; g * rp < npairs; g += gridDim.x) {
    const int pair0 = a.pair_begin + g * rp;
    float2 res[NF][8];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const float2* __restrict__ in = NF == 1 ? a.in[field] : a.in[f];
      for (int idx = threadIdx.x; idx < rp * (half + 1); idx += blockDim.x) {
        const int b = fdiv(idx, inv_h1), k = idx - b * (half + 1);
        const int pair = pair0 + b;
        float2 va = make_float2(0.f, 0.f), vb = va;
        if (pair < a.pair_end) {
          const size_t off = row_off(a.map, 2 * (size_t)pair, a.nxp);
          va = __ldg(in + off + k), vb = __ldg(in + off + a.nxp + k);
        }
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
      const float2* R = fft_batch<0>(A, B, rp, pl, tab, +1);
      if (b_own < rp) {
#pragma unroll
        for (int m = 0; m < 8; ++m) res[f][m] = R[(size_t)b_own * n + t_own + m * T];
      }
      __syncthreads();
    }
    const int pair = pair0 + b_own;
    if (b_own < rp && pair < a.pair_end) {
      const size_t row0 = 2 * (size_t)pair;
      const int y = (int)(row0 % a.ny), z = (int)(row0 / a.ny);
      epi.template apply<0>(res, field, t_own, row0, y, z, nullptr, n);
    }
  }
}

// in-place complex transform along y or z of [..][..][NXP] (k_col)
template <int W> static __global__ void __launch_bounds__(kThreads) g_col(ColArgs a, GenPlan pl, int dir) {
  extern __shared__ float2 gsm[];
  const int n = pl.n;
  float2* tab = gsm;
  float2* A = gsm + n;
  float2* B = A + (size_t)n * W;
  fill_table(tab, n);
  float2* __restrict__ data = a.data[blockIdx.y];
  __syncthreads();
  for (int tile = a.tile_begin + blockIdx.x; tile < a.tile_end; tile += gridDim.x) {
    float2* base = data + (size_t)(tile / a.ngroups) * a.outer_stride + (size_t)(tile % a.ngroups) * W;
    for (int idx = threadIdx.x; idx < n * W; idx += blockDim.x) {
      const int i = idx / W, c = idx - i * W;
      A[idx] = base[(size_t)i * a.stride + c];
    }
    __syncthreads();
    const float2* R = fft_batch<W>(A, B, W, pl, tab, dir);
    for (int idx = threadIdx.x; idx < n * W; idx += blockDim.x) {
      const int i = idx / W, c = idx - i * W;
      base[(size_t)i * a.stride + c] = R[idx];
    }
    __syncthreads();
  }
}

// forward z -> k-space operator -> inverse z (k_zmid; same operator semantics, axis as a run-time argument)
template <int W> static __global__ void __launch_bounds__(kThreads) g_zmid(ZMidArgs a, GenPlan pl) {
  extern __shared__ float2 gsm[];
  const int n = pl.n, axis = a.axis;
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
    for (int idx = threadIdx.x; idx < n * W; idx += blockDim.x) {
      const int i = idx / W, c = idx - i * W;
      A[idx] = __ldg(in + base + (size_t)i * a.plane + c);
    }
    __syncthreads();
    float2* S = fft_batch<W>(A, B, W, pl, tab, -1);
    float2* O = S == A ? B : A;  // the other buffer
    if (axis == 3) {
      for (int idx = threadIdx.x; idx < n * W; idx += blockDim.x) {
        const int i = idx / W, c = idx - i * W;
        const float m = mul ? __ldg(mul + base + (size_t)i * a.plane + c) * scal : scal;
        const float2 v = S[idx];
        E[idx] = make_float2(v.x * m, v.y * m);
      }
      __syncthreads();
      for (int f = 0; f < 3; ++f) {
        for (int idx = threadIdx.x; idx < n * W; idx += blockDim.x) {
          const int i = idx / W, c = idx - i * W;
          const float2 w = f == 0 ? __ldg(a.f.vec + kx0 + c) : f == 1 ? __ldg(a.f.vec_y + y) : __ldg(a.f.vec_z + i);
          S[idx] = gmul(E[idx], w);
        }
        __syncthreads();
        const float2* R = fft_batch<W>(S, O, W, pl, tab, +1);
        float2* __restrict__ out = f == 0 ? a.f.out : f == 1 ? a.f.out_y : a.f.out_z;
        for (int idx = threadIdx.x; idx < n * W; idx += blockDim.x) {
          const int i = idx / W, c = idx - i * W;
          out[base + (size_t)i * a.plane + c] = R[idx];
        }
        __syncthreads();
      }
    } else {
      for (int idx = threadIdx.x; idx < n * W; idx += blockDim.x) {
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
      const float2* R = fft_batch<W>(S, O, W, pl, tab, +1);
      for (int idx = threadIdx.x; idx < n * W; idx += blockDim.x) {
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

static void xfwd(const XFwdArgs& a, int nfields, cudaStream_t st) {
  const GenPlan pl = plan_of(a.n);
  const int rp = rows_per_cta(a.n);
  const size_t smem = ((size_t)a.n + 2 * (size_t)rp * a.n) * sizeof(float2);
  opt_in(g_xfwd, smem);
  const int groups = (a.pair_end - a.pair_begin + rp - 1) / rp;
  g_xfwd<<<dim3(grid_for(groups, 4), nfields), kThreads, smem, st>>>(a, pl, rp);
}
template <int NF, class Epi> static void xinv(const XInvArgs<NF>& a, const Epi& e, int gy, cudaStream_t st) {
  const GenPlan pl = plan_of(a.n);
  const int T = a.n / 8;
  const int rp = kThreads / T < 1 ? 1 : kThreads / T;
  const size_t smem = ((size_t)a.n + 2 * (size_t)rp * a.n) * sizeof(float2);
  opt_in(g_xinv<NF, Epi>, smem);
  const int groups = (a.pair_end - a.pair_begin + rp - 1) / rp;
  g_xinv<NF, Epi><<<dim3(grid_for(groups, 4), gy), T * rp, smem, st>>>(a, e, pl);
}
static void xinv_store(const XInvArgs<1>& a, const EpiStore& e, int nfields, cudaStream_t st) { xinv<1>(a, e, nfields, st); }
static void xinv_add(const XInvArgs<1>& a, const EpiAdd& e, cudaStream_t st) { xinv<1>(a, e, 1, st); }
static void xinv_velocity(const XInvArgs<1>& a, const EpiVelocity& e, int nfields, cudaStream_t st) { xinv<1>(a, e, nfields, st); }
static void xinv_density(const XInvArgs<3>& a, const EpiDensity& e, cudaStream_t st) { xinv<3>(a, e, 1, st); }
static void xinv_psum(const XInvArgs<2>& a, const EpiPressureSum& e, cudaStream_t st) { xinv<2>(a, e, 1, st); }

static void col(const ColArgs& a, int dir, int nfields, cudaStream_t st) {
  const GenPlan pl = plan_of(a.n);
  const int W = tile_w(a.n);
  const size_t smem = ((size_t)a.n + 2 * (size_t)a.n * W) * sizeof(float2);
  const dim3 grid(grid_for(a.tile_end - a.tile_begin, ctas_per_sm(smem)), nfields);
  auto go = [&](auto kernel) {
    opt_in(kernel, smem);
    kernel<<<grid, kThreads, smem, st>>>(a, pl, dir);
  };
  W == 16 ? go(g_col<16>) : W == 8 ? go(g_col<8>) : go(g_col<4>);
}
static void zmid(const ZMidArgs& a, cudaStream_t st) {
  const GenPlan pl = plan_of(a.n);
  const int W = tile_w(a.n);
  const size_t smem = ((size_t)a.n + (a.axis == 3 ? 3 : 2) * (size_t)a.n * W) * sizeof(float2);
  const int grid = grid_for(a.ntiles, ctas_per_sm(smem));
  auto go = [&](auto kernel) {
    opt_in(kernel, smem);
    kernel<<<grid, kThreads, smem, st>>>(a, pl);
  };
  W == 16 ? go(g_zmid<16>) : W == 8 ? go(g_zmid<8>) : go(g_zmid<4>);
}

}  // namespace generic

bool generic_length_supported(int n) {
  generic::GenPlan pl;
  return generic::make_plan(n, &pl);
}

const FftOps* get_generic_fft_ops(int n) {
  if (!generic_length_supported(n)) return nullptr;
  static std::mutex mu;
  static std::map<int, FftOps> tables;
  std::lock_guard<std::mutex> lk(mu);
  auto it = tables.find(n);
  if (it == tables.end()) {
    const int w = generic::tile_w(n);
    FftOps ops{n, w, 1, w, generic::xfwd, generic::xinv_store, generic::xinv_add, generic::xinv_velocity, generic::xinv_density,
               generic::xinv_psum, generic::col, generic::zmid};
    it = tables.emplace(n, ops).first;
  }
  return &it->second;
}

}  // namespace kw
