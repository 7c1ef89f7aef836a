// The four kernel shapes every 3-D transform of the time step is built from (all FP32, sm_100a, no cuFFT):
//
//   k_xfwd   real [z][y][x]  -> half spectrum along x [z][y][NXP]       (two real rows per complex transform)
//   k_col    in-place complex transform along y (or z) of [z][y][NXP]   (W neighbouring kx per worker)
//   k_zmid   forward z transform -> k-space operator -> inverse z transform, one HBM round trip
//   k_xinv   half spectrum -> real rows, with the real-space update fused as an epilogue functor
//
// NXP = (Nx/2+1) rounded up to 16 complex values so that column tiles are 128-byte segments.
// Together they replace cufftExecR2C/C2R (MatrixClasses/CufftComplexMatrix.cpp:511,527) and the k-space kernels
// cudaComputePressureGradient / cudaComputeVelocityGradient / cudaComputeAbsorbtionTerm / cudaComputeSourceGradient
// (KSpaceSolver/SolverCudaKernels.cu:1139,1210,1812,740).
#pragma once
#include "fft_core.cuh"
#include "fft_core2.cuh"
#include "tma.cuh"

namespace kw {

constexpr int kMaxFields = 3;
constexpr int kXThreads = 256;

// ---------------------------------------------------------------------------------------------------------------------
// Where row (z, y) of a half-spectrum array lives.  Single GPU: [z][y][NXP], i.e. row * NXP.  Slab-decomposed runs keep
// the x/y-local spectra in the y-blocked "exchange" layout [q][z_local][y_local][NXP] (q = rank that owns y after the
// transpose, y = q*nyl + y_local), so that the block sent to rank q by the all-to-all is one contiguous range.
struct RowMap {
  int ny_log2;   // log2(Ny); < 0: Ny is not a power of two (the same layouts, addressed with divisions)
  int ysh;       // log2(nyl), nyl = Ny / nranks  (== ny_log2 on one GPU)
  size_t blk;    // elements of one block: nzl * nyl * NXP
  int ny, nyl;   // Ny and Ny / nranks (used when Ny is not a power of two)
  __device__ __forceinline__ size_t off(size_t row, int nxp) const {
    if (ny_log2 < 0) {
      if (nyl == ny) return row * (size_t)nxp;
      const size_t z = row / (unsigned)ny;
      const unsigned y = (unsigned)(row - z * (unsigned)ny), q = y / (unsigned)nyl;
      return (size_t)q * blk + (z * (unsigned)nyl + (y - q * (unsigned)nyl)) * (size_t)nxp;
    }
    const size_t z = row >> ny_log2;
    const unsigned y = (unsigned)row & ((1u << ny_log2) - 1u);
    return (size_t)(y >> ysh) * blk + ((z << ysh) + (y & ((1u << ysh) - 1u))) * (size_t)nxp;
  }
};

struct XFwdArgs {
  const float* in[kMaxFields];
  float2* out[kMaxFields];
  const float2* tab;  // forward twiddle table of length Nx
  int pair_begin, pair_end;  // range of row pairs (row = z*Ny + y) this launch transforms
  int nxp;
  RowMap map;
  int n;  // Nx (read by the run-time-length kernels of fft_generic.cu; the tuned kernels have it as a template argument)
};

// One row pair (rows row0, row0+1) of one field: real -> half spectrum.  T = N/8 threads cooperate through `ex`.
// `out_off` = element offset of row0's spectrum row in `out` (row0 + 1 follows nxp elements later).
template <int N, class EX> __device__ __forceinline__ void xfwd_rows(const float* __restrict__ in, float2* __restrict__ out, int nxp,
                                                                     size_t row0, size_t out_off, bool valid, int t, const RegTw& twp, EX& ex) {
  constexpr int T = N / 8;
  const float* ra = in + row0 * N;
  float2 v[1][8];
#pragma unroll
  for (int r = 0; r < 8; ++r) v[0][r] = make_float2(__ldg(ra + t + r * T), __ldg(ra + N + t + r * T));
  fft_worker<N, -1, 1>(v, t, 0, twp, ex);
  // mirror exchange: Z[N-k] lives in worker T-t
#pragma unroll
  for (int m = 0; m < 8; ++m) ex.put(0, t + m * T, v[0][m]);
  ex.sync();
  float2 zp[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) zp[m] = ex.get(0, (N - (t + m * T)) & (N - 1));
  ex.sync();
  if (valid) {
    float2* oa = out + out_off;
    float2* ob = oa + nxp;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const float2 z = v[0][m], q = zp[m];
      oa[t + m * T] = make_float2(0.5f * (z.x + q.x), 0.5f * (z.y - q.y));
      ob[t + m * T] = make_float2(0.5f * (z.y + q.y), -0.5f * (z.x - q.x));
    }
    if (t == 0) {  // Nyquist: Z[N/2] is its own mirror
      oa[N / 2] = make_float2(v[0][4].x, 0.f);
      ob[N / 2] = make_float2(v[0][4].y, 0.f);
    }
  }
}

template <int N> __global__ void __launch_bounds__(kXThreads, 3) k_xfwd(XFwdArgs a) {
  using P = Plan<N>;
  constexpr int T = P::T;
  constexpr int RP = kXThreads / T;  // row pairs per CTA iteration
  __shared__ float2 sbuf[RP * N];
  const int t = threadIdx.x % T, rp = threadIdx.x / T;
  float2 twr[P::NTW > 0 ? P::NTW : 1];
  const float2* tab = a.tab;
  load_twiddles<N>(twr, t, [tab](int m) { return __ldg(tab + m); });
  RegTw twp{twr};
  RowExchange<T> ex{sbuf, rp * N, 1 + rp};
  const float* __restrict__ in = a.in[blockIdx.y];
  float2* __restrict__ out = a.out[blockIdx.y];
  const int npairs = a.pair_end - a.pair_begin;
  for (int g = blockIdx.x; g * RP < npairs; g += gridDim.x) {
    const int pair = a.pair_begin + g * RP + rp;
    const bool valid = pair < a.pair_end;
    const size_t row0 = 2 * (size_t)(valid ? pair : a.pair_begin);
    xfwd_rows<N>(in, out, a.nxp, row0, a.map.off(row0, a.nxp), valid, t, twp, ex);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
template <int NF> struct XInvArgs {
  const float2* in[kMaxFields];  // NF == 1: field selected by blockIdx.y; NF > 1: the NF fields of one voxel
  const float2* tab;
  int pair_begin, pair_end, nxp, ny;
  RowMap map;
  int field0;  // NF == 1: the launch covers fields field0 .. field0 + gridDim.y - 1 (per-field launches of pipelined runs)
  int n;       // Nx (fft_generic.cu)
};

// The half-spectrum values one thread needs for one row pair of one field: 4 + 4 points of the two rows, and (thread 0 of the
// transform) the two Nyquist bins.  They are loaded one transform AHEAD -- the next field of the voxel, or the first field of the
// CTA's next row pair -- so that their HBM latency hides behind the butterflies of the current transform (ncu, round 1:
// k_xinv<512,3,EpiDensity> stalled 2.9 cycles per issue on long_scoreboard with the loads at the head of each transform).
struct SpecRegs {
  float2 A[4], B[4], nA, nB;
};
template <int N> __device__ __forceinline__ SpecRegs load_spec(const float2* __restrict__ in, size_t off, int nxp, int t) {
  constexpr int T = N / 8;
  SpecRegs s;
  const float2* ia = in + off;
  const float2* ib = ia + nxp;
#pragma unroll
  for (int m = 0; m < 4; ++m) s.A[m] = __ldg(ia + t + m * T), s.B[m] = __ldg(ib + t + m * T);
  s.nA = s.nB = make_float2(0.f, 0.f);
  if (t == 0) s.nA = __ldg(ia + N / 2), s.nB = __ldg(ib + N / 2);
  return s;
}
// Hermitian merge of two rows + inverse transform: v[m] = (row a, row b) at x = t + m*T
// loop-invariant twiddles of a row transform kept in shared memory (entry n of worker t at [n * T + t]): frees the 2 * NTW registers
// RegTw holds, which is what lets k_xinv carry the prefetched spectrum of the next transform without spilling
struct SmemTw {
  const float2* tw;  // table + t
  int stride;        // T
  __device__ __forceinline__ float2 get(int n, int) const { return tw[n * stride]; }
};
template <int N, class TW, class EX> __device__ __forceinline__ void xinv_transform(const SpecRegs& s, float2 (&v)[1][8], int t, const TW& twp, EX& ex) {
  constexpr int T = N / 8;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int k = t + m * T;
    float2 A = s.A[m], B = s.B[m];
    if (m == 0 && t == 0) A.y = 0.f, B.y = 0.f;  // C2R ignores the imaginary part of DC
    v[0][m] = make_float2(A.x - B.y, A.y + B.x);  // Z[k] = A + iB
    if (!(m == 0 && t == 0)) ex.put(0, N - k, make_float2(A.x + B.y, B.x - A.y));  // Z[N-k] = conj(A) + i conj(B)
  }
  if (t == 0) ex.put(0, N / 2, make_float2(s.nA.x, s.nB.x));  // imaginary parts of the Nyquist bin ignored
  ex.sync();
#pragma unroll
  for (int m = 4; m < 8; ++m) v[0][m] = ex.get(0, t + m * T);
  ex.sync();
  fft_worker<N, +1, 1>(v, t, 0, twp, ex);
}

// Epilogue contract:  epi.apply<N>(res, field, t, row0, y, z)  where res[f][m] = (row a, row b) values at x = t + m*T of
// field f, row0 = flattened (z*Ny + y) index of row a (row b = row0 + 1, same z).
template <int N, int NF, class Epi> __global__ void __launch_bounds__(kXThreads, Epi::kMinBlocks) k_xinv(XInvArgs<NF> a, Epi epi) {
  using P = Plan<N>;
  constexpr int T = P::T;
  constexpr int RP = kXThreads / T;
  __shared__ float2 sbuf[RP * N];
  __shared__ float2 stw[(P::NTW > 0 ? P::NTW : 1) * T];
  extern __shared__ float stage_smem[];  // Epi::kStage floats per thread (slot s of thread i at [s * kXThreads + i])
  const int t = threadIdx.x % T, rp = threadIdx.x / T;
  if (rp == 0) {
    float2 twr[P::NTW > 0 ? P::NTW : 1];
    const float2* tab = a.tab;
    load_twiddles<N>(twr, t, [tab](int m) { return __ldg(tab + m); });
#pragma unroll
    for (int n = 0; n < P::NTW; ++n) stw[n * T + t] = twr[n];
  }
  __syncthreads();
  SmemTw twp{stw + t, T};
  RowExchange<T> ex{sbuf, rp * N, 1 + rp};
  float* const stg = stage_smem + threadIdx.x;
  const int npairs = a.pair_end - a.pair_begin;
  const int field = blockIdx.y + a.field0;  // NF == 1: the field of this CTA
  auto row_of = [&](int g, bool& valid) -> size_t {
    const int pair = a.pair_begin + g * RP + rp;
    valid = pair < a.pair_end;
    return 2 * (size_t)(valid ? pair : a.pair_begin);
  };
  int g = blockIdx.x;
  if (g * RP >= npairs) return;
  bool valid;
  size_t row0 = row_of(g, valid);
  SpecRegs cur = load_spec<N>(NF == 1 ? a.in[field] : a.in[0], a.map.off(row0, a.nxp), a.nxp, t);
  for (; g * RP < npairs; g += gridDim.x) {
    // real-space operands of the epilogue start their way into (thread-private slots of) shared memory now and land while
    // the transforms run: their latency is hidden without holding registers for them
    if constexpr (Epi::kStage > 0) {
      if (valid) epi.template stage<N>(stg, t, row0);
      cp_async_commit();
    }
    const int gn = g + gridDim.x;
    const bool more = gn * RP < npairs;
    bool valid_n = false;
    const size_t row_n = more ? row_of(gn, valid_n) : row0;
    float2 res[NF][8];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      SpecRegs nxt = cur;
      if (f + 1 < NF) nxt = load_spec<N>(a.in[f + 1], a.map.off(row0, a.nxp), a.nxp, t);
      else if (more) nxt = load_spec<N>(NF == 1 ? a.in[field] : a.in[0], a.map.off(row_n, a.nxp), a.nxp, t);
      float2 v[1][8];
      xinv_transform<N>(cur, v, t, twp, ex);
#pragma unroll
      for (int m = 0; m < 8; ++m) res[f][m] = v[0][m];
      cur = nxt;
    }
    if constexpr (Epi::kStage > 0) cp_async_wait<0>();
    if (valid) {
      const int y = (int)(row0 % a.ny), z = (int)(row0 / a.ny);
      if constexpr (Epi::kStage > 0) epi.template apply<N>(res, field, t, row0, y, z, stg);
      else epi.template apply<N>(res, field, t, row0, y, z);
    }
    row0 = row_n, valid = valid_n;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// column passes (y and z axes)
//
// A tile is W = 16 neighbouring kx (128-byte segments) times all N points of the transform axis.  One worker = 16 lanes;
// a thread owns E = 8..32 points of one kx in registers and the transform is two register-resident butterflies
// (radix 8/16/32, compile-time twiddles) around ONE shared-memory exchange (fft_core2.cuh).  Points go from global
// memory straight into registers and back; shared memory only carries the exchange.
// WV = kx values per tile.  16 (128-byte segments) everywhere, except the fused z pass of N = 1024 (ZCfg below): its
// 32 points per thread need ~250 registers, i.e. a 256-thread slot, i.e. 8 kx per tile (measured on B200, isolated:
// fused z pass 1024: W=8 2.2 TB/s vs W=16 (128 registers, spilling) 1.5 TB/s; plain passes: W=16 3.9-4.0 vs W=8 2.9-3.5).
template <int N, int WV = 16> struct ColCfg {
  using P = Plan2<N>;
  static constexpr int W = WV;
  static constexpr int WK = P::WK;                                         // workers (threads per kx) of a tile
  static constexpr int SLOT = W * WK;                                      // threads of a tile slot
  static constexpr int TPC = (SLOT >= 256) ? 1 : 256 / SLOT;               // tile slots per CTA
  static constexpr int THREADS = SLOT * TPC;
  static constexpr int MINB = (THREADS * 128 <= 32768) ? 2 : 1;            // CTAs per SM at <= 128 registers
  static constexpr size_t SMEM = (P::R2 > 1) ? (size_t)TPC * N * W * sizeof(float2) : 0;
  // + landing area of the real multiplier + the z-indexed 1-D operator
  static constexpr size_t SMEM_ZMID = SMEM + (size_t)TPC * N * W * sizeof(float) + (size_t)N * sizeof(float2);
  // TMA variant: the next tile lands in its own buffer, the multiplier in alternating ones; + one mbarrier
  static constexpr size_t SMEM_ZMID_DB = 2 * SMEM + 2 * (size_t)TPC * N * W * sizeof(float) + (size_t)N * sizeof(float2) + 16;
  // barrier flavour of a slot: whole CTA, named barrier (slot spans whole warps), or none (single-stage plans)
  static constexpr int BAR_THREADS = (P::R2 == 1) ? -1 : (TPC == 1 ? 0 : SLOT);
};
template <int N> using ZCfg = ColCfg<N, (N >= 1024) ? 8 : 16>;  // tiles of the fused z pass (cp.async / double-buffered TMA forms)
// MODE of k_zmid: 0 = cp.async into the exchange buffer (late), 1 = TMA, double-buffered (a tile ahead), 2 = TMA into the exchange
// buffer (late, single-buffered).  MODE 2 exists for N = 1024: with no per-thread copy addresses to keep, 32 points per thread fit
// into 128 registers with < 130 bytes of spills (cp.async form: 530), so the tile can be 16 kx wide with 512 threads = 16 warps per SM
// instead of 8-wide with 8 warps; 128 KB exchange + 64 KB multiplier leave no room for a second tile buffer.
template <int N, int MODE> using ZCfgM = ColCfg<N, (N >= 1024 && MODE != 2) ? 8 : 16>;


template <int W, int NTHREADS> struct ColExchange2 {
  float2* buf;  // tile buffer + lane
  int bar;      // named barrier id of this slot
  // W = 16: a worker's 16 lanes cover all 32 banks, any row index is conflict free.  W = 8 (fused z pass of N = 1024, radix 32 x 32):
  // a warp holds 4 workers of 64 bytes each; the stage-1 outputs of neighbouring workers lie 32 rows apart -- the same half of
  // the banks, a 2-way conflict on every store (ncu: 36 % of the shared-memory wavefronts of k_zmid<1024> were conflicts,
  // profiles/r02_j_zmid1024.summary.txt).  Swapping odd and even rows in every other group of 32 rows separates them and
  // leaves the other access patterns (neighbouring workers = neighbouring rows) conflict free.
  __device__ __forceinline__ static int row(int i) { return W == 8 ? (i ^ ((i >> 5) & 1)) : i; }
  __device__ __forceinline__ float2* at(int i) const { return buf + row(i) * W; }
  __device__ __forceinline__ void put(int i, float2 x) { buf[row(i) * W] = x; }
  __device__ __forceinline__ float2 get(int i) const { return buf[row(i) * W]; }
  __device__ __forceinline__ void sync() {
    if (NTHREADS == 0) __syncthreads();
    else if (NTHREADS > 0) asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(NTHREADS > 0 ? NTHREADS : 32) : "memory");
  }
};

// twiddles of the transform length of this translation unit, in constant memory (index is uniform per worker)
#ifdef KW_N
__constant__ float2 c_tw[KW_N];
struct ConstTab {
  __device__ __forceinline__ float2 operator()(int m) const { return c_tw[m]; }
};
#endif

struct ColArgs {
  float2* data[kMaxFields];
  size_t stride;        // elements between consecutive points of the transform axis
  size_t outer_stride;  // elements between consecutive "outer" tiles
  int ngroups;          // NXP / W
  int tile_begin, tile_end;  // range of tiles (tile = outer * ngroups + group) this launch transforms
  // y-blocked exchange layout (slab-decomposed runs, see RowMap): point e of a worker (axis index w + WK*e) lives at
  // (e >> blk_es) * blk + (e & ((1 << blk_es) - 1)) * WK * stride;  requires nyl >= WK.  Unused by k_col<.., false>.
  int blk_es;
  size_t blk;
  int n;  // length of the transform axis (fft_generic.cu)
  // kx >= nvalid are padding columns (NXP - (Nx/2+1) <= 15 of them): they are transformed like the others (their lanes carry whatever shared
  // memory holds) but neither loaded nor stored -- 5.5 % of the traffic of a column pass at Nx = 512.  0 = every column is moved.
  int nvalid;
  int nyl;  // fft_generic.cu, y-blocked layout with a length that is not a power of two: point i lives at (i / nyl) * blk + (i % nyl) * stride; 0 = plain
};

#ifdef KW_N
template <int N, int DIR, bool BLOCKED> __global__ void __launch_bounds__(ColCfg<N>::THREADS, ColCfg<N>::MINB) k_col(ColArgs a) {
  using C = ColCfg<N>;
  using P = Plan2<N>;
  constexpr int W = C::W, WK = C::WK, E = P::E;
  extern __shared__ float2 smem[];
  const int lane = threadIdx.x, w = threadIdx.y, tz = threadIdx.z;
  ColExchange2<W, C::BAR_THREADS> ex{smem + (size_t)tz * N * W + lane, 1 + tz};
  float2* __restrict__ data = a.data[blockIdx.y];
  const int ntiles = a.tile_end - a.tile_begin;
  const int niter = (ntiles + C::TPC - 1) / C::TPC;
  const size_t estride = (size_t)WK * a.stride;
  auto poff = [&](int e) -> size_t {
    if constexpr (BLOCKED) return (size_t)(e >> a.blk_es) * a.blk + (size_t)(e & ((1 << a.blk_es) - 1)) * estride;
    else return e * estride;
  };
  const int nvalid = a.nvalid ? a.nvalid : a.ngroups * W;
  auto tile_ptr = [&](int it, bool& valid, bool& moved) -> float2* {
    const int tile = a.tile_begin + it * C::TPC + tz;
    valid = tile < a.tile_end;
    const int tl = valid ? tile : a.tile_begin;
    const int kx = (tl % a.ngroups) * W + lane;
    moved = kx < nvalid;  // padding columns are neither loaded nor stored
    return data + (size_t)(tl / a.ngroups) * a.outer_stride + kx + (size_t)w * a.stride;
  };
  // The points of the next tile are copied asynchronously (LDGSTS) into the exchange buffer as soon as the current
  // transform has read its last exchange back, each thread fetching exactly the points it will own: the global-load
  // latency hides behind the second butterfly and the stores.  Single-stage plans (N <= 32) load straight to registers.
  auto prefetch = [&](int it) {
    if constexpr (P::R2 > 1) {
      bool valid, moved;
      const float2* p = tile_ptr(it, valid, moved);
      if (moved) {
#pragma unroll
        for (int e = 0; e < E; ++e) cp_async8(ex.buf + (w + WK * e) * W, p + poff(e));
      }
      cp_async_commit();
    }
  };
  int it = blockIdx.x;
  if (it < niter) prefetch(it);
  for (; it < niter; it += gridDim.x) {
    bool valid, moved;
    float2* p = tile_ptr(it, valid, moved);
    float2 v[E];
    if constexpr (P::R2 > 1) {
      cp_async_wait<0>();
#pragma unroll
      for (int e = 0; e < E; ++e) v[e] = ex.get(w + WK * e);
      ex.sync();  // landing slots of other workers must be consumed before stage-1 outputs overwrite them
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) v[e] = moved ? p[poff(e)] : make_float2(0.f, 0.f);
    }
    const int nxt = it + gridDim.x;
    fft2_worker<N, DIR>(v, w, ex, ConstTab(), [&] {
      if (nxt < niter) prefetch(nxt);
    });
    if (valid && moved) {
#pragma unroll
      for (int e = 0; e < E; ++e) p[poff(e)] = v[e];
    }
  }
}
#endif

// fused z pass:  out = IFFT_z( (FFT_z(in) * (mul * scal)) (x) vec[coord(AXIS)] ), one field per launch.
// AXIS (compile time): -1 no 1-D operator, 0: vec indexed by kx, 1: by ky, 2: by kz (staged in shared memory),
// 3: gradient -- e = FFT_z(in)*mul*scal is kept in registers and three products e (x) vec_{x,y,z} are inverse-transformed
// and stored (cudaComputePressureGradient fused between the z transforms, SolverCudaKernels.cu:1139-1157).
struct ZField {
  const float2* in;
  float2* out;
  const float* mul;   // real multiplier on the padded reduced grid [kz][ky][NXP], or nullptr
  float scal;         // scalar folded into the multiplier (fftDivider where the reference folds it there)
  const float2* vec;  // 1-D complex operator (AXIS >= 0)
  // AXIS == 3 (gradient): one forward transform feeds three inverse transforms, out/out_y/out_z with the x/y/z operators
  float2* out_y;
  float2* out_z;
  const float2* vec_y;
  const float2* vec_z;
};
struct ZMidArgs {
  ZField f;
  int axis;
  int nxp, ngroups, ntiles;  // ntiles = Ny * ngroups
  unsigned plane;            // Ny * NXP
  int n;                     // Nz (fft_generic.cu)
  int nvalid;                // as in ColArgs: Nx/2 + 1, or 0
};

#ifdef KW_N
#ifndef KW_ZMID_MINB
// E = 32 points per thread need ~150 registers to stay spill free (ncu/ptxas: 128 registers spill 300+ bytes and run
// 25-55% slower than one CTA per SM at 254 registers)
#define KW_ZMID_MINB ((MODE == 2) ? ZCfgM<N, MODE>::MINB : (Plan2<N>::E >= 32 || AXIS == 3) ? 1 : ZCfgM<N, MODE>::MINB)
#endif
#ifndef KW_ZMID_MULMODE
#define KW_ZMID_MULMODE 0  // 0: multiplier lands in shared memory through cp.async; 1: plain loads at the point of use
#endif
// MODE 1 (one tile slot per CTA, N >= 256): the next tile is requested as soon as the current one sits in registers -- a whole
// tile of compute ahead -- by ONE thread through the TMA unit (box copies of 256 rows x 128 bytes, completion on an
// mbarrier) into a landing buffer of its own, its multiplier into the other of two multiplier buffers.  MODE 0: every
// thread copies its own points with cp.async once the last inverse butterfly has freed the exchange buffer.  MODE 2: see ZCfgM.
template <int N, int AXIS, int MODE = 0> __global__ void __launch_bounds__(ZCfgM<N, MODE>::THREADS, KW_ZMID_MINB) k_zmid(ZMidArgs a, const __grid_constant__ ZMaps maps) {
  using C = ZCfgM<N, MODE>;
  using P = Plan2<N>;
  constexpr bool DB = MODE == 1, TS = MODE == 2, TMA = DB || TS;
  static_assert(!TMA || (P::R2 > 1 && C::TPC == 1), "the TMA variants need a plan with an exchange and one tile slot per CTA");
  static_assert(!TS || C::W == 16, "single-buffered TMA tiles land in the exchange buffer, whose rows are only linear for W = 16");
  constexpr int W = C::W, WK = C::WK, E = P::E;
  constexpr size_t TILES = (size_t)C::TPC * N * W;  // points of all tile slots of the CTA
  extern __shared__ float2 smem[];
  const int lane = threadIdx.x, w = threadIdx.y, tz = threadIdx.z;
  ColExchange2<W, C::BAR_THREADS> ex{smem + (size_t)tz * N * W + lane, 1 + tz};
  float2* const land = DB ? ex.buf + TILES : ex.buf;
  float2* const land0 = DB ? smem + TILES : smem;  // where the TMA boxes go
  float* const mulbase = reinterpret_cast<float*>(smem + (DB ? 2 : 1) * C::SMEM / sizeof(float2));
  float* const mul0 = mulbase + (size_t)tz * N * W + w * W + lane;
  float2* const svec = smem + ((DB ? 2 : 1) * (C::SMEM + TILES * sizeof(float))) / sizeof(float2);  // N entries (AXIS == 2)
  uint64_t* const bar = reinterpret_cast<uint64_t*>(svec + N);
  const bool leader = TMA && lane == 0 && w == 0 && tz == 0;
  if (leader) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (TMA) __syncthreads();
  const float2* __restrict__ in = a.f.in;
  const float* __restrict__ mul = a.f.mul;
  const int niter = (a.ntiles + C::TPC - 1) / C::TPC;
  const unsigned estride = (unsigned)WK * a.plane;
  if (AXIS == 2 || AXIS == 3) {
    const float2* zv = AXIS == 2 ? a.f.vec : a.f.vec_z;
    for (int i = threadIdx.x + W * (threadIdx.y + WK * threadIdx.z); i < N; i += C::THREADS) svec[i] = __ldg(zv + i);
    __syncthreads();
  }
  const int nvalid = a.nvalid ? a.nvalid : a.nxp;  // kx >= nvalid: padding columns, neither loaded nor stored
  auto tile_base = [&](int it, bool& valid, int& y, int& kx) -> unsigned {
    const int tile = it * C::TPC + tz;
    valid = tile < a.ntiles;
    const int tl = valid ? tile : 0;
    y = tl / a.ngroups, kx = (tl % a.ngroups) * W + lane;
    return (unsigned)y * a.nxp + kx + (unsigned)w * a.plane;  // point e of this worker: base + e*estride
  };
  auto prefetch = [&](int it, int par) {  // see k_col
    if constexpr (P::R2 > 1) {
      bool valid;
      int y, kx;
      const unsigned b = tile_base(it, valid, y, kx);
      if constexpr (TMA) {
        if (leader) {  // lane 0: kx is the first kx of the tile
          constexpr int ZB = N < 256 ? N : 256;  // rows per box
          mbar_expect_tx(bar, (uint32_t)(TILES * (mul ? 12 : 8)));
#pragma unroll
          for (int zb = 0; zb < N / ZB; ++zb) {
            tma_load_3d(land0 + (size_t)zb * ZB * W, &maps.in, 2 * kx, y, zb * ZB, bar);
            if (mul) tma_load_3d(mulbase + (DB ? par * TILES : 0) + (size_t)zb * ZB * W, &maps.mul, kx, y, zb * ZB, bar);
          }
        }
      } else {
        const float2* p = in + b;
        if (kx < nvalid) {
#pragma unroll
          for (int e = 0; e < E; ++e) cp_async8(ex.at(w + WK * e), p + e * estride);
        }
        cp_async_commit();
      }
    }
  };
  int it = blockIdx.x, par = 0;
  if (it < niter) prefetch(it, 0);
  for (; it < niter; it += gridDim.x, par ^= 1) {  // par = parity of the CTA's tile count = phase of the mbarrier
    bool valid;
    int y, kx;
    const unsigned base = tile_base(it, valid, y, kx);
    const float* const mulbuf = mul0 + (DB ? par * TILES : 0);
    const int nxt = it + gridDim.x;
    // the real multiplier of this tile lands in shared memory (cp.async, thread-private slots) while the forward
    // transform runs: no registers, no exposed latency.  Issued before the points are taken into registers.
    const bool moved = kx < nvalid;
    if constexpr (!TMA) {
      if (mul && KW_ZMID_MULMODE == 0 && moved) {
#pragma unroll
        for (int e = 0; e < E; ++e) cp_async4(mul0 + e * (WK * W), mul + base + e * estride);
      }
      cp_async_commit();
    }
    float2 v[E];
    if constexpr (P::R2 > 1) {
      if constexpr (TMA) mbar_wait(bar, (uint32_t)par);
      else cp_async_wait<1>();  // the tile (older group) has landed; the multiplier may still be in flight
#pragma unroll
      for (int e = 0; e < E; ++e) v[e] = DB ? land[(w + WK * e) * W] : ex.get(w + WK * e);  // TMA boxes land row by row
      ex.sync();
      if constexpr (DB) {
        if (nxt < niter) prefetch(nxt, par ^ 1);
      }
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) v[e] = moved ? __ldg(in + base + e * estride) : make_float2(0.f, 0.f);
    }
    fft2_worker<N, -1>(v, w, ex, ConstTab());
    if constexpr (!TMA) cp_async_wait<0>();
    asm volatile("" ::: "memory");  // keep the phases apart: interleaving them only lengthens live ranges
    const float scal = a.f.scal;
    if constexpr (AXIS == 3) {
      float2 ev[E];
#pragma unroll
      for (int e = 0; e < E; ++e) ev[e] = cscale(v[e], mul ? mulbuf[e * (WK * W)] * scal : scal);
      const float2 wx = __ldg(a.f.vec + kx), wy = __ldg(a.f.vec_y + y);
#pragma unroll
      for (int f = 0; f < 3; ++f) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = cmul(ev[e], f == 0 ? wx : f == 1 ? wy : svec[w + WK * e]);
        if (f < 2) fft2_worker<N, +1>(v, w, ex, ConstTab());
        else fft2_worker<N, +1>(v, w, ex, ConstTab(), [&] {
          if (!DB && nxt < niter) prefetch(nxt, 0);
        });
        if (valid && moved) {
          float2* __restrict__ p = (f == 0 ? a.f.out : f == 1 ? a.f.out_y : a.f.out_z) + base;
#pragma unroll
          for (int e = 0; e < E; ++e) p[e * estride] = v[e];
        }
      }
    } else {
      float2 w01 = make_float2(1.f, 0.f);
      if (AXIS == 0) w01 = __ldg(a.f.vec + kx);
      if (AXIS == 1) w01 = __ldg(a.f.vec + y);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float m = mul ? (KW_ZMID_MULMODE == 0 ? mulbuf[e * (WK * W)] : (moved ? __ldg(mul + base + e * estride) : 0.f)) * scal : scal;
        float2 x = cscale(v[e], m);
        if (AXIS == 0 || AXIS == 1) x = cmul(x, w01);
        if (AXIS == 2) x = cmul(x, svec[w + WK * e]);
        v[e] = x;
      }
      fft2_worker<N, +1>(v, w, ex, ConstTab(), [&] {
        if (!DB && nxt < niter) prefetch(nxt, 0);
      });
      if (valid && moved) {
        float2* __restrict__ p = a.f.out + base;
#pragma unroll
        for (int e = 0; e < E; ++e) p[e * estride] = v[e];
      }
    }
  }
}
#endif

}  // namespace kw
