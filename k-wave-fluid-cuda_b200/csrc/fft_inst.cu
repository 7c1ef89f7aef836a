// One translation unit per transform length: nvcc ... -DKW_N=<N> fft_inst.cu -o fft_inst_<N>.o
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "ops.h"

#ifndef KW_N
#error "compile with -DKW_N=<transform length>"
#endif

namespace kw {
#define KW_CAT2(a, b) a##b
#define KW_CAT(a, b) KW_CAT2(a, b)
// a distinctly named namespace per length so that every translation unit gets distinct symbols
namespace KW_CAT(inst_, KW_N) {

constexpr int N = KW_N;

template <class K> static int blocks_per_sm(K kernel, int threads, size_t smem) {
  int b = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, threads, smem);
  return b > 0 ? b : 1;
}

// Per-device launch state of one kernel: the dynamic shared-memory opt-in (cudaFuncSetAttribute is per device) and the
// occupancy-derived CTAs per SM.  A process may hold contexts on several devices.
struct PerDev {
  bool once[64] = {};
  int per_sm[64] = {};
};
template <class K> static int kernel_setup(PerDev& pd, K kernel, int threads, size_t smem) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!pd.once[dev]) {
    if (smem > 0) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pd.per_sm[dev] = blocks_per_sm(kernel, threads, smem);
    pd.once[dev] = true;
  }
  return pd.per_sm[dev];
}

static int x_grid(int npairs, int per_sm) {
  constexpr int RP = kXThreads / (N / 8);
  const int groups = (npairs + RP - 1) / RP;
  const int cap = sm_count() * per_sm;
  return groups < cap ? groups : cap;
}

static void xfwd(const XFwdArgs& a, int nfields, cudaStream_t st) {
  static PerDev pd;
  const int per_sm = kernel_setup(pd, k_xfwd<N>, kXThreads, 0);
  k_xfwd<N><<<dim3(x_grid(a.pair_end - a.pair_begin, per_sm), nfields), kXThreads, 0, st>>>(a);
}
template <int NF, class Epi> static void xinv(const XInvArgs<NF>& a, const Epi& e, int gy, cudaStream_t st) {
  constexpr size_t smem = (size_t)Epi::kStage * kXThreads * sizeof(float);  // staged epilogue operands
  static PerDev pd;
  const int per_sm = kernel_setup(pd, k_xinv<N, NF, Epi>, kXThreads, smem);
  k_xinv<N, NF, Epi><<<dim3(x_grid(a.pair_end - a.pair_begin, per_sm), gy), kXThreads, smem, st>>>(a, e);
}
static void xinv_store(const XInvArgs<1>& a, const EpiStore& e, int nf, cudaStream_t st) { xinv<1>(a, e, nf, st); }
static void xinv_add(const XInvArgs<1>& a, const EpiAdd& e, cudaStream_t st) { xinv<1>(a, e, 1, st); }
static void xinv_velocity(const XInvArgs<1>& a, const EpiVelocity& e, int nf, cudaStream_t st) { xinv<1>(a, e, nf, st); }
static void xinv_density(const XInvArgs<3>& a, const EpiDensity& e, cudaStream_t st) { xinv<3>(a, e, 1, st); }
static void xinv_psum(const XInvArgs<2>& a, const EpiPressureSum& e, cudaStream_t st) { xinv<2>(a, e, 1, st); }

// ID distinguishes kernels of identical function type (the statics below are per kernel)
static void upload_twiddles() {  // forward table e^{-2 pi i m/N} in double precision -> constant memory, once per device
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || done[dev]) return;
  float2 h[N];
  for (int m = 0; m < N; ++m) {
    const double a = -2.0 * 3.14159265358979323846 * (double)m / (double)N;
    h[m] = make_float2((float)cos(a), (float)sin(a));
  }
  cudaMemcpyToSymbol(c_tw, h, sizeof(h));
  done[dev] = true;
}

template <int ID, class K> static int col_grid(K kernel, int ntiles, size_t smem) {
  using C = ColCfg<N>;
  static PerDev pd;
  upload_twiddles();
  const int per_sm = kernel_setup(pd, kernel, C::THREADS, smem);
  const int groups = (ntiles + C::TPC - 1) / C::TPC;
  const int cap = sm_count() * per_sm;
  return groups < cap ? groups : cap;
}

static void col(const ColArgs& a, int dir, int nfields, cudaStream_t st) {
  using C = ColCfg<N>;
  const dim3 block(C::W, C::WK, C::TPC);
  const bool blocked = a.blk != 0;
  if (dir < 0 && !blocked) {
    const int g = col_grid<0>(k_col<N, -1, false>, a.tile_end - a.tile_begin, C::SMEM);
    k_col<N, -1, false><<<dim3(g, nfields), block, C::SMEM, st>>>(a);
  } else if (dir > 0 && !blocked) {
    const int g = col_grid<1>(k_col<N, +1, false>, a.tile_end - a.tile_begin, C::SMEM);
    k_col<N, +1, false><<<dim3(g, nfields), block, C::SMEM, st>>>(a);
  } else if (dir < 0) {
    const int g = col_grid<2>(k_col<N, -1, true>, a.tile_end - a.tile_begin, C::SMEM);
    k_col<N, -1, true><<<dim3(g, nfields), block, C::SMEM, st>>>(a);
  } else {
    const int g = col_grid<3>(k_col<N, +1, true>, a.tile_end - a.tile_begin, C::SMEM);
    k_col<N, +1, true><<<dim3(g, nfields), block, C::SMEM, st>>>(a);
  }
}
#ifndef KW_ZMID_1024_DEFAULT
#define KW_ZMID_1024_DEFAULT 2
#endif
// Variants of the fused z pass: KW_ZMID_VARIANT = 0 every thread copies its points with cp.async into the exchange buffer once the
// last inverse butterfly has freed it; 1 = one thread requests the next tile through the TMA unit a whole tile ahead (N >= 256).
// (A 16-points-per-thread "paired" plan -- radix-32 / 64 butterflies shared by 2 / 4 threads of a warp through shuffles, 16 warps
//  per SM -- was built and measured in round 2: 3.2 / 1.5 TB/s at N = 512 / 1024 against 3.9 / 2.3 for this plan, 4.2 with TMA
//  against 4.7; removed again, profiles/r02_d_zmid_variants.log keeps the numbers.)
static int zmid_variant() {
  static const int v = getenv("KW_ZMID_VARIANT") ? atoi(getenv("KW_ZMID_VARIANT")) : -1;
  return v;
}
// ID distinguishes kernels of identical function type (the statics are per instantiation); W = kx values per tile
template <int ID, bool DB, class K> static void zmid_go(K kernel, const ZMidArgs& a, int W, dim3 block, int tiles_per_cta, size_t smem, cudaStream_t st) {
  static PerDev pd;
  upload_twiddles();
  const int per_sm = kernel_setup(pd, kernel, (int)(block.x * block.y * block.z), smem);
  ZMaps maps{};
  if (DB) {  // box = 256 rows (kz) of W neighbouring kx of one ky
    const uint64_t ny = a.plane / a.nxp;
    const uint32_t zb = N < 256 ? N : 256;
    // the tensors end at the last valid kx: the padding columns of a box are filled with zeros without being read (ZMidArgs::nvalid)
    const uint64_t nv = a.nvalid ? (uint64_t)a.nvalid : (uint64_t)a.nxp;
    bool ok = make_tensor_map_3d(&maps.in, a.f.in, 2ull * nv, ny, N, 8ull * a.nxp, 8ull * a.plane, 2u * W, 1, zb);
    if (ok && a.f.mul) ok = make_tensor_map_3d(&maps.mul, a.f.mul, nv, ny, N, 4ull * a.nxp, 4ull * a.plane, (uint32_t)W, 1, zb);
    if (!ok) {
      fprintf(stderr, "kwave_b200: cuTensorMapEncodeTiled failed for the fused z pass (N = %d)\n", N);
      abort();
    }
  }
  const int groups = (a.ntiles + tiles_per_cta - 1) / tiles_per_cta;
  const int cap = sm_count() * per_sm;
  kernel<<<groups < cap ? groups : cap, block, smem, st>>>(a, maps);
}
template <int NN, int AXIS, bool = (NN >= 256)> struct TmaZ {
  static bool launch(const ZMidArgs&, int, cudaStream_t) { return false; }
};
template <int NN, int AXIS> struct TmaZ<NN, AXIS, true> {
  template <int MODE> static void go(ZMidArgs a, cudaStream_t st) {
    using C = ZCfgM<NN, MODE>;
    const int ny = a.ntiles / a.ngroups;  // the caller sized the tiles for FftOps::zmid_w
    a.ngroups = a.nxp / C::W, a.ntiles = ny * a.ngroups;
    const size_t smem = MODE == 1 ? C::SMEM_ZMID_DB : C::SMEM_ZMID + 16;
    zmid_go<10 * MODE + AXIS + 1, true>(k_zmid<NN, AXIS, MODE>, a, C::W, dim3(C::W, C::WK, C::TPC), C::TPC, smem, st);
  }
  static bool launch(const ZMidArgs& a, int variant, cudaStream_t st) {
    if (variant == 1) go<1>(a, st);
    else if (variant == 2) go<2>(a, st);
    else return false;
    return true;
  }
};
template <int AXIS> static void zmid_axis(const ZMidArgs& a, cudaStream_t st) {
  using C = ZCfg<N>;
  int variant = zmid_variant();
  // default: TMA double-buffered tiles where they win (profiles/r02_d_zmid_variants.log: N = 256 +20 %, N = 512 +19 %, N = 1024 -5 %);
  // N = 1024: 16-wide single-buffered TMA tiles (see ZCfgM)
  if (variant < 0) variant = (N == 256 || N == 512) ? 1 : (N == 1024 ? KW_ZMID_1024_DEFAULT : 0);
  if (TmaZ<N, AXIS>::launch(a, variant, st)) return;
  zmid_go<AXIS + 1, false>(k_zmid<N, AXIS, 0>, a, C::W, dim3(C::W, C::WK, C::TPC), C::TPC, C::SMEM_ZMID, st);
}
static void zmid(const ZMidArgs& a, cudaStream_t st) {
  switch (a.axis) {
    case 0: zmid_axis<0>(a, st); break;
    case 1: zmid_axis<1>(a, st); break;
    case 2: zmid_axis<2>(a, st); break;
    case 3: zmid_axis<3>(a, st); break;
    default: zmid_axis<-1>(a, st); break;
  }
}

}  // namespace inst_<N>

#define KW_OPS_NAME2(n) fft_ops_##n
#define KW_OPS_NAME(n) KW_OPS_NAME2(n)
extern const FftOps KW_OPS_NAME(KW_N) = {KW_N, ColCfg<KW_N>::W, ColCfg<KW_N>::WK, ZCfg<KW_N>::W, KW_CAT(inst_, KW_N)::xfwd, KW_CAT(inst_, KW_N)::xinv_store, KW_CAT(inst_, KW_N)::xinv_add, KW_CAT(inst_, KW_N)::xinv_velocity,
                                            KW_CAT(inst_, KW_N)::xinv_density, KW_CAT(inst_, KW_N)::xinv_psum, KW_CAT(inst_, KW_N)::col, KW_CAT(inst_, KW_N)::zmid};

}  // namespace kw
