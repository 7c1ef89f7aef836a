// One translation unit per transform length: nvcc ... -DKW_N=<N> fft_inst.cu -o fft_inst_<N>.o
#include <cmath>

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

static int x_grid(int npairs, int per_sm) {
  constexpr int RP = kXThreads / (N / 8);
  const int groups = (npairs + RP - 1) / RP;
  const int cap = sm_count() * per_sm;
  return groups < cap ? groups : cap;
}

static void xfwd(const XFwdArgs& a, int nfields, cudaStream_t st) {
  static const int per_sm = blocks_per_sm(k_xfwd<N>, kXThreads, 0);
  k_xfwd<N><<<dim3(x_grid(a.pair_end - a.pair_begin, per_sm), nfields), kXThreads, 0, st>>>(a);
}
template <int NF, class Epi> static void xinv(const XInvArgs<NF>& a, const Epi& e, int gy, cudaStream_t st) {
  constexpr size_t smem = (size_t)Epi::kStage * kXThreads * sizeof(float);  // staged epilogue operands
  static const int per_sm = [] {
    if (smem > 0) cudaFuncSetAttribute(k_xinv<N, NF, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return blocks_per_sm(k_xinv<N, NF, Epi>, kXThreads, smem);
  }();
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
  static bool once = false;
  static int per_sm = 1;
  upload_twiddles();
  if (!once) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    per_sm = blocks_per_sm(kernel, C::THREADS, smem);
    once = true;
  }
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
template <int AXIS> static void zmid_axis(const ZMidArgs& a, cudaStream_t st) {
  using C = ZCfg<N>;
  const int g = col_grid<10 + AXIS>(k_zmid<N, AXIS>, a.ntiles, C::SMEM_ZMID);
  k_zmid<N, AXIS><<<g, dim3(C::W, C::WK, C::TPC), C::SMEM_ZMID, st>>>(a);
}
// ---- plane-fused x/y passes (fft_xy.cuh) ----------------------------------------------------------------------------
// Fills the scheduling part of the arguments: items per plane, lag and ring depth from the grid that will run.
template <int ID, class K> static int xy_setup(K kernel, PipeState& ps, PipeArgs* q, int planes, int i1, int i2, size_t slot_elems) {
  using X = XYCfg<N>;
  static bool once = false;
  static int per_sm = 1;
  upload_twiddles();
  if (!once) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)X::SMEM);
    per_sm = blocks_per_sm(kernel, X::THREADS, X::SMEM);
    once = true;
  }
  const int total = planes * (i1 + i2);
  int grid = sm_count() * per_sm;
  if (grid > total) grid = total;
  const int per = i1 + i2;
  int lag = (5 * grid + 2 * per - 1) / (2 * per) + 1;  // ~2.5 waves of CTAs between the two passes of a plane
  int ring = lag + (grid + per - 1) / per + 2;
  const int cap = (int)(ps.ring_elems / slot_elems);
  if (ring > cap) ring = cap, lag = ring > 4 ? ring - 3 : ring - 1;
  if (lag > planes) lag = planes;
  if (ring <= lag) ring = lag + 1;  // only when planes is tiny; cap >= 2 is guaranteed by the allocation
  q->ctr = ps.ctr[ps.cur], q->ctr_other = ps.ctr[ps.cur ^ 1], q->nctr = ps.nctr;
  q->P = planes, q->I1 = i1, q->I2 = i2, q->L = lag, q->R = ring, q->err = ps.err;
  ps.cur ^= 1;
  return grid;
}
static bool xy_ok(const PipeState& ps, int planes, size_t slot_elems) {
  return ps.ring && 2 * planes + 1 <= ps.nctr && ps.ring_elems / slot_elems >= 4;
}
static bool xy_fwd(XYFwdArgs& a, int nfields, PipeState& ps, cudaStream_t st) {
  using X = XYCfg<N>;
  using C = ColCfg<N>;
  const size_t plane_c = (size_t)N * a.nxp;
  const int planes = nfields * a.nz;
  if (!xy_ok(ps, planes, plane_c)) return false;
  const int i2 = (a.nxp / C::W + C::TPC - 1) / C::TPC;
  a.ring = ps.ring;
  const int grid = xy_setup<0>(k_xy_fwd<N>, ps, &a.pipe, planes, X::XI, i2, plane_c);
  k_xy_fwd<N><<<grid, dim3(C::W, C::WK, C::TPC), X::SMEM, st>>>(a);
  return true;
}
template <int ID, int NF, class Epi> static bool yx_inv(YXInvArgs<NF>& a, const Epi& e, PipeState& ps, cudaStream_t st) {
  using X = XYCfg<N>;
  using C = ColCfg<N>;
  const size_t plane_c = (size_t)N * a.nxp;
  const int planes = (NF == 1 ? a.nfields : 1) * a.nz;
  if (!xy_ok(ps, planes, plane_c * NF)) return false;
  const int i1 = NF * ((a.nxp / C::W + C::TPC - 1) / C::TPC);
  a.ring = ps.ring;
  const int grid = xy_setup<10 + ID>(k_yx_inv<N, NF, Epi>, ps, &a.pipe, planes, i1, X::XI, plane_c * NF);
  k_yx_inv<N, NF, Epi><<<grid, dim3(C::W, C::WK, C::TPC), X::SMEM, st>>>(a, e);
  return true;
}
static bool yx_store(YXInvArgs<1>& a, const EpiStore& e, PipeState& ps, cudaStream_t st) { return yx_inv<0, 1>(a, e, ps, st); }
static bool yx_add(YXInvArgs<1>& a, const EpiAdd& e, PipeState& ps, cudaStream_t st) { return yx_inv<1, 1>(a, e, ps, st); }
static bool yx_velocity(YXInvArgs<1>& a, const EpiVelocity& e, PipeState& ps, cudaStream_t st) { return yx_inv<2, 1>(a, e, ps, st); }
static bool yx_density(YXInvArgs<3>& a, const EpiDensity& e, PipeState& ps, cudaStream_t st) { return yx_inv<3, 3>(a, e, ps, st); }
static bool yx_psum(YXInvArgs<2>& a, const EpiPressureSum& e, PipeState& ps, cudaStream_t st) { return yx_inv<4, 2>(a, e, ps, st); }

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
                                            KW_CAT(inst_, KW_N)::xinv_density, KW_CAT(inst_, KW_N)::xinv_psum, KW_CAT(inst_, KW_N)::col, KW_CAT(inst_, KW_N)::zmid,
                                            KW_CAT(inst_, KW_N)::xy_fwd, KW_CAT(inst_, KW_N)::yx_store, KW_CAT(inst_, KW_N)::yx_add, KW_CAT(inst_, KW_N)::yx_velocity,
                                            KW_CAT(inst_, KW_N)::yx_density, KW_CAT(inst_, KW_N)::yx_psum};

}  // namespace kw
