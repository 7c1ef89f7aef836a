// Host side of the C ABI (include/kwave_b200.h): context, pre-processing, per-step sequencing, output streams.
//
// Mirrors, for the hot path only, KSpaceFirstOrderSolver::{preProcessing, computeMainLoop, storeSensorData,
// postProcessing} (KSpaceSolver/KSpaceFirstOrderSolver.cpp:784-857, :864-943, :1060-1093, :950-973) and the parts of
// MatrixContainer / OutputStreamContainer that decide which arrays and streams exist.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/kwave_b200.h"
#include "nccl_dl.h"
#include "peer_dl.h"
#include "ops.h"

namespace kw {

// ---------------------------------------------------------------------------------------------------------------------
thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define KW_CUDA(call)                                                                                    \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(e_ == cudaErrorMemoryAllocation ? KW_ERR_ALLOC : KW_ERR_CUDA,                          \
                  std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
  } while (0)
#define KW_TRY(expr)          \
  do {                        \
    int r_ = (expr);          \
    if (r_ != KW_OK) return r_; \
  } while (0)

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// cuTensorMapEncodeTiled through the runtime's driver entry-point lookup: the library links no libcuda and still loads (and
// exports its symbols) on a machine without a driver
bool make_tensor_map_3d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                        uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<Fn>(p);
  }();
  if (!fn) return false;
  const cuuint64_t dims[3] = {d0, d1, d2}, strides[2] = {stride1_bytes, stride2_bytes};
  const cuuint32_t box[3] = {b0, b1, b2}, estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

const FftOps* get_fft_ops(int n) {
  switch (n) {
    case 16: return &fft_ops_16;
    case 32: return &fft_ops_32;
    case 64: return &fft_ops_64;
    case 128: return &fft_ops_128;
    case 256: return &fft_ops_256;
    case 512: return &fft_ops_512;
    case 1024: return &fft_ops_1024;
    default: return get_generic_fft_ops(n);  // other lengths 8 m with factors 2, 3, 5, 7: run-time-length kernels (fft_generic.cu)
  }
}

// forward twiddle table e^{-2 pi i m/N}, generated in double precision, one device copy per (device, N)
static int twiddle_table(int n, const float2** out) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, float2*> cache;
  int dev = 0;
  KW_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find({dev, n});
  if (it == cache.end()) {
    std::vector<float2> h(n);
    for (int m = 0; m < n; ++m) {
      const double a = -2.0 * M_PI * (double)m / (double)n;
      h[m] = make_float2((float)cos(a), (float)sin(a));
    }
    float2* d = nullptr;
    KW_CUDA(cudaMalloc(&d, n * sizeof(float2)));
    KW_CUDA(cudaMemcpy(d, h.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
    it = cache.emplace(std::make_pair(dev, n), d).first;
  }
  *out = it->second;
  return KW_OK;
}

static inline int ew_grid(size_t n) {
  const size_t b = (n + 255) / 256;
  const size_t cap = (size_t)sm_count() * 8;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

// ---------------------------------------------------------------------------------------------------------------------
struct Stream {
  bool enabled = false;
  int op = kOpNone;
  int src = 0;       // 0: p, 1..3: ux,uy,uz (staggered)
  bool all = false;  // whole-domain aggregate
  float* dbuf = nullptr;
  size_t row = 0, cap_rows = 0, rows = 0;
  // asynchronous output (kw_stream_async): series streams own two device buffers; a full one travels to pinned host memory on the
  // output stream while sampling continues into the other (IndexOutputStream::flushBufferToFile :583-591 writes every step instead)
  float* dalt = nullptr;       // the other device buffer
  float* hbuf = nullptr;       // pinned host copy of the buffer in flight
  size_t pending_rows = 0;     // rows of the buffer in flight / landed, not yet fetched
  cudaEvent_t landed = nullptr;
  bool fused_this_step = false;  // accumulated by the kernel that produced the field
  // compressed streams (kOpC): two overlapped accumulators per (sensor, harmonic); frames are appended to dbuf
  bool nosave = false;          // created only because an I_avg_c stream needs it (OutputStreamContainer.cpp:273-323)
  bool shifted_bases = false;   // u*_non_staggered_c: time-shifted bases, exponent offset 114 (BaseOutputStream.cpp:63-102)
  float* sbuf = nullptr;        // raw samples of the current step
  void *acc1 = nullptr, *acc2 = nullptr;
  void* cur = nullptr;          // the accumulator completed at this step (between flush and postSample2)
  size_t acc_bytes = 0;
  uint64_t sampled = 0, compressed = 0;
  bool saving = false;
};

// Grid + slab decomposition.  Rank r of P owns the real-space planes z in [z0, z0 + nzl) of every
// field ("x/y-local" side) and, after the all-to-all of a transform, the ky range [y0, y0 + nyl) of every spectrum for
// all kz ("z-local" side).  P == 1: nzl = nz, nyl = ny, and every layout below reduces to the plain [z][y][x] one.
struct Geometry {
  int nx = 0, ny = 0, nz = 0, nxr = 0, nxp = 0;
  int rank = 0, nranks = 1;
  int nzl = 0, nyl = 0, z0 = 0, y0 = 0;  // local extents / offsets
  int ny_log2 = 0, ysh = 0;              // log2(ny), log2(nyl)
  size_t n = 0, nc = 0;                  // LOCAL real voxels, LOCAL padded complex elements (nxp*ny*nzl == nxp*nyl*nz)
  size_t nca = 0;                        // LOCAL complex elements without the padding, (nx/2+1)*ny*nzl: the algorithmic bytes of SURVEY 8(d)
  size_t ntot = 0;                       // global real voxels
  size_t blk = 0;                        // complex elements exchanged with one peer per field: nxp*nyl*nzl
  const FftOps *ox = nullptr, *oy = nullptr, *oz = nullptr;
  const float2 *tx = nullptr, *ty = nullptr, *tz = nullptr;
  RowMap row_map() const { return RowMap{ny_log2, ysh, blk, ny, nyl}; }
  int init(uint64_t nx_, uint64_t ny_, uint64_t nz_, int rank_ = 0, int nranks_ = 1) {
    nx = (int)nx_, ny = (int)ny_, nz = (int)nz_;
    // Nz == 1 (2-D simulations, Parameters.h:88-94): the z transforms are identities; everything else is the 3-D path with a
    // z extent of one (the z components stay exactly zero, so the sums of the reference's 2-D kernels are reproduced)
    ox = get_fft_ops(nx), oy = get_fft_ops(ny), oz = nz == 1 ? nullptr : get_fft_ops(nz);
    if (!ox || !oy || (!oz && nz != 1))
      return fail(KW_ERR_INVALID, "grid sizes must be multiples of 8 in [16,2048] with prime factors 2, 3, 5, 7 (powers of two up to 1024 run the tuned "
                                  "kernels, other lengths the run-time-length ones); got " +
                                      std::to_string(nx_) + "x" + std::to_string(ny_) + "x" + std::to_string(nz_));
    rank = rank_, nranks = nranks_ < 1 ? 1 : nranks_;
    if (rank < 0 || rank >= nranks || (nranks & (nranks - 1)) || nz % nranks || ny % nranks)
      return fail(KW_ERR_INVALID, "slab decomposition needs a power-of-two number of ranks dividing Ny and Nz");
    nzl = nz / nranks, nyl = ny / nranks, z0 = rank * nzl, y0 = rank * nyl;
    if (nranks > 1 && (nyl < oy->col_wk || (nyl & 1)))
      return fail(KW_ERR_INVALID, "slab decomposition: Ny / nranks = " + std::to_string(nyl) + " is below the y-pass worker count " +
                                      std::to_string(oy->col_wk) + " of this Ny (use fewer ranks)");
    for (ny_log2 = 0; (1 << ny_log2) < ny; ++ny_log2) {}
    for (ysh = 0; (1 << ysh) < nyl; ++ysh) {}
    if (ny & (ny - 1)) ny_log2 = -1, ysh = 0;  // RowMap: plain [z][y][NXP]
    nxr = nx / 2 + 1;
    nxp = (nxr + 15) / 16 * 16;
    ntot = (size_t)nx * ny * nz;
    n = (size_t)nx * ny * nzl;
    nc = (size_t)nxp * ny * nzl;
    nca = (size_t)nxr * ny * nzl;
    blk = (size_t)nxp * nyl * nzl;
    KW_TRY(twiddle_table(nx, &tx));
    KW_TRY(twiddle_table(ny, &ty));
    if (nz > 1) KW_TRY(twiddle_table(nz, &tz));
    return KW_OK;
  }
};

}  // namespace kw

using namespace kw;

// per-kernel device timing (CUDA events on the solver stream) and algorithmic bytes, for bench.py's roofline
struct ProfStat {
  uint64_t launches = 0;
  double ms = 0.0, bytes = 0.0;
};
struct ProfPending {
  const char* name;
  cudaEvent_t e0, e1;
  double bytes;
  bool solver_stream;  // launched on the solver stream: gaps between consecutive ones are reported as "idle before <name>"
};

struct kw_ctx {
  bool prof_on = false;
  std::map<std::string, ProfStat> prof;
  std::vector<ProfPending> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  kw_config cfg{};
  Geometry g;
  cudaStream_t st = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool preprocessed = false, finished = false;
  uint64_t t = 0;
  uint64_t launches = 0;
  float last_ms = 0.f;
  // host copies of inputs needed by pre-processing
  std::vector<float> h_in[KW_ARRAY_COUNT];
  std::vector<uint64_t> h_idx[KW_ARRAY_COUNT];
  // device arrays (nullptr when absent); float arrays indexed by kw_array
  float* d[KW_ARRAY_COUNT] = {};
  uint64_t* di[KW_ARRAY_COUNT] = {};
  size_t count[KW_ARRAY_COUNT] = {};
  float scalar[KW_ARRAY_COUNT] = {};
  float2* S[4] = {};
  float2* R[4] = {};  // receive side of the all-to-all (slab-decomposed runs only)
  float *tA = nullptr, *tB = nullptr, *tNL = nullptr, *tSrc = nullptr;
  uint64_t* cub_offsets = nullptr;
  uint64_t* cub_corners = nullptr;  // cuboid corners clipped to the local slab, local z
  size_t nsens = 0;  // LOCAL sensor points (index mask) or LOCAL cuboid points
  size_t nsens_total = 0;
  int ncuboids = 0;
  // slab-decomposed runs: positions of the locally kept entries of each index list in the original list
  uint64_t* dpos[KW_ARRAY_COUNT] = {};
  size_t count_total[KW_ARRAY_COUNT] = {};
  std::vector<uint64_t> sens_pos;     // index mask: position of every local sensor point in the global row
  std::vector<uint64_t> sens_ranges;  // cuboid mask: (start in the global row, length) per cuboid part held locally
  KwNcclComm comm = nullptr;
  double comm_bytes = 0.0;
  // slab-decomposed runs: exchanges run on their own stream, ordered against the solver stream by events
  cudaStream_t cs = nullptr;
  cudaStream_t ws = nullptr;  // local copy of the own block (peer path): keeps the copy stream for the NVLink pushes
  std::vector<cudaEvent_t> ev_ring;
  size_t ev_next = 0;
  bool async_out = false;     // kw_stream_async
  cudaStream_t os = nullptr;  // output stream: device -> pinned host copies of full row buffers
  PeerLink peer;        // copy-engine pushes into peer memory (falls back to NCCL send/recv when unavailable)
  char* arena = nullptr;  // S[0..3], R[0..3] in one allocation (one IPC handle per rank)
  char nccl_id[128] = {};
  Stream streams[KW_STREAM_COUNT];
  // compression (Compression/CompressHelper.cpp:48-65): bases [0] plain, [1] time shifted
  int c_osize = 0, c_bsize = 0, c_H = 0;
  bool c_no_overlap = false;
  float2 *d_be[2] = {}, *d_be1[2] = {};
  // non-staggered velocity (cpp:2714-2735): shift operators extended to the full ky / kz range
  float2* shift_full[3] = {};
  bool need_shifted = false;
  std::vector<void*> owned;

  Fld fld(int id) const { return Fld{count[id] > 1 ? d[id] : nullptr, scalar[id]}; }
};

namespace kw {

static cudaEvent_t prof_event(kw_ctx* c) {
  if (!c->prof_pool.empty()) {
    cudaEvent_t e = c->prof_pool.back();
    c->prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
// every kernel launch of the time loop goes through here: counts it and, in profiling mode, brackets it with events
template <class F> static void launch(kw_ctx* c, const char* name, double bytes, F&& f, cudaStream_t stream = nullptr) {
  if (!stream || stream == c->st) c->launches++;  // kernels (and the few copies) of the solver stream; exchanges are not kernels
  if (!c->prof_on) {
    f();
    return;
  }
  if (!stream) stream = c->st;
  ProfPending p{name, prof_event(c), prof_event(c), bytes, stream == c->st};
  cudaEventRecord(p.e0, stream);
  f();
  cudaEventRecord(p.e1, stream);
  c->prof_pending.push_back(p);
}
// an event marking everything enqueued on `s` so far (ring of reusable events: a wait captures the record made before it)
static cudaEvent_t mark(kw_ctx* c, cudaStream_t s) {
  if (c->ev_ring.empty()) {
    c->ev_ring.resize(256);
    for (auto& e : c->ev_ring) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  }
  cudaEvent_t e = c->ev_ring[c->ev_next++ % c->ev_ring.size()];
  cudaEventRecord(e, s);
  return e;
}
static void prof_resolve(kw_ctx* c) {  // stream must be idle
  const ProfPending* prev = nullptr;
  for (auto& p : c->prof_pending) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.e0, p.e1);
    ProfStat& st = c->prof[p.name];
    st.launches++, st.ms += ms, st.bytes += p.bytes;
    if (p.solver_stream && c->g.nranks > 1) {  // time the solver stream spent waiting (for an exchange) before this launch
      if (prev) {
        float gap = 0.f;
        cudaEventElapsedTime(&gap, prev->e1, p.e0);
        if (gap > 0.002f) {
          ProfStat& idle = c->prof[std::string("idle_before_") + p.name];
          idle.launches++, idle.ms += gap;
        }
      }
      prev = &p;
    }
  }
  for (auto& p : c->prof_pending) c->prof_pool.push_back(p.e0), c->prof_pool.push_back(p.e1);
  c->prof_pending.clear();
}

static int dalloc(kw_ctx* c, void** p, size_t bytes, bool zero = true) {
  KW_CUDA(cudaMalloc(p, bytes ? bytes : 4));
  c->owned.push_back(*p);
  if (zero) KW_CUDA(cudaMemsetAsync(*p, 0, bytes ? bytes : 4, c->st));
  return KW_OK;
}

static bool is_index_array(int id) {
  return id == KW_SENSOR_MASK_INDEX || id == KW_SENSOR_MASK_CORNERS || id == KW_P_SOURCE_INDEX || id == KW_U_SOURCE_INDEX ||
         id == KW_DELAY_MASK;
}
static bool is_complex_vec(int id) {
  return (id >= KW_DDX_K_SHIFT_POS_R && id <= KW_DDZ_K_SHIFT_NEG) || (id >= KW_X_SHIFT_NEG_R && id <= KW_Z_SHIFT_NEG_R);
}
static bool is_reduced_real(int id) {
  return id == KW_KAPPA || id == KW_SOURCE_KAPPA || id == KW_ABSORB_NABLA1 || id == KW_ABSORB_NABLA2;
}

// upload a host float array into a (new or existing) device array
static int upload_f(kw_ctx* c, int id, const float* h, size_t n) {
  if (!c->d[id] || c->count[id] != n) KW_TRY(dalloc(c, (void**)&c->d[id], n * sizeof(float), false));
  c->count[id] = n;
  KW_CUDA(cudaMemcpyAsync(c->d[id], h, n * sizeof(float), cudaMemcpyHostToDevice, c->st));
  KW_CUDA(cudaStreamSynchronize(c->st));
  return KW_OK;
}
// reduced-grid real operator given in the reference layout [nz][ny][nxr] -> padded device array
static int upload_reduced(kw_ctx* c, int id, const float* h) {
  const Geometry& g = c->g;
  const size_t rows = (size_t)g.ny * g.nz;
  float* tmp = nullptr;
  KW_CUDA(cudaMalloc(&tmp, rows * g.nxr * sizeof(float)));
  KW_CUDA(cudaMemcpyAsync(tmp, h, rows * g.nxr * sizeof(float), cudaMemcpyHostToDevice, c->st));
  if (!c->d[id]) KW_TRY(dalloc(c, (void**)&c->d[id], g.nc * sizeof(float), false));
  c->count[id] = rows * g.nxr;
  k_pad_real<<<ew_grid(g.nc), 256, 0, c->st>>>(c->d[id], tmp, g.nxr, g.nxp, rows);
  c->launches++;
  KW_CUDA(cudaStreamSynchronize(c->st));
  KW_CUDA(cudaFree(tmp));
  return KW_OK;
}

// ---- host pre-processing, FP32 as in the reference -----------------------------------------------------------------
// KSpaceFirstOrderSolver.cpp:2404-2452 (kappa), :2460-2506 (source kappa), :2514-2577 (kappa + nablas)
static void generate_k_operators(const kw_config& cf, const Geometry& g, std::vector<float>* kappa, std::vector<float>* n1,
                                 std::vector<float>* n2, std::vector<float>* skappa) {
  // 2-D (Nz == 1): dz is not part of the input file and the z term is dropped (cpp:2404-2452, k2D branches)
  const float dx2 = 1.0f / (cf.dx * cf.dx), dy2 = 1.0f / (cf.dy * cf.dy), dz2 = g.nz == 1 ? 0.0f : 1.0f / (cf.dz * cf.dz);
  const float cRefDtPi = cf.c_ref * cf.dt * static_cast<float>(M_PI);
  const float cRefDt2 = cf.c_ref * cf.dt * 0.5f;
  const float pi2 = static_cast<float>(M_PI) * 2.0f;
  const float nxRec = 1.0f / static_cast<float>(g.nx), nyRec = 1.0f / static_cast<float>(g.ny),
              nzRec = 1.0f / static_cast<float>(g.nz);
  const float ap = cf.alpha_power;
  // output: the z-local side of this rank, [kz][ky_local][NXP] (zero padded) -- what k_zmid multiplies with
  const size_t tot = (size_t)g.nxp * g.nyl * g.nz;
  if (kappa) kappa->assign(tot, 0.f);
  if (n1) n1->assign(tot, 0.f), n2->assign(tot, 0.f);
  if (skappa) skappa->assign(tot, 0.f);
#pragma omp parallel for schedule(static)
  for (int z = 0; z < g.nz; z++) {
    float zPart = 0.5f - fabsf(0.5f - (float)z * nzRec);
    zPart = (zPart * zPart) * dz2;
    for (int yl = 0; yl < g.nyl; yl++) {
      const int y = g.y0 + yl;
      float yPart = 0.5f - fabsf(0.5f - (float)y * nyRec);
      yPart = (yPart * yPart) * dy2;
      const float yzPart = zPart + yPart;
      for (int x = 0; x < g.nxr; x++) {
        float xPart = 0.5f - fabsf(0.5f - (float)x * nxRec);
        xPart = (xPart * xPart) * dx2;
        const size_t i = ((size_t)z * g.nyl + yl) * g.nxp + x;
        const float root = sqrtf(xPart + yzPart);
        if (n1) {  // absorbing: kappa from pi2 * root * cRefDt2 (cpp:2556-2561)
          const float k = pi2 * root;
          const float cRefK = cRefDt2 * k;
          if (kappa) (*kappa)[i] = (cRefK == 0.0f) ? 1.0f : sinf(cRefK) / cRefK;
          float a = powf(k, ap - 2.0f), b = powf(k, ap - 1.0f);
          if (a == std::numeric_limits<float>::infinity()) a = 0.0f;
          if (b == std::numeric_limits<float>::infinity()) b = 0.0f;
          (*n1)[i] = a;
          (*n2)[i] = b;
        } else if (kappa) {  // lossless: kappa from cRefDtPi * root (cpp:2440-2446)
          const float k = cRefDtPi * root;
          (*kappa)[i] = (k == 0.0f) ? 1.0f : sinf(k) / k;
        }
        if (skappa) (*skappa)[i] = cosf(cRefDtPi * root);
      }
    }
  }
}

}  // namespace kw

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" {

int kw_abi_version(void) { return KW_ABI_VERSION; }
const char* kw_last_error(void) { return g_err.c_str(); }

__global__ void k_code_version(int* v) {
#ifdef __CUDA_ARCH__
  *v = __CUDA_ARCH__ / 10;
#endif
}
int kw_cuda_code_version(int* version) {
  int* d = nullptr;
  KW_CUDA(cudaMalloc(&d, sizeof(int)));
  KW_CUDA(cudaMemset(d, 0, sizeof(int)));
  k_code_version<<<1, 1>>>(d);
  KW_CUDA(cudaGetLastError());
  KW_CUDA(cudaMemcpy(version, d, sizeof(int), cudaMemcpyDeviceToHost));
  KW_CUDA(cudaFree(d));
  return KW_OK;
}

int kw_device_memory(size_t* free_bytes, size_t* total_bytes) {
  size_t f = 0, t = 0;
  KW_CUDA(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return KW_OK;
}

int kw_ctx_create(const kw_config* cfg, kw_ctx** out) {
  if (!cfg || !out) return fail(KW_ERR_INVALID, "null argument");
  if (cfg->abi_version != KW_ABI_VERSION || cfg->struct_size != sizeof(kw_config))
    return fail(KW_ERR_INVALID, "kw_config ABI mismatch");
  if (cfg->nonuniform_grid_flag) return fail(KW_ERR_INVALID, "nonuniform_grid_flag must be 0 (main.cpp:460)");
  if (cfg->nz < 1) return fail(KW_ERR_INVALID, "Nz must be at least 1");
  if (cfg->nz == 1 && cfg->nranks > 1) return fail(KW_ERR_INVALID, "2-D simulations (Nz == 1) run on one GPU");
  if (cfg->absorbing_flag && cfg->alpha_power == 1.0f)
    return fail(KW_ERR_INVALID, "alpha_power == 1 is not supported (Parameters.cpp:421-424)");
  if (cfg->nranks > 1 && !cfg->nccl_unique_id) return fail(KW_ERR_INVALID, "nranks > 1 needs the shared ncclUniqueId (kw_nccl_unique_id)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(KW_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  // CudaParameters::selectDevice (Parameters/CudaParameters.cpp:81-167): a given index must exist and be usable; without one the
  // first device that accepts a context and runs this library's sm_100a code is taken (a device in exclusive-process mode that
  // another process holds fails the probe and is skipped).  Unlike the reference the library never resets a device: the process
  // may hold other contexts (PyTorch in the benchmark).
  auto usable = [](int d) {
    int v = 0;
    const bool ok = cudaSetDevice(d) == cudaSuccess && cudaFree(nullptr) == cudaSuccess && kw_cuda_code_version(&v) == KW_OK && v >= 100;
    cudaGetLastError();
    return ok;
  };
  if (cfg->device >= 0) {
    if (cfg->device >= ndev)
      return fail(KW_ERR_INVALID, "Wrong CUDA device id " + std::to_string(cfg->device) + ". Allowed devices <0, " + std::to_string(ndev - 1) + ">.");
    if (!usable(cfg->device)) return fail(KW_ERR_CUDA, "CUDA device id " + std::to_string(cfg->device) + " is busy or unavailable.");
  } else {
    int found = -1;
    for (int d = 0; d < ndev && found < 0; ++d)
      if (usable(d)) found = d;
    if (found < 0) return fail(KW_ERR_CUDA, "All CUDA-capable devices are busy or unavailable.");
  }
  kw_ctx* c = new kw_ctx();
  c->cfg = *cfg;
  int r = c->g.init(cfg->nx, cfg->ny, cfg->nz, cfg->rank, cfg->nranks);
  if (r != KW_OK) {
    delete c;
    return r;
  }
  if (c->g.nranks > 1) {  // one communicator per context, over NVLink / NVSwitch
    NcclApi& nc = nccl_api();
    if (!nc.ok) {
      delete c;
      return fail(KW_ERR_COMM, "NCCL unavailable: " + nc.error);
    }
    KwNcclUniqueId id;
    memcpy(&id, cfg->nccl_unique_id, sizeof(id));
    const int e = nc.CommInitRank(&c->comm, c->g.nranks, id, c->g.rank);
    if (e != kNcclSuccess) {
      const std::string msg = nc.GetErrorString(e);
      delete c;
      return fail(KW_ERR_COMM, "ncclCommInitRank: " + msg);
    }
    memcpy(c->nccl_id, cfg->nccl_unique_id, sizeof(c->nccl_id));
    c->cfg.nccl_unique_id = nullptr;
    KW_CUDA(cudaStreamCreateWithFlags(&c->cs, cudaStreamNonBlocking));
    static const bool own = !getenv("KW_SELF_COPY_STREAM") || atoi(getenv("KW_SELF_COPY_STREAM")) != 0;
    if (own) KW_CUDA(cudaStreamCreateWithFlags(&c->ws, cudaStreamNonBlocking));  // local copy of the own block of every exchange
  }
  KW_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
  KW_CUDA(cudaEventCreate(&c->ev0));
  KW_CUDA(cudaEventCreate(&c->ev1));
  *out = c;
  return KW_OK;
}

int kw_ctx_destroy(kw_ctx* c) {
  if (!c) return KW_OK;
  cudaStreamSynchronize(c->st);
  if (c->cs) cudaStreamSynchronize(c->cs);
  if (c->ws) cudaStreamSynchronize(c->ws), cudaStreamDestroy(c->ws);
  if (c->peer.shm) c->peer.teardown();
  if (c->comm) nccl_api().CommDestroy(c->comm);
  if (c->os) cudaStreamSynchronize(c->os), cudaStreamDestroy(c->os);
  for (auto& s : c->streams) {
    if (s.hbuf) cudaFreeHost(s.hbuf);
    if (s.landed) cudaEventDestroy(s.landed);
  }
  for (cudaEvent_t e : c->ev_ring) cudaEventDestroy(e);
  if (c->cs) cudaStreamDestroy(c->cs);
  for (void* p : c->owned) cudaFree(p);
  prof_resolve(c);
  for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->st) cudaStreamDestroy(c->st);
  delete c;
  return KW_OK;
}

int kw_set_array(kw_ctx* c, int id, const void* host, uint64_t count) {
  if (!c || !host || id < 0 || id >= KW_ARRAY_COUNT || count == 0) return fail(KW_ERR_INVALID, "kw_set_array: bad argument");
  const Geometry& g = c->g;
  if (is_index_array(id)) {
    const uint64_t* h = static_cast<const uint64_t*>(host);
    c->h_idx[id].assign(h, h + count);
    return KW_OK;  // shifted to 0-based and uploaded by kw_preprocess
  }
  if (is_complex_vec(id)) {  // count complex values; x vectors are padded to nxp
    const bool xvec = (id == KW_DDX_K_SHIFT_POS_R || id == KW_DDX_K_SHIFT_NEG_R || id == KW_X_SHIFT_NEG_R);
    const size_t want = xvec ? g.nxr
                             : (id == KW_DDY_K_SHIFT_POS || id == KW_DDY_K_SHIFT_NEG) ? g.ny
                             : (id == KW_DDZ_K_SHIFT_POS || id == KW_DDZ_K_SHIFT_NEG) ? g.nz
                             : (id == KW_Y_SHIFT_NEG_R) ? g.ny / 2 + 1 : g.nz / 2 + 1;
    if (count != want) return fail(KW_ERR_INVALID, "kw_set_array: wrong length for complex vector " + std::to_string(id));
    const size_t padded = xvec ? g.nxp : want;
    std::vector<float> tmp(2 * padded, 0.f);
    memcpy(tmp.data(), host, 2 * count * sizeof(float));
    KW_TRY(upload_f(c, id, tmp.data(), 2 * padded));
    if (id >= KW_X_SHIFT_NEG_R && id <= KW_Z_SHIFT_NEG_R) c->h_in[id] = tmp;  // extended to full length by kw_preprocess
    return KW_OK;
  }
  const float* h = static_cast<const float*>(host);
  if (is_reduced_real(id)) {
    if (g.nranks > 1) return fail(KW_ERR_INVALID, "kw_set_array: k-space operators are generated per rank in slab-decomposed runs");
    if (count != (size_t)g.nxr * g.ny * g.nz) return fail(KW_ERR_INVALID, "kw_set_array: wrong size for reduced-grid operator");
    return upload_reduced(c, id, h);
  }
  switch (id) {
    case KW_C0: case KW_ALPHA_COEFF: case KW_RHO0_SGX: case KW_RHO0_SGY: case KW_RHO0_SGZ:
      // transformed by kw_preprocess on the host, as the reference does
      if (count != 1 && count != g.n) return fail(KW_ERR_INVALID, "kw_set_array: medium array must have 1 or Nx*Ny*Nz elements");
      c->h_in[id].assign(h, h + count);
      return KW_OK;
    case KW_RHO0: case KW_BONA: case KW_ABSORB_TAU: case KW_ABSORB_ETA:
      if (count != 1 && count != g.n) return fail(KW_ERR_INVALID, "kw_set_array: medium array must have 1 or Nx*Ny*Nz elements");
      if (count == 1) { c->scalar[id] = h[0]; c->count[id] = 1; return KW_OK; }
      return upload_f(c, id, h, count);
    case KW_P: case KW_RHOX: case KW_RHOY: case KW_RHOZ: case KW_UX_SGX: case KW_UY_SGY: case KW_UZ_SGZ: case KW_P0_SOURCE_INPUT:
      if (count != g.n) return fail(KW_ERR_INVALID, "kw_set_array: field must have Nx*Ny*Nz elements");
      return upload_f(c, id, h, count);
    case KW_PML_X_SGX: case KW_PML_X: if (count != (size_t)g.nx) return fail(KW_ERR_INVALID, "pml x length"); return upload_f(c, id, h, count);
    case KW_PML_Y_SGY: case KW_PML_Y: if (count != (size_t)g.ny) return fail(KW_ERR_INVALID, "pml y length"); return upload_f(c, id, h, count);
    case KW_PML_Z_SGZ: case KW_PML_Z: if (count != (size_t)g.nz) return fail(KW_ERR_INVALID, "pml z length"); return upload_f(c, id, h, count);
    case KW_P_SOURCE_INPUT: case KW_TRANSDUCER_SOURCE_INPUT: case KW_UX_SOURCE_INPUT: case KW_UY_SOURCE_INPUT: case KW_UZ_SOURCE_INPUT:
      return upload_f(c, id, h, count);
    default:
      return fail(KW_ERR_INVALID, "kw_set_array: array id " + std::to_string(id) + " is not an input");
  }
}

int kw_set_source_row(kw_ctx* c, int id, uint64_t t, const float* row, uint64_t count) {
  if (!c || !row) return fail(KW_ERR_INVALID, "null argument");
  if (id != KW_P_SOURCE_INPUT && id != KW_UX_SOURCE_INPUT && id != KW_UY_SOURCE_INPUT && id != KW_UZ_SOURCE_INPUT)
    return fail(KW_ERR_INVALID, "kw_set_source_row: not a source input");
  if (!c->d[id] || (t + 1) * count > c->count[id]) return fail(KW_ERR_INVALID, "kw_set_source_row: row outside the signal");
  KW_CUDA(cudaMemcpyAsync(c->d[id] + t * count, row, count * sizeof(float), cudaMemcpyHostToDevice, c->st));
  return KW_OK;
}

static const struct { int op; int src; bool all; bool supported; } kStreamTable[KW_STREAM_COUNT] = {
    /* P_RAW */ {kOpNone, 0, false, true}, /* P_C */ {kOpC, 0, false, true}, /* P_RMS */ {kOpRms, 0, false, true},
    /* P_MAX */ {kOpMax, 0, false, true}, /* P_MIN */ {kOpMin, 0, false, true}, /* P_MAX_ALL */ {kOpMax, 0, true, true},
    /* P_MIN_ALL */ {kOpMin, 0, true, true},
    /* U*_RAW */ {kOpNone, 1, false, true}, {kOpNone, 2, false, true}, {kOpNone, 3, false, true},
    /* U*_C */ {kOpC, 1, false, true}, {kOpC, 2, false, true}, {kOpC, 3, false, true},
    /* U*_NS_RAW */ {kOpNone, 4, false, true}, {kOpNone, 5, false, true}, {kOpNone, 6, false, true},
    /* U*_NS_C */ {kOpC, 4, false, true}, {kOpC, 5, false, true}, {kOpC, 6, false, true},
    /* U*_RMS */ {kOpRms, 1, false, true}, {kOpRms, 2, false, true}, {kOpRms, 3, false, true},
    /* U*_MAX */ {kOpMax, 1, false, true}, {kOpMax, 2, false, true}, {kOpMax, 3, false, true},
    /* U*_MIN */ {kOpMin, 1, false, true}, {kOpMin, 2, false, true}, {kOpMin, 3, false, true},
    /* U*_MAX_ALL */ {kOpMax, 1, true, true}, {kOpMax, 2, true, true}, {kOpMax, 3, true, true},
    /* U*_MIN_ALL */ {kOpMin, 1, true, true}, {kOpMin, 2, true, true}, {kOpMin, 3, true, true},
    /* I*_AVG */ {0, 0, false, false}, {0, 0, false, false}, {0, 0, false, false},
    /* I*_AVG_C */ {kOpIAvgC, 0, false, true}, {kOpIAvgC, 1, false, true}, {kOpIAvgC, 2, false, true},
    /* Q_TERM */ {0, 0, false, false}, /* Q_TERM_C */ {kOpQTermC, 0, false, true}};

int kw_stream_enable(kw_ctx* c, int sid) {
  if (!c || sid < 0 || sid >= KW_STREAM_COUNT) return fail(KW_ERR_INVALID, "kw_stream_enable: bad stream id");
  if (c->preprocessed) return fail(KW_ERR_STATE, "kw_stream_enable must precede kw_preprocess");
  if (!kStreamTable[sid].supported) return fail(KW_ERR_INVALID, "stream " + std::to_string(sid) + " is not available in this build");
  Stream& s = c->streams[sid];
  s.enabled = true;
  s.op = kStreamTable[sid].op;
  s.src = kStreamTable[sid].src;
  s.all = kStreamTable[sid].all;
  return KW_OK;
}

// second device buffer + pinned host buffer of a series stream (asynchronous output only)
static int async_buffers(kw_ctx* c, Stream& s) {
  if (!c->async_out || s.cap_rows == 0) return KW_OK;
  // (a rank of a slab-decomposed run that holds no sensor point still switches buffers in step with the others: kw_run is collective)
  const size_t bytes = std::max<size_t>(s.cap_rows * s.row * sizeof(float), 4);
  KW_TRY(dalloc(c, (void**)&s.dalt, bytes));
  KW_CUDA(cudaHostAlloc((void**)&s.hbuf, bytes, cudaHostAllocDefault));
  KW_CUDA(cudaEventCreateWithFlags(&s.landed, cudaEventDisableTiming));
  if (!c->os) KW_CUDA(cudaStreamCreateWithFlags(&c->os, cudaStreamNonBlocking));
  return KW_OK;
}

int kw_stream_async(kw_ctx* c, int enable) {
  if (!c) return fail(KW_ERR_INVALID, "null context");
  if (c->preprocessed) return fail(KW_ERR_STATE, "kw_stream_async must precede kw_preprocess");
  c->async_out = enable != 0;
  return KW_OK;
}

int kw_preprocess(kw_ctx* c) {
  if (!c) return fail(KW_ERR_INVALID, "null context");
  if (c->preprocessed) return fail(KW_ERR_STATE, "kw_preprocess called twice");
  const kw_config& cf = c->cfg;
  const Geometry& g = c->g;
  // --- 1. index arrays: 1-based -> 0-based (cpp:787-813; IndexMatrix.cpp:161-169)
  for (int id : {KW_SENSOR_MASK_INDEX, KW_SENSOR_MASK_CORNERS, KW_P_SOURCE_INDEX, KW_U_SOURCE_INDEX, KW_DELAY_MASK}) {
    auto& h = c->h_idx[id];
    if (h.empty()) continue;
    for (auto& v : h) {
      if (v == 0) return fail(KW_ERR_INVALID, "index arrays are 1-based in the input file; found 0");
      v -= 1;
    }
    if (id != KW_DELAY_MASK && id != KW_SENSOR_MASK_CORNERS)
      for (auto v : h)
        if (v >= g.ntot) return fail(KW_ERR_INVALID, "index outside the grid in array " + std::to_string(id));
    c->count_total[id] = h.size();
  }
  // slab-decomposed runs keep the points of the local slab only (global 0-based linear index -> owner = z / nzl), with
  // their position in the original list so that signals are looked up and rows are assembled in list order
  {
    const uint64_t lo = (uint64_t)g.z0 * g.nx * g.ny, hi = lo + g.n;
    auto keep_local = [&](int id, std::vector<uint64_t>* pos) {
      auto& h = c->h_idx[id];
      std::vector<uint64_t> loc;
      for (size_t j = 0; j < h.size(); ++j)
        if (h[j] >= lo && h[j] < hi) loc.push_back(h[j] - lo), pos->push_back(j);
      h.swap(loc);
    };
    if (g.nranks > 1) {
      std::vector<uint64_t> pos;
      auto upload_pos = [&](int id) -> int {
        KW_TRY(dalloc(c, (void**)&c->dpos[id], pos.size() * sizeof(uint64_t), false));
        KW_CUDA(cudaMemcpyAsync(c->dpos[id], pos.data(), pos.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, c->st));
        KW_CUDA(cudaStreamSynchronize(c->st));
        return KW_OK;
      };
      if (!c->h_idx[KW_P_SOURCE_INDEX].empty()) {
        keep_local(KW_P_SOURCE_INDEX, &pos);
        KW_TRY(upload_pos(KW_P_SOURCE_INDEX));
      }
      if (!c->h_idx[KW_U_SOURCE_INDEX].empty()) {
        pos.clear();
        keep_local(KW_U_SOURCE_INDEX, &pos);
        KW_TRY(upload_pos(KW_U_SOURCE_INDEX));
        auto& dm = c->h_idx[KW_DELAY_MASK];
        if (!dm.empty()) {  // delay of every kept transducer point
          std::vector<uint64_t> loc(pos.size());
          for (size_t j = 0; j < pos.size(); ++j) loc[j] = dm[pos[j]];
          dm.swap(loc);
        }
      }
      if (!c->h_idx[KW_SENSOR_MASK_INDEX].empty()) keep_local(KW_SENSOR_MASK_INDEX, &c->sens_pos);
    }
    for (int id : {KW_SENSOR_MASK_INDEX, KW_P_SOURCE_INDEX, KW_U_SOURCE_INDEX, KW_DELAY_MASK}) {
      auto& h = c->h_idx[id];
      if (c->count_total[id] == 0) continue;
      KW_TRY(dalloc(c, (void**)&c->di[id], h.size() * sizeof(uint64_t), false));
      KW_CUDA(cudaMemcpyAsync(c->di[id], h.data(), h.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, c->st));
      c->count[id] = h.size();
    }
    KW_CUDA(cudaStreamSynchronize(c->st));
  }
  // --- Nz == 1: the input file of a 2-D simulation has no z arrays (main.cpp:446-563); give the 3-D path neutral ones
  if (g.nz == 1) {
    const float one = 1.0f, zero2[2] = {0.f, 0.f};
    for (int id : {KW_PML_Z, KW_PML_Z_SGZ})
      if (!c->d[id]) KW_TRY(upload_f(c, id, &one, 1));
    for (int id : {KW_DDZ_K_SHIFT_POS, KW_DDZ_K_SHIFT_NEG})
      if (!c->d[id]) KW_TRY(upload_f(c, id, zero2, 2));
    if (c->h_in[KW_RHO0_SGZ].empty()) c->h_in[KW_RHO0_SGZ].assign(1, 1.0f);
    if (c->h_in[KW_Z_SHIFT_NEG_R].empty() && !c->h_in[KW_X_SHIFT_NEG_R].empty()) c->h_in[KW_Z_SHIFT_NEG_R] = {1.0f, 0.0f};
  }
  // --- 2. dt / rho0_sg (cpp:825-830)
  for (int k = 0; k < 3; ++k) {
    const int id = KW_RHO0_SGX + k;
    auto& h = c->h_in[id];
    if (h.empty()) return fail(KW_ERR_INVALID, "rho0_sg* missing");
    for (auto& v : h) v = cf.dt / v;
    if (h.size() == 1) c->scalar[id] = h[0], c->count[id] = 1;
    else KW_TRY(upload_f(c, id, h.data(), h.size()));
  }
  // --- 3. k-space operators (cpp:835-843) unless supplied through kw_set_array
  auto& c0 = c->h_in[KW_C0];
  if (c0.empty()) return fail(KW_ERR_INVALID, "c0 missing");
  {
    std::vector<float> kappa, n1, n2, sk;
    const bool need_kappa = !c->d[KW_KAPPA];
    const bool need_nabla = cf.absorbing_flag && !c->d[KW_ABSORB_NABLA1];
    const bool need_sk = ((cf.p_source_flag && cf.p_source_mode == KW_SRC_ADDITIVE) ||
                          ((cf.ux_source_flag || cf.uy_source_flag || cf.uz_source_flag) && cf.u_source_mode == KW_SRC_ADDITIVE)) &&
                         !c->d[KW_SOURCE_KAPPA];
    if (need_kappa || need_nabla || need_sk)
      generate_k_operators(cf, g, need_kappa ? &kappa : nullptr, cf.absorbing_flag ? &n1 : nullptr,
                           cf.absorbing_flag ? &n2 : nullptr, need_sk ? &sk : nullptr);
    auto upload_padded = [&](int id, const std::vector<float>& h) -> int {
      KW_TRY(upload_f(c, id, h.data(), h.size()));
      c->count[id] = (size_t)g.nxr * g.nyl * g.nz;
      return KW_OK;
    };
    if (need_kappa) KW_TRY(upload_padded(KW_KAPPA, kappa));
    if (need_nabla) {
      KW_TRY(upload_padded(KW_ABSORB_NABLA1, n1));
      KW_TRY(upload_padded(KW_ABSORB_NABLA2, n2));
    }
    if (need_sk) KW_TRY(upload_padded(KW_SOURCE_KAPPA, sk));
  }
  if (cf.absorbing_flag && c->count[KW_ABSORB_TAU] == 0) {  // tau, eta (cpp:2584-2643); c0 still unsquared here
    auto& al = c->h_in[KW_ALPHA_COEFF];
    if (al.empty()) return fail(KW_ERR_INVALID, "alpha_coeff missing");
    const float ap = cf.alpha_power;
    const float tanPi2 = tanf(static_cast<float>(M_PI_2) * ap);
    const float neper = (100.0f * powf(1.0e-6f / (2.0f * static_cast<float>(M_PI)), ap)) / (20.0f * static_cast<float>(M_LOG10E));
    const size_t m = std::max(al.size(), c0.size());
    std::vector<float> tau(m), eta(m);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)m; ++i) {
      const float a2 = 2.0f * neper * (al.size() == 1 ? al[0] : al[i]);
      const float cc = c0.size() == 1 ? c0[0] : c0[i];
      tau[i] = (-a2) * powf(cc, ap - 1.0f);
      eta[i] = a2 * powf(cc, ap) * tanPi2;
    }
    if (m == 1) {
      c->scalar[KW_ABSORB_TAU] = tau[0], c->scalar[KW_ABSORB_ETA] = eta[0];
      c->count[KW_ABSORB_TAU] = c->count[KW_ABSORB_ETA] = 1;
    } else {
      KW_TRY(upload_f(c, KW_ABSORB_TAU, tau.data(), m));
      KW_TRY(upload_f(c, KW_ABSORB_ETA, eta.data(), m));
    }
  }
  // --- 5. c2 = c0^2 (cpp:2690-2703)
  for (auto& v : c0) v = v * v;
  if (c0.size() == 1) c->scalar[KW_C0] = c0[0], c->count[KW_C0] = 1;
  else KW_TRY(upload_f(c, KW_C0, c0.data(), c0.size()));
  for (int id = 0; id < KW_ARRAY_COUNT; ++id)
    if (id < KW_X_SHIFT_NEG_R || id > KW_Z_SHIFT_NEG_R) std::vector<float>().swap(c->h_in[id]);  // the shift vectors are extended below
  // --- validation of what the loop needs
  if (c->count[KW_RHO0] == 0) return fail(KW_ERR_INVALID, "rho0 missing");
  if (cf.nonlinear_flag && c->count[KW_BONA] == 0) return fail(KW_ERR_INVALID, "BonA missing (nonlinear_flag = 1)");
  for (int id : {KW_DDX_K_SHIFT_POS_R, KW_DDY_K_SHIFT_POS, KW_DDZ_K_SHIFT_POS, KW_DDX_K_SHIFT_NEG_R, KW_DDY_K_SHIFT_NEG,
                 KW_DDZ_K_SHIFT_NEG, KW_PML_X_SGX, KW_PML_Y_SGY, KW_PML_Z_SGZ, KW_PML_X, KW_PML_Y, KW_PML_Z})
    if (!c->d[id]) return fail(KW_ERR_INVALID, "k-space shift vector or PML vector missing (array id " + std::to_string(id) + ")");
  if (cf.p_source_flag && (!c->di[KW_P_SOURCE_INDEX] || !c->d[KW_P_SOURCE_INPUT])) return fail(KW_ERR_INVALID, "p source arrays missing");
  if ((cf.ux_source_flag || cf.uy_source_flag || cf.uz_source_flag || cf.transducer_source_flag) && !c->di[KW_U_SOURCE_INDEX])
    return fail(KW_ERR_INVALID, "u_source_index missing");
  if (cf.ux_source_flag && !c->d[KW_UX_SOURCE_INPUT]) return fail(KW_ERR_INVALID, "ux_source_input missing");
  if (cf.uy_source_flag && !c->d[KW_UY_SOURCE_INPUT]) return fail(KW_ERR_INVALID, "uy_source_input missing");
  if (cf.uz_source_flag && !c->d[KW_UZ_SOURCE_INPUT]) return fail(KW_ERR_INVALID, "uz_source_input missing");
  if (cf.transducer_source_flag && (!c->d[KW_TRANSDUCER_SOURCE_INPUT] || !c->di[KW_DELAY_MASK]))
    return fail(KW_ERR_INVALID, "transducer source arrays missing");
  if (cf.p0_source_flag && !c->d[KW_P0_SOURCE_INPUT]) return fail(KW_ERR_INVALID, "p0_source_input missing");
  {  // every signal must cover the steps it is read at: signal[t] (single), signal[t * Nsrc + j] (many), signal[delay[j] + t]
     // (SolverCudaKernels.cu:463-471, :504-527, :570-629); the device kernels do not check bounds
    const uint64_t nt = cf.nt;
    auto need = [&](uint64_t flag, int many, int index_id) -> uint64_t {
      const uint64_t steps = std::min<uint64_t>(flag, nt);
      return many ? steps * c->count_total[index_id] : steps;
    };
    if (cf.p_source_flag && c->count[KW_P_SOURCE_INPUT] < need(cf.p_source_flag, cf.p_source_many, KW_P_SOURCE_INDEX))
      return fail(KW_ERR_INVALID, "p_source_input shorter than min(p_source_flag, Nt) [* Nsrc]");
    const uint64_t uflags[3] = {cf.ux_source_flag, cf.uy_source_flag, cf.uz_source_flag};
    for (int k = 0; k < 3; ++k)
      if (uflags[k] && c->count[KW_UX_SOURCE_INPUT + k] < need(uflags[k], cf.u_source_many, KW_U_SOURCE_INDEX))
        return fail(KW_ERR_INVALID, "u" + std::string(1, "xyz"[k]) + "_source_input shorter than min(flag, Nt) [* Nsrc]");
    if (cf.transducer_source_flag) {
      if (c->count_total[KW_DELAY_MASK] != c->count_total[KW_U_SOURCE_INDEX]) return fail(KW_ERR_INVALID, "delay_mask and u_source_index differ in length");
      uint64_t dmax = 0;
      for (auto v : c->h_idx[KW_DELAY_MASK]) dmax = std::max(dmax, v);
      if (!c->h_idx[KW_DELAY_MASK].empty() && c->count[KW_TRANSDUCER_SOURCE_INPUT] < dmax + std::min<uint64_t>(cf.transducer_source_flag, nt))
        return fail(KW_ERR_INVALID, "transducer_source_input shorter than max(delay_mask) + min(transducer_source_flag, Nt)");
    }
  }
  // --- state and temporaries (state starts at zero, BaseFloatMatrix.cpp:144-145)
  for (int id : {KW_P, KW_RHOX, KW_RHOY, KW_RHOZ, KW_UX_SGX, KW_UY_SGY, KW_UZ_SGZ})
    if (!c->d[id]) {
      KW_TRY(dalloc(c, (void**)&c->d[id], g.n * sizeof(float)));
      c->count[id] = g.n;
    }
  if (g.nranks == 1) {
    for (int k = 0; k < 4; ++k) KW_TRY(dalloc(c, (void**)&c->S[k], g.nc * sizeof(float2)));
  } else {  // one arena: S[0..3] then R[0..3], so that one IPC handle per rank maps every exchange buffer
    const size_t per = (g.nc * sizeof(float2) + 255) / 256 * 256;
    KW_TRY(dalloc(c, (void**)&c->arena, kPeerBufs * per));
    for (int k = 0; k < 4; ++k) c->S[k] = reinterpret_cast<float2*>(c->arena + k * per), c->R[k] = reinterpret_cast<float2*>(c->arena + (4 + k) * per);
    KW_CUDA(cudaStreamSynchronize(c->st));
    const char* env = getenv("KW_PEER");
    const bool want = !env || atoi(env) != 0;
    int* dflag = nullptr;
    KW_CUDA(cudaMalloc(&dflag, sizeof(int)));
    int agree_err = 0;
    auto agree = [&](bool ok) -> bool {  // min over the ranks of `ok`: the run's NCCL communicator is the one thing all ranks share for sure
      int v = ok ? 1 : 0;
      if (cudaMemcpyAsync(dflag, &v, sizeof v, cudaMemcpyHostToDevice, c->cs) != cudaSuccess) agree_err = 1;
      if (nccl_api().AllReduce(dflag, dflag, 1, kNcclInt32, kNcclMin, c->comm, c->cs) != kNcclSuccess) agree_err = 1;
      if (cudaMemcpyAsync(&v, dflag, sizeof v, cudaMemcpyDeviceToHost, c->cs) != cudaSuccess) agree_err = 1;
      if (cudaStreamSynchronize(c->cs) != cudaSuccess) agree_err = 1;
      return !agree_err && v == 1;
    };
    const bool peer_ok = c->peer.setup(c->nccl_id, sizeof(c->nccl_id), g.rank, g.nranks, c->arena, want, agree);
    cudaFree(dflag);
    if (agree_err) return fail(KW_ERR_COMM, "NCCL all-reduce failed while agreeing on the exchange path");
    if (!peer_ok && getenv("KW_PEER_VERBOSE"))
      fprintf(stderr, "kwave_b200 rank %d: peer-memory exchange unavailable (%s); using NCCL send/recv\n", g.rank, c->peer.error.c_str());
  }
  if (cf.absorbing_flag) {
    KW_TRY(dalloc(c, (void**)&c->tA, g.n * sizeof(float)));
    KW_TRY(dalloc(c, (void**)&c->tB, g.n * sizeof(float)));
    if (cf.nonlinear_flag) KW_TRY(dalloc(c, (void**)&c->tNL, g.n * sizeof(float)));
  }
  if (c->d[KW_SOURCE_KAPPA]) KW_TRY(dalloc(c, (void**)&c->tSrc, g.n * sizeof(float)));
  // --- streams that exist only because others need them (OutputStreamContainer.cpp:273-323): I_avg_c reads the
  //     frames of p_c and u*_non_staggered_c
  if (c->streams[KW_S_Q_TERM_C].enabled)  // Q_term_c is computed from the three I_avg_c (cpp:1013-1020)
    for (int k = 0; k < (g.nz == 1 ? 2 : 3); ++k) {
      Stream& d = c->streams[KW_S_IX_AVG_C + k];
      if (d.enabled) continue;
      d.enabled = true, d.nosave = true, d.op = kOpIAvgC, d.src = k, d.all = false;
    }
  for (int k = 0; k < 3; ++k)
    if (c->streams[KW_S_IX_AVG_C + k].enabled)
      for (int dep : {(int)KW_S_P_C, (int)KW_S_UX_NS_C + k}) {
        Stream& d = c->streams[dep];
        if (d.enabled) continue;
        d.enabled = true, d.nosave = true, d.op = kStreamTable[dep].op, d.src = kStreamTable[dep].src, d.all = false;
      }
  bool any_c = false;
  for (int sid = 0; sid < KW_STREAM_COUNT; ++sid) {
    Stream& s = c->streams[sid];
    if (!s.enabled) continue;
    any_c |= s.op == kOpC;
    c->need_shifted |= s.src >= 4 && s.op != kOpIAvgC;
    s.shifted_bases = sid >= KW_S_UX_NS_C && sid <= KW_S_UZ_NS_C;
  }
  if (c->need_shifted) {  // non-staggered velocity (cpp:2714-2735)
    const int nfull[3] = {g.nxp, g.ny, g.nz};
    for (int k = 0; k < 3; ++k) {
      auto& h = c->h_in[KW_X_SHIFT_NEG_R + k];
      if (h.empty()) return fail(KW_ERR_INVALID, "x/y/z_shift_neg_r missing (needed by non-staggered velocity outputs, MatrixContainer.cpp:375-384)");
      std::vector<float2> full(nfull[k], make_float2(0.f, 0.f));
      if (k == 0) {
        for (int i = 0; i < g.nxr; ++i) full[i] = make_float2(h[2 * i], h[2 * i + 1]);
      } else {
        // the reference transforms real lines (R2C / C2R along y or z): for the complex lines of the half spectrum the same
        // operator is its Hermitian extension, with the real part only where C2R ignores the imaginary one (DC, Nyquist)
        const int n = nfull[k], half = n / 2;
        for (int i = 0; i <= half; ++i) {
          const bool edge = i == 0 || (i == half && n % 2 == 0);
          full[i] = make_float2(h[2 * i], edge ? 0.f : h[2 * i + 1]);
          if (i > 0 && i < n - i) full[n - i] = make_float2(h[2 * i], -h[2 * i + 1]);
        }
      }
      KW_TRY(dalloc(c, (void**)&c->shift_full[k], full.size() * sizeof(float2), false));
      KW_CUDA(cudaMemcpyAsync(c->shift_full[k], full.data(), full.size() * sizeof(float2), cudaMemcpyHostToDevice, c->st));
      KW_CUDA(cudaStreamSynchronize(c->st));
      KW_TRY(dalloc(c, (void**)&c->d[KW_UX_SHIFTED + k], g.n * sizeof(float)));
      c->count[KW_UX_SHIFTED + k] = g.n;
    }
  }
  const uint64_t nsamp_c = cf.nt > cf.sampling_start_index ? cf.nt - cf.sampling_start_index : 0;
  if (any_c) {  // CompressHelper::init + generateFunctions (Compression/CompressHelper.cpp:48-65, :672-778), FP32
    if (!(cf.c_period > 0.f) || cf.c_mos == 0 || cf.c_harmonics == 0) return fail(KW_ERR_INVALID, "compression needs c_period > 0, c_mos >= 1, c_harmonics >= 1");
    c->c_osize = (int)(cf.c_period * (float)cf.c_mos);
    if (c->c_osize < 1) return fail(KW_ERR_INVALID, "compression: period * mos must be at least one step");
    c->c_bsize = 2 * c->c_osize + 1;
    c->c_H = (int)cf.c_harmonics;
    c->c_no_overlap = cf.c_no_overlap != 0 || cf.c_period >= (float)nsamp_c;  // Parameters.cpp:141-145
    const int os = c->c_osize, bs = c->c_bsize, H = c->c_H;
    std::vector<float> win(bs);
    for (int x = 0; x < bs; ++x) win[x] = x < os ? (float)x / (float)os : 2.0f - (float)x / (float)os;
    for (int sh = 0; sh < 2; ++sh) {
      std::vector<float2> e((size_t)H * bs), be((size_t)H * bs), be1((size_t)H * bs);
      for (int ih = 0; ih < H; ++ih) {
        const float w = 2.0f * (float)M_PI / (cf.c_period / (float)(ih + 1));
        const float sa = (float)M_PI / (cf.c_period / (float)(ih + 1));
        const float cs = cosf(sa), sn = sinf(sa);
        for (int x = 0; x < bs; ++x) {
          const float ang = w * (float)x;
          float2 v = make_float2(cosf(ang), -sinf(ang));  // exp(-i w x)
          if (sh) v = make_float2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);  // * exp(+i pi h / period)
          e[(size_t)ih * bs + x] = v;
        }
        const float norm = 2.0f / (float)os;
        for (int x = 0; x < bs; ++x) {
          const int x1 = (x + os) % (bs - 1);
          const float2 a = e[(size_t)ih * bs + x], b = e[(size_t)ih * bs + x1];
          be[(size_t)ih * bs + x] = make_float2(win[x] * a.x * norm, win[x] * a.y * norm);
          be1[(size_t)ih * bs + x] = make_float2(win[x1] * b.x * norm, win[x1] * b.y * norm);
        }
      }
      KW_TRY(dalloc(c, (void**)&c->d_be[sh], be.size() * sizeof(float2), false));
      KW_TRY(dalloc(c, (void**)&c->d_be1[sh], be1.size() * sizeof(float2), false));
      KW_CUDA(cudaMemcpyAsync(c->d_be[sh], be.data(), be.size() * sizeof(float2), cudaMemcpyHostToDevice, c->st));
      KW_CUDA(cudaMemcpyAsync(c->d_be1[sh], be1.data(), be1.size() * sizeof(float2), cudaMemcpyHostToDevice, c->st));
      KW_CUDA(cudaStreamSynchronize(c->st));
    }
  }
  // --- sensors and streams
  bool any_sensor_stream = false;
  for (int s = 0; s < KW_STREAM_COUNT; ++s) any_sensor_stream |= c->streams[s].enabled && !c->streams[s].all;
  if (any_sensor_stream) {
    if (cf.sensor_mask_type == 0) {
      if (!c->di[KW_SENSOR_MASK_INDEX]) return fail(KW_ERR_INVALID, "sensor_mask_index missing");
      c->nsens = c->count[KW_SENSOR_MASK_INDEX];
      c->nsens_total = c->count_total[KW_SENSOR_MASK_INDEX];
    } else {
      auto& h = c->h_idx[KW_SENSOR_MASK_CORNERS];
      if (h.empty() || h.size() % 6) return fail(KW_ERR_INVALID, "sensor_mask_corners missing or not a multiple of 6");
      c->ncuboids = (int)(h.size() / 6);
      // every cuboid is clipped to the local slab (a z-range of a cuboid is a contiguous range of its x-fastest buffer)
      std::vector<uint64_t> off(c->ncuboids + 1, 0), loc(h.size(), 0);
      uint64_t goff = 0;
      for (int k = 0; k < c->ncuboids; ++k) {
        const uint64_t* q = &h[6 * k];
        if (q[3] < q[0] || q[4] < q[1] || q[5] < q[2] || q[3] >= (uint64_t)g.nx || q[4] >= (uint64_t)g.ny || q[5] >= (uint64_t)g.nz)
          return fail(KW_ERR_INVALID, "sensor_mask_corners: cuboid outside the grid");
        const uint64_t cxy = (q[3] - q[0] + 1) * (q[4] - q[1] + 1);
        const uint64_t zlo = std::max<uint64_t>(q[2], (uint64_t)g.z0), zhi = std::min<uint64_t>(q[5], (uint64_t)(g.z0 + g.nzl - 1));
        uint64_t* l = &loc[6 * k];
        l[0] = q[0], l[1] = q[1], l[3] = q[3], l[4] = q[4];
        if (zlo <= zhi) {
          l[2] = zlo - g.z0, l[5] = zhi - g.z0;
          off[k + 1] = off[k] + cxy * (zhi - zlo + 1);
          c->sens_ranges.push_back(goff + cxy * (zlo - q[2]));
          c->sens_ranges.push_back(cxy * (zhi - zlo + 1));
        } else {
          l[2] = l[5] = 0;  // empty part: zero length in the offsets, never addressed
          off[k + 1] = off[k];
        }
        goff += cxy * (q[5] - q[2] + 1);
      }
      c->nsens = off.back();
      c->nsens_total = goff;
      KW_TRY(dalloc(c, (void**)&c->cub_offsets, off.size() * sizeof(uint64_t), false));
      KW_TRY(dalloc(c, (void**)&c->cub_corners, loc.size() * sizeof(uint64_t), false));
      KW_CUDA(cudaMemcpyAsync(c->cub_offsets, off.data(), off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, c->st));
      KW_CUDA(cudaMemcpyAsync(c->cub_corners, loc.data(), loc.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, c->st));
      KW_CUDA(cudaStreamSynchronize(c->st));
      c->h_idx[KW_SENSOR_MASK_CORNERS] = loc;  // fused_p_sample reads the (local) corners of a single cuboid from here
      c->count[KW_SENSOR_MASK_CORNERS] = loc.size();
    }
  }
  const uint64_t nsamp = cf.nt > cf.sampling_start_index ? cf.nt - cf.sampling_start_index : 0;
  for (int sid = 0; sid < KW_STREAM_COUNT; ++sid) {
    Stream& s = c->streams[sid];
    if (!s.enabled) continue;
    s.row = s.all ? g.n : c->nsens;
    if (s.op == kOpC) {
      const size_t H = c->c_H;
      // BaseOutputStream.cpp:98-101 / IndexOutputStream.cpp:87-125: ceil(Nsens * complexSize) * harmonics floats per frame
      s.row = cf.c_40bit ? (size_t)ceilf((float)c->nsens * 1.25f) * H : 2 * c->nsens * H;
      s.acc_bytes = std::max<size_t>(s.row * sizeof(float), c->nsens * H * (cf.c_40bit ? 5 : 8));
      KW_TRY(dalloc(c, (void**)&s.sbuf, c->nsens * sizeof(float)));
      KW_TRY(dalloc(c, &s.acc1, s.acc_bytes));
      if (c->c_no_overlap) s.acc2 = s.acc1;  // BaseOutputStream.cpp:246-257
      else KW_TRY(dalloc(c, &s.acc2, s.acc_bytes));
      const uint64_t nframes = std::max<uint64_t>(nsamp / (uint64_t)c->c_osize, 1);
      uint64_t cap = cf.raw_rows_capacity ? cf.raw_rows_capacity : std::max<uint64_t>(1, std::min<uint64_t>(nframes, (256ull << 20) / (s.row * sizeof(float) + 1)));
      s.cap_rows = s.nosave ? 0 : cap;
      if (!s.nosave) KW_TRY(dalloc(c, (void**)&s.dbuf, s.cap_rows * s.row * sizeof(float)));
      if (!s.nosave) KW_TRY(async_buffers(c, s));
    } else if (s.op == kOpIAvgC || s.op == kOpQTermC) {
      s.cap_rows = 1;
      KW_TRY(dalloc(c, (void**)&s.dbuf, s.row * sizeof(float)));
    } else if (s.op == kOpNone) {
      uint64_t cap = cf.raw_rows_capacity;
      if (cap == 0) cap = std::max<uint64_t>(1, std::min<uint64_t>(nsamp ? nsamp : 1, (256ull << 20) / (s.row * sizeof(float) + 1)));
      s.cap_rows = cap;
      KW_TRY(dalloc(c, (void**)&s.dbuf, s.cap_rows * s.row * sizeof(float)));
      KW_TRY(async_buffers(c, s));
    } else {
      s.cap_rows = 1;
      KW_TRY(dalloc(c, (void**)&s.dbuf, s.row * sizeof(float)));
      const float init = s.op == kOpMax ? -FLT_MAX : s.op == kOpMin ? FLT_MAX : 0.f;  // BaseOutputStream.cpp:338-366
      k_fill<<<ew_grid(s.row), 256, 0, c->st>>>(s.dbuf, init, s.row);
      c->launches++;
    }
  }
  KW_CUDA(cudaStreamSynchronize(c->st));
  KW_CUDA(cudaGetLastError());
  c->preprocessed = true;
  return KW_OK;
}

}  // extern "C"

// =====================================================================================================================
// the time step
// =====================================================================================================================
namespace kw {

// ---- building blocks of a (possibly slab-decomposed) 3-D transform -------------------------------------------------
// forward:  real slab --x pass, y pass--> x/y-local spectrum (y-blocked layout) --all-to-all--> z-local spectrum
//           [kz][ky_local][NXP] --k_zmid (forward z, operator, inverse z)--> --all-to-all--> --y pass, x pass + epilogue
// On one GPU the exchanges vanish and both layouts are the plain [z][y][NXP].

static ColArgs ycol_args(const Geometry& g, float2* const* data, int nf) {
  ColArgs ca{};
  for (int f = 0; f < nf; ++f) ca.data[f] = data[f];
  ca.stride = g.nxp, ca.outer_stride = (size_t)g.nyl * g.nxp, ca.ngroups = g.nxp / g.oy->col_w;
  ca.tile_begin = 0, ca.tile_end = g.nzl * ca.ngroups, ca.n = g.ny, ca.nvalid = g.nxr;
  if (g.nranks > 1) {
    int wk_log2 = 0;
    while ((1 << wk_log2) < g.oy->col_wk) ++wk_log2;
    ca.blk_es = g.ysh - wk_log2, ca.blk = g.blk, ca.nyl = g.nyl;
  }
  return ca;
}

// forward x and y passes of `nf` real fields into spectral buffers (x/y-local side)
static void forward_xy(kw_ctx* c, const float* const* in, float2* const* out, int nf) {
  const Geometry& g = c->g;
  XFwdArgs xa{};
  for (int f = 0; f < nf; ++f) xa.in[f] = in[f], xa.out[f] = out[f];
  xa.tab = g.tx, xa.nxp = g.nxp, xa.map = g.row_map(), xa.n = g.nx;
  // (z-chunked launches that keep a chunk's half spectra in L2 between the x and the y pass were measured and lose: 16.8 / 13.5 /
  //  12.1 ms per step at 24 / 48 / 96 MB chunks against 10.4 ms for whole-grid launches, profiles/r02_e_chunk_sweep.log)
  xa.pair_begin = 0, xa.pair_end = g.nzl * g.ny / 2;
  const ColArgs ca = ycol_args(g, out, nf);
  launch(c, "xfwd", nf * (4.0 * g.n + 8.0 * g.nca), [&] { g.ox->xfwd(xa, nf, c->st); });
  launch(c, "ycol_fwd", nf * 16.0 * g.nca, [&] { g.oy->col(ca, -1, nf, c->st); });
}
// inverse y pass then the x inverse with its fused epilogue; xinv(pair_begin, pair_end) launches it
template <class F>
static void inverse_yx(kw_ctx* c, float2* const* data, int nf, const char* name, double xinv_bytes, F&& xinv) {
  const Geometry& g = c->g;
  const ColArgs ca = ycol_args(g, data, nf);
  launch(c, "ycol_inv", nf * 16.0 * g.nca, [&] { g.oy->col(ca, +1, nf, c->st); });
  launch(c, name, xinv_bytes, [&] { xinv(0, g.nzl * g.ny / 2); });
}

// ---- all-to-all between the x/y-local and the z-local side --------------------------------------------------------
// The same contiguous-block exchange in both directions: block q of the source goes to rank q, block q of the destination
// comes from rank q.  Exchange buffers are addressed by id (S[k] -> k, R[k] -> 4 + k).  Exchanges run on the
// communication stream `cs`; they start once everything enqueued on the solver stream so far has finished and hand back
// an event the consumer waits for.
//  * peer path (PeerLink): each rank pushes its blocks into the destination buffers of its peers with copy-engine copies
//    over NVLink and announces them with stream memory operations on flags every rank maps; no SM is involved.  A rank
//    may only be pushed to once it has said (release_buffer) that it no longer reads the previous content.
//  * fallback: one NCCL group of send/recv pairs.
static float2* xbuf(kw_ctx* c, int id) { return id < 4 ? c->S[id] : c->R[id - 4]; }
static int xbuf_id(kw_ctx* c, const float2* p) {
  for (int k = 0; k < 4; ++k) {
    if (p == c->S[k]) return k;
    if (p == c->R[k]) return 4 + k;
  }
  return -1;
}
static int exchange_async(kw_ctx* c, int src_id, int dst_id, cudaEvent_t* done) {
  const Geometry& g = c->g;
  const int me = g.rank, P = g.nranks;
  float2 *src = xbuf(c, src_id), *dst = xbuf(c, dst_id);
  const size_t bytes = g.blk * sizeof(float2);
  const cudaEvent_t start = mark(c, c->st);
  cudaStreamWaitEvent(c->cs, start, 0);
  int e = 0;
  std::string what;
  const double moved = 16.0 * (double)g.blk * (P - 1);
  if (c->peer.active) {
    PeerLink& pl = c->peer;
    const uint32_t n = ++pl.use[dst_id];
    const size_t dst_off = reinterpret_cast<char*>(dst) - c->arena;
    // (profiling brackets the four parts separately: credit wait / copies / signal / arrival wait)
    launch(c, "all_to_all_credit_wait", 0.0, [&] {
      if (n > 1)  // every receiver has released the previous content of its destination buffer
        e = pl.flag_ops(c->cs, true, n - 1, [&](int q) { return &pl.shm->credit[me][dst_id][q]; });
    }, c->cs);
    cudaEvent_t self_done = nullptr;
    if (!e && c->ws && src != dst) {  // the own block is a local copy: off the copy stream that feeds NVLink
      cudaStreamWaitEvent(c->ws, start, 0);
      e = (int)cudaMemcpyAsync(dst + (size_t)me * g.blk, src + (size_t)me * g.blk, bytes, cudaMemcpyDeviceToDevice, c->ws);
      self_done = mark(c, c->ws);
    }
    launch(c, "all_to_all", moved, [&] {
      for (int k = 1; k < P && !e; ++k) {  // staggered: at any time every rank receives from one sender
        const int q = (me + k) % P;
        e = (int)cudaMemcpyAsync(pl.peer_base[q] + dst_off + (size_t)me * bytes, src + (size_t)q * g.blk, bytes, cudaMemcpyDeviceToDevice, c->cs);
      }
      if (!e && !c->ws && src != dst) e = (int)cudaMemcpyAsync(dst + (size_t)me * g.blk, src + (size_t)me * g.blk, bytes, cudaMemcpyDeviceToDevice, c->cs);
    }, c->cs);
    launch(c, "all_to_all_arrival_wait", 0.0, [&] {
      e = pl.flag_ops(c->cs, false, n, [&](int q) { return &pl.shm->arrived[q][dst_id][me]; });
      if (!e) e = pl.flag_ops(c->cs, true, n, [&](int r) { return &pl.shm->arrived[me][dst_id][r]; });
      if (self_done) cudaStreamWaitEvent(c->cs, self_done, 0);
    }, c->cs);
    if (e) what = "peer-memory all-to-all: CUDA error " + std::to_string(e);
  } else {
    NcclApi& nc = nccl_api();
    launch(c, "all_to_all", moved, [&] {
      e = nc.GroupStart();
      for (int q = 0; q < P && e == kNcclSuccess; ++q) {
        e = nc.Send(src + (size_t)q * g.blk, 2 * g.blk, kNcclFloat, q, c->comm, c->cs);
        if (e == kNcclSuccess) e = nc.Recv(dst + (size_t)q * g.blk, 2 * g.blk, kNcclFloat, q, c->comm, c->cs);
      }
      const int e2 = nc.GroupEnd();
      if (e == kNcclSuccess) e = e2;
    }, c->cs);
    if (e != kNcclSuccess) what = std::string("NCCL all-to-all: ") + nc.GetErrorString(e);
  }
  if (e) return fail(KW_ERR_COMM, what);
  c->comm_bytes += 8.0 * (double)g.blk * (P - 1);
  *done = mark(c, c->cs);
  return KW_OK;
}
// this rank has finished reading the current content of exchange buffer `id` (work enqueued on `s` so far): its peers may
// push the next content.  Must follow every use of a buffer as a destination.
static int release_buffer(kw_ctx* c, int id, cudaStream_t s) {
  if (!c->peer.active) return KW_OK;
  PeerLink& pl = c->peer;
  const uint32_t n = pl.use[id];
  if (pl.flag_ops(s, false, n, [&](int r) { return &pl.shm->credit[r][id][pl.rank]; }))
    return fail(KW_ERR_COMM, "peer-memory all-to-all: cuStreamWriteValue32 failed");
  return KW_OK;
}
// blocking form used by the rarely taken paths (additive sources): result[f] = the buffers that hold the exchanged spectra
static int exchange(kw_ctx* c, float2* const* src, float2* const* dst, int nf, float2** result) {
  const Geometry& g = c->g;
  if (g.nranks == 1) {
    for (int f = 0; f < nf; ++f) result[f] = src[f];
    return KW_OK;
  }
  for (int f = 0; f < nf; ++f) {
    cudaEvent_t ev;
    KW_TRY(exchange_async(c, xbuf_id(c, src[f]), xbuf_id(c, dst[f]), &ev));
    cudaStreamWaitEvent(c->st, ev, 0);
    result[f] = dst[f];
  }
  return KW_OK;
}

// Nz == 1: out = (in * (mul * scal)) (x) vec[coordinate of the axis]; the z-indexed operator has one entry
struct ZMulArgs {
  ZField f;
  int axis, nxp;
  size_t n;  // NXP * Ny
};
static __global__ void k_zmul(ZMulArgs a) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x * blockDim.x) {
    const int kx = (int)(i % a.nxp), y = (int)(i / a.nxp);
    const float2 e = cscale(a.f.in[i], a.f.mul ? a.f.mul[i] * a.f.scal : a.f.scal);
    if (a.axis == 3) {
      a.f.out[i] = cmul(e, a.f.vec[kx]);
      a.f.out_y[i] = cmul(e, a.f.vec_y[y]);
      a.f.out_z[i] = cmul(e, a.f.vec_z[0]);
    } else {
      a.f.out[i] = a.axis < 0 ? e : cmul(e, a.f.vec[a.axis == 0 ? kx : a.axis == 1 ? y : 0]);
    }
  }
}

// one fused z pass per field on the z-local side (axis < 0: no 1-D operator)
static void zmid_launch(kw_ctx* c, ZField f, int axis) {
  const Geometry& g = c->g;
  if (g.nz == 1) {  // 2-D: forward z, operator, inverse z collapse to the operator
    ZMulArgs ma{f, axis, g.nxp, (size_t)g.nxp * g.ny};
    launch(c, axis == 3 ? "zmul_grad" : "zmul", (axis == 3 ? 32.0 : 16.0) * g.nca + (f.mul ? 4.0 * g.nca : 0.0),
           [&] { k_zmul<<<ew_grid(ma.n), 256, 0, c->st>>>(ma); });
    return;
  }
  ZMidArgs za{};
  // the ky-indexed operators are looked up with the local ky: shift them to this rank's range
  if (axis == 1 && f.vec) f.vec += g.y0;
  if (axis == 3 && f.vec_y) f.vec_y += g.y0;
  za.f = f, za.axis = axis;
  za.nxp = g.nxp, za.ngroups = g.nxp / g.oz->zmid_w, za.ntiles = g.nyl * za.ngroups, za.plane = (unsigned)((size_t)g.nyl * g.nxp), za.n = g.nz, za.nvalid = g.nxr;
  launch(c, axis == 3 ? "zmid_grad" : "zmid", (axis == 3 ? 32.0 : 16.0) * g.nca + (f.mul ? 4.0 * g.nca : 0.0), [&] { g.oz->zmid(za, c->st); });
}
template <int NF> static XInvArgs<NF> xinv_args(kw_ctx* c, float2* const* in, int pb, int pe, int nfields = NF) {
  XInvArgs<NF> a{};
  for (int f = 0; f < nfields; ++f) a.in[f] = in[f];
  a.tab = c->g.tx, a.pair_begin = pb, a.pair_end = pe, a.nxp = c->g.nxp, a.ny = c->g.ny, a.map = c->g.row_map(), a.n = c->g.nx;
  return a;
}

// F[p]*kappa*ddk_pos after the fused z pass; grad[0..2] are ready for the inverse y and x passes  (cpp:2087-2101)
static int pressure_gradient_spectra(kw_ctx* c, float2** grad) {
  const float* in[1] = {c->d[KW_P]};
  float2* out[1] = {c->S[3]};
  forward_xy(c, in, out, 1);
  float2* zin[1];
  KW_TRY(exchange(c, out, &c->R[3], 1, zin));
  ZField zf{zin[0], c->S[0], c->d[KW_KAPPA], 1.0f, reinterpret_cast<const float2*>(c->d[KW_DDX_K_SHIFT_POS_R]),
            c->S[1], c->S[2], reinterpret_cast<const float2*>(c->d[KW_DDY_K_SHIFT_POS]),
            reinterpret_cast<const float2*>(c->d[KW_DDZ_K_SHIFT_POS])};
  zmid_launch(c, zf, 3);
  return exchange(c, c->S, c->R, 3, grad);
}

// additive source: scaled = IFFT(FFT(scatter) * (source_kappa * fd)), added to the targets  (cpp:2339-2352)
static int add_scaled_source(kw_ctx* c, const float* signal, int index_id, int many, float* const* targets, int ntargets) {
  const Geometry& g = c->g;
  const size_t nsrc = c->count[index_id];
  launch(c, "memset_source_grid", 4.0 * g.n, [&] { cudaMemsetAsync(c->tSrc, 0, g.n * sizeof(float), c->st); });
  launch(c, "insert_source", 16.0 * nsrc, [&] {
    k_insert_source<<<ew_grid(nsrc), 256, 0, c->st>>>(c->tSrc, signal, c->di[index_id], c->dpos[index_id], nsrc, c->count_total[index_id], c->t, many);
  });
  const float* in[1] = {c->tSrc};
  float2* out[1] = {c->S[3]};
  forward_xy(c, in, out, 1);
  float2 *zb[1], *back[1];
  KW_TRY(exchange(c, out, &c->R[3], 1, zb));
  zmid_launch(c, ZField{zb[0], zb[0], c->d[KW_SOURCE_KAPPA], 1.0f / (float)g.ntot, nullptr}, -1);
  KW_TRY(exchange(c, zb, &c->S[3], 1, back));
  if (g.nranks > 1) KW_TRY(release_buffer(c, 7, c->cs));  // R[3] has been pushed back
  EpiAdd e{};
  for (int k = 0; k < ntargets; ++k) e.out[k] = targets[k];
  e.ntargets = ntargets;
  inverse_yx(c, back, 1, "xinv_add_source", 8.0 * g.nca + 8.0 * g.n * ntargets,
             [&](int pb, int pe) { g.ox->xinv_add(xinv_args<1>(c, back, pb, pe), e, c->st); });
  if (g.nranks > 1) KW_TRY(release_buffer(c, 3, c->st));
  return KW_OK;
}

static TermsArgs terms_args(kw_ctx* c) {
  TermsArgs ta{};
  ta.rho[0] = c->d[KW_RHOX], ta.rho[1] = c->d[KW_RHOY], ta.rho[2] = c->d[KW_RHOZ];
  ta.rho0 = c->fld(KW_RHO0), ta.bona = c->fld(KW_BONA), ta.c2 = c->fld(KW_C0);
  ta.nonlinear = c->cfg.nonlinear_flag, ta.absorbing = c->cfg.absorbing_flag;
  ta.outB = c->tB, ta.outNL = c->tNL, ta.p = c->d[KW_P];
  return ta;
}

template <int OP> static void sample_one(kw_ctx* c, Stream& s, const float* src, float* dst) {
  const Geometry& g = c->g;
  const double per = OP == kOpNone ? 8.0 : 12.0;  // src read + buffer write (+ buffer read for aggregates)
  if (s.all) {
    launch(c, "sample_all", per * g.n, [&] { k_sample_all<OP><<<ew_grid(g.n), 256, 0, c->st>>>(dst, src, g.n); });
  } else if (c->nsens == 0) {
    return;  // no sensor point in this slab
  } else if (c->cfg.sensor_mask_type == 0) {
    launch(c, "sample_index", (per + 8.0) * c->nsens,
           [&] { k_sample_index<OP><<<ew_grid(c->nsens), 256, 0, c->st>>>(dst, src, c->di[KW_SENSOR_MASK_INDEX], c->nsens); });
  } else {
    CuboidArgs ca{c->cub_corners, c->cub_offsets, c->ncuboids, g.nx, g.ny};
    launch(c, "sample_cuboid", per * c->nsens, [&] { k_sample_cuboid<OP><<<ew_grid(c->nsens), 256, 0, c->st>>>(dst, src, ca, c->nsens); });
  }
}

// Aggregates of p that the pressure-producing epilogue can accumulate itself.  Returns false when nothing is fused.
static bool fused_p_sample(kw_ctx* c, FusedSample* fs, double* extra_bytes) {
  memset(fs, 0, sizeof(*fs));
  const kw_config& cf = c->cfg;
  if (c->t < cf.sampling_start_index) return false;
  if (c->t == 0 && cf.p0_source_flag == 1) return false;  // p is overwritten by the initial pressure afterwards
  bool any = false;
  *extra_bytes = 0;
  auto take = [&](int sid, float** slot, double bytes) {
    Stream& s = c->streams[sid];
    if (!s.enabled) return;
    *slot = s.dbuf, s.fused_this_step = true, any = true, *extra_bytes += bytes;
  };
  take(KW_S_P_MAX_ALL, &fs->max_all, 8.0 * c->g.n);
  take(KW_S_P_MIN_ALL, &fs->min_all, 8.0 * c->g.n);
  if (cf.sensor_mask_type == 1 && c->ncuboids == 1 && c->nsens > 0) {
    take(KW_S_P_RMS, &fs->rms, 8.0 * c->nsens);
    take(KW_S_P_MAX, &fs->mx, 8.0 * c->nsens);
    take(KW_S_P_MIN, &fs->mn, 8.0 * c->nsens);
    if (fs->rms || fs->mx || fs->mn) {
      const auto& h = c->h_idx[KW_SENSOR_MASK_CORNERS];  // clipped to the slab, local z
      fs->x0 = (int)h[0], fs->y0 = (int)h[1], fs->z0 = (int)h[2], fs->x1 = (int)h[3], fs->y1 = (int)h[4], fs->z1 = (int)h[5];
      fs->cub = (c->nsens == c->g.n) ? 1 : 2;
    }
  }
  return any;
}

// computeShiftedVelocity (cpp:2714-2735): u_i shifted by half a cell onto the pressure grid.  The reference runs 1-D
// R2C / multiply / C2R along each axis; here the same multiplier is applied between the z transforms of the 3-D pipeline
// (it commutes with the transforms along the other two axes), which also covers slab-decomposed grids.
static int add_scaled_source(kw_ctx* c, const float* signal, int index_id, int many, float* const* targets, int ntargets);
static int compute_shifted_velocity(kw_ctx* c) {
  const Geometry& g = c->g;
  const float fd = 1.0f / (float)g.ntot;
  for (int f = 0; f < 3; ++f) {
    const float* in[1] = {c->d[KW_UX_SGX + f]};
    float2* out[1] = {c->S[3]};
    forward_xy(c, in, out, 1);
    float2 *zb[1], *back[1];
    KW_TRY(exchange(c, out, &c->R[3], 1, zb));
    zmid_launch(c, ZField{zb[0], zb[0], nullptr, fd, c->shift_full[f]}, f);
    KW_TRY(exchange(c, zb, &c->S[3], 1, back));
    if (g.nranks > 1) KW_TRY(release_buffer(c, 7, c->cs));
    EpiStore e{};
    e.out[0] = c->d[KW_UX_SHIFTED + f], e.scale = 1.0f;
    inverse_yx(c, back, 1, "xinv_shifted_velocity", 8.0 * g.nca + 4.0 * g.n,
               [&](int pb, int pe) { g.ox->xinv_store(xinv_args<1>(c, back, pb, pe), e, 1, c->st); });
    if (g.nranks > 1) KW_TRY(release_buffer(c, 3, c->st));
  }
  return KW_OK;
}

static const float* stream_source(kw_ctx* c, const Stream& s) {
  return s.src == 0 ? c->d[KW_P] : s.src <= 3 ? c->d[KW_UX_SGX + (s.src - 1)] : c->d[KW_UX_SHIFTED + (s.src - 4)];
}

// OutputStreamContainer::sampleStreams + flushRawStreams (Containers/OutputStreamContainer.cpp:364-403): sampling in enum
// order; then, for the compressed streams, every flushRaw, every postSample (I_avg_c) and every postSample2 -- all on the
// device, in the step that produced the samples (the reference does this on the host, one step later).
static int sample_streams(kw_ctx* c) {
  const kw_config& cf = c->cfg;
  if (c->need_shifted) KW_TRY(compute_shifted_velocity(c));
  const uint64_t nsamp = cf.nt - cf.sampling_start_index;
  for (int sid = 0; sid < KW_STREAM_COUNT; ++sid) {
    Stream& s = c->streams[sid];
    if (!s.enabled || s.op == kOpIAvgC || s.op == kOpQTermC) continue;
    if (s.fused_this_step) {
      s.fused_this_step = false;
      continue;
    }
    const float* src = stream_source(c, s);
    switch (s.op) {
      case kOpNone: sample_one<kOpNone>(c, s, src, s.dbuf + s.rows * s.row); s.rows++; break;
      case kOpRms: sample_one<kOpRms>(c, s, src, s.dbuf); break;
      case kOpMax: sample_one<kOpMax>(c, s, src, s.dbuf); break;
      case kOpMin: sample_one<kOpMin>(c, s, src, s.dbuf); break;
      case kOpC: {  // IndexOutputStream::flushRaw :373-470 / CuboidOutputStream::flushRaw :431-532
        sample_one<kOpNone>(c, s, src, s.sbuf);
        const int step_local = (int)(s.sampled % (uint64_t)(c->c_bsize - 1));
        s.saving = (step_local + 1) % c->c_osize == 0;
        const bool odd = (s.compressed + 1) % 2 == 0;
        const bool mirror = s.compressed == 0 && s.saving && !c->c_no_overlap;
        const size_t n = c->nsens * (size_t)c->c_H;
        if (n) {
          CompressArgs a{};
          a.x = s.sbuf, a.n = n, a.H = c->c_H, a.bsize = c->c_bsize, a.step_local = step_local, a.mirror = mirror;
          a.be = c->d_be[s.shifted_bases], a.be1 = c->d_be1[s.shifted_bases];
          a.e = s.shifted_bases ? 114 : 138;  // CompressHelper::kMaxExpU / kMaxExpP
          if (cf.c_40bit) {
            a.q1 = static_cast<uint8_t*>(s.acc1), a.q2 = static_cast<uint8_t*>(s.acc2);
            launch(c, "compress", 4.0 * c->nsens + 20.0 * n, [&] { k_compress40<<<ew_grid(n), 256, 0, c->st>>>(a); });
          } else {
            a.acc1 = static_cast<float2*>(s.acc1), a.acc2 = static_cast<float2*>(s.acc2);
            launch(c, "compress", 4.0 * c->nsens + 32.0 * n, [&] { k_compress<<<ew_grid(n), 256, 0, c->st>>>(a); });
          }
        }
        const bool last = (nsamp - s.sampled == 1) && nsamp <= (uint64_t)c->c_osize;
        if (s.saving || last) {
          s.cur = odd ? s.acc1 : s.acc2;
          if (!s.nosave) {
            launch(c, "store_frame", 8.0 * s.row, [&] {
              cudaMemcpyAsync(s.dbuf + s.rows * s.row, s.cur, s.row * sizeof(float), cudaMemcpyDeviceToDevice, c->st);
            });
            s.rows++;
          }
          s.compressed++;
        }
        s.sampled++;
        break;
      }
      default: break;
    }
  }
  // postSample: I_avg_c (IndexOutputStream.cpp:299-342)
  for (int k = 0; k < 3; ++k) {
    Stream& s = c->streams[KW_S_IX_AVG_C + k];
    if (!s.enabled) continue;
    Stream &sp = c->streams[KW_S_P_C], &su = c->streams[KW_S_UX_NS_C + k];
    if (!sp.cur || !su.cur) continue;
    if (c->nsens) {
      const bool q = cf.c_40bit != 0;
      launch(c, "intensity_c", (8.0 + 16.0 * c->c_H) * c->nsens, [&] {
        k_intensity_c<<<ew_grid(c->nsens), 256, 0, c->st>>>(s.dbuf, q ? nullptr : static_cast<const float2*>(sp.cur), q ? nullptr : static_cast<const float2*>(su.cur),
                                                            q ? static_cast<const uint8_t*>(sp.cur) : nullptr, q ? static_cast<const uint8_t*>(su.cur) : nullptr,
                                                            c->nsens, c->c_H, 138, 114);
      });
    }
    s.compressed++;
  }
  // postSample2: the completed accumulator starts the next frame from zero (BaseOutputStream.cpp:117-133)
  for (auto& s : c->streams) {
    if (!s.enabled || s.op != kOpC) continue;
    if (s.saving && s.cur) launch(c, "zero_frame", 4.0 * s.row, [&] { cudaMemsetAsync(s.cur, 0, s.acc_bytes, c->st); });
    s.cur = nullptr;
  }
  return KW_OK;
}

static EpiVelocity velocity_epilogue(kw_ctx* c, float* const* u, float fd, int init) {
  EpiVelocity e{};
  for (int k = 0; k < 3; ++k) e.u[k] = u[k], e.dtrho[k] = c->fld(KW_RHO0_SGX + k), e.pml_sg[k] = c->d[KW_PML_X_SGX + k];
  e.pml_sg[2] += c->g.z0;  // the epilogues see local plane numbers
  e.fd = fd, e.init = init;
  return e;
}

static int step(kw_ctx* c) {
  const kw_config& cf = c->cfg;
  const Geometry& g = c->g;
  const uint64_t t = c->t;
  const float fd = 1.0f / (float)g.ntot;  // fftDivider, CudaParameters.cpp:259
  float* u[3] = {c->d[KW_UX_SGX], c->d[KW_UY_SGY], c->d[KW_UZ_SGZ]};
  float* rho[3] = {c->d[KW_RHOX], c->d[KW_RHOY], c->d[KW_RHOZ]};
  float2* sp[3];  // spectra ready for the inverse y/x passes of the current stage

  // ---- computeVelocity (cpp:2087-2119)
  KW_TRY(pressure_gradient_spectra(c, sp));
  {
    const EpiVelocity e = velocity_epilogue(c, u, fd, 0);
    const double het = c->count[KW_RHO0_SGX] > 1 ? 4.0 : 0.0;
    inverse_yx(c, sp, 3, "xinv_velocity", 3 * (8.0 * g.nca + (8.0 + het) * g.n),
               [&](int pb, int pe) { g.ox->xinv_velocity(xinv_args<1>(c, sp, pb, pe, 3), e, 3, c->st); });
  }
  // ---- addVelocitySource (cpp:2252-2303), transducer (cpp:894-897)
  const uint64_t uflag[3] = {cf.ux_source_flag, cf.uy_source_flag, cf.uz_source_flag};
  for (int k = 0; k < 3; ++k) {
    if (uflag[k] <= t) continue;
    const size_t nsrc = c->count[KW_U_SOURCE_INDEX];
    if (cf.u_source_mode != KW_SRC_ADDITIVE) {
      if (nsrc == 0) continue;
      SourceArgs sa{};
      sa.target[0] = u[k], sa.ntargets = 1, sa.signal = c->d[KW_UX_SOURCE_INPUT + k], sa.index = c->di[KW_U_SOURCE_INDEX];
      sa.pos = c->dpos[KW_U_SOURCE_INDEX], sa.nsrc = nsrc, sa.nsrc_total = c->count_total[KW_U_SOURCE_INDEX];
      sa.t = t, sa.many = cf.u_source_many, sa.mode = cf.u_source_mode;
      launch(c, "add_u_source", 20.0 * nsrc, [&] { k_add_source<<<ew_grid(nsrc), 256, 0, c->st>>>(sa); });
    } else {
      float* tg[1] = {u[k]};
      KW_TRY(add_scaled_source(c, c->d[KW_UX_SOURCE_INPUT + k], KW_U_SOURCE_INDEX, cf.u_source_many, tg, 1));
    }
  }
  if (cf.transducer_source_flag > t && c->count[KW_U_SOURCE_INDEX] > 0) {
    const size_t nsrc = c->count[KW_U_SOURCE_INDEX];
    launch(c, "add_transducer", 28.0 * nsrc, [&] {
      k_add_transducer<<<ew_grid(nsrc), 256, 0, c->st>>>(u[0], c->di[KW_U_SOURCE_INDEX], c->d[KW_TRANSDUCER_SOURCE_INPUT],
                                                          c->di[KW_DELAY_MASK], nsrc, t);
    });
  }
  // ---- computeVelocityGradient (cpp:2126-2150) + computeDensity (cpp:2157/2169) [+ pressure terms / lossless p]
  {
    forward_xy(c, u, c->S, 3);
    float2* zb[3];
    KW_TRY(exchange(c, c->S, c->R, 3, zb));
    const int vec[3] = {KW_DDX_K_SHIFT_NEG_R, KW_DDY_K_SHIFT_NEG, KW_DDZ_K_SHIFT_NEG};
    for (int f = 0; f < 3; ++f)
      zmid_launch(c, ZField{zb[f], zb[f], c->d[KW_KAPPA], fd, reinterpret_cast<const float2*>(c->d[vec[f]])}, f);
    KW_TRY(exchange(c, zb, c->S, 3, sp));
    const bool p_src = cf.p_source_flag > t;
    EpiDensity e{};
    for (int k = 0; k < 3; ++k) e.rho[k] = rho[k], e.pml[k] = c->d[KW_PML_X + k];
    e.pml[2] += g.z0;
    e.rho0 = c->fld(KW_RHO0), e.bona = c->fld(KW_BONA), e.c2 = c->fld(KW_C0);
    e.dt = cf.dt, e.nonlinear = cf.nonlinear_flag, e.absorbing = cf.absorbing_flag;
    e.defer_terms = p_src && cf.p_source_mode == KW_SRC_ADDITIVE;
    e.outA = c->tA, e.outB = c->tB, e.outNL = c->tNL, e.p = c->d[KW_P];
    double fused_bytes = 0;
    if (!cf.absorbing_flag && !p_src) e.sample = fused_p_sample(c, &e.fs, &fused_bytes);
    {
      double per = 24.0 + (c->count[KW_RHO0] > 1 ? 4.0 : 0.0);  // rho r/w, rho0
      if (cf.absorbing_flag) per += 4.0 + (e.defer_terms ? 0.0 : 4.0 + (cf.nonlinear_flag ? 4.0 : 0.0));  // A, B, NL
      else if (!e.defer_terms) per += 4.0 + (c->count[KW_C0] > 1 ? 4.0 : 0.0);                               // p, c2
      if (cf.nonlinear_flag && !e.defer_terms && c->count[KW_BONA] > 1) per += 4.0;
      inverse_yx(c, sp, 3, "xinv_density", 24.0 * g.nca + per * g.n + fused_bytes,
                 [&](int pb, int pe) { g.ox->xinv_density(xinv_args<3>(c, sp, pb, pe), e, c->st); });
    }
    // ---- addPressureSource (cpp:2310-2334)
    if (p_src) {
      const size_t nsrc = c->count[KW_P_SOURCE_INDEX];
      TermsArgs ta = terms_args(c);
      if (cf.p_source_mode != KW_SRC_ADDITIVE) {
        if (nsrc > 0) {
          SourceArgs sa{};
          for (int k = 0; k < 3; ++k) sa.target[k] = rho[k];
          sa.ntargets = g.nz == 1 ? 2 : 3 /* 2-D: rhox, rhoy only (SolverCudaKernels.cu:570-629) */, sa.signal = c->d[KW_P_SOURCE_INPUT], sa.index = c->di[KW_P_SOURCE_INDEX];
          sa.pos = c->dpos[KW_P_SOURCE_INDEX], sa.nsrc = nsrc, sa.nsrc_total = c->count_total[KW_P_SOURCE_INDEX];
          sa.t = t, sa.many = cf.p_source_many, sa.mode = cf.p_source_mode;
          launch(c, "add_p_source", 36.0 * nsrc, [&] { k_add_source<<<ew_grid(nsrc), 256, 0, c->st>>>(sa); });
          // the fused epilogue computed the sum-of-density terms before the source landed: redo them at the source voxels
          ta.index = c->di[KW_P_SOURCE_INDEX], ta.n = nsrc;
          launch(c, "pressure_terms_fixup", 40.0 * nsrc, [&] { k_pressure_terms<<<ew_grid(nsrc), 256, 0, c->st>>>(ta); });
        }
      } else {
        KW_TRY(add_scaled_source(c, c->d[KW_P_SOURCE_INPUT], KW_P_SOURCE_INDEX, cf.p_source_many, rho, g.nz == 1 ? 2 : 3));
        ta.index = nullptr, ta.n = g.n;
        launch(c, "pressure_terms", 28.0 * g.n, [&] { k_pressure_terms<<<ew_grid(g.n), 256, 0, c->st>>>(ta); });
      }
    }
  }
  // ---- computePressure, absorbing branch (cpp:2180-2246)
  if (cf.absorbing_flag) {
    const float* in[2] = {c->tA, c->tB};
    forward_xy(c, in, c->S, 2);
    float2* zb[2];
    KW_TRY(exchange(c, c->S, c->R, 2, zb));
    zmid_launch(c, ZField{zb[0], zb[0], c->d[KW_ABSORB_NABLA1], 1.0f, nullptr}, -1);
    zmid_launch(c, ZField{zb[1], zb[1], c->d[KW_ABSORB_NABLA2], 1.0f, nullptr}, -1);
    KW_TRY(exchange(c, zb, c->S, 2, sp));
    EpiPressureSum e{};
    e.p = c->d[KW_P], e.base = cf.nonlinear_flag ? c->tNL : c->tB;
    e.c2 = c->fld(KW_C0), e.tau = c->fld(KW_ABSORB_TAU), e.eta = c->fld(KW_ABSORB_ETA), e.fd = fd;
    const double per = 8.0 + (c->count[KW_C0] > 1 ? 4.0 : 0.0) + (c->count[KW_ABSORB_TAU] > 1 ? 8.0 : 0.0);
    double fused_bytes = 0;
    e.sample = fused_p_sample(c, &e.fs, &fused_bytes);
    inverse_yx(c, sp, 2, "xinv_pressure_sum", 16.0 * g.nca + per * g.n + fused_bytes,
               [&](int pb, int pe) { g.ox->xinv_psum(xinv_args<2>(c, sp, pb, pe), e, c->st); });
  }
  // ---- addInitialPressureSource (cpp:2359-2396)
  if (t == 0 && cf.p0_source_flag == 1) {
    launch(c, "initial_pressure", 24.0 * g.n, [&] {
      k_initial_pressure<<<ew_grid(g.n), 256, 0, c->st>>>(c->d[KW_P], rho[0], rho[1], rho[2], c->d[KW_P0_SOURCE_INPUT], c->fld(KW_C0), g.n, g.nz == 1 ? 2 : 3);
    });
    KW_TRY(pressure_gradient_spectra(c, sp));
    const EpiVelocity e = velocity_epilogue(c, u, fd, 1);
    inverse_yx(c, sp, 3, "xinv_initial_velocity", 3 * (8.0 * g.nca + 8.0 * g.n),
               [&](int pb, int pe) { g.ox->xinv_velocity(xinv_args<1>(c, sp, pb, pe, 3), e, 3, c->st); });
  }
  // ---- storeSensorData (cpp:1060-1093)
  if (t >= cf.sampling_start_index) KW_TRY(sample_streams(c));
  c->t++;
  return KW_OK;
}

// ---- the pipelined step of slab-decomposed runs -------------------------------------------------------------------
// Same arithmetic as step(), but every field travels on its own: x/y passes of field f, its exchange (on the communication
// stream), its z pass, the exchange back, its inverse y pass.  While field f is on the wire the solver stream works on
// the other fields, and the forward transform of u_f starts as soon as the velocity update of component f is done, i.e.
// while the gradient components f+1.. are still arriving.
static void ycol1(kw_ctx* c, float2* buf, int dir) {
  float2* b[1] = {buf};
  const ColArgs ca = ycol_args(c->g, b, 1);
  launch(c, dir < 0 ? "ycol_fwd" : "ycol_inv", 16.0 * c->g.nca, [&] { c->g.oy->col(ca, dir, 1, c->st); });
}
static void forward1(kw_ctx* c, const float* in, float2* out) {
  const float* i[1] = {in};
  float2* o[1] = {out};
  forward_xy(c, i, o, 1);
}

static int step_sharded(kw_ctx* c) {
  const kw_config& cf = c->cfg;
  const Geometry& g = c->g;
  const uint64_t t = c->t;
  const float fd = 1.0f / (float)g.ntot;
  float* u[3] = {c->d[KW_UX_SGX], c->d[KW_UY_SGY], c->d[KW_UZ_SGZ]};
  float* rho[3] = {c->d[KW_RHOX], c->d[KW_RHOY], c->d[KW_RHOZ]};
  const int neg[3] = {KW_DDX_K_SHIFT_NEG_R, KW_DDY_K_SHIFT_NEG, KW_DDZ_K_SHIFT_NEG};
  const uint64_t uflag[3] = {cf.ux_source_flag, cf.uy_source_flag, cf.uz_source_flag};
  const int npairs = g.nzl * g.ny / 2;
  cudaEvent_t ev, ef[3], eb[3];

  // computeVelocity (cpp:2087-2119) [+ velocity sources + forward transforms of the new velocity]
  auto velocity_phase = [&](int init, bool forward_u) -> int {
    forward1(c, c->d[KW_P], c->S[3]);
    KW_TRY(exchange_async(c, 3, 7, &ev));
    cudaStreamWaitEvent(c->st, ev, 0);
    ZField zf{c->R[3], c->S[0], c->d[KW_KAPPA], 1.0f, reinterpret_cast<const float2*>(c->d[KW_DDX_K_SHIFT_POS_R]),
              c->S[1], c->S[2], reinterpret_cast<const float2*>(c->d[KW_DDY_K_SHIFT_POS]),
              reinterpret_cast<const float2*>(c->d[KW_DDZ_K_SHIFT_POS])};
    zmid_launch(c, zf, 3);
    KW_TRY(release_buffer(c, 7, c->st));
    cudaEvent_t eg[3];
    for (int f = 0; f < 3; ++f) KW_TRY(exchange_async(c, f, 4 + f, &eg[f]));
    const EpiVelocity e = velocity_epilogue(c, u, fd, init);
    const double het = c->count[KW_RHO0_SGX] > 1 ? 4.0 : 0.0;
    for (int f = 0; f < 3; ++f) {
      cudaStreamWaitEvent(c->st, eg[f], 0);
      ycol1(c, c->R[f], +1);
      XInvArgs<1> xa = xinv_args<1>(c, c->R, 0, npairs, 3);
      xa.field0 = f;
      launch(c, init ? "xinv_initial_velocity" : "xinv_velocity", 8.0 * g.nca + (init ? 8.0 : 8.0 + het) * g.n,
             [&] { g.ox->xinv_velocity(xa, e, 1, c->st); });
      KW_TRY(release_buffer(c, 4 + f, c->st));
      if (!forward_u) continue;
      // addVelocitySource (cpp:2252-2303), transducer (cpp:894-897) of this component
      if (uflag[f] > t) {
        const size_t nsrc = c->count[KW_U_SOURCE_INDEX];
        if (cf.u_source_mode != KW_SRC_ADDITIVE) {
          if (nsrc > 0) {
            SourceArgs sa{};
            sa.target[0] = u[f], sa.ntargets = 1, sa.signal = c->d[KW_UX_SOURCE_INPUT + f], sa.index = c->di[KW_U_SOURCE_INDEX];
            sa.pos = c->dpos[KW_U_SOURCE_INDEX], sa.nsrc = nsrc, sa.nsrc_total = c->count_total[KW_U_SOURCE_INDEX];
            sa.t = t, sa.many = cf.u_source_many, sa.mode = cf.u_source_mode;
            launch(c, "add_u_source", 20.0 * nsrc, [&] { k_add_source<<<ew_grid(nsrc), 256, 0, c->st>>>(sa); });
          }
        } else {
          float* tg[1] = {u[f]};
          KW_TRY(add_scaled_source(c, c->d[KW_UX_SOURCE_INPUT + f], KW_U_SOURCE_INDEX, cf.u_source_many, tg, 1));
        }
      }
      if (f == 0 && cf.transducer_source_flag > t && c->count[KW_U_SOURCE_INDEX] > 0) {
        const size_t nsrc = c->count[KW_U_SOURCE_INDEX];
        launch(c, "add_transducer", 28.0 * nsrc, [&] {
          k_add_transducer<<<ew_grid(nsrc), 256, 0, c->st>>>(u[0], c->di[KW_U_SOURCE_INDEX], c->d[KW_TRANSDUCER_SOURCE_INPUT],
                                                              c->di[KW_DELAY_MASK], nsrc, t);
        });
      }
      forward1(c, u[f], c->S[f]);  // computeVelocityGradient starts for this component
      KW_TRY(exchange_async(c, f, 4 + f, &ef[f]));
    }
    return KW_OK;
  };
  KW_TRY(velocity_phase(0, true));

  // ---- computeVelocityGradient (cpp:2126-2150) + computeDensity (cpp:2157/2169) [+ pressure terms / lossless p]
  for (int f = 0; f < 3; ++f) {
    cudaStreamWaitEvent(c->st, ef[f], 0);
    zmid_launch(c, ZField{c->R[f], c->R[f], c->d[KW_KAPPA], fd, reinterpret_cast<const float2*>(c->d[neg[f]])}, f);
    KW_TRY(exchange_async(c, 4 + f, f, &eb[f]));
    KW_TRY(release_buffer(c, 4 + f, c->cs));
  }
  for (int f = 0; f < 3; ++f) {
    cudaStreamWaitEvent(c->st, eb[f], 0);
    ycol1(c, c->S[f], +1);
  }
  const bool p_src = cf.p_source_flag > t;
  {
    EpiDensity e{};
    for (int k = 0; k < 3; ++k) e.rho[k] = rho[k], e.pml[k] = c->d[KW_PML_X + k];
    e.pml[2] += g.z0;
    e.rho0 = c->fld(KW_RHO0), e.bona = c->fld(KW_BONA), e.c2 = c->fld(KW_C0);
    e.dt = cf.dt, e.nonlinear = cf.nonlinear_flag, e.absorbing = cf.absorbing_flag;
    e.defer_terms = p_src && cf.p_source_mode == KW_SRC_ADDITIVE;
    e.outA = c->tA, e.outB = c->tB, e.outNL = c->tNL, e.p = c->d[KW_P];
    double fused_bytes = 0;
    if (!cf.absorbing_flag && !p_src) e.sample = fused_p_sample(c, &e.fs, &fused_bytes);
    double per = 24.0 + (c->count[KW_RHO0] > 1 ? 4.0 : 0.0);
    if (cf.absorbing_flag) per += 4.0 + (e.defer_terms ? 0.0 : 4.0 + (cf.nonlinear_flag ? 4.0 : 0.0));
    else if (!e.defer_terms) per += 4.0 + (c->count[KW_C0] > 1 ? 4.0 : 0.0);
    if (cf.nonlinear_flag && !e.defer_terms && c->count[KW_BONA] > 1) per += 4.0;
    launch(c, "xinv_density", 24.0 * g.nca + per * g.n + fused_bytes, [&] { g.ox->xinv_density(xinv_args<3>(c, c->S, 0, npairs), e, c->st); });
    for (int f = 0; f < 3; ++f) KW_TRY(release_buffer(c, f, c->st));
  }
  // ---- addPressureSource (cpp:2310-2334)
  if (p_src) {
    const size_t nsrc = c->count[KW_P_SOURCE_INDEX];
    TermsArgs ta = terms_args(c);
    if (cf.p_source_mode != KW_SRC_ADDITIVE) {
      if (nsrc > 0) {
        SourceArgs sa{};
        for (int k = 0; k < 3; ++k) sa.target[k] = rho[k];
        sa.ntargets = g.nz == 1 ? 2 : 3 /* 2-D: rhox, rhoy only (SolverCudaKernels.cu:570-629) */, sa.signal = c->d[KW_P_SOURCE_INPUT], sa.index = c->di[KW_P_SOURCE_INDEX];
        sa.pos = c->dpos[KW_P_SOURCE_INDEX], sa.nsrc = nsrc, sa.nsrc_total = c->count_total[KW_P_SOURCE_INDEX];
        sa.t = t, sa.many = cf.p_source_many, sa.mode = cf.p_source_mode;
        launch(c, "add_p_source", 36.0 * nsrc, [&] { k_add_source<<<ew_grid(nsrc), 256, 0, c->st>>>(sa); });
        ta.index = c->di[KW_P_SOURCE_INDEX], ta.n = nsrc;
        launch(c, "pressure_terms_fixup", 40.0 * nsrc, [&] { k_pressure_terms<<<ew_grid(nsrc), 256, 0, c->st>>>(ta); });
      }
    } else {
      KW_TRY(add_scaled_source(c, c->d[KW_P_SOURCE_INPUT], KW_P_SOURCE_INDEX, cf.p_source_many, rho, g.nz == 1 ? 2 : 3));
      ta.index = nullptr, ta.n = g.n;
      launch(c, "pressure_terms", 28.0 * g.n, [&] { k_pressure_terms<<<ew_grid(g.n), 256, 0, c->st>>>(ta); });
    }
  }
  // ---- computePressure, absorbing branch (cpp:2180-2246)
  if (cf.absorbing_flag) {
    const float* in[2] = {c->tA, c->tB};
    const int nab[2] = {KW_ABSORB_NABLA1, KW_ABSORB_NABLA2};
    for (int f = 0; f < 2; ++f) {
      forward1(c, in[f], c->S[f]);
      KW_TRY(exchange_async(c, f, 4 + f, &ef[f]));
    }
    for (int f = 0; f < 2; ++f) {
      cudaStreamWaitEvent(c->st, ef[f], 0);
      zmid_launch(c, ZField{c->R[f], c->R[f], c->d[nab[f]], 1.0f, nullptr}, -1);
      KW_TRY(exchange_async(c, 4 + f, f, &eb[f]));
      KW_TRY(release_buffer(c, 4 + f, c->cs));
    }
    for (int f = 0; f < 2; ++f) {
      cudaStreamWaitEvent(c->st, eb[f], 0);
      ycol1(c, c->S[f], +1);
    }
    EpiPressureSum e{};
    e.p = c->d[KW_P], e.base = cf.nonlinear_flag ? c->tNL : c->tB;
    e.c2 = c->fld(KW_C0), e.tau = c->fld(KW_ABSORB_TAU), e.eta = c->fld(KW_ABSORB_ETA), e.fd = fd;
    const double per = 8.0 + (c->count[KW_C0] > 1 ? 4.0 : 0.0) + (c->count[KW_ABSORB_TAU] > 1 ? 8.0 : 0.0);
    double fused_bytes = 0;
    e.sample = fused_p_sample(c, &e.fs, &fused_bytes);
    launch(c, "xinv_pressure_sum", 16.0 * g.nca + per * g.n + fused_bytes, [&] { g.ox->xinv_psum(xinv_args<2>(c, c->S, 0, npairs), e, c->st); });
    for (int f = 0; f < 2; ++f) KW_TRY(release_buffer(c, f, c->st));
  }
  // ---- addInitialPressureSource (cpp:2359-2396)
  if (t == 0 && cf.p0_source_flag == 1) {
    launch(c, "initial_pressure", 24.0 * g.n, [&] {
      k_initial_pressure<<<ew_grid(g.n), 256, 0, c->st>>>(c->d[KW_P], rho[0], rho[1], rho[2], c->d[KW_P0_SOURCE_INPUT], c->fld(KW_C0), g.n, g.nz == 1 ? 2 : 3);
    });
    KW_TRY(velocity_phase(1, false));
  }
  // ---- storeSensorData (cpp:1060-1093)
  if (t >= cf.sampling_start_index) KW_TRY(sample_streams(c));
  c->t++;
  return KW_OK;
}

// Q_term_c = -(dIx/dx + dIy/dy + dIz/dz) of the time-averaged intensity (computeQTerm, KSpaceFirstOrderSolver.cpp:1783-2080):
// the I_avg_c values are scattered onto a zero grid, differentiated spectrally along their own axis (multiplier i*k,
// k = 2 pi / d * shift / N; the reference runs 1-D R2C / C2R per axis, here the multiplier sits between the z transforms of
// the 3-D pipeline like the other operators), summed in the order x, y, z, negated and gathered at the sensor points.
static int compute_q_term(kw_ctx* c, const float* const* intensity, float* q_out) {
  const Geometry& g = c->g;
  const kw_config& cf = c->cfg;
  Stream& q = c->streams[KW_S_P_RAW];  // only its "not whole-domain" flag matters to sample_one below
  // two full-grid scratch arrays and the i*k vectors live only for this call
  struct Scratch {
    std::vector<void*> p;
    ~Scratch() {
      for (void* q : p) cudaFree(q);
    }
  } scratch;
  auto salloc = [&](void** p, size_t bytes) -> int {
    KW_CUDA(cudaMalloc(p, bytes));
    scratch.p.push_back(*p);
    return KW_OK;
  };
  float *grid = nullptr, *acc = nullptr;
  KW_TRY(salloc((void**)&grid, g.n * sizeof(float)));
  KW_TRY(salloc((void**)&acc, g.n * sizeof(float)));
  KW_CUDA(cudaMemsetAsync(acc, 0, g.n * sizeof(float), c->st));
  const int ncomp = g.nz == 1 ? 2 : 3;
  const int len[3] = {g.nxp, g.ny, g.nz}, n[3] = {g.nx, g.ny, g.nz};
  const float d[3] = {cf.dx, cf.dy, cf.dz};
  const float fd = 1.0f / (float)g.ntot;
  for (int f = 0; f < ncomp; ++f) {
    // i*k of this axis: the half range for x, the full (Hermitian) range for y and z; the Nyquist entry multiplies a
    // value whose imaginary part C2R ignores, i.e. it contributes nothing (cpp:1903-1921)
    std::vector<float2> ik(len[f], make_float2(0.f, 0.f));
    const int count = f == 0 ? g.nxr : n[f];
    const float pi2 = static_cast<float>(M_PI) * 2.0f;
    for (int i = 0; i < count; ++i) {
      const long long shift = (long long)((i + n[f] / 2) % n[f]) - n[f] / 2;
      const bool nyquist = n[f] % 2 == 0 && i == n[f] / 2;
      ik[i] = make_float2(0.f, nyquist && f != 0 ? 0.f : (pi2 / d[f]) * ((float)shift / (float)n[f]));
    }
    float2* dik = nullptr;
    KW_TRY(salloc((void**)&dik, ik.size() * sizeof(float2)));
    KW_CUDA(cudaMemcpyAsync(dik, ik.data(), ik.size() * sizeof(float2), cudaMemcpyHostToDevice, c->st));
    KW_CUDA(cudaStreamSynchronize(c->st));
    const float* I = intensity[f];
    KW_CUDA(cudaMemsetAsync(grid, 0, g.n * sizeof(float), c->st));
    if (c->nsens) {
      if (cf.sensor_mask_type == 0) {
        k_scatter_index<<<ew_grid(c->nsens), 256, 0, c->st>>>(grid, I, c->di[KW_SENSOR_MASK_INDEX], c->nsens);
      } else {
        CuboidArgs ca{c->cub_corners, c->cub_offsets, c->ncuboids, g.nx, g.ny};
        k_scatter_cuboid<<<ew_grid(c->nsens), 256, 0, c->st>>>(grid, I, ca, c->nsens);
      }
      c->launches++;
    }
    const float* in[1] = {grid};
    float2* out[1] = {c->S[3]};
    forward_xy(c, in, out, 1);
    float2 *zb[1], *back[1];
    KW_TRY(exchange(c, out, &c->R[3], 1, zb));
    zmid_launch(c, ZField{zb[0], zb[0], nullptr, fd, dik}, f);
    KW_TRY(exchange(c, zb, &c->S[3], 1, back));
    if (g.nranks > 1) KW_TRY(release_buffer(c, 7, c->cs));
    EpiAdd e{};
    e.out[0] = acc, e.ntargets = 1;
    inverse_yx(c, back, 1, "xinv_q_term", 8.0 * g.nca + 8.0 * g.n,
               [&](int pb, int pe) { g.ox->xinv_add(xinv_args<1>(c, back, pb, pe), e, c->st); });
    if (g.nranks > 1) KW_TRY(release_buffer(c, 3, c->st));
  }
  if (c->nsens) {
    Stream gather = q;
    gather.all = false;
    sample_one<kOpNone>(c, gather, acc, q_out);
    k_negate<<<ew_grid(c->nsens), 256, 0, c->st>>>(q_out, c->nsens);
    c->launches++;
  }
  KW_CUDA(cudaStreamSynchronize(c->st));  // the scratch arrays are released on return
  if (c->cs) KW_CUDA(cudaStreamSynchronize(c->cs));
  return KW_OK;
}
static int compute_q_term_c(kw_ctx* c) {
  const float* I[3] = {c->streams[KW_S_IX_AVG_C].dbuf, c->streams[KW_S_IY_AVG_C].dbuf, c->streams[KW_S_IZ_AVG_C].dbuf};
  return compute_q_term(c, I, c->streams[KW_S_Q_TERM_C].dbuf);
}

// computeAverageIntensities (KSpaceFirstOrderSolver.cpp:1231-1534) for a block of n sensor points with `steps` stored samples
// each (series laid out [step][point], as in the output file): the velocity is shifted by half a time step -- the
// reference's R2C / * exp(i pi shift / steps) / C2R along time is a circular convolution with the real kernel h built
// below in double precision (any number of steps, no FFT length restriction) -- and I = sum_t p * u_shifted / steps.
// partial[c][i] = sum over the time chunk c of p * u_shifted: chunks are summed afterwards in a fixed order, so the result does
// not depend on the order in which blocks run (a restart must reproduce the uninterrupted run bit for bit).
// HS: the kernel h sits in shared memory (steps * 4 bytes fit), otherwise it is read through the read-only cache.
template <bool HS>
static __global__ void k_intensity_avg(const float* __restrict__ p, const float* __restrict__ u, const float* __restrict__ h, float* partial, size_t n,
                                       int steps, int tchunk) {
  extern __shared__ float sh[];
  if (HS) {
    for (int m = threadIdx.x; m < steps; m += blockDim.x) sh[m] = h[m];
    __syncthreads();
  }
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t0 = blockIdx.y * tchunk, t1 = min(steps, t0 + tchunk);
  float acc = 0.f;
  for (int t = t0; t < t1; ++t) {
    float us = 0.f;
    int m = t;  // (t - s) mod steps, walked downwards
    for (int s = 0; s < steps; ++s) {
      us = fmaf(__ldg(u + (size_t)s * n + i), HS ? sh[m] : __ldg(h + m), us);
      m = m == 0 ? steps - 1 : m - 1;
    }
    acc = fmaf(__ldg(p + (size_t)t * n + i), us, acc);
  }
  partial[(size_t)blockIdx.y * n + i] = acc;
}
// I[i] = (sum_c partial[c][i]) / steps, chunks in ascending order
static __global__ void k_intensity_reduce(const float* __restrict__ partial, float* I, size_t n, int nchunks, float steps) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < nchunks; ++c) acc += partial[(size_t)c * n + i];
    I[i] = acc / steps;
  }
}

}  // namespace kw

extern "C" {

int kw_run(kw_ctx* c, uint64_t nsteps, uint64_t* steps_done, int sync) {
  if (steps_done) *steps_done = 0;
  if (!c) return fail(KW_ERR_INVALID, "null context");
  if (!c->preprocessed) return fail(KW_ERR_STATE, "kw_run before kw_preprocess");
  uint64_t done = 0;
  int rc = KW_OK;
  KW_CUDA(cudaEventRecord(c->ev0, c->st));
  for (; done < nsteps && c->t < c->cfg.nt; ++done) {
    bool full = false;
    if (c->t >= c->cfg.sampling_start_index)
      for (auto& s : c->streams) {
        if (!(s.enabled && !s.nosave && (s.op == kOpNone || s.op == kOpC) && s.rows >= s.cap_rows)) continue;
        if (s.dalt && s.pending_rows == 0) {  // asynchronous output: the full buffer leaves on the output stream, sampling goes on in the other
          const cudaEvent_t filled = mark(c, c->st);
          cudaStreamWaitEvent(c->os, filled, 0);
          if (s.row) KW_CUDA(cudaMemcpyAsync(s.hbuf, s.dbuf, s.rows * s.row * sizeof(float), cudaMemcpyDeviceToHost, c->os));
          KW_CUDA(cudaEventRecord(s.landed, c->os));
          s.pending_rows = s.rows, s.rows = 0;
          std::swap(s.dbuf, s.dalt);
          continue;
        }
        full = true;
      }
    if (full) {
      rc = fail(KW_ERR_STREAM_FULL, "a raw stream buffer is full: fetch it with kw_stream_fetch");
      break;
    }
    rc = c->g.nranks > 1 ? step_sharded(c) : step(c);
    if (rc != KW_OK) break;  // steps_done and the end event are still recorded: the host's checkpoint bookkeeping relies on them
  }
  KW_CUDA(cudaEventRecord(c->ev1, c->st));
  if (steps_done) *steps_done = done;
  if (rc != KW_OK && rc != KW_ERR_STREAM_FULL) return rc;
  KW_CUDA(cudaGetLastError());
  if (sync) {
    KW_CUDA(cudaStreamSynchronize(c->st));
    if (c->cs) KW_CUDA(cudaStreamSynchronize(c->cs));
    if (c->ws) KW_CUDA(cudaStreamSynchronize(c->ws));
    KW_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    prof_resolve(c);
  }
  return rc;
}

int kw_time_index(kw_ctx* c, uint64_t* t) {
  if (!c || !t) return fail(KW_ERR_INVALID, "null argument");
  *t = c->t;
  return KW_OK;
}
// ---- checkpoint / restart support (KSpaceFirstOrderSolver::saveCheckpointData cpp:1176-1224, loadInputData :186-228;
//      BaseOutputStream::checkpoint / reopen, OutputStreams/BaseOutputStream.cpp:528-606) ------------------------------
int kw_set_time_index(kw_ctx* c, uint64_t t) {
  if (!c) return fail(KW_ERR_INVALID, "null context");
  if (!c->preprocessed) return fail(KW_ERR_STATE, "kw_set_time_index before kw_preprocess");
  if (t > c->cfg.nt) return fail(KW_ERR_INVALID, "time index beyond Nt");
  c->t = t;
  // the step counters of the streams follow from the time index, as in IndexOutputStream::reopen (IndexOutputStream.cpp:203-213):
  // sampled = t - start, compressed = floor(sampled / oSize); a blob restored afterwards (kw_stream_state_set) overrides them
  const uint64_t sampled = t > c->cfg.sampling_start_index ? t - c->cfg.sampling_start_index : 0;
  for (auto& s : c->streams) {
    if (!s.enabled) continue;
    s.rows = 0;
    if (s.op == kOpC || s.op == kOpIAvgC) {
      s.sampled = sampled;
      s.compressed = c->c_osize > 0 ? sampled / (uint64_t)c->c_osize : 0;
    }
  }
  return KW_OK;
}
// The buffers BaseOutputStream::checkpoint / reopen move through files (OutputStreams/BaseOutputStream.cpp:528-606, IndexOutputStream.cpp:177-247,
// :536-557): which = 0 the accumulator of an aggregate stream (rms / max / min [_all], I_avg_c: flushed into / reloaded from the OUTPUT
// file), which = 1 / 2 the two compression accumulators of a *_c stream (the reference's Temp_<name>_1 / _2 datasets of the checkpoint file).
static int stream_buffer(kw_ctx* c, int sid, int which, void** ptr, size_t* floats) {
  if (!c || sid < 0 || sid >= KW_STREAM_COUNT || !c->streams[sid].enabled) return fail(KW_ERR_INVALID, std::string("stream_buffer: stream ") + std::to_string(sid) + " is not enabled (or a null argument)");
  if (!c->preprocessed) return fail(KW_ERR_STATE, "stream buffers exist after kw_preprocess");
  Stream& s = c->streams[sid];
  if (which == 0 && s.op != kOpNone && s.op != kOpC && s.dbuf) *ptr = s.dbuf, *floats = s.row;
  else if ((which == 1 || which == 2) && s.op == kOpC) *ptr = which == 1 ? s.acc1 : s.acc2, *floats = s.acc_bytes / sizeof(float);
  else return fail(KW_ERR_INVALID, "kw_stream_buffer: this stream has no such buffer");
  return KW_OK;
}
int kw_stream_buffer_get(kw_ctx* c, int sid, int which, float* host, uint64_t capacity_floats, uint64_t* floats) {
  void* p = nullptr;
  size_t n = 0;
  KW_TRY(stream_buffer(c, sid, which, &p, &n));
  if (floats) *floats = n;
  if (!host) return KW_OK;
  if (capacity_floats < n) return fail(KW_ERR_INVALID, "kw_stream_buffer_get: buffer too small");
  KW_CUDA(cudaStreamSynchronize(c->st));
  KW_CUDA(cudaMemcpy(host, p, n * sizeof(float), cudaMemcpyDeviceToHost));
  return KW_OK;
}
int kw_stream_buffer_set(kw_ctx* c, int sid, int which, const float* host, uint64_t floats) {
  void* p = nullptr;
  size_t n = 0;
  KW_TRY(stream_buffer(c, sid, which, &p, &n));
  if (!host || floats != n) return fail(KW_ERR_INVALID, "kw_stream_buffer_set: wrong size");
  KW_CUDA(cudaStreamSynchronize(c->st));
  KW_CUDA(cudaMemcpy(p, host, n * sizeof(float), cudaMemcpyHostToDevice));
  return KW_OK;
}
namespace {
struct StreamStateHeader {
  uint64_t magic, op, sampled, compressed, payload_bytes, reserved;
};
constexpr uint64_t kStateMagic = 0x4b57535452454d31ull;  // "KWSTREM1"
size_t stream_payload_bytes(const kw_ctx* c, const Stream& s) {
  if (s.op == kOpC) return s.acc_bytes * (s.acc2 != s.acc1 ? 2 : 1);
  if (s.op == kOpNone) return 0;  // raw rows are flushed before a checkpoint
  return s.row * sizeof(float);   // rms / max / min (+ _all) accumulators, I_avg_c
}
}  // namespace
int kw_stream_state_size(kw_ctx* c, int sid, uint64_t* bytes) {
  if (!c || !bytes || sid < 0 || sid >= KW_STREAM_COUNT) return fail(KW_ERR_INVALID, "bad argument");
  if (!c->preprocessed) return fail(KW_ERR_STATE, "stream state before kw_preprocess");
  const Stream& s = c->streams[sid];
  *bytes = s.enabled ? sizeof(StreamStateHeader) + stream_payload_bytes(c, s) : 0;  // 0: stream does not exist in this run
  return KW_OK;
}
int kw_stream_state_get(kw_ctx* c, int sid, void* buf, uint64_t bytes) {
  uint64_t need = 0;
  KW_TRY(kw_stream_state_size(c, sid, &need));
  if (!buf || need == 0 || bytes < need) return fail(KW_ERR_INVALID, "kw_stream_state_get: stream not enabled or buffer too small");
  Stream& s = c->streams[sid];
  if (s.rows || s.pending_rows) return fail(KW_ERR_STATE, "kw_stream_state_get: fetch the buffered rows first");
  KW_CUDA(cudaStreamSynchronize(c->st));
  StreamStateHeader h{kStateMagic, (uint64_t)s.op, s.sampled, s.compressed, need - sizeof(StreamStateHeader), 0};
  memcpy(buf, &h, sizeof h);
  char* out = static_cast<char*>(buf) + sizeof h;
  if (s.op == kOpC) {
    KW_CUDA(cudaMemcpy(out, s.acc1, s.acc_bytes, cudaMemcpyDeviceToHost));
    if (s.acc2 != s.acc1) KW_CUDA(cudaMemcpy(out + s.acc_bytes, s.acc2, s.acc_bytes, cudaMemcpyDeviceToHost));
  } else if (h.payload_bytes) {
    KW_CUDA(cudaMemcpy(out, s.dbuf, h.payload_bytes, cudaMemcpyDeviceToHost));
  }
  return KW_OK;
}
int kw_stream_state_set(kw_ctx* c, int sid, const void* buf, uint64_t bytes) {
  uint64_t need = 0;
  KW_TRY(kw_stream_state_size(c, sid, &need));
  if (!buf || need == 0 || bytes < need) return fail(KW_ERR_INVALID, "kw_stream_state_set: stream not enabled or state too short");
  Stream& s = c->streams[sid];
  StreamStateHeader h;
  memcpy(&h, buf, sizeof h);
  if (h.magic != kStateMagic || h.op != (uint64_t)s.op || h.payload_bytes != need - sizeof h)
    return fail(KW_ERR_INVALID, "kw_stream_state_set: the state was saved by a different stream configuration");
  s.sampled = h.sampled, s.compressed = h.compressed, s.rows = 0;
  const char* in = static_cast<const char*>(buf) + sizeof h;
  if (s.op == kOpC) {
    KW_CUDA(cudaMemcpy(s.acc1, in, s.acc_bytes, cudaMemcpyHostToDevice));
    if (s.acc2 != s.acc1) KW_CUDA(cudaMemcpy(s.acc2, in + s.acc_bytes, s.acc_bytes, cudaMemcpyHostToDevice));
  } else if (h.payload_bytes) {
    KW_CUDA(cudaMemcpy(s.dbuf, in, h.payload_bytes, cudaMemcpyHostToDevice));
  }
  return KW_OK;
}

int kw_synchronize(kw_ctx* c) {
  if (!c) return fail(KW_ERR_INVALID, "null context");
  KW_CUDA(cudaStreamSynchronize(c->st));
  if (c->cs) KW_CUDA(cudaStreamSynchronize(c->cs));
  if (c->ws) KW_CUDA(cudaStreamSynchronize(c->ws));
  cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1);
  prof_resolve(c);
  KW_CUDA(cudaGetLastError());
  return KW_OK;
}
int kw_profile(kw_ctx* c, int enable, int reset) {
  if (!c) return fail(KW_ERR_INVALID, "null context");
  KW_CUDA(cudaStreamSynchronize(c->st));
  if (c->cs) KW_CUDA(cudaStreamSynchronize(c->cs));
  if (c->ws) KW_CUDA(cudaStreamSynchronize(c->ws));
  prof_resolve(c);
  c->prof_on = enable != 0;
  if (reset) c->prof.clear();
  return KW_OK;
}
int kw_profile_report(kw_ctx* c, char* buf, uint64_t cap) {
  if (!c || !buf || cap == 0) return fail(KW_ERR_INVALID, "null argument");
  std::string out = "{";
  bool first = true;
  for (auto& kv : c->prof) {
    char line[256];
    snprintf(line, sizeof line, "%s\"%s\": {\"launches\": %llu, \"ms\": %.6f, \"bytes\": %.1f}", first ? "" : ", ", kv.first.c_str(),
             (unsigned long long)kv.second.launches, kv.second.ms, kv.second.bytes);
    out += line;
    first = false;
  }
  out += "}";
  if (out.size() + 1 > cap) return fail(KW_ERR_INVALID, "kw_profile_report: buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return KW_OK;
}
int kw_last_run_ms(kw_ctx* c, float* ms) {
  if (!c || !ms) return fail(KW_ERR_INVALID, "null argument");
  *ms = c->last_ms;
  return KW_OK;
}
int kw_launch_count(kw_ctx* c, uint64_t* n) {
  if (!c || !n) return fail(KW_ERR_INVALID, "null argument");
  *n = c->launches;
  return KW_OK;
}

int kw_get_array(kw_ctx* c, int id, void* host, uint64_t count) {
  if (!c || !host || id < 0 || id >= KW_ARRAY_COUNT) return fail(KW_ERR_INVALID, "kw_get_array: bad argument");
  const Geometry& g = c->g;
  KW_CUDA(cudaStreamSynchronize(c->st));
  if (is_reduced_real(id)) {
    if (!c->d[id]) return fail(KW_ERR_INVALID, "array not present");
    const size_t rows = (size_t)g.nyl * g.nz;  // z-local side: [kz][ky_local][kx]
    if (count < rows * g.nxr) return fail(KW_ERR_INVALID, "kw_get_array: buffer too small");
    KW_CUDA(cudaMemcpy2D(host, g.nxr * sizeof(float), c->d[id], g.nxp * sizeof(float), g.nxr * sizeof(float), rows, cudaMemcpyDeviceToHost));
    return KW_OK;
  }
  if (c->count[id] == 1 && !c->d[id]) {
    static_cast<float*>(host)[0] = c->scalar[id];
    return KW_OK;
  }
  if (!c->d[id] || is_index_array(id)) return fail(KW_ERR_INVALID, "array not present on the device");
  size_t n = c->count[id];
  if (is_complex_vec(id)) n = std::min<size_t>(n, 2 * count);
  else if (count < n) return fail(KW_ERR_INVALID, "kw_get_array: buffer too small");
  KW_CUDA(cudaMemcpy(host, c->d[id], n * sizeof(float), cudaMemcpyDeviceToHost));
  return KW_OK;
}

int kw_nccl_unique_id(void* out, uint64_t capacity) {
  if (!out || capacity < sizeof(KwNcclUniqueId)) return fail(KW_ERR_INVALID, "kw_nccl_unique_id: need a 128-byte buffer");
  NcclApi& nc = nccl_api();
  if (!nc.ok) return fail(KW_ERR_COMM, "NCCL unavailable: " + nc.error);
  KwNcclUniqueId id;
  const int e = nc.GetUniqueId(&id);
  if (e != kNcclSuccess) return fail(KW_ERR_COMM, std::string("ncclGetUniqueId: ") + nc.GetErrorString(e));
  memcpy(out, &id, sizeof(id));
  return KW_OK;
}
int kw_local_slab(kw_ctx* c, uint64_t* z_begin, uint64_t* z_count) {
  if (!c) return fail(KW_ERR_INVALID, "null context");
  if (z_begin) *z_begin = (uint64_t)c->g.z0;
  if (z_count) *z_count = (uint64_t)c->g.nzl;
  return KW_OK;
}
int kw_sensor_layout(kw_ctx* c, uint64_t* total, uint64_t* local, uint64_t* positions, uint64_t capacity) {
  if (!c || !c->preprocessed) return fail(KW_ERR_STATE, "kw_sensor_layout before kw_preprocess");
  if (total) *total = c->nsens_total;
  if (local) *local = c->nsens;
  if (!positions) return KW_OK;
  if (capacity < c->nsens) return fail(KW_ERR_INVALID, "kw_sensor_layout: buffer too small");
  if (c->cfg.sensor_mask_type == 0) {
    if (c->g.nranks == 1) for (size_t j = 0; j < c->nsens; ++j) positions[j] = j;
    else memcpy(positions, c->sens_pos.data(), c->nsens * sizeof(uint64_t));
  } else {
    size_t k = 0;
    for (size_t r = 0; r + 1 < c->sens_ranges.size(); r += 2)
      for (uint64_t j = 0; j < c->sens_ranges[r + 1]; ++j) positions[k++] = c->sens_ranges[r] + j;
  }
  return KW_OK;
}
// ---- compression helpers exposed for hosts and tests (CompressHelper's public statics) -----------------------------
__global__ void k_c40_encode(const float2* in, uint8_t* out, size_t n, int e) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) c40_encode(in[i], out + 5 * i, e);
}
__global__ void k_c40_decode(const uint8_t* in, float2* out, size_t n, int e) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = c40_decode(in + 5 * i, e);
}
static int c40_host(const void* in, void* out, uint64_t n, int e, bool encode) {
  if (!in || !out || n == 0) return fail(KW_ERR_INVALID, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(KW_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  void *din = nullptr, *dout = nullptr;
  const size_t bin = encode ? n * sizeof(float2) : n * 5, bout = encode ? n * 5 : n * sizeof(float2);
  KW_CUDA(cudaMalloc(&din, bin));
  KW_CUDA(cudaMalloc(&dout, bout));
  KW_CUDA(cudaMemcpy(din, in, bin, cudaMemcpyHostToDevice));
  if (encode) k_c40_encode<<<ew_grid(n), 256>>>(static_cast<const float2*>(din), static_cast<uint8_t*>(dout), n, e);
  else k_c40_decode<<<ew_grid(n), 256>>>(static_cast<const uint8_t*>(din), static_cast<float2*>(dout), n, e);
  KW_CUDA(cudaGetLastError());
  KW_CUDA(cudaMemcpy(out, dout, bout, cudaMemcpyDeviceToHost));
  cudaFree(din), cudaFree(dout);
  return KW_OK;
}
int kw_c40_encode(const float* complex_pairs, uint64_t n, int max_exp, uint8_t* bytes) { return c40_host(complex_pairs, bytes, n, max_exp, true); }
int kw_c40_decode(const uint8_t* bytes, uint64_t n, int max_exp, float* complex_pairs) { return c40_host(bytes, complex_pairs, n, max_exp, false); }
int kw_compression_bases(kw_ctx* c, int shifted, float* be, float* be1, uint64_t capacity_complex, uint64_t* osize, uint64_t* bsize) {
  if (!c || !c->preprocessed || !c->d_be[0]) return fail(KW_ERR_STATE, "no compressed stream is enabled");
  if (osize) *osize = (uint64_t)c->c_osize;
  if (bsize) *bsize = (uint64_t)c->c_bsize;
  const size_t n = (size_t)c->c_H * c->c_bsize;
  if (!be && !be1) return KW_OK;
  if (capacity_complex < n) return fail(KW_ERR_INVALID, "kw_compression_bases: buffer too small");
  KW_CUDA(cudaStreamSynchronize(c->st));
  if (be) KW_CUDA(cudaMemcpy(be, c->d_be[shifted ? 1 : 0], n * sizeof(float2), cudaMemcpyDeviceToHost));
  if (be1) KW_CUDA(cudaMemcpy(be1, c->d_be1[shifted ? 1 : 0], n * sizeof(float2), cudaMemcpyDeviceToHost));
  return KW_OK;
}

// ---- post-processing of stored raw series (--I_avg, --Q_term): cpp:1231-1534, :1783-2080 ----------------------------------
int kw_intensity_avg_block(const float* p, const float* const* u, int ncomp, uint64_t n, uint64_t steps, float* const* intensity) {
  if (!p || !u || !intensity || ncomp < 1 || ncomp > 3 || n == 0 || steps == 0) return fail(KW_ERR_INVALID, "kw_intensity_avg_block: bad argument");
  if (steps > 0x7fffffffull) return fail(KW_ERR_INVALID, "kw_intensity_avg_block: too many stored steps");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(KW_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  std::vector<float> h(steps);
  {  // h[m] = (1 + 2 sum_{k=1}^{(steps-1)/2} cos(2 pi k m / steps + pi k / steps)) / steps; the Nyquist bin contributes nothing
    const long long K = ((long long)steps - 1) / 2;
#pragma omp parallel for schedule(static)
    for (long long m = 0; m < (long long)steps; ++m) {
      double acc = 1.0;
      for (long long k = 1; k <= K; ++k) acc += 2.0 * cos((2.0 * M_PI * (double)k * (double)m + M_PI * (double)k) / (double)steps);
      h[m] = (float)(acc / (double)steps);
    }
  }
  struct Scratch {  // released on every return path
    std::vector<void*> p;
    ~Scratch() {
      for (void* q : p) cudaFree(q);
    }
  } sc;
  auto alloc = [&](float** q, size_t bytes) -> int {
    KW_CUDA(cudaMalloc(q, bytes));
    sc.p.push_back(*q);
    return KW_OK;
  };
  // time chunks: enough of them to fill the device when the block has few points, never more than one per 32 steps
  const int max_chunks = (int)((steps + 31) / 32);
  int nchunks = (int)std::min<uint64_t>((uint64_t)max_chunks, std::max<uint64_t>(1, ((uint64_t)sm_count() * 8 * 128 + n - 1) / n));
  const int tchunk = (int)((steps + nchunks - 1) / nchunks);
  nchunks = (int)((steps + tchunk - 1) / tchunk);
  float *dp = nullptr, *du = nullptr, *dh = nullptr, *dI = nullptr, *dpart = nullptr;
  const size_t bytes = n * steps * sizeof(float);
  KW_TRY(alloc(&dp, bytes));
  KW_TRY(alloc(&du, bytes));
  KW_TRY(alloc(&dh, steps * sizeof(float)));
  KW_TRY(alloc(&dI, n * sizeof(float)));
  KW_TRY(alloc(&dpart, (size_t)nchunks * n * sizeof(float)));
  KW_CUDA(cudaMemcpy(dp, p, bytes, cudaMemcpyHostToDevice));
  KW_CUDA(cudaMemcpy(dh, h.data(), steps * sizeof(float), cudaMemcpyHostToDevice));
  const dim3 grid((unsigned)((n + 127) / 128), (unsigned)nchunks);
  const bool hs = steps * sizeof(float) <= 200 * 1024;  // the kernel h in shared memory when it fits
  if (hs) KW_CUDA(cudaFuncSetAttribute(k_intensity_avg<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(steps * sizeof(float))));
  for (int f = 0; f < ncomp; ++f) {
    KW_CUDA(cudaMemcpy(du, u[f], bytes, cudaMemcpyHostToDevice));
    if (hs) k_intensity_avg<true><<<grid, 128, steps * sizeof(float)>>>(dp, du, dh, dpart, n, (int)steps, tchunk);
    else k_intensity_avg<false><<<grid, 128>>>(dp, du, dh, dpart, n, (int)steps, tchunk);
    KW_CUDA(cudaGetLastError());
    k_intensity_reduce<<<ew_grid(n), 256>>>(dpart, dI, n, nchunks, (float)steps);
    KW_CUDA(cudaMemcpy(intensity[f], dI, n * sizeof(float), cudaMemcpyDeviceToHost));
  }
  return KW_OK;
}
int kw_q_term(kw_ctx* c, const float* const* intensity, int ncomp, float* q_out, uint64_t capacity) {
  if (!c || !intensity || !q_out) return fail(KW_ERR_INVALID, "null argument");
  if (!c->preprocessed) return fail(KW_ERR_STATE, "kw_q_term before kw_preprocess");
  if (ncomp != (c->g.nz == 1 ? 2 : 3)) return fail(KW_ERR_INVALID, "kw_q_term: one intensity per dimension");
  if (capacity < c->nsens) return fail(KW_ERR_INVALID, "kw_q_term: buffer too small");
  if (c->nsens == 0 && c->g.nranks == 1) return KW_OK;  // (slab-decomposed: collective -- a rank without sensor points still transforms its slab)
  float* dbuf = nullptr;  // [ncomp intensities | q]
  KW_CUDA(cudaMalloc(&dbuf, std::max<size_t>((size_t)(ncomp + 1) * c->nsens, 1) * sizeof(float)));
  float* dI[3] = {};
  for (int f = 0; f < ncomp; ++f) {
    dI[f] = dbuf + (size_t)f * c->nsens;
    if (c->nsens) cudaMemcpyAsync(dI[f], intensity[f], c->nsens * sizeof(float), cudaMemcpyHostToDevice, c->st);
  }
  float* dq = dbuf + (size_t)ncomp * c->nsens;
  int rc = compute_q_term(c, dI, dq);
  if (rc == KW_OK && c->nsens && cudaMemcpyAsync(q_out, dq, c->nsens * sizeof(float), cudaMemcpyDeviceToHost, c->st) != cudaSuccess) rc = fail(KW_ERR_CUDA, "kw_q_term: copy failed");
  cudaStreamSynchronize(c->st);
  cudaFree(dbuf);
  return rc;
}

int kw_comm_mode(kw_ctx* c, int* mode) {
  if (!c || !mode) return fail(KW_ERR_INVALID, "null argument");
  *mode = c->g.nranks == 1 ? 0 : (c->peer.active ? 2 : 1);
  return KW_OK;
}
int kw_comm_bytes(kw_ctx* c, double* bytes_sent) {
  if (!c || !bytes_sent) return fail(KW_ERR_INVALID, "null argument");
  *bytes_sent = c->comm_bytes;
  return KW_OK;
}

int kw_stream_info(kw_ctx* c, int sid, uint64_t* row_floats, uint64_t* rows) {
  if (!c || sid < 0 || sid >= KW_STREAM_COUNT || !c->streams[sid].enabled) return fail(KW_ERR_INVALID, std::string("kw_stream_info: stream ") + std::to_string(sid) + " is not enabled (or a null argument)");
  if (row_floats) *row_floats = c->streams[sid].row;
  const Stream& si = c->streams[sid];
  if (rows) *rows = (si.op == kOpNone || si.op == kOpC) ? (si.pending_rows ? si.pending_rows : si.rows) : 1;
  return KW_OK;
}

int kw_stream_fetch(kw_ctx* c, int sid, float* host, uint64_t cap, uint64_t* rows_fetched) {
  if (rows_fetched) *rows_fetched = 0;
  if (!c || !host || sid < 0 || sid >= KW_STREAM_COUNT || !c->streams[sid].enabled) return fail(KW_ERR_INVALID, std::string("kw_stream_fetch: stream ") + std::to_string(sid) + " is not enabled (or a null argument)");
  Stream& s = c->streams[sid];
  if (s.nosave) return fail(KW_ERR_INVALID, "stream exists only as an input of I_avg_c (not stored)");
  const bool series = s.op == kOpNone || s.op == kOpC;
  if (series && s.pending_rows) {  // asynchronous output: the older chunk, already on its way to (or in) pinned host memory
    const uint64_t n = s.pending_rows;
    if (cap < n * s.row) return fail(KW_ERR_INVALID, "kw_stream_fetch: host buffer too small");
    KW_CUDA(cudaEventSynchronize(s.landed));
    memcpy(host, s.hbuf, n * s.row * sizeof(float));
    s.pending_rows = 0;
    if (rows_fetched) *rows_fetched = n;
    return KW_OK;
  }
  const uint64_t rows = series ? s.rows : 1;
  if (cap < rows * s.row) return fail(KW_ERR_INVALID, "kw_stream_fetch: host buffer too small");
  if (rows) {
    KW_CUDA(cudaMemcpyAsync(host, s.dbuf, rows * s.row * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    KW_CUDA(cudaStreamSynchronize(c->st));
  }
  if (series) s.rows = 0;
  if (rows_fetched) *rows_fetched = rows;
  return KW_OK;
}

int kw_stream_pending(kw_ctx* c, int sid, uint64_t* rows) {
  if (!c || !rows || sid < 0 || sid >= KW_STREAM_COUNT || !c->streams[sid].enabled) return fail(KW_ERR_INVALID, std::string("kw_stream_pending: stream ") + std::to_string(sid) + " is not enabled (or a null argument)");
  *rows = c->streams[sid].pending_rows;
  return KW_OK;
}

int kw_stream_peek(kw_ctx* c, int sid, uint64_t offset, float* host, uint64_t count) {
  if (!c || !host || sid < 0 || sid >= KW_STREAM_COUNT || !c->streams[sid].enabled) return fail(KW_ERR_INVALID, std::string("kw_stream_peek: stream ") + std::to_string(sid) + " is not enabled (or a null argument)");
  Stream& s = c->streams[sid];
  if (s.op == kOpNone || s.op == kOpC || !s.dbuf) return fail(KW_ERR_INVALID, "kw_stream_peek reads aggregate streams (use kw_stream_fetch for series)");
  if (offset + count > s.row) return fail(KW_ERR_INVALID, "kw_stream_peek: range outside the accumulator");
  KW_CUDA(cudaMemcpyAsync(host, s.dbuf + offset, count * sizeof(float), cudaMemcpyDeviceToHost, c->st));
  KW_CUDA(cudaStreamSynchronize(c->st));
  return KW_OK;
}

int kw_finish(kw_ctx* c) {
  if (!c) return fail(KW_ERR_INVALID, "null context");
  if (c->finished) return KW_OK;
  const uint64_t nsamp = c->cfg.nt - c->cfg.sampling_start_index;
  for (auto& s : c->streams)
    if (s.enabled && s.op == kOpRms) {  // BaseOutputStream.cpp:172-178
      k_post_rms<<<ew_grid(s.row), 256, 0, c->st>>>(s.dbuf, 1.0f / (float)nsamp, s.row);
      c->launches++;
    } else if (s.enabled && s.op == kOpIAvgC && s.compressed > 0) {  // IndexOutputStream::postProcess :477-490
      k_divide<<<ew_grid(s.row), 256, 0, c->st>>>(s.dbuf, (float)s.compressed, s.row);
      c->launches++;
      s.compressed = 0;
    }
  if (c->streams[KW_S_Q_TERM_C].enabled) KW_TRY(compute_q_term_c(c));
  KW_CUDA(cudaStreamSynchronize(c->st));
  if (c->cs) KW_CUDA(cudaStreamSynchronize(c->cs));
  KW_CUDA(cudaGetLastError());
  c->finished = true;
  return KW_OK;
}

// ---- stand-alone 3-D transforms on host buffers (cuFFT layout) ----------------------------------------------------
static int fft3d_host(uint64_t nx, uint64_t ny, uint64_t nz, const float* in, float* out, bool forward) {
  if (!in || !out) return fail(KW_ERR_INVALID, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(KW_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  Geometry g;
  KW_TRY(g.init(nx, ny, nz));
  const size_t rows = (size_t)g.ny * g.nz;
  float* dreal = nullptr;
  float2 *dspec = nullptr, *dnat = nullptr;
  KW_CUDA(cudaMalloc(&dreal, g.n * sizeof(float)));
  KW_CUDA(cudaMalloc(&dspec, g.nc * sizeof(float2)));
  KW_CUDA(cudaMalloc(&dnat, rows * g.nxr * sizeof(float2)));
  KW_CUDA(cudaMemset(dspec, 0, g.nc * sizeof(float2)));
  ColArgs cy{}, cz{};
  cy.data[0] = cz.data[0] = dspec;
  cy.stride = g.nxp, cy.outer_stride = (size_t)g.ny * g.nxp, cy.ngroups = g.nxp / g.oy->col_w, cy.tile_begin = 0, cy.tile_end = g.nz * cy.ngroups, cy.n = g.ny, cy.nvalid = g.nxr;
  if (g.nz > 1) cz.stride = (size_t)g.ny * g.nxp, cz.outer_stride = g.nxp, cz.ngroups = g.nxp / g.oz->col_w, cz.tile_begin = 0, cz.tile_end = g.ny * cz.ngroups, cz.n = g.nz, cz.nvalid = g.nxr;
  if (forward) {
    KW_CUDA(cudaMemcpy(dreal, in, g.n * sizeof(float), cudaMemcpyHostToDevice));
    XFwdArgs xa{};
    xa.in[0] = dreal, xa.out[0] = dspec, xa.tab = g.tx, xa.pair_begin = 0, xa.pair_end = (int)(rows / 2), xa.nxp = g.nxp, xa.map = g.row_map(), xa.n = g.nx;
    g.ox->xfwd(xa, 1, 0);
    g.oy->col(cy, -1, 1, 0);
    if (g.nz > 1) g.oz->col(cz, -1, 1, 0);
    k_pad_complex<<<ew_grid(g.nc), 256>>>(dnat, dspec, g.nxr, g.nxp, rows, 0);
    KW_CUDA(cudaGetLastError());
    KW_CUDA(cudaMemcpy(out, dnat, rows * g.nxr * sizeof(float2), cudaMemcpyDeviceToHost));
  } else {
    KW_CUDA(cudaMemcpy(dnat, in, rows * g.nxr * sizeof(float2), cudaMemcpyHostToDevice));
    k_pad_complex<<<ew_grid(g.nc), 256>>>(dspec, dnat, g.nxr, g.nxp, rows, 1);
    if (g.nz > 1) g.oz->col(cz, +1, 1, 0);
    g.oy->col(cy, +1, 1, 0);
    XInvArgs<1> xa{};
    xa.in[0] = dspec, xa.tab = g.tx, xa.pair_begin = 0, xa.pair_end = (int)(rows / 2), xa.nxp = g.nxp, xa.ny = g.ny, xa.map = g.row_map(), xa.n = g.nx;
    EpiStore e{};
    e.out[0] = dreal, e.scale = 1.0f;
    g.ox->xinv_store(xa, e, 1, 0);
    KW_CUDA(cudaGetLastError());
    KW_CUDA(cudaMemcpy(out, dreal, g.n * sizeof(float), cudaMemcpyDeviceToHost));
  }
  KW_CUDA(cudaDeviceSynchronize());
  cudaFree(dreal), cudaFree(dspec), cudaFree(dnat);
  return KW_OK;
}
// micro-benchmark of one column pass (axis 1: y, 2: z; fused != 0: the z pass with its forward+multiply+inverse body)
int kw_bench_col(uint64_t nx, uint64_t ny, uint64_t nz, int axis, int fused, int iters, float* ms_per_pass) {
  Geometry g;
  KW_TRY(g.init(nx, ny, nz));
  float2* d = nullptr;
  float* mul = nullptr;
  KW_CUDA(cudaMalloc(&d, g.nc * sizeof(float2)));
  KW_CUDA(cudaMalloc(&mul, g.nc * sizeof(float)));
  KW_CUDA(cudaMemset(d, 0, g.nc * sizeof(float2)));
  KW_CUDA(cudaMemset(mul, 0, g.nc * sizeof(float)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  ColArgs ca{};
  ca.data[0] = d;
  const FftOps* op = axis == 1 ? g.oy : g.oz;
  if (axis == 1) ca.stride = g.nxp, ca.outer_stride = (size_t)g.ny * g.nxp, ca.ngroups = g.nxp / op->col_w, ca.tile_end = g.nz * ca.ngroups, ca.n = g.ny, ca.nvalid = g.nxr;
  else ca.stride = (size_t)g.ny * g.nxp, ca.outer_stride = g.nxp, ca.ngroups = g.nxp / op->col_w, ca.tile_end = g.ny * ca.ngroups, ca.n = g.nz, ca.nvalid = g.nxr;
  ZMidArgs za{};
  za.f = ZField{d, d, mul, 1.0f, nullptr}, za.axis = -1;
  za.nxp = g.nxp, za.ngroups = g.nxp / g.oz->zmid_w, za.ntiles = g.ny * za.ngroups, za.plane = (unsigned)((size_t)g.ny * g.nxp), za.n = g.nz, za.nvalid = g.nxr;
  for (int i = 0; i < iters + 2; ++i) {
    if (i == 2) cudaEventRecord(e0);
    if (fused) g.oz->zmid(za, 0);
    else op->col(ca, -1, 1, 0);
  }
  cudaEventRecord(e1);
  KW_CUDA(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  *ms_per_pass = ms / iters;
  cudaFree(d), cudaFree(mul);
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  KW_CUDA(cudaGetLastError());
  return KW_OK;
}

// The fused z pass on host buffers (unit-test entry of k_zmid, the kernel that carries every k-space operator of the step):
//   e      = FFT_z(in) * (mul * scal)                       in, mul: [nz][ny][nx/2+1] (cuFFT layout; mul real or NULL)
//   axis -1: out0 = IFFT_z(e)          axis 0/1/2: out0 = IFFT_z(e (x) vec_{x|y|z}[coordinate])
//   axis  3: out0/1/2 = IFFT_z(e (x) vec_x[kx]), IFFT_z(e (x) vec_y[ky]), IFFT_z(e (x) vec_z[kz])   (pressure gradient form)
// vec_x has nx/2+1, vec_y ny, vec_z nz complex entries.  Transforms are unnormalised in both directions.
int kw_fft_zmid(uint64_t nx, uint64_t ny, uint64_t nz, int axis, const float* in, const float* mul, float scal, const float* vec_x,
                const float* vec_y, const float* vec_z, float* out0, float* out1, float* out2) {
  if (!in || !out0 || axis < -1 || axis > 3) return fail(KW_ERR_INVALID, "kw_fft_zmid: bad argument");
  if ((axis == 0 || axis == 3) && !vec_x) return fail(KW_ERR_INVALID, "kw_fft_zmid: vec_x missing");
  if ((axis == 1 || axis == 3) && !vec_y) return fail(KW_ERR_INVALID, "kw_fft_zmid: vec_y missing");
  if ((axis == 2 || axis == 3) && !vec_z) return fail(KW_ERR_INVALID, "kw_fft_zmid: vec_z missing");
  if (axis == 3 && (!out1 || !out2)) return fail(KW_ERR_INVALID, "kw_fft_zmid: three outputs needed");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(KW_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  Geometry g;
  KW_TRY(g.init(nx, ny, nz));
  if (g.nz == 1) return fail(KW_ERR_INVALID, "kw_fft_zmid: Nz must be a transform length");
  const size_t rows = (size_t)g.ny * g.nz, nnat = rows * g.nxr;
  struct Scratch {
    std::vector<void*> p;
    ~Scratch() {
      for (void* q : p) cudaFree(q);
    }
  } sc;
  auto alloc = [&](void** p, size_t bytes) -> int {
    KW_CUDA(cudaMalloc(p, bytes));
    sc.p.push_back(*p);
    return KW_OK;
  };
  float2 *dnat = nullptr, *din = nullptr, *dout[3] = {}, *dv[3] = {};
  float *dmul_nat = nullptr, *dmul = nullptr;
  KW_TRY(alloc((void**)&dnat, nnat * sizeof(float2)));
  KW_TRY(alloc((void**)&din, g.nc * sizeof(float2)));
  const int nout = axis == 3 ? 3 : 1;
  // in place for the single-output forms, out of place for the gradient form: the way the time step launches them
  if (axis == 3) for (int k = 0; k < nout; ++k) KW_TRY(alloc((void**)&dout[k], g.nc * sizeof(float2)));
  else dout[0] = din;
  KW_CUDA(cudaMemcpy(dnat, in, nnat * sizeof(float2), cudaMemcpyHostToDevice));
  k_pad_complex<<<ew_grid(g.nc), 256>>>(din, dnat, g.nxr, g.nxp, rows, 1);
  if (mul) {
    KW_TRY(alloc((void**)&dmul_nat, nnat * sizeof(float)));
    KW_TRY(alloc((void**)&dmul, g.nc * sizeof(float)));
    KW_CUDA(cudaMemcpy(dmul_nat, mul, nnat * sizeof(float), cudaMemcpyHostToDevice));
    k_pad_real<<<ew_grid(g.nc), 256>>>(dmul, dmul_nat, g.nxr, g.nxp, rows);
  }
  const float* hv[3] = {vec_x, vec_y, vec_z};
  const size_t lv[3] = {(size_t)g.nxr, (size_t)g.ny, (size_t)g.nz}, lp[3] = {(size_t)g.nxp, (size_t)g.ny, (size_t)g.nz};
  for (int k = 0; k < 3; ++k)
    if (hv[k]) {
      KW_TRY(alloc((void**)&dv[k], lp[k] * sizeof(float2)));
      KW_CUDA(cudaMemset(dv[k], 0, lp[k] * sizeof(float2)));
      KW_CUDA(cudaMemcpy(dv[k], hv[k], lv[k] * sizeof(float2), cudaMemcpyHostToDevice));
    }
  ZMidArgs za{};
  za.f = ZField{din, dout[0], dmul, scal, axis == 3 ? dv[0] : axis >= 0 ? dv[axis] : nullptr, dout[1], dout[2], dv[1], dv[2]};
  za.axis = axis;
  za.nxp = g.nxp, za.ngroups = g.nxp / g.oz->zmid_w, za.ntiles = g.ny * za.ngroups, za.plane = (unsigned)((size_t)g.ny * g.nxp), za.n = g.nz, za.nvalid = g.nxr;
  g.oz->zmid(za, 0);
  KW_CUDA(cudaGetLastError());
  float* ho[3] = {out0, out1, out2};
  for (int k = 0; k < nout; ++k) {
    k_pad_complex<<<ew_grid(g.nc), 256>>>(dnat, dout[k], g.nxr, g.nxp, rows, 0);
    KW_CUDA(cudaMemcpy(ho[k], dnat, nnat * sizeof(float2), cudaMemcpyDeviceToHost));
  }
  KW_CUDA(cudaDeviceSynchronize());
  return KW_OK;
}

int kw_length_supported(uint64_t n) {
  if (n > 2048) return 0;
  switch ((int)n) {
    case 16: case 32: case 64: case 128: case 256: case 512: case 1024: return 2;
    default: return generic_length_supported((int)n) ? 1 : 0;
  }
}

int kw_fft_r2c_3d(uint64_t nx, uint64_t ny, uint64_t nz, const float* host_real, float* host_complex) {
  return fft3d_host(nx, ny, nz, host_real, host_complex, true);
}
int kw_fft_c2r_3d(uint64_t nx, uint64_t ny, uint64_t nz, const float* host_complex, float* host_real) {
  return fft3d_host(nx, ny, nz, host_complex, host_real, false);
}

}  // extern "C"
