#include "KSpaceFirstOrderSolver.h"

#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <functional>
#include <limits>
#include <new>
#include <stdexcept>
#include <thread>

namespace kwhost {

namespace {
struct NamedArray {
  const char* name;  // dataset name in the input file (Utils/MatrixNames.h)
  int id;            // kw_array
  bool isIndex;
};
// MatrixContainer::init (Containers/MatrixContainer.cpp:94-410): what is loaded from the input file
const NamedArray kMedium[] = {{"c0", KW_C0, false}, {"rho0", KW_RHO0, false}, {"rho0_sgx", KW_RHO0_SGX, false},
                              {"rho0_sgy", KW_RHO0_SGY, false}, {"rho0_sgz", KW_RHO0_SGZ, false}};
const NamedArray kOperators[] = {{"ddx_k_shift_pos_r", KW_DDX_K_SHIFT_POS_R, false}, {"ddy_k_shift_pos", KW_DDY_K_SHIFT_POS, false},
                                 {"ddz_k_shift_pos", KW_DDZ_K_SHIFT_POS, false}, {"ddx_k_shift_neg_r", KW_DDX_K_SHIFT_NEG_R, false},
                                 {"ddy_k_shift_neg", KW_DDY_K_SHIFT_NEG, false}, {"ddz_k_shift_neg", KW_DDZ_K_SHIFT_NEG, false},
                                 {"pml_x_sgx", KW_PML_X_SGX, false}, {"pml_y_sgy", KW_PML_Y_SGY, false}, {"pml_z_sgz", KW_PML_Z_SGZ, false},
                                 {"pml_x", KW_PML_X, false}, {"pml_y", KW_PML_Y, false}, {"pml_z", KW_PML_Z, false}};
const NamedArray kShifts[] = {{"x_shift_neg_r", KW_X_SHIFT_NEG_R, false}, {"y_shift_neg_r", KW_Y_SHIFT_NEG_R, false},
                              {"z_shift_neg_r", KW_Z_SHIFT_NEG_R, false}};

std::string formatSeconds(double s) {
  char buf[64];
  snprintf(buf, sizeof buf, "%8.2fs", s);
  return buf;
}
}  // namespace

KSpaceFirstOrderSolver::KSpaceFirstOrderSolver(const CommandLine& commandLine, const Team* team) : mCmd(commandLine), mTeam(team) { mTotalTime.start(); }

KSpaceFirstOrderSolver::~KSpaceFirstOrderSolver() { freeMemory(); }

void KSpaceFirstOrderSolver::log(int level, const char* fmt, ...) const {
  if (mCmd.verbose + 1 < level || !root()) return;  // 0 basic, 1 advanced, 2 full (Logger/Logger.h); rank 0 speaks for the team
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stdout, fmt, ap);
  va_end(ap);
  fflush(stdout);
}

// the C ABI reports status codes; the host turns them into the exceptions the reference throws
void KSpaceFirstOrderSolver::check(int status) const {
  if (status == KW_OK) return;
  const std::string msg = std::string("Error: ") + kw_last_error();
  switch (status) {
    case KW_ERR_ALLOC: throw std::bad_alloc();
    case KW_ERR_INVALID: throw std::invalid_argument(msg);
    default: throw std::runtime_error(msg);
  }
}

void KSpaceFirstOrderSolver::printFullCodeNameAndLicense() const {
  printf("+---------------------------------------------------------------+\n");
  printf("| %-61s |\n", getCodeName().c_str());
  printf("| B200-native time-step engine behind the k-Wave CUDA interface |\n");
  printf("| Input/output files, flags and outputs as kspaceFirstOrder-CUDA|\n");
  int version = 0;
  if (kw_cuda_code_version(&version) == KW_OK) printf("| GPU code built for compute capability %d.%d                     |\n", version / 10, version % 10);
  printf("+---------------------------------------------------------------+\n");
}

// ---------------------------------------------------------------------------------------------------------------------
void KSpaceFirstOrderSolver::readScalars() {  // Parameters::readScalarsFromInputFile (Parameters.cpp:194-553)
  const hid_t root = mInputFile.root();
  const std::string type = mInputFile.getStringAttribute(root, "/", "file_type");
  if (!type.empty() && type != "input") throw std::ios::failure("Error: The input file has not a valid format (file_type = \"" + type + "\").");
  const std::string major = mInputFile.getStringAttribute(root, "/", "major_version"), minor = mInputFile.getStringAttribute(root, "/", "minor_version");
  if (!major.empty() && (major != "1" || (minor != "0" && minor != "1")))
    throw std::ios::failure("Error: Unsupported file format version " + major + "." + minor + " (1.0 and 1.1 are supported).");
  FileScalars& s = mScalars;
  auto u = [&](const char* n) { return mInputFile.readIndexScalar(root, n); };
  auto f = [&](const char* n) { return mInputFile.readFloatScalar(root, n); };
  s.nx = u("Nx"), s.ny = u("Ny"), s.nz = u("Nz"), s.nt = u("Nt");
  s.dt = f("dt"), s.dx = f("dx"), s.dy = f("dy");
  if (s.nz > 1) s.dz = f("dz");
  s.cRef = f("c_ref");
  s.pmlXSize = u("pml_x_size"), s.pmlYSize = u("pml_y_size"), s.pmlXAlpha = f("pml_x_alpha"), s.pmlYAlpha = f("pml_y_alpha");
  if (s.nz > 1) s.pmlZSize = u("pml_z_size"), s.pmlZAlpha = f("pml_z_alpha");
  s.sensorMaskType = u("sensor_mask_type");
  if (s.sensorMaskType > 1) throw std::ios::failure("Error: The sensor mask type specified in the input file is not supported.");
  s.uxSourceFlag = u("ux_source_flag"), s.uySourceFlag = u("uy_source_flag");
  if (s.nz > 1) s.uzSourceFlag = u("uz_source_flag");
  s.transducerSourceFlag = u("transducer_source_flag"), s.pSourceFlag = u("p_source_flag"), s.p0SourceFlag = u("p0_source_flag");
  s.nonuniformGridFlag = u("nonuniform_grid_flag"), s.absorbingFlag = u("absorbing_flag"), s.nonlinearFlag = u("nonlinear_flag");
  if (s.uxSourceFlag || s.uySourceFlag || s.uzSourceFlag) s.uSourceMany = u("u_source_many"), s.uSourceMode = u("u_source_mode");
  if (s.pSourceFlag) s.pSourceMany = u("p_source_many"), s.pSourceMode = u("p_source_mode");
  if (s.absorbingFlag) {
    s.alphaPower = f("alpha_power");
    if (s.alphaPower == 1.0f) throw std::invalid_argument("Error: The value of alpha_power = 1.0 is not supported (Parameters.cpp:421-424).");
  }
  if (s.nonuniformGridFlag) throw std::invalid_argument("Error: Non-uniform grids are not supported (main.cpp:460).");
}

void KSpaceFirstOrderSolver::allocateMemory() {
  mInputFile.open(mCmd.inputFile, true);
  readScalars();
  FileScalars& s = mScalars;
  if (mCmd.benchmark) s.nt = mCmd.benchmarkSteps;  // Parameters.cpp:130-133
  if (mCmd.samplingStartIndex >= s.nt) throw std::invalid_argument("Error: The beginning of data sampling is out of the simulation time span <1, " + std::to_string(s.nt) + ">.");
  mSamplingSteps = s.nt - mCmd.samplingStartIndex;

  kw_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = KW_ABI_VERSION, cfg.struct_size = sizeof cfg;
  cfg.nx = s.nx, cfg.ny = s.ny, cfg.nz = s.nz, cfg.nt = s.nt;
  cfg.dt = s.dt, cfg.dx = s.dx, cfg.dy = s.dy, cfg.dz = s.dz, cfg.c_ref = s.cRef, cfg.alpha_power = s.alphaPower;
  cfg.nonlinear_flag = (int)s.nonlinearFlag, cfg.absorbing_flag = (int)s.absorbingFlag, cfg.nonuniform_grid_flag = (int)s.nonuniformGridFlag;
  cfg.p_source_flag = s.pSourceFlag, cfg.ux_source_flag = s.uxSourceFlag, cfg.uy_source_flag = s.uySourceFlag, cfg.uz_source_flag = s.uzSourceFlag;
  cfg.transducer_source_flag = s.transducerSourceFlag, cfg.p0_source_flag = (int)s.p0SourceFlag;
  cfg.p_source_mode = (int)s.pSourceMode, cfg.p_source_many = (int)s.pSourceMany, cfg.u_source_mode = (int)s.uSourceMode, cfg.u_source_many = (int)s.uSourceMany;
  cfg.sensor_mask_type = (int)s.sensorMaskType;
  cfg.sampling_start_index = mCmd.samplingStartIndex;
  if (mCmd.anyCompressed()) {
    // CompressHelper is initialised with the period in time steps; --frequency is converted with dt (main.cpp)
    cfg.c_period = mCmd.period > 0.f ? mCmd.period : 1.0f / (mCmd.frequency * s.dt);
    cfg.c_mos = (uint32_t)mCmd.mos, cfg.c_harmonics = (uint32_t)mCmd.harmonics;
    cfg.c_no_overlap = mCmd.noOverlap, cfg.c_40bit = mCmd.c40bit;
    const uint64_t oSize = (uint64_t)(cfg.c_period * (float)cfg.c_mos);
    mCompressedSteps = oSize ? std::max<uint64_t>(mSamplingSteps / oSize, 1) : 1;
  }
  cfg.device = mCmd.gpuDevice;
  cfg.raw_rows_capacity = 0;  // the library sizes the device-side row buffers (<= 256 MB per stream)
  cfg.rank = 0, cfg.nranks = 1;
  char ncclId[128] = {};
  if (multi()) {  // slab decomposition: one process per GPU, rank r on device (-g or 0) + r, planes [r Nz/P, (r+1) Nz/P)
    const int P = mTeam->size;
    if (s.nz % P || s.ny % P) throw std::invalid_argument("Error: --gpus " + std::to_string(P) + " must divide Ny and Nz.");
    if (mCmd.c40bit) throw std::invalid_argument("Error: --40-bit_complex runs on one GPU.");
    cfg.rank = mTeam->rank, cfg.nranks = P;
    cfg.device = (mCmd.gpuDevice < 0 ? 0 : mCmd.gpuDevice) + mTeam->rank;
    if (root()) check(kw_nccl_unique_id(ncclId, sizeof ncclId));
    mTeam->bcast(ncclId, sizeof ncclId);
    cfg.nccl_unique_id = ncclId;
    // every rank must hit a full row buffer at the same step (kw_run is collective): one capacity for all, from the complete row
    uint64_t points = 1;  // sensor points of the whole mask, from the input file
    const hid_t in = mInputFile.root();
    if (s.sensorMaskType == 0 && mInputFile.exists(in, "sensor_mask_index")) points = mInputFile.elementCount(in, "sensor_mask_index");
    if (s.sensorMaskType == 1 && mInputFile.exists(in, "sensor_mask_corners")) {
      const auto c = mInputFile.readIndices(in, "sensor_mask_corners");
      points = 0;
      for (size_t k = 0; k + 5 < c.size(); k += 6) points += (c[k + 3] - c[k] + 1) * (c[k + 4] - c[k + 1] + 1) * (c[k + 5] - c[k + 2] + 1);
    }
    const uint64_t perRow = 4 * (points / P + 1) * (mCmd.anyCompressed() ? 2 * std::max<uint64_t>(mCmd.harmonics, 1) : 1);
    cfg.raw_rows_capacity = std::max<uint64_t>(1, std::min<uint64_t>(mSamplingSteps, (64ull << 20) / perRow));
    mNzLocal = s.nz / P, mZ0 = mTeam->rank * mNzLocal;
  } else {
    mNzLocal = s.nz, mZ0 = 0;
  }
  check(kw_ctx_create(&cfg, &mCtx));
}

void KSpaceFirstOrderSolver::freeMemory() {
  for (auto& st : mStreams) {
    if (st.dataset >= 0) mOutputFile.closeDataset(st.dataset);
    for (hid_t d : st.cuboidDatasets) mOutputFile.closeDataset(d);
    if (st.group >= 0) mOutputFile.closeGroup(st.group);
  }
  mStreams.clear();
  if (mCtx) kw_ctx_destroy(mCtx);
  mCtx = nullptr;
  mInputFile.close();
  mOutputFile.close();
}

void KSpaceFirstOrderSolver::loadArray(const std::string& name, int arrayId, bool isIndex, bool required) {
  const hid_t root = mInputFile.root();
  if (!mInputFile.exists(root, name)) {
    if (required) throw std::ios::failure("Error: dataset \"" + name + "\" is missing in the input file.");
    return;
  }
  if (isIndex) {
    const auto v = mInputFile.readIndices(root, name);
    mHostBytes = std::max(mHostBytes, v.size() * sizeof(uint64_t));
    check(kw_set_array(mCtx, arrayId, v.data(), v.size()));
    if (arrayId == KW_SENSOR_MASK_CORNERS) mCorners = v;
    if (arrayId == KW_SENSOR_MASK_INDEX) mSensorPoints = v.size();
  } else if (multi() && mInputFile.elementCount(root, name) == mScalars.nx * mScalars.ny * mScalars.nz && mScalars.nx * mScalars.ny * mScalars.nz > 1) {
    // a full-grid array: this rank loads its z-slab only (hyperslab read; minih5 serves it from the file mapping)
    const FileScalars& s = mScalars;
    std::vector<float> v(s.nx * s.ny * mNzLocal);
    const hid_t d = mInputFile.openDataset(root, name);
    mInputFile.readHyperslab(d, {mZ0, 0, 0}, {mNzLocal, s.ny, s.nx}, v.data());
    mInputFile.closeDataset(d);
    mHostBytes = std::max(mHostBytes, v.size() * sizeof(float));
    check(kw_set_array(mCtx, arrayId, v.data(), v.size()));
  } else {
    const auto v = mInputFile.readFloats(root, name);
    mHostBytes = std::max(mHostBytes, v.size() * sizeof(float));
    const std::string domain = mInputFile.getStringAttribute(root, name, "domain_type");
    const bool isComplex = domain == "complex";
    check(kw_set_array(mCtx, arrayId, v.data(), isComplex ? v.size() / 2 : v.size()));
  }
}

void KSpaceFirstOrderSolver::loadInputData() {
  mDataLoadTime.start();
  const FileScalars& s = mScalars;
  const bool is3D = s.nz > 1;  // the input file of a 2-D simulation carries no z arrays (main.cpp:446-563)
  auto isZArray = [](int id) {
    return id == KW_RHO0_SGZ || id == KW_DDZ_K_SHIFT_POS || id == KW_DDZ_K_SHIFT_NEG || id == KW_PML_Z || id == KW_PML_Z_SGZ || id == KW_Z_SHIFT_NEG_R;
  };
  for (const auto& a : kMedium) loadArray(a.name, a.id, a.isIndex, is3D || !isZArray(a.id));
  if (s.nonlinearFlag) loadArray("BonA", KW_BONA, false, true);
  if (s.absorbingFlag) loadArray("alpha_coeff", KW_ALPHA_COEFF, false, true);
  for (const auto& a : kOperators) loadArray(a.name, a.id, a.isIndex, is3D || !isZArray(a.id));
  const bool needShift = mCmd.uNonStaggeredRaw || mCmd.uNonStaggeredC || mCmd.iAvgC || mCmd.qTermC || mCmd.iAvg || mCmd.qTerm;
  for (const auto& a : kShifts) loadArray(a.name, a.id, a.isIndex, needShift && (is3D || !isZArray(a.id)));
  if (s.sensorMaskType == 0) loadArray("sensor_mask_index", KW_SENSOR_MASK_INDEX, true, true);
  else loadArray("sensor_mask_corners", KW_SENSOR_MASK_CORNERS, true, true);
  if (s.p0SourceFlag) loadArray("p0_source_input", KW_P0_SOURCE_INPUT, false, true);
  if (s.pSourceFlag) {
    loadArray("p_source_index", KW_P_SOURCE_INDEX, true, true);
    loadArray("p_source_input", KW_P_SOURCE_INPUT, false, true);
  }
  if (s.uxSourceFlag || s.uySourceFlag || s.uzSourceFlag || s.transducerSourceFlag) loadArray("u_source_index", KW_U_SOURCE_INDEX, true, true);
  if (s.uxSourceFlag) loadArray("ux_source_input", KW_UX_SOURCE_INPUT, false, true);
  if (s.uySourceFlag) loadArray("uy_source_input", KW_UY_SOURCE_INPUT, false, true);
  if (s.uzSourceFlag) loadArray("uz_source_input", KW_UZ_SOURCE_INPUT, false, true);
  if (s.transducerSourceFlag) {
    loadArray("delay_mask", KW_DELAY_MASK, true, true);
    loadArray("transducer_source_input", KW_TRANSDUCER_SOURCE_INPUT, false, true);
  }
  if (s.sensorMaskType == 1) {
    if (mCorners.empty() || mCorners.size() % 6) throw std::ios::failure("Error: sensor_mask_corners has not a valid format.");
    mSensorPoints = 0;
    for (size_t k = 0; k < mCorners.size(); k += 6)
      mSensorPoints += (mCorners[k + 3] - mCorners[k] + 1) * (mCorners[k + 4] - mCorners[k + 1] + 1) * (mCorners[k + 5] - mCorners[k + 2] + 1);
  }
  // cpp:185-240: a run with checkpointing enabled whose checkpoint file exists continues that run
  mRecover = mCmd.isCheckpointEnabled() && Hdf5File::canAccess(mCmd.checkpointFile);
  if (!root()) {  // rank 0 owns the output file
  } else if (mRecover) {
    if (!Hdf5File::canAccess(mCmd.outputFile)) throw std::ios::failure("Error: The output file of the checkpointed run \"" + mCmd.outputFile + "\" is missing.");
    mOutputFile.open(mCmd.outputFile, false);
    const std::string type = mOutputFile.getStringAttribute(mOutputFile.root(), "/", "file_type");
    if (type != "output") throw std::ios::failure("Error: \"" + mCmd.outputFile + "\" is not an output file of a checkpointed run.");
  } else if (mCmd.post) {  // cpp:231-239: post-processing of an existing output file, opened read-write
    if (!Hdf5File::canAccess(mCmd.outputFile)) throw std::ios::failure("Error: --post: the output file \"" + mCmd.outputFile + "\" does not exist.");
    mOutputFile.open(mCmd.outputFile, false);
  } else {
    mOutputFile.create(mCmd.outputFile);  // cpp:230-235
  }
  mStepsToCheckpoint = mCmd.checkpointTimeSteps ? mCmd.checkpointTimeSteps : ~0ull;  // Parameters.cpp:147-151
  mDataLoadTime.stop();
}

// ---------------------------------------------------------------------------------------------------------------------
// ---- slab-decomposed runs: host-side gathers towards rank 0 (the data plane between the GPUs lives in the engine library) ----------
void KSpaceFirstOrderSolver::exchangeSensorLayout() {
  uint64_t total = 0, local = 0;
  check(kw_sensor_layout(mCtx, &total, &local, nullptr, 0));
  mLocalPoints = local;
  if (!multi()) return;
  mPositions.assign(local, 0);
  if (local) check(kw_sensor_layout(mCtx, &total, &local, mPositions.data(), mPositions.size()));
  if (total) mSensorPoints = total;
  if (root()) {
    mRankPositions.assign(mTeam->size, {});
    mRankPositions[0] = mPositions;
    for (int r = 1; r < mTeam->size; ++r) mRankPositions[r] = mTeam->recvVec<uint64_t>(r);
  } else {
    mTeam->sendVec(0, mPositions);
  }
}
// `rows` rows of `localFloats` floats (perPoint floats per local sensor point) -> rows of fullFloats floats in mask order on rank 0
std::vector<float> KSpaceFirstOrderSolver::gatherPoints(const float* local, uint64_t rows, uint64_t localFloats, uint64_t perPoint, uint64_t fullFloats) {
  if (!multi()) return std::vector<float>(local, local + rows * localFloats);
  if (!root()) {
    mTeam->send(0, &rows, sizeof rows);
    if (rows && localFloats) mTeam->send(0, local, rows * localFloats * sizeof(float));
    return {};
  }
  std::vector<float> full(rows * fullFloats, 0.f), part;
  for (int r = 0; r < mTeam->size; ++r) {
    const std::vector<uint64_t>& pos = mRankPositions[r];
    const uint64_t lf = pos.size() * perPoint;
    const float* src = local;
    if (r > 0) {
      uint64_t theirs = 0;
      mTeam->recv(r, &theirs, sizeof theirs);
      if (theirs != rows) throw std::runtime_error("Error: rank " + std::to_string(r) + " holds a different number of buffered rows.");
      part.resize(rows * lf);
      if (rows && lf) mTeam->recv(r, part.data(), rows * lf * sizeof(float));
      src = part.data();
    }
    for (uint64_t row = 0; row < rows; ++row)
      for (size_t j = 0; j < pos.size(); ++j) memcpy(&full[row * fullFloats + pos[j] * perPoint], src + row * lf + j * perPoint, perPoint * sizeof(float));
  }
  return full;
}
// this rank's z-slab of a grid -> the whole grid on rank 0 (slabs are consecutive plane ranges)
std::vector<float> KSpaceFirstOrderSolver::gatherSlabs(const float* local) {
  const FileScalars& s = mScalars;
  const uint64_t slab = s.nx * s.ny * mNzLocal;
  if (!multi()) return std::vector<float>(local, local + slab);
  if (!root()) {
    mTeam->send(0, local, slab * sizeof(float));
    return {};
  }
  std::vector<float> full(s.nx * s.ny * s.nz);
  memcpy(full.data(), local, slab * sizeof(float));
  for (int r = 1; r < mTeam->size; ++r) mTeam->recv(r, full.data() + (uint64_t)r * slab, slab * sizeof(float));
  return full;
}
// the entries of a complete per-sensor buffer that belong to this rank's points (read from a file every rank can open)
void KSpaceFirstOrderSolver::scatterPointsFrom(const std::vector<float>& full, uint64_t perPoint, std::vector<float>* local) const {
  if (!multi()) {
    *local = full;
    return;
  }
  local->resize(mPositions.size() * perPoint);
  for (size_t j = 0; j < mPositions.size(); ++j) memcpy(&(*local)[j * perPoint], &full[mPositions[j] * perPoint], perPoint * sizeof(float));
}

// rank 0 holds a complete buffer (per-sensor values or a whole grid): every rank gets the part it owns
std::vector<float> KSpaceFirstOrderSolver::distribute(const std::vector<float>& full, bool wholeDomain, uint64_t perPoint, uint64_t fullFloats) {
  const FileScalars& s = mScalars;
  if (wholeDomain) {
    const uint64_t slab = s.nx * s.ny * mNzLocal;
    std::vector<float> mine(slab);
    if (root()) {
      for (int r = 1; r < mTeam->size; ++r) mTeam->send(r, full.data() + (uint64_t)r * slab, slab * sizeof(float));
      memcpy(mine.data(), full.data(), slab * sizeof(float));
    } else {
      mTeam->recv(0, mine.data(), slab * sizeof(float));
    }
    return mine;
  }
  std::vector<float> all(fullFloats);
  if (root()) all = full;
  mTeam->bcast(all.data(), all.size() * sizeof(float));
  std::vector<float> mine;
  scatterPointsFrom(all, perPoint, &mine);
  return mine;
}

void KSpaceFirstOrderSolver::createStreams() {  // OutputStreamContainer::init (Containers/OutputStreamContainer.cpp:70-325)
  using K = OutputStream::Kind;
  auto add = [&](bool on, int id, const std::string& name, K kind, bool shifted = false) {
    if (!on) return;
    check(kw_stream_enable(mCtx, id));
    OutputStream st{};
    st.id = id, st.name = name, st.kind = kind, st.shifted = shifted;
    mStreams.push_back(st);
  };
  const char* axes[3] = {"x", "y", "z"};
  add(mCmd.pRaw, KW_S_P_RAW, "p", K::kSeries);
  add(mCmd.pC, KW_S_P_C, "p_c", K::kCompressed);
  add(mCmd.pRms, KW_S_P_RMS, "p_rms", K::kAggregate);
  add(mCmd.pMax, KW_S_P_MAX, "p_max", K::kAggregate);
  add(mCmd.pMin, KW_S_P_MIN, "p_min", K::kAggregate);
  add(mCmd.pMaxAll, KW_S_P_MAX_ALL, "p_max_all", K::kWholeDomain);
  add(mCmd.pMinAll, KW_S_P_MIN_ALL, "p_min_all", K::kWholeDomain);
  const int ncomp = mScalars.nz > 1 ? 3 : 2;  // 2-D: no z component streams (OutputStreamContainer.cpp: is3DSimulation)
  for (int k = 0; k < ncomp; ++k) {
    const std::string u = std::string("u") + axes[k];
    add(mCmd.uRaw, KW_S_UX_RAW + k, u, K::kSeries);
    add(mCmd.uC, KW_S_UX_C + k, u + "_c", K::kCompressed);
    add(mCmd.uNonStaggeredRaw, KW_S_UX_NS_RAW + k, u + "_non_staggered", K::kSeries);
    add(mCmd.uNonStaggeredC, KW_S_UX_NS_C + k, u + "_non_staggered_c", K::kCompressed, true);
    add(mCmd.uRms, KW_S_UX_RMS + k, u + "_rms", K::kAggregate);
    add(mCmd.uMax, KW_S_UX_MAX + k, u + "_max", K::kAggregate);
    add(mCmd.uMin, KW_S_UX_MIN + k, u + "_min", K::kAggregate);
    add(mCmd.uMaxAll, KW_S_UX_MAX_ALL + k, u + "_max_all", K::kWholeDomain);
    add(mCmd.uMinAll, KW_S_UX_MIN_ALL + k, u + "_min_all", K::kWholeDomain);
    add(mCmd.iAvgC, KW_S_IX_AVG_C + k, std::string("I") + axes[k] + "_avg_c", K::kAggregate);
  }
  add(mCmd.qTermC, KW_S_Q_TERM_C, "Q_term_c", K::kAggregate);
  // flush order = OutputStreamIdx order (Containers/OutputStreamContainer.h:59-150)
  std::stable_sort(mStreams.begin(), mStreams.end(), [](const OutputStream& a, const OutputStream& b) { return a.id < b.id; });
}

void KSpaceFirstOrderSolver::createOutputDatasets() {
  using K = OutputStream::Kind;
  if (mCmd.post) return;  // --post works on the datasets an earlier run stored (postProcessOnly)
  const hid_t root = mOutputFile.root();
  const bool cuboids = mScalars.sensorMaskType == 1;
  const unsigned deflate = mCmd.compressionLevel;
  for (auto& st : mStreams) {
    uint64_t rowFloats = 0, rows = 0;
    check(kw_stream_info(mCtx, st.id, &rowFloats, &rows));
    st.localFloats = rowFloats;
    st.perPoint = st.kind == K::kCompressed ? 2 * mCmd.harmonics : 1;
    // the row in the file covers every sensor point (slab-decomposed runs: this rank holds mLocalPoints of them)
    st.rowFloats = !multi() ? rowFloats : st.kind == K::kWholeDomain ? mScalars.nx * mScalars.ny * mScalars.nz : mSensorPoints * st.perPoint;
    rowFloats = st.rowFloats;
    if (!this->root()) continue;
    if (st.kind == K::kWholeDomain) {  // WholeDomainOutputStream::create (:78-99): (Nx, Ny, Nz), chunk (Nx, Ny, 1); reopen on recovery
      const FileScalars& s = mScalars;
      st.dataset = mRecover ? mOutputFile.openDataset(root, st.name)
                            : mOutputFile.createDataset(root, st.name, {s.nz, s.ny, s.nx}, {1, s.ny, s.nx}, true, deflate);
      continue;
    }
    const bool series = st.kind == K::kSeries || st.kind == K::kCompressed;
    const uint64_t nRows = st.kind == K::kSeries ? mSamplingSteps : st.kind == K::kCompressed ? mCompressedSteps : 0;
    auto compressionAttributes = [&](hid_t loc, const std::string& name) {  // IndexOutputStream.cpp:147-157
      mOutputFile.setLongLongAttribute(loc, name, "c_harmonics", (long long)mCmd.harmonics);
      mOutputFile.setStringAttribute(loc, name, "c_type", "c");
      float period = mCmd.period > 0.f ? mCmd.period : 1.0f / (mCmd.frequency * mScalars.dt);
      mOutputFile.setFloatAttribute(loc, name, "c_period", period);
      mOutputFile.setLongLongAttribute(loc, name, "c_mos", (long long)mCmd.mos);
      mOutputFile.setLongLongAttribute(loc, name, "c_shift", st.shifted ? 1 : 0);
      mOutputFile.setFloatAttribute(loc, name, "c_complex_size", mCmd.c40bit ? 1.25f : 2.0f);
      mOutputFile.setLongLongAttribute(loc, name, "c_max_exp", st.shifted ? 114 : 138);
    };
    if (!cuboids) {  // IndexOutputStream::create (:87-160): (Nsens, Nt - s, 1), i.e. [1][rows][Nsens] on disk
      const hsize_t width = rowFloats;
      const std::vector<hsize_t> dims = series ? std::vector<hsize_t>{1, nRows, width} : std::vector<hsize_t>{1, 1, width};
      std::vector<hsize_t> chunk = {1, 1, width > (1ull << 23) ? (1ull << 20) : width};
      if (mRecover) {
        st.dataset = mOutputFile.openDataset(root, st.name);  // IndexOutputStream::reopen (:177-247)
      } else {
        st.dataset = mOutputFile.createDataset(root, st.name, dims, chunk, true, deflate);
        if (st.kind == K::kCompressed) compressionAttributes(root, st.name);
      }
    } else {  // CuboidOutputStream::create (:80-150): group /<name>, one dataset per cuboid named 1..Ncub
      if (st.kind == K::kCompressed && mCmd.c40bit)
        throw std::invalid_argument("Error: --40-bit_complex with a cuboid sensor mask is not available in this build.");
      st.group = mRecover ? mOutputFile.openGroup(root, st.name) : mOutputFile.createGroup(root, st.name);
      const uint64_t factor = st.kind == K::kCompressed ? 2 * mCmd.harmonics : 1;
      for (size_t k = 0; k < mCorners.size() / 6; ++k) {
        const uint64_t* c = &mCorners[6 * k];
        const hsize_t cx = (c[3] - c[0] + 1) * factor, cy = c[4] - c[1] + 1, cz = c[5] - c[2] + 1;
        const std::string name = std::to_string(k + 1);
        std::vector<hsize_t> dims, chunk;
        if (series) dims = {nRows, cz, cy, cx}, chunk = {1, cz, cy, cx};
        else dims = {cz, cy, cx}, chunk = {cz, cy, cx};
        if (cx * cy * cz > (1ull << 23)) {  // >= 32 MB per step: ~4 MB slabs (CuboidOutputStream.cpp:676-684)
          hsize_t slabs = 1;
          while (slabs * cx * cy < (1ull << 20)) ++slabs;
          chunk[series ? 1 : 0] = std::min<hsize_t>(slabs, cz);
        }
        if (mRecover) {
          st.cuboidDatasets.push_back(mOutputFile.openDataset(st.group, name));  // CuboidOutputStream::reopen (:156-257)
          continue;
        }
        st.cuboidDatasets.push_back(mOutputFile.createDataset(st.group, name, dims, chunk, true, deflate));
        if (st.kind == K::kCompressed) compressionAttributes(st.group, name);
      }
    }
  }
}

// rows buffered on the device -> output file (IndexOutputStream::flushBufferToFile :583-591, CuboidOutputStream :560-620)
// onlyLanded: only the chunks the library has already sent towards pinned host memory (asynchronous output, kw_stream_async): the
// time loop is not held up for rows that still sit in a half-filled device buffer
void KSpaceFirstOrderSolver::flushSeries(bool onlyLanded) {
  using K = OutputStream::Kind;
  for (auto& st : mStreams) {
    if (st.kind != K::kSeries && st.kind != K::kCompressed) continue;
   for (;;) {  // a landed chunk first, then (unless onlyLanded) the rows still on the device
    uint64_t rowFloats = 0, rows = 0, pending = 0;
    check(kw_stream_pending(mCtx, st.id, &pending));
    if (onlyLanded && pending == 0) break;
    check(kw_stream_info(mCtx, st.id, &rowFloats, &rows));
    if (rows == 0) break;
    if (mRowBuffer.size() < std::max<uint64_t>(rows * rowFloats, 1)) mRowBuffer.resize(std::max<uint64_t>(rows * rowFloats, 1));
    uint64_t got = 0;
    check(kw_stream_fetch(mCtx, st.id, mRowBuffer.data(), mRowBuffer.size(), &got));  // (a rank without sensor points still counts rows)
    if (multi()) {  // rows of the local sensor points -> complete rows on rank 0
      std::vector<float> full = gatherPoints(mRowBuffer.data(), got, st.localFloats, st.perPoint, st.rowFloats);
      if (!root()) {
        st.rowsWritten += got;
        continue;  // next chunk of this stream
      }
      mRowBuffer.swap(full);
      rowFloats = st.rowFloats;
    }
    if (st.dataset >= 0) {
      mOutputFile.writeHyperslab(st.dataset, {0, st.rowsWritten, 0}, {1, got, rowFloats}, mRowBuffer.data());
    } else {
      const uint64_t factor = st.kind == K::kCompressed ? 2 * mCmd.harmonics : 1;
      uint64_t offset = 0;
      std::vector<float> part;
      for (size_t k = 0; k < st.cuboidDatasets.size(); ++k) {
        const uint64_t* c = &mCorners[6 * k];
        const hsize_t cx = (c[3] - c[0] + 1) * factor, cy = c[4] - c[1] + 1, cz = c[5] - c[2] + 1;
        const uint64_t n = cx * cy * cz;
        part.resize(got * n);
        for (uint64_t r = 0; r < got; ++r) memcpy(&part[r * n], &mRowBuffer[r * rowFloats + offset], n * sizeof(float));
        mOutputFile.writeHyperslab(st.cuboidDatasets[k], {st.rowsWritten, 0, 0, 0}, {got, cz, cy, cx}, part.data());
        offset += n;
      }
    }
    st.rowsWritten += got;
   }
  }
}

// the accumulator of one aggregate / whole-domain stream -> its dataset(s) in the output file
void KSpaceFirstOrderSolver::writeStreamBuffer(OutputStream& st, const float* buf) {
  using K = OutputStream::Kind;
  const FileScalars& s = mScalars;
  if (st.kind == K::kWholeDomain) {
    mOutputFile.writeHyperslab(st.dataset, {0, 0, 0}, {s.nz, s.ny, s.nx}, buf);
  } else if (st.dataset >= 0) {
    mOutputFile.writeHyperslab(st.dataset, {0, 0, 0}, {1, 1, st.rowFloats}, buf);
  } else {
    uint64_t offset = 0;
    for (size_t k = 0; k < st.cuboidDatasets.size(); ++k) {
      const uint64_t* c = &mCorners[6 * k];
      const hsize_t cx = c[3] - c[0] + 1, cy = c[4] - c[1] + 1, cz = c[5] - c[2] + 1;
      mOutputFile.writeHyperslab(st.cuboidDatasets[k], {0, 0, 0}, {cz, cy, cx}, buf + offset);
      offset += cx * cy * cz;
    }
  }
}
void KSpaceFirstOrderSolver::readStreamBuffer(OutputStream& st, float* buf) {
  using K = OutputStream::Kind;
  const FileScalars& s = mScalars;
  if (st.kind == K::kWholeDomain) {
    mOutputFile.readHyperslab(st.dataset, {0, 0, 0}, {s.nz, s.ny, s.nx}, buf);
  } else if (st.dataset >= 0) {
    mOutputFile.readHyperslab(st.dataset, {0, 0, 0}, {1, 1, st.rowFloats}, buf);
  } else {
    uint64_t offset = 0;
    for (size_t k = 0; k < st.cuboidDatasets.size(); ++k) {
      const uint64_t* c = &mCorners[6 * k];
      const hsize_t cx = c[3] - c[0] + 1, cy = c[4] - c[1] + 1, cz = c[5] - c[2] + 1;
      mOutputFile.readHyperslab(st.cuboidDatasets[k], {0, 0, 0}, {cz, cy, cx}, buf + offset);
      offset += cx * cy * cz;
    }
  }
}

void KSpaceFirstOrderSolver::writeAggregates() {
  using K = OutputStream::Kind;
  const FileScalars& s = mScalars;
  const hid_t root = mOutputFile.root();
  for (auto& st : mStreams) {
    if (st.kind != K::kAggregate && st.kind != K::kWholeDomain) continue;
    std::vector<float> buf(std::max<uint64_t>(st.localFloats, 1));
    uint64_t got = 0;
    if (st.localFloats) check(kw_stream_fetch(mCtx, st.id, buf.data(), buf.size(), &got));
    if (multi()) buf = st.kind == K::kWholeDomain ? gatherSlabs(buf.data()) : gatherPoints(buf.data(), 1, st.localFloats, st.perPoint, st.rowFloats);
    if (this->root()) writeStreamBuffer(st, buf.data());
  }
  // p_final / u*_final (cpp:952-973; RealMatrix::writeData :88-121)
  std::vector<float> field;
  auto final_field = [&](bool on, int id, const char* name) {
    if (!on) return;
    field.resize(s.nx * s.ny * mNzLocal);
    check(kw_get_array(mCtx, id, field.data(), field.size()));
    if (multi()) field = gatherSlabs(field.data());
    if (this->root()) mOutputFile.writeWhole(root, name, {s.nz, s.ny, s.nx}, {1, s.ny, s.nx}, field.data(), true, mCmd.compressionLevel);
  };
  final_field(mCmd.pFinal, KW_P, "p_final");
  final_field(mCmd.uFinal, KW_UX_SGX, "ux_final");
  final_field(mCmd.uFinal, KW_UY_SGY, "uy_final");
  final_field(mCmd.uFinal && s.nz > 1, KW_UZ_SGZ, "uz_final");
}

// one value per sensor point in the layout of an aggregated stream: (Nsens,1,1), or one 3-D dataset per cuboid in a group
void KSpaceFirstOrderSolver::writeSensorValues(const std::string& name, const float* data) {
  const hid_t root = mOutputFile.root();
  if (mScalars.sensorMaskType == 0) {
    mOutputFile.writeWhole(root, name, {1, 1, mSensorPoints}, {1, 1, mSensorPoints}, data, true, mCmd.compressionLevel);
    return;
  }
  const hid_t group = mOutputFile.createGroup(root, name);
  uint64_t offset = 0;
  for (size_t k = 0; k < mCorners.size() / 6; ++k) {
    const uint64_t* c = &mCorners[6 * k];
    const hsize_t cx = c[3] - c[0] + 1, cy = c[4] - c[1] + 1, cz = c[5] - c[2] + 1;
    mOutputFile.writeWhole(group, std::to_string(k + 1), {cz, cy, cx}, {cz, cy, cx}, data + offset, true, mCmd.compressionLevel);
    offset += cx * cy * cz;
  }
  mOutputFile.closeGroup(group);
}

// Average intensity from the stored raw series (computeAverageIntensities, cpp:1231-1534): p and the non-staggered velocity
// are read back from the output file block by block (all stored steps of a block of sensor points), the device shifts the
// velocity by half a time step and accumulates p * u; then the Q term from the three intensities (computeQTerm, :1783-2080).
void KSpaceFirstOrderSolver::computeAverageIntensities() {
  const int ncomp = mScalars.nz > 1 ? 3 : 2;
  const uint64_t steps = mSamplingSteps;
  const OutputStream* sp = nullptr;
  const OutputStream* su[3] = {};
  for (const auto& st : mStreams) {
    if (st.id == KW_S_P_RAW) sp = &st;
    for (int k = 0; k < ncomp; ++k)
      if (st.id == KW_S_UX_NS_RAW + k) su[k] = &st;
  }
  if (!sp || !su[0] || !su[1] || (ncomp == 3 && !su[2])) throw std::runtime_error("Error: the raw series needed by --I_avg / --Q_term were not stored.");
  // slab-decomposed runs: the stored series live in rank 0's output file, so rank 0 alone forms the intensities
  // (kw_intensity_avg_block needs no context); the Q term is a 3-D transform and runs on every rank (writeQTerm)
  std::vector<std::vector<float>> intensity(ncomp, std::vector<float>(root() ? mSensorPoints : 0, 0.f));
  // block size: --block_size points, else what keeps the five host / device buffers of a block near 1 GB (cpp:1281-1300)
  uint64_t maxPoints = mCmd.blockSize ? mCmd.blockSize : std::max<uint64_t>(1, (48ull << 20) / steps);
  std::vector<float> bp, bu[3];
  auto process = [&](uint64_t first, uint64_t n, const std::function<void(const OutputStream&, float*)>& read) {
    bp.resize(n * steps);
    read(*sp, bp.data());
    const float* up[3] = {};
    float* ip[3] = {};
    for (int k = 0; k < ncomp; ++k) {
      bu[k].resize(n * steps);
      read(*su[k], bu[k].data());
      up[k] = bu[k].data(), ip[k] = intensity[k].data() + first;
    }
    check(kw_intensity_avg_block(bp.data(), up, ncomp, n, steps, ip));
  };
  if (!root()) {
  } else if (mScalars.sensorMaskType == 0) {
    for (uint64_t i = 0; i < mSensorPoints; i += maxPoints) {
      const uint64_t n = std::min<uint64_t>(maxPoints, mSensorPoints - i);
      process(i, n, [&](const OutputStream& st, float* dst) { mOutputFile.readHyperslab(st.dataset, {0, 0, i}, {1, steps, n}, dst); });
    }
  } else {
    uint64_t offset = 0;
    for (size_t k = 0; k < mCorners.size() / 6; ++k) {
      const uint64_t* c = &mCorners[6 * k];
      const hsize_t cx = c[3] - c[0] + 1, cy = c[4] - c[1] + 1, cz = c[5] - c[2] + 1;
      const uint64_t slab = std::max<uint64_t>(1, maxPoints / (cx * cy));  // whole z slices of the cuboid per block
      for (uint64_t z = 0; z < cz; z += slab) {
        const uint64_t zc = std::min<uint64_t>(slab, cz - z);
        process(offset + z * cx * cy, zc * cx * cy,
                [&](const OutputStream& st, float* dst) { mOutputFile.readHyperslab(st.cuboidDatasets[k], {0, z, 0, 0}, {steps, zc, cy, cx}, dst); });
      }
      offset += cx * cy * cz;
    }
  }
  const char* names[3] = {"Ix_avg", "Iy_avg", "Iz_avg"};
  if (mCmd.iAvg && root())
    for (int k = 0; k < ncomp; ++k) replaceSensorValues(names[k], intensity[k].data());
  if (mCmd.qTerm) writeQTerm(intensity, "Q_term");
}

// Q = -div I of per-sensor intensities held by rank 0 (complete rows in mask order) -> dataset `name`.  kw_q_term is collective on a
// slab-decomposed context: every rank gets the values of its own sensor points, the results return to rank 0 in mask order.
void KSpaceFirstOrderSolver::writeQTerm(const std::vector<std::vector<float>>& intensity, const char* name) {
  const int ncomp = mScalars.nz > 1 ? 3 : 2;
  if (!multi()) {
    std::vector<float> q(mSensorPoints, 0.f);
    const float* ip[3] = {intensity[0].data(), intensity[1].data(), ncomp == 3 ? intensity[2].data() : nullptr};
    check(kw_q_term(mCtx, ip, ncomp, q.data(), q.size()));
    replaceSensorValues(name, q.data());
    return;
  }
  std::vector<float> local[3];
  const float* ip[3] = {};
  for (int k = 0; k < ncomp; ++k) {
    local[k] = distribute(intensity[k], false, 1, mSensorPoints);
    local[k].resize(std::max<uint64_t>(mLocalPoints, 1));
    ip[k] = local[k].data();
  }
  std::vector<float> ql(std::max<uint64_t>(mLocalPoints, 1), 0.f);
  check(kw_q_term(mCtx, ip, ncomp, ql.data(), ql.size()));
  const std::vector<float> q = gatherPoints(ql.data(), 1, mLocalPoints, 1, mSensorPoints);
  if (root()) replaceSensorValues(name, q.data());
}

// replaces a dataset of per-sensor values (or the per-cuboid group) when an earlier run already stored it
void KSpaceFirstOrderSolver::replaceSensorValues(const std::string& name, const float* data) {
  const hid_t root = mOutputFile.root();
  if (H5Lexists(root, name.c_str(), H5P_DEFAULT) <= 0) {
    writeSensorValues(name, data);
    return;
  }
  if (mScalars.sensorMaskType == 0) {
    const hid_t d = mOutputFile.openDataset(root, name);
    mOutputFile.writeHyperslab(d, {0, 0, 0}, {1, 1, mSensorPoints}, data);
    mOutputFile.closeDataset(d);
    return;
  }
  const hid_t group = mOutputFile.openGroup(root, name);
  uint64_t offset = 0;
  for (size_t k = 0; k < mCorners.size() / 6; ++k) {
    const uint64_t* c = &mCorners[6 * k];
    const hsize_t cx = c[3] - c[0] + 1, cy = c[4] - c[1] + 1, cz = c[5] - c[2] + 1;
    const hid_t d = mOutputFile.openDataset(group, std::to_string(k + 1));
    mOutputFile.writeHyperslab(d, {0, 0, 0}, {cz, cy, cx}, data + offset);
    mOutputFile.closeDataset(d);
    offset += cx * cy * cz;
  }
  mOutputFile.closeGroup(group);
}

// computeAverageIntensitiesC (KSpaceFirstOrderSolver.cpp:1543-1775): the time-averaged intensity from the STORED compression
// coefficients p_c and u?_non_staggered_c -- for every sensor point the mean over the stored frames of
// sum_h Re(P_h conj(U_h)) / 2, accumulated frame by frame and harmonic by harmonic as the reference does.
void KSpaceFirstOrderSolver::computeAverageIntensitiesC(std::vector<std::vector<float>>& intensity) {
  const int ncomp = mScalars.nz > 1 ? 3 : 2;
  const uint64_t H = mCmd.harmonics;
  const hid_t root = mOutputFile.root();
  const char* unames[3] = {"ux_non_staggered_c", "uy_non_staggered_c", "uz_non_staggered_c"};
  intensity.assign(ncomp, std::vector<float>(mSensorPoints, 0.f));
  if (mCmd.c40bit) throw std::invalid_argument("Error: --post with --40-bit_complex is not available (the reference has no such path either, cpp:1580).");
  auto accumulate = [&](const std::vector<float>& p, const std::vector<float>* u, uint64_t frames, uint64_t npts, uint64_t first) {
    // frame layout: npts * H complex values (re, im), sensor-major then harmonic
    for (uint64_t fr = 0; fr < frames; ++fr)
      for (uint64_t x = 0; x < npts; ++x)
        for (uint64_t ih = 0; ih < H; ++ih) {
          const uint64_t q = 2 * (fr * npts * H + x * H + ih);
          for (int k = 0; k < ncomp; ++k) intensity[k][first + x] += (p[q] * u[k][q] + p[q + 1] * u[k][q + 1]) / 2.0f;
        }
    for (uint64_t x = 0; x < npts; ++x)
      for (int k = 0; k < ncomp; ++k) intensity[k][first + x] /= (float)frames;
  };
  std::vector<float> u[3];
  if (mScalars.sensorMaskType == 0) {
    for (const char* n : {"p_c", unames[0], unames[1]})
      if (!mOutputFile.exists(root, n)) throw std::ios::failure(std::string("Error: --post: dataset \"") + n + "\" is missing in the output file (store --p_c --u_non_staggered_c).");
    const std::vector<float> p = mOutputFile.readFloats(root, "p_c");
    for (int k = 0; k < ncomp; ++k) u[k] = mOutputFile.readFloats(root, unames[k]);
    const uint64_t frames = p.size() / (2 * H * mSensorPoints);
    if (!frames || p.size() != frames * 2 * H * mSensorPoints) throw std::ios::failure("Error: --post: p_c does not match the sensor mask and --harmonics.");
    for (int k = 0; k < ncomp; ++k)
      if (u[k].size() != p.size()) throw std::ios::failure("Error: --post: the stored velocity coefficients do not match p_c.");
    accumulate(p, u, frames, mSensorPoints, 0);
  } else {
    uint64_t offset = 0;
    for (size_t c = 0; c < mCorners.size() / 6; ++c) {
      const uint64_t* q = &mCorners[6 * c];
      const uint64_t npts = (q[3] - q[0] + 1) * (q[4] - q[1] + 1) * (q[5] - q[2] + 1);
      const std::string ds = std::to_string(c + 1);
      const hid_t gp = mOutputFile.openGroup(root, "p_c");
      const std::vector<float> p = mOutputFile.readFloats(gp, ds);
      mOutputFile.closeGroup(gp);
      for (int k = 0; k < ncomp; ++k) {
        const hid_t gu = mOutputFile.openGroup(root, unames[k]);
        u[k] = mOutputFile.readFloats(gu, ds);
        mOutputFile.closeGroup(gu);
        if (u[k].size() != p.size()) throw std::ios::failure("Error: --post: the stored velocity coefficients do not match p_c.");
      }
      const uint64_t frames = p.size() / (2 * H * npts);
      if (!frames || p.size() != frames * 2 * H * npts) throw std::ios::failure("Error: --post: p_c does not match the sensor mask and --harmonics.");
      accumulate(p, u, frames, npts, offset);
      offset += npts;
    }
  }
}

// --post (KSpaceFirstOrderSolver.cpp:231-239, postProcessing :975-1030 without the time loop): the intensities and Q terms of an
// existing output file, from its stored raw series (--I_avg, --Q_term) and / or its stored coefficients (--I_avg_c, --Q_term_c)
void KSpaceFirstOrderSolver::postProcessOnly() {
  using K = OutputStream::Kind;
  const hid_t root = mOutputFile.root();
  const int ncomp = mScalars.nz > 1 ? 3 : 2;
  mPostProcessingTime.start();
  if (mCmd.iAvg || mCmd.qTerm) {  // the raw series the earlier run stored (rank 0 owns the file)
    for (auto& st : mStreams) {
      if (st.kind != K::kSeries || !this->root()) continue;
      if (mScalars.sensorMaskType == 0) {
        st.dataset = mOutputFile.openDataset(root, st.name);
      } else {
        st.group = mOutputFile.openGroup(root, st.name);
        for (size_t k = 0; k < mCorners.size() / 6; ++k) st.cuboidDatasets.push_back(mOutputFile.openDataset(st.group, std::to_string(k + 1)));
      }
    }
    // the number of stored steps comes from the file, not from -s / Nt of this command line
    const OutputStream* sp = nullptr;
    for (const auto& st : mStreams)
      if (st.id == KW_S_P_RAW) sp = &st;
    if (!sp) throw std::runtime_error("Error: --post: the raw pressure series is not part of this command line.");
    uint64_t stored = 0;
    if (this->root())
      stored = mScalars.sensorMaskType == 0 ? mOutputFile.elementCount(root, "p") / std::max<uint64_t>(mSensorPoints, 1)
                                            : mOutputFile.elementCount(sp->group, "1") /
                                                  std::max<uint64_t>((mCorners[3] - mCorners[0] + 1) * (mCorners[4] - mCorners[1] + 1) * (mCorners[5] - mCorners[2] + 1), 1);
    if (multi()) mTeam->bcast(&stored, sizeof stored);
    mSamplingSteps = stored;
    computeAverageIntensities();
  }
  if (mCmd.iAvgC || mCmd.qTermC) {
    std::vector<std::vector<float>> intensity(ncomp);
    if (this->root()) computeAverageIntensitiesC(intensity);
    const char* names[3] = {"Ix_avg_c", "Iy_avg_c", "Iz_avg_c"};
    if (mCmd.iAvgC && this->root())
      for (int k = 0; k < ncomp; ++k) replaceSensorValues(names[k], intensity[k].data());
    if (mCmd.qTermC) writeQTerm(intensity, "Q_term_c");
  }
  mPostProcessingTime.stop();
  mTotalTime.stop();
  if (this->root()) mOutputFile.close();
  if (multi()) mTeam->barrier();
  log(1, "Post-processing phase (--post): %s\n", formatSeconds(getPostProcessingTime()).c_str());
}

void KSpaceFirstOrderSolver::saveScalarsToOutputFile() {  // Parameters::saveScalarsToOutputFile (Parameters.cpp:559-650)
  const FileScalars& s = mScalars;
  const hid_t root = mOutputFile.root();
  Hdf5File& o = mOutputFile;
  o.writeScalar(root, "Nx", s.nx), o.writeScalar(root, "Ny", s.ny), o.writeScalar(root, "Nz", s.nz), o.writeScalar(root, "Nt", s.nt);
  const bool is3D = s.nz > 1;
  o.writeScalar(root, "dt", s.dt), o.writeScalar(root, "dx", s.dx), o.writeScalar(root, "dy", s.dy);
  if (is3D) o.writeScalar(root, "dz", s.dz);
  o.writeScalar(root, "c_ref", s.cRef);
  o.writeScalar(root, "pml_x_size", s.pmlXSize), o.writeScalar(root, "pml_y_size", s.pmlYSize);
  if (is3D) o.writeScalar(root, "pml_z_size", s.pmlZSize);
  o.writeScalar(root, "pml_x_alpha", s.pmlXAlpha), o.writeScalar(root, "pml_y_alpha", s.pmlYAlpha);
  if (is3D) o.writeScalar(root, "pml_z_alpha", s.pmlZAlpha);
  o.writeScalar(root, "p_source_flag", s.pSourceFlag), o.writeScalar(root, "p0_source_flag", s.p0SourceFlag);
  o.writeScalar(root, "transducer_source_flag", s.transducerSourceFlag);
  o.writeScalar(root, "ux_source_flag", s.uxSourceFlag), o.writeScalar(root, "uy_source_flag", s.uySourceFlag);
  if (is3D) o.writeScalar(root, "uz_source_flag", s.uzSourceFlag);
  o.writeScalar(root, "nonuniform_grid_flag", s.nonuniformGridFlag), o.writeScalar(root, "absorbing_flag", s.absorbingFlag);
  o.writeScalar(root, "nonlinear_flag", s.nonlinearFlag);
  if (s.uxSourceFlag || s.uySourceFlag || s.uzSourceFlag) o.writeScalar(root, "u_source_many", s.uSourceMany), o.writeScalar(root, "u_source_mode", s.uSourceMode);
  if (s.pSourceFlag) o.writeScalar(root, "p_source_many", s.pSourceMany), o.writeScalar(root, "p_source_mode", s.pSourceMode);
  if (s.absorbingFlag) o.writeScalar(root, "alpha_power", s.alphaPower);
  o.writeScalar(root, "t_index", (uint64_t)timeIndex());
  if (mCmd.copySensorMask && !o.exists(root, s.sensorMaskType == 0 ? "sensor_mask_index" : "sensor_mask_corners")) {  // cpp:1100-1116
    o.writeScalar(root, "sensor_mask_type", s.sensorMaskType);
    if (s.sensorMaskType == 0) {
      const auto idx = mInputFile.readIndices(mInputFile.root(), "sensor_mask_index");
      o.writeWhole(root, "sensor_mask_index", {1, 1, idx.size()}, {}, idx.data(), false, 0);
    } else {
      o.writeWhole(root, "sensor_mask_corners", {1, mCorners.size() / 6, 6}, {}, mCorners.data(), false, 0);
    }
  }
}

void KSpaceFirstOrderSolver::writeOutputHeader() {  // Hdf5FileHeader (Hdf5/Hdf5FileHeader.cpp:70-87); statistics cpp:1132-1168
  const hid_t root = mOutputFile.root();
  Hdf5File& o = mOutputFile;
  char host[256] = "unknown";
  gethostname(host, sizeof host - 1);
  char date[64];
  const time_t now = time(nullptr);
  strftime(date, sizeof date, "%d-%b-%Y-%H-%M-%S", localtime(&now));
  o.setStringAttribute(root, "/", "created_by", getCodeName());
  o.setStringAttribute(root, "/", "creation_date", date);
  o.setStringAttribute(root, "/", "file_description", "Output data created by the B200 time-step engine");
  o.setStringAttribute(root, "/", "file_type", "output");
  o.setStringAttribute(root, "/", "major_version", "1");
  o.setStringAttribute(root, "/", "minor_version", "1");
  o.setStringAttribute(root, "/", "host_names", host);
  o.setStringAttribute(root, "/", "number_of_cpu_cores", std::to_string(mCmd.numberOfThreads > 0 ? mCmd.numberOfThreads : (long)std::thread::hardware_concurrency()));
  o.setStringAttribute(root, "/", "total_memory_in_use", std::to_string(getHostMemoryUsage() >> 20) + " MB");
  o.setStringAttribute(root, "/", "peak_core_memory_in_use", std::to_string(getDeviceMemoryUsage() >> 20) + " MB");
  o.setStringAttribute(root, "/", "total_execution_time", formatSeconds(getTotalTime()));
  o.setStringAttribute(root, "/", "data_loading_phase_execution_time", formatSeconds(getDataLoadTime()));
  o.setStringAttribute(root, "/", "pre-processing_phase_execution_time", formatSeconds(getPreProcessingTime()));
  o.setStringAttribute(root, "/", "simulation_phase_execution_time", formatSeconds(getSimulationTime()));
  o.setStringAttribute(root, "/", "post-processing_phase_execution_time", formatSeconds(getPostProcessingTime()));
}

bool KSpaceFirstOrderSolver::isTimeToCheckpoint() const {  // Parameters::isTimeToCheckpoint (Parameters.cpp:683-692)
  if (!mCmd.isCheckpointEnabled()) return false;
  TimeMeasure t = mTotalTime;
  t.stop();
  return mStepsToCheckpoint == 0 || (mCmd.checkpointInterval > 0 && t.getElapsedTime() > (double)mCmd.checkpointInterval);
}

namespace {
const struct { const char* name; int id; } kCheckpointArrays[] = {  // the kCheckpoint records, Containers/MatrixContainer.cpp:101-115
    {"p", KW_P}, {"ux_sgx", KW_UX_SGX}, {"uy_sgy", KW_UY_SGY}, {"uz_sgz", KW_UZ_SGZ}, {"rhox", KW_RHOX}, {"rhoy", KW_RHOY}, {"rhoz", KW_RHOZ}};

// name of the reference's stream object behind a kw_stream id (OutputStreamContainer.cpp:70-325): the Temp_<name> datasets
std::string streamObjectName(int sid) {
  const char* axes[3] = {"x", "y", "z"};
  if (sid == KW_S_P_C) return "p_c";
  for (int k = 0; k < 3; ++k) {
    if (sid == KW_S_UX_C + k) return std::string("u") + axes[k] + "_c";
    if (sid == KW_S_UX_NS_C + k) return std::string("u") + axes[k] + "_non_staggered_c";
    if (sid == KW_S_IX_AVG_C + k) return std::string("I") + axes[k] + "_avg_c";
  }
  return "";
}
}  // namespace

// The reference's checkpoint (KSpaceFirstOrderSolver::saveCheckpointData cpp:1176-1224, OutputStreamContainer::checkpointStreams,
// BaseOutputStream.cpp:528-606): the checkpoint file holds the seven state arrays, t_index, Nx, Ny, Nz, the header and the
// compression accumulators Temp_<name>_1 / _2 (and Temp_<I?_avg_c>); the accumulators of the aggregate streams are flushed
// into their datasets of the OUTPUT file.  Either code can resume a run the other one interrupted.
void KSpaceFirstOrderSolver::saveCheckpointData() {
  using K = OutputStream::Kind;
  const FileScalars& s = mScalars;
  Hdf5File ck;  // rank 0 owns the checkpoint file; the other ranks hand it their slabs and sensor points
  const bool io = this->root();
  if (io) ck.create(mCmd.checkpointFile);  // overwrites the one of the previous leg
  const hid_t root = io ? ck.root() : -1;
  std::vector<float> field(s.nx * s.ny * mNzLocal);
  for (const auto& a : kCheckpointArrays) {
    if (s.nz == 1 && (a.id == KW_UZ_SGZ || a.id == KW_RHOZ)) continue;  // 2-D runs have no z components (MatrixContainer.cpp:101-115)
    field.resize(s.nx * s.ny * mNzLocal);
    check(kw_get_array(mCtx, a.id, field.data(), field.size()));
    if (multi()) field = gatherSlabs(field.data());
    if (io) ck.writeWhole(root, a.name, {s.nz, s.ny, s.nx}, {1, s.ny, s.nx}, field.data(), true, mCmd.compressionLevel);
  }
  if (io) {
    ck.writeScalar(root, "t_index", (uint64_t)timeIndex());
    ck.writeScalar(root, "Nx", s.nx), ck.writeScalar(root, "Ny", s.ny), ck.writeScalar(root, "Nz", s.nz);
  }
  if (timeIndex() > mCmd.samplingStartIndex) {  // cpp:1214-1220: nothing was sampled before
    std::vector<float> buf;
    for (auto& st : mStreams) {  // IndexOutputStream::checkpoint (:536-557): aggregates are flushed into the output file
      if ((st.kind != K::kAggregate && st.kind != K::kWholeDomain) || st.id == KW_S_Q_TERM_C) continue;
      uint64_t n = 0;
      check(kw_stream_buffer_get(mCtx, st.id, 0, nullptr, 0, &n));
      buf.resize(std::max<uint64_t>(n, 1));
      if (n) check(kw_stream_buffer_get(mCtx, st.id, 0, buf.data(), buf.size(), &n));
      if (multi()) buf = st.kind == K::kWholeDomain ? gatherSlabs(buf.data()) : gatherPoints(buf.data(), 1, n, st.perPoint, st.rowFloats);
      if (io) writeStreamBuffer(st, buf.data());
    }
    // IndexOutputStream::reopen / CuboidOutputStream::reopen read the attributes min, max, min_index, max_index of every raw / compressed
    // dataset (IndexOutputStream.cpp:244, BaseOutputStream.cpp:492-506) -- the reference's index streams never store them (the calls are
    // commented out, :505-509, :551-555), so it cannot resume its OWN index-mask runs; ours carry placeholders so that it can resume ours
    if (io)
      for (auto& st : mStreams) {
        if (st.kind != K::kSeries && st.kind != K::kCompressed) continue;
        auto mark = [&](hid_t loc, const std::string& name) {
          mOutputFile.setFloatAttribute(loc, name, "min", std::numeric_limits<float>::max());
          mOutputFile.setFloatAttribute(loc, name, "max", std::numeric_limits<float>::lowest());
          mOutputFile.setLongLongAttribute(loc, name, "min_index", 0);
          mOutputFile.setLongLongAttribute(loc, name, "max_index", 0);
        };
        if (st.dataset >= 0) mark(mOutputFile.root(), st.name);
        for (size_t k = 0; k < st.cuboidDatasets.size(); ++k) mark(st.group, std::to_string(k + 1));
      }
    for (int sid = 0; sid < KW_STREAM_COUNT; ++sid) {  // storeCheckpointCompressionCoefficients, do-not-save streams included
      const std::string name = streamObjectName(sid);
      uint64_t bytes = 0;
      if (name.empty() || kw_stream_state_size(mCtx, sid, &bytes) != KW_OK || !bytes) continue;
      const bool intensity = sid >= KW_S_IX_AVG_C && sid <= KW_S_IZ_AVG_C;
      for (int which = intensity ? 0 : 1; which <= (intensity ? 0 : 2); ++which) {
        uint64_t n = 0;
        check(kw_stream_buffer_get(mCtx, sid, which, nullptr, 0, &n));
        buf.resize(std::max<uint64_t>(n, 1));
        if (n) check(kw_stream_buffer_get(mCtx, sid, which, buf.data(), buf.size(), &n));
        if (multi()) {  // accumulators are per sensor point (2 * harmonics floats each, one for the intensities)
          const uint64_t per = intensity ? 1 : 2 * mCmd.harmonics;
          buf = gatherPoints(buf.data(), 1, n, per, mSensorPoints * per);
          n = mSensorPoints * per;
        }
        const std::string ds = "Temp_" + name + (intensity ? "" : which == 1 ? "_1" : "_2");
        if (io) ck.writeWhole(root, ds, {1, 1, n}, {1, 1, n}, buf.data(), true, mCmd.compressionLevel);
      }
    }
  }
  if (!io) return;
  char date[64];
  const time_t now = time(nullptr);
  strftime(date, sizeof date, "%d-%b-%Y-%H-%M-%S", localtime(&now));
  ck.setStringAttribute(root, "/", "created_by", getCodeName());
  ck.setStringAttribute(root, "/", "creation_date", date);
  ck.setStringAttribute(root, "/", "file_description", "Checkpoint data created by the B200 time-step engine");
  ck.setStringAttribute(root, "/", "file_type", "checkpoint");
  ck.setStringAttribute(root, "/", "major_version", "1");
  ck.setStringAttribute(root, "/", "minor_version", "1");
  ck.close();
}

void KSpaceFirstOrderSolver::recoverFromCheckpoint() {
  using K = OutputStream::Kind;
  const FileScalars& s = mScalars;
  Hdf5File ck;
  ck.open(mCmd.checkpointFile, true);
  const hid_t root = ck.root();
  if (ck.getStringAttribute(root, "/", "file_type") != "checkpoint") throw std::ios::failure("Error: \"" + mCmd.checkpointFile + "\" is not a checkpoint file.");
  if (ck.readIndexScalar(root, "Nx") != s.nx || ck.readIndexScalar(root, "Ny") != s.ny || ck.readIndexScalar(root, "Nz") != s.nz)
    throw std::ios::failure("Error: The checkpoint file was created for a different domain size (cpp:2846-2890).");
  for (const auto& a : kCheckpointArrays) {  // every rank reads its own slab
    if (s.nz == 1 && (a.id == KW_UZ_SGZ || a.id == KW_RHOZ)) continue;
    std::vector<float> v(s.nx * s.ny * mNzLocal);
    const hid_t d = ck.openDataset(root, a.name);
    ck.readHyperslab(d, {mZ0, 0, 0}, {mNzLocal, s.ny, s.nx}, v.data());
    ck.closeDataset(d);
    check(kw_set_array(mCtx, a.id, v.data(), v.size()));
  }
  const uint64_t t = ck.readIndexScalar(root, "t_index");
  check(kw_set_time_index(mCtx, t));  // also the sampled / compressed step counters of the streams (IndexOutputStream::reopen :203-213)
  const uint64_t sampled = t > mCmd.samplingStartIndex ? t - mCmd.samplingStartIndex : 0;
  uint64_t frameSteps = 0;
  if (mCmd.anyCompressed()) {
    const float period = mCmd.period > 0.f ? mCmd.period : 1.0f / (mCmd.frequency * s.dt);
    frameSteps = (uint64_t)(period * (float)mCmd.mos);
  }
  for (auto& st : mStreams) {
    if (st.kind == K::kSeries) st.rowsWritten = sampled;
    else if (st.kind == K::kCompressed) st.rowsWritten = frameSteps ? sampled / frameSteps : 0;
  }
  if (sampled > 0) {
    std::vector<float> buf;
    for (auto& st : mStreams) {  // aggregated quantities are reloaded from the output file (IndexOutputStream::reopen :215-230)
      if ((st.kind != K::kAggregate && st.kind != K::kWholeDomain) || st.id == KW_S_Q_TERM_C) continue;
      if (st.id >= KW_S_IX_AVG_C && st.id <= KW_S_IZ_AVG_C) continue;  // restored from Temp_<name> below
      uint64_t n = 0;
      check(kw_stream_buffer_get(mCtx, st.id, 0, nullptr, 0, &n));
      if (!multi()) {
        buf.resize(n);
        readStreamBuffer(st, buf.data());
      } else {  // rank 0 reads the complete buffer from the output file and hands every rank its part
        std::vector<float> full;
        if (this->root()) {
          full.resize(st.rowFloats);
          readStreamBuffer(st, full.data());
        }
        buf = distribute(full, st.kind == K::kWholeDomain, st.perPoint, st.rowFloats);
      }
      if (n) check(kw_stream_buffer_set(mCtx, st.id, 0, buf.data(), n));
    }
    for (int sid = 0; sid < KW_STREAM_COUNT; ++sid) {  // loadCheckpointCompressionCoefficients (BaseOutputStream.cpp:528-545)
      const std::string name = streamObjectName(sid);
      uint64_t bytes = 0;
      if (name.empty() || kw_stream_state_size(mCtx, sid, &bytes) != KW_OK || !bytes) continue;
      const bool intensity = sid >= KW_S_IX_AVG_C && sid <= KW_S_IZ_AVG_C;
      for (int which = intensity ? 0 : 1; which <= (intensity ? 0 : 2); ++which) {
        const std::string ds = "Temp_" + name + (intensity ? "" : which == 1 ? "_1" : "_2");
        if (!ck.exists(root, ds)) throw std::ios::failure("Error: The checkpoint file was created with different output flags (" + ds + " is missing).");
        std::vector<float> v = ck.readFloats(root, ds);
        if (multi()) {
          std::vector<float> mine;
          scatterPointsFrom(v, intensity ? 1 : 2 * mCmd.harmonics, &mine);
          v.swap(mine);
        }
        if (!v.empty()) check(kw_stream_buffer_set(mCtx, sid, which, v.data(), v.size()));
      }
    }
  }
  ck.close();
  log(1, "Recovered from the checkpoint at time step %llu\n", (unsigned long long)timeIndex());
}

uint64_t KSpaceFirstOrderSolver::timeIndex() const {
  uint64_t t = 0;
  if (mCtx) kw_time_index(mCtx, &t);
  return t;
}

size_t KSpaceFirstOrderSolver::getHostMemoryUsage() const { return mHostBytes + mRowBuffer.size() * sizeof(float); }
size_t KSpaceFirstOrderSolver::getAvailableDeviceMemory() const {
  size_t freeB = 0, total = 0;
  kw_device_memory(&freeB, &total);
  return freeB;
}
size_t KSpaceFirstOrderSolver::getDeviceMemoryUsage() const {
  size_t freeB = 0, total = 0;
  kw_device_memory(&freeB, &total);
  return total - freeB;
}

// ---------------------------------------------------------------------------------------------------------------------
void KSpaceFirstOrderSolver::compute() {
  // preProcessing (cpp:784-857)
  mPreProcessingTime.start();
  createStreams();
  check(kw_stream_async(mCtx, 1));  // double-buffered rows: full buffers travel to the host while the loop goes on (OutputStreams async path)
  check(kw_preprocess(mCtx));
  exchangeSensorLayout();
  createOutputDatasets();
  mPreProcessingTime.stop();
  log(1, "Pre-processing phase: %s, device memory in use: %zu MB\n", formatSeconds(getPreProcessingTime()).c_str(), getDeviceMemoryUsage() >> 20);

  if (mRecover) recoverFromCheckpoint();
  if (mCmd.post) {
    postProcessOnly();
    return;
  }

  // computeMainLoop (cpp:864-943): the library runs until Nt, until a device-side row buffer is full, or until it is time
  // to checkpoint (cpp:885)
  mSimulationTime.start();
  const uint64_t nt = mScalars.nt;
  uint64_t nextReport = 0;
  auto timeToCheckpoint = [&]() {  // the wall clock of rank 0 decides for the whole team (kw_run is collective)
    int yes = isTimeToCheckpoint() ? 1 : 0;
    if (multi()) mTeam->bcast(&yes, sizeof yes);
    return yes != 0;
  };
  while (timeIndex() < nt && !timeToCheckpoint()) {
    uint64_t done = 0;
    uint64_t chunk = std::max<uint64_t>(1, nt * (uint64_t)mCmd.progressInterval / 100);
    chunk = std::min<uint64_t>(chunk, nt - timeIndex());
    if (mCmd.checkpointTimeSteps) chunk = std::min<uint64_t>(chunk, mStepsToCheckpoint);
    const int status = kw_run(mCtx, chunk, &done, 1);
    mStepsToCheckpoint -= std::min<uint64_t>(done, mStepsToCheckpoint);  // Parameters::incrementTimeIndex (:698-702)
    if (status == KW_ERR_STREAM_FULL) {
      flushSeries(false);
      continue;
    }
    check(status);
    flushSeries(true);  // chunks that left the device while the loop ran: written without stalling it
    if (mCmd.verbose > 0 && timeIndex() >= nextReport) {
      log(2, "  %3llu%% done, step %llu of %llu\n", (unsigned long long)(100 * timeIndex() / nt), (unsigned long long)timeIndex(), (unsigned long long)nt);
      nextReport = timeIndex() + chunk;
    }
  }
  check(kw_synchronize(mCtx));
  flushSeries(false);
  mSimulationTime.stop();
  log(1, "Simulation phase: %s (%llu of %llu steps)\n", formatSeconds(getSimulationTime()).c_str(), (unsigned long long)timeIndex(), (unsigned long long)nt);

  if (timeIndex() < nt) {  // interrupted to checkpoint (cpp:373-398): store the state, keep the output file for the next leg
    mPostProcessingTime.start();
    saveCheckpointData();
    if (root()) saveScalarsToOutputFile();  // writeOutputDataInfo runs on every leg (cpp:1099-1168): a resuming run checks Nx, Ny, Nz of the output file
    mPostProcessingTime.stop();
    mTotalTime.stop();
    if (root()) writeOutputHeader();
    mOutputFile.close();
    if (multi()) mTeam->barrier();
    log(1, "Checkpoint created after %llu time steps; run the same command again to continue.\n", (unsigned long long)timeIndex());
    return;
  }

  // postProcessing (cpp:950-973) + writeOutputDataInfo (cpp:1099-1168)
  mPostProcessingTime.start();
  check(kw_finish(mCtx));
  writeAggregates();
  if (mCmd.iAvg || mCmd.qTerm) computeAverageIntensities();
  if (root()) saveScalarsToOutputFile();
  mPostProcessingTime.stop();
  mTotalTime.stop();
  if (root()) writeOutputHeader();
  mOutputFile.close();
  if (multi()) mTeam->barrier();  // nobody removes the checkpoint file while a rank may still read it
  if (root() && mCmd.isCheckpointEnabled()) std::remove(mCmd.checkpointFile.c_str());  // cpp:409-413
  log(1, "Post-processing phase: %s, total: %s\n", formatSeconds(getPostProcessingTime()).c_str(), formatSeconds(getTotalTime()).c_str());
}

}  // namespace kwhost
