// C++ host of the B200 time-step engine: the public interface of the reference's solver class
// (KSpaceSolver/KSpaceFirstOrderSolver.h:56-227 -- same method names, call order and exception types), implemented over
// the C ABI of include/kwave_b200.h.  The class owns the kw_ctx (device arrays, streams) and the input / output files; all
// device work happens inside libkwave_b200.so.
//
//   KSpaceFirstOrderSolver solver(commandLine);
//   solver.allocateMemory();   // reads the scalars, creates the device context            (cpp:124-151)
//   solver.loadInputData();    // input datasets -> kw_set_array, output file created         (cpp:157-240)
//   solver.compute();          // pre-processing, time loop with in-step sampling, outputs    (cpp:246-439)
//
// Errors: std::bad_alloc (device / host memory), std::ios::failure (files), std::runtime_error (CUDA),
// std::invalid_argument (unsupported input) -- as thrown by the reference and turned into error exits by main().
#pragma once
#include <chrono>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/kwave_b200.h"
#include "Hdf5Io.h"
#include "Parameters.h"
#include "Team.h"

namespace kwhost {

class TimeMeasure {  // Utils/TimeMeasure.h
 public:
  void start() { mStart = std::chrono::steady_clock::now(); }
  void stop() { mElapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - mStart).count(); }
  double getElapsedTime() const { return mElapsed; }
  double getElapsedTimeOverAllLegs() const { return mElapsed; }  // no checkpoint legs in this build

 private:
  std::chrono::steady_clock::time_point mStart;
  double mElapsed = 0.0;
};

class KSpaceFirstOrderSolver {
 public:
  // `team`: the processes of a slab-decomposed run (--gpus N, host/Team.h); nullptr = one GPU
  explicit KSpaceFirstOrderSolver(const CommandLine& commandLine, const Team* team = nullptr);
  KSpaceFirstOrderSolver(const KSpaceFirstOrderSolver&) = delete;
  virtual ~KSpaceFirstOrderSolver();
  KSpaceFirstOrderSolver& operator=(const KSpaceFirstOrderSolver&) = delete;

  virtual void allocateMemory();
  virtual void freeMemory();
  virtual void loadInputData();
  virtual void compute();

  size_t getHostMemoryUsage() const;
  size_t getDeviceMemoryUsage() const;
  size_t getAvailableDeviceMemory() const;
  const std::string getCodeName() const { return "kspaceFirstOrder-B200 v1.0 (hot path of kspaceFirstOrder-CUDA v1.3)"; }
  void printFullCodeNameAndLicense() const;

  double getTotalTime() const { return mTotalTime.getElapsedTime(); }
  double getPreProcessingTime() const { return mPreProcessingTime.getElapsedTime(); }
  double getDataLoadTime() const { return mDataLoadTime.getElapsedTime(); }
  double getSimulationTime() const { return mSimulationTime.getElapsedTime(); }
  double getPostProcessingTime() const { return mPostProcessingTime.getElapsedTime(); }
  double getCumulatedTotalTime() const { return mTotalTime.getElapsedTimeOverAllLegs(); }
  double getCumulatedPreProcessingTime() const { return mPreProcessingTime.getElapsedTimeOverAllLegs(); }
  double getCumulatedDataLoadTime() const { return mDataLoadTime.getElapsedTimeOverAllLegs(); }
  double getCumulatedSimulationTime() const { return mSimulationTime.getElapsedTimeOverAllLegs(); }
  double getCumulatedPostProcessingTime() const { return mPostProcessingTime.getElapsedTimeOverAllLegs(); }

  const FileScalars& scalars() const { return mScalars; }
  uint64_t timeIndex() const;

 protected:
  // one output stream of the container (OutputStreamContainer.cpp:70-325): which kw_stream feeds which dataset
  struct OutputStream {
    int id;                // kw_stream
    std::string name;      // dataset (index mask) or group (cuboid mask) name, Utils/MatrixNames.h
    enum Kind { kSeries, kCompressed, kAggregate, kWholeDomain } kind;
    bool shifted = false;  // compression attributes c_shift / c_max_exp
    hid_t dataset = -1;    // index mask: the dataset; cuboids: unused
    hid_t group = -1;
    std::vector<hid_t> cuboidDatasets;
    uint64_t rowsWritten = 0;
    uint64_t rowFloats = 0;       // floats of one row in the FILE (all sensor points)
    uint64_t localFloats = 0;     // floats of one row held by this rank (== rowFloats on one GPU)
    uint64_t perPoint = 1;        // floats per sensor point (2 * harmonics for compressed series)
  };
  // slab-decomposed runs: what this rank holds of a per-sensor / whole-grid buffer -> the complete buffer on rank 0
  std::vector<float> gatherPoints(const float* local, uint64_t rows, uint64_t localFloats, uint64_t perPoint, uint64_t fullFloats);
  std::vector<float> gatherSlabs(const float* local);
  std::vector<float> distribute(const std::vector<float>& full, bool wholeDomain, uint64_t perPoint, uint64_t fullFloats);
  void scatterPointsFrom(const std::vector<float>& full, uint64_t perPoint, std::vector<float>* local) const;
  bool root() const { return !mTeam || mTeam->root(); }
  bool multi() const { return mTeam && mTeam->multi(); }
  void check(int status) const;  // C-ABI status -> the reference's exception types
  void readScalars();
  void loadArray(const std::string& name, int arrayId, bool isIndex, bool required);
  void createStreams();
  void exchangeSensorLayout();
  void postProcessOnly();             // --post (cpp:231-239, :975-1030)
  void computeAverageIntensitiesC(std::vector<std::vector<float>>& intensity);  // cpp:1543-1775
  void replaceSensorValues(const std::string& name, const float* data);
  void writeQTerm(const std::vector<std::vector<float>>& intensity, const char* name);  // computeQTerm (cpp:1783-2080), collective on slabs
  void writeStreamBuffer(OutputStream& st, const float* buf);  // accumulator of an aggregate stream -> output file
  void readStreamBuffer(OutputStream& st, float* buf);
  void createOutputDatasets();
  bool isTimeToCheckpoint() const;   // Parameters::isTimeToCheckpoint (Parameters.cpp:683-692)
  void saveCheckpointData();         // cpp:1176-1224
  void recoverFromCheckpoint();      // cpp:186-228 (state) + OutputStreamContainer::reopenStreams
  void flushSeries(bool final);
  void writeAggregates();
  void computeAverageIntensities();  // cpp:1231-1534 (+ computeQTerm :1783-2080): --I_avg / --Q_term from the stored series
  void writeSensorValues(const std::string& name, const float* data);
  void writeOutputHeader();
  void saveScalarsToOutputFile();
  void log(int level, const char* fmt, ...) const;

  CommandLine mCmd;
  const Team* mTeam = nullptr;
  uint64_t mZ0 = 0, mNzLocal = 0;                   // this rank's planes
  std::vector<uint64_t> mPositions;                  // positions of this rank's sensor points in the complete row
  std::vector<std::vector<uint64_t>> mRankPositions;  // rank 0: the same for every rank
  uint64_t mLocalPoints = 0;
  FileScalars mScalars;
  Hdf5File mInputFile, mOutputFile;
  kw_ctx* mCtx = nullptr;
  std::vector<OutputStream> mStreams;
  std::vector<uint64_t> mCorners;  // sensor_mask_corners (1-based, 6 per cuboid) for the output layout
  uint64_t mSensorPoints = 0;
  uint64_t mSamplingSteps = 0;
  uint64_t mCompressedSteps = 0;
  uint64_t mRowsCapacity = 0;
  std::vector<float> mRowBuffer;
  bool mRecover = false;             // a checkpoint file exists: continue that run
  uint64_t mStepsToCheckpoint = 0;   // Parameters::mTimeStepsToCheckpoint
  double mElapsedBefore[5] = {};     // total, load, pre-processing, simulation, post-processing of the previous legs
  size_t mHostBytes = 0;
  TimeMeasure mTotalTime, mPreProcessingTime, mDataLoadTime, mSimulationTime, mPostProcessingTime;
};

}  // namespace kwhost
