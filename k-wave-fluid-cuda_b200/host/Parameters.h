// Command-line flags and input-file scalars of the solver.  Flag names, meaning and validation follow the reference
// (Parameters/CommandLineParameters.cpp:264-292, :888-964; Parameters/Parameters.cpp:111-163, :194-553): a run
// scripted for kspaceFirstOrder-CUDA is accepted unchanged.  Flags whose feature lies outside the time-step hot path
// are still parsed.
#pragma once
#include <cstdint>
#include <string>

namespace kwhost {

struct CommandLine {
  std::string inputFile, outputFile, checkpointFile;
  long numberOfThreads = 0;   // -t (host threads: file I/O and pre-processing only)
  int gpuDevice = -1;         // -g
  int gpus = 1;               // --gpus: slab decomposition over N GPUs (no counterpart in the single-GPU reference)
  long progressInterval = 5;  // -r
  unsigned compressionLevel = 0;  // -c
  uint64_t checkpointInterval = 0;   // --checkpoint_interval <seconds>
  uint64_t checkpointTimeSteps = 0;  // --checkpoint_timesteps <steps>
  bool isCheckpointEnabled() const { return checkpointInterval > 0 || checkpointTimeSteps > 0; }  // CommandLineParameters.h:348
  bool benchmark = false;
  uint64_t benchmarkSteps = 0;
  uint64_t samplingStartIndex = 0;  // -s, stored 0-based (CommandLineParameters.cpp:424)
  int verbose = 0;
  bool copySensorMask = false;
  bool post = false;  // --post: only post-process an existing output file (I_avg / Q_term from stored series or coefficients)
  bool printVersion = false, printHelp = false;
  // outputs
  bool pRaw = false, pC = false, pRms = false, pMax = false, pMin = false, pMaxAll = false, pMinAll = false, pFinal = false;
  bool uRaw = false, uC = false, uRms = false, uMax = false, uMin = false, uMaxAll = false, uMinAll = false, uFinal = false;
  bool uNonStaggeredRaw = false, uNonStaggeredC = false, iAvgC = false, qTermC = false, iAvg = false, qTerm = false;
  // compression
  float frequency = 0.f, period = 0.f;
  uint64_t mos = 1, harmonics = 1;
  bool noOverlap = false, c40bit = false;
  uint64_t blockSize = 0;

  // throws std::invalid_argument with the message to print (the caller prints usage and exits with EXIT_FAILURE)
  void parse(int argc, char** argv);
  bool anyCompressed() const { return pC || uC || uNonStaggeredC || iAvgC || qTermC; }
  static std::string usage();
};

struct FileScalars {
  uint64_t nx = 0, ny = 0, nz = 0, nt = 0;
  float dt = 0, dx = 0, dy = 0, dz = 0, cRef = 0, alphaPower = 0;
  uint64_t pmlXSize = 0, pmlYSize = 0, pmlZSize = 0;
  float pmlXAlpha = 0, pmlYAlpha = 0, pmlZAlpha = 0;
  uint64_t sensorMaskType = 0;
  uint64_t pSourceFlag = 0, p0SourceFlag = 0, transducerSourceFlag = 0, uxSourceFlag = 0, uySourceFlag = 0, uzSourceFlag = 0;
  uint64_t nonuniformGridFlag = 0, absorbingFlag = 0, nonlinearFlag = 0;
  uint64_t uSourceMany = 0, uSourceMode = 0, pSourceMany = 0, pSourceMode = 0;
};

}  // namespace kwhost
