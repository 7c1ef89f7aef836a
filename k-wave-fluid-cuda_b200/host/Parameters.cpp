#include "Parameters.h"

#include <getopt.h>

#include <stdexcept>
#include <string>

namespace kwhost {

std::string CommandLine::usage() {
  return "Usage: kspaceFirstOrder-B200 -i <input_file> -o <output_file> [options]\n"
         "  -i, -o                      input / output file (k-Wave HDF5 layout)\n"
         "  -t <n>  -g <id>  -r <pct>  -c <0-9>  -s <step>  --benchmark <steps>  --verbose <0-2>  --copy_sensor_mask\n"
         "  -p|--p_raw --p_c --p_rms --p_max --p_min --p_max_all --p_min_all --p_final\n"
         "  -u|--u_raw --u_c --u_non_staggered_raw --u_non_staggered_c --u_rms --u_max --u_min --u_max_all --u_min_all --u_final\n"
         "  --I_avg  --Q_term  --block_size <n>  (from the stored raw series)\n"
         "  --I_avg_c  --Q_term_c  --period <steps> | --frequency <Hz>  --mos <n>  --harmonics <n>  --no_overlap  --40-bit_complex\n"
         "  --checkpoint_file <file> with --checkpoint_interval <seconds> and/or --checkpoint_timesteps <steps>\n"
         "  --post                      only post-process an existing output file (with --I_avg / --Q_term / --I_avg_c / --Q_term_c)\n"
         "  --gpus <1|2|4|8>            slab-decompose the grid along z over that many GPUs (one process per GPU, devices -g .. -g + N - 1)\n"
         "  -h|--help  --version\n"
         "";
}

static long toLong(const char* s, const char* what, long minValue) {
  try {
    size_t pos = 0;
    const long v = std::stol(s, &pos);
    if (pos != std::string(s).size() || v < minValue) throw std::invalid_argument(what);
    return v;
  } catch (...) {
    throw std::invalid_argument(std::string("Error: Invalid value of ") + what + ".");
  }
}
static float toFloat(const char* s, const char* what, float minValue) {
  try {
    const float v = std::stof(s);
    if (v < minValue) throw std::invalid_argument(what);
    return v;
  } catch (...) {
    throw std::invalid_argument(std::string("Error: Invalid value of ") + what + ".");
  }
}

void CommandLine::parse(int argc, char** argv) {
  // same short options and long-option codes as the reference (CommandLineParameters.cpp:264-292)
  const char* shortOpts = "i:o:r:c:t:g:puhs:";
  const struct option longOpts[] = {{"benchmark", required_argument, nullptr, 1}, {"copy_sensor_mask", no_argument, nullptr, 2},
    {"checkpoint_file", required_argument, nullptr, 3}, {"checkpoint_interval", required_argument, nullptr, 4},
    {"checkpoint_timesteps", required_argument, nullptr, 5}, {"help", no_argument, nullptr, 'h'}, {"verbose", required_argument, nullptr, 6},
    {"version", no_argument, nullptr, 7}, {"p_raw", no_argument, nullptr, 'p'}, {"p_c", no_argument, nullptr, 9}, {"p_rms", no_argument, nullptr, 10},
    {"p_max", no_argument, nullptr, 11}, {"p_min", no_argument, nullptr, 12}, {"p_max_all", no_argument, nullptr, 13},
    {"p_min_all", no_argument, nullptr, 14}, {"p_final", no_argument, nullptr, 15}, {"frequency", required_argument, nullptr, 16},
    {"period", required_argument, nullptr, 17}, {"mos", required_argument, nullptr, 18}, {"harmonics", required_argument, nullptr, 19},
    {"u_raw", no_argument, nullptr, 'u'}, {"u_rms", no_argument, nullptr, 20}, {"u_max", no_argument, nullptr, 21},
    {"u_min", no_argument, nullptr, 22}, {"u_max_all", no_argument, nullptr, 23}, {"u_min_all", no_argument, nullptr, 24},
    {"u_final", no_argument, nullptr, 25}, {"u_non_staggered_raw", no_argument, nullptr, 26}, {"u_c", no_argument, nullptr, 27},
    {"u_non_staggered_c", no_argument, nullptr, 28}, {"I_avg", no_argument, nullptr, 29}, {"I_avg_c", no_argument, nullptr, 30},
    {"Q_term", no_argument, nullptr, 31}, {"Q_term_c", no_argument, nullptr, 32}, {"post", no_argument, nullptr, 33},
    {"block_size", required_argument, nullptr, 34}, {"no_overlap", no_argument, nullptr, 35}, {"40-bit_complex", no_argument, nullptr, 36},
    {"gpus", required_argument, nullptr, 37},
    {nullptr, no_argument, nullptr, 0}};  // clang-format on
  optind = 1;
  int opt, idx = -1;
  while ((opt = getopt_long(argc, argv, shortOpts, longOpts, &idx)) != -1) {
    switch (opt) {
      case 'i': inputFile = optarg; break;
      case 'o': outputFile = optarg; break;
      case 'r': progressInterval = toLong(optarg, "-r (1-100)", 1); if (progressInterval > 100) throw std::invalid_argument("Error: Invalid value of -r (1-100)."); break;
      case 'c': compressionLevel = (unsigned)toLong(optarg, "-c (0-9)", 0); if (compressionLevel > 9) throw std::invalid_argument("Error: Invalid value of -c (0-9)."); break;
      case 't': numberOfThreads = toLong(optarg, "-t", 1); break;
      case 'g': gpuDevice = (int)toLong(optarg, "-g", 0); break;
      case 'p': pRaw = true; break;
      case 'u': uRaw = true; break;
      case 'h': printHelp = true; break;
      case 's': samplingStartIndex = (uint64_t)toLong(optarg, "-s (must be >= 1)", 1) - 1; break;  // 1-based on the command line
      case 1: benchmark = true; benchmarkSteps = (uint64_t)toLong(optarg, "--benchmark", 1); break;
      case 2: copySensorMask = true; break;
      case 3: checkpointFile = optarg; break;
      case 4: checkpointInterval = (uint64_t)toLong(optarg, "--checkpoint_interval", 1); break;
      case 5: checkpointTimeSteps = (uint64_t)toLong(optarg, "--checkpoint_timesteps", 1); break;
      case 6: verbose = (int)toLong(optarg, "--verbose (0-2)", 0); if (verbose > 2) throw std::invalid_argument("Error: Invalid value of --verbose (0-2)."); break;
      case 7: printVersion = true; break;
      case 9: pC = true; break;
      case 10: pRms = true; break;
      case 11: pMax = true; break;
      case 12: pMin = true; break;
      case 13: pMaxAll = true; break;
      case 14: pMinAll = true; break;
      case 15: pFinal = true; break;
      case 16: frequency = toFloat(optarg, "--frequency", 1.0f); break;
      case 17: period = toFloat(optarg, "--period", 1.0f); break;
      case 18: mos = (uint64_t)toLong(optarg, "--mos", 1); break;
      case 19: harmonics = (uint64_t)toLong(optarg, "--harmonics", 1); break;
      case 20: uRms = true; break;
      case 21: uMax = true; break;
      case 22: uMin = true; break;
      case 23: uMaxAll = true; break;
      case 24: uMinAll = true; break;
      case 25: uFinal = true; break;
      case 26: uNonStaggeredRaw = true; break;
      case 27: uC = true; break;
      case 28: uNonStaggeredC = true; break;
      case 30: iAvgC = true; break;
      case 32: qTermC = true; break;
      case 29: iAvg = true; break;
      case 31: qTerm = true; break;
      case 33: post = true; break;
      case 34: blockSize = (uint64_t)toLong(optarg, "--block_size", 1); break;
      case 35: noOverlap = true; break;
      case 36: c40bit = true; break;
      case 37: gpus = (int)toLong(optarg, "--gpus (1, 2, 4, 8)", 1); if (gpus != 1 && gpus != 2 && gpus != 4 && gpus != 8) throw std::invalid_argument("Error: Invalid value of --gpus (1, 2, 4, 8)."); break;
      default: throw std::invalid_argument("Error: Unknown command line switch or missing argument.");
    }
  }
  if (printHelp || printVersion) return;
  // validation (CommandLineParameters.cpp:888-947)
  if (inputFile.empty()) throw std::invalid_argument("Error: Input file was not specified.");
  if (outputFile.empty()) throw std::invalid_argument("Error: Output file was not specified.");
  if (isCheckpointEnabled() && checkpointFile.empty()) throw std::invalid_argument("Error: Checkpoint file was not specified.");  // :905-913
  if (!checkpointFile.empty() && !isCheckpointEnabled())
    throw std::invalid_argument("Error: Checkpoint interval or the number of time steps to checkpoint was not specified.");
  if (anyCompressed() && period == 0.f && frequency == 0.f)
    throw std::invalid_argument("Error: Compression (--p_c, --u_c, --u_non_staggered_c, --I_avg_c, --Q_term_c) needs --period or --frequency.");
  // --I_avg / --Q_term are computed from the stored raw series of p and the non-staggered velocity, which are therefore
  // stored too (OutputStreamContainer.cpp:229-262)
  if (iAvg || qTerm) pRaw = uNonStaggeredRaw = true;
  // --post: no simulation, only the post-processing of an existing output file (KSpaceFirstOrderSolver.cpp:231-239, :990-994)
  if (post && !(iAvg || qTerm || iAvgC || qTermC))
    throw std::invalid_argument("Error: --post needs at least one of --I_avg, --I_avg_c, --Q_term, --Q_term_c.");
  if (post && isCheckpointEnabled()) throw std::invalid_argument("Error: --post cannot be combined with checkpointing.");
  // nothing selected: this fork of the reference stores nothing (CommandLineParameters.cpp:938-947 sets
  // mStorePressureRawFlag = false where upstream k-Wave defaults to --p_raw); the scalars and the header are still written
}

}  // namespace kwhost
