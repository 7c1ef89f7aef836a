// kspaceFirstOrder-B200: command-line front end with the flags, files and exit behaviour of kspaceFirstOrder-CUDA
// (main.cpp of the reference: parse -> allocateMemory -> loadInputData -> compute; any error prints a boxed message to
// stderr and exits with EXIT_FAILURE, Logger/Logger.cpp:82-89).
extern "C" void omp_set_num_threads(int);  // libgomp (the engine library links it); -t bounds its host loops

#include <cstdio>
#include <cstdlib>
#include <unistd.h>
#include <exception>
#include <new>
#include <stdexcept>

#include "KSpaceFirstOrderSolver.h"

static void errorAndTerminate(const std::string& message) {
  fprintf(stderr, "+---------------------------------------------------------------+\n");
  fprintf(stderr, "|            !!! K-Wave experienced a fatal error !!!           |\n");
  fprintf(stderr, "+---------------------------------------------------------------+\n");
  fprintf(stderr, "%s\n", message.c_str());
  fprintf(stderr, "+---------------------------------------------------------------+\n");
  fprintf(stderr, "|                      Execution terminated                     |\n");
  fprintf(stderr, "+---------------------------------------------------------------+\n");
  exit(EXIT_FAILURE);
}

int main(int argc, char** argv) {
  kwhost::CommandLine cmd;
  try {
    cmd.parse(argc, argv);
  } catch (const std::exception& e) {
    fputs(kwhost::CommandLine::usage().c_str(), stderr);
    errorAndTerminate(e.what());
  }
  if (cmd.printHelp) {
    fputs(kwhost::CommandLine::usage().c_str(), stdout);
    return EXIT_SUCCESS;
  }
  if (cmd.numberOfThreads > 0) omp_set_num_threads((int)cmd.numberOfThreads);
  // --gpus N: one process per GPU, forked BEFORE anything touches CUDA; rank 0 (this process) owns the files
  kwhost::Team team;
  if (cmd.gpus > 1 && !cmd.printVersion) {
    try {
      team.spawn(cmd.gpus);
    } catch (const std::exception& e) {
      errorAndTerminate(e.what());
    }
  }
  kwhost::KSpaceFirstOrderSolver solver(cmd, &team);
  if (cmd.printVersion) {
    solver.printFullCodeNameAndLicense();
    return EXIT_SUCCESS;
  }
  if (cmd.verbose > 0 && team.root()) solver.printFullCodeNameAndLicense();
  try {
    solver.allocateMemory();
    solver.loadInputData();
  } catch (const std::bad_alloc&) {
    errorAndTerminate("Error: Not enough CPU or GPU memory to run this simulation.");
  } catch (const std::exception& e) {
    errorAndTerminate(e.what());
  }
  try {
    solver.compute();
  } catch (const std::bad_alloc&) {
    errorAndTerminate("Error: Not enough CPU or GPU memory to run this simulation.");
  } catch (const std::exception& e) {
    errorAndTerminate(e.what());
  }
  if (!team.root()) _exit(EXIT_SUCCESS);  // workers: done (no static destructors of the parent's state)
  if (!team.join()) errorAndTerminate("Error: a worker process of the slab-decomposed run failed.");
  if (cmd.verbose >= 0)
    printf("Total execution time: %.2fs (load %.2fs, pre-processing %.2fs, simulation %.2fs, post-processing %.2fs)\n", solver.getTotalTime(),
           solver.getDataLoadTime(), solver.getPreProcessingTime(), solver.getSimulationTime(), solver.getPostProcessingTime());
  return EXIT_SUCCESS;
}
