// CPU test of the host's file layer (host/Hdf5Io.h over csrc/minih5): what checkpoint / restart and the post-processing of
// stored series rely on -- datasets created, written row by row, the file closed, reopened read-write, the dataset
// continued, hyperslabs read back; groups with per-cuboid datasets; scalars and attributes.
#include <cstdio>
#include <vector>

#include "Hdf5Io.h"

using kwhost::Hdf5File;

#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) {                                                      \
      fprintf(stderr, "FAILED %s (line %d)\n", #cond, __LINE__);        \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main(int argc, char** argv) {
  const std::string path = argc > 1 ? argv[1] : "/tmp/hdf5io_roundtrip.h5";
  const hsize_t rows = 6, n = 4;
  {
    Hdf5File f;
    f.create(path);
    const hid_t d = f.createDataset(f.root(), "p", {1, rows, n}, {1, 1, n}, true, 0);
    for (hsize_t r = 0; r < 3; ++r) {
      std::vector<float> row(n);
      for (hsize_t i = 0; i < n; ++i) row[i] = 10.f * r + i;
      f.writeHyperslab(d, {0, r, 0}, {1, 1, n}, row.data());
    }
    f.closeDataset(d);
    const hid_t g = f.createGroup(f.root(), "p_max");
    std::vector<float> cub(2 * 3 * 2, 7.f);
    f.writeWhole(g, "1", {2, 3, 2}, {2, 3, 2}, cub.data(), true, 0);
    f.closeGroup(g);
    f.writeScalar(f.root(), "t_index", (uint64_t)3);
    f.setStringAttribute(f.root(), "/", "file_type", "output");
    f.close();
  }
  CHECK(Hdf5File::canAccess(path));
  {
    Hdf5File f;
    f.open(path, false);  // read-write, as a restarted run does
    CHECK(f.getStringAttribute(f.root(), "/", "file_type") == "output");
    CHECK(f.readIndexScalar(f.root(), "t_index") == 3);
    const hid_t d = f.openDataset(f.root(), "p");
    for (hsize_t r = 3; r < rows; ++r) {
      std::vector<float> row(n);
      for (hsize_t i = 0; i < n; ++i) row[i] = 10.f * r + i;
      f.writeHyperslab(d, {0, r, 0}, {1, 1, n}, row.data());
    }
    // all steps of the points 1..2 (the block read of computeAverageIntensities)
    std::vector<float> block(rows * 2);
    f.readHyperslab(d, {0, 0, 1}, {1, rows, 2}, block.data());
    for (hsize_t r = 0; r < rows; ++r)
      for (hsize_t i = 0; i < 2; ++i) CHECK(block[r * 2 + i] == 10.f * r + (i + 1));
    f.closeDataset(d);
    const hid_t g = f.openGroup(f.root(), "p_max");
    CHECK(f.elementCount(g, "1") == 12);
    f.closeGroup(g);
    f.close();
  }
  {
    Hdf5File f;
    f.open(path, true);
    const auto all = f.readFloats(f.root(), "p");
    CHECK(all.size() == rows * n);
    for (hsize_t r = 0; r < rows; ++r)
      for (hsize_t i = 0; i < n; ++i) CHECK(all[r * n + i] == 10.f * r + i);
    CHECK(f.exists(f.root(), "p") && !f.exists(f.root(), "missing"));
  }
  printf("hdf5io round trip ok\n");
  return 0;
}
