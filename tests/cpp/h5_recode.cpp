// CPU test tool: reads the named datasets of an HDF5 file through the HDF5 API subset of csrc/minih5 and writes them into a new
// file as chunked (one chunk per index of the slowest dimension, clipped second dimension) + deflate datasets, copying the
// k-Wave attributes.  tests/test_hdf5_cpu.py compares both directions against the independent Python implementation.
//   usage: h5_recode in.h5 out.h5 /name1 /group/name2 ...
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "hdf5_hl.h"

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const hid_t fin = H5Fopen(argv[1], H5F_ACC_RDONLY, H5P_DEFAULT);
  if (fin < 0) {
    fprintf(stderr, "cannot open %s\n", argv[1]);
    return 1;
  }
  const hid_t fout = H5Fcreate(argv[2], H5F_ACC_TRUNC, H5P_DEFAULT, H5P_DEFAULT);
  if (fout < 0) return 1;
  char buf[256];
  for (const char* a : {"file_type", "major_version", "created_by"})
    if (H5LTget_attribute_string(fin, "/", a, buf) >= 0) H5LTset_attribute_string(fout, "/", a, buf);
  for (int i = 3; i < argc; ++i) {
    const std::string name = argv[i];
    int rank = 0;
    if (H5LTget_dataset_ndims(fin, name.c_str(), &rank) < 0 || rank < 1 || rank > 4) {
      fprintf(stderr, "no dataset %s\n", name.c_str());
      return 1;
    }
    hsize_t dims[4] = {1, 1, 1, 1};
    H5T_class_t cls;
    size_t tsize = 0;
    if (H5LTget_dataset_info(fin, name.c_str(), dims, &cls, &tsize) < 0) return 1;
    size_t n = 1;
    for (int d = 0; d < rank; ++d) n *= dims[d];
    const bool is_float = tsize == 4;
    std::vector<unsigned char> data(n * tsize);
    if (H5LTread_dataset(fin, name.c_str(), is_float ? H5T_NATIVE_FLOAT : H5T_STD_U64LE, data.data()) < 0) return 1;
    // parent groups
    for (size_t p = name.find('/', 1); p != std::string::npos; p = name.find('/', p + 1)) {
      const std::string g = name.substr(0, p);
      if (H5Lexists(fout, g.c_str(), H5P_DEFAULT) <= 0) H5Gclose(H5Gcreate(fout, g.c_str(), H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT));
    }
    hsize_t chunk[4];
    for (int d = 0; d < rank; ++d) chunk[d] = dims[d];
    if (rank > 1) chunk[0] = 1;
    if (rank > 2 && dims[1] > 3) chunk[1] = 3;  // clipped edge chunks
    const hid_t space = H5Screate_simple(rank, dims, nullptr);
    const hid_t pl = H5Pcreate(H5P_DATASET_CREATE);
    H5Pset_chunk(pl, rank, chunk);
    H5Pset_deflate(pl, 4);
    const hid_t ds = H5Dcreate(fout, name.c_str(), is_float ? H5T_NATIVE_FLOAT : H5T_STD_U64LE, space, H5P_DEFAULT, pl, H5P_DEFAULT);
    if (ds < 0 || H5Dwrite(ds, is_float ? H5T_NATIVE_FLOAT : H5T_STD_U64LE, H5S_ALL, H5S_ALL, H5P_DEFAULT, data.data()) < 0) return 1;
    H5Dclose(ds), H5Pclose(pl), H5Sclose(space);
    for (const char* a : {"data_type", "domain_type"})
      if (H5LTget_attribute_string(fin, name.c_str(), a, buf) >= 0) H5LTset_attribute_string(fout, name.c_str(), a, buf);
    long long ll;
    if (H5LTget_attribute_long_long(fin, name.c_str(), "c_harmonics", &ll) >= 0) H5LTset_attribute_long_long(fout, name.c_str(), "c_harmonics", &ll, 1);
    float fl;
    if (H5LTget_attribute_float(fin, name.c_str(), "c_period", &fl) >= 0) H5LTset_attribute_float(fout, name.c_str(), "c_period", &fl, 1);
  }
  H5Fclose(fin);
  return H5Fclose(fout) < 0 ? 1 : 0;
}
