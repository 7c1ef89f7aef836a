// Thin RAII layer over the HDF5 C API calls the k-Wave file format needs (the same calls the reference's Hdf5File makes,
// Hdf5/Hdf5File.cpp:97-1086).  Builds against the real hdf5.h / hdf5_hl.h, or against csrc/minih5 where no HDF5 library
// exists (this image).  Errors surface as std::ios::failure, as in the reference.
#pragma once
#include <hdf5.h>
#include <hdf5_hl.h>

#include <unistd.h>

#include <cstdint>
#include <ios>
#include <string>
#include <vector>

namespace kwhost {

class Hdf5File {
 public:
  Hdf5File() = default;
  ~Hdf5File() { close(); }
  Hdf5File(const Hdf5File&) = delete;
  Hdf5File& operator=(const Hdf5File&) = delete;

  void open(const std::string& name, bool readOnly = true) {
    mName = name;
    mFile = H5Fopen(name.c_str(), readOnly ? H5F_ACC_RDONLY : H5F_ACC_RDWR, H5P_DEFAULT);
    if (mFile < 0) throw std::ios::failure("Error: File \"" + name + "\" could not be opened.");
  }
  void create(const std::string& name) {
    mName = name;
    mFile = H5Fcreate(name.c_str(), H5F_ACC_TRUNC, H5P_DEFAULT, H5P_DEFAULT);
    if (mFile < 0) throw std::ios::failure("Error: File \"" + name + "\" could not be created.");
  }
  void close() {
    if (mFile >= 0) H5Fclose(mFile);
    mFile = -1;
  }
  bool isOpen() const { return mFile >= 0; }
  hid_t root() const { return mFile; }
  const std::string& name() const { return mName; }

  bool exists(hid_t loc, const std::string& name) const { return H5LTfind_dataset(loc, name.c_str()) == 1; }
  hid_t createGroup(hid_t loc, const std::string& name) {
    const hid_t g = H5Gcreate(loc, name.c_str(), H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
    if (g < 0) throw std::ios::failure("Error: cannot create group \"" + name + "\".");
    return g;
  }
  hid_t openGroup(hid_t loc, const std::string& name) {
    const hid_t g = H5Gopen(loc, name.c_str(), H5P_DEFAULT);
    if (g < 0) throw std::ios::failure("Error: cannot open group \"" + name + "\" of \"" + mName + "\".");
    return g;
  }
  void closeGroup(hid_t g) { H5Gclose(g); }
  hid_t openDataset(hid_t loc, const std::string& name) {
    const hid_t d = H5Dopen(loc, name.c_str(), H5P_DEFAULT);
    if (d < 0) throw std::ios::failure("Error: cannot open dataset \"" + name + "\" of \"" + mName + "\".");
    return d;
  }
  static bool canAccess(const std::string& name) { return access(name.c_str(), F_OK) == 0; }  // Hdf5File::canAccess

  // number of elements of a dataset (any rank)
  uint64_t elementCount(hid_t loc, const std::string& name) const {
    int rank = 0;
    if (H5LTget_dataset_ndims(loc, name.c_str(), &rank) < 0) throw std::ios::failure("Error: dataset \"" + name + "\" not found in \"" + mName + "\".");
    std::vector<hsize_t> dims(rank > 0 ? rank : 1, 1);
    H5T_class_t cls;
    size_t tsize = 0;
    if (H5LTget_dataset_info(loc, name.c_str(), dims.data(), &cls, &tsize) < 0) throw std::ios::failure("Error: cannot read the size of \"" + name + "\".");
    uint64_t n = 1;
    for (int i = 0; i < rank; ++i) n *= dims[i];
    return n;
  }
  std::vector<float> readFloats(hid_t loc, const std::string& name) const {
    std::vector<float> v(elementCount(loc, name));
    if (H5LTread_dataset(loc, name.c_str(), H5T_NATIVE_FLOAT, v.data()) < 0) throw std::ios::failure("Error: cannot read dataset \"" + name + "\".");
    return v;
  }
  std::vector<uint64_t> readIndices(hid_t loc, const std::string& name) const {
    std::vector<uint64_t> v(elementCount(loc, name));
    if (H5LTread_dataset(loc, name.c_str(), H5T_STD_U64LE, v.data()) < 0) throw std::ios::failure("Error: cannot read dataset \"" + name + "\".");
    return v;
  }
  float readFloatScalar(hid_t loc, const std::string& name) const {
    const auto v = readFloats(loc, name);
    if (v.size() != 1) throw std::ios::failure("Error: \"" + name + "\" is not a scalar.");
    return v[0];
  }
  uint64_t readIndexScalar(hid_t loc, const std::string& name) const {
    const auto v = readIndices(loc, name);
    if (v.size() != 1) throw std::ios::failure("Error: \"" + name + "\" is not a scalar.");
    return v[0];
  }

  // datasets are created with the dimension order of the reference: (nt,) nz, ny, nx on disk (Hdf5File.cpp:301-328)
  hid_t createDataset(hid_t loc, const std::string& name, const std::vector<hsize_t>& dims, const std::vector<hsize_t>& chunk, bool isFloat,
                      unsigned deflate) {
    const hid_t space = H5Screate_simple((int)dims.size(), dims.data(), nullptr);
    const hid_t plist = H5Pcreate(H5P_DATASET_CREATE);
    bool chunked = !chunk.empty();
    for (hsize_t c : chunk) chunked = chunked && c > 0;
    if (chunked) H5Pset_chunk(plist, (int)chunk.size(), chunk.data());
    if (chunked) H5Pset_deflate(plist, deflate);
    const hid_t d = H5Dcreate(loc, name.c_str(), isFloat ? H5T_NATIVE_FLOAT : H5T_STD_U64LE, space, H5P_DEFAULT, plist, H5P_DEFAULT);
    H5Pclose(plist);
    H5Sclose(space);
    if (d < 0) throw std::ios::failure("Error: cannot create dataset \"" + name + "\".");
    setStringAttribute(loc, name, "domain_type", "real");
    setStringAttribute(loc, name, "data_type", isFloat ? "float" : "long");
    return d;
  }
  void closeDataset(hid_t d) { H5Dclose(d); }
  // write `count` elements at `start` (both in dataset dimension order)
  void writeHyperslab(hid_t dataset, const std::vector<hsize_t>& start, const std::vector<hsize_t>& count, const void* data, bool isFloat = true) {
    const hid_t fspace = H5Dget_space(dataset);
    H5Sselect_hyperslab(fspace, H5S_SELECT_SET, start.data(), nullptr, count.data(), nullptr);
    const hid_t mspace = H5Screate_simple((int)count.size(), count.data(), nullptr);
    const herr_t e = H5Dwrite(dataset, isFloat ? H5T_NATIVE_FLOAT : H5T_STD_U64LE, mspace, fspace, H5P_DEFAULT, data);
    H5Sclose(mspace);
    H5Sclose(fspace);
    if (e < 0) throw std::ios::failure("Error: cannot write to the output file.");
  }
  // read `count` elements at `start` (dataset dimension order) into a dense buffer (Hdf5File::readHyperSlab, Hdf5File.cpp:559-600)
  void readHyperslab(hid_t dataset, const std::vector<hsize_t>& start, const std::vector<hsize_t>& count, float* data) const {
    const hid_t fspace = H5Dget_space(dataset);
    H5Sselect_hyperslab(fspace, H5S_SELECT_SET, start.data(), nullptr, count.data(), nullptr);
    const hid_t mspace = H5Screate_simple((int)count.size(), count.data(), nullptr);
    const herr_t e = H5Dread(dataset, H5T_NATIVE_FLOAT, mspace, fspace, H5P_DEFAULT, data);
    H5Sclose(mspace);
    H5Sclose(fspace);
    if (e < 0) throw std::ios::failure("Error: cannot read from \"" + mName + "\".");
  }
  void writeWhole(hid_t loc, const std::string& name, const std::vector<hsize_t>& dims, const std::vector<hsize_t>& chunk, const void* data,
                  bool isFloat, unsigned deflate) {
    const hid_t d = createDataset(loc, name, dims, chunk, isFloat, deflate);
    const herr_t e = H5Dwrite(d, isFloat ? H5T_NATIVE_FLOAT : H5T_STD_U64LE, H5S_ALL, H5S_ALL, H5P_DEFAULT, data);
    H5Dclose(d);
    if (e < 0) throw std::ios::failure("Error: cannot write dataset \"" + name + "\".");
  }
  // Hdf5File::writeScalarValue (Hdf5File.cpp:695-748): an existing scalar (later leg of a checkpointed run) is overwritten
  void writeScalar(hid_t loc, const std::string& name, float v) { writeScalarImpl(loc, name, &v, true); }
  void writeScalar(hid_t loc, const std::string& name, uint64_t v) { writeScalarImpl(loc, name, &v, false); }

  void setStringAttribute(hid_t loc, const std::string& obj, const std::string& attr, const std::string& value) {
    H5LTset_attribute_string(loc, obj.c_str(), attr.c_str(), value.c_str());
  }
  void setLongLongAttribute(hid_t loc, const std::string& obj, const std::string& attr, long long v) {
    H5LTset_attribute_long_long(loc, obj.c_str(), attr.c_str(), &v, 1);
  }
  void setFloatAttribute(hid_t loc, const std::string& obj, const std::string& attr, float v) {
    H5LTset_attribute_float(loc, obj.c_str(), attr.c_str(), &v, 1);
  }
  std::string getStringAttribute(hid_t loc, const std::string& obj, const std::string& attr) const {
    char buf[512] = {};
    if (H5LTget_attribute_string(loc, obj.c_str(), attr.c_str(), buf) < 0) return "";
    return buf;
  }

 private:
  void writeScalarImpl(hid_t loc, const std::string& name, const void* v, bool isFloat) {
    if (!exists(loc, name)) {
      writeWhole(loc, name, {1, 1, 1}, {}, v, isFloat, 0);
      return;
    }
    const hid_t d = openDataset(loc, name);
    const herr_t e = H5Dwrite(d, isFloat ? H5T_NATIVE_FLOAT : H5T_STD_U64LE, H5S_ALL, H5S_ALL, H5P_DEFAULT, v);
    H5Dclose(d);
    if (e < 0) throw std::ios::failure("Error: cannot write dataset \"" + name + "\".");
  }
  hid_t mFile = -1;
  std::string mName;
};

}  // namespace kwhost
