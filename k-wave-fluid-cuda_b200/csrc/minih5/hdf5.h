/*
 * hdf5.h -- "minih5": the subset of the HDF5 1.8 C API that the k-Wave CUDA code base uses
 * (Hdf5/Hdf5File.cpp:97-1086 of the reference), implemented in this repository because no HDF5 library exists in the
 * build image.  It lets (a) the reference's own sources be compiled unchanged as the parity oracle /
 * reported baseline (oracle/ref_build), and (b) the C++ host of this repository read and write k-Wave files through the
 * same calls.
 *
 * Storage back end: REAL HDF5 files.  A file is parsed completely on H5Fopen and written completely on H5Fclose (objects live
 * in memory in between), in the subset of the HDF5 1.8 file format that libhdf5 / MATLAB produce with default settings and that
 * the reference itself writes: superblock 0, version-1 object headers, symbol-table groups, contiguous and chunked (version-1
 * B-tree) datasets with the deflate filter, float32 / uint64 data, fixed-string / float / long long attributes (minih5.cpp).
 * tools/h5lite.py is the independent Python implementation of the same subset used by the tests.
 */
#ifndef MINIH5_HDF5_H
#define MINIH5_HDF5_H

#include <stddef.h>
#include <stdint.h>
/* headers the real hdf5.h drags in and the reference relies on */
#ifdef __cplusplus
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <limits>
#endif
#include <immintrin.h>
#include <unistd.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hid_t;
typedef int herr_t;
typedef int htri_t;
typedef unsigned long long hsize_t;
typedef long long hssize_t;

#define H5I_BADID ((hid_t)(-1))
#define H5I_INVALID_HID ((hid_t)(-1))
#define H5P_DEFAULT ((hid_t)0)
#define H5S_ALL ((hid_t)0)
#define H5E_DEFAULT ((hid_t)0)
#define H5F_ACC_RDONLY 0x0000u
#define H5F_ACC_RDWR 0x0001u
#define H5F_ACC_TRUNC 0x0002u
#define H5F_ACC_EXCL 0x0004u
#define H5P_DATASET_CREATE ((hid_t)1)
#define H5T_NATIVE_FLOAT ((hid_t)0x7001)
#define H5T_STD_U64LE ((hid_t)0x7002)
#define H5T_NATIVE_UINT64 H5T_STD_U64LE

typedef enum H5S_seloper_t { H5S_SELECT_SET = 0 } H5S_seloper_t;
typedef enum H5G_obj_t { H5G_UNKNOWN = -1, H5G_GROUP = 0, H5G_DATASET = 1, H5G_TYPE = 2 } H5G_obj_t;
typedef enum H5O_type_t { H5O_TYPE_UNKNOWN = -1, H5O_TYPE_GROUP = 0, H5O_TYPE_DATASET = 1 } H5O_type_t;
typedef struct H5O_info_t {
  H5O_type_t type;
} H5O_info_t;
typedef int H5T_class_t;
typedef herr_t (*H5E_auto_t)(hid_t, void*);

herr_t H5Eset_auto(hid_t estack, H5E_auto_t func, void* client_data);

hid_t H5Fcreate(const char* name, unsigned flags, hid_t fcpl, hid_t fapl);
hid_t H5Fopen(const char* name, unsigned flags, hid_t fapl);
herr_t H5Fclose(hid_t file);
htri_t H5Fis_hdf5(const char* name);
herr_t H5Fget_filesize(hid_t file, hsize_t* size);

hid_t H5Gcreate(hid_t loc, const char* name, hid_t lcpl, hid_t gcpl, hid_t gapl);
hid_t H5Gopen(hid_t loc, const char* name, hid_t gapl);
herr_t H5Gclose(hid_t group);

htri_t H5Lexists(hid_t loc, const char* name, hid_t lapl);
htri_t H5Oexists_by_name(hid_t loc, const char* name, hid_t lapl);
herr_t H5Oget_info_by_name(hid_t loc, const char* name, H5O_info_t* info, hid_t lapl);
ssize_t H5Iget_name(hid_t id, char* name, size_t size);

hid_t H5Dcreate(hid_t loc, const char* name, hid_t type, hid_t space, hid_t lcpl, hid_t dcpl, hid_t dapl);
hid_t H5Dopen(hid_t loc, const char* name, hid_t dapl);
herr_t H5Dclose(hid_t dset);
hid_t H5Dget_space(hid_t dset);
herr_t H5Dread(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t plist, void* buf);
herr_t H5Dwrite(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t plist, const void* buf);

hid_t H5Screate_simple(int rank, const hsize_t* dims, const hsize_t* maxdims);
herr_t H5Sclose(hid_t space);
int H5Sget_simple_extent_ndims(hid_t space);
herr_t H5Sselect_hyperslab(hid_t space, H5S_seloper_t op, const hsize_t* start, const hsize_t* stride, const hsize_t* count,
                           const hsize_t* block);
herr_t H5Sselect_elements(hid_t space, H5S_seloper_t op, size_t nelem, const hsize_t* coord);

hid_t H5Pcreate(hid_t cls);
herr_t H5Pclose(hid_t plist);
herr_t H5Pset_chunk(hid_t plist, int ndims, const hsize_t* dim);
herr_t H5Pset_deflate(hid_t plist, unsigned level);

#ifdef __cplusplus
}
#endif
#endif
