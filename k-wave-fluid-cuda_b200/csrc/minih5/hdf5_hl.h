/* hdf5_hl.h -- the H5LT ("lite") calls used by Hdf5/Hdf5File.cpp:716-1070 of the reference; see hdf5.h. */
#ifndef MINIH5_HDF5_HL_H
#define MINIH5_HDF5_HL_H
#include "hdf5.h"
#ifdef __cplusplus
extern "C" {
#endif
herr_t H5LTread_dataset(hid_t loc, const char* name, hid_t type, void* buffer);
herr_t H5LTget_dataset_info(hid_t loc, const char* name, hsize_t* dims, H5T_class_t* cls, size_t* type_size);
herr_t H5LTget_dataset_ndims(hid_t loc, const char* name, int* rank);
herr_t H5LTfind_dataset(hid_t loc, const char* name);
herr_t H5LTset_attribute_string(hid_t loc, const char* obj, const char* attr, const char* value);
herr_t H5LTset_attribute_long_long(hid_t loc, const char* obj, const char* attr, const long long* value, size_t n);
herr_t H5LTset_attribute_float(hid_t loc, const char* obj, const char* attr, const float* value, size_t n);
herr_t H5LTget_attribute_string(hid_t loc, const char* obj, const char* attr, char* value);
herr_t H5LTget_attribute_long_long(hid_t loc, const char* obj, const char* attr, long long* value);
herr_t H5LTget_attribute_float(hid_t loc, const char* obj, const char* attr, float* value);
#ifdef __cplusplus
}
#endif
#endif
