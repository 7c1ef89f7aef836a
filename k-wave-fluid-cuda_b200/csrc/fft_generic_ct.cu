// Per-length instantiations of the run-time-length kernels (fft_generic.cuh, CN > 0) for the grid sizes k-Wave users pick most often:
// 2^a 3^b 5^c multiples of 8 and 16.  One translation unit per group: nvcc ... -DKW_CT_GROUP=<0..3> fft_generic_ct.cu
#include "fft_generic.cuh"

#ifndef KW_CT_GROUP
#error "compile with -DKW_CT_GROUP=<0..3>"
#endif

namespace kw {
namespace generic {
#define KW_CT_CASE(N) case N: *out = make_ops<N>(n); return true;
#if KW_CT_GROUP == 0
bool ct_ops_group0(int n, FftOps* out) {
  switch (n) {
    KW_CT_CASE(96) KW_CT_CASE(120) KW_CT_CASE(144) KW_CT_CASE(160) KW_CT_CASE(192) KW_CT_CASE(200) KW_CT_CASE(216) KW_CT_CASE(240) KW_CT_CASE(288)
    default: return false;
  }
}
#elif KW_CT_GROUP == 1
bool ct_ops_group1(int n, FftOps* out) {
  switch (n) {
    KW_CT_CASE(320) KW_CT_CASE(360) KW_CT_CASE(384) KW_CT_CASE(400) KW_CT_CASE(432) KW_CT_CASE(480) KW_CT_CASE(576) KW_CT_CASE(600)
    default: return false;
  }
}
#elif KW_CT_GROUP == 2
bool ct_ops_group2(int n, FftOps* out) {
  switch (n) {
    KW_CT_CASE(640) KW_CT_CASE(720) KW_CT_CASE(768) KW_CT_CASE(800) KW_CT_CASE(864) KW_CT_CASE(960) KW_CT_CASE(1000) KW_CT_CASE(1152)
    default: return false;
  }
}
#else
bool ct_ops_group3(int n, FftOps* out) {
  switch (n) {
    KW_CT_CASE(1200) KW_CT_CASE(1280) KW_CT_CASE(1296) KW_CT_CASE(1440) KW_CT_CASE(1536) KW_CT_CASE(1600) KW_CT_CASE(1728) KW_CT_CASE(1920) KW_CT_CASE(2048)
    default: return false;
  }
}
#endif
}  // namespace generic
}  // namespace kw
