// Run-time-length FFT kernels: registry of the launch tables (device code: fft_generic.cuh; the per-length instantiations of the
// common sizes live in fft_generic_ct.cu, compiled in four groups so that the build parallelises).
#include "fft_generic.cuh"

namespace kw {
namespace generic {
bool ct_ops_group0(int n, FftOps* out);
bool ct_ops_group1(int n, FftOps* out);
bool ct_ops_group2(int n, FftOps* out);
bool ct_ops_group3(int n, FftOps* out);
// lengths with kernels of their own, else the run-time-length instantiation (CN = 0)
static FftOps ops_for(int n) {
  FftOps ops{};
#ifndef KW_GENERIC_NO_COMMON
  if (ct_ops_group0(n, &ops) || ct_ops_group1(n, &ops) || ct_ops_group2(n, &ops) || ct_ops_group3(n, &ops)) return ops;
#endif
  return make_ops<0>(n);
}
}  // namespace generic

bool generic_length_supported(int n) {
  generic::GenPlan pl;
  return generic::make_plan(n, &pl);
}

const FftOps* get_generic_fft_ops(int n) {
  if (!generic_length_supported(n)) return nullptr;
  static std::mutex mu;
  static std::map<int, FftOps> tables;
  std::lock_guard<std::mutex> lk(mu);
  auto it = tables.find(n);
  if (it == tables.end()) {
    it = tables.emplace(n, generic::ops_for(n)).first;
  }
  return &it->second;
}

}  // namespace kw
