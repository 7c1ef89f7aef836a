// Driver around the REFERENCE's own Compression/CompressHelper.cpp (compiled from where it lies under /root/reference by
// oracle/Makefile; no reference source is copied into this repository).  It prints, as JSON, the known answers the
// Python restatement (oracle/compress_oracle.py) and the CUDA compression streams are pinned against:
//   * basis functions bE, bE_1 and their half-step-shifted variants (bit patterns) for a few (period, mos, harmonics)
//   * the 40-bit complex codec: encode and decode of a deterministic value list, for both exponent constants
// TEST INFRASTRUCTURE ONLY.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include <Compression/CompressHelper.h>

static uint32_t bits(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}

static void dump_basis(const char* name, const FloatComplex* b, size_t n, bool last) {
  std::printf("      \"%s\": [", name);
  for (size_t i = 0; i < n; ++i) std::printf("%s[%u, %u]", i ? ", " : "", bits(b[i].real()), bits(b[i].imag()));
  std::printf("]%s\n", last ? "" : ",");
}

int main() {
  struct Cfg { float period; hsize_t mos, harmonics; } cfgs[] = {{50.0f, 1, 2}, {50.0f, 2, 3}, {33.3f, 1, 1}, {8.0f, 1, 4}};
  std::printf("{\n  \"bases\": [\n");
  for (size_t k = 0; k < sizeof(cfgs) / sizeof(cfgs[0]); ++k) {
    CompressHelper& h = CompressHelper::getInstance();
    h.init(cfgs[k].period, cfgs[k].mos, cfgs[k].harmonics, true);
    const size_t n = h.getHarmonics() * h.getBSize();
    std::printf("    {\"period_bits\": %u, \"mos\": %llu, \"harmonics\": %llu, \"oSize\": %llu, \"bSize\": %llu,\n", bits(cfgs[k].period),
                (unsigned long long)cfgs[k].mos, (unsigned long long)cfgs[k].harmonics, (unsigned long long)h.getOSize(),
                (unsigned long long)h.getBSize());
    dump_basis("bE", h.getBE(), n, false);
    dump_basis("bE_1", h.getBE_1(), n, false);
    dump_basis("bE_shifted", h.getBEShifted(), n, false);
    dump_basis("bE_1_shifted", h.getBE_1Shifted(), n, true);
    std::printf("    }%s\n", k + 1 < sizeof(cfgs) / sizeof(cfgs[0]) ? "," : "");
    // the singleton deletes itself in its destructor; leak it on purpose instead (process exits right after)
  }
  std::printf("  ],\n  \"codec\": [\n");
  // deterministic value list: powers, tiny, huge, mixed signs, zeros
  std::vector<float> vals = {0.0f, -0.0f, 1.0f, -1.0f, 0.5f, 12345.678f, -0.5f, 3.0e-2f, 3.2e-2f, 1.0e-3f, 1.0e8f, 1.34e8f, 2.0e9f,
                             7.99f, 8.01f, 1.0e-9f, 1.9e-9f, 123.456f, -98765.4321f, 6.1e-5f, 2.5f, -1.0e-6f, 65535.0f, 65536.0f,
                             131071.0f, 0.0312500000f, 0.03f, 4.0e-3f};
  uint32_t lcg = 12345u;
  for (int i = 0; i < 40; ++i) {
    lcg = lcg * 1664525u + 1013904223u;
    const float mant = (float)(lcg >> 8) / 16777216.0f * 2.0f - 1.0f;
    lcg = lcg * 1664525u + 1013904223u;
    const int ex = (int)(lcg >> 27) - 12;
    vals.push_back(std::ldexp(mant, ex * 2));
  }
  bool first = true;
  for (int e : {CompressHelper::kMaxExpP, CompressHelper::kMaxExpU})
    for (size_t i = 0; i < vals.size(); ++i) {
      const size_t j = (i * 7 + 3) % vals.size();
      FloatComplex c(vals[i], vals[j]), d;
      uint8_t b[5];
      CompressHelper::convertFloatCTo40b(c, b, e);
      CompressHelper::convert40bToFloatC(b, d, e);
      std::printf("%s    {\"e\": %d, \"re\": %u, \"im\": %u, \"bytes\": [%u, %u, %u, %u, %u], \"dre\": %u, \"dim\": %u}", first ? "" : ",\n", e,
                  bits(c.real()), bits(c.imag()), b[0], b[1], b[2], b[3], b[4], bits(d.real()), bits(d.imag()));
      first = false;
    }
  std::printf("\n  ]\n}\n");
  std::fflush(stdout);
  std::_Exit(0);
}
