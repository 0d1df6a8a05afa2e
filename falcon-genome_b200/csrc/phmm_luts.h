// phmm_luts.h — quality lookup tables (host copies; uploaded once per device).
#pragma once
namespace fcsphmm {
constexpr int kNumQual = 128;                                  // quals are masked with & 127
constexpr int kMmSize = (kNumQual * (kNumQual + 1)) / 2;       // 8256, index ((max*(max+1))>>1)+min
struct Luts {
  float ph2pr_f[kNumQual];
  double ph2pr_d[kNumQual];
  float mm_f[kMmSize];
  double mm_d[kMmSize];
};
const Luts& luts();
inline int mm_index(int i, int d) {
  int mn = i < d ? i : d, mx = i < d ? d : i;
  return ((mx * (mx + 1)) >> 1) + mn;
}
}  // namespace fcsphmm
