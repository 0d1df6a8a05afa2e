// phmm_luts.cpp — host-side construction of the quality -> probability tables the kernels
// read (ph2pr from shared memory, matchToMatch from L2).  Product code: built here, NOT
// taken from oracle/.  Follows SURVEY.md Appendix A.3 [upstream GATK PairHMMModel /
// GKL ContextBase]: the match->match prior goes through approximateLog10SumLog10 with the
// 0.0001-step Jacobian log table; GKL's float context keeps that table in float.
#include "phmm_luts.h"

#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

namespace fcsphmm {

namespace {
constexpr double kJacTol = 8.0;
constexpr double kJacStep = 0.0001;
constexpr double kJacInvStep = 1.0 / kJacStep;
constexpr int kJacSize = 80001;  // (int)(8.0 / 0.0001) + 1

int fast_round(double d) { return d > 0.0 ? (int)(d + 0.5) : (int)(d - 0.5); }

template <typename J>
double approx_log10_sum(double a, double b, const std::vector<J>& jac) {
  double small = std::min(a, b), big = std::max(a, b);
  if (std::isinf(small) && small < 0) return big;
  double diff = big - small;
  if (diff >= kJacTol) return big;
  return big + (double)jac[fast_round(diff * kJacInvStep)];
}

Luts* g_luts = nullptr;
std::once_flag g_once;

void build() {
  Luts* L = new Luts();
  std::vector<double> jac_d(kJacSize);
  std::vector<float> jac_f(kJacSize);
  for (int k = 0; k < kJacSize; ++k) {
    jac_d[k] = std::log10(1.0 + std::pow(10.0, -((double)k) * kJacStep));
    jac_f[k] = (float)jac_d[k];
  }
  for (int q = 0; q < kNumQual; ++q) {
    L->ph2pr_d[q] = std::pow(10.0, -((double)q) / 10.0);
    L->ph2pr_f[q] = (float)L->ph2pr_d[q];
  }
  const double inv_ln10 = 1.0 / std::log(10.0);
  for (int i = 0, offset = 0; i < kNumQual; offset += ++i) {
    for (int j = 0; j <= i; ++j) {
      const double sd = approx_log10_sum(-0.1 * i, -0.1 * j, jac_d);
      const double sf = approx_log10_sum(-0.1 * i, -0.1 * j, jac_f);
      L->mm_d[offset + j] = std::pow(10.0, std::log1p(-std::min(1.0, std::pow(10.0, sd))) * inv_ln10);
      L->mm_f[offset + j] = (float)std::pow(10.0, std::log1p(-std::min(1.0, std::pow(10.0, sf))) * inv_ln10);
    }
  }
  g_luts = L;
}
}  // namespace

const Luts& luts() {
  std::call_once(g_once, build);
  return *g_luts;
}

}  // namespace fcsphmm
