// phmm_prepost.cpp — the steps GATK's PairHMMLikelihoodCalculationEngine performs either side of the
// native PairHMM call [upstream GATK4; restated from the published
// algorithm — none of it is in /root/reference], so that a caller can hand the library raw reads:
//
//  before:  insertion / deletion quals default to 45 when the BAM has no BI/BD tags; PCR indel error model: at
//           every base the tandem-repeat length around it lowers both gap-open quals to
//           max(10, round(40 - exp(repeatLength / (rateFactor * pi)) + 1)); then capMinimumReadQualities:
//           base quals capped by the mapping quality, "q < 18 -> 6", insertion / deletion quals floored at
//           MIN_USABLE_Q_SCORE (6); gap continuation constant 10.
//  after:   per read, likelihoods are capped at best + log10(global mismapping rate = 10^-4.5);
//           reads whose best likelihood is below  min(2, ceil(len * 0.02)) * -4.0  are flagged as
//           poorly modelled.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "phmm_engine.h"

namespace fcsphmm {

namespace {
constexpr int kMaxStrUnit = 8;
constexpr int kMaxRepeatLen = 20;

// number of consecutive copies of `unit` at the start (leading) or end (trailing) of seq[0..n)
int count_repeats(const uint8_t* unit, int ulen, const uint8_t* seq, int n, bool leading) {
  int reps = 0;
  if (leading) {
    for (int s = 0; s + ulen <= n; s += ulen) {
      if (std::memcmp(seq + s, unit, (size_t)ulen) != 0) break;
      ++reps;
    }
  } else {
    for (int e = n; e - ulen >= 0; e -= ulen) {
      if (std::memcmp(seq + e - ulen, unit, (size_t)ulen) != 0) break;
      ++reps;
    }
  }
  return reps;
}

// GATK findTandemRepeatUnits(readBases, offset).getRight(): repeat count of the best unit around offset.
// As published, the best backward / forward unit starts as the single base at offset / offset + 1 and is
// replaced only by the first unit length whose repeat count exceeds 1 (the assignment sits inside
// `if (maxBW > 1)`); the count variable keeps the value of the last length tried.
int tandem_repeat_length(const uint8_t* b, int n, int offset) {
  int max_bw = 0;
  const uint8_t* best_bw = b + offset;
  int best_bw_len = 1;
  for (int str = 1; str <= kMaxStrUnit; ++str) {
    if (offset + 1 - str < 0) break;
    const uint8_t* unit = b + offset - str + 1;
    max_bw = count_repeats(unit, str, b, offset + 1, false);
    if (max_bw > 1) { best_bw = unit; best_bw_len = str; break; }
  }
  int max_rl = max_bw;
  if (offset < n - 1) {
    const uint8_t* best_fw = b + offset + 1;
    int best_fw_len = 1, max_fw = 0;
    for (int str = 1; str <= kMaxStrUnit; ++str) {
      if (offset + str + 1 > n) break;
      const uint8_t* unit = b + offset + 1;
      max_fw = count_repeats(unit, str, b + offset + 1, n - offset - 1, true);
      if (max_fw > 1) { best_fw = unit; best_fw_len = str; break; }
    }
    if (best_fw_len == best_bw_len && std::memcmp(best_fw, best_bw, (size_t)best_fw_len) == 0) {
      max_rl = max_bw + max_fw;
    } else {
      max_bw = count_repeats(best_fw, best_fw_len, b, offset + 1, false);
      max_rl = max_fw + max_bw;
    }
  }
  return std::min(max_rl, kMaxRepeatLen);
}

int fast_round(double d) { return d > 0.0 ? (int)(d + 0.5) : (int)(d - 0.5); }
}  // namespace

int prepare_read(const uint8_t* bases, const uint8_t* raw_q, int32_t len, int32_t mapq, const uint8_t* bam_ins, const uint8_t* bam_del,
                 const fcs_phmm_prep_params* pp, uint8_t* out_q, uint8_t* out_i, uint8_t* out_d, uint8_t* out_c) {
  if (len < 0 || (len > 0 && (!bases || !raw_q || !out_q || !out_i || !out_d || !out_c))) return set_error(FCS_PHMM_EINVAL, "null array");
  fcs_phmm_prep_params p = {18, 6, 45, 10, 3};
  if (pp) p = *pp;
  for (int32_t k = 0; k < len; ++k) {
    int q = std::min<int>(raw_q[k], mapq < 0 ? 255 : mapq);
    out_q[k] = (uint8_t)(q < p.base_q_threshold ? p.min_usable_q : q);
    out_i[k] = bam_ins ? bam_ins[k] : (uint8_t)p.default_indel_q;
    out_d[k] = bam_del ? bam_del[k] : (uint8_t)p.default_indel_q;
    out_c[k] = (uint8_t)p.gcp;
  }
  if (p.pcr_model != 0 && len > 1) {
    const double rate = p.pcr_model == 1 ? 1.0 : (p.pcr_model == 2 ? 2.0 : 3.0);  // HOSTILE / AGGRESSIVE / CONSERVATIVE
    uint8_t cache[kMaxRepeatLen + 1];
    for (int r = 0; r <= kMaxRepeatLen; ++r)
      cache[r] = (uint8_t)std::max(10, fast_round(40.0 - std::exp(r / (rate * M_PI)) + 1.0));
    for (int32_t k = 1; k < len; ++k) {
      const int rl = tandem_repeat_length(bases, len, k - 1);
      out_i[k - 1] = std::min(out_i[k - 1], cache[rl]);
      out_d[k - 1] = std::min(out_d[k - 1], cache[rl]);
    }
  }
  // capMinimumReadQualities, second half: after the PCR model (GATK's order: applyPCRErrorModel, then the caps)
  // insertion and deletion qualities below MIN_USABLE_Q_SCORE are set to it (setToFixedValueIfTooLow(q, 6, 6));
  // only BAM-supplied BI/BD values can be that low.
  for (int32_t k = 0; k < len; ++k) {
    if (out_i[k] < p.min_usable_q) out_i[k] = (uint8_t)p.min_usable_q;
    if (out_d[k] < p.min_usable_q) out_d[k] = (uint8_t)p.min_usable_q;
  }
  return FCS_PHMM_OK;
}

int finalize_region(double* l, int32_t n_reads, int32_t n_haps, const int32_t* read_len, double log10_mismap, double err_rate,
                    uint8_t* poorly) {
  if (n_reads < 0 || n_haps < 0 || (n_reads > 0 && n_haps > 0 && !l)) return set_error(FCS_PHMM_EINVAL, "bad matrix");
  for (int32_t r = 0; r < n_reads; ++r) {
    double* row = l + (size_t)r * (size_t)n_haps;
    double best = -INFINITY;
    for (int32_t h = 0; h < n_haps; ++h) best = std::max(best, row[h]);
    const double cap = best + log10_mismap;
    for (int32_t h = 0; h < n_haps; ++h)
      if (row[h] < cap) row[h] = cap;
    if (poorly) {
      const double max_err = read_len ? std::min(2.0, std::ceil(read_len[r] * err_rate)) : 2.0;
      poorly[r] = (n_haps > 0 && best < max_err * -4.0) ? 1 : 0;
    }
  }
  return FCS_PHMM_OK;
}

}  // namespace fcsphmm
