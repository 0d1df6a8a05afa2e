// FP64 rerun kernel, general form, register tier 1.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f64_tier1, kTierF64T1, QUEUE, double, true, 0, 1, PHMM_F64_TIER1, PHMM_F64_TIER1_N, PHMM_CLASSDESC_F64)
}
