// FP64 rerun kernel, uniform gap-continuation form, register tier 1.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f64u_tier1, kTierF64UT1, QUEUE, double, true, 1, 1, PHMM_F64U_TIER1, PHMM_F64U_TIER1_N, PHMM_CLASSDESC_F64)
}
