// FP64 rerun kernel, uniform gap-continuation form, register tier 2.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f64u_tier2, kTierF64UT2, QUEUE, double, true, 1, 2, PHMM_F64U_TIER2, PHMM_F64U_TIER2_N, PHMM_CLASSDESC_F64)
}
