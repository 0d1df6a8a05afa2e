// FP64 rerun kernel, uniform gap-continuation form, register tier 0.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f64u_tier0, kTierF64UT0, QUEUE, double, true, 1, 0, PHMM_F64U_TIER0, PHMM_F64U_TIER0_N, PHMM_CLASSDESC_F64)
}
