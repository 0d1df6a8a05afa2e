// FP32 wavefront kernel, uniform gap-continuation form, register tier 2.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f32u_tier2, kTierF32UT2, TASK, float, false, 1, 2, PHMM_F32U_TIER2, PHMM_F32U_TIER2_N, PHMM_CLASSDESC_F32)
}
