// FP32 wavefront kernel, uniform gap-continuation form, register tier 0.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f32u_tier0, kTierF32UT0, TASK, float, false, 1, 0, PHMM_F32U_TIER0, PHMM_F32U_TIER0_N, PHMM_CLASSDESC_F32)
}
