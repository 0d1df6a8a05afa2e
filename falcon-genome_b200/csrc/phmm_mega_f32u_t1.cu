// FP32 wavefront kernel, uniform gap-continuation form, register tier 1.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f32u_tier1, kTierF32UT1, TASK, float, false, 1, 1, PHMM_F32U_TIER1, PHMM_F32U_TIER1_N, PHMM_CLASSDESC_F32)
}
