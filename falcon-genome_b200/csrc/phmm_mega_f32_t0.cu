// FP32 wavefront kernel, general form, register tier 0.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f32_tier0, kTierF32T0, TASK, float, false, 0, 0, PHMM_F32_TIER0, PHMM_F32_TIER0_N, PHMM_CLASSDESC_F32)
}
