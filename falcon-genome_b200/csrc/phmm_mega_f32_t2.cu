// FP32 wavefront kernel, general form, register tier 2.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f32_tier2, kTierF32T2, TASK, float, false, 0, 2, PHMM_F32_TIER2, PHMM_F32_TIER2_N, PHMM_CLASSDESC_F32)
}
