// FP32 wavefront kernel, general form, register tier 1.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f32_tier1, kTierF32T1, TASK, float, false, 0, 1, PHMM_F32_TIER1, PHMM_F32_TIER1_N, PHMM_CLASSDESC_F32)
}
