// FP32 wavefront kernel, all-uniform form (constant insertion / deletion / continuation qualities), register tier 2.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f32a_tier2, kTierF32AT2, TASK, float, false, 2, 2, PHMM_F32A_TIER2, PHMM_F32A_TIER2_N, PHMM_CLASSDESC_F32A)
}
