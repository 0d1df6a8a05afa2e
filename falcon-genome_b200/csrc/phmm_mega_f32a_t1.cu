// FP32 wavefront kernel, all-uniform form (constant insertion / deletion / continuation qualities), register tier 1.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f32a_tier1, kTierF32AT1, TASK, float, false, 2, 1, PHMM_F32A_TIER1, PHMM_F32A_TIER1_N, PHMM_CLASSDESC_F32A)
}
