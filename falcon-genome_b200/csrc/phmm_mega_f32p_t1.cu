// FP32 wavefront kernel, haplotype-pair form (uniform gap-continuation quality, packed f32x2 arithmetic), register tier 1.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f32p_tier1, kTierF32PT1, PAIR, float, false, 3, 1, PHMM_F32P_TIER1, PHMM_F32P_TIER1_N, PHMM_CLASSDESC_F32)
}
