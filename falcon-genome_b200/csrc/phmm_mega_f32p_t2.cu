// FP32 wavefront kernel, haplotype-pair form (uniform gap-continuation quality, packed f32x2 arithmetic), register tier 2.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f32p_tier2, kTierF32PT2, PAIR, float, false, 3, 2, PHMM_F32P_TIER2, PHMM_F32P_TIER2_N, PHMM_CLASSDESC_F32)
}
