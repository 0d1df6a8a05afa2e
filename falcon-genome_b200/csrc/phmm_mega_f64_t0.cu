// FP64 rerun kernel, general form, register tier 0.
#include "phmm_mega_inst.cuh"
namespace fcsphmm {
PHMM_TIER_UNIT(phmm_f64_tier0, kTierF64T0, QUEUE, double, true, 0, 0, PHMM_F64_TIER0, PHMM_F64_TIER0_N, PHMM_CLASSDESC_F64)
}
