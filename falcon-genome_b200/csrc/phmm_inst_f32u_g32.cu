// FP32 wavefront kernels (uniform gap-continuation quality), G = 32 lanes per read.
#include "phmm_classes.h"
#include "phmm_inst.cuh"
namespace fcsphmm {
extern const KernelEntry kEntriesF32UG32[] = {PHMM_F32U_G32(PHMM_ENTRY_F32U) PHMM_ENTRY_END};
}
