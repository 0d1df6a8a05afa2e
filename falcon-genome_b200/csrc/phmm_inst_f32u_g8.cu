// FP32 wavefront kernels (uniform gap-continuation quality), G = 8 lanes per read.
#include "phmm_classes.h"
#include "phmm_inst.cuh"
namespace fcsphmm {
extern const KernelEntry kEntriesF32UG8[] = {PHMM_F32U_G8(PHMM_ENTRY_F32U) PHMM_ENTRY_END};
}
