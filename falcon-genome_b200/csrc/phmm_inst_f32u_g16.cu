// FP32 wavefront kernels (uniform gap-continuation quality), G = 16 lanes per read.
#include "phmm_classes.h"
#include "phmm_inst.cuh"
namespace fcsphmm {
extern const KernelEntry kEntriesF32UG16[] = {PHMM_F32U_G16(PHMM_ENTRY_F32U) PHMM_ENTRY_END};
}
