// FP32 wavefront kernels (general form), G = 8 lanes per read.
#include "phmm_classes.h"
#include "phmm_inst.cuh"
namespace fcsphmm {
extern const KernelEntry kEntriesF32G8[] = {PHMM_F32_G8(PHMM_ENTRY_F32) PHMM_ENTRY_END};
}
