// FP32 wavefront kernels (general form), G = 32 lanes per read.
#include "phmm_classes.h"
#include "phmm_inst.cuh"
namespace fcsphmm {
extern const KernelEntry kEntriesF32G32[] = {PHMM_F32_G32(PHMM_ENTRY_F32) PHMM_ENTRY_END};
}
