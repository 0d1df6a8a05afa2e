// FP32 wavefront kernels (general form), G = 4 lanes per read.
#include "phmm_classes.h"
#include "phmm_inst.cuh"
namespace fcsphmm {
extern const KernelEntry kEntriesF32G4[] = {PHMM_F32_G4(PHMM_ENTRY_F32) PHMM_ENTRY_END};
}
