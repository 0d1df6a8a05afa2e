// FP64 rerun kernels (general and uniform-GCP forms), G = 16 lanes per read.
#include "phmm_classes.h"
#include "phmm_inst.cuh"
namespace fcsphmm {
extern const KernelEntry kEntriesF64G16[] = {PHMM_F64_G16(PHMM_ENTRY_F64) PHMM_F64_G16(PHMM_ENTRY_F64U) PHMM_ENTRY_END};
}
