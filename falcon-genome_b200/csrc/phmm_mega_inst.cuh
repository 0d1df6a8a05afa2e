// phmm_mega_inst.cuh — defines one tier kernel and its TierKernel descriptor.
#pragma once
#include "phmm_mega.cuh"
#include "phmm_registry.h"

namespace fcsphmm {

template <typename T, int G, int R, bool LIST, int FORM>
size_t class_smem_bytes(uint32_t hs_cap, uint32_t hap_stage_bytes, uint32_t n_sym) {
  return Layout<T, G, R, LIST, FORM>::smem_bytes(hs_cap, hap_stage_bytes, n_sym);
}

// the shared-memory layout of the general and uniform-GCP forms is the same (form 0)
#define PHMM_CLASSDESC_F32(I, G, R) {G, R, &class_smem_bytes<float, G, R, false, 0>},
#define PHMM_CLASSDESC_F32A(I, G, R) {G, R, &class_smem_bytes<float, G, R, false, 2>},
#define PHMM_CLASSDESC_F64(I, G, R) {G, R, &class_smem_bytes<double, G, R, true, 0>},

// KERNEL: kernel symbol; DESC: exported TierKernel; KIND: TASK (FP32) or QUEUE (FP64)
#define PHMM_TIER_UNIT(KERNEL, DESC, KIND, T_, F64_, FORM_, TIER_, LIST_MACRO, N_, CLASSDESC)                          \
  PHMM_DEFINE_##KIND##_KERNEL(KERNEL, T_, FORM_, kTierMinBlocks[TIER_], LIST_MACRO)                                    \
  static const ClassDesc DESC##_classes[] = {LIST_MACRO(CLASSDESC)};                                                  \
  static cudaError_t DESC##_launch(const KParams& p, unsigned grid, size_t smem, cudaStream_t s) {                    \
    KERNEL<<<grid, 32, smem, s>>>(p);                                                                                 \
    return cudaGetLastError();                                                                                        \
  }                                                                                                                   \
  static cudaError_t DESC##_set_max_smem(size_t bytes) {                                                              \
    cudaError_t e = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);            \
    if (e != cudaSuccess) return e;                                                                                   \
    return cudaFuncSetAttribute(KERNEL, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
  }                                                                                                                   \
  extern const TierKernel DESC = {F64_, FORM_, TIER_, kTierMinBlocks[TIER_], N_, DESC##_classes, &DESC##_launch, &DESC##_set_max_smem};

}  // namespace fcsphmm
