// phmm_inst.cuh — macro that instantiates one kernel class and registers its launcher.
#pragma once
#include "phmm_kernel.cuh"
#include "phmm_registry.h"

namespace fcsphmm {

template <typename T, int G, int R, bool LIST, bool UG>
struct Launcher {
  static constexpr int MINB = min_blocks_for(R, (int)sizeof(T), UG);
  static cudaError_t launch(const KParams& p, unsigned grid, size_t smem, cudaStream_t s) {
    phmm_kernel<T, G, R, LIST, UG, MINB><<<grid, 32, smem, s>>>(p);
    return cudaGetLastError();
  }
  static size_t smem_bytes(uint32_t hs_cap, uint32_t hap_stage_bytes) {
    return Layout<T, G, R, LIST>::smem_bytes(hs_cap, hap_stage_bytes);
  }
  static cudaError_t set_max_smem(size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(phmm_kernel<T, G, R, LIST, UG, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(phmm_kernel<T, G, R, LIST, UG, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
};

#define PHMM_ENTRY(T, F64, LIST, UG, G, R) \
  {G, R, F64, UG, &Launcher<T, G, R, LIST, UG>::launch, &Launcher<T, G, R, LIST, UG>::smem_bytes, \
   &Launcher<T, G, R, LIST, UG>::set_max_smem, Launcher<T, G, R, LIST, UG>::MINB},
#define PHMM_ENTRY_F32(G, R) PHMM_ENTRY(float, false, false, false, G, R)
#define PHMM_ENTRY_F32U(G, R) PHMM_ENTRY(float, false, false, true, G, R)
#define PHMM_ENTRY_F64(G, R) PHMM_ENTRY(double, true, true, false, G, R)
#define PHMM_ENTRY_F64U(G, R) PHMM_ENTRY(double, true, true, true, G, R)
#define PHMM_ENTRY_END {0, 0, false, false, nullptr, nullptr, nullptr, 0}

}  // namespace fcsphmm
