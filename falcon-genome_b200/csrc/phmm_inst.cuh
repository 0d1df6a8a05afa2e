// phmm_inst.cuh — macro that instantiates one kernel class and registers its launcher.
#pragma once
#include "phmm_kernel.cuh"
#include "phmm_registry.h"

namespace fcsphmm {

template <typename T, int G, int R, bool LIST>
struct Launcher {
  static constexpr int MINB = min_blocks_for(R, (int)sizeof(T));
  static cudaError_t launch(const KParams& p, unsigned grid, size_t smem, cudaStream_t s) {
    phmm_kernel<T, G, R, LIST, MINB><<<grid, 32, smem, s>>>(p);
    return cudaGetLastError();
  }
  static size_t smem_bytes(uint32_t hs_cap, uint32_t hap_stage_bytes) {
    return Layout<T, G, R, LIST>::smem_bytes(hs_cap, hap_stage_bytes);
  }
  static cudaError_t set_max_smem(size_t bytes) {
    return cudaFuncSetAttribute(phmm_kernel<T, G, R, LIST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  }
};

#define PHMM_ENTRY_F32(G, R) \
  {G, R, false, &Launcher<float, G, R, false>::launch, &Launcher<float, G, R, false>::smem_bytes, \
   &Launcher<float, G, R, false>::set_max_smem, Launcher<float, G, R, false>::MINB},
#define PHMM_ENTRY_F64(G, R) \
  {G, R, true, &Launcher<double, G, R, true>::launch, &Launcher<double, G, R, true>::smem_bytes, \
   &Launcher<double, G, R, true>::set_max_smem, Launcher<double, G, R, true>::MINB},

}  // namespace fcsphmm
