// Striped generic kernels (FP32 from the host-built pair list, FP64 from the rerun queue).
#include "phmm_generic.cuh"
#include "phmm_registry.h"

namespace fcsphmm {

cudaError_t launch_generic_f32(const KParams& p, unsigned grid, cudaStream_t s) {
  phmm_generic<float, false><<<grid, 32, generic_smem_bytes<float>(p.n_sym), s>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_generic_f64(const KParams& p, unsigned grid, cudaStream_t s) {
  phmm_generic<double, true><<<grid, 32, generic_smem_bytes<double>(p.n_sym), s>>>(p);
  return cudaGetLastError();
}

}  // namespace fcsphmm
