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

// Finalize epilogue (GATK's normalizeLikelihoods + the poorly-modelled-read test): one thread per read.
// O(pairs) byte work next to O(cells) DP: a few microseconds per chunk, but it saves the caller a pass over the
// matrix.  Statement for statement the arithmetic of finalize_region (phmm_prepost.cpp): same bits.
__global__ void phmm_finalize_rows(double* __restrict__ out, const ReadMeta* __restrict__ rmeta, const uint32_t* __restrict__ rnh,
                                   uint8_t* __restrict__ poorly, uint32_t n_reads, double log10_mismap, double err_rate) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const ReadMeta rm = rmeta[r];
  const uint32_t nh = rnh[r];
  double* row = out + rm.out_off;
  double best = -INFINITY;
  for (uint32_t h = 0; h < nh; ++h) best = fmax(best, row[h]);
  const double cap = __dadd_rn(best, log10_mismap);
  for (uint32_t h = 0; h < nh; ++h)
    if (row[h] < cap) row[h] = cap;
  const double max_err = fmin(2.0, ceil(__dmul_rn((double)(int)read_len_of(rm), err_rate)));
  poorly[r] = (nh > 0 && best < __dmul_rn(max_err, -4.0)) ? 1 : 0;
}

cudaError_t launch_finalize(double* out, const ReadMeta* rmeta, const uint32_t* rnh, uint8_t* poorly, uint32_t n_reads, double log10_mismap,
                            double err_rate, cudaStream_t s) {
  if (!n_reads) return cudaSuccess;
  phmm_finalize_rows<<<(n_reads + 127) / 128, 128, 0, s>>>(out, rmeta, rnh, poorly, n_reads, log10_mismap, err_rate);
  return cudaGetLastError();
}

}  // namespace fcsphmm
