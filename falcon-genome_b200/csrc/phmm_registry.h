// phmm_registry.h — table of compiled kernel classes (lanes per read G x rows per lane R).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "phmm_types.h"

namespace fcsphmm {

struct KernelEntry {
  int G, R;
  bool f64;  // double-precision rerun (queue-driven) kernel
  bool ug;   // uniform gap-continuation quality: pGM / pXX come from the constant bank
  cudaError_t (*launch)(const KParams& p, unsigned grid, size_t smem, cudaStream_t s);
  size_t (*smem_bytes)(uint32_t hs_cap, uint32_t hap_stage_bytes);
  cudaError_t (*set_max_smem)(size_t bytes);
  int min_blocks;  // __launch_bounds__ residency target (CTAs of one warp per SM)
};

// Every class compiled into the library; terminated by G == 0.
const KernelEntry* kernel_table();
const KernelEntry* find_kernel(bool f64, bool ug, int G, int R);

// Rows a read of length len needs: len + 1 (one boundary-replica row on top).
// Picks the cheapest compiled class; returns nullptr when none covers the read.
const KernelEntry* select_kernel(bool f64, bool ug, int read_len);

// Residency target (one-warp CTAs per SM).  The register file is split per SM sub-partition
// (16384 registers each), so the per-thread budget moves in steps: 2 warps per sub-partition
// (8 CTAs/SM) allow 255 registers, 3 (12/SM) allow 168, 4 (16/SM) allow 128.  A tile needs about
// 8 registers per row (7 with uniform GCP; twice that in double) plus ~18.
PHMM_HD inline constexpr int min_blocks_for(int R, int esz, bool ug) {
#ifdef PHMM_FORCE_MINB
  return PHMM_FORCE_MINB;
#else
  const int regs = (ug ? 7 : 8) * R * (esz / 4) + 18;
  return regs <= 126 ? 16 : (regs <= 170 ? 12 : 8);
#endif
}

}  // namespace fcsphmm
