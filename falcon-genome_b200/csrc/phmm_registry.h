// phmm_registry.h — the compiled kernels: one per (precision, form, register tier), each
// holding several (lanes per read G, rows per lane R) classes (phmm_tiers.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "phmm_types.h"

namespace fcsphmm {

struct ClassDesc {
  int G, R;
  size_t (*smem_bytes)(uint32_t hs_cap, uint32_t hap_stage_bytes, uint32_t n_sym);
};

struct TierKernel {
  bool f64;        // double-precision, queue-driven rerun kernel
  int form;        // 0 general; 1 uniform gap-continuation quality (pGM / pXX from the constant bank);
                   // 2 all-uniform (pMM / pMX / pMY as well; FP32 only)
                   // 3 haplotype pairs: uniform gap-continuation quality, two haplotypes per lane in packed f32x2 arithmetic (FP32 only)
  int tier;        // 0: 16 CTAs/SM (<=128 regs), 1: 12 (<=168), 2: 8 (<=255)
  int min_blocks;  // __launch_bounds__ residency target (one-warp CTAs per SM)
  int n_classes;
  const ClassDesc* classes;
  cudaError_t (*launch)(const KParams& p, unsigned grid, size_t smem, cudaStream_t s);
  cudaError_t (*set_max_smem)(size_t bytes);
};

struct ClassRef {
  const TierKernel* tk;
  int cls;  // index inside tk->classes == Task::cls / KParams::seg_cls
  int G, R;
  const ClassRef* twin = nullptr;  // the same (G, R) in the other GCP form (general <-> uniform)
  size_t smem_bytes(uint32_t hs_cap, uint32_t hap_stage, uint32_t n_sym) const { return tk->classes[cls].smem_bytes(hs_cap, hap_stage, n_sym); }
};

// All compiled kernels (16).
const TierKernel* const* tier_kernels(int* n);
// Cheapest class covering a read of this length (rows needed = len + 1) when every lane group of the
// warp is filled, or nullptr.
// coarse: only classes on the coarse grid of rows per lane (<= 8, then multiples of 4)
const ClassRef* select_class(bool f64, int form, int read_len, bool coarse = false);
// Cheapest class per read served when only n_reads (>= 1) reads are left to fill the 32/G lane groups
// against haplotypes of about avg_hap_len columns: favours wide groups (large G, small R) for leftovers.
const ClassRef* select_class_for(bool f64, int form, int read_len, int n_reads, int avg_hap_len, bool coarse = false);
const ClassRef* find_class(bool f64, int form, int G, int R);
// Latency policy (under-filled calls): the class with at least min_G lanes per read and the fewest rows
// per lane that covers the read -- the shortest serial chain per haplotype column.
const ClassRef* select_class_wide(bool f64, int form, int read_len, int min_G);
// FP64 rerun queues are keyed by the (G, R) of the general-form FP64 class of the read.
int f64_queue_count();
int f64_queue_id(int G, int R);
const ClassRef* f64_queue_class(int qid, int form);

#ifndef PHMM_TIER1_MINB
#define PHMM_TIER1_MINB 12
#endif
constexpr int kTierMinBlocks[3] = {16, PHMM_TIER1_MINB, 8};  // (developer builds vary the middle tier: -DPHMM_TIER1_MINB=10)

// striped generic kernels (phmm_generic.cuh): any read / haplotype length
cudaError_t launch_generic_f32(const KParams& p, unsigned grid, cudaStream_t s);
cudaError_t launch_generic_f64(const KParams& p, unsigned grid, cudaStream_t s);
// finalize epilogue (phmm_generic_inst.cu): per read, cap the row at best + log10_mismap and flag poorly-modelled reads
cudaError_t launch_finalize(double* out, const ReadMeta* rmeta, const uint32_t* rnh, uint8_t* poorly, uint32_t n_reads, double log10_mismap,
                            double err_rate, cudaStream_t s);
constexpr int kGenericMaxSinglePassRead = 383;  // longer reads take the striped path (FP64 single-pass tiles end at 384 rows)
constexpr int kGenericMinHapLen = 2001;         // regions with a longer haplotype take the striped path

}  // namespace fcsphmm
