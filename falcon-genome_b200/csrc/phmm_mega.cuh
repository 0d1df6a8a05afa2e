// phmm_mega.cuh — one kernel per register tier.  All (G, R) classes of a tier are instances of the
// same tile code; a task (FP32) or a queue segment (FP64) carries its class index and the CTA
// switches to the matching instance.  One launch per (tier, form) per chunk.
#pragma once
#include "phmm_kernel.cuh"
#include "phmm_tiers.h"

namespace fcsphmm {

#define PHMM_CASE_TASK(I, G, R) case I: run_task<T, G, R, FORM>(p, task, smem); break;
#define PHMM_CASE_PAIR(I, G, R) case I: run_task_pairs<G, R>(p, task, smem); break;
#define PHMM_CASE_QUEUE(I, G, R) case I: run_queue<T, G, R, FORM>(p, qid, cta, nctas, smem); break;

#define PHMM_DEFINE_TASK_KERNEL(NAME, T_, FORM_, MINB_, LIST_MACRO)                                   \
  __global__ void __launch_bounds__(32, MINB_) NAME(const __grid_constant__ KParams p) {          \
    using T = T_;                                                                                  \
    constexpr int FORM = FORM_;                                                                    \
    extern __shared__ __align__(128) uint8_t smem[];                                               \
    const Task task = p.tasks[blockIdx.x];                                                         \
    switch (task.cls) { LIST_MACRO(PHMM_CASE_TASK) default: break; }                               \
  }

// haplotype-pair form (FP32, uniform gap-continuation quality): T_ / FORM_ are fixed, kept for a uniform macro signature
#define PHMM_DEFINE_PAIR_KERNEL(NAME, T_, FORM_, MINB_, LIST_MACRO)                                   \
  __global__ void __launch_bounds__(32, MINB_) NAME(const __grid_constant__ KParams p) {          \
    extern __shared__ __align__(128) uint8_t smem[];                                               \
    const Task task = p.tasks[blockIdx.x];                                                         \
    switch (task.cls) { LIST_MACRO(PHMM_CASE_PAIR) default: break; }                               \
  }

#define PHMM_DEFINE_QUEUE_KERNEL(NAME, T_, FORM_, MINB_, LIST_MACRO)                                  \
  __global__ void __launch_bounds__(32, MINB_) NAME(const __grid_constant__ KParams p) {          \
    using T = T_;                                                                                  \
    constexpr int FORM = FORM_;                                                                    \
    extern __shared__ __align__(128) uint8_t smem[];                                               \
    uint32_t k = 0;                                                                                \
    while (k + 1 < p.n_seg && blockIdx.x >= p.seg_cta0[k + 1]) ++k;                                \
    const uint32_t qid = p.seg_qid[k], cta = blockIdx.x - p.seg_cta0[k];                           \
    const uint32_t nctas = p.seg_cta0[k + 1] - p.seg_cta0[k];                                      \
    const uint32_t qlen = p.rerun_count[qid];                                                      \
    if (qlen < p.seg_min[k] || qlen > p.seg_max[k]) return;                                        \
    switch (p.seg_cls[k]) { LIST_MACRO(PHMM_CASE_QUEUE) default: break; }                          \
  }

}  // namespace fcsphmm
