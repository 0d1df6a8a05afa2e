/*
 * oracle/pairhmm_cpu_simd.c — the CPU baseline timed beside the GPU number.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (bench.py cpu_baseline and --impl reference);
 * never linked into the product library.
 *
 * Label: "AVX-512/AVX2 + OpenMP C PairHMM, GKL-equivalent semantics
 * (float first with 2^120 scaling, double rerun when the raw float sum < 1e-28f) —
 * not GKL itself".  The real thing (GATK + GKL as launched by
 * /root/reference/src/workers/HTCWorker.cpp:51-58,85) needs a JVM and the GATK jar,
 * neither of which exists in this image.
 *
 * Vectorisation is across READS of one region (16 reads per vector, one haplotype at a
 * time): lane l owns read l, rows are bottom-aligned so every lane finishes on the same
 * row, and the rows above a shorter read replicate the row-0 boundary exactly
 * (M = X = 0, Y = K/Lh).  Every real cell executes the statement sequence of
 * pairhmm_oracle.c, so results are bit-identical to the scalar float twin
 * (tests/test_oracle.py checks this).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <xmmintrin.h>
#include <pmmintrin.h>

float phmm_oracle_ph2pr_f(int q);
float phmm_oracle_mm_f(int i, int d);
void phmm_oracle_init(void);
double phmm_oracle_log10_double(const uint8_t* rs, const uint8_t* q, const uint8_t* iq, const uint8_t* dq,
                                const uint8_t* gq, int Lr, const uint8_t* hap, int Lh);
double phmm_oracle_float_sum_to_log10(float S);

#include "pairhmm_cpu_dp.h"

static dp_hap_fn pick_dp(void) {
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
      __builtin_cpu_supports("avx512dq"))
    return dp_hap_avx512;
  if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return dp_hap_avx2;
  return dp_hap_generic;
}

const char* phmm_cpu_isa(void) {
  dp_hap_fn f = pick_dp();
  return f == dp_hap_avx512 ? "avx512" : (f == dp_hap_avx2 ? "avx2+fma" : "generic");
}

/*
 * Whole batch.  Flat layout shared with the Python generators:
 *   reads:   five byte planes + rd_off[int64] + rd_len[int32]
 *   haps:    one byte plane   + hp_off[int64] + hp_len[int32]
 *   regions: reg_read0, reg_nreads, reg_hap0, reg_nhaps (int32), reg_out0 (int64)
 *   out[reg_out0 + r*nhaps + h]
 * Returns the number of pairs that took the double path.
 *
 * ftz != 0 sets flush-to-zero + denormals-are-zero in every worker thread, as GKL's
 * initNative does [upstream] (most of the DP matrix off the alignment diagonal underflows
 * binary32, and x86 handles subnormals with microcode assists ~100x slower).  That is the
 * setting the bench times.  ftz == 0 keeps IEEE subnormals: bit-identical to the scalar
 * float twin and to the CUDA kernel (the tests use this).
 */
int64_t phmm_cpu_batch(const uint8_t* rbases, const uint8_t* rq, const uint8_t* ri, const uint8_t* rd,
                       const uint8_t* rcq, const int64_t* rd_off, const int32_t* rd_len, const uint8_t* hbases,
                       const int64_t* hp_off, const int32_t* hp_len, const int32_t* reg_read0,
                       const int32_t* reg_nreads, const int32_t* reg_hap0, const int32_t* reg_nhaps,
                       const int64_t* reg_out0, int n_regions, double* out, uint8_t* used_double,
                       float* raw_float, int nthreads, int ftz) {
  phmm_oracle_init();
  const dp_hap_fn dp_hap = pick_dp();
  /* task list: (region, first read of a W-wide group) */
  int64_t ntasks = 0;
  for (int g = 0; g < n_regions; g++) ntasks += (reg_nreads[g] + W - 1) / W;
  int32_t* t_reg = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntasks ? ntasks : 1));
  int32_t* t_r0 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntasks ? ntasks : 1));
  int64_t k = 0;
  for (int g = 0; g < n_regions; g++)
    for (int r0 = 0; r0 < reg_nreads[g]; r0 += W) { t_reg[k] = g; t_r0[k] = r0; k++; }
  int64_t n_double = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : n_double)
  {
    unsigned int saved_csr = _mm_getcsr();
    if (ftz) {
      _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
      _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
    }
    row_consts* rows = NULL;
    size_t rows_cap = 0;
    float* buf = NULL;
    size_t buf_cap = 0;
#pragma omp for schedule(dynamic, 1)
    for (int64_t t = 0; t < ntasks; t++) {
      int g = t_reg[t], r0 = t_r0[t];
      int nr = reg_nreads[g] - r0;
      if (nr > W) nr = W;
      int Lmax = 0, Lhmax = 0;
      for (int l = 0; l < nr; l++) {
        int L = rd_len[reg_read0[g] + r0 + l];
        if (L > Lmax) Lmax = L;
      }
      for (int h = 0; h < reg_nhaps[g]; h++)
        if (hp_len[reg_hap0[g] + h] > Lhmax) Lhmax = hp_len[reg_hap0[g] + h];
      if (Lmax == 0 || Lhmax == 0) continue;
      if ((size_t)Lmax > rows_cap) { free(rows); rows_cap = (size_t)Lmax; rows = (row_consts*)malloc(sizeof(row_consts) * rows_cap); }
      size_t need = (size_t)6 * (Lhmax + 1) * W;
      if (need > buf_cap) { free(buf); buf_cap = need; buf = (float*)malloc(sizeof(float) * buf_cap); }
      for (int r = 0; r < Lmax; r++) {
        for (int l = 0; l < W; l++) {
          int pos = -1;
          int64_t off = 0;
          if (l < nr) {
            int ridx = reg_read0[g] + r0 + l;
            pos = r - (Lmax - rd_len[ridx]);
            off = rd_off[ridx];
          }
          row_consts* rc = &rows[r];
          if (pos < 0) { /* boundary replica: M = X = 0, Y stays K/Lh */
            rc->pMM[l] = rc->pGM[l] = rc->pMX[l] = rc->pXX[l] = rc->pMY[l] = 0.0f;
            rc->pYY[l] = 1.0f; rc->pm[l] = rc->px[l] = 0.0f; rc->y0mask[l] = 1.0f; rc->rb[l] = 0;
          } else {
            int qi = rq[off + pos] & 127, ii = ri[off + pos] & 127, di = rd[off + pos] & 127, ci = rcq[off + pos] & 127;
            float e = phmm_oracle_ph2pr_f(qi);
            rc->pMM[l] = phmm_oracle_mm_f(ii, di);
            rc->pGM[l] = 1.0f - phmm_oracle_ph2pr_f(ci);
            rc->pMX[l] = phmm_oracle_ph2pr_f(ii);
            rc->pXX[l] = phmm_oracle_ph2pr_f(ci);
            rc->pMY[l] = phmm_oracle_ph2pr_f(di);
            rc->pYY[l] = phmm_oracle_ph2pr_f(ci);
            rc->pm[l] = 1.0f - e;
            rc->px[l] = e / 3.0f;
            rc->y0mask[l] = 0.0f;
            rc->rb[l] = rbases[off + pos];
          }
        }
      }
      for (int h = 0; h < reg_nhaps[g]; h++) {
        int hidx = reg_hap0[g] + h;
        float S[W];
        dp_hap(rows, Lmax, hbases + hp_off[hidx], hp_len[hidx], buf, S);
        for (int l = 0; l < nr; l++) {
          int ridx = reg_read0[g] + r0 + l;
          int64_t o = reg_out0[g] + (int64_t)(r0 + l) * reg_nhaps[g] + h;
          if (raw_float) raw_float[o] = S[l];
          if (rd_len[ridx] == 0 || hp_len[hidx] == 0) { /* degenerate: log10(0) */
            out[o] = -INFINITY;
            if (used_double) used_double[o] = 1;
            n_double++;
          } else if (S[l] < 1e-28f) {
            out[o] = phmm_oracle_log10_double(rbases + rd_off[ridx], rq + rd_off[ridx], ri + rd_off[ridx],
                                              rd + rd_off[ridx], rcq + rd_off[ridx], rd_len[ridx],
                                              hbases + hp_off[hidx], hp_len[hidx]);
            if (used_double) used_double[o] = 1;
            n_double++;
          } else {
            out[o] = phmm_oracle_float_sum_to_log10(S[l]);
            if (used_double) used_double[o] = 0;
          }
        }
      }
    }
    free(rows);
    free(buf);
    _mm_setcsr(saved_csr);
  }
  free(t_reg);
  free(t_r0);
  return n_double;
}

int phmm_cpu_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
