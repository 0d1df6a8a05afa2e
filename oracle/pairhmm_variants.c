/*
 * oracle/pairhmm_variants.c — arithmetic VARIANTS of the float path, to bound "parity unpinned".
 *
 * TEST INFRASTRUCTURE ONLY (tools/variant_table.py, tests/test_oracle.py); never linked into the product.
 *
 * The reference's PairHMM arithmetic lives in an un-vendored dependency (Intel GKL inside the GATK jar,
 * /root/reference/src/config.cpp:285-286, call site src/workers/HTCWorker.cpp:85), so the oracle's float twin
 * pins ONE of the functions a GKL build may compute.  The places where a real GKL binary can differ from the
 * pinned contract are enumerated here as flags of one scalar restatement; tools/variant_table.py counts, per
 * BASELINE config, how many pairs change their float->double fallback decision or move by more than 1e-6 in
 * log10 L under each variant (profiles/r02_oracle_variants.md).
 *
 *   PHMM_VAR_NOFMA      every multiply and add rounded separately (GKL's AVX build: -mavx has no FMA; the pinned
 *                       contract is the contraction of an FMA-capable build, e.g. the AVX-512 one)
 *   PHMM_VAR_FTZ        flush-to-zero + denormals-are-zero, as GKL's initNative sets in MXCSR [upstream]
 *   PHMM_VAR_POWF       ph2pr[q] = powf(10.f, -(float)q / 10.f) as GKL Context<float> computes it [upstream],
 *                       instead of the correctly rounded (float)pow(10.0, -q / 10.0)
 *   PHMM_VAR_LOG10F     libm log10f(S) - log10f(2^120) instead of the correctly rounded float log10
 *   PHMM_VAR_SPLITSUM   last-row sums of M and X accumulated separately and added at the end (the accumulation
 *                       order of GKL's vector kernels) instead of S += (M + X) per column
 *
 * (matchToMatch: GKL computes the table in double from a NUMBER-typed Jacobian table and casts -- "called only once
 * during library load - don't bother to optimize with single precision fp" [upstream] -- which is what
 * pairhmm_oracle.c does, so it is not a variant.)
 *
 * Compile with -ffp-contract=off (oracle/Makefile): the only fused operations are the explicit fmaf() calls.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <xmmintrin.h>
#include <pmmintrin.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PHMM_VAR_NOFMA 1
#define PHMM_VAR_FTZ 2
#define PHMM_VAR_POWF 4
#define PHMM_VAR_LOG10F 8
#define PHMM_VAR_SPLITSUM 16

float phmm_oracle_ph2pr_f(int q);
float phmm_oracle_mm_f(int i, int d);
void phmm_oracle_init(void);
double phmm_oracle_log10_double(const uint8_t* rs, const uint8_t* q, const uint8_t* iq, const uint8_t* dq,
                                const uint8_t* gq, int Lr, const uint8_t* hap, int Lh);
double phmm_oracle_float_sum_to_log10(float S);

static float g_ph2pr_powf[128];
static int g_var_init;
static void var_init(void) {
  if (g_var_init) return;
  phmm_oracle_init();
  for (int x = 0; x < 128; x++) g_ph2pr_powf[x] = powf(10.f, -((float)x) / 10.f);
  g_var_init = 1;
}

/* how many of the 128 ph2pr entries differ between the two ways of computing the table (reported in the table) */
int phmm_variant_ph2pr_diffs(void) {
  var_init();
  int n = 0;
  for (int x = 0; x < 128; x++) n += g_ph2pr_powf[x] != phmm_oracle_ph2pr_f(x);
  return n;
}

float phmm_variant_ph2pr_powf(int q) {
  var_init();
  return g_ph2pr_powf[q & 127];
}

static inline int vmatch(uint8_t r, uint8_t h) { return r == h || r == 'N' || h == 'N'; }

/* the float path under `flags`; the caller has set MXCSR for PHMM_VAR_FTZ */
static float variant_sum_float(int flags, const uint8_t* rs, const uint8_t* q, const uint8_t* iq, const uint8_t* dq,
                               const uint8_t* gq, int Lr, const uint8_t* hap, int Lh, float* buf) {
  const float* ph = (flags & PHMM_VAR_POWF) ? g_ph2pr_powf : NULL;
  const float K = ldexpf(1.0f, 120);
  size_t n = (size_t)Lh + 1;
  float *Mp = buf, *Xp = buf + n, *Yp = buf + 2 * n, *Mc = buf + 3 * n, *Xc = buf + 4 * n, *Yc = buf + 5 * n;
  const float y0 = K / (float)Lh;
  for (int c = 0; c <= Lh; c++) { Mp[c] = 0.0f; Xp[c] = 0.0f; Yp[c] = y0; }
  const int nofma = flags & PHMM_VAR_NOFMA;
  for (int r = 1; r <= Lr; r++) {
    int qi = q[r - 1] & 127, ii = iq[r - 1] & 127, di = dq[r - 1] & 127, ci = gq[r - 1] & 127;
    const float e = ph ? ph[qi] : phmm_oracle_ph2pr_f(qi);
    const float pc = ph ? ph[ci] : phmm_oracle_ph2pr_f(ci);
    const float pMM = phmm_oracle_mm_f(ii, di), pGM = 1.0f - pc;
    const float pMX = ph ? ph[ii] : phmm_oracle_ph2pr_f(ii), pXX = pc, pMY = ph ? ph[di] : phmm_oracle_ph2pr_f(di), pYY = pc;
    const float pm = 1.0f - e, px = e / 3.0f;
    const uint8_t rb = rs[r - 1];
    Mc[0] = 0.0f; Xc[0] = 0.0f; Yc[0] = 0.0f;
    if (nofma) {
      for (int c = 1; c <= Lh; c++) {
        const float prior = vmatch(rb, hap[c - 1]) ? pm : px;
        float t = Mp[c - 1] * pMM;
        t = t + Xp[c - 1] * pGM;
        t = t + Yp[c - 1] * pGM;
        Mc[c] = t * prior;
        Xc[c] = Mp[c] * pMX + Xp[c] * pXX;
        Yc[c] = Mc[c - 1] * pMY + Yc[c - 1] * pYY;
      }
    } else {
      for (int c = 1; c <= Lh; c++) {
        const float prior = vmatch(rb, hap[c - 1]) ? pm : px;
        float t = Mp[c - 1] * pMM;
        t = fmaf(Xp[c - 1], pGM, t);
        t = fmaf(Yp[c - 1], pGM, t);
        Mc[c] = t * prior;
        Xc[c] = fmaf(Xp[c], pXX, Mp[c] * pMX);
        Yc[c] = fmaf(Yc[c - 1], pYY, Mc[c - 1] * pMY);
      }
    }
    float* t0;
    t0 = Mp; Mp = Mc; Mc = t0;
    t0 = Xp; Xp = Xc; Xc = t0;
    t0 = Yp; Yp = Yc; Yc = t0;
  }
  if (flags & PHMM_VAR_SPLITSUM) {
    float sM = 0.0f, sX = 0.0f;
    for (int c = 1; c <= Lh; c++) { sM += Mp[c]; sX += Xp[c]; }
    return sM + sX;
  }
  float S = 0.0f;
  for (int c = 1; c <= Lh; c++) S += (Mp[c] + Xp[c]);
  return S;
}

/*
 * Whole batch under one variant (flat layout of pairhmm_cpu_simd.c).  out = final log10 L with the float-first /
 * double-fallback rule, used_double = the decision, raw_float = the raw float sums.  OpenMP over reads.
 */
int64_t phmm_variant_batch(int flags, const uint8_t* rbases, const uint8_t* rq, const uint8_t* ri, const uint8_t* rd,
                           const uint8_t* rcq, const int64_t* rd_off, const int32_t* rd_len, const uint8_t* hbases,
                           const int64_t* hp_off, const int32_t* hp_len, const int32_t* reg_read0,
                           const int32_t* reg_nreads, const int32_t* reg_hap0, const int32_t* reg_nhaps,
                           const int64_t* reg_out0, int n_regions, double* out, uint8_t* used_double,
                           float* raw_float, int nthreads) {
  var_init();
  int64_t ntasks = 0;
  for (int g = 0; g < n_regions; g++) ntasks += reg_nreads[g];
  int32_t* t_reg = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntasks ? ntasks : 1));
  int32_t* t_r = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntasks ? ntasks : 1));
  int64_t k = 0;
  int maxlh = 1;
  for (int g = 0; g < n_regions; g++) {
    for (int r = 0; r < reg_nreads[g]; r++) { t_reg[k] = g; t_r[k] = r; k++; }
    for (int h = 0; h < reg_nhaps[g]; h++)
      if (hp_len[reg_hap0[g] + h] > maxlh) maxlh = hp_len[reg_hap0[g] + h];
  }
  int64_t n_double = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : n_double)
  {
    const unsigned int saved_csr = _mm_getcsr();
    float* buf = (float*)malloc(sizeof(float) * 6 * ((size_t)maxlh + 1));
#pragma omp for schedule(dynamic, 8)
    for (int64_t t = 0; t < ntasks; t++) {
      const int g = t_reg[t], r = t_r[t];
      const int64_t ro = rd_off[reg_read0[g] + r];
      const int Lr = rd_len[reg_read0[g] + r];
      for (int h = 0; h < reg_nhaps[g]; h++) {
        const int64_t ho = hp_off[reg_hap0[g] + h];
        const int Lh = hp_len[reg_hap0[g] + h];
        const int64_t o = reg_out0[g] + (int64_t)r * reg_nhaps[g] + h;
        if (flags & PHMM_VAR_FTZ) {
          _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
          _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
        }
        const float S = (Lr > 0 && Lh > 0) ? variant_sum_float(flags, rbases + ro, rq + ro, ri + ro, rd + ro, rcq + ro, Lr, hbases + ho, Lh, buf) : 0.0f;
        /* (under PHMM_VAR_FTZ the double rerun keeps the mode too: initNative sets MXCSR once for the thread) */
        if (raw_float) raw_float[o] = S;
        if (S < 1e-28f) {
          out[o] = phmm_oracle_log10_double(rbases + ro, rq + ro, ri + ro, rd + ro, rcq + ro, Lr, hbases + ho, Lh);
          if (used_double) used_double[o] = 1;
          n_double++;
        } else {
          out[o] = (flags & PHMM_VAR_LOG10F) ? (double)(log10f(S) - log10f(ldexpf(1.0f, 120))) : phmm_oracle_float_sum_to_log10(S);
          if (used_double) used_double[o] = 0;
        }
        _mm_setcsr(saved_csr);
      }
    }
    free(buf);
  }
  free(t_reg);
  free(t_r);
  return n_double;
}

/* Every pair in double precision (the north_star's reference arithmetic): out[o] = log10 L.  OpenMP over reads. */
void phmm_double_batch(const uint8_t* rbases, const uint8_t* rq, const uint8_t* ri, const uint8_t* rd, const uint8_t* rcq,
                       const int64_t* rd_off, const int32_t* rd_len, const uint8_t* hbases, const int64_t* hp_off,
                       const int32_t* hp_len, const int32_t* reg_read0, const int32_t* reg_nreads, const int32_t* reg_hap0,
                       const int32_t* reg_nhaps, const int64_t* reg_out0, int n_regions, double* out, int nthreads) {
  phmm_oracle_init();
  int64_t ntasks = 0;
  for (int g = 0; g < n_regions; g++) ntasks += reg_nreads[g];
  int32_t* t_reg = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntasks ? ntasks : 1));
  int32_t* t_r = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntasks ? ntasks : 1));
  int64_t k = 0;
  for (int g = 0; g < n_regions; g++)
    for (int r = 0; r < reg_nreads[g]; r++) { t_reg[k] = g; t_r[k] = r; k++; }
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t t = 0; t < ntasks; t++) {
    const int g = t_reg[t], r = t_r[t];
    const int64_t ro = rd_off[reg_read0[g] + r];
    const int Lr = rd_len[reg_read0[g] + r];
    for (int h = 0; h < reg_nhaps[g]; h++) {
      const int64_t ho = hp_off[reg_hap0[g] + h];
      out[reg_out0[g] + (int64_t)r * reg_nhaps[g] + h] =
          phmm_oracle_log10_double(rbases + ro, rq + ro, ri + ro, rd + ro, rcq + ro, Lr, hbases + ho, hp_len[reg_hap0[g] + h]);
    }
  }
  free(t_reg);
  free(t_r);
}
