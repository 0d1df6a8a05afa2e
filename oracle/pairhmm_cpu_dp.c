/*
 * oracle/pairhmm_cpu_dp.c — inner DP of the CPU baseline (see pairhmm_cpu_simd.c).
 * Compiled three times by the Makefile (-DDP_NAME=dp_hap_avx512 / _avx2 / _generic with the
 * matching -m flags); pairhmm_cpu_simd.c picks one at run time with __builtin_cpu_supports.
 * TEST / BENCH INFRASTRUCTURE ONLY.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include "pairhmm_cpu_dp.h"

/* One haplotype against W bottom-aligned reads.  */
void DP_NAME(
    const row_consts* rows, int Lmax, const uint8_t* hap, int Lh, float* buf, float* S_out) {
  const float K = ldexpf(1.0f, 120);
  const float y0 = K / (float)Lh;
  size_t n = (size_t)(Lh + 1) * W;
  float *Mp = buf, *Xp = buf + n, *Yp = buf + 2 * n, *Mc = buf + 3 * n, *Xc = buf + 4 * n, *Yc = buf + 5 * n;
  for (int c = 0; c <= Lh; c++)
    for (int l = 0; l < W; l++) { Mp[c * W + l] = 0.0f; Xp[c * W + l] = 0.0f; Yp[c * W + l] = y0; }
  for (int r = 0; r < Lmax; r++) {
    const row_consts* rc = &rows[r];
    for (int l = 0; l < W; l++) { Mc[l] = 0.0f; Xc[l] = 0.0f; Yc[l] = rc->y0mask[l] * y0; }
    for (int c = 1; c <= Lh; c++) {
      const int32_t hb = hap[c - 1];
      const int hn = hb == 'N';
      const float *mp = Mp + (size_t)(c - 1) * W, *xp = Xp + (size_t)(c - 1) * W, *yp = Yp + (size_t)(c - 1) * W;
      const float *mu = Mp + (size_t)c * W, *xu = Xp + (size_t)c * W;
      const float *ml = Mc + (size_t)(c - 1) * W, *yl = Yc + (size_t)(c - 1) * W;
      float *mo = Mc + (size_t)c * W, *xo = Xc + (size_t)c * W, *yo = Yc + (size_t)c * W;
#pragma omp simd
      for (int l = 0; l < W; l++) {
        int match = (rc->rb[l] == hb) | (rc->rb[l] == 'N') | hn;
        float prior = match ? rc->pm[l] : rc->px[l];
        float t = mp[l] * rc->pMM[l];
        t = __builtin_fmaf(xp[l], rc->pGM[l], t);
        t = __builtin_fmaf(yp[l], rc->pGM[l], t);
        mo[l] = t * prior;
        xo[l] = __builtin_fmaf(xu[l], rc->pXX[l], mu[l] * rc->pMX[l]);
        yo[l] = __builtin_fmaf(yl[l], rc->pYY[l], ml[l] * rc->pMY[l]);
      }
    }
    float* t0;
    t0 = Mp; Mp = Mc; Mc = t0;
    t0 = Xp; Xp = Xc; Xc = t0;
    t0 = Yp; Yp = Yc; Yc = t0;
  }
  float S[W];
  for (int l = 0; l < W; l++) S[l] = 0.0f;
  for (int c = 1; c <= Lh; c++)
    for (int l = 0; l < W; l++) S[l] += (Mp[(size_t)c * W + l] + Xp[(size_t)c * W + l]);
  for (int l = 0; l < W; l++) S_out[l] = S[l];
}

