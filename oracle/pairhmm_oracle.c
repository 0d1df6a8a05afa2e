/*
 * oracle/pairhmm_oracle.c — CPU restatement of the PairHMM forward likelihood.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under falcon-genome_b200/ may include, link
 * or call this file; it is the checker for tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * PARITY UNPINNED: /root/reference (falcon-computing/falcon-genome) contains no
 * PairHMM arithmetic and no likelihood fixtures.  It shells out to an un-vendored
 * GATK jar (src/workers/HTCWorker.cpp:51-58, --native-pair-hmm-threads at :85;
 * src/workers/Mutect2Worker.cpp:113-121; jar paths src/config.cpp:285-286), inside
 * which PairHMMLikelihoodCalculationEngine -> VectorLoglessPairHMM -> Intel GKL
 * computeLikelihoodsNative does the work.  GATK pin seen in the reference:
 * "3.7-2-g53263cf" (test/TestLog.cpp:63) / "GATK-3.8" (test/global.bash:18); the
 * GATK4 jar is unpinned.  This file restates the PUBLISHED algorithm of those
 * dependencies (GATK LoglessPairHMM + PairHMMModel, GKL Context<> and
 * compute_full_prob<>) as written down in SURVEY.md Appendix A, and is pinned only
 * by hand-derivable known answers and brute-force path enumeration (tests/).
 *
 * Arithmetic contract (shared bit-for-bit with the CUDA kernels):
 *   per cell, with d = (r-1,c-1), u = (r-1,c), l = (r,c-1):
 *     t  = M_d * pMM;  t = fma(X_d, pGM, t);  t = fma(Y_d, pGM, t);  M = t * prior
 *     X  = fma(X_u, pXX, M_u * pMX)
 *     Y  = fma(Y_l, pYY, M_l * pMY)
 *   i.e. GKL's  distm*(M*p_MM + X*p_GapM + Y*p_GapM),  M*p_MX + X*p_XX,
 *   M*p_MY + Y*p_YY  with the contraction a -mfma build performs, fixed explicitly
 *   so the float twin is a deterministic function (compile with -ffp-contract=off).
 *   Final sum: S = 0; for c ascending: S += (M[Lr][c] + X[Lr][c]).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAX_QUAL 127 /* native code masks quals with & 127 */
#define MM_SIZE (((MAX_QUAL + 1) * (MAX_QUAL + 2)) >> 1)
#define JAC_TOL 8.0
#define JAC_STEP 0.0001
#define JAC_INV_STEP (1.0 / JAC_STEP)
#define JAC_SIZE 80001 /* (int)(8.0 / 0.0001) + 1 */

static double g_ph2pr_d[128];
static float g_ph2pr_f[128];
static double g_mm_d[MM_SIZE];
static float g_mm_f[MM_SIZE];
static double g_jac_d[80001];
static float g_jac_f[80001];
static int g_init = 0;

static int fast_round(double d) { return d > 0.0 ? (int)(d + 0.5) : (int)(d - 0.5); }

/* MathUtils.approximateLog10SumLog10 with the Jacobian log table. */
static double approx_log10_sum_log10(double a, double b, int use_float_table) {
  double small = a, big = b;
  if (small > big) { small = b; big = a; }
  if (isinf(small) && small < 0) return big;
  if (isinf(big) && big < 0) return big;
  double diff = big - small;
  if (diff >= JAC_TOL) return big;
  int ind = fast_round(diff * JAC_INV_STEP);
  return big + (use_float_table ? (double)g_jac_f[ind] : g_jac_d[ind]);
}

void phmm_oracle_init(void) {
  if (g_init) return;
  for (int k = 0; k < JAC_SIZE; k++) {
    g_jac_d[k] = log10(1.0 + pow(10.0, -((double)k) * JAC_STEP));
    g_jac_f[k] = (float)g_jac_d[k]; /* GKL Context<float>: table is NUMBER-typed */
  }
  for (int k = 0; k < 128; k++) {
    g_ph2pr_d[k] = pow(10.0, -((double)k) / 10.0);
    g_ph2pr_f[k] = (float)g_ph2pr_d[k];
  }
  const double inv_ln10 = 1.0 / log(10.0);
  for (int i = 0, offset = 0; i <= MAX_QUAL; offset += ++i) {
    for (int j = 0; j <= i; j++) {
      double s_d = approx_log10_sum_log10(-0.1 * i, -0.1 * j, 0);
      double s_f = approx_log10_sum_log10(-0.1 * i, -0.1 * j, 1);
      g_mm_d[offset + j] = pow(10.0, log1p(-fmin(1.0, pow(10.0, s_d))) * inv_ln10);
      g_mm_f[offset + j] = (float)pow(10.0, log1p(-fmin(1.0, pow(10.0, s_f))) * inv_ln10);
    }
  }
  g_init = 1;
}

static int mm_index(int i, int d) {
  int mn = i < d ? i : d, mx = i < d ? d : i;
  return ((mx * (mx + 1)) >> 1) + mn;
}

/* LUT accessors so the tests can compare the product's device LUTs bit for bit. */
double phmm_oracle_ph2pr_d(int q) { phmm_oracle_init(); return g_ph2pr_d[q & 127]; }
float phmm_oracle_ph2pr_f(int q) { phmm_oracle_init(); return g_ph2pr_f[q & 127]; }
double phmm_oracle_mm_d(int i, int d) { phmm_oracle_init(); return g_mm_d[mm_index(i & 127, d & 127)]; }
float phmm_oracle_mm_f(int i, int d) { phmm_oracle_init(); return g_mm_f[mm_index(i & 127, d & 127)]; }

static inline int base_match(uint8_t r, uint8_t h) { return r == h || r == 'N' || h == 'N'; }

/*
 * SURVEY A.2 in double, K = 2^1020 (LoglessPairHMM.INITIAL_CONDITION / GKL
 * Context<double>).  Returns the raw scaled sum S (not the log).
 */
double phmm_oracle_sum_double(const uint8_t* rs, const uint8_t* q, const uint8_t* iq, const uint8_t* dq,
                              const uint8_t* gq, int Lr, const uint8_t* hap, int Lh) {
  phmm_oracle_init();
  if (Lr <= 0 || Lh <= 0) return 0.0;
  const double K = ldexp(1.0, 1020);
  size_t n = (size_t)Lh + 1;
  double* buf = (double*)calloc(6 * n, sizeof(double));
  double *Mp = buf, *Xp = buf + n, *Yp = buf + 2 * n, *Mc = buf + 3 * n, *Xc = buf + 4 * n, *Yc = buf + 5 * n;
  const double y0 = K / (double)Lh;
  for (int c = 0; c <= Lh; c++) { Mp[c] = 0.0; Xp[c] = 0.0; Yp[c] = y0; }
  for (int r = 1; r <= Lr; r++) {
    int qi = q[r - 1] & 127, ii = iq[r - 1] & 127, di = dq[r - 1] & 127, ci = gq[r - 1] & 127;
    double pMM = g_mm_d[mm_index(ii, di)], pGM = 1.0 - g_ph2pr_d[ci];
    double pMX = g_ph2pr_d[ii], pXX = g_ph2pr_d[ci], pMY = g_ph2pr_d[di], pYY = g_ph2pr_d[ci];
    double e = g_ph2pr_d[qi], pm = 1.0 - e, px = e / 3.0;
    uint8_t rb = rs[r - 1];
    Mc[0] = 0.0; Xc[0] = 0.0; Yc[0] = 0.0;
    for (int c = 1; c <= Lh; c++) {
      double prior = base_match(rb, hap[c - 1]) ? pm : px;
      double t = Mp[c - 1] * pMM;
      t = fma(Xp[c - 1], pGM, t);
      t = fma(Yp[c - 1], pGM, t);
      Mc[c] = t * prior;
      Xc[c] = fma(Xp[c], pXX, Mp[c] * pMX);
      Yc[c] = fma(Yc[c - 1], pYY, Mc[c - 1] * pMY);
    }
    double* t0;
    t0 = Mp; Mp = Mc; Mc = t0;
    t0 = Xp; Xp = Xc; Xc = t0;
    t0 = Yp; Yp = Yc; Yc = t0;
  }
  double S = 0.0;
  for (int c = 1; c <= Lh; c++) S += (Mp[c] + Xp[c]);
  free(buf);
  return S;
}

/* The float twin: same statement sequence in binary32, K = 2^120 (GKL Context<float>). */
float phmm_oracle_sum_float(const uint8_t* rs, const uint8_t* q, const uint8_t* iq, const uint8_t* dq,
                            const uint8_t* gq, int Lr, const uint8_t* hap, int Lh) {
  phmm_oracle_init();
  if (Lr <= 0 || Lh <= 0) return 0.0f;
  const float K = ldexpf(1.0f, 120);
  size_t n = (size_t)Lh + 1;
  float* buf = (float*)calloc(6 * n, sizeof(float));
  float *Mp = buf, *Xp = buf + n, *Yp = buf + 2 * n, *Mc = buf + 3 * n, *Xc = buf + 4 * n, *Yc = buf + 5 * n;
  const float y0 = K / (float)Lh;
  for (int c = 0; c <= Lh; c++) { Mp[c] = 0.0f; Xp[c] = 0.0f; Yp[c] = y0; }
  for (int r = 1; r <= Lr; r++) {
    int qi = q[r - 1] & 127, ii = iq[r - 1] & 127, di = dq[r - 1] & 127, ci = gq[r - 1] & 127;
    float pMM = g_mm_f[mm_index(ii, di)], pGM = 1.0f - g_ph2pr_f[ci];
    float pMX = g_ph2pr_f[ii], pXX = g_ph2pr_f[ci], pMY = g_ph2pr_f[di], pYY = g_ph2pr_f[ci];
    float e = g_ph2pr_f[qi], pm = 1.0f - e, px = e / 3.0f;
    uint8_t rb = rs[r - 1];
    Mc[0] = 0.0f; Xc[0] = 0.0f; Yc[0] = 0.0f;
    for (int c = 1; c <= Lh; c++) {
      float prior = base_match(rb, hap[c - 1]) ? pm : px;
      float t = Mp[c - 1] * pMM;
      t = fmaf(Xp[c - 1], pGM, t);
      t = fmaf(Yp[c - 1], pGM, t);
      Mc[c] = t * prior;
      Xc[c] = fmaf(Xp[c], pXX, Mp[c] * pMX);
      Yc[c] = fmaf(Yc[c - 1], pYY, Mc[c - 1] * pMY);
    }
    float* t0;
    t0 = Mp; Mp = Mc; Mc = t0;
    t0 = Xp; Xp = Xc; Xc = t0;
    t0 = Yp; Yp = Yc; Yc = t0;
  }
  float S = 0.0f;
  for (int c = 1; c <= Lh; c++) S += (Mp[c] + Xp[c]);
  free(buf);
  return S;
}

/* log10 L of the double path: log10(S) - log10(2^1020)  (SURVEY A.2 "result"). */
double phmm_oracle_log10_double(const uint8_t* rs, const uint8_t* q, const uint8_t* iq, const uint8_t* dq,
                                const uint8_t* gq, int Lr, const uint8_t* hap, int Lh) {
  double S = phmm_oracle_sum_double(rs, q, iq, dq, gq, Lr, hap, Lh);
  return log10(S) - log10(ldexp(1.0, 1020));
}

/*
 * Float result -> log10 L.  GKL writes (double)(log10f(S) - log10f(2^120)); libm's
 * log10f is not correctly rounded on every libc, so the contract here (and in the
 * kernels) is the correctly-rounded value (float)log10((double)S), which every
 * <=1-ulp double log10 produces except in astronomically rare double-rounding cases.
 */
double phmm_oracle_float_sum_to_log10(float S) {
  float l = (float)log10((double)S);
  float k = (float)log10((double)ldexpf(1.0f, 120));
  return (double)(l - k);
}

/*
 * GKL computeLikelihoodsNative per-pair policy.  float first; when the raw
 * float sum is below MIN_ACCEPTED = 1e-28f the pair is recomputed in double.
 * *used_double receives the decision.  force_double mirrors initNative(use_double).
 */
double phmm_oracle_pair(const uint8_t* rs, const uint8_t* q, const uint8_t* iq, const uint8_t* dq,
                        const uint8_t* gq, int Lr, const uint8_t* hap, int Lh, int force_double,
                        int* used_double, float* raw_float_sum) {
  float f = 0.0f;
  if (!force_double) f = phmm_oracle_sum_float(rs, q, iq, dq, gq, Lr, hap, Lh);
  if (raw_float_sum) *raw_float_sum = f;
  if (force_double || f < 1e-28f) {
    if (used_double) *used_double = 1;
    return phmm_oracle_log10_double(rs, q, iq, dq, gq, Lr, hap, Lh);
  }
  if (used_double) *used_double = 0;
  return phmm_oracle_float_sum_to_log10(f);
}

/*
 * One region: reads x haps -> out[r * n_haps + h] (read-major).
 * Read r occupies rd_len[r] bytes at offset rd_off[r] of each of the five byte planes.
 */
void phmm_oracle_region(const uint8_t* rbases, const uint8_t* rq, const uint8_t* ri, const uint8_t* rd,
                        const uint8_t* rc, const int64_t* rd_off, const int32_t* rd_len, int n_reads,
                        const uint8_t* hbases, const int64_t* hp_off, const int32_t* hp_len, int n_haps,
                        int force_double, double* out, uint8_t* used_double, float* raw_float) {
  for (int r = 0; r < n_reads; r++) {
    for (int h = 0; h < n_haps; h++) {
      int ud = 0;
      float rf = 0.0f;
      size_t o = (size_t)r * n_haps + h;
      out[o] = phmm_oracle_pair(rbases + rd_off[r], rq + rd_off[r], ri + rd_off[r], rd + rd_off[r],
                                rc + rd_off[r], rd_len[r], hbases + hp_off[h], hp_len[h], force_double, &ud, &rf);
      if (used_double) used_double[o] = (uint8_t)ud;
      if (raw_float) raw_float[o] = rf;
    }
  }
}

/*
 * Brute force for the known-answer tests: explicit sum over all
 * alignment paths in long double, no dynamic programming.  States: M consumes one
 * read base and one hap base, X (insertion) consumes a read base, Y (deletion)
 * consumes a hap base.  The path starts in the row-0 deletion state at any column
 * (weight 1/Lh), so the first read base is entered from Y with p_GapM; it ends after
 * the last read base in M or X.  Exponential: only for Lr, Lh <= 6.
 */
static long double bf_rec(int state, int r, int c, const uint8_t* rs, const uint8_t* q, const uint8_t* iq,
                          const uint8_t* dq, const uint8_t* gq, int Lr, const uint8_t* hap, int Lh) {
  /* state: 0=M 1=X 2=Y, currently at cell (r,c) (1-based, r bases of the read consumed). */
  long double total = 0.0L;
  if (r == Lr && state != 2) total += 1.0L;
  /* outgoing transitions use the quals of the destination row for M/X, of the current row for Y */
  if (r < Lr) {
    int rr = r + 1;
    long double ei = powl(10.0L, -(long double)(iq[rr - 1] & 127) / 10.0L);
    long double ed = powl(10.0L, -(long double)(dq[rr - 1] & 127) / 10.0L);
    long double ec = powl(10.0L, -(long double)(gq[rr - 1] & 127) / 10.0L);
    long double e = powl(10.0L, -(long double)(q[rr - 1] & 127) / 10.0L);
    (void)ed;
    /* -> M at (r+1, c+1) */
    if (c < Lh) {
      long double tr = (state == 0) ? (long double)g_mm_d[mm_index(iq[rr - 1] & 127, dq[rr - 1] & 127)] : (1.0L - ec);
      long double prior = base_match(rs[rr - 1], hap[c]) ? (1.0L - e) : e / 3.0L;
      total += tr * prior * bf_rec(0, rr, c + 1, rs, q, iq, dq, gq, Lr, hap, Lh);
    }
    /* -> X at (r+1, c): from M with p_MX, from X with p_XX (never from Y) */
    if (state == 0 && c >= 1) total += ei * bf_rec(1, rr, c, rs, q, iq, dq, gq, Lr, hap, Lh);
    if (state == 1 && c >= 1) total += ec * bf_rec(1, rr, c, rs, q, iq, dq, gq, Lr, hap, Lh);
  }
  /* -> Y at (r, c+1): from M with p_MY, from Y with p_YY, quals of row r (r >= 1) */
  if (r >= 1 && c < Lh) {
    long double ed = powl(10.0L, -(long double)(dq[r - 1] & 127) / 10.0L);
    long double ec = powl(10.0L, -(long double)(gq[r - 1] & 127) / 10.0L);
    if (state == 0) total += ed * bf_rec(2, r, c + 1, rs, q, iq, dq, gq, Lr, hap, Lh);
    if (state == 2) total += ec * bf_rec(2, r, c + 1, rs, q, iq, dq, gq, Lr, hap, Lh);
  }
  return total;
}

double phmm_oracle_bruteforce_log10(const uint8_t* rs, const uint8_t* q, const uint8_t* iq, const uint8_t* dq,
                                    const uint8_t* gq, int Lr, const uint8_t* hap, int Lh) {
  phmm_oracle_init();
  long double total = 0.0L;
  /* start: row 0, deletion state, column c0 in 0..Lh-1 means the first match lands on hap[c0] */
  for (int c0 = 0; c0 <= Lh; c0++)
    total += (1.0L / (long double)Lh) * bf_rec(2, 0, c0, rs, q, iq, dq, gq, Lr, hap, Lh);
  return (double)log10l(total);
}
