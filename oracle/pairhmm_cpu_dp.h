/* oracle/pairhmm_cpu_dp.h — shared between pairhmm_cpu_simd.c and pairhmm_cpu_dp.c. */
#ifndef PAIRHMM_CPU_DP_H
#define PAIRHMM_CPU_DP_H
#include <stdint.h>
#define W 16
typedef struct {
  float pMM[W], pGM[W], pMX[W], pXX[W], pMY[W], pYY[W], pm[W], px[W], y0mask[W];
  int32_t rb[W];
} row_consts;
typedef void (*dp_hap_fn)(const row_consts* rows, int Lmax, const uint8_t* hap, int Lh, float* buf, float* S_out);
void dp_hap_avx512(const row_consts* rows, int Lmax, const uint8_t* hap, int Lh, float* buf, float* S_out);
void dp_hap_avx2(const row_consts* rows, int Lmax, const uint8_t* hap, int Lh, float* buf, float* S_out);
void dp_hap_generic(const row_consts* rows, int Lmax, const uint8_t* hap, int Lh, float* buf, float* S_out);
#endif
