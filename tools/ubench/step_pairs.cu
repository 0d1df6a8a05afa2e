// Microbenchmark: the PairHMM tile step with TWO haplotype columns per lane in packed f32x2 arithmetic
// (FFMA2 / FMUL2, sm_100a): the same read rows run against two haplotypes at once, the row constants are
// broadcast operands (R.F32 / UR.F32), the state lives in aligned register pairs.  Compared with the scalar
// step (shape A of step_shapes.cu).  cycles per warp-cell on one SM sub-partition; 8.0 = FMA-pipe peak.
// Developer tool.
#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define STEPS 4096

__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }

// FORM 1: uniform gap-continuation (per-row pMM, pMX, pMY, pYY).  FORM 2: all-uniform (M * cMX kept in P).
template <int R, int FORM, int UNROLL, int MINB, int GW, bool PACKED>
__global__ void __launch_bounds__(32, MINB) k(float* out, const float* in, float cGM, float cXX, float cMM, float cMX, int nsteps) {
  constexpr int TS = (((R + 3) / 4) | 1) * 4;  // lane stride in floats: odd multiple of 16 B
  constexpr int NV = (R + 3) / 4;
  extern __shared__ __align__(16) float tab[];  // 5 symbol rows x 32 lanes x TS, then the haplotype stream
  uint32_t* hs = reinterpret_cast<uint32_t*>(tab + 5 * 32 * TS);
  for (int i = threadIdx.x; i < 5 * 32 * TS; i += 32) tab[i] = 0.5f + 0.0001f * (i % 977);
  for (int i = threadIdx.x; i < 1024; i += 32) hs[i] = (uint32_t)((i * 7 + 3) % 5) | ((uint32_t)((i * 11 + 1) % 5) << 16);
  __syncwarp();
  const float* tl = tab + threadIdx.x * TS;
  if constexpr (PACKED) {
    float2 M[R], X[R], Y[R], P[FORM == 2 ? R : 1];
    float pMM[FORM == 2 ? 1 : R], pMX[FORM == 2 ? 1 : R], pMY[FORM == 2 ? 1 : R], pYY[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      M[k] = make_float2(in[threadIdx.x + k], in[threadIdx.x + k + 1]);
      X[k] = make_float2(in[32 + threadIdx.x + k], in[33 + threadIdx.x + k]);
      Y[k] = make_float2(in[64 + threadIdx.x + k], in[65 + threadIdx.x + k]);
      if constexpr (FORM == 2) P[k] = make_float2(0.f, 0.f);
      else { pMM[k] = in[96 + k]; pMX[k] = in[128 + k] * 0.01f; pMY[k] = in[160 + k] * 0.01f; }
      pYY[k] = in[192 + k] * 0.2f;
    }
    if constexpr (FORM == 2) { pMM[0] = cMM; pMX[0] = cMX; pMY[0] = cMX; }
    float2 dM = bc(0.f), dX = bc(0.f), dY = bc(0.f), acc = bc(0.f);
#pragma unroll(UNROLL)
    for (int t = 0; t < nsteps; ++t) {
      const uint32_t h2 = hs[t & 1023];
      const float* prowA = tl + (h2 & 0xffffu) * (32 * TS);
      const float* prowB = tl + (h2 >> 16) * (32 * TS);
      float prA[NV * 4], prB[NV * 4];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 f = *reinterpret_cast<const float4*>(prowA + v * 4);
        prA[v * 4] = f.x; prA[v * 4 + 1] = f.y; prA[v * 4 + 2] = f.z; prA[v * 4 + 3] = f.w;
        const float4 g = *reinterpret_cast<const float4*>(prowB + v * 4);
        prB[v * 4] = g.x; prB[v * 4 + 1] = g.y; prB[v * 4 + 2] = g.z; prB[v * 4 + 3] = g.w;
      }
      float2 uM, uX, uY;
      uM.x = __shfl_up_sync(0xffffffffu, M[R - 1].x, 1, GW); uM.y = __shfl_up_sync(0xffffffffu, M[R - 1].y, 1, GW);
      uX.x = __shfl_up_sync(0xffffffffu, X[R - 1].x, 1, GW); uX.y = __shfl_up_sync(0xffffffffu, X[R - 1].y, 1, GW);
      uY.x = __shfl_up_sync(0xffffffffu, Y[R - 1].x, 1, GW); uY.y = __shfl_up_sync(0xffffffffu, Y[R - 1].y, 1, GW);
      float2 nM[R], nX[R], nY[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const float2 md = k ? M[k - 1] : dM, xd = k ? X[k - 1] : dX, yd = k ? Y[k - 1] : dY;
        float2 s = __fmul2_rn(md, bc(FORM == 2 ? cMM : pMM[FORM == 2 ? 0 : k]));
        s = __ffma2_rn(xd, bc(cGM), s);
        s = __ffma2_rn(yd, bc(cGM), s);
        nM[k] = make_float2(__fmul_rn(s.x, prA[k]), __fmul_rn(s.y, prB[k]));
        if constexpr (FORM == 2) nY[k] = __ffma2_rn(Y[k], bc(pYY[k]), P[k]);
        else nY[k] = __ffma2_rn(Y[k], bc(pYY[k]), __fmul2_rn(M[k], bc(pMY[k])));
      }
      nX[0] = __ffma2_rn(uX, bc(pYY[0]), __fmul2_rn(uM, bc(pMX[0])));
      if constexpr (FORM == 2) {
#pragma unroll
        for (int k = 0; k < R; ++k) P[k] = __fmul2_rn(nM[k], bc(cMX));
#pragma unroll
        for (int k = 1; k < R; ++k) nX[k] = __ffma2_rn(nX[k - 1], bc(cXX), P[k - 1]);
      } else {
#pragma unroll
        for (int k = 1; k < R; ++k) nX[k] = __ffma2_rn(nX[k - 1], bc(cXX), __fmul2_rn(nM[k - 1], bc(pMX[k])));
      }
      acc = __fadd2_rn(acc, __fadd2_rn(nM[R - 1], nX[R - 1]));
      dM = uM; dX = uX; dY = uY;
#pragma unroll
      for (int k = 0; k < R; ++k) { M[k] = nM[k]; X[k] = nX[k]; Y[k] = nY[k]; }
    }
    float s = acc.x + acc.y;
#pragma unroll
    for (int k = 0; k < R; ++k) s += M[k].x + X[k].x + Y[k].x + M[k].y + X[k].y + Y[k].y;
    out[blockIdx.x * 32 + threadIdx.x] = s;
  } else {
    float M[R], X[R], Y[R], P[FORM == 2 ? R : 1];
    float pMM[FORM == 2 ? 1 : R], pMX[FORM == 2 ? 1 : R], pMY[FORM == 2 ? 1 : R], pYY[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      M[k] = in[threadIdx.x + k]; X[k] = in[32 + threadIdx.x + k]; Y[k] = in[64 + threadIdx.x + k];
      if constexpr (FORM == 2) P[k] = 0.f;
      else { pMM[k] = in[96 + k]; pMX[k] = in[128 + k] * 0.01f; pMY[k] = in[160 + k] * 0.01f; }
      pYY[k] = in[192 + k] * 0.2f;
    }
    if constexpr (FORM == 2) { pMM[0] = cMM; pMX[0] = cMX; pMY[0] = cMX; }
    float dM = 0, dX = 0, dY = 0, acc = 0;
#pragma unroll(UNROLL)
    for (int t = 0; t < nsteps; ++t) {
      const uint32_t h2 = hs[t & 1023];
      const float* prow = tl + (h2 & 0xffffu) * (32 * TS);
      float pr[NV * 4];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 f = *reinterpret_cast<const float4*>(prow + v * 4);
        pr[v * 4] = f.x; pr[v * 4 + 1] = f.y; pr[v * 4 + 2] = f.z; pr[v * 4 + 3] = f.w;
      }
      const float uM = __shfl_up_sync(0xffffffffu, M[R - 1], 1, GW), uX = __shfl_up_sync(0xffffffffu, X[R - 1], 1, GW),
                  uY = __shfl_up_sync(0xffffffffu, Y[R - 1], 1, GW);
      float nM[R], nX[R], nY[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const float md = k ? M[k - 1] : dM, xd = k ? X[k - 1] : dX, yd = k ? Y[k - 1] : dY;
        float s = __fmul_rn(md, FORM == 2 ? cMM : pMM[FORM == 2 ? 0 : k]);
        s = __fmaf_rn(xd, cGM, s);
        s = __fmaf_rn(yd, cGM, s);
        nM[k] = __fmul_rn(s, pr[k]);
        if constexpr (FORM == 2) nY[k] = __fmaf_rn(Y[k], pYY[k], P[k]);
        else nY[k] = __fmaf_rn(Y[k], pYY[k], __fmul_rn(M[k], pMY[k]));
      }
      nX[0] = __fmaf_rn(uX, pYY[0], __fmul_rn(uM, pMX[0]));
      if constexpr (FORM == 2) {
#pragma unroll
        for (int k = 0; k < R; ++k) P[k] = __fmul_rn(nM[k], cMX);
#pragma unroll
        for (int k = 1; k < R; ++k) nX[k] = __fmaf_rn(nX[k - 1], cXX, P[k - 1]);
      } else {
#pragma unroll
        for (int k = 1; k < R; ++k) nX[k] = __fmaf_rn(nX[k - 1], cXX, __fmul_rn(nM[k - 1], pMX[k]));
      }
      acc = __fadd_rn(acc, __fadd_rn(nM[R - 1], nX[R - 1]));
      dM = uM; dX = uX; dY = uY;
#pragma unroll
      for (int k = 0; k < R; ++k) { M[k] = nM[k]; X[k] = nX[k]; Y[k] = nY[k]; }
    }
    float s = acc;
#pragma unroll
    for (int k = 0; k < R; ++k) s += M[k] + X[k] + Y[k];
    out[blockIdx.x * 32 + threadIdx.x] = s;
  }
}


// Row-pair packing (single haplotype per lane, same tile as the scalar step): the state lives in ODD-aligned
// register pairs P_j = (row 2j-1, row 2j), row -1 being the value received from the lane above; the M phase
// of rows (2j, 2j+1) reads P_j as its diagonal operand, the Y phase runs on the pairs themselves, the
// products of the X chain are new-P_j x (pMX[2j], pMX[2j+1]); the prior multiply and the X chain stay scalar
// (they write single halves, which re-aligns the pairs for free).
template <int R, int FORM, int UNROLL, int MINB, int GW>
__global__ void __launch_bounds__(32, MINB) krp(float* out, const float* in, float cGM, float cXX, float cMM, float cMX, int nsteps) {
  constexpr int TS = (((R + 3) / 4) | 1) * 4;
  constexpr int NV = (R + 3) / 4;
  constexpr int NP = R / 2 + 1;          // pairs P_0..P_{NP-1} cover rows -1..2*NP-2 (>= R-1)
  constexpr int NE = (R + 1) / 2;        // even-aligned parameter pairs (rows 2j, 2j+1)
  extern __shared__ __align__(16) float tab[];
  uint32_t* hs = reinterpret_cast<uint32_t*>(tab + 5 * 32 * TS);
  for (int i = threadIdx.x; i < 5 * 32 * TS; i += 32) tab[i] = 0.5f + 0.0001f * (i % 977);
  for (int i = threadIdx.x; i < 1024; i += 32) hs[i] = (uint32_t)((i * 7 + 3) % 5) | ((uint32_t)((i * 11 + 1) % 5) << 16);
  __syncwarp();
  const float* tl = tab + threadIdx.x * TS;
  float2 M[NP], X[NP], Y[NP], P[FORM == 2 ? NP : 1];
  float2 pMMe[FORM == 2 ? 1 : NE], pMXe[FORM == 2 ? 1 : NE];  // even-aligned: (row 2j, row 2j+1)
  float2 pMYo[FORM == 2 ? 1 : NP], pYYo[NP];                  // odd-aligned: (row 2j-1, row 2j)
  float xx0, mx0;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    M[j] = make_float2(in[threadIdx.x + j], in[threadIdx.x + j + 1]);
    X[j] = make_float2(in[32 + threadIdx.x + j], in[33 + threadIdx.x + j]);
    Y[j] = make_float2(in[64 + threadIdx.x + j], in[65 + threadIdx.x + j]);
    pYYo[j] = make_float2(in[192 + j] * 0.2f, in[193 + j] * 0.21f);
    if constexpr (FORM == 2) P[j] = make_float2(0.f, 0.f);
    else pMYo[j] = make_float2(in[160 + j] * 0.01f, in[161 + j] * 0.011f);
  }
  if constexpr (FORM != 2) {
#pragma unroll
    for (int j = 0; j < NE; ++j) { pMMe[j] = make_float2(in[96 + j], in[97 + j] * 0.99f); pMXe[j] = make_float2(in[128 + j] * 0.01f, in[129 + j] * 0.011f); }
  }
  xx0 = in[250]; mx0 = in[251] * 0.01f;
  float acc = 0.f;
#pragma unroll(UNROLL)
  for (int t = 0; t < nsteps; ++t) {
    const uint32_t h2 = hs[t & 1023];
    const float* prow = tl + (h2 & 0xffffu) * (32 * TS);
    float pr[NV * 4];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 f = *reinterpret_cast<const float4*>(prow + v * 4);
      pr[v * 4] = f.x; pr[v * 4 + 1] = f.y; pr[v * 4 + 2] = f.z; pr[v * 4 + 3] = f.w;
    }
    // bottom row R-1 lives in pair (R-1+1)/2, half (R-1+1)&1
    constexpr int BJ = R / 2, BH = R & 1;  // row R-1 = 2*BJ-1+BH
    const float bM = BH ? M[BJ].y : M[BJ].x, bX = BH ? X[BJ].y : X[BJ].x, bY = BH ? Y[BJ].y : Y[BJ].x;
    const float uM = __shfl_up_sync(0xffffffffu, bM, 1, GW), uX = __shfl_up_sync(0xffffffffu, bX, 1, GW), uY = __shfl_up_sync(0xffffffffu, bY, 1, GW);
    float2 nM[NP], nX[NP], nY[NP];
    // ---- M phase: rows (2j, 2j+1) from the old pair P_j
#pragma unroll
    for (int j = 0; j < NE; ++j) {
      float2 s;
      if constexpr (FORM == 2) s = __fmul2_rn(M[j], bc(cMM));
      else s = __fmul2_rn(M[j], pMMe[j]);
      s = __ffma2_rn(X[j], bc(cGM), s);
      s = __ffma2_rn(Y[j], bc(cGM), s);
      // row 2j -> pair j half y; row 2j+1 -> pair j+1 half x
      nM[j].y = __fmul_rn(s.x, pr[2 * j]);
      if (2 * j + 1 < R) nM[j + 1].x = __fmul_rn(s.y, pr[2 * j + 1]);
    }
    nM[0].x = uM;
    // ---- Y phase on the pairs themselves (half x of pair 0 is the boundary slot: overwritten below)
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if constexpr (FORM == 2) nY[j] = __ffma2_rn(Y[j], pYYo[j], P[j]);
      else nY[j] = __ffma2_rn(Y[j], pYYo[j], __fmul2_rn(M[j], pMYo[j]));
    }
    nY[0].x = uY;
    // ---- X chain: products packed, chain scalar
    float2 q[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if constexpr (FORM == 2) { q[j] = __fmul2_rn(nM[j], bc(cMX)); }
      else q[j] = __fmul2_rn(nM[j], j < NE ? pMXe[j] : bc(0.f));
    }
    if constexpr (FORM == 2) q[0].x = __fmul_rn(uM, mx0);
    float xprev = __fmaf_rn(uX, xx0, q[0].x);  // row 0
    nX[0].x = uX;
    nX[0].y = xprev;
#pragma unroll
    for (int k = 1; k < R; ++k) {
      // row k = pair (k+1)/2 half (k+1)&1 ; product M[k-1]*pMX[k] = q[k/2] half (k&1)
      const float prod = (k & 1) ? q[k / 2].y : q[k / 2].x;
      xprev = __fmaf_rn(xprev, cXX, prod);
      if ((k + 1) & 1) nX[(k + 1) / 2].y = xprev; else nX[(k + 1) / 2].x = xprev;
    }
    if constexpr (FORM == 2) {
#pragma unroll
      for (int j = 0; j < NP; ++j) P[j] = q[j];
    }
    const float nbM = BH ? nM[BJ].y : nM[BJ].x, nbX = BH ? nX[BJ].y : nX[BJ].x;
    acc = __fadd_rn(acc, __fadd_rn(nbM, nbX));
#pragma unroll
    for (int j = 0; j < NP; ++j) { M[j] = nM[j]; X[j] = nX[j]; Y[j] = nY[j]; }
  }
  float s = acc;
#pragma unroll
  for (int j = 0; j < NP; ++j) s += M[j].x + X[j].x + Y[j].x + M[j].y + X[j].y + Y[j].y;
  out[blockIdx.x * 32 + threadIdx.x] = s;
}

// Two READS per lane group (same haplotype column): everything is a natural pair -- state, per-row parameters,
// and the prior table (interleaved by read), so that all eight operations per cell pair are packed.
template <int R, int FORM, int UNROLL, int MINB, int GW>
__global__ void __launch_bounds__(32, MINB) k2r(float* out, const float* in, float cGM, float cXX, float cMM, float cMX, int nsteps) {
  constexpr int TS = (((2 * R + 3) / 4) | 1) * 4;  // lane stride in floats (2 reads interleaved)
  constexpr int NV = (2 * R + 3) / 4;
  extern __shared__ __align__(16) float tab[];
  uint32_t* hs = reinterpret_cast<uint32_t*>(tab + 5 * 32 * TS);
  for (int i = threadIdx.x; i < 5 * 32 * TS; i += 32) tab[i] = 0.5f + 0.0001f * (i % 977);
  for (int i = threadIdx.x; i < 1024; i += 32) hs[i] = (uint32_t)((i * 7 + 3) % 5) | ((uint32_t)((i * 11 + 1) % 5) << 16);
  __syncwarp();
  const float* tl = tab + threadIdx.x * TS;
  float2 M[R], X[R], Y[R], P[FORM == 2 ? R : 1];
  float2 pMM[FORM == 2 ? 1 : R], pMX[FORM == 2 ? 1 : R], pMY[FORM == 2 ? 1 : R], pYY[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    M[k] = make_float2(in[threadIdx.x + k], in[threadIdx.x + k + 1]);
    X[k] = make_float2(in[32 + threadIdx.x + k], in[33 + threadIdx.x + k]);
    Y[k] = make_float2(in[64 + threadIdx.x + k], in[65 + threadIdx.x + k]);
    pYY[k] = make_float2(in[192 + k] * 0.2f, in[193 + k] * 0.21f);
    if constexpr (FORM == 2) P[k] = make_float2(0.f, 0.f);
    else {
      pMM[k] = make_float2(in[96 + k], in[97 + k] * 0.99f); pMX[k] = make_float2(in[128 + k] * 0.01f, in[129 + k] * 0.011f);
      pMY[k] = make_float2(in[160 + k] * 0.01f, in[161 + k] * 0.011f);
    }
  }
  const float2 xx0 = make_float2(in[250], in[251]), mx0 = make_float2(in[252] * 0.01f, in[253] * 0.01f);
  float2 dM = bc(0.f), dX = bc(0.f), dY = bc(0.f), acc = bc(0.f);
#pragma unroll(UNROLL)
  for (int t = 0; t < nsteps; ++t) {
    const uint32_t h2 = hs[t & 1023];
    const float* prow = tl + (h2 & 0xffffu) * (32 * TS);
    float2 pr[NV * 2];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 f = *reinterpret_cast<const float4*>(prow + v * 4);
      pr[v * 2] = make_float2(f.x, f.y); pr[v * 2 + 1] = make_float2(f.z, f.w);
    }
    float2 uM, uX, uY;
    uM.x = __shfl_up_sync(0xffffffffu, M[R - 1].x, 1, GW); uM.y = __shfl_up_sync(0xffffffffu, M[R - 1].y, 1, GW);
    uX.x = __shfl_up_sync(0xffffffffu, X[R - 1].x, 1, GW); uX.y = __shfl_up_sync(0xffffffffu, X[R - 1].y, 1, GW);
    uY.x = __shfl_up_sync(0xffffffffu, Y[R - 1].x, 1, GW); uY.y = __shfl_up_sync(0xffffffffu, Y[R - 1].y, 1, GW);
    float2 nM[R], nX[R], nY[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const float2 md = k ? M[k - 1] : dM, xd = k ? X[k - 1] : dX, yd = k ? Y[k - 1] : dY;
      float2 s = FORM == 2 ? __fmul2_rn(md, bc(cMM)) : __fmul2_rn(md, pMM[FORM == 2 ? 0 : k]);
      s = __ffma2_rn(xd, bc(cGM), s);
      s = __ffma2_rn(yd, bc(cGM), s);
      nM[k] = __fmul2_rn(s, pr[k]);
      if constexpr (FORM == 2) nY[k] = __ffma2_rn(Y[k], pYY[k], P[k]);
      else nY[k] = __ffma2_rn(Y[k], pYY[k], __fmul2_rn(M[k], pMY[k]));
    }
    nX[0] = __ffma2_rn(uX, xx0, __fmul2_rn(uM, mx0));
    if constexpr (FORM == 2) {
#pragma unroll
      for (int k = 0; k < R; ++k) P[k] = __fmul2_rn(nM[k], bc(cMX));
#pragma unroll
      for (int k = 1; k < R; ++k) nX[k] = __ffma2_rn(nX[k - 1], bc(cXX), P[k - 1]);
    } else {
#pragma unroll
      for (int k = 1; k < R; ++k) nX[k] = __ffma2_rn(nX[k - 1], bc(cXX), __fmul2_rn(nM[k - 1], pMX[k]));
    }
    acc = __fadd2_rn(acc, __fadd2_rn(nM[R - 1], nX[R - 1]));
    dM = uM; dX = uX; dY = uY;
#pragma unroll
    for (int k = 0; k < R; ++k) { M[k] = nM[k]; X[k] = nX[k]; Y[k] = nY[k]; }
  }
  float s = acc.x + acc.y;
#pragma unroll
  for (int k = 0; k < R; ++k) s += M[k].x + X[k].x + Y[k].x + M[k].y + X[k].y + Y[k].y;
  out[blockIdx.x * 32 + threadIdx.x] = s;
}

// All-uniform form with LIGHT packing: only the two per-row operations that need no re-alignment run packed on natural
// (even, odd) row pairs -- Y' = Y * pYY + P (FFMA2) and P' = M' * cMX (FMUL2); the M phase and the X chain stay scalar and
// read / write the halves.  6.67 issue slots per cell instead of 7.67 (the scalar all-uniform loop is issue bound).
template <int R, int FORM, int UNROLL, int MINB, int GW>
__global__ void __launch_bounds__(32, MINB) kual(float* out, const float* in, float cGM, float cXX, float cMM, float cMX, int nsteps) {
  static_assert(R % 2 == 0, "even rows per lane");
  constexpr int TS = (((R + 3) / 4) | 1) * 4;
  constexpr int NV = (R + 3) / 4;
  constexpr int H = R / 2;
  extern __shared__ __align__(16) float tab[];
  uint32_t* hs = reinterpret_cast<uint32_t*>(tab + 5 * 32 * TS);
  for (int i = threadIdx.x; i < 5 * 32 * TS; i += 32) tab[i] = 0.5f + 0.0001f * (i % 977);
  for (int i = threadIdx.x; i < 1024; i += 32) hs[i] = (uint32_t)((i * 7 + 3) % 5) | ((uint32_t)((i * 11 + 1) % 5) << 16);
  __syncwarp();
  const float* tl = tab + threadIdx.x * TS;
  float2 M[H], Y[H], P[H], pYY[H];
  float X[R];
#pragma unroll
  for (int j = 0; j < H; ++j) {
    M[j] = make_float2(in[threadIdx.x + j], in[threadIdx.x + j + 1]);
    Y[j] = make_float2(in[64 + threadIdx.x + j], in[65 + threadIdx.x + j]);
    P[j] = make_float2(0.f, 0.f);
    pYY[j] = make_float2(in[192 + j] * 0.2f, in[193 + j] * 0.21f);
    X[2 * j] = in[32 + threadIdx.x + j]; X[2 * j + 1] = in[33 + threadIdx.x + j];
  }
  const float xx0 = in[250], mx0 = in[251] * 0.01f;
  float dM = 0, dX = 0, dY = 0, acc = 0;
  auto half = [](const float2& v, int k) { return (k & 1) ? v.y : v.x; };
#pragma unroll(UNROLL)
  for (int t = 0; t < nsteps; ++t) {
    const uint32_t h2 = hs[t & 1023];
    const float* prow = tl + (h2 & 0xffffu) * (32 * TS);
    float pr[NV * 4];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 f = *reinterpret_cast<const float4*>(prow + v * 4);
      pr[v * 4] = f.x; pr[v * 4 + 1] = f.y; pr[v * 4 + 2] = f.z; pr[v * 4 + 3] = f.w;
    }
    const float uM = __shfl_up_sync(0xffffffffu, M[H - 1].y, 1, GW), uX = __shfl_up_sync(0xffffffffu, X[R - 1], 1, GW),
                uY = __shfl_up_sync(0xffffffffu, Y[H - 1].y, 1, GW);
    float2 nM[H], nY[H];
    float nX[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const float md = k ? half(M[(k - 1) / 2], k - 1) : dM, xd = k ? X[k - 1] : dX, yd = k ? half(Y[(k - 1) / 2], k - 1) : dY;
      float s = __fmul_rn(md, cMM);
      s = __fmaf_rn(xd, cGM, s);
      s = __fmaf_rn(yd, cGM, s);
      const float m = __fmul_rn(s, pr[k]);
      if (k & 1) nM[k / 2].y = m; else nM[k / 2].x = m;
    }
#pragma unroll
    for (int j = 0; j < H; ++j) nY[j] = __ffma2_rn(Y[j], pYY[j], P[j]);
    nX[0] = __fmaf_rn(uX, xx0, __fmul_rn(uM, mx0));
#pragma unroll
    for (int j = 0; j < H; ++j) P[j] = __fmul2_rn(nM[j], bc(cMX));
#pragma unroll
    for (int k = 1; k < R; ++k) nX[k] = __fmaf_rn(nX[k - 1], cXX, half(P[(k - 1) / 2], k - 1));
    acc = __fadd_rn(acc, __fadd_rn(nM[H - 1].y, nX[R - 1]));
    dM = uM; dX = uX; dY = uY;
#pragma unroll
    for (int j = 0; j < H; ++j) { M[j] = nM[j]; Y[j] = nY[j]; }
#pragma unroll
    for (int k = 0; k < R; ++k) X[k] = nX[k];
  }
  float s = acc;
#pragma unroll
  for (int j = 0; j < H; ++j) s += M[j].x + M[j].y + Y[j].x + Y[j].y + X[2 * j] + X[2 * j + 1];
  out[blockIdx.x * 32 + threadIdx.x] = s;
}

static int g_sel = -1, g_idx = 0;
template <int R, int FORM, int UNROLL, int MINB, int GW, int MODE>
void run(const char* name, int ctas_per_sm, float* out, float* in) {
  if (!(g_sel < 0 || g_sel == g_idx++)) return;
  constexpr int TS = MODE == 3 ? (((2 * R + 3) / 4) | 1) * 4 : (((R + 3) / 4) | 1) * 4;
  const size_t smem = (size_t)5 * 32 * TS * 4 + 1024 * 4;
  const int grid = 148 * ctas_per_sm;
  constexpr bool PACKED = MODE == 1 || MODE == 3;
  void (*kern)(float*, const float*, float, float, float, float, int);
  if constexpr (MODE == 2) kern = krp<R, FORM, UNROLL, MINB, GW>; else if constexpr (MODE == 3) kern = k2r<R, FORM, UNROLL, MINB, GW>; else if constexpr (MODE == 4) kern = kual<R, FORM, UNROLL, MINB, GW>; else kern = k<R, FORM, UNROLL, MINB, GW, MODE == 1>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, smem);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  kern<<<grid, 32, smem>>>(out, in, 0.9f, 0.1f, 0.9998f, 3e-5f, STEPS);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<grid, 32, smem>>>(out, in, 0.9f, 0.1f, 0.9998f, 3e-5f, STEPS);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double per = PACKED ? 2.0 : 1.0;
  const double cells = (double)grid * 32 * R * STEPS * per;
  const double clk = 1.965e9;
  const double cyc_per_cell = ms * 1e-3 * clk / ((double)ctas_per_sm / 4.0 * R * STEPS * per);
  printf("%-30s form %d R=%2d unroll=%d CTAs/SM=%2d (occupancy %2d, %3d regs, %zu B local)  %.3f ms  %.0f GCUPS-eq  %.2f cyc/warp-cell %s\n", name, FORM, R, UNROLL,
         ctas_per_sm, occ, fa.numRegs, (size_t)fa.localSizeBytes, ms, cells / (ms * 1e-3) / 1e9, cyc_per_cell, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main(int argc, char** argv) {
  if (argc > 1) g_sel = atoi(argv[1]);
  float *out, *in;
  cudaMalloc(&out, 148 * 16 * 32 * 4);
  cudaMalloc(&in, 4096);
  float h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = 0.5f + 0.0003f * i;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  // scalar baselines
  run<19, 1, 4, 12, 8, 0>("scalar UG G=8", 12, out, in);
  run<19, 1, 2, 12, 8, 0>("scalar UG G=8", 12, out, in);
  run<38, 2, 2, 8, 4, 0>("scalar UA G=4", 8, out, in);
  // packed, two haplotypes per lane
  run<19, 1, 2, 8, 8, 1>("packed UG x2 haps G=8", 8, out, in);
  run<19, 1, 1, 8, 8, 1>("packed UG x2 haps G=8", 8, out, in);
  run<19, 1, 4, 8, 8, 1>("packed UG x2 haps G=8", 8, out, in);
  run<16, 1, 2, 8, 16, 1>("packed UG x2 haps G=16", 8, out, in);
  run<16, 1, 2, 12, 16, 1>("packed UG x2 haps G=16", 12, out, in);
  run<13, 1, 2, 12, 8, 1>("packed UG x2 haps G=8", 12, out, in);
  run<13, 1, 4, 12, 8, 1>("packed UG x2 haps G=8", 12, out, in);
  run<10, 1, 4, 16, 16, 1>("packed UG x2 haps G=16", 16, out, in);
  run<19, 2, 2, 8, 8, 1>("packed UA x2 haps G=8", 8, out, in);
  run<19, 2, 2, 12, 8, 1>("packed UA x2 haps G=8", 12, out, in);
  run<19, 2, 4, 8, 8, 1>("packed UA x2 haps G=8", 8, out, in);
  run<24, 2, 2, 8, 8, 1>("packed UA x2 haps G=8", 8, out, in);
  run<19, 1, 2, 12, 8, 2>("row-pair UG G=8", 12, out, in);
  run<19, 1, 4, 12, 8, 2>("row-pair UG G=8", 12, out, in);
  run<19, 1, 2, 8, 8, 2>("row-pair UG G=8", 8, out, in);
  run<16, 1, 2, 12, 16, 2>("row-pair UG G=16", 12, out, in);
  run<24, 1, 2, 8, 8, 2>("row-pair UG G=8", 8, out, in);
  run<38, 2, 2, 8, 4, 2>("row-pair UA G=4", 8, out, in);
  run<38, 2, 1, 8, 4, 2>("row-pair UA G=4", 8, out, in);
  run<19, 2, 2, 12, 8, 2>("row-pair UA G=8", 12, out, in);
  run<13, 1, 2, 16, 8, 2>("row-pair UG G=8", 16, out, in);
  run<10, 1, 2, 12, 16, 3>("two reads UG G=16", 12, out, in);
  run<10, 1, 4, 12, 16, 3>("two reads UG G=16", 12, out, in);
  run<10, 1, 2, 16, 16, 3>("two reads UG G=16", 16, out, in);
  run<12, 1, 2, 12, 16, 3>("two reads UG G=16", 12, out, in);
  run<16, 1, 2, 8, 16, 3>("two reads UG G=16", 8, out, in);
  run<19, 2, 2, 8, 8, 3>("two reads UA G=8", 8, out, in);
  run<19, 2, 2, 12, 8, 3>("two reads UA G=8", 12, out, in);
  run<19, 2, 1, 8, 8, 3>("two reads UA G=8", 8, out, in);
  run<10, 2, 2, 16, 16, 3>("two reads UA G=16", 16, out, in);
  run<38, 2, 2, 8, 4, 4>("light-packed UA G=4", 8, out, in);
  run<38, 2, 1, 8, 4, 4>("light-packed UA G=4", 8, out, in);
  run<38, 2, 4, 8, 4, 4>("light-packed UA G=4", 8, out, in);
  run<32, 2, 2, 8, 4, 4>("light-packed UA G=4", 8, out, in);
  run<20, 2, 4, 12, 8, 4>("light-packed UA G=8", 12, out, in);
  run<32, 2, 2, 8, 4, 0>("scalar UA G=4", 8, out, in);
  run<20, 2, 4, 12, 8, 0>("scalar UA G=8", 12, out, in);
  return 0;
}
