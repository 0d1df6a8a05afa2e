// Microbenchmark: cycles per DP cell of the PairHMM tile step in isolation (uniform-GCP form),
// one column per step (shape A, what the kernel does) vs two columns per step (shape B: the two
// columns of a row share the row's constants, so every second instruction can take one operand
// from the operand-reuse cache).  Developer tool.
#include <algorithm>
#include <cstdio>
#include <cuda_runtime.h>

#define STEPS 4096

template <int R, int SHAPE, int UNROLL, int MINB, int GW>
__global__ void __launch_bounds__(32, MINB) k(float* out, const float* in, long long* cyc, float cGM, float cXX, float cMM, float cMX, float cMY, int nsteps) {
  constexpr int TS = (((R + 3) / 4) | 1) * 4;  // lane stride in floats: odd multiple of 16 B
  __shared__ __align__(16) float tab[4 * 32 * TS];
  for (int i = threadIdx.x; i < 4 * 32 * TS; i += 32) tab[i] = 0.5f + 0.0001f * i;
  __syncwarp();
  float M[R], X[R], Y[R], pMM[R], pMX[R], pMY[R], pYY[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    M[k] = in[threadIdx.x + k]; X[k] = in[32 + threadIdx.x + k]; Y[k] = in[64 + threadIdx.x + k];
    pMM[k] = in[96 + k]; pMX[k] = in[128 + k] * 0.01f; pMY[k] = in[160 + k] * 0.01f; pYY[k] = 0.1f + 0.001f * k;
  }
  float dM = 0, dX = 0, dY = 0, acc = 0, b0M = 0, b0X = 0, b0Y = 0;
  const float* tl = tab + threadIdx.x * TS;
  long long t0 = clock64();
  if (SHAPE == 0) {
#pragma unroll(UNROLL)
    for (int t = 0; t < nsteps; ++t) {
      const float* prow = tl + (t & 3) * (32 * TS);
      float pr[R];
#pragma unroll
      for (int v = 0; v < (R + 3) / 4; ++v) {
        float4 f = *reinterpret_cast<const float4*>(prow + v * 4);
        pr[v * 4] = f.x; if (v * 4 + 1 < R) pr[v * 4 + 1] = f.y; if (v * 4 + 2 < R) pr[v * 4 + 2] = f.z; if (v * 4 + 3 < R) pr[v * 4 + 3] = f.w;
      }
      const float uM = __shfl_up_sync(0xffffffffu, M[R - 1], 1, GW), uX = __shfl_up_sync(0xffffffffu, X[R - 1], 1, GW),
                  uY = __shfl_up_sync(0xffffffffu, Y[R - 1], 1, GW);
      float nM[R], nX[R], nY[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const float md = k ? M[k - 1] : dM, xd = k ? X[k - 1] : dX, yd = k ? Y[k - 1] : dY;
        float s = __fmul_rn(md, pMM[k]);
        s = __fmaf_rn(xd, cGM, s);
        s = __fmaf_rn(yd, cGM, s);
        nM[k] = __fmul_rn(s, pr[k]);
        nY[k] = __fmaf_rn(Y[k], pYY[k], __fmul_rn(M[k], pMY[k]));
      }
      nX[0] = __fmaf_rn(uX, cXX, __fmul_rn(uM, pMX[0]));
#pragma unroll
      for (int k = 1; k < R; ++k) nX[k] = __fmaf_rn(nX[k - 1], cXX, __fmul_rn(nM[k - 1], pMX[k]));
      acc = __fadd_rn(acc, __fadd_rn(nM[R - 1], nX[R - 1]));
      dM = uM; dX = uX; dY = uY;
#pragma unroll
      for (int k = 0; k < R; ++k) { M[k] = nM[k]; X[k] = nX[k]; Y[k] = nY[k]; }
    }
  } else if (SHAPE == 2) {
#pragma unroll(UNROLL)
    for (int t = 0; t < nsteps; ++t) {
      const float* prow = tl + (t & 3) * (32 * TS);
      float pr[R + 3];
#pragma unroll
      for (int v = 0; v < (R + 3) / 4; ++v) {
        float4 f = *reinterpret_cast<const float4*>(prow + v * 4);
        pr[v * 4] = f.x; pr[v * 4 + 1] = f.y; pr[v * 4 + 2] = f.z; pr[v * 4 + 3] = f.w;
      }
      const float uM = __shfl_up_sync(0xffffffffu, M[R - 1], 1, GW), uX = __shfl_up_sync(0xffffffffu, X[R - 1], 1, GW),
                  uY = __shfl_up_sync(0xffffffffu, Y[R - 1], 1, GW);
      float nM[R], nX[R], nY[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const float md = k ? M[k - 1] : dM, xd = k ? X[k - 1] : dX, yd = k ? Y[k - 1] : dY;
        float s = __fmul_rn(md, cMM);
        s = __fmaf_rn(xd, cGM, s);
        s = __fmaf_rn(yd, cGM, s);
        nM[k] = __fmul_rn(s, pr[k]);
        nY[k] = __fmaf_rn(Y[k], pYY[k], __fmul_rn(M[k], cMY));
      }
      nX[0] = __fmaf_rn(uX, pMM[0], __fmul_rn(uM, pMX[0]));
#pragma unroll
      for (int k = 1; k < R; ++k) nX[k] = __fmaf_rn(nX[k - 1], cXX, __fmul_rn(nM[k - 1], cMX));
      acc = __fadd_rn(acc, __fadd_rn(nM[R - 1], nX[R - 1]));
      dM = uM; dX = uX; dY = uY;
#pragma unroll
      for (int k = 0; k < R; ++k) { M[k] = nM[k]; X[k] = nX[k]; Y[k] = nY[k]; }
    }
  } else {
#pragma unroll(UNROLL)
    for (int t = 0; t < nsteps; t += 2) {
      const float* prow0 = tl + (t & 3) * (32 * TS);
      const float* prow1 = tl + ((t + 1) & 3) * (32 * TS);
      float pr0[R], pr1[R];
#pragma unroll
      for (int v = 0; v < (R + 3) / 4; ++v) {
        float4 f = *reinterpret_cast<const float4*>(prow0 + v * 4);
        pr0[v * 4] = f.x; if (v * 4 + 1 < R) pr0[v * 4 + 1] = f.y; if (v * 4 + 2 < R) pr0[v * 4 + 2] = f.z; if (v * 4 + 3 < R) pr0[v * 4 + 3] = f.w;
        float4 g = *reinterpret_cast<const float4*>(prow1 + v * 4);
        pr1[v * 4] = g.x; if (v * 4 + 1 < R) pr1[v * 4 + 1] = g.y; if (v * 4 + 2 < R) pr1[v * 4 + 2] = g.z; if (v * 4 + 3 < R) pr1[v * 4 + 3] = g.w;
      }
      // lane above: bottom row at column c0 (saved last step) and at column c1 (its current state)
      const float u0M = __shfl_up_sync(0xffffffffu, b0M, 1, GW), u0X = __shfl_up_sync(0xffffffffu, b0X, 1, GW), u0Y = __shfl_up_sync(0xffffffffu, b0Y, 1, GW);
      const float u1M = __shfl_up_sync(0xffffffffu, M[R - 1], 1, GW), u1X = __shfl_up_sync(0xffffffffu, X[R - 1], 1, GW), u1Y = __shfl_up_sync(0xffffffffu, Y[R - 1], 1, GW);
      float aM[R], aX[R], aY[R], bM[R], bX[R], bY[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const float md0 = k ? M[k - 1] : dM, xd0 = k ? X[k - 1] : dX, yd0 = k ? Y[k - 1] : dY;
        const float md1 = k ? aM[k - 1] : u0M, xd1 = k ? aX[k - 1] : u0X, yd1 = k ? aY[k - 1] : u0Y;
        float s0 = __fmul_rn(md0, pMM[k]);
        float s1 = __fmul_rn(md1, pMM[k]);
        s0 = __fmaf_rn(xd0, cGM, s0);
        s1 = __fmaf_rn(xd1, cGM, s1);
        s0 = __fmaf_rn(yd0, cGM, s0);
        s1 = __fmaf_rn(yd1, cGM, s1);
        aM[k] = __fmul_rn(s0, pr0[k]);
        bM[k] = __fmul_rn(s1, pr1[k]);
        aY[k] = __fmaf_rn(Y[k], pYY[k], __fmul_rn(M[k], pMY[k]));
        bY[k] = __fmaf_rn(aY[k], pYY[k], __fmul_rn(aM[k], pMY[k]));
        const float um0 = k ? aM[k - 1] : u0M, ux0 = k ? aX[k - 1] : u0X;
        const float um1 = k ? bM[k - 1] : u1M, ux1 = k ? bX[k - 1] : u1X;
        aX[k] = __fmaf_rn(ux0, cXX, __fmul_rn(um0, pMX[k]));
        bX[k] = __fmaf_rn(ux1, cXX, __fmul_rn(um1, pMX[k]));
      }
      acc = __fadd_rn(acc, __fadd_rn(aM[R - 1], aX[R - 1]));
      acc = __fadd_rn(acc, __fadd_rn(bM[R - 1], bX[R - 1]));
      dM = u1M; dX = u1X; dY = u1Y;
      b0M = aM[R - 1]; b0X = aX[R - 1]; b0Y = aY[R - 1];
#pragma unroll
      for (int k = 0; k < R; ++k) { M[k] = bM[k]; X[k] = bX[k]; Y[k] = bY[k]; }
    }
  }
  long long t1 = clock64();
  float s = acc;
#pragma unroll
  for (int k = 0; k < R; ++k) s += M[k] + X[k] + Y[k];
  out[blockIdx.x * 32 + threadIdx.x] = s;
  if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = t0; cyc[2 * blockIdx.x + 1] = t1; }
}

template <int R, int SHAPE, int UNROLL, int MINB, int GW>
void run(const char* name, int ctas_per_sm, float* out, float* in, long long* cyc) {
  static long long h[148 * 16 * 2];
  const int grid = 148 * ctas_per_sm;
  k<R, SHAPE, UNROLL, MINB, GW><<<grid, 32>>>(out, in, cyc, 0.9f, 0.1f, 0.9998f, 3e-5f, 3e-5f, STEPS);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<R, SHAPE, UNROLL, MINB, GW><<<grid, 32>>>(out, in, cyc, 0.9f, 0.1f, 0.9998f, 3e-5f, 3e-5f, STEPS);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cells = (double)grid * 32 * R * STEPS;
  // cycles per warp-cell on one sub-partition: (grid/148/4 warps per SMSP) x R x STEPS warp-cells share the time
  const double clk = 1.965e9;
  const double cyc_per_cell = ms * 1e-3 * clk / ((double)ctas_per_sm / 4.0 * R * STEPS);
  printf("%-26s R=%2d unroll=%d CTAs/SM=%2d  %.3f ms  %.0f GCUPS-equivalent  %.2f cycles per warp-cell (8.0 = FMA-pipe peak) %s\n", name, R, UNROLL,
         ctas_per_sm, ms, cells / (ms * 1e-3) / 1e9, cyc_per_cell, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  float *out, *in;
  long long* cyc;
  cudaMalloc(&out, 148 * 16 * 32 * 4);
  cudaMalloc(&in, 4096);
  cudaMalloc(&cyc, 148 * 16 * 16);
  float h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = 0.5f + 0.0003f * i;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<19, 0, 4, 12, 8>("A: general UG, G=8", 12, out, in, cyc);
  run<16, 2, 4, 16, 4>("C: all-uniform, G=4", 16, out, in, cyc);
  run<20, 2, 4, 16, 4>("C: all-uniform, G=4", 16, out, in, cyc);
  run<20, 2, 4, 12, 4>("C: all-uniform, G=4", 12, out, in, cyc);
  run<24, 2, 4, 16, 4>("C: all-uniform, G=4", 16, out, in, cyc);
  run<24, 2, 4, 12, 4>("C: all-uniform, G=4", 12, out, in, cyc);
  run<26, 2, 4, 16, 4>("C: all-uniform, G=4", 16, out, in, cyc);
  run<26, 2, 4, 12, 4>("C: all-uniform, G=4", 12, out, in, cyc);
  run<26, 2, 4, 8, 4>("C: all-uniform, G=4", 8, out, in, cyc);
  run<28, 2, 4, 12, 4>("C: all-uniform, G=4", 12, out, in, cyc);
  run<28, 2, 4, 8, 4>("C: all-uniform, G=4", 8, out, in, cyc);
  run<32, 2, 4, 12, 4>("C: all-uniform, G=4", 12, out, in, cyc);
  run<32, 2, 4, 8, 4>("C: all-uniform, G=4", 8, out, in, cyc);
  run<36, 2, 4, 12, 4>("C: all-uniform, G=4", 12, out, in, cyc);
  run<36, 2, 4, 8, 4>("C: all-uniform, G=4", 8, out, in, cyc);
  run<38, 2, 4, 12, 4>("C: all-uniform, G=4", 12, out, in, cyc);
  run<38, 2, 4, 8, 4>("C: all-uniform, G=4", 8, out, in, cyc);
  run<40, 2, 4, 8, 4>("C: all-uniform, G=4", 8, out, in, cyc);
  run<16, 2, 4, 16, 8>("C: all-uniform, G=8", 16, out, in, cyc);
  run<20, 2, 4, 16, 8>("C: all-uniform, G=8", 16, out, in, cyc);
  run<20, 2, 4, 12, 8>("C: all-uniform, G=8", 12, out, in, cyc);
  run<24, 2, 4, 16, 8>("C: all-uniform, G=8", 16, out, in, cyc);
  run<24, 2, 4, 12, 8>("C: all-uniform, G=8", 12, out, in, cyc);
  run<26, 2, 4, 16, 8>("C: all-uniform, G=8", 16, out, in, cyc);
  run<26, 2, 4, 12, 8>("C: all-uniform, G=8", 12, out, in, cyc);
  run<26, 2, 4, 8, 8>("C: all-uniform, G=8", 8, out, in, cyc);
  run<28, 2, 4, 12, 8>("C: all-uniform, G=8", 12, out, in, cyc);
  run<28, 2, 4, 8, 8>("C: all-uniform, G=8", 8, out, in, cyc);
  run<32, 2, 4, 12, 8>("C: all-uniform, G=8", 12, out, in, cyc);
  run<32, 2, 4, 8, 8>("C: all-uniform, G=8", 8, out, in, cyc);
  run<36, 2, 4, 12, 8>("C: all-uniform, G=8", 12, out, in, cyc);
  run<36, 2, 4, 8, 8>("C: all-uniform, G=8", 8, out, in, cyc);
  run<38, 2, 4, 12, 8>("C: all-uniform, G=8", 12, out, in, cyc);
  run<38, 2, 4, 8, 8>("C: all-uniform, G=8", 8, out, in, cyc);
  run<40, 2, 4, 8, 8>("C: all-uniform, G=8", 8, out, in, cyc);
  return 0;
}
