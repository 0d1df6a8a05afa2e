// Microbenchmark: FP32 issue rates on sm_100a (per SM sub-partition) for the instruction
// shapes the PairHMM inner loop can be built from.  Developer tool, not part of the library.
#include <cstdio>
#include <algorithm>
#include <cuda_runtime.h>
#define N 16
#define ITERS 2000

template <int V>
__global__ void k(float* out, const float* in, long long* cyc, float kc) {
  float a[N], x[N], y[N];
  float2 a2[N], x2[N], y2[N];
  for (int i = 0; i < N; ++i) {
    a[i] = in[threadIdx.x + i]; x[i] = in[threadIdx.x + 32 + i]; y[i] = in[threadIdx.x + 64 + i];
    a2[i] = make_float2(a[i], a[i] + 1.f); x2[i] = make_float2(x[i], x[i] * 0.5f); y2[i] = make_float2(y[i], y[i] * 0.25f);
  }
  float c = in[200], d = in[201];
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (V == 0) a[i] = __fmaf_rn(x[i], y[i], a[i]);                       // 3 distinct regs
      if (V == 1) a[i] = __fmaf_rn(a[i], c, d);                              // shared c, d (reuse cache)
      if (V == 2) a[i] = __fmul_rn(a[i], x[i]);                              // FMUL 2 regs
      if (V == 3) a2[i] = __ffma2_rn(x2[i], y2[i], a2[i]);                   // FFMA2 3 distinct pairs
      if (V == 4) a2[i] = __fmul2_rn(a2[i], x2[i]);                          // FMUL2
      if (V == 5) { a[i] = __fmaf_rn(x[i], y[i], a[i]); x[i] = __fmul_rn(x[i], y[(i + 1) % N]); }  // FFMA+FMUL mix
      if (V == 6) { a2[i] = __ffma2_rn(x2[i], y2[i], a2[i]); x2[i] = __fmul2_rn(x2[i], y2[(i + 1) % N]); }
      if (V == 7) a[i] = __fmaf_rn(x[i], c, a[i]);                           // 2 distinct + shared c
      if (V == 8) a[i] = __fadd_rn(a[i], x[i]);
      if (V == 11) { a[i] = __fmul_rn(a[i], y[i]); x[i] = __fmul_rn(x[i], y[i]); }      // two FMULs sharing y[i]
      if (V == 12) { a[i] = __fmaf_rn(a[i], y[i], kc); x[i] = __fmaf_rn(x[i], y[i], kc); } // two FFMAs sharing y[i], const addend
      if (V == 13) { float t0 = __fmul_rn(a[i], y[i]); float t1 = __fmul_rn(x[i], y[i]); a[i] = __fmaf_rn(x[i], kc, t0); x[i] = __fmaf_rn(a[i], kc, t1); }
      if (V == 9) a[i] = __fmaf_rn(x[i], kc, a[i]);                          // constant-bank operand
      if (V == 10) { a[i] = __fmaf_rn(x[i], kc, a[i]); x[i] = __fmul_rn(x[i], y[i]); }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < N; ++i) s += a[i] + x[i] + a2[i].x + a2[i].y + x2[i].x + x2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) { cyc[2 * (blockIdx.x * 32 + (threadIdx.x >> 5))] = t0; cyc[2 * (blockIdx.x * 32 + (threadIdx.x >> 5)) + 1] = t1; }
}

template <int V>
void run(const char* name, int inst_per_iter, float* out, float* in, long long* cyc) {
  static long long h[148 * 64];
  for (int warps : {4, 8, 12, 16}) {
    k<V><<<148, warps * 32>>>(out, in, cyc, 0.999f);
    cudaDeviceSynchronize();
    k<V><<<148, warps * 32>>>(out, in, cyc, 0.999f);
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = e2;
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int b = 0; b < 148; ++b) {
      long long lo = h[2 * (b * 32)], hi = h[2 * (b * 32) + 1];
      for (int w = 0; w < warps; ++w) { lo = std::min(lo, h[2 * (b * 32 + w)]); hi = std::max(hi, h[2 * (b * 32 + w) + 1]); }
      worst = std::max(worst, (double)(hi - lo));
    }
    double ipc_smsp = (double)ITERS * inst_per_iter * (warps / 4.0) / worst;
    printf("%-34s warps/SM=%2d  cycles=%8.0f  warp-inst/clk/SMSP=%.3f %s\n", name, warps, worst, ipc_smsp, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  float *out, *in;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&in, 4096);
  cudaMalloc(&cyc, 148 * 64 * 8);
  float h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = 0.5f + 0.001f * i;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("FFMA  r,r,r (3 distinct)", N, out, in, cyc);
  run<7>("FFMA  r,c,r (2 distinct + shared)", N, out, in, cyc);
  run<1>("FFMA  r,c,d (shared c,d)", N, out, in, cyc);
  run<2>("FMUL  r,r", N, out, in, cyc);
  run<8>("FADD  r,r", N, out, in, cyc);
  run<9>("FFMA  r,const,r", N, out, in, cyc);
  run<10>("FFMA r,const,r + FMUL r,r", 2 * N, out, in, cyc);
  run<11>("2x FMUL sharing one operand", 2 * N, out, in, cyc);
  run<12>("2x FFMA r,shared,const", 2 * N, out, in, cyc);
  run<13>("pairs: 2 FMUL shared + 2 FFMA const", 4 * N, out, in, cyc);
  run<3>("FFMA2 rr,rr,rr", N, out, in, cyc);
  run<4>("FMUL2 rr,rr", N, out, in, cyc);
  run<5>("FFMA+FMUL mix", 2 * N, out, in, cyc);
  run<6>("FFMA2+FMUL2 mix", 2 * N, out, in, cyc);
  return 0;
}
