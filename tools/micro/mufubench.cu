// Microbenchmark: throughput of the activation math of the fused epilogues, per SM.
//   mufubench [warps per CTA]   -- one CTA per SM; prints cycles per 32-value warp-pass for several variants
// variants: 0 tanh only | 1 fma, tanh, fma (Swish as in the kernels) | 2 ex2 only | 3 ex2 + rcp (sigmoid the long way)
//           4 packed: fma.f32x2, 2 tanh, fma.f32x2 | 5 rsqrt only
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ float tanh_approx(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsq_approx(float x) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int V>
__global__ void __launch_bounds__(1024, 1) k(int iters, float seed, float* out, long long* cyc) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = seed * (float)(j + threadIdx.x);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (V == 4) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        unsigned long long h, o, half = 0x3f0000003f000000ull, b = 0x3dcccccd3dcccccdull, vv;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(vv) : "f"(v[j]), "f"(v[j + 1]));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(h) : "l"(vv), "l"(half), "l"(b));
        float h0, h1;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(h0), "=f"(h1) : "l"(h));
        float t0_ = tanh_approx(h0), t1_ = tanh_approx(h1);
        unsigned long long tt;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(tt) : "f"(t0_), "f"(t1_));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(o) : "l"(h), "l"(tt));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(v[j]), "=f"(v[j + 1]) : "l"(o));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (V == 0) v[j] = tanh_approx(v[j]);
        if (V == 1) { const float h = fmaf(v[j], 0.5f, 0.1f); v[j] = fmaf(h, tanh_approx(h), h); }
        if (V == 2) v[j] = ex2_approx(v[j]);
        if (V == 3) v[j] = v[j] * rcp_approx(1.0f + ex2_approx(-1.442695f * v[j]));
        if (V == 5) v[j] = rsq_approx(v[j]);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main(int argc, char** argv) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024); cudaMalloc(&cyc, sizeof(long long) * sms);
  const int iters = 200;
  const char* names[6] = {"tanh", "fma tanh fma", "ex2", "ex2 rcp mul fma", "fma2 2tanh fma2", "rsqrt"};
  for (int w = 4; w <= 32; w *= 2) {
    if (argc > 1 && atoi(argv[1]) != w) continue;
    for (int v = 0; v < 6; ++v) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (v) {
          case 0: k<0><<<sms, w * 32>>>(iters, 0.001f, out, cyc); break;
          case 1: k<1><<<sms, w * 32>>>(iters, 0.001f, out, cyc); break;
          case 2: k<2><<<sms, w * 32>>>(iters, 0.001f, out, cyc); break;
          case 3: k<3><<<sms, w * 32>>>(iters, 0.001f, out, cyc); break;
          case 4: k<4><<<sms, w * 32>>>(iters, 0.001f, out, cyc); break;
          case 5: k<5><<<sms, w * 32>>>(iters, 0.001f, out, cyc); break;
        }
        cudaDeviceSynchronize();
      }
      long long c;
      cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
      // elements per SM = w * 32 lanes * 32 values * iters
      printf("warps %2d  %-18s  %8.1f cycles per 32-value pass of all warps, %6.2f elements / clk / SM\n", w, names[v], (double)c / iters,
             (double)w * 32 * 32 * iters / (double)c);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
