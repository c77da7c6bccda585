// Microbenchmark: issue rate of the CUDA-core FMA forms a depthwise convolution can be written in, per SM.
//   fmabench    -- one CTA per SM, 4 / 8 / 12 / 16 warps; prints cycles per warp-instruction per scheduler
// variants: 0 fma.rn.f32 | 1 fma.rn.f32x2 (FFMA2, 64-bit operands) | 2 fma.rn.f16x2 (HFMA2) | 3 fma.rn.bf16x2
// Pattern as in the convolution: 8 independent accumulators, a[o] = w[j] * x + a[o] with x shared by 8 consecutive instructions
// and w[j] from a register array of 31 taps.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
template <int V>
__global__ void __launch_bounds__(512, 1) k(int iters, float seed, float* out, long long* cyc) {
  long long t0, t1;
  float s = 0.0f;
  if (V == 0) {
    float w[31], a[8], x = seed;
#pragma unroll
    for (int j = 0; j < 31; ++j) w[j] = seed * (float)(j + 1);
#pragma unroll
    for (int o = 0; o < 8; ++o) a[o] = seed * (float)(threadIdx.x + o);
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 31; ++j) {
#pragma unroll
        for (int o = 0; o < 8; ++o) a[o] = fmaf(w[(j + o) % 31], x, a[o]);
        x += 1.0f;
      }
    }
    t1 = clock64();
#pragma unroll
    for (int o = 0; o < 8; ++o) s += a[o];
  } else if (V == 1) {
    unsigned long long w[31], a[8], x;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(x) : "f"(seed));
#pragma unroll
    for (int j = 0; j < 31; ++j) { float f = seed * (float)(j + 1); asm volatile("mov.b64 %0, {%1, %1};" : "=l"(w[j]) : "f"(f)); }
#pragma unroll
    for (int o = 0; o < 8; ++o) { float f = seed * (float)(threadIdx.x + o); asm volatile("mov.b64 %0, {%1, %1};" : "=l"(a[o]) : "f"(f)); }
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 31; ++j) {
#pragma unroll
        for (int o = 0; o < 8; ++o) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[o]) : "l"(w[(j + o) % 31]), "l"(x));
        x += 0x0000000100000001ull;
      }
    }
    t1 = clock64();
#pragma unroll
    for (int o = 0; o < 8; ++o) { float f0, f1; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(f0), "=f"(f1) : "l"(a[o])); s += f0 + f1; }
  } else {
    unsigned w[31], a[8], x = 0x3c003c00u;
#pragma unroll
    for (int j = 0; j < 31; ++j) w[j] = 0x2e662e66u + (unsigned)j;
#pragma unroll
    for (int o = 0; o < 8; ++o) a[o] = threadIdx.x + o;
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 31; ++j) {
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          if (V == 2) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(a[o]) : "r"(w[(j + o) % 31]), "r"(x));
          else asm volatile("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(a[o]) : "r"(w[(j + o) % 31]), "r"(x));
        }
        x += 0x00010001u;
      }
    }
    t1 = clock64();
#pragma unroll
    for (int o = 0; o < 8; ++o) s += (float)a[o];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 512); cudaMalloc(&cyc, sizeof(long long) * sms);
  const int iters = 100;
  const char* names[4] = {"fma.f32", "fma.f32x2", "fma.f16x2", "fma.bf16x2"};
  for (int w = 4; w <= 16; w += 4) {
    for (int v = 0; v < 4; ++v) {
      for (int rep = 0; rep < 2; ++rep) {
        if (v == 0) k<0><<<sms, w * 32>>>(iters, 1e-3f, out, cyc);
        if (v == 1) k<1><<<sms, w * 32>>>(iters, 1e-3f, out, cyc);
        if (v == 2) k<2><<<sms, w * 32>>>(iters, 1e-3f, out, cyc);
        if (v == 3) k<3><<<sms, w * 32>>>(iters, 1e-3f, out, cyc);
        cudaDeviceSynchronize();
      }
      long long c0; cudaMemcpy(&c0, cyc, sizeof(c0), cudaMemcpyDeviceToHost);
      const double instr_per_sched = (double)iters * 248 * (w / 4);
      printf("%2d warps %-10s: %8lld cycles, %.2f cycles per warp-instruction per scheduler, %.1f FMA/clk/SM\n", w, names[v], c0,
             c0 / instr_per_sched, (double)iters * 248 * w * 32 * (v == 0 ? 1 : 2) / c0);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
