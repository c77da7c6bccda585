// Microbenchmark 2: are the packed half-precision MUFU forms faster per element?  (one CTA per SM, 16 warps)
#include <cuda_runtime.h>
#include <cstdio>
template <int V>
__global__ void __launch_bounds__(1024, 1) k(int iters, unsigned seed, unsigned* out, long long* cyc) {
  unsigned v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = seed + 0x00010001u * (j + threadIdx.x);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (V == 0) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(v[j]));
      if (V == 1) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(v[j]));
      if (V == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[j]));
      if (V == 3) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[j]));
      if (V == 4) { float f = __uint_as_float(v[j]); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f)); v[j] = __float_as_uint(f); }
    }
  }
  const long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s ^= v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* out; long long* cyc;
  cudaMalloc(&out, 4 * sms * 1024); cudaMalloc(&cyc, 8 * sms);
  const int iters = 400, w = 16;
  const char* names[5] = {"tanh.f16x2", "tanh.bf16x2", "ex2.f16x2", "ex2.bf16x2", "tanh.f32 (16 regs)"};
  for (int v = 0; v < 5; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (v) {
        case 0: k<0><<<sms, w * 32>>>(iters, 0x38003800u, out, cyc); break;
        case 1: k<1><<<sms, w * 32>>>(iters, 0x3f003f00u, out, cyc); break;
        case 2: k<2><<<sms, w * 32>>>(iters, 0x38003800u, out, cyc); break;
        case 3: k<3><<<sms, w * 32>>>(iters, 0x3f003f00u, out, cyc); break;
        case 4: k<4><<<sms, w * 32>>>(iters, 0x3f000000u, out, cyc); break;
      }
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double regs = (double)w * 32 * 16 * iters;
    printf("%-20s %8.1f cycles per 16-register pass, %6.2f registers / clk / SM (%s elements / clk / SM: %.2f)\n", names[v], (double)c / iters, regs / c,
           v < 4 ? "x2" : "x1", (v < 4 ? 2 : 1) * regs / c);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
