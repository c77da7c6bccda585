// Microbenchmark: tcgen05.ld throughput (TMEM -> registers) and MUFU.TANH throughput per SM, as the epilogue warps of the
// fused kernels use them: NW warps (4 per TMEM lane quadrant... NW/4 per quadrant) loop over 32-column pieces.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/tmem_bench tools/micro/tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>  // 0: ld only, 1: ld + 32 tanh per piece, 2: tanh only (no ld), 3: ld + 32 FFMA-only "epilogue"
__global__ void __launch_bounds__(1024, 1) bench(int iters, unsigned long long* out, float* sink) {
  __shared__ uint32_t tmem_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
  float acc = 0.0f;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = (float)(lane + j) * 1e-3f;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE != 2) {
      ld32(tmem + lane_sel + (uint32_t)(((warp >> 2) * 32 + i * 128) & 511), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { const float h = 0.5f * v[j]; v[j] = fmaf(h, tanh_approx(h), h); }
    }
    if (MODE == 3) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], 1.0001f, 0.5f);
    }
    if (MODE != 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += v[j];
    } else {
      acc += v[i & 31];
    }
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
  if (acc == 123.456f) *sink = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  unsigned long long* d;
  float* sink;
  cudaMalloc(&d, 8);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  const char* names[4] = {"tcgen05.ld x32 only", "ld + 32 swish (MUFU.TANH)", "32 swish only", "ld + 32 FFMA"};
  for (int mode = 0; mode < 4; ++mode)
    for (int nw = 4; nw <= 32; nw *= 2) {
      unsigned long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) bench<0><<<148, nw * 32>>>(iters, d, sink);
        if (mode == 1) bench<1><<<148, nw * 32>>>(iters, d, sink);
        if (mode == 2) bench<2><<<148, nw * 32>>>(iters, d, sink);
        if (mode == 3) bench<3><<<148, nw * 32>>>(iters, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double per_piece = (double)h / iters;  // cycles per iteration = nw pieces of 32x32 fp32 (4 KB each)
      printf("%-28s warps %2d: %7.1f cycles per round of %2d pieces -> %6.1f B/clk/SM TMEM read, %5.2f elements/clk/SM\n", names[mode], nw,
             per_piece, nw, mode == 2 ? 0.0 : nw * 4096.0 / per_piece, nw * 1024.0 / per_piece);
    }
  return 0;
}
