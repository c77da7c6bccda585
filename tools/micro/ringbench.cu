// Microbenchmark: throughput of a cp.async.bulk + mbarrier weight ring as the fused kernels use it.
// One CTA per SM; one producer thread and one consumer thread per CTA.  The consumer emulates the MMA issuer: waits for
// the step's full barrier, "works" for W cycles, then releases the slots (plain mbarrier arrive).
//   ringbench <bytes per copy> <copies per step> <ring slots (8 KB each)> <work cycles> <steps> [distinct: 0 same image for all CTAs, 1 per-CTA image]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__global__ void __launch_bounds__(64, 1) ring(const uint8_t* img, size_t img_bytes, int distinct, uint32_t bytes, int copies, int slots, int work, int steps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[32], empty_bar[32];
  const int tid = threadIdx.x;
  if (tid == 0) { for (int i = 0; i < 32; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const uint8_t* base = img + (distinct ? (size_t)blockIdx.x * (img_bytes / gridDim.x) : 0);
  const size_t span = distinct ? img_bytes / gridDim.x : img_bytes;
  const int nsteps_ring = slots / copies;  // steps that fit the ring
  if (tid == 0) {  // producer
    size_t off = (size_t)(blockIdx.x % 8) * 65536 % span;
    for (int st = 0; st < steps; ++st) {
      const int r = st % nsteps_ring, use = st / nsteps_ring;
      if (use > 0) for (int u = 0; u < copies; ++u) mbar_wait(&empty_bar[r * copies + u], (use - 1) & 1);
      mbar_expect(&full_bar[r * copies], bytes * copies);
      for (int u = 0; u < copies; ++u) {
        bulk(smem + (size_t)(r * copies + u) * 8192, base + off, bytes, &full_bar[r * copies]);
        off += 8192; if (off + 8192 > span) off = 0;
      }
    }
  } else if (tid == 32) {  // consumer
    const long long t0 = clock64();
    long long waited = 0;
    for (int st = 0; st < steps; ++st) {
      const int r = st % nsteps_ring, use = st / nsteps_ring;
      const long long c0 = clock64();
      mbar_wait(&full_bar[r * copies], use & 1);
      const long long c1 = clock64();
      waited += c1 - c0;
      while (clock64() - c1 < work) {}
      for (int u = 0; u < copies; ++u) mbar_arrive(&empty_bar[r * copies + u]);
    }
    out[blockIdx.x * 2] = clock64() - t0;
    out[blockIdx.x * 2 + 1] = waited;
  }
}
int main(int argc, char** argv) {
  const uint32_t bytes = argc > 1 ? atoi(argv[1]) : 8192;
  const int copies = argc > 2 ? atoi(argv[2]) : 2, slots = argc > 3 ? atoi(argv[3]) : 16, work = argc > 4 ? atoi(argv[4]) : 500;
  const int steps = argc > 5 ? atoi(argv[5]) : 480, distinct = argc > 6 ? atoi(argv[6]) : 0;
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t img_bytes = distinct ? (size_t)sms * (1 << 20) : (1 << 20);
  uint8_t* img; cudaMalloc(&img, img_bytes); cudaMemset(img, 1, img_bytes);
  long long* out; cudaMalloc(&out, sms * 16);
  const size_t smem = (size_t)slots * 8192 + 1024;
  cudaFuncSetAttribute(ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 3; ++rep) ring<<<sms, 64, smem>>>(img, img_bytes, distinct, bytes, copies, slots, work, steps, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long* h = (long long*)malloc(sms * 16); cudaMemcpy(h, out, sms * 16, cudaMemcpyDeviceToHost);
  double tot = 0, wt = 0; for (int i = 0; i < sms; ++i) { tot += h[2 * i]; wt += h[2 * i + 1]; }
  printf("bytes/copy %u copies/step %d slots %d work %d distinct %d : %.0f cycles/step (waited %.0f/step), %.1f B/clk/SM  [%s]\n", bytes, copies, slots, work, distinct,
         tot / sms / steps, wt / sms / steps, (double)bytes * copies / (tot / sms / steps), cudaGetErrorString(e));
  return 0;
}
