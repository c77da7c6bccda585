// tcgen05 arm of libsmx, part 1: weight packing and the UMMA self-test GEMM (diagnostic entry point
// smx_debug_tc_gemm) that pins the shared-memory operand layouts, descriptors and TMEM read-back used by
// the fused kernels.
#include "smx_tc.h"
#include "smx_tc_common.cuh"

namespace smx {


static inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

size_t tc_packed_bytes(int N, int K, int NT) {
  return (size_t)round_up(N, NT) * (size_t)round_up(K, 64) * 2;
}

__global__ void pack_weight_kernel(const float* w, int64_t stride_n, int64_t stride_k, int N, int K, int NT, int layout,
                                   int Npad, int Kpad, __nv_bfloat16* out) {
  // one thread per 16-byte chunk (8 bf16 along K) of the padded (Npad x Kpad) matrix
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int chunks_per_row = Kpad / 8;
  if (i >= (int64_t)Npad * chunks_per_row) return;
  int n = (int)(i / chunks_per_row), ck = (int)(i % chunks_per_row);
  uint32_t v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int k0 = ck * 8 + 2 * e;
    float a = (n < N && k0 < K) ? w[(int64_t)n * stride_n + (int64_t)k0 * stride_k] : 0.0f;
    float b = (n < N && k0 + 1 < K) ? w[(int64_t)n * stride_n + (int64_t)(k0 + 1) * stride_k] : 0.0f;
    v[e] = tc::pack_bf16x2(a, b);
  }
  int tile = n / NT, r = n % NT;
  size_t off;
  if (layout == 0) {
    int kb = ck / 8, c16 = ck % 8;
    off = ((size_t)tile * (Kpad / 64) + kb) * tc::kblock_bytes(NT) + tc::sw128_offset(r, c16);
  } else {
    off = (size_t)tile * NT * Kpad * 2 + (size_t)ck * NT * 16 + (size_t)r * 16;
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + off) = make_uint4(v[0], v[1], v[2], v[3]);
}

int tc_pack_weight(const float* w, int64_t stride_n, int64_t stride_k, int N, int K, int NT, int layout,
                   __nv_bfloat16* out, cudaStream_t st) {
  if (NT % 8 || NT <= 0) return fail(SMX_ERR_BAD_ARG, "pack: NT must be a positive multiple of 8");
  int Npad = round_up(N, NT), Kpad = round_up(K, 64);
  int64_t n = (int64_t)Npad * (Kpad / 8);
  pack_weight_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, stride_n, stride_k, N, K, NT, layout, Npad, Kpad, out);
  count_launch();
  return check_launch("pack_weight_kernel");
}

// ---------------------------------------------------------------------------------------------
// UMMA self-test: C (M x N fp32) = A (M x K bf16, row-major) * W^T, W packed by tc_pack_weight.
// One CTA per 128-row tile.  layout 0 = 128B swizzle, 1 = no swizzle.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tc_gemm_test_kernel(const __nv_bfloat16* __restrict__ A,
                                                           const __nv_bfloat16* __restrict__ Wp, float* __restrict__ C,
                                                           int M, int N, int K, int NT, int layout, uint32_t tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int Kpad = (K + 63) / 64 * 64;
  const int nkb = Kpad / 64;
  uint8_t* sA = smem;                                  // 128 x Kpad bf16
  uint8_t* sB = smem + (size_t)128 * Kpad * 2;         // NT x Kpad bf16
  __shared__ __align__(8) uint64_t full_bar, mma_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int m0 = blockIdx.x * 128;

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, tmem_cols);
  if (tid == 0) {
    tc::mbar_init(&full_bar, 1);
    tc::mbar_init(&mma_bar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  const uint32_t bytesB = (uint32_t)NT * Kpad * 2;
  if (tid == 0) {
    tc::mbar_arrive_expect_tx(&full_bar, bytesB);
    tc::bulk_g2s(sB, Wp, bytesB, &full_bar);
  }
  // A tile: global (row-major) -> operand layout in shared memory, 16 bytes per step
  const int chunks = 128 * (Kpad / 8);
  for (int i = tid; i < chunks; i += 128) {
    int r = i / (Kpad / 8), ck = i % (Kpad / 8);
    uint4 v = make_uint4(0, 0, 0, 0);
    int m = m0 + r;
    if (m < M && ck * 8 < K) v = *reinterpret_cast<const uint4*>(A + (size_t)m * K + ck * 8);  // K % 8 == 0 required
    size_t off;
    if (layout == 0) off = (size_t)(ck / 8) * tc::kblock_bytes(128) + tc::sw128_offset(r, ck % 8);
    else off = (size_t)ck * 128 * 16 + (size_t)r * 16;
    *reinterpret_cast<uint4*>(sA + off) = v;
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    tc::mbar_wait(&full_bar, 0);
    tc::tc_fence_after();
    const uint32_t idesc = tc::make_idesc_bf16(128, (uint32_t)NT);
    const uint32_t a0 = tc::smem_u32(sA), b0 = tc::smem_u32(sB);
    for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {  // 4 x K=16 per 64-column block
        uint64_t ad, bd;
        if (layout == 0) {
          ad = tc::make_desc_sw128(a0 + kb * tc::kblock_bytes(128) + ks * 32);
          bd = tc::make_desc_sw128(b0 + kb * tc::kblock_bytes(NT) + ks * 32);
        } else {
          int chunk = kb * 8 + ks * 2;
          ad = tc::make_desc_nosw(a0 + chunk * 128 * 16, 128 * 16, 128);
          bd = tc::make_desc_nosw(b0 + chunk * NT * 16, NT * 16, 128);
        }
        tc::umma_bf16(tmem, ad, bd, idesc, (kb | ks) ? 1u : 0u);
      }
    }
    tc::umma_commit(&mma_bar);
  }
  __syncwarp();
  tc::mbar_wait(&mma_bar, 0);
  tc::tc_fence_after();
  const int row = m0 + warp * 32 + lane;
  for (int c0 = 0; c0 < NT; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tc::tmem_ld_wait();
    if (row < M) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < N) C[(size_t)row * N + c0 + j] = v[j];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, tmem_cols);
}

}  // namespace smx

using namespace smx;

extern "C" SMX_API int smx_debug_tc_gemm(int layout, int32_t M, int32_t N, int32_t K, const void* a_bf16,
                                         const float* w_f32, float* c_f32, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  if (!a_bf16 || !w_f32 || !c_f32 || !workspace) return fail(SMX_ERR_BAD_ARG, "debug_tc_gemm: NULL pointer");
  if (M <= 0 || N <= 0 || N > 256 || K <= 0 || K % 8) return fail(SMX_ERR_BAD_ARG, "debug_tc_gemm: need 0<N<=256, K%%8==0");
  if (layout != 0 && layout != 1) return fail(SMX_ERR_BAD_ARG, "debug_tc_gemm: layout 0 or 1");
  const int NT = (N + 15) / 16 * 16;
  const int Kpad = (K + 63) / 64 * 64;
  size_t need = tc_packed_bytes(N, K, NT);
  if (workspace_bytes < need) return fail(SMX_ERR_WORKSPACE, "debug_tc_gemm: workspace needs %zu bytes", need);
  cudaStream_t st = (cudaStream_t)stream;
  SMX_TRY(tc_pack_weight(w_f32, K, 1, N, K, NT, layout, (__nv_bfloat16*)workspace, st));
  size_t smem = (size_t)(128 + NT) * Kpad * 2 + 1024;
  if (smem > 227 * 1024) return fail(SMX_ERR_UNSUPPORTED, "debug_tc_gemm: tile does not fit shared memory");
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  uint32_t cols = 32;
  while (cols < (uint32_t)NT) cols <<= 1;
  tc_gemm_test_kernel<<<(M + 127) / 128, 128, smem, st>>>((const __nv_bfloat16*)a_bf16, (const __nv_bfloat16*)workspace,
                                                         c_f32, M, N, K, NT, layout, cols);
  count_launch();
  return check_launch("tc_gemm_test_kernel");
}

// debug: device buffer (>= 1024 x uint64, zeroed by the caller) receiving clock64() timelines of CTA 0 of the fused
// kernels; NULL switches tracing off.  Not thread-safe; diagnostics only.
extern "C" SMX_API int smx_debug_set_trace(void* device_u64_buffer) {
  tc_set_trace(device_u64_buffer);
  tc_set_trace_cell3(device_u64_buffer);
  tc_set_trace_cell4(device_u64_buffer);
  tc_set_trace_ffn(device_u64_buffer ? (char*)device_u64_buffer + 4096 : nullptr);  // entries 512..1023
  tc_set_trace_ffn3(device_u64_buffer ? (char*)device_u64_buffer + 4096 : nullptr);
  tc_set_trace_conv(device_u64_buffer ? (char*)device_u64_buffer + 7680 : nullptr); // entries 960..1023
  return SMX_OK;
}

// tuning knob: thread-block cluster size (1, 2 or 4) of the persistent FFN kernel (weight multicast)
extern "C" SMX_API int smx_debug_set_ffn_cluster(int cluster_size) {
  tc_set_ffn_cluster(cluster_size);
  return SMX_OK;
}

namespace smx {
static std::atomic<int> g_pdl{1};
bool tc_pdl_enabled() { return g_pdl != 0; }
void tc_set_pdl(int on) { g_pdl = on ? 1 : 0; }
}  // namespace smx
// programmatic dependent launch of the fused kernels on (default) / off
extern "C" SMX_API int smx_debug_set_pdl(int on) {
  tc_set_pdl(on);
  return SMX_OK;
}

// fp32 I/O: large linears on the tensor cores with split-bf16 operands (default on) or on the CUDA-core GEMM (exact fp32 products)
extern "C" SMX_API int smx_debug_set_f32_tc(int on) {
  tc_set_f32_tc(on);
  return SMX_OK;
}

// A-B switch between the fused cell / GLU-pass generations (3: operands in tensor memory, 1: first generation)
extern "C" SMX_API int smx_debug_set_cell_version(int version) {
  tc_set_cell_version(version);
  return SMX_OK;
}

// A-B switch between the fused FFN generations (3: hidden activation in tensor memory, 2: in shared memory)
extern "C" SMX_API int smx_debug_set_ffn_version(int version) {
  tc_set_ffn_version(version);
  return SMX_OK;
}
