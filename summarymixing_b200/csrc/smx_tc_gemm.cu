// tcgen05 arm of libsmx, part 11: K-GEMM, the general linear kernel, and the fused CSGU gate of the Branchformer block.
//
//   out[M, N] (bf16) = epilogue( A[M, K] (bf16 rows, any row pitch) @ W[N, K]^T (bf16, nn.Linear layout) )
//   epilogue: + bias[n] + rowbias[row / rows_per_group][n] -> activation -> * rowmask[row] -> resid + alpha * v
//
// Unlike K-LIN (smx_tc_lin.cu: the whole 128 x K activation tile staged in shared memory, K <= 512, LayerNorm prologue),
// both operands STREAM through one shared-memory ring by tensor-map TMA (cp.async.bulk.tensor.2d, 128-byte swizzle = the
// UMMA operand image), so K is unbounded: the Branchformer's 1536 -> D and 2D -> D projections (Branchformer.py:78-84,
// 220-226), the frontend's 640 -> D input projection (TransformerASR.py:353-358).  Persistent over 128 x NT output tiles
// (tile order n-major so that concurrently running CTAs share a weight panel in L2), two TMEM accumulators: the eight
// epilogue warps drain tile i while the tensor pipe works on tile i + 1.
// Warp roles: 0 TMA producer | 1 MMA issuer | 2-9 epilogue (quadrant = warp % 4, column half = (warp - 2) / 4)
#include <cuda.h>
#include <stdlib.h>

#include "smx_tc.h"
#include "smx_tc_common.cuh"

namespace smx {

using tc::kblock_bytes;

constexpr int GM_THREADS = 320;
constexpr int GM_MAX_STAGES = 8;

struct GemmTcP {
  int M, N, K, NT, m_tiles, n_tiles;
  const float* bias;
  const float* rowbias; int64_t rowbias_ld; int rows_per_group;
  int act;
  const uint8_t* rowmask;
  const __nv_bfloat16* resid; int64_t ldr; float alpha;
  __nv_bfloat16* out; int64_t ldo;
  float* out_f32; const float* resid_f32;   // fp32 output / residual (split3 arm); out / resid are NULL then
  int split3;                                // A and W hold [hi | lo] halves of width K: accumulate hi*hi + hi*lo + lo*hi
  int ksplit, kb_per_split;                  // split-K: slice ks covers K-blocks [ks kb_per_split, ...) and writes out_f32 + ks * out_split_stride
  int64_t out_split_stride;
  int glu;                                   // GLU epilogue: W / bias rows are interleaved per N tile ([NT/2 value rows | their NT/2 gate rows], tc_glu_dense_bf16);
                                             // out (width N/2) = (v + bv) * sigmoid(g + bg)                                   Conformer.py:322-324
  int bd_in, bd_out;                         // block-diagonal W (ParallelLinear as dense with zero blocks): head widths; 0 = dense.  An N tile
                                             // only visits the K-blocks of the heads it covers (the others are zero)
  int n_stages; uint32_t stage_bytes; uint32_t tmem_cols;
};
// K-block range [kb0, kb0 + nloc) of one (split-K slice, N tile)
__device__ __forceinline__ void gm_krange(const GemmTcP& p, int ks, int nt, int nkb1, int& kb0, int& nloc) {
  kb0 = ks * p.kb_per_split;
  nloc = min(p.kb_per_split, nkb1 - kb0);
  if (p.bd_in) {
    const int h0 = (nt * p.NT) / p.bd_out, h1 = (nt * p.NT + p.NT - 1) / p.bd_out;
    kb0 = (h0 * p.bd_in) >> 6;
    nloc = min(nkb1, ((h1 + 1) * p.bd_in + 63) >> 6) - kb0;
  }
}

__device__ __forceinline__ void gm_tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void gm_ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void gm_stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

template <bool GLU>   // (the GLU epilogue is its own instantiation: the plain epilogue's code and register allocation stay as they were)
__global__ void __launch_bounds__(GM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmW, const GemmTcP p) {
  extern __shared__ __align__(1024) uint8_t smem[];  // ring stages: [A block 128 x 64 | W block NT x 64]
  __shared__ __align__(8) uint64_t full_bar[GM_MAX_STAGES], empty_bar[GM_MAX_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nkb1 = p.K >> 6, NT = p.NT;
  const int nseg = p.split3 ? 3 : 1;           // split3: iteration = (segment, K-block): segments (hi,hi) (hi,lo) (lo,hi)
  const int mn_tiles = p.m_tiles * p.n_tiles;
  const int n_tiles_total = mn_tiles * p.ksplit;

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, p.tmem_cols);
  if (tid == 32) {
    for (int s = 0; s < p.n_stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 8); }
    tc::fence_barrier_init();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  tc::pdl_wait();               // A (and the residual) come from the preceding kernel
  tc::pdl_launch_dependents();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
        const int ks = tile / mn_tiles, t2 = tile - ks * mn_tiles;
        const int mt = t2 / p.n_tiles, nt = t2 - mt * p.n_tiles;   // N tiles of one row tile run side by side (on neighbouring CTAs): its A tile leaves HBM once
        int kb0, nloc;
        gm_krange(p, ks, nt, nkb1, kb0, nloc);
#pragma unroll 1
        for (int kb = 0; kb < nseg * nloc; ++kb) {
          tc::mbar_wait(&empty_bar[s], ph ^ 1u);
          tc::mbar_arrive_expect_tx(&full_bar[s], p.stage_bytes);
          uint8_t* dst = smem + (size_t)s * p.stage_bytes;
          const int seg = kb / nloc, k1 = kb0 + kb - seg * nloc;
          gm_tma_load_2d(dst, &tmA, k1 * 64 + (seg == 2 ? p.K : 0), mt * 128, &full_bar[s]);          // rows >= M are zero-filled
          gm_tma_load_2d(dst + kblock_bytes(128), &tmW, k1 * 64 + (seg == 1 ? p.K : 0), nt * NT, &full_bar[s]);
          if (++s == p.n_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    int s = 0;
    uint32_t ph = 0;
    const uint32_t s0 = tc::smem_u32(smem);
    const uint32_t idesc = tc::make_idesc_bf16(128, (uint32_t)NT);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      tc::mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator (first two uses: fresh)
      tc::tc_fence_after();
      const uint32_t d_addr = tmem + (uint32_t)buf * (uint32_t)NT;
      const int ks = tile / mn_tiles, nt_i = (tile - ks * mn_tiles) % p.n_tiles;
      int kb0_i, nloc_i;
      gm_krange(p, ks, nt_i, nkb1, kb0_i, nloc_i);
      const int nkb = nseg * nloc_i;
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb) {
        tc::mbar_wait_spin(&full_bar[s], ph);
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t a_addr = s0 + (uint32_t)s * p.stage_bytes;
          const uint64_t ad = tc::make_desc_sw128(a_addr), bd = tc::make_desc_sw128(a_addr + kblock_bytes(128));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc::umma_bf16(d_addr, ad + 2u * ks, bd + 2u * ks, idesc, (kb == 0 && ks == 0) ? 0u : 1u);
          tc::umma_commit(&empty_bar[s]);
          if (kb == nkb - 1) tc::umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++s == p.n_stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const int npc = NT >> 6;  // 32-column pieces per column half (NT = 64: one piece per half)
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int ks = tile / mn_tiles, t2 = tile - ks * mn_tiles;
      const int mt = t2 / p.n_tiles, nt = t2 - mt * p.n_tiles;
      const int64_t grow = (int64_t)mt * 128 + r;
      const bool live = grow < p.M;
      const float rscale = (live && p.rowmask) ? (float)p.rowmask[grow] : 1.0f;
      const float* rb = (live && p.rowbias) ? p.rowbias + (grow / p.rows_per_group) * p.rowbias_ld : nullptr;
      tc::mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc::tc_fence_after();
      if constexpr (GLU) {
#pragma unroll 1
        for (int pc = 0; pc < (NT >> 7); ++pc) {
          const int c = half * (NT >> 2) + pc * 32;       // channel inside the tile's NT/2 channels: value column c, gate column NT/2 + c
          float v[32], gt[32];
          tc::tmem_ld32(tmem + lane_sel + (uint32_t)buf * (uint32_t)NT + c, v);
          tc::tmem_ld32(tmem + lane_sel + (uint32_t)buf * (uint32_t)NT + (NT >> 1) + c, gt);
          tc::tmem_ld_wait();
          const float4* bv = reinterpret_cast<const float4*>(p.bias + nt * NT + c);
          const float4* bg = reinterpret_cast<const float4*>(p.bias + nt * NT + (NT >> 1) + c);
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = __ldg(bv + j), g4 = __ldg(bg + j);
            const float av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w};
            float r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float hv = 0.5f * (v[4 * j + e] + av[e]);
              r[e] = fmaf(hv, tc::tanh_approx(0.5f * (gt[4 * j + e] + gv[e])), hv);   // a * sigmoid(z) = a/2 * tanh(z/2) + a/2
            }
            o[2 * j] = tc::pack_bf16x2(r[0], r[1]);
            o[2 * j + 1] = tc::pack_bf16x2(r[2], r[3]);
          }
          if (live) {
            __nv_bfloat16* op = p.out + grow * p.ldo + nt * (NT >> 1) + c;
            gm_stg256(op, o);
            gm_stg256(op + 16, o + 8);
          }
        }
      } else
#pragma unroll 1
      for (int pc = 0; pc < npc; ++pc) {
        const int col = half * (NT >> 1) + pc * 32;       // column inside the tile
        const int gcol = nt * NT + col;
        uint32_t rres[16];
        if (p.resid && live) { gm_ldg256(p.resid + grow * p.ldr + gcol, rres); gm_ldg256(p.resid + grow * p.ldr + gcol + 16, rres + 8); }
        float v[32];
        tc::tmem_ld32(tmem + lane_sel + (uint32_t)buf * (uint32_t)NT + col, v);
        tc::tmem_ld_wait();
        if (p.bias) {
          const float4* bp = reinterpret_cast<const float4*>(p.bias + gcol);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float4 bb = __ldg(bp + j); v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
        }
        if (rb) {
          const float4* bp = reinterpret_cast<const float4*>(rb + gcol);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float4 bb = __ldcg(bp + j); v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
        }
        tc::act_apply<32>(p.act, v);
        if (p.rowmask) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= rscale;
        }
        if (p.resid && live) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rres[i]));
            v[2 * i] = fmaf(p.alpha, v[2 * i], f.x); v[2 * i + 1] = fmaf(p.alpha, v[2 * i + 1], f.y);
          }
        }
        if (p.resid_f32 && live) {
          const float4* rp = reinterpret_cast<const float4*>(p.resid_f32 + grow * p.ldr + gcol);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 f = __ldcg(rp + j);
            v[4 * j] = fmaf(p.alpha, v[4 * j], f.x); v[4 * j + 1] = fmaf(p.alpha, v[4 * j + 1], f.y);
            v[4 * j + 2] = fmaf(p.alpha, v[4 * j + 2], f.z); v[4 * j + 3] = fmaf(p.alpha, v[4 * j + 3], f.w);
          }
        }
        if (live && p.out_f32) {
          float* op = p.out_f32 + (int64_t)ks * p.out_split_stride + grow * p.ldo + gcol;
#pragma unroll
          for (int j = 0; j < 4; ++j) gm_stg256(op + 8 * j, reinterpret_cast<const uint32_t*>(v + 8 * j));
        } else if (live) {
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = tc::pack_bf16x2(v[2 * i], v[2 * i + 1]);
          gm_stg256(p.out + grow * p.ldo + gcol, o);
          gm_stg256(p.out + grow * p.ldo + gcol + 16, o + 8);
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*gm_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static gm_encode_fn gm_encoder() {
  static gm_encode_fn fn = nullptr;
  static std::atomic<int> state{0};
  if (state.load() == 0) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess && f) { fn = (gm_encode_fn)f; state = 1; }
    else { (void)cudaGetLastError(); state = 2; }
  }
  return state.load() == 1 ? fn : nullptr;
}
static int gm_map_2d(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_bytes, uint32_t box_rows) {
  gm_encode_fn enc = gm_encoder();
  if (!enc) return fail(SMX_ERR_UNSUPPORTED, "tc gemm: cuTensorMapEncodeTiled is not available");
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstr[1] = {pitch_bytes};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SMX_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return SMX_OK;
}

bool tc_gemm_supported(int K, int N) { return K >= 64 && K % 64 == 0 && N >= 64 && N % 64 == 0; }

static int gm_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int tc_gemm_launch(const GemmTc& g, cudaStream_t st) {
  if (!tc_gemm_supported(g.K, g.N) || g.M <= 0) return fail(SMX_ERR_UNSUPPORTED, "tc gemm: M=%lld N=%d K=%d", (long long)g.M, g.N, g.K);
  if (g.M > 0x7fffffff) return fail(SMX_ERR_UNSUPPORTED, "tc gemm: too many rows");
  if (((uintptr_t)g.a % 16) || ((uintptr_t)g.w % 16) || ((uintptr_t)g.out % 32) || ((uintptr_t)g.out_f32 % 32) || (g.lda % 8) ||
      (g.ldo % (g.out_f32 ? 8 : 16)) || (g.resid && (((uintptr_t)g.resid % 32) || g.ldr % 16)) ||
      (g.resid_f32 && (((uintptr_t)g.resid_f32 % 16) || g.ldr % 4)))
    return fail(SMX_ERR_ALIGNMENT, "tc gemm: operand alignment");
  if ((g.out != nullptr) == (g.out_f32 != nullptr)) return fail(SMX_ERR_BAD_ARG, "tc gemm: exactly one of out / out_f32");
  GemmTcP p{};
  p.M = (int)g.M; p.N = g.N; p.K = g.K;
  p.NT = g.N % 256 == 0 ? 256 : (g.N % 128 == 0 ? 128 : 64);
  p.m_tiles = (int)((g.M + 127) / 128); p.n_tiles = g.N / p.NT;
  p.bias = g.bias; p.rowbias = g.rowbias; p.rowbias_ld = g.rowbias_ld; p.rows_per_group = g.rows_per_group > 0 ? g.rows_per_group : 1;
  p.act = g.act; p.rowmask = g.rowmask; p.resid = g.resid; p.ldr = g.ldr; p.alpha = g.alpha;
  p.out = g.out; p.ldo = g.ldo; p.out_f32 = g.out_f32; p.resid_f32 = g.resid_f32; p.split3 = g.split3;
  if (g.glu) {
    if (p.NT < 128 || !g.out || !g.bias || g.split3 || g.ksplit > 1 || g.resid || g.rowbias || g.rowmask || g.ldo % 16)
      return fail(SMX_ERR_UNSUPPORTED, "tc gemm: GLU epilogue needs N %% 128 == 0, bf16 output, an (interleaved) bias and no other epilogue term");
    p.glu = 1;
  }
  {
    const int nkb1 = g.K / 64;
    int want = g.ksplit > 1 ? g.ksplit : 1;
    if (want > 1 && !g.out_f32) return fail(SMX_ERR_BAD_ARG, "tc gemm: split-K needs fp32 partial outputs");
    if (want > nkb1) want = nkb1;
    p.kb_per_split = (nkb1 + want - 1) / want;
    p.ksplit = (nkb1 + p.kb_per_split - 1) / p.kb_per_split;   // no empty slice
    p.out_split_stride = g.out_split_stride;
  }
  if (g.bd_in > 0 && g.bd_out > 0 && p.ksplit == 1 && g.bd_in % 64 == 0 && g.K % g.bd_in == 0 && g.N % g.bd_out == 0 && g.K / g.bd_in == g.N / g.bd_out) {
    p.bd_in = g.bd_in; p.bd_out = g.bd_out;
  }
  p.stage_bytes = kblock_bytes(128) + (uint32_t)p.NT * 128u;
  int stages = (int)((227 * 1024 - 2048) / p.stage_bytes);
  if (stages > GM_MAX_STAGES) stages = GM_MAX_STAGES;
  p.n_stages = stages;
  p.tmem_cols = p.NT == 256 ? 512 : (p.NT == 128 ? 256 : 128);
  CUtensorMap tmA, tmW;
  const uint64_t kw = g.split3 ? 2 * (uint64_t)g.K : (uint64_t)g.K;  // split3: rows of A and W are [hi | lo], 2K wide
  SMX_TRY(gm_map_2d(&tmA, g.a, kw, (uint64_t)g.M, (uint64_t)g.lda * 2, 128));
  SMX_TRY(gm_map_2d(&tmW, g.w, kw, (uint64_t)g.N, kw * 2, (uint32_t)p.NT));
  const size_t smem = (size_t)stages * p.stage_bytes + 1024;
  cudaError_t e = p.glu ? cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                        : cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(gemm_tc_kernel): %s", cudaGetErrorString(e));
  const int total = p.m_tiles * p.n_tiles * p.ksplit;
  const unsigned grid = (unsigned)(total < gm_sms() ? total : gm_sms());
  e = p.glu ? launch_pdl(gemm_tc_kernel<true>, dim3(grid), dim3(GM_THREADS), smem, st, 1u, tmA, tmW, p)
            : launch_pdl(gemm_tc_kernel<false>, dim3(grid), dim3(GM_THREADS), smem, st, 1u, tmA, tmW, p);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaLaunchKernelEx(gemm_tc_kernel): %s", cudaGetErrorString(e));
  count_tc_launch();
  return check_launch("gemm_tc_kernel");
}

// dense bf16 (N, K) row-major copy of columns [k_offset, k_offset + K) of an smx_linear (block-diagonal weights get their zeros)
__global__ void dense_bf16_kernel(const float* __restrict__ w, int in_dim, int out_dim, int n_split, int k_offset, int K, int N,
                                  __nv_bfloat16* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * K) return;
  const int n = (int)(i / K), k = (int)(i % K) + k_offset;
  float val;
  if (n_split <= 1) {
    val = w[(int64_t)n * in_dim + k];
  } else {
    const int hi = in_dim / n_split, ho = out_dim / n_split, m = n / ho;
    val = (k / hi == m) ? w[((int64_t)m * hi + (k - m * hi)) * ho + (n - m * ho)] : 0.0f;
  }
  out[i] = __float2bfloat16(val);
}
int tc_dense_bf16(const smx_linear& L, int k_offset, int K, void* out, cudaStream_t st) {
  const int64_t n = (int64_t)L.out_dim * K;
  dense_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L.w, L.in_dim, L.out_dim, L.n_split, k_offset, K, L.out_dim,
                                                                 (__nv_bfloat16*)out);
  count_launch();
  return check_launch("dense_bf16_kernel");
}

// ---- fp32 on the tensor cores: split-bf16 ("bf16x3") operands ------------------------------------------------------------------
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): x w ~= hi_x hi_w + hi_x lo_w + lo_x hi_w (the dropped lo lo term is 2^-16
// relative).  Rows are stored [hi | lo] (2K bf16); K-GEMM's split3 mode walks the three (A, W) half pairings as one 3K-long
// accumulation in fp32, so an fp32 linear costs three bf16 MMAs and stays ~1e-5 relative to the fp32 result.
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int K, __nv_bfloat16* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 consecutive elements
  const int k4 = K >> 2;
  if (i >= rows * k4) return;
  const int64_t r = i / k4;
  const int c = (int)(i - r * k4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
  const float f[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) { hi[e] = __float2bfloat16(f[e]); lo[e] = __float2bfloat16(f[e] - __bfloat162float(hi[e])); }
  __nv_bfloat16* o = out + r * (2 * (int64_t)K) + c;
  *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(o + K) = *reinterpret_cast<const uint2*>(lo);
}
// W' (N, 2K) = [hi | lo] of columns [k_offset, k_offset + K) of the dense (N, in_dim) view of an smx_linear
// GLU weights for the K-GEMM GLU epilogue: row nt * NT + j of the image is value row nt * NT/2 + j (j < NT/2) or gate row
// N/2 + nt * NT/2 + (j - NT/2) of the dense (N, K) linear; the bias likewise (fp32).  NT as tc_gemm_launch picks it for N.
__global__ void glu_dense_bf16_kernel(const float* __restrict__ w, const float* __restrict__ b, int K, int N, int NT, __nv_bfloat16* __restrict__ out,
                                      float* __restrict__ bias_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * K) return;
  const int n = (int)(i / K), k = (int)(i - (int64_t)n * K);
  const int nt = n / NT, j = n - nt * NT, h = NT >> 1;
  const int src = j < h ? nt * h + j : (N >> 1) + nt * h + (j - h);
  out[i] = __float2bfloat16_rn(w[(int64_t)src * K + k]);
  if (k == 0) bias_out[n] = b[src];
}
size_t tc_glu_dense_bytes(int K, int N) { return align_up((size_t)N * K * 2, 1024) + align_up((size_t)N * 4, 1024); }
int tc_glu_dense_bf16(const smx_linear& L, void* out, cudaStream_t st) {
  const int N = L.out_dim, K = L.in_dim;
  if (L.n_split > 1 || !L.b || N % 256) return fail(SMX_ERR_UNSUPPORTED, "GLU image: dense linear with a bias and out_dim %% 256 == 0");
  const int NT = 256;
  const int64_t n = (int64_t)N * K;
  glu_dense_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L.w, L.b, K, N, NT, (__nv_bfloat16*)out,
                                                                     (float*)((char*)out + align_up((size_t)N * K * 2, 1024)));
  count_launch();
  return check_launch("glu_dense_bf16_kernel");
}

__global__ void __launch_bounds__(256) split_weight_kernel(const float* __restrict__ w, int in_dim, int out_dim, int n_split, int k_offset, int K, int N,
                                                           __nv_bfloat16* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * K) return;
  const int n = (int)(i / K), k = (int)(i % K) + k_offset;
  float val;
  if (n_split <= 1) {
    val = w[(int64_t)n * in_dim + k];
  } else {
    const int hi_ = in_dim / n_split, ho = out_dim / n_split, m = n / ho;
    val = (k / hi_ == m) ? w[((int64_t)m * hi_ + (k - m * hi_)) * ho + (n - m * ho)] : 0.0f;
  }
  const __nv_bfloat16 h = __float2bfloat16(val);
  out[(int64_t)n * 2 * K + (k - k_offset)] = h;
  out[(int64_t)n * 2 * K + K + (k - k_offset)] = __float2bfloat16(val - __bfloat162float(h));
}
size_t tc_split3_scratch_bytes(int64_t rows, int K, int N) { return align_up((size_t)rows * 2 * K * 2, 1024) + align_up((size_t)N * 2 * K * 2, 1024); }
bool tc_split3_ok(int64_t rows, int K, int N) { return rows >= 128 && tc_gemm_supported(K, N); }  // (one full row tile: the arithmetic of a row must not depend on how many rows ride along -- utterance sharding)
static std::atomic<int> g_f32_tc{1};
void tc_set_f32_tc(int on) { g_f32_tc = on ? 1 : 0; }
bool tc_f32_tc_enabled() { return g_f32_tc.load() != 0; }
// out = epilogue(A (rows, K) fp32 @ W^T) through the split-bf16 GEMM; g carries the epilogue (a / w / K / split3 are set here)
int tc_linear_split3(const smx_linear& L, int k_offset, int K, const float* A, int64_t lda, int64_t rows, GemmTc g, void* scratch, cudaStream_t st) {
  if (lda % 4 || ((uintptr_t)A % 16)) return fail(SMX_ERR_ALIGNMENT, "split3: A alignment");
  __nv_bfloat16* a2 = (__nv_bfloat16*)scratch;
  __nv_bfloat16* w2 = (__nv_bfloat16*)((char*)scratch + align_up((size_t)rows * 2 * K * 2, 1024));
  const int64_t na = rows * (K / 4), nw = (int64_t)L.out_dim * K;
  split_rows_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(A, lda, rows, K, a2);
  count_launch();
  SMX_TRY(check_launch("split_rows_kernel"));
  split_weight_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(L.w, L.in_dim, L.out_dim, L.n_split, k_offset, K, L.out_dim, w2);
  count_launch();
  SMX_TRY(check_launch("split_weight_kernel"));
  g.a = a2; g.lda = 2 * (int64_t)K; g.M = rows; g.N = L.out_dim; g.K = K; g.w = w2; g.split3 = 1;
  return tc_gemm_launch(g, st);
}

// ---- backward linears on the tensor cores (split-bf16, fp32 accuracy) -----------------------------------------------------------
// out (C, 2 Kp) = [hi | lo] of x^T for x (rows, C) fp32; Kp = rows rounded up to 64 (zero padded): the operands of a weight gradient
// dW = dZ^T X are the TRANSPOSED activations, contracted over the rows.
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int C, int64_t Kp,
                                                              __nv_bfloat16* __restrict__ out) {
  // a block transposes 64 rows x 32 columns; every thread then stores ROW PAIRS as bf16x2 (a warp writes 128 contiguous bytes of one
  // output row; 2-byte stores, 64 bytes per warp, ran at 2.6 TB/s of combined traffic)
  __shared__ float tile[32][66];   // [column][row]
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t r = r0 + ty + 8 * j;
    const int c = c0 + tx;
    tile[tx][ty + 8 * j] = (r < rows && c < C) ? x[r * ldx + c] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cl = ty + 8 * j, c = c0 + cl;
    const int64_t r = r0 + 2 * tx;   // (Kp is a multiple of 64: the pair is inside the padded row)
    if (c < C) {
      const float2 v = *reinterpret_cast<const float2*>(&tile[cl][2 * tx]);
      const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
      const float2 hf = __bfloat1622float2(h);
      *reinterpret_cast<__nv_bfloat162*>(out + (int64_t)c * 2 * Kp + r) = h;
      *reinterpret_cast<__nv_bfloat162*>(out + (int64_t)c * 2 * Kp + Kp + r) = __floats2bfloat162_rn(v.x - hf.x, v.y - hf.y);
    }
  }
}
// W'' (Kin, 2 out) = [hi | lo] of the TRANSPOSED dense view: W''[i][o] = W[o][k_offset + i] (data gradient dX = dZ W)
__global__ void __launch_bounds__(256) split_weight_t_kernel(const float* __restrict__ w, int in_dim, int out_dim, int n_split, int k_offset, int Kin,
                                                             __nv_bfloat16* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Kin * out_dim) return;
  const int i = (int)(idx / out_dim), o = (int)(idx % out_dim), k = i + k_offset;
  float val;
  if (n_split <= 1) {
    val = w[(int64_t)o * in_dim + k];
  } else {
    const int hi_ = in_dim / n_split, ho = out_dim / n_split, m = o / ho;
    val = (k / hi_ == m) ? w[((int64_t)m * hi_ + (k - m * hi_)) * ho + (o - m * ho)] : 0.0f;
  }
  const __nv_bfloat16 h = __float2bfloat16(val);
  out[(int64_t)i * 2 * out_dim + o] = h;
  out[(int64_t)i * 2 * out_dim + out_dim + o] = __float2bfloat16(val - __bfloat162float(h));
}
// dX (rows, Kin) = dZ (rows, out) @ W[:, k_offset : k_offset + Kin]; g carries the output / residual
int tc_dgrad_split3(const smx_linear& L, int k_offset, int Kin, const float* dZ, int64_t ldz, int64_t rows, GemmTc g, void* scratch, cudaStream_t st) {
  if (ldz % 4 || ((uintptr_t)dZ % 16)) return fail(SMX_ERR_ALIGNMENT, "split3 dgrad: dZ alignment");
  const int out = L.out_dim;
  __nv_bfloat16* a2 = (__nv_bfloat16*)scratch;
  __nv_bfloat16* w2 = (__nv_bfloat16*)((char*)scratch + align_up((size_t)rows * 2 * out * 2, 1024));
  const int64_t na = rows * (out / 4), nw = (int64_t)Kin * out;
  split_rows_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(dZ, ldz, rows, out, a2);
  count_launch();
  SMX_TRY(check_launch("split_rows_kernel"));
  split_weight_t_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(L.w, L.in_dim, L.out_dim, L.n_split, k_offset, Kin, w2);
  count_launch();
  SMX_TRY(check_launch("split_weight_t_kernel"));
  g.a = a2; g.lda = 2 * (int64_t)out; g.M = rows; g.N = Kin; g.K = out; g.w = w2; g.split3 = 1;
  return tc_gemm_launch(g, st);
}
// P (ksplit, M, N) fp32 partials of A^T B over the rows: A (rows, M), B (rows, N); returns the number of slices in *ns
size_t tc_wgrad_scratch_bytes(int64_t rows, int M, int N) {
  const int64_t Kp = (rows + 63) / 64 * 64;
  return align_up((size_t)M * 2 * Kp * 2, 1024) + align_up((size_t)N * 2 * Kp * 2, 1024);
}
int tc_wgrad_slices(int64_t rows) { const int64_t nkb = (rows + 63) / 64; return (int)(nkb < 64 ? nkb : 64); }
int tc_wgrad_split3(const float* A, int64_t lda, int M, const float* Bm, int64_t ldb, int N, int64_t rows, float* P, int* ns, void* scratch,
                    cudaStream_t st) {
  const int64_t Kp = (rows + 63) / 64 * 64;
  if (Kp > 0x7fffffff) return fail(SMX_ERR_UNSUPPORTED, "split3 wgrad: too many rows");
  __nv_bfloat16* a2 = (__nv_bfloat16*)scratch;
  __nv_bfloat16* b2 = (__nv_bfloat16*)((char*)scratch + align_up((size_t)M * 2 * Kp * 2, 1024));
  dim3 ga((unsigned)(Kp / 64), (M + 31) / 32), gb((unsigned)(Kp / 64), (N + 31) / 32);
  split_transpose_kernel<<<ga, 256, 0, st>>>(A, lda, rows, M, Kp, a2);
  count_launch();
  SMX_TRY(check_launch("split_transpose_kernel"));
  split_transpose_kernel<<<gb, 256, 0, st>>>(Bm, ldb, rows, N, Kp, b2);
  count_launch();
  SMX_TRY(check_launch("split_transpose_kernel"));
  GemmTc g{};
  g.a = a2; g.lda = 2 * Kp; g.M = M; g.N = N; g.K = (int)Kp; g.w = b2; g.split3 = 1;
  g.act = SMX_ACT_IDENTITY; g.alpha = 1.0f;
  g.out_f32 = P; g.ldo = N; g.ksplit = tc_wgrad_slices(rows); g.out_split_stride = (int64_t)M * N;
  SMX_TRY(tc_gemm_launch(g, st));
  const int nkb1 = (int)(Kp / 64);
  const int kbps = (nkb1 + g.ksplit - 1) / g.ksplit;
  *ns = (nkb1 + kbps - 1) / kbps;
  return SMX_OK;
}

// =============================================================================================
// CSGU (speechbrain ConvolutionalSpatialGatingUnit, used by Branchformer.py:78-84), fused:
//   u (rows, 2H) bf16 = [value | gate];  out (rows, H) = gate_act( dwconv_reflect( LN(gate) ) + b ) * value
// Pass 1: per-row LayerNorm statistics of the gate half.  Pass 2: a block owns 32 frames x 128 channels of one utterance: the
// 32 + k - 1 normalised gate rows (reflect padding at the utterance edges) are staged in shared memory once, every thread
// convolves one channel over 16 frames with its taps in registers.
// =============================================================================================
__global__ void __launch_bounds__(256) csgu_stats_kernel(const __nv_bfloat16* __restrict__ u, int64_t rows, int H, float2* __restrict__ stats) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const __nv_bfloat16* g = u + row * (2 * (int64_t)H) + H;
  float s = 0.0f, q = 0.0f;
  const float x0 = __bfloat162float(g[0]);  // shift: no cancellation for rows with a large mean
  for (int c = lane * 8; c < H; c += 256) {
    const uint4 raw = *reinterpret_cast<const uint4*>(g + c);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(h[e]);
      const float a = f.x - x0, b = f.y - x0;
      s += a + b; q = fmaf(a, a, fmaf(b, b, q));
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if (lane == 0) {
    const float m0 = s / (float)H;
    const float rstd = rsqrtf(fmaxf(q / (float)H - m0 * m0, 0.0f) + 1e-5f);
    stats[row] = make_float2(rstd, -(m0 + x0) * rstd);
  }
}

constexpr int CS_FR = 32, CS_CH = 128;
template <int KS>
__global__ void __launch_bounds__(256) csgu_kernel(const __nv_bfloat16* __restrict__ u, const float2* __restrict__ stats,
                                                   const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                   const float* __restrict__ dw_w, const float* __restrict__ dw_b, int gate_act, int T,
                                                   int H, __nv_bfloat16* __restrict__ out, int64_t ldo) {
  constexpr int PAD = (KS - 1) / 2, NIN = CS_FR + KS - 1;
  __shared__ float sG[NIN][CS_CH];
  __shared__ float sW[CS_CH * KS];  // this block's taps, fetched with coalesced loads (per-thread rows 4 KS bytes apart throttled the load/store queue)
  const int b = blockIdx.z, t0 = blockIdx.x * CS_FR, c0 = blockIdx.y * CS_CH;
  const int tid = threadIdx.x;
  const int64_t ubase = (int64_t)b * T;
  {
    const int nch = H - c0 < CS_CH ? H - c0 : CS_CH;
    for (int i = tid; i < nch * KS; i += 256) sW[i] = dw_w[(size_t)c0 * KS + i];
  }
  for (int idx = tid; idx < NIN * (CS_CH / 8); idx += 256) {
    const int row = idx / (CS_CH / 8), ch = (idx % (CS_CH / 8)) * 8;
    int t = t0 - PAD + row;
    if (t < 0) t = -t;                       // reflect padding (F.pad(..., mode="reflect")): edge not repeated
    if (t >= T) t = 2 * (T - 1) - t;
    float v[8];
    if (t >= 0 && t < T && c0 + ch < H) {
      const float2 st = stats[ubase + t];
      const uint4 raw = *reinterpret_cast<const uint4*>(u + (ubase + t) * (2 * (int64_t)H) + H + c0 + ch);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        v[2 * e] = fmaf(fmaf(f.x, st.x, st.y), ln_w[c0 + ch + 2 * e], ln_b[c0 + ch + 2 * e]);
        v[2 * e + 1] = fmaf(fmaf(f.y, st.x, st.y), ln_w[c0 + ch + 2 * e + 1], ln_b[c0 + ch + 2 * e + 1]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.0f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sG[row][ch + e] = v[e];
  }
  __syncthreads();
  const int c = tid % CS_CH, fh = tid / CS_CH;  // channel, frame half (16 frames each)
  if (c0 + c >= H) return;
  float w[KS];
#pragma unroll
  for (int j = 0; j < KS; ++j) w[j] = sW[c * KS + j];
  const float bias = dw_b ? dw_b[c0 + c] : 0.0f;
#pragma unroll 1
  for (int fb = 0; fb < 2; ++fb) {   // two blocks of 8 frames
    float a[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) a[o] = bias;
    const int f0 = fh * 16 + fb * 8;
#pragma unroll
    for (int i = 0; i < KS + 7; ++i) {
      const float x = sG[f0 + i][c];
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int j = i - o;
        if (j >= 0 && j < KS) a[o] = fmaf(w[j], x, a[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int t = t0 + f0 + o;
      if (t < T) {
        const float val = __bfloat162float(u[(ubase + t) * (2 * (int64_t)H) + c0 + c]);
        out[(ubase + t) * ldo + c0 + c] = __float2bfloat16(tc::act_fast(gate_act, a[o]) * val);
      }
    }
  }
}


// ---- CSGU gate, second generation: 128 frames x 256 channels per block (halo overhead 158 / 128 instead of 62 / 32), normalised
// gate rows staged as fp32 pairs, a thread owns one channel PAIR and blocks of eight frames: taps and accumulators in registers,
// one packed fma.rn.f32x2 per tap and frame for both channels (the depthwise phase of K-CONV, smx_tc_conv.cu).  The first
// generation (32 x 128 tiles, scalar FMAs) ran at 13 % of the fp32 rate: 329 us for 32 000 x 1536 -- a third of a
// Branchformer-lite layer (profiles/r02_notes.md).
constexpr int CS2_FR = 128, CS2_CH = 128, CS2_THREADS = 256;   // 98 KB of shared memory per block: two blocks per SM (one stages while the other convolves), 4 frame blocks per thread
__device__ __forceinline__ float2 cs2_fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
template <int KS>
__global__ void __launch_bounds__(CS2_THREADS, 2) csgu2_kernel(const __nv_bfloat16* __restrict__ u, const float2* __restrict__ stats,
                                                               const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                               const float* __restrict__ dw_w, const float* __restrict__ dw_b, int gate_act,
                                                               int T, int H, __nv_bfloat16* __restrict__ out, int64_t ldo) {
  constexpr int PAD = (KS - 1) / 2, NIN = CS2_FR + KS - 1;
  extern __shared__ __align__(16) uint8_t cs2_smem[];
  float* sG = reinterpret_cast<float*>(cs2_smem);                 // [NIN][CS2_CH] normalised gate rows
  float* sW = sG + (size_t)NIN * CS2_CH;                          // [CS2_CH][KS] taps, then (ln_w | ln_b) of the block's channels
  float* sLw = sW + CS2_CH * KS;
  float* sLb = sLw + CS2_CH;
  const int b = blockIdx.z, t0 = blockIdx.x * CS2_FR, c0 = blockIdx.y * CS2_CH;
  const int tid = threadIdx.x;
  const int nch = H - c0 < CS2_CH ? H - c0 : CS2_CH;             // channels of this block (a multiple of 8)
  const int64_t ubase = (int64_t)b * T;
  for (int i = tid; i < nch * KS; i += CS2_THREADS) sW[i] = dw_w[(size_t)c0 * KS + i];
  for (int i = tid; i < CS2_CH; i += CS2_THREADS) { sLw[i] = i < nch ? ln_w[c0 + i] : 0.0f; sLb[i] = i < nch ? ln_b[c0 + i] : 0.0f; }
  __syncthreads();
  // stage LN(gate) for frames [t0 - PAD, t0 + 128 + PAD) with reflect padding (F.pad(..., mode="reflect"): the edge is not repeated);
  // four chunks per thread and round, all loads issued before the first is used (one block per SM: nothing else hides the latency)
  for (int base = 0; base < NIN * (CS2_CH / 8); base += 4 * CS2_THREADS) {
    uint4 raw[4];
    float2 st[4];
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = base + k * CS2_THREADS + tid;
      const int row = idx / (CS2_CH / 8), ch = (idx % (CS2_CH / 8)) * 8;
      int t = t0 - PAD + row;
      if (t < 0) t = -t;
      if (t >= T) t = 2 * (T - 1) - t;
      ok[k] = idx < NIN * (CS2_CH / 8) && t >= 0 && t < T && ch < nch;
      if (ok[k]) {
        st[k] = stats[ubase + t];
        raw[k] = *reinterpret_cast<const uint4*>(u + (ubase + t) * (2 * (int64_t)H) + H + c0 + ch);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = base + k * CS2_THREADS + tid;
      if (idx >= NIN * (CS2_CH / 8)) continue;
      const int row = idx / (CS2_CH / 8), ch = (idx % (CS2_CH / 8)) * 8;
      float v[8];
      if (ok[k]) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[k]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h[e]);
          v[2 * e] = fmaf(fmaf(f.x, st[k].x, st[k].y), sLw[ch + 2 * e], sLb[ch + 2 * e]);
          v[2 * e + 1] = fmaf(fmaf(f.y, st[k].x, st[k].y), sLw[ch + 2 * e + 1], sLb[ch + 2 * e + 1]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.0f;
      }
      float4* dst = reinterpret_cast<float4*>(sG + (size_t)row * CS2_CH + ch);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  __syncthreads();
  constexpr int pairs = CS2_CH / 2, nfbp = CS2_THREADS / pairs;  // 128 channel pairs, 3 frame-block groups
  const int pr = tid % pairs, fb0 = tid / pairs, cc = pr * 2;
  if (cc >= nch) return;
  float2 w2[KS];
#pragma unroll
  for (int j = 0; j < KS; ++j) w2[j] = make_float2(sW[cc * KS + j], sW[(cc + 1) * KS + j]);
  const float2 bias2 = make_float2(dw_b ? dw_b[c0 + cc] : 0.0f, dw_b ? dw_b[c0 + cc + 1] : 0.0f);
#pragma unroll 1
  for (int fb = fb0; fb < CS2_FR / 8; fb += nfbp) {
    if (t0 + fb * 8 >= T) break;
    float2 a[8];
    uint32_t valr[8];  // the value half of the block's eight frames, requested before the convolution: the loads land under the FMAs
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      a[o] = bias2;
      const int t = t0 + fb * 8 + o;
      valr[o] = t < T ? *reinterpret_cast<const uint32_t*>(u + (ubase + t) * (2 * (int64_t)H) + c0 + cc) : 0u;
    }
#pragma unroll
    for (int i = 0; i < KS + 7; ++i) {
      const float2 x = *reinterpret_cast<const float2*>(sG + (size_t)(fb * 8 + i) * CS2_CH + cc);
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int j = i - o;
        if (j >= 0 && j < KS) a[o] = cs2_fma2(w2[j], x, a[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int t = t0 + fb * 8 + o;
      if (t < T) {
        const float2 val = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&valr[o]));
        *reinterpret_cast<__nv_bfloat162*>(out + (ubase + t) * ldo + c0 + cc) =
            __floats2bfloat162_rn(tc::act_fast(gate_act, a[o].x) * val.x, tc::act_fast(gate_act, a[o].y) * val.y);
      }
    }
  }
}

size_t tc_csgu_workspace_bytes(int64_t rows) { return align_up((size_t)rows * sizeof(float2)); }
int tc_csgu_fwd(const __nv_bfloat16* u, int B, int T, int H, const float* ln_w, const float* ln_b, const float* dw_w, const float* dw_b,
                int kernel_size, int gate_act, __nv_bfloat16* out, int64_t ldo, void* stats_ws, cudaStream_t st) {
  if (kernel_size != 31 || H % 8) return fail(SMX_ERR_UNSUPPORTED, "csgu: kernel_size=%d H=%d", kernel_size, H);
  if ((kernel_size - 1) / 2 >= T) return fail(SMX_ERR_BAD_ARG, "reflect padding %d needs T > pad (T=%d)", (kernel_size - 1) / 2, T);
  const int64_t rows = (int64_t)B * T;
  csgu_stats_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(u, rows, H, (float2*)stats_ws);
  count_launch();
  SMX_TRY(check_launch("csgu_stats_kernel"));
  if (H % 8 == 0 && ldo % 2 == 0 && !getenv("SMX_CSGU_V1")) {
    constexpr int NIN = CS2_FR + 30;
    const size_t smem = ((size_t)NIN * CS2_CH + (size_t)CS2_CH * 31 + 2 * CS2_CH) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(csgu2_kernel<31>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(csgu2_kernel): %s", cudaGetErrorString(e));
    dim3 grid2((T + CS2_FR - 1) / CS2_FR, (H + CS2_CH - 1) / CS2_CH, B);
    csgu2_kernel<31><<<grid2, CS2_THREADS, smem, st>>>(u, (const float2*)stats_ws, ln_w, ln_b, dw_w, dw_b, gate_act, T, H, out, ldo);
    count_launch();
    return check_launch("csgu2_kernel");
  }
  dim3 grid((T + CS_FR - 1) / CS_FR, (H + CS_CH - 1) / CS_CH, B);
  csgu_kernel<31><<<grid, 256, 0, st>>>(u, (const float2*)stats_ws, ln_w, ln_b, dw_w, dw_b, gate_act, T, H, out, ldo);
  count_launch();
  return check_launch("csgu_kernel");
}

}  // namespace smx
