// tcgen05 arm of libsmx, part 13: K-GLU v4 -- the first half of the ConvolutionModule as one persistent kernel
//
//   g = (LN(x) W_v^T + b_v) * sigmoid(LN(x) W_g^T + b_g)                     Conformer.py:322-324 (LayerNorm, bottleneck conv, GLU)
//
// What changed against the GLU pass of smx_tc_cell3.cu (24 us at the bench shape; ncu: 40 % of its stall samples were waits
// for weight steps and accumulators), and why (profiles/r02_notes.md):
//   * a CTA keeps BOTH of its row tiles resident (tensor-map TMA, raw rows) and walks the 256 KB of weights ONCE: every 16 KB
//     weight unit (64 value rows + 64 gate rows of one K-block) feeds the MMAs of both tiles -- the old pass streamed the whole
//     image per tile through its ring and was bound by that stream;
//   * the LayerNorm is folded into the GEMM (gamma into the packed weights, beta into the bias, gw[n] = sum_k bf16(gamma_k W[n,k]);
//     the epilogue applies rstd and -mean rstd per row, like K-FFN): no in-place normalisation pass in front of the first MMA,
//     the raw bf16 rows are exact operands; the row statistics are computed by the prologue warps while the first MMAs run;
//   * the output is produced in four quarters of 64 channels (value + gate = one 128-column accumulator per tile and quarter,
//     double-buffered in TMEM: 2 tiles x 2 buffers x 128 = 512 columns): the epilogue of quarter q runs under the MMAs of q + 1;
//   * results leave through a 16 KB staging tile per epilogue group and a bulk tensor store (rows past the end are clipped).
// Warp roles:  0-7 epilogue of the CTA's first tile | 8-15 epilogue of its second tile | 16-19 row statistics
//              | 20 TMA producer (x tiles, weight ring) | 21 MMA issuer
// Built for D = 256 (the fused Conformer configuration); other widths / more than two tiles per CTA take the cell3 GLU pass.
#include <cuda.h>

#include "smx_tc.h"
#include "smx_tc_common.cuh"

namespace smx {

using tc::kblock_bytes;

constexpr int G4_D = 256, G4_NKB = G4_D / 64, G4_NQ = G4_D / 64;   // K-blocks of x; output quarters of 64 channels
constexpr int G4_THREADS = 22 * 32, G4_PRO_WARP0 = 16, G4_PROD_WARP = 20, G4_MMA_WARP = 21;
constexpr int G4_SLOTS = 3;
constexpr uint32_t G4_UNIT = 16384;                                 // one weight unit: 128 rows (value | gate) x 64 K
constexpr uint32_t G4_XTILE = G4_NKB * 16384u;                      // 64 KB
constexpr uint32_t G4_OFF_RING = 2 * G4_XTILE, G4_OFF_STAGE = G4_OFF_RING + G4_SLOTS * G4_UNIT, G4_OFF_PAR = G4_OFF_STAGE + 2 * 16384u,
                   G4_OFF_STAT = G4_OFF_PAR + 2 * 2 * G4_D * 4u, G4_SMEM = G4_OFF_STAT + 2 * 128 * 8u + 1024u;

struct Glu4P {
  int64_t rows;
  int n_tiles;
  const uint8_t* img;     // 16 units of 16 KB in issue order: quarter q, K-block kb -> [value rows 64 q.. | gate rows D + 64 q..]
  const float* gw;        // [2 D] gw[n] = sum_k bf16(gamma_k W[n,k]) (zeros without LayerNorm)
  const float* bias;      // [2 D] b[n] + sum_k beta_k W[n,k]
  int has_ln;
};

__device__ __forceinline__ void g4_tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void g4_tma_store_2d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0), "r"(c1),
               "r"(tc::smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void g4_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void g4_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void g4_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(G4_THREADS, 1) glu4_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_g,
                                                             const Glu4P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sRing = smem + G4_OFF_RING;
  uint8_t* sStage = smem + G4_OFF_STAGE;                              // [2 (tile)][128 rows][128 B], 128-byte swizzle
  float* sGw = reinterpret_cast<float*>(smem + G4_OFF_PAR);          // [2 D] gw / 2
  float* sBh = sGw + 2 * G4_D;                                       // [2 D] bias / 2
  float2* sStat = reinterpret_cast<float2*>(smem + G4_OFF_STAT);     // [2 (tile)][128] per-row (1/std, -mean/std)
  __shared__ __align__(8) uint64_t full_bar[G4_SLOTS], empty_bar[G4_SLOTS], x_land[2], stat_full[2], acc_full[2][2], acc_free[2][2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // this CTA's tiles: blockIdx.x and blockIdx.x + gridDim.x (the host guarantees n_tiles <= 2 gridDim.x)
  const int ntl = (int)blockIdx.x + (int)gridDim.x < p.n_tiles ? 2 : 1;

  if (warp == G4_PROD_WARP) {
    tc::tmem_alloc(&tmem_base_s, 512);
    if (lane == 0) {  // the x tiles are the first thing on the critical path: requested before the rest of the set-up
      tc::mbar_init(&x_land[0], 1); tc::mbar_init(&x_land[1], 1);
      tc::fence_barrier_init();
      tc::fence_proxy_async();
      tc::pdl_wait();  // x comes from the preceding kernel
      for (int t = 0; t < ntl; ++t) {
        const int row0 = ((int)blockIdx.x + t * (int)gridDim.x) * 128;
        tc::mbar_arrive_expect_tx(&x_land[t], G4_XTILE);
        for (int kb = 0; kb < G4_NKB; ++kb) g4_tma_load_2d(smem + (size_t)t * G4_XTILE + (size_t)kb * 16384, &tmap_x, kb * 64, row0, &x_land[t]);
      }
    }
    __syncwarp();
  }
  if (tid == 0) {
    for (int s = 0; s < G4_SLOTS; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int t = 0; t < 2; ++t) {
      tc::mbar_init(&stat_full[t], 128);
      for (int u = 0; u < 2; ++u) { tc::mbar_init(&acc_full[t][u], 1); tc::mbar_init(&acc_free[t][u], 8); }
    }
    tc::fence_barrier_init();
  }
  for (int i = tid; i < 2 * G4_D; i += G4_THREADS) { sGw[i] = 0.5f * p.gw[i]; sBh[i] = 0.5f * p.bias[i]; }  // (halved: see the epilogue)
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  tc::pdl_launch_dependents();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);

  if (warp == G4_PROD_WARP) {
    // =============================== weight ring: 16 units of 16 KB, once per CTA ===============================
    if (lane == 0) {
      int s = 0;
      uint32_t pe = 0;
      for (int u = 0; u < G4_NQ * G4_NKB; ++u) {
        tc::mbar_wait(&empty_bar[s], ((pe >> s) & 1u) ^ 1u);
        pe ^= 1u << s;
        tc::mbar_arrive_expect_tx(&full_bar[s], G4_UNIT);
        tc::bulk_g2s(sRing + (size_t)s * G4_UNIT, p.img + (size_t)u * G4_UNIT, G4_UNIT, &full_bar[s]);
        if (++s == G4_SLOTS) s = 0;
      }
    }
  } else if (warp == G4_MMA_WARP) {
    // =============================== MMA issuer ===============================
    // Quarter q accumulates into buffer q & 1 of each tile: [value 64 | gate 64] columns, K = 256 in four units.  A unit's weights
    // are used by the MMAs of both tiles before its ring slot is released.  Compile-time schedule (operands = base + immediate).
    constexpr uint32_t ID128 = tc::make_idesc_bf16(128, 128);
    constexpr uint32_t KB16 = 16384u >> 4;
    const uint64_t ring_d = tc::make_desc_sw128(tc::smem_u32(sRing));
    const uint64_t x_d = tc::make_desc_sw128(tc::smem_u32(smem));
    int s = 0;
    uint32_t pf = 0;
    for (int t = 0; t < ntl; ++t) tc::mbar_wait(&x_land[t], 0);
#pragma unroll
    for (int q = 0; q < G4_NQ; ++q) {
      const int buf = q & 1;
      if (q >= 2) {  // the epilogue of quarter q - 2 has drained this buffer
        for (int t = 0; t < ntl; ++t) tc::mbar_wait_spin(&acc_free[t][buf], 0);
      }
#pragma unroll
      for (int kb = 0; kb < G4_NKB; ++kb) {
        tc::mbar_wait_spin(&full_bar[s], (pf >> s) & 1u);
        pf ^= 1u << s;
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint64_t bd = ring_d + (uint64_t)((uint32_t)s * KB16);
          for (int t = 0; t < ntl; ++t) {
            const uint64_t ad = x_d + (uint64_t)((uint32_t)t * (G4_XTILE >> 4) + kb * KB16);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc::umma_bf16(tmem + t * 256 + buf * 128, ad + 2 * ks, bd + 2 * ks, ID128, (kb | ks) ? 1u : 0u);
          }
          tc::umma_commit(&empty_bar[s]);
          if (kb == G4_NKB - 1) {
            for (int t = 0; t < ntl; ++t) tc::umma_commit(&acc_full[t][buf]);
          }
        }
        __syncwarp();
        if (++s == G4_SLOTS) s = 0;
      }
    }
  } else if (warp >= G4_PRO_WARP0) {
    // =============================== row statistics (one thread per row, one pass shifted by the row's first element) ===============
    const int pw = warp - G4_PRO_WARP0;
    for (int t = 0; t < ntl; ++t) {
      tc::mbar_wait(&x_land[t], 0);
      float rs = 1.0f, nm = 0.0f;
      if (p.has_ln) {
        const int row = pw * 32 + lane;
        const uint8_t* rp = smem + (size_t)t * G4_XTILE + row * 128;
        const float x0 = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(rp + ((row & 7) << 4)));
        float2 s1 = make_float2(0.0f, 0.0f), s2 = make_float2(0.0f, 0.0f);
        const float2 sh = make_float2(-x0, -x0);
#pragma unroll 1
        for (int kb = 0; kb < G4_NKB; ++kb) {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint4 raw = *reinterpret_cast<const uint4*>(rp + (size_t)kb * 16384 + ((ch ^ (row & 7)) << 4));
            float2 v[4];
            tc::unpack_bf16x8_pairs(raw, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 d = tc::add2(v[e], sh); s1 = tc::add2(s1, d); s2 = tc::fma2(d, d, s2); }
          }
        }
        const float inv = 1.0f / (float)G4_D;
        const float m1 = (s1.x + s1.y) * inv;
        const float var = fmaxf((s2.x + s2.y) * inv - m1 * m1, 0.0f);
        rs = rsqrtf(var + 1e-5f);
        nm = -(x0 + m1) * rs;
      }
      sStat[t * 128 + pw * 32 + lane] = make_float2(rs, nm);
      tc::mbar_arrive(&stat_full[t]);  // every thread publishes its own row (release / acquire through the barrier)
    }
  } else {
    // =============================== epilogue: group t (eight warps) owns tile t ===============================
    const int t = warp >> 3;
    if (t < ntl) {
      const int qd = warp & 3, ch = (warp >> 2) & 1;  // TMEM lane quadrant; which 32 of the quarter's 64 channels
      const int r = qd * 32 + lane;
      const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
      const int row0 = ((int)blockIdx.x + t * (int)gridDim.x) * 128;
      uint8_t* const stage = sStage + (size_t)t * 16384;
      tc::mbar_wait(&stat_full[t], 0);
      const float2 st = sStat[t * 128 + r];
      // g = a * sigmoid(z) with sigmoid(z) = 1/2 + 1/2 tanh(z/2):  h = a/2, zh = z/2 (the staged parameters are halved, the row
      // scale too), g = h + h tanh(zh): two fma + one MUFU per output on top of the two affine LayerNorm corrections
      const float rs2 = 0.5f * st.x, nm = st.y;
#pragma unroll 1
      for (int q = 0; q < G4_NQ; ++q) {
        const int buf = q & 1;
        tc::mbar_wait(&acc_full[t][buf], (uint32_t)(q >> 1) & 1u);
        tc::tc_fence_after();
        if (q > 0) {  // the bulk store of the previous quarter has read the staging tile
          if ((tid & 255) == 0) g4_bulk_wait_read0();
          tc::named_bar_sync(7 + t, 256);
        }
#pragma unroll 1
        for (int pc = 0; pc < 2; ++pc) {
          const int c0 = ch * 32 + pc * 16;  // channel inside the quarter
          float a[16], z[16];
          g4_ld16(tmem + lane_sel + t * 256 + buf * 128 + c0, a);
          g4_ld16(tmem + lane_sel + t * 256 + buf * 128 + 64 + c0, z);
          tc::tmem_ld_wait();
          const int nv = q * 64 + c0, ng = G4_D + nv;
          const float4* gv = reinterpret_cast<const float4*>(sGw + nv);
          const float4* bv = reinterpret_cast<const float4*>(sBh + nv);
          const float4* gg = reinterpret_cast<const float4*>(sGw + ng);
          const float4* bg = reinterpret_cast<const float4*>(sBh + ng);
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 g1 = gv[i], b1 = bv[i], g2 = gg[i], b2 = bg[i];
            const float h0 = fmaf(rs2, a[4 * i], fmaf(nm, g1.x, b1.x)), h1 = fmaf(rs2, a[4 * i + 1], fmaf(nm, g1.y, b1.y));
            const float h2 = fmaf(rs2, a[4 * i + 2], fmaf(nm, g1.z, b1.z)), h3 = fmaf(rs2, a[4 * i + 3], fmaf(nm, g1.w, b1.w));
            const float z0 = fmaf(rs2, z[4 * i], fmaf(nm, g2.x, b2.x)), z1 = fmaf(rs2, z[4 * i + 1], fmaf(nm, g2.y, b2.y));
            const float z2 = fmaf(rs2, z[4 * i + 2], fmaf(nm, g2.z, b2.z)), z3 = fmaf(rs2, z[4 * i + 3], fmaf(nm, g2.w, b2.w));
            o[2 * i] = tc::pack_bf16x2(fmaf(h0, tc::tanh_approx(z0), h0), fmaf(h1, tc::tanh_approx(z1), h1));
            o[2 * i + 1] = tc::pack_bf16x2(fmaf(h2, tc::tanh_approx(z2), h2), fmaf(h3, tc::tanh_approx(z3), h3));
          }
          uint8_t* const rowp = stage + r * 128;
          const int k0 = c0 >> 3;  // first of this piece's two 16-byte chunks
          *reinterpret_cast<uint4*>(rowp + (((k0) ^ (r & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(rowp + (((k0 + 1) ^ (r & 7)) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_free[t][buf]);
        tc::fence_proxy_async();
        tc::named_bar_sync(7 + t, 256);
        if ((tid & 255) == 0) {
          g4_tma_store_2d(&tmap_g, stage, q * 64, row0);
          g4_bulk_commit();
        }
      }
      if ((tid & 255) == 0) g4_bulk_wait_read0();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == G4_PROD_WARP) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int g4_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
bool tc_glu4_supported(const smx_convmod_weights* w) {
  return w && w->bottleneck.in_dim == G4_D && w->bottleneck.out_dim == 2 * G4_D && w->bottleneck.n_split <= 1 && w->bottleneck.w && w->bottleneck.b;
}
bool tc_glu4_fits(int64_t rows) { return (rows + 127) / 128 <= 2 * (int64_t)g4_sms() && rows < 0x7fffff00; }

// image: [16 units x 16 KB] [gw f32 2D] [bias f32 2D] [pack-time scratch: W gamma fp32 (2D x D) | its chunk-major bf16 image]
struct Glu4Image { size_t gw, bias, s_wg, s_img, total; };
static Glu4Image g4_image() {
  Glu4Image im{};
  size_t off = (size_t)G4_NQ * G4_NKB * G4_UNIT;
  im.gw = off; off += align_up((size_t)2 * G4_D * 4, 1024);
  im.bias = off; off += align_up((size_t)2 * G4_D * 4, 1024);
  im.s_wg = off; off += align_up((size_t)2 * G4_D * G4_D * 4, 1024);
  im.s_img = off; off += align_up((size_t)2 * G4_D * G4_D * 2, 1024);
  im.total = off;
  return im;
}
size_t tc_glu4_packed_bytes(const smx_convmod_weights* w) { return tc_glu4_supported(w) ? g4_image().total : 0; }

// unit (q, kb) = 8 KB blocks (chunk q, kb) and (chunk D/64 + q, kb) of the chunk-major image ([64-row chunk][K-block])
__global__ void glu4_gather_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst) {
  const int u = blockIdx.x >> 1, part = blockIdx.x & 1, q = u / G4_NKB, kb = u % G4_NKB;
  const uint4* s = src + (size_t)((part * G4_NQ + q) * G4_NKB + kb) * 512;
  uint4* d = dst + (size_t)blockIdx.x * 512;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) d[i] = s[i];
}
__global__ void glu4_bias_kernel(const float* __restrict__ b, const float* __restrict__ bw, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (b ? b[i] : 0.0f) + bw[i];
}
int tc_glu4_pack(const smx_convmod_weights* w, void* out, cudaStream_t st) {
  if (!tc_glu4_supported(w)) return fail(SMX_ERR_UNSUPPORTED, "glu v4: configuration not handled");
  const Glu4Image im = g4_image();
  uint8_t* b = (uint8_t*)out;
  float* Wg = (float*)(b + im.s_wg);
  float* gw = (float*)(b + im.gw);
  float* bias = (float*)(b + im.bias);
  SMX_TRY(tc_fold_ln(w->bottleneck.w, G4_D, G4_D, 2 * G4_D, w->ln_w, w->ln_b, Wg, gw, bias, st));  // (bias holds W beta for a moment)
  glu4_bias_kernel<<<(2 * G4_D + 255) / 256, 256, 0, st>>>(w->bottleneck.b, bias, bias, 2 * G4_D);
  count_launch();
  SMX_TRY(check_launch("glu4_bias_kernel"));
  smx_linear Lg{};
  Lg.w = Wg; Lg.b = nullptr; Lg.in_dim = G4_D; Lg.out_dim = 2 * G4_D; Lg.n_split = 1;
  SMX_TRY(tc_pack_linear_nt(Lg, 0, G4_D, 64, b + im.s_img, st));
  glu4_gather_kernel<<<2 * G4_NQ * G4_NKB, 128, 0, st>>>((const uint4*)(b + im.s_img), (uint4*)b);
  count_launch();
  return check_launch("glu4_gather_kernel");
}

int tc_glu4_fwd(const smx_convmod_weights* w, const void* img, int64_t rows, const __nv_bfloat16* x, __nv_bfloat16* g, cudaStream_t st) {
  if (!tc_glu4_supported(w) || !tc_glu4_fits(rows)) return fail(SMX_ERR_UNSUPPORTED, "glu v4: shape not handled");
  const Glu4Image im = g4_image();
  Glu4P p{};
  p.rows = rows; p.n_tiles = (int)((rows + 127) / 128);
  p.img = (const uint8_t*)img;
  p.gw = (const float*)((const uint8_t*)img + im.gw);
  p.bias = (const float*)((const uint8_t*)img + im.bias);
  p.has_ln = w->ln_w != nullptr ? 1 : 0;
  CUtensorMap tx, tg;
  {
    const uint64_t dims[2] = {(uint64_t)G4_D, (uint64_t)rows}, strides[1] = {(uint64_t)G4_D * 2};
    const uint32_t box[2] = {64, 128};
    if (!tc_encode_tmap_bf16(&tx, x, 2, dims, strides, box) || !tc_encode_tmap_bf16(&tg, g, 2, dims, strides, box))
      return fail(SMX_ERR_CUDA, "glu v4: cuTensorMapEncodeTiled failed");
  }
  const unsigned grid = (unsigned)(p.n_tiles < g4_sms() ? p.n_tiles : g4_sms());
  cudaError_t e = cudaFuncSetAttribute(glu4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G4_SMEM);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(glu4_kernel): %s", cudaGetErrorString(e));
  e = launch_pdl(glu4_kernel, dim3(grid), dim3(G4_THREADS), (size_t)G4_SMEM, st, 1u, tx, tg, p);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaLaunchKernelEx(glu4_kernel): %s", cudaGetErrorString(e));
  count_tc_launch();
  return check_launch("glu4_kernel");
}

}  // namespace smx
