// tcgen05 arm of libsmx, part 4: K-FFN, the fused macaron feed-forward half-step
//
//   y = x + 0.5 * ( W2 @ act( W1 @ LN(x) + b1 ) + b2 )        [optionally y = LN_out(y)]
//   (Conformer.py:470-484, :518, :547)
//
// One CTA per 128-row tile.  LN(x) is written once to shared memory as the A operand.  The hidden
// dimension is walked in 64-wide chunks: GEMM1 (K=D, N=64) fills a TMEM buffer; the epilogue warps add
// b1, activate and write the bf16 chunk to shared memory as the A operand of GEMM2 (K=64, N=D), which
// accumulates the output tile in TMEM.  The d_ffn-wide hidden activation never leaves the SM.
// Per chunk, W1's rows and W2's columns arrive as ONE cp.async.bulk from the packed image.
#include "smx_tc.h"
#include "smx_tc_common.cuh"

namespace smx {

using tc::kblock_bytes;

constexpr int FFN_THREADS = 320;  // warps 0..7: epilogue, warp 8: producer, warp 9: MMA issuer (highest id = arbiter priority)
constexpr int FFN_HC = 64;        // hidden chunk width

struct FfnP {
  const __nv_bfloat16* x; __nv_bfloat16* y; int64_t rows;
  int D, F;
  const uint8_t* wp;           // packed chunks
  const float* ln_w; const float* ln_b;
  const float* b1; const float* b2;
  const float* oln_w; const float* oln_b; float oln_eps;
  int act;
  uint32_t stage_bytes, w2_off, tmem_cols;   // stage_bytes is a multiple of 1024 (128B-swizzle operands)
};

template <bool OLN>
__global__ void __launch_bounds__(FFN_THREADS, 1) ffn_kernel(const FfnP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int D = p.D, nkb = D / 64, nc = p.F / FFN_HC;
  uint8_t* sA = smem;                                   // LN(x) operand: nkb K-blocks of 128 rows
  uint8_t* sH = sA + (size_t)nkb * kblock_bytes(128);   // 2 hidden chunks (one K-block each)
  uint8_t* sW = sH + 2 * kblock_bytes(128);             // 2 weight stages
  __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2], acc1_full[2], acc1_empty[2], h_full[2], h_empty[2], acc2_full;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler (role dispatch below)
  const int64_t row0 = (int64_t)blockIdx.x * 128;
  const int nrows = p.rows - row0 < 128 ? (int)(p.rows - row0) : 128;

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, p.tmem_cols);
  if (tid == 32) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1);
      tc::mbar_init(&acc1_full[i], 1); tc::mbar_init(&acc1_empty[i], 8);
      tc::mbar_init(&h_full[i], 8); tc::mbar_init(&h_empty[i], 1);
    }
    tc::mbar_init(&acc2_full, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (tid == 256) {  // first two weight chunks stream in while the tile is normalised
    for (int j = 0; j < 2 && j < nc; ++j) {
      tc::mbar_arrive_expect_tx(&full_bar[j], p.stage_bytes);
      tc::bulk_g2s(sW + (size_t)j * p.stage_bytes, p.wp + (size_t)j * p.stage_bytes, p.stage_bytes, &full_bar[j]);
    }
  }

  // ---- prologue: LN(x tile) -> A operand ----------------------------------------------------------
  {
    const int nchunk = D / 8;
    constexpr int NW = FFN_THREADS / 32;
    for (int rb = warp; rb < 128; rb += 4 * NW) {  // four rows per round trip to memory
    uint4 rawb[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rb + j * NW;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ck = lane + 32 * c;
        rawb[j][c] = make_uint4(0, 0, 0, 0);
        if (ck < nchunk && r < nrows) rawb[j][c] = *reinterpret_cast<const uint4*>(p.x + (row0 + r) * D + ck * 8);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rb + j * NW;
      if (r >= 128) break;
      const bool live = r < nrows;
      float v[2][8];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ck = lane + 32 * c;
        if (ck < nchunk && live) {
          uint4 raw = rawb[j][c];
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(h[e]); v[c][2 * e] = f.x; v[c][2 * e + 1] = f.y; }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[c][e] = 0.0f;
        }
      }
      float s = 0.0f;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[c][e];
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / (float)D;
      float q = 0.0f;
#pragma unroll
      for (int c = 0; c < 2; ++c)
        if (lane + 32 * c < nchunk) {
#pragma unroll
          for (int e = 0; e < 8; ++e) { float d = v[c][e] - mean; q += d * d; }
        }
#pragma unroll
      for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q / (float)D + 1e-5f);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ck = lane + 32 * c;
        if (ck < nchunk) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = live ? (v[c][e] - mean) * rstd * p.ln_w[ck * 8 + e] + p.ln_b[ck * 8 + e] : 0.0f;
          *reinterpret_cast<uint4*>(sA + (size_t)(ck >> 3) * kblock_bytes(128) + tc::sw128_offset(r, ck & 7)) =
              make_uint4(tc::pack_bf16x2(o[0], o[1]), tc::pack_bf16x2(o[2], o[3]), tc::pack_bf16x2(o[4], o[5]), tc::pack_bf16x2(o[6], o[7]));
        }
      }
    }
    }
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  const uint32_t t_acc2 = tmem, t_acc1 = tmem + D;  // acc1 buffers at D and D+64

  if (warp == 8) {
    // =============================== producer ===============================
    if (lane == 0) {
      for (int j = 2; j < nc; ++j) {
        const int s = j & 1;
        tc::mbar_wait(&empty_bar[s], ((j >> 1) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&full_bar[s], p.stage_bytes);
        tc::bulk_g2s(sW + (size_t)s * p.stage_bytes, p.wp + (size_t)j * p.stage_bytes, p.stage_bytes, &full_bar[s]);
      }
    }
  } else if (warp == 9) {
    // =============================== MMA issuer ===============================
    // the whole warp walks the (warp-uniform) schedule; one elected lane issues tcgen05.mma / tcgen05.commit
    {
      const uint32_t a0 = tc::smem_u32(sA), h0 = tc::smem_u32(sH), w0 = tc::smem_u32(sW);
      const uint32_t idesc1 = tc::make_idesc_bf16(128, FFN_HC), idesc2 = tc::make_idesc_bf16(128, (uint32_t)D);
      auto gemm1 = [&](int j) {  // acc1[j&1] = LN(x) @ W1[j]^T
        const int s = j & 1;
        tc::mbar_wait(&full_bar[s], (j >> 1) & 1);
        tc::mbar_wait(&acc1_empty[s], ((j >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t wb = w0 + s * p.stage_bytes;
        if (tc::elect_one()) {
          for (int kb = 0; kb < nkb; ++kb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc::umma_bf16(t_acc1 + s * FFN_HC, tc::make_desc_sw128(a0 + kb * kblock_bytes(128) + ks * 32),
                            tc::make_desc_sw128(wb + kb * kblock_bytes(FFN_HC) + ks * 32), idesc1, (kb | ks) ? 1u : 0u);
          tc::umma_commit(&acc1_full[s]);
        }
        __syncwarp();
      };
      gemm1(0);
      for (int j = 0; j < nc; ++j) {
        if (j + 1 < nc) gemm1(j + 1);
        const int s = j & 1;
        tc::mbar_wait(&h_full[s], (j >> 1) & 1);
        tc::tc_fence_after();
        const uint32_t wb = w0 + s * p.stage_bytes + p.w2_off;
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)  // acc2 += H[j] @ W2[:, j]^T
            tc::umma_bf16(t_acc2, tc::make_desc_sw128(h0 + s * kblock_bytes(128) + ks * 32), tc::make_desc_sw128(wb + ks * 32),
                          idesc2, (j | ks) ? 1u : 0u);
          tc::umma_commit(&empty_bar[s]);
          tc::umma_commit(&h_empty[s]);
        }
        __syncwarp();
      }
      if (tc::elect_one()) tc::umma_commit(&acc2_full);
      __syncwarp();
    }
  } else {
    // =============================== epilogue ===============================
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const bool live = r < nrows;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    for (int j = 0; j < nc; ++j) {
      const int s = j & 1;
      tc::mbar_wait(&acc1_full[s], (j >> 1) & 1);
      tc::tc_fence_after();
      float v[32];
      tc::tmem_ld32(t_acc1 + lane_sel + s * FFN_HC + hf * 32, v);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc1_empty[s]);
      const float4* b1 = reinterpret_cast<const float4*>(p.b1 + j * FFN_HC + hf * 32);  // warp-uniform: broadcast loads
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = __ldg(b1 + i);
        v[4 * i] = tc::act_fast(p.act, v[4 * i] + b.x);
        v[4 * i + 1] = tc::act_fast(p.act, v[4 * i + 1] + b.y);
        v[4 * i + 2] = tc::act_fast(p.act, v[4 * i + 2] + b.z);
        v[4 * i + 3] = tc::act_fast(p.act, v[4 * i + 3] + b.w);
      }
      tc::mbar_wait(&h_empty[s], ((j >> 1) & 1) ^ 1);  // GEMM2 of chunk j-2 has finished reading this buffer
      uint8_t* hrow = sH + (size_t)s * kblock_bytes(128);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(hrow + tc::sw128_offset(r, hf * 4 + c)) =
            make_uint4(tc::pack_bf16x2(v[c * 8], v[c * 8 + 1]), tc::pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]),
                       tc::pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]), tc::pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]));
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&h_full[s]);
    }
    // ---- final: y = x + 0.5*(acc2 + b2)  [-> LN_out] ------------------------------------------------
    tc::mbar_wait(&acc2_full, 0);
    tc::tc_fence_after();
    const int half_cols = D / 2, npieces = half_cols / 32;
    const int64_t row = row0 + r;
    float keep[OLN ? 4 : 1][32];
    float s1 = 0.0f;
#pragma unroll
    for (int pc = 0; pc < 6; ++pc) {
      if (pc >= npieces) break;
      const int col = hf * half_cols + pc * 32;
      float v[32];
      tc::tmem_ld32(t_acc2 + lane_sel + col, v);
      tc::tmem_ld_wait();
      if (live) {
        const uint4* xp = reinterpret_cast<const uint4*>(p.x + row * D + col);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 raw = xp[c];
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 f = __bfloat1622float2(h[e]);
            const int i0 = c * 8 + 2 * e;
            v[i0] = fmaf(0.5f, v[i0] + p.b2[col + i0], f.x);
            v[i0 + 1] = fmaf(0.5f, v[i0 + 1] + p.b2[col + i0 + 1], f.y);
          }
        }
      }
      if (OLN) {
        if (pc < 4) {
#pragma unroll
          for (int i = 0; i < 32; ++i) { keep[pc][i] = v[i]; s1 += v[i]; }
        }
      } else if (live) {
        uint4* op = reinterpret_cast<uint4*>(p.y + row * D + col);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          op[c] = make_uint4(tc::pack_bf16x2(v[c * 8], v[c * 8 + 1]), tc::pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]),
                             tc::pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]), tc::pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]));
      }
    }
    if (OLN) {  // LayerNorm over the output row; the two column halves of a row live in two warps
      float* red = reinterpret_cast<float*>(sH);  // [2][2][128] — the hidden buffers are idle now
      red[(0 * 2 + hf) * 128 + r] = s1;
      tc::named_bar_sync(1, 256);
      const float mean = (red[r] + red[128 + r]) / (float)D;
      float s2 = 0.0f;
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) {
        if (pc >= npieces) break;
#pragma unroll
        for (int i = 0; i < 32; ++i) { float d = keep[pc][i] - mean; s2 += d * d; }
      }
      red[(2 + hf) * 128 + r] = s2;
      tc::named_bar_sync(1, 256);
      const float rstd = rsqrtf((red[256 + r] + red[384 + r]) / (float)D + p.oln_eps);
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) {
        if (pc >= npieces) break;
        const int col = hf * half_cols + pc * 32;
        if (live) {
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = (keep[pc][i] - mean) * rstd * p.oln_w[col + i] + p.oln_b[col + i];
          uint4* op = reinterpret_cast<uint4*>(p.y + row * D + col);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            op[c] = make_uint4(tc::pack_bf16x2(o[c * 8], o[c * 8 + 1]), tc::pack_bf16x2(o[c * 8 + 2], o[c * 8 + 3]),
                               tc::pack_bf16x2(o[c * 8 + 4], o[c * 8 + 5]), tc::pack_bf16x2(o[c * 8 + 6], o[c * 8 + 7]));
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static uint32_t ffn_stage_bytes(int D) { return (uint32_t)((D / 64) * kblock_bytes(FFN_HC) + kblock_bytes(D)); }

static bool ffn_fused_ok(const smx_ffn_weights* w) {
  const int D = w->w1.in_dim, F = w->w1.out_dim;
  if (w->w1.n_split > 1 || w->w2.n_split > 1 || w->w2.in_dim != F || w->w2.out_dim != D) return false;
  if (D % 64 || D < 64 || D > 256 || F % 64 || F < 64) return false;  // acc2 (D cols) + 2x64 acc1 cols fit 512 TMEM cols
  if (!w->w1.w || !w->w1.b || !w->w2.w || !w->w2.b || !w->ln_w || !w->ln_b) return false;
  return true;
}
// Wide models (conformer_large: D = 512, d_ffn = 2048, conformer_summarymixing.yaml:113-125): the output tile no longer fits the
// fused kernels' TMEM budget; three launches instead -- the input LayerNorm, K-GEMM (D -> F, activation) and K-GEMM (F -> D with
// the scaled residual as epilogue) -- with the normalised rows and the hidden activation making one bf16 round trip through
// memory.  (The first GEMM used to be K-LIN with its LayerNorm prologue: 546 us at B=32, T=1000 against 70 us for the same
// FLOPs on K-GEMM, whose operands stream by TMA under double-buffered accumulators: profiles/r02_notes.md.)
static bool ffn_wide_ok(const smx_ffn_weights* w) {
  const int D = w->w1.in_dim, F = w->w1.out_dim;
  if (w->w1.n_split > 1 || w->w2.n_split > 1 || w->w2.in_dim != F || w->w2.out_dim != D) return false;
  if (!w->w1.w || !w->w1.b || !w->w2.w || !w->w2.b || !w->ln_w || !w->ln_b) return false;
  return D > 256 && tc_gemm_supported(D, F) && tc_gemm_supported(F, D);
}
bool tc_ffn_supported(const smx_ffn_weights* w) { return ffn_fused_ok(w) || ffn_wide_ok(w); }
size_t tc_ffn_packed_bytes(const smx_ffn_weights* w) {
  if (!tc_ffn_supported(w)) return 0;
  if (ffn_wide_ok(w)) return 2 * align_up((size_t)w->w2.out_dim * w->w2.in_dim * 2, 1024);  // dense bf16 W1 | W2
  if (tc_ffn3_supported(w)) return tc_ffn3_packed_bytes(w);
  if (tc_ffn2_supported(w)) return tc_ffn2_packed_bytes(w);
  return (size_t)(w->w1.out_dim / FFN_HC) * ffn_stage_bytes(w->w1.in_dim);
}

// chunk j image: [W1 rows j*64.., K=D as D/64 K-blocks of 64 rows][W2 all D rows, K cols j*64.. (one K-block)]
__global__ void ffn_pack_kernel(const float* w1, const float* w2, int D, int F, uint32_t stage_bytes, uint32_t w2_off,
                                uint8_t* out) {
  const int j = blockIdx.y;
  const int nkb = D / 64;
  const int n1 = FFN_HC * (D / 8);  // 16-byte chunks of the W1 part
  const int n2 = D * 8;             // 16-byte chunks of the W2 part
  uint8_t* base = out + (size_t)j * stage_bytes;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
    if (i < n1) {
      const int r = i / (D / 8), ck = i % (D / 8);
      const float* src = w1 + (size_t)(j * FFN_HC + r) * D + ck * 8;
      *reinterpret_cast<uint4*>(base + (size_t)(ck >> 3) * kblock_bytes(FFN_HC) + tc::sw128_offset(r, ck & 7)) =
          make_uint4(tc::pack_bf16x2(src[0], src[1]), tc::pack_bf16x2(src[2], src[3]), tc::pack_bf16x2(src[4], src[5]), tc::pack_bf16x2(src[6], src[7]));
    } else {
      const int i2 = i - n1, r = i2 / 8, c16 = i2 % 8;
      const float* src = w2 + (size_t)r * F + j * FFN_HC + c16 * 8;
      *reinterpret_cast<uint4*>(base + w2_off + tc::sw128_offset(r, c16)) =
          make_uint4(tc::pack_bf16x2(src[0], src[1]), tc::pack_bf16x2(src[2], src[3]), tc::pack_bf16x2(src[4], src[5]), tc::pack_bf16x2(src[6], src[7]));
    }
  }
  (void)nkb;
}

int tc_ffn_pack(const smx_ffn_weights* w, void* packed, cudaStream_t st) {
  if (!tc_ffn_supported(w)) return fail(SMX_ERR_UNSUPPORTED, "ffn configuration not handled by the tensor-core arm");
  if (ffn_wide_ok(w)) {
    SMX_TRY(tc_dense_bf16(w->w1, 0, w->w1.in_dim, packed, st));
    return tc_dense_bf16(w->w2, 0, w->w2.in_dim, (char*)packed + align_up((size_t)w->w1.out_dim * w->w1.in_dim * 2, 1024), st);
  }
  if (tc_ffn3_supported(w)) return tc_ffn3_pack(w, packed, st);
  if (tc_ffn2_supported(w)) return tc_ffn2_pack(w, packed, st);
  const int D = w->w1.in_dim, F = w->w1.out_dim;
  const uint32_t sb = ffn_stage_bytes(D), w2_off = (D / 64) * kblock_bytes(FFN_HC);
  dim3 grid(16, F / FFN_HC);
  ffn_pack_kernel<<<grid, 256, 0, st>>>(w->w1.w, w->w2.w, D, F, sb, w2_off, (uint8_t*)packed);
  count_launch();
  return check_launch("ffn_pack_kernel");
}

size_t tc_ffn_workspace_bytes(const smx_ffn_weights* w, int64_t rows) {
  return ffn_wide_ok(w) ? align_up((size_t)rows * w->w1.out_dim * 2) + align_up((size_t)rows * w->w1.in_dim * 2) : 0;  // the fused kernels need no scratch
}

int tc_ffn_fwd(const smx_ffn_weights* w, const void* packed, int act, int64_t rows, const __nv_bfloat16* x,
               const float* oln_w, const float* oln_b, float oln_eps, __nv_bfloat16* y, Arena& ws, cudaStream_t st) {
  if (ffn_wide_ok(w)) {
    const int D = w->w1.in_dim, F = w->w1.out_dim;
    const size_t m0 = ws.mark();
    __nv_bfloat16* h = (__nv_bfloat16*)ws.take((size_t)rows * F * 2);
    __nv_bfloat16* xn = (__nv_bfloat16*)ws.take((size_t)rows * D * 2);
    if (!h || !xn) return fail(SMX_ERR_WORKSPACE, "workspace too small (wide FFN)");
    if (!ws.dry) {
      SMX_TRY(layernorm(x, SMX_BF16, D, w->ln_w, w->ln_b, 1e-5f, SMX_ACT_IDENTITY, xn, SMX_BF16, D, rows, D, st));  // Conformer.py:470
      GemmTc g1{};
      g1.a = xn; g1.lda = D; g1.M = rows; g1.N = F; g1.K = D;
      g1.w = (const __nv_bfloat16*)packed;
      g1.bias = w->w1.b; g1.act = act; g1.alpha = 1.0f; g1.out = h; g1.ldo = F;
      SMX_TRY(tc_gemm_launch(g1, st));                                                  // :471-473
      GemmTc g{};
      g.a = h; g.lda = F; g.M = rows; g.N = D; g.K = F;
      g.w = (const __nv_bfloat16*)((const char*)packed + align_up((size_t)F * D * 2, 1024));
      g.bias = w->w2.b; g.act = SMX_ACT_IDENTITY; g.resid = x; g.ldr = D; g.alpha = 0.5f; g.out = y; g.ldo = D;
      SMX_TRY(tc_gemm_launch(g, st));                                                   // x + 0.5 * ffn(x), :518
      if (oln_w) SMX_TRY(layernorm(y, SMX_BF16, D, oln_w, oln_b, oln_eps, SMX_ACT_IDENTITY, y, SMX_BF16, D, rows, D, st));  // norm2, :547
    }
    ws.release(m0);
    return SMX_OK;
  }
  if (ws.dry) return SMX_OK;
  // v3 (hidden activation in tensor memory) moves rows with 256-bit global accesses: 32-byte aligned x / y
  if (tc_ffn_version() == 3 && tc_ffn3_supported(w) && ((uintptr_t)x % 32 == 0) && ((uintptr_t)y % 32 == 0))
    return tc_ffn3_fwd(w, packed, act, rows, x, oln_w, oln_b, oln_eps, y, st);
  if (tc_ffn2_supported(w)) return tc_ffn2_fwd(w, packed, act, rows, x, oln_w, oln_b, oln_eps, y, st);
  const int D = w->w1.in_dim, F = w->w1.out_dim;
  FfnP p{};
  p.x = x; p.y = y; p.rows = rows; p.D = D; p.F = F;
  p.wp = (const uint8_t*)packed;
  p.ln_w = w->ln_w; p.ln_b = w->ln_b; p.b1 = w->w1.b; p.b2 = w->w2.b;
  p.oln_w = oln_w; p.oln_b = oln_b; p.oln_eps = oln_eps;
  p.act = act;
  p.stage_bytes = ffn_stage_bytes(D);
  p.w2_off = (D / 64) * kblock_bytes(FFN_HC);
  uint32_t cols = 32;
  while (cols < (uint32_t)(D + 2 * FFN_HC)) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = (size_t)(D / 64) * kblock_bytes(128) + 2 * kblock_bytes(128) + 2 * (size_t)p.stage_bytes;
  const unsigned grid = (unsigned)((rows + 127) / 128);
  cudaError_t e;
  if (oln_w) {
    e = cudaFuncSetAttribute(ffn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(ffn_kernel): %s", cudaGetErrorString(e));
    ffn_kernel<true><<<grid, FFN_THREADS, smem, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(ffn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(ffn_kernel): %s", cudaGetErrorString(e));
    ffn_kernel<false><<<grid, FFN_THREADS, smem, st>>>(p);
  }
  count_tc_launch();
  return check_launch("ffn_kernel");
}

}  // namespace smx
