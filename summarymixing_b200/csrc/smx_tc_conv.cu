// tcgen05 arm of libsmx, part 7: K-CONV, the second half of the ConvolutionModule fused into one persistent kernel
//
//   y = x + mask * ( W_out @ act( LN( dwconv_k(g) ) ) + b_out )                       Conformer.py:325-338, :543
//
// where g = GLU(W_bottleneck @ LN(x)) is produced by the GLU pass of the fused cell skeleton (smx_tc_cell.cu, PHASE 2).
// One CTA per SM walks utterance-aligned 128-frame tiles:
//   warps 0-11 (1) stage the (128 + k - 1) frames of g the tile needs in shared memory (zero rows outside the utterance:
//                  the reference's zero padding, Conformer.py:142-151);
//              (2) depthwise conv: a thread owns one channel pair and blocks of eight consecutive frames, taps and
//                  accumulators in registers as fp32 pairs; one packed fma.rn.f32x2 (FFMA2) updates both channels
//                  (8*k FFMA2 per k+7 shared-memory loads); LayerNorm statistics per frame
//                  by warp shuffles + one small shared-memory exchange; normalise, activate and write the bf16 A
//                  operand (128B swizzle) for the tensor core;
//              (4) epilogue of the output GEMM: + bias, * mask, + residual (parked in the idle g buffer with coalesced
//                  loads issued while the GEMM runs), staged tile, coalesced stores
//   warp 12    weight producer (32 KB steps of the schedule-ordered W_out image through a shared-memory ring: one
//              cp.async.bulk, one mbarrier pair and one tcgen05.commit per step)
//   warp 13    (3) MMA issuer: tcgen05.mma, N = up to 256 per instruction, accumulator in TMEM
// The depthwise output, its LayerNorm and the activation never exist in global memory.
#include "smx_tc.h"
#include "smx_tc_common.cuh"

namespace smx {

using tc::kblock_bytes;

constexpr int CV_NCW = 12;                    // compute warps, three per TMEM lane quadrant (16 would need < 100 registers: spills)
constexpr int CV_CT = CV_NCW * 32;            // compute threads
constexpr int CV_THREADS = CV_CT + 64;
constexpr int CV_PROD_WARP = CV_NCW, CV_MMA_WARP = CV_NCW + 1;
constexpr int CV_STAGES = 2;                  // ring slots of 32 KB (one step each)
constexpr uint32_t CV_BLOCK = 8192, CV_SLOT = 32768;

struct ConvFP {
  const __nv_bfloat16* g;
  const float* dw_w; const float* dw_b; const float* ln_w; const float* ln_b;
  const uint8_t* w_img; const float* b_out;
  const uint8_t* mask; const __nv_bfloat16* resid; __nv_bfloat16* y;
  int B, T, D, tpu, n_tiles, act, gw;
  int al32;  // y is 32-byte aligned: rows leave with 256-bit stores (else two 128-bit stores)
  unsigned long long* trace;
  uint32_t off_a, off_ring, off_par, off_stat;
};

// both channels of a pair in one instruction: d = a * b + c on (x, y) fp32 pairs (FFMA2)
__device__ __forceinline__ float2 cv_fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
// warp-collective 16-column fp32 TMEM load (lane i <-> TMEM lane 32*(w%4)+i)
__device__ __forceinline__ void cv_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void cv_stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ uint4 cv_pack8(const float* v) {
  return make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void cv_unpack8(const uint4& raw, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
}

#define CV_TRACE(it, ev)                                                                                  \
  do {                                                                                                    \
    if (p.trace && blockIdx.x == 0 && tid == 0 && (it) < 4) p.trace[(it) * 16 + (ev)] = clock64();          \
  } while (0)

template <int K, int ACT, int D>  // D (64, 128, 256) is compile time: immediate shared-memory offsets, unrolled statistics
__global__ void __launch_bounds__(CV_THREADS, 1) conv_kernel(const ConvFP p) {
  constexpr int PAD = (K - 1) / 2, NIN = 128 + K - 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  __nv_bfloat16* sG = reinterpret_cast<__nv_bfloat16*>(smem);   // [NIN][D] staged GLU output; later the residual/output tile
  uint8_t* sStage = smem;
  uint8_t* sA = smem + p.off_a;
  uint8_t* sRing = smem + p.off_ring;
  float* sPar = reinterpret_cast<float*>(smem + p.off_par);     // [b_out | ln_w | ln_b], 256 floats each
  float* sStat = reinterpret_cast<float*>(smem + p.off_stat);   // [2][CV_NCW warps][8 frames][2]
  __shared__ __align__(8) uint64_t full_bar[CV_STAGES], empty_bar[CV_STAGES], a_full, acc_full;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  constexpr int nkb = D / 64;
  const int act = ACT >= 0 ? ACT : p.act;

  tc::pdl_launch_dependents();
  if (warp == CV_PROD_WARP) tc::tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    for (int s = 0; s < CV_STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&a_full, CV_NCW); tc::mbar_init(&acc_full, 1);
    tc::fence_barrier_init();
  }
  for (int i = tid; i < 256; i += CV_THREADS) {
    sPar[i] = i < D ? p.b_out[i] : 0.0f;
    sPar[256 + i] = i < D ? p.ln_w[i] : 1.0f;
    sPar[512 + i] = i < D ? p.ln_b[i] : 0.0f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  const int first_tile = blockIdx.x, tile_step = gridDim.x;

  // W_out streams through a step-granular ring: its image is in schedule order ([K-block][64-row chunk]), a step is up to
  // four 8 KB blocks = (4 / gw) K-blocks of the N = D wide GEMM: one bulk copy, one barrier pair, one commit per step.
  constexpr int UPS = 4 / nkb;                    // K-blocks (MMA units) per step: gw == nkb chunks per unit
  constexpr int NSTEP = (nkb + UPS - 1) / UPS;
  constexpr uint32_t STEP_BYTES = (uint32_t)(UPS < nkb ? UPS : nkb) * nkb * CV_BLOCK;
  if (warp == CV_PROD_WARP) {
    // =============================== weight producer ===============================
    if (lane == 0) {
      int s = 0;
      uint32_t pe = 0;
      for (int tile = first_tile; tile < p.n_tiles; tile += tile_step)
        for (int st = 0; st < NSTEP; ++st) {
          tc::mbar_wait(&empty_bar[s], ((pe >> s) & 1u) ^ 1u);  // suspending wait: a polling producer floods the SM sub-partition's shared-memory queue
          pe ^= 1u << s;
          tc::mbar_arrive_expect_tx(&full_bar[s], STEP_BYTES);
          tc::bulk_g2s(sRing + (size_t)s * CV_SLOT, p.w_img + (size_t)st * STEP_BYTES, STEP_BYTES, &full_bar[s]);
          if (++s == CV_STAGES) s = 0;
        }
    }
  } else if (warp == CV_MMA_WARP) {
    // =============================== MMA issuer ===============================
    int s = 0, it = 0;
    uint32_t pf = 0;
    const uint32_t a0 = tc::smem_u32(sA), r0 = tc::smem_u32(sRing);
    const uint32_t idesc = tc::make_idesc_bf16(128, (uint32_t)D);
    for (int tile = first_tile; tile < p.n_tiles; tile += tile_step, ++it) {
      tc::mbar_wait(&a_full, it & 1);  // also: the previous tile's accumulator has been drained (same warps, program order)
      tc::tc_fence_after();
      for (int st = 0; st < NSTEP; ++st) {
        tc::mbar_wait_spin(&full_bar[s], (pf >> s) & 1u);
        pf ^= 1u << s;
        tc::tc_fence_after();
        if (tc::elect_one()) {
#pragma unroll
          for (int u = 0; u < UPS; ++u) {
            const int kb = st * UPS + u;
            if (kb < nkb) {
              const uint64_t ad = tc::make_desc_sw128(a0 + kb * kblock_bytes(128));
              const uint64_t bd = tc::make_desc_sw128(r0 + (uint32_t)s * CV_SLOT + (uint32_t)u * nkb * CV_BLOCK);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) tc::umma_bf16(tmem, ad + 2u * ks, bd + 2u * ks, idesc, (kb == 0 && ks == 0) ? 0u : 1u);
            }
          }
          tc::umma_commit(&empty_bar[s]);
          if (st == NSTEP - 1) tc::umma_commit(&acc_full);
        }
        __syncwarp();
        if (++s == CV_STAGES) s = 0;
      }
    }
  } else if (warp < CV_NCW) {
    // =============================== compute warps ===============================
    const int q = warp & 3, grp = warp >> 2;  // TMEM lane quadrant; 64-column chunk (chunks grp, grp + 4, ...)
    const int r = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    constexpr int pairs = D / 2, nfbp = CV_CT / pairs, wpf = pairs / 32;
    const int pr = tid % pairs, fb0 = tid / pairs;
    const int c0 = pr * 2;                      // this thread's channel pair
    const int wig = (tid % pairs) >> 5;         // warp index inside its frame-block group
    float2 w2[K];  // taps of both channels
    {
      // The D x K taps pass through shared memory (the A-operand buffer, idle until the first tile's convolution has been
      // computed): one coalesced sweep by all compute threads, then every thread picks up its 2 K values.  (Read straight from
      // global memory -- 62 scalar loads per thread, 124 bytes apart between lanes -- they kept the load/store queue
      // throttled for a quarter of the kernel's stall samples: profiles/r02_notes.md.)
      float* sTap = reinterpret_cast<float*>(sA);
      if ((reinterpret_cast<uintptr_t>(p.dw_w) & 15) == 0 && (D * K) % 4 == 0) {  // all chunks in flight at once
        for (int i = tid; i < D * K / 4; i += CV_CT) tc::cp_async16(sTap + 4 * i, p.dw_w + 4 * i, 16u);
        tc::cp_async_commit();
        tc::cp_async_wait_all();
      } else {
        for (int i = tid; i < D * K; i += CV_CT) sTap[i] = p.dw_w[i];
      }
      tc::named_bar_sync(1, CV_CT);
#pragma unroll
      for (int j = 0; j < K; ++j) w2[j] = make_float2(sTap[c0 * K + j], sTap[(c0 + 1) * K + j]);
      tc::named_bar_sync(1, CV_CT);  // before the first A-operand rows are written
    }
    const float2 bias2 = make_float2(p.dw_b ? p.dw_b[c0] : 0.0f, p.dw_b ? p.dw_b[c0 + 1] : 0.0f);
    tc::pdl_wait();  // g / residual come from the preceding kernels (the taps above are parameters)
    const float2 lw2 = make_float2(sPar[256 + c0], sPar[256 + c0 + 1]), lb2 = make_float2(sPar[512 + c0], sPar[512 + c0 + 1]);
    constexpr float invD = 1.0f / (float)D;
    constexpr int cpr = D / 8;
    const int rr0 = tid / cpr, ch0 = tid - rr0 * cpr, drr = CV_CT / cpr, dch = CV_CT - drr * cpr;
    constexpr int NST = (128 * 32 + CV_CT - 1) / CV_CT;       // 16-byte chunks of a D = 256 tile per thread (residual / output staging)
    uint32_t sp = 0;  // statistics buffer parity
    int it = 0;
    for (int tile = first_tile; tile < p.n_tiles; tile += tile_step, ++it) {
      const int b = tile / p.tpu, t0 = (tile % p.tpu) * 128;
      const int64_t row0 = (int64_t)b * p.T + t0;
      const int nrows = p.T - t0 < 128 ? p.T - t0 : 128;
      CV_TRACE(it, 0);
      // ---- (1) stage g[t0 - PAD, t0 + 128 + PAD) -----------------------------------------------------
      for (int idx = tid; idx < NIN * cpr; idx += CV_CT) {
        const int row = idx / cpr, ch = idx - row * cpr, u = t0 - PAD + row;
        const bool in = u >= 0 && u < p.T;
        tc::cp_async16(sG + (size_t)idx * 8, in ? (const void*)(p.g + ((int64_t)b * p.T + u) * D + ch * 8) : (const void*)p.g, in ? 16u : 0u);
      }
      tc::cp_async_commit();
      tc::cp_async_wait_all();
      tc::named_bar_sync(1, CV_CT);
      CV_TRACE(it, 1);
      // ---- (2) depthwise conv -> LayerNorm -> activation -> A operand --------------------------------
      for (int fb = fb0; fb < 16; fb += nfbp) {
        float2 a[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) a[o] = bias2;
#pragma unroll
        for (int i = 0; i < K + 7; ++i) {
          const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sG + (size_t)(fb * 8 + i) * D + c0));
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const int j = i - o;
            if (j >= 0 && j < K) a[o] = cv_fma2(w2[j], x, a[o]);
          }
        }
        // per-frame statistics over the D channels: this thread's pair -> warp -> the group's warps
        // 16 partial sums (8 frames x {sum, sum of squares}) reduced over the warp with a halving butterfly: 16 shuffles
        // instead of 80; lane l (even) ends with the warp total of value (l >> 1)
        float pv[16];
#pragma unroll
        for (int o = 0; o < 8; ++o) { pv[o] = a[o].x + a[o].y; pv[8 + o] = fmaf(a[o].x, a[o].x, a[o].y * a[o].y); }
#pragma unroll
        for (int sh = 16, n = 8; sh >= 2; sh >>= 1, n >>= 1) {
          const bool up = (lane & sh) != 0;
#pragma unroll
          for (int j = 0; j < n; ++j) {
            const float send = up ? pv[j] : pv[j + n];
            const float keepv = up ? pv[j + n] : pv[j];
            pv[j] = keepv + __shfl_xor_sync(0xffffffffu, send, sh);
          }
        }
        pv[0] += __shfl_xor_sync(0xffffffffu, pv[0], 1);
        float* st = sStat + (size_t)sp * (CV_NCW * 16);
        if ((lane & 1) == 0) {  // value index = lane bits 4..1: bit 4 selects {sum, sumsq}, bits 3..1 the frame
          const int vi = lane >> 1;
          st[(warp * 8 + (vi & 7)) * 2 + (vi >> 3)] = pv[0];
        }
        if (wpf > 1) tc::named_bar_sync(2 + fb0, wpf * 32); else __syncwarp();
        // lane l < 16 adds the group's warp totals of value l (frame l & 7, statistic l >> 3); lanes 0..7 then hold the
        // frame's mean and 1/std, which every lane fetches with one shuffle per frame
        const int wbase = warp - wig;  // first warp of this group
        float tot = 0.0f;
#pragma unroll
        for (int w = 0; w < wpf; ++w) tot += st[((wbase + w) * 8 + (lane & 7)) * 2 + ((lane >> 3) & 1)];
        const float sq = __shfl_xor_sync(0xffffffffu, tot, 8);
        const float mean_l = tot * invD;
        const float rstd_l = rsqrtf(fmaxf(sq * invD - mean_l * mean_l, 0.0f) + 1e-5f);
        uint8_t* const arow = sA + (size_t)(c0 >> 6) * kblock_bytes(128) + (c0 & 7) * 2;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const float mean = __shfl_sync(0xffffffffu, mean_l, o), rstd = __shfl_sync(0xffffffffu, rstd_l, o);
          const float2 d = make_float2(a[o].x - mean, a[o].y - mean);
          const float2 t = cv_fma2(d, make_float2(rstd * lw2.x, rstd * lw2.y), lb2);
          float v[2] = {t.x, t.y};
          tc::act_apply<2>(act, v);
          const int row = fb * 8 + o;
          *reinterpret_cast<uint32_t*>(arow + tc::sw128_offset(row, (c0 & 63) >> 3)) = tc::pack_bf16x2(v[0], v[1]);
        }
        sp ^= 1u;
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&a_full);
      CV_TRACE(it, 2);
      // ---- (4) epilogue of the output GEMM -----------------------------------------------------------
      tc::named_bar_sync(1, CV_CT);  // every warp has finished reading the staged g tile
      CV_TRACE(it, 3);
      {  // the residual tile goes straight to shared memory (swizzled staging layout) while the output GEMM runs
        int rr = rr0, ch = ch0;
#pragma unroll 4
        for (int k = 0; k < NST; ++k) {
          if (rr < 128) {
            const bool in = p.resid && rr < nrows;
            tc::cp_async16(sStage + (size_t)(ch >> 3) * kblock_bytes(128) + tc::sw128_offset(rr, ch & 7),
                           in ? (const void*)(p.resid + (row0 + rr) * D + ch * 8) : (const void*)p.g, in ? 16u : 0u);
          }
          rr += drr; ch += dch;
          if (ch >= cpr) { ch -= cpr; ++rr; }
        }
        tc::cp_async_commit();
      }
      const bool live = r < nrows;
      const float rscale = live ? (p.mask ? (float)p.mask[row0 + r] : 1.0f) : 0.0f;
      tc::mbar_wait(&acc_full, it & 1);
      tc::tc_fence_after();
      tc::cp_async_wait_all();
      CV_TRACE(it, 4);
      tc::named_bar_sync(1, CV_CT);   // the residual tile is in shared memory
      // this thread: row r, 16-column pieces grp, grp + 3, ...: TMEM -> + bias, * mask, + residual -> one 256-bit store
      for (int pc = grp; pc < D / 16; pc += CV_NCW / 4) {
        const int col = pc * 16;
        float v[16];
        cv_ld16(tmem + lane_sel + col, v);
        tc::tmem_ld_wait();
        const float4* bp = reinterpret_cast<const float4*>(sPar + col);
        float f[16];
        cv_unpack8(*reinterpret_cast<const uint4*>(sStage + (size_t)(col >> 6) * kblock_bytes(128) + tc::sw128_offset(r, (col & 63) >> 3)), f);
        cv_unpack8(*reinterpret_cast<const uint4*>(sStage + (size_t)(col >> 6) * kblock_bytes(128) + tc::sw128_offset(r, ((col & 63) >> 3) + 1)), f + 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 bb = bp[i];
          v[4 * i] = fmaf(v[4 * i] + bb.x, rscale, f[4 * i]);          // out*mask (:338) then x + out (:543)
          v[4 * i + 1] = fmaf(v[4 * i + 1] + bb.y, rscale, f[4 * i + 1]);
          v[4 * i + 2] = fmaf(v[4 * i + 2] + bb.z, rscale, f[4 * i + 2]);
          v[4 * i + 3] = fmaf(v[4 * i + 3] + bb.w, rscale, f[4 * i + 3]);
        }
        if (live) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = tc::pack_bf16x2(v[2 * i], v[2 * i + 1]);
          __nv_bfloat16* dst = p.y + (row0 + r) * D + col;
          if (p.al32) cv_stg256(dst, o);
          else { *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]); *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]); }
        }
      }
      tc::tc_fence_before();
      CV_TRACE(it, 5);
      tc::named_bar_sync(1, CV_CT);   // the staged tile (= the next tile's g buffer) is free; the accumulator is drained
      CV_TRACE(it, 6);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == CV_PROD_WARP) tc::tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool tc_convf_supported(const smx_convmod_weights* w, int chunk) {
  const int D = w->bottleneck.in_dim;
  if (chunk > 0 || w->causal) return false;
  if (w->kernel_size != 31) return false;
  if (D != 64 && D != 128 && D != 256) return false;
  if (w->bottleneck.out_dim != 2 * D || w->out.in_dim != D || w->out.out_dim != D) return false;
  if (w->bottleneck.n_split > 1 || w->out.n_split > 1) return false;
  if (!w->bottleneck.w || !w->bottleneck.b || !w->out.w || !w->out.b || !w->dw_w || !w->ln_w || !w->ln_b || !w->after_ln_w ||
      !w->after_ln_b)
    return false;
  return true;
}

static std::atomic<unsigned long long*> g_trace3{nullptr};
void tc_set_trace_conv(void* p) { g_trace3 = (unsigned long long*)p; }

static int convf_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int ACT, int D>
static int launch_conv_d(const ConvFP& p, unsigned grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(conv_kernel<31, ACT, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(conv_kernel): %s", cudaGetErrorString(e));
  e = launch_pdl(conv_kernel<31, ACT, D>, dim3(grid), dim3(CV_THREADS), smem, st, 1u, p);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaLaunchKernelEx(conv_kernel): %s", cudaGetErrorString(e));
  count_tc_launch();
  return check_launch("conv_kernel");
}
template <int ACT>
static int launch_conv(const ConvFP& p, unsigned grid, size_t smem, cudaStream_t st) {
  switch (p.D) {
    case 256: return launch_conv_d<ACT, 256>(p, grid, smem, st);
    case 128: return launch_conv_d<ACT, 128>(p, grid, smem, st);
    default: return launch_conv_d<ACT, 64>(p, grid, smem, st);
  }
}

// g: GLU output (B,T,D) bf16; img_out: packed after_conv.2 weight (64 x 64 blocks) in schedule order ([K-block][chunk])
int tc_convf_second_half(const smx_convmod_weights* w, const void* img_out, int act, int B, int T, const __nv_bfloat16* g,
                         const uint8_t* mask, const __nv_bfloat16* residual, __nv_bfloat16* y, cudaStream_t st) {
  const int D = w->bottleneck.in_dim;
  ConvFP p{};
  p.g = g; p.dw_w = w->dw_w; p.dw_b = w->dw_b; p.ln_w = w->after_ln_w; p.ln_b = w->after_ln_b;
  p.w_img = (const uint8_t*)img_out; p.b_out = w->out.b; p.mask = mask; p.resid = residual; p.y = y;
  p.trace = g_trace3;
  p.al32 = ((uintptr_t)y % 32 == 0) ? 1 : 0;
  p.B = B; p.T = T; p.D = D; p.tpu = (T + 127) / 128; p.n_tiles = B * p.tpu; p.act = act;
  const int nkb = D / 64;
  p.gw = nkb % 4 == 0 ? 4 : (nkb % 2 == 0 ? 2 : 1);
  const size_t gbytes = align_up((size_t)(128 + 30) * D * 2, 1024), stage = (size_t)nkb * kblock_bytes(128);
  p.off_a = (uint32_t)(gbytes > stage ? gbytes : stage);
  p.off_ring = p.off_a + (uint32_t)stage;
  p.off_par = p.off_ring + CV_STAGES * CV_SLOT;
  p.off_stat = p.off_par + 3072;
  const size_t smem = (size_t)p.off_stat + 2048;
  const unsigned grid = (unsigned)(p.n_tiles < convf_sms() ? p.n_tiles : convf_sms());
  switch (act) {
    case SMX_ACT_SWISH: return launch_conv<SMX_ACT_SWISH>(p, grid, smem, st);
    case SMX_ACT_GELU: return launch_conv<SMX_ACT_GELU>(p, grid, smem, st);
    case SMX_ACT_RELU: return launch_conv<SMX_ACT_RELU>(p, grid, smem, st);
    default: return launch_conv<-1>(p, grid, smem, st);
  }
}

}  // namespace smx
