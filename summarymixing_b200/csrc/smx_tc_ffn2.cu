// tcgen05 arm of libsmx, part 6: K-FFN v2, the persistent fused macaron feed-forward half-step
//
//   y = x + 0.5 * ( W2 @ act( W1 @ LN(x) + b1 ) + b2 )        [optionally y = LN_out(y)]
//   (Conformer.py:470-484, :518, :547)
//
// One CTA per SM walks 128-row tiles.  The hidden dimension is processed in 128-wide chunks: GEMM1 (K = D,
// N = 128) fills one of two TMEM accumulators; the epilogue warps add b1, activate and write the bf16 chunk to
// shared memory as the A operand of GEMM2 (K = 128, N = D), which accumulates the output tile in TMEM — the
// d_ffn-wide hidden activation never leaves the SM.  Warp roles as in the fused cell (smx_tc_cell.cu):
//   warps 0-7   epilogue (two column groups x four TMEM lane quadrants)
//   warps 8-11  prologue: LayerNorm of the next tile into the A operand (as soon as the last GEMM1 released it)
//   warp 12     weight producer: 8 KB blocks of W1 / W2 through a shared-memory ring (cp.async.bulk + mbarrier)
//   warp 13     MMA issuer
// The final epilogue parks the residual tile (coalesced loads, issued while the last GEMMs run) in the idle hidden
// buffers, adds it in fp32, optionally applies the output LayerNorm (norm2) and leaves with coalesced stores.
#include "smx_tc.h"
#include "smx_tc_common.cuh"

namespace smx {

using tc::kblock_bytes;

constexpr int F2_THREADS = 448;       // 14 warps (measured: 4 prologue warps beat 2 even with a 128-register cap)
constexpr int F2_PRO_WARP0 = 8, F2_NPW = 4, F2_PROD_WARP = 12, F2_MMA_WARP = 13;
constexpr int F2_RPW = 128 / F2_NPW;  // rows per prologue warp
constexpr int F2_HC = 128;            // hidden chunk width
constexpr int F2_STAGES = 8;          // ring slots of 8 KB
constexpr uint32_t F2_BLOCK = 8192;

struct Ffn2P {
  const __nv_bfloat16* x; __nv_bfloat16* y; int64_t rows;
  int D, F, n_tiles;
  const uint8_t* w1; const uint8_t* w2;   // packed images, 64 x 64 blocks: w1 [F/64][D/64], w2 [D/64][F/64]
  const float* ln_w; const float* ln_b; const float* b1; const float* b2;
  const float* oln_w; const float* oln_b; float oln_eps;
  int act;
  int cl;                                  // thread-block cluster size (1, 2 or 4): weight blocks are multicast
  int gw2;                                 // GEMM2 column group width in 64-col chunks (1, 2 or 4; divides D/64)
  unsigned long long* trace;
  uint32_t off_h, off_ring, off_par, off_red;
};

#define F2_TRACE(role, it, ev)                                                                                \
  do {                                                                                                        \
    if (p.trace && blockIdx.x == 0 && lane == 0 && (it) < 2) p.trace[(((role)*2 + (it)) * 32) + (ev)] = clock64(); \
  } while (0)

__device__ __forceinline__ uint4 f2_pack8(const float* v) {
  return make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void f2_unpack8(const uint4& raw, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
}

template <bool OLN, int ACT>  // ACT >= 0: compile-time activation (smx_act); -1: runtime p.act
__global__ void __launch_bounds__(F2_THREADS, 1) ffn2_kernel(const Ffn2P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sX = smem;
  uint8_t* sH = smem + p.off_h;       // two hidden buffers of 2 K-blocks (32 KB each); also the output staging tile
  uint8_t* sRing = smem + p.off_ring;
  float* sPar = reinterpret_cast<float*>(smem + p.off_par);  // [b1 (F) | b2 | oln_w | oln_b | ln_w | ln_b (256 each)]
  float* sRed = reinterpret_cast<float*>(smem + p.off_red);  // [2 stats][2 groups][128 rows]
  __shared__ __align__(8) uint64_t full_bar[F2_STAGES], empty_bar[F2_STAGES];
  __shared__ __align__(8) uint64_t x_full, x_free, acc1_full[2], acc1_empty[2], h_full[2], h_empty[2], acc2_full, epi_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int D = p.D, nkbD = D / 64, nj = p.F / F2_HC, nkbF = p.F / 64;
  const int act = ACT >= 0 ? ACT : p.act;

  // Cluster of cl CTAs: every CTA fetches 1/cl of each weight block and multicasts it to all of them, so the L2
  // (whose bandwidth is what bounds this kernel) serves each block once per cluster instead of once per CTA.
  const int cl = p.cl;
  const uint32_t crank = cl > 1 ? tc::cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << cl) - 1u);
  if (warp == F2_PROD_WARP) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int s = 0; s < F2_STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], (uint32_t)cl); }
    tc::mbar_init(&x_full, F2_NPW); tc::mbar_init(&x_free, 1); tc::mbar_init(&acc2_full, 1); tc::mbar_init(&epi_done, 8);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&acc1_full[i], 1); tc::mbar_init(&acc1_empty[i], 8);
      tc::mbar_init(&h_full[i], 8); tc::mbar_init(&h_empty[i], 1);
    }
    tc::fence_barrier_init();
  }
  for (int i = tid; i < p.F; i += F2_THREADS) sPar[i] = p.b1[i];
  float* sB2 = sPar + p.F; float* sOw = sB2 + 256; float* sOb = sOw + 256; float* sLw = sOb + 256; float* sLb = sLw + 256;
  for (int i = tid; i < 256; i += F2_THREADS) {
    sB2[i] = i < D ? p.b2[i] : 0.0f;
    sOw[i] = (OLN && i < D) ? p.oln_w[i] : 1.0f;
    sOb[i] = (OLN && i < D) ? p.oln_b[i] : 0.0f;
    sLw[i] = i < D ? p.ln_w[i] : 1.0f;
    sLb[i] = i < D ? p.ln_b[i] : 0.0f;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (cl > 1) tc::cluster_sync();  // peers' barriers are initialised before any multicast / remote arrive
  tc::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  const uint32_t t_acc2 = tmem, t_acc1 = tmem + 256;  // acc1 buffers at +256 and +384
  // cluster c takes tile groups c, c + n_clusters, ...; CTA rank r the r-th tile of the group.  All CTAs of a cluster
  // run the same number of iterations (the ring protocol is collective); tiles past the end are empty (nrows <= 0).
  const int first_base = (cl > 1 ? (int)tc::cluster_id_x() : (int)blockIdx.x) * cl;
  const int base_step = (cl > 1 ? (int)tc::cluster_count_x() : (int)gridDim.x) * cl;
  const int gw2 = p.gw2, ng2 = nkbD / gw2;  // GEMM2 column groups
  // CTAs walk the hidden chunks in rotated order so that at any moment different SMs ask the L2 for different weight
  // blocks (all SMs fetching the same lines at once serialises on the L2 slices that hold them)
  const int rot = (int)(blockIdx.x % (unsigned)nj);
  auto chunk_of = [&](int j) { int c = j + rot; return c >= nj ? c - nj : c; };

  if (warp == F2_PROD_WARP) {
    // =============================== weight producer ===============================
    // ring order == issue order: W1[0], W1[1], W2[0], W1[2], W2[1], ..., W2[nj-1]; a step takes gw consecutive,
    // gw-aligned slots: data arrival on the first slot's full barrier, consumption on every slot's empty barrier
    if (lane == 0) {
      int s = 0;
      uint32_t pe = 0;
      auto load = [&](const uint8_t* img, int nkb_img, int c0, int gw, int kb) {
        s = (s + gw - 1) & ~(gw - 1);
        if (s >= F2_STAGES) s = 0;
        for (int u = 0; u < gw; ++u) {
          tc::mbar_wait(&empty_bar[s + u], ((pe >> (s + u)) & 1u) ^ 1u);
          pe ^= 1u << (s + u);
        }
        tc::mbar_arrive_expect_tx(&full_bar[s], F2_BLOCK * gw);  // all slices of the step land here, whoever fetched them
        for (int u = 0; u < gw; ++u) {
          uint8_t* dst = sRing + (size_t)(s + u) * F2_BLOCK;
          const uint8_t* src = img + (size_t)((c0 + u) * nkb_img + kb) * F2_BLOCK;
          if (cl == 1) {
            tc::bulk_g2s(dst, src, F2_BLOCK, &full_bar[s]);
          } else {
            const uint32_t slice = F2_BLOCK / (uint32_t)cl;
            tc::bulk_g2s_multicast(dst + crank * slice, src + crank * slice, slice, &full_bar[s], cmask);
          }
        }
        s += gw;
      };
      auto load_g1 = [&](int j) { const int c = chunk_of(j); for (int kb = 0; kb < nkbD; ++kb) load(p.w1, nkbD, 2 * c, 2, kb); };
      auto load_g2 = [&](int j) {
        const int c = chunk_of(j);
        for (int g = 0; g < ng2; ++g)
          for (int u = 0; u < 2; ++u) load(p.w2, nkbF, g * gw2, gw2, 2 * c + u);
      };
      for (int base = first_base; base < p.n_tiles; base += base_step) {
        load_g1(0);
        for (int j = 0; j < nj; ++j) {
          if (j + 1 < nj) load_g1(j + 1);
          load_g2(j);
        }
      }
    }
  } else if (warp == F2_MMA_WARP) {
    // =============================== MMA issuer ===============================
    int s = 0;
    uint32_t pf = 0;
    uint32_t ph_a1e = 0, ph_hf = 0;  // per-buffer parity bits of acc1_empty / h_full
    const uint32_t x0 = tc::smem_u32(sX), h0 = tc::smem_u32(sH), r0 = tc::smem_u32(sRing);
    const uint32_t idesc1 = tc::make_idesc_bf16(128, F2_HC), idesc2 = tc::make_idesc_bf16(128, 64u * gw2);
    int it = 0;
    int tr = -1;  // fine-grained trace cursor (chunk 2 of the first tile)
    auto stamp = [&]() { if (tr >= 0 && tr < 40 && p.trace && blockIdx.x == 0 && lane == 0) p.trace[320 + tr] = clock64(); if (tr >= 0) ++tr; };
    auto step = [&](int gw, uint32_t a_addr, uint32_t d_addr, uint32_t idesc, bool first) {
      s = (s + gw - 1) & ~(gw - 1);
      if (s >= F2_STAGES) s = 0;
      stamp();
      tc::mbar_wait(&full_bar[s], (pf >> s) & 1u);
      stamp();
      pf ^= 1u << s;
      tc::tc_fence_after();
      const uint32_t b_addr = r0 + s * F2_BLOCK;
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc::umma_bf16(d_addr, tc::make_desc_sw128(a_addr + ks * 32), tc::make_desc_sw128(b_addr + ks * 32), idesc,
                        (first && ks == 0) ? 0u : 1u);
        for (int u = 0; u < gw; ++u) {
          if (cl == 1) tc::umma_commit(&empty_bar[s + u]);
          else tc::umma_commit_multicast(&empty_bar[s + u], cmask);  // every CTA's producer waits for all consumers
        }
      }
      __syncwarp();
      s += gw;
    };
    auto gemm1 = [&](int j) {  // acc1[j&1] = LN(x) @ W1[chunk j]^T
      const int bsel = j & 1;
      stamp();
      tc::mbar_wait(&acc1_empty[bsel], ((ph_a1e >> bsel) & 1u) ^ 1u);
      stamp();
      ph_a1e ^= 1u << bsel;
      tc::tc_fence_after();
      for (int kb = 0; kb < nkbD; ++kb) step(2, x0 + kb * kblock_bytes(128), t_acc1 + bsel * F2_HC, idesc1, kb == 0);
      if (tc::elect_one()) {
        tc::umma_commit(&acc1_full[bsel]);
        if (j == nj - 1) tc::umma_commit(&x_free);
      }
      __syncwarp();
    };
    for (int base = first_base; base < p.n_tiles; base += base_step, ++it) {
      const uint32_t par = it & 1;
      tc::mbar_wait(&x_full, par);
      tc::tc_fence_after();
      F2_TRACE(1, it, 0);
      gemm1(0);
      for (int j = 0; j < nj; ++j) {
        tr = (it == 0 && j == 2) ? 0 : -1;
        if (j + 1 < nj) gemm1(j + 1);
        const int bsel = j & 1;
        if (j == 0 && it > 0) tc::mbar_wait(&epi_done, par ^ 1);  // previous tile's output accumulator is drained
        stamp();
        tc::mbar_wait(&h_full[bsel], (ph_hf >> bsel) & 1u);
        stamp();
        ph_hf ^= 1u << bsel;
        tc::tc_fence_after();
        for (int g = 0; g < ng2; ++g)
          for (int u = 0; u < 2; ++u)  // acc2[:, group g] += H[j][:, K-block u] @ W2[group g, K-block 2j+u]^T
            step(gw2, h0 + (bsel * 2 + u) * kblock_bytes(128), t_acc2 + g * gw2 * 64, idesc2, j == 0 && u == 0);
        if (tc::elect_one()) {
          tc::umma_commit(&h_empty[bsel]);
          if (j == nj - 1) tc::umma_commit(&acc2_full);
        }
        __syncwarp();
        if (j < 4) F2_TRACE(1, it, 1 + j);
      }
      F2_TRACE(1, it, 8);
    }
  } else if (warp >= F2_PRO_WARP0) {
    // =============================== prologue: x tile -> LN -> A operand ===============================
    const int pw = warp - F2_PRO_WARP0;
    int it = 0;
    for (int base = first_base; base < p.n_tiles; base += base_step, ++it) {
      const int64_t row0 = (int64_t)(base + (int)crank) * 128;
      const int nrows = p.rows - row0 < 128 ? (int)(p.rows - row0) : 128;
      if (it > 0) tc::mbar_wait(&x_free, (it - 1) & 1);
      if (pw == 0) F2_TRACE(2, it, 0);
      // all rows of this warp in flight at once (cp.async straight into the operand image), then LayerNorm in place
      tc::stage_ln_rows(sX, p.x, D, row0, nrows, D, pw, lane, true, sLw, sLb,
                        (p.trace && blockIdx.x == 0 && pw == 0 && it < 2) ? p.trace + ((2 * 2 + it) * 32) + 4 : nullptr);
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&x_full);
      if (pw == 0) F2_TRACE(2, it, 1);
    }
  } else {
    // =============================== epilogue ===============================
    const int grp = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const int etid = tid;  // 0..255
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    uint32_t ph_a1f = 0, ph_he = 0;
    const int cpr = D / 8;
    const int rr0 = etid / cpr, ch0 = etid - rr0 * cpr, drr = 256 / cpr, dch = 256 - drr * cpr;
    int it = 0;
    for (int base = first_base; base < p.n_tiles; base += base_step, ++it) {
      const uint32_t par = it & 1;
      const int64_t row0 = (int64_t)(base + (int)crank) * 128;
      const int nrows = p.rows - row0 < 128 ? (int)(p.rows - row0) : 128;
      for (int j = 0; j < nj; ++j) {
        const int bsel = j & 1;
        tc::mbar_wait(&acc1_full[bsel], (ph_a1f >> bsel) & 1u);
        ph_a1f ^= 1u << bsel;
        tc::tc_fence_after();
        if (q == 0 && j < 4) F2_TRACE(3 + grp, it, 2 * j);
        float v[2][32];
        tc::tmem_ld32(t_acc1 + lane_sel + bsel * F2_HC + grp * 64, v[0]);
        tc::tmem_ld32(t_acc1 + lane_sel + bsel * F2_HC + grp * 64 + 32, v[1]);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc1_empty[bsel]);
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {
          const float4* bp = reinterpret_cast<const float4*>(sPar + chunk_of(j) * F2_HC + grp * 64 + pc * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float4 bb = bp[i]; v[pc][4 * i] += bb.x; v[pc][4 * i + 1] += bb.y; v[pc][4 * i + 2] += bb.z; v[pc][4 * i + 3] += bb.w; }
          tc::act_apply<32>(act, v[pc]);
        }
        tc::mbar_wait(&h_empty[bsel], ((ph_he >> bsel) & 1u) ^ 1u);  // GEMM2 of chunk j-2 has finished reading this buffer
        ph_he ^= 1u << bsel;
        uint8_t* hk = sH + (size_t)(bsel * 2 + grp) * kblock_bytes(128);
#pragma unroll
        for (int pc = 0; pc < 2; ++pc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(hk + tc::sw128_offset(r, pc * 4 + k)) = f2_pack8(v[pc] + 8 * k);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&h_full[bsel]);
        if (q == 0 && j < 4) F2_TRACE(3 + grp, it, 2 * j + 1);
      }
      // ---- final: y = x + 0.5*(acc2 + b2)  [-> LN_out]; residual parked in the idle hidden buffers
      uint4 rres[16];
      {
        int rr = rr0, ch = ch0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          rres[k] = make_uint4(0, 0, 0, 0);
          if (rr < nrows) rres[k] = *reinterpret_cast<const uint4*>(p.x + (row0 + rr) * D + ch * 8);
          rr += drr; ch += dch;
          if (ch >= cpr) { ch -= cpr; ++rr; }
        }
      }
      tc::mbar_wait(&acc2_full, par);  // every MMA of the tile has completed: the hidden buffers are free
      tc::tc_fence_after();
      if (q == 0) F2_TRACE(3 + grp, it, 10);
      {
        int rr = rr0, ch = ch0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (rr < 128) *reinterpret_cast<uint4*>(sH + (size_t)(ch >> 3) * kblock_bytes(128) + tc::sw128_offset(rr, ch & 7)) = rres[k];
          rr += drr; ch += dch;
          if (ch >= cpr) { ch -= cpr; ++rr; }
        }
      }
      tc::named_bar_sync(1, 256);
      // columns of this thread's row: groups split the 64-column chunks (chunk c belongs to group c & 1)
      float s1 = 0.0f;
      for (int c = grp; c < nkbD; c += 2) {
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {
          const int col = c * 64 + pc * 32;
          float v[32];
          tc::tmem_ld32(t_acc2 + lane_sel + col, v);
          tc::tmem_ld_wait();
          const float4* bp = reinterpret_cast<const float4*>(sB2 + col);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint4* sp = reinterpret_cast<uint4*>(sH + (size_t)c * kblock_bytes(128) + tc::sw128_offset(r, pc * 4 + k));
            float f[8];
            f2_unpack8(*sp, f);
            const float4 ba = bp[2 * k], bb = bp[2 * k + 1];
            const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) { v[8 * k + e] = fmaf(0.5f, v[8 * k + e] + bv[e], f[e]); s1 += v[8 * k + e]; }
            if (!OLN) *sp = f2_pack8(v + 8 * k);
          }
          if (OLN) tc::tmem_st32(t_acc2 + lane_sel + col, v);  // park the pre-norm row in TMEM for the LayerNorm passes
        }
      }
      if (OLN) {
        tc::tmem_st_wait();
        sRed[grp * 128 + r] = s1;
        tc::named_bar_sync(1, 256);
        const float mean = (sRed[r] + sRed[128 + r]) / (float)D;
        float s2 = 0.0f;
        for (int c = grp; c < nkbD; c += 2) {
#pragma unroll
          for (int pc = 0; pc < 2; ++pc) {
            float v[32];
            tc::tmem_ld32(t_acc2 + lane_sel + c * 64 + pc * 32, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) { const float d = v[e] - mean; s2 = fmaf(d, d, s2); }
          }
        }
        sRed[256 + grp * 128 + r] = s2;
        tc::named_bar_sync(1, 256);
        const float rstd = rsqrtf((sRed[256 + r] + sRed[384 + r]) / (float)D + p.oln_eps);
        for (int c = grp; c < nkbD; c += 2) {
#pragma unroll
          for (int pc = 0; pc < 2; ++pc) {
            const int col = c * 64 + pc * 32;
            float v[32];
            tc::tmem_ld32(t_acc2 + lane_sel + col, v);
            tc::tmem_ld_wait();
            const float4* wp = reinterpret_cast<const float4*>(sOw + col);
            const float4* bp = reinterpret_cast<const float4*>(sOb + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 ww = wp[i], bb = bp[i];
              v[4 * i] = (v[4 * i] - mean) * rstd * ww.x + bb.x;
              v[4 * i + 1] = (v[4 * i + 1] - mean) * rstd * ww.y + bb.y;
              v[4 * i + 2] = (v[4 * i + 2] - mean) * rstd * ww.z + bb.z;
              v[4 * i + 3] = (v[4 * i + 3] - mean) * rstd * ww.w + bb.w;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<uint4*>(sH + (size_t)c * kblock_bytes(128) + tc::sw128_offset(r, pc * 4 + k)) = f2_pack8(v + 8 * k);
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&epi_done);
      if (q == 0) F2_TRACE(3 + grp, it, 11);
      tc::named_bar_sync(1, 256);
      {
        int rr = rr0, ch = ch0;
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
          if (rr < nrows) {
            const uint4 val = *reinterpret_cast<const uint4*>(sH + (size_t)(ch >> 3) * kblock_bytes(128) + tc::sw128_offset(rr, ch & 7));
            *reinterpret_cast<uint4*>(p.y + (row0 + rr) * D + ch * 8) = val;
          }
          rr += drr; ch += dch;
          if (ch >= cpr) { ch -= cpr; ++rr; }
        }
      }
      tc::named_bar_sync(1, 256);
      if (q == 0) F2_TRACE(3 + grp, it, 12);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (cl > 1) tc::cluster_sync();  // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == F2_PROD_WARP) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool tc_ffn2_supported(const smx_ffn_weights* w) {
  const int D = w->w1.in_dim, F = w->w1.out_dim;
  if (w->w1.n_split > 1 || w->w2.n_split > 1 || w->w2.in_dim != F || w->w2.out_dim != D) return false;
  if (D % 64 || D < 64 || D > 256 || F % F2_HC || F < F2_HC || F > 2048) return false;
  if (!w->w1.w || !w->w1.b || !w->w2.w || !w->w2.b || !w->ln_w || !w->ln_b) return false;
  return true;
}
size_t tc_ffn2_packed_bytes(const smx_ffn_weights* w) { return 2 * align_up((size_t)w->w1.in_dim * w->w1.out_dim * 2, 1024); }

int tc_ffn2_pack(const smx_ffn_weights* w, void* packed, cudaStream_t st) {
  const int D = w->w1.in_dim, F = w->w1.out_dim;
  SMX_TRY(tc_pack_linear_nt(w->w1, 0, D, 64, packed, st));
  return tc_pack_linear_nt(w->w2, 0, F, 64, (char*)packed + align_up((size_t)D * F * 2, 1024), st);
}

static std::atomic<int> g_ffn_cluster{1};  // measured on B200: multicast clusters couple the CTAs' rings and run slower (85/101/120 us for 1/2/4)
void tc_set_ffn_cluster(int cl) { g_ffn_cluster = (cl == 4 || cl == 2) ? cl : 1; }
static std::atomic<unsigned long long*> g_trace2{nullptr};
void tc_set_trace_ffn(void* p) { g_trace2 = (unsigned long long*)p; }

static int ffn2_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <bool OLN>
static int launch_ffn2(const Ffn2P& p, unsigned grid, size_t smem, cudaStream_t st) {
  cudaError_t e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(F2_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
#define SMX_FFN2_LAUNCH(A)                                                                                   \
  e = cudaFuncSetAttribute(ffn2_kernel<OLN, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(ffn2_kernel): %s", cudaGetErrorString(e)); \
  e = cudaLaunchKernelEx(&cfg, ffn2_kernel<OLN, A>, p);                                                      \
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaLaunchKernelEx(ffn2_kernel): %s", cudaGetErrorString(e));
  switch (p.act) {
    case SMX_ACT_SWISH: SMX_FFN2_LAUNCH(SMX_ACT_SWISH); break;
    case SMX_ACT_GELU: SMX_FFN2_LAUNCH(SMX_ACT_GELU); break;
    case SMX_ACT_RELU: SMX_FFN2_LAUNCH(SMX_ACT_RELU); break;
    default: SMX_FFN2_LAUNCH(-1); break;
  }
#undef SMX_FFN2_LAUNCH
  count_tc_launch();
  return check_launch("ffn2_kernel");
}

int tc_ffn2_fwd(const smx_ffn_weights* w, const void* packed, int act, int64_t rows, const __nv_bfloat16* x,
                const float* oln_w, const float* oln_b, float oln_eps, __nv_bfloat16* y, cudaStream_t st) {
  const int D = w->w1.in_dim, F = w->w1.out_dim;
  Ffn2P p{};
  p.x = x; p.y = y; p.rows = rows; p.D = D; p.F = F;
  p.n_tiles = (int)((rows + 127) / 128);
  p.w1 = (const uint8_t*)packed;
  p.w2 = p.w1 + align_up((size_t)D * F * 2, 1024);
  p.ln_w = w->ln_w; p.ln_b = w->ln_b; p.b1 = w->w1.b; p.b2 = w->w2.b;
  p.oln_w = oln_w; p.oln_b = oln_b; p.oln_eps = oln_eps;
  p.act = act;
  const int nc = D / 64;
  p.gw2 = nc % 4 == 0 ? 4 : (nc % 2 == 0 ? 2 : 1);
  p.trace = g_trace2;
  const uint32_t xb = (uint32_t)nc * kblock_bytes(128);
  p.off_h = xb;
  p.off_ring = xb + 4 * kblock_bytes(128);
  p.off_par = p.off_ring + F2_STAGES * F2_BLOCK;
  p.off_red = p.off_par + (uint32_t)align_up((size_t)(F + 1280) * 4, 1024);
  const size_t smem = (size_t)p.off_red + 2048;
  if (smem > 227 * 1024 - 1024) return fail(SMX_ERR_UNSUPPORTED, "ffn: tile does not fit shared memory");
  // cluster size: pairs (or quads) of CTAs share every weight block; single CTAs when there is too little work
  p.cl = g_ffn_cluster;
  while (p.cl > 1 && p.n_tiles < 2 * p.cl) p.cl >>= 1;
  unsigned grid = (unsigned)(p.n_tiles < ffn2_sms() ? p.n_tiles : ffn2_sms());
  grid = (grid + p.cl - 1) / p.cl * p.cl;
  if ((int)grid > ffn2_sms()) grid = (unsigned)(ffn2_sms() / p.cl * p.cl);
  return oln_w ? launch_ffn2<true>(p, grid, smem, st) : launch_ffn2<false>(p, grid, smem, st);
}

}  // namespace smx
