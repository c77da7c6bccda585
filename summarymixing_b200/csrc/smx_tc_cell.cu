// tcgen05 arm of libsmx, part 5: K-SM, the fused SummaryMixing cell (mode "SummaryMixing", whole-utterance
// mean)                                                                   summary_mixing.py:198-253
//
// Two persistent, warp-specialised passes over utterance-aligned 128-frame tiles (one CTA per SM):
//
//   pass A (summary):  X = LN1(x tile)  ->  S = act(act(X W_s1 + b) W_s2 + b) * mask  ->  column sums of the tile
//   finalise (tiny):   per utterance: mean over valid frames (fixed-order sum of the tile partials: no atomics,
//                      deterministic), LN_s, c[b] = W_c[:, D_l:] mean + b_c
//   pass B (local):    X = LN1(x tile)  ->  L = LN_l(act(act(X W_f1 + b) W_f2 + b) * mask)
//                      ->  y = act(L W_c[:, :D_l]^T + c[b]) (+ residual)
//
// The frame tile is read from HBM once per pass (the second read hits L2) and only y is written: hidden
// activations, S, L and the concatenation never exist in global memory.  Inside a CTA:
//   warps 0-7   epilogue: two groups of four warps (one per TMEM lane quadrant) take alternate 64-column chunks:
//               tcgen05.ld -> bias/activation/mask -> next GEMM's A operand in shared memory (or LN_l through
//               TMEM, column sums, or the staged output tile which leaves with coalesced 16-byte stores)
//   warps 8-11  prologue: coalesced 16-byte loads of the next tile, LayerNorm (norm1), bf16 A operand (128B swizzle)
//   warp 12     weight producer: 8 KB (64 x 64 bf16) weight blocks stream from the packed image (L2) through a
//               shared-memory ring with cp.async.bulk + mbarrier; zero blocks of block-diagonal weights are skipped
//   warp 13     MMA issuer: tcgen05.mma (M=128, N=64, K=16) into 64-column TMEM accumulator chunks
// so the tensor pipe, the TMA engine, the prologue loads and the epilogue math of neighbouring chunks/tiles overlap.
#include "smx_tc.h"
#include "smx_tc_common.cuh"

namespace smx {

using tc::kblock_bytes;

constexpr int CF_THREADS = 448;      // 14 warps (measured: 4 prologue warps beat 2 even with a 128-register cap)
// Warp roles.  The SM's warp arbiter favours the highest warp id among eligible warps, so the two single-thread,
// latency-critical roles (MMA issuer, weight producer) sit in the top warps and the bulk math below them.
constexpr int CF_EPI_WARP0 = 0;      // warps 0..7   epilogue
constexpr int CF_PRO_WARP0 = 8;      // warps 8..11  prologue
constexpr int CF_NPW = 4;            // prologue warps, 128 / CF_NPW rows each
constexpr int CF_PROD_WARP = 12;     // weight producer (also owns the TMEM allocation)
constexpr int CF_MMA_WARP = 13;      // MMA issuer
constexpr int CF_MAX_STAGES = 12;     // ring slots (a multiple of 4 is used)
constexpr uint32_t CF_BLOCK_BYTES = 8192;  // one 64 x 64 bf16 weight block

struct CfGemm {
  const uint8_t* img;  // packed image: [chunk][kblock] blocks of 64 rows x 128 B (128B swizzle)
  int nkb, nc;         // K/64, N/64
  int nheads, cph, kph;   // block-diagonal structure: heads, chunks per head, K-blocks per head (dense: 1, nc, nkb)
  int gw;                 // chunks per MMA column group (1, 2 or 4; divides cph)
};

struct CellFP {
  const __nv_bfloat16* x; int64_t ldx;
  const float* pre_w; const float* pre_b;   // norm1 (NULL: none)
  const uint8_t* mask;                       // (B,T) or NULL
  const __nv_bfloat16* resid; int64_t ldr;
  __nv_bfloat16* y; int64_t ldy;
  int B, T, tpu, n_tiles;
  int D;                                     // enc_dim
  CfGemm g[3];                               // A: s1, s2     B: f1, f2, combiner(local part)
  const float* b1; const float* b2;          // biases of the two MLP blocks of this pass (GLU pass: value / gate halves)
  int nb1, nb2;                              // their lengths (<= 256)
  const float* ln_w; const float* ln_b;      // A: summary_norm   B: local_norm      (NULL: no LayerNorm)
  const float* Wc; const float* bc; int Dl, Ds, Dout;   // merge weight (fp32, (Dout, Dl+Ds)) and bias
  int act;
  float* colsum;        // [n_tiles][Ds]
  float* rowbias;       // [B][Dout]
  unsigned long long* trace;  // debug timeline of CTA 0 (NULL: off)
  int n_stages;
  uint32_t off_y, off_ring, off_par, off_red;   // shared-memory carve-up (bytes from the 1024-aligned base)
};

// debug timeline: role (0 producer, 1 issuer, 2 prologue, 3/4 epilogue groups) x tile iteration (< 4) x event (< 16)
#define CF_TRACE(role, it, ev)                                                                          \
  do {                                                                                                  \
    if (p.trace && blockIdx.x == 0 && lane == 0 && (it) < 4) p.trace[(((role)*4 + (it)) * 16) + (ev)] = clock64(); \
  } while (0)

__device__ __forceinline__ uint4 cf_pack8(const float* v) {
  return make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]), tc::pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void cf_unpack8(const uint4& raw, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
}

// sum over the 32 lanes of v[j] for each j; lane l ends up holding column l's total
__device__ __forceinline__ float cf_column_sums(float* v, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      float send = up ? v[j] : v[j + s];
      float keep = up ? v[j + s] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// Per-utterance finalisation between the two passes: CTA (b, j) recomputes the utterance mean from the per-tile
// column sums (fixed order: deterministic), applies LN_s and produces 64 entries of
//   c[b] = W_c[:, D_l:] @ LN_s( sum_t s[b,t] / sum_t mask[b,t] ) + b_c           summary_mixing.py:229-231, 248-253
// Eight warps x eight outputs each; a warp streams its eight weight rows with all loads in flight at once.
struct CellFinP {
  const float* colsum; const uint8_t* mask; const float* ln_w; const float* ln_b; const float* Wc; const float* bc;
  float* rowbias;
  int T, tpu, Ds, Dl, Dout;
};
__global__ void __launch_bounds__(256) cell_finalize2_kernel(const CellFinP p) {
  __shared__ float mu[256];
  __shared__ float red[8];
  __shared__ float stat[2];
  const int b = blockIdx.x, n0 = blockIdx.y * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ds = p.Ds;
  tc::pdl_launch_dependents();
  tc::pdl_wait();  // the column sums come from pass A
  float cnt;
  if (p.mask) {  // number of valid frames (integer-valued float, like torch.sum(mask) in the reference)
    float c = 0.0f;
    for (int t = tid; t < p.T; t += 256) c += (float)p.mask[(size_t)b * p.T + t];
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) red[warp] = c;
    __syncthreads();
    cnt = 0.0f;
    for (int i = 0; i < 8; ++i) cnt += red[i];
    __syncthreads();
  } else {
    cnt = (float)p.T;
  }
  if (tid < Ds) {
    const float* cs = p.colsum + (size_t)b * p.tpu * Ds + tid;
    float s = 0.0f;
    int i = 0;
    for (; i + 8 <= p.tpu; i += 8) {  // eight loads in flight, summed in tile order
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = cs[(size_t)(i + u) * Ds];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; i < p.tpu; ++i) s += cs[(size_t)i * Ds];
    mu[tid] = s / cnt;
  }
  __syncthreads();
  if (p.ln_w) {
    float v = tid < Ds ? mu[tid] : 0.0f, s = v;
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) { float t = 0.0f; for (int i = 0; i < 8; ++i) t += red[i]; stat[0] = t / (float)Ds; }
    __syncthreads();
    const float mean = stat[0];
    float d = tid < Ds ? v - mean : 0.0f, q = d * d;
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    __syncthreads();
    if (lane == 0) red[warp] = q;
    __syncthreads();
    if (tid == 0) { float t = 0.0f; for (int i = 0; i < 8; ++i) t += red[i]; stat[1] = rsqrtf(t / (float)Ds + 1e-5f); }
    __syncthreads();
    if (tid < Ds) mu[tid] = d * stat[1] * p.ln_w[tid] + p.ln_b[tid];
    __syncthreads();
  }
  // Ds <= 256: lane covers k = 4*lane..+3 and 128 + 4*lane..+3
  const int ldw = p.Dl + Ds;
  const bool h0 = 4 * lane < Ds, h1 = 128 + 4 * lane < Ds;
  float4 w0[8], w1[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int n = n0 + warp * 8 + u;
    const float* wr = p.Wc + (size_t)n * ldw + p.Dl;
    w0[u] = (n < p.Dout && h0) ? *reinterpret_cast<const float4*>(wr + 4 * lane) : make_float4(0, 0, 0, 0);
    w1[u] = (n < p.Dout && h1) ? *reinterpret_cast<const float4*>(wr + 128 + 4 * lane) : make_float4(0, 0, 0, 0);
  }
  const float4 m0 = h0 ? *reinterpret_cast<const float4*>(mu + 4 * lane) : make_float4(0, 0, 0, 0);
  const float4 m1 = h1 ? *reinterpret_cast<const float4*>(mu + 128 + 4 * lane) : make_float4(0, 0, 0, 0);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    float acc = w0[u].x * m0.x + w0[u].y * m0.y + w0[u].z * m0.z + w0[u].w * m0.w;
    acc += w1[u].x * m1.x + w1[u].y * m1.y + w1[u].z * m1.z + w1[u].w * m1.w;
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const int n = n0 + warp * 8 + u;
    if (lane == 0 && n < p.Dout) p.rowbias[(size_t)b * p.Dout + n] = acc + p.bc[n];
  }
}

template <int PHASE, int ACT>  // PHASE 0: pass A (summary), 1: pass B (local + combiner), 2: GLU pass; ACT >= 0: compile-time smx_act
__global__ void __launch_bounds__(CF_THREADS, 1) cell_kernel(const CellFP p) {
  extern __shared__ __align__(1024) uint8_t smem[];  // no pointer arithmetic through integers: keeps LDS/STS addressing
  uint8_t* sX = smem;
  uint8_t* sY = smem + p.off_y;
  uint8_t* sRing = smem + p.off_ring;
  float* sPar = reinterpret_cast<float*>(smem + p.off_par);   // [b1 | b2 | ln_w | ln_b | c[b] | norm1 w | norm1 b], 256 floats each
  float* sRed = reinterpret_cast<float*>(smem + p.off_red);   // 1024 floats: column partials / LN statistics / finalize
  __shared__ __align__(8) uint64_t full_bar[CF_MAX_STAGES], empty_bar[CF_MAX_STAGES];
  __shared__ __align__(8) uint64_t x_full, x_free, a2_full, epi_done;
  __shared__ __align__(8) uint64_t acc1_full[8], a1_full[4], acc2_full[4], acc3_full[4];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler (role dispatch below)
  constexpr int NG = PHASE == 0 ? 2 : (PHASE == 1 ? 3 : 1);
  const int act = ACT >= 0 ? ACT : p.act;  // compile-time activation: one tight loop per epilogue instead of a 7-way switch

  tc::pdl_launch_dependents();
  if (warp == CF_PROD_WARP) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int s = 0; s < p.n_stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&x_full, CF_NPW); tc::mbar_init(&x_free, 1); tc::mbar_init(&a2_full, 8); tc::mbar_init(&epi_done, 8);
    for (int c = 4; c < 8; ++c) tc::mbar_init(&acc1_full[c], 1);
    for (int c = 0; c < 4; ++c) {
      tc::mbar_init(&acc1_full[c], 1); tc::mbar_init(&a1_full[c], 4);
      tc::mbar_init(&acc2_full[c], 1); tc::mbar_init(&acc3_full[c], 1);
    }
    tc::fence_barrier_init();
  }
  for (int i = tid; i < 256; i += CF_THREADS) {
    sPar[i] = i < p.nb1 ? p.b1[i] : 0.0f;
    sPar[256 + i] = i < p.nb2 ? p.b2[i] : 0.0f;
    sPar[512 + i] = (p.ln_w && i < p.nb2) ? p.ln_w[i] : 1.0f;
    sPar[768 + i] = (p.ln_b && i < p.nb2) ? p.ln_b[i] : 0.0f;
    sPar[1280 + i] = (p.pre_w && i < p.D) ? p.pre_w[i] : 1.0f;   // norm1 (prologue LayerNorm)
    sPar[1536 + i] = (p.pre_b && i < p.D) ? p.pre_b[i] : 0.0f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  tc::pdl_wait();  // x / column sums / c[b] come from the preceding kernels (everything above touched only parameters)
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  const int first_tile = blockIdx.x, tile_step = gridDim.x;

  if (warp == CF_PROD_WARP) {
    // =============================== weight producer ===============================
    // Schedule (identical in the MMA issuer): block-diagonal weights are walked head by head; head m owns chunks
    // [m*cph, (m+1)*cph) and K-blocks [m*kph, (m+1)*kph) (dense = one head).  Inside a head, gw consecutive chunks
    // form a column group that one MMA covers (N = 64*gw); its K-block step takes gw consecutive, gw-aligned ring
    // slots (one 8 KB block per chunk): the data arrival is tracked on the first slot's full barrier, consumption
    // on every slot's own empty barrier (slots are also used singly).  Ring parity is tracked per slot.
    if (lane == 0) {
      int s = 0;
      uint32_t pe = 0;  // per-slot parity of the next empty-barrier wait
      for (int tile = first_tile; tile < p.n_tiles; tile += tile_step) {
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) {
          const CfGemm g = p.g[gi];
          for (int m = 0; m < g.nheads; ++m)
            for (int j = 0; j < g.cph; j += g.gw)
              for (int kb = m * g.kph; kb < (m + 1) * g.kph; ++kb) {
                s = (s + g.gw - 1) & ~(g.gw - 1);
                if (s >= p.n_stages) s = 0;
                for (int u = 0; u < g.gw; ++u) {  // every slot of the group must have been consumed
                  tc::mbar_wait(&empty_bar[s + u], ((pe >> (s + u)) & 1u) ^ 1u);
                  pe ^= 1u << (s + u);
                }
                tc::mbar_arrive_expect_tx(&full_bar[s], CF_BLOCK_BYTES * g.gw);
                for (int u = 0; u < g.gw; ++u)
                  tc::bulk_g2s(sRing + (size_t)(s + u) * CF_BLOCK_BYTES,
                               g.img + (size_t)((m * g.cph + j + u) * g.nkb + kb) * CF_BLOCK_BYTES, CF_BLOCK_BYTES, &full_bar[s]);
                s += g.gw;
              }
        }
      }
    }
  } else if (warp == CF_MMA_WARP) {
    // =============================== MMA issuer ===============================
    int s = 0;
    uint32_t pf = 0;  // per-slot parity of the next full-barrier wait
    const uint32_t x0 = tc::smem_u32(sX), y0 = tc::smem_u32(sY), r0 = tc::smem_u32(sRing);
    int it = 0;
    for (int tile = first_tile; tile < p.n_tiles; tile += tile_step, ++it) {
      const uint32_t par = it & 1;
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        const CfGemm g = p.g[gi];
        const uint32_t a0 = gi == 0 ? x0 : y0;
        const uint32_t dcol = gi == 0 ? 0u : 256u;
        uint64_t* accbar = gi == 0 ? acc1_full : (gi == 1 ? acc2_full : acc3_full);
        const uint32_t idesc = tc::make_idesc_bf16(128, 64u * g.gw);
        if (gi == 0) {
          tc::mbar_wait(&x_full, par);
          if (PHASE == 2 && it > 0) tc::mbar_wait(&epi_done, par ^ 1);  // the single accumulator set is drained
        } else if (gi == 1) {
          if (it > 0) tc::mbar_wait(&epi_done, par ^ 1);  // previous tile's accumulators in [256,512) are drained
        } else {
          tc::mbar_wait(&a2_full, par);
        }
        tc::tc_fence_after();
        CF_TRACE(1, it, gi * 2);
        for (int m = 0; m < g.nheads; ++m) {
          if (gi == 1) {  // A1 K-blocks of this head (every K-block belongs to exactly one head)
            for (int kb = m * g.kph; kb < (m + 1) * g.kph; ++kb) tc::mbar_wait(&a1_full[kb], par);
            tc::tc_fence_after();
          }
          for (int j = 0; j < g.cph; j += g.gw) {
            const int c0 = m * g.cph + j;
            const uint32_t d_addr = tmem + dcol + c0 * 64;
            for (int kb = m * g.kph; kb < (m + 1) * g.kph; ++kb) {
              s = (s + g.gw - 1) & ~(g.gw - 1);
              if (s >= p.n_stages) s = 0;
              tc::mbar_wait(&full_bar[s], (pf >> s) & 1u);
              pf ^= 1u << s;
              tc::tc_fence_after();
              const uint32_t a_addr = a0 + kb * kblock_bytes(128);
              const uint32_t b_addr = r0 + s * CF_BLOCK_BYTES;
              if (tc::elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  tc::umma_bf16(d_addr, tc::make_desc_sw128(a_addr + ks * 32), tc::make_desc_sw128(b_addr + ks * 32), idesc,
                                (kb == m * g.kph && ks == 0) ? 0u : 1u);
                for (int u = 0; u < g.gw; ++u) tc::umma_commit(&empty_bar[s + u]);
                if (kb == (m + 1) * g.kph - 1)
                  for (int u = 0; u < g.gw; ++u) tc::umma_commit(&accbar[c0 + u]);
              }
              __syncwarp();
              s += g.gw;
            }
          }
        }
        if (gi == 0 && tc::elect_one()) tc::umma_commit(&x_free);
        __syncwarp();
        CF_TRACE(1, it, gi * 2 + 1);
      }
    }
  } else if (warp >= CF_PRO_WARP0) {
    // =============================== prologue: x tile -> LN1 -> A operand ===============================
    const int pw = warp - CF_PRO_WARP0;
    int it = 0;
    for (int tile = first_tile; tile < p.n_tiles; tile += tile_step, ++it) {
      const int b = tile / p.tpu, t0 = (tile % p.tpu) * 128;
      const int64_t row0 = (int64_t)b * p.T + t0;
      const int nrows = p.T - t0 < 128 ? p.T - t0 : 128;
      if (it > 0) tc::mbar_wait(&x_free, (it - 1) & 1);
      if (pw == 0) CF_TRACE(2, it, 0);
      // all rows of this warp in flight at once (cp.async straight into the operand image), then LayerNorm in place, a thread per row
      tc::stage_ln_rows(sX, p.x, p.ldx, row0, nrows, p.D, pw, lane, p.pre_w != nullptr, sPar + 1280, sPar + 1536,
                        (p.trace && blockIdx.x == 0 && pw == 0 && it < 4) ? p.trace + ((2 * 4 + it) * 16) + 4 : nullptr);
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&x_full);
      if (pw == 0) CF_TRACE(2, it, 1);
    }
  } else {
    // =============================== epilogue ===============================
    const int e = warp - CF_EPI_WARP0;
    const int grp = e >> 2;            // chunk parity this warp handles
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;       // row inside the tile
    const int etid = e * 32 + lane;    // 0..255
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const float* sB1 = sPar; const float* sB2 = sPar + 256; const float* sLw = sPar + 512; const float* sLb = sPar + 768;
    float* sRB = sPar + 1024;
    const int nc1 = p.g[0].nc, nc2 = p.g[1].nc;
    int it = 0;
    for (int tile = first_tile; tile < p.n_tiles; tile += tile_step, ++it) {
      const uint32_t par = it & 1;
      const int b = tile / p.tpu, t0 = (tile % p.tpu) * 128;
      const int64_t row0 = (int64_t)b * p.T + t0;
      const int nrows = p.T - t0 < 128 ? p.T - t0 : 128;
      const bool live = r < nrows;
      const float rscale = live ? (p.mask ? (float)p.mask[row0 + r] : 1.0f) : 0.0f;

      if (PHASE == 2) {
        // ---- GLU pass: out = (acc[value] + b) * sigmoid(acc[gate] + b) -> staged tile -> coalesced stores
        // (value / gate 64-column chunks are interleaved in the packed weight image: chunk 2c / 2c+1)     Conformer.py:322-324
        const int npair = p.g[0].nc / 2;
        for (int c = grp; c < npair; c += 2) {
          tc::mbar_wait(&acc1_full[2 * c + 1], par);
          tc::tc_fence_after();
#pragma unroll
          for (int pc = 0; pc < 2; ++pc) {
            float v[32], gt[32];
            tc::tmem_ld32(tmem + lane_sel + (2 * c) * 64 + pc * 32, v);
            tc::tmem_ld32(tmem + lane_sel + (2 * c + 1) * 64 + pc * 32, gt);
            tc::tmem_ld_wait();
            const float4* ba = reinterpret_cast<const float4*>(sB1 + c * 64 + pc * 32);
            const float4* bg = reinterpret_cast<const float4*>(sB2 + c * 64 + pc * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 a4 = ba[j], g4 = bg[j];
              v[4 * j] += a4.x; v[4 * j + 1] += a4.y; v[4 * j + 2] += a4.z; v[4 * j + 3] += a4.w;
              gt[4 * j] += g4.x; gt[4 * j + 1] += g4.y; gt[4 * j + 2] += g4.z; gt[4 * j + 3] += g4.w;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= tc::act_sigmoid(gt[j]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<uint4*>(sY + (size_t)c * kblock_bytes(128) + tc::sw128_offset(r, pc * 4 + k)) = cf_pack8(v + 8 * k);
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&epi_done);
        tc::named_bar_sync(1, 256);
        {
          const int cpr = p.Dout / 8;
          const int rr0 = etid / cpr, ch0 = etid - rr0 * cpr, drr = 256 / cpr, dch = 256 - drr * cpr;
          int rr = rr0, ch = ch0;
#pragma unroll 4
          for (int k = 0; k < 16; ++k) {
            if (rr < nrows) {
              const uint4 val = *reinterpret_cast<const uint4*>(sY + (size_t)(ch >> 3) * kblock_bytes(128) + tc::sw128_offset(rr, ch & 7));
              *reinterpret_cast<uint4*>(p.y + (row0 + rr) * p.ldy + ch * 8) = val;
            }
            rr += drr; ch += dch;
            if (ch >= cpr) { ch -= cpr; ++rr; }
          }
        }
        tc::named_bar_sync(1, 256);
        continue;
      }
      // ---- E1: hidden = act(acc1 + b1) -> A operand of the second GEMM (K-block c of Y)
      if (q == 0) CF_TRACE(3 + grp, it, 0);
      for (int c = grp; c < nc1; c += 2) {
        tc::mbar_wait(&acc1_full[c], par);
        tc::tc_fence_after();
        if (q == 0 && c == grp) CF_TRACE(3 + grp, it, 1);
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {
          float v[32];
          tc::tmem_ld32(tmem + lane_sel + c * 64 + pc * 32, v);
          tc::tmem_ld_wait();
          const float4* bp = reinterpret_cast<const float4*>(sB1 + c * 64 + pc * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float4 bb = bp[j]; v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
          tc::act_apply<32>(act, v);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            *reinterpret_cast<uint4*>(sY + (size_t)c * kblock_bytes(128) + tc::sw128_offset(r, pc * 4 + k)) = cf_pack8(v + 8 * k);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&a1_full[c]);
      }

      if (q == 0) CF_TRACE(3 + grp, it, 2);
      if (PHASE == 0) {
        // ---- E2': S = act(acc2 + b2) * mask -> column sums of this tile                      :221, 229-231
        for (int c = grp; c < nc2; c += 2) {
          tc::mbar_wait(&acc2_full[c], par);
          tc::tc_fence_after();
#pragma unroll
          for (int pc = 0; pc < 2; ++pc) {
            float v[32];
            tc::tmem_ld32(tmem + lane_sel + 256 + c * 64 + pc * 32, v);
            tc::tmem_ld_wait();
            const float4* bp = reinterpret_cast<const float4*>(sB2 + c * 64 + pc * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 bb = bp[j]; v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
            tc::act_apply<32>(act, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= rscale;
            const float tot = cf_column_sums(v, lane);
            sRed[q * 256 + c * 64 + pc * 32 + lane] = tot;
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&epi_done);
        if (q == 0) CF_TRACE(3 + grp, it, 3);
        tc::named_bar_sync(1, 256);
        if (etid < p.Ds)  // fixed-order reduction over the four row quadrants: deterministic
          p.colsum[(size_t)tile * p.Ds + etid] = (sRed[etid] + sRed[256 + etid]) + (sRed[512 + etid] + sRed[768 + etid]);
        tc::named_bar_sync(1, 256);  // sRed is rewritten by the next tile
        if (q == 0) CF_TRACE(3 + grp, it, 5);
      } else {
        // ---- E2: L = LN_l(act(acc2 + b2) * mask) -> A operand of the combiner (Y, in place)     :215-218
        const int Dl = nc2 * 64;
        if (p.ln_w) {
          float s1 = 0.0f;
          for (int c = grp; c < nc2; c += 2) {
            tc::mbar_wait(&acc2_full[c], par);
            tc::tc_fence_after();
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
              float v[32];
              const uint32_t ta = tmem + lane_sel + 256 + c * 64 + pc * 32;
              tc::tmem_ld32(ta, v);
              tc::tmem_ld_wait();
              const float4* bp = reinterpret_cast<const float4*>(sB2 + c * 64 + pc * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) { const float4 bb = bp[j]; v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
              tc::act_apply<32>(act, v);
#pragma unroll
              for (int j = 0; j < 32; ++j) { v[j] *= rscale; s1 += v[j]; }
              tc::tmem_st32(ta, v);  // park the fp32 values in TMEM for the two LayerNorm passes
            }
          }
          tc::tmem_st_wait();
          if (q == 0) CF_TRACE(3 + grp, it, 3);
          sRed[grp * 128 + r] = s1;
          tc::named_bar_sync(1, 256);
          const float mean = (sRed[r] + sRed[128 + r]) / (float)Dl;
          float s2 = 0.0f;
          for (int c = grp; c < nc2; c += 2) {
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
              float v[32];
              tc::tmem_ld32(tmem + lane_sel + 256 + c * 64 + pc * 32, v);
              tc::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) { const float d = v[j] - mean; s2 = fmaf(d, d, s2); }
            }
          }
          sRed[256 + grp * 128 + r] = s2;
          tc::named_bar_sync(1, 256);
          const float rstd = rsqrtf((sRed[256 + r] + sRed[384 + r]) / (float)Dl + 1e-5f);
          for (int c = grp; c < nc2; c += 2) {
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
              float v[32];
              tc::tmem_ld32(tmem + lane_sel + 256 + c * 64 + pc * 32, v);
              tc::tmem_ld_wait();
              const float4* wp = reinterpret_cast<const float4*>(sLw + c * 64 + pc * 32);
              const float4* bp = reinterpret_cast<const float4*>(sLb + c * 64 + pc * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 ww = wp[j], bb = bp[j];
                v[4 * j] = (v[4 * j] - mean) * rstd * ww.x + bb.x;
                v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * ww.y + bb.y;
                v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * ww.z + bb.z;
                v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * ww.w + bb.w;
              }
#pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<uint4*>(sY + (size_t)c * kblock_bytes(128) + tc::sw128_offset(r, pc * 4 + k)) = cf_pack8(v + 8 * k);
            }
          }
        } else {
          for (int c = grp; c < nc2; c += 2) {
            tc::mbar_wait(&acc2_full[c], par);
            tc::tc_fence_after();
          }
          // without LayerNorm the values go straight to Y; all G2 chunks must have finished reading Y first
          tc::named_bar_sync(1, 256);
          for (int c = grp; c < nc2; c += 2) {
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
              float v[32];
              tc::tmem_ld32(tmem + lane_sel + 256 + c * 64 + pc * 32, v);
              tc::tmem_ld_wait();
              const float4* bp = reinterpret_cast<const float4*>(sB2 + c * 64 + pc * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) { const float4 bb = bp[j]; v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
              tc::act_apply<32>(act, v);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= rscale;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<uint4*>(sY + (size_t)c * kblock_bytes(128) + tc::sw128_offset(r, pc * 4 + k)) = cf_pack8(v + 8 * k);
            }
          }
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&a2_full);
        if (q == 0) CF_TRACE(3 + grp, it, 4);

        // ---- E3: y = act(acc3 + c[b]) (+ residual) -> staged tile -> coalesced stores            :251-253, :541
        // While the combiner GEMM runs, the residual tile is fetched with coalesced 16-byte loads (thread etid takes
        // chunk idx = etid + 256k of the tile) and c[b] goes to shared memory; once Y is free the residual is parked
        // there (swizzled, conflict-free) so that each thread can add its own row in fp32 and round once.
        const int nc3 = p.g[2].nc;
        const int cpr = p.Dout / 8;           // 16-byte chunks per output row
        const int rr0 = etid / cpr, ch0 = etid - rr0 * cpr, drr = 256 / cpr, dch = 256 - drr * cpr;
        uint4 rres[16];
        if (p.resid) {
          int rr = rr0, ch = ch0;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            rres[k] = make_uint4(0, 0, 0, 0);
            if (rr < nrows) rres[k] = *reinterpret_cast<const uint4*>(p.resid + (row0 + rr) * p.ldr + ch * 8);
            rr += drr; ch += dch;
            if (ch >= cpr) { ch -= cpr; ++rr; }
          }
        }
        if (etid < p.Dout) sRB[etid] = __ldcg(p.rowbias + (size_t)b * p.Dout + etid);
        tc::mbar_wait(&acc3_full[nc3 - 1], par);  // commits complete in order: every combiner chunk is done, Y is free
        tc::tc_fence_after();
        if (q == 0) CF_TRACE(3 + grp, it, 5);
        if (p.resid) {
          int rr = rr0, ch = ch0;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (rr < 128) *reinterpret_cast<uint4*>(sY + (size_t)(ch >> 3) * kblock_bytes(128) + tc::sw128_offset(rr, ch & 7)) = rres[k];
            rr += drr; ch += dch;
            if (ch >= cpr) { ch -= cpr; ++rr; }
          }
        }
        tc::named_bar_sync(1, 256);
        for (int c = grp; c < nc3; c += 2) {
#pragma unroll
          for (int pc = 0; pc < 2; ++pc) {
            const int col = c * 64 + pc * 32;
            float v[32];
            tc::tmem_ld32(tmem + lane_sel + 256 + col, v);
            tc::tmem_ld_wait();
            const float4* bp = reinterpret_cast<const float4*>(sRB + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 bb = bp[j]; v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
            tc::act_apply<32>(act, v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              uint4* sp = reinterpret_cast<uint4*>(sY + (size_t)c * kblock_bytes(128) + tc::sw128_offset(r, pc * 4 + k));
              if (p.resid) {
                float f[8];
                cf_unpack8(*sp, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[8 * k + j] += f[j];
              }
              *sp = cf_pack8(v + 8 * k);
            }
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&epi_done);
        if (q == 0) CF_TRACE(3 + grp, it, 6);
        tc::named_bar_sync(1, 256);
        {
          int rr = rr0, ch = ch0;
#pragma unroll 4
          for (int k = 0; k < 16; ++k) {
            if (rr < nrows) {
              const uint4 val = *reinterpret_cast<const uint4*>(sY + (size_t)(ch >> 3) * kblock_bytes(128) + tc::sw128_offset(rr, ch & 7));
              *reinterpret_cast<uint4*>(p.y + (row0 + rr) * p.ldy + ch * 8) = val;
            }
            rr += drr; ch += dch;
            if (ch >= cpr) { ch -= cpr; ++rr; }
          }
        }
        tc::named_bar_sync(1, 256);
        if (q == 0) CF_TRACE(3 + grp, it, 7);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == CF_PROD_WARP) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool dim_ok(int d) { return d >= 64 && d <= 256 && d % 64 == 0; }

bool tc_cellf_supported(const smx_cell_weights* w) {
  if (w->mode != SMX_MODE_FULL || w->n_local != 2 || w->n_summary != 2) return false;
  const int D = w->enc_dim;
  if (!dim_ok(D)) return false;
  for (int i = 0; i < 2; ++i) {
    const smx_linear* L[2] = {&w->local[i], &w->summary[i]};
    for (const smx_linear* l : L) {
      if (!l->w || !l->b || !dim_ok(l->in_dim) || !dim_ok(l->out_dim)) return false;
      if (l->n_split > 1 && (l->in_dim % l->n_split || l->out_dim % l->n_split)) return false;
    }
  }
  if (w->local[0].in_dim != D || w->summary[0].in_dim != D) return false;
  if (w->local[1].in_dim != w->local[0].out_dim || w->summary[1].in_dim != w->summary[0].out_dim) return false;
  if (w->local[1].out_dim != w->local_out_dim || w->summary[1].out_dim != w->summary_out_dim) return false;
  if (w->merge.n_split > 1 || !w->merge.w || !w->merge.b) return false;
  if (w->merge.in_dim != w->local_out_dim + w->summary_out_dim || !dim_ok(w->merge.out_dim)) return false;
  if (w->use_layernorm && (!w->local_norm_w || !w->local_norm_b || !w->summary_norm_w || !w->summary_norm_b)) return false;
  return true;
}

// block-diagonal structure the kernel can skip zero blocks for (head dims multiples of 64); anything else is
// packed with its zeros and run as dense
static CfGemm make_gemm(const smx_linear& L, const void* img, int K) {
  CfGemm g{};
  g.img = (const uint8_t*)img;
  g.nkb = K / 64;
  g.nc = L.out_dim / 64;
  g.nheads = 1; g.cph = g.nc; g.kph = g.nkb;
  if (L.n_split > 1) {
    const int a = L.in_dim / L.n_split, b = L.out_dim / L.n_split;
    if (a % 64 == 0 && b % 64 == 0 && a * L.n_split == K) { g.nheads = L.n_split; g.kph = a / 64; g.cph = b / 64; }
  }
  g.gw = g.cph % 4 == 0 ? 4 : (g.cph % 2 == 0 ? 2 : 1);
  return g;
}

size_t tc_cellf_workspace_bytes(const smx_cell_weights* w, int B, int T) {
  const int tpu = (T + 127) / 128;
  return align_up((size_t)B * tpu * w->summary_out_dim * 4) + align_up((size_t)B * w->merge.out_dim * 4);
}

static std::atomic<unsigned long long*> g_trace{nullptr};  // set by smx_debug_set_trace
void tc_set_trace(void* p) { g_trace = (unsigned long long*)p; }

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int PHASE, int ACT>
static int launch_cell_act(const CellFP& p, unsigned grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(cell_kernel<PHASE, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(cell_kernel): %s", cudaGetErrorString(e));
  e = launch_pdl(cell_kernel<PHASE, ACT>, dim3(grid), dim3(CF_THREADS), smem, st, 1u, p);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaLaunchKernelEx(cell_kernel): %s", cudaGetErrorString(e));
  count_tc_launch();
  return check_launch("cell_kernel");
}
template <int PHASE>
static int launch_cell(const CellFP& p, unsigned grid, size_t smem, cudaStream_t st) {
  switch (p.act) {
    case SMX_ACT_SWISH: return launch_cell_act<PHASE, SMX_ACT_SWISH>(p, grid, smem, st);
    case SMX_ACT_GELU: return launch_cell_act<PHASE, SMX_ACT_GELU>(p, grid, smem, st);
    case SMX_ACT_RELU: return launch_cell_act<PHASE, SMX_ACT_RELU>(p, grid, smem, st);
    default: return launch_cell_act<PHASE, -1>(p, grid, smem, st);
  }
}

// per-utterance finalisation between the two passes (shared by both kernel generations)
int tc_cell_finalize(const smx_cell_weights* w, int B, int T, const float* colsum, const uint8_t* mask, float* rowbias, cudaStream_t st) {
  CellFinP f{};
  f.colsum = colsum; f.mask = mask; f.Wc = w->merge.w; f.bc = w->merge.b; f.rowbias = rowbias;
  f.ln_w = w->use_layernorm ? w->summary_norm_w : nullptr;
  f.ln_b = w->use_layernorm ? w->summary_norm_b : nullptr;
  f.T = T; f.tpu = (T + 127) / 128; f.Ds = w->summary_out_dim; f.Dl = w->local_out_dim; f.Dout = w->merge.out_dim;
  cudaError_t e = launch_pdl(cell_finalize2_kernel, dim3(B, (f.Dout + 63) / 64), dim3(256), 0, st, 1u, f);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaLaunchKernelEx(cell_finalize2_kernel): %s", cudaGetErrorString(e));
  count_launch();
  return check_launch("cell_finalize2_kernel");
}

// images: [s1][s2][f1][f2][merge local part], each N*K*2 bytes in 64x64 blocks (NT = 64)
int tc_cellf_fwd(const smx_cell_weights* w, const void* img_s1, const void* img_s2, const void* img_f1, const void* img_f2,
                 const void* img_c, int B, int T, const __nv_bfloat16* x, const float* pre_ln_w, const float* pre_ln_b,
                 const uint8_t* mask, const __nv_bfloat16* residual, __nv_bfloat16* y, Arena& ws, cudaStream_t st) {
  const int tpu = (T + 127) / 128;
  const int Ds = w->summary_out_dim, Dl = w->local_out_dim, Dout = w->merge.out_dim, D = w->enc_dim;
  const size_t m0 = ws.mark();
  float* colsum = ws.f32((size_t)B * tpu * Ds);
  float* rowbias = ws.f32((size_t)B * Dout);
  if (!colsum || !rowbias) return fail(SMX_ERR_WORKSPACE, "workspace too small (fused cell)");

  CellFP p{};
  p.x = x; p.ldx = D; p.pre_w = pre_ln_w; p.pre_b = pre_ln_b; p.mask = mask;
  p.resid = residual; p.ldr = Dout; p.y = y; p.ldy = Dout;
  p.B = B; p.T = T; p.tpu = tpu; p.n_tiles = B * tpu; p.D = D;
  p.Wc = w->merge.w; p.bc = w->merge.b; p.Dl = Dl; p.Ds = Ds; p.Dout = Dout;
  p.act = w->act; p.colsum = colsum; p.rowbias = rowbias;

  auto carve = [&](int ycols) {
    const uint32_t xb = (uint32_t)(D / 64) * kblock_bytes(128), yb = (uint32_t)(ycols / 64) * kblock_bytes(128);
    p.off_y = xb;
    p.off_ring = xb + yb;
    const size_t fixed = (size_t)xb + yb + 8192 /*params*/ + 4096 /*reductions*/ + 1024 /*align*/ + 1024 /*static*/;
    int stages = (int)((227 * 1024 - fixed) / CF_BLOCK_BYTES);
    if (stages > CF_MAX_STAGES) stages = CF_MAX_STAGES;
    stages &= ~3;  // column groups take up to 4 aligned consecutive slots
    p.n_stages = stages;
    p.off_par = p.off_ring + stages * CF_BLOCK_BYTES;
    p.off_red = p.off_par + 8192;
    return (size_t)p.off_red + 4096 + 1024;
  };
  const unsigned grid = (unsigned)(p.n_tiles < num_sms() ? p.n_tiles : num_sms());

  unsigned long long* const trace0 = g_trace.load();
  p.trace = trace0;
  {  // pass A
    p.g[0] = make_gemm(w->summary[0], img_s1, D);
    p.g[1] = make_gemm(w->summary[1], img_s2, w->summary[0].out_dim);
    p.b1 = w->summary[0].b; p.b2 = w->summary[1].b; p.nb1 = w->summary[0].out_dim; p.nb2 = w->summary[1].out_dim;
    p.ln_w = w->use_layernorm ? w->summary_norm_w : nullptr;
    p.ln_b = w->use_layernorm ? w->summary_norm_b : nullptr;
    const size_t smem = carve(w->summary[0].out_dim);
    if (p.n_stages < 4) return fail(SMX_ERR_UNSUPPORTED, "fused cell: tile does not fit shared memory");
    SMX_TRY(launch_cell<0>(p, grid, smem, st));
  }
  SMX_TRY(tc_cell_finalize(w, B, T, colsum, mask, rowbias, st));  // per-utterance mean -> LN_s -> summary share of the combiner
  if (trace0) p.trace = trace0 + 512;
  {  // pass B
    p.g[0] = make_gemm(w->local[0], img_f1, D);
    p.g[1] = make_gemm(w->local[1], img_f2, w->local[0].out_dim);
    smx_linear mg = w->merge;
    mg.n_split = 1;
    p.g[2] = make_gemm(mg, img_c, Dl);
    p.b1 = w->local[0].b; p.b2 = w->local[1].b; p.nb1 = w->local[0].out_dim; p.nb2 = w->local[1].out_dim;
    p.ln_w = w->use_layernorm ? w->local_norm_w : nullptr;
    p.ln_b = w->use_layernorm ? w->local_norm_b : nullptr;
    int ycols = w->local[0].out_dim;
    if (Dl > ycols) ycols = Dl;
    if (Dout > ycols) ycols = Dout;
    const size_t smem = carve(ycols);
    if (p.n_stages < 4) return fail(SMX_ERR_UNSUPPORTED, "fused cell: tile does not fit shared memory");
    SMX_TRY(launch_cell<1>(p, grid, smem, st));
  }
  ws.release(m0);
  return SMX_OK;
}


// GLU pass: out (rows, D) = value * sigmoid(gate), [value | gate] = LN(x) @ W^T + b with W (2D, D) packed with its
// value / gate 64-row blocks interleaved (tc_pack_linear_nt(..., glu_interleave = 1)).          Conformer.py:322-324
int tc_glu_fwd(const smx_linear& L, const void* img, const float* ln_w, const float* ln_b, int64_t rows,
               const __nv_bfloat16* x, __nv_bfloat16* out, cudaStream_t st) {
  const int D = L.in_dim;
  if (L.out_dim != 2 * D || !dim_ok(D) || rows > 0x7fffffff) return fail(SMX_ERR_UNSUPPORTED, "glu pass: D=%d", D);
  CellFP p{};
  p.x = x; p.ldx = D; p.pre_w = ln_w; p.pre_b = ln_b;
  p.y = out; p.ldy = D;
  p.B = 1; p.T = (int)rows; p.tpu = (int)((rows + 127) / 128); p.n_tiles = p.tpu; p.D = D;
  p.Dout = D; p.Dl = D; p.Ds = D;
  smx_linear Ld = L;
  Ld.n_split = 1;
  p.g[0] = make_gemm(Ld, img, D);
  p.b1 = L.b; p.b2 = L.b + D; p.nb1 = D; p.nb2 = D;
  p.trace = nullptr;
  const uint32_t xb = (uint32_t)(D / 64) * kblock_bytes(128);
  p.off_y = xb;
  p.off_ring = 2 * xb;
  p.n_stages = 8;
  p.off_par = p.off_ring + p.n_stages * CF_BLOCK_BYTES;
  p.off_red = p.off_par + 8192;
  const size_t smem = (size_t)p.off_red + 4096;
  const unsigned grid = (unsigned)(p.n_tiles < num_sms() ? p.n_tiles : num_sms());
  return launch_cell_act<2, 0>(p, grid, smem, st);  // the GLU pass has no runtime activation (sigmoid gate only)
}

}  // namespace smx
