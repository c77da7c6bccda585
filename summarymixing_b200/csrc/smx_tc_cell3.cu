// tcgen05 arm of libsmx, part 9: K-SM v3 -- the fused SummaryMixing cell (mode "SummaryMixing", whole-utterance mean,
// summary_mixing.py:198-253) and the GLU pass of the convolution module (Conformer.py:322-324) with every intermediate
// operand resident in TENSOR MEMORY.
//
//   pass A (summary):  X = LN1(x tile)  ->  S = act(act(X W_s1 + b) W_s2 + b) * mask  ->  column sums of the tile
//   finalise (tiny):   per utterance: mean over valid frames, LN_s, c[b] = W_c[:, D_l:] mean + b_c   (smx_tc_cell.cu)
//   pass B (local):    X = LN1(x tile)  ->  L = LN_l(act(act(X W_f1 + b) W_f2 + b) * mask)
//                      ->  y = act(L W_c[:, :D_l]^T + c[b]) (+ residual)
//   GLU pass:          X = LN(x tile)   ->  g = (X W_v + b_v) * sigmoid(X W_g + b_g)
//
// What changed against the first generation (smx_tc_cell.cu), and why (timelines / ncu in profiles/r01_notes.md):
//   * the hidden activation H (and, in pass B, the normalised local branch L) are written back to tensor memory as packed
//     bf16 and consumed from there as the A operand of the next tcgen05.mma: no swizzled shared-memory stores, no operand
//     reads from shared memory, and the 64 KB operand buffer disappears (its space goes to the weight ring);
//   * the weight ring is step-granular: a step is up to four 8 KB blocks that are contiguous in the (schedule-ordered)
//     image: ONE bulk copy, one full/empty barrier pair and one tcgen05.commit per step, one accumulator barrier per GEMM.
//     The single-thread roles pay ~100+ cycles per mbarrier / commit operation, so their count sets the pace;
//   * 16 epilogue warps (four per TMEM lane quadrant, each owning two 32-column pieces of every 256-wide accumulator)
//     instead of 8: the epilogues are latency chains (tcgen05.ld -> math -> tcgen05.st), more warps shorten them directly;
//   * rows leave (and the residual arrives) with 256-bit global accesses straight from / to registers: no staging tile,
//     no block-wide barriers around it.
// TMEM map (512 columns): region A = [0, 256): accumulator of GEMM 1, then H (bf16 pairs) in [0, 128) and, pass B, L in
// [128, 256); region B = [256, 512): accumulator of GEMM 2 (LayerNorm parks its fp32 values there) and of the combiner.
// The GLU pass uses A and B as one 512-column accumulator.
// Warp roles:  0-15 epilogue | 16-19 prologue (cp.async staging + in-place LayerNorm, a thread per row)
//              | 20 weight producer | 21 MMA issuer
#include "smx_tc.h"
#include "smx_tc_common.cuh"

namespace smx {

using tc::kblock_bytes;

constexpr int C3_NEW = 16;
constexpr int C3_PRO_WARP0 = 16, C3_NPW = 4, C3_PROD_WARP = 20, C3_MMA_WARP = 21;
constexpr int C3_THREADS = 22 * 32;
constexpr int C3_MAX_SLOTS = 5;
constexpr uint32_t C3_SLOT = 32768, C3_BLOCK = 8192;

struct C3Gemm {
  const uint8_t* img;     // blocks of 64 rows x 128 B (128B swizzle) in SCHEDULE order (tc_cell3_reorder)
  int nheads, cph, kph;   // block-diagonal structure: heads, 64-column chunks per head, K-blocks per head (dense: 1, nc, nkb)
  int gw;                 // chunks per MMA (1, 2 or 4; divides cph): a unit = gw blocks = one K-block step of one column group
  int n_units, n_steps;   // units = nheads * (cph / gw) * kph; steps of (4 / gw) units (the last one may be shorter)
};

struct Cell3P {
  const __nv_bfloat16* x; int64_t ldx;
  const float* pre_w; const float* pre_b;   // norm1 / conv-module LayerNorm (NULL: none)
  const uint8_t* mask;                       // (B,T) or NULL
  const __nv_bfloat16* resid; int64_t ldr;
  __nv_bfloat16* y; int64_t ldy;
  int B, T, tpu, n_tiles;
  int D;                                     // enc_dim
  C3Gemm g[3];                               // A: s1, s2     B: f1, f2, combiner (local part)     GLU: bottleneck
  const float* b1; const float* b2;          // biases of the two MLP blocks (GLU pass: value / gate halves)
  int nb1, nb2;
  const float* ln_w; const float* ln_b;      // B: local_norm (NULL: no LayerNorm)
  int Ds, Dout;
  int act;
  float* colsum;        // [n_tiles][Ds]
  const float* rowbias; // [B][Dout]
  int nslots;
  uint32_t off_ring, off_par, off_red;
  unsigned long long* trace;  // debug timeline of CTA 0 (NULL: off): role x tile iteration (< 4) x event (< 16)
};

#define C3_TRACE(role, it, ev)                                                                          \
  do {                                                                                                  \
    if (p.trace && blockIdx.x == 0 && lane == 0 && (it) < 4) p.trace[(((role)*4 + (it)) * 16) + (ev)] = clock64(); \
  } while (0)

// warp-collective 16-column TMEM accesses (lane i <-> TMEM lane 32*(w%4)+i)
__device__ __forceinline__ void c3_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void c3_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// D[tmem] (+)= A[tmem: 128 lanes x K bf16, two per 32-bit column] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void c3_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void c3_ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void c3_stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ float2 c3_bf2(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u)); }

// sum over the 32 lanes of v[j] for each j; lane l ends up holding column l's total
__device__ __forceinline__ float c3_column_sums(float* v, int lane) {
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

template <int PHASE, int ACT>  // PHASE 0: pass A (summary), 1: pass B (local + combiner), 2: GLU pass; ACT >= 0: compile-time smx_act
__global__ void __launch_bounds__(C3_THREADS, 1) cell3_kernel(const Cell3P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sX = smem;
  uint8_t* sRing = smem + p.off_ring;
  float* sPar = reinterpret_cast<float*>(smem + p.off_par);   // [b1 | b2 | ln_w | ln_b | c[b] | norm1 w | norm1 b], 256 floats each
  float* sRed = reinterpret_cast<float*>(smem + p.off_red);   // 1536 floats: column partials / LayerNorm statistics
  float2* sStat = reinterpret_cast<float2*>(smem + p.off_red + 6144);  // per-row (1/std, -mean/std) of the prologue LayerNorm
  __shared__ __align__(8) uint64_t full_bar[C3_MAX_SLOTS], empty_bar[C3_MAX_SLOTS];
  __shared__ __align__(8) uint64_t x_full, x_free, x_copied, acc_full[3], op_full[2], epi_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler (role dispatch below)
  constexpr int NG = PHASE == 0 ? 2 : (PHASE == 1 ? 3 : 1);
  // Who normalises a tile.  Pass B's tiles are long (three GEMMs, three epilogues): the four prologue warps copy AND
  // normalise the next tile in their shadow (the CTA's first tile is done by the 16 idle epilogue warps).  Pass A and the
  // GLU pass have short tiles: four solo warps competing with the busy epilogue warps need ~14 k cycles and would sit on
  // the critical path between two tiles, so there the prologue warps only copy (early) and the 16 epilogue warps
  // normalise in place right after the previous tile's epilogue.
  constexpr bool EPI_LN = PHASE != 1;
  const int act = ACT >= 0 ? ACT : p.act;

  // Programmatic dependent launch.  Pass A and the GLU pass read the output of their immediate predecessor: they wait for it
  // below and only then let their own dependents launch.  Pass B reads x, which is three launches old (FFN1 -> pass A ->
  // finalise -> pass B) and therefore complete once pass B can run at all (pass A's CTAs trigger only after their own
  // wait); what pass B needs from its predecessors is c[b], so its epilogue warps wait right before they fetch it (and
  // before the kernel's first global store) -- the first tile's GEMMs and epilogues overlap pass A's tail and the
  // finalisation kernel.
  if (PHASE == 1) tc::pdl_launch_dependents();
  if (warp == C3_PROD_WARP) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int s = 0; s < C3_MAX_SLOTS; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&x_full, C3_NPW); tc::mbar_init(&x_free, 1); tc::mbar_init(&x_copied, C3_NPW * 32); tc::mbar_init(&epi_done, C3_NEW);
    for (int i = 0; i < 3; ++i) tc::mbar_init(&acc_full[i], 1);
    for (int i = 0; i < 2; ++i) tc::mbar_init(&op_full[i], C3_NEW);
    tc::fence_barrier_init();
  }
  for (int i = tid; i < 256; i += C3_THREADS) {
    sPar[i] = i < p.nb1 ? p.b1[i] : 0.0f;
    sPar[256 + i] = i < p.nb2 ? p.b2[i] : 0.0f;
    sPar[512 + i] = (p.ln_w && i < p.nb2) ? p.ln_w[i] : 1.0f;
    sPar[768 + i] = (p.ln_b && i < p.nb2) ? p.ln_b[i] : 0.0f;
    if (i < p.D) {  // norm1 / conv-module LayerNorm parameters, padded layout (tc::ln_pad_index)
      sPar[1280 + tc::ln_pad_index(i, p.D)] = p.pre_w ? p.pre_w[i] : 1.0f;
      sPar[1552 + tc::ln_pad_index(i, p.D)] = p.pre_b ? p.pre_b[i] : 0.0f;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (PHASE != 1) {
    tc::pdl_wait();  // x comes from the preceding kernel (everything above touched only parameters)
    tc::pdl_launch_dependents();
  }
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  const uint32_t regA = tmem, regB = tmem + 256;
  const int first_tile = blockIdx.x, tile_step = gridDim.x;
  const int nslots = p.nslots;

  if (warp == C3_PROD_WARP) {
    // =============================== weight producer: one bulk copy per step ===============================
    if (lane == 0) {
      int s = 0;
      uint32_t pe = 0;
      for (int tile = first_tile; tile < p.n_tiles; tile += tile_step) {
#pragma unroll 1
        for (int gi = 0; gi < NG; ++gi) {
          const C3Gemm& g = p.g[gi];
          const int ups = 4 / g.gw;  // units per step
          for (int st = 0; st < g.n_steps; ++st) {
            const int nu = g.n_units - st * ups < ups ? g.n_units - st * ups : ups;
            const uint32_t bytes = (uint32_t)(nu * g.gw) * C3_BLOCK;
            tc::mbar_wait(&empty_bar[s], ((pe >> s) & 1u) ^ 1u);  // suspending wait: a polling producer floods the SM sub-partition's shared-memory queue
            pe ^= 1u << s;
            tc::mbar_arrive_expect_tx(&full_bar[s], bytes);
            tc::bulk_g2s(sRing + (size_t)s * C3_SLOT, g.img + (size_t)st * C3_SLOT, bytes, &full_bar[s]);
            if (++s == nslots) s = 0;
          }
        }
      }
    }
  } else if (warp == C3_MMA_WARP) {
    // =============================== MMA issuer ===============================
    // The tensor pipe executes in issue order, which protects every TMEM hand-over inside a tile: GEMM 2 reads H before
    // the next tile's GEMM 1 overwrites region A, the combiner reads L likewise.  Cross-warp hand-overs use barriers:
    //   x_full (prologue -> G1), op_full[0] (H stored -> G2), op_full[1] (L stored, region B drained -> combiner),
    //   epi_done (previous tile's last accumulator drained -> first write to region B / to the GLU accumulator).
    int s = 0;
    uint32_t pf = 0;
    const uint32_t x0 = tc::smem_u32(sX), r0 = tc::smem_u32(sRing);
    int it = 0;
    for (int tile = first_tile; tile < p.n_tiles; tile += tile_step, ++it) {
      const uint32_t par = it & 1;
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        const C3Gemm g = p.g[gi];
        const uint32_t idesc = tc::make_idesc_bf16(128, 64u * g.gw);
        const uint32_t dbase = (gi == 0) ? regA : regB;
        if (gi == 0) {
          tc::mbar_wait(&x_full, par);
          if (PHASE == 2 && it > 0) tc::mbar_wait(&epi_done, par ^ 1);
        } else if (gi == 1) {
          tc::mbar_wait_spin(&op_full[0], par);
          if (it > 0) tc::mbar_wait(&epi_done, par ^ 1);
        } else {
          tc::mbar_wait_spin(&op_full[1], par);
        }
        tc::tc_fence_after();
        C3_TRACE(1, it, gi * 2);
        // The issue loop is a serial scalar instruction stream on one warp: keep it short.  Per step: one barrier wait by
        // the warp, then ONE elected lane walks the step's units (up to four 4-MMA groups) in straight-line code with
        // incrementally maintained schedule coordinates (no divisions, no per-unit elect/sync; descriptors advance by adding
        // to their encoded form: +2 in the 16-byte address field per 32-byte K step), then one commit.
        const int ups = 4 / g.gw;  // units per step (1, 2 or 4)
        const uint32_t unit_bytes = (uint32_t)g.gw * C3_BLOCK;
        int m = 0, j = 0, kb = 0, units_left = g.n_units;
        for (int st = 0; st < g.n_steps; ++st) {
          tc::mbar_wait_spin(&full_bar[s], (pf >> s) & 1u);
          pf ^= 1u << s;
          tc::tc_fence_after();
          const int nu = units_left < ups ? units_left : ups;
          if (tc::elect_one()) {
            uint32_t b_addr = r0 + (uint32_t)s * C3_SLOT;
            int mm = m, jj = j, kk = kb;
            for (int u = 0; u < nu; ++u) {
              const uint64_t bd = tc::make_desc_sw128(b_addr);
              const uint32_t d_addr = dbase + (uint32_t)(mm * g.cph + jj) * 64u;
              const uint32_t kba = (uint32_t)(mm * g.kph + kk);  // K-block of the A operand
              if (gi == 0) {
                const uint64_t ad = tc::make_desc_sw128(x0 + kba * kblock_bytes(128));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) tc::umma_bf16(d_addr, ad + 2u * ks, bd + 2u * ks, idesc, (kk == 0 && ks == 0) ? 0u : 1u);
              } else {
                const uint32_t at = regA + (gi == 2 ? 128u : 0u) + kba * 32u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) c3_umma_ts(d_addr, at + 8u * ks, bd + 2u * ks, idesc, (kk == 0 && ks == 0) ? 0u : 1u);
              }
              b_addr += unit_bytes;
              if (++kk == g.kph) { kk = 0; jj += g.gw; if (jj == g.cph) { jj = 0; ++mm; } }
            }
            tc::umma_commit(&empty_bar[s]);  // one commit per step
          }
          __syncwarp();
          // every lane advances the schedule coordinates by the step's units (warp-uniform state)
          for (int u = 0; u < nu; ++u) { if (++kb == g.kph) { kb = 0; j += g.gw; if (j == g.cph) { j = 0; ++m; } } }
          units_left -= nu;
          if (++s == nslots) s = 0;
        }
        if (tc::elect_one()) {
          tc::umma_commit(&acc_full[gi]);
          if (gi == 0) tc::umma_commit(&x_free);
        }
        __syncwarp();
        C3_TRACE(1, it, gi * 2 + 1);
      }
    }
  } else if (warp >= C3_PRO_WARP0) {
    // =============================== prologue: x tile -> LN -> A operand ===============================
    const int pw = warp - C3_PRO_WARP0;
    int it = EPI_LN ? 0 : 1;  // (pass B: the CTA's first tile is staged by the 16 epilogue warps, idle at that point)
    for (int tile = first_tile + (EPI_LN ? 0 : tile_step); tile < p.n_tiles; tile += tile_step, ++it) {
      const int b = tile / p.tpu, t0 = (tile % p.tpu) * 128;
      const int64_t row0 = (int64_t)b * p.T + t0;
      const int nrows = p.T - t0 < 128 ? p.T - t0 : 128;
      if (it > 0) tc::mbar_wait(&x_free, (it - 1) & 1);
      if (pw == 0) C3_TRACE(2, it, 0);
      if (EPI_LN) {  // copy only
#pragma unroll 1
        for (int i = 0; i < 4; ++i) tc::rows8_copy(sX, p.x, p.ldx, row0, nrows, p.D, pw * 4 + i, lane);
        tc::cp_async_commit();
        tc::cp_async_wait_all();
        __syncwarp();
        tc::mbar_arrive(&x_copied);  // every lane: each thread publishes its own cp.async writes (count = all copying threads)
      } else {
        tc::stage_ln_rows_wide(sX, p.x, p.ldx, row0, nrows, p.D, pw * 4, 4, lane, p.pre_w != nullptr, sPar + 1280, sPar + 1552, sStat);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&x_full);
      }
      if (pw == 0) C3_TRACE(2, it, 1);
    }
  } else {
    // =============================== epilogue ===============================
    const int q = warp & 3, k = warp >> 2;   // TMEM lane quadrant; this warp's 32-column pieces are k and k + 4
    const int r = q * 32 + lane;             // row inside the tile
    const int etid = tid;                    // 0..511
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const float* sB1 = sPar; const float* sB2 = sPar + 256; const float* sLw = sPar + 512; const float* sLb = sPar + 768;
    float* sRB = sPar + 1024;
    const int H1 = p.g[0].nheads * p.g[0].cph * 64;                      // width of GEMM 1's output
    const int H2 = PHASE == 2 ? 0 : p.g[1].nheads * p.g[1].cph * 64;     // width of GEMM 2's output (D_s / D_l)
    int it = 0;
    for (int tile = first_tile; tile < p.n_tiles; tile += tile_step, ++it) {
      const uint32_t par = it & 1;
      const int b = tile / p.tpu, t0 = (tile % p.tpu) * 128;
      const int64_t row0 = (int64_t)b * p.T + t0;
      const int nrows = p.T - t0 < 128 ? p.T - t0 : 128;
      if (warp == 0) C3_TRACE(3, it, 0);
      if (EPI_LN) {  // the tile has been copied into the operand image: LayerNorm in place, 8 rows per warp
        tc::mbar_wait(&x_copied, par);
        if (warp == 0) C3_TRACE(3, it, 10);
        if (p.pre_w) tc::rows8_ln(sX, nrows, p.D, warp, lane, sPar + 1280, sPar + 1552, sStat,
                                  (p.trace && blockIdx.x == 0 && warp == 0 && it < 4) ? p.trace + ((3 * 4 + it) * 16) + 9 : nullptr);
        if (warp == 0) C3_TRACE(3, it, 11);
        tc::fence_proxy_async();
        if (warp == 0) C3_TRACE(3, it, 12);
        tc::named_bar_sync(5, C3_NEW * 32);
        if (warp < C3_NPW && lane == 0) tc::mbar_arrive(&x_full);
        if (warp == 0) C3_TRACE(3, it, 1);
      } else if (it == 0) {  // first tile: all 16 epilogue warps stage and normalise it (8 rows each)
        tc::stage_ln_rows_wide(sX, p.x, p.ldx, row0, nrows, p.D, warp, 1, lane, p.pre_w != nullptr, sPar + 1280, sPar + 1552, sStat);
        tc::fence_proxy_async();
        tc::named_bar_sync(5, C3_NEW * 32);
        if (warp < C3_NPW && lane == 0) tc::mbar_arrive(&x_full);
        if (warp == 0) C3_TRACE(3, it, 1);
      }
      const bool live = r < nrows;
      const float rscale = live ? (p.mask ? (float)p.mask[row0 + r] : 1.0f) : 0.0f;

      if (PHASE == 2) {
        // ---- GLU: g = (acc[value] + b_v) * sigmoid(acc[gate] + b_g); value / gate 64-column chunks are interleaved
        // (chunk 2c / 2c+1).  This warp: output columns [64k, 64k + 64) in four pieces of 16, one 256-bit store each.
        tc::mbar_wait(&acc_full[0], par);
        tc::tc_fence_after();
        if (warp == 0) C3_TRACE(3, it, 2);
        if (k * 64 < p.Dout) {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float v[16], gt[16];
            c3_ld16(tmem + lane_sel + (uint32_t)(2 * k) * 64u + h * 16, v);
            c3_ld16(tmem + lane_sel + (uint32_t)(2 * k + 1) * 64u + h * 16, gt);
            tc::tmem_ld_wait();
            const float4* ba = reinterpret_cast<const float4*>(sB1 + k * 64 + h * 16);
            const float4* bg = reinterpret_cast<const float4*>(sB2 + k * 64 + h * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 a4 = ba[i], g4 = bg[i];
              v[4 * i] = (v[4 * i] + a4.x) * tc::act_sigmoid(gt[4 * i] + g4.x);
              v[4 * i + 1] = (v[4 * i + 1] + a4.y) * tc::act_sigmoid(gt[4 * i + 1] + g4.y);
              v[4 * i + 2] = (v[4 * i + 2] + a4.z) * tc::act_sigmoid(gt[4 * i + 2] + g4.z);
              v[4 * i + 3] = (v[4 * i + 3] + a4.w) * tc::act_sigmoid(gt[4 * i + 3] + g4.w);
            }
            if (live) {
              uint32_t o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = tc::pack_bf16x2(v[2 * i], v[2 * i + 1]);
              c3_stg256(p.y + (row0 + r) * p.ldy + k * 64 + h * 16, o);
            }
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&epi_done);
        if (warp == 0) C3_TRACE(3, it, 3);
        continue;
      }

      // ---- E1: H = act(acc1 + b1) as packed bf16 into region A [0, 128) (the A operand of GEMM 2).  The overlay is
      // safe in two rounds: round 0 covers accumulator columns [0, 128) -- once every warp of the lane quadrant holds its
      // piece (barrier), their H pairs go to columns [0, 64); round 1 covers [128, 256) and writes to [64, 128), columns
      // that round 0 has already consumed.
      tc::mbar_wait(&acc_full[0], par);
      tc::tc_fence_after();
      if (warp == 0) C3_TRACE(3, it, 2);
#pragma unroll 1
      for (int rd = 0; rd < 2; ++rd) {
        const int piece = k + 4 * rd, col = piece * 32;
        float v[32];
        if (col < H1) {
          tc::tmem_ld32(regA + lane_sel + col, v);
          tc::tmem_ld_wait();
        }
        if (rd == 0) tc::named_bar_sync(1 + q, 128);
        if (col < H1) {
          const float4* bp = reinterpret_cast<const float4*>(sB1 + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float4 bb = bp[j]; v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
          tc::act_apply<32>(act, v);
          uint32_t hp[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) hp[i] = tc::pack_bf16x2(v[2 * i], v[2 * i + 1]);
          c3_st16(regA + lane_sel + piece * 16, hp);
        }
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&op_full[0]);
      if (warp == 0) C3_TRACE(3, it, 3);

      if (PHASE == 0) {
        // ---- E2': S = act(acc2 + b2) * mask -> column sums of this tile                      :221, 229-231
        tc::mbar_wait(&acc_full[1], par);
        tc::tc_fence_after();
        if (warp == 0) C3_TRACE(3, it, 4);
#pragma unroll 1
        for (int rd = 0; rd < 2; ++rd) {
          const int col = (k + 4 * rd) * 32;
          if (col < H2) {
            float v[32];
            tc::tmem_ld32(regB + lane_sel + col, v);
            tc::tmem_ld_wait();
            const float4* bp = reinterpret_cast<const float4*>(sB2 + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 bb = bp[j]; v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
            tc::act_apply<32>(act, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= rscale;
            const float tot = c3_column_sums(v, lane);
            sRed[q * 256 + col + lane] = tot;
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&epi_done);
        tc::named_bar_sync(5, C3_NEW * 32);
        if (etid < p.Ds)  // fixed-order reduction over the four row quadrants: deterministic
          p.colsum[(size_t)tile * p.Ds + etid] = (sRed[etid] + sRed[256 + etid]) + (sRed[512 + etid] + sRed[768 + etid]);
        tc::named_bar_sync(5, C3_NEW * 32);  // sRed is rewritten by the next tile
        if (warp == 0) C3_TRACE(3, it, 5);
      } else {
        // ---- E2: L = LN_l(act(acc2 + b2) * mask) as packed bf16 into region A [128, 256) (A operand of the combiner)
        tc::mbar_wait(&acc_full[1], par);
        tc::tc_fence_after();
        if (warp == 0) C3_TRACE(3, it, 4);
        float mean = 0.0f, rstd = 1.0f;
        if (p.ln_w) {
          // pass 1: activated, masked values parked as fp32 in region B; per-thread (mean, M2) over its <= 64 values
          float mean_t = 0.0f, m2_t = 0.0f, n_t = 0.0f;
#pragma unroll 1
          for (int rd = 0; rd < 2; ++rd) {
            const int col = (k + 4 * rd) * 32;
            if (col < H2) {
              float v[32];
              tc::tmem_ld32(regB + lane_sel + col, v);
              tc::tmem_ld_wait();
              const float4* bp = reinterpret_cast<const float4*>(sB2 + col);
#pragma unroll
              for (int j = 0; j < 8; ++j) { const float4 bb = bp[j]; v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
              tc::act_apply<32>(act, v);
              float sa[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // four independent chains (fixed association: deterministic)
#pragma unroll
              for (int j = 0; j < 32; ++j) { v[j] *= rscale; sa[j & 3] += v[j]; }
              const float mh = ((sa[0] + sa[1]) + (sa[2] + sa[3])) * (1.0f / 32.0f);
              float qa[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
              for (int j = 0; j < 32; ++j) { const float d = v[j] - mh; qa[j & 3] = fmaf(d, d, qa[j & 3]); }
              const float qh = (qa[0] + qa[1]) + (qa[2] + qa[3]);
              if (n_t == 0.0f) { mean_t = mh; m2_t = qh; n_t = 32.0f; }
              else { const float dl = mh - mean_t; mean_t += 0.5f * dl; m2_t += qh + dl * dl * 16.0f; n_t = 64.0f; }
              tc::tmem_st32(regB + lane_sel + col, v);
            }
          }
          tc::tmem_st_wait();
          if (warp == 0) C3_TRACE(3, it, 8);
          sRed[(k * 128 + r) * 3] = mean_t; sRed[(k * 128 + r) * 3 + 1] = m2_t; sRed[(k * 128 + r) * 3 + 2] = n_t;
          tc::named_bar_sync(1 + q, 128);
          // Chan's merge of the (up to) four per-thread partials of this row, in fixed order
          float n = 0.0f, m2 = 0.0f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float mi = sRed[(i * 128 + r) * 3], qi = sRed[(i * 128 + r) * 3 + 1], ni = sRed[(i * 128 + r) * 3 + 2];
            if (ni > 0.0f) {
              const float dl = mi - mean, nn = n + ni;
              mean += dl * (ni / nn);
              m2 += qi + dl * dl * (n * ni / nn);
              n = nn;
            }
          }
          rstd = rsqrtf(m2 / n + 1e-5f);
          if (warp == 0) C3_TRACE(3, it, 9);
        }
#pragma unroll 1
        for (int rd = 0; rd < 2; ++rd) {
          const int piece = k + 4 * rd, col = piece * 32;
          if (col < H2) {
            float v[32];
            tc::tmem_ld32(regB + lane_sel + col, v);
            tc::tmem_ld_wait();
            if (p.ln_w) {
              const float4* wp = reinterpret_cast<const float4*>(sLw + col);
              const float4* bp = reinterpret_cast<const float4*>(sLb + col);
              const float shift = -mean * rstd;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 ww = wp[j], bb = bp[j];
                v[4 * j] = fmaf(fmaf(v[4 * j], rstd, shift), ww.x, bb.x);
                v[4 * j + 1] = fmaf(fmaf(v[4 * j + 1], rstd, shift), ww.y, bb.y);
                v[4 * j + 2] = fmaf(fmaf(v[4 * j + 2], rstd, shift), ww.z, bb.z);
                v[4 * j + 3] = fmaf(fmaf(v[4 * j + 3], rstd, shift), ww.w, bb.w);
              }
            } else {
              const float4* bp = reinterpret_cast<const float4*>(sB2 + col);
#pragma unroll
              for (int j = 0; j < 8; ++j) { const float4 bb = bp[j]; v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w; }
              tc::act_apply<32>(act, v);
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= rscale;
            }
            uint32_t lp[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) lp[i] = tc::pack_bf16x2(v[2 * i], v[2 * i + 1]);
            c3_st16(regA + 128 + lane_sel + piece * 16, lp);
          }
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        if (it == 0) tc::pdl_wait();  // pass A and the finalisation kernel have completed: c[b] is there, y may be written
        if (etid < p.Dout) sRB[etid] = __ldcg(p.rowbias + (size_t)b * p.Dout + etid);  // c[b] for E3
        tc::named_bar_sync(5, C3_NEW * 32);  // c[b] is in shared memory; sRed may be rewritten
        if (lane == 0) tc::mbar_arrive(&op_full[1]);
        if (warp == 0) C3_TRACE(3, it, 5);

        // ---- E3: y = act(acc3 + c[b]) (+ residual); this thread: row r, output columns [64k, 64k + 64) in four pieces
        // of 16 (one 256-bit residual load and one 256-bit store each)                           :251-253, :541
        const bool active = k * 64 < p.Dout;
        uint32_t rres[32];
        if (active && p.resid) {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            if (live) c3_ldg256(p.resid + (row0 + r) * p.ldr + k * 64 + h * 16, rres + 8 * h);
            else {
#pragma unroll
              for (int e = 0; e < 8; ++e) rres[8 * h + e] = 0u;
            }
          }
        }
        tc::mbar_wait(&acc_full[2], par);
        tc::tc_fence_after();
        if (warp == 0) C3_TRACE(3, it, 6);
        if (active) {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int col = k * 64 + h * 16;
            float v[16];
            c3_ld16(regB + lane_sel + col, v);
            tc::tmem_ld_wait();
            const float4* bp = reinterpret_cast<const float4*>(sRB + col);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float4 bb = bp[i]; v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w; }
            tc::act_apply<16>(act, v);
            if (p.resid) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { const float2 f = c3_bf2(rres[h * 8 + i]); v[2 * i] += f.x; v[2 * i + 1] += f.y; }
            }
            if (live) {
              uint32_t o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = tc::pack_bf16x2(v[2 * i], v[2 * i + 1]);
              c3_stg256(p.y + (row0 + r) * p.ldy + col, o);
            }
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&epi_done);
        if (warp == 0) C3_TRACE(3, it, 7);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == C3_PROD_WARP) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static std::atomic<unsigned long long*> g_trace_c3{nullptr};  // set by smx_debug_set_trace
void tc_set_trace_cell3(void* p) { g_trace_c3 = (unsigned long long*)p; }
static std::atomic<int> g_cell_ver{4};  // smx_debug_set_cell_version: 1 = first generation (smx_tc_cell.cu), 3 = this file, 4 = smx_tc_cell4.cu (default)
void tc_set_cell_version(int v) { g_cell_ver = (v == 1 || v == 3) ? v : 4; }
int tc_cell_version() { return g_cell_ver; }

static bool c3_dim_ok(int d) { return d >= 64 && d <= 256 && d % 64 == 0; }

// schedule description of one linear layer (the same rules as the first generation: heads whose dims are multiples of 64
// are walked block-diagonally, anything else as dense)
static C3Gemm c3_make_gemm(const smx_linear& L, int K, int n_split) {
  C3Gemm g{};
  const int nkb = K / 64, nc = L.out_dim / 64;
  g.nheads = 1; g.cph = nc; g.kph = nkb;
  if (n_split > 1) {
    const int a = L.in_dim / n_split, b = L.out_dim / n_split;
    if (a % 64 == 0 && b % 64 == 0 && a * n_split == K) { g.nheads = n_split; g.kph = a / 64; g.cph = b / 64; }
  }
  g.gw = g.cph % 4 == 0 ? 4 : (g.cph % 2 == 0 ? 2 : 1);
  g.n_units = g.nheads * (g.cph / g.gw) * g.kph;
  const int ups = 4 / g.gw;
  g.n_steps = (g.n_units + ups - 1) / ups;
  return g;
}

// image in schedule order <- image in [chunk][K-block] order (tc_pack_linear_nt, NT = 64); zero blocks of block-diagonal
// weights are dropped.  One CTA per scheduled 8 KB block.
__global__ void cell3_reorder_kernel(const uint4* src, uint4* dst, int cph, int kph, int gw, int nkb) {
  const int blk = blockIdx.x, per_head = cph * kph;
  const int m = blk / per_head, rem = blk % per_head;
  const int jg = rem / (kph * gw), rem2 = rem % (kph * gw), kbl = rem2 / gw, u = rem2 % gw;
  const int c = m * cph + jg * gw + u, kb = m * kph + kbl;
  const uint4* s = src + (size_t)(c * nkb + kb) * 512;
  uint4* d = dst + (size_t)blk * 512;
  for (int i = threadIdx.x; i < 512; i += blockDim.x) d[i] = s[i];
}
int tc_cell3_reorder(const smx_linear& L, int K, int n_split, const void* img_chunk_major, void* img_sched, cudaStream_t st) {
  const C3Gemm g = c3_make_gemm(L, K, n_split);
  cell3_reorder_kernel<<<g.n_units * g.gw, 128, 0, st>>>((const uint4*)img_chunk_major, (uint4*)img_sched, g.cph, g.kph, g.gw, K / 64);
  count_launch();
  return check_launch("cell3_reorder_kernel");
}

bool tc_cell3_supported(const smx_cell_weights* w) {
  if (!tc_cellf_supported(w)) return false;
  return c3_dim_ok(w->enc_dim);
}

static int c3_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static size_t c3_carve(Cell3P& p, int D) {
  const uint32_t xb = (uint32_t)(D / 64) * kblock_bytes(128);
  p.off_ring = xb;
  const size_t fixed = (size_t)xb + 8192 /*params*/ + 7168 /*reductions + row statistics*/ + 1024 /*align*/ + 1024 /*static*/;
  int slots = (int)((227 * 1024 - fixed) / C3_SLOT);
  if (slots > C3_MAX_SLOTS) slots = C3_MAX_SLOTS;
  p.nslots = slots;
  p.off_par = p.off_ring + (uint32_t)slots * C3_SLOT;
  p.off_red = p.off_par + 8192;
  return (size_t)p.off_red + 7168 + 1024;
}

template <int PHASE, int ACT>
static int launch_cell3_act(const Cell3P& p, unsigned grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(cell3_kernel<PHASE, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(cell3_kernel): %s", cudaGetErrorString(e));
  e = launch_pdl(cell3_kernel<PHASE, ACT>, dim3(grid), dim3(C3_THREADS), smem, st, 1u, p);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaLaunchKernelEx(cell3_kernel): %s", cudaGetErrorString(e));
  count_tc_launch();
  return check_launch("cell3_kernel");
}
template <int PHASE>
static int launch_cell3(const Cell3P& p, unsigned grid, size_t smem, cudaStream_t st) {
  switch (p.act) {
    case SMX_ACT_SWISH: return launch_cell3_act<PHASE, SMX_ACT_SWISH>(p, grid, smem, st);
    case SMX_ACT_GELU: return launch_cell3_act<PHASE, SMX_ACT_GELU>(p, grid, smem, st);
    case SMX_ACT_RELU: return launch_cell3_act<PHASE, SMX_ACT_RELU>(p, grid, smem, st);
    default: return launch_cell3_act<PHASE, -1>(p, grid, smem, st);
  }
}

// images (schedule order): [s1][s2][f1][f2][merge local part]
int tc_cell3_fwd(const smx_cell_weights* w, const void* img_s1, const void* img_s2, const void* img_f1, const void* img_f2,
                 const void* img_c, int B, int T, const __nv_bfloat16* x, const float* pre_ln_w, const float* pre_ln_b,
                 const uint8_t* mask, const __nv_bfloat16* residual, __nv_bfloat16* y, Arena& ws, cudaStream_t st) {
  const int tpu = (T + 127) / 128;
  const int Ds = w->summary_out_dim, Dl = w->local_out_dim, Dout = w->merge.out_dim, D = w->enc_dim;
  const size_t m0 = ws.mark();
  float* colsum = ws.f32((size_t)B * tpu * Ds);
  float* rowbias = ws.f32((size_t)B * Dout);
  if (!colsum || !rowbias) return fail(SMX_ERR_WORKSPACE, "workspace too small (fused cell)");

  Cell3P p{};
  p.x = x; p.ldx = D; p.pre_w = pre_ln_w; p.pre_b = pre_ln_b; p.mask = mask;
  p.resid = residual; p.ldr = Dout; p.y = y; p.ldy = Dout;
  p.B = B; p.T = T; p.tpu = tpu; p.n_tiles = B * tpu; p.D = D;
  p.Ds = Ds; p.Dout = Dout;
  p.act = w->act; p.colsum = colsum; p.rowbias = rowbias;
  const size_t smem = c3_carve(p, D);
  if (p.nslots < 2) return fail(SMX_ERR_UNSUPPORTED, "fused cell: tile does not fit shared memory");
  const unsigned grid = (unsigned)(p.n_tiles < c3_sms() ? p.n_tiles : c3_sms());

  unsigned long long* const trace0 = g_trace_c3.load();
  p.trace = trace0;
  {  // pass A
    p.g[0] = c3_make_gemm(w->summary[0], D, w->summary[0].n_split); p.g[0].img = (const uint8_t*)img_s1;
    p.g[1] = c3_make_gemm(w->summary[1], w->summary[0].out_dim, w->summary[1].n_split); p.g[1].img = (const uint8_t*)img_s2;
    p.b1 = w->summary[0].b; p.b2 = w->summary[1].b; p.nb1 = w->summary[0].out_dim; p.nb2 = w->summary[1].out_dim;
    p.ln_w = nullptr; p.ln_b = nullptr;
    SMX_TRY(launch_cell3<0>(p, grid, smem, st));
  }
  SMX_TRY(tc_cell_finalize(w, B, T, colsum, mask, rowbias, st));  // per-utterance mean -> LN_s -> summary share of the combiner
  if (trace0) p.trace = trace0 + 512;
  {  // pass B
    p.g[0] = c3_make_gemm(w->local[0], D, w->local[0].n_split); p.g[0].img = (const uint8_t*)img_f1;
    p.g[1] = c3_make_gemm(w->local[1], w->local[0].out_dim, w->local[1].n_split); p.g[1].img = (const uint8_t*)img_f2;
    p.g[2] = c3_make_gemm(w->merge, Dl, 1); p.g[2].img = (const uint8_t*)img_c;
    // the combiner's packed image covers only its local part: K = D_l
    p.b1 = w->local[0].b; p.b2 = w->local[1].b; p.nb1 = w->local[0].out_dim; p.nb2 = w->local[1].out_dim;
    p.ln_w = w->use_layernorm ? w->local_norm_w : nullptr;
    p.ln_b = w->use_layernorm ? w->local_norm_b : nullptr;
    SMX_TRY(launch_cell3<1>(p, grid, smem, st));
  }
  ws.release(m0);
  return SMX_OK;
}

// GLU pass: out (rows, D) = value * sigmoid(gate), [value | gate] = LN(x) @ W^T + b with W (2D, D) packed with its
// value / gate 64-row blocks interleaved and then put in schedule order.                     Conformer.py:322-324
int tc_glu3_fwd(const smx_linear& L, const void* img_sched, const float* ln_w, const float* ln_b, int64_t rows,
                const __nv_bfloat16* x, __nv_bfloat16* out, cudaStream_t st) {
  const int D = L.in_dim;
  if (L.out_dim != 2 * D || !c3_dim_ok(D) || rows > 0x7fffffff) return fail(SMX_ERR_UNSUPPORTED, "glu pass: D=%d", D);
  Cell3P p{};
  p.x = x; p.ldx = D; p.pre_w = ln_w; p.pre_b = ln_b;
  p.y = out; p.ldy = D;
  p.B = 1; p.T = (int)rows; p.tpu = (int)((rows + 127) / 128); p.n_tiles = p.tpu; p.D = D;
  p.Dout = D; p.Ds = D;
  p.g[0] = c3_make_gemm(L, D, 1); p.g[0].img = (const uint8_t*)img_sched;
  p.b1 = L.b; p.b2 = L.b + D; p.nb1 = D; p.nb2 = D;
  p.trace = g_trace_c3.load();  // (diagnostics) the GLU pass writes its timeline where pass A would
  const size_t smem = c3_carve(p, D);
  const unsigned grid = (unsigned)(p.n_tiles < c3_sms() ? p.n_tiles : c3_sms());
  return launch_cell3_act<2, 0>(p, grid, smem, st);  // the GLU pass has no runtime activation (sigmoid gate only)
}

}  // namespace smx
