// tcgen05 arm of libsmx, part 8: K-FFN v3, the persistent fused macaron feed-forward half-step with the hidden
// activation resident in TENSOR MEMORY
//
//   y = x + 0.5 * ( W2 @ act( W1 @ LN(x) + b1 ) + b2 )        [optionally y = LN_out(y)]
//   (Conformer.py:470-484, :518, :547)
//
// One CTA per SM walks 128-row tiles; the hidden dimension is processed in 128-wide chunks.  Differences from v2
// (smx_tc_ffn2.cu), each aimed at what the v2 timelines and ncu captures showed (profiles/r01_notes.md):
//   * GEMM1 (K = D, N = 128) fills one of two fp32 TMEM accumulators; the epilogue warps add b1, activate and write the
//     chunk back as packed bf16 INTO THE SAME TMEM COLUMNS (each warp over the first half of its own 32): GEMM2 takes it from there as its
//     A operand (tcgen05.mma with A in tensor memory).  The hidden activation costs no shared-memory write, no
//     shared-memory operand read and no shared-memory space -- v2 was bound by shared-memory bandwidth plus the L2
//     latency its 64 KB weight ring could not cover.
//   * the freed 64 KB go to the weight ring: 16 slots of 8 KB (128 KB in flight / being consumed).
//   * 16 epilogue warps (four per TMEM lane quadrant, 32 hidden columns each) instead of 8: the chunk epilogue is a
//     latency chain (tcgen05.ld -> bias/activation -> pack -> tcgen05.st), so more warps shorten it directly.
//   * the final epilogue needs no staging: a thread owns 64 consecutive output columns of its row and moves them with
//     256-bit global loads (residual) and stores (y); the output LayerNorm combines per-thread (mean, M2) pairs.
// Warp roles:  0-15 epilogue | 16-19 prologue (cp.async tile staging + in-place LayerNorm, a thread per row)
//              | 20 weight producer (cp.async.bulk + mbarrier ring) | 21 MMA issuer
#include "smx_tc.h"
#include "smx_tc_common.cuh"

namespace smx {

using tc::kblock_bytes;

constexpr int F3_NEW = 16;                       // epilogue warps
constexpr int F3_PRO_WARP0 = 16, F3_NPW = 4, F3_PROD_WARP = 20, F3_MMA_WARP = 21;
constexpr int F3_THREADS = 22 * 32;
constexpr int F3_HC = 128;                       // hidden chunk width
constexpr int F3_STAGES = 8;                     // ring slots: 4 x 32 KB (one CTA per tile) or 8 x 16 KB (CTA pairs)
constexpr uint32_t F3_RING_BYTES = 131072;
constexpr uint32_t F3_BLOCK = 8192;              // one 64 x 64 bf16 weight block

struct Ffn3P {
  const __nv_bfloat16* x; __nv_bfloat16* y; int64_t rows;
  int D, F, n_tiles;
  const uint8_t* w1; const uint8_t* w2;   // v3 images of 64 x 64 blocks in step order: w1 [F/128][D/64][2], w2 [F/64][D/64]
  const float* b1; const float* b2;        // b1: first-layer bias with the LayerNorm's beta share folded in (tc_ffn3_pack)
  const float* gw1;                        // [F] gw1[n] = sum_k bf16(gamma_k W1[n,k]): the mean's share of GEMM 1 (zeros without LayerNorm)
  int has_ln;
  const float* oln_w; const float* oln_b; float oln_eps;
  int act;
  unsigned long long* trace;
  uint32_t off_ring, off_par, off_red;
  int cl2;                                 // 1: CTA pairs (cta_group::2)
  uint32_t ring_bytes;                     // pair mode: bytes of the weight ring in use (the rest of the 128 KB region stages the output tile)
};

#define F3_TRACE(role, it, ev)                                                                                \
  do {                                                                                                        \
    if (p.trace && blockIdx.x == 0 && lane == 0 && (it) < 2) p.trace[(((role)*2 + (it)) * 32) + (ev)] = clock64(); \
  } while (0)

// D[tmem] (+)= A[tmem: 128 lanes x K bf16, two per 32-bit column] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void f3_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// warp-collective: lane i of warp w writes 16 consecutive 32-bit columns of TMEM lane 32*(w%4)+i
__device__ __forceinline__ void f3_tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// warp-collective 16-column fp32 load / store (lane i <-> TMEM lane 32*(w%4)+i)
__device__ __forceinline__ void f3_tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void f3_tmem_st16f(uint32_t taddr, const float* v) { f3_tmem_st16(taddr, reinterpret_cast<const uint32_t*>(v)); }
__device__ __forceinline__ void f3_ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void f3_stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void f3_tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(tc::smem_u32(bar))
               : "memory");
}
// One 64 x 64 block of a packed weight image into this CTA's ring, completion (complete_tx) on the barrier at the same offset in
// the PAIR LEADER's shared memory (.cta_group::2: the mbarrier may live in either CTA of the pair) -- the leader's issuer learns
// that both halves of a step have landed from its own barrier, without a relay through the peer's warps.
__device__ __forceinline__ void f3_tma_block_pair(uint32_t smem_dst, const CUtensorMap* tmap, int block, uint32_t leader_bar) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_dst),
               "l"(tmap), "r"(0), "r"(block * 64), "r"(leader_bar)
               : "memory");
}
__device__ __forceinline__ uint32_t f3_leader_addr(const void* p) {  // the same shared-memory offset in CTA 0 of the cluster
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(tc::smem_u32(p)), "r"(0u));
  return ra;
}
__device__ __forceinline__ void f3_tma_store_2d(const CUtensorMap* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(smem_src) : "memory");
}
__device__ __forceinline__ float2 f3_bf2(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u)); }

template <bool OLN, int ACT, bool CL2>  // ACT >= 0: compile-time activation (smx_act); -1: runtime p.act; CL2: CTA pairs
__global__ void __launch_bounds__(F3_THREADS, 1) ffn3_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                                                              const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_y, const Ffn3P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sX = smem;
  uint8_t* sRing = smem + p.off_ring;
  float* sPar = reinterpret_cast<float*>(smem + p.off_par);  // [b1 (F) | b2 | oln_w | oln_b | ln_w | ln_b (256 each)]
  float* sRed = reinterpret_cast<float*>(smem + p.off_red);  // [4 column quarters][128 rows][2]: per-thread (mean, M2)
  float2* sStat = reinterpret_cast<float2*>(smem + p.off_red + 4096);  // [2 (tile parity)][128] per-row (1/std, -mean/std) of the input LayerNorm
  __shared__ __align__(8) uint64_t full_bar[F3_STAGES], empty_bar[F3_STAGES];
  __shared__ __align__(8) uint64_t x_full, x_free, x_landed, stat_full, acc1_full[2], h_full[2], acc2_full, epi_done;
  __shared__ __align__(8) uint64_t res_full[4];  // pair mode: the residual's 64-column block k has landed in the staging tile
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int D = p.D, nkbD = D / 64, nj = p.F / F3_HC, nkbF = p.F / 64;
  const int act = ACT >= 0 ? ACT : p.act;
  // CTA pair (CL2): the two CTAs of a cluster take the two 128-row tiles of a 256-row pair; every tcgen05.mma is issued by
  // the leader (rank 0) with cta_group::2 and spans both SMs: M = 256 (each CTA its own A operand and accumulators), each
  // CTA streams only HALF of every weight step (N/2 rows of B) from L2 into its own ring -- half the L2 requests and
  // shared-memory fill per SM, and the 16-slot ring covers two hidden chunks instead of one.  Cross-CTA protocol:
  //   leader <- both CTAs : x_full, h_full, epi_done (arrivals from the peer are remote mbarrier arrives), full_bar[s] (the peer's
  //                         weight copies are cp.async.bulk.tensor with .cta_group::2 and complete on the leader's barrier)
  //   leader -> both CTAs : empty_bar[s], acc1_full, acc2_full, x_free (tcgen05.commit multicast to the pair)
  constexpr uint32_t NCTA = CL2 ? 2u : 1u;
  const uint32_t crank = CL2 ? tc::cluster_ctarank() : 0u;
  const bool leader = crank == 0;

  tc::pdl_launch_dependents();  // the next kernel may start taking over SMs as CTAs of this grid leave them
  if (warp == F3_PROD_WARP) { if (CL2) tc::tmem_alloc2(&tmem_base_s, 512); else tc::tmem_alloc(&tmem_base_s, 512); }
  if (tid == 0) {
    for (int s = 0; s < F3_STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&x_full, F3_NPW * NCTA); tc::mbar_init(&x_free, 1); tc::mbar_init(&stat_full, F3_NPW * 32); tc::mbar_init(&x_landed, 1); tc::mbar_init(&acc2_full, 1); tc::mbar_init(&epi_done, F3_NEW * NCTA);
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc1_full[i], 1); tc::mbar_init(&h_full[i], F3_NEW * NCTA); }
    for (int i = 0; i < 4; ++i) tc::mbar_init(&res_full[i], 1);
    tc::fence_barrier_init();
  }
  for (int i = tid; i < p.F; i += F3_THREADS) sPar[i] = p.b1[i];
  float* sB2 = sPar + p.F; float* sOw = sB2 + 256; float* sOb = sOw + 256; float* sGw = sOb + 256;
  for (int i = tid; i < p.F; i += F3_THREADS) sGw[i] = p.gw1[i];
  for (int i = tid; i < 256; i += F3_THREADS) {
    sB2[i] = i < D ? p.b2[i] : 0.0f;
    sOw[i] = (OLN && i < D) ? p.oln_w[i] : 1.0f;
    sOb[i] = (OLN && i < D) ? p.oln_b[i] : 0.0f;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (CL2) tc::cluster_sync();  // the peer's barriers are initialised before any remote arrive / multicast commit
  tc::tc_fence_after();
  tc::pdl_wait();  // the producer of x has completed (everything above touched only parameters)
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  const uint32_t t_acc2 = tmem, t_acc1 = tmem + 256;  // acc1 buffers at +256 and +384; the bf16 chunk H[b] overlays acc1[b] (16 of every 32 columns)
  // tile walk: CTA (or pair) g takes tile groups g, g + n_groups, ...; in a pair, rank r takes the r-th tile of the group
  // (both CTAs run the same number of iterations: the protocol is collective; a tile past the end has nrows <= 0)
  const int first_base = CL2 ? (int)tc::cluster_id_x() * 2 : (int)blockIdx.x;
  const int base_step = CL2 ? (int)tc::cluster_count_x() * 2 : (int)gridDim.x;
  // Weight ring: step-granular slots (one full and one empty barrier, one tcgen05.commit and -- single CTA -- one bulk copy
  // per step: the single-thread roles pay ~100+ cycles of latency for every mbarrier / commit operation, so their count
  // per step, not the bytes, sets the pace of the ring).  Steps per hidden chunk: ceil(nkbD / 2) GEMM1 steps of up to two
  // K-blocks (N = 128 rows each) and two GEMM2 steps of one K-block (N = D rows).  The v3 weight images are laid out in
  // step order (tc_ffn3_pack), so a step is one contiguous run of 8 KB blocks.
  const int ng1 = (nkbD + 1) / 2;
  const uint32_t slot_bytes = CL2 ? 16384u : 32768u;
  const int nslots = (int)((CL2 ? p.ring_bytes : F3_RING_BYTES) / slot_bytes);
  // Every CTA walks the hidden chunks in the same order: a row's result does not depend on which CTA / tile computes it
  // (bit-exact batch invariance).  A per-CTA rotation of the order, meant to spread L2 requests, measured no gain
  // (replicating the weight images showed there is no hot-line effect to avoid).
  auto chunk_of = [&](int j) { return j; };
  // barriers owned by the leader: an arrival from the peer CTA is a remote arrive
  auto arrive_leader = [&](uint64_t* bar) { if (CL2) tc::mbar_arrive_remote(bar, 0); else tc::mbar_arrive(bar); };
  // hand-overs of tensor-memory contents only (hidden chunks, the drained output accumulator): no memory fence
  auto arrive_leader_tmem = [&](uint64_t* bar) { if (CL2) tc::mbar_arrive_remote_relaxed(bar, 0); else tc::mbar_arrive(bar); };

  if (warp == F3_PROD_WARP) {
    // =============================== weight producer ===============================
    // ring order == issue order: W1[0], W1[1], W2[0], W1[2], W2[1], ..., W2[nj-1]
    if (lane == 0) {
      int s = 0;
      uint32_t pe = 0;
      long long tw_empty = 0, t_tile0 = 0;
      const bool tracing = p.trace != nullptr && blockIdx.x == 0;
      // bytes: what the step's full barrier has to see -- in a pair BOTH halves, counted on the leader's barrier only (the peer's
      // copies complete there; a copy that lands before the leader has armed the phase only drives the transaction count negative)
      auto next_slot = [&](uint32_t bytes) -> uint8_t* {  // wait until the slot is free, arm its full barrier
        const long long c0t = tracing ? clock64() : 0;
        tc::mbar_wait(&empty_bar[s], ((pe >> s) & 1u) ^ 1u);  // suspending wait: a polling producer floods the SM sub-partition's shared-memory queue
        pe ^= 1u << s;
        if (tracing) tw_empty += clock64() - c0t;
        if (leader) tc::mbar_arrive_expect_tx(&full_bar[s], bytes);
        return sRing + (size_t)s * slot_bytes;
      };
      const uint32_t full0 = CL2 ? f3_leader_addr(&full_bar[0]) : 0u;  // the leader's full_bar[0] in the cluster window
      auto load_g1 = [&](int j) {
        const int c = chunk_of(j);
        for (int h = 0; h < ng1; ++h) {
          const int nk = nkbD - 2 * h < 2 ? nkbD - 2 * h : 2;
          const uint8_t* src = p.w1 + (size_t)((c * nkbD + 2 * h) * 2) * F3_BLOCK;  // blocks [c][kb][u]
          if (!CL2) {
            uint8_t* dst = next_slot((uint32_t)nk * 2u * F3_BLOCK);
            tc::bulk_g2s(dst, src, (uint32_t)nk * 2u * F3_BLOCK, &full_bar[s]);
          } else {  // this CTA's 64 of the 128 rows of every K-block
            const uint32_t dst = tc::smem_u32(next_slot((uint32_t)nk * 2u * F3_BLOCK));
            for (int kbl = 0; kbl < nk; ++kbl)
              f3_tma_block_pair(dst + (uint32_t)kbl * F3_BLOCK, &tmap_w1, (c * nkbD + 2 * h + kbl) * 2 + (int)crank, full0 + 8u * (uint32_t)s);
          }
          if (++s == nslots) s = 0;
        }
      };
      auto load_g2 = [&](int j) {
        const int c = chunk_of(j);
        for (int u = 0; u < 2; ++u) {
          const uint8_t* src = p.w2 + (size_t)((2 * c + u) * nkbD) * F3_BLOCK;  // blocks [kb of F][n-chunk]
          const uint32_t bytes = (uint32_t)nkbD * F3_BLOCK;
          uint8_t* dst = next_slot(bytes);
          if (!CL2) tc::bulk_g2s(dst, src, bytes, &full_bar[s]);
          else {  // this CTA's half of the D rows
            const int nb = nkbD / 2;
            for (int b = 0; b < nb; ++b)
              f3_tma_block_pair(tc::smem_u32(dst) + (uint32_t)b * F3_BLOCK, &tmap_w2, (2 * c + u) * nkbD + (int)crank * nb + b, full0 + 8u * (uint32_t)s);
          }
          if (++s == nslots) s = 0;
        }
      };
      int itp = 0;
      for (int base = first_base; base < p.n_tiles; base += base_step, ++itp) {
        if (tracing) t_tile0 = clock64();
        load_g1(0);
        for (int j = 0; j < nj; ++j) {
          if (j + 1 < nj) load_g1(j + 1);
          load_g2(j);
        }
        if (tracing && itp < 2) {  // producer: cycles blocked on empty slots / total cycles for this tile's loads
          p.trace[((0 * 2 + itp) * 32) + 16] = (unsigned long long)tw_empty;
          p.trace[((0 * 2 + itp) * 32) + 17] = (unsigned long long)(clock64() - t_tile0);
        }
        tw_empty = 0;
      }
    }
  } else if (warp == F3_MMA_WARP) {
    // =============================== MMA issuer (leader) / "my half has landed" relay (peer) ===============================
    // Issue order G1(0), G1(1), G2(0), G1(2), G2(1), ...: the tensor pipe executes in issue order, so G1(j+2), which
    // overwrites acc1[j&1] (and with it H[j&1]), runs after G2(j) has read H[j&1]; G2(j) itself is issued only after
    // the epilogue has loaded acc1[j&1] and stored H[j&1] (h_full).  No separate "accumulator drained" barrier.
    if (CL2 && !leader) {
      // Peer CTA: this warp has no MMAs to issue (and nothing to relay: the peer's weight copies complete on the leader's barriers)
    } else {
    int s = 0;
    uint32_t pf = 0;
    uint32_t ph_hf = 0;  // per-buffer parity of h_full
    const uint32_t x0 = tc::smem_u32(sX), r0 = tc::smem_u32(sRing);
    const uint32_t idesc1 = tc::make_idesc_bf16(128 * NCTA, F3_HC), idesc2 = tc::make_idesc_bf16(128 * NCTA, (uint32_t)D);
    const uint32_t g1_kb_bytes = 2u * F3_BLOCK / NCTA;  // one K-block of a GEMM1 step in this CTA's slot (128 or 64 rows)
    int it = 0;
    long long tw_ring = 0, tw_h = 0, tw_epi = 0;  // cycles this warp waited for weights / hidden chunks / the drained accumulator
    const bool tracing = p.trace != nullptr && blockIdx.x == 0;
    auto ring_next = [&]() -> uint32_t {  // wait for the next step (both halves); returns its shared-memory address
      const long long c0 = tracing ? clock64() : 0;
      tc::mbar_wait_spin(&full_bar[s], (pf >> s) & 1u);
      if (tracing) tw_ring += clock64() - c0;
      pf ^= 1u << s;
      tc::tc_fence_after();
      return r0 + (uint32_t)s * slot_bytes;
    };
    auto release = [&]() {  // one commit per step: the slot is free when the MMAs issued so far have completed
      if (leader && tc::elect_one()) { if (CL2) tc::umma2_commit(&empty_bar[s]); else tc::umma_commit(&empty_bar[s]); }
      __syncwarp();
      if (++s == nslots) s = 0;
    };
    auto gemm1 = [&](int j) {  // acc1[j&1] = LN(x) @ W1[chunk j]^T
      const int bsel = j & 1;
      for (int h = 0; h < ng1; ++h) {
        const int nk = nkbD - 2 * h < 2 ? nkbD - 2 * h : 2;
        const uint32_t b_addr = ring_next();
        if (leader && tc::elect_one()) {
          for (int kbl = 0; kbl < nk; ++kbl) {
            const uint32_t a_addr = x0 + (uint32_t)(2 * h + kbl) * kblock_bytes(128), bk = b_addr + (uint32_t)kbl * g1_kb_bytes;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t acc = (h == 0 && kbl == 0 && ks == 0) ? 0u : 1u;
              if (CL2) tc::umma2_bf16(t_acc1 + bsel * F3_HC, tc::make_desc_sw128(a_addr + ks * 32), tc::make_desc_sw128(bk + ks * 32), idesc1, acc);
              else tc::umma_bf16(t_acc1 + bsel * F3_HC, tc::make_desc_sw128(a_addr + ks * 32), tc::make_desc_sw128(bk + ks * 32), idesc1, acc);
            }
          }
        }
        __syncwarp();
        release();
      }
      if (leader && tc::elect_one()) {
        if (CL2) { tc::umma2_commit(&acc1_full[bsel]); if (j == nj - 1) tc::umma2_commit(&x_free); }
        else { tc::umma_commit(&acc1_full[bsel]); if (j == nj - 1) tc::umma_commit(&x_free); }
      }
      __syncwarp();
    };
    for (int base = first_base; base < p.n_tiles; base += base_step, ++it) {
      const uint32_t par = it & 1;
      if (leader) { if (CL2) tc::mbar_wait_cluster(&x_full, par); else tc::mbar_wait(&x_full, par); }
      tc::tc_fence_after();
      F3_TRACE(1, it, 0);
      gemm1(0);
      for (int j = 0; j < nj; ++j) {
        if (j + 1 < nj) gemm1(j + 1);
        const int bsel = j & 1;
        if (leader) {
          long long c0 = tracing ? clock64() : 0;
          if (j == 0 && it > 0) { if (CL2) tc::mbar_wait_cluster(&epi_done, par ^ 1); else tc::mbar_wait(&epi_done, par ^ 1); }  // acc2 drained
          if (tracing) { const long long c1 = clock64(); tw_epi += c1 - c0; c0 = c1; }
          tc::mbar_wait_spin(&h_full[bsel], (ph_hf >> bsel) & 1u);  // (tensor-memory payload: no cluster-scope acquire, which costs an L1 invalidate per wait)
          if (tracing) tw_h += clock64() - c0;
          ph_hf ^= 1u << bsel;
        }
        tc::tc_fence_after();
        const uint32_t a_tmem = t_acc1 + bsel * F3_HC;  // H[bsel]: K = 128 bf16 in 4 x 16 columns (the first half of every epilogue warp's 32)
        for (int u = 0; u < 2; ++u) {  // acc2 += H[j][:, K-block u] @ W2[:, K-block 2j+u]^T
          const uint32_t b_addr = ring_next();
          if (leader && tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t acc = (j == 0 && u == 0 && ks == 0) ? 0u : 1u;
              if (CL2) tc::umma2_bf16_ts(t_acc2, a_tmem + (u * 2 + (ks >> 1)) * 32 + (ks & 1) * 8, tc::make_desc_sw128(b_addr + ks * 32), idesc2, acc);
              else f3_umma_ts(t_acc2, a_tmem + (u * 2 + (ks >> 1)) * 32 + (ks & 1) * 8, tc::make_desc_sw128(b_addr + ks * 32), idesc2, acc);
            }
          }
          __syncwarp();
          release();
        }
        if (j == nj - 1) {
          if (leader && tc::elect_one()) { if (CL2) tc::umma2_commit(&acc2_full); else tc::umma_commit(&acc2_full); }
          __syncwarp();
        }
        if (j < 4) F3_TRACE(1, it, 1 + j);
      }
      F3_TRACE(1, it, 8);
      if (tracing && lane == 0 && it < 2) {
        p.trace[((1 * 2 + it) * 32) + 16] = (unsigned long long)tw_ring;
        p.trace[((1 * 2 + it) * 32) + 17] = (unsigned long long)tw_h;
        p.trace[((1 * 2 + it) * 32) + 18] = (unsigned long long)tw_epi;
      }
      tw_ring = tw_h = tw_epi = 0;
    }
    }
  } else if (warp >= F3_PRO_WARP0) {
    // =============================== prologue: x tile -> A operand image, row statistics ===============================
    // The input LayerNorm is NOT applied to the tile: gamma is folded into the packed W1, beta into b1 (tc_ffn3_pack), and the
    // first epilogue applies the two per-row scalars -- LN(x) W1^T = rstd (x (W1 gamma)^T - mean gw1) + W1 beta.  GEMM 1 of a tile
    // therefore starts as soon as its rows have landed (under the final epilogue of the previous tile), and the raw bf16 rows
    // are exact GEMM operands (the normalised rows were rounded to bf16 once more).  The statistics (one thread per row, one
    // pass shifted by the row's first element) are only needed by the first epilogue, some 3 k cycles later.
    // The copy of tile t+1 starts as soon as the last GEMM1 of tile t has released X.
    const int pw = warp - F3_PRO_WARP0;
    int it = 0;
    for (int base = first_base; base < p.n_tiles; base += base_step, ++it) {
      const int64_t row0 = (int64_t)(base + (int)crank) * 128;
      const int nrows = p.rows - row0 < 128 ? (int)(p.rows - row0) : 128;
      if (it > 0) tc::mbar_wait(&x_free, (it - 1) & 1);
      if (pw == 0) F3_TRACE(2, it, 0);
      // the tile arrives by tensor-map TMA (one 64-column box per K-block, 128-byte swizzle: the operand image as is; rows past
      // the end are zero-filled) -- four bulk copies instead of 4096 16-byte cp.async competing with the epilogue warps for
      // issue slots and the load/store queue (the copy of a CTA's second tile took 9.7 k cycles that way)
      if (pw == 0 && lane == 0) {
        tc::mbar_arrive_expect_tx(&x_landed, (uint32_t)nkbD * kblock_bytes(128));
        for (int kb = 0; kb < nkbD; ++kb) f3_tma_load_2d(sX + (size_t)kb * kblock_bytes(128), &tmap_x, kb * 64, (int)row0, &x_landed);
      }
      tc::mbar_wait(&x_landed, it & 1);
      if (lane == 0) arrive_leader(&x_full);
      if (pw == 0) F3_TRACE(2, it, 1);
      float rs = 1.0f, nm = 0.0f;
      if (p.has_ln) {
        const int row = pw * 32 + lane;  // copied by this warp
        const uint8_t* rp = sX + row * 128;
        const float x0 = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(rp + ((row & 7) << 4)));  // element 0 (chunk 0 lies at chunk position row & 7)
        float2 s1 = make_float2(0.0f, 0.0f), s2 = make_float2(0.0f, 0.0f);
        const float2 sh = make_float2(-x0, -x0);
#pragma unroll 1
        for (int kb = 0; kb < nkbD; ++kb) {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint4 raw = *reinterpret_cast<const uint4*>(rp + (size_t)kb * kblock_bytes(128) + ((ch ^ (row & 7)) << 4));
            float2 v[4];
            tc::unpack_bf16x8_pairs(raw, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 d = tc::add2(v[e], sh); s1 = tc::add2(s1, d); s2 = tc::fma2(d, d, s2); }
          }
        }
        const float inv = 1.0f / (float)D;
        const float m1 = (s1.x + s1.y) * inv;
        const float var = fmaxf((s2.x + s2.y) * inv - m1 * m1, 0.0f);
        rs = rsqrtf(var + 1e-5f);
        nm = -(x0 + m1) * rs;
      }
      sStat[(it & 1) * 128 + pw * 32 + lane] = make_float2(rs, nm);
      tc::mbar_arrive(&stat_full);  // every thread publishes its own row (release / acquire through the barrier)
    }
  } else {
    // =============================== epilogue ===============================
    const int q = warp & 3, k = warp >> 2;   // TMEM lane quadrant; column quarter (32 hidden columns / 64 output columns)
    const int r = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    uint32_t ph_a1f = 0;
    int it = 0;
    for (int base = first_base; base < p.n_tiles; base += base_step, ++it) {
      const uint32_t par = it & 1;
      const int64_t row0 = (int64_t)(base + (int)crank) * 128;
      const int nrows = p.rows - row0 < 128 ? (int)(p.rows - row0) : 128;
      // Pair mode: the final epilogue goes through a staging tile in shared memory (the half of the ring region the pair does not
      // need: each CTA streams half of every weight step).  The residual's 64-column block k is fetched into it by tensor-map TMA,
      // the four warps of column quarter k add the accumulator in place and one of them stores the block with a bulk tensor store
      // (rows past the end are clipped): LDS/STS.128 that are conflict-free under the 128-byte swizzle and whole lines on the
      // way out, instead of per-thread row accesses with lanes 512 bytes apart.  Block k's owner (warp 4k, lane 0) issues the
      // loads: for the first tile at once, for a later tile after the first hidden chunks (by then the previous tile's store has
      // long read the block: the wait costs nothing).
      uint8_t* const sStage = sRing + 65536;
      const bool stage_owner = CL2 && q == 0 && lane == 0 && k * 64 < D;
      auto fetch_residual = [&]() {
        tc::bulk_wait_read0();
        tc::mbar_arrive_expect_tx(&res_full[k], kblock_bytes(128));
        f3_tma_load_2d(sStage + (size_t)k * kblock_bytes(128), &tmap_x, k * 64, (int)row0, &res_full[k]);
      };
      if (stage_owner && it == 0) fetch_residual();
      // this row's LayerNorm scalars (the prologue warps computed them while GEMM 1 of the first chunks ran)
      tc::mbar_wait(&stat_full, par);
      const float2 rst = sStat[par * 128 + r];
      const float rs = rst.x, nm = rst.y;
      for (int j = 0; j < nj; ++j) {
        const int bsel = j & 1;
        tc::mbar_wait(&acc1_full[bsel], (ph_a1f >> bsel) & 1u);
        ph_a1f ^= 1u << bsel;
        tc::tc_fence_after();
        if (warp == 0 && j < 4) F3_TRACE(3, it, 2 * j);
        float v[32];
        tc::tmem_ld32(t_acc1 + lane_sel + bsel * F3_HC + k * 32, v);
        tc::tmem_ld_wait();
        // H[bsel] overlays acc1[bsel]: this warp writes its 32 activations (16 packed columns) over the FIRST HALF OF ITS OWN 32
        // accumulator columns, which it has just loaded -- no other warp reads or writes them, so no barrier between the warps of
        // a lane quadrant is needed.  GEMM 2 takes K-slice s (16 hidden units, 8 columns) from columns 32 (s / 2) + 8 (s % 2).
        const float4* bp = reinterpret_cast<const float4*>(sPar + chunk_of(j) * F3_HC + k * 32);
        const float4* gp = reinterpret_cast<const float4*>(sGw + chunk_of(j) * F3_HC + k * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // LN(x) W1^T + b1 = rstd acc + (-mean rstd) gw1 + b1'
          const float4 bb = bp[i], gg = gp[i];
          v[4 * i] = fmaf(rs, v[4 * i], fmaf(nm, gg.x, bb.x)); v[4 * i + 1] = fmaf(rs, v[4 * i + 1], fmaf(nm, gg.y, bb.y));
          v[4 * i + 2] = fmaf(rs, v[4 * i + 2], fmaf(nm, gg.z, bb.z)); v[4 * i + 3] = fmaf(rs, v[4 * i + 3], fmaf(nm, gg.w, bb.w));
        }
        tc::act_apply<32>(act, v);
        uint32_t hp[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) hp[i] = tc::pack_bf16x2(v[2 * i], v[2 * i + 1]);
        f3_tmem_st16(t_acc1 + lane_sel + bsel * F3_HC + k * 32, hp);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader_tmem(&h_full[bsel]);
        if (warp == 0 && j < 4) F3_TRACE(3, it, 2 * j + 1);
        if (stage_owner && it > 0 && j == 1) fetch_residual();
      }
      // ---- final: y = x + 0.5*(acc2 + b2)  [-> LN_out]; this thread: row r, output columns [64k, 64k + 64), in four
      // pieces of 16 columns (one 256-bit residual load and one 256-bit store each)
      const bool active = k * 64 < D;
      const bool live = r < nrows;
      uint32_t rres[CL2 ? 1 : 32];  // the residual (bf16 pairs), fetched while the last GEMMs run (pair mode: read from the staging tile)
      uint8_t* const srow = sStage + (size_t)k * kblock_bytes(128) + r * 128;  // this thread's 128 bytes of staging block k
      if (!CL2 && active) {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          if (live) f3_ldg256(p.x + (row0 + r) * D + k * 64 + h * 16, rres + 8 * h);
          else {
#pragma unroll
            for (int e = 0; e < 8; ++e) rres[8 * h + e] = 0u;
          }
        }
      }
      if (CL2 && active) tc::mbar_wait(&res_full[k], par);
      tc::mbar_wait(&acc2_full, par);
      tc::tc_fence_after();
      if (warp == 0) F3_TRACE(3, it, 10);
      float mean_t = 0.0f, m2_t = 0.0f;
      if (active) {
        float vbuf[2][16];  // the next piece's tcgen05.ld is in flight while this piece is processed
        f3_tmem_ld16(t_acc2 + lane_sel + k * 64, vbuf[0]);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int col = k * 64 + h * 16;
          tc::tmem_ld_wait();
          if (h + 1 < 4) f3_tmem_ld16(t_acc2 + lane_sel + col + 16, vbuf[(h + 1) & 1]);
          float* v = vbuf[h & 1];
          const float4* bp = reinterpret_cast<const float4*>(sB2 + col);
          uint32_t rr[8];
          if (CL2) {
            const uint4 r0 = *reinterpret_cast<const uint4*>(srow + (((2 * h) ^ (r & 7)) << 4));
            const uint4 r1 = *reinterpret_cast<const uint4*>(srow + (((2 * h + 1) ^ (r & 7)) << 4));
            rr[0] = r0.x; rr[1] = r0.y; rr[2] = r0.z; rr[3] = r0.w; rr[4] = r1.x; rr[5] = r1.y; rr[6] = r1.z; rr[7] = r1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) rr[e] = rres[(CL2 ? 0 : h * 8) + (CL2 ? 0 : e)];
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 bb = bp[i];
            const float2 ra = f3_bf2(rr[2 * i]), rb = f3_bf2(rr[2 * i + 1]);
            v[4 * i] = fmaf(0.5f, v[4 * i] + bb.x, ra.x);
            v[4 * i + 1] = fmaf(0.5f, v[4 * i + 1] + bb.y, ra.y);
            v[4 * i + 2] = fmaf(0.5f, v[4 * i + 2] + bb.z, rb.x);
            v[4 * i + 3] = fmaf(0.5f, v[4 * i + 3] + bb.w, rb.y);
          }
          if (OLN) {
            // running (mean, M2) of this thread's 64 values: exact two-pass inside a piece, Chan's merge across pieces
            float sa[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // four independent chains (fixed association: deterministic)
#pragma unroll
            for (int e = 0; e < 16; ++e) sa[e & 3] += v[e];
            const float mh = ((sa[0] + sa[1]) + (sa[2] + sa[3])) * (1.0f / 16.0f);
            float qa[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int e = 0; e < 16; ++e) { const float d = v[e] - mh; qa[e & 3] = fmaf(d, d, qa[e & 3]); }
            const float qh = (qa[0] + qa[1]) + (qa[2] + qa[3]);
            if (h == 0) { mean_t = mh; m2_t = qh; }
            else {
              const float dl = mh - mean_t, na = 16.0f * h;
              mean_t += dl * (16.0f / (na + 16.0f));
              m2_t += qh + dl * dl * (na * 16.0f / (na + 16.0f));
            }
            f3_tmem_st16f(t_acc2 + lane_sel + col, v);  // park the pre-norm values for the normalisation pass
          } else if (CL2 || live) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = tc::pack_bf16x2(v[2 * i], v[2 * i + 1]);
            if (CL2) {
              *reinterpret_cast<uint4*>(srow + (((2 * h) ^ (r & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(srow + (((2 * h + 1) ^ (r & 7)) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
            } else f3_stg256(p.y + (row0 + r) * D + col, o);
          }
        }
      }
      if (warp == 0) F3_TRACE(3, it, 12);
      if (OLN) {
        tc::tmem_st_wait();
        if (active) { sRed[(k * 128 + r) * 2] = mean_t; sRed[(k * 128 + r) * 2 + 1] = m2_t; }
        tc::named_bar_sync(1 + q, 128);
        const int nq = D / 64;  // column quarters in use (64 values each)
        float mean = 0.0f;
        for (int i = 0; i < nq; ++i) mean += sRed[(i * 128 + r) * 2];
        mean /= (float)nq;
        float m2 = 0.0f;
        for (int i = 0; i < nq; ++i) { const float dl = sRed[(i * 128 + r) * 2] - mean; m2 += sRed[(i * 128 + r) * 2 + 1] + 64.0f * dl * dl; }
        const float rstd = rsqrtf(m2 / (float)D + p.oln_eps);
        const float shift = -mean * rstd;
        if (warp == 0) F3_TRACE(3, it, 13);
        if (active) {
#pragma unroll 1
          for (int h = 0; h < 4; ++h) {
            const int col = k * 64 + h * 16;
            float v[16];
            f3_tmem_ld16(t_acc2 + lane_sel + col, v);
            tc::tmem_ld_wait();
            const float4* wp = reinterpret_cast<const float4*>(sOw + col);
            const float4* bp = reinterpret_cast<const float4*>(sOb + col);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 ww = wp[i], bb = bp[i];
              v[4 * i] = fmaf(fmaf(v[4 * i], rstd, shift), ww.x, bb.x);
              v[4 * i + 1] = fmaf(fmaf(v[4 * i + 1], rstd, shift), ww.y, bb.y);
              v[4 * i + 2] = fmaf(fmaf(v[4 * i + 2], rstd, shift), ww.z, bb.z);
              v[4 * i + 3] = fmaf(fmaf(v[4 * i + 3], rstd, shift), ww.w, bb.w);
            }
            if (CL2 || live) {
              uint32_t o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = tc::pack_bf16x2(v[2 * i], v[2 * i + 1]);
              if (CL2) {
                *reinterpret_cast<uint4*>(srow + (((2 * h) ^ (r & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(srow + (((2 * h + 1) ^ (r & 7)) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
              } else f3_stg256(p.y + (row0 + r) * D + col, o);
            }
          }
        }
        tc::named_bar_sync(1 + q, 128);  // sRed is rewritten by the next tile
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader_tmem(&epi_done);
      if (CL2 && active) {  // block k of the output tile is complete in the staging tile: one bulk tensor store
        tc::fence_proxy_async();
        tc::named_bar_sync(5 + k, 128);
        if (stage_owner) { f3_tma_store_2d(&tmap_y, tc::smem_u32(sStage + (size_t)k * kblock_bytes(128)), k * 64, (int)row0); tc::bulk_commit(); }
      }
      if (warp == 0) F3_TRACE(3, it, 11);
    }
  }
  if (CL2 && warp < F3_NEW && (warp & 3) == 0 && lane == 0) tc::bulk_wait_read0();  // the stores have read the staging tile before the CTA retires
  tc::tc_fence_before();
  __syncthreads();
  if (CL2) tc::cluster_sync();  // no CTA leaves (or frees tensor memory) while the pair's MMAs, multicast commits or remote arrives may still land
  if (warp == F3_PROD_WARP) { if (CL2) tc::tmem_dealloc2(tmem, 512); else tc::tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------
// host side (weights packed by tc_ffn2_pack: the images are shared with v2)
// ---------------------------------------------------------------------------------------------
static std::atomic<int> g_ffn_ver{3};   // smx_debug_set_ffn_version: 2 = shared-memory hidden (v2), 3 = TMEM hidden, single CTAs,
static std::atomic<int> g_ffn_pair{1};  //                            4 = TMEM hidden + CTA pairs (cta_group::2; bit-identical to 3; default where it applies)
void tc_set_ffn_version(int v) { g_ffn_ver = v == 2 ? 2 : 3; g_ffn_pair = v == 4 ? 1 : 0; }  // 4 also restores the default
int tc_ffn_version() { return g_ffn_ver; }

// v3 images = the v2 images with their 8 KB blocks reordered into ring-step order:
//   W1: [hidden chunk c][K-block kb][row half u]  <-  v2 block ((2c + u) * nkbD + kb)
//   W2: [K-block kb of F][64-row chunk n of D]    <-  v2 block (n * nkbF + kb)
__global__ void ffn3_reorder_kernel(const uint4* w1v2, const uint4* w2v2, uint4* w1v3, uint4* w2v3, int nkbD, int nkbF) {
  const int blk = blockIdx.x, n1 = nkbD * nkbF;  // blocks per image
  const uint4* src; uint4* dst;
  if (blk < n1) {
    const int u = blk & 1, kb = (blk >> 1) % nkbD, c = (blk >> 1) / nkbD;
    src = w1v2 + (size_t)((2 * c + u) * nkbD + kb) * 512; dst = w1v3 + (size_t)blk * 512;
  } else {
    const int b2 = blk - n1, n = b2 % nkbD, kb = b2 / nkbD;
    src = w2v2 + (size_t)(n * nkbF + kb) * 512; dst = w2v3 + (size_t)b2 * 512;
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) dst[i] = src[i];
}
// image: [v2 W1 | v2 W2 (the v2 kernel's)] [v3 W1, LayerNorm folded | v3 W2 (step order)] [gw1 f32 F] [b1' f32 F]
//        [pack-time scratch: v2-order image of the folded W1 | W1 gamma in fp32]
struct Ffn3Image { size_t img, gw, b1, s_img, s_wg, total; };
static Ffn3Image ffn3_image(const smx_ffn_weights* w) {
  Ffn3Image im{};
  const size_t D = w->w1.in_dim, F = w->w1.out_dim;
  im.img = align_up(D * F * 2, 1024);
  size_t off = 4 * im.img;
  im.gw = off; off += align_up(F * 4, 1024);
  im.b1 = off; off += align_up(F * 4, 1024);
  im.s_img = off; off += im.img;
  im.s_wg = off; off += align_up(D * F * 4, 1024);
  im.total = off;
  return im;
}
__global__ void ffn3_bias_kernel(const float* __restrict__ b1, const float* __restrict__ bw, float* __restrict__ out, int F) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < F) out[i] = (b1 ? b1[i] : 0.0f) + bw[i];
}
size_t tc_ffn3_packed_bytes(const smx_ffn_weights* w) { return ffn3_image(w).total; }
int tc_ffn3_pack(const smx_ffn_weights* w, void* packed, cudaStream_t st) {
  SMX_TRY(tc_ffn2_pack(w, packed, st));
  const int D = w->w1.in_dim, F = w->w1.out_dim, nkbD = D / 64, nkbF = F / 64;
  const Ffn3Image im = ffn3_image(w);
  uint8_t* b = (uint8_t*)packed;
  // the input LayerNorm folded into the first layer (see the kernel's prologue): W1 gamma -> image, gw1, b1' = b1 + W1 beta
  float* Wg = (float*)(b + im.s_wg);
  float* gw = (float*)(b + im.gw);
  float* b1f = (float*)(b + im.b1);
  SMX_TRY(tc_fold_ln(w->w1.w, D, D, F, w->ln_w, w->ln_b, Wg, gw, b1f, st));  // (b1f holds W1 beta for a moment)
  ffn3_bias_kernel<<<(F + 255) / 256, 256, 0, st>>>(w->w1.b, b1f, b1f, F);
  count_launch();
  SMX_TRY(check_launch("ffn3_bias_kernel"));
  smx_linear Lg{};
  Lg.w = Wg; Lg.b = nullptr; Lg.in_dim = D; Lg.out_dim = F; Lg.n_split = 1;
  SMX_TRY(tc_pack_linear_nt(Lg, 0, D, 64, b + im.s_img, st));
  ffn3_reorder_kernel<<<2 * nkbD * nkbF, 128, 0, st>>>((const uint4*)(b + im.s_img), (const uint4*)(b + im.img), (uint4*)(b + 2 * im.img),
                                                      (uint4*)(b + 3 * im.img), nkbD, nkbF);
  count_launch();
  return check_launch("ffn3_reorder_kernel");
}

bool tc_ffn3_supported(const smx_ffn_weights* w) {
  if (!tc_ffn2_supported(w)) return false;
  const int D = w->w1.in_dim;
  return D % 64 == 0 && D <= 256;
}

static std::atomic<unsigned long long*> g_trace3f{nullptr};
void tc_set_trace_ffn3(void* p) { g_trace3f = (unsigned long long*)p; }

static int ffn3_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <bool OLN, bool CL2>
static int launch_ffn3(const CUtensorMap& tm, const CUtensorMap& tw1, const CUtensorMap& tw2, const CUtensorMap& ty, const Ffn3P& p, unsigned grid, size_t smem, cudaStream_t st) {
  cudaError_t e;
#define SMX_FFN3_LAUNCH(A)                                                                                   \
  e = cudaFuncSetAttribute(ffn3_kernel<OLN, A, CL2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(ffn3_kernel): %s", cudaGetErrorString(e)); \
  e = launch_pdl(ffn3_kernel<OLN, A, CL2>, dim3(grid), dim3(F3_THREADS), smem, st, CL2 ? 2u : 1u, tm, tw1, tw2, ty, p);     \
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaLaunchKernelEx(ffn3_kernel): %s", cudaGetErrorString(e));
  switch (p.act) {
    case SMX_ACT_SWISH: SMX_FFN3_LAUNCH(SMX_ACT_SWISH); break;
    case SMX_ACT_GELU: SMX_FFN3_LAUNCH(SMX_ACT_GELU); break;
    case SMX_ACT_RELU: SMX_FFN3_LAUNCH(SMX_ACT_RELU); break;
    default: SMX_FFN3_LAUNCH(-1); break;
  }
#undef SMX_FFN3_LAUNCH
  count_tc_launch();
  return check_launch("ffn3_kernel");
}

int tc_ffn3_fwd(const smx_ffn_weights* w, const void* packed, int act, int64_t rows, const __nv_bfloat16* x,
                const float* oln_w, const float* oln_b, float oln_eps, __nv_bfloat16* y, cudaStream_t st) {
  const int D = w->w1.in_dim, F = w->w1.out_dim;
  Ffn3P p{};
  p.x = x; p.y = y; p.rows = rows; p.D = D; p.F = F;
  p.n_tiles = (int)((rows + 127) / 128);
  p.w1 = (const uint8_t*)packed + 2 * align_up((size_t)D * F * 2, 1024);  // the v3 images follow the v2 images
  p.w2 = p.w1 + align_up((size_t)D * F * 2, 1024);
  const Ffn3Image im = ffn3_image(w);
  p.b1 = (const float*)((const uint8_t*)packed + im.b1); p.gw1 = (const float*)((const uint8_t*)packed + im.gw); p.b2 = w->w2.b;
  p.has_ln = w->ln_w != nullptr ? 1 : 0;
  p.oln_w = oln_w; p.oln_b = oln_b; p.oln_eps = oln_eps;
  p.act = act;
  const int nc = D / 64;
  p.trace = g_trace3f;
  p.off_ring = (uint32_t)nc * kblock_bytes(128);
  p.off_par = p.off_ring + F3_RING_BYTES;
  p.off_red = p.off_par + (uint32_t)align_up((size_t)(2 * F + 768 + 32) * 4, 1024);
  const size_t smem = (size_t)p.off_red + 4096 + 2048;
  if (smem > 227 * 1024 - 1024) return fail(SMX_ERR_UNSUPPORTED, "ffn: tile does not fit shared memory");
  CUtensorMap tm;
  {
    const uint64_t dims[2] = {(uint64_t)D, (uint64_t)rows}, strides[1] = {(uint64_t)D * 2};
    const uint32_t box[2] = {64, 128};
    if (!tc_encode_tmap_bf16(&tm, x, 2, dims, strides, box)) return fail(SMX_ERR_CUDA, "ffn: cuTensorMapEncodeTiled failed");
  }
  // CTA pairs when the weight steps split evenly over two CTAs (D a multiple of 128) and there is more than one tile
  p.cl2 = (g_ffn_pair && D % 128 == 0 && p.n_tiles >= 2) ? 1 : 0;
  p.ring_bytes = 65536u;
  CUtensorMap tw1 = tm, tw2 = tm;  // weight images as block tensors (pair mode only: the peer's copies complete on the leader's barriers)
  CUtensorMap ty = tm;
  if (p.cl2) {
    const uint64_t dims[2] = {(uint64_t)D, (uint64_t)rows}, strides[1] = {(uint64_t)D * 2};
    const uint32_t box[2] = {64, 128};
    if (!tc_encode_tmap_bf16(&ty, y, 2, dims, strides, box)) return fail(SMX_ERR_CUDA, "ffn: cuTensorMapEncodeTiled (y) failed");
  }
  if (p.cl2 && !(tc_encode_tmap_image(&tw1, p.w1, (uint64_t)D * F / 4096) && tc_encode_tmap_image(&tw2, p.w2, (uint64_t)D * F / 4096)))
    return fail(SMX_ERR_CUDA, "ffn: cuTensorMapEncodeTiled (weight images) failed");
  if (p.cl2) {
    const int n_pairs = (p.n_tiles + 1) / 2, max_pairs = ffn3_sms() / 2;
    const unsigned grid = 2u * (unsigned)(n_pairs < max_pairs ? n_pairs : max_pairs);
    return oln_w ? launch_ffn3<true, true>(tm, tw1, tw2, ty, p, grid, smem, st) : launch_ffn3<false, true>(tm, tw1, tw2, ty, p, grid, smem, st);
  }
  const unsigned grid = (unsigned)(p.n_tiles < ffn3_sms() ? p.n_tiles : ffn3_sms());
  return oln_w ? launch_ffn3<true, false>(tm, tw1, tw2, ty, p, grid, smem, st) : launch_ffn3<false, false>(tm, tw1, tw2, ty, p, grid, smem, st);
}

}  // namespace smx
