// sm_100a building blocks for the tcgen05 arm of libsmx: mbarrier, bulk async copy (TMA engine),
// tcgen05.mma / TMEM wrappers and the shared-memory operand layouts.  Inline PTX only.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace smx {
namespace tc {

// ---------------------------------------------------------------------------------------------
// Operand layout: K-major, 128-byte swizzle (UMMA LayoutType::SWIZZLE_128B).
// An operand "K-block" holds R rows x 64 bf16 (128 bytes per row).  Rows are grouped by 8 into
// 1024-byte atoms; inside an atom the 16-byte chunk c of row r is stored at chunk (c ^ (r & 7)).
// The same image is produced by the weight packer (global memory, copied verbatim by cp.async.bulk)
// and by the kernels that write activation operands from registers.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t chunk16) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((chunk16 ^ (r & 7u)) << 4);
}
constexpr uint32_t KBLOCK_COLS = 64;                              // bf16 elements per K-block row
__host__ __device__ constexpr uint32_t kblock_bytes(uint32_t rows) { return rows * 128u; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- programmatic dependent launch --------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// make generic-proxy writes to shared memory visible to the async proxy (TMA engine, tcgen05.mma)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.  The retry loop lives inside one
// asm block so that the surrounding C++ control flow stays warp-uniform for the compiler (the MMA issuer's address
// arithmetic then stays in uniform registers).  try_wait suspends the thread in hardware for a bounded time per
// attempt (~0.5 us observed), so 2^23 attempts bound the wait to a few seconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 n;\n\t"
      "mov.u32 n, 0;\n\t"
      "MBAR_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MBAR_DONE_%=;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 p, n, 8388608;\n\t"
      "@p bra MBAR_WAIT_%=;\n\t"
      "trap;\n\t"
      "MBAR_DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// The same without hardware suspension: test_wait polls.  For the single-thread roles on the critical path (weight
// producer, MMA issuer) whose wake-up latency out of a suspended try_wait would otherwise be paid once per ring step.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 n;\n\t"
      "mov.u32 n, 0;\n\t"
      "MBAR_SPIN_%=:\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MBAR_SPUN_%=;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 p, n, 268435456;\n\t"
      "@p bra MBAR_SPIN_%=;\n\t"
      "trap;\n\t"
      "MBAR_SPUN_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_spin_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 n;\n\t"
      "mov.u32 n, 0;\n\t"
      "MBAR_SPINC_%=:\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MBAR_SPUNC_%=;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 p, n, 268435456;\n\t"
      "@p bra MBAR_SPINC_%=;\n\t"
      "trap;\n\t"
      "MBAR_SPUNC_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- bulk async copies (TMA engine, 1-D) ---------------------------------------------------------
// global -> shared, completion signalled on an mbarrier (complete_tx).  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the same, delivered to the same CTA-relative offset (data and mbarrier) of every CTA in cta_mask of the cluster
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// ---- thread-block clusters ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {  // all threads of all CTAs of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared -> global (bulk group completion)
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- cp.async (LDGSTS): 16-byte global -> shared copies that bypass registers -----------------------
// src_bytes == 0 writes 16 zero bytes (nothing is read)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, uint32_t src_bytes = 16) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp as alloc
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// warp-collective: lane i of warp w receives 32 consecutive fp32 columns of TMEM lane 32*(w%4)+i
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// warp-collective store of 32 fp32 columns per lane (inverse of tmem_ld32)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- tcgen05.mma ----------------------------------------------------------------------------------
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address >> 4 in [0,14),
// leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48),
// layout type in [61,64) (0 none/interleave, 2 = 128B swizzle).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;  // SBO: 8 rows x 128 B between row groups
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}
// no-swizzle K-major: core matrices of 8 rows x 16 B; lbo = bytes between the two K chunks of one
// K=16 instruction, sbo = bytes between 8-row groups
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B bf16, both K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// the same, arriving on the barrier at this CTA-relative offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- CTA pairs (cta_group::2): one tcgen05.mma spans both SMs of a 2-CTA cluster -------------------------------
// M = 256 (128 rows per CTA, each CTA's own A operand and accumulator), the B operand is split: each CTA holds N/2 of its
// rows in its own shared memory at the same CTA-relative address.  Issued by one thread of the LEADER CTA (rank 0).
// Every tcgen05 alloc / mma / commit of such a kernel uses cta_group::2.
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {  // the same warp of BOTH CTAs, same smem_dst offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in tensor memory (each CTA's own 128 lanes, bf16 pairs per 32-bit column)
__device__ __forceinline__ void umma2_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this CTA-relative offset in both CTAs of the pair when all previously issued MMAs completed
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrive (release at cluster scope) on the barrier at this CTA-relative offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// The same arrival WITHOUT memory ordering: for hand-overs whose payload lives in tensor memory only (ordered by tcgen05.wait +
// tcgen05.fence::before_thread_sync on the arriving side).  The release form above is lowered to MEMBAR.ALL.GPU + ERRBAR per arrive.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// bounded wait with acquire at cluster scope (the arrivals may come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 n;\n\t"
      "mov.u32 n, 0;\n\t"
      "MBAR_WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MBAR_DONEC_%=;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 p, n, 8388608;\n\t"
      "@p bra MBAR_WAITC_%=;\n\t"
      "trap;\n\t"
      "MBAR_DONEC_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// named barrier among a subset of warps (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- activations on the fast path -------------------------------------------------------------------
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// act codes mirror smx_act.  swish/sigmoid use one MUFU (tanh.approx, |rel err| ~ 2^-11, below the bf16
// rounding the result receives); gelu(erf) keeps erff for parity with nn.GELU().
__device__ __forceinline__ float act_swish(float x) { float h = 0.5f * x; return fmaf(h, tanh_approx(h), h); }
__device__ __forceinline__ float act_sigmoid(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
// exact (erf) GELU of nn.GELU(): x * Phi(x) with erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7 on erf: three orders of magnitude
// below the bf16 rounding the result receives), branch-free: one rcp, one ex2, ten FMA-pipe instructions.  erff() costs about twice
// the instructions with divergent ranges; as the epilogue of the Branchformer's 512 -> 3072 pre-projection (3072 activations per frame
// against 512 MACs each) it bounded that GEMM: 229 us against 145 us of tensor time at 43.7 k frames.
__device__ __forceinline__ float act_gelu_erf(float x) {
  const float u = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, u, 1.0f)));
  float pl = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  pl = fmaf(pl, t, 0.5f * 1.421413741f);
  pl = fmaf(pl, t, 0.5f * -0.284496736f);
  pl = fmaf(pl, t, 0.5f * 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * u * u));
  const float q = pl * t * e;                 // 0.5 * erfc(|x| / sqrt 2) = Phi(-|x|)
  return x * (x >= 0.0f ? 1.0f - q : q);
}
__device__ __forceinline__ float act_gelu_tanh(float x) {
  float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  float h = 0.5f * x;
  return fmaf(h, tanh_approx(u), h);
}
// scalar form (cold paths only: the switch sits inside the caller's element loop)
__device__ __forceinline__ float act_fast(int act, float x) {
  switch (act) {
    case 1: return act_swish(x);
    case 2: return act_gelu_erf(x);
    case 3: return fmaxf(x, 0.0f);
    case 4: return x >= 0.0f ? x : 0.01f * x;
    case 5: return tanh_approx(x);
    case 6: return act_sigmoid(x);
    case 7: return act_gelu_tanh(x);
    default: return x;
  }
}
// in-place activation of N values; the dispatch is OUTSIDE the element loops (one tight loop per case)
template <int N>
__device__ __forceinline__ void act_apply(int act, float* v) {
  switch (act) {
    case 1:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = act_swish(v[i]);
      break;
    case 2:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = act_gelu_erf(v[i]);
      break;
    case 3:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = fmaxf(v[i], 0.0f);
      break;
    case 4:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = v[i] >= 0.0f ? v[i] : 0.01f * v[i];
      break;
    case 5:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = tanh_approx(v[i]);
      break;
    case 6:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = act_sigmoid(v[i]);
      break;
    case 7:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = act_gelu_tanh(v[i]);
      break;
    default: break;
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---- row-tile prologue shared by the fused kernels ----------------------------------------------------
// Prologue warp pw (of four) brings rows [32 pw, 32 pw + 32) of the (128 x D) bf16 tile straight into the A operand image
// (K-blocks of 128 rows x 128 B, 128B swizzle) with 16-byte cp.async copies: lane l copies the 16-byte chunk l of every
// row, all 32 rows in flight at once, no registers held.  LayerNorm (eps 1e-5, fp32) then runs in place with ONE THREAD PER
// ROW: lane l owns row 32 pw + l and walks its D / 8 chunks (the swizzle makes the row-per-lane accesses conflict-free), so
// there are no shuffles and every chunk is independent work.  Statistics are shifted by the row's first element
// (var = E[(x-x0)^2] - E[x-x0]^2: one read, no cancellation for rows with a large mean); a second read normalises.
// Rows >= nrows become zero rows.  D % 8 == 0, D <= 256.  sW / sB: LN weight / bias in shared memory (D floats each).
// The caller orders the A operand for the async proxy afterwards (fence_proxy_async).
__device__ __forceinline__ void unpack_bf16x8(const uint4& raw, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
}
__device__ __forceinline__ void stage_ln_rows(uint8_t* sX, const __nv_bfloat16* x, int64_t ldx, int64_t row0, int nrows, int D,
                                              int pw, int lane, bool do_ln, const float* sW, const float* sB,
                                              unsigned long long* tr = nullptr) {
  const int nch = D >> 3;
  {
    const bool has = lane < nch;
    uint8_t* const base = sX + (size_t)(lane >> 3) * kblock_bytes(128);
    const __nv_bfloat16* src = x + (row0 + pw * 32) * ldx + lane * 8;
    if (has) {
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const int r = pw * 32 + j;
        const bool in = r < nrows;
        cp_async16(base + sw128_offset(r, lane & 7), in ? (const void*)(src + (int64_t)j * ldx) : (const void*)x, in ? 16u : 0u);
      }
    }
    cp_async_commit();
  }
  if (tr && lane == 0) tr[0] = clock64();
  cp_async_wait_all();
  __syncwarp();  // the row a lane normalises was copied by all lanes of this warp
  if (tr && lane == 0) tr[1] = clock64();
  if (!do_ln) return;
  const int r = pw * 32 + lane;
  uint8_t* const rowp = sX + (r >> 3) * 1024 + (r & 7) * 128;
  const uint32_t rx = (uint32_t)(r & 7);
  float v[8];
  unpack_bf16x8(*reinterpret_cast<const uint4*>(rowp + (rx << 4)), v);  // chunk 0
  const float x0 = v[0];
  float a1[8], a2[8];  // eight independent accumulator chains
#pragma unroll
  for (int e = 0; e < 8; ++e) { a1[e] = 0.0f; a2[e] = 0.0f; }
#pragma unroll 4
  for (int c = 0; c < nch; ++c) {
    unpack_bf16x8(*reinterpret_cast<const uint4*>(rowp + (size_t)(c >> 3) * kblock_bytes(128) + ((((uint32_t)c & 7u) ^ rx) << 4)), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float d = v[e] - x0; a1[e] += d; a2[e] = fmaf(d, d, a2[e]); }
  }
  const float s1 = ((a1[0] + a1[1]) + (a1[2] + a1[3])) + ((a1[4] + a1[5]) + (a1[6] + a1[7]));
  const float s2 = ((a2[0] + a2[1]) + (a2[2] + a2[3])) + ((a2[4] + a2[5]) + (a2[6] + a2[7]));
  const float invD = 1.0f / (float)D;
  const float m0 = s1 * invD;                                  // mean - x0
  const float rstd = rsqrtf(fmaxf(s2 * invD - m0 * m0, 0.0f) + 1e-5f);
  const float shift = -(m0 + x0) * rstd;                       // (x - mean) * rstd = x * rstd + shift
  const bool live = r < nrows;
#pragma unroll 4
  for (int c = 0; c < nch; ++c) {
    uint4* const cp = reinterpret_cast<uint4*>(rowp + (size_t)(c >> 3) * kblock_bytes(128) + ((((uint32_t)c & 7u) ^ rx) << 4));
    unpack_bf16x8(*cp, v);
    const float4 w0 = *reinterpret_cast<const float4*>(sW + c * 8), w1 = *reinterpret_cast<const float4*>(sW + c * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(sB + c * 8), b1 = *reinterpret_cast<const float4*>(sB + c * 8 + 4);
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = live ? fmaf(fmaf(v[e], rstd, shift), wv[e], bv[e]) : 0.0f;
    *cp = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// The same tile prologue spread over 16 warps (the epilogue warps, idle while a CTA's FIRST tile is being staged): warp w
// takes rows [8w, 8w + 8); for the copy lane l moves chunk l of each row, for the LayerNorm lane l owns row 8w + (l & 7)
// and the column quarter l >> 3 (D / 32 chunks), the four partial statistics of a row meeting through two shuffles.
// The 8 lanes of each 128-bit shared-memory phase hold 8 different rows of one quarter: conflict-free under the swizzle.
// D % 32 == 0, D <= 256.
__device__ __forceinline__ void rows8_copy(uint8_t* sX, const __nv_bfloat16* x, int64_t ldx, int64_t row0, int nrows, int D, int w, int lane) {
  const int nch = D >> 3;
  if (lane < nch) {
    uint8_t* const base = sX + (size_t)(lane >> 3) * kblock_bytes(128);
    const __nv_bfloat16* src = x + (row0 + w * 8) * ldx + lane * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = w * 8 + j;
      const bool in = r < nrows;
      cp_async16(base + sw128_offset(r, lane & 7), in ? (const void*)(src + (int64_t)j * ldx) : (const void*)x, in ? 16u : 0u);
    }
  }
}
// packed fp32 pairs (Blackwell FFMA2 / FADD2: one issue slot for two independent IEEE operations)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ void unpack_bf16x8_pairs(const uint4& raw, float2* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = __bfloat1622float2(h[e]);
}
// LayerNorm of rows [8w, 8w + 8) of the staged tile, in place, by one warp.
//   pass 1 (statistics): lane l owns row 8w + (l & 7) and the column quarter l >> 3; shifted one-pass sums, the four
//     partials of a row meet through two shuffles; lanes 0..7 publish (1/std, -mean/std) of their row in sStat;
//   pass 2 (normalise): lane l owns 16-byte chunk l of every row: its eight weights / biases sit in registers for the
//     whole call, so the pass costs one shared-memory load and one store per row and lane (reading the parameters per
//     row quarter instead made the shared-memory pipe, not the math, the limit of this phase).
// Both mappings are conflict-free under the 128-byte swizzle.  sStat: 128 float2 (one per row of the tile).
__device__ __forceinline__ void rows8_ln(uint8_t* sX, int nrows, int D, int w, int lane, const float* sW, const float* sB, float2* sStat,
                                         unsigned long long* tr = nullptr) {
  const int nch = D >> 3;
  {
    const int r = w * 8 + (lane & 7), qq = lane >> 3, cq = nch >> 2;  // chunks per quarter
    const uint32_t rx = (uint32_t)(r & 7);
    const uint8_t* const rowp = sX + (r >> 3) * 1024 + (r & 7) * 128;
    float2 v[4];
    unpack_bf16x8_pairs(*reinterpret_cast<const uint4*>(rowp + (rx << 4)), v);  // chunk 0 of the row: the common shift
    const float x0 = v[0].x;
    const float2 nx0 = make_float2(-x0, -x0);
    float2 a1[4], a2[4];  // four independent pair-accumulator chains
#pragma unroll
    for (int e = 0; e < 4; ++e) { a1[e] = make_float2(0.0f, 0.0f); a2[e] = make_float2(0.0f, 0.0f); }
#pragma unroll 2
    for (int i = 0; i < cq; ++i) {
      const int c = qq * cq + i;
      unpack_bf16x8_pairs(*reinterpret_cast<const uint4*>(rowp + (size_t)(c >> 3) * kblock_bytes(128) + ((((uint32_t)c & 7u) ^ rx) << 4)), v);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 d = add2(v[e], nx0); a1[e] = add2(a1[e], d); a2[e] = fma2(d, d, a2[e]); }
    }
    const float2 s1p = add2(add2(a1[0], a1[1]), add2(a1[2], a1[3])), s2p = add2(add2(a2[0], a2[1]), add2(a2[2], a2[3]));
    float s1 = s1p.x + s1p.y, s2 = s2p.x + s2p.y;
    s1 += __shfl_xor_sync(0xffffffffu, s1, 8); s2 += __shfl_xor_sync(0xffffffffu, s2, 8);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 16); s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
    const float invD = 1.0f / (float)D;
    const float m0 = s1 * invD;
    // rows past the end of the tile become zero rows: scale and shift vanish (and pass 2 drops the bias)
    const float rstd = r < nrows ? rsqrtf(fmaxf(s2 * invD - m0 * m0, 0.0f) + 1e-5f) : 0.0f;
    if (lane < 8) sStat[r] = make_float2(rstd, -(m0 + x0) * rstd);
  }
  __syncwarp();
  if (tr && lane == 0) *tr = clock64();
  if (lane < nch) {
    const int c = lane;
    const int po = c * 8 + 4 * (c / (nch >> 2));  // padded parameter layout (ln_pad_index)
    const float4 w0 = *reinterpret_cast<const float4*>(sW + po), w1 = *reinterpret_cast<const float4*>(sW + po + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(sB + po), b1 = *reinterpret_cast<const float4*>(sB + po + 4);
    const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
    const float2 bv[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
    const float2 z2 = make_float2(0.0f, 0.0f);
    uint8_t* const kb = sX + (size_t)(c >> 3) * kblock_bytes(128);
#pragma unroll 4
    for (int j = 0; j < 8; ++j) {
      const int r = w * 8 + j;
      const float2 st = sStat[r];
      const float lv = r < nrows ? 1.0f : 0.0f;
      const float2 rs2 = make_float2(st.x, st.x), sh2 = make_float2(st.y, st.y), lv2 = make_float2(lv, lv);
      uint4* const cp = reinterpret_cast<uint4*>(kb + (r >> 3) * 1024 + (r & 7) * 128 + ((((uint32_t)c & 7u) ^ (uint32_t)(r & 7)) << 4));
      float2 v[4];
      unpack_bf16x8_pairs(*cp, v);
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 y = fma2(fma2(v[e], rs2, sh2), wv[e], fma2(bv[e], lv2, z2));  // (x*rstd + shift) * w + b; b -> 0 on dead rows
        o[e] = pack_bf16x2(y.x, y.y);
      }
      *cp = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncwarp();
}
// shared-memory index of LN parameter i (of D) in the padded layout rows8_ln reads: 4 floats of padding per column quarter
__device__ __forceinline__ int ln_pad_index(int i, int D) { return i + 4 * (i / (D >> 2)); }
// n_groups consecutive 8-row groups starting at group w0 (1 for each of the 16 epilogue warps on a CTA's first tile, 4 for
// each of the 4 prologue warps afterwards): every copy in flight first, then the LayerNorm passes.  One implementation for
// both callers keeps the kernels' code small (each kernel's code is fetched cold at every launch of a layer's chain).
static __device__ __noinline__ void stage_ln_rows_wide(uint8_t* sX, const __nv_bfloat16* x, int64_t ldx, int64_t row0, int nrows, int D,
                                                   int w0, int n_groups, int lane, bool do_ln, const float* sW, const float* sB, float2* sStat) {
#pragma unroll 1
  for (int i = 0; i < n_groups; ++i) rows8_copy(sX, x, ldx, row0, nrows, D, w0 + i, lane);
  cp_async_commit();
  cp_async_wait_all();
  __syncwarp();
  if (!do_ln) return;
#pragma unroll 1
  for (int i = 0; i < n_groups; ++i) rows8_ln(sX, nrows, D, w0 + i, lane, sW, sB, sStat);
}

}  // namespace tc
}  // namespace smx
