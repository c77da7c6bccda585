// tcgen05 arm of libsmx, part 2: K-LIN, the row-tile linear kernel.
//
//   out[128-row tile] = epilogue( LN?(x tile) @ W^T )
//
// One CTA per 128-row tile.  The x tile is normalised by all warps and written to shared memory as the
// UMMA A operand (bf16, K-major, 128B swizzle).  W streams from its packed image through a ring of
// shared-memory stages filled by cp.async.bulk (TMA engine) and tracked with mbarriers; one thread
// issues tcgen05.mma into a double-buffered TMEM accumulator; eight epilogue warps drain TMEM with
// tcgen05.ld and apply the fused epilogue (bias, row-group bias, activation, row mask, residual, GLU,
// output LayerNorm, or masked column sums for the SummaryMixing time reduction).
// Block-diagonal (ParallelLinear) weights skip their zero K-blocks.
#include "smx_tc.h"
#include "smx_tc_common.cuh"

namespace smx {

using tc::kblock_bytes;

enum { LIN_PLAIN = TC_LIN_PLAIN, LIN_GLU = TC_LIN_GLU, LIN_OLN = TC_LIN_OLN, LIN_COLSUM = TC_LIN_COLSUM };
constexpr int LIN_THREADS = 320;  // warps 0..7: epilogue, warp 8: producer, warp 9: MMA issuer (highest id = arbiter priority)
constexpr int LIN_MAX_STAGES = 8;

// (nt, kb) -> sub-range of the n-tile's columns that K-block kb contributes to
__device__ __forceinline__ bool lin_chunk(const LinP& p, int nt, int kb, int& rel_lo, int& n_cnt, int& first) {
  if (p.head_in == 0) {
    rel_lo = 0; n_cnt = p.NT; first = (kb == 0);
    return true;
  }
  const int m = (kb * 64) / p.head_in;
  int lo = m * p.head_out, hi = lo + p.head_out;
  const int t0 = nt * p.NT, t1 = t0 + p.NT;
  lo = lo > t0 ? lo : t0;
  hi = hi < t1 ? hi : t1;
  if (lo >= hi) return false;
  rel_lo = lo - t0; n_cnt = hi - lo; first = ((kb * 64) % p.head_in == 0);
  return true;
}
__device__ __forceinline__ int lin_order(const LinP& p, int i) {
  if (!p.glu) return i;
  return (i & 1) ? (i >> 1) + p.n_tiles / 2 : (i >> 1);
}

// sum over the 32 lanes of v[j] for each j; lane l ends up holding column l's total in v[0]
__device__ __forceinline__ float warp_column_sums(float* v, int lane) {
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

template <int MODE>
__global__ void __launch_bounds__(LIN_THREADS, 1) lin_kernel(const LinP p) {
  extern __shared__ __align__(1024) uint8_t smem[];  // no pointer arithmetic through integers: keeps LDS/STS addressing
  uint8_t* sA = smem;
  uint8_t* sW = smem + (size_t)128 * p.K * 2;
  __shared__ __align__(8) uint64_t full_bar[LIN_MAX_STAGES], empty_bar[LIN_MAX_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[2][2][128];   // [sum|sq][column half][row]   (LIN_OLN)
  __shared__ float colred[4][256];   // [quadrant][column]           (LIN_COLSUM)

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler (role dispatch below)
  // tile -> first row and number of valid rows (utterance-aligned tiles never straddle two utterances)
  int64_t row0;
  int nrows;
  if (p.utt_tiles) {
    const int tpu = (p.T + 127) / 128;
    const int b = blockIdx.x / tpu, t0 = (blockIdx.x % tpu) * 128;
    row0 = (int64_t)b * p.T + t0;
    nrows = p.T - t0 < 128 ? p.T - t0 : 128;
  } else {
    row0 = (int64_t)blockIdx.x * 128;
    nrows = p.rows - row0 < 128 ? (int)(p.rows - row0) : 128;
  }
  const int nkb = p.K / 64;
  const int NT = p.NT;

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, p.tmem_cols);
  if (tid == 32) {
    for (int s = 0; s < p.n_stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&acc_full[0], 1); tc::mbar_init(&acc_full[1], 1);
    tc::mbar_init(&acc_empty[0], 8); tc::mbar_init(&acc_empty[1], 8);
    tc::fence_barrier_init();
  }
  __syncthreads();

  // ---- producer state (also used for the pre-issue before the prologue) ------------------------
  int pr_i = 0, pr_kb = 0, pr_s = 0, pr_ph = 0, pr_issued = 0;
  auto produce = [&](int max_chunks) {  // lane 0 of warp 0 only
    while (pr_i < p.n_tiles && max_chunks > 0) {
      const int nt = lin_order(p, pr_i);
      int rel_lo, n_cnt, first;
      if (lin_chunk(p, nt, pr_kb, rel_lo, n_cnt, first)) {
        tc::mbar_wait(&empty_bar[pr_s], pr_ph ^ 1);
        const uint32_t bytes = (uint32_t)n_cnt * 128u;
        tc::mbar_arrive_expect_tx(&full_bar[pr_s], bytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wp) + ((size_t)nt * nkb + pr_kb) * kblock_bytes(NT) + (size_t)rel_lo * 128;
        tc::bulk_g2s(sW + (size_t)pr_s * p.stage_bytes + (size_t)rel_lo * 128, src, bytes, &full_bar[pr_s]);
        if (++pr_s == p.n_stages) { pr_s = 0; pr_ph ^= 1; }
        --max_chunks; ++pr_issued;
      }
      if (++pr_kb == nkb) { pr_kb = 0; ++pr_i; }
    }
  };
  if (tid == 256) produce(p.n_stages);  // weights start streaming while the tile is normalised

  // ---- prologue: x tile -> (LayerNorm) -> A operand ------------------------------------------------
  {
    const int nchunk = p.K / 8;  // 16-byte chunks per row
    constexpr int NW = LIN_THREADS / 32;
    for (int rb = warp; rb < 128; rb += 4 * NW) {  // four rows per round trip to memory
    uint4 rawb[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rb + j * NW;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ck = lane + 32 * c;
        rawb[j][c] = make_uint4(0, 0, 0, 0);
        if (ck < nchunk && r < nrows) rawb[j][c] = *reinterpret_cast<const uint4*>(p.x + (row0 + r) * p.ldx + ck * 8);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rb + j * NW;
      if (r >= 128) break;
      float v[2][8];
      const bool live = r < nrows;
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
      if (p.ln_w) {
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int e = 0; e < 8; ++e) s += v[c][e];
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)p.K;
        float q = 0.0f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int ck = lane + 32 * c;
          if (ck < nchunk) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { float d = v[c][e] - mean; q += d * d; }
          }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q / (float)p.K + p.ln_eps);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int ck = lane + 32 * c;
          if (ck < nchunk) {
            const float4 w0 = *reinterpret_cast<const float4*>(p.ln_w + ck * 8), w1 = *reinterpret_cast<const float4*>(p.ln_w + ck * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(p.ln_b + ck * 8), b1 = *reinterpret_cast<const float4*>(p.ln_b + ck * 8 + 4);
            const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) v[c][e] = live ? (v[c][e] - mean) * rstd * ww[e] + bb[e] : 0.0f;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ck = lane + 32 * c;
        if (ck < nchunk) {
          uint4 o = make_uint4(tc::pack_bf16x2(v[c][0], v[c][1]), tc::pack_bf16x2(v[c][2], v[c][3]),
                               tc::pack_bf16x2(v[c][4], v[c][5]), tc::pack_bf16x2(v[c][6], v[c][7]));
          *reinterpret_cast<uint4*>(sA + (size_t)(ck >> 3) * kblock_bytes(128) + tc::sw128_offset(r, ck & 7)) = o;
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

  if (warp == 8) {
    // =============================== producer ===============================
    if (lane == 0) produce(1 << 30);
  } else if (warp == 9) {
    // =============================== MMA issuer ===============================
    // the whole warp walks the (warp-uniform) schedule; one elected lane issues tcgen05.mma / tcgen05.commit
    {
      int s = 0, ph = 0;
      const uint32_t a0 = tc::smem_u32(sA), w0 = tc::smem_u32(sW);
      for (int i = 0; i < p.n_tiles; ++i) {
        const int nt = lin_order(p, i), buf = i & 1;
        tc::mbar_wait(&acc_empty[buf], ((i >> 1) & 1) ^ 1);
        tc::tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          int rel_lo, n_cnt, first;
          if (!lin_chunk(p, nt, kb, rel_lo, n_cnt, first)) continue;
          tc::mbar_wait(&full_bar[s], ph);
          tc::tc_fence_after();
          const uint32_t idesc = tc::make_idesc_bf16(128, (uint32_t)n_cnt);
          const uint32_t a_addr = a0 + kb * kblock_bytes(128);
          const uint32_t b_addr = w0 + s * p.stage_bytes + rel_lo * 128;
          const uint32_t d_addr = tmem + buf * NT + rel_lo;
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc::umma_bf16(d_addr, tc::make_desc_sw128(a_addr + ks * 32), tc::make_desc_sw128(b_addr + ks * 32), idesc,
                            (first && ks == 0) ? 0u : 1u);
            tc::umma_commit(&empty_bar[s]);
          }
          __syncwarp();
          if (++s == p.n_stages) { s = 0; ph ^= 1; }
        }
        if (tc::elect_one()) tc::umma_commit(&acc_full[buf]);
        __syncwarp();
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int hf = warp >> 2;           // column half
    const int r = q * 32 + lane;        // row inside the tile
    const int64_t row = row0 + r;
    const bool live = r < nrows;
    const float rmask = (p.rowmask && live) ? (float)p.rowmask[row] : 1.0f;
    const int64_t grp = live ? row / p.T : 0;
    const int half_cols = NT / 2;       // multiple of 32
    const int npieces = half_cols / 32; // 1..4
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const int etid = tid;               // 0..255 inside the epilogue group

    const int n_iter = (MODE == LIN_GLU) ? p.n_tiles / 2 : p.n_tiles;
    for (int it = 0; it < n_iter; ++it) {
      int nt, buf;
      if (MODE == LIN_GLU) {
        nt = it; buf = 0;
        tc::mbar_wait(&acc_full[0], it & 1);
        tc::mbar_wait(&acc_full[1], it & 1);
      } else {
        nt = it; buf = it & 1;
        tc::mbar_wait(&acc_full[buf], (it >> 1) & 1);
      }
      tc::tc_fence_after();
      float keep[(MODE == LIN_OLN) ? 4 : 1][32];
      float s1 = 0.0f;
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) {
        if (pc >= npieces) break;
        const int ct = hf * half_cols + pc * 32;  // column inside the n-tile
        const int col = nt * NT + ct;             // global output column
        float v[32];
        tc::tmem_ld32(tmem + lane_sel + buf * NT + ct, v);
        float g[(MODE == LIN_GLU) ? 32 : 1];
        if (MODE == LIN_GLU) tc::tmem_ld32(tmem + lane_sel + NT + ct, g);
        tc::tmem_ld_wait();
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(p.bias + col + j);
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (MODE == LIN_GLU) {
          const float* gb = p.bias + p.N / 2 + col;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] * tc::act_fast(SMX_ACT_SIGMOID, g[j] + gb[j]);
        }
        if (p.rowbias) {
          const float* rb = p.rowbias + grp * p.rowbias_ld + col;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(rb + j);
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (p.act != SMX_ACT_IDENTITY) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tc::act_fast(p.act, v[j]);
        }
        if (p.rowmask) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= rmask;
        }
        if (p.resid && live) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.resid + row * p.ldr + col);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 raw = rp[c];
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 f = __bfloat1622float2(h[e]);
              v[c * 8 + 2 * e] = fmaf(p.alpha, v[c * 8 + 2 * e], f.x);
              v[c * 8 + 2 * e + 1] = fmaf(p.alpha, v[c * 8 + 2 * e + 1], f.y);
            }
          }
        }
        if (MODE == LIN_OLN) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { keep[pc][j] = v[j]; s1 += v[j]; }
        } else if (MODE == LIN_COLSUM) {
          if (!live) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.0f;
          }
          const float tot = warp_column_sums(v, lane);
          colred[q][ct + lane] = tot;
        } else if (live) {
          uint4* op = reinterpret_cast<uint4*>(p.out + row * p.ldo + ((MODE == LIN_GLU) ? col : col));
#pragma unroll
          for (int c = 0; c < 4; ++c)
            op[c] = make_uint4(tc::pack_bf16x2(v[c * 8], v[c * 8 + 1]), tc::pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]),
                               tc::pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]), tc::pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]));
        }
      }
      // accumulator drained: hand the TMEM buffer(s) back to the MMA issuer
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(&acc_empty[buf]);
        if (MODE == LIN_GLU) tc::mbar_arrive(&acc_empty[1]);
      }
      if (MODE == LIN_OLN) {  // LayerNorm over the full output row (two column halves live in two warps)
        red[0][hf][r] = s1;
        tc::named_bar_sync(1, 256);
        const float mean = (red[0][0][r] + red[0][1][r]) / (float)p.N;
        float s2 = 0.0f;
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          if (pc >= npieces) break;
#pragma unroll
          for (int j = 0; j < 32; ++j) { float d = keep[pc][j] - mean; s2 += d * d; }
        }
        red[1][hf][r] = s2;
        tc::named_bar_sync(1, 256);
        const float rstd = rsqrtf((red[1][0][r] + red[1][1][r]) / (float)p.N + p.oln_eps);
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          if (pc >= npieces) break;
          const int col = hf * half_cols + pc * 32;
          if (live) {
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = (keep[pc][j] - mean) * rstd * p.oln_w[col + j] + p.oln_b[col + j];
            uint4* op = reinterpret_cast<uint4*>(p.out + row * p.ldo + col);
#pragma unroll
            for (int c = 0; c < 4; ++c)
              op[c] = make_uint4(tc::pack_bf16x2(o[c * 8], o[c * 8 + 1]), tc::pack_bf16x2(o[c * 8 + 2], o[c * 8 + 3]),
                                 tc::pack_bf16x2(o[c * 8 + 4], o[c * 8 + 5]), tc::pack_bf16x2(o[c * 8 + 6], o[c * 8 + 7]));
          }
        }
      }
      if (MODE == LIN_COLSUM) {  // fixed-order reduction over the four row quadrants: deterministic
        tc::named_bar_sync(1, 256);
        if (etid < NT)
          p.colsum[((size_t)blockIdx.x * p.N) + nt * NT + etid] =
              (colred[0][etid] + colred[1][etid]) + (colred[2][etid] + colred[3][etid]);
        tc::named_bar_sync(1, 256);
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
// n-tile width: the largest of 256/192/128/64 dividing N (GLU: dividing N/2, so a-tiles pair with gate tiles)
int tc_pick_nt(int N, int glu) {
  const int n = glu ? N / 2 : N;
  for (int nt = 256; nt >= 64; nt -= 64)
    if (n % nt == 0) return nt;
  return 0;
}

bool tc_linear_supported(int K, int N) { return K >= 64 && K <= 512 && K % 64 == 0 && N >= 64 && N % 64 == 0; }

size_t tc_linear_packed_bytes(int K, int N) { return (size_t)N * (size_t)K * 2; }

// block-diagonal aware gather: element (n,k) of the dense (N x K) view of an smx_linear
// glu_il: image row block 2c holds value rows [64c, 64c+64), block 2c+1 the gate rows [N/2 + 64c, ...) (NT == 64)
__global__ void pack_linear_kernel(const float* w, int in_dim, int out_dim, int n_split, int k_offset, int K, int N,
                                   int NT, __nv_bfloat16* out, int glu_il = 0) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunks_per_row = K / 8;
  if (i >= (int64_t)N * chunks_per_row) return;
  const int n_img = (int)(i / chunks_per_row), ck = (int)(i % chunks_per_row);
  const int n = glu_il ? ((n_img / 64) & 1) * (N / 2) + (n_img / 128) * 64 + (n_img % 64) : n_img;
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = ck * 8 + e;
    float val;
    if (n_split <= 1) {
      val = w[(int64_t)n * in_dim + k_offset + k];  // nn.Linear (out,in)
    } else {
      const int hi = in_dim / n_split, ho = out_dim / n_split;
      const int m = n / ho;
      val = (k / hi == m) ? w[((int64_t)m * hi + (k - m * hi)) * ho + (n - m * ho)] : 0.0f;  // (h, in/h, out/h)
    }
    f[e] = val;
  }
  const int tile = n_img / NT, r = n_img % NT, kb = ck / 8, c16 = ck % 8;
  const size_t off = ((size_t)tile * (K / 64) + kb) * kblock_bytes(NT) + tc::sw128_offset(r, c16);
  *reinterpret_cast<uint4*>(reinterpret_cast<char*>(out) + off) =
      make_uint4(tc::pack_bf16x2(f[0], f[1]), tc::pack_bf16x2(f[2], f[3]), tc::pack_bf16x2(f[4], f[5]), tc::pack_bf16x2(f[6], f[7]));
}

int tc_pack_linear(const smx_linear& L, int k_offset, int K, int glu, void* out, cudaStream_t st) {
  const int N = L.out_dim;
  if (!tc_linear_supported(K, N)) return fail(SMX_ERR_UNSUPPORTED, "tc pack: K=%d N=%d not supported", K, N);
  const int NT = tc_pick_nt(N, glu);
  if (NT == 0) return fail(SMX_ERR_UNSUPPORTED, "tc pack: N=%d has no tile width", N);
  int64_t n = (int64_t)N * (K / 8);
  pack_linear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L.w, L.in_dim, L.out_dim, L.n_split, k_offset, K, N, NT,
                                                                  (__nv_bfloat16*)out);
  count_launch();
  return check_launch("pack_linear_kernel");
}

int tc_pack_linear_nt(const smx_linear& L, int k_offset, int K, int NT, void* out, cudaStream_t st, int glu_interleave) {
  const int N = L.out_dim;
  if (K % 64 || N % NT || NT % 8) return fail(SMX_ERR_UNSUPPORTED, "tc pack: K=%d N=%d NT=%d not supported", K, N, NT);
  if (glu_interleave && (NT != 64 || N % 128)) return fail(SMX_ERR_UNSUPPORTED, "tc pack: GLU interleave needs NT=64, N%%128==0");
  int64_t n = (int64_t)N * (K / 8);
  pack_linear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L.w, L.in_dim, L.out_dim, L.n_split, k_offset, K, N, NT,
                                                                  (__nv_bfloat16*)out, glu_interleave);
  count_launch();
  return check_launch("pack_linear_kernel");
}

int tc_linear_launch(LinP p, int mode, cudaStream_t st) {
  if (!tc_linear_supported(p.K, p.N)) return fail(SMX_ERR_UNSUPPORTED, "tc linear: K=%d N=%d not supported", p.K, p.N);
  p.NT = tc_pick_nt(p.N, mode == LIN_GLU);
  if (p.NT == 0) return fail(SMX_ERR_UNSUPPORTED, "tc linear: N=%d has no tile width", p.N);
  p.n_tiles = p.N / p.NT;
  if (p.head_in) {  // block-diagonal: heads must align with 64-wide K-blocks and 16-wide MMA N
    if (p.head_in % 64 || p.head_out % 16) return fail(SMX_ERR_UNSUPPORTED, "tc linear: head dims %d->%d", p.head_in, p.head_out);
  }
  if (mode == LIN_GLU) {
    if (p.n_tiles % 2) return fail(SMX_ERR_UNSUPPORTED, "tc linear GLU: N=%d needs an even number of %d-wide tiles", p.N, p.NT);
    p.glu = 1;
  }
  if (mode == LIN_OLN && p.n_tiles != 1) return fail(SMX_ERR_UNSUPPORTED, "tc linear output-LN: N=%d > 256", p.N);
  if (p.T <= 0) p.T = 1;
  p.stage_bytes = p.NT * 128;
  const size_t a_bytes = (size_t)128 * p.K * 2;
  const size_t budget = 227 * 1024 - 1024 /*align*/ - 8 * 1024 /*static*/;
  int stages = (int)((budget - a_bytes) / p.stage_bytes);
  if (stages > LIN_MAX_STAGES) stages = LIN_MAX_STAGES;
  if (stages < 2) return fail(SMX_ERR_UNSUPPORTED, "tc linear: tile does not fit shared memory");
  p.n_stages = stages;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * p.NT)) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = a_bytes + (size_t)stages * p.stage_bytes + 1024;
  unsigned grid = (unsigned)((p.rows + 127) / 128);
  if (p.utt_tiles) grid = (unsigned)((p.rows / p.T) * ((p.T + 127) / 128));
  cudaError_t e;
#define SMX_LAUNCH_LIN(M)                                                                                      \
  e = cudaFuncSetAttribute(lin_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(lin_kernel): %s", cudaGetErrorString(e)); \
  lin_kernel<M><<<grid, LIN_THREADS, smem, st>>>(p);
  switch (mode) {
    case LIN_PLAIN: SMX_LAUNCH_LIN(LIN_PLAIN); break;
    case LIN_GLU: SMX_LAUNCH_LIN(LIN_GLU); break;
    case LIN_OLN: SMX_LAUNCH_LIN(LIN_OLN); break;
    case LIN_COLSUM: SMX_LAUNCH_LIN(LIN_COLSUM); break;
    default: return fail(SMX_ERR_BAD_ARG, "tc linear: bad mode");
  }
#undef SMX_LAUNCH_LIN
  count_tc_launch();
  return check_launch("lin_kernel");
}

}  // namespace smx
