// tcgen05 arm of libsmx, part 3: the reference's module forwards composed from tensor-core kernels
// (bf16 activations, fp32 accumulation).  Small per-utterance work (mean finalisation) and the
// depthwise convolution run on CUDA cores.
#include "smx_tc.h"
#include "smx_tc_common.cuh"

namespace smx {

// =============================================================================================
// SummaryMixing cell, mode "SummaryMixing", whole-utterance mean          summary_mixing.py:198-253
// =============================================================================================
static bool mlp_ok(const smx_linear* blk, int n, int in_dim) {
  if (n < 1 || n > SMX_MAX_BLOCKS) return false;
  int cur = in_dim;
  for (int i = 0; i < n; ++i) {
    const smx_linear& L = blk[i];
    if (!L.w || !L.b || L.in_dim != cur) return false;
    if (!tc_linear_supported(L.in_dim, L.out_dim)) return false;
    if (L.n_split > 1) {
      if (L.in_dim % L.n_split || L.out_dim % L.n_split) return false;
      if ((L.in_dim / L.n_split) % 64 || (L.out_dim / L.n_split) % 16) return false;
    }
    cur = L.out_dim;
  }
  return true;
}

// modes "SummaryMixing-lite" (summary_mixing.py:300-324) and "SummaryMixing-fast" (:255-298) on the row-tile linear kernels
static bool lite_ok(const smx_cell_weights* w) {
  return mlp_ok(w->summary, w->n_summary, w->enc_dim) && w->summary[w->n_summary - 1].out_dim == w->summary_out_dim &&
         w->summary_out_dim <= 1024;
}
static bool fast_ok(const smx_cell_weights* w) {
  const smx_linear& g = w->global_proj;
  const int Dl = w->local_out_dim;
  if (!g.w || !g.b || g.n_split > 1 || g.in_dim != w->enc_dim || g.out_dim != 2 * Dl) return false;
  if (!tc_linear_supported(g.in_dim, Dl)) return false;
  if (w->merge.n_split > 1 || w->merge.in_dim != 2 * Dl || !w->merge.w || !w->merge.b) return false;
  return tc_linear_supported(Dl, w->merge.out_dim) && Dl <= 1024;
}
bool tc_cell_supported(const smx_cell_weights* w, int has_sum_mask) {
  if (w->mode == SMX_MODE_LITE) return lite_ok(w);             // (lite ignores sum_mask, :300-324)
  if (w->mode == SMX_MODE_FAST) return !has_sum_mask && fast_ok(w);
  if (w->mode != SMX_MODE_FULL || has_sum_mask) return false;
  if (tc_cellf_supported(w)) return true;  // fused kernels: any n_split (heads not aligned to 64 are packed as dense)
  if (!mlp_ok(w->local, w->n_local, w->enc_dim) || !mlp_ok(w->summary, w->n_summary, w->enc_dim)) return false;
  const int Dl = w->local_out_dim, Ds = w->summary_out_dim;
  if (w->local[w->n_local - 1].out_dim != Dl || w->summary[w->n_summary - 1].out_dim != Ds) return false;
  if (w->merge.n_split > 1 || w->merge.in_dim != Dl + Ds || !w->merge.w || !w->merge.b) return false;
  if (!tc_linear_supported(Dl, w->merge.out_dim)) return false;
  if (Ds > 1024) return false;
  if (w->use_layernorm && (!w->local_norm_w || !w->summary_norm_w)) return false;
  return true;
}

struct CellLayout {
  size_t local[SMX_MAX_BLOCKS], summary[SMX_MAX_BLOCKS], merge, total;
  size_t local_s[2], summary_s[2], merge_s;  // the same images in schedule order (K-SM v3)
  size_t v4;                                 // stream-order image + W_cs^T of K-SM v4 (0 bytes when v4 does not apply)
  size_t local_d[SMX_MAX_BLOCKS], summary_d[SMX_MAX_BLOCKS], merge_d;  // dense (out, in) bf16 copies for K-GEMM (unfused full mode; 0 = not used)
  int gemm;                                  // the unfused full-mode cell runs its plain linears on K-GEMM
};
// Unfused full-mode cells (conformer_large: D = 512) run every linear that needs no column-sum epilogue on K-GEMM (operands by
// TMA, double-buffered accumulators; block-diagonal weights as dense with zero blocks): K-LIN took 160-180 us per 512 x 512
// linear at B=32, T=1000, K-GEMM takes ~25 (profiles/r02_notes.md).
static bool cell_gemm_ok(const smx_cell_weights* w) {
  if (w->mode != SMX_MODE_FULL || tc_cellf_supported(w)) return false;
  for (int i = 0; i < w->n_local; ++i) if (!tc_gemm_supported(w->local[i].in_dim, w->local[i].out_dim)) return false;
  for (int i = 0; i < w->n_summary; ++i) if (!tc_gemm_supported(w->summary[i].in_dim, w->summary[i].out_dim)) return false;
  return tc_gemm_supported(w->local_out_dim, w->merge.out_dim);
}
static bool lite_gemm_ok(const smx_cell_weights* w) {
  if (w->mode != SMX_MODE_LITE || w->n_summary < 2 || w->enc_dim <= 256) return false;
  for (int i = 0; i < w->n_summary; ++i) if (!tc_gemm_supported(w->summary[i].in_dim, w->summary[i].out_dim)) return false;
  return true;
}
static CellLayout cell_layout(const smx_cell_weights* w) {
  CellLayout l{};
  size_t off = 0;
  if (w->mode == SMX_MODE_LITE) {
    for (int i = 0; i < w->n_summary; ++i) { l.summary[i] = off; off += align_up(tc_linear_packed_bytes(w->summary[i].in_dim, w->summary[i].out_dim)); }
    if (lite_gemm_ok(w)) {  // every block but the last (column-sum epilogue) on K-GEMM
      l.gemm = 1;
      for (int i = 0; i < w->n_summary; ++i) { l.summary_d[i] = off; off += align_up((size_t)w->summary[i].in_dim * w->summary[i].out_dim * 2, 1024); }
    }
    l.total = off;
    return l;
  }
  if (w->mode == SMX_MODE_FAST) {  // [global_proj rows 0..Dl (local half)] [rows Dl..2Dl (summary half)] [merge local part]
    l.local[0] = off; off += align_up(tc_linear_packed_bytes(w->enc_dim, w->local_out_dim));
    l.summary[0] = off; off += align_up(tc_linear_packed_bytes(w->enc_dim, w->local_out_dim));
    l.merge = off; off += align_up(tc_linear_packed_bytes(w->local_out_dim, w->merge.out_dim));
    l.total = off;
    return l;
  }
  for (int i = 0; i < w->n_local; ++i) { l.local[i] = off; off += align_up(tc_linear_packed_bytes(w->local[i].in_dim, w->local[i].out_dim)); }
  for (int i = 0; i < w->n_summary; ++i) { l.summary[i] = off; off += align_up(tc_linear_packed_bytes(w->summary[i].in_dim, w->summary[i].out_dim)); }
  l.merge = off; off += align_up(tc_linear_packed_bytes(w->local_out_dim, w->merge.out_dim));
  if (tc_cellf_supported(w)) {
    for (int i = 0; i < 2; ++i) { l.local_s[i] = off; off += align_up(tc_linear_packed_bytes(w->local[i].in_dim, w->local[i].out_dim)); }
    for (int i = 0; i < 2; ++i) { l.summary_s[i] = off; off += align_up(tc_linear_packed_bytes(w->summary[i].in_dim, w->summary[i].out_dim)); }
    l.merge_s = off; off += align_up(tc_linear_packed_bytes(w->local_out_dim, w->merge.out_dim));
    l.v4 = off; off += align_up(tc_cell4_packed_bytes(w), 1024);
  }
  if (cell_gemm_ok(w)) {
    l.gemm = 1;
    for (int i = 0; i < w->n_local; ++i) { l.local_d[i] = off; off += align_up((size_t)w->local[i].in_dim * w->local[i].out_dim * 2, 1024); }
    for (int i = 0; i < w->n_summary; ++i) { l.summary_d[i] = off; off += align_up((size_t)w->summary[i].in_dim * w->summary[i].out_dim * 2, 1024); }
    l.merge_d = off; off += align_up((size_t)w->local_out_dim * w->merge.out_dim * 2, 1024);
  }
  l.total = off;
  return l;
}
size_t tc_cell_packed_bytes(const smx_cell_weights* w) { return tc_cell_supported(w, 0) ? cell_layout(w).total : 0; }

int tc_cell_pack(const smx_cell_weights* w, void* packed, cudaStream_t st) {
  if (!tc_cell_supported(w, 0)) return fail(SMX_ERR_UNSUPPORTED, "cell configuration not handled by the tensor-core arm");
  const CellLayout l = cell_layout(w);
  char* base = (char*)packed;
  if (w->mode == SMX_MODE_LITE) {
    for (int i = 0; i < w->n_summary; ++i) SMX_TRY(tc_pack_linear(w->summary[i], 0, w->summary[i].in_dim, 0, base + l.summary[i], st));
    if (l.gemm) for (int i = 0; i < w->n_summary; ++i) SMX_TRY(tc_dense_bf16(w->summary[i], 0, w->summary[i].in_dim, base + l.summary_d[i], st));
    return SMX_OK;
  }
  if (w->mode == SMX_MODE_FAST) {
    smx_linear half = w->global_proj;  // dense (2 D_l, D): rows [0, D_l) feed the local half, rows [D_l, 2 D_l) the summary half
    half.out_dim = w->local_out_dim;
    SMX_TRY(tc_pack_linear(half, 0, half.in_dim, 0, base + l.local[0], st));
    half.w = w->global_proj.w + (size_t)w->local_out_dim * w->global_proj.in_dim;
    half.b = w->global_proj.b + w->local_out_dim;
    SMX_TRY(tc_pack_linear(half, 0, half.in_dim, 0, base + l.summary[0], st));
    return tc_pack_linear(w->merge, 0, w->local_out_dim, 0, base + l.merge, st);  // W_c[:, :D_l]
  }
  if (tc_cellf_supported(w)) {  // fused persistent cell: 64 x 64 weight blocks
    for (int i = 0; i < 2; ++i) SMX_TRY(tc_pack_linear_nt(w->local[i], 0, w->local[i].in_dim, 64, base + l.local[i], st));
    for (int i = 0; i < 2; ++i) SMX_TRY(tc_pack_linear_nt(w->summary[i], 0, w->summary[i].in_dim, 64, base + l.summary[i], st));
    SMX_TRY(tc_pack_linear_nt(w->merge, 0, w->local_out_dim, 64, base + l.merge, st));
    for (int i = 0; i < 2; ++i) {
      SMX_TRY(tc_cell3_reorder(w->local[i], w->local[i].in_dim, w->local[i].n_split, base + l.local[i], base + l.local_s[i], st));
      SMX_TRY(tc_cell3_reorder(w->summary[i], w->summary[i].in_dim, w->summary[i].n_split, base + l.summary[i], base + l.summary_s[i], st));
    }
    SMX_TRY(tc_cell3_reorder(w->merge, w->local_out_dim, 1, base + l.merge, base + l.merge_s, st));
    if (tc_cell4_packed_bytes(w))
      SMX_TRY(tc_cell4_pack(w, base + l.summary[0], base + l.summary[1], base + l.local[0], base + l.local[1], base + l.merge, base + l.v4, st));
    return SMX_OK;
  }
  for (int i = 0; i < w->n_local; ++i) SMX_TRY(tc_pack_linear(w->local[i], 0, w->local[i].in_dim, 0, base + l.local[i], st));
  for (int i = 0; i < w->n_summary; ++i) SMX_TRY(tc_pack_linear(w->summary[i], 0, w->summary[i].in_dim, 0, base + l.summary[i], st));
  SMX_TRY(tc_pack_linear(w->merge, 0, w->local_out_dim, 0, base + l.merge, st));  // W_c[:, :D_l]
  if (l.gemm) {
    for (int i = 0; i < w->n_local; ++i) SMX_TRY(tc_dense_bf16(w->local[i], 0, w->local[i].in_dim, base + l.local_d[i], st));
    for (int i = 0; i < w->n_summary; ++i) SMX_TRY(tc_dense_bf16(w->summary[i], 0, w->summary[i].in_dim, base + l.summary_d[i], st));
    SMX_TRY(tc_dense_bf16(w->merge, 0, w->local_out_dim, base + l.merge_d, st));  // W_c[:, :D_l]
  }
  return SMX_OK;
}

// The LayerNorm in front of the cell (Conformer.py:520: norm1) folded into the packed image of the one-kernel cell: a no-op for
// configurations that kernel does not take.  Records the folded parameters in the weights struct; the forward uses the folded
// image only when it is called with exactly these parameters.
int tc_cell_pack_prenorm(smx_cell_weights* w, const float* norm_w, const float* norm_b, cudaStream_t st) {
  w->prenorm_w = nullptr; w->prenorm_b = nullptr;
  if (!w->packed || !norm_w || !norm_b || w->mode != SMX_MODE_FULL || !tc_cell_supported(w, 0) || !tc_cellf_supported(w) || !tc_cell4_prenorm_ok(w))
    return SMX_OK;
  const CellLayout l = cell_layout(w);
  SMX_TRY(tc_cell4_pack_prenorm(w, (char*)w->packed + l.v4, norm_w, norm_b, st));
  w->prenorm_w = norm_w; w->prenorm_b = norm_b;
  return SMX_OK;
}

// per-utterance finalisation: mean over time, LayerNorm, and the summary's share of the combiner
//   c[b] = W_c[:, D_l:] @ LN_s( sum_t s[b,t] / sum_t mask[b,t] ) + b_c          summary_mixing.py:229-231,248-253
__global__ void __launch_bounds__(256) cell_finalize_kernel(const float* __restrict__ colsum, int tiles_per_utt, int T,
                                                            const uint8_t* __restrict__ mask, int Ds, int Dl, int Dout,
                                                            const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                            const float* __restrict__ Wc, const float* __restrict__ bc,
                                                            float* __restrict__ rowbias) {
  __shared__ float mu[1024];
  __shared__ float red[8];
  __shared__ float stat[2];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // number of valid frames (an integer-valued float, like torch.sum(mask) in the reference)
  float cnt = 0.0f;
  if (mask) {
    for (int t = tid; t < T; t += 256) cnt += (float)mask[(size_t)b * T + t];
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) red[warp] = cnt;
    __syncthreads();
    cnt = 0.0f;
    for (int i = 0; i < 8; ++i) cnt += red[i];
    __syncthreads();
  } else {
    cnt = (float)T;
  }
  for (int d = tid; d < Ds; d += 256) {
    float s = 0.0f;
    for (int i = 0; i < tiles_per_utt; ++i) s += colsum[((size_t)b * tiles_per_utt + i) * Ds + d];  // fixed order
    mu[d] = s / cnt;
  }
  __syncthreads();
  if (ln_w) {
    float s = 0.0f;
    for (int d = tid; d < Ds; d += 256) s += mu[d];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) { float t = 0.0f; for (int i = 0; i < 8; ++i) t += red[i]; stat[0] = t / (float)Ds; }
    __syncthreads();
    const float mean = stat[0];
    float q = 0.0f;
    for (int d = tid; d < Ds; d += 256) { float e = mu[d] - mean; q += e * e; }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    __syncthreads();
    if (lane == 0) red[warp] = q;
    __syncthreads();
    if (tid == 0) { float t = 0.0f; for (int i = 0; i < 8; ++i) t += red[i]; stat[1] = rsqrtf(t / (float)Ds + 1e-5f); }
    __syncthreads();
    const float rstd = stat[1];
    for (int d = tid; d < Ds; d += 256) mu[d] = (mu[d] - mean) * rstd * ln_w[d] + ln_b[d];
    __syncthreads();
  }
  const int ldw = Dl + Ds;
  // one warp per output, lanes along k: coalesced weight rows; the outputs are split over gridDim.y blocks per utterance (each
  // repeats the cheap mean / LayerNorm above): 32 blocks alone left the 512 x 512 GEMV at 111 us
  for (int n = warp + 8 * blockIdx.y; n < Dout; n += 8 * gridDim.y) {
    const float* wr = Wc + (size_t)n * ldw + Dl;
    float acc = 0.0f;
    for (int k = lane; k < Ds; k += 32) acc = fmaf(wr[k], mu[k], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) rowbias[(size_t)b * Dout + n] = acc + bc[n];
  }
}

// lite mode: y[b] = sum_t s[b,t] mask[b,t] / sum_t mask[b,t] from the per-tile column sums (fixed order)   summary_mixing.py:318-322
__global__ void __launch_bounds__(256) cell_mean_kernel(const float* __restrict__ colsum, int tiles_per_utt, int T,
                                                        const uint8_t* __restrict__ mask, int Ds, __nv_bfloat16* __restrict__ y) {
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float cnt = (float)T;
  if (mask) {
    cnt = 0.0f;
    for (int t = tid; t < T; t += 256) cnt += (float)mask[(size_t)b * T + t];
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) red[warp] = cnt;
    __syncthreads();
    cnt = 0.0f;
    for (int i = 0; i < 8; ++i) cnt += red[i];
  }
  for (int d = tid; d < Ds; d += 256) {
    float s = 0.0f;
    for (int i = 0; i < tiles_per_utt; ++i) s += colsum[((size_t)b * tiles_per_utt + i) * Ds + d];
    y[(size_t)b * Ds + d] = __float2bfloat16(s / cnt);
  }
}

// out[r, :] = a[r, :] + s[r / T, :]  (bf16; the `x + skip` of a lite cell, whose output is one row per utterance)
__global__ void __launch_bounds__(256) add_bcast_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ s, int64_t n8,
                                                             int T, int D8, __nv_bfloat16* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const int64_t row = i / D8;
  const int c8 = (int)(i - row * D8);
  const uint4 av = reinterpret_cast<const uint4*>(a)[i];
  const uint4 sv = reinterpret_cast<const uint4*>(s)[(row / T) * D8 + c8];
  const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&av);
  const __nv_bfloat162* sh = reinterpret_cast<const __nv_bfloat162*>(&sv);
  uint32_t o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 x = __bfloat1622float2(ah[e]), y = __bfloat1622float2(sh[e]);
    o[e] = tc::pack_bf16x2(x.x + y.x, x.y + y.y);
  }
  reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}
int tc_add_bcast(const __nv_bfloat16* a, const __nv_bfloat16* s, int64_t rows, int T, int D, __nv_bfloat16* out, cudaStream_t st) {
  if (D % 8) return fail(SMX_ERR_UNSUPPORTED, "add_bcast: D=%d", D);
  const int64_t n8 = rows * (D / 8);
  add_bcast_bf16_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(a, s, n8, T, D / 8, out);
  count_launch();
  return check_launch("add_bcast_bf16_kernel");
}

// per-tile column sums of S (B, T, Ds) bf16 (already masked): out[b][tile][d] = sum over the tile's frames, in frame order
// (the K-GEMM arm of the unfused cells: K-LIN's fused column-sum epilogue cost 165 us for a 512 x 512 block, K-GEMM + this ~40)
__global__ void __launch_bounds__(256) colsum_tiles_kernel(const __nv_bfloat16* __restrict__ S, int T, int Ds, int tpu, float* __restrict__ out) {
  const int b = blockIdx.y, tile = blockIdx.x, t0 = tile * 128;
  const int nrows = T - t0 < 128 ? T - t0 : 128;
  for (int d2 = threadIdx.x; d2 < Ds / 2; d2 += 256) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(S + ((size_t)b * T + t0) * Ds) + d2;
    float2 acc = make_float2(0.0f, 0.0f);
    for (int r = 0; r < nrows; ++r) { const float2 v = __bfloat1622float2(p[(size_t)r * (Ds / 2)]); acc.x += v.x; acc.y += v.y; }
    *reinterpret_cast<float2*>(out + ((size_t)b * tpu + tile) * Ds + 2 * d2) = acc;
  }
}
static int colsum_tiles(const __nv_bfloat16* S, int B, int T, int Ds, float* out, cudaStream_t st) {
  const int tpu = (T + 127) / 128;
  colsum_tiles_kernel<<<dim3(tpu, B), 256, 0, st>>>(S, T, Ds, tpu, out);
  count_launch();
  return check_launch("colsum_tiles_kernel");
}

struct CellWs {
  __nv_bfloat16 *h, *L, *xn;
  float *colsum, *rowbias;
};
static int cell_ws(const smx_cell_weights* w, int B, int T, Arena& ws, CellWs& o) {
  const int64_t rows = (int64_t)B * T;
  if (w->mode == SMX_MODE_LITE || w->mode == SMX_MODE_FAST) {
    const int tpu = (T + 127) / 128;
    int maxh = 1;
    for (int i = 0; i + 1 < w->n_summary; ++i) maxh = w->summary[i].out_dim > maxh ? w->summary[i].out_dim : maxh;
    if (lite_gemm_ok(w)) maxh = w->summary_out_dim > maxh ? w->summary_out_dim : maxh;  // S itself makes a round trip on the K-GEMM arm
    const int Dsum = w->mode == SMX_MODE_LITE ? w->summary_out_dim : w->local_out_dim;
    o.h = (__nv_bfloat16*)ws.take((size_t)rows * maxh * 2 * 2);
    o.L = (__nv_bfloat16*)ws.take(w->mode == SMX_MODE_FAST ? (size_t)rows * w->local_out_dim * 2 : 16);
    o.colsum = ws.f32((size_t)B * tpu * Dsum);
    o.rowbias = ws.f32((size_t)B * (w->mode == SMX_MODE_FAST ? w->merge.out_dim : 1));
    o.xn = lite_gemm_ok(w) ? (__nv_bfloat16*)ws.take((size_t)rows * w->enc_dim * 2) : nullptr;
    if (!o.h || !o.L || !o.colsum || !o.rowbias || (lite_gemm_ok(w) && !o.xn)) return fail(SMX_ERR_WORKSPACE, "workspace too small (tc cell lite/fast)");
    return SMX_OK;
  }
  int maxh = 0;
  for (int i = 0; i + 1 < w->n_local; ++i) maxh = w->local[i].out_dim > maxh ? w->local[i].out_dim : maxh;
  for (int i = 0; i + 1 < w->n_summary; ++i) maxh = w->summary[i].out_dim > maxh ? w->summary[i].out_dim : maxh;
  if (cell_gemm_ok(w)) maxh = w->summary_out_dim > maxh ? w->summary_out_dim : maxh;  // S itself makes a round trip on the K-GEMM arm
  const int tpu = (T + 127) / 128;
  // two hidden buffers (ping-pong through the MLP chain) + L
  o.h = (__nv_bfloat16*)ws.take((size_t)rows * (maxh > 0 ? maxh : 1) * 2 * 2);
  o.L = (__nv_bfloat16*)ws.take((size_t)rows * w->local_out_dim * 2);
  o.colsum = ws.f32((size_t)B * tpu * w->summary_out_dim);
  o.rowbias = ws.f32((size_t)B * w->merge.out_dim);
  o.xn = cell_gemm_ok(w) ? (__nv_bfloat16*)ws.take((size_t)rows * w->enc_dim * 2) : nullptr;  // norm1(x), shared by both branches
  if (!o.h || !o.L || !o.colsum || !o.rowbias || (cell_gemm_ok(w) && !o.xn)) return fail(SMX_ERR_WORKSPACE, "workspace too small (tc cell)");
  return SMX_OK;
}
size_t tc_cell_workspace_bytes(const smx_cell_weights* w, int B, int T) {
  if (!tc_cell_supported(w, 0)) return 0;
  if (w->mode == SMX_MODE_FULL && tc_cellf_supported(w)) {
    const size_t a = tc_cellf_workspace_bytes(w, B, T), b = tc_cell4_supported(w) ? tc_cell4_workspace_bytes(w, B, T) : 0;
    return a > b ? a : b;
  }
  Arena a(nullptr, 0, true);
  CellWs o;
  cell_ws(w, B, T, a, o);
  return a.peak;
}

static LinP lin_base(int B, int T) {
  LinP p{};
  p.rows = (int64_t)B * T;
  p.T = T;
  p.utt_tiles = 1;
  p.alpha = 1.0f;
  p.ln_eps = 1e-5f;
  p.oln_eps = 1e-5f;
  p.act = SMX_ACT_IDENTITY;
  return p;
}
static void lin_weight(LinP& p, const smx_linear& L, const void* packed, int K) {
  p.K = K; p.N = L.out_dim; p.wp = (const __nv_bfloat16*)packed; p.bias = L.b;
  if (L.n_split > 1) { p.head_in = L.in_dim / L.n_split; p.head_out = L.out_dim / L.n_split; }
  else { p.head_in = p.head_out = 0; }
}

int tc_cell_fwd(const smx_cell_weights* w, const void* packed, int B, int T, const __nv_bfloat16* x,
                const float* pre_ln_w, const float* pre_ln_b, const uint8_t* mask, const __nv_bfloat16* residual,
                __nv_bfloat16* y, Arena& ws, cudaStream_t st) {
  const CellLayout l = cell_layout(w);
  const char* pk = (const char*)packed;
  const size_t m0 = ws.mark();
  if (ws.dry) { ws.take(tc_cell_workspace_bytes(w, B, T)); ws.release(m0); return SMX_OK; }
  if (w->mode == SMX_MODE_LITE || w->mode == SMX_MODE_FAST) {
    const bool lite = w->mode == SMX_MODE_LITE;
    if (lite && residual) return fail(SMX_ERR_BAD_ARG, "lite mode returns (B,D_s): no fused residual");
    CellWs o;
    SMX_TRY(cell_ws(w, B, T, ws, o));
    const int64_t rows = (int64_t)B * T;
    const int Dl = w->local_out_dim, Ds = w->summary_out_dim, tpu = (T + 127) / 128;
    if (lite) {
      // s(x) * mask -> per-tile column sums (fused epilogue of the last block) -> mean over valid frames        :318-322
      int maxh = 1;
      for (int i = 0; i + 1 < w->n_summary; ++i) maxh = w->summary[i].out_dim > maxh ? w->summary[i].out_dim : maxh;
      if (l.gemm) maxh = Ds > maxh ? Ds : maxh;
      __nv_bfloat16* hb[2] = {o.h, o.h + (size_t)rows * maxh};
      const __nv_bfloat16* cur = x; int64_t ld = w->enc_dim;
      bool ln_in_kernel = true;
      if (l.gemm && pre_ln_w) {  // LayerNorm as its own pass, then K-GEMM for every block but the last
        SMX_TRY(layernorm(x, SMX_BF16, w->enc_dim, pre_ln_w, pre_ln_b, 1e-5f, SMX_ACT_IDENTITY, o.xn, SMX_BF16, w->enc_dim, rows, w->enc_dim, st));
        cur = o.xn;
        ln_in_kernel = false;
      }
      for (int i = 0; i < w->n_summary; ++i) {
        const bool last = (i == w->n_summary - 1);
        if (l.gemm) {
          GemmTc g{};
          g.a = cur; g.lda = ld; g.M = rows; g.N = w->summary[i].out_dim; g.K = w->summary[i].in_dim;
          g.w = (const __nv_bfloat16*)(pk + l.summary_d[i]); g.bias = w->summary[i].b; g.act = w->act; g.alpha = 1.0f;
          g.rowmask = last ? mask : nullptr;
          g.out = hb[i & 1]; g.ldo = g.N;
          if (w->summary[i].n_split > 1) { g.bd_in = g.K / w->summary[i].n_split; g.bd_out = g.N / w->summary[i].n_split; }  // skip the zero blocks
          SMX_TRY(tc_gemm_launch(g, st));
          cur = g.out; ld = g.ldo;
          if (last) SMX_TRY(colsum_tiles(cur, B, T, Ds, o.colsum, st));
          continue;
        }
        LinP p = lin_base(B, T);
        p.x = cur; p.ldx = ld;
        lin_weight(p, w->summary[i], pk + l.summary[i], w->summary[i].in_dim);
        p.act = w->act;
        if (i == 0 && ln_in_kernel) { p.ln_w = pre_ln_w; p.ln_b = pre_ln_b; }
        if (last) {
          p.rowmask = mask; p.colsum = o.colsum;
          SMX_TRY(tc_linear_launch(p, TC_LIN_COLSUM, st));
        } else {
          p.out = hb[i & 1]; p.ldo = w->summary[i].out_dim;
          SMX_TRY(tc_linear_launch(p, TC_LIN_PLAIN, st));
          cur = p.out; ld = p.ldo;
        }
      }
      cell_mean_kernel<<<B, 256, 0, st>>>(o.colsum, tpu, T, mask, Ds, y);
      count_launch();
      SMX_TRY(check_launch("cell_mean_kernel"));
    } else {
      // G = act(W_g x + b_g) * mask; local = G[:, :D_l] (bf16), summary half -> per-tile column sums             :271-281
      smx_linear half = w->global_proj;
      half.out_dim = Dl;
      {
        LinP p = lin_base(B, T);
        p.x = x; p.ldx = w->enc_dim;
        lin_weight(p, half, pk + l.local[0], w->enc_dim);
        p.ln_w = pre_ln_w; p.ln_b = pre_ln_b;
        p.act = w->act; p.rowmask = mask;
        p.out = o.L; p.ldo = Dl;
        SMX_TRY(tc_linear_launch(p, TC_LIN_PLAIN, st));
      }
      {
        half.b = w->global_proj.b + Dl;
        LinP p = lin_base(B, T);
        p.x = x; p.ldx = w->enc_dim;
        lin_weight(p, half, pk + l.summary[0], w->enc_dim);
        p.ln_w = pre_ln_w; p.ln_b = pre_ln_b;
        p.act = w->act; p.rowmask = mask; p.colsum = o.colsum;
        SMX_TRY(tc_linear_launch(p, TC_LIN_COLSUM, st));
      }
      // mean (no LayerNorm in fast mode) and the summary's share of the combiner: c[b] = W_c[:, D_l:] mean + b_c   :278-281, 296-298
      cell_finalize_kernel<<<dim3(B, 8), 256, 0, st>>>(o.colsum, tpu, T, mask, Dl, Dl, w->merge.out_dim, nullptr, nullptr, w->merge.w, w->merge.b, o.rowbias);
      count_launch();
      SMX_TRY(check_launch("cell_finalize_kernel"));
      LinP p = lin_base(B, T);
      p.x = o.L; p.ldx = Dl;
      lin_weight(p, w->merge, pk + l.merge, Dl);
      p.bias = nullptr;
      p.head_in = p.head_out = 0;
      p.rowbias = o.rowbias; p.rowbias_ld = w->merge.out_dim;
      p.act = w->act;
      p.resid = residual; p.ldr = w->merge.out_dim;
      p.out = y; p.ldo = w->merge.out_dim;
      SMX_TRY(tc_linear_launch(p, TC_LIN_PLAIN, st));
    }
    ws.release(m0);
    return SMX_OK;
  }
  if (tc_cellf_supported(w)) {
    // v3 moves rows with 256-bit global accesses: 32-byte aligned x / y / residual
    const bool al = ((uintptr_t)x % 32 == 0) && ((uintptr_t)y % 32 == 0) && ((uintptr_t)residual % 32 == 0);
    if (tc_cell_version() == 4 && tc_cell4_packed_bytes(w) && tc_cell4_fits(B, T) && al)
      return tc_cell4_fwd(w, pk + l.v4, B, T, x, pre_ln_w, pre_ln_b, mask, residual, y, ws, st);
    if (tc_cell_version() >= 3 && tc_cell3_supported(w) && al)
      return tc_cell3_fwd(w, pk + l.summary_s[0], pk + l.summary_s[1], pk + l.local_s[0], pk + l.local_s[1], pk + l.merge_s, B, T, x,
                          pre_ln_w, pre_ln_b, mask, residual, y, ws, st);
    return tc_cellf_fwd(w, pk + l.summary[0], pk + l.summary[1], pk + l.local[0], pk + l.local[1], pk + l.merge, B, T, x,
                        pre_ln_w, pre_ln_b, mask, residual, y, ws, st);
  }
  CellWs o;
  SMX_TRY(cell_ws(w, B, T, ws, o));
  const int64_t rows = (int64_t)B * T;
  int maxh = 1;
  for (int i = 0; i + 1 < w->n_local; ++i) maxh = w->local[i].out_dim > maxh ? w->local[i].out_dim : maxh;
  for (int i = 0; i + 1 < w->n_summary; ++i) maxh = w->summary[i].out_dim > maxh ? w->summary[i].out_dim : maxh;
  if (l.gemm) maxh = w->summary_out_dim > maxh ? w->summary_out_dim : maxh;
  __nv_bfloat16* hb[2] = {o.h, o.h + (size_t)rows * maxh};
  const int Dl = w->local_out_dim, Ds = w->summary_out_dim, Dout = w->merge.out_dim;
  const bool use_gemm = l.gemm != 0;
  auto gemm_lin = [&](const smx_linear& L, const void* wd, int K, const __nv_bfloat16* a, int64_t lda, int act_code, const uint8_t* rowmask,
                      const float* rowbias, const __nv_bfloat16* resid, bool with_bias, __nv_bfloat16* out) -> int {
    GemmTc g{};
    g.a = a; g.lda = lda; g.M = rows; g.N = L.out_dim; g.K = K;
    g.w = (const __nv_bfloat16*)wd;
    g.bias = with_bias ? L.b : nullptr;
    g.rowbias = rowbias; g.rowbias_ld = L.out_dim; g.rows_per_group = T;
    g.act = act_code; g.rowmask = rowmask;
    g.resid = resid; g.ldr = L.out_dim; g.alpha = 1.0f;
    g.out = out; g.ldo = L.out_dim;
    if (L.n_split > 1 && K == L.in_dim) { g.bd_in = L.in_dim / L.n_split; g.bd_out = L.out_dim / L.n_split; }  // skip the zero blocks
    return tc_gemm_launch(g, st);
  };
  const __nv_bfloat16* x_in = x;      // input of the first block of both branches
  bool ln_in_kernel = true;           // ... normalised by the K-LIN prologue (else: already normalised)
  if (use_gemm && pre_ln_w) {
    SMX_TRY(layernorm(x, SMX_BF16, w->enc_dim, pre_ln_w, pre_ln_b, 1e-5f, SMX_ACT_IDENTITY, o.xn, SMX_BF16, w->enc_dim, rows, w->enc_dim, st));
    x_in = o.xn;
    ln_in_kernel = false;
  }

  // ---- s(): summary branch -> masked column sums per tile                               :221, 229-231
  {
    const __nv_bfloat16* cur = x_in; int64_t ld = w->enc_dim;
    for (int i = 0; i < w->n_summary; ++i) {
      const bool last = (i == w->n_summary - 1);
      if (use_gemm) {
        SMX_TRY(gemm_lin(w->summary[i], pk + l.summary_d[i], w->summary[i].in_dim, cur, ld, w->act, last ? mask : nullptr, nullptr, nullptr, true, hb[i & 1]));
        cur = hb[i & 1]; ld = w->summary[i].out_dim;
        if (last) SMX_TRY(colsum_tiles(cur, B, T, Ds, o.colsum, st));
        continue;
      }
      LinP p = lin_base(B, T);
      p.x = cur; p.ldx = ld;
      lin_weight(p, w->summary[i], pk + l.summary[i], w->summary[i].in_dim);
      p.act = w->act;
      if (i == 0 && ln_in_kernel) { p.ln_w = pre_ln_w; p.ln_b = pre_ln_b; }
      if (last) {
        p.rowmask = mask; p.colsum = o.colsum;
        SMX_TRY(tc_linear_launch(p, TC_LIN_COLSUM, st));
      } else {
        p.out = hb[i & 1]; p.ldo = w->summary[i].out_dim;
        SMX_TRY(tc_linear_launch(p, TC_LIN_PLAIN, st));
        cur = p.out; ld = p.ldo;
      }
    }
  }
  // ---- mean, LN_s, summary share of the combiner                                        :229-233, 248-253
  cell_finalize_kernel<<<dim3(B, 8), 256, 0, st>>>(o.colsum, (T + 127) / 128, T, mask, Ds, Dl, Dout,
                                          w->use_layernorm ? w->summary_norm_w : nullptr,
                                          w->use_layernorm ? w->summary_norm_b : nullptr, w->merge.w, w->merge.b, o.rowbias);
  count_launch();
  SMX_TRY(check_launch("cell_finalize_kernel"));
  // ---- f(): local branch -> mask -> LN_l                                                :215-218
  {
    const __nv_bfloat16* cur = x_in; int64_t ld = w->enc_dim;
    for (int i = 0; i < w->n_local; ++i) {
      const bool last = (i == w->n_local - 1);
      if (use_gemm) {
        __nv_bfloat16* dst = last ? o.L : hb[i & 1];
        SMX_TRY(gemm_lin(w->local[i], pk + l.local_d[i], w->local[i].in_dim, cur, ld, w->act, last ? mask : nullptr, nullptr, nullptr, true, dst));
        if (last && w->use_layernorm)
          SMX_TRY(layernorm(o.L, SMX_BF16, Dl, w->local_norm_w, w->local_norm_b, 1e-5f, SMX_ACT_IDENTITY, o.L, SMX_BF16, Dl, rows, Dl, st));
        cur = dst; ld = w->local[i].out_dim;
        continue;
      }
      LinP p = lin_base(B, T);
      p.x = cur; p.ldx = ld;
      lin_weight(p, w->local[i], pk + l.local[i], w->local[i].in_dim);
      p.act = w->act;
      if (i == 0 && ln_in_kernel) { p.ln_w = pre_ln_w; p.ln_b = pre_ln_b; }
      if (last) {
        p.rowmask = mask; p.out = o.L; p.ldo = Dl;
        if (w->use_layernorm && Dl <= 256) {
          p.oln_w = w->local_norm_w; p.oln_b = w->local_norm_b;
          SMX_TRY(tc_linear_launch(p, TC_LIN_OLN, st));
        } else {
          SMX_TRY(tc_linear_launch(p, TC_LIN_PLAIN, st));
          if (w->use_layernorm)
            SMX_TRY(layernorm(o.L, SMX_BF16, Dl, w->local_norm_w, w->local_norm_b, 1e-5f, SMX_ACT_IDENTITY, o.L, SMX_BF16, Dl, rows, Dl, st));
        }
      } else {
        p.out = hb[i & 1]; p.ldo = w->local[i].out_dim;
        SMX_TRY(tc_linear_launch(p, TC_LIN_PLAIN, st));
        cur = p.out; ld = p.ldo;
      }
    }
  }
  // ---- combiner: y = act(W_c[:, :D_l] @ local + c[b]) (+ residual)                         :251-253
  if (use_gemm) {
    smx_linear mc = w->merge;  // (b_c is inside c[b])
    SMX_TRY(gemm_lin(mc, pk + l.merge_d, Dl, o.L, Dl, w->act, nullptr, o.rowbias, residual, false, y));
  } else {
    LinP p = lin_base(B, T);
    p.x = o.L; p.ldx = Dl;
    lin_weight(p, w->merge, pk + l.merge, Dl);
    p.bias = nullptr;  // b_c is inside c[b]
    p.head_in = p.head_out = 0;
    p.rowbias = o.rowbias; p.rowbias_ld = Dout;
    p.act = w->act;
    p.resid = residual; p.ldr = Dout;
    p.out = y; p.ldo = Dout;
    SMX_TRY(tc_linear_launch(p, TC_LIN_PLAIN, st));
  }
  ws.release(m0);
  return SMX_OK;
}

// =============================================================================================
// ConvolutionModule                                                         Conformer.py:322-338
// =============================================================================================
bool tc_convmod_supported(const smx_convmod_weights* w, int chunk) {
  const int D = w->bottleneck.in_dim;
  if (chunk > 0 || w->causal) return false;
  if (w->bottleneck.out_dim != 2 * D || w->out.in_dim != D || w->out.out_dim != D) return false;
  if (!tc_linear_supported(D, 2 * D) || !tc_linear_supported(D, D)) return false;
  if (w->kernel_size < 1 || w->kernel_size > 63 || w->kernel_size % 2 == 0) return false;
  if (!w->bottleneck.b || !w->out.b || !w->dw_w) return false;
  return true;
}
// unfused path (D = 512): LayerNorm pass + K-GEMM with the GLU epilogue instead of K-LIN (which stages a whole activation tile and has
// no operand pipelining: 102 us against 15 + ~40 us at 32 000 x 512)
static bool convmod_glu_gemm(const smx_convmod_weights* w) {
  const int D = w->bottleneck.in_dim;
  return !tc_convf_supported(w, 0) && w->bottleneck.n_split <= 1 && (2 * D) % 256 == 0 && tc_gemm_supported(D, 2 * D) && w->ln_w && w->ln_b;
}
size_t tc_convmod_packed_bytes(const smx_convmod_weights* w) {
  if (!tc_convmod_supported(w, 0)) return 0;
  const int D = w->bottleneck.in_dim;
  return 2 * align_up(tc_linear_packed_bytes(D, 2 * D)) + 2 * align_up(tc_linear_packed_bytes(D, D))  // [GLU][out][GLU, out in schedule order]
         + (tc_convf_supported(w, 0) ? tc_glu4_packed_bytes(w) : 0)                                     // [K-GLU v4 image]
         + (convmod_glu_gemm(w) ? tc_glu_dense_bytes(D, 2 * D) : 0);                                    // [unfused path: GLU image for K-GEMM]
}
static size_t convmod_glu4_offset(int D) { return 2 * align_up(tc_linear_packed_bytes(D, 2 * D)) + 2 * align_up(tc_linear_packed_bytes(D, D)); }
int tc_convmod_pack(const smx_convmod_weights* w, void* packed, cudaStream_t st) {
  if (!tc_convmod_supported(w, 0)) return fail(SMX_ERR_UNSUPPORTED, "conv module not handled by the tensor-core arm");
  const int D = w->bottleneck.in_dim;
  if (tc_convf_supported(w, 0)) {  // fused path: 64 x 64 blocks, value/gate blocks interleaved for the GLU pass
    SMX_TRY(tc_pack_linear_nt(w->bottleneck, 0, D, 64, packed, st, 1));
    SMX_TRY(tc_pack_linear_nt(w->out, 0, D, 64, (char*)packed + align_up(tc_linear_packed_bytes(D, 2 * D)), st));
    char* sched = (char*)packed + align_up(tc_linear_packed_bytes(D, 2 * D)) + align_up(tc_linear_packed_bytes(D, D));
    SMX_TRY(tc_cell3_reorder(w->bottleneck, D, 1, packed, sched, st));
    SMX_TRY(tc_cell3_reorder(w->out, D, 1, (char*)packed + align_up(tc_linear_packed_bytes(D, 2 * D)),
                             sched + align_up(tc_linear_packed_bytes(D, 2 * D)), st));
    if (tc_glu4_supported(w)) SMX_TRY(tc_glu4_pack(w, (char*)packed + convmod_glu4_offset(D), st));
    return SMX_OK;
  }
  SMX_TRY(tc_pack_linear(w->bottleneck, 0, D, 1, packed, st));
  SMX_TRY(tc_pack_linear(w->out, 0, D, 0, (char*)packed + align_up(tc_linear_packed_bytes(D, 2 * D)), st));
  if (tc_gemm_supported(D, D))  // dense copy of the output linear for K-GEMM (in the area the fused path uses for its schedule-order images)
    SMX_TRY(tc_dense_bf16(w->out, 0, D, (char*)packed + align_up(tc_linear_packed_bytes(D, 2 * D)) + align_up(tc_linear_packed_bytes(D, D)), st));
  if (convmod_glu_gemm(w)) SMX_TRY(tc_glu_dense_bf16(w->bottleneck, (char*)packed + convmod_glu4_offset(D), st));
  return SMX_OK;
}
size_t tc_convmod_workspace_bytes(const smx_convmod_weights* w, int B, int T) {
  const int D = w->bottleneck.in_dim;
  return 2 * align_up((size_t)B * T * D * 2);
}

// depthwise conv over time (zero padding at utterance edges) + LayerNorm over channels + activation.
// One warp per output frame; a lane owns 8 channels (two chunks when D > 256).
constexpr int DW_FRAMES = 64;  // frames per block
__global__ void __launch_bounds__(256) dwconv_ln_act_kernel(const __nv_bfloat16* __restrict__ g, const float* __restrict__ dw_w,
                                                            const float* __restrict__ dw_b, const float* __restrict__ ln_w,
                                                            const float* __restrict__ ln_b, int act, int T, int D, int k,
                                                            __nv_bfloat16* __restrict__ out) {
  extern __shared__ float wT[];  // [k][D] transposed taps
  const int b = blockIdx.y, t0 = blockIdx.x * DW_FRAMES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < k * D; i += 256) {
    int c = i / k, j = i % k;
    wT[j * D + c] = dw_w[i];
  }
  __syncthreads();
  const int pad = (k - 1) / 2;
  const int nch = D / 8;  // 16-byte chunks per row
  for (int f = warp; f < DW_FRAMES; f += 8) {
    const int t = t0 + f;
    if (t >= T) break;
    float acc[2][8];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int ck = lane + 32 * c;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[c][e] = (ck < nch && dw_b) ? dw_b[ck * 8 + e] : 0.0f;
    }
    for (int j = 0; j < k; ++j) {
      const int u = t + j - pad;
      if (u < 0 || u >= T) continue;
      const __nv_bfloat16* row = g + ((size_t)b * T + u) * D;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ck = lane + 32 * c;
        if (ck < nch) {
          const uint4 raw = *reinterpret_cast<const uint4*>(row + ck * 8);
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
          const float4 w0 = *reinterpret_cast<const float4*>(wT + j * D + ck * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wT + j * D + ck * 8 + 4);
          float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]), f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
          acc[c][0] = fmaf(w0.x, f0.x, acc[c][0]); acc[c][1] = fmaf(w0.y, f0.y, acc[c][1]);
          acc[c][2] = fmaf(w0.z, f1.x, acc[c][2]); acc[c][3] = fmaf(w0.w, f1.y, acc[c][3]);
          acc[c][4] = fmaf(w1.x, f2.x, acc[c][4]); acc[c][5] = fmaf(w1.y, f2.y, acc[c][5]);
          acc[c][6] = fmaf(w1.z, f3.x, acc[c][6]); acc[c][7] = fmaf(w1.w, f3.y, acc[c][7]);
        }
      }
    }
    // LayerNorm over the D channels of this frame (the warp holds the whole row)
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < 2; ++c)
      if (lane + 32 * c < nch) {
#pragma unroll
        for (int e = 0; e < 8; ++e) s += acc[c][e];
      }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)D;
    float q = 0.0f;
#pragma unroll
    for (int c = 0; c < 2; ++c)
      if (lane + 32 * c < nch) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { float d = acc[c][e] - mean; q += d * d; }
      }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)D + 1e-5f);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int ck = lane + 32 * c;
      if (ck < nch) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = tc::act_fast(act, (acc[c][e] - mean) * rstd * ln_w[ck * 8 + e] + ln_b[ck * 8 + e]);
        *reinterpret_cast<uint4*>(out + ((size_t)b * T + t) * D + ck * 8) =
            make_uint4(tc::pack_bf16x2(o[0], o[1]), tc::pack_bf16x2(o[2], o[3]), tc::pack_bf16x2(o[4], o[5]), tc::pack_bf16x2(o[6], o[7]));
      }
    }
  }
}

// Tiled variant for kernel size K and D % 128 == 0, D <= 512.  A block takes DWT_FRAMES output frames of one utterance:
// the (DWT_FRAMES + K - 1) input frames are staged once in shared memory; phase 1 gives every thread one channel pair and
// blocks of eight consecutive frames (taps and accumulators in registers: 8*K*2 FMAs per K+7 shared-memory loads);
// phase 2 normalises each frame over the channels (one warp per frame) and applies the activation.
constexpr int DWT_FRAMES = 24;   // 104 KB of shared memory at D = 512: two blocks per SM (one stages while the other computes)
template <int K>
__global__ void __launch_bounds__(256, 2) dwconv_tiled_kernel(const __nv_bfloat16* __restrict__ g, const float* __restrict__ dw_w,
                                                              const float* __restrict__ dw_b, const float* __restrict__ ln_w,
                                                              const float* __restrict__ ln_b, int act, int B, int T, int D,
                                                              __nv_bfloat16* __restrict__ out) {
  constexpr int PAD = (K - 1) / 2, NIN = DWT_FRAMES + K - 1;
  extern __shared__ __align__(16) uint8_t dsm[];
  __nv_bfloat16* sIn = reinterpret_cast<__nv_bfloat16*>(dsm);                          // [NIN][D]
  float* sOut = reinterpret_cast<float*>(dsm + (size_t)NIN * D * sizeof(__nv_bfloat16));  // [DWT_FRAMES][D]
  const int tiles_t = (T + DWT_FRAMES - 1) / DWT_FRAMES, n_tiles = tiles_t * B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cpr = D / 8;
  const int pairs = D / 2, pr = tid % pairs, fb0 = tid / pairs, nfbp = 256 / pairs;
  const int c0 = pr * 2;
  float w0[K], w1[K];
  {
    // the D x K taps pass through the (still unused) shared memory with coalesced loads: read per thread straight from global
    // memory -- rows 4 K bytes apart between lanes -- they throttle the load/store queue (profiles/r02_notes.md)
    float* sTap = reinterpret_cast<float*>(dsm);
    for (int i = tid; i < D * K; i += 256) sTap[i] = dw_w[i];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < K; ++j) { w0[j] = sTap[c0 * K + j]; w1[j] = sTap[(c0 + 1) * K + j]; }
    __syncthreads();  // every thread holds its taps before the input rows overwrite them
  }
  // persistent over the (utterance, frame tile) pairs: the taps (D x K x 4 bytes per block) are fetched once per block, not per tile
#pragma unroll 1
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  const int b = tile / tiles_t, t0 = (tile - b * tiles_t) * DWT_FRAMES;
  // four 16-byte chunks per thread and round, all four loads issued before the first is stored (a load-then-store loop is a chain of
  // L2 round trips: ~14 per tile at D = 512)
  for (int base = 0; base < NIN * cpr; base += 4 * 256) {
    uint4 raw[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = base + k * 256 + tid;
      const int row = idx / cpr, ch = idx - row * cpr, u = t0 - PAD + row;
      raw[k] = make_uint4(0, 0, 0, 0);  // zero padding at the utterance edges (Conformer.py:142-151)
      if (idx < NIN * cpr && u >= 0 && u < T) raw[k] = *reinterpret_cast<const uint4*>(g + ((size_t)b * T + u) * D + ch * 8);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = base + k * 256 + tid;
      if (idx < NIN * cpr) *reinterpret_cast<uint4*>(sIn + (size_t)idx * 8) = raw[k];   // (row * D + ch * 8 == idx * 8)
    }
  }
  __syncthreads();
  {
    const float b0 = dw_b ? dw_b[c0] : 0.0f, b1 = dw_b ? dw_b[c0 + 1] : 0.0f;
    for (int fb = fb0; fb < DWT_FRAMES / 8; fb += nfbp) {
      float a0[8], a1[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) { a0[o] = b0; a1[o] = b1; }
#pragma unroll
      for (int i = 0; i < K + 7; ++i) {
        const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sIn + (size_t)(fb * 8 + i) * D + c0));
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const int j = i - o;
          if (j >= 0 && j < K) { a0[o] = fmaf(w0[j], x.x, a0[o]); a1[o] = fmaf(w1[j], x.y, a1[o]); }
        }
      }
#pragma unroll
      for (int o = 0; o < 8; ++o) *reinterpret_cast<float2*>(sOut + (size_t)(fb * 8 + o) * D + c0) = make_float2(a0[o], a1[o]);
    }
  }
  __syncthreads();
  const int nseg = D / 128;  // a lane owns 4 channels in each 128-channel segment: conflict-free 16-byte accesses
  const float invD = 1.0f / (float)D;
  for (int f = warp; f < DWT_FRAMES; f += 8) {
    const int t = t0 + f;
    if (t >= T) break;
    float v[4][4];
    float s = 0.0f;
#pragma unroll
    for (int sg = 0; sg < 4; ++sg)
      if (sg < nseg) {
        const float4 x = *reinterpret_cast<const float4*>(sOut + (size_t)f * D + sg * 128 + lane * 4);
        v[sg][0] = x.x; v[sg][1] = x.y; v[sg][2] = x.z; v[sg][3] = x.w;
        s += (x.x + x.y) + (x.z + x.w);
      }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * invD;
    float q = 0.0f;
#pragma unroll
    for (int sg = 0; sg < 4; ++sg)
      if (sg < nseg) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float d = v[sg][e] - mean; q = fmaf(d, d, q); }
      }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * invD + 1e-5f);
#pragma unroll
    for (int sg = 0; sg < 4; ++sg)
      if (sg < nseg) {
        const int c = sg * 128 + lane * 4;
        const float4 lw = *reinterpret_cast<const float4*>(ln_w + c), lb = *reinterpret_cast<const float4*>(ln_b + c);
        float o[4] = {(v[sg][0] - mean) * rstd * lw.x + lb.x, (v[sg][1] - mean) * rstd * lw.y + lb.y,
                      (v[sg][2] - mean) * rstd * lw.z + lb.z, (v[sg][3] - mean) * rstd * lw.w + lb.w};
        tc::act_apply<4>(act, o);
        *reinterpret_cast<uint2*>(out + ((size_t)b * T + t) * D + c) = make_uint2(tc::pack_bf16x2(o[0], o[1]), tc::pack_bf16x2(o[2], o[3]));
      }
  }
  }  // tile loop (the next tile's staging barrier also orders this tile's reads of sOut before its overwrite)
}

int tc_dwconv_ln_act(const __nv_bfloat16* g, const float* dw_w, const float* dw_b, const float* ln_w,
                     const float* ln_b, int act, int B, int T, int D, int k, __nv_bfloat16* out, cudaStream_t st) {
  if (D % 8 || D > 512) return fail(SMX_ERR_UNSUPPORTED, "dwconv: D=%d", D);
  if (k == 31 && D % 128 == 0) {
    size_t smem = (size_t)(DWT_FRAMES + 30) * D * 2 + (size_t)DWT_FRAMES * D * 4;
    if (smem < (size_t)D * 31 * 4) smem = (size_t)D * 31 * 4;   // (the taps pass through it first)
    cudaError_t e = cudaFuncSetAttribute(dwconv_tiled_kernel<31>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(dwconv_tiled): %s", cudaGetErrorString(e));
    const int n_tiles = ((T + DWT_FRAMES - 1) / DWT_FRAMES) * B;
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const unsigned grid = (unsigned)(n_tiles < 2 * sms ? n_tiles : 2 * sms);   // two resident blocks per SM
    dwconv_tiled_kernel<31><<<grid, 256, smem, st>>>(g, dw_w, dw_b, ln_w, ln_b, act, B, T, D, out);
    count_launch();
    return check_launch("dwconv_tiled_kernel");
  }
  const size_t smem = (size_t)k * D * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(dwconv_ln_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(dwconv): %s", cudaGetErrorString(e));
  dim3 grid((T + DW_FRAMES - 1) / DW_FRAMES, B);
  dwconv_ln_act_kernel<<<grid, 256, smem, st>>>(g, dw_w, dw_b, ln_w, ln_b, act, T, D, k, out);
  count_launch();
  return check_launch("dwconv_ln_act_kernel");
}

int tc_convmod_fwd(const smx_convmod_weights* w, const void* packed, int act, int B, int T, const __nv_bfloat16* x,
                   const uint8_t* mask, const __nv_bfloat16* residual, __nv_bfloat16* y, Arena& ws, cudaStream_t st) {
  const int D = w->bottleneck.in_dim;
  const int64_t rows = (int64_t)B * T;
  const size_t m0 = ws.mark();
  if (ws.dry) { ws.take(tc_convmod_workspace_bytes(w, B, T)); ws.release(m0); return SMX_OK; }
  if (tc_convf_supported(w, 0)) {
    __nv_bfloat16* gb = (__nv_bfloat16*)ws.take((size_t)rows * D * 2);
    if (!gb) return fail(SMX_ERR_WORKSPACE, "workspace too small (tc conv module)");
    if (tc_cell_version() >= 4 && tc_glu4_supported(w) && tc_glu4_fits(rows) && ((uintptr_t)x % 32 == 0) && ((uintptr_t)gb % 32 == 0))  // :322-324
      SMX_TRY(tc_glu4_fwd(w, (const char*)packed + convmod_glu4_offset(D), rows, x, gb, st));
    else if (tc_cell_version() >= 3 && ((uintptr_t)x % 32 == 0) && ((uintptr_t)gb % 32 == 0))
      SMX_TRY(tc_glu3_fwd(w->bottleneck, (const char*)packed + align_up(tc_linear_packed_bytes(D, 2 * D)) + align_up(tc_linear_packed_bytes(D, D)),
                          w->ln_w, w->ln_b, rows, x, gb, st));
    else
      SMX_TRY(tc_glu_fwd(w->bottleneck, packed, w->ln_w, w->ln_b, rows, x, gb, st));
    SMX_TRY(tc_convf_second_half(w, (const char*)packed + 2 * align_up(tc_linear_packed_bytes(D, 2 * D)) + align_up(tc_linear_packed_bytes(D, D)),
                                 act, B, T, gb, mask, residual, y, st));                                   // :325-338, :543
    ws.release(m0);
    return SMX_OK;
  }
  __nv_bfloat16* g = (__nv_bfloat16*)ws.take((size_t)rows * D * 2);
  __nv_bfloat16* c = (__nv_bfloat16*)ws.take((size_t)rows * D * 2);
  if (!g || !c) return fail(SMX_ERR_WORKSPACE, "workspace too small (tc conv module)");
  if (convmod_glu_gemm(w) && ((uintptr_t)x % 16) == 0) {  // LN pass (into c, free until the depthwise stage) -> K-GEMM with the GLU epilogue   :322-324
    SMX_TRY(layernorm(x, SMX_BF16, D, w->ln_w, w->ln_b, 1e-5f, SMX_ACT_IDENTITY, c, SMX_BF16, D, rows, D, st));
    const char* img = (const char*)packed + convmod_glu4_offset(D);
    GemmTc gm{};
    gm.a = c; gm.lda = D; gm.M = rows; gm.N = 2 * D; gm.K = D;
    gm.w = (const __nv_bfloat16*)img; gm.bias = (const float*)(img + align_up((size_t)2 * D * D * 2, 1024));
    gm.act = SMX_ACT_IDENTITY; gm.alpha = 1.0f; gm.glu = 1; gm.out = g; gm.ldo = D;
    SMX_TRY(tc_gemm_launch(gm, st));
  } else {  // LN -> pointwise conv (D -> 2D) -> GLU                                       :322-324
    LinP p = lin_base(B, T);
    p.utt_tiles = 0;
    p.x = x; p.ldx = D;
    lin_weight(p, w->bottleneck, packed, D);
    p.ln_w = w->ln_w; p.ln_b = w->ln_b;
    p.out = g; p.ldo = D;
    SMX_TRY(tc_linear_launch(p, TC_LIN_GLU, st));
  }
  // depthwise conv -> LN -> act                                                           :325, :331-332
  SMX_TRY(tc_dwconv_ln_act(g, w->dw_w, w->dw_b, w->after_ln_w, w->after_ln_b, act, B, T, D, w->kernel_size, c, st));
  if (tc_gemm_supported(D, D)) {  // Linear D -> D, * mask, + residual on K-GEMM                            :332-338, :543
    GemmTc gm{};
    gm.a = c; gm.lda = D; gm.M = rows; gm.N = D; gm.K = D;
    gm.w = (const __nv_bfloat16*)((const char*)packed + align_up(tc_linear_packed_bytes(D, 2 * D)) + align_up(tc_linear_packed_bytes(D, D)));
    gm.bias = w->out.b; gm.act = SMX_ACT_IDENTITY; gm.rowmask = mask; gm.resid = residual; gm.ldr = D; gm.alpha = 1.0f; gm.out = y; gm.ldo = D;
    SMX_TRY(tc_gemm_launch(gm, st));
  } else {  // Linear D -> D, * mask, + residual                                           :332-338, :543
    LinP p = lin_base(B, T);
    p.utt_tiles = 0;
    p.x = c; p.ldx = D;
    lin_weight(p, w->out, (const char*)packed + align_up(tc_linear_packed_bytes(D, 2 * D)), D);
    p.rowmask = mask;
    p.resid = residual; p.ldr = D;
    p.out = y; p.ldo = D;
    SMX_TRY(tc_linear_launch(p, TC_LIN_PLAIN, st));
  }
  ws.release(m0);
  return SMX_OK;
}

}  // namespace smx
