// tcgen05 arm of libsmx, part 12: BranchformerEncoderLayer.forward with attention_type "SummaryMixing" (Branchformer.py:243-334)
// on the tensor cores, bf16 activations / fp32 accumulation:
//
//   x1 = SummaryMixing(norm_mhsa(x))                      the cell's tensor-core arm (LayerNorm as GEMM prologue)        :317-322
//   u  = act(norm_conv(x) W_pre^T + b)                    K-LIN with the LayerNorm prologue, D -> U                      :292, 86-90
//   g  = gate_act(dwconv_reflect(LN(u[:, U/2:]))) * u[:, :U/2]          fused CSGU gate (smx_tc_gemm.cu)                 :91-94
//   x2 = g W_post^T + b                                   K-GEMM, U/2 -> D                                               :95-96
//   y  = x + merge_proj([x1 ; x2])                        K-GEMMs; first block: lite -> x1 is one row per utterance, so its
//                                                         share W[:, :D_s] x1 + b becomes a per-utterance row bias         :279, 220-226
#include "smx_tc.h"
#include "smx_tc_common.cuh"

namespace smx {

struct BranchLayout {
  size_t pre, post, m0a, m0b, m[SMX_MAX_BLOCKS], total;
};
static int branch_dx1(const smx_branchformer_layer_weights* w) {
  return w->cell.mode == SMX_MODE_LITE ? w->cell.summary_out_dim : w->cell.merge.out_dim;
}
static BranchLayout branch_layout(const smx_branchformer_layer_weights* w) {
  BranchLayout l{};
  const smx_convbranch_weights& br = w->branch;
  const int D = br.pre.in_dim, U = br.pre.out_dim, H = U / 2, Dx1 = branch_dx1(w);
  size_t off = 0;
  l.pre = off; off += align_up((size_t)U * D * 2, 1024);   // dense (U, D) bf16 for K-GEMM
  l.post = off; off += align_up((size_t)D * H * 2, 1024);
  l.m0a = off; off += align_up((size_t)w->merge[0].out_dim * Dx1 * 2, 1024);   // first merge block, columns of x1 (kept for the full cell: dense cat GEMM uses m[0])
  l.m0b = off; off += align_up((size_t)w->merge[0].out_dim * D * 2, 1024);     // first merge block, columns of x2
  for (int i = 0; i < w->n_merge; ++i) { l.m[i] = off; off += align_up((size_t)w->merge[i].out_dim * w->merge[i].in_dim * 2, 1024); }
  l.total = off;
  return l;
}

bool tc_branchformer_supported(const smx_branchformer_layer_weights* w, int has_sum_mask) {
  const smx_convbranch_weights& br = w->branch;
  const int D = br.pre.in_dim, U = br.pre.out_dim, H = U / 2;
  if (U % 2 || br.post.in_dim != H || br.post.out_dim != D) return false;
  if (!br.pre.w || !br.pre.b || !br.post.w || !br.post.b || br.pre.n_split > 1 || br.post.n_split > 1) return false;
  if (br.csgu_linear.w) return false;                        // use_linear_after_conv: not on this arm
  if (br.kernel_size != 31 || H % 64) return false;
  if (!tc_gemm_supported(D, U) || !tc_gemm_supported(H, D)) return false;
  if (!w->cell.packed || !tc_cell_supported(&w->cell, has_sum_mask)) return false;
  if (w->cell.mode != SMX_MODE_LITE && w->cell.mode != SMX_MODE_FULL && w->cell.mode != SMX_MODE_FAST) return false;
  const int Dx1 = branch_dx1(w);
  if (w->n_merge < 1 || w->n_merge > SMX_MAX_BLOCKS || w->merge[0].in_dim != Dx1 + D) return false;
  if (w->merge[w->n_merge - 1].out_dim != D) return false;
  for (int i = 0; i < w->n_merge; ++i) {
    const smx_linear& L = w->merge[i];
    if (!L.w || !L.b || L.n_split > 1 || !tc_gemm_supported(L.in_dim, L.out_dim)) return false;
    if (i > 0 && L.in_dim != w->merge[i - 1].out_dim) return false;
  }
  if (Dx1 % 64 || D % 64) return false;
  return true;
}
size_t tc_branchformer_packed_bytes(const smx_branchformer_layer_weights* w) {
  return tc_branchformer_supported(w, 0) ? branch_layout(w).total : 0;
}
int tc_branchformer_pack(const smx_branchformer_layer_weights* w, void* packed, cudaStream_t st) {
  if (!tc_branchformer_supported(w, 0)) return fail(SMX_ERR_UNSUPPORTED, "branchformer layer not handled by the tensor-core arm");
  const BranchLayout l = branch_layout(w);
  const smx_convbranch_weights& br = w->branch;
  const int D = br.pre.in_dim, Dx1 = branch_dx1(w);
  char* base = (char*)packed;
  SMX_TRY(tc_dense_bf16(br.pre, 0, D, base + l.pre, st));
  SMX_TRY(tc_dense_bf16(br.post, 0, br.post.in_dim, base + l.post, st));
  SMX_TRY(tc_dense_bf16(w->merge[0], 0, Dx1, base + l.m0a, st));
  SMX_TRY(tc_dense_bf16(w->merge[0], Dx1, D, base + l.m0b, st));
  for (int i = 0; i < w->n_merge; ++i) SMX_TRY(tc_dense_bf16(w->merge[i], 0, w->merge[i].in_dim, base + l.m[i], st));
  return SMX_OK;
}

// rowbias[b][n] = sum_k W[n][k] x1[b][k] + bias[n]  (lite: the first merge block's share of the per-utterance summary;
// W is the fp32 (N, ldw) weight, its first Dx1 columns)
__global__ void __launch_bounds__(256) branch_rowbias_kernel(const __nv_bfloat16* __restrict__ x1, int Dx1, const float* __restrict__ W, int ldw,
                                                             const float* __restrict__ bias, int N, float* __restrict__ rowbias) {
  __shared__ float sx[1024];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < Dx1; k += 256) sx[k] = __bfloat162float(x1[(size_t)b * Dx1 + k]);
  __syncthreads();
  for (int n = warp + 8 * blockIdx.y; n < N; n += 8 * gridDim.y) {  // (outputs split over gridDim.y blocks per utterance)
    const float* wr = W + (size_t)n * ldw;
    float acc = 0.0f;
    for (int k = lane; k < Dx1; k += 32) acc = fmaf(wr[k], sx[k], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) rowbias[(size_t)b * N + n] = acc + bias[n];
  }
}

int tc_branchformer_layer_fwd(const smx_branchformer_layer_weights* w, int B, int T, const __nv_bfloat16* x, const uint8_t* mask,
                              __nv_bfloat16* y, Arena& ws, cudaStream_t st) {
  const smx_convbranch_weights& br = w->branch;
  const int D = br.pre.in_dim, U = br.pre.out_dim, H = U / 2, Dx1 = branch_dx1(w), Dcat = Dx1 + D;
  const int64_t rows = (int64_t)B * T;
  const bool lite = w->cell.mode == SMX_MODE_LITE;
  const BranchLayout l = branch_layout(w);
  const char* pk = (const char*)w->packed;
  const size_t m0 = ws.mark();
  int maxm = D;
  for (int i = 0; i < w->n_merge; ++i) maxm = w->merge[i].out_dim > maxm ? w->merge[i].out_dim : maxm;
  __nv_bfloat16* u = (__nv_bfloat16*)ws.take((size_t)rows * U * 2);
  __nv_bfloat16* g = (__nv_bfloat16*)ws.take((size_t)rows * H * 2);
  __nv_bfloat16* cat = (__nv_bfloat16*)ws.take((size_t)rows * (lite ? D : Dcat) * 2);   // lite: x2 only
  __nv_bfloat16* hb = (__nv_bfloat16*)ws.take((size_t)rows * maxm * 2 * 2);              // ping-pong of the merge MLP
  __nv_bfloat16* x1 = (__nv_bfloat16*)ws.take(lite ? (size_t)B * Dx1 * 2 : (size_t)rows * Dx1 * 2);
  float* rowbias = ws.f32((size_t)B * w->merge[0].out_dim);
  void* stats = ws.take(tc_csgu_workspace_bytes(rows));
  if (!u || !g || !cat || !hb || !x1 || !rowbias || !stats) return fail(SMX_ERR_WORKSPACE, "workspace too small (tc branchformer layer)");
  // branch 1: the cell on norm_mhsa(x)                                                          :317-322
  SMX_TRY(tc_cell_fwd(&w->cell, w->cell.packed, B, T, x, w->norm_mhsa_w, w->norm_mhsa_b, mask, nullptr, x1, ws, st));
  if (ws.dry) { ws.release(m0); return SMX_OK; }
  // branch 2: u = act(LN_conv(x) W_pre^T + b)   (no mask on this branch, :276)                    :292, 86-90
  // (LayerNorm as its own pass into the idle merge buffer, then K-GEMM: the K-LIN kernel with its LayerNorm prologue took 543 us
  // for this D -> 3072 projection at B=32, T=1000, K-GEMM takes ~100)
  {
    __nv_bfloat16* xn = hb;
    SMX_TRY(layernorm(x, SMX_BF16, D, w->norm_conv_w, w->norm_conv_b, 1e-5f, SMX_ACT_IDENTITY, xn, SMX_BF16, D, rows, D, st));
    GemmTc gp{};
    gp.a = xn; gp.lda = D; gp.M = rows; gp.N = U; gp.K = D; gp.w = (const __nv_bfloat16*)(pk + l.pre); gp.bias = br.pre.b;
    gp.act = br.act; gp.alpha = 1.0f; gp.out = u; gp.ldo = U;
    SMX_TRY(tc_gemm_launch(gp, st));
  }
  SMX_TRY(tc_csgu_fwd(u, B, T, H, br.csgu_ln_w, br.csgu_ln_b, br.csgu_dw_w, br.csgu_dw_b, br.kernel_size, br.gate_act, g, H, stats, st));
  __nv_bfloat16* x2 = lite ? cat : cat + Dx1;
  const int64_t ldx2 = lite ? D : Dcat;
  {
    GemmTc gm{};
    gm.a = g; gm.lda = H; gm.M = rows; gm.N = D; gm.K = H; gm.w = (const __nv_bfloat16*)(pk + l.post); gm.bias = br.post.b;
    gm.act = SMX_ACT_IDENTITY; gm.alpha = 1.0f; gm.out = x2; gm.ldo = ldx2;
    SMX_TRY(tc_gemm_launch(gm, st));
  }
  // y = x + merge_proj([x1 ; x2]): activation after every block (VanillaNN.py:196)                :279
  const __nv_bfloat16* cur; int64_t ldc; int Kc;
  GemmTc gm{};
  gm.M = rows; gm.alpha = 1.0f; gm.act = w->act;
  if (lite) {
    branch_rowbias_kernel<<<dim3(B, 8), 256, 0, st>>>(x1, Dx1, w->merge[0].w, w->merge[0].in_dim, w->merge[0].b, w->merge[0].out_dim, rowbias);
    count_launch();
    SMX_TRY(check_launch("branch_rowbias_kernel"));
    gm.a = x2; gm.lda = ldx2; gm.K = D; gm.w = (const __nv_bfloat16*)(pk + l.m0b);
    gm.rowbias = rowbias; gm.rowbias_ld = w->merge[0].out_dim; gm.rows_per_group = T;
  } else {
    cudaError_t e = cudaMemcpy2DAsync(cat, (size_t)Dcat * 2, x1, (size_t)Dx1 * 2, (size_t)Dx1 * 2, (size_t)rows, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaMemcpy2DAsync: %s", cudaGetErrorString(e));
    gm.a = cat; gm.lda = Dcat; gm.K = Dcat; gm.w = (const __nv_bfloat16*)(pk + l.m[0]); gm.bias = w->merge[0].b;
  }
  for (int i = 0; i < w->n_merge; ++i) {
    const bool last = i == w->n_merge - 1;
    if (i > 0) {
      gm.a = cur; gm.lda = ldc; gm.K = Kc; gm.w = (const __nv_bfloat16*)(pk + l.m[i]); gm.bias = w->merge[i].b;
      gm.rowbias = nullptr;
    }
    gm.N = w->merge[i].out_dim;
    if (last) { gm.resid = x; gm.ldr = D; gm.out = y; gm.ldo = D; }
    else { gm.out = hb + (size_t)(i & 1) * rows * maxm; gm.ldo = gm.N; }
    SMX_TRY(tc_gemm_launch(gm, st));
    cur = gm.out; ldc = gm.ldo; Kc = gm.N;
  }
  ws.release(m0);
  return SMX_OK;
}

}  // namespace smx
