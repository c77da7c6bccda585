// Host orchestration of the generic (fp32-math) arm: composes the kernels of smx_simt.cu into the
// reference's module forwards.  Intermediates live in the caller's workspace as fp32; the external
// x / y / residual tensors carry their own dtype tag.  Every function also runs "dry" (ws.dry) to size
// the workspace: same control flow, no launches.
#include "smx_internal.h"
#include "smx_tc.h"

namespace smx {

static GemmP base_gemm() {
  GemmP p{};
  p.alpha = 1.0f;
  p.rowbias_div = 1;
  p.batches = 1;
  p.act = SMX_ACT_IDENTITY;
  return p;
}

struct LinOpts {
  int act = SMX_ACT_IDENTITY;
  const uint8_t* rowmask = nullptr;
  const float* rowbias = nullptr; int64_t rowbias_ld = 0; int rowbias_div = 1;
  const void* residual = nullptr; int r_dt = SMX_F32; int64_t ldr = 0; float alpha = 1.0f;
  bool use_bias = true;
  int k_offset = 0;  // dense only: use input columns [k_offset, k_offset+K) of the weight
  int K = -1;        // dense only: reduced K
  void* scratch = nullptr; size_t scratch_bytes = 0;  // split_scratch(): lets large fp32 linears run on the tensor cores (split-bf16)
};

// Scratch for the split-bf16 tensor-core form of the largest of `ls` at `rows` rows (0 bytes when none qualifies); taken from
// the arena by every module function, so the sizing (dry) runs account for it.
static void* split_scratch(Arena& ws, int64_t rows, std::initializer_list<const smx_linear*> ls, size_t& bytes) {
  bytes = 0;
  for (const smx_linear* l : ls)
    if (l && l->w && tc_split3_ok(rows, l->in_dim, l->out_dim)) {
      const size_t b = tc_split3_scratch_bytes(rows, l->in_dim, l->out_dim);
      bytes = b > bytes ? b : bytes;
    }
  return bytes ? ws.take(bytes) : nullptr;
}

// C = epilogue(A @ L) for one smx_linear (dense nn.Linear or block-diagonal ParallelLinear, VanillaNN.py:99-117)
static int linear(const smx_linear& L, const void* A, int a_dt, int64_t lda, int64_t rows, void* C, int c_dt,
                  int64_t ldc, const LinOpts& o, cudaStream_t st) {
  if (!L.w) return fail(SMX_ERR_BAD_ARG, "linear: NULL weight");
  if (rows > 0x7fffffff) return fail(SMX_ERR_UNSUPPORTED, "linear: more than 2^31 rows");
  // fp32 activations, large row counts: the tensor cores with split-bf16 operands (three bf16 MMAs per fp32 product, ~1e-5
  // relative to the fp32 result; smx_tc_gemm.cu) instead of the CUDA-core GEMM
  if (o.scratch && a_dt == SMX_F32 && tc_f32_tc_enabled()) {
    const bool dense = L.n_split <= 1;
    const int K = dense ? (o.K > 0 ? o.K : L.in_dim - o.k_offset) : L.in_dim;
    const bool split_ok = dense || (L.in_dim % L.n_split == 0 && L.out_dim % L.n_split == 0 && !o.k_offset && o.K <= 0);
    const bool al = lda % 4 == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)C % 32) == 0 && ldc % (c_dt == SMX_F32 ? 8 : 16) == 0 &&
                    (!o.residual || (((uintptr_t)o.residual % 32) == 0 && o.ldr % (o.r_dt == SMX_F32 ? 4 : 16) == 0));
    if (split_ok && al && tc_split3_ok(rows, K, L.out_dim) && o.scratch_bytes >= tc_split3_scratch_bytes(rows, K, L.out_dim)) {
      GemmTc g{};
      g.bias = (o.use_bias && L.b) ? L.b : nullptr;
      g.rowbias = o.rowbias; g.rowbias_ld = o.rowbias_ld; g.rows_per_group = o.rowbias_div;
      g.act = o.act; g.rowmask = o.rowmask; g.alpha = o.alpha; g.ldr = o.ldr;
      if (o.residual) { if (o.r_dt == SMX_F32) g.resid_f32 = (const float*)o.residual; else g.resid = (const __nv_bfloat16*)o.residual; }
      if (c_dt == SMX_F32) g.out_f32 = (float*)C; else g.out = (__nv_bfloat16*)C;
      g.ldo = ldc;
      return tc_linear_split3(L, dense ? o.k_offset : 0, K, (const float*)A, lda, rows, g, o.scratch, st);
    }
  }
  GemmP p = base_gemm();
  p.A = A; p.a_dtype = a_dt; p.lda = lda;
  p.C = C; p.c_dtype = c_dt; p.ldc = ldc;
  p.M = (int)rows; p.act = o.act; p.rowmask = o.rowmask;
  p.rowbias = o.rowbias; p.rowbias_ld = o.rowbias_ld; p.rowbias_div = o.rowbias_div;
  p.residual = o.residual; p.r_dtype = o.r_dt; p.ldr = o.ldr; p.alpha = o.alpha;
  if (L.n_split <= 1) {
    p.K = o.K > 0 ? o.K : L.in_dim - o.k_offset;
    p.N = L.out_dim;
    p.W = L.w + o.k_offset; p.w_sk = 1; p.w_sn = L.in_dim;
    p.bias = (o.use_bias && L.b) ? L.b : nullptr;
  } else {
    const int h = L.n_split;
    if (L.in_dim % h || L.out_dim % h)
      return fail(SMX_ERR_BAD_ARG, "input_size and n_neurons must be dividible by n_split!");
    if (o.k_offset || o.K > 0) return fail(SMX_ERR_BAD_ARG, "linear: k-slicing a split linear");
    p.K = L.in_dim / h; p.N = L.out_dim / h; p.batches = h;
    p.a_bs = p.K; p.c_bs = p.N;
    p.W = L.w; p.w_sk = p.N; p.w_sn = 1; p.w_bs = (int64_t)p.K * p.N;
    p.bias = (o.use_bias && L.b) ? L.b : nullptr; p.bias_bs = p.N;
  }
  return gemm(p, st);
}

// VanillaNN.forward: blocks x (linear, act), act after EVERY block (VanillaNN.py:168-196)
int vanilla_generic(const smx_linear* blocks, int n, int act, const void* x, int x_dt, int64_t ldx, int64_t rows,
                    const uint8_t* rowmask, const void* residual, int r_dt, int64_t ldr, void* y, int y_dt,
                    int64_t ldy, Arena& ws, cudaStream_t st) {
  if (n < 1 || n > SMX_MAX_BLOCKS)
    return fail(SMX_ERR_UNSUPPORTED, "VanillaNN with %d blocks (library handles 1..%d)", n, SMX_MAX_BLOCKS);
  const size_t m0 = ws.mark();
  size_t sc_bytes = 0;
  void* sc = split_scratch(ws, rows, {&blocks[0], n > 1 ? &blocks[1] : nullptr, n > 2 ? &blocks[2] : nullptr, n > 3 ? &blocks[3] : nullptr}, sc_bytes);
  const void* cur = x; int cur_dt = x_dt; int64_t cur_ld = ldx;
  if (x_dt != SMX_F32 && sc && tc_f32_tc_enabled() && ldx == blocks[0].in_dim && tc_split3_ok(rows, blocks[0].in_dim, blocks[0].out_dim)) {
    // bf16 rows: one conversion pass puts the first block on the split-bf16 tensor-core GEMM too (the CUDA-core GEMM reads bf16 directly)
    float* xf = ws.f32((size_t)rows * blocks[0].in_dim);
    if (!xf) return fail(SMX_ERR_WORKSPACE, "workspace too small (VanillaNN)");
    if (!ws.dry) SMX_TRY(convert(x, x_dt, xf, SMX_F32, rows * blocks[0].in_dim, st));
    cur = xf; cur_dt = SMX_F32;
  }
  for (int i = 0; i < n; ++i) {
    const bool last = (i == n - 1);
    if (i > 0 && blocks[i].in_dim != blocks[i - 1].out_dim)
      return fail(SMX_ERR_BAD_ARG, "VanillaNN block %d: in_dim %d != previous out_dim %d", i, blocks[i].in_dim,
                  blocks[i - 1].out_dim);
    void* out; int out_dt; int64_t out_ld;
    if (last) { out = y; out_dt = y_dt; out_ld = ldy; }
    else {
      out = ws.f32((size_t)rows * blocks[i].out_dim); out_dt = SMX_F32; out_ld = blocks[i].out_dim;
      if (!out) return fail(SMX_ERR_WORKSPACE, "workspace too small (VanillaNN)");
    }
    LinOpts o; o.act = act; o.scratch = sc; o.scratch_bytes = sc_bytes;
    if (last) { o.rowmask = rowmask; o.residual = residual; o.r_dt = r_dt; o.ldr = ldr; }
    if (!ws.dry) SMX_TRY(linear(blocks[i], cur, cur_dt, cur_ld, rows, out, out_dt, out_ld, o, st));
    cur = out; cur_dt = out_dt; cur_ld = out_ld;
  }
  ws.release(m0);
  return SMX_OK;
}

// ---------------------------------------------------------------------------------------------
// SummaryMixing cell                                             summary_mixing.py:169-324
// ---------------------------------------------------------------------------------------------
int cell_generic(const smx_cell_weights* w, int B, int T, const void* x, int x_dt, const uint8_t* mask,
                 const float* sum_mask, const void* residual, int r_dt, void* y, int y_dt, int64_t ldy, Arena& ws,
                 cudaStream_t st) {
  const int64_t rows = (int64_t)B * T;
  const int D = w->enc_dim, Dl = w->local_out_dim, Ds = w->summary_out_dim;
  const int mode = w->mode;
  if (mode < SMX_MODE_FULL || mode > SMX_MODE_EXPDECAY)
    return fail(SMX_ERR_BAD_ARG,
                "The SummaryMixing mode should either be 'SummaryMixing', 'SummaryMixing-lite', "
                "'SummaryMixing-fast' or 'SummaryMixing-expdecay'");
  const size_t m0 = ws.mark();

  if (mode == SMX_MODE_LITE) {  // summary_mixing.py:300-324 (sum_mask ignored, no LN, no combiner)
    if (residual) return fail(SMX_ERR_BAD_ARG, "lite mode returns (B,D_s): no fused residual");
    float* S = ws.f32((size_t)rows * Ds);
    if (!S) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell lite)");
    SMX_TRY(vanilla_generic(w->summary, w->n_summary, w->act, x, x_dt, D, rows, mask, nullptr, 0, 0, S, SMX_F32, Ds, ws, st));
    if (!ws.dry) SMX_TRY(masked_mean(S, Ds, mask, B, T, Ds, y, y_dt, st));
    ws.release(m0);
    return SMX_OK;
  }

  const float* local; int64_t ld_local;     // (rows, Dl) fp32, masked (+LN)
  const float* S; int64_t ld_S; int Dsum;   // (rows, Dsum) fp32, masked
  const bool use_ln = (mode != SMX_MODE_FAST) && w->use_layernorm;
  if (mode == SMX_MODE_FAST) {  // summary_mixing.py:255-298
    if (w->global_proj.out_dim != 2 * Dl)
      return fail(SMX_ERR_BAD_ARG, "fast mode: global_proj must map to 2*local_proj_out_dim");
    float* G = ws.f32((size_t)rows * 2 * Dl);
    if (!G) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell fast)");
    SMX_TRY(vanilla_generic(&w->global_proj, 1, w->act, x, x_dt, D, rows, mask, nullptr, 0, 0, G, SMX_F32, 2 * Dl, ws, st));
    local = G; ld_local = 2 * Dl; S = G + Dl; ld_S = 2 * Dl; Dsum = Dl;
  } else {  // full / expdecay, summary_mixing.py:198-253
    float* L = ws.f32((size_t)rows * Dl);
    float* Sb = ws.f32((size_t)rows * Ds);
    if (!L || !Sb) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell)");
    SMX_TRY(vanilla_generic(w->local, w->n_local, w->act, x, x_dt, D, rows, mask, nullptr, 0, 0, L, SMX_F32, Dl, ws, st));
    if (use_ln && !ws.dry)
      SMX_TRY(layernorm(L, SMX_F32, Dl, w->local_norm_w, w->local_norm_b, 1e-5f, SMX_ACT_IDENTITY, L, SMX_F32, Dl, rows, Dl, st));
    SMX_TRY(vanilla_generic(w->summary, w->n_summary, w->act, x, x_dt, D, rows, mask, nullptr, 0, 0, Sb, SMX_F32, Ds, ws, st));
    local = L; ld_local = Dl; S = Sb; ld_S = Ds; Dsum = Ds;
  }
  if (w->merge.in_dim != Dl + Dsum || w->merge.n_split > 1)
    return fail(SMX_ERR_BAD_ARG, "summary_local_merging must be dense with in_dim == D_l + D_s (%d vs %d)",
                w->merge.in_dim, Dl + Dsum);
  const int Dout = w->merge.out_dim;

  const bool per_frame = (sum_mask != nullptr) || (mode == SMX_MODE_EXPDECAY);
  float* cbias;  // summary contribution to the combiner pre-activation, + bias
  int cdiv;
  if (!per_frame) {  // summary_mixing.py:226-233
    float* mean = ws.f32((size_t)B * Dsum);
    cbias = ws.f32((size_t)B * Dout);
    if (!mean || !cbias) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell mean)");
    if (!ws.dry) {
      SMX_TRY(masked_mean(S, ld_S, mask, B, T, Dsum, mean, SMX_F32, st));
      if (use_ln)
        SMX_TRY(layernorm(mean, SMX_F32, Dsum, w->summary_norm_w, w->summary_norm_b, 1e-5f, SMX_ACT_IDENTITY, mean, SMX_F32, Dsum, B, Dsum, st));
      LinOpts o; o.k_offset = Dl; o.K = Dsum;  // W_c[:, D_l:] @ mean + b_c
      SMX_TRY(linear(w->merge, mean, SMX_F32, Dsum, B, cbias, SMX_F32, Dout, o, st));
    }
    cdiv = T;
  } else {  // summary_mixing.py:223-224, 235-246: (T,T) weights, per-frame summaries
    const float* Mx = sum_mask;
    if (mode == SMX_MODE_EXPDECAY) {
      float* lap = ws.f32((size_t)T * T);
      if (!lap) return fail(SMX_ERR_WORKSPACE, "workspace too small (laplace)");
      if (!ws.dry) SMX_TRY(laplace(w->decay_constant, sum_mask, T, lap, st));
      Mx = lap;
    }
    float* rs = ws.f32((size_t)T);
    float* Sm = ws.f32((size_t)rows * Dsum);
    cbias = ws.f32((size_t)rows * Dout);
    // Dynamic-chunk masks (TransformerASR.py:85-110) have rows that are single runs of ones: per-frame summaries are then
    // differences of prefix sums over time, O(T D) per utterance.  The structure is checked on the device; any other mask
    // (weights, holes, the Laplace matrix of -expdecay) takes the (T,T) @ (T,D) product below instead.
    const bool try_intervals = (mode != SMX_MODE_EXPDECAY);
    int* iv = try_intervals ? (int*)ws.take((size_t)(2 * T + 1) * sizeof(int)) : nullptr;
    void* pws = try_intervals ? ws.take(interval_means_workspace_bytes(B, T, Dsum)) : nullptr;
    if (!rs || !Sm || !cbias || (try_intervals && (!iv || !pws))) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell sum_mask)");
    if (!ws.dry) {
      int* not_interval = nullptr;
      if (try_intervals) {
        not_interval = iv + 2 * T;
        SMX_TRY(interval_detect(Mx, T, iv, iv + T, not_interval, st));
        SMX_TRY(interval_means(S, ld_S, B, T, Dsum, iv, iv + T, not_interval, Sm, pws, st));
      }
      SMX_TRY(rowsum(Mx, T, T, rs, st));
      GemmP p = base_gemm();  // Sm[b] = (Mx @ S[b]) / rowsum(Mx)   (padding NOT removed from the denominator, :239-246)
      p.A = Mx; p.a_dtype = SMX_F32; p.lda = T; p.a_bs = 0;
      p.W = S; p.w_sk = ld_S; p.w_sn = 1; p.w_bs = (int64_t)T * ld_S;
      p.rowdiv = rs;
      p.C = Sm; p.c_dtype = SMX_F32; p.ldc = Dsum; p.c_bs = (int64_t)T * Dsum;
      p.M = T; p.N = Dsum; p.K = T; p.batches = B;
      p.run_if_nonzero = not_interval;  // (NULL for -expdecay: always)
      SMX_TRY(gemm(p, st));
      if (use_ln)
        SMX_TRY(layernorm(Sm, SMX_F32, Dsum, w->summary_norm_w, w->summary_norm_b, 1e-5f, SMX_ACT_IDENTITY, Sm, SMX_F32, Dsum, rows, Dsum, st));
      LinOpts o; o.k_offset = Dl; o.K = Dsum;
      SMX_TRY(linear(w->merge, Sm, SMX_F32, Dsum, rows, cbias, SMX_F32, Dout, o, st));
    }
    cdiv = 1;
  }
  size_t sc_bytes = 0;
  void* sc = split_scratch(ws, rows, {&w->merge}, sc_bytes);
  if (!ws.dry) {  // y = act(W_c[:, :D_l] @ local + cbias) (+ residual)      summary_mixing.py:251-253
    LinOpts o; o.act = w->act; o.use_bias = false; o.K = Dl; o.scratch = sc; o.scratch_bytes = sc_bytes;
    o.rowbias = cbias; o.rowbias_ld = Dout; o.rowbias_div = cdiv;
    o.residual = residual; o.r_dt = r_dt; o.ldr = Dout;
    SMX_TRY(linear(w->merge, local, SMX_F32, ld_local, rows, y, y_dt, ldy, o, st));
  }
  ws.release(m0);
  return SMX_OK;
}

// ---------------------------------------------------------------------------------------------
// macaron half-step FFN                                         Conformer.py:470-484, 518, 547
// ---------------------------------------------------------------------------------------------
int ffn_generic(const smx_ffn_weights* w, int act, int64_t rows, const void* x, int x_dt, const float* oln_w,
                const float* oln_b, float oln_eps, void* y, int y_dt, Arena& ws, cudaStream_t st) {
  const int D = w->w1.in_dim, F = w->w1.out_dim;
  if (w->w2.in_dim != F || w->w2.out_dim != D) return fail(SMX_ERR_BAD_ARG, "ffn: inconsistent dims");
  const size_t m0 = ws.mark();
  float* t = ws.f32((size_t)rows * D);
  float* h = ws.f32((size_t)rows * F);
  size_t sc_bytes = 0;
  void* sc = split_scratch(ws, rows, {&w->w1, &w->w2}, sc_bytes);
  if (!t || !h) return fail(SMX_ERR_WORKSPACE, "workspace too small (ffn)");
  if (!ws.dry) {
    SMX_TRY(layernorm(x, x_dt, D, w->ln_w, w->ln_b, 1e-5f, SMX_ACT_IDENTITY, t, SMX_F32, D, rows, D, st));
    LinOpts o1; o1.act = act; o1.scratch = sc; o1.scratch_bytes = sc_bytes;
    SMX_TRY(linear(w->w1, t, SMX_F32, D, rows, h, SMX_F32, F, o1, st));
    LinOpts o2; o2.residual = x; o2.r_dt = x_dt; o2.ldr = D; o2.alpha = 0.5f; o2.scratch = sc; o2.scratch_bytes = sc_bytes;
    if (oln_w) {  // t is free again: reuse it for the pre-norm sum
      SMX_TRY(linear(w->w2, h, SMX_F32, F, rows, t, SMX_F32, D, o2, st));
      SMX_TRY(layernorm(t, SMX_F32, D, oln_w, oln_b, oln_eps, SMX_ACT_IDENTITY, y, y_dt, D, rows, D, st));
    } else {
      SMX_TRY(linear(w->w2, h, SMX_F32, F, rows, y, y_dt, D, o2, st));
    }
  }
  ws.release(m0);
  return SMX_OK;
}

// ---------------------------------------------------------------------------------------------
// ConvolutionModule                                                   Conformer.py:166-340
// ---------------------------------------------------------------------------------------------
int convmod_generic(const smx_convmod_weights* w, int act, int B, int T, int chunk, const void* x, int x_dt,
                    const uint8_t* mask, const void* residual, int r_dt, void* y, int y_dt, Arena& ws,
                    cudaStream_t st) {
  const int64_t rows = (int64_t)B * T;
  const int D = w->bottleneck.in_dim;
  if (w->bottleneck.out_dim != 2 * D || w->out.in_dim != D || w->out.out_dim != D)
    return fail(SMX_ERR_BAD_ARG, "conv module: inconsistent dims");
  if (chunk > 0 && w->causal) return fail(SMX_ERR_BAD_ARG, "Chunked convolution not supported with causal padding");
  if (w->kernel_size < 1) return fail(SMX_ERR_BAD_ARG, "conv module: kernel_size < 1");
  const size_t m0 = ws.mark();
  float* t = ws.f32((size_t)rows * D);
  float* p = ws.f32((size_t)rows * 2 * D);
  float* g = ws.f32((size_t)rows * D);
  size_t sc_bytes = 0;
  void* sc = split_scratch(ws, rows, {&w->bottleneck, &w->out}, sc_bytes);
  if (!t || !p || !g) return fail(SMX_ERR_WORKSPACE, "workspace too small (conv module)");
  if (!ws.dry) {
    SMX_TRY(layernorm(x, x_dt, D, w->ln_w, w->ln_b, 1e-5f, SMX_ACT_IDENTITY, t, SMX_F32, D, rows, D, st));
    LinOpts o1; o1.scratch = sc; o1.scratch_bytes = sc_bytes;
    SMX_TRY(linear(w->bottleneck, t, SMX_F32, D, rows, p, SMX_F32, 2 * D, o1, st));
    SMX_TRY(glu(p, rows, D, g, st));
    const int pad_mode = chunk > 0 ? SMX_CONV_CHUNKED : (w->causal ? SMX_CONV_CAUSAL : SMX_CONV_SAME_ZERO);
    SMX_TRY(dwconv(g, D, w->dw_w, w->dw_b, B, T, D, w->kernel_size, pad_mode, chunk, t, D, st));
    SMX_TRY(layernorm(t, SMX_F32, D, w->after_ln_w, w->after_ln_b, 1e-5f, act, t, SMX_F32, D, rows, D, st));
    LinOpts o2; o2.rowmask = mask; o2.residual = residual; o2.r_dt = r_dt; o2.ldr = D;  // out*mask (:338) then x + out
    o2.scratch = sc; o2.scratch_bytes = sc_bytes;
    SMX_TRY(linear(w->out, t, SMX_F32, D, rows, y, y_dt, D, o2, st));
  }
  ws.release(m0);
  return SMX_OK;
}

// ---------------------------------------------------------------------------------------------
// ConformerEncoderLayer                                               Conformer.py:490-548
// fp32: generic arm with fp32 intermediates.  bf16: bf16 intermediates; each module runs on the
// tcgen05 arm when its weights carry a packed image and the configuration is supported, else on the
// generic arm (which reads/writes bf16 through the dtype tags).
// ---------------------------------------------------------------------------------------------
// y = x + SummaryMixing(LayerNorm(x)): the mixing block of the Conformer layer (skip = x; x = norm1(x); x = mha_layer(x);
// x = x + skip, Conformer.py:520-541).  bf16 with a packed cell: the fused tcgen05 kernel(s) with the LayerNorm as prologue
// and the skip as epilogue; otherwise LayerNorm + generic cell.
int mixing_block_generic(const smx_cell_weights* cw, const float* norm_w, const float* norm_b, int dtype, int B, int T, const void* x1,
                         const uint8_t* mask, const float* sum_mask, void* x2, Arena& ws, cudaStream_t st) {
  const int64_t rows = (int64_t)B * T;
  const int D = cw->enc_dim;
  const int idt = dtype;
  if (dtype == SMX_BF16 && cw->packed && tc_cell_supported(cw, sum_mask != nullptr)) {
    if (cw->mode == SMX_MODE_LITE) {  // the cell returns one row per utterance (:318-322); x2 = x1 + that row
      if (cw->summary_out_dim != D) return fail(SMX_ERR_BAD_ARG, "mixing block: lite summary_out_dim != d_model");
      const size_t m1 = ws.mark();
      __nv_bfloat16* mean = (__nv_bfloat16*)ws.take((size_t)B * D * 2);
      if (!mean) return fail(SMX_ERR_WORKSPACE, "workspace too small (mixing block lite)");
      SMX_TRY(tc_cell_fwd(cw, cw->packed, B, T, (const __nv_bfloat16*)x1, norm_w, norm_b, mask, nullptr, mean, ws, st));
      if (!ws.dry) SMX_TRY(tc_add_bcast((const __nv_bfloat16*)x1, mean, rows, T, D, (__nv_bfloat16*)x2, st));
      ws.release(m1);
      return SMX_OK;
    }
    if (cw->merge.out_dim != D) return fail(SMX_ERR_BAD_ARG, "mixing block: cell output dim != d_model");
    return tc_cell_fwd(cw, cw->packed, B, T, (const __nv_bfloat16*)x1, norm_w, norm_b, mask, (const __nv_bfloat16*)x1,
                       (__nv_bfloat16*)x2, ws, st);
  }
  const size_t m1 = ws.mark();
  float* n1 = ws.f32((size_t)rows * D);
  if (!n1) return fail(SMX_ERR_WORKSPACE, "workspace too small (mixing block norm1)");
  if (!ws.dry) SMX_TRY(layernorm(x1, idt, D, norm_w, norm_b, 1e-5f, SMX_ACT_IDENTITY, n1, SMX_F32, D, rows, D, st));
  if (cw->mode == SMX_MODE_LITE) {
    if (cw->summary_out_dim != D) return fail(SMX_ERR_BAD_ARG, "mixing block: lite summary_out_dim != d_model");
    float* mean = ws.f32((size_t)B * D);
    float* x1f = (idt == SMX_F32) ? (float*)x1 : ws.f32((size_t)rows * D);
    float* x2f = (idt == SMX_F32) ? (float*)x2 : ws.f32((size_t)rows * D);
    if (!mean || !x1f || !x2f) return fail(SMX_ERR_WORKSPACE, "workspace too small (mixing block lite)");
    SMX_TRY(cell_generic(cw, B, T, n1, SMX_F32, mask, sum_mask, nullptr, 0, mean, SMX_F32, D, ws, st));
    if (!ws.dry) {
      if (idt != SMX_F32) SMX_TRY(convert(x1, idt, x1f, SMX_F32, rows * D, st));
      SMX_TRY(add_bcast(x1f, mean, rows, T, D, x2f, st));
      if (idt != SMX_F32) SMX_TRY(convert(x2f, SMX_F32, x2, idt, rows * D, st));
    }
  } else {
    if (cw->merge.out_dim != D) return fail(SMX_ERR_BAD_ARG, "mixing block: cell output dim != d_model");
    SMX_TRY(cell_generic(cw, B, T, n1, SMX_F32, mask, sum_mask, x1, idt, x2, idt, D, ws, st));
  }
  ws.release(m1);
  return SMX_OK;
}

int conformer_layer_generic(const smx_conformer_layer_weights* w, int dtype, int B, int T, int chunk, const void* x,
                            const uint8_t* mask, const float* sum_mask, void* y, Arena& ws, cudaStream_t st) {
  const int64_t rows = (int64_t)B * T;
  const int D = w->ffn1.w1.in_dim;
  const int idt = dtype;  // dtype of the inter-module activations
  const size_t es = elem_size(idt);
  const size_t m0 = ws.mark();
  void* x1 = ws.take((size_t)rows * D * es);
  void* x2 = ws.take((size_t)rows * D * es);
  if (!x1 || !x2) return fail(SMX_ERR_WORKSPACE, "workspace too small (conformer layer)");
  const bool bf = (dtype == SMX_BF16);
  // x1 = x + 0.5*ffn1(x)                                                              :518
  if (bf && w->ffn1.packed && tc_ffn_supported(&w->ffn1))
    SMX_TRY(tc_ffn_fwd(&w->ffn1, w->ffn1.packed, w->act, rows, (const __nv_bfloat16*)x, nullptr, nullptr, 0.f, (__nv_bfloat16*)x1, ws, st));
  else
    SMX_TRY(ffn_generic(&w->ffn1, w->act, rows, x, dtype, nullptr, nullptr, 0.f, x1, idt, ws, st));
  // x2 = cell(norm1(x1)) + x1                                                          :520-541
  SMX_TRY(mixing_block_generic(&w->cell, w->norm1_w, w->norm1_b, dtype, B, T, x1, mask, sum_mask, x2, ws, st));
  // x3 = x2 + conv_module(x2)*mask   (into x1, which is dead)                           :543-545
  if (bf && w->conv.packed && tc_convmod_supported(&w->conv, chunk))
    SMX_TRY(tc_convmod_fwd(&w->conv, w->conv.packed, w->act, B, T, (const __nv_bfloat16*)x2, mask, (const __nv_bfloat16*)x2,
                           (__nv_bfloat16*)x1, ws, st));
  else
    SMX_TRY(convmod_generic(&w->conv, w->act, B, T, chunk, x2, idt, mask, x2, idt, x1, idt, ws, st));
  // y = norm2(x3 + 0.5*ffn2(x3))                                                        :547
  if (bf && w->ffn2.packed && tc_ffn_supported(&w->ffn2))
    SMX_TRY(tc_ffn_fwd(&w->ffn2, w->ffn2.packed, w->act, rows, (const __nv_bfloat16*)x1, w->norm2_w, w->norm2_b, 1e-5f, (__nv_bfloat16*)y, ws, st));
  else
    SMX_TRY(ffn_generic(&w->ffn2, w->act, rows, x1, idt, w->norm2_w, w->norm2_b, 1e-5f, y, dtype, ws, st));
  ws.release(m0);
  return SMX_OK;
}

// ---------------------------------------------------------------------------------------------
// BranchformerEncoderLayer                                          Branchformer.py:243-334
// ---------------------------------------------------------------------------------------------
int branchformer_layer_generic(const smx_branchformer_layer_weights* w, int dtype, int B, int T, const void* x,
                               const uint8_t* mask, const float* sum_mask, void* y, Arena& ws, cudaStream_t st) {
  const int64_t rows = (int64_t)B * T;
  const smx_convbranch_weights& br = w->branch;
  const int D = br.pre.in_dim, U = br.pre.out_dim, H = U / 2;
  if (U % 2) return fail(SMX_ERR_BAD_ARG, "Input size must be divisible by 2!");
  if (br.post.in_dim != H || br.post.out_dim != D) return fail(SMX_ERR_BAD_ARG, "convolution branch: inconsistent dims");
  const int Dx1 = (w->cell.mode == SMX_MODE_LITE) ? w->cell.summary_out_dim : w->cell.merge.out_dim;
  const int Dcat = Dx1 + D;
  if (w->n_merge < 1 || w->merge[0].in_dim != Dcat)
    return fail(SMX_ERR_BAD_ARG, "merge_proj expects %d inputs but the branches provide %d", w->n_merge < 1 ? -1 : w->merge[0].in_dim, Dcat);
  if (dtype == SMX_BF16 && w->packed && tc_branchformer_supported(w, sum_mask != nullptr))   // tensor-core arm (smx_tc_branch.cu)
    return tc_branchformer_layer_fwd(w, B, T, (const __nv_bfloat16*)x, mask, (__nv_bfloat16*)y, ws, st);
  const size_t m0 = ws.mark();
  float* n = ws.f32((size_t)rows * D);
  float* cat = ws.f32((size_t)rows * Dcat);
  if (!n || !cat) return fail(SMX_ERR_WORKSPACE, "workspace too small (branchformer layer)");
  // branch 1: x1 = cell(norm_mhsa(x)) -> cat[:, :Dx1]                                   :317-322
  if (!ws.dry)
    SMX_TRY(layernorm(x, dtype, D, w->norm_mhsa_w, w->norm_mhsa_b, 1e-5f, SMX_ACT_IDENTITY, n, SMX_F32, D, rows, D, st));
  if (w->cell.mode == SMX_MODE_LITE) {
    float* mean = ws.f32((size_t)B * Dx1);
    if (!mean) return fail(SMX_ERR_WORKSPACE, "workspace too small (branchformer lite)");
    SMX_TRY(cell_generic(&w->cell, B, T, n, SMX_F32, mask, sum_mask, nullptr, 0, mean, SMX_F32, Dx1, ws, st));
    if (!ws.dry) SMX_TRY(broadcast_rows(mean, B, T, Dx1, cat, Dcat, st));
  } else {
    SMX_TRY(cell_generic(&w->cell, B, T, n, SMX_F32, mask, sum_mask, nullptr, 0, cat, SMX_F32, Dcat, ws, st));
  }
  // branch 2: x2 = conv_branch(norm_conv(x)) -> cat[:, Dx1:]   (no mask, :276)            :292-293, :86-97
  float* u = ws.f32((size_t)rows * U);
  float* g = ws.f32((size_t)rows * H);
  float* g2 = ws.f32((size_t)rows * H);
  size_t sc_bytes = 0;
  void* sc = split_scratch(ws, rows, {&br.pre, &br.post}, sc_bytes);
  if (!u || !g || !g2) return fail(SMX_ERR_WORKSPACE, "workspace too small (convolution branch)");
  if (!ws.dry) {
    SMX_TRY(layernorm(x, dtype, D, w->norm_conv_w, w->norm_conv_b, 1e-5f, SMX_ACT_IDENTITY, n, SMX_F32, D, rows, D, st));
    LinOpts o1; o1.act = br.act; o1.scratch = sc; o1.scratch_bytes = sc_bytes;
    SMX_TRY(linear(br.pre, n, SMX_F32, D, rows, u, SMX_F32, U, o1, st));
    // CSGU: gate half = u[:, H:], LN -> depthwise conv (reflect) -> [linear] -> gate_act -> * u[:, :H]
    SMX_TRY(layernorm(u + H, SMX_F32, U, br.csgu_ln_w, br.csgu_ln_b, 1e-5f, SMX_ACT_IDENTITY, g, SMX_F32, H, rows, H, st));
    SMX_TRY(dwconv(g, H, br.csgu_dw_w, br.csgu_dw_b, B, T, H, br.kernel_size, SMX_CONV_SAME_REFLECT, 0, g2, H, st));
    const float* gate = g2;
    if (br.csgu_linear.w) {
      LinOpts ol;
      SMX_TRY(linear(br.csgu_linear, g2, SMX_F32, H, rows, g, SMX_F32, H, ol, st));
      gate = g;
    }
    float* prod = (gate == g) ? g2 : g;
    SMX_TRY(gate_mul(gate, H, u, U, br.gate_act, rows, H, prod, st));
    LinOpts o2; o2.scratch = sc; o2.scratch_bytes = sc_bytes;
    SMX_TRY(linear(br.post, prod, SMX_F32, H, rows, cat + Dx1, SMX_F32, Dcat, o2, st));
  }
  // y = x + merge_proj(cat)                                                              :279
  SMX_TRY(vanilla_generic(w->merge, w->n_merge, w->act, cat, SMX_F32, Dcat, rows, nullptr, x, dtype, D, y, dtype, D, ws, st));
  ws.release(m0);
  return SMX_OK;
}

}  // namespace smx
