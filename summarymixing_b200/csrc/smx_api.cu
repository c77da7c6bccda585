// libsmx C ABI: argument validation, workspace sizing and dispatch between the tcgen05 arm
// (bf16, smx_tc_*.cu) and the generic fp32-math arm (smx_generic.cu).  See include/smx.h.
#include "smx_internal.h"
#include "smx_tc.h"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace smx {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<uint64_t> g_tc_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
void count_tc_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); g_tc_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return SMX_OK;
}

static int check_dtype(int dtype) {
  if (dtype != SMX_F32 && dtype != SMX_BF16) return fail(SMX_ERR_BAD_ARG, "unknown dtype %d", dtype);
  return SMX_OK;
}
static int check_ptr(const void* p, const char* name) {
  if (!p) return fail(SMX_ERR_BAD_ARG, "%s is NULL", name);
  if ((uintptr_t)p & 15) return fail(SMX_ERR_ALIGNMENT, "%s is not 16-byte aligned", name);
  return SMX_OK;
}
static int check_bt(int B, int T) {
  if (B <= 0 || T <= 0) return fail(SMX_ERR_BAD_ARG, "B and T must be positive (got %d, %d)", B, T);
  return SMX_OK;
}
static int check_arch() {
  static int cached = 0;  // 0 unknown, 1 ok, -1 bad
  if (cached == 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      cudaGetLastError();
      return fail(SMX_ERR_CUDA, "no usable CUDA device");
    }
    cached = (major == 10) ? 1 : -1;
  }
  if (cached < 0) return fail(SMX_ERR_ARCH, "libsmx is built for sm_100a (B200) only");
  return SMX_OK;
}

}  // namespace smx

using namespace smx;

extern "C" {

int smx_version(void) { return SMX_VERSION; }
const char* smx_last_error(void) { return g_err; }
uint64_t smx_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
uint64_t smx_tc_launch_count(void) { return g_tc_launches.load(std::memory_order_relaxed); }
size_t smx_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(smx_linear);
    case 1: return sizeof(smx_cell_weights);
    case 2: return sizeof(smx_ffn_weights);
    case 3: return sizeof(smx_convmod_weights);
    case 4: return sizeof(smx_conformer_layer_weights);
    case 5: return sizeof(smx_convbranch_weights);
    case 6: return sizeof(smx_branchformer_layer_weights);
    case 7: return sizeof(smx_cell_grads);
    case 8: return sizeof(smx_ffn_grads);
    case 9: return sizeof(smx_convmod_grads);
    case 10: return sizeof(smx_convbranch_grads);
    default: return 0;
  }
}

// ---- weight packing for the tcgen05 arm ------------------------------------------------------------
static int check_packed(const void* packed, size_t given, size_t need) {
  if (need == 0) return fail(SMX_ERR_UNSUPPORTED, "this configuration is not handled by the tensor-core arm");
  if (!packed) return fail(SMX_ERR_BAD_ARG, "packed is NULL");
  if ((uintptr_t)packed & 1023) return fail(SMX_ERR_ALIGNMENT, "packed must be 1024-byte aligned");
  if (given < need) return fail(SMX_ERR_WORKSPACE, "packed buffer too small: need %zu bytes, got %zu", need, given);
  return SMX_OK;
}
size_t smx_cell_packed_bytes(const smx_cell_weights* w) { return w ? tc_cell_packed_bytes(w) : 0; }
int smx_cell_pack(const smx_cell_weights* w, void* packed, size_t packed_bytes, void* stream) {
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  SMX_TRY(check_packed(packed, packed_bytes, tc_cell_packed_bytes(w)));
  SMX_TRY(check_arch());
  return tc_cell_pack(w, packed, (cudaStream_t)stream);
}
int smx_cell_pack_prenorm(smx_cell_weights* w, const float* norm_w, const float* norm_b, void* stream) {
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  SMX_TRY(check_arch());
  return tc_cell_pack_prenorm(w, norm_w, norm_b, (cudaStream_t)stream);
}
size_t smx_ffn_packed_bytes(const smx_ffn_weights* w) { return w ? tc_ffn_packed_bytes(w) : 0; }
int smx_ffn_pack(const smx_ffn_weights* w, void* packed, size_t packed_bytes, void* stream) {
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  SMX_TRY(check_packed(packed, packed_bytes, tc_ffn_packed_bytes(w)));
  SMX_TRY(check_arch());
  return tc_ffn_pack(w, packed, (cudaStream_t)stream);
}
size_t smx_branchformer_packed_bytes(const smx_branchformer_layer_weights* w) { return w ? tc_branchformer_packed_bytes(w) : 0; }
int smx_branchformer_pack(const smx_branchformer_layer_weights* w, void* packed, size_t packed_bytes, void* stream) {
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  SMX_TRY(check_packed(packed, packed_bytes, tc_branchformer_packed_bytes(w)));
  SMX_TRY(check_arch());
  return tc_branchformer_pack(w, packed, (cudaStream_t)stream);
}
size_t smx_convmod_packed_bytes(const smx_convmod_weights* w) { return w ? tc_convmod_packed_bytes(w) : 0; }
int smx_convmod_pack(const smx_convmod_weights* w, void* packed, size_t packed_bytes, void* stream) {
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  SMX_TRY(check_packed(packed, packed_bytes, tc_convmod_packed_bytes(w)));
  SMX_TRY(check_arch());
  return tc_convmod_pack(w, packed, (cudaStream_t)stream);
}

// ---- LayerNorm -----------------------------------------------------------------------------
int smx_layernorm_fwd(int dtype, int64_t rows, int32_t D, const void* x, const float* w, const float* b, float eps,
                      void* y, void* stream) {
  SMX_TRY(check_dtype(dtype));
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  if (!w || !b) return fail(SMX_ERR_BAD_ARG, "layernorm: NULL weight/bias");
  if (rows <= 0 || D <= 0) return fail(SMX_ERR_BAD_ARG, "layernorm: rows and D must be positive");
  SMX_TRY(check_arch());
  return layernorm(x, dtype, D, w, b, eps, SMX_ACT_IDENTITY, y, dtype, D, rows, D, (cudaStream_t)stream);
}

// ---- VanillaNN -------------------------------------------------------------------------------
size_t smx_vanilla_nn_workspace_bytes(const smx_linear* blocks, int32_t n_blocks, int dtype, int64_t rows) {
  if (!blocks || n_blocks < 1 || n_blocks > SMX_MAX_BLOCKS || rows <= 0) return 0;
  Arena a(nullptr, 0, true);
  vanilla_generic(blocks, n_blocks, 0, nullptr, dtype, blocks[0].in_dim, rows, nullptr, nullptr, 0, 0, nullptr, dtype,
                  blocks[n_blocks - 1].out_dim, a, nullptr);
  return a.peak;
}
int smx_vanilla_nn_fwd(const smx_linear* blocks, int32_t n_blocks, int act, int dtype, int64_t rows, const void* x,
                       void* y, void* workspace, size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!blocks) return fail(SMX_ERR_BAD_ARG, "blocks is NULL");
  if (rows <= 0) return fail(SMX_ERR_BAD_ARG, "rows must be positive");
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  size_t need = smx_vanilla_nn_workspace_bytes(blocks, n_blocks, dtype, rows);
  if (need > workspace_bytes || (need && !workspace))
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  return vanilla_generic(blocks, n_blocks, act, x, dtype, blocks[0].in_dim, rows, nullptr, nullptr, 0, 0, y, dtype,
                         blocks[n_blocks - 1].out_dim, a, (cudaStream_t)stream);
}

// ---- SummaryMixing cell ------------------------------------------------------------------------
size_t smx_summary_mixing_workspace_bytes(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, int has_sum_mask) {
  if (!w || B <= 0 || T <= 0) return 0;
  Arena a(nullptr, 0, true);
  const float* sm = has_sum_mask ? (const float*)(uintptr_t)256 : nullptr;
  int Dout = w->mode == SMX_MODE_LITE ? w->summary_out_dim : w->merge.out_dim;
  cell_generic(w, B, T, nullptr, dtype, nullptr, sm, nullptr, dtype, nullptr, dtype, Dout, a, nullptr);
  size_t tc = (dtype == SMX_BF16 && w->packed) ? tc_cell_workspace_bytes(w, B, T) : 0;
  if (tc && !has_sum_mask && tc_cell_supported(w, 0)) return tc;  // the tensor-core arm runs: only its scratch is needed
  return a.peak > tc ? a.peak : tc;
}
int smx_summary_mixing_fwd(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x,
                           const uint8_t* padding_mask, const float* sum_mask, const void* residual, void* y,
                           void* workspace, size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  size_t need = smx_summary_mixing_workspace_bytes(w, dtype, B, T, sum_mask != nullptr);
  if (need > workspace_bytes || (need && !workspace))
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  int Dout = w->mode == SMX_MODE_LITE ? w->summary_out_dim : w->merge.out_dim;
  if (dtype == SMX_BF16 && w->packed && tc_cell_supported(w, sum_mask != nullptr))
    return tc_cell_fwd(w, w->packed, B, T, (const __nv_bfloat16*)x, nullptr, nullptr, padding_mask,
                       (const __nv_bfloat16*)residual, (__nv_bfloat16*)y, a, (cudaStream_t)stream);
  return cell_generic(w, B, T, x, dtype, padding_mask, sum_mask, residual, dtype, y, dtype, Dout, a, (cudaStream_t)stream);
}

// ---- SummaryMixing cell, backward ----------------------------------------------------------------
static void all_grads_wanted(smx_cell_grads& g) {
  float* dummy = (float*)(uintptr_t)256;  // sizing only: never dereferenced
  for (int i = 0; i < SMX_MAX_BLOCKS; ++i) { g.local[i] = {dummy, dummy}; g.summary[i] = {dummy, dummy}; }
  g.merge = {dummy, dummy};
  g.local_norm_dw = g.local_norm_db = g.summary_norm_dw = g.summary_norm_db = dummy;
  g.global_proj = {dummy, dummy};
}
size_t smx_summary_mixing_bwd_workspace_bytes(const smx_cell_weights* w, int dtype, int32_t B, int32_t T) {
  if (!w || B <= 0 || T <= 0) return 0;
  smx_cell_grads g;
  all_grads_wanted(g);
  Arena a(nullptr, 0, true);
  if (cell_bwd_generic(w, B, T, nullptr, dtype, nullptr, nullptr, dtype, (void*)(uintptr_t)256, dtype, &g, a, nullptr) != SMX_OK)
    return 0;
  return a.peak;
}
int smx_summary_mixing_bwd(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x,
                           const uint8_t* padding_mask, const void* dy, void* dx, const smx_cell_grads* grads,
                           void* workspace, size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  if (!grads) return fail(SMX_ERR_BAD_ARG, "grads is NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(dy, "dy"));
  if (dx) SMX_TRY(check_ptr(dx, "dx"));
  SMX_TRY(check_arch());
  Arena dry(nullptr, 0, true);
  SMX_TRY(cell_bwd_generic(w, B, T, x, dtype, padding_mask, dy, dtype, dx, dtype, grads, dry, nullptr));
  if (dry.peak > workspace_bytes || (dry.peak && !workspace))
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", dry.peak, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  return cell_bwd_generic(w, B, T, x, dtype, padding_mask, dy, dtype, dx, dtype, grads, a, (cudaStream_t)stream);
}

// ---- LayerNorm / FFN / conv module, backward ------------------------------------------------------
static float* const kDummy = (float*)(uintptr_t)256;  // sizing runs only: never dereferenced
size_t smx_vanilla_nn_bwd_workspace_bytes(const smx_linear* blocks, int32_t n_blocks, int dtype, int64_t rows) {
  if (!blocks || n_blocks < 1 || n_blocks > SMX_MAX_BLOCKS || rows <= 0) return 0;
  smx_linear_grad g[SMX_MAX_BLOCKS];
  for (int i = 0; i < SMX_MAX_BLOCKS; ++i) g[i] = {kDummy, kDummy};
  Arena a(nullptr, 0, true);
  if (vanilla_bwd_generic(blocks, n_blocks, 0, rows, nullptr, dtype, nullptr, dtype, kDummy, dtype, g, a, nullptr) != SMX_OK) return 0;
  return a.peak;
}
int smx_vanilla_nn_bwd(const smx_linear* blocks, int32_t n_blocks, int act, int dtype, int64_t rows, const void* x, const void* dy,
                       void* dx, const smx_linear_grad* grads, void* workspace, size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!blocks || !grads) return fail(SMX_ERR_BAD_ARG, "blocks or grads is NULL");
  if (rows <= 0) return fail(SMX_ERR_BAD_ARG, "rows must be positive");
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(dy, "dy"));
  if (dx) SMX_TRY(check_ptr(dx, "dx"));
  SMX_TRY(check_arch());
  Arena dry(nullptr, 0, true);
  SMX_TRY(vanilla_bwd_generic(blocks, n_blocks, act, rows, x, dtype, dy, dtype, dx, dtype, grads, dry, nullptr));
  if (dry.peak > workspace_bytes || (dry.peak && !workspace))
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", dry.peak, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  return vanilla_bwd_generic(blocks, n_blocks, act, rows, x, dtype, dy, dtype, dx, dtype, grads, a, (cudaStream_t)stream);
}
size_t smx_layernorm_bwd_workspace_bytes(int dtype, int64_t rows, int32_t D) {
  if (rows <= 0 || D <= 0) return 0;
  Arena a(nullptr, 0, true);
  layernorm_bwd_generic(nullptr, dtype, rows, D, nullptr, 0.f, nullptr, dtype, kDummy, dtype, kDummy, kDummy, a, nullptr);
  return a.peak;
}
int smx_layernorm_bwd(int dtype, int64_t rows, int32_t D, const void* x, const float* w, float eps, const void* dy, void* dx,
                      float* dw, float* db, void* workspace, size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(dy, "dy"));
  if (dx) SMX_TRY(check_ptr(dx, "dx"));
  if (!w) return fail(SMX_ERR_BAD_ARG, "layernorm backward: NULL weight");
  if (rows <= 0 || D <= 0) return fail(SMX_ERR_BAD_ARG, "layernorm backward: rows and D must be positive");
  SMX_TRY(check_arch());
  size_t need = smx_layernorm_bwd_workspace_bytes(dtype, rows, D);
  if (need > workspace_bytes || (need && !workspace))
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  return layernorm_bwd_generic(x, dtype, rows, D, w, eps, dy, dtype, dx, dtype, dw, db, a, (cudaStream_t)stream);
}
size_t smx_ffn_bwd_workspace_bytes(const smx_ffn_weights* w, int dtype, int64_t rows, int has_out_ln) {
  if (!w || rows <= 0) return 0;
  smx_ffn_grads g{kDummy, kDummy, {kDummy, kDummy}, {kDummy, kDummy}, kDummy, kDummy};
  Arena a(nullptr, 0, true);
  if (ffn_bwd_generic(w, 0, rows, nullptr, dtype, has_out_ln ? kDummy : nullptr, has_out_ln ? kDummy : nullptr, 0.f, nullptr, dtype,
                      kDummy, dtype, &g, a, nullptr) != SMX_OK)
    return 0;
  return a.peak;
}
int smx_ffn_bwd(const smx_ffn_weights* w, int act, int dtype, int64_t rows, const void* x, const float* out_ln_w,
                const float* out_ln_b, float out_ln_eps, const void* dy, void* dx, const smx_ffn_grads* grads, void* workspace,
                size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!w || !grads) return fail(SMX_ERR_BAD_ARG, "weights or grads is NULL");
  if (rows <= 0) return fail(SMX_ERR_BAD_ARG, "rows must be positive");
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(dy, "dy"));
  if (dx) SMX_TRY(check_ptr(dx, "dx"));
  SMX_TRY(check_arch());
  size_t need = smx_ffn_bwd_workspace_bytes(w, dtype, rows, out_ln_w != nullptr);
  if (need == 0) return SMX_ERR_BAD_ARG;  // message set by the sizing run
  if (need > workspace_bytes || !workspace)
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  return ffn_bwd_generic(w, act, rows, x, dtype, out_ln_w, out_ln_b, out_ln_eps, dy, dtype, dx, dtype, grads, a, (cudaStream_t)stream);
}
size_t smx_conv_module_bwd_workspace_bytes(const smx_convmod_weights* w, int dtype, int32_t B, int32_t T) {
  if (!w || B <= 0 || T <= 0) return 0;
  smx_convmod_grads g{kDummy, kDummy, {kDummy, kDummy}, kDummy, kDummy, kDummy, kDummy, {kDummy, kDummy}};
  Arena a(nullptr, 0, true);
  if (convmod_bwd_generic(w, 0, B, T, nullptr, dtype, nullptr, nullptr, dtype, kDummy, dtype, &g, a, nullptr) != SMX_OK) return 0;
  return a.peak;
}
int smx_conv_module_bwd(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, const void* x,
                        const uint8_t* padding_mask, const void* dy, void* dx, const smx_convmod_grads* grads, void* workspace,
                        size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!w || !grads) return fail(SMX_ERR_BAD_ARG, "weights or grads is NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(dy, "dy"));
  if (dx) SMX_TRY(check_ptr(dx, "dx"));
  SMX_TRY(check_arch());
  size_t need = smx_conv_module_bwd_workspace_bytes(w, dtype, B, T);
  if (need == 0) return SMX_ERR_BAD_ARG;
  if (need > workspace_bytes || !workspace)
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  return convmod_bwd_generic(w, act, B, T, x, dtype, padding_mask, dy, dtype, dx, dtype, grads, a, (cudaStream_t)stream);
}

// ---- training-mode forward / backward with dropout ------------------------------------------------
static int check_drop(const smx_dropout* d) {
  if (d && !(d->p >= 0.0f && d->p < 1.0f)) return fail(SMX_ERR_BAD_ARG, "dropout p must be in [0, 1), got %g", (double)d->p);
  return SMX_OK;
}
static const smx_dropout kSizingDrop{0.5f, 0};  // sizing runs take the dropout branches (the larger workspace)
size_t smx_ffn_train_workspace_bytes(const smx_ffn_weights* w, int dtype, int64_t rows, int has_out_ln) {
  if (!w || rows <= 0) return 0;
  smx_ffn_grads g{kDummy, kDummy, {kDummy, kDummy}, {kDummy, kDummy}, kDummy, kDummy};
  Arena a(nullptr, 0, true), b(nullptr, 0, true);
  float* const oln = has_out_ln ? kDummy : nullptr;
  if (ffn_bwd_generic(w, 0, rows, nullptr, dtype, oln, oln, 0.f, nullptr, dtype, kDummy, dtype, &g, a, nullptr, &kSizingDrop) != SMX_OK) return 0;
  if (ffn_bwd_generic(w, 0, rows, nullptr, dtype, oln, oln, 0.f, nullptr, dtype, nullptr, dtype, &g, b, nullptr, &kSizingDrop, kDummy, dtype) != SMX_OK) return 0;
  return a.peak > b.peak ? a.peak : b.peak;
}
static int ffn_train(const smx_ffn_weights* w, int act, int dtype, int64_t rows, const void* x, const float* out_ln_w, const float* out_ln_b,
                     float out_ln_eps, const smx_dropout* drop, const void* dy, void* dx, const smx_ffn_grads* grads, void* y, void* workspace,
                     size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  if (rows <= 0) return fail(SMX_ERR_BAD_ARG, "rows must be positive");
  SMX_TRY(check_drop(drop));
  SMX_TRY(check_ptr(x, "x"));
  if (y) SMX_TRY(check_ptr(y, "y")); else { if (!grads) return fail(SMX_ERR_BAD_ARG, "grads is NULL"); SMX_TRY(check_ptr(dy, "dy")); }
  if (dx) SMX_TRY(check_ptr(dx, "dx"));
  SMX_TRY(check_arch());
  const size_t need = smx_ffn_train_workspace_bytes(w, dtype, rows, out_ln_w != nullptr);
  if (need == 0) return SMX_ERR_BAD_ARG;
  if (need > workspace_bytes || !workspace) return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  smx_ffn_grads none{};
  return ffn_bwd_generic(w, act, rows, x, dtype, out_ln_w, out_ln_b, out_ln_eps, dy, dtype, dx, dtype, grads ? grads : &none, a, (cudaStream_t)stream,
                         drop, y, dtype);
}
int smx_ffn_train_fwd(const smx_ffn_weights* w, int act, int dtype, int64_t rows, const void* x, const float* out_ln_w, const float* out_ln_b,
                      float out_ln_eps, const smx_dropout* drop, void* y, void* workspace, size_t workspace_bytes, void* stream) {
  if (!y) return fail(SMX_ERR_BAD_ARG, "y is NULL");
  return ffn_train(w, act, dtype, rows, x, out_ln_w, out_ln_b, out_ln_eps, drop, nullptr, nullptr, nullptr, y, workspace, workspace_bytes, stream);
}
int smx_ffn_train_bwd(const smx_ffn_weights* w, int act, int dtype, int64_t rows, const void* x, const float* out_ln_w, const float* out_ln_b,
                      float out_ln_eps, const smx_dropout* drop, const void* dy, void* dx, const smx_ffn_grads* grads, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return ffn_train(w, act, dtype, rows, x, out_ln_w, out_ln_b, out_ln_eps, drop, dy, dx, grads, nullptr, workspace, workspace_bytes, stream);
}
size_t smx_conv_module_train_workspace_bytes(const smx_convmod_weights* w, int dtype, int32_t B, int32_t T) {
  return smx_conv_module_bwd_workspace_bytes(w, dtype, B, T);  // the forward-only run uses a prefix of the backward's buffers
}
static int convmod_train(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, const void* x, const uint8_t* padding_mask,
                         const smx_dropout* drop, const void* dy, void* dx, const smx_convmod_grads* grads, void* y, void* workspace,
                         size_t workspace_bytes, void* stream, int chunk = 0) {
  if (chunk < 0) return fail(SMX_ERR_BAD_ARG, "chunk_size must be non-negative");
  SMX_TRY(check_dtype(dtype));
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_drop(drop));
  SMX_TRY(check_ptr(x, "x"));
  if (y) SMX_TRY(check_ptr(y, "y")); else { if (!grads) return fail(SMX_ERR_BAD_ARG, "grads is NULL"); SMX_TRY(check_ptr(dy, "dy")); }
  if (dx) SMX_TRY(check_ptr(dx, "dx"));
  SMX_TRY(check_arch());
  const size_t need = smx_conv_module_train_workspace_bytes(w, dtype, B, T);
  if (need == 0) return SMX_ERR_BAD_ARG;
  if (need > workspace_bytes || !workspace) return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  smx_convmod_grads none{};
  return convmod_bwd_generic(w, act, B, T, x, dtype, padding_mask, dy, dtype, dx, dtype, grads ? grads : &none, a, (cudaStream_t)stream, drop, y, dtype, chunk);
}
int smx_conv_module_dcc_train_fwd(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, int32_t chunk_size, const void* x,
                                  const uint8_t* padding_mask, const smx_dropout* drop, void* y, void* workspace, size_t workspace_bytes, void* stream) {
  if (!y) return fail(SMX_ERR_BAD_ARG, "y is NULL");
  return convmod_train(w, act, dtype, B, T, x, padding_mask, drop, nullptr, nullptr, nullptr, y, workspace, workspace_bytes, stream, chunk_size);
}
int smx_conv_module_dcc_train_bwd(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, int32_t chunk_size, const void* x,
                                  const uint8_t* padding_mask, const smx_dropout* drop, const void* dy, void* dx, const smx_convmod_grads* grads,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  return convmod_train(w, act, dtype, B, T, x, padding_mask, drop, dy, dx, grads, nullptr, workspace, workspace_bytes, stream, chunk_size);
}
int smx_conv_module_train_fwd(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, const void* x, const uint8_t* padding_mask,
                              const smx_dropout* drop, void* y, void* workspace, size_t workspace_bytes, void* stream) {
  if (!y) return fail(SMX_ERR_BAD_ARG, "y is NULL");
  return convmod_train(w, act, dtype, B, T, x, padding_mask, drop, nullptr, nullptr, nullptr, y, workspace, workspace_bytes, stream);
}
int smx_conv_module_train_bwd(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, const void* x, const uint8_t* padding_mask,
                              const smx_dropout* drop, const void* dy, void* dx, const smx_convmod_grads* grads, void* workspace,
                              size_t workspace_bytes, void* stream) {
  return convmod_train(w, act, dtype, B, T, x, padding_mask, drop, dy, dx, grads, nullptr, workspace, workspace_bytes, stream);
}
size_t smx_conv_branch_train_workspace_bytes(const smx_convbranch_weights* w, int dtype, int32_t B, int32_t T) {
  if (!w || B <= 0 || T <= 0) return 0;
  smx_convbranch_grads g{{kDummy, kDummy}, {kDummy, kDummy}, kDummy, kDummy, kDummy, kDummy, {kDummy, kDummy}};
  Arena a(nullptr, 0, true), b(nullptr, 0, true);
  if (convbranch_bwd_generic(w, B, T, nullptr, dtype, nullptr, dtype, kDummy, dtype, &g, a, nullptr, &kSizingDrop) != SMX_OK) return 0;
  if (convbranch_bwd_generic(w, B, T, nullptr, dtype, nullptr, dtype, nullptr, dtype, &g, b, nullptr, &kSizingDrop, kDummy, dtype) != SMX_OK) return 0;
  return a.peak > b.peak ? a.peak : b.peak;
}
static int convbranch_train(const smx_convbranch_weights* w, int dtype, int32_t B, int32_t T, const void* x, const smx_dropout* drop,
                            const void* dy, void* dx, const smx_convbranch_grads* grads, void* y, void* workspace, size_t workspace_bytes,
                            void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_drop(drop));
  SMX_TRY(check_ptr(x, "x"));
  if (y) SMX_TRY(check_ptr(y, "y")); else { if (!grads) return fail(SMX_ERR_BAD_ARG, "grads is NULL"); SMX_TRY(check_ptr(dy, "dy")); }
  if (dx) SMX_TRY(check_ptr(dx, "dx"));
  SMX_TRY(check_arch());
  const size_t need = smx_conv_branch_train_workspace_bytes(w, dtype, B, T);
  if (need == 0) {  // the sizing run failed: repeat it for its message (bad dims, T too short for the reflect padding)
    Arena a(nullptr, 0, true);
    smx_convbranch_grads none{};
    const int rc = convbranch_bwd_generic(w, B, T, nullptr, dtype, nullptr, dtype, nullptr, dtype, &none, a, nullptr, nullptr, kDummy, dtype);
    return rc != SMX_OK ? rc : SMX_ERR_BAD_ARG;
  }
  if (need > workspace_bytes || !workspace) return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  smx_convbranch_grads none{};
  return convbranch_bwd_generic(w, B, T, x, dtype, dy, dtype, dx, dtype, grads ? grads : &none, a, (cudaStream_t)stream, drop, y, dtype);
}
int smx_conv_branch_train_fwd(const smx_convbranch_weights* w, int dtype, int32_t B, int32_t T, const void* x, const smx_dropout* drop, void* y,
                              void* workspace, size_t workspace_bytes, void* stream) {
  if (!y) return fail(SMX_ERR_BAD_ARG, "y is NULL");
  return convbranch_train(w, dtype, B, T, x, drop, nullptr, nullptr, nullptr, y, workspace, workspace_bytes, stream);
}
int smx_conv_branch_train_bwd(const smx_convbranch_weights* w, int dtype, int32_t B, int32_t T, const void* x, const smx_dropout* drop,
                              const void* dy, void* dx, const smx_convbranch_grads* grads, void* workspace, size_t workspace_bytes, void* stream) {
  return convbranch_train(w, dtype, B, T, x, drop, dy, dx, grads, nullptr, workspace, workspace_bytes, stream);
}
int smx_dropout_apply(const smx_dropout* drop, int32_t site, int dtype, int64_t n, const void* x, void* y, void* stream) {
  SMX_TRY(check_dtype(dtype));
  SMX_TRY(check_drop(drop));
  if (n < 0) return fail(SMX_ERR_BAD_ARG, "n must be non-negative");
  if (n == 0) return SMX_OK;
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  return dropout_apply(drop, site, dtype, n, x, y, (cudaStream_t)stream);
}
size_t smx_summary_mixing_train_workspace_bytes(const smx_cell_weights* w, int dtype, int32_t B, int32_t T) {
  if (!w || B <= 0 || T <= 0) return 0;
  smx_cell_grads g;
  all_grads_wanted(g);
  Arena a(nullptr, 0, true);
  if (cell_bwd_generic(w, B, T, nullptr, dtype, nullptr, nullptr, dtype, (void*)(uintptr_t)256, dtype, &g, a, nullptr, &kSizingDrop) != SMX_OK) return 0;
  return a.peak;
}
size_t smx_summary_mixing_masked_train_workspace_bytes(const smx_cell_weights* w, int dtype, int32_t B, int32_t T) {
  if (!w || B <= 0 || T <= 0) return 0;
  smx_cell_grads g;
  all_grads_wanted(g);
  Arena a(nullptr, 0, true);
  if (cell_bwd_generic(w, B, T, nullptr, dtype, nullptr, nullptr, dtype, (void*)(uintptr_t)256, dtype, &g, a, nullptr, &kSizingDrop, nullptr, 0,
                       (const float*)(uintptr_t)256) != SMX_OK) return 0;
  const size_t plain = smx_summary_mixing_train_workspace_bytes(w, dtype, B, T);
  return a.peak > plain ? a.peak : plain;
}
static int cell_train(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x, const uint8_t* padding_mask,
                      const smx_dropout* drop, const void* dy, void* dx, const smx_cell_grads* grads, void* y, void* workspace,
                      size_t workspace_bytes, void* stream, const float* sum_mask = nullptr) {
  SMX_TRY(check_dtype(dtype));
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_drop(drop));
  SMX_TRY(check_ptr(x, "x"));
  if (y) SMX_TRY(check_ptr(y, "y")); else { if (!grads) return fail(SMX_ERR_BAD_ARG, "grads is NULL"); SMX_TRY(check_ptr(dy, "dy")); }
  if (dx) SMX_TRY(check_ptr(dx, "dx"));
  SMX_TRY(check_arch());
  if (sum_mask && w->mode == SMX_MODE_LITE) sum_mask = nullptr;  // (lite ignores sum_mask, summary_mixing.py:300-324)
  if (sum_mask) SMX_TRY(check_ptr(sum_mask, "sum_mask"));
  const size_t need = sum_mask ? smx_summary_mixing_masked_train_workspace_bytes(w, dtype, B, T) : smx_summary_mixing_train_workspace_bytes(w, dtype, B, T);
  if (need == 0) return SMX_ERR_BAD_ARG;
  if (need > workspace_bytes || !workspace) return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  smx_cell_grads none{};
  return cell_bwd_generic(w, B, T, x, dtype, padding_mask, dy, dtype, dx, dtype, grads ? grads : &none, a, (cudaStream_t)stream, drop, y, dtype, sum_mask);
}
int smx_summary_mixing_masked_train_fwd(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x, const uint8_t* padding_mask,
                                        const float* sum_mask, const smx_dropout* drop, void* y, void* workspace, size_t workspace_bytes, void* stream) {
  if (!y) return fail(SMX_ERR_BAD_ARG, "y is NULL");
  return cell_train(w, dtype, B, T, x, padding_mask, drop, nullptr, nullptr, nullptr, y, workspace, workspace_bytes, stream, sum_mask);
}
int smx_summary_mixing_masked_train_bwd(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x, const uint8_t* padding_mask,
                                        const float* sum_mask, const smx_dropout* drop, const void* dy, void* dx, const smx_cell_grads* grads,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  return cell_train(w, dtype, B, T, x, padding_mask, drop, dy, dx, grads, nullptr, workspace, workspace_bytes, stream, sum_mask);
}
int smx_summary_mixing_train_fwd(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x, const uint8_t* padding_mask,
                                 const smx_dropout* drop, void* y, void* workspace, size_t workspace_bytes, void* stream) {
  if (!y) return fail(SMX_ERR_BAD_ARG, "y is NULL");
  return cell_train(w, dtype, B, T, x, padding_mask, drop, nullptr, nullptr, nullptr, y, workspace, workspace_bytes, stream);
}
int smx_summary_mixing_train_bwd(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x, const uint8_t* padding_mask,
                                 const smx_dropout* drop, const void* dy, void* dx, const smx_cell_grads* grads, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  return cell_train(w, dtype, B, T, x, padding_mask, drop, dy, dx, grads, nullptr, workspace, workspace_bytes, stream);
}
int smx_dropout_keep_mask(const smx_dropout* drop, int32_t site, int64_t n, uint8_t* keep, void* stream) {
  SMX_TRY(check_drop(drop));
  if (n < 0 || site < 0) return fail(SMX_ERR_BAD_ARG, "n and site must be non-negative");
  if (n > 0) SMX_TRY(check_ptr(keep, "keep"));
  SMX_TRY(check_arch());
  return dropout_keep_mask(drop, site, n, keep, (cudaStream_t)stream);
}

// ---- ConvolutionModule ---------------------------------------------------------------------------
size_t smx_conv_module_workspace_bytes(const smx_convmod_weights* w, int dtype, int32_t B, int32_t T) {
  if (!w || B <= 0 || T <= 0) return 0;
  Arena a(nullptr, 0, true);
  convmod_generic(w, 0, B, T, 0, nullptr, dtype, nullptr, nullptr, dtype, nullptr, dtype, a, nullptr);
  size_t tc = (dtype == SMX_BF16 && w->packed) ? tc_convmod_workspace_bytes(w, B, T) : 0;
  return a.peak > tc ? a.peak : tc;
}
int smx_conv_module_fwd(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, int32_t chunk_size,
                        const void* x, const uint8_t* padding_mask, const void* residual, void* y, void* workspace,
                        size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  size_t need = smx_conv_module_workspace_bytes(w, dtype, B, T);
  if (need > workspace_bytes || (need && !workspace))
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  if (dtype == SMX_BF16 && w->packed && tc_convmod_supported(w, chunk_size))
    return tc_convmod_fwd(w, w->packed, act, B, T, (const __nv_bfloat16*)x, padding_mask, (const __nv_bfloat16*)residual,
                          (__nv_bfloat16*)y, a, (cudaStream_t)stream);
  return convmod_generic(w, act, B, T, chunk_size, x, dtype, padding_mask, residual, dtype, y, dtype, a, (cudaStream_t)stream);
}

// ---- FFN -----------------------------------------------------------------------------------------
size_t smx_ffn_workspace_bytes(const smx_ffn_weights* w, int dtype, int64_t rows) {
  if (!w || rows <= 0) return 0;
  if (dtype == SMX_BF16 && w->packed && tc_ffn_supported(w)) return tc_ffn_workspace_bytes(w, rows);  // fused: no scratch
  Arena a(nullptr, 0, true);
  ffn_generic(w, 0, rows, nullptr, dtype, nullptr, nullptr, 0.f, nullptr, dtype, a, nullptr);
  return a.peak;
}
int smx_ffn_fwd(const smx_ffn_weights* w, int act, int dtype, int64_t rows, const void* x, const float* out_ln_w,
                const float* out_ln_b, float out_ln_eps, void* y, void* workspace, size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  if (rows <= 0) return fail(SMX_ERR_BAD_ARG, "rows must be positive");
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  size_t need = smx_ffn_workspace_bytes(w, dtype, rows);
  if (need > workspace_bytes || (need && !workspace))
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  if (dtype == SMX_BF16 && w->packed && tc_ffn_supported(w))
    return tc_ffn_fwd(w, w->packed, act, rows, (const __nv_bfloat16*)x, out_ln_w, out_ln_b, out_ln_eps, (__nv_bfloat16*)y, a,
                      (cudaStream_t)stream);
  return ffn_generic(w, act, rows, x, dtype, out_ln_w, out_ln_b, out_ln_eps, y, dtype, a, (cudaStream_t)stream);
}

// Zeroes `bytes` at the start of the workspace (the fused cell kernels' per-utterance counters, smx_tc_cell4.cu) with ONE
// memset ahead of the whole kernel chain and hands the area down to tc_cell4_fwd (thread-local); released on scope exit.
struct PresyncGuard {
  size_t bytes;
  int status;
  PresyncGuard(void* ws, size_t n, bool active, cudaStream_t st) : bytes(n), status(SMX_OK) {
    if (!active || !ws || n == 0) { tc_cell4_set_presync(nullptr, 0); return; }
    cudaError_t e = cudaMemsetAsync(ws, 0, n, st);
    if (e != cudaSuccess) { status = fail(SMX_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e)); return; }
    tc_cell4_set_presync(ws, n);
  }
  ~PresyncGuard() { tc_cell4_set_presync(nullptr, 0); }
};

// ---- Conformer layer / encoder -------------------------------------------------------------------
size_t smx_mixing_block_workspace_bytes(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, int has_sum_mask) {
  if (!w || B <= 0 || T <= 0) return 0;
  Arena a(nullptr, 0, true);
  const float* sm = has_sum_mask ? (const float*)(uintptr_t)256 : nullptr;
  if (mixing_block_generic(w, kDummy, kDummy, dtype, B, T, nullptr, nullptr, sm, nullptr, a, nullptr) != SMX_OK) return 0;
  return a.peak + tc_cell4_sync_bytes(B);
}
int smx_mixing_block_fwd(const smx_cell_weights* w, const float* norm_w, const float* norm_b, int dtype, int32_t B, int32_t T,
                         const void* x, const uint8_t* padding_mask, const float* sum_mask, void* y, void* workspace,
                         size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!w || !norm_w || !norm_b) return fail(SMX_ERR_BAD_ARG, "weights / norm parameters are NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  size_t need = smx_mixing_block_workspace_bytes(w, dtype, B, T, sum_mask != nullptr);
  if (need == 0) return fail(SMX_ERR_BAD_ARG, "mixing block: unsupported configuration (summary_out_dim must equal enc_dim)");
  if (need > workspace_bytes || !workspace)
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  PresyncGuard presync(workspace, tc_cell4_sync_bytes(B), dtype == SMX_BF16, (cudaStream_t)stream);
  if (presync.status != SMX_OK) return presync.status;
  Arena a((char*)workspace + presync.bytes, workspace_bytes - presync.bytes, false);
  return mixing_block_generic(w, norm_w, norm_b, dtype, B, T, x, padding_mask, sum_mask, y, a, (cudaStream_t)stream);
}

int smx_mixing_block_fwd_batch(const smx_cell_weights* w, const float* norm_w, const float* norm_b, int dtype, int32_t B, int32_t T,
                               int32_t n, const void* const* xs, const uint8_t* padding_mask, void* const* ys, void* workspace,
                               size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!w || !norm_w || !norm_b || !xs || !ys || n <= 0) return fail(SMX_ERR_BAD_ARG, "weights / norm parameters / batch lists are NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_arch());
  const size_t one = smx_mixing_block_workspace_bytes(w, dtype, B, T, 0);
  if (one == 0) return fail(SMX_ERR_BAD_ARG, "mixing block: unsupported configuration (summary_out_dim must equal enc_dim)");
  const size_t sync_all = (size_t)n * tc_cell4_sync_bytes(B);
  if (one + sync_all > workspace_bytes || !workspace)
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", one + sync_all, workspace_bytes);
  PresyncGuard presync(workspace, sync_all, dtype == SMX_BF16, (cudaStream_t)stream);
  if (presync.status != SMX_OK) return presync.status;
  for (int i = 0; i < n; ++i) {
    SMX_TRY(check_ptr(xs[i], "xs[i]")); SMX_TRY(check_ptr(ys[i], "ys[i]"));
    Arena a((char*)workspace + presync.bytes, workspace_bytes - presync.bytes, false);
    SMX_TRY(mixing_block_generic(w, norm_w, norm_b, dtype, B, T, xs[i], padding_mask, nullptr, ys[i], a, (cudaStream_t)stream));
  }
  return SMX_OK;
}

size_t smx_conformer_layer_workspace_bytes(const smx_conformer_layer_weights* w, int dtype, int32_t B, int32_t T,
                                           int has_sum_mask) {
  if (!w || B <= 0 || T <= 0) return 0;
  Arena a(nullptr, 0, true);
  const float* sm = has_sum_mask ? (const float*)(uintptr_t)256 : nullptr;
  conformer_layer_generic(w, dtype, B, T, 0, nullptr, nullptr, sm, nullptr, a, nullptr);
  return a.peak + tc_cell4_sync_bytes(B);
}
int smx_conformer_layer_fwd(const smx_conformer_layer_weights* w, int dtype, int32_t B, int32_t T, int32_t chunk_size,
                            const void* x, const uint8_t* padding_mask, const float* sum_mask, void* y, void* workspace,
                            size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  size_t need = smx_conformer_layer_workspace_bytes(w, dtype, B, T, sum_mask != nullptr);
  if (need > workspace_bytes || (need && !workspace))
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  // the fused cell's per-utterance counters are zeroed here, ahead of the layer's kernel chain (a memset node right in
  // front of the cell kernel would cut its programmatic-launch edge to the FFN kernel)
  PresyncGuard presync(workspace, tc_cell4_sync_bytes(B), dtype == SMX_BF16, (cudaStream_t)stream);
  if (presync.status != SMX_OK) return presync.status;
  Arena a((char*)workspace + presync.bytes, workspace_bytes - presync.bytes, false);
  return conformer_layer_generic(w, dtype, B, T, chunk_size, x, padding_mask, sum_mask, y, a, (cudaStream_t)stream);
}

size_t smx_conformer_encoder_workspace_bytes(const smx_conformer_layer_weights* layers, int32_t n_layers, int dtype,
                                             int32_t B, int32_t T, int has_sum_mask) {
  if (!layers || n_layers <= 0 || B <= 0 || T <= 0) return 0;
  size_t m = 0;
  for (int i = 0; i < n_layers; ++i) {
    size_t s = smx_conformer_layer_workspace_bytes(&layers[i], dtype, B, T, has_sum_mask);
    if (s > m) m = s;
  }
  return m + (size_t)n_layers * tc_cell4_sync_bytes(B);
}
int smx_conformer_encoder_fwd(const smx_conformer_layer_weights* layers, int32_t n_layers, const float* final_norm_w,
                              const float* final_norm_b, int dtype, int32_t B, int32_t T, int32_t chunk_size,
                              const void* x, const uint8_t* padding_mask, const float* sum_mask, void* y,
                              void* const* hidden, void* workspace, size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!layers || n_layers <= 0) return fail(SMX_ERR_BAD_ARG, "layers is NULL or n_layers <= 0");
  if (!final_norm_w || !final_norm_b) return fail(SMX_ERR_BAD_ARG, "final norm weights are NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  size_t need = smx_conformer_encoder_workspace_bytes(layers, n_layers, dtype, B, T, sum_mask != nullptr);
  if (need > workspace_bytes || (need && !workspace))
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int D = layers[0].ffn1.w1.in_dim;
  const int64_t rows = (int64_t)B * T;
  const void* cur = x;
  PresyncGuard presync(workspace, (size_t)n_layers * tc_cell4_sync_bytes(B), dtype == SMX_BF16, st);  // one memset for all layers' counters
  if (presync.status != SMX_OK) return presync.status;
  for (int i = 0; i < n_layers; ++i) {
    void* out = hidden ? hidden[i] : y;
    if (hidden) SMX_TRY(check_ptr(out, "hidden[i]"));
    Arena a((char*)workspace + presync.bytes, workspace_bytes - presync.bytes, false);
    SMX_TRY(conformer_layer_generic(&layers[i], dtype, B, T, chunk_size, cur, padding_mask, sum_mask, out, a, st));
    cur = out;
  }
  SMX_TRY(layernorm(cur, dtype, D, final_norm_w, final_norm_b, 1e-6f, SMX_ACT_IDENTITY, y, dtype, D, rows, D, st));
  if (hidden && hidden[n_layers - 1] != y)  // hidden_lst[-1] = output (Conformer.py:824)
    SMX_TRY(convert(y, dtype, hidden[n_layers - 1], dtype, rows * D, st));
  return SMX_OK;
}

// ---- Branchformer layer / encoder -------------------------------------------------------------------
size_t smx_branchformer_layer_workspace_bytes(const smx_branchformer_layer_weights* w, int dtype, int32_t B, int32_t T,
                                              int has_sum_mask) {
  if (!w || B <= 0 || T <= 0) return 0;
  Arena a(nullptr, 0, true);
  const float* sm = has_sum_mask ? (const float*)(uintptr_t)256 : nullptr;
  branchformer_layer_generic(w, dtype, B, T, nullptr, nullptr, sm, nullptr, a, nullptr);
  return a.peak;
}
int smx_branchformer_layer_fwd(const smx_branchformer_layer_weights* w, int dtype, int32_t B, int32_t T, const void* x,
                               const uint8_t* padding_mask, const float* sum_mask, void* y, void* workspace,
                               size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!w) return fail(SMX_ERR_BAD_ARG, "weights is NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  size_t need = smx_branchformer_layer_workspace_bytes(w, dtype, B, T, sum_mask != nullptr);
  if (need > workspace_bytes || (need && !workspace))
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  Arena a(workspace, workspace_bytes, false);
  return branchformer_layer_generic(w, dtype, B, T, x, padding_mask, sum_mask, y, a, (cudaStream_t)stream);
}
size_t smx_branchformer_encoder_workspace_bytes(const smx_branchformer_layer_weights* layers, int32_t n_layers,
                                                int dtype, int32_t B, int32_t T, int has_sum_mask) {
  if (!layers || n_layers <= 0 || B <= 0 || T <= 0) return 0;
  size_t m = 0;
  for (int i = 0; i < n_layers; ++i) {
    size_t s = smx_branchformer_layer_workspace_bytes(&layers[i], dtype, B, T, has_sum_mask);
    if (s > m) m = s;
  }
  return m;
}
int smx_branchformer_encoder_fwd(const smx_branchformer_layer_weights* layers, int32_t n_layers,
                                 const float* final_norm_w, const float* final_norm_b, int dtype, int32_t B, int32_t T,
                                 const void* x, const uint8_t* padding_mask, const float* sum_mask, void* y,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(dtype));
  if (!layers || n_layers <= 0) return fail(SMX_ERR_BAD_ARG, "layers is NULL or n_layers <= 0");
  if (!final_norm_w || !final_norm_b) return fail(SMX_ERR_BAD_ARG, "final norm weights are NULL");
  SMX_TRY(check_bt(B, T));
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  size_t need = smx_branchformer_encoder_workspace_bytes(layers, n_layers, dtype, B, T, sum_mask != nullptr);
  if (need > workspace_bytes || (need && !workspace))
    return fail(SMX_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int D = layers[0].branch.pre.in_dim;
  const int64_t rows = (int64_t)B * T;
  const void* cur = x;
  for (int i = 0; i < n_layers; ++i) {
    Arena a(workspace, workspace_bytes, false);
    SMX_TRY(branchformer_layer_generic(&layers[i], dtype, B, T, cur, padding_mask, sum_mask, y, a, st));
    cur = y;
  }
  return layernorm(cur, dtype, D, final_norm_w, final_norm_b, 1e-6f, SMX_ACT_IDENTITY, y, dtype, D, rows, D, st);
}

// ---- acoustic frontend ---------------------------------------------------------------------------------------------------
int32_t smx_fbank_frames(const smx_fbank_desc* d, int32_t n_samples) {
  if (!d || n_samples <= 0) return 0;
  const int hop = (int)lrintf((float)d->sample_rate / 1000.0f * d->hop_length_ms);
  return hop > 0 ? fbank_frames(n_samples, hop) : 0;
}
size_t smx_fbank_workspace_bytes(const smx_fbank_desc* d, int32_t B) { return (d && B > 0) ? fbank_workspace_bytes(B, d->n_mels) : 0; }
int smx_fbank_fwd(const smx_fbank_desc* d, int32_t B, int32_t n_samples, const float* wav, float* feats, void* workspace,
                  size_t workspace_bytes, void* stream) {
  if (!d || B <= 0 || n_samples <= 0) return fail(SMX_ERR_BAD_ARG, "fbank: NULL descriptor or non-positive size");
  SMX_TRY(check_ptr(wav, "wav")); SMX_TRY(check_ptr(feats, "feats"));
  SMX_TRY(check_arch());
  if (!workspace || workspace_bytes < smx_fbank_workspace_bytes(d, B)) return fail(SMX_ERR_WORKSPACE, "fbank: workspace too small");
  return fbank_fwd(d, B, n_samples, wav, feats, workspace, (cudaStream_t)stream);
}
int smx_input_norm_fwd(int64_t rows, int32_t F, const float* x, const float* mean, const float* std, float* y, void* stream) {
  if (rows <= 0 || F <= 0 || !mean || !std) return fail(SMX_ERR_BAD_ARG, "input_norm: bad size or NULL statistics");
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  return input_norm_fwd(rows, F, x, mean, std, y, (cudaStream_t)stream);
}
size_t smx_spec_drop_workspace_bytes(void) { return spec_drop_workspace_bytes(); }
int smx_spec_drop_fwd(int32_t B, int32_t T, int32_t F, float* x, int32_t dim, int32_t n_masks, const int32_t* pos, const int32_t* len,
                      int32_t replace_mean, void* workspace, size_t workspace_bytes, void* stream) {
  if (B <= 0 || T <= 0 || F <= 0 || n_masks < 0 || (n_masks && (!pos || !len))) return fail(SMX_ERR_BAD_ARG, "spec_drop: bad arguments");
  SMX_TRY(check_ptr(x, "x"));
  SMX_TRY(check_arch());
  if (!workspace || workspace_bytes < spec_drop_workspace_bytes()) return fail(SMX_ERR_WORKSPACE, "spec_drop: workspace too small");
  return spec_drop_fwd(B, T, F, x, dim, n_masks, pos, len, replace_mean, workspace, (cudaStream_t)stream);
}
int smx_time_warp_fwd(int32_t B, int32_t T, int32_t F, const float* x, int32_t c, int32_t w, float* y, void* stream) {
  if (B <= 0 || T <= 0 || F <= 0) return fail(SMX_ERR_BAD_ARG, "time_warp: non-positive size");
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  if (x == y) return fail(SMX_ERR_BAD_ARG, "time_warp: y must not alias x");
  SMX_TRY(check_arch());
  return time_warp_fwd(B, T, F, x, c, w, y, (cudaStream_t)stream);
}
int smx_conv_frontend_block_fwd(int32_t B, int32_t T, int32_t F, int32_t Cin, int32_t Cout, int32_t kernel, int32_t stride,
                                const float* x, const float* conv_w, const float* conv_b, const float* ln_w, const float* ln_b,
                                float* y, void* stream) {
  if (B <= 0 || T <= 0 || F <= 0 || Cin <= 0 || Cout <= 0 || kernel <= 0 || stride <= 0 || !conv_w || !ln_w || !ln_b)
    return fail(SMX_ERR_BAD_ARG, "conv frontend block: bad arguments");
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  return conv_block_fwd(B, T, F, Cin, Cout, kernel, stride, x, conv_w, conv_b, ln_w, ln_b, y, (cudaStream_t)stream);
}
size_t smx_input_proj_workspace_bytes(int32_t B, int32_t T, int32_t D) {
  return (B > 0 && T > 0 && D > 0) ? align_up((size_t)B * T * D * 4) + (1 << 16) : 0;
}
int smx_input_proj_fwd(const smx_linear* proj, int32_t B, int32_t T, int32_t max_len, const float* x, int out_dtype, void* y,
                       void* workspace, size_t workspace_bytes, void* stream) {
  SMX_TRY(check_dtype(out_dtype));
  if (!proj || !proj->w) return fail(SMX_ERR_BAD_ARG, "input_proj: NULL weights");
  SMX_TRY(check_bt(B, T));
  if (max_len > 0 && T > max_len)  /* the reference's table has max_len rows: the broadcast add fails there (Transformer.py:339) */
    return fail(SMX_ERR_BAD_ARG, "input_proj: sequence length %d exceeds the positional table (%d)", T, max_len);
  SMX_TRY(check_ptr(x, "x")); SMX_TRY(check_ptr(y, "y"));
  SMX_TRY(check_arch());
  const int D = proj->out_dim;
  if (!workspace || workspace_bytes < smx_input_proj_workspace_bytes(B, T, D)) return fail(SMX_ERR_WORKSPACE, "input_proj: workspace too small");
  Arena a(workspace, workspace_bytes, false);
  float* v = a.f32((size_t)B * T * D);
  if (!v) return fail(SMX_ERR_WORKSPACE, "input_proj: workspace too small");
  SMX_TRY(vanilla_generic(proj, 1, SMX_ACT_IDENTITY, x, SMX_F32, proj->in_dim, (int64_t)B * T, nullptr, nullptr, 0, 0, v, SMX_F32, D, a,
                          (cudaStream_t)stream));
  return posenc_add(v, B, T, D, y, out_dtype, (cudaStream_t)stream);
}

}  // extern "C"
