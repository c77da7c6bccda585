// Internal declarations shared by the libsmx translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include "smx.h"

namespace smx {

// ---- error plumbing ---------------------------------------------------------------------
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);
void count_tc_launch(int n = 1);
int check_launch(const char* what);  // cudaGetLastError -> SMX_ERR_CUDA

#define SMX_TRY(expr)                \
  do {                               \
    int _s = (expr);                 \
    if (_s != SMX_OK) return _s;     \
  } while (0)

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  size_t peak;
  bool dry;  // size computation only: nothing is launched, take() returns a dummy
  Arena(void* p, size_t c, bool dry_run) : base((char*)p), cap(c), off(0), peak(0), dry(dry_run) {}
  void* take(size_t bytes) {
    size_t o = off;
    off += align_up(bytes);
    if (off > peak) peak = off;
    if (dry) return (void*)(uintptr_t)256;  // non-null dummy, never dereferenced
    if (off > cap) return nullptr;
    return base + o;
  }
  float* f32(size_t n) { return (float*)take(n * sizeof(float)); }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }  // stream order makes reuse after release safe
};

inline size_t elem_size(int dtype) { return dtype == SMX_BF16 ? 2 : 4; }

// ---- generic (fp32 math) kernels: smx_simt.cu -------------------------------------------
struct GemmP {
  const void* A;  int a_dtype;  int64_t lda;  int64_t a_bs;     // A[m*lda + k*a_sk] (+ batch*a_bs)
  int64_t a_sk;                                                  // k stride of A (0 means 1); != 1: A is read transposed
  int k_total;                                                   // > 0: split-K over batches, batch i covers k in [i*K, min((i+1)*K, k_total))
  const float* W;  int64_t w_sk;  int64_t w_sn;  int64_t w_bs;  // W[k*w_sk + n*w_sn] (+ batch*w_bs)
  const float* bias;  int64_t bias_bs;                           // bias[n] (+ batch*bias_bs)
  const float* rowbias;  int64_t rowbias_ld;  int32_t rowbias_div;  // + rowbias[(m/div)*ld + n]
  const float* rowdiv;                                           // acc /= rowdiv[m] (before bias)
  const uint8_t* rowmask;                                        // (after act) *= rowmask[m]
  const void* residual;  int r_dtype;  int64_t ldr;  float alpha; // out = residual + alpha*v
  void* C;  int c_dtype;  int64_t ldc;  int64_t c_bs;
  int M, N, K, batches, act;
  // optional outer batch level (0 means 1): blockIdx.z = outer * batches + inner; the outer index adds these offsets
  int batches2;  int64_t a_bs2;  int64_t w_bs2;  int64_t c_bs2;
  const int* run_if_nonzero;                                     // device flag (NULL: always run): the kernel returns at once when *flag == 0
};
int gemm(const GemmP& p, cudaStream_t st);
int layernorm(const void* x, int x_dtype, int64_t ldx, const float* w, const float* b, float eps, int act,
              void* y, int y_dtype, int64_t ldy, int64_t rows, int D, cudaStream_t st);
int masked_mean(const float* s, int64_t lds, const uint8_t* mask, int B, int T, int D, void* out, int out_dtype,
                cudaStream_t st);
int glu(const float* p, int64_t rows, int D, float* out, cudaStream_t st);
int dwconv(const float* in, int64_t ldin, const float* w, const float* b, int B, int T, int C, int k, int pad_mode,
           int chunk, float* out, int64_t ldout, cudaStream_t st);
bool dwconv_window(const float* in, int64_t ldin, const float* w, const float* b, int B, int T, int C, int k, int pad, int reflect, int flip,
                   float* out, int64_t ldout, cudaStream_t st, int* status);
int gate_mul(const float* gate, int64_t ldg, const float* other, int64_t ldo, int gate_act, int64_t rows, int C,
             float* out, cudaStream_t st);
int broadcast_rows(const float* src, int B, int T, int D, float* dst, int64_t lddst, cudaStream_t st);
int add_bcast(const float* a, const float* s, int64_t rows, int div, int D, float* out, cudaStream_t st);
int laplace(float decay, const float* binary, int T, float* out, cudaStream_t st);
int rowsum(const float* m, int rows, int cols, float* out, cudaStream_t st);
// Per-frame summaries for a (T,T) sum mask whose rows are runs of ones (the dynamic-chunk masks of TransformerASR.py:85-110):
// interval_detect finds every row's [lo, hi) and clears *not_interval... sets it to 1 when some row is not a single run of
// exact ones; interval_means then gives Sm[b,t,:] = sum_{j in [lo_t, hi_t)} S[b,j,:] / (hi_t - lo_t) from prefix sums over time
// (fp64 running sums) -- O(T D) per utterance instead of the (T,T) @ (T,D) product.  Both are no-ops once the flag is set.
int interval_detect(const float* M, int T, int* lo, int* hi, int* not_interval, cudaStream_t st);
size_t interval_means_workspace_bytes(int B, int T, int D);
int interval_transpose(const int* lo, const int* hi, int T, int* tlo, int* thi, int* not_interval, cudaStream_t st);
int interval_sums(const float* S, int64_t ldS, int B, int T, int D, const int* lo, const int* hi, const int* not_interval, float* out, int64_t ldo,
                  void* workspace, cudaStream_t st);  // the gradient form: sums over the (transposed) intervals, workspace as interval_means
int interval_means(const float* S, int64_t ldS, int B, int T, int D, const int* lo, const int* hi, const int* not_interval, float* Sm,
                   void* workspace, cudaStream_t st);
int convert(const void* src, int s_dtype, void* dst, int d_dtype, int64_t n, cudaStream_t st);

// ---- frontend: smx_frontend.cu -------------------------------------------------------------------
int fbank_frames(int n_samples, int hop);
size_t fbank_workspace_bytes(int B, int n_mels);
int fbank_fwd(const smx_fbank_desc* d, int B, int n_samples, const float* wav, float* feats, void* workspace, cudaStream_t st);
int input_norm_fwd(int64_t rows, int F, const float* x, const float* mean, const float* stdv, float* y, cudaStream_t st);
size_t spec_drop_workspace_bytes();
int spec_drop_fwd(int B, int T, int F, float* x, int dim, int n_masks, const int* pos, const int* len, int replace_mean, void* workspace,
                  cudaStream_t st);
int time_warp_fwd(int B, int T, int F, const float* x, int c, int w, float* y, cudaStream_t st);
int conv_block_fwd(int B, int T, int F, int Cin, int Cout, int ks, int stride, const float* x, const float* cw, const float* cb, const float* lw,
                   const float* lb, float* y, cudaStream_t st);
int posenc_add(const float* v, int B, int T, int D, void* y, int y_dt, cudaStream_t st);

// host orchestration of the generic path (smx_generic.cu).  x/y/residual carry their own dtype tags.
int vanilla_generic(const smx_linear* blocks, int n, int act, const void* x, int x_dt, int64_t ldx, int64_t rows,
                    const uint8_t* rowmask, const void* residual, int r_dt, int64_t ldr, void* y, int y_dt,
                    int64_t ldy, Arena& ws, cudaStream_t st);
int cell_generic(const smx_cell_weights* w, int B, int T, const void* x, int x_dt, const uint8_t* mask,
                 const float* sum_mask, const void* residual, int r_dt, void* y, int y_dt, int64_t ldy, Arena& ws,
                 cudaStream_t st);
// backward of the cell, mode "SummaryMixing" (smx_bwd.cu); recomputes the forward intermediates from x
int cell_bwd_generic(const smx_cell_weights* w, int B, int T, const void* x, int x_dt, const uint8_t* mask, const void* dy, int dy_dt,
                     void* dx, int dx_dt, const smx_cell_grads* g, Arena& ws, cudaStream_t st, const smx_dropout* drop = nullptr,
                     void* y_fwd = nullptr, int y_dt = 0,   // drop: training-mode dropout; y_fwd: forward only (smx_*_train_fwd)
                     const float* sum_mask = nullptr);      // (T,T) fp32: Dynamic Chunk Training (per-frame summaries)
int vanilla_bwd_generic(const smx_linear* blocks, int n, int act, int64_t rows, const void* x, int x_dt, const void* dy, int dy_dt,
                        void* dx, int dx_dt, const smx_linear_grad* g, Arena& ws, cudaStream_t st);
int layernorm_bwd_generic(const void* x, int x_dt, int64_t rows, int D, const float* w, float eps, const void* dy, int dy_dt, void* dx,
                          int dx_dt, float* dw, float* db, Arena& ws, cudaStream_t st);
int ffn_bwd_generic(const smx_ffn_weights* w, int act, int64_t rows, const void* x, int x_dt, const float* oln_w, const float* oln_b,
                    float oln_eps, const void* dy, int dy_dt, void* dx, int dx_dt, const smx_ffn_grads* g, Arena& ws, cudaStream_t st,
                    const smx_dropout* drop = nullptr, void* y_fwd = nullptr, int y_dt = 0);
int convbranch_bwd_generic(const smx_convbranch_weights* w, int B, int T, const void* x, int x_dt, const void* dy, int dy_dt, void* dx,
                           int dx_dt, const smx_convbranch_grads* g, Arena& ws, cudaStream_t st, const smx_dropout* drop = nullptr,
                           void* y_fwd = nullptr, int y_dt = 0);
int dropout_apply(const smx_dropout* drop, int site, int dt, int64_t n, const void* x, void* y, cudaStream_t st);
int convmod_bwd_generic(const smx_convmod_weights* w, int act, int B, int T, const void* x, int x_dt, const uint8_t* mask, const void* dy,
                        int dy_dt, void* dx, int dx_dt, const smx_convmod_grads* g, Arena& ws, cudaStream_t st,
                        const smx_dropout* drop = nullptr, void* y_fwd = nullptr, int y_dt = 0, int chunk = 0);
int dropout_keep_mask(const smx_dropout* drop, int site, int64_t n, uint8_t* keep, cudaStream_t st);
int ffn_generic(const smx_ffn_weights* w, int act, int64_t rows, const void* x, int x_dt, const float* oln_w,
                const float* oln_b, float oln_eps, void* y, int y_dt, Arena& ws, cudaStream_t st);
int convmod_generic(const smx_convmod_weights* w, int act, int B, int T, int chunk, const void* x, int x_dt,
                    const uint8_t* mask, const void* residual, int r_dt, void* y, int y_dt, Arena& ws,
                    cudaStream_t st);
int mixing_block_generic(const smx_cell_weights* cw, const float* norm_w, const float* norm_b, int dtype, int B, int T, const void* x1,
                         const uint8_t* mask, const float* sum_mask, void* x2, Arena& ws, cudaStream_t st);
int conformer_layer_generic(const smx_conformer_layer_weights* w, int dtype, int B, int T, int chunk, const void* x,
                            const uint8_t* mask, const float* sum_mask, void* y, Arena& ws, cudaStream_t st);
int branchformer_layer_generic(const smx_branchformer_layer_weights* w, int dtype, int B, int T, const void* x,
                               const uint8_t* mask, const float* sum_mask, void* y, Arena& ws, cudaStream_t st);

}  // namespace smx
