// Declarations of the tcgen05 (bf16 tensor-core) arm of libsmx.
#pragma once
#include <cuda.h>
#include <utility>
#include "smx_internal.h"

namespace smx {

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------
// The fused kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization: every kernel executes
// griddepcontrol.launch_dependents when it starts and griddepcontrol.wait before its first access to activations, so the
// next kernel's CTAs take over SMs as soon as this kernel's CTAs leave them and run their set-up (barrier init, TMEM
// allocation, parameter staging) and their launch latency under the tail of this kernel instead of after it.
// Weights / parameters are never written by a kernel, so reading them before the wait is safe.
bool tc_pdl_enabled();
void tc_set_pdl(int on);
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster_x,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (tc_pdl_enabled()) { attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
  if (cluster_x > 1) { attr[n].id = cudaLaunchAttributeClusterDimension; attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1; ++n; }
  cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- smx_tc.cu: packing for the self-test -------------------------------------------------------
size_t tc_packed_bytes(int N, int K, int NT);
int tc_pack_weight(const float* w, int64_t stride_n, int64_t stride_k, int N, int K, int NT, int layout,
                   __nv_bfloat16* out, cudaStream_t st);

// ---- smx_tc_lin.cu: K-LIN, the row-tile linear kernel -------------------------------------------
struct LinP {
  const __nv_bfloat16* x; int64_t ldx;   // input rows (bf16)
  int64_t rows;                          // number of rows (B*T)
  int T;                                 // frames per utterance: row group for rowbias; tile alignment if utt_tiles
  int utt_tiles;                         // 1: tiles are utterance-aligned (ceil(T/128) tiles per utterance)
  int K, N;                              // GEMM dims
  int head_in, head_out;                 // block-diagonal structure (0 = dense)
  const __nv_bfloat16* wp;               // packed weight image (tc_pack_linear)
  const float* ln_w; const float* ln_b; float ln_eps;   // prologue LayerNorm over K (NULL = none)
  const float* bias;                     // [N]
  const float* rowbias; int64_t rowbias_ld;             // [rows/T][N] added before the activation
  int act;
  const uint8_t* rowmask;                // [rows] multiplies the activated value
  const __nv_bfloat16* resid; int64_t ldr; float alpha; // out = resid + alpha * v
  const float* oln_w; const float* oln_b; float oln_eps; // LIN_OLN: LayerNorm over the output row
  __nv_bfloat16* out; int64_t ldo;
  float* colsum;                         // LIN_COLSUM: [tiles][N] masked column sums
  // filled by tc_linear_launch
  int NT, n_tiles, glu, stage_bytes, n_stages; uint32_t tmem_cols;
};
enum { TC_LIN_PLAIN = 0, TC_LIN_GLU = 1, TC_LIN_OLN = 2, TC_LIN_COLSUM = 3 };
bool tc_linear_supported(int K, int N);
size_t tc_linear_packed_bytes(int K, int N);
int tc_pack_linear(const smx_linear& L, int k_offset, int K, int glu, void* out, cudaStream_t st);
int tc_pack_linear_nt(const smx_linear& L, int k_offset, int K, int NT, void* out, cudaStream_t st,
                      int glu_interleave = 0);  // explicit n-tile width; value/gate 64-row blocks interleaved for GLU
int tc_pick_nt(int N, int glu);
int tc_linear_launch(LinP p, int mode, cudaStream_t st);

// ---- smx_tc_path.cu: module forwards composed from tcgen05 kernels --------------------------------
// Each *_supported() says whether the bf16 tensor-core arm handles the configuration; *_packed_bytes /
// *_pack build the bf16 operand images once per weight set; *_ws sizes the workspace.
bool tc_cell_supported(const smx_cell_weights* w, int has_sum_mask);
size_t tc_cell_packed_bytes(const smx_cell_weights* w);
int tc_cell_pack(const smx_cell_weights* w, void* packed, cudaStream_t st);
size_t tc_cell_workspace_bytes(const smx_cell_weights* w, int B, int T);
// x: rows to mix (bf16); pre_ln_*: LayerNorm applied to x first (norm1 of the encoder layer) or NULL
int tc_cell_fwd(const smx_cell_weights* w, const void* packed, int B, int T, const __nv_bfloat16* x,
                const float* pre_ln_w, const float* pre_ln_b, const uint8_t* mask, const __nv_bfloat16* residual,
                __nv_bfloat16* y, Arena& ws, cudaStream_t st);

int tc_add_bcast(const __nv_bfloat16* a, const __nv_bfloat16* s, int64_t rows, int T, int D, __nv_bfloat16* out, cudaStream_t st);

// ---- smx_tc_cell.cu: K-SM, the fused persistent cell (two passes) ---------------------------------
void tc_set_trace(void* p);  // debug timeline buffer (>= 1024 x u64) or NULL
bool tc_cellf_supported(const smx_cell_weights* w);
size_t tc_cellf_workspace_bytes(const smx_cell_weights* w, int B, int T);
int tc_cellf_fwd(const smx_cell_weights* w, const void* img_s1, const void* img_s2, const void* img_f1, const void* img_f2,
                 const void* img_c, int B, int T, const __nv_bfloat16* x, const float* pre_ln_w, const float* pre_ln_b,
                 const uint8_t* mask, const __nv_bfloat16* residual, __nv_bfloat16* y, Arena& ws, cudaStream_t st);

int tc_glu_fwd(const smx_linear& L, const void* img, const float* ln_w, const float* ln_b, int64_t rows,
               const __nv_bfloat16* x, __nv_bfloat16* out, cudaStream_t st);
int tc_cell_finalize(const smx_cell_weights* w, int B, int T, const float* colsum, const uint8_t* mask, float* rowbias, cudaStream_t st);

// ---- smx_tc_cell3.cu: K-SM v3 (operands resident in tensor memory, step-granular weight ring, 16 epilogue warps) -----
// images in SCHEDULE order (tc_cell3_reorder from the [chunk][K-block] images of tc_pack_linear_nt)
bool tc_cell3_supported(const smx_cell_weights* w);
int tc_cell3_reorder(const smx_linear& L, int K, int n_split, const void* img_chunk_major, void* img_sched, cudaStream_t st);
int tc_cell3_fwd(const smx_cell_weights* w, const void* img_s1, const void* img_s2, const void* img_f1, const void* img_f2,
                 const void* img_c, int B, int T, const __nv_bfloat16* x, const float* pre_ln_w, const float* pre_ln_b,
                 const uint8_t* mask, const __nv_bfloat16* residual, __nv_bfloat16* y, Arena& ws, cudaStream_t st);
// smx_tc_glu4.cu: K-GLU v4 (D = 256, <= 2 row tiles per CTA): both tiles resident, one pass over the weights, LayerNorm folded
bool tc_cell4_prenorm_ok(const smx_cell_weights* w);
int tc_cell4_pack_prenorm(const smx_cell_weights* w, void* img, const float* norm_w, const float* norm_b, cudaStream_t st);
int tc_cell_pack_prenorm(smx_cell_weights* w, const float* norm_w, const float* norm_b, cudaStream_t st);
bool tc_glu4_supported(const smx_convmod_weights* w);
bool tc_glu4_fits(int64_t rows);
size_t tc_glu4_packed_bytes(const smx_convmod_weights* w);
int tc_glu4_pack(const smx_convmod_weights* w, void* out, cudaStream_t st);
int tc_glu4_fwd(const smx_convmod_weights* w, const void* img, int64_t rows, const __nv_bfloat16* x, __nv_bfloat16* g, cudaStream_t st);
int tc_glu3_fwd(const smx_linear& L, const void* img_sched, const float* ln_w, const float* ln_b, int64_t rows,
                const __nv_bfloat16* x, __nv_bfloat16* out, cudaStream_t st);
void tc_set_cell_version(int v);  // 1, 3 or 4 (diagnostics / A-B timing)
void tc_set_trace_cell3(void* p);
int tc_cell_version();

// ---- smx_tc_cell4.cu: K-SM v4 (one persistent kernel: x tile by tensor-map TMA, normalised once and resident through both
// phases, per-utterance counter / flag instead of a kernel boundary, half-width chains ping-ponging in TMEM) --------------
bool tc_cell4_supported(const smx_cell_weights* w);
bool tc_cell4_fits(int B, int T);  // at most two tiles per CTA, one CTA per SM
size_t tc_cell4_packed_bytes(const smx_cell_weights* w);
int tc_cell4_pack(const smx_cell_weights* w, const void* img_s1, const void* img_s2, const void* img_f1, const void* img_f2,
                  const void* img_c, void* out, cudaStream_t st);  // inputs: chunk-major images (tc_pack_linear_nt, NT = 64)
size_t tc_cell4_workspace_bytes(const smx_cell_weights* w, int B, int T);
size_t tc_cell4_sync_bytes(int B);
void tc_cell4_set_presync(void* zeroed, size_t bytes);  // (thread-local) counters zeroed ahead of the kernel chain by the caller
int tc_cell4_fwd(const smx_cell_weights* w, const void* img, int B, int T, const __nv_bfloat16* x, const float* pre_ln_w,
                 const float* pre_ln_b, const uint8_t* mask, const __nv_bfloat16* residual, __nv_bfloat16* y, Arena& ws, cudaStream_t st);
void tc_set_trace_cell4(void* p);

// ---- smx_tc_conv.cu: K-CONV, depthwise conv + LN + act + output GEMM + mask/residual, persistent -------
bool tc_convf_supported(const smx_convmod_weights* w, int chunk);
int tc_convf_second_half(const smx_convmod_weights* w, const void* img_out, int act, int B, int T, const __nv_bfloat16* g,
                         const uint8_t* mask, const __nv_bfloat16* residual, __nv_bfloat16* y, cudaStream_t st);

bool tc_ffn_supported(const smx_ffn_weights* w);
size_t tc_ffn_packed_bytes(const smx_ffn_weights* w);
int tc_ffn_pack(const smx_ffn_weights* w, void* packed, cudaStream_t st);
size_t tc_ffn_workspace_bytes(const smx_ffn_weights* w, int64_t rows);
int tc_ffn_fwd(const smx_ffn_weights* w, const void* packed, int act, int64_t rows, const __nv_bfloat16* x,
               const float* oln_w, const float* oln_b, float oln_eps, __nv_bfloat16* y, Arena& ws, cudaStream_t st);

// smx_tc_ffn2.cu: persistent K-FFN (preferred when supported)
bool tc_ffn2_supported(const smx_ffn_weights* w);
size_t tc_ffn2_packed_bytes(const smx_ffn_weights* w);
int tc_ffn2_pack(const smx_ffn_weights* w, void* packed, cudaStream_t st);
int tc_ffn2_fwd(const smx_ffn_weights* w, const void* packed, int act, int64_t rows, const __nv_bfloat16* x,
                const float* oln_w, const float* oln_b, float oln_eps, __nv_bfloat16* y, cudaStream_t st);
// smx_tc_ffn3.cu: K-FFN v3, hidden activation resident in tensor memory (preferred; same packed images as v2)
// bf16 tensor map (cuTensorMapEncodeTiled through the runtime's driver entry point; smx_tc_cell4.cu): rank 2 or 3, dims / box
// innermost first, strides in bytes for dims 1.., 128-byte swizzle (box[0] = 64 columns: one UMMA K-block).  false: unavailable / rejected
bool tc_encode_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);
// a packed operand image (8 KB blocks of 64 rows x 128 bytes, already in the 128-byte-swizzled UMMA layout) as a [64 * n_blocks][64] bf16
// tensor without swizzle, box = one block: lets cp.async.bulk.tensor (with .cta_group::2: completion on the pair leader's barrier) copy blocks as they are
bool tc_encode_tmap_image(CUtensorMap* m, const void* image, uint64_t n_blocks);
// W gamma (fp32), gw[n] = sum_k bf16(W[n,k] gamma_k), bw[n] = sum_k beta_k W[n,k]  (smx_tc_cell4.cu; gamma / beta NULL = no LayerNorm)
int tc_fold_ln(const float* W, int K, int ldw, int N, const float* gamma, const float* beta, float* Wg, float* gw, float* bw, cudaStream_t st);
bool tc_ffn3_supported(const smx_ffn_weights* w);
size_t tc_ffn3_packed_bytes(const smx_ffn_weights* w);  // v2 images + the same blocks in ring-step order
int tc_ffn3_pack(const smx_ffn_weights* w, void* packed, cudaStream_t st);
int tc_ffn3_fwd(const smx_ffn_weights* w, const void* packed, int act, int64_t rows, const __nv_bfloat16* x,
                const float* oln_w, const float* oln_b, float oln_eps, __nv_bfloat16* y, cudaStream_t st);
void tc_set_ffn_version(int v);  // 2 or 3 (diagnostics / A-B timing)
int tc_ffn_version();
void tc_set_trace_ffn3(void* p);
void tc_set_trace_ffn(void* p);
void tc_set_trace_conv(void* p);
void tc_set_ffn_cluster(int cl);  // 1, 2 or 4 CTAs share each weight block (diagnostics / tuning)

bool tc_convmod_supported(const smx_convmod_weights* w, int chunk);
size_t tc_convmod_packed_bytes(const smx_convmod_weights* w);
int tc_convmod_pack(const smx_convmod_weights* w, void* packed, cudaStream_t st);
size_t tc_convmod_workspace_bytes(const smx_convmod_weights* w, int B, int T);
int tc_convmod_fwd(const smx_convmod_weights* w, const void* packed, int act, int B, int T, const __nv_bfloat16* x,
                   const uint8_t* mask, const __nv_bfloat16* residual, __nv_bfloat16* y, Arena& ws, cudaStream_t st);

// depthwise conv (zero 'same' padding) + LayerNorm + activation over bf16 rows (smx_tc_path.cu)
int tc_dwconv_ln_act(const __nv_bfloat16* g, const float* dw_w, const float* dw_b, const float* ln_w,
                     const float* ln_b, int act, int B, int T, int D, int k, __nv_bfloat16* out, cudaStream_t st);

}  // namespace smx

namespace smx {
// ---- smx_tc_gemm.cu: K-GEMM (both operands streamed by tensor-map TMA, any K) and the fused CSGU gate ------------------
struct GemmTc {
  const __nv_bfloat16* a; int64_t lda; int64_t M;   // A (M, K) bf16 rows, row pitch lda elements
  int N, K;
  const __nv_bfloat16* w;                            // W (N, K) bf16 row-major (tc_dense_bf16)
  const float* bias;                                 // [N] or NULL
  const float* rowbias; int64_t rowbias_ld; int rows_per_group;   // + rowbias[row / rows_per_group][n] before the activation, or NULL
  int act;
  const uint8_t* rowmask;                            // [M] multiplies the activated value, or NULL
  const __nv_bfloat16* resid; int64_t ldr; float alpha;   // out = resid + alpha * v, or NULL
  __nv_bfloat16* out; int64_t ldo;
  float* out_f32 = nullptr; const float* resid_f32 = nullptr;   // fp32 output / residual instead of out / resid
  int split3 = 0;                                     // set by tc_linear_split3
  int ksplit = 1; int64_t out_split_stride = 0;       // split-K: fp32 partials out_f32 + ks * out_split_stride (summed by the caller)
  int glu = 0;                                        // GLU epilogue over an interleaved image (tc_glu_dense_bf16): N = 2 x output width
  int bd_in = 0, bd_out = 0;                          // W is block-diagonal (ParallelLinear as dense with zero blocks): per-head input / output width
};
bool tc_gemm_supported(int K, int N);
// fp32 linears on the tensor cores with split-bf16 operands (smx_tc_gemm.cu)
size_t tc_split3_scratch_bytes(int64_t rows, int K, int N);
bool tc_split3_ok(int64_t rows, int K, int N);
bool tc_f32_tc_enabled();
void tc_set_f32_tc(int on);
int tc_linear_split3(const smx_linear& L, int k_offset, int K, const float* A, int64_t lda, int64_t rows, GemmTc g, void* scratch, cudaStream_t st);
int tc_dgrad_split3(const smx_linear& L, int k_offset, int Kin, const float* dZ, int64_t ldz, int64_t rows, GemmTc g, void* scratch, cudaStream_t st);
size_t tc_wgrad_scratch_bytes(int64_t rows, int M, int N);
int tc_wgrad_slices(int64_t rows);
int tc_wgrad_split3(const float* A, int64_t lda, int M, const float* Bm, int64_t ldb, int N, int64_t rows, float* P, int* ns, void* scratch,
                    cudaStream_t st);  // P[slice][M][N] = partial sums over row slices of A^T B
int tc_gemm_launch(const GemmTc& g, cudaStream_t st);
int tc_dense_bf16(const smx_linear& L, int k_offset, int K, void* out, cudaStream_t st);
size_t tc_glu_dense_bytes(int K, int N);
int tc_glu_dense_bf16(const smx_linear& L, void* out, cudaStream_t st);  // [bf16 image | fp32 bias], rows interleaved per 256-wide N tile for the GLU epilogue  // (out_dim, K) bf16 copy of columns [k_offset, +K)
size_t tc_csgu_workspace_bytes(int64_t rows);
int tc_csgu_fwd(const __nv_bfloat16* u, int B, int T, int H, const float* ln_w, const float* ln_b, const float* dw_w, const float* dw_b,
                int kernel_size, int gate_act, __nv_bfloat16* out, int64_t ldo, void* stats_ws, cudaStream_t st);

// ---- smx_tc_branch.cu: BranchformerEncoderLayer on the tensor-core arm ------------------------------------------------
bool tc_branchformer_supported(const smx_branchformer_layer_weights* w, int has_sum_mask);
size_t tc_branchformer_packed_bytes(const smx_branchformer_layer_weights* w);
int tc_branchformer_pack(const smx_branchformer_layer_weights* w, void* packed, cudaStream_t st);
int tc_branchformer_layer_fwd(const smx_branchformer_layer_weights* w, int B, int T, const __nv_bfloat16* x, const uint8_t* mask,
                              __nv_bfloat16* y, Arena& ws, cudaStream_t st);
}  // namespace smx
