/*
 * smx.h — C ABI of libsmx, the B200 (sm_100a) SummaryMixing encoder-path library.
 *
 * The reference (SamsungLabs/SummaryMixing @ d1b1f42) has no FFI: its operator API is the Python
 * nn.Module surface.  This header is the boundary a maintainer binds from those modules (ctypes stub in
 * INTEGRATION.md); every entry point cites the reference code it replaces (paths relative to the
 * reference root).
 *
 * Conventions (all entry points)
 *  - extern "C", plain pointers and sizes.  Every pointer except the `_host` entry points' buffers is
 *    a DEVICE pointer.  The library allocates nothing persistent and keeps no pointer after return.
 *  - Enqueue-only on `stream` (a cudaStream_t passed as void*); no host synchronisation; safe to
 *    capture in a CUDA graph.  Re-entrant: no entry point keeps state between calls.  The only process-wide
 *    state is the smx_debug_set_* switches at the end of this header (atomics; a concurrent call sees the old or the
 *    new value) and the launch counters.
 *  - Returns SMX_OK (0) or a negative smx_status; never throws, never aborts.  smx_last_error()
 *    returns a thread-local message for the last failure.
 *  - Activations x/y: row-major (B,T,D) contiguous, fp32 (SMX_F32) or bf16 (SMX_BF16), 16-byte aligned.
 *    Weights: fp32, in the reference's own state_dict layouts (nn.Linear: (out,in); ParallelLinear:
 *    (h, in/h, out/h), VanillaNN.py:85-88).  The bf16 tensor-core path additionally takes a packed
 *    image of the weights built once by smx_*_pack().
 *  - padding mask: uint8 (B,T), 1 = valid frame (TransformerASR.py:158-162, 348-349); NULL = all valid.
 *  - sum mask: fp32 (T,T), row t = weights of the frames visible to frame t (summary_mixing.py:188-189,
 *    235-246); NULL = whole-utterance mean.
 */
#ifndef SMX_H_
#define SMX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SMX_API __attribute__((visibility("default")))
#else
#define SMX_API
#endif

#define SMX_VERSION 100 /* 0.1.0 */
#define SMX_MAX_BLOCKS 4 /* max linear blocks per VanillaNN handled by the library */

typedef enum {
  SMX_OK = 0,
  SMX_ERR_BAD_ARG = -1,      /* NULL pointer, non-positive size, unknown enum */
  SMX_ERR_UNSUPPORTED = -2,  /* valid in the reference but not handled by this build (message says what) */
  SMX_ERR_ALIGNMENT = -3,    /* pointer not 16-byte aligned */
  SMX_ERR_WORKSPACE = -4,    /* workspace too small */
  SMX_ERR_CUDA = -5,         /* a CUDA runtime call or launch failed (message has cudaGetErrorString) */
  SMX_ERR_ARCH = -6          /* device is not compute capability 10.x */
} smx_status;

typedef enum { SMX_F32 = 0, SMX_BF16 = 1 } smx_dtype;

/* activation of VanillaNN / FFN / conv module (a torch.nn.Module class in the reference,
 * summary_mixing.py:86; Conformer.py:117,404; Branchformer.py:68-69) */
typedef enum {
  SMX_ACT_IDENTITY = 0,
  SMX_ACT_SWISH = 1,      /* speechbrain Swish / nn.SiLU: x*sigmoid(x) */
  SMX_ACT_GELU = 2,       /* nn.GELU(): exact erf form */
  SMX_ACT_RELU = 3,
  SMX_ACT_LEAKY_RELU = 4, /* negative_slope 0.01 */
  SMX_ACT_TANH = 5,
  SMX_ACT_SIGMOID = 6,
  SMX_ACT_GELU_TANH = 7   /* nn.GELU(approximate="tanh") */
} smx_act;

/* SummaryMixing.mode, summary_mixing.py:93-101 */
typedef enum { SMX_MODE_FULL = 0, SMX_MODE_LITE = 1, SMX_MODE_FAST = 2, SMX_MODE_EXPDECAY = 3 } smx_mode;

/* depthwise-convolution boundary rule */
typedef enum {
  SMX_CONV_SAME_ZERO = 0, /* Conformer.py:142-151: zero padding (k-1)/2 both sides */
  SMX_CONV_CAUSAL = 1,    /* Conformer.py:130-131, 327-329: left padding k-1, chomp */
  SMX_CONV_CHUNKED = 2,   /* Conformer.py:197-320: Dynamic Chunk Convolution (future beyond own chunk masked) */
  SMX_CONV_SAME_REFLECT = 3 /* speechbrain CSGU depthwise conv: reflect padding */
} smx_conv_pad;

/* One linear block: n_split == 1 → dense nn.Linear, w (out_dim,in_dim); n_split > 1 → ParallelLinear,
 * w (n_split, in_dim/n_split, out_dim/n_split) (VanillaNN.py:85-88, 112).  b has out_dim entries. */
typedef struct {
  const float* w;
  const float* b;
  int32_t in_dim;
  int32_t out_dim;
  int32_t n_split;
  int32_t _pad;
} smx_linear;

/* SummaryMixing cell parameters (summary_mixing.py:78-167). */
typedef struct {
  int32_t mode;            /* smx_mode */
  int32_t act;             /* smx_act */
  int32_t use_layernorm;   /* summary_mixing.py:163-165 */
  int32_t enc_dim;
  int32_t local_out_dim;   /* D_l */
  int32_t summary_out_dim; /* D_s */
  int32_t n_local;         /* blocks in local_proj (FULL/EXPDECAY) */
  int32_t n_summary;       /* blocks in summary_proj (FULL/EXPDECAY/LITE) */
  smx_linear local[SMX_MAX_BLOCKS];
  smx_linear summary[SMX_MAX_BLOCKS];
  smx_linear global_proj;  /* FAST: dense D -> 2*D_l (summary_mixing.py:133-140) */
  smx_linear merge;        /* summary_local_merging: dense (D_l+D_s) -> D_s (FULL/EXPDECAY/FAST) */
  const float* local_norm_w;
  const float* local_norm_b;
  const float* summary_norm_w;
  const float* summary_norm_b;
  const void* packed;      /* bf16 operand image from smx_cell_pack(), or NULL (generic arm) */
  float decay_constant;    /* EXPDECAY (summary_mixing.py:158-161) */
  int32_t _pad;
  const float* prenorm_w;  /* set by smx_cell_pack_prenorm(): the LayerNorm in front of the cell whose gamma / beta are folded */
  const float* prenorm_b;  /* into `packed`; NULL = none.  Used only by calls that pass exactly these pointers as the pre-norm */
} smx_cell_weights;

/* ffn_module{1,2}: LayerNorm + PositionalwiseFeedForward (Conformer.py:470-484). */
typedef struct {
  const float* ln_w;
  const float* ln_b;
  smx_linear w1; /* D -> d_ffn */
  smx_linear w2; /* d_ffn -> D */
  const void* packed; /* bf16 operand image from smx_ffn_pack(), or NULL (generic arm) */
} smx_ffn_weights;

/* ConvolutionModule (Conformer.py:80-340). */
typedef struct {
  const float* ln_w;
  const float* ln_b;
  smx_linear bottleneck;  /* pointwise Conv1d D -> 2D, weight (2D,D,1) viewed (2D,D) */
  const float* dw_w;      /* depthwise weight (D,1,k) viewed (D,k) */
  const float* dw_b;      /* (D) or NULL */
  const float* after_ln_w;
  const float* after_ln_b;
  smx_linear out;         /* after_conv.2: D -> D */
  const void* packed;     /* bf16 operand image from smx_convmod_pack(), or NULL (generic arm) */
  int32_t kernel_size;
  int32_t causal;
} smx_convmod_weights;

/* ConformerEncoderLayer with attention_type == "SummaryMixing" (Conformer.py:343-548). */
typedef struct {
  smx_ffn_weights ffn1;
  smx_ffn_weights ffn2;
  const float* norm1_w;
  const float* norm1_b;
  const float* norm2_w;
  const float* norm2_b;
  smx_cell_weights cell;
  smx_convmod_weights conv;
  int32_t act; /* shared by FFN, cell and conv module (Conformer.py:446-484) */
  int32_t _pad;
} smx_conformer_layer_weights;

/* ConvolutionBranch + CSGU (Branchformer.py:31-97; speechbrain ConvolutionalSpatialGatingUnit). */
typedef struct {
  smx_linear pre;          /* pre_channel_proj D -> U */
  smx_linear post;         /* post_channel_proj U/2 -> D */
  const float* csgu_ln_w;  /* (U/2) */
  const float* csgu_ln_b;
  const float* csgu_dw_w;  /* (U/2,1,k) */
  const float* csgu_dw_b;
  smx_linear csgu_linear;  /* optional (use_linear_after_conv); w == NULL when absent */
  int32_t kernel_size;
  int32_t act;             /* activation after pre_channel_proj */
  int32_t gate_act;
  int32_t _pad;
} smx_convbranch_weights;

/* BranchformerEncoderLayer with attention_type == "SummaryMixing" (Branchformer.py:100-334). */
typedef struct {
  const float* norm_mhsa_w;
  const float* norm_mhsa_b;
  const float* norm_conv_w;
  const float* norm_conv_b;
  smx_cell_weights cell;
  smx_convbranch_weights branch;
  int32_t n_merge;                   /* blocks in merge_proj (Branchformer.py:220-226) */
  int32_t act;
  smx_linear merge[SMX_MAX_BLOCKS];
  const void* packed;                /* bf16 operand images from smx_branchformer_pack(), or NULL (generic arm) */
} smx_branchformer_layer_weights;

/* ---- library info -------------------------------------------------------------------------- */
SMX_API int smx_version(void);
SMX_API const char* smx_last_error(void);
/* sizeof() of the ABI structs as compiled (0 smx_linear, 1 smx_cell_weights, 2 smx_ffn_weights,
 * 3 smx_convmod_weights, 4 smx_conformer_layer_weights, 5 smx_convbranch_weights,
 * 6 smx_branchformer_layer_weights, 7 smx_cell_grads, 8 smx_ffn_grads, 9 smx_convmod_grads, 10 smx_convbranch_grads) so a
 * binding can verify its mirror of this header. */
SMX_API size_t smx_struct_size(int which);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
SMX_API uint64_t smx_launch_count(void);
/* of those, launches of tcgen05 (tensor-core arm) kernels */
SMX_API uint64_t smx_tc_launch_count(void);

/* ---- weight packing for the bf16 tensor-core (tcgen05) arm ----------------------------------------
 * With dtype SMX_BF16 a module runs on the tensor-core arm when its weights struct carries a `packed`
 * image; otherwise (or for configurations the arm does not handle: *_packed_bytes() == 0) the generic
 * fp32-math arm runs.  Pack once per weight set; `packed` must be 1024-byte aligned device memory. */
SMX_API size_t smx_cell_packed_bytes(const smx_cell_weights* w);
SMX_API int smx_cell_pack(const smx_cell_weights* w, void* packed, size_t packed_bytes, void* stream);
/* Optional, after smx_cell_pack() and with w->packed set: folds the LayerNorm that precedes the cell in a Conformer layer (norm1,
 * Conformer.py:520) into the image of the one-kernel cell -- gamma into the first blocks' weights, beta into their biases; the
 * kernel then feeds the raw rows to the tensor cores and applies the two per-row statistics in its first epilogue.  Sets
 * w->prenorm_w / w->prenorm_b (a no-op that leaves them NULL for configurations that kernel does not take).  smx_mixing_block_fwd /
 * smx_conformer_layer_fwd use the folded image when called with exactly these norm parameters, the plain one otherwise. */
SMX_API int smx_cell_pack_prenorm(smx_cell_weights* w, const float* norm_w, const float* norm_b, void* stream);
SMX_API size_t smx_ffn_packed_bytes(const smx_ffn_weights* w);
SMX_API int smx_ffn_pack(const smx_ffn_weights* w, void* packed, size_t packed_bytes, void* stream);
SMX_API size_t smx_branchformer_packed_bytes(const smx_branchformer_layer_weights* w); /* w->cell.packed must be set first */
SMX_API int smx_branchformer_pack(const smx_branchformer_layer_weights* w, void* packed, size_t packed_bytes, void* stream);
SMX_API size_t smx_convmod_packed_bytes(const smx_convmod_weights* w);
SMX_API int smx_convmod_pack(const smx_convmod_weights* w, void* packed, size_t packed_bytes, void* stream);

/* ---- primitives (each replaces one reference module call) ---------------------------------- */

/* nn.LayerNorm over the last dim (rows x D).  summary_mixing.py:218,249; Conformer.py:521,547,821. */
SMX_API int smx_layernorm_fwd(int dtype, int64_t rows, int32_t D, const void* x, const float* w, const float* b,
                      float eps, void* y, void* stream);

/* VanillaNN.forward: n_blocks x (linear, act) (VanillaNN.py:168-196).  x (rows,in) -> y (rows,out).
 * workspace: smx_vanilla_nn_workspace_bytes(). */
SMX_API size_t smx_vanilla_nn_workspace_bytes(const smx_linear* blocks, int32_t n_blocks, int dtype, int64_t rows);
SMX_API int smx_vanilla_nn_fwd(const smx_linear* blocks, int32_t n_blocks, int act, int dtype, int64_t rows,
                       const void* x, void* y, void* workspace, size_t workspace_bytes, void* stream);

/* SummaryMixing.forward (summary_mixing.py:169-324), eval mode.
 * y: (B,T,D_s); for SMX_MODE_LITE y is (B,D_s) — the reference returns a stride-0 expand over T
 * (summary_mixing.py:322) and the caller does the same.  `residual` (B,T,D_s) or NULL: when given,
 * y = cell(x) + residual (the `x + skip` of Conformer.py:541 fused; not for LITE). */
SMX_API size_t smx_summary_mixing_workspace_bytes(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, int has_sum_mask);
SMX_API int smx_summary_mixing_fwd(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x,
                           const uint8_t* padding_mask, const float* sum_mask, const void* residual, void* y,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Gradient of SummaryMixing.forward with respect to x and every parameter — what torch.autograd computes for
 * summary_mixing.py:198-253 (mode "SummaryMixing", whole-utterance mean: sum_mask == None) and :300-324
 * (mode "SummaryMixing-lite": y and dy are (B,D_s), only grads->summary is written) and :255-298 (mode
 * "SummaryMixing-fast": grads->global_proj and grads->merge), dropout off.
 * The call is self-contained: it recomputes the forward intermediates from x in fp32 (nothing is saved by
 * smx_summary_mixing_fwd), then back-propagates dy.  x, dy, dx are (B,T,*) in `dtype`; parameter gradients are fp32
 * in the parameters' own layouts (dense (out,in) / ParallelLinear (n_split,in/n_split,out/n_split)), OVERWRITTEN,
 * not accumulated; any gradient pointer may be NULL (not wanted).  dx may be NULL.
 * Mode -expdecay and sum_mask != None: SMX_ERR_UNSUPPORTED. */
typedef struct {
  float* dw;
  float* db;
} smx_linear_grad;
/* Training-mode dropout (torch.nn.Dropout semantics: an element is zeroed with probability p, the others scaled by 1/(1-p)).  The
 * masks are counter-based: Philox4x32-10 keyed by `seed`, counter = (element index / 4, site), word = element index % 4, an
 * element is dropped when its word < p * 2^32 -- a function of (seed, site, element index) only, so the backward call that
 * receives the same smx_dropout regenerates the forward's masks (nothing is stored).  Sites: FFN 0 = inside
 * PositionalwiseFeedForward, after the activation, over (rows, d_ffn); FFN 1 = the nn.Dropout after the block, over (rows, D)
 * (Conformer.py:470-484); cell 0 = on cat([local, summary]) over (B*T, D_l + D_s) (summary_mixing.py:252, :297); conv module
 * 0 = the nn.Dropout ending after_conv, over (B*T, D) (Conformer.py:163).  torch's own random stream is NOT reproduced. */
typedef struct {
  float p;       /* [0, 1); 0 = no dropout */
  uint64_t seed; /* drawn by the caller once per module call */
} smx_dropout;
typedef struct {
  smx_linear_grad local[SMX_MAX_BLOCKS];
  smx_linear_grad summary[SMX_MAX_BLOCKS];
  smx_linear_grad merge;
  float* local_norm_dw;
  float* local_norm_db;
  float* summary_norm_dw;
  float* summary_norm_db;
  smx_linear_grad global_proj; /* mode "SummaryMixing-fast" (summary_mixing.py:133-140) */
} smx_cell_grads;
SMX_API size_t smx_summary_mixing_bwd_workspace_bytes(const smx_cell_weights* w, int dtype, int32_t B, int32_t T);
SMX_API int smx_summary_mixing_bwd(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x,
                           const uint8_t* padding_mask, const void* dy, void* dx, const smx_cell_grads* grads,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Gradients of the other blocks of ConformerEncoderLayer.forward (Conformer.py:490-548), same conventions as
 * smx_summary_mixing_bwd: self-contained (forward intermediates recomputed from x in fp32), x / dy / dx in `dtype`,
 * parameter gradients fp32, overwritten, NULL = not wanted, dropout off.
 *   smx_layernorm_bwd    nn.LayerNorm (norm1, the encoder's final norm)
 *   smx_ffn_bwd          y = x + 0.5 * FFN(LN(x)), optionally followed by LN_out (Conformer.py:470-484, 518, 547)
 *   smx_conv_module_bwd  y = conv_module(x) * mask (Conformer.py:322-338; chunk_size == 0: no Dynamic Chunk Convolution) */
typedef struct {
  float* ln_dw;
  float* ln_db;
  smx_linear_grad w1;
  smx_linear_grad w2;
  float* out_ln_dw;
  float* out_ln_db;
} smx_ffn_grads;
typedef struct {
  float* ln_dw;
  float* ln_db;
  smx_linear_grad bottleneck;
  float* dw_dw;        /* depthwise weight gradient (D,1,k) */
  float* dw_db;
  float* after_ln_dw;
  float* after_ln_db;
  smx_linear_grad out;
} smx_convmod_grads;
/* VanillaNN backward (VanillaNN.py:168-196): grads[i] = gradients of block i, n_blocks entries. */
SMX_API size_t smx_vanilla_nn_bwd_workspace_bytes(const smx_linear* blocks, int32_t n_blocks, int dtype, int64_t rows);
SMX_API int smx_vanilla_nn_bwd(const smx_linear* blocks, int32_t n_blocks, int act, int dtype, int64_t rows, const void* x,
                       const void* dy, void* dx, const smx_linear_grad* grads, void* workspace, size_t workspace_bytes,
                       void* stream);
SMX_API size_t smx_layernorm_bwd_workspace_bytes(int dtype, int64_t rows, int32_t D);
SMX_API int smx_layernorm_bwd(int dtype, int64_t rows, int32_t D, const void* x, const float* w, float eps, const void* dy,
                      void* dx, float* dw, float* db, void* workspace, size_t workspace_bytes, void* stream);
SMX_API size_t smx_ffn_bwd_workspace_bytes(const smx_ffn_weights* w, int dtype, int64_t rows, int has_out_ln);
SMX_API int smx_ffn_bwd(const smx_ffn_weights* w, int act, int dtype, int64_t rows, const void* x, const float* out_ln_w,
                const float* out_ln_b, float out_ln_eps, const void* dy, void* dx, const smx_ffn_grads* grads,
                void* workspace, size_t workspace_bytes, void* stream);
SMX_API size_t smx_conv_module_bwd_workspace_bytes(const smx_convmod_weights* w, int dtype, int32_t B, int32_t T);
SMX_API int smx_conv_module_bwd(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, const void* x,
                        const uint8_t* padding_mask, const void* dy, void* dx, const smx_convmod_grads* grads,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Training-mode forward and backward WITH dropout (the smx_*_fwd entry points are the inference path: dropout is the identity
 * there).  Same contracts as smx_*_bwd: self-contained calls on the fp32-math arm (linears as split-bf16 tcgen05 GEMMs), the
 * *_train_bwd call recomputes the forward -- including the masks -- from x and the same smx_dropout.  drop == NULL or p == 0:
 * the same function as the inference forward / smx_*_bwd.  workspace: the matching smx_*_train_workspace_bytes() (covers both).
 * smx_ffn_train_fwd: y = x + 0.5 * drop1(W2 drop0(act(W1 LN(x) + b1)) + b2) [-> out LayerNorm];
 * smx_conv_module_train_fwd: y = drop0(conv_module(x)) * mask; smx_summary_mixing_train_fwd: modes "SummaryMixing" and
 * "SummaryMixing-fast", y = act(Wc drop0(cat[local, summary]) + bc) ("-lite" has no dropout: use the plain entry points). */
SMX_API size_t smx_ffn_train_workspace_bytes(const smx_ffn_weights* w, int dtype, int64_t rows, int has_out_ln);
SMX_API int smx_ffn_train_fwd(const smx_ffn_weights* w, int act, int dtype, int64_t rows, const void* x, const float* out_ln_w,
                const float* out_ln_b, float out_ln_eps, const smx_dropout* drop, void* y, void* workspace, size_t workspace_bytes,
                void* stream);
SMX_API int smx_ffn_train_bwd(const smx_ffn_weights* w, int act, int dtype, int64_t rows, const void* x, const float* out_ln_w,
                const float* out_ln_b, float out_ln_eps, const smx_dropout* drop, const void* dy, void* dx,
                const smx_ffn_grads* grads, void* workspace, size_t workspace_bytes, void* stream);
SMX_API size_t smx_conv_module_train_workspace_bytes(const smx_convmod_weights* w, int dtype, int32_t B, int32_t T);
SMX_API int smx_conv_module_train_fwd(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, const void* x,
                const uint8_t* padding_mask, const smx_dropout* drop, void* y, void* workspace, size_t workspace_bytes, void* stream);
SMX_API int smx_conv_module_train_bwd(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, const void* x,
                const uint8_t* padding_mask, const smx_dropout* drop, const void* dy, void* dx, const smx_convmod_grads* grads,
                void* workspace, size_t workspace_bytes, void* stream);
SMX_API size_t smx_summary_mixing_train_workspace_bytes(const smx_cell_weights* w, int dtype, int32_t B, int32_t T);
SMX_API int smx_summary_mixing_train_fwd(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x,
                const uint8_t* padding_mask, const smx_dropout* drop, void* y, void* workspace, size_t workspace_bytes, void* stream);
SMX_API int smx_summary_mixing_train_bwd(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x,
                const uint8_t* padding_mask, const smx_dropout* drop, const void* dy, void* dx, const smx_cell_grads* grads,
                void* workspace, size_t workspace_bytes, void* stream);
/* Dynamic Chunk Training (TransformerASR.py:85-110 builds the masks): the same training-mode calls with the chunk structure.
 * smx_summary_mixing_masked_train_*: sum_mask is the (T,T) fp32 src_mask (summary_mixing.py:235-246, :292-294; modes
 * "SummaryMixing" and "-fast"; NULL = the functions above; "-lite" ignores it); smx_conv_module_dcc_train_*: chunk_size > 0 selects
 * Dynamic Chunk Convolution (Conformer.py:197-320; 0 = the functions above).  drop may be NULL.  workspace:
 * smx_summary_mixing_masked_train_workspace_bytes() / smx_conv_module_train_workspace_bytes(). */
SMX_API size_t smx_summary_mixing_masked_train_workspace_bytes(const smx_cell_weights* w, int dtype, int32_t B, int32_t T);
SMX_API int smx_summary_mixing_masked_train_fwd(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x,
                const uint8_t* padding_mask, const float* sum_mask, const smx_dropout* drop, void* y, void* workspace,
                size_t workspace_bytes, void* stream);
SMX_API int smx_summary_mixing_masked_train_bwd(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, const void* x,
                const uint8_t* padding_mask, const float* sum_mask, const smx_dropout* drop, const void* dy, void* dx,
                const smx_cell_grads* grads, void* workspace, size_t workspace_bytes, void* stream);
SMX_API int smx_conv_module_dcc_train_fwd(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, int32_t chunk_size,
                const void* x, const uint8_t* padding_mask, const smx_dropout* drop, void* y, void* workspace, size_t workspace_bytes,
                void* stream);
SMX_API int smx_conv_module_dcc_train_bwd(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, int32_t chunk_size,
                const void* x, const uint8_t* padding_mask, const smx_dropout* drop, const void* dy, void* dx,
                const smx_convmod_grads* grads, void* workspace, size_t workspace_bytes, void* stream);
/* ConvolutionBranch (Branchformer.py:86-97: pre_channel_proj -> activation -> CSGU -> post_channel_proj; the CSGU is SpeechBrain's
 * ConvolutionalSpatialGatingUnit: split halves, LayerNorm + reflect-padded depthwise conv [+ linear] + gate activation on the second
 * half, product with the first, dropout).  x is the branch's input (the layer applies norm_conv before, Branchformer.py:292-293; no
 * padding mask, :276).  smx_conv_branch_train_fwd is the module's forward on the fp32-math arm (drop == NULL or p == 0: the
 * inference function; site 0 = the CSGU's dropout on the product); smx_conv_branch_train_bwd is self-contained like the other
 * smx_*_bwd calls (recomputes the forward from x, regenerates the mask).  workspace: smx_conv_branch_train_workspace_bytes(). */
typedef struct {
  smx_linear_grad pre;
  smx_linear_grad post;
  float* csgu_ln_dw;
  float* csgu_ln_db;
  float* csgu_dw_dw;   /* depthwise weight gradient (U/2,1,k) */
  float* csgu_dw_db;
  smx_linear_grad csgu_linear; /* used when the branch has use_linear_after_conv */
} smx_convbranch_grads;
SMX_API size_t smx_conv_branch_train_workspace_bytes(const smx_convbranch_weights* w, int dtype, int32_t B, int32_t T);
SMX_API int smx_conv_branch_train_fwd(const smx_convbranch_weights* w, int dtype, int32_t B, int32_t T, const void* x,
                const smx_dropout* drop, void* y, void* workspace, size_t workspace_bytes, void* stream);
SMX_API int smx_conv_branch_train_bwd(const smx_convbranch_weights* w, int dtype, int32_t B, int32_t T, const void* x,
                const smx_dropout* drop, const void* dy, void* dx, const smx_convbranch_grads* grads, void* workspace,
                size_t workspace_bytes, void* stream);
/* y = x * keep(site) / (1 - p): one nn.Dropout call of the layer surface (Branchformer.py:279, :294, :334) with the counter-based
 * mask; its own backward (dx = the same function of dy).  n elements; x and y may alias. */
SMX_API int smx_dropout_apply(const smx_dropout* drop, int32_t site, int dtype, int64_t n, const void* x, void* y, void* stream);
/* keep[e] = 1 when element e of `site` survives under `drop` (the mask the calls above apply); n elements, device pointer. */
SMX_API int smx_dropout_keep_mask(const smx_dropout* drop, int32_t site, int64_t n, uint8_t* keep, void* stream);

/* ConvolutionModule.forward (Conformer.py:166-340): y = conv_module(x) * mask (+ residual if given).
 * chunk_size > 0 selects Dynamic Chunk Convolution (Conformer.py:197-320). */
SMX_API size_t smx_conv_module_workspace_bytes(const smx_convmod_weights* w, int dtype, int32_t B, int32_t T);
SMX_API int smx_conv_module_fwd(const smx_convmod_weights* w, int act, int dtype, int32_t B, int32_t T, int32_t chunk_size,
                        const void* x, const uint8_t* padding_mask, const void* residual, void* y,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Macaron half-step FFN: y = x + 0.5 * FFN(LN(x)); if out_ln_w != NULL, y = LN_out(y)
 * (Conformer.py:518 and :547). */
SMX_API size_t smx_ffn_workspace_bytes(const smx_ffn_weights* w, int dtype, int64_t rows);
SMX_API int smx_ffn_fwd(const smx_ffn_weights* w, int act, int dtype, int64_t rows, const void* x,
                const float* out_ln_w, const float* out_ln_b, float out_ln_eps, void* y,
                void* workspace, size_t workspace_bytes, void* stream);

/* ---- encoder blocks ------------------------------------------------------------------------- */

/* The mixing block of ConformerEncoderLayer.forward: y = x + SummaryMixing(LayerNorm(x)) -- skip = x; x = norm1(x);
 * x = mha_layer(x, ...); x = x + skip (Conformer.py:520-541).  This is the unit the fused cell kernel executes inside a
 * layer (norm1 as its prologue, the skip as its epilogue) and the one bench.py reports the K-SM roofline on.
 * Requires summary_out_dim == enc_dim (the layer passes summary_out_dim=d_model, Conformer.py:446-457). */
SMX_API size_t smx_mixing_block_workspace_bytes(const smx_cell_weights* w, int dtype, int32_t B, int32_t T, int has_sum_mask);
SMX_API int smx_mixing_block_fwd(const smx_cell_weights* w, const float* norm_w, const float* norm_b, int dtype, int32_t B, int32_t T,
                         const void* x, const uint8_t* padding_mask, const float* sum_mask, void* y, void* workspace,
                         size_t workspace_bytes, void* stream);

/* The same block over n independent padded batches (same B, T; xs[i] -> ys[i]), enqueued back to back: the kernels chain by
 * programmatic dependent launch and the fused cell's per-utterance counters for all n calls are cleared by one memset ahead of
 * the chain (as smx_conformer_encoder_fwd does for its layers).  workspace: smx_mixing_block_workspace_bytes() +
 * n * 4096 + n * 8 * B bytes. */
SMX_API int smx_mixing_block_fwd_batch(const smx_cell_weights* w, const float* norm_w, const float* norm_b, int dtype, int32_t B,
                               int32_t T, int32_t n, const void* const* xs, const uint8_t* padding_mask, void* const* ys,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ConformerEncoderLayer.forward (Conformer.py:490-548). */
SMX_API size_t smx_conformer_layer_workspace_bytes(const smx_conformer_layer_weights* w, int dtype, int32_t B, int32_t T,
                                           int has_sum_mask);
SMX_API int smx_conformer_layer_fwd(const smx_conformer_layer_weights* w, int dtype, int32_t B, int32_t T, int32_t chunk_size,
                            const void* x, const uint8_t* padding_mask, const float* sum_mask, void* y,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ConformerEncoder.forward, eval mode (Conformer.py:765-827): n_layers layers then LayerNorm(eps=1e-6).
 * x and y may alias.  hidden (optional, may be NULL): n_layers pointers receiving each layer's output
 * (output_hidden_states, Conformer.py:819-825; the last one receives the normalised output). */
SMX_API size_t smx_conformer_encoder_workspace_bytes(const smx_conformer_layer_weights* layers, int32_t n_layers, int dtype,
                                             int32_t B, int32_t T, int has_sum_mask);
SMX_API int smx_conformer_encoder_fwd(const smx_conformer_layer_weights* layers, int32_t n_layers, const float* final_norm_w,
                              const float* final_norm_b, int dtype, int32_t B, int32_t T, int32_t chunk_size,
                              const void* x, const uint8_t* padding_mask, const float* sum_mask, void* y,
                              void* const* hidden, void* workspace, size_t workspace_bytes, void* stream);

/* BranchformerEncoderLayer.forward (Branchformer.py:243-334) and BranchformerEncoder.forward (:447-491). */
SMX_API size_t smx_branchformer_layer_workspace_bytes(const smx_branchformer_layer_weights* w, int dtype, int32_t B, int32_t T,
                                              int has_sum_mask);
SMX_API int smx_branchformer_layer_fwd(const smx_branchformer_layer_weights* w, int dtype, int32_t B, int32_t T, const void* x,
                               const uint8_t* padding_mask, const float* sum_mask, void* y, void* workspace,
                               size_t workspace_bytes, void* stream);
SMX_API size_t smx_branchformer_encoder_workspace_bytes(const smx_branchformer_layer_weights* layers, int32_t n_layers,
                                                int dtype, int32_t B, int32_t T, int has_sum_mask);
SMX_API int smx_branchformer_encoder_fwd(const smx_branchformer_layer_weights* layers, int32_t n_layers,
                                 const float* final_norm_w, const float* final_norm_b, int dtype, int32_t B, int32_t T,
                                 const void* x, const uint8_t* padding_mask, const float* sum_mask, void* y,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* ---- acoustic frontend (the blocks the recipes put in front of the encoder) --------------------------------------------------
 * The reference configures these by YAML name only (conformer_summarymixing.yaml:145-152, 198-200, 298-330); the classes are
 * SpeechBrain v1.0's (un-vendored): semantics restated from its published behaviour, "parity unpinned" (oracle/frontend_oracle.py).
 * All tensors fp32 device memory unless noted. */
typedef struct {
  int32_t sample_rate;   /* 16000 */
  int32_t n_fft;         /* 512 (the implemented size) */
  int32_t n_mels;        /* 80 */
  float win_length_ms;   /* 32 */
  float hop_length_ms;   /* 10 */
  float f_min;           /* 0 */
  float f_max;           /* <= 0: sample_rate / 2 */
  float amin;            /* 1e-10 */
  float top_db;          /* 80; <= 0: no clamp */
} smx_fbank_desc;
/* speechbrain.lobes.features.Fbank (conformer_summarymixing.yaml:326-330): wav (B, n_samples) -> feats (B, T', n_mels) with
 * T' = smx_fbank_frames() = 1 + n_samples / hop: hamming-window STFT (center, zero padding), power spectrum, triangular mel
 * filterbank, 10 log10, per-utterance top_db clamp. */
SMX_API int32_t smx_fbank_frames(const smx_fbank_desc* d, int32_t n_samples);
SMX_API size_t smx_fbank_workspace_bytes(const smx_fbank_desc* d, int32_t B);
SMX_API int smx_fbank_fwd(const smx_fbank_desc* d, int32_t B, int32_t n_samples, const float* wav, float* feats, void* workspace,
                  size_t workspace_bytes, void* stream);
/* InputNormalization(norm_type="global") at inference (yaml:198-200): y = (x - mean[f]) / std[f]; x, y (rows, F); may alias. */
SMX_API int smx_input_norm_fwd(int64_t rows, int32_t F, const float* x, const float* mean, const float* std, float* y, void* stream);
/* SpectrogramDrop (yaml:298-312), in place on x (B, T, F): positions [pos[b][m], pos[b][m] + len[b][m]) along dim (1 = time,
 * 2 = frequency) are replaced by the mean of the whole tensor (replace_mean != 0) or by zero.  pos / len: device int32 (B, n_masks),
 * drawn by the caller.  workspace: smx_spec_drop_workspace_bytes(). */
SMX_API size_t smx_spec_drop_workspace_bytes(void);
SMX_API int smx_spec_drop_fwd(int32_t B, int32_t T, int32_t F, float* x, int32_t dim, int32_t n_masks, const int32_t* pos, const int32_t* len,
                      int32_t replace_mean, void* workspace, size_t workspace_bytes, void* stream);
/* Warping (yaml:315; dim = 1, bicubic, align_corners): frames [0, c) of x (B, T, F) are resampled to w frames, frames [c, T) to
 * T - w frames; y must not alias x.  c, w in (0, T), drawn by the caller. */
SMX_API int smx_time_warp_fwd(int32_t B, int32_t T, int32_t F, const float* x, int32_t c, int32_t w, float* y, void* stream);
/* One block of ConvolutionFrontEnd (yaml:145-152; num_layers_per_block = 1, no residual): Conv2d(kernel, stride, reflect "same")
 * over (time, frequency) of x (B, T, F, Cin) channels-last -> LayerNorm over (F', Cout) -> LeakyReLU(0.01); y (B, T', F', Cout),
 * T' = ceil(T / stride), F' = ceil(F / stride).  conv_w (Cout, Cin, k, k), conv_b (Cout), ln_w / ln_b (F', Cout). */
SMX_API int smx_conv_frontend_block_fwd(int32_t B, int32_t T, int32_t F, int32_t Cin, int32_t Cout, int32_t kernel, int32_t stride,
                                const float* x, const float* conv_w, const float* conv_b, const float* ln_w, const float* ln_b,
                                float* y, void* stream);
/* custom_src_module + positional encoding (TransformerASR.py:353-358, 405-406; Transformer.py:288-339), dropout off:
 * y (B, T, D) in out_dtype = x (B, T, in_dim) W^T + b + pe[t, :].  T must not exceed max_len (the reference's table, 2500). */
SMX_API size_t smx_input_proj_workspace_bytes(int32_t B, int32_t T, int32_t D);
SMX_API int smx_input_proj_fwd(const smx_linear* proj, int32_t B, int32_t T, int32_t max_len, const float* x, int out_dtype, void* y,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- mask builders (TransformerASR.py:50-180) ------------------------------------------------ */

/* padding mask from relative lengths: abs_len = round(wav_len*T); mask[b,t] = t < abs_len[b]
 * (TransformerASR.py:157-162, masked_false_or_true == False). */
SMX_API int smx_padding_mask_from_wav_len(const float* wav_len, int32_t B, int32_t T, uint8_t* mask, void* stream);
/* dynamic-chunk sum mask (T,T) fp32: 1 where frame j is visible to frame i (TransformerASR.py:85-110,
 * masked_false_or_true == False).  left_context_chunks < 0 = infinite left context. */
SMX_API int smx_chunk_mask(int32_t T, int32_t chunk_size, int32_t left_context_chunks, float* mask, void* stream);

/* ---- diagnostics ---------------------------------------------------------------------------------- */

/* UMMA self-test: c (M,N) fp32 = a (M,K) bf16 @ w (N,K) fp32->bf16 transposed, computed by the tcgen05
 * building blocks the fused kernels use (layout 0: 128B-swizzled operands, 1: unswizzled).
 * N <= 256, K % 8 == 0.  workspace: N*K*2 bytes rounded up to tiles (1 MiB is always enough). */
SMX_API int smx_debug_tc_gemm(int layout, int32_t M, int32_t N, int32_t K, const void* a_bf16, const float* w_f32,
                              float* c_f32, void* workspace, size_t workspace_bytes, void* stream);

/* Timeline of CTA 0 of the fused persistent kernels: device buffer of >= 1024 uint64 (zeroed by the caller)
 * receiving clock64() stamps per warp role / tile / event; NULL switches tracing off.  Diagnostics only. */
SMX_API int smx_debug_set_trace(void* device_u64_buffer);
/* Thread-block cluster size (1, 2 or 4) of the persistent FFN kernel: CTAs of a cluster multicast weight blocks
 * to each other.  Tuning / diagnostics. */
SMX_API int smx_debug_set_ffn_cluster(int cluster_size);
/* Fused FFN kernel generation: 4 (default where D is a multiple of 128 and there are at least two row tiles: hidden activation
 * resident in tensor memory, CTA pairs with cta_group::2 MMAs, each CTA streams half of every weight step, final epilogue staged
 * through shared memory), 3 (the same on single CTAs; bit-identical results; what other shapes run) or 2 (hidden activation
 * staged through shared memory).  All compute the same function; diagnostics / A-B timing. */
SMX_API int smx_debug_set_ffn_version(int version);
/* Programmatic dependent launch of the fused kernels (default on): a kernel's set-up overlaps the tail of its
 * predecessor; results are identical.  Diagnostics / A-B timing. */
SMX_API int smx_debug_set_pdl(int on);
/* Fused SummaryMixing cell / GLU pass generation: 3 (default; hidden activations and the normalised local branch stay in
 * tensor memory as MMA A operands, step-granular weight ring, 16 epilogue warps) or 1 (first generation, operands staged
 * through shared memory).  Same function; diagnostics / A-B timing. */
SMX_API int smx_debug_set_cell_version(int version);

/* fp32 (SMX_F32) arm: linears with >= 128 rows and dims that are multiples of 64 run on the tensor cores with split-bf16
 * operands (x = hi + lo in bf16; hi*hi + hi*lo + lo*hi accumulated in fp32: ~1e-5 relative to the fp32 product); 0 keeps every
 * product on the CUDA-core fp32 GEMM.  Default on.  Diagnostics / A-B. */
SMX_API int smx_debug_set_f32_tc(int on);

#ifdef __cplusplus
}
#endif
#endif /* SMX_H_ */
