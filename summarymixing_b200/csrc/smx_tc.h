// Declarations of the tcgen05 (bf16 tensor-core) arm of libsmx.
#pragma once
#include "smx_internal.h"

namespace smx {

// bytes of workspace the tcgen05 cell path needs for this problem, or 0 when it does not apply
size_t tc_cell_workspace_bytes(const smx_cell_weights* w, int dtype, int B, int T, int has_sum_mask);

// Pack an fp32 weight (element (n,k) at w[n*stride_n + k*stride_k]) into bf16 operand images:
// layout 0: [n_tile][k_block] tiles of (NT rows x 64 cols) in the 128B-swizzled K-major layout
// layout 1: [n_tile] tiles of (NT rows x Kpad cols) in the no-swizzle core-matrix layout
// Rows >= N and cols >= K are zero.  Returns bytes written via *bytes.
size_t tc_packed_bytes(int N, int K, int NT);
int tc_pack_weight(const float* w, int64_t stride_n, int64_t stride_k, int N, int K, int NT, int layout,
                   __nv_bfloat16* out, cudaStream_t st);

}  // namespace smx
