// libsmx, backward of the SummaryMixing cell (fp32-math arm): what torch.autograd derives from
// summary_mixing.py:198-253 (mode "SummaryMixing", whole-utterance mean), written out by hand.
//
//   forward (recomputed here from x, fp32):
//     f-branch  a_0 = x;  z_i = lin_i(a_i);  a_{i+1} = act(z_i);  Lm = a_n * mask;  L = LN_l(Lm)            :215-218
//     s-branch  same blocks of summary_proj;  Sm = a_n * mask;  mean_b = sum_t Sm / sum_t mask;  mu = LN_s(mean)  :221-249
//     combiner  zc = L Wc[:, :D_l]^T + (mu_b Wc[:, D_l:]^T + b_c);  y = act(zc)                                  :251-253
//   backward:
//     dzc = dy * act'(zc);  dWc = [dzc^T L | dcb^T mu] with dcb_b = sum_t dzc[b,t];  db_c = sum dzc
//     dL = dzc Wc[:, :D_l];  dmu = dcb Wc[:, D_l:];  LayerNorm backward on both;  dSm[b,t] = dmean_b / count_b
//     MLP backward per branch: dz_i = da_{i+1} * act'(z_i) (* mask on the last block), dW_i = a_i^T dz_i (split-K over
//     row slices, fixed-order reduction: deterministic), db_i = sum dz_i, da_i = dz_i W_i;  dx = da_0(f) + da_0(s)
//
// First-correct implementation on the generic kernels (strided fp32 GEMM of smx_simt.cu + the elementwise / reduction
// kernels below); not tuned.  Parity: tests/test_backward_gpu.py against torch.autograd of the oracle restatement.
#include "smx_internal.h"
#include "smx_tc.h"
#include <math.h>

namespace smx {

namespace {

__device__ __forceinline__ float bw_ld(const void* p, int dt, int64_t i) {
  return dt == SMX_BF16 ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : ((const float*)p)[i];
}

__device__ __forceinline__ float bw_act(int act, float x) {
  switch (act) {
    case SMX_ACT_SWISH: return x / (1.0f + expf(-x));
    case SMX_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
    case SMX_ACT_RELU: return fmaxf(x, 0.0f);
    case SMX_ACT_LEAKY_RELU: return x >= 0.0f ? x : 0.01f * x;
    case SMX_ACT_TANH: return tanhf(x);
    case SMX_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case SMX_ACT_GELU_TANH: {
      float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      return 0.5f * x * (1.0f + tanhf(u));
    }
    default: return x;
  }
}

// d act(x) / dx
__device__ __forceinline__ float bw_dact(int act, float x) {
  switch (act) {
    case SMX_ACT_SWISH: {
      const float s = 1.0f / (1.0f + expf(-x));
      return s * (1.0f + x * (1.0f - s));
    }
    case SMX_ACT_GELU: {
      const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
      const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
      return cdf + x * pdf;
    }
    case SMX_ACT_RELU: return x > 0.0f ? 1.0f : 0.0f;
    case SMX_ACT_LEAKY_RELU: return x > 0.0f ? 1.0f : 0.01f;
    case SMX_ACT_TANH: {
      const float t = tanhf(x);
      return 1.0f - t * t;
    }
    case SMX_ACT_SIGMOID: {
      const float s = 1.0f / (1.0f + expf(-x));
      return s * (1.0f - s);
    }
    case SMX_ACT_GELU_TANH: {
      const float x2 = x * x;
      const float u = 0.7978845608028654f * (x + 0.044715f * x * x2);
      const float t = tanhf(u);
      const float du = 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x2);
      return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
    }
    default: return 1.0f;
  }
}

// ---- training-mode dropout: counter-based masks (Philox4x32-10), a function of (seed, site, element index) only, so that the
// backward regenerates the forward's mask.  Element e of site s: counter (e / 4 [64 bit], s, 0), key = seed, word e % 4;
// dropped when word < p * 2^32, kept values scaled by 1 / (1 - p) (torch.nn.Dropout's semantics; its random stream is not matched).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ uint4 drop_words(uint64_t seed, uint32_t site, int64_t quad) {
  return philox4x32_10(make_uint4((uint32_t)quad, (uint32_t)((uint64_t)quad >> 32), site, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// buf[e] = keep(e) ? buf[e] * scale : 0, in place (fp32, contiguous; one Philox call per four elements)
__global__ void __launch_bounds__(256) dropout_kernel(float* buf, int64_t n, uint64_t seed, uint32_t site, uint32_t thresh, float scale) {
  const int64_t quad = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e0 = quad * 4;
  if (e0 >= n) return;
  const uint4 r = drop_words(seed, site, quad);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (e0 + i < n) buf[e0 + i] = w[i] >= thresh ? buf[e0 + i] * scale : 0.0f;
}
__global__ void __launch_bounds__(256) dropout_mask_kernel(uint8_t* keep, int64_t n, uint64_t seed, uint32_t site, uint32_t thresh) {
  const int64_t quad = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e0 = quad * 4;
  if (e0 >= n) return;
  const uint4 r = drop_words(seed, site, quad);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (e0 + i < n) keep[e0 + i] = w[i] >= thresh ? 1 : 0;
}
// cat[row] = [ L[row, 0:Dl) | mu[row / T, 0:Ds) ]: the combiner's input written out (needed only when dropout acts on it)
__global__ void __launch_bounds__(256) concat_bcast_kernel(const float* L, int64_t ldl, const float* mu, int T, int Dl, int Ds, int64_t n, float* cat) {
  const int W = Dl + Ds;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t row = i / W;
    const int c = (int)(i - row * W);
    cat[i] = c < Dl ? L[row * ldl + c] : mu[(row / T) * Ds + (c - Dl)];
  }
}
__global__ void __launch_bounds__(256) take_cols_kernel(const float* src, int64_t lds, int ncols, int64_t n, float* dst, int64_t ldd) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t row = i / ncols;
    const int c = (int)(i - row * ncols);
    dst[row * ldd + c] = src[row * lds + c];
  }
}
// y = x + alpha * u
__global__ void __launch_bounds__(256) axpy_kernel(const float* x, const float* u, float alpha, int64_t n, float* y) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) y[i] = fmaf(alpha, u[i], x[i]);
}

// out = act(z) (* rowmask[row])
__global__ void __launch_bounds__(256) act_fwd_kernel(const float* z, int64_t n, int ncols, int act, const uint8_t* rowmask,
                                                      float* out) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    float v = bw_act(act, z[i]);
    if (rowmask) v *= (float)rowmask[i / ncols];
    out[i] = v;
  }
}

// dz = da * act'(z) (* rowmask[row]);  dz may alias z or da
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* z, const void* da, int da_dt, int64_t n, int ncols, int act,
                                                      const uint8_t* rowmask, float* dz) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    float v = bw_ld(da, da_dt, i) * bw_dact(act, z[i]);
    if (rowmask) v *= (float)rowmask[i / ncols];
    dz[i] = v;
  }
}

// dst[slice*N + n] = sum over the slice's rows of src[row*ld + n]; fixed order
__global__ void __launch_bounds__(256) colsum_kernel(const float* src, int64_t ld, int64_t rows, int rows_per_slice, int N,
                                                     float* dst) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x % 32, ry = threadIdx.x / 32;
  const int n = blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_slice;
  const int64_t r1 = r0 + rows_per_slice < rows ? r0 + rows_per_slice : rows;
  float acc = 0.0f;
  if (n < N)
    for (int64_t r = r0 + ry; r < r1; r += 8) acc += src[r * ld + n];
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && n < N) {
    float tot = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r) tot += red[r][cx];
    dst[(int64_t)blockIdx.y * N + n] = tot;
  }
}

// dst[r*ldd + c] = sum_s P[s][r][c]; fixed order
__global__ void __launch_bounds__(256) sum_slices_kernel(const float* P, int ns, int nrows, int ncols, float* dst, int64_t ldd) {
  const int64_t n = (int64_t)nrows * ncols;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    float tot = 0.0f;
    for (int s = 0; s < ns; ++s) tot += P[(int64_t)s * n + i];
    dst[(i / ncols) * ldd + (i % ncols)] = tot;
  }
}

// LayerNorm backward, one warp per row.  v: the LayerNorm input; g (in): dy, (out): dv;  t (out): dy * xhat (its column
// sums are the weight gradient; the column sums of dy, taken BEFORE this kernel, are the bias gradient).
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* v, int64_t rows, int D, const float* w, float eps, float* g,
                                                     float* t, int64_t ldv, int64_t ldg) {
  const int64_t row = (int64_t)blockIdx.x * 8 + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= rows) return;
  const float* vr = v + row * ldv;
  float* gr = g + row * ldg;
  float* tr = t + row * D;
  float s = 0.0f;
  for (int c = lane; c < D; c += 32) s += vr[c];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)D;
  float q = 0.0f;
  for (int c = lane; c < D; c += 32) { const float d = vr[c] - mean; q += d * d; }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.0f / sqrtf(q / (float)D + eps);
  float m1 = 0.0f, m2 = 0.0f;
  for (int c = lane; c < D; c += 32) {
    const float gw = gr[c] * w[c];
    m1 += gw;
    m2 += gw * (vr[c] - mean) * rstd;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o); }
  m1 /= (float)D; m2 /= (float)D;
  for (int c = lane; c < D; c += 32) {
    const float xh = (vr[c] - mean) * rstd;
    const float dy = gr[c];
    tr[c] = dy * xh;
    gr[c] = rstd * (dy * w[c] - m1 - xh * m2);
  }
}

// inv[b] = 1 / (number of valid frames of utterance b)            summary_mixing.py:229-231
__global__ void inv_count_kernel(const uint8_t* mask, int T, float* inv) {
  const int b = blockIdx.x;
  float c = 0.0f;
  for (int t = threadIdx.x; t < T; t += 32) c += mask ? (float)mask[(int64_t)b * T + t] : 1.0f;
#pragma unroll
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (threadIdx.x == 0) inv[b] = 1.0f / c;
}

// dst[(b*T + t)*D + d] = src[b*D + d] * inv[b]   (gradient of the mean, broadcast back over the frames)
__global__ void __launch_bounds__(256) bcast_scale_kernel(const float* src, const float* inv, int T, int D, int64_t n, float* dst,
                                                          int64_t ldd = 0) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t row = i / D;
    const int b = (int)(row / T), d = (int)(i % D);
    dst[ldd ? row * ldd + d : i] = src[(int64_t)b * D + d] * inv[b];
  }
}

unsigned ew_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

struct Drop { bool on; uint64_t seed; uint32_t thresh; float scale; };
Drop make_drop(const smx_dropout* d) {
  Drop r{false, 0, 0, 1.0f};
  if (d && d->p > 0.0f) {
    const double t = (double)d->p * 4294967296.0;
    r.on = true; r.seed = d->seed; r.thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t; r.scale = 1.0f / (1.0f - d->p);
  }
  return r;
}
int dropout_inplace(float* buf, int64_t n, const Drop& d, uint32_t site, cudaStream_t st) {
  if (!d.on || n <= 0) return SMX_OK;
  const int64_t quads = (n + 3) / 4;
  dropout_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, st>>>(buf, n, d.seed, site, d.thresh, d.scale);
  count_launch();
  return check_launch("dropout_kernel");
}

// Tensor-core form of the three linear primitives below (recompute, data gradient, weight gradient): split-bf16 operands,
// fp32 accumulation (smx_tc_gemm.cu), ~1e-5 relative to the fp32 products -- the golden-gradient tolerances hold.  The scratch
// for the two per-row primitives is reserved once per backward entry point (BwScratch) so that the sizing runs see it.
thread_local void* t_bw_sc = nullptr;
thread_local size_t t_bw_sc_bytes = 0;
struct BwScratch {
  BwScratch(Arena& ws, int64_t rows, int maxdim) {
    t_bw_sc = nullptr; t_bw_sc_bytes = 0;
    if (!tc_f32_tc_enabled() || !tc_split3_ok(rows, 64, 64)) return;
    const int md = (maxdim + 63) / 64 * 64;
    t_bw_sc_bytes = tc_split3_scratch_bytes(rows, md, md);
    t_bw_sc = ws.take(t_bw_sc_bytes);
    if (!t_bw_sc) t_bw_sc_bytes = 0;
  }
  ~BwScratch() { t_bw_sc = nullptr; t_bw_sc_bytes = 0; }
};
bool bw_tc_ok(int64_t rows, int K, int N, const void* a, int64_t lda, const void* c, int64_t ldc, bool c_f32) {
  return t_bw_sc && tc_f32_tc_enabled() && tc_split3_ok(rows, K, N) && t_bw_sc_bytes >= tc_split3_scratch_bytes(rows, K, N) &&
         lda % 4 == 0 && ((uintptr_t)a % 16) == 0 && ((uintptr_t)c % 32) == 0 && ldc % (c_f32 ? 8 : 16) == 0;
}

GemmP bw_gemm() {
  GemmP p{};
  p.alpha = 1.0f;
  p.rowbias_div = 1;
  p.batches = 1;
  p.act = SMX_ACT_IDENTITY;
  p.a_dtype = SMX_F32;
  p.c_dtype = SMX_F32;
  return p;
}

// z = A @ L (+ b): pre-activation of one block (dense columns [k_offset, k_offset + K) when K > 0)
int lin_fwd(const smx_linear& L, const float* A, int64_t lda, int64_t rows, float* C, int64_t ldc, bool use_bias, int k_offset,
            int K, const float* rowbias, int rowbias_div, cudaStream_t st) {
  {
    const bool dense = L.n_split <= 1;
    const int Kt = dense ? (K > 0 ? K : L.in_dim - k_offset) : L.in_dim;
    if ((dense || (L.in_dim % L.n_split == 0 && L.out_dim % L.n_split == 0)) && bw_tc_ok(rows, Kt, L.out_dim, A, lda, C, ldc, true)) {
      GemmTc g{};
      g.bias = (use_bias && L.b) ? L.b : nullptr;
      g.rowbias = rowbias; g.rowbias_ld = L.out_dim; g.rows_per_group = rowbias_div > 0 ? rowbias_div : 1;
      g.act = SMX_ACT_IDENTITY; g.alpha = 1.0f; g.out_f32 = C; g.ldo = ldc;
      return tc_linear_split3(L, dense ? k_offset : 0, Kt, A, lda, rows, g, t_bw_sc, st);
    }
  }
  GemmP p = bw_gemm();
  p.A = A; p.lda = lda; p.C = C; p.ldc = ldc; p.M = (int)rows;
  p.rowbias = rowbias; p.rowbias_ld = L.out_dim; p.rowbias_div = rowbias_div;
  if (L.n_split <= 1) {
    p.K = K > 0 ? K : L.in_dim - k_offset; p.N = L.out_dim;
    p.W = L.w + k_offset; p.w_sk = 1; p.w_sn = L.in_dim;
    p.bias = (use_bias && L.b) ? L.b : nullptr;
  } else {
    const int h = L.n_split;
    p.K = L.in_dim / h; p.N = L.out_dim / h; p.batches = h;
    p.a_bs = p.K; p.c_bs = p.N;
    p.W = L.w; p.w_sk = p.N; p.w_sn = 1; p.w_bs = (int64_t)p.K * p.N;
    p.bias = (use_bias && L.b) ? L.b : nullptr; p.bias_bs = p.N;
  }
  return gemm(p, st);
}

// dX = dZ @ W (+ residual), the gradient with respect to the block's input (dense: columns [k_offset, k_offset + K))
int lin_dgrad(const smx_linear& L, const float* dZ, int64_t ldz, int64_t rows, void* dX, int dx_dt, int64_t ldx, int k_offset, int K,
              const float* residual, cudaStream_t st) {
  {
    const bool dense = L.n_split <= 1;
    const int Kin = dense ? (K > 0 ? K : L.in_dim - k_offset) : L.in_dim;
    if ((dense || (L.in_dim % L.n_split == 0 && L.out_dim % L.n_split == 0)) &&
        bw_tc_ok(rows, L.out_dim, Kin, dZ, ldz, dX, ldx, dx_dt == SMX_F32) && (!residual || (((uintptr_t)residual % 16) == 0 && ldx % 4 == 0))) {
      GemmTc g{};
      g.act = SMX_ACT_IDENTITY; g.alpha = 1.0f; g.resid_f32 = residual; g.ldr = ldx; g.ldo = ldx;
      if (dx_dt == SMX_F32) g.out_f32 = (float*)dX; else g.out = (__nv_bfloat16*)dX;
      return tc_dgrad_split3(L, dense ? k_offset : 0, Kin, dZ, ldz, rows, g, t_bw_sc, st);
    }
  }
  GemmP p = bw_gemm();
  p.A = dZ; p.lda = ldz; p.C = dX; p.c_dtype = dx_dt; p.ldc = ldx; p.M = (int)rows;
  p.residual = residual; p.r_dtype = SMX_F32; p.ldr = ldx;
  if (L.n_split <= 1) {
    p.K = L.out_dim; p.N = K > 0 ? K : L.in_dim - k_offset;
    p.W = L.w + k_offset; p.w_sk = L.in_dim; p.w_sn = 1;
  } else {
    const int h = L.n_split, a = L.in_dim / h, b = L.out_dim / h;
    p.K = b; p.N = a; p.batches = h;
    p.a_bs = b; p.c_bs = a;
    p.W = L.w; p.w_sk = 1; p.w_sn = b; p.w_bs = (int64_t)a * b;
  }
  return gemm(p, st);
}

// dW[m][i][o] = T[m a + i][m b + o]: the per-head blocks of the dense (in, out) product, in ParallelLinear's layout
__global__ void gather_heads_kernel(const float* __restrict__ Tm, int ldt, int h, int a, int b, float* __restrict__ dW) {
  const int64_t n = (int64_t)h * a * b;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % b), ii = (int)((i / b) % a), m = (int)(i / ((int64_t)a * b));
    dW[i] = Tm[(int64_t)(m * a + ii) * ldt + m * b + o];
  }
}

int sum_slices(const float* P, int ns, int nrows, int ncols, float* dst, int64_t ldd, cudaStream_t st) {
  sum_slices_kernel<<<ew_grid((int64_t)nrows * ncols), 256, 0, st>>>(P, ns, nrows, ncols, dst, ldd);
  count_launch();
  return check_launch("sum_slices_kernel");
}

void slice_plan(int64_t rows, int& chunk, int& ns) {
  int64_t want = (rows + 255) / 256;
  if (want > 64) want = 64;
  if (want < 1) want = 1;
  chunk = (int)((rows + want - 1) / want);
  ns = (int)((rows + chunk - 1) / chunk);
}

// dW = X^T dZ in the parameter's own layout; split-K over row slices, then a fixed-order reduction.
int lin_wgrad(const smx_linear& L, const float* dZ, int64_t ldz, const float* X, int64_t ldx, int64_t rows, float* dW, int k_offset,
              int K, Arena& ws, cudaStream_t st) {
  if (tc_f32_tc_enabled() && rows >= 128) {
    // tensor-core form: partial products over row slices by the split-bf16 GEMM (contraction over the rows: both operands are
    // transposed activations), then the same fixed-order reduction as below
    const bool dense = L.n_split <= 1;
    const int Kin = dense ? (K > 0 ? K : L.in_dim - k_offset) : L.in_dim;
    const int M = dense ? L.out_dim : Kin, N = dense ? Kin : L.out_dim;   // dense: dW (out, in); split: product as (in, out), gathered per head
    if (M % 64 == 0 && N % 64 == 0 && (dense || (L.in_dim % L.n_split == 0 && L.out_dim % L.n_split == 0))) {
      const size_t m1 = ws.mark();
      const int nsl = tc_wgrad_slices(rows);
      float* P = ws.f32((size_t)nsl * M * N);
      float* Tm = dense ? nullptr : ws.f32((size_t)M * N);
      void* sc = ws.take(tc_wgrad_scratch_bytes(rows, M, N));
      if (!P || !sc || (!dense && !Tm)) return fail(SMX_ERR_WORKSPACE, "workspace too small (weight gradient, tensor-core form)");
      if (!ws.dry) {
        int ns2 = 0;
        if (dense) {
          SMX_TRY(tc_wgrad_split3(dZ, ldz, M, X, ldx, N, rows, P, &ns2, sc, st));
          SMX_TRY(sum_slices(P, ns2, M, N, dW + k_offset, L.in_dim, st));
        } else {
          SMX_TRY(tc_wgrad_split3(X, ldx, M, dZ, ldz, N, rows, P, &ns2, sc, st));
          SMX_TRY(sum_slices(P, ns2, M, N, Tm, N, st));
          const int h = L.n_split, a = L.in_dim / h, b = L.out_dim / h;
          gather_heads_kernel<<<ew_grid((int64_t)h * a * b), 256, 0, st>>>(Tm, N, h, a, b, dW);
          count_launch();
          SMX_TRY(check_launch("gather_heads_kernel"));
        }
      }
      ws.release(m1);
      return SMX_OK;
    }
  }
  int chunk, ns;
  slice_plan(rows, chunk, ns);
  const size_t m0 = ws.mark();
  if (L.n_split <= 1) {
    const int Kin = K > 0 ? K : L.in_dim - k_offset;
    float* P = ws.f32((size_t)ns * L.out_dim * Kin);
    if (!P) return fail(SMX_ERR_WORKSPACE, "workspace too small (weight gradient)");
    if (!ws.dry) {
      GemmP p = bw_gemm();  // C[o][i] = sum_m dZ[m,o] X[m,i]
      p.A = dZ; p.lda = 1; p.a_sk = ldz; p.a_bs = (int64_t)chunk * ldz;
      p.W = X; p.w_sk = ldx; p.w_sn = 1; p.w_bs = (int64_t)chunk * ldx;
      p.C = P; p.ldc = Kin; p.c_bs = (int64_t)L.out_dim * Kin;
      p.M = L.out_dim; p.N = Kin; p.K = chunk; p.k_total = (int)rows; p.batches = ns;
      SMX_TRY(gemm(p, st));
      SMX_TRY(sum_slices(P, ns, L.out_dim, Kin, dW + k_offset, L.in_dim, st));
    }
  } else {
    const int h = L.n_split, a = L.in_dim / h, b = L.out_dim / h;
    float* P = ws.f32((size_t)ns * h * a * b);  // [slice][head][a][b]
    if (!P) return fail(SMX_ERR_WORKSPACE, "workspace too small (weight gradient)");
    if (!ws.dry) {  // C[head][i][o] = sum_m X[m, head*a + i] dZ[m, head*b + o]: one launch over slices x heads
      GemmP p = bw_gemm();
      p.A = X; p.lda = 1; p.a_sk = ldx; p.a_bs = (int64_t)chunk * ldx; p.a_bs2 = a;
      p.W = dZ; p.w_sk = ldz; p.w_sn = 1; p.w_bs = (int64_t)chunk * ldz; p.w_bs2 = b;
      p.C = P; p.ldc = b; p.c_bs = (int64_t)h * a * b; p.c_bs2 = (int64_t)a * b;
      p.M = a; p.N = b; p.K = chunk; p.k_total = (int)rows; p.batches = ns; p.batches2 = h;
      SMX_TRY(gemm(p, st));
      SMX_TRY(sum_slices(P, ns, h * a, b, dW, b, st));
    }
  }
  ws.release(m0);
  return SMX_OK;
}

// dst[n] = sum over all rows of src[row*ld + n]
int colsum_all(const float* src, int64_t ld, int64_t rows, int N, float* dst, Arena& ws, cudaStream_t st) {
  const int rps = 512;
  const int ns = (int)((rows + rps - 1) / rps);
  const size_t m0 = ws.mark();
  float* P = ws.f32((size_t)ns * N);
  if (!P) return fail(SMX_ERR_WORKSPACE, "workspace too small (column sums)");
  if (!ws.dry) {
    colsum_kernel<<<dim3((N + 31) / 32, ns), 256, 0, st>>>(src, ld, rows, rps, N, P);
    count_launch();
    SMX_TRY(check_launch("colsum_kernel"));
    SMX_TRY(sum_slices(P, ns, 1, N, dst, N, st));
  }
  ws.release(m0);
  return SMX_OK;
}

int act_fwd(const float* z, int64_t rows, int ncols, int act, const uint8_t* rowmask, float* out, cudaStream_t st) {
  act_fwd_kernel<<<ew_grid(rows * ncols), 256, 0, st>>>(z, rows * ncols, ncols, act, rowmask, out);
  count_launch();
  return check_launch("act_fwd_kernel");
}
int act_bwd(const float* z, const void* da, int da_dt, int64_t rows, int ncols, int act, const uint8_t* rowmask, float* dz,
            cudaStream_t st) {
  act_bwd_kernel<<<ew_grid(rows * ncols), 256, 0, st>>>(z, da, da_dt, rows * ncols, ncols, act, rowmask, dz);
  count_launch();
  return check_launch("act_bwd_kernel");
}

// LayerNorm backward over `rows` rows: g (dy -> dv, in place), parameter gradients into dw / db (either may be NULL)
int ln_bwd(const float* v, int64_t rows, int D, const float* w, float* g, float* dw, float* db, Arena& ws, cudaStream_t st,
           float eps = 1e-5f, int64_t ldv = 0, int64_t ldg = 0) {
  if (ldv == 0) ldv = D;
  if (ldg == 0) ldg = D;
  const size_t m0 = ws.mark();
  float* t = ws.f32((size_t)rows * D);
  if (!t) return fail(SMX_ERR_WORKSPACE, "workspace too small (LayerNorm backward)");
  if (db) SMX_TRY(colsum_all(g, ldg, rows, D, db, ws, st));  // before g is overwritten
  if (!ws.dry) {
    ln_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(v, rows, D, w, eps, g, t, ldv, ldg);
    count_launch();
    SMX_TRY(check_launch("ln_bwd_kernel"));
  }
  if (dw) SMX_TRY(colsum_all(t, D, rows, D, dw, ws, st));
  ws.release(m0);
  return SMX_OK;
}

struct BranchFwd {
  float* z[SMX_MAX_BLOCKS];      // pre-activations
  const float* a[SMX_MAX_BLOCKS + 1];  // block inputs; a[n] = act(z[n-1]) * mask
};

int branch_fwd(const smx_linear* blocks, int n, int act, const float* x32, int64_t rows, const uint8_t* mask, BranchFwd& f, Arena& ws,
               cudaStream_t st) {
  f.a[0] = x32;
  for (int i = 0; i < n; ++i) {
    if (i > 0 && blocks[i].in_dim != blocks[i - 1].out_dim) return fail(SMX_ERR_BAD_ARG, "VanillaNN block %d: inconsistent dims", i);
    const int N = blocks[i].out_dim;
    f.z[i] = ws.f32((size_t)rows * N);
    float* an = ws.f32((size_t)rows * N);
    if (!f.z[i] || !an) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell backward, forward recomputation)");
    if (!ws.dry) {
      SMX_TRY(lin_fwd(blocks[i], f.a[i], blocks[i].in_dim, rows, f.z[i], N, true, 0, 0, nullptr, 1, st));
      SMX_TRY(act_fwd(f.z[i], rows, N, act, i == n - 1 ? mask : nullptr, an, st));
    }
    f.a[i + 1] = an;
  }
  return SMX_OK;
}

// da_n (rows, out_{n-1}; overwritten) -> parameter gradients, and the gradient with respect to x: dx_out = dX (+ dx_add)
int branch_bwd(const smx_linear* blocks, const smx_linear_grad* g, int n, int act, const BranchFwd& f, int64_t rows, const uint8_t* mask,
               float* da, void* dx_out, int dx_dt, const float* dx_add, bool want_dx, Arena& ws, cudaStream_t st) {
  float* cur = da;
  for (int i = n - 1; i >= 0; --i) {
    const smx_linear& L = blocks[i];
    if (!ws.dry) SMX_TRY(act_bwd(f.z[i], cur, SMX_F32, rows, L.out_dim, act, i == n - 1 ? mask : nullptr, cur, st));
    if (g[i].dw) SMX_TRY(lin_wgrad(L, cur, L.out_dim, f.a[i], L.in_dim, rows, g[i].dw, 0, 0, ws, st));
    if (g[i].db) SMX_TRY(colsum_all(cur, L.out_dim, rows, L.out_dim, g[i].db, ws, st));
    if (i > 0) {
      float* prev = ws.f32((size_t)rows * L.in_dim);
      if (!prev) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell backward)");
      if (!ws.dry) SMX_TRY(lin_dgrad(L, cur, L.out_dim, rows, prev, SMX_F32, L.in_dim, 0, 0, nullptr, st));
      cur = prev;
    } else if (want_dx) {
      if (!ws.dry) SMX_TRY(lin_dgrad(L, cur, L.out_dim, rows, dx_out, dx_dt, L.in_dim, 0, 0, dx_add, st));
    }
  }
  return SMX_OK;
}

// per-frame summaries under a (T,T) sum mask (Dynamic Chunk Training, summary_mixing.py:235-246, :292-294):
// Sf[b] = (M @ S[b]) / rowsum(M) -- the denominator keeps padded frames, like the reference -- and the gradient
// dS[b] = M^T @ (dSf[b] / rowsum(M)[:, None]) (dSf is scaled in place; dS has row stride ldd).  O(T^2 D) products on the CUDA-core GEMM.
__global__ void __launch_bounds__(256) div_rows_kernel(float* v, const float* rs, int T, int D, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) v[i] = v[i] / rs[(i / D) % T];
}
// iv (4 T + 1 ints: lo, hi, tlo, thi, flag) + pws (interval_means_workspace_bytes): the prefix-sum form for masks whose rows are runs of
// ones (the dynamic-chunk masks, TransformerASR.py:85-110), O(T D) per utterance; the structure is checked on the device and any other
// mask (weights, holes, the Laplace matrix) takes the (T,T) products.  iv == NULL: products only.
struct SumMaskWs { int* iv; void* pws; };
SumMaskWs summask_ws(Arena& ws, int B, int T, int D, bool try_intervals) {
  SumMaskWs r{nullptr, nullptr};
  if (try_intervals) {
    r.iv = (int*)ws.take((size_t)(4 * T + 1) * sizeof(int));
    r.pws = ws.take(interval_means_workspace_bytes(B, T, D));
    if (!r.iv || !r.pws) r = SumMaskWs{nullptr, nullptr};
  }
  return r;
}
int summask_fwd(const float* M, float* rs, const float* S, int64_t ldS, int B, int T, int Ds, float* Sf, const SumMaskWs& sw, cudaStream_t st) {
  int* flag = nullptr;
  if (sw.iv) {
    flag = sw.iv + 4 * T;
    SMX_TRY(interval_detect(M, T, sw.iv, sw.iv + T, flag, st));
    SMX_TRY(interval_means(S, ldS, B, T, Ds, sw.iv, sw.iv + T, flag, Sf, sw.pws, st));
  }
  SMX_TRY(rowsum(M, T, T, rs, st));
  GemmP p = bw_gemm();
  p.A = M; p.lda = T; p.a_bs = 0;
  p.W = S; p.w_sk = ldS; p.w_sn = 1; p.w_bs = (int64_t)T * ldS;
  p.rowdiv = rs;
  p.C = Sf; p.ldc = Ds; p.c_bs = (int64_t)T * Ds;
  p.M = T; p.N = Ds; p.K = T; p.batches = B;
  p.run_if_nonzero = flag;
  return gemm(p, st);
}
int summask_bwd(const float* M, const float* rs, float* dSf, int B, int T, int Ds, float* dS, int64_t ldd, const SumMaskWs& sw, cudaStream_t st) {
  const int64_t n = (int64_t)B * T * Ds;
  div_rows_kernel<<<ew_grid(n), 256, 0, st>>>(dSf, rs, T, Ds, n);
  count_launch();
  SMX_TRY(check_launch("div_rows_kernel"));
  int* flag = nullptr;
  if (sw.iv) {  // (summask_fwd has filled lo / hi and the flag)
    flag = sw.iv + 4 * T;
    SMX_TRY(interval_transpose(sw.iv, sw.iv + T, T, sw.iv + 2 * T, sw.iv + 3 * T, flag, st));
    SMX_TRY(interval_sums(dSf, Ds, B, T, Ds, sw.iv + 2 * T, sw.iv + 3 * T, flag, dS, ldd, sw.pws, st));
  }
  GemmP p = bw_gemm();
  p.A = M; p.lda = 1; p.a_sk = T; p.a_bs = 0;   // A[m][k] = M[k][m]
  p.W = dSf; p.w_sk = Ds; p.w_sn = 1; p.w_bs = (int64_t)T * Ds;
  p.C = dS; p.ldc = ldd; p.c_bs = (int64_t)T * ldd;
  p.M = T; p.N = Ds; p.K = T; p.batches = B;
  p.run_if_nonzero = flag;
  return gemm(p, st);
}

}  // namespace

int cell_bwd_generic(const smx_cell_weights* w, int B, int T, const void* x, int x_dt, const uint8_t* mask, const void* dy, int dy_dt,
                     void* dx, int dx_dt, const smx_cell_grads* g, Arena& ws, cudaStream_t st, const smx_dropout* drop, void* y_fwd, int y_dt,
                     const float* sum_mask) {
  // sum_mask ((T,T) fp32, modes "SummaryMixing" / "-fast"; "-lite" ignores it like the reference): per-frame summaries
  // (M @ S) / rowsum(M) instead of the utterance mean; the combiner then runs on the written-out concatenation.
  // drop (p > 0): dropout on the combiner's input cat = [local | summary] (summary_mixing.py:252, :297) -- the concatenation is
  // then written out per frame (the per-utterance bias shortcut no longer holds: every frame draws its own mask over the summary).
  // y_fwd: training-mode FORWARD only (same recomputation, result written to y_fwd, no gradients).
  const Drop dr = make_drop(drop);
  if ((dr.on || y_fwd) && w->mode == SMX_MODE_LITE && y_fwd)
    return fail(SMX_ERR_UNSUPPORTED, "cell training forward: mode 'SummaryMixing-lite' has no dropout (use smx_summary_mixing_fwd)");
  if (w->mode != SMX_MODE_FULL && w->mode != SMX_MODE_LITE && w->mode != SMX_MODE_FAST && w->mode != SMX_MODE_EXPDECAY)
    return fail(SMX_ERR_UNSUPPORTED, "smx_summary_mixing_bwd: unknown mode %d", w->mode);
  const int64_t rows = (int64_t)B * T;
  if (rows > 0x7fffffff) return fail(SMX_ERR_UNSUPPORTED, "cell backward: more than 2^31 frames");
  int maxdim = w->enc_dim;
  {
    const smx_linear* ls[2 * SMX_MAX_BLOCKS + 2];
    int nl = 0;
    for (int i = 0; i < w->n_local; ++i) ls[nl++] = &w->local[i];
    for (int i = 0; i < w->n_summary; ++i) ls[nl++] = &w->summary[i];
    if (w->mode == SMX_MODE_FAST) ls[nl++] = &w->global_proj;
    if (w->mode != SMX_MODE_LITE) ls[nl++] = &w->merge;
    for (int i = 0; i < nl; ++i) { maxdim = ls[i]->in_dim > maxdim ? ls[i]->in_dim : maxdim; maxdim = ls[i]->out_dim > maxdim ? ls[i]->out_dim : maxdim; }
  }
  BwScratch bw_scratch(ws, rows, maxdim);
  if (w->mode == SMX_MODE_FAST) {
    // G = act(W_g x + b_g) * mask (rows, 2 D_l); local = G[:, :D_l], S = G[:, D_l:]; mean_b = sum_t S / count_b;
    // y = act(local Wc[:, :D_l]^T + mean_b Wc[:, D_l:]^T + b_c); no LayerNorms            summary_mixing.py:255-298
    const int D = w->enc_dim, Dl = w->local_out_dim, Dout = w->merge.out_dim, act = w->act;
    if (w->global_proj.in_dim != D || w->global_proj.out_dim != 2 * Dl || w->global_proj.n_split > 1 || w->merge.in_dim != 2 * Dl ||
        w->merge.n_split > 1)
      return fail(SMX_ERR_BAD_ARG, "cell backward (fast): global_proj must be dense D -> 2*D_l and the merger 2*D_l -> D_s");
    const size_t m0 = ws.mark();
#define BWF_BUF(name, n) float* name = ws.f32((size_t)(n)); if (!name) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell backward)")
    const float* x32 = (const float*)x;
    if (x_dt != SMX_F32) {
      BWF_BUF(xc, rows * D);
      if (!ws.dry) SMX_TRY(convert(x, x_dt, xc, SMX_F32, rows * D, st));
      x32 = xc;
    }
    BranchFwd fg{};
    SMX_TRY(branch_fwd(&w->global_proj, 1, act, x32, rows, mask, fg, ws, st));
    const float* G = fg.a[1];
    BWF_BUF(mean, (size_t)B * Dl);
    BWF_BUF(cbias, (size_t)B * Dout);
    BWF_BUF(zc, rows * Dout);
    BWF_BUF(dcb, (size_t)B * Dout);
    BWF_BUF(dG, rows * 2 * Dl);
    BWF_BUF(dmean, (size_t)B * Dl);
    BWF_BUF(inv, (size_t)B);
    if (dr.on || y_fwd || sum_mask) {
      // written-out combiner: cat = dropout([G[:, :D_l] | mean_b]) (sum_mask: [G[:, :D_l] | Sf], per-frame summaries), zc = cat Wc^T + b_c
      BWF_BUF(cat, rows * 2 * Dl);
      float* rsm = nullptr; float* Sf = nullptr;
      SumMaskWs smw{nullptr, nullptr};
      if (sum_mask) {
        rsm = ws.f32((size_t)T); Sf = ws.f32((size_t)rows * Dl);
        smw = summask_ws(ws, B, T, Dl, true);
        if (!rsm || !Sf) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell backward, sum_mask)");
      }
      if (!ws.dry) {
        if (sum_mask) SMX_TRY(summask_fwd(sum_mask, rsm, G + Dl, 2 * Dl, B, T, Dl, Sf, smw, st));
        else SMX_TRY(masked_mean(G + Dl, 2 * Dl, mask, B, T, Dl, mean, SMX_F32, st));
        concat_bcast_kernel<<<ew_grid(rows * 2 * Dl), 256, 0, st>>>(G, 2 * Dl, sum_mask ? Sf : mean, sum_mask ? 1 : T, Dl, Dl, rows * 2 * Dl, cat);
        count_launch();
        SMX_TRY(check_launch("concat_bcast_kernel"));
        SMX_TRY(dropout_inplace(cat, rows * 2 * Dl, dr, 0, st));
        SMX_TRY(lin_fwd(w->merge, cat, 2 * Dl, rows, zc, Dout, true, 0, 0, nullptr, 1, st));
      }
      if (y_fwd) {
        if (!ws.dry) {
          SMX_TRY(act_fwd(zc, rows, Dout, act, nullptr, zc, st));
          SMX_TRY(convert(zc, SMX_F32, y_fwd, y_dt, rows * Dout, st));
        }
        ws.release(m0);
        return SMX_OK;
      }
      if (!ws.dry) SMX_TRY(act_bwd(zc, dy, dy_dt, rows, Dout, act, nullptr, zc, st));  // zc now holds dzc
      if (g->merge.dw) SMX_TRY(lin_wgrad(w->merge, zc, Dout, cat, 2 * Dl, rows, g->merge.dw, 0, 0, ws, st));
      if (g->merge.db) SMX_TRY(colsum_all(zc, Dout, rows, Dout, g->merge.db, ws, st));
      float* dcat = cat;  // cat is dead
      if (!ws.dry) {
        SMX_TRY(lin_dgrad(w->merge, zc, Dout, rows, dcat, SMX_F32, 2 * Dl, 0, 0, nullptr, st));
        SMX_TRY(dropout_inplace(dcat, rows * 2 * Dl, dr, 0, st));
        take_cols_kernel<<<ew_grid(rows * Dl), 256, 0, st>>>(dcat, 2 * Dl, Dl, rows * Dl, dG, 2 * Dl);  // dG[:, :D_l]
        count_launch();
        SMX_TRY(check_launch("take_cols_kernel"));
        if (sum_mask) {  // dG[:, D_l:] = M^T (dcat[:, D_l:] / rowsum)
          take_cols_kernel<<<ew_grid(rows * Dl), 256, 0, st>>>(dcat + Dl, 2 * Dl, Dl, rows * Dl, Sf, Dl);  // (Sf is dead)
          count_launch();
          SMX_TRY(check_launch("take_cols_kernel"));
          SMX_TRY(summask_bwd(sum_mask, rsm, Sf, B, T, Dl, dG + Dl, 2 * Dl, smw, st));
        } else {
        colsum_kernel<<<dim3((Dl + 31) / 32, B), 256, 0, st>>>(dcat + Dl, 2 * Dl, rows, T, Dl, dmean);  // d mean_b = sum_t dcat[b,t,D_l:]
        count_launch();
        SMX_TRY(check_launch("colsum_kernel"));
        inv_count_kernel<<<B, 32, 0, st>>>(mask, T, inv);
        count_launch();
        SMX_TRY(check_launch("inv_count_kernel"));
        bcast_scale_kernel<<<ew_grid(rows * Dl), 256, 0, st>>>(dmean, inv, T, Dl, rows * Dl, dG + Dl, 2 * Dl);  // dG[:, D_l:]
        count_launch();
        SMX_TRY(check_launch("bcast_scale_kernel"));
        }
      }
      SMX_TRY(branch_bwd(&w->global_proj, &g->global_proj, 1, act, fg, rows, mask, dG, dx, dx_dt, nullptr, dx != nullptr, ws, st));
      ws.release(m0);
      return SMX_OK;
    }
    if (!ws.dry) {
      SMX_TRY(masked_mean(G + Dl, 2 * Dl, mask, B, T, Dl, mean, SMX_F32, st));
      SMX_TRY(lin_fwd(w->merge, mean, Dl, B, cbias, Dout, true, Dl, Dl, nullptr, 1, st));
      SMX_TRY(lin_fwd(w->merge, G, 2 * Dl, rows, zc, Dout, false, 0, Dl, cbias, T, st));
      SMX_TRY(act_bwd(zc, dy, dy_dt, rows, Dout, act, nullptr, zc, st));  // zc now holds dzc
      colsum_kernel<<<dim3((Dout + 31) / 32, B), 256, 0, st>>>(zc, Dout, rows, T, Dout, dcb);
      count_launch();
      SMX_TRY(check_launch("colsum_kernel"));
    }
    if (g->merge.dw) {
      SMX_TRY(lin_wgrad(w->merge, zc, Dout, G, 2 * Dl, rows, g->merge.dw, 0, Dl, ws, st));
      SMX_TRY(lin_wgrad(w->merge, dcb, Dout, mean, Dl, B, g->merge.dw, Dl, Dl, ws, st));
    }
    if (g->merge.db) SMX_TRY(colsum_all(dcb, Dout, B, Dout, g->merge.db, ws, st));
    if (!ws.dry) {
      SMX_TRY(lin_dgrad(w->merge, zc, Dout, rows, dG, SMX_F32, 2 * Dl, 0, Dl, nullptr, st));      // dG[:, :D_l]
      SMX_TRY(lin_dgrad(w->merge, dcb, Dout, B, dmean, SMX_F32, Dl, Dl, Dl, nullptr, st));
      inv_count_kernel<<<B, 32, 0, st>>>(mask, T, inv);
      count_launch();
      SMX_TRY(check_launch("inv_count_kernel"));
      bcast_scale_kernel<<<ew_grid(rows * Dl), 256, 0, st>>>(dmean, inv, T, Dl, rows * Dl, dG + Dl, 2 * Dl);  // dG[:, D_l:]
      count_launch();
      SMX_TRY(check_launch("bcast_scale_kernel"));
    }
    SMX_TRY(branch_bwd(&w->global_proj, &g->global_proj, 1, act, fg, rows, mask, dG, dx, dx_dt, nullptr, dx != nullptr, ws, st));
#undef BWF_BUF
    ws.release(m0);
    return SMX_OK;
  }
  if (w->mode == SMX_MODE_LITE) {
    // y[b] = sum_t (s(x)[b,t] * mask) / count_b  (summary_mixing.py:300-324; y and dy are (B, D_s): the caller owns the
    // stride-0 expand over T and its gradient).  dS[b,t] = dy[b] / count_b, then the MLP backward of summary_proj.
    const int D = w->enc_dim, Ds = w->summary_out_dim, nsm = w->n_summary;
    if (nsm < 1 || nsm > SMX_MAX_BLOCKS || w->summary[0].in_dim != D || w->summary[nsm - 1].out_dim != Ds)
      return fail(SMX_ERR_BAD_ARG, "cell backward (lite): summary_proj dims");
    const size_t m0 = ws.mark();
    const float* x32 = (const float*)x;
    if (x_dt != SMX_F32) {
      float* xc = ws.f32((size_t)rows * D);
      if (!xc) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell backward)");
      if (!ws.dry) SMX_TRY(convert(x, x_dt, xc, SMX_F32, rows * D, st));
      x32 = xc;
    }
    BranchFwd fs{};
    SMX_TRY(branch_fwd(w->summary, nsm, w->act, x32, rows, mask, fs, ws, st));
    float* dy32 = ws.f32((size_t)B * Ds);
    float* inv = ws.f32((size_t)B);
    float* dS = ws.f32((size_t)rows * Ds);
    if (!dy32 || !inv || !dS) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell backward)");
    if (!ws.dry) {
      SMX_TRY(convert(dy, dy_dt, dy32, SMX_F32, (int64_t)B * Ds, st));
      inv_count_kernel<<<B, 32, 0, st>>>(mask, T, inv);
      count_launch();
      SMX_TRY(check_launch("inv_count_kernel"));
      bcast_scale_kernel<<<ew_grid(rows * Ds), 256, 0, st>>>(dy32, inv, T, Ds, rows * Ds, dS);
      count_launch();
      SMX_TRY(check_launch("bcast_scale_kernel"));
    }
    SMX_TRY(branch_bwd(w->summary, g->summary, nsm, w->act, fs, rows, mask, dS, dx, dx_dt, nullptr, dx != nullptr, ws, st));
    ws.release(m0);
    return SMX_OK;
  }
  const int D = w->enc_dim, Dl = w->local_out_dim, Ds = w->summary_out_dim, Dout = w->merge.out_dim;
  const int nl = w->n_local, nsm = w->n_summary, act = w->act;
  if (nl < 1 || nl > SMX_MAX_BLOCKS || nsm < 1 || nsm > SMX_MAX_BLOCKS) return fail(SMX_ERR_BAD_ARG, "cell backward: block counts");
  if (w->local[0].in_dim != D || w->summary[0].in_dim != D || w->local[nl - 1].out_dim != Dl || w->summary[nsm - 1].out_dim != Ds)
    return fail(SMX_ERR_BAD_ARG, "cell backward: projection dims do not match enc_dim / local_out_dim / summary_out_dim");
  if (w->merge.in_dim != Dl + Ds || w->merge.n_split > 1)
    return fail(SMX_ERR_BAD_ARG, "summary_local_merging must be dense with in_dim == D_l + D_s");
  const bool use_ln = w->use_layernorm != 0;
  if (use_ln && (!w->local_norm_w || !w->local_norm_b || !w->summary_norm_w || !w->summary_norm_b))
    return fail(SMX_ERR_BAD_ARG, "cell backward: use_layernorm without LayerNorm parameters");
  const size_t m0 = ws.mark();
#define BW_RUN(expr) do { if (!ws.dry) SMX_TRY(expr); } while (0)
#define BW_BUF(name, n) float* name = ws.f32((size_t)(n)); if (!name) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell backward)")

  // "SummaryMixing-expdecay" (summary_mixing.py:223-224): the sum mask is the Laplace-shaped weight matrix decay^|t - t'| (times the
  // caller's binary mask, if any); decay_constant is not trainable (requires_grad=False, :159-161), so this is the sum_mask path
  bool try_intervals = sum_mask != nullptr;
  if (w->mode == SMX_MODE_EXPDECAY) {
    BW_BUF(lap, (size_t)T * T);
    BW_RUN(laplace(w->decay_constant, sum_mask, T, lap, st));
    sum_mask = lap;
    try_intervals = false;
  }
  // ---- forward recomputation ----
  const float* x32 = (const float*)x;
  if (x_dt != SMX_F32) {
    BW_BUF(xc, rows * D);
    BW_RUN(convert(x, x_dt, xc, SMX_F32, rows * D, st));
    x32 = xc;
  }
  BranchFwd ff{}, fs{};
  SMX_TRY(branch_fwd(w->local, nl, act, x32, rows, mask, ff, ws, st));
  SMX_TRY(branch_fwd(w->summary, nsm, act, x32, rows, mask, fs, ws, st));
  const float* Lm = ff.a[nl];
  const float* Sm = fs.a[nsm];
  const float* Lmat = Lm;
  if (use_ln) {
    BW_BUF(Lb, rows * Dl);
    BW_RUN(layernorm(Lm, SMX_F32, Dl, w->local_norm_w, w->local_norm_b, 1e-5f, SMX_ACT_IDENTITY, Lb, SMX_F32, Dl, rows, Dl, st));
    Lmat = Lb;
  }
  // the summary: one row per utterance (the masked mean), or one per frame under a sum mask
  const int64_t srows = sum_mask ? rows : B;
  float* rsm = nullptr;
  SumMaskWs smw{nullptr, nullptr};
  BW_BUF(mean, (size_t)srows * Ds);
  if (sum_mask) {
    rsm = ws.f32((size_t)T);
    smw = summask_ws(ws, B, T, Ds, try_intervals);
    if (!rsm) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell backward, sum_mask)");
    BW_RUN(summask_fwd(sum_mask, rsm, Sm, Ds, B, T, Ds, mean, smw, st));
  } else {
    BW_RUN(masked_mean(Sm, Ds, mask, B, T, Ds, mean, SMX_F32, st));
  }
  const float* mu = mean;
  if (use_ln) {
    BW_BUF(mub, (size_t)srows * Ds);
    BW_RUN(layernorm(mean, SMX_F32, Ds, w->summary_norm_w, w->summary_norm_b, 1e-5f, SMX_ACT_IDENTITY, mub, SMX_F32, Ds, srows, Ds, st));
    mu = mub;
  }
  BW_BUF(zc, rows * Dout);
  BW_BUF(dL, rows * Dl);
  BW_BUF(dmu, (size_t)srows * Ds);
  float* dzc = zc;
  if (dr.on || y_fwd || sum_mask) {
    // written-out combiner: cat = dropout([L | mu_b]) (sum_mask: mu per frame), zc = cat Wc^T + b_c
    BW_BUF(cat, rows * (Dl + Ds));
    if (!ws.dry) {
      concat_bcast_kernel<<<ew_grid(rows * (Dl + Ds)), 256, 0, st>>>(Lmat, Dl, mu, sum_mask ? 1 : T, Dl, Ds, rows * (Dl + Ds), cat);
      count_launch();
      SMX_TRY(check_launch("concat_bcast_kernel"));
      SMX_TRY(dropout_inplace(cat, rows * (Dl + Ds), dr, 0, st));
      SMX_TRY(lin_fwd(w->merge, cat, Dl + Ds, rows, zc, Dout, true, 0, 0, nullptr, 1, st));
    }
    if (y_fwd) {
      if (!ws.dry) {
        SMX_TRY(act_fwd(zc, rows, Dout, act, nullptr, zc, st));
        SMX_TRY(convert(zc, SMX_F32, y_fwd, y_dt, rows * Dout, st));
      }
      ws.release(m0);
      return SMX_OK;
    }
    BW_RUN(act_bwd(zc, dy, dy_dt, rows, Dout, act, nullptr, dzc, st));
    if (g->merge.dw) SMX_TRY(lin_wgrad(w->merge, dzc, Dout, cat, Dl + Ds, rows, g->merge.dw, 0, 0, ws, st));
    if (g->merge.db) SMX_TRY(colsum_all(dzc, Dout, rows, Dout, g->merge.db, ws, st));
    float* dcat = cat;  // cat is dead
    if (!ws.dry) {
      SMX_TRY(lin_dgrad(w->merge, dzc, Dout, rows, dcat, SMX_F32, Dl + Ds, 0, 0, nullptr, st));
      SMX_TRY(dropout_inplace(dcat, rows * (Dl + Ds), dr, 0, st));
      take_cols_kernel<<<ew_grid(rows * Dl), 256, 0, st>>>(dcat, Dl + Ds, Dl, rows * Dl, dL, Dl);
      count_launch();
      SMX_TRY(check_launch("take_cols_kernel"));
      if (sum_mask) {  // d mu[b,t] = dcat[b,t,D_l:]
        take_cols_kernel<<<ew_grid(rows * Ds), 256, 0, st>>>(dcat + Dl, Dl + Ds, Ds, rows * Ds, dmu, Ds);
        count_launch();
        SMX_TRY(check_launch("take_cols_kernel"));
      } else {
      colsum_kernel<<<dim3((Ds + 31) / 32, B), 256, 0, st>>>(dcat + Dl, Dl + Ds, rows, T, Ds, dmu);  // d mu_b = sum_t dcat[b,t,D_l:]
      count_launch();
      SMX_TRY(check_launch("colsum_kernel"));
      }
    }
  } else {
  BW_BUF(cbias, (size_t)B * Dout);
  BW_RUN(lin_fwd(w->merge, mu, Ds, B, cbias, Dout, true, Dl, Ds, nullptr, 1, st));
  BW_RUN(lin_fwd(w->merge, Lmat, Dl, rows, zc, Dout, false, 0, Dl, cbias, T, st));

  // ---- combiner ----
  BW_RUN(act_bwd(zc, dy, dy_dt, rows, Dout, act, nullptr, dzc, st));
  BW_BUF(dcb, (size_t)B * Dout);
  if (!ws.dry) {  // per-utterance column sums
    colsum_kernel<<<dim3((Dout + 31) / 32, B), 256, 0, st>>>(dzc, Dout, rows, T, Dout, dcb);
    count_launch();
    SMX_TRY(check_launch("colsum_kernel"));
  }
  if (g->merge.dw) {
    SMX_TRY(lin_wgrad(w->merge, dzc, Dout, Lmat, Dl, rows, g->merge.dw, 0, Dl, ws, st));
    SMX_TRY(lin_wgrad(w->merge, dcb, Dout, mu, Ds, B, g->merge.dw, Dl, Ds, ws, st));
  }
  if (g->merge.db) SMX_TRY(colsum_all(dcb, Dout, B, Dout, g->merge.db, ws, st));
  BW_RUN(lin_dgrad(w->merge, dzc, Dout, rows, dL, SMX_F32, Dl, 0, Dl, nullptr, st));
  BW_RUN(lin_dgrad(w->merge, dcb, Dout, B, dmu, SMX_F32, Ds, Dl, Ds, nullptr, st));
  }

  // ---- summary branch: LN_s, mean, MLP ----
  if (use_ln) SMX_TRY(ln_bwd(mean, srows, Ds, w->summary_norm_w, dmu, g->summary_norm_dw, g->summary_norm_db, ws, st));
  BW_BUF(inv, (size_t)B);
  BW_BUF(dS, rows * Ds);
  if (sum_mask) {
    BW_RUN(summask_bwd(sum_mask, rsm, dmu, B, T, Ds, dS, Ds, smw, st));
  } else if (!ws.dry) {
    inv_count_kernel<<<B, 32, 0, st>>>(mask, T, inv);
    count_launch();
    SMX_TRY(check_launch("inv_count_kernel"));
    bcast_scale_kernel<<<ew_grid(rows * Ds), 256, 0, st>>>(dmu, inv, T, Ds, rows * Ds, dS);
    count_launch();
    SMX_TRY(check_launch("bcast_scale_kernel"));
  }
  float* dxs = nullptr;
  if (dx) {
    dxs = ws.f32((size_t)rows * D);
    if (!dxs) return fail(SMX_ERR_WORKSPACE, "workspace too small (cell backward)");
  }
  SMX_TRY(branch_bwd(w->summary, g->summary, nsm, act, fs, rows, mask, dS, dxs, SMX_F32, nullptr, dx != nullptr, ws, st));

  // ---- local branch: LN_l, MLP; dx = dx(local) + dx(summary) ----
  if (use_ln) SMX_TRY(ln_bwd(Lm, rows, Dl, w->local_norm_w, dL, g->local_norm_dw, g->local_norm_db, ws, st));
  SMX_TRY(branch_bwd(w->local, g->local, nl, act, ff, rows, mask, dL, dx, dx_dt, dxs, dx != nullptr, ws, st));
#undef BW_RUN
#undef BW_BUF
  ws.release(m0);
  return SMX_OK;
}


// =============================================================================================
// macaron half-step FFN, LayerNorm and convolution module: backward          Conformer.py:470-484, 322-338
// =============================================================================================
namespace {

// dst (fp32) = alpha * src
__global__ void __launch_bounds__(256) scale_kernel(const void* src, int s_dt, float alpha, int64_t n, float* dst) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) dst[i] = alpha * bw_ld(src, s_dt, i);
}
// dst (any dtype) = a + b (* rowmask)
__global__ void __launch_bounds__(256) add_kernel(const float* a, const float* b, int64_t n, void* dst, int d_dt) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float v = a[i] + (b ? b[i] : 0.0f);
    if (d_dt == SMX_BF16) ((__nv_bfloat16*)dst)[i] = __float2bfloat16_rn(v);
    else ((float*)dst)[i] = v;
  }
}
// dst (fp32) = src * rowmask[row]
__global__ void __launch_bounds__(256) mask_rows_kernel(const void* src, int s_dt, const uint8_t* rowmask, int ncols, int64_t n, float* dst) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    dst[i] = bw_ld(src, s_dt, i) * (rowmask ? (float)rowmask[i / ncols] : 1.0f);
}
// GLU backward: p = [value | gate] (rows, 2D); dp[:, :D] = dg * sigmoid(gate); dp[:, D:] = dg * value * s * (1 - s)
__global__ void __launch_bounds__(256) glu_bwd_kernel(const float* p, const float* dg, int64_t rows, int D, float* dp) {
  const int64_t n = rows * D;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t r = i / D;
    const int c = (int)(i % D);
    const float a = p[r * 2 * D + c], gt = p[r * 2 * D + D + c];
    const float s = 1.0f / (1.0f + expf(-gt));
    const float d = dg[i];
    dp[r * 2 * D + c] = d * s;
    dp[r * 2 * D + D + c] = d * a * s * (1.0f - s);
  }
}
// CSGU gate backward (u = [value | gate-half input], gp = gate pre-activation, dprod = gradient of value * gate_act(gp)):
// du[:, :H] = dprod * gate_act(gp);  dg = dprod * value * gate_act'(gp)
__global__ void __launch_bounds__(256) csgu_gate_bwd_kernel(const float* dprod, const float* gp, const float* u, int U, int H, int gate_act,
                                                            int64_t n, float* du, float* dg) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t row = i / H;
    const int c = (int)(i - row * H);
    const float d = dprod[i], z = gp[i];
    du[row * U + c] = d * bw_act(gate_act, z);
    dg[i] = d * u[row * U + c] * bw_dact(gate_act, z);
  }
}
// y[e] = keep(e) ? x[e] * scale : 0 in the I/O dtype (x and y may alias)
__global__ void __launch_bounds__(256) dropout_apply_kernel(const void* x, void* y, int dt, int64_t n, uint64_t seed, uint32_t site, uint32_t thresh,
                                                            float scale) {
  const int64_t quad = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e0 = quad * 4;
  if (e0 >= n) return;
  const uint4 r = drop_words(seed, site, quad);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (e0 + i < n) {
      const float v = w[i] >= thresh ? bw_ld(x, dt, e0 + i) * scale : 0.0f;
      if (dt == SMX_BF16) ((__nv_bfloat16*)y)[e0 + i] = __float2bfloat16_rn(v);
      else ((float*)y)[e0 + i] = v;
    }
}
// reflect-padded depthwise conv, data gradient, the part the zero-padded form misses: frames s <= pad and s >= T-1-pad also collect the
// taps that read their mirror images (padded frame -s and padded frame 2(T-1)-s).  din += ...; one thread per (utterance, boundary frame, channel).
__global__ void __launch_bounds__(256) dwconv_reflect_fix_kernel(const float* dout, const float* w, int B, int T, int C, int k, int pad, float* din,
                                                                 int64_t ldo) {
  const int nb = T <= 2 * pad + 2 ? T : 2 * pad + 2;
  const int64_t n = (int64_t)B * nb * C;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % C);
    const int r = (int)((i / C) % nb);
    const int b = (int)(i / ((int64_t)C * nb));
    const int t = T <= 2 * pad + 2 ? r : (r <= pad ? r : T - 1 - (r - pad - 1));
    const float* base = dout + (int64_t)b * T * C + c;
    float acc = 0.0f;
    for (int j = 0; j < k; ++j) {
      const float wj = w[(int64_t)c * k + j];
      const int ul = pad - j - t;
      if (t >= 1 && ul >= 0 && ul < T) acc = fmaf(wj, base[(int64_t)ul * C], acc);
      const int ur = 2 * (T - 1) - t - j + pad;
      if (t <= T - 2 && ur >= 0 && ur < T) acc = fmaf(wj, base[(int64_t)ur * C], acc);
    }
    din[((int64_t)b * T + t) * ldo + c] += acc;
  }
}
// depthwise conv, weight gradient, register-window form: a thread owns one channel and K accumulators, a warp a strip of 64 frames of one
// utterance (blocks of eight frames: 8 dout values against K + 7 input frames), a CTA eight consecutive strips; the warps' sums are
// added in fixed order through shared memory: P[slice = (utterance, 512-frame span)][c][j].
template <int K>
__global__ void __launch_bounds__(256) dwconv_wgrad_win_kernel(const float* __restrict__ dout, const float* __restrict__ in, int B, int T, int C,
                                                               int pad, int reflect, int spans, float* __restrict__ P) {
  __shared__ float red[8][K][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ncg = (C + 31) / 32;
  const int cg = blockIdx.x % ncg;
  const int slice = blockIdx.x / ncg;
  const int b = slice / spans, span = slice % spans;
  const int c = cg * 32 + lane;
  float acc[K];
#pragma unroll
  for (int j = 0; j < K; ++j) acc[j] = 0.0f;
  if (c < C) {
    const float* dob = dout + (int64_t)b * T * C + c;
    const float* inb = in + (int64_t)b * T * C + c;
    const int ts = span * 512 + warp * 64;
#pragma unroll 1
    for (int blk = 0; blk < 8; ++blk) {
      const int t0 = ts + blk * 8;
      if (t0 >= T) break;
      float d[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) d[o] = t0 + o < T ? dob[(int64_t)(t0 + o) * C] : 0.0f;
#pragma unroll
      for (int i = 0; i < K + 7; ++i) {
        int u = t0 + i - pad;
        if (reflect) u = u < 0 ? -u : (u >= T ? 2 * (T - 1) - u : u);
        const float x = (u >= 0 && u < T) ? inb[(int64_t)u * C] : 0.0f;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const int j = i - o;
          if (j >= 0 && j < K) acc[j] = fmaf(d[o], x, acc[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < K; ++j) red[warp][j][lane] = acc[j];
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * K; e += 256) {
    const int l = e / K, j = e - l * K;   // consecutive threads -> consecutive j of one channel: P rows are contiguous
    float tot = 0.0f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) tot += red[wv][j][l];
    if (cg * 32 + l < C) P[((int64_t)slice * C + cg * 32 + l) * K + j] = tot;
  }
}
// depthwise conv, gradient with respect to the input: din[b,t,c] = sum_j w[c,j] * dout[b, t - j + pad, c]; reflect != 0: the input
// was reflect-padded (frame -s and frame T-1+s read frames s and T-1-s: those taps' gradients land there too; T > pad).
// din has row stride ldo.
__global__ void __launch_bounds__(256) dwconv_bwd_data_kernel(const float* dout, const float* w, int B, int T, int C, int k, int pad,
                                                              int reflect, float* din, int64_t ldo, int chunk = 0) {
  const int64_t n = (int64_t)B * T * C;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % C);
    const int t = (int)((i / C) % T);
    const int b = (int)(i / ((int64_t)C * T));
    const float* base = dout + (int64_t)b * T * C + c;
    float acc = 0.0f;
    for (int j = 0; j < k; ++j) {
      const float wj = w[(int64_t)c * k + j];
      const int u = t - j + pad;   // the output frame whose tap j read input frame t
      // chunk > 0 (Dynamic Chunk Convolution): output u sees input frames below the end of ITS chunk only, i.e. t's chunk <= u's chunk
      if (u >= 0 && u < T && (chunk <= 0 || u >= (t / chunk) * chunk)) acc = fmaf(wj, base[(int64_t)u * C], acc);
      if (reflect) {
        const int ul = pad - j - t;                   // output frame whose tap j read padded frame -t
        if (t >= 1 && ul >= 0 && ul < T) acc = fmaf(wj, base[(int64_t)ul * C], acc);
        const int ur = 2 * (T - 1) - t - j + pad;     // ... read padded frame 2(T-1) - t
        if (t <= T - 2 && ur >= 0 && ur < T) acc = fmaf(wj, base[(int64_t)ur * C], acc);
      }
    }
    din[((int64_t)b * T + t) * ldo + c] = acc;
  }
}
// depthwise conv, weight gradient partials: P[slice][c][j] = sum over the slice's utterances and all t of
// dout[b,t,c] * in[b, t + j - pad, c].  grid (ceil(C/32), k, slices), block 32 x 8 (channels x frame lanes).
__global__ void __launch_bounds__(256) dwconv_bwd_w_kernel(const float* dout, const float* in, int B, int T, int C, int k, int pad,
                                                           int utt_per_slice, float* P, int reflect = 0, int chunk = 0) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x % 32, ry = threadIdx.x / 32;
  const int c = blockIdx.x * 32 + cx, j = blockIdx.y;
  const int b0 = blockIdx.z * utt_per_slice;
  const int b1 = b0 + utt_per_slice < B ? b0 + utt_per_slice : B;
  float acc = 0.0f;
  if (c < C) {
    for (int b = b0; b < b1; ++b) {
      const float* dob = dout + (int64_t)b * T * C + c;
      const float* inb = in + (int64_t)b * T * C + c;
      for (int t = ry; t < T; t += 8) {
        int u = t + j - pad;
        if (reflect) u = u < 0 ? -u : (u >= T ? 2 * (T - 1) - u : u);
        if (u >= 0 && u < T && (chunk <= 0 || u < (t / chunk + 1) * chunk)) acc = fmaf(dob[(int64_t)t * C], inb[(int64_t)u * C], acc);
      }
    }
  }
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < C) {
    float tot = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r) tot += red[r][cx];
    P[((int64_t)blockIdx.z * C + c) * k + j] = tot;
  }
}

}  // namespace

// din (row stride ldo) = gradient of the depthwise conv input; dw (C,1,k) = its weight gradient.  Register-window kernels for the
// common kernel sizes (k = 31, 15), the per-element kernels otherwise.
int dwconv_bwd_data(const float* dout, const float* w, int B, int T, int C, int k, int pad, int reflect, float* din, int64_t ldo,
                    cudaStream_t st, int chunk = 0) {
  int status = SMX_OK;
  if (chunk <= 0 && dwconv_window(dout, C, w, nullptr, B, T, C, k, k - 1 - pad, 0, 1, din, ldo, st, &status)) {
    SMX_TRY(status);
    if (reflect) {
      const int nb = T <= 2 * pad + 2 ? T : 2 * pad + 2;
      dwconv_reflect_fix_kernel<<<ew_grid((int64_t)B * nb * C), 256, 0, st>>>(dout, w, B, T, C, k, pad, din, ldo);
      count_launch();
      SMX_TRY(check_launch("dwconv_reflect_fix_kernel"));
    }
    return SMX_OK;
  }
  dwconv_bwd_data_kernel<<<ew_grid((int64_t)B * T * C), 256, 0, st>>>(dout, w, B, T, C, k, pad, reflect, din, ldo, chunk);
  count_launch();
  return check_launch("dwconv_bwd_data_kernel");
}
int dwconv_wgrad(const float* dout, const float* in, int B, int T, int C, int k, int pad, int reflect, float* dw, Arena& ws, cudaStream_t st,
                 int chunk = 0) {
  const size_t m0 = ws.mark();
  if (chunk <= 0 && (k == 31 || k == 15)) {
    const int spans = (T + 511) / 512, ns = B * spans, ncg = (C + 31) / 32;
    float* P = ws.f32((size_t)ns * C * k);
    if (!P) return fail(SMX_ERR_WORKSPACE, "workspace too small (depthwise weight gradient)");
    if (!ws.dry) {
      if (k == 31) dwconv_wgrad_win_kernel<31><<<(unsigned)(ns * ncg), 256, 0, st>>>(dout, in, B, T, C, pad, reflect, spans, P);
      else dwconv_wgrad_win_kernel<15><<<(unsigned)(ns * ncg), 256, 0, st>>>(dout, in, B, T, C, pad, reflect, spans, P);
      count_launch();
      SMX_TRY(check_launch("dwconv_wgrad_win_kernel"));
      SMX_TRY(sum_slices(P, ns, C, k, dw, k, st));
    }
  } else {
    const int ups = 4, ns = (B + ups - 1) / ups;
    float* P = ws.f32((size_t)ns * C * k);
    if (!P) return fail(SMX_ERR_WORKSPACE, "workspace too small (depthwise weight gradient)");
    if (!ws.dry) {
      dwconv_bwd_w_kernel<<<dim3((C + 31) / 32, k, ns), 256, 0, st>>>(dout, in, B, T, C, k, pad, ups, P, reflect, chunk);
      count_launch();
      SMX_TRY(check_launch("dwconv_bwd_w_kernel"));
      SMX_TRY(sum_slices(P, ns, C, k, dw, k, st));
    }
  }
  ws.release(m0);
  return SMX_OK;
}

#define BW_RUN(expr) do { if (!ws.dry) SMX_TRY(expr); } while (0)
#define BW_BUF(name, n) float* name = ws.f32((size_t)(n)); if (!name) return fail(SMX_ERR_WORKSPACE, "workspace too small (backward)")
#define BW_LAUNCH(what, ...) do { if (!ws.dry) { __VA_ARGS__; count_launch(); SMX_TRY(check_launch(what)); } } while (0)

// VanillaNN backward: n x (linear, act), activation after every block                 VanillaNN.py:168-196
int vanilla_bwd_generic(const smx_linear* blocks, int n, int act, int64_t rows, const void* x, int x_dt, const void* dy, int dy_dt,
                        void* dx, int dx_dt, const smx_linear_grad* g, Arena& ws, cudaStream_t st) {
  if (n < 1 || n > SMX_MAX_BLOCKS) return fail(SMX_ERR_UNSUPPORTED, "VanillaNN with %d blocks (library handles 1..%d)", n, SMX_MAX_BLOCKS);
  if (rows > 0x7fffffff) return fail(SMX_ERR_UNSUPPORTED, "VanillaNN backward: more than 2^31 rows");
  for (int i = 0; i < n; ++i)
    if (blocks[i].n_split > 1 && (blocks[i].in_dim % blocks[i].n_split || blocks[i].out_dim % blocks[i].n_split))
      return fail(SMX_ERR_BAD_ARG, "input_size and n_neurons must be dividible by n_split!");
  const size_t m0 = ws.mark();
  const int D = blocks[0].in_dim, N = blocks[n - 1].out_dim;
  int maxdim = D;
  for (int i = 0; i < n; ++i) maxdim = blocks[i].out_dim > maxdim ? blocks[i].out_dim : maxdim;
  BwScratch bw_scratch(ws, rows, maxdim);
  const float* x32 = (const float*)x;
  if (x_dt != SMX_F32) {
    BW_BUF(xc, rows * D);
    BW_RUN(convert(x, x_dt, xc, SMX_F32, rows * D, st));
    x32 = xc;
  }
  BranchFwd f{};
  SMX_TRY(branch_fwd(blocks, n, act, x32, rows, nullptr, f, ws, st));
  BW_BUF(da, rows * N);
  BW_RUN(convert(dy, dy_dt, da, SMX_F32, rows * N, st));
  SMX_TRY(branch_bwd(blocks, g, n, act, f, rows, nullptr, da, dx, dx_dt, nullptr, dx != nullptr, ws, st));
  ws.release(m0);
  return SMX_OK;
}

// nn.LayerNorm backward: dx (dtype tag) and fp32 parameter gradients
int layernorm_bwd_generic(const void* x, int x_dt, int64_t rows, int D, const float* w, float eps, const void* dy, int dy_dt, void* dx,
                          int dx_dt, float* dw, float* db, Arena& ws, cudaStream_t st) {
  const size_t m0 = ws.mark();
  const float* x32 = (const float*)x;
  if (x_dt != SMX_F32) {
    BW_BUF(xc, rows * D);
    BW_RUN(convert(x, x_dt, xc, SMX_F32, rows * D, st));
    x32 = xc;
  }
  BW_BUF(g, rows * D);
  BW_LAUNCH("scale_kernel", scale_kernel<<<ew_grid(rows * D), 256, 0, st>>>(dy, dy_dt, 1.0f, rows * D, g));
  SMX_TRY(ln_bwd(x32, rows, D, w, g, dw, db, ws, st, eps));
  if (dx) BW_LAUNCH("add_kernel", add_kernel<<<ew_grid(rows * D), 256, 0, st>>>(g, nullptr, rows * D, dx, dx_dt));
  ws.release(m0);
  return SMX_OK;
}

// y = x + 0.5 * W2 act(W1 LN(x) + b1) + 0.5 * b2;  y = LN_out(y) when out_ln_w               Conformer.py:470-484, 518, 547
int ffn_bwd_generic(const smx_ffn_weights* w, int act, int64_t rows, const void* x, int x_dt, const float* oln_w, const float* oln_b,
                    float oln_eps, const void* dy, int dy_dt, void* dx, int dx_dt, const smx_ffn_grads* g, Arena& ws, cudaStream_t st,
                    const smx_dropout* drop, void* y_fwd, int y_dt) {
  // drop (p > 0): site 0 = the dropout inside PositionalwiseFeedForward (after the activation), site 1 = the nn.Dropout that follows
  // the block in ffn_module (Conformer.py:470-484).  y_fwd: training-mode FORWARD only.
  const Drop dr = make_drop(drop);
  const int D = w->w1.in_dim, F = w->w1.out_dim;
  if (w->w2.in_dim != F || w->w2.out_dim != D || w->w1.n_split > 1 || w->w2.n_split > 1)
    return fail(SMX_ERR_BAD_ARG, "ffn backward: inconsistent dims");
  if (rows > 0x7fffffff) return fail(SMX_ERR_UNSUPPORTED, "ffn backward: more than 2^31 rows");
  const size_t m0 = ws.mark();
  BwScratch bw_scratch(ws, rows, D > F ? D : F);
  const float* x32 = (const float*)x;
  if (x_dt != SMX_F32) {
    BW_BUF(xc, rows * D);
    BW_RUN(convert(x, x_dt, xc, SMX_F32, rows * D, st));
    x32 = xc;
  }
  // forward recomputation
  BW_BUF(xn, rows * D);
  BW_RUN(layernorm(x32, SMX_F32, D, w->ln_w, w->ln_b, 1e-5f, SMX_ACT_IDENTITY, xn, SMX_F32, D, rows, D, st));
  BW_BUF(z1, rows * F);
  BW_RUN(lin_fwd(w->w1, xn, D, rows, z1, F, true, 0, 0, nullptr, 1, st));
  BW_BUF(h, rows * F);
  BW_RUN(act_fwd(z1, rows, F, act, nullptr, h, st));
  BW_RUN(dropout_inplace(h, rows * F, dr, 0, st));
  float* ypre_d = nullptr;  // x + 0.5 * dropout(h W2^T + b2), written out when dropout acts on u or only the forward is wanted
  if (y_fwd || (oln_w && dr.on)) {
    BW_BUF(u, rows * D);
    BW_RUN(lin_fwd(w->w2, h, F, rows, u, D, true, 0, 0, nullptr, 1, st));
    BW_RUN(dropout_inplace(u, rows * D, dr, 1, st));
    BW_LAUNCH("axpy_kernel", axpy_kernel<<<ew_grid(rows * D), 256, 0, st>>>(x32, u, 0.5f, rows * D, u));
    ypre_d = u;
    if (y_fwd) {
      if (oln_w) BW_RUN(layernorm(u, SMX_F32, D, oln_w, oln_b, oln_eps, SMX_ACT_IDENTITY, y_fwd, y_dt, D, rows, D, st));
      else BW_RUN(convert(u, SMX_F32, y_fwd, y_dt, rows * D, st));
      ws.release(m0);
      return SMX_OK;
    }
  }
  BW_BUF(gy, rows * D);  // gradient with respect to the pre-norm sum x + 0.5 u
  BW_LAUNCH("scale_kernel", scale_kernel<<<ew_grid(rows * D), 256, 0, st>>>(dy, dy_dt, 1.0f, rows * D, gy));
  if (oln_w && ypre_d) {
    SMX_TRY(ln_bwd(ypre_d, rows, D, oln_w, gy, g->out_ln_dw, g->out_ln_db, ws, st, oln_eps));
  } else if (oln_w) {
    BW_BUF(ypre, rows * D);
    if (!ws.dry && bw_tc_ok(rows, F, D, h, F, ypre, D, true) && ((uintptr_t)x32 % 16) == 0) {  // ypre = x + 0.5 * (h W2^T + b2), split-bf16 tensor-core GEMM
      GemmTc gt{};
      gt.bias = w->w2.b; gt.act = SMX_ACT_IDENTITY; gt.alpha = 0.5f; gt.resid_f32 = x32; gt.ldr = D; gt.out_f32 = ypre; gt.ldo = D;
      gt.rows_per_group = 1;
      SMX_TRY(tc_linear_split3(w->w2, 0, F, h, F, rows, gt, t_bw_sc, st));
    } else if (!ws.dry) {
      GemmP p = bw_gemm();
      p.A = h; p.lda = F; p.C = ypre; p.ldc = D; p.M = (int)rows; p.K = F; p.N = D;
      p.W = w->w2.w; p.w_sk = 1; p.w_sn = F; p.bias = w->w2.b;
      p.residual = x32; p.r_dtype = SMX_F32; p.ldr = D; p.alpha = 0.5f;
      SMX_TRY(gemm(p, st));
    }
    SMX_TRY(ln_bwd(ypre, rows, D, oln_w, gy, g->out_ln_dw, g->out_ln_db, ws, st, oln_eps));
  }
  BW_BUF(du, rows * D);
  BW_LAUNCH("scale_kernel", scale_kernel<<<ew_grid(rows * D), 256, 0, st>>>(gy, SMX_F32, 0.5f, rows * D, du));
  BW_RUN(dropout_inplace(du, rows * D, dr, 1, st));
  if (g->w2.dw) SMX_TRY(lin_wgrad(w->w2, du, D, h, F, rows, g->w2.dw, 0, 0, ws, st));
  if (g->w2.db) SMX_TRY(colsum_all(du, D, rows, D, g->w2.db, ws, st));
  BW_BUF(dh, rows * F);
  BW_RUN(lin_dgrad(w->w2, du, D, rows, dh, SMX_F32, F, 0, 0, nullptr, st));
  BW_RUN(dropout_inplace(dh, rows * F, dr, 0, st));
  BW_RUN(act_bwd(z1, dh, SMX_F32, rows, F, act, nullptr, dh, st));
  if (g->w1.dw) SMX_TRY(lin_wgrad(w->w1, dh, F, xn, D, rows, g->w1.dw, 0, 0, ws, st));
  if (g->w1.db) SMX_TRY(colsum_all(dh, F, rows, F, g->w1.db, ws, st));
  float* dxn = du;  // du is dead
  BW_RUN(lin_dgrad(w->w1, dh, F, rows, dxn, SMX_F32, D, 0, 0, nullptr, st));
  SMX_TRY(ln_bwd(x32, rows, D, w->ln_w, dxn, g->ln_dw, g->ln_db, ws, st));
  if (dx) BW_LAUNCH("add_kernel", add_kernel<<<ew_grid(rows * D), 256, 0, st>>>(gy, dxn, rows * D, dx, dx_dt));
  ws.release(m0);
  return SMX_OK;
}

// y = (Linear(act(LN_after(dwconv(GLU(pointwise(LN(x))))))) ) * mask                        Conformer.py:322-338
int convmod_bwd_generic(const smx_convmod_weights* w, int act, int B, int T, const void* x, int x_dt, const uint8_t* mask, const void* dy,
                        int dy_dt, void* dx, int dx_dt, const smx_convmod_grads* g, Arena& ws, cudaStream_t st, const smx_dropout* drop,
                        void* y_fwd, int y_dt, int chunk) {
  // chunk > 0: Dynamic Chunk Convolution (Conformer.py:197-320): a frame sees the past and the frames of its own chunk only.
  // drop (p > 0): site 0 = the nn.Dropout that ends after_conv, before the padding mask (Conformer.py:163, :334-337).
  // y_fwd: training-mode FORWARD only.
  const Drop dr = make_drop(drop);
  const int64_t rows = (int64_t)B * T;
  const int D = w->bottleneck.in_dim, k = w->kernel_size;
  if (w->bottleneck.out_dim != 2 * D || w->out.in_dim != D || w->out.out_dim != D || k < 1)
    return fail(SMX_ERR_BAD_ARG, "conv module backward: inconsistent dims");
  if (rows > 0x7fffffff) return fail(SMX_ERR_UNSUPPORTED, "conv module backward: more than 2^31 frames");
  const int pad = w->causal ? (k - 1) : (k - 1) / 2;
  if (chunk > 0 && w->causal) return fail(SMX_ERR_BAD_ARG, "Chunked convolution not supported with causal padding");
  const size_t m0 = ws.mark();
  BwScratch bw_scratch(ws, rows, 2 * D);
  const float* x32 = (const float*)x;
  if (x_dt != SMX_F32) {
    BW_BUF(xc, rows * D);
    BW_RUN(convert(x, x_dt, xc, SMX_F32, rows * D, st));
    x32 = xc;
  }
  // forward recomputation
  BW_BUF(xn, rows * D);
  BW_RUN(layernorm(x32, SMX_F32, D, w->ln_w, w->ln_b, 1e-5f, SMX_ACT_IDENTITY, xn, SMX_F32, D, rows, D, st));
  BW_BUF(p, rows * 2 * D);
  BW_RUN(lin_fwd(w->bottleneck, xn, D, rows, p, 2 * D, true, 0, 0, nullptr, 1, st));
  BW_BUF(gl, rows * D);
  BW_RUN(glu(p, rows, D, gl, st));
  BW_BUF(c, rows * D);
  BW_RUN(dwconv(gl, D, w->dw_w, w->dw_b, B, T, D, k, chunk > 0 ? SMX_CONV_CHUNKED : (w->causal ? SMX_CONV_CAUSAL : SMX_CONV_SAME_ZERO), chunk, c, D, st));
  BW_BUF(cn, rows * D);
  BW_RUN(layernorm(c, SMX_F32, D, w->after_ln_w, w->after_ln_b, 1e-5f, SMX_ACT_IDENTITY, cn, SMX_F32, D, rows, D, st));
  BW_BUF(a, rows * D);
  BW_RUN(act_fwd(cn, rows, D, act, nullptr, a, st));
  BW_BUF(dout, rows * D);
  if (y_fwd) {
    BW_RUN(lin_fwd(w->out, a, D, rows, dout, D, true, 0, 0, nullptr, 1, st));
    BW_RUN(dropout_inplace(dout, rows * D, dr, 0, st));
    BW_LAUNCH("mask_rows_kernel", mask_rows_kernel<<<ew_grid(rows * D), 256, 0, st>>>(dout, SMX_F32, mask, D, rows * D, dout));
    BW_RUN(convert(dout, SMX_F32, y_fwd, y_dt, rows * D, st));
    ws.release(m0);
    return SMX_OK;
  }
  // backward
  BW_LAUNCH("mask_rows_kernel", mask_rows_kernel<<<ew_grid(rows * D), 256, 0, st>>>(dy, dy_dt, mask, D, rows * D, dout));
  BW_RUN(dropout_inplace(dout, rows * D, dr, 0, st));
  if (g->out.dw) SMX_TRY(lin_wgrad(w->out, dout, D, a, D, rows, g->out.dw, 0, 0, ws, st));
  if (g->out.db) SMX_TRY(colsum_all(dout, D, rows, D, g->out.db, ws, st));
  float* da = a;  // a is dead
  BW_RUN(lin_dgrad(w->out, dout, D, rows, da, SMX_F32, D, 0, 0, nullptr, st));
  BW_RUN(act_bwd(cn, da, SMX_F32, rows, D, act, nullptr, da, st));
  SMX_TRY(ln_bwd(c, rows, D, w->after_ln_w, da, g->after_ln_dw, g->after_ln_db, ws, st));
  float* dc = da;
  if (g->dw_db) SMX_TRY(colsum_all(dc, D, rows, D, g->dw_db, ws, st));
  if (g->dw_dw) SMX_TRY(dwconv_wgrad(dc, gl, B, T, D, k, pad, 0, g->dw_dw, ws, st, chunk));
  float* dgl = dout;  // dout is dead
  BW_RUN(dwconv_bwd_data(dc, w->dw_w, B, T, D, k, pad, 0, dgl, D, st, chunk));
  BW_BUF(dp, rows * 2 * D);
  BW_LAUNCH("glu_bwd_kernel", glu_bwd_kernel<<<ew_grid(rows * D), 256, 0, st>>>(p, dgl, rows, D, dp));
  if (g->bottleneck.dw) SMX_TRY(lin_wgrad(w->bottleneck, dp, 2 * D, xn, D, rows, g->bottleneck.dw, 0, 0, ws, st));
  if (g->bottleneck.db) SMX_TRY(colsum_all(dp, 2 * D, rows, 2 * D, g->bottleneck.db, ws, st));
  float* dxn = c;  // c is dead (LN_after backward has run)
  BW_RUN(lin_dgrad(w->bottleneck, dp, 2 * D, rows, dxn, SMX_F32, D, 0, 0, nullptr, st));
  SMX_TRY(ln_bwd(x32, rows, D, w->ln_w, dxn, g->ln_dw, g->ln_db, ws, st));
  if (dx) BW_LAUNCH("add_kernel", add_kernel<<<ew_grid(rows * D), 256, 0, st>>>(dxn, nullptr, rows * D, dx, dx_dt));
  ws.release(m0);
  return SMX_OK;
}

// ConvolutionBranch: y = post( drop0( value * gate_act( [lin]( dwconv_reflect( LN(gate half) ) ) ) ) ),  [value | gate half] = act(pre(x))
// (Branchformer.py:86-97 + SpeechBrain's ConvolutionalSpatialGatingUnit).  y_fwd: forward only; else the backward from dy.
int convbranch_bwd_generic(const smx_convbranch_weights* w, int B, int T, const void* x, int x_dt, const void* dy, int dy_dt, void* dx,
                           int dx_dt, const smx_convbranch_grads* g, Arena& ws, cudaStream_t st, const smx_dropout* drop, void* y_fwd,
                           int y_dt) {
  const Drop dr = make_drop(drop);
  const int64_t rows = (int64_t)B * T;
  const int D = w->pre.in_dim, U = w->pre.out_dim, H = U / 2, k = w->kernel_size;
  if (U % 2) return fail(SMX_ERR_BAD_ARG, "Input size must be divisible by 2!");
  if (w->post.in_dim != H || w->post.out_dim != D || k < 1 || (k % 2) == 0) return fail(SMX_ERR_BAD_ARG, "convolution branch: inconsistent dims");
  if (w->csgu_linear.w && (w->csgu_linear.in_dim != H || w->csgu_linear.out_dim != H)) return fail(SMX_ERR_BAD_ARG, "convolution branch: csgu linear dims");
  const int pad = (k - 1) / 2;
  if (T <= pad) return fail(SMX_ERR_UNSUPPORTED, "convolution branch: reflect padding needs T > (kernel_size - 1) / 2 (T = %d)", T);
  if (rows > 0x7fffffff) return fail(SMX_ERR_UNSUPPORTED, "convolution branch: more than 2^31 frames");
  const size_t m0 = ws.mark();
  BwScratch bw_scratch(ws, rows, U);
  const float* x32 = (const float*)x;
  if (x_dt != SMX_F32) {
    BW_BUF(xc, rows * D);
    BW_RUN(convert(x, x_dt, xc, SMX_F32, rows * D, st));
    x32 = xc;
  }
  // forward recomputation
  BW_BUF(z, rows * U);
  BW_RUN(lin_fwd(w->pre, x32, D, rows, z, U, true, 0, 0, nullptr, 1, st));
  BW_BUF(u, rows * U);
  BW_RUN(act_fwd(z, rows, U, w->act, nullptr, u, st));
  BW_BUF(gl, rows * H);
  BW_RUN(layernorm(u + H, SMX_F32, U, w->csgu_ln_w, w->csgu_ln_b, 1e-5f, SMX_ACT_IDENTITY, gl, SMX_F32, H, rows, H, st));
  BW_BUF(gc, rows * H);
  BW_RUN(dwconv(gl, H, w->csgu_dw_w, w->csgu_dw_b, B, T, H, k, SMX_CONV_SAME_REFLECT, 0, gc, H, st));
  float* gp = gc;  // gate pre-activation
  if (w->csgu_linear.w) {
    BW_BUF(gpl, rows * H);
    BW_RUN(lin_fwd(w->csgu_linear, gc, H, rows, gpl, H, true, 0, 0, nullptr, 1, st));
    gp = gpl;
  }
  BW_BUF(prod, rows * H);
  BW_RUN(gate_mul(gp, H, u, U, w->gate_act, rows, H, prod, st));
  BW_RUN(dropout_inplace(prod, rows * H, dr, 0, st));
  if (y_fwd) {
    if (y_dt == SMX_F32 && ((uintptr_t)y_fwd % 32) == 0) {
      BW_RUN(lin_fwd(w->post, prod, H, rows, (float*)y_fwd, D, true, 0, 0, nullptr, 1, st));
    } else {
      BW_BUF(yo, rows * D);
      BW_RUN(lin_fwd(w->post, prod, H, rows, yo, D, true, 0, 0, nullptr, 1, st));
      BW_RUN(convert(yo, SMX_F32, y_fwd, y_dt, rows * D, st));
    }
    ws.release(m0);
    return SMX_OK;
  }
  // backward
  BW_BUF(dy32, rows * D);
  BW_LAUNCH("scale_kernel", scale_kernel<<<ew_grid(rows * D), 256, 0, st>>>(dy, dy_dt, 1.0f, rows * D, dy32));
  if (g->post.dw) SMX_TRY(lin_wgrad(w->post, dy32, D, prod, H, rows, g->post.dw, 0, 0, ws, st));
  if (g->post.db) SMX_TRY(colsum_all(dy32, D, rows, D, g->post.db, ws, st));
  float* dprod = prod;  // prod is dead
  BW_RUN(lin_dgrad(w->post, dy32, D, rows, dprod, SMX_F32, H, 0, 0, nullptr, st));
  BW_RUN(dropout_inplace(dprod, rows * H, dr, 0, st));
  BW_BUF(du, rows * U);
  BW_BUF(dg, rows * H);
  BW_LAUNCH("csgu_gate_bwd_kernel", csgu_gate_bwd_kernel<<<ew_grid(rows * H), 256, 0, st>>>(dprod, gp, u, U, H, w->gate_act, rows * H, du, dg));
  float* dgc = dg;
  if (w->csgu_linear.w) {
    if (g->csgu_linear.dw) SMX_TRY(lin_wgrad(w->csgu_linear, dg, H, gc, H, rows, g->csgu_linear.dw, 0, 0, ws, st));
    if (g->csgu_linear.db) SMX_TRY(colsum_all(dg, H, rows, H, g->csgu_linear.db, ws, st));
    dgc = dprod;  // dprod is dead
    BW_RUN(lin_dgrad(w->csgu_linear, dg, H, rows, dgc, SMX_F32, H, 0, 0, nullptr, st));
  }
  if (g->csgu_dw_db) SMX_TRY(colsum_all(dgc, H, rows, H, g->csgu_dw_db, ws, st));
  if (g->csgu_dw_dw) SMX_TRY(dwconv_wgrad(dgc, gl, B, T, H, k, pad, 1, g->csgu_dw_dw, ws, st));
  // d(LN output) straight into the gate half of du, then the LayerNorm backward in place there (input: the gate half of u)
  BW_RUN(dwconv_bwd_data(dgc, w->csgu_dw_w, B, T, H, k, pad, 1, du + H, U, st));
  SMX_TRY(ln_bwd(u + H, rows, H, w->csgu_ln_w, du + H, g->csgu_ln_dw, g->csgu_ln_db, ws, st, 1e-5f, U, U));
  BW_RUN(act_bwd(z, du, SMX_F32, rows, U, w->act, nullptr, du, st));
  if (g->pre.dw) SMX_TRY(lin_wgrad(w->pre, du, U, x32, D, rows, g->pre.dw, 0, 0, ws, st));
  if (g->pre.db) SMX_TRY(colsum_all(du, U, rows, U, g->pre.db, ws, st));
  if (dx) BW_RUN(lin_dgrad(w->pre, du, U, rows, dx, dx_dt, D, 0, 0, nullptr, st));
  ws.release(m0);
  return SMX_OK;
}
#undef BW_RUN
#undef BW_BUF
#undef BW_LAUNCH

int dropout_keep_mask(const smx_dropout* drop, int site, int64_t n, uint8_t* keep, cudaStream_t st) {
  const Drop dr = make_drop(drop);
  if (n <= 0) return SMX_OK;
  const int64_t quads = (n + 3) / 4;
  dropout_mask_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, st>>>(keep, n, dr.seed, (uint32_t)site, dr.on ? dr.thresh : 0u);
  count_launch();
  return check_launch("dropout_mask_kernel");
}

int dropout_apply(const smx_dropout* drop, int site, int dt, int64_t n, const void* x, void* y, cudaStream_t st) {
  const Drop dr = make_drop(drop);
  if (n <= 0) return SMX_OK;
  const int64_t quads = (n + 3) / 4;
  dropout_apply_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, st>>>(x, y, dt, n, dr.seed, (uint32_t)site, dr.on ? dr.thresh : 0u, dr.on ? dr.scale : 1.0f);
  count_launch();
  return check_launch("dropout_apply_kernel");
}

}  // namespace smx
