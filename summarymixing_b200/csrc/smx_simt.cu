// Generic fp32-math kernels of libsmx: the exact-precision arm of the SummaryMixing encoder path.
// Every tensor carries a runtime dtype tag (fp32 or bf16 storage); arithmetic is fp32 throughout.
// These kernels handle every shape and mode the reference accepts; the tcgen05 arm (smx_tc_*.cu)
// takes over for the bf16 shapes it supports.
#include <cstdio>
#include <cstdlib>
#include "smx_internal.h"
#include <math.h>

namespace smx {

__device__ __forceinline__ float ld_any(const void* p, int dt, int64_t i) {
  return dt == SMX_BF16 ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void st_any(void* p, int dt, int64_t i, float v) {
  if (dt == SMX_BF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
  else ((float*)p)[i] = v;
}

__device__ __forceinline__ float apply_act(int act, float x) {
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

// ---------------------------------------------------------------------------------------------
// batched strided GEMM with fused epilogue.  64x64 tile, BK=16, 256 threads, 4x4 per thread.
// ---------------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) gemm_kernel(GemmP p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  if (p.run_if_nonzero && *p.run_if_nonzero == 0) return;  // (uniform over the grid: the flag was written by an earlier kernel)
  const int batch = blockIdx.z % p.batches, outer = blockIdx.z / p.batches;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t a_off = (int64_t)batch * p.a_bs + (int64_t)outer * p.a_bs2;
  const float* W = p.W + (int64_t)batch * p.w_bs + (int64_t)outer * p.w_bs2;
  const int64_t ask = p.a_sk ? p.a_sk : 1;
  int K = p.K;
  if (p.k_total > 0) { const int rem = p.k_total - batch * p.K; K = rem < p.K ? rem : p.K; }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = tid + e * 256;  // 0..1023
      int kk, mm;  // the faster-varying index follows the unit stride of A
      if (ask == 1) { kk = idx % BK; mm = idx / BK; } else { mm = idx % BM; kk = idx / BM; }
      int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < p.M && k < K) ? ld_any(p.A, p.a_dtype, a_off + (int64_t)m * p.lda + (int64_t)k * ask) : 0.0f;
      // W: choose the faster-varying index according to the strides so that loads coalesce
      int kk2, nn2;
      if (p.w_sn == 1) { nn2 = idx % BN; kk2 = idx / BN; } else { kk2 = idx % BK; nn2 = idx / BK; }
      int n = n0 + nn2, k2 = k0 + kk2;
      Ws[kk2][nn2] = (n < p.N && k2 < K) ? W[(int64_t)k2 * p.w_sk + (int64_t)n * p.w_sn] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const float* bias = p.bias ? p.bias + (int64_t)batch * p.bias_bs : nullptr;
  const int64_t c_off = (int64_t)batch * p.c_bs + (int64_t)outer * p.c_bs2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    float rd = p.rowdiv ? p.rowdiv[m] : 1.0f;
    float rm = p.rowmask ? (float)p.rowmask[m] : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.rowdiv) v = v / rd;
      if (bias) v += bias[n];
      if (p.rowbias) v += p.rowbias[(int64_t)(m / p.rowbias_div) * p.rowbias_ld + n];
      v = apply_act(p.act, v);
      if (p.rowmask) v *= rm;
      if (p.residual) v = ld_any(p.residual, p.r_dtype, (int64_t)m * p.ldr + c_off + n) + p.alpha * v;
      st_any(p.C, p.c_dtype, (int64_t)m * p.ldc + c_off + n, v);
    }
  }
}


int gemm(const GemmP& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.batches <= 0) return fail(SMX_ERR_BAD_ARG, "gemm: empty problem");
  const int nz = p.batches * (p.batches2 > 0 ? p.batches2 : 1);
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, nz);
  if (grid.y > 65535) {  // split rows over several launches
    GemmP q = p;
    const int chunk = 65535 * BM;
    for (int m0 = 0; m0 < p.M; m0 += chunk) {
      q.M = (p.M - m0 < chunk) ? p.M - m0 : chunk;
      q.A = (const char*)p.A + (int64_t)m0 * p.lda * (p.a_dtype == SMX_BF16 ? 2 : 4);
      q.C = (char*)p.C + (int64_t)m0 * p.ldc * (p.c_dtype == SMX_BF16 ? 2 : 4);
      if (p.residual) q.residual = (const char*)p.residual + (int64_t)m0 * p.ldr * (p.r_dtype == SMX_BF16 ? 2 : 4);
      if (p.rowmask) q.rowmask = p.rowmask + m0;
      if (p.rowdiv) q.rowdiv = p.rowdiv + m0;
      if (p.rowbias) {
        if (m0 % p.rowbias_div) return fail(SMX_ERR_UNSUPPORTED, "gemm: row split not aligned to rowbias_div");
        q.rowbias = p.rowbias + (int64_t)(m0 / p.rowbias_div) * p.rowbias_ld;
      }
      SMX_TRY(gemm(q, st));
    }
    return SMX_OK;
  }
  gemm_kernel<<<grid, 256, 0, st>>>(p);
  count_launch();
  return check_launch("gemm_kernel");
}

// ---------------------------------------------------------------------------------------------
// LayerNorm (+ optional activation) over rows of length D; one warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_kernel(const void* x, int x_dt, int64_t ldx, const float* w, const float* b,
                                                 float eps, int act, void* y, int y_dt, int64_t ldy, int64_t rows, int D) {
  int64_t row = (int64_t)blockIdx.x * 8 + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (row >= rows) return;
  float s = 0.0f;
  for (int c = lane; c < D; c += 32) s += ld_any(x, x_dt, row * ldx + c);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float mean = s / (float)D;
  float v = 0.0f;
  for (int c = lane; c < D; c += 32) {
    float d = ld_any(x, x_dt, row * ldx + c) - mean;
    v += d * d;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  float rstd = 1.0f / sqrtf(v / (float)D + eps);
  for (int c = lane; c < D; c += 32) {
    float o = (ld_any(x, x_dt, row * ldx + c) - mean) * rstd * w[c] + b[c];
    st_any(y, y_dt, row * ldy + c, apply_act(act, o));
  }
}

// bf16 -> bf16 rows of D = 256 NCH elements: one warp per row, the row held in registers (128-bit loads and stores, one pass over
// memory; two-pass statistics like the generic kernel).  The generic kernel's 2-byte accesses and three passes ran at 1.7 TB/s
// (38 us for 32 000 x 512), a fifth of the D = 512 layers' time.  The pass is latency-bound: bytes in flight per SM = resident warps x
// 2 rows, so the plain (no activation) instance is compiled without the activation switch and for 5 / 4 resident blocks
// (48 / 64 registers): 15.1 -> 8.4 us at 32 000 x 256, 20.8 -> 15.3 us at x 512 (3.9 / 4.3 TB/s; tools/ln_time.py).
template <int NCH, bool IDENT>   // IDENT: no activation after the normalisation (the common case; the activation switch costs registers, i.e. resident warps)
__global__ void __launch_bounds__(256, (!IDENT ? 2 : (NCH == 1 ? 5 : (NCH == 2 ? 4 : 2)))) ln_bf16_rows_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                           const float* __restrict__ b, float eps, int act, __nv_bfloat16* __restrict__ y,
                                                           int64_t ldy, int64_t rows) {
  constexpr int RPW = 2;  // rows per warp: both rows' loads are in flight before the first reduction
  const int64_t row0 = ((int64_t)blockIdx.x * 8 + threadIdx.x / 32) * RPW;
  const int lane = threadIdx.x % 32;
  if (row0 >= rows) return;
  uint4 raw[RPW][NCH];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      raw[r][c] = row0 + r < rows ? *reinterpret_cast<const uint4*>(x + (row0 + r) * ldx + (c * 32 + lane) * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    if (row0 + r >= rows) break;
    float v[NCH][8];
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[r][c]);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); v[c][2 * e] = f.x; v[c][2 * e + 1] = f.y; s += f.x + f.y; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)(NCH * 256);
    float q = 0.0f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[c][e] - mean; q = fmaf(d, d, q); }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.0f / sqrtf(q / (float)(NCH * 256) + eps);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = (c * 32 + lane) * 8;
      const float4 w0 = *reinterpret_cast<const float4*>(w + col), w1 = *reinterpret_cast<const float4*>(w + col + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(b + col), b1 = *reinterpret_cast<const float4*>(b + col + 4);
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float t = (v[c][e] - mean) * rstd * ww[e] + bb[e];
        o[e] = IDENT ? t : apply_act(act, t);
      }
      uint4 out;
      __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
      for (int e = 0; e < 4; ++e) ho[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
      *reinterpret_cast<uint4*>(y + (row0 + r) * ldy + col) = out;
    }
  }
}

int layernorm(const void* x, int x_dtype, int64_t ldx, const float* w, const float* b, float eps, int act, void* y,
              int y_dtype, int64_t ldy, int64_t rows, int D, cudaStream_t st) {
  if (rows <= 0 || D <= 0) return fail(SMX_ERR_BAD_ARG, "layernorm: empty problem");
  if (x_dtype == SMX_BF16 && y_dtype == SMX_BF16 && D % 256 == 0 && D <= 1024 && ldx % 8 == 0 && ldy % 8 == 0 &&
      ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)b % 16 == 0)) {
    const unsigned grid = (unsigned)((rows + 15) / 16);  // 8 warps x 2 rows per block
    const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
    __nv_bfloat16* yb = (__nv_bfloat16*)y;
    switch (D / 256) {
#define SMX_LN_CASE(n) \
      if (act == SMX_ACT_IDENTITY) ln_bf16_rows_kernel<n, true><<<grid, 256, 0, st>>>(xb, ldx, w, b, eps, act, yb, ldy, rows); \
      else ln_bf16_rows_kernel<n, false><<<grid, 256, 0, st>>>(xb, ldx, w, b, eps, act, yb, ldy, rows); \
      break
      case 1: SMX_LN_CASE(1);
      case 2: SMX_LN_CASE(2);
      case 3: SMX_LN_CASE(3);
      default: SMX_LN_CASE(4);
#undef SMX_LN_CASE
    }
    count_launch();
    return check_launch("ln_bf16_rows_kernel");
  }
  ln_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, x_dtype, ldx, w, b, eps, act, y, y_dtype, ldy, rows, D);
  count_launch();
  return check_launch("ln_kernel");
}

// ---------------------------------------------------------------------------------------------
// masked temporal mean: out[b,d] = sum_t s[b,t,d] / sum_t mask[b,t]   (s is already masked)
// summary_mixing.py:229-231.  Fixed reduction order: deterministic.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) masked_mean_kernel(const float* s, int64_t lds, const uint8_t* mask, int T, int D,
                                                          void* out, int out_dt) {
  __shared__ float red[8][33];
  __shared__ float cnt[8];
  int b = blockIdx.y;
  int cx = threadIdx.x % 32, ry = threadIdx.x / 32;
  int d = blockIdx.x * 32 + cx;
  float acc = 0.0f, c = 0.0f;
  for (int t = ry; t < T; t += 8) {
    if (d < D) acc += s[((int64_t)b * T + t) * lds + d];
    c += mask ? (float)mask[(int64_t)b * T + t] : 1.0f;
  }
  red[ry][cx] = acc;
  if (cx == 0) cnt[ry] = c;
  __syncthreads();
  if (ry == 0 && d < D) {
    float tot = 0.0f, n = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r) { tot += red[r][cx]; n += cnt[r]; }
    st_any(out, out_dt, (int64_t)b * D + d, tot / n);
  }
}

int masked_mean(const float* s, int64_t lds, const uint8_t* mask, int B, int T, int D, void* out, int out_dtype,
                cudaStream_t st) {
  dim3 grid((D + 31) / 32, B);
  masked_mean_kernel<<<grid, 256, 0, st>>>(s, lds, mask, T, D, out, out_dtype);
  count_launch();
  return check_launch("masked_mean_kernel");
}

// ---------------------------------------------------------------------------------------------
// GLU over the channel dim: out[r,c] = p[r,c] * sigmoid(p[r,c+D])          Conformer.py:139
// ---------------------------------------------------------------------------------------------
__global__ void glu_kernel(const float* p, int64_t rows, int D, float* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  int64_t r = i / D;
  int c = (int)(i % D);
  float a = p[r * 2 * D + c], g = p[r * 2 * D + D + c];
  out[i] = a / (1.0f + expf(-g));
}
int glu(const float* p, int64_t rows, int D, float* out, cudaStream_t st) {
  int64_t n = rows * D;
  glu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, rows, D, out);
  count_launch();
  return check_launch("glu_kernel");
}

// ---------------------------------------------------------------------------------------------
// depthwise conv along T with the four boundary rules of smx_conv_pad.
// ---------------------------------------------------------------------------------------------
__global__ void dwconv_kernel(const float* in, int64_t ldin, const float* w, const float* bias, int B, int T, int C,
                              int k, int pad_mode, int chunk, float* out, int64_t ldout) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * T * C) return;
  int c = (int)(i % C);
  int t = (int)((i / C) % T);
  int b = (int)(i / ((int64_t)C * T));
  int pad = (pad_mode == SMX_CONV_CAUSAL) ? (k - 1) : (k - 1) / 2;
  int hi = T;  // exclusive upper bound of visible frames
  if (pad_mode == SMX_CONV_CHUNKED) {
    int ce = (t / chunk + 1) * chunk;
    hi = ce < T ? ce : T;
  }
  float acc = bias ? bias[c] : 0.0f;
  const float* base = in + (int64_t)b * T * ldin + c;
  for (int j = 0; j < k; ++j) {
    int u = t + j - pad;
    if (pad_mode == SMX_CONV_SAME_REFLECT) {
      if (u < 0) u = -u;
      if (u >= T) u = 2 * (T - 1) - u;
    } else if (u < 0 || u >= hi) {
      continue;
    }
    acc = fmaf(w[(int64_t)c * k + j], base[(int64_t)u * ldin], acc);
  }
  out[((int64_t)b * T + t) * ldout + c] = acc;
}
// Register-window form for the common kernel sizes: a thread owns one channel (its K taps in registers) and blocks of eight
// consecutive frames (8 accumulators, K + 7 loads per 8 K FMAs; lanes = 32 consecutive channels: coalesced rows).  A warp walks
// four frame blocks of one utterance.  flip: taps reversed (the data gradient of a zero-padded convolution is the correlation with
// the reversed taps and pad' = K - 1 - pad).  reflect: frames outside [0, T) read their mirror image, else zero.
template <int K>
__global__ void __launch_bounds__(256) dwconv_win_kernel(const float* __restrict__ in, int64_t ldin, const float* __restrict__ w,
                                                         const float* __restrict__ bias, int B, int T, int C, int pad, int reflect, int flip,
                                                         float* __restrict__ out, int64_t ldout, int strips) {
  const int lane = threadIdx.x & 31;
  const int64_t wg = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int ncg = (C + 31) / 32;
  const int cg = (int)(wg % ncg);
  const int strip = (int)((wg / ncg) % strips);
  const int b = (int)(wg / ((int64_t)ncg * strips));
  const int c = cg * 32 + lane;
  if (b >= B || c >= C) return;
  float wt[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wt[j] = w[(int64_t)c * K + (flip ? K - 1 - j : j)];
  const float bv = bias ? bias[c] : 0.0f;
  const float* base = in + (int64_t)b * T * ldin + c;
  float* obase = out + (int64_t)b * T * ldout + c;
#pragma unroll 1
  for (int blk = 0; blk < 4; ++blk) {
    const int t0 = strip * 32 + blk * 8;
    if (t0 >= T) break;
    float a[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) a[o] = bv;
#pragma unroll
    for (int i = 0; i < K + 7; ++i) {
      int u = t0 + i - pad;
      if (reflect) u = u < 0 ? -u : (u >= T ? 2 * (T - 1) - u : u);
      const float x = (u >= 0 && u < T) ? base[(int64_t)u * ldin] : 0.0f;
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int j = i - o;
        if (j >= 0 && j < K) a[o] = fmaf(wt[j], x, a[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o)
      if (t0 + o < T) obase[(int64_t)(t0 + o) * ldout] = a[o];
  }
}
template <int K>
static int launch_dwconv_win(const float* in, int64_t ldin, const float* w, const float* b, int B, int T, int C, int pad, int reflect, int flip,
                             float* out, int64_t ldout, cudaStream_t st) {
  const int strips = (T + 31) / 32, ncg = (C + 31) / 32;
  const int64_t warps = (int64_t)B * strips * ncg;
  dwconv_win_kernel<K><<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(in, ldin, w, b, B, T, C, pad, reflect, flip, out, ldout, strips);
  count_launch();
  return check_launch("dwconv_win_kernel");
}
// out[b,t,c] = bias[c] + sum_j w[c, flip ? k-1-j : j] * in[b, t + j - pad, c] with zero or mirrored frames outside [0, T); false when
// this kernel size has no register-window instance (the caller keeps its generic kernel)
bool dwconv_window(const float* in, int64_t ldin, const float* w, const float* b, int B, int T, int C, int k, int pad, int reflect, int flip,
                   float* out, int64_t ldout, cudaStream_t st, int* status) {
  switch (k) {
    case 31: *status = launch_dwconv_win<31>(in, ldin, w, b, B, T, C, pad, reflect, flip, out, ldout, st); return true;
    case 15: *status = launch_dwconv_win<15>(in, ldin, w, b, B, T, C, pad, reflect, flip, out, ldout, st); return true;
    default: return false;
  }
}

int dwconv(const float* in, int64_t ldin, const float* w, const float* b, int B, int T, int C, int k, int pad_mode,
           int chunk, float* out, int64_t ldout, cudaStream_t st) {
  if (pad_mode == SMX_CONV_SAME_REFLECT && (k - 1) / 2 >= T)
    return fail(SMX_ERR_BAD_ARG, "reflect padding %d needs T > pad (T=%d)", (k - 1) / 2, T);
  if (pad_mode == SMX_CONV_CHUNKED && chunk <= 0) return fail(SMX_ERR_BAD_ARG, "chunked conv needs chunk_size > 0");
  if (pad_mode != SMX_CONV_CHUNKED) {
    int status = SMX_OK;
    const int pad = (pad_mode == SMX_CONV_CAUSAL) ? (k - 1) : (k - 1) / 2;
    if (dwconv_window(in, ldin, w, b, B, T, C, k, pad, pad_mode == SMX_CONV_SAME_REFLECT, 0, out, ldout, st, &status)) return status;
  }
  int64_t n = (int64_t)B * T * C;
  dwconv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, ldin, w, b, B, T, C, k, pad_mode, chunk, out, ldout);
  count_launch();
  return check_launch("dwconv_kernel");
}

// out = act(gate) * other                                   (CSGU gating)
__global__ void gate_mul_kernel(const float* gate, int64_t ldg, const float* other, int64_t ldo, int act, int64_t rows,
                                int C, float* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  int64_t r = i / C;
  int c = (int)(i % C);
  out[i] = apply_act(act, gate[r * ldg + c]) * other[r * ldo + c];
}
int gate_mul(const float* gate, int64_t ldg, const float* other, int64_t ldo, int gate_act, int64_t rows, int C,
             float* out, cudaStream_t st) {
  int64_t n = rows * C;
  gate_mul_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gate, ldg, other, ldo, gate_act, rows, C, out);
  count_launch();
  return check_launch("gate_mul_kernel");
}

// dst[(b*T+t)*ld + d] = src[b*D + d]
__global__ void broadcast_rows_kernel(const float* src, int B, int T, int D, float* dst, int64_t ld) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * T * D) return;
  int d = (int)(i % D);
  int64_t bt = i / D;
  int b = (int)(bt / T);
  dst[bt * ld + d] = src[(int64_t)b * D + d];
}
int broadcast_rows(const float* src, int B, int T, int D, float* dst, int64_t lddst, cudaStream_t st) {
  int64_t n = (int64_t)B * T * D;
  broadcast_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, B, T, D, dst, lddst);
  count_launch();
  return check_launch("broadcast_rows_kernel");
}

// out[r,c] = a[r,c] + s[(r/div),c]
__global__ void add_bcast_kernel(const float* a, const float* s, int64_t rows, int div, int D, float* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  int64_t r = i / D;
  int c = (int)(i % D);
  out[i] = a[i] + s[(r / div) * D + c];
}
int add_bcast(const float* a, const float* s, int64_t rows, int div, int D, float* out, cudaStream_t st) {
  int64_t n = rows * D;
  add_bcast_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, s, rows, div, D, out);
  count_launch();
  return check_launch("add_bcast_kernel");
}

// Laplace weights decay^|i-j| (* binary mask)                  summary_mixing.py:330-379
__global__ void laplace_kernel(float log_decay, const float* binary, int T, float* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T * T) return;
  int r = (int)(i / T), c = (int)(i % T);
  float v = expf((float)abs(r - c) * log_decay);
  out[i] = binary ? v * binary[i] : v;
}
int laplace(float decay, const float* binary, int T, float* out, cudaStream_t st) {
  int64_t n = (int64_t)T * T;
  laplace_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(logf(decay), binary, T, out);
  count_launch();
  return check_launch("laplace_kernel");
}

// ---- chunk-prefix path of the sum-mask summaries ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) interval_detect_kernel(const float* __restrict__ M, int T, int* __restrict__ lo, int* __restrict__ hi,
                                                              int* __restrict__ not_interval) {
  const int row = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= T) return;
  int first = T, last = -1, cnt = 0, bad = 0;
  for (int c = lane; c < T; c += 32) {
    const float v = M[(int64_t)row * T + c];
    if (v != 0.0f) {
      if (v != 1.0f) bad = 1;
      first = c < first ? c : first;
      last = c > last ? c : last;
      ++cnt;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const int f2 = __shfl_xor_sync(0xffffffffu, first, o), l2 = __shfl_xor_sync(0xffffffffu, last, o);
    first = f2 < first ? f2 : first;
    last = l2 > last ? l2 : last;
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if (lane == 0) {
    if (bad || cnt == 0 || cnt != last - first + 1) atomicExch(not_interval, 1);
    lo[row] = first; hi[row] = last + 1;
  }
}
int interval_detect(const float* M, int T, int* lo, int* hi, int* not_interval, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(not_interval, 0, sizeof(int), st);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  interval_detect_kernel<<<(T + 7) / 8, 256, 0, st>>>(M, T, lo, hi, not_interval);
  count_launch();
  return check_launch("interval_detect_kernel");
}
// P[b][t][d] = sum_{j < t} S[b][j][d] (exclusive, fp64), one thread per (b, d): coalesced over d
__global__ void __launch_bounds__(256) prefix_time_kernel(const float* __restrict__ S, int64_t ldS, int B, int T, int D, const int* __restrict__ skip,
                                                          double* __restrict__ P) {
  if (*skip) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * D) return;
  const int b = (int)(i / D), d = (int)(i % D);
  const float* s = S + (int64_t)b * T * ldS + d;
  double* p = P + (int64_t)b * (T + 1) * D + d;
  double acc = 0.0;
  p[0] = 0.0;
#pragma unroll 4
  for (int t = 0; t < T; ++t) {
    acc += (double)s[(int64_t)t * ldS];
    p[(int64_t)(t + 1) * D] = acc;
  }
}
__global__ void __launch_bounds__(256) interval_mean_kernel(const double* __restrict__ P, const int* __restrict__ lo, const int* __restrict__ hi, int B,
                                                            int T, int D, const int* __restrict__ skip, float* __restrict__ Sm, int64_t ldo = 0,
                                                            int as_sum = 0) {
  if (*skip) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * T * D) return;
  const int d = (int)(i % D);
  const int64_t bt = i / D;
  const int t = (int)(bt % T), b = (int)(bt / T);
  const int l = lo[t], h = hi[t];
  const double* p = P + (int64_t)b * (T + 1) * D + d;
  // the reference divides by the row sum of the mask: padding is NOT removed from the denominator (summary_mixing.py:239-246)
  if (as_sum) { Sm[bt * ldo + d] = h > l ? (float)(p[(int64_t)h * D] - p[(int64_t)l * D]) : 0.0f; return; }  // (the gradient form)
  Sm[i] = (float)((p[(int64_t)h * D] - p[(int64_t)l * D]) / (double)(h - l));
}
// The transposed intervals of a mask whose rows t are runs [lo_t, hi_t) with lo and hi non-decreasing in t (the dynamic-chunk masks):
// column t' is covered by the rows [tlo, thi) with tlo = first t with hi_t > t', thi = first t with lo_t > t'.  Any other structure sets
// *not_interval (one block).
__global__ void __launch_bounds__(256) interval_transpose_kernel(const int* __restrict__ lo, const int* __restrict__ hi, int T, int* __restrict__ tlo,
                                                                 int* __restrict__ thi, int* __restrict__ not_interval) {
  if (*not_interval) return;
  for (int t = threadIdx.x + 1; t < T; t += 256)
    if (lo[t] < lo[t - 1] || hi[t] < hi[t - 1]) atomicExch(not_interval, 1);
  for (int c = threadIdx.x; c < T; c += 256) {
    int a = 0, b = T;   // first t with hi[t] > c
    while (a < b) { const int m = (a + b) >> 1; if (hi[m] > c) b = m; else a = m + 1; }
    tlo[c] = a;
    a = 0; b = T;       // first t with lo[t] > c
    while (a < b) { const int m = (a + b) >> 1; if (lo[m] > c) b = m; else a = m + 1; }
    thi[c] = a;
  }
}
int interval_transpose(const int* lo, const int* hi, int T, int* tlo, int* thi, int* not_interval, cudaStream_t st) {
  interval_transpose_kernel<<<1, 256, 0, st>>>(lo, hi, T, tlo, thi, not_interval);
  count_launch();
  return check_launch("interval_transpose_kernel");
}
// out[b,t,:] (row stride ldo) = sum_{j in [lo_t, hi_t)} S[b,j,:]; skipped when *not_interval
int interval_sums(const float* S, int64_t ldS, int B, int T, int D, const int* lo, const int* hi, const int* not_interval, float* out, int64_t ldo,
                  void* workspace, cudaStream_t st) {
  double* P = (double*)workspace;
  const int64_t n1 = (int64_t)B * D, n2 = (int64_t)B * T * D;
  prefix_time_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(S, ldS, B, T, D, not_interval, P);
  count_launch();
  SMX_TRY(check_launch("prefix_time_kernel"));
  interval_mean_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(P, lo, hi, B, T, D, not_interval, out, ldo, 1);
  count_launch();
  return check_launch("interval_mean_kernel");
}
size_t interval_means_workspace_bytes(int B, int T, int D) { return align_up((size_t)B * (T + 1) * D * sizeof(double)); }
int interval_means(const float* S, int64_t ldS, int B, int T, int D, const int* lo, const int* hi, const int* not_interval, float* Sm,
                   void* workspace, cudaStream_t st) {
  double* P = (double*)workspace;
  const int64_t n1 = (int64_t)B * D, n2 = (int64_t)B * T * D;
  prefix_time_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(S, ldS, B, T, D, not_interval, P);
  count_launch();
  SMX_TRY(check_launch("prefix_time_kernel"));
  interval_mean_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(P, lo, hi, B, T, D, not_interval, Sm);
  count_launch();
  return check_launch("interval_mean_kernel");
}

__global__ void rowsum_kernel(const float* m, int rows, int cols, float* out) {
  int row = blockIdx.x * 8 + threadIdx.x / 32;
  int lane = threadIdx.x % 32;
  if (row >= rows) return;
  float s = 0.0f;
  for (int c = lane; c < cols; c += 32) s += m[(int64_t)row * cols + c];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}
int rowsum(const float* m, int rows, int cols, float* out, cudaStream_t st) {
  rowsum_kernel<<<(rows + 7) / 8, 256, 0, st>>>(m, rows, cols, out);
  count_launch();
  return check_launch("rowsum_kernel");
}

__global__ void convert_kernel(const void* src, int s_dt, void* dst, int d_dt, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) st_any(dst, d_dt, i, ld_any(src, s_dt, i));
}
int convert(const void* src, int s_dtype, void* dst, int d_dtype, int64_t n, cudaStream_t st) {
  if (n <= 0) return SMX_OK;
  convert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, s_dtype, dst, d_dtype, n);
  count_launch();
  return check_launch("convert_kernel");
}

// mask builders ------------------------------------------------------------------------------
__global__ void padding_mask_kernel(const float* wav_len, int B, int T, uint8_t* mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * T) return;
  int b = (int)(i / T), t = (int)(i % T);
  float abs_len = rintf(wav_len[b] * (float)T);  // torch.round: half to even
  mask[i] = (float)t < abs_len ? 1 : 0;
}
__global__ void chunk_mask_kernel(int T, int chunk, int left, float* mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T * T) return;
  int r = (int)(i / T), c = (int)(i % T);
  int end = (r / chunk + 1) * chunk;
  bool vis = c < end;
  if (left >= 0) vis = vis && (c >= end - chunk * (left + 1));
  mask[i] = vis ? 1.0f : 0.0f;
}

}  // namespace smx

extern "C" int smx_padding_mask_from_wav_len(const float* wav_len, int32_t B, int32_t T, uint8_t* mask, void* stream) {
  if (!wav_len || !mask || B <= 0 || T <= 0) return smx::fail(SMX_ERR_BAD_ARG, "padding_mask: bad argument");
  int64_t n = (int64_t)B * T;
  smx::padding_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(wav_len, B, T, mask);
  smx::count_launch();
  return smx::check_launch("padding_mask_kernel");
}

extern "C" int smx_chunk_mask(int32_t T, int32_t chunk_size, int32_t left_context_chunks, float* mask, void* stream) {
  if (!mask || T <= 0 || chunk_size <= 0) return smx::fail(SMX_ERR_BAD_ARG, "chunk_mask: bad argument");
  int64_t n = (int64_t)T * T;
  smx::chunk_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, chunk_size,
                                                                                       left_context_chunks, mask);
  smx::count_launch();
  return smx::check_launch("chunk_mask_kernel");
}
