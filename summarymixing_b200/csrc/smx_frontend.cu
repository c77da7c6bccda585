// libsmx frontend: the acoustic front of the encoder as the recipes configure it (conformer_summarymixing.yaml:145-152,
// 198-200, 298-330; TransformerASR.py:353-358, 405-406; Transformer.py:288-339).  All kernels are HBM-bound streaming work on
// CUDA cores (fp32 arithmetic): coalesced 128-bit accesses, shared-memory staging, grids sized from the data.
//
//   K-FBANK   waveform -> hamming-windowed frames -> 512-point FFT in shared memory -> power -> triangular mel filters -> dB
//             (one CTA per group of frames; the (T', 80) features are written once; the per-utterance top_db clamp is a second,
//             tiny pass because it needs the utterance maximum)
//   K-NORM    (x - mean[f]) / std[f]                     InputNormalization(global), inference
//   K-DROP    SpectrogramDrop: spans along time / frequency replaced by zero or the batch mean (positions drawn by the caller)
//   K-WARP    Warping: bicubic resampling of the two segments around a warp centre (align_corners = True, A = -0.75)
//   K-CNN     ConvolutionFrontEnd block: Conv2d(k=3, stride 2, reflect 'same') + LayerNorm over (F', C) + LeakyReLU, one CTA
//             per output frame
//   K-POSENC  y = x W^T + b + sinusoidal positional encoding (the GEMM itself is the generic / tensor-core linear)
#include "smx_internal.h"
#include "smx_tc.h"

namespace smx {

// =============================================================================================
// K-FBANK
// =============================================================================================
constexpr int FB_NFFT = 512, FB_BINS = 257;

// one CTA = 256 threads works through frames [f0, f1) of one utterance: radix-2 DIT FFT of the windowed frame in shared memory
__global__ void __launch_bounds__(256) fbank_kernel(const float* __restrict__ wav, int n_samples, int hop, int win, int n_frames,
                                                    int frames_per_cta, const float* __restrict__ melw, int n_mels, float amin,
                                                    float* __restrict__ out) {
  __shared__ float2 buf[FB_NFFT];
  __shared__ float2 tw[FB_NFFT / 2];
  __shared__ float pw[FB_BINS + 7];
  __shared__ float window[FB_NFFT];
  const int b = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < FB_NFFT / 2; i += 256) {
    float s, c;
    sincospif(-2.0f * (float)i / (float)FB_NFFT, &s, &c);
    tw[i] = make_float2(c, s);
  }
  for (int i = tid; i < FB_NFFT; i += 256)  // torch.hamming_window(win) (periodic), centred in the n_fft frame
    window[i] = (i < win) ? 0.54f - 0.46f * cospif(2.0f * (float)i / (float)win) : 0.0f;
  __syncthreads();
  const float* w = wav + (size_t)b * n_samples;
  const int f0 = blockIdx.x * frames_per_cta, f1 = min(n_frames, f0 + frames_per_cta);
  const int woff = (FB_NFFT - win) / 2;
  for (int f = f0; f < f1; ++f) {
    // frame f covers samples [f hop - n_fft/2, f hop + n_fft/2) (center=True, zero padding); bit-reversed load
    for (int i = tid; i < FB_NFFT; i += 256) {
      const int n = f * hop - FB_NFFT / 2 + i;
      const int wi = i - woff;
      float v = (n >= 0 && n < n_samples && wi >= 0 && wi < win) ? w[n] * window[wi] : 0.0f;
      buf[__brev((unsigned)i) >> 23] = make_float2(v, 0.0f);
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 1; s < FB_NFFT; s <<= 1) {   // 9 stages, 256 butterflies each
      const int j = tid & (s - 1), base = ((tid - j) << 1) + j;
      const float2 t = tw[j * (FB_NFFT / 2 / s)];
      const float2 a = buf[base], c = buf[base + s];
      const float2 m = make_float2(c.x * t.x - c.y * t.y, c.x * t.y + c.y * t.x);
      buf[base] = make_float2(a.x + m.x, a.y + m.y);
      buf[base + s] = make_float2(a.x - m.x, a.y - m.y);
      __syncthreads();
    }
    for (int i = tid; i < FB_BINS; i += 256) pw[i] = buf[i].x * buf[i].x + buf[i].y * buf[i].y;
    __syncthreads();
    if (tid < n_mels) {  // mel filter tid: dense (257 x n_mels) weights, mostly zeros (triangles): 257 FMAs per output
      float acc = 0.0f;
      for (int k = 0; k < FB_BINS; ++k) acc = fmaf(pw[k], melw[(size_t)k * n_mels + tid], acc);
      out[((size_t)b * n_frames + f) * n_mels + tid] = 10.0f * log10f(fmaxf(acc, amin));
    }
    __syncthreads();
  }
}
// per-utterance maximum (stage 1: per-CTA partials; stage 2 folded into the clamp kernel)
__global__ void __launch_bounds__(256) utt_max_kernel(const float* __restrict__ x, int64_t per_utt, float* __restrict__ umax) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  float m = -INFINITY;
  for (int64_t i = threadIdx.x; i < per_utt; i += 256) m = fmaxf(m, x[(size_t)b * per_utt + i]);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    umax[b] = m;
  }
}
__global__ void __launch_bounds__(256) top_db_kernel(float* __restrict__ x, int64_t per_utt, int64_t n, const float* __restrict__ umax, float top_db) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  x[i] = fmaxf(x[i], umax[i / per_utt] - top_db);
}
// (n_fft/2+1, n_mels) triangular mel weights, SpeechBrain Filterbank (symmetric triangles, slope from the band left of the centre)
__global__ void mel_weights_kernel(int n_mels, int sample_rate, float f_min, float f_max, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= FB_BINS * n_mels) return;
  const int k = i / n_mels, m = i % n_mels;
  const double mel_lo = 2595.0 * log10(1.0 + (double)f_min / 700.0), mel_hi = 2595.0 * log10(1.0 + (double)f_max / 700.0);
  auto hz = [&](int j) { return 700.0 * (pow(10.0, (mel_lo + (mel_hi - mel_lo) * (double)j / (double)(n_mels + 1)) / 2595.0) - 1.0); };
  const double fc = hz(m + 1), band = hz(m + 1) - hz(m);
  const double fr = (double)(sample_rate / 2) * (double)k / (double)(FB_BINS - 1);
  const double slope = (fr - fc) / band;
  out[i] = (float)fmax(0.0, fmin(slope + 1.0, -slope + 1.0));
}

int fbank_frames(int n_samples, int hop) { return 1 + n_samples / hop; }
size_t fbank_workspace_bytes(int B, int n_mels) { return align_up((size_t)FB_BINS * n_mels * 4) + align_up((size_t)B * 4); }
int fbank_fwd(const smx_fbank_desc* d, int B, int n_samples, const float* wav, float* feats, void* workspace, cudaStream_t st) {
  if (d->n_fft != FB_NFFT) return fail(SMX_ERR_UNSUPPORTED, "fbank: n_fft=%d (512 is implemented)", d->n_fft);
  const int win = (int)lrintf((float)d->sample_rate / 1000.0f * d->win_length_ms), hop = (int)lrintf((float)d->sample_rate / 1000.0f * d->hop_length_ms);
  if (win < 1 || win > FB_NFFT || hop < 1 || d->n_mels < 1 || d->n_mels > 256) return fail(SMX_ERR_BAD_ARG, "fbank: win=%d hop=%d n_mels=%d", win, hop, d->n_mels);
  const int nf = fbank_frames(n_samples, hop);
  float* melw = (float*)workspace;
  float* umax = (float*)((char*)workspace + align_up((size_t)FB_BINS * d->n_mels * 4));
  const float f_max = d->f_max > 0 ? d->f_max : (float)d->sample_rate / 2;
  mel_weights_kernel<<<(FB_BINS * d->n_mels + 255) / 256, 256, 0, st>>>(d->n_mels, d->sample_rate, d->f_min, f_max, melw);
  count_launch();
  SMX_TRY(check_launch("mel_weights_kernel"));
  const int fpc = 16;
  dim3 grid((nf + fpc - 1) / fpc, B);
  fbank_kernel<<<grid, 256, 0, st>>>(wav, n_samples, hop, win, nf, fpc, melw, d->n_mels, d->amin, feats);
  count_launch();
  SMX_TRY(check_launch("fbank_kernel"));
  if (d->top_db > 0) {
    const int64_t per = (int64_t)nf * d->n_mels, n = per * B;
    utt_max_kernel<<<B, 256, 0, st>>>(feats, per, umax);
    count_launch();
    SMX_TRY(check_launch("utt_max_kernel"));
    top_db_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(feats, per, n, umax, d->top_db);
    count_launch();
    SMX_TRY(check_launch("top_db_kernel"));
  }
  return SMX_OK;
}

// =============================================================================================
// K-NORM, K-DROP, K-WARP
// =============================================================================================
__global__ void __launch_bounds__(256) input_norm_kernel(const float* __restrict__ x, int64_t n, int F, const float* __restrict__ mean,
                                                         const float* __restrict__ stdv, float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int f = (int)(i % F);
  y[i] = (x[i] - mean[f]) / stdv[f];
}
int input_norm_fwd(int64_t rows, int F, const float* x, const float* mean, const float* stdv, float* y, cudaStream_t st) {
  const int64_t n = rows * F;
  input_norm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, F, mean, stdv, y);
  count_launch();
  return check_launch("input_norm_kernel");
}

// mean of the whole tensor: fixed-order two-stage reduction (deterministic)
__global__ void __launch_bounds__(256) sum_partial_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ part) {
  __shared__ double red[8];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) s += (double)x[i];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) s += red[i];
    part[blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(256) spec_drop_kernel(float* __restrict__ x, int B, int T, int F, int dim, int n_masks, const int* __restrict__ pos,
                                                        const int* __restrict__ len, const double* __restrict__ part, int n_part, int replace_mean) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)B * T * F;
  if (i >= n) return;
  const int f = (int)(i % F), t = (int)((i / F) % T), b = (int)(i / ((int64_t)F * T));
  const int c = dim == 1 ? t : f;
  bool hit = false;
  for (int m = 0; m < n_masks; ++m) {
    const int p = pos[b * n_masks + m], l = len[b * n_masks + m];
    hit |= (c >= p && c < p + l);
  }
  if (!hit) return;
  float val = 0.0f;
  if (replace_mean) {
    double s = 0.0;
    for (int k = 0; k < n_part; ++k) s += part[k];   // fixed order
    val = (float)(s / (double)n);
  }
  x[i] = val;
}
size_t spec_drop_workspace_bytes() { return 256 * sizeof(double); }
int spec_drop_fwd(int B, int T, int F, float* x, int dim, int n_masks, const int* pos, const int* len, int replace_mean, void* workspace,
                  cudaStream_t st) {
  if (dim != 1 && dim != 2) return fail(SMX_ERR_BAD_ARG, "spec_drop: dim must be 1 (time) or 2 (frequency)");
  const int64_t n = (int64_t)B * T * F;
  double* part = (double*)workspace;
  const int n_part = 128;
  if (replace_mean) {
    sum_partial_kernel<<<n_part, 256, 0, st>>>(x, n, part);
    count_launch();
    SMX_TRY(check_launch("sum_partial_kernel"));
  }
  spec_drop_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, B, T, F, dim, n_masks, pos, len, part, n_part, replace_mean);
  count_launch();
  return check_launch("spec_drop_kernel");
}

// cubic convolution weights, A = -0.75 (torch bicubic)
__device__ __forceinline__ void cubic_w(float t, float* w) {
  const float A = -0.75f;
  const float x0 = t + 1.0f, x1 = t, x2 = 1.0f - t, x3 = 2.0f - t;
  w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  w[1] = ((A + 2.0f) * x1 - (A + 3.0f)) * x1 * x1 + 1.0f;
  w[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}
// y[b, t, :]: t < w from x[b, 0:c] resampled to w frames; t >= w from x[b, c:T] resampled to T - w frames (align_corners)
__global__ void __launch_bounds__(256) time_warp_kernel(const float* __restrict__ x, int B, int T, int F, int c, int w, float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)B * T * F;
  if (i >= n) return;
  const int f = (int)(i % F), t = (int)((i / F) % T), b = (int)(i / ((int64_t)F * T));
  int in0, in_len, out_len, to;
  if (t < w) { in0 = 0; in_len = c; out_len = w; to = t; }
  else { in0 = c; in_len = T - c; out_len = T - w; to = t - w; }
  const float scale = out_len > 1 ? (float)(in_len - 1) / (float)(out_len - 1) : 0.0f;
  const float src = scale * (float)to;
  const int i0 = (int)floorf(src);
  float cw[4];
  cubic_w(src - (float)i0, cw);
  const float* xb = x + ((size_t)b * T + in0) * F + f;
  float acc = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int j = i0 - 1 + k;
    j = j < 0 ? 0 : (j > in_len - 1 ? in_len - 1 : j);
    acc = fmaf(cw[k], xb[(size_t)j * F], acc);
  }
  y[i] = acc;
}
int time_warp_fwd(int B, int T, int F, const float* x, int c, int w, float* y, cudaStream_t st) {
  if (c < 1 || c >= T || w < 1 || w >= T) return fail(SMX_ERR_BAD_ARG, "time_warp: centre %d / new position %d outside (0, %d)", c, w, T);
  const int64_t n = (int64_t)B * T * F;
  time_warp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, B, T, F, c, w, y);
  count_launch();
  return check_launch("time_warp_kernel");
}

// =============================================================================================
// K-CNN: one ConvolutionFrontEnd block.  x (B, T, F, Cin) channels-last -> y (B, T', F', Cout), T' = ceil(T / s), F' = ceil(F / s)
// One CTA per output frame: the k input frames it needs (reflect padding in time and frequency) are staged in shared memory,
// every thread produces outputs (f', co) in a strided loop, LayerNorm statistics over all F' * Cout outputs of the frame are
// reduced in the block (two-pass, fixed order), then LeakyReLU.
// =============================================================================================
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
__global__ void __launch_bounds__(256) conv_block_kernel(const float* __restrict__ x, int T, int F, int Cin, int To, int Fo, int Cout, int ks, int stride,
                                                         const float* __restrict__ cw, const float* __restrict__ cb, const float* __restrict__ lw,
                                                         const float* __restrict__ lb, float slope, float* __restrict__ y) {
  extern __shared__ float sm[];
  float* sIn = sm;                             // [ks][F][Cin]
  float* sOut = sm + (size_t)ks * F * Cin;     // [Fo][Cout]
  __shared__ float red[8];
  __shared__ float stat[2];
  const int b = blockIdx.y, to = blockIdx.x, tid = threadIdx.x;
  const int pad = ks / 2;
  for (int i = tid; i < ks * F * Cin; i += 256) {
    const int kt = i / (F * Cin), rem = i % (F * Cin);
    const int t = reflect_idx(to * stride - pad + kt, T);
    sIn[i] = x[((size_t)b * T + t) * F * Cin + rem];
  }
  __syncthreads();
  const int n_out = Fo * Cout;
  float lsum = 0.0f;
  for (int o = tid; o < n_out; o += 256) {
    const int fo = o / Cout, co = o % Cout;
    float acc = cb ? cb[co] : 0.0f;
    for (int kt = 0; kt < ks; ++kt)
      for (int kf = 0; kf < ks; ++kf) {
        const int fi = reflect_idx(fo * stride - pad + kf, F);
        const float* in = sIn + ((size_t)kt * F + fi) * Cin;
        const float* wv = cw + (((size_t)co * Cin) * ks + kt) * ks + kf;   // weight (Cout, Cin, kt, kf)
        for (int ci = 0; ci < Cin; ++ci) acc = fmaf(in[ci], wv[(size_t)ci * ks * ks], acc);
      }
    sOut[o] = acc;
    lsum += acc;
  }
  // LayerNorm over the frame's F' * Cout values
#pragma unroll
  for (int o = 16; o; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if ((tid & 31) == 0) red[tid >> 5] = lsum;
  __syncthreads();
  if (tid == 0) { float s = 0.0f; for (int i = 0; i < 8; ++i) s += red[i]; stat[0] = s / (float)n_out; }
  __syncthreads();
  const float mean = stat[0];
  float lq = 0.0f;
  for (int o = tid; o < n_out; o += 256) { const float d = sOut[o] - mean; lq = fmaf(d, d, lq); }
#pragma unroll
  for (int o = 16; o; o >>= 1) lq += __shfl_xor_sync(0xffffffffu, lq, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = lq;
  __syncthreads();
  if (tid == 0) { float s = 0.0f; for (int i = 0; i < 8; ++i) s += red[i]; stat[1] = rsqrtf(s / (float)n_out + 1e-5f); }
  __syncthreads();
  const float rstd = stat[1];
  float* yo = y + ((size_t)b * To + to) * n_out;
  for (int o = tid; o < n_out; o += 256) {
    const float v = (sOut[o] - mean) * rstd * lw[o] + lb[o];
    yo[o] = v >= 0.0f ? v : slope * v;
  }
}
int conv_block_fwd(int B, int T, int F, int Cin, int Cout, int ks, int stride, const float* x, const float* cw, const float* cb, const float* lw,
                   const float* lb, float* y, cudaStream_t st) {
  if (ks % 2 == 0 || ks / 2 >= T || ks / 2 >= F) return fail(SMX_ERR_BAD_ARG, "conv block: kernel %d vs T=%d F=%d", ks, T, F);
  const int To = (T + stride - 1) / stride, Fo = (F + stride - 1) / stride;
  const size_t smem = ((size_t)ks * F * Cin + (size_t)Fo * Cout) * sizeof(float);
  if (smem > 200 * 1024) return fail(SMX_ERR_UNSUPPORTED, "conv block: frame does not fit shared memory");
  cudaError_t e = cudaFuncSetAttribute(conv_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(SMX_ERR_CUDA, "cudaFuncSetAttribute(conv_block_kernel): %s", cudaGetErrorString(e));
  dim3 grid(To, B);
  conv_block_kernel<<<grid, 256, smem, st>>>(x, T, F, Cin, To, Fo, Cout, ks, stride, cw, cb, lw, lb, 0.01f, y);
  count_launch();
  return check_launch("conv_block_kernel");
}

// =============================================================================================
// K-POSENC: y[b, t, d] = v[b, t, d] + pe[t, d], pe = sin / cos table of Transformer.py:288-339 (computed in place, fp32 math as the
// reference: exp(2i * -(ln 10000 / D)), sin / cos of position * that)
// =============================================================================================
__global__ void __launch_bounds__(256) posenc_kernel(const float* __restrict__ v, int64_t n, int T, int D, void* __restrict__ y, int y_dt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = (int)(i % D), t = (int)((i / D) % T);
  const float den = expf((float)(d & ~1) * (-(logf(10000.0f) / (float)D)));
  const float a = (float)t * den;
  const float pe = (d & 1) ? cosf(a) : sinf(a);
  const float r = v[i] + pe;
  if (y_dt == SMX_BF16) ((__nv_bfloat16*)y)[i] = __float2bfloat16(r);
  else ((float*)y)[i] = r;
}
int posenc_add(const float* v, int B, int T, int D, void* y, int y_dt, cudaStream_t st) {
  const int64_t n = (int64_t)B * T * D;
  posenc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, n, T, D, y, y_dt);
  count_launch();
  return check_launch("posenc_kernel");
}

}  // namespace smx
