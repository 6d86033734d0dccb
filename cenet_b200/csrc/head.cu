// Segmentation head tail (out.py:74 + metrics_eval.py:52) and the fused Dice+CE loss (utils/core.py:57-80,176-188).
#include "common.cuh"
#include <algorithm>

namespace {
constexpr int kMaxCls = 16;

// one thread per output pixel: bilinear x2 (align_corners=False, scale_factor=2 -> source scale 0.5) of every class,
// logits written NCHW (coalesced along w for each class), softmax -> argmax with lowest-index tie-break.
template <int NC>
__global__ void __launch_bounds__(256) head_upsample_argmax_kernel(const float* __restrict__ y, float* __restrict__ logits,
                                                                   long long* __restrict__ labels, int B, int h, int w,
                                                                   int ncls_rt) {
  const int ncls = NC > 0 ? NC : ncls_rt;
  const int Ho = 2 * h, Wo = 2 * w;
  const long long total = (long long)B * Ho * Wo;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wo = (int)(idx % Wo), ho = (int)((idx / Wo) % Ho);
  const long long b = idx / ((long long)Ho * Wo);
  int h0, h1, w0, w1;
  float lh, lw;
  bilin_src(ho, 0.5f, h, h0, h1, lh);
  bilin_src(wo, 0.5f, w, w0, w1, lw);
  const float* p00 = y + ((b * h + h0) * w + w0) * ncls;
  const float* p01 = y + ((b * h + h0) * w + w1) * ncls;
  const float* p10 = y + ((b * h + h1) * w + w0) * ncls;
  const float* p11 = y + ((b * h + h1) * w + w1) * ncls;
  float v[NC > 0 ? NC : kMaxCls];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < (NC > 0 ? NC : kMaxCls); c++) {
    if (c < ncls) {
      v[c] = (1.f - lh) * ((1.f - lw) * p00[c] + lw * p01[c]) + lh * ((1.f - lw) * p10[c] + lw * p11[c]);
      mx = fmaxf(mx, v[c]);
      if (logits) logits[((b * ncls + c) * Ho + ho) * Wo + wo] = v[c];
    }
  }
  if (labels) {
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < (NC > 0 ? NC : kMaxCls); c++)
      if (c < ncls) { v[c] = expf(v[c] - mx); sum += v[c]; }
    int best = 0;
    float pb = -1.f;
#pragma unroll
    for (int c = 0; c < (NC > 0 ? NC : kMaxCls); c++)
      if (c < ncls) { const float p = v[c] / sum; if (p > pb) { pb = p; best = c; } }
    labels[idx] = best;
  }
}

// ---- Dice + CE ---------------------------------------------------------------------------------------------------
constexpr int kLossThreads = 256;
constexpr int kLossPixPerBlock = 4096;

// pass 1: per-block partial sums [I_i, Z_i, Y_i]*ncls + CE  (fixed order -> deterministic)
__global__ void __launch_bounds__(kLossThreads) dice_ce_partial_kernel(const float* __restrict__ logits,
                                                                       const long long* __restrict__ labels,
                                                                       float* __restrict__ part, int ncls, int HW,
                                                                       long long npix) {
  __shared__ float red[kLossThreads / 32][3 * kMaxCls + 1];
  float acc[3 * kMaxCls + 1];
#pragma unroll
  for (int i = 0; i < 3 * kMaxCls + 1; i++) acc[i] = 0.f;
  const long long p0 = (long long)blockIdx.x * kLossPixPerBlock;
  const long long p1 = p0 + kLossPixPerBlock < npix ? p0 + kLossPixPerBlock : npix;
  for (long long p = p0 + threadIdx.x; p < p1; p += kLossThreads) {
    const long long b = p / HW, q = p % HW;
    const float* lp = logits + b * ncls * HW + q;
    float v[kMaxCls];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < kMaxCls; c++) if (c < ncls) { v[c] = lp[(long long)c * HW]; mx = fmaxf(mx, v[c]); }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxCls; c++) if (c < ncls) { v[c] = expf(v[c] - mx); sum += v[c]; }
    const int t = (int)labels[p];
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < kMaxCls; c++) {
      if (c < ncls) {
        const float pc = v[c] * inv;
        const float tc = (c == t) ? 1.f : 0.f;
        acc[3 * c + 0] += pc * tc;
        acc[3 * c + 1] += pc * pc;
        acc[3 * c + 2] += tc;
        if (c == t) acc[3 * kMaxCls] += -logf(fmaxf(pc, 1e-45f));
      }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 3 * kMaxCls + 1; i++) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) red[wid][i] = s;
  }
  __syncthreads();
  const int nvals = 3 * ncls + 1;
  if (threadIdx.x < nvals) {
    const int src = threadIdx.x < 3 * ncls ? threadIdx.x : 3 * kMaxCls;
    float s = 0.f;
    for (int k = 0; k < kLossThreads / 32; k++) s += red[k][src];
    part[(long long)blockIdx.x * nvals + threadIdx.x] = s;
  }
}

// pass 2: one block sums the partials in block order and writes totals + the loss
__global__ void __launch_bounds__(64) dice_ce_finalize_kernel(const float* __restrict__ part, float* __restrict__ tot,
                                                              float* __restrict__ loss_out, int nblk, int ncls,
                                                              long long npix, float w_dice, float w_ce) {
  __shared__ float s_tot[3 * kMaxCls + 1];
  const int nvals = 3 * ncls + 1;
  if (threadIdx.x < nvals) {
    double s = 0.0;
    for (int k = 0; k < nblk; k++) s += (double)part[(long long)k * nvals + threadIdx.x];
    s_tot[threadIdx.x] = (float)s;
    tot[threadIdx.x] = (float)s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float dice = 0.f;
    for (int c = 0; c < ncls; c++) {
      const float I = s_tot[3 * c], Z = s_tot[3 * c + 1], Y = s_tot[3 * c + 2];
      const float sc = (2.f * I + 1e-5f) / (Z + Y + 1e-5f);
      dice += 1.f - sc;
      loss_out[1 + c] = sc;
    }
    dice /= (float)ncls;
    const float ce = s_tot[3 * ncls] / (float)npix;
    loss_out[0] = w_dice * dice + w_ce * ce;
    tot[nvals] = dice;
    tot[nvals + 1] = ce;
  }
}

// pass 3: dL/dlogits
__global__ void __launch_bounds__(256) dice_ce_grad_kernel(const float* __restrict__ logits,
                                                           const long long* __restrict__ labels,
                                                           const float* __restrict__ tot, float* __restrict__ dlogits,
                                                           int ncls, int HW, long long npix, float w_dice, float w_ce,
                                                           float grad_scale) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const long long b = p / HW, q = p % HW;
  const float* lp = logits + b * ncls * HW + q;
  float v[kMaxCls], g[kMaxCls];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < kMaxCls; c++) if (c < ncls) { v[c] = lp[(long long)c * HW]; mx = fmaxf(mx, v[c]); }
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxCls; c++) if (c < ncls) { v[c] = expf(v[c] - mx); sum += v[c]; }
  const int t = (int)labels[p];
  const float inv = 1.f / sum;
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxCls; c++) {
    if (c < ncls) {
      v[c] *= inv;
      const float I = tot[3 * c], D = tot[3 * c + 1] + tot[3 * c + 2] + 1e-5f;
      const float tc = (c == t) ? 1.f : 0.f;
      g[c] = -w_dice / (float)ncls * (2.f * tc * D - (2.f * I + 1e-5f) * 2.f * v[c]) / (D * D);
      dot += g[c] * v[c];
    }
  }
  float* dp = dlogits + b * ncls * HW + q;
  const float cw = w_ce / (float)npix;
#pragma unroll
  for (int c = 0; c < kMaxCls; c++) {
    if (c < ncls) {
      const float tc = (c == t) ? 1.f : 0.f;
      dp[(long long)c * HW] = grad_scale * (v[c] * (g[c] - dot) + cw * (v[c] - tc));
    }
  }
}
}  // namespace

extern "C" int cenet_head_upsample_argmax(const float* y, float* logits_nchw, long long* labels, int B, int h, int w,
                                          int ncls, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(y && (logits_nchw || labels), "cenet_head_upsample_argmax: null pointer");
  CENET_REQUIRE(ncls >= 1 && ncls <= kMaxCls, "cenet_head_upsample_argmax: 1..%d classes supported, got %d", kMaxCls, ncls);
  const long long total = (long long)B * 4 * h * w;
  const int grid = cdiv(total, 256);
  if (ncls == 9) head_upsample_argmax_kernel<9><<<grid, 256, 0, to_stream(s)>>>(y, logits_nchw, labels, B, h, w, ncls);
  else if (ncls == 4) head_upsample_argmax_kernel<4><<<grid, 256, 0, to_stream(s)>>>(y, logits_nchw, labels, B, h, w, ncls);
  else if (ncls == 2) head_upsample_argmax_kernel<2><<<grid, 256, 0, to_stream(s)>>>(y, logits_nchw, labels, B, h, w, ncls);
  else head_upsample_argmax_kernel<0><<<grid, 256, 0, to_stream(s)>>>(y, logits_nchw, labels, B, h, w, ncls);
  CENET_LAUNCH_CHECK("head_upsample_argmax");
  return 0;
}

extern "C" int cenet_loss_nblocks(long long npix) { return (int)((npix + kLossPixPerBlock - 1) / kLossPixPerBlock); }

extern "C" int cenet_dice_ce(const float* logits, const long long* labels, float* loss_out, float* dlogits, float* ws,
                             int B, int ncls, int HW, float w_dice, float w_ce, float grad_scale, cenet_stream_t s) {
  CENET_REQUIRE(logits && labels && loss_out && ws, "cenet_dice_ce: null pointer");
  CENET_REQUIRE(B >= 1 && ncls >= 1 && ncls <= kMaxCls, "cenet_dice_ce: 1..%d classes supported, got %d", kMaxCls, ncls);
  const long long npix = (long long)B * HW;
  const int nblk = cenet_loss_nblocks(npix);
  const int nvals = 3 * ncls + 1;
  float* part = ws;
  float* tot = ws + (long long)nblk * nvals;
  dice_ce_partial_kernel<<<nblk, kLossThreads, 0, to_stream(s)>>>(logits, labels, part, ncls, HW, npix);
  CENET_LAUNCH_CHECK("dice_ce_partial");
  dice_ce_finalize_kernel<<<1, 64, 0, to_stream(s)>>>(part, tot, loss_out, nblk, ncls, npix, w_dice, w_ce);
  CENET_LAUNCH_CHECK("dice_ce_finalize");
  if (dlogits) {
    dice_ce_grad_kernel<<<cdiv(npix, 256), 256, 0, to_stream(s)>>>(logits, labels, tot, dlogits, ncls, HW, npix, w_dice,
                                                                   w_ce, grad_scale);
    CENET_LAUNCH_CHECK("dice_ce_grad");
  }
  return 0;
}
