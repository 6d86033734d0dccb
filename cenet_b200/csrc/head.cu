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
                                                                   int ncls_rt, int ldy) {
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
  const float* p00 = y + ((b * h + h0) * w + w0) * ldy;
  const float* p01 = y + ((b * h + h0) * w + w1) * ldy;
  const float* p10 = y + ((b * h + h1) * w + w0) * ldy;
  const float* p11 = y + ((b * h + h1) * w + w1) * ldy;
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

// ---- Dice + CE + Boundary-DoU ------------------------------------------------------------------------------------
// One reduction pass, one finalize block, one element-wise gradient pass for Criterion's weighted sum (core.py:179-188)
// of DiceLoss (57-80), CrossEntropyLoss (176) and BoundaryDoULoss (83-131).  Per class c the partial sums are
//   I_c = sum p_c t_c, Z_c = sum p_c^2, Y_c = sum t_c            (nacc == 3: Dice / CE only)
//   C_c = #{pixels of class c with a 4-neighbour of another class or outside the image}   (nacc == 4: + Boundary-DoU;
//         = count_nonzero of the reference's cross-kernel conv of the one-hot map with the 5s zeroed, core.py:99-107)
// Boundary-DoU: alpha_c = min(2 (1 - (C_c + s)/(Y_c + s)) - 1, 0.8);  loss_c = (Z+Y-2I+s) / (Z+Y-(1+alpha_c) I+s), s = 1e-5.
// Counts are sums of exact small integers in fp32 per block and double across blocks -> bit-exact integers.
constexpr int kLossThreads = 256;
constexpr int kLossPixPerBlock = 4096;
constexpr int kAccMax = 4 * kMaxCls + 1;

__global__ void __launch_bounds__(kLossThreads) seg_loss_partial_kernel(const float* __restrict__ logits,
                                                                        const long long* __restrict__ labels,
                                                                        float* __restrict__ part, int ncls, int nacc, int H,
                                                                        int W, long long npix) {
  __shared__ float red[kLossThreads / 32][kAccMax];
  float acc[kAccMax];
#pragma unroll
  for (int i = 0; i < kAccMax; i++) acc[i] = 0.f;
  const int HW = H * W;
  const long long p0 = (long long)blockIdx.x * kLossPixPerBlock;
  const long long p1 = p0 + kLossPixPerBlock < npix ? p0 + kLossPixPerBlock : npix;
  for (long long p = p0 + threadIdx.x; p < p1; p += kLossThreads) {
    const long long b = p / HW;
    const int q = (int)(p % HW);
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
    float edge = 0.f;
    if (nacc == 4) {
      const int h = q / W, w = q % W;
      const long long* lb = labels + b * HW;
      const bool inner = h > 0 && h < H - 1 && w > 0 && w < W - 1 && (int)lb[q - W] == t && (int)lb[q + W] == t &&
                         (int)lb[q - 1] == t && (int)lb[q + 1] == t;
      edge = inner ? 0.f : 1.f;
    }
#pragma unroll
    for (int c = 0; c < kMaxCls; c++) {
      if (c < ncls) {
        const float pc = v[c] * inv;
        const float tc = (c == t) ? 1.f : 0.f;
        acc[nacc * c + 0] += pc * tc;
        acc[nacc * c + 1] += pc * pc;
        acc[nacc * c + 2] += tc;
        if (nacc == 4) acc[nacc * c + 3] += tc * edge;
        if (c == t) acc[kAccMax - 1] += -logf(fmaxf(pc, 1e-45f));
      }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kAccMax; i++) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) red[wid][i] = s;
  }
  __syncthreads();
  const int nvals = nacc * ncls + 1;
  if (threadIdx.x < nvals) {
    const int src = threadIdx.x < nacc * ncls ? threadIdx.x : kAccMax - 1;
    float s = 0.f;
    for (int k = 0; k < kLossThreads / 32; k++) s += red[k][src];
    part[(long long)blockIdx.x * nvals + threadIdx.x] = s;
  }
}

// pass 2: one block sums the partials in block order and writes totals + the loss
//   tot: [nvals sums | dice | ce | boundary | alpha_c (ncls, nacc == 4 only)]
__global__ void __launch_bounds__(128) seg_loss_finalize_kernel(const float* __restrict__ part, float* __restrict__ tot,
                                                                float* __restrict__ loss_out, int nblk, int ncls, int nacc,
                                                                long long npix, float w_dice, float w_ce, float w_bd) {
  __shared__ float s_tot[kAccMax];
  const int nvals = nacc * ncls + 1;
  if (threadIdx.x < nvals) {
    double s = 0.0;
    for (int k = 0; k < nblk; k++) s += (double)part[(long long)k * nvals + threadIdx.x];
    s_tot[threadIdx.x] = (float)s;
    tot[threadIdx.x] = (float)s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float dice = 0.f, bd = 0.f;
    const float sm = 1e-5f;
    for (int c = 0; c < ncls; c++) {
      const float I = s_tot[nacc * c], Z = s_tot[nacc * c + 1], Y = s_tot[nacc * c + 2];
      const float sc = (2.f * I + sm) / (Z + Y + sm);
      dice += 1.f - sc;
      loss_out[1 + c] = sc;
      if (nacc == 4) {
        const float Cb = s_tot[nacc * c + 3];
        float alpha = 2.f * (1.f - (Cb + sm) / (Y + sm)) - 1.f;
        alpha = fminf(alpha, 0.8f);
        bd += (Z + Y - 2.f * I + sm) / (Z + Y - (1.f + alpha) * I + sm);
        tot[nvals + 3 + c] = alpha;
      }
    }
    dice /= (float)ncls;
    bd /= (float)ncls;
    const float ce = s_tot[nacc * ncls] / (float)npix;
    loss_out[0] = w_dice * dice + w_ce * ce + w_bd * bd;
    tot[nvals] = dice;
    tot[nvals + 1] = ce;
    if (nacc == 4) tot[nvals + 2] = bd;              // (the 3-sum layout of cenet_dice_ce ends at tot[nvals + 1])
  }
}

// pass 3: dL/dlogits
__global__ void __launch_bounds__(256) seg_loss_grad_kernel(const float* __restrict__ logits,
                                                            const long long* __restrict__ labels,
                                                            const float* __restrict__ tot, float* __restrict__ dlogits,
                                                            int ncls, int nacc, int HW, long long npix, float w_dice, float w_ce,
                                                            float w_bd, float grad_scale) {
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
  const float sm = 1e-5f;
  const int nvals = nacc * ncls + 1;
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxCls; c++) {
    if (c < ncls) {
      v[c] *= inv;
      const float I = tot[nacc * c], Z = tot[nacc * c + 1], Y = tot[nacc * c + 2];
      const float D = Z + Y + sm;
      const float tc = (c == t) ? 1.f : 0.f;
      g[c] = -w_dice / (float)ncls * (2.f * tc * D - (2.f * I + sm) * 2.f * v[c]) / (D * D);
      if (nacc == 4) {
        const float alpha = tot[nvals + 3 + c];
        const float Nb = Z + Y - 2.f * I + sm, Db = Z + Y - (1.f + alpha) * I + sm;
        const float dN = 2.f * v[c] - 2.f * tc, dD = 2.f * v[c] - (1.f + alpha) * tc;
        g[c] += w_bd / (float)ncls * (dN * Db - Nb * dD) / (Db * Db);
      }
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
                                          int ncls, int ldy, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(y && (logits_nchw || labels), "cenet_head_upsample_argmax: null pointer");
  CENET_REQUIRE(ncls >= 1 && ncls <= kMaxCls, "cenet_head_upsample_argmax: 1..%d classes supported, got %d", kMaxCls, ncls);
  if (ldy <= 0) ldy = ncls;
  CENET_REQUIRE(ldy >= ncls, "cenet_head_upsample_argmax: pitch %d < %d classes", ldy, ncls);
  const long long total = (long long)B * 4 * h * w;
  const int grid = cdiv(total, 256);
  if (ncls == 9) head_upsample_argmax_kernel<9><<<grid, 256, 0, to_stream(s)>>>(y, logits_nchw, labels, B, h, w, ncls, ldy);
  else if (ncls == 4) head_upsample_argmax_kernel<4><<<grid, 256, 0, to_stream(s)>>>(y, logits_nchw, labels, B, h, w, ncls, ldy);
  else if (ncls == 2) head_upsample_argmax_kernel<2><<<grid, 256, 0, to_stream(s)>>>(y, logits_nchw, labels, B, h, w, ncls, ldy);
  else head_upsample_argmax_kernel<0><<<grid, 256, 0, to_stream(s)>>>(y, logits_nchw, labels, B, h, w, ncls, ldy);
  CENET_LAUNCH_CHECK("head_upsample_argmax");
  return 0;
}

extern "C" int cenet_loss_nblocks(long long npix) { return (int)((npix + kLossPixPerBlock - 1) / kLossPixPerBlock); }

static int seg_loss_launch(const float* logits, const long long* labels, float* loss_out, float* dlogits, float* ws, int B,
                           int ncls, int H, int W, int nacc, float w_dice, float w_ce, float w_bd, float grad_scale,
                           cudaStream_t s) {
  const int HW = H * W;
  const long long npix = (long long)B * HW;
  const int nblk = cenet_loss_nblocks(npix);
  const int nvals = nacc * ncls + 1;
  float* part = ws;
  float* tot = ws + (long long)nblk * nvals;
  seg_loss_partial_kernel<<<nblk, kLossThreads, 0, s>>>(logits, labels, part, ncls, nacc, H, W, npix);
  CENET_LAUNCH_CHECK("seg_loss_partial");
  seg_loss_finalize_kernel<<<1, 128, 0, s>>>(part, tot, loss_out, nblk, ncls, nacc, npix, w_dice, w_ce, w_bd);
  CENET_LAUNCH_CHECK("seg_loss_finalize");
  if (dlogits) {
    seg_loss_grad_kernel<<<cdiv(npix, 256), 256, 0, s>>>(logits, labels, tot, dlogits, ncls, nacc, HW, npix, w_dice, w_ce, w_bd,
                                                         grad_scale);
    CENET_LAUNCH_CHECK("seg_loss_grad");
  }
  return 0;
}

extern "C" int cenet_dice_ce(const float* logits, const long long* labels, float* loss_out, float* dlogits, float* ws,
                             int B, int ncls, int HW, float w_dice, float w_ce, float grad_scale, cenet_stream_t s) {
  CENET_REQUIRE(logits && labels && loss_out && ws, "cenet_dice_ce: null pointer");
  CENET_REQUIRE(B >= 1 && ncls >= 1 && ncls <= kMaxCls, "cenet_dice_ce: 1..%d classes supported, got %d", kMaxCls, ncls);
  return seg_loss_launch(logits, labels, loss_out, dlogits, ws, B, ncls, HW, 1, 3, w_dice, w_ce, 0.f, grad_scale, to_stream(s));
}

extern "C" int cenet_seg_loss(const float* logits, const long long* labels, float* loss_out, float* dlogits, float* ws,
                              int B, int ncls, int H, int W, float w_dice, float w_ce, float w_boundary, float grad_scale,
                              cenet_stream_t s) {
  CENET_REQUIRE(logits && labels && loss_out && ws, "cenet_seg_loss: null pointer");
  CENET_REQUIRE(B >= 1 && H >= 1 && W >= 1 && ncls >= 1 && ncls <= kMaxCls, "cenet_seg_loss: 1..%d classes supported, got %d",
                kMaxCls, ncls);
  return seg_loss_launch(logits, labels, loss_out, dlogits, ws, B, ncls, H, W, 4, w_dice, w_ce, w_boundary, grad_scale, to_stream(s));
}
