// CFAM statistics kernels (cfam.py): CCU channel gate, SRM spatial gate, image-pooling branch.
#include "common.cuh"
#include <algorithm>

namespace {
constexpr int kCcuChunk = 128;   // pixels per partial-statistics chunk

// ---- CCU pass 1: per (b, chunk, c) partial [max, mean, M2] of (x*scale+shift) ---------------------------------
// block = 64 channels x 4 pixel lanes; channel-contiguous loads (128 B per pixel row for bf16).
template <typename T>
__global__ void __launch_bounds__(256) ccu_partial_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, float* __restrict__ ws,
                                                          int HW, int C, int nchunk) {
  __shared__ float s_max[4][64], s_mean[4][64], s_m2[4][64], s_n[4][64];
  const int tc = threadIdx.x & 63, tp = threadIdx.x >> 6;
  const int c = blockIdx.z * 64 + tc, b = blockIdx.y, chunk = blockIdx.x;
  const int p0 = chunk * kCcuChunk, p1 = min(p0 + kCcuChunk, HW);
  float mx = -INFINITY, mean = 0.f, m2 = 0.f, n = 0.f;
  if (c < C) {
    const float sc = scale ? scale[c] : 1.f, sh = shift ? shift[c] : 0.f;
    for (int p = p0 + tp; p < p1; p += 4) {
      const float v = ldf(x + ((long long)b * HW + p) * C + c) * sc + sh;
      mx = fmaxf(mx, v);
      n += 1.f;
      const float d = v - mean;
      mean += d / n;
      m2 += d * (v - mean);
    }
  }
  s_max[tp][tc] = mx; s_mean[tp][tc] = mean; s_m2[tp][tc] = m2; s_n[tp][tc] = n;
  __syncthreads();
  if (tp == 0 && c < C) {
    for (int k = 1; k < 4; k++) {
      const float nb = s_n[k][tc];
      if (nb > 0.f) {
        const float d = s_mean[k][tc] - mean, nt = n + nb;
        m2 += s_m2[k][tc] + d * d * n * nb / nt;
        mean += d * nb / nt;
        n = nt;
        mx = fmaxf(mx, s_max[k][tc]);
      }
    }
    float* o = ws + (((long long)b * nchunk + chunk) * C + c) * 3;
    o[0] = mx; o[1] = mean; o[2] = m2;
  }
}

// ---- CCU pass 2: merge chunks, 3->3->1 MLP per channel, optional BN1d affine, sigmoid --------------------------
__global__ void __launch_bounds__(256) ccu_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ fc1,
                                                           const float* __restrict__ fc2, const float* __restrict__ bns,
                                                           const float* __restrict__ bnt, float* __restrict__ gate,
                                                           int B, int HW, int C, int nchunk) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  const int c = idx % C, b = idx / C;
  float mx = -INFINITY, mean = 0.f, m2 = 0.f, n = 0.f;
  for (int k = 0; k < nchunk; k++) {
    const float* o = ws + (((long long)b * nchunk + k) * C + c) * 3;
    const float nb = (float)(min((k + 1) * kCcuChunk, HW) - k * kCcuChunk);
    const float d = o[1] - mean, nt = n + nb;
    m2 += o[2] + d * d * n * nb / nt;
    mean += d * nb / nt;
    n = nt;
    mx = fmaxf(mx, o[0]);
  }
  const float u[3] = {mx, mean, sqrtf(m2 / n)};
  float z = 0.f;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float h = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) h = fmaf(fc1[(c * 3 + j) * 3 + k], u[k], h);
    z = fmaf(fc2[c * 3 + j], fmaxf(h, 0.f), z);
  }
  if (bns) z = z * bns[c] + bnt[c];
  gate[idx] = 1.f / (1.f + expf(-z));
}

// ---- SRM gate: one thread per pixel ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) srm_gate_kernel(const float* __restrict__ u, float* __restrict__ gate,
                                                       const float* __restrict__ pw3, const float* __restrict__ dw27,
                                                       float bn_scale, float bn_shift, int B, int H, int W) {
  __shared__ float sdw[27], spw[3];
  if (threadIdx.x < 27) sdw[threadIdx.x] = dw27[threadIdx.x];
  if (threadIdx.x < 3) spw[threadIdx.x] = pw3[threadIdx.x];
  __syncthreads();
  const float pw0 = spw[0], pw1 = spw[1], pw2 = spw[2];
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W) return;
  const int w = (int)(idx % W), h = (int)((idx / W) % H);
  const long long b = idx / ((long long)H * W);
  const float* uc = u + idx * 3;
  float f = pw0 * uc[0] + pw1 * uc[1] + pw2 * uc[2];
  for (int dh = -1; dh <= 1; dh++) {
    const int hh = h + dh;
    if (hh < 0 || hh >= H) continue;
    for (int dwi = -1; dwi <= 1; dwi++) {
      const int ww = w + dwi;
      if (ww < 0 || ww >= W) continue;
      const float* un = u + ((b * H + hh) * W + ww) * 3;
      const int t = (dh + 1) * 3 + (dwi + 1);
      f += sdw[t] * un[0] + sdw[9 + t] * un[1] + sdw[18 + t] * un[2];
    }
  }
  f = gelu_erf(f) * bn_scale + bn_shift;
  gate[idx] = 1.f / (1.f + expf(-f));
}

// ---- pooling branch, step 1: AdaptiveAvgPool(7) -> 1x1 (r->r) -> BN -> LeakyReLU; CTA per (bin, b) ------------
template <typename T>
__global__ void __launch_bounds__(128) pool7_conv_kernel(const T* __restrict__ x, long long ldx, int coff,
                                                         const float* __restrict__ w_rr, const float* __restrict__ bns,
                                                         const float* __restrict__ bnt, float slope,
                                                         float* __restrict__ pooled, int H, int W, int r) {
  extern __shared__ float avg[];   // r floats
  const int bin = blockIdx.x, b = blockIdx.y;
  const int bi = bin / 7, bj = bin % 7;
  const int h0 = (bi * H) / 7, h1 = ((bi + 1) * H + 6) / 7;
  const int w0 = (bj * W) / 7, w1 = ((bj + 1) * W + 6) / 7;
  const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
  for (int c = threadIdx.x; c < r; c += blockDim.x) {
    float s = 0.f;
    for (int h = h0; h < h1; h++)
      for (int w = w0; w < w1; w++) s += ldf(x + (((long long)b * H + h) * W + w) * ldx + coff + c);
    avg[c] = s * inv;
  }
  __syncthreads();
  for (int co = threadIdx.x; co < r; co += blockDim.x) {
    float a = 0.f;
    for (int ci = 0; ci < r; ci++) a = fmaf(w_rr[co * r + ci], avg[ci], a);
    a = a * bns[co] + bnt[co];
    pooled[((long long)b * 49 + bin) * r + co] = a > 0.f ? a : a * slope;
  }
}

// ---- pooling branch, step 2: 7x7 -> (49x49, align_corners=True) -> (H,W, align_corners=False), composed ----------
template <typename T>
__global__ void __launch_bounds__(256) pool_upsample_kernel(const float* __restrict__ pooled, T* __restrict__ y,
                                                            long long ldy, int coff_y, int B, int H, int W, int r) {
  const long long total = (long long)B * H * W * r;
  const bool same = (H == 49 && W == 49);
  const float s2h = 49.f / (float)H, s2w = 49.f / (float)W;   // size= semantics: in/out
  const float s1 = 6.f / 48.f;                                // align_corners=True: (7-1)/(49-1)
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % r);
    long long p = idx / r;
    const int w = (int)(p % W);
    p /= W;
    const int h = (int)(p % H);
    const int b = (int)(p / H);
    int ih[2], iw[2];
    float lh, lw;
    if (same) { ih[0] = ih[1] = h; iw[0] = iw[1] = w; lh = lw = 0.f; }
    else { bilin_src(h, s2h, 49, ih[0], ih[1], lh); bilin_src(w, s2w, 49, iw[0], iw[1], lw); }
    const float* pb = pooled + (long long)b * 49 * r + c;
    float v[2][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int q = 0; q < 2; q++) {
        int h0, h1, w0, w1;
        float mh, mw;
        bilin_src_ac(ih[a], s1, 7, h0, h1, mh);
        bilin_src_ac(iw[q], s1, 7, w0, w1, mw);
        v[a][q] = (1.f - mh) * ((1.f - mw) * pb[(h0 * 7 + w0) * r] + mw * pb[(h0 * 7 + w1) * r]) +
                  mh * ((1.f - mw) * pb[(h1 * 7 + w0) * r] + mw * pb[(h1 * 7 + w1) * r]);
      }
    const float o = (1.f - lh) * ((1.f - lw) * v[0][0] + lw * v[0][1]) + lh * ((1.f - lw) * v[1][0] + lw * v[1][1]);
    stf(y + (((long long)b * H + h) * W + w) * ldy + coff_y + c, o);
  }
}
}  // namespace

extern "C" int cenet_ccu_nchunk(int HW) { return (HW + kCcuChunk - 1) / kCcuChunk; }

extern "C" int cenet_ccu_gate(const void* x, int x_dtype, const float* scale, const float* shift, const float* fc1_c33,
                              const float* fc2_c3, const float* bn_scale, const float* bn_shift, float* gate_bc,
                              float* ws, int B, int HW, int C, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(x && fc1_c33 && fc2_c3 && gate_bc && ws, "cenet_ccu_gate: null pointer");
  CENET_REQUIRE((scale == nullptr) == (shift == nullptr) && (bn_scale == nullptr) == (bn_shift == nullptr),
                "cenet_ccu_gate: scale/shift pairs must come together");
  const int nchunk = cenet_ccu_nchunk(HW);
  dim3 grid(nchunk, B, cdiv(C, 64));
  CENET_DISPATCH(x_dtype, T, (ccu_partial_kernel<T><<<grid, 256, 0, to_stream(s)>>>((const T*)x, scale, shift, ws, HW, C, nchunk)));
  CENET_LAUNCH_CHECK("ccu_partial");
  ccu_finalize_kernel<<<cdiv((long long)B * C, 256), 256, 0, to_stream(s)>>>(ws, fc1_c33, fc2_c3, bn_scale, bn_shift,
                                                                             gate_bc, B, HW, C, nchunk);
  CENET_LAUNCH_CHECK("ccu_finalize");
  return 0;
}

extern "C" int cenet_srm_gate(const float* u, float* gate, const float* pw3, const float* dw27, float bn_scale,
                              float bn_shift, int B, int H, int W, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(u && gate && pw3 && dw27, "cenet_srm_gate: null pointer");
  const long long total = (long long)B * H * W;
  srm_gate_kernel<<<cdiv(total, 256), 256, 0, to_stream(s)>>>(u, gate, pw3, dw27, bn_scale, bn_shift, B, H, W);
  CENET_LAUNCH_CHECK("srm_gate");
  return 0;
}

extern "C" int cenet_pool_branch(const void* x, int x_dtype, long long ldx, int coff, void* y, int y_dtype,
                                 long long ldy, int coff_y, const float* w_rr, const float* bn_scale,
                                 const float* bn_shift, float slope, float* pooled_ws, int B, int H, int W, int r,
                                 cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(x && y && w_rr && bn_scale && bn_shift && pooled_ws, "cenet_pool_branch: null pointer");
  CENET_REQUIRE(r >= 1 && r <= 1024 && H >= 1 && W >= 1, "cenet_pool_branch: bad shape");
  dim3 g1(49, B);
  CENET_DISPATCH(x_dtype, T, (pool7_conv_kernel<T><<<g1, 128, r * sizeof(float), to_stream(s)>>>(
      (const T*)x, ldx, coff, w_rr, bn_scale, bn_shift, slope, pooled_ws, H, W, r)));
  CENET_LAUNCH_CHECK("pool7_conv");
  const long long total = (long long)B * H * W * r;
  const int grid = (int)std::min<long long>(cdiv(total, 256), (long long)kNumSMs * 32);
  CENET_DISPATCH(y_dtype, T, (pool_upsample_kernel<T><<<grid, 256, 0, to_stream(s)>>>(pooled_ws, (T*)y, ldy, coff_y, B, H, W, r)));
  CENET_LAUNCH_CHECK("pool_upsample");
  return 0;
}
