// Depthwise 3x3 family on channels-last tensors (HBM-bound): one thread = one pixel x V channels, neighbours come
// from L1/L2 (each input element is re-read up to 9 times on-chip, once from HBM).
#include "common.cuh"
#include <algorithm>

namespace {
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const TI* __restrict__ x, long long ldx, TO* __restrict__ y,
                                                        long long ldy, const float* __restrict__ w9c,
                                                        const float* __restrict__ bias, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, int B, int H, int W, int C,
                                                        int dil, int up2, int act, float slope) {
  const int cv = C / V;
  const long long total = (long long)B * H * W * cv;
  const int Hi = up2 ? H / 2 : H, Wi = up2 ? W / 2 : W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * V;
    long long p = idx / cv;
    const int w = (int)(p % W);
    p /= W;
    const int h = (int)(p % H);
    const int b = (int)(p / H);
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; v++) acc[v] = bias ? bias[c + v] : 0.f;
#pragma unroll
    for (int dh = -1; dh <= 1; dh++) {
      const int hh = h + dh * dil;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int dw = -1; dw <= 1; dw++) {
        const int ww = w + dw * dil;
        if (ww < 0 || ww >= W) continue;
        const int hs = up2 ? hh >> 1 : hh, ws = up2 ? ww >> 1 : ww;   // nearest x2: src = floor(dst/2)
        float xv[V], wv[V];
        ldv<V>(x + (((long long)b * Hi + hs) * Wi + ws) * ldx + c, xv);
        ldv<V>(w9c + ((dh + 1) * 3 + (dw + 1)) * C + c, wv);
#pragma unroll
        for (int v = 0; v < V; v++) acc[v] = fmaf(xv[v], wv[v], acc[v]);
      }
    }
    float o[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      float t = acc[v];
      if (scale) t = t * scale[c + v] + shift[c + v];
      o[v] = apply_act(t, act, slope);
    }
    stv<V>(y + (((long long)b * H + h) * W + w) * ldy + c, o);
  }
}
}  // namespace

extern "C" int cenet_dwconv3x3(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, long long ldy,
                               const float* w9c, const float* bias, const float* scale, const float* shift, int B,
                               int H, int W, int C, int dil, int up2, int act, float slope, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(x && y && w9c, "cenet_dwconv3x3: null pointer");
  CENET_REQUIRE((scale == nullptr) == (shift == nullptr), "cenet_dwconv3x3: scale and shift come together");
  CENET_REQUIRE(!up2 || (H % 2 == 0 && W % 2 == 0), "cenet_dwconv3x3: up2 needs even output size");
  CENET_REQUIRE(ldx >= C && ldy >= C && dil >= 1, "cenet_dwconv3x3: bad pitch / dilation");
  int V = pick_vec({C, ldx, ldy, ptr_align_elems(x, dtype_size(x_dtype)), ptr_align_elems(y, dtype_size(y_dtype)),
                    ptr_align_elems(w9c, 4) * 2});
  if (V > 4 && (x_dtype == CENET_F32 || y_dtype == CENET_F32)) V = 4;   // keep fp32 accesses at 16 bytes
  const long long total = (long long)B * H * W * (C / V);
  const int grid = (int)std::min<long long>(cdiv(total, 256), (long long)kNumSMs * 32);
#define LAUNCH(VV)                                                                                              \
  CENET_DISPATCH(x_dtype, TI, CENET_DISPATCH(y_dtype, TO, (dwconv3x3_kernel<TI, TO, VV><<<grid, 256, 0, to_stream(s)>>>( \
      (const TI*)x, ldx, (TO*)y, ldy, w9c, bias, scale, shift, B, H, W, C, dil, up2, act, slope))))
  if (V == 8) LAUNCH(8); else if (V == 4) LAUNCH(4); else if (V == 2) LAUNCH(2); else LAUNCH(1);
#undef LAUNCH
  CENET_LAUNCH_CHECK("dwconv3x3");
  return 0;
}
