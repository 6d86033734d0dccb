// Depthwise 3x3 family on channels-last tensors (HBM-bound).
// Thread = one pixel x V channels (16 bytes of bf16); a block covers PT consecutive pixels of one image row x CVT
// channel vectors, so a warp reads/writes contiguous channel runs and the 3x3 neighbourhood is served by L1/L2 (each
// input element comes from HBM once).  Index math is 32-bit with one division per thread.
// Algorithmic bytes: B*H*W*C*(sizeof(in)+sizeof(out)) (+ 9*C weights).
#include "common.cuh"
#include <algorithm>

namespace {
// GELU(x) = 0.5 x (1 + erf(x/sqrt2)); erf by Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, i.e. fp32-level for the
// bf16 / 1e-4 tolerances of this path) -- 1 MUFU.RCP + 1 MUFU.EX2 + 7 FMA instead of erff()'s branchy ~40 instructions
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);      // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}
__device__ __forceinline__ float dw_act(float v, int act, float slope) {
  return act == CENET_ACT_GELU ? gelu_fast(v) : apply_act(v, act, slope);
}

template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const TI* __restrict__ x, int ldx, TO* __restrict__ y, int ldy,
                                                        const float* __restrict__ w9c, const float* __restrict__ bias,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        int H, int W, int C, int dil, int up2, int act, float slope) {
  const int cvi = blockIdx.x * blockDim.x + threadIdx.x;     // channel-vector index
  const int w = blockIdx.y * blockDim.y + threadIdx.y;
  if (cvi * V >= C || w >= W) return;
  const int c = cvi * V;
  const int b = blockIdx.z / H, h = blockIdx.z - b * H;
  const int Hi = up2 ? H >> 1 : H, Wi = up2 ? W >> 1 : W;
  const TI* xb = x + (size_t)b * Hi * Wi * ldx + c;
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; v++) acc[v] = bias ? bias[c + v] : 0.f;
#pragma unroll
  for (int dh = -1; dh <= 1; dh++) {
    const int hh = h + dh * dil;
    if (hh < 0 || hh >= H) continue;
    const int hs = up2 ? hh >> 1 : hh;                       // nearest x2: src = floor(dst/2)
#pragma unroll
    for (int dw = -1; dw <= 1; dw++) {
      const int ww = w + dw * dil;
      if (ww < 0 || ww >= W) continue;
      const int ws = up2 ? ww >> 1 : ww;
      float xv[V], wv[V];
      ldv<V>(xb + (size_t)(hs * Wi + ws) * ldx, xv);
      ldv<V>(w9c + ((dh + 1) * 3 + (dw + 1)) * C + c, wv);
#pragma unroll
      for (int v = 0; v < V; v++) acc[v] = fmaf(xv[v], wv[v], acc[v]);
    }
  }
  float o[V];
#pragma unroll
  for (int v = 0; v < V; v++) {
    float t = acc[v];
    if (scale) t = fmaf(t, scale[c + v], shift[c + v]);
    o[v] = dw_act(t, act, slope);
  }
  stv<V>(y + ((size_t)(b * H + h) * W + w) * ldy + c, o);
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
  CENET_REQUIRE((long long)B * H <= 65535, "cenet_dwconv3x3: B*H=%lld exceeds the grid limit", (long long)B * H);
  int V = pick_vec({C, ldx, ldy, ptr_align_elems(x, dtype_size(x_dtype)), ptr_align_elems(y, dtype_size(y_dtype)),
                    ptr_align_elems(w9c, 4) * 2});
  if (V > 4 && (x_dtype == CENET_F32 || y_dtype == CENET_F32)) V = 4;   // keep fp32 accesses at 16 bytes
  const int cv = C / V;
  int tx = 1;
  while (tx < cv && tx < 64) tx <<= 1;                      // channel vectors per block (power of two <= 64)
  const int ty = 256 / tx;
  dim3 block(tx, ty), grid(cdiv(cv, tx), cdiv(W, ty), B * H);
#define LAUNCH(VV)                                                                                              \
  CENET_DISPATCH(x_dtype, TI, CENET_DISPATCH(y_dtype, TO, (dwconv3x3_kernel<TI, TO, VV><<<grid, block, 0, to_stream(s)>>>( \
      (const TI*)x, (int)ldx, (TO*)y, (int)ldy, w9c, bias, scale, shift, H, W, C, dil, up2, act, slope))))
  if (V == 8) LAUNCH(8); else if (V == 4) LAUNCH(4); else if (V == 2) LAUNCH(2); else LAUNCH(1);
#undef LAUNCH
  CENET_LAUNCH_CHECK("dwconv3x3");
  return 0;
}
