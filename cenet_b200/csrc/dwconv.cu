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

// bf16 storage: tanh-form GELU on the MUFU pipe (1 MUFU.TANH + 5 FMA; |diff to the erf form| < 5e-4, below the bf16
// rounding of the value it produces).  fp32 storage keeps the erf form (A&S 7.1.26, 1.5e-7).
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = 0.7978845608f * fmaf(0.044715f * x, x * x, x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
template <bool FAST>
__device__ __forceinline__ float dw_act2(float v, int act, float slope) {
  if (act == CENET_ACT_GELU) return FAST ? gelu_tanh(v) : gelu_fast(v);
  return apply_act(v, act, slope);
}

// thread = V channels x PW consecutive output pixels of one row: the 3 x (PW+2) input vectors are loaded and unpacked
// once and feed all PW outputs; filter taps are loaded once per thread.
template <typename TI, typename TO, int V, int PW, int RB>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const TI* __restrict__ x, int ldx, TO* __restrict__ y, int ldy,
                                                        const float* __restrict__ w9c, const float* __restrict__ bias,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        int H, int W, int C, int dil, int up2, int act, float slope,
                                                        int nrows, TO* __restrict__ zout) {
  constexpr bool FAST = sizeof(TO) == 2;
  // block = TX channel vectors x (RB rows x blockDim.y/RB pixel groups): neighbouring rows of the 3x3 window are served
  // by the same SM's L1, so each input element crosses the L2->SM fabric ~2x instead of ~4x
  const int cvi = blockIdx.x * blockDim.x + threadIdx.x;     // channel-vector index
  const int pgroups = blockDim.y / RB;
  const int w0 = (blockIdx.y * pgroups + threadIdx.y % pgroups) * PW;
  const int hrow = blockIdx.z * RB + threadIdx.y / pgroups;   // row index over B*H
  if (cvi * V >= C || w0 >= W || hrow >= nrows) return;
  const int c = cvi * V;
  const int b = hrow / H, h = hrow - b * H;
  const int Hi = up2 ? H >> 1 : H, Wi = up2 ? W >> 1 : W;
  const TI* xb = x + (size_t)b * Hi * Wi * ldx + c;
  float acc[PW][V];
#pragma unroll
  for (int v = 0; v < V; v++) {
    const float bv = bias ? bias[c + v] : 0.f;
#pragma unroll
    for (int p = 0; p < PW; p++) acc[p][v] = bv;
  }
  if (dil == 1) {
    // dense 3x3: sliding window over PW+2 input columns
#pragma unroll
    for (int dh = -1; dh <= 1; dh++) {
      const int hh = h + dh;
      if (hh < 0 || hh >= H) continue;
      const int hs = up2 ? hh >> 1 : hh;
      float wv[3][V];
#pragma unroll
      for (int t = 0; t < 3; t++) ldv<V>(w9c + ((dh + 1) * 3 + t) * C + c, wv[t]);
#pragma unroll
      for (int q = 0; q < PW + 2; q++) {
        const int ww = w0 + q - 1;
        if (ww < 0 || ww >= W) continue;
        const int ws = up2 ? ww >> 1 : ww;
        float xv[V];
        ldv<V>(xb + (size_t)(hs * Wi + ws) * ldx, xv);
#pragma unroll
        for (int p = 0; p < PW; p++) {
          const int t = q - p;                 // tap column index (0..2) of input column q for output pixel p
          if (t < 0 || t > 2) continue;
#pragma unroll
          for (int v = 0; v < V; v++) acc[p][v] = fmaf(xv[v], wv[t][v], acc[p][v]);
        }
      }
    }
  } else {
#pragma unroll
    for (int dh = -1; dh <= 1; dh++) {
      const int hh = h + dh * dil;
      if (hh < 0 || hh >= H) continue;
      const int hs = up2 ? hh >> 1 : hh;
#pragma unroll
      for (int dw = -1; dw <= 1; dw++) {
        float wv[V];
        ldv<V>(w9c + ((dh + 1) * 3 + (dw + 1)) * C + c, wv);
#pragma unroll
        for (int p = 0; p < PW; p++) {
          const int ww = w0 + p + dw * dil;
          if (ww < 0 || ww >= W || w0 + p >= W) continue;
          const int ws = up2 ? ww >> 1 : ww;
          float xv[V];
          ldv<V>(xb + (size_t)(hs * Wi + ws) * ldx, xv);
#pragma unroll
          for (int v = 0; v < V; v++) acc[p][v] = fmaf(xv[v], wv[v], acc[p][v]);
        }
      }
    }
  }
  float sc[V], sh[V];
#pragma unroll
  for (int v = 0; v < V; v++) { sc[v] = scale ? scale[c + v] : 1.f; sh[v] = scale ? shift[c + v] : 0.f; }
#pragma unroll
  for (int p = 0; p < PW; p++) {
    if (w0 + p >= W) continue;
    float o[V];
#pragma unroll
    for (int v = 0; v < V; v++) o[v] = fmaf(acc[p][v], sc[v], sh[v]);
    if (zout) stv<V>(zout + ((size_t)(b * H + h) * W + w0 + p) * C + c, o);     // training: keep the pre-activation
#pragma unroll
    for (int v = 0; v < V; v++) o[v] = dw_act2<FAST>(o[v], act, slope);
    stv<V>(y + ((size_t)(b * H + h) * W + w0 + p) * ldy + c, o);
  }
}
}  // namespace

static int dwconv_launch(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, long long ldy,
                         const float* w9c, const float* bias, const float* scale, const float* shift, int B,
                         int H, int W, int C, int dil, int up2, int act, float slope, void* zout, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(x && y && w9c, "cenet_dwconv3x3: null pointer");
  CENET_REQUIRE((scale == nullptr) == (shift == nullptr), "cenet_dwconv3x3: scale and shift come together");
  CENET_REQUIRE(!up2 || (H % 2 == 0 && W % 2 == 0), "cenet_dwconv3x3: up2 needs even output size");
  CENET_REQUIRE(ldx >= C && ldy >= C && dil >= 1, "cenet_dwconv3x3: bad pitch / dilation");
  CENET_REQUIRE((long long)B * H <= 65535, "cenet_dwconv3x3: B*H=%lld exceeds the grid limit", (long long)B * H);
  int V = pick_vec({C, ldx, ldy, ptr_align_elems(x, dtype_size(x_dtype)), ptr_align_elems(y, dtype_size(y_dtype)),
                    ptr_align_elems(w9c, 4) * 2, zout ? ptr_align_elems(zout, dtype_size(y_dtype)) : 8});
  if (V > 4 && (x_dtype == CENET_F32 || y_dtype == CENET_F32)) V = 4;   // keep fp32 accesses at 16 bytes
  const int cv = C / V;
  constexpr int PW = 2;
  constexpr int RB = 1;                                     // image rows per block (RB = 4 measured 7% slower on B200)
  int tx = 1;
  while (tx < cv && tx < 64) tx <<= 1;                      // channel vectors per block (power of two <= 64)
  const int ty = 256 / tx;                                  // = RB rows x (ty/RB) pixel groups
  const int pgroups = ty / RB;
  const int nrows = B * H;
  dim3 block(tx, ty), grid(cdiv(cv, tx), cdiv(cdiv(W, PW), pgroups), cdiv(nrows, RB));
  CENET_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "cenet_dwconv3x3: grid too large");
#define LAUNCH(VV)                                                                                              \
  CENET_DISPATCH(x_dtype, TI, CENET_DISPATCH(y_dtype, TO, (dwconv3x3_kernel<TI, TO, VV, PW, RB><<<grid, block, 0, to_stream(s)>>>( \
      (const TI*)x, (int)ldx, (TO*)y, (int)ldy, w9c, bias, scale, shift, H, W, C, dil, up2, act, slope, nrows, (TO*)zout))))
  if (V == 8) LAUNCH(8); else if (V == 4) LAUNCH(4); else if (V == 2) LAUNCH(2); else LAUNCH(1);
#undef LAUNCH
  CENET_LAUNCH_CHECK("dwconv3x3");
  return 0;
}

extern "C" int cenet_dwconv3x3(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, long long ldy,
                               const float* w9c, const float* bias, const float* scale, const float* shift, int B,
                               int H, int W, int C, int dil, int up2, int act, float slope, cenet_stream_t s) {
  return dwconv_launch(x, x_dtype, ldx, y, y_dtype, ldy, w9c, bias, scale, shift, B, H, W, C, dil, up2, act, slope, nullptr, s);
}

// training forward: also stores the pre-activation z (contiguous [B,H,W,C], dtype z_dtype == y_dtype) for the backward pass
extern "C" int cenet_dwconv3x3_train(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, long long ldy, void* zout,
                                     const float* w9c, const float* bias, int B, int H, int W, int C, int dil, int act,
                                     int z_dtype, float slope, cenet_stream_t s) {
  CENET_REQUIRE(zout != nullptr && z_dtype == y_dtype, "cenet_dwconv3x3_train: zout must be given with the dtype of y");
  return dwconv_launch(x, x_dtype, ldx, y, y_dtype, ldy, w9c, bias, nullptr, nullptr, B, H, W, C, dil, 0, act, slope, zout, s);
}
