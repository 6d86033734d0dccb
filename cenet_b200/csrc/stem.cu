// OutHead stem (out.py:41-44 / unet.py:201-209): the first UnetResBlock convolution on the RAW input image.
//   o1 = LeakyReLU( BN1(conv5x5(x)) )        Cin in {1..4} -> 32 channels, K = 25*Cin is far too shallow for an MMA tile
//   r  = BN3(conv1x1(x))                     the block's residual branch (reads the same pixels)
// One thread = one pixel x 32 output channels; a 16x16 pixel tile with its 2-pixel halo sits in shared memory, the
// (BN-folded) filters are read as float4 broadcasts.  CUDA-core kernel: 800*Cin FMA per pixel against 128 B written,
// i.e. FMA-bound at ~0.2 ms for 64 x 224^2 pixels; HBM sees x once and o1, r once.
#include "common.cuh"

namespace {
constexpr int OC = 32, TS = 16, KS = 5, HALO = 2, TIN = TS + 2 * HALO;

template <typename TI, typename TO, int CIN>
__global__ void __launch_bounds__(256) stem5x5_kernel(const TI* __restrict__ x, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, const float* __restrict__ w3,
                                                      const float* __restrict__ b3, TO* __restrict__ o1, TO* __restrict__ r,
                                                      int H, int W, float slope) {
  pdl_prologue();
  __shared__ __align__(16) float sw[KS * KS * CIN][OC];     // [tap*CIN + ci][oc]
  __shared__ float sx[TIN][TIN][CIN];
  const int tid = threadIdx.y * TS + threadIdx.x;
  const int b = blockIdx.z, h0 = blockIdx.y * TS, w0 = blockIdx.x * TS;
  for (int i = tid; i < KS * KS * CIN * OC; i += 256) {
    const int oc = i % OC, k = i / OC;                      // w1 is [oc][k] (k = (kh,kw,ci))
    sw[k][oc] = w1[oc * (KS * KS * CIN) + k];
  }
  for (int i = tid; i < TIN * TIN * CIN; i += 256) {
    const int ci = i % CIN, p = i / CIN, ww = p % TIN, hh = p / TIN;
    const int h = h0 + hh - HALO, w = w0 + ww - HALO;
    sx[hh][ww][ci] = (h >= 0 && h < H && w >= 0 && w < W) ? ldf(x + (((size_t)b * H + h) * W + w) * CIN + ci) : 0.f;
  }
  __syncthreads();
  const int h = h0 + threadIdx.y, w = w0 + threadIdx.x;
  float acc[OC];
#pragma unroll
  for (int o = 0; o < OC; o++) acc[o] = b1[o];
#pragma unroll
  for (int kh = 0; kh < KS; kh++)
#pragma unroll
    for (int kw = 0; kw < KS; kw++)
#pragma unroll
      for (int ci = 0; ci < CIN; ci++) {
        const float xv = sx[threadIdx.y + kh][threadIdx.x + kw][ci];
        const float4* wr = reinterpret_cast<const float4*>(sw[(kh * KS + kw) * CIN + ci]);
#pragma unroll
        for (int q = 0; q < OC / 4; q++) {
          const float4 wv = wr[q];
          acc[4 * q] = fmaf(xv, wv.x, acc[4 * q]);
          acc[4 * q + 1] = fmaf(xv, wv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(xv, wv.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(xv, wv.w, acc[4 * q + 3]);
        }
      }
  if (h >= H || w >= W) return;
  const size_t pix = ((size_t)b * H + h) * W + w;
#pragma unroll
  for (int q = 0; q < OC / 8; q++) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { const float t = acc[8 * q + j]; v[j] = t > 0.f ? t : t * slope; }
    stv<8>(o1 + pix * OC + 8 * q, v);
  }
  if (r) {
    float xc[CIN];
#pragma unroll
    for (int ci = 0; ci < CIN; ci++) xc[ci] = sx[threadIdx.y + HALO][threadIdx.x + HALO][ci];
#pragma unroll
    for (int q = 0; q < OC / 8; q++) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        float t = b3[8 * q + j];
#pragma unroll
        for (int ci = 0; ci < CIN; ci++) t = fmaf(xc[ci], w3[(8 * q + j) * CIN + ci], t);
        v[j] = t;
      }
      stv<8>(r + pix * OC + 8 * q, v);
    }
  }
}
}  // namespace

extern "C" int cenet_stem5x5(const void* x, int x_dtype, const float* w1, const float* b1, const float* w3,
                             const float* b3, void* o1, void* r, int o_dtype, int B, int H, int W, int Cin, float slope,
                             cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(x && w1 && b1 && o1, "cenet_stem5x5: null pointer");
  CENET_REQUIRE((r == nullptr) || (w3 && b3), "cenet_stem5x5: residual output needs w3/b3");
  CENET_REQUIRE(Cin >= 1 && Cin <= 4, "cenet_stem5x5: Cin=%d not in 1..4", Cin);
  CENET_REQUIRE(B <= 65535, "cenet_stem5x5: batch too large for the grid");
  dim3 block(TS, TS), grid(cdiv(W, TS), cdiv(H, TS), B);
#define LAUNCH(CI)                                                                                                   \
  CENET_DISPATCH(x_dtype, TI, CENET_DISPATCH(o_dtype, TO, (stem5x5_kernel<TI, TO, CI><<<grid, block, 0, to_stream(s)>>>( \
      (const TI*)x, w1, b1, w3, b3, (TO*)o1, (TO*)r, H, W, slope))))
  switch (Cin) {
    case 1: LAUNCH(1); break;
    case 2: LAUNCH(2); break;
    case 3: LAUNCH(3); break;
    default: LAUNCH(4); break;
  }
#undef LAUNCH
  CENET_LAUNCH_CHECK("stem5x5");
  return 0;
}
