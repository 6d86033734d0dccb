// OutHead stem (out.py:41-44 / unet.py:201-209): the first UnetResBlock convolution on the RAW input image.
//   o1 = LeakyReLU( BN1(conv5x5(x)) )        Cin in {1..4} -> 32 channels, K = 25*Cin is far too shallow for an MMA tile
//   r  = BN3(conv1x1(x))                     the block's residual branch (reads the same pixels)
// One thread = one pixel x 32 output channels; a 16x16 pixel tile with its 2-pixel halo sits in shared memory, the
// (BN-folded) filters are read as float4 broadcasts.  CUDA-core kernel: 800*Cin FMA per pixel against 128 B written,
// i.e. FMA-bound at ~0.2 ms for 64 x 224^2 pixels; HBM sees x once and o1, r once.
#include "common.cuh"

namespace {
constexpr int OC = 32, TS = 16, KS = 5, HALO = 2, TIN = TS + 2 * HALO;

// Thread = TWO horizontally adjacent pixels x 32 output channels in packed fp32 pairs (FFMA2): a filter float4 read from
// shared memory feeds 2 pixels x 4 channels = 4 FFMA2, i.e. 42 issue slots per tap and pixel pair instead of 82 with scalar FMAs
// and one pixel per thread (the kernel is issue-bound: 800*Cin FMA per pixel).  Tile: 32 (w) x 16 (h) pixels per 256 threads.
constexpr int TW = 2 * TS, TINW = TW + 2 * HALO;

template <typename TI, typename TO, int CIN>
__global__ void __launch_bounds__(256) stem5x5_kernel(const TI* __restrict__ x, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, const float* __restrict__ w3,
                                                      const float* __restrict__ b3, TO* __restrict__ o1, TO* __restrict__ r,
                                                      int H, int W, float slope) {
  __shared__ __align__(16) float sw[KS * KS * CIN][OC];     // [tap*CIN + ci][oc]
  __shared__ float sx[TIN][TINW][CIN];
  const int tid = threadIdx.y * TS + threadIdx.x;
  const int b = blockIdx.z, h0 = blockIdx.y * TS, w0 = blockIdx.x * TW;
  for (int i = tid; i < KS * KS * CIN * OC; i += 256) {
    const int oc = i % OC, k = i / OC;                      // w1 is [oc][k] (k = (kh,kw,ci))
    sw[k][oc] = w1[oc * (KS * KS * CIN) + k];
  }
  for (int i = tid; i < TIN * TINW * CIN; i += 256) {
    const int ci = i % CIN, p = i / CIN, ww = p % TINW, hh = p / TINW;
    const int h = h0 + hh - HALO, w = w0 + ww - HALO;
    sx[hh][ww][ci] = (h >= 0 && h < H && w >= 0 && w < W) ? ldf(x + (((size_t)b * H + h) * W + w) * CIN + ci) : 0.f;
  }
  __syncthreads();
  const int h = h0 + threadIdx.y, wl = 2 * threadIdx.x, w = w0 + wl;
  f32x2 acc[2][OC / 2];
#pragma unroll
  for (int o = 0; o < OC / 2; o++) acc[0][o] = acc[1][o] = pk2(b1[2 * o], b1[2 * o + 1]);
#pragma unroll
  for (int kh = 0; kh < KS; kh++)
#pragma unroll
    for (int kw = 0; kw < KS; kw++)
#pragma unroll
      for (int ci = 0; ci < CIN; ci++) {
        const float xa = sx[threadIdx.y + kh][wl + kw][ci], xb = sx[threadIdx.y + kh][wl + 1 + kw][ci];
        const f32x2 xa2 = pk2(xa, xa), xb2 = pk2(xb, xb);
        const float4* wr = reinterpret_cast<const float4*>(sw[(kh * KS + kw) * CIN + ci]);
#pragma unroll
        for (int q = 0; q < OC / 4; q++) {
          const float4 wv = wr[q];
          const f32x2 wlo = pk2(wv.x, wv.y), whi = pk2(wv.z, wv.w);
          acc[0][2 * q] = ffma2(xa2, wlo, acc[0][2 * q]);
          acc[0][2 * q + 1] = ffma2(xa2, whi, acc[0][2 * q + 1]);
          acc[1][2 * q] = ffma2(xb2, wlo, acc[1][2 * q]);
          acc[1][2 * q + 1] = ffma2(xb2, whi, acc[1][2 * q + 1]);
        }
      }
#pragma unroll
  for (int px = 0; px < 2; px++) {
    if (h >= H || w + px >= W) continue;
    const size_t pix = ((size_t)b * H + h) * W + w + px;
#pragma unroll
    for (int q = 0; q < OC / 8; q++) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float t0, t1;
        upk2(acc[px][4 * q + j], t0, t1);
        v[2 * j] = t0 > 0.f ? t0 : t0 * slope;
        v[2 * j + 1] = t1 > 0.f ? t1 : t1 * slope;
      }
      stv<8>(o1 + pix * OC + 8 * q, v);
    }
    if (r) {
      float xc[CIN];
#pragma unroll
      for (int ci = 0; ci < CIN; ci++) xc[ci] = sx[threadIdx.y + HALO][wl + px + HALO][ci];
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
  dim3 block(TS, TS), grid(cdiv(W, TW), cdiv(H, TS), B);
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
