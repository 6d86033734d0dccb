// Layout / resampling helpers on channels-last activations (all HBM-bound, coalesced on both sides).
#include "common.cuh"
#include <algorithm>

namespace {
// ---- [B,HW,C] (pitch ldx) -> [B,Ctot,HW] at channel offset coff : 32x32 smem tile transpose -------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const TI* __restrict__ x, long long ldx, TO* __restrict__ y,
                                                           int HW, int C, int Ctot, int coff) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    int p = p0 + i, c = c0 + tx;
    tile[i][tx] = (p < HW && c < C) ? ldf(x + ((long long)b * HW + p) * ldx + c) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int c = c0 + i, p = p0 + tx;
    if (c < C && p < HW) stf(y + ((long long)b * Ctot + coff + c) * HW + p, tile[tx][i]);
  }
}
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long ldy,
                                                           int HW, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (p < HW && c < C) ? ldf(x + ((long long)b * C + c) * HW + p) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int p = p0 + i, c = c0 + tx;
    if (c < C && p < HW) stf(y + ((long long)b * HW + p) * ldy + c, tile[tx][i]);
  }
}

// ---- bilinear x2, align_corners=True (UpConv) -----------------------------------------------------------------
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256) upsample2x_ac_kernel(const TI* __restrict__ x, TO* __restrict__ y, int B, int H,
                                                            int W, int C) {
  const int Ho = 2 * H, Wo = 2 * W, cv = C / V;
  const float sh = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float sw = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const long long total = (long long)B * Ho * Wo * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * V;
    long long p = idx / cv;
    const int wo = (int)(p % Wo);
    p /= Wo;
    const int ho = (int)(p % Ho);
    const int b = (int)(p / Ho);
    int h0, h1, w0, w1;
    float lh, lw;
    bilin_src_ac(ho, sh, H, h0, h1, lh);
    bilin_src_ac(wo, sw, W, w0, w1, lw);
    const TI* base = x + (long long)b * H * W * C + c;
    float a[V], bb[V], cc[V], d[V], o[V];
    ldv<V>(base + ((long long)h0 * W + w0) * C, a);
    ldv<V>(base + ((long long)h0 * W + w1) * C, bb);
    ldv<V>(base + ((long long)h1 * W + w0) * C, cc);
    ldv<V>(base + ((long long)h1 * W + w1) * C, d);
#pragma unroll
    for (int v = 0; v < V; v++)
      o[v] = (1.f - lh) * ((1.f - lw) * a[v] + lw * bb[v]) + lh * ((1.f - lw) * cc[v] + lw * d[v]);
    stv<V>(y + (((long long)b * Ho + ho) * Wo + wo) * C + c, o);
  }
}

// ---- MaxPool2d(2) then per-channel scale, written into a channel slice ----------------------------------------
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256) maxpool2_scale_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long ldy,
                                                             int coff, const float* __restrict__ wch, int B, int H,
                                                             int W, int C) {
  const int Ho = H / 2, Wo = W / 2, cv = C / V;
  const long long total = (long long)B * Ho * Wo * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * V;
    long long p = idx / cv;
    const int wo = (int)(p % Wo);
    p /= Wo;
    const int ho = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const TI* base = x + (((long long)b * H + 2 * ho) * W + 2 * wo) * C + c;
    float a[V], bb[V], cc[V], d[V], o[V];
    ldv<V>(base, a);
    ldv<V>(base + C, bb);
    ldv<V>(base + (long long)W * C, cc);
    ldv<V>(base + (long long)W * C + C, d);
#pragma unroll
    for (int v = 0; v < V; v++) o[v] = fmaxf(fmaxf(a[v], bb[v]), fmaxf(cc[v], d[v])) * wch[c + v];
    stv<V>(y + (((long long)b * Ho + ho) * Wo + wo) * ldy + coff + c, o);
  }
}

// ---- y = (x*scale[c]+shift[c]) * gate[b,c] -----------------------------------------------------------------------
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256) affine_gate_kernel(const TI* __restrict__ x, TO* __restrict__ y,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          const float* __restrict__ gate, int B, int HW, int C) {
  const int cv = C / V;
  const long long total = (long long)B * HW * cv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * V;
    const long long p = idx / cv;
    const int b = (int)(p / HW);
    float v[V], o[V];
    ldv<V>(x + p * C + c, v);
#pragma unroll
    for (int i = 0; i < V; i++) {
      float t = v[i];
      if (scale) t = t * scale[c + i] + shift[c + i];
      o[i] = gate ? t * gate[(long long)b * C + c + i] : t;
    }
    stv<V>(y + p * C + c, o);
  }
}

// ---- im2col for strided / non-overlapping convolutions: out[m, (kh,kw,ci)] (row pitch Kpad, zero padded) ---------
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256) im2col_kernel(const TI* __restrict__ x, TO* __restrict__ out, int B, int H, int W,
                                                     int Cin, int KH, int KW, int stride, int pad, int Ho, int Wo,
                                                     int Kpad) {
  const int kv = Kpad / V;
  const int K = KH * KW * Cin;
  const long long total = (long long)B * Ho * Wo * kv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % kv) * V;
    long long m = idx / kv;
    const int wo = (int)(m % Wo);
    long long t = m / Wo;
    const int ho = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float v[V];
#pragma unroll
    for (int i = 0; i < V; i++) v[i] = 0.f;
    if (k < K) {
      const int ci = k % Cin, tap = k / Cin, kw = tap % KW, kh = tap / KW;
      const int h = ho * stride - pad + kh, w = wo * stride - pad + kw;
      if (h >= 0 && h < H && w >= 0 && w < W) ldv<V>(x + (((long long)b * H + h) * W + w) * Cin + ci, v);
    }
    stv<V>(out + m * Kpad + k, v);
  }
}

// Few input channels (the 1- or 3-channel image: patch_embed1 7x7 s4, the 5x5 stem of the output head in training): a V-wide vector
// of the generic kernel would be one element (V divides Cin), i.e. one 64-bit index decomposition and one 2-byte store per
// element.  Here a thread produces 8 consecutive k of one output row: one decomposition, 8 scalar gathers (L1), one 16-byte store.
template <typename TI>
__global__ void __launch_bounds__(256) im2col_gather8_kernel(const TI* __restrict__ x, bf16* __restrict__ out, int B, int H, int W,
                                                             int Cin, int KH, int KW, int stride, int pad, int Ho, int Wo, int Kpad) {
  const int kv = Kpad >> 3;
  const int K = KH * KW * Cin;
  const long long total = (long long)B * Ho * Wo * kv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / kv;
    const int k0 = (int)(idx - m * kv) * 8;
    const int mi = (int)m;                                   // B * Ho * Wo < 2^31 (checked by the launcher)
    const int wo = mi % Wo, t = mi / Wo, ho = t % Ho, b = t / Ho;
    int ci = k0 % Cin, tap = k0 / Cin, kw = tap % KW, kh = tap / KW;
    const int hb = ho * stride - pad, wb = wo * stride - pad;
    const TI* xb = x + (long long)b * H * W * Cin;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int h = hb + kh, w = wb + kw;
      v[i] = (k0 + i < K && h >= 0 && h < H && w >= 0 && w < W) ? ldf(xb + ((long long)h * W + w) * Cin + ci) : 0.f;
      if (++ci == Cin) { ci = 0; if (++kw == KW) { kw = 0; kh++; } }
    }
    stv<8>(out + m * Kpad + k0, v);
  }
}

inline int ew_grid(long long total) { return (int)std::min<long long>(cdiv(total, 256), (long long)kNumSMs * 32); }
}  // namespace

#define DISPATCH_V(V, MAXV, ...)                                  \
  do {                                                            \
    int vv__ = (V) > (MAXV) ? (MAXV) : (V);                       \
    if (vv__ == 8) { constexpr int VV = 8; __VA_ARGS__; }         \
    else if (vv__ == 4) { constexpr int VV = 4; __VA_ARGS__; }    \
    else if (vv__ == 2) { constexpr int VV = 2; __VA_ARGS__; }    \
    else { constexpr int VV = 1; __VA_ARGS__; }                   \
  } while (0)

extern "C" int cenet_nhwc_to_nchw(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, int B, int HW,
                                  int C, int Ctot, int coff, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(x && y && coff + C <= Ctot && ldx >= C, "cenet_nhwc_to_nchw: bad arguments");
  dim3 grid(cdiv(HW, 32), cdiv(C, 32), B);
  CENET_DISPATCH(x_dtype, TI, CENET_DISPATCH(y_dtype, TO, (nhwc_to_nchw_kernel<TI, TO><<<grid, 256, 0, to_stream(s)>>>(
      (const TI*)x, ldx, (TO*)y, HW, C, Ctot, coff))));
  CENET_LAUNCH_CHECK("nhwc_to_nchw");
  return 0;
}

extern "C" int cenet_nchw_to_nhwc(const void* x, int x_dtype, void* y, int y_dtype, long long ldy, int B, int HW,
                                  int C, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(x && y && ldy >= C, "cenet_nchw_to_nhwc: bad arguments");
  dim3 grid(cdiv(HW, 32), cdiv(C, 32), B);
  CENET_DISPATCH(x_dtype, TI, CENET_DISPATCH(y_dtype, TO, (nchw_to_nhwc_kernel<TI, TO><<<grid, 256, 0, to_stream(s)>>>(
      (const TI*)x, (TO*)y, ldy, HW, C))));
  CENET_LAUNCH_CHECK("nchw_to_nhwc");
  return 0;
}

extern "C" int cenet_upsample2x_ac(const void* x, int x_dtype, void* y, int y_dtype, int B, int H, int W, int C,
                                   cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(x && y, "cenet_upsample2x_ac: null pointer");
  int V = pick_vec({C});
  const int maxv = (x_dtype == CENET_F32 || y_dtype == CENET_F32) ? 4 : 8;
  const long long total = (long long)B * 4 * H * W * C;
  DISPATCH_V(V, maxv, CENET_DISPATCH(x_dtype, TI, CENET_DISPATCH(y_dtype, TO,
      (upsample2x_ac_kernel<TI, TO, VV><<<ew_grid(total / VV), 256, 0, to_stream(s)>>>((const TI*)x, (TO*)y, B, H, W, C)))));
  CENET_LAUNCH_CHECK("upsample2x_ac");
  return 0;
}

extern "C" int cenet_maxpool2_scale(const void* x, int x_dtype, void* y, int y_dtype, long long ldy, int coff,
                                    const float* wch, int B, int H, int W, int C, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(x && y && wch && H % 2 == 0 && W % 2 == 0 && ldy >= coff + C, "cenet_maxpool2_scale: bad arguments");
  int V = pick_vec({C, ldy, coff});
  const int maxv = (x_dtype == CENET_F32 || y_dtype == CENET_F32) ? 4 : 8;
  const long long total = (long long)B * (H / 2) * (W / 2) * C;
  DISPATCH_V(V, maxv, CENET_DISPATCH(x_dtype, TI, CENET_DISPATCH(y_dtype, TO,
      (maxpool2_scale_kernel<TI, TO, VV><<<ew_grid(total / VV), 256, 0, to_stream(s)>>>((const TI*)x, (TO*)y, ldy, coff, wch,
                                                                                       B, H, W, C)))));
  CENET_LAUNCH_CHECK("maxpool2_scale");
  return 0;
}

extern "C" int cenet_affine_gate(const void* x, int x_dtype, void* y, int y_dtype, const float* scale,
                                 const float* shift, const float* gate_bc, int B, int HW, int C, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(x && y && ((scale == nullptr) == (shift == nullptr)), "cenet_affine_gate: bad arguments");
  int V = pick_vec({C});
  const int maxv = (x_dtype == CENET_F32 || y_dtype == CENET_F32) ? 4 : 8;
  const long long total = (long long)B * HW * C;
  DISPATCH_V(V, maxv, CENET_DISPATCH(x_dtype, TI, CENET_DISPATCH(y_dtype, TO,
      (affine_gate_kernel<TI, TO, VV><<<ew_grid(total / VV), 256, 0, to_stream(s)>>>((const TI*)x, (TO*)y, scale, shift,
                                                                                    gate_bc, B, HW, C)))));
  CENET_LAUNCH_CHECK("affine_gate");
  return 0;
}

extern "C" int cenet_im2col(const void* x, int x_dtype, void* out, int o_dtype, int B, int H, int W, int Cin, int KH,
                            int KW, int stride, int pad, int Ho, int Wo, int Kpad, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(x && out && Kpad >= KH * KW * Cin, "cenet_im2col: bad arguments");
  // a V-wide vector must not straddle a filter tap: V divides Cin (and Kpad)
  int V = pick_vec({Cin, Kpad});
  const int maxv = (x_dtype == CENET_F32 || o_dtype == CENET_F32) ? 4 : 8;
  const long long total = (long long)B * Ho * Wo * Kpad;
  if (V < 4 && o_dtype == CENET_BF16 && Kpad % 8 == 0 && (((uintptr_t)out) & 15) == 0 && (long long)B * Ho * Wo < (1LL << 31)) {
    CENET_DISPATCH(x_dtype, TI, (im2col_gather8_kernel<TI><<<ew_grid(total / 8), 256, 0, to_stream(s)>>>(
        (const TI*)x, (bf16*)out, B, H, W, Cin, KH, KW, stride, pad, Ho, Wo, Kpad)));
    CENET_LAUNCH_CHECK("im2col_gather8");
    return 0;
  }
  DISPATCH_V(V, maxv, CENET_DISPATCH(x_dtype, TI, CENET_DISPATCH(o_dtype, TO,
      (im2col_kernel<TI, TO, VV><<<ew_grid(total / VV), 256, 0, to_stream(s)>>>((const TI*)x, (TO*)out, B, H, W, Cin, KH, KW,
                                                                               stride, pad, Ho, Wo, Kpad)))));
  CENET_LAUNCH_CHECK("im2col");
  return 0;
}
