// Depthwise 3x3 family on channels-last tensors (HBM-bound).
// Thread = one pixel x V channels (16 bytes of bf16); a block covers PT consecutive pixels of one image row x CVT
// channel vectors, so a warp reads/writes contiguous channel runs and the 3x3 neighbourhood is served by L1/L2 (each
// input element comes from HBM once).  Index math is 32-bit with one division per thread.
// Algorithmic bytes: B*H*W*C*(sizeof(in)+sizeof(out)) (+ 9*C weights).
#include "common.cuh"
#include <algorithm>
#include <atomic>
#include <cstdlib>

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

// ------------------------------------------------------------------------------------------------ staged variant
// Shared-memory staged depthwise 3x3 (dilation 1, bf16, C % 64 == 0, W % PW == 0): the Mix-FFN / CFAM-Mlp case.
// A CTA owns 64 channels (128 B per pixel) of a band of image rows and streams the rows through a ring of NR row
// buffers with cp.async (16 B per thread, zero-fill for rows outside the image; the w = -1 / w = W halo columns are
// zeroed once): the loads of rows h+2, h+3 are in flight while row h is computed and cost no registers.  ncu on the
// first version of this kernel (profiles/r1_dwconv_ncu.md): memory stalls gone, but 62 instructions per output element,
// 60 % of them integer / branch overhead -> this version is written for instruction count:
//   * thread = 4 channels x PW consecutive pixels, its 36 filter taps live in registers as packed float pairs;
//   * all arithmetic is packed fp32 (FFMA2 / FMUL2, sm_100 `fma.rn.f32x2`): 2 channels per instruction;
//   * shared-memory reads are `ld.shared.v2.b32` from one 32-bit base with immediate offsets, no per-load address math;
//   * no bounds predicates in the inner loops (W % PW == 0, the halo columns exist in shared memory).
constexpr int SG_C = 64, SG_T = 256, SG_PF = 2, SG_NR = 3 + SG_PF;
// tanh-form GELU on two channels: 0.5 x (1 + tanh(x (k0 + k1 x^2)))
__device__ __forceinline__ f32x2 gelu2(f32x2 x) {
  const f32x2 k0 = pk2(0.7978845608f, 0.7978845608f), k1 = pk2(0.0356774081f, 0.0356774081f), hf = pk2(0.5f, 0.5f);
  const f32x2 u = fmul2(x, ffma2(fmul2(x, x), k1, k0));
  float ua, ub; upk2(u, ua, ub);
  float ta, tb;
  asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(ua));
  asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(ub));
  const f32x2 hx = fmul2(x, hf);
  return ffma2(hx, pk2(ta, tb), hx);
}

template <int PW>
__global__ void __launch_bounds__(SG_T, 3) dwconv3x3_staged_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ y, int ldy,
                                                                   const float* __restrict__ w9c, const float* __restrict__ bias,
                                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                                   int H, int W, int C, int act, float slope, bf16* __restrict__ zout,
                                                                   int rows_per_cta) {
  extern __shared__ __align__(16) unsigned char sg_smem[];
  const int tid = threadIdx.x;
  const int rowB = (W + 2) * 128;                              // bytes of one staged row: pixels -1 .. W, 64 channels
  const int c_base = blockIdx.x * SG_C;
  const int h0 = blockIdx.y * rows_per_cta, h1 = min(H, h0 + rows_per_cta);
  const int b = blockIdx.z;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sg_smem);
  for (int i = tid; i < SG_NR * 16; i += SG_T) {              // halo columns of every ring slot
    const int slot = i >> 4, side = (i >> 3) & 1, ch = i & 7;
    *reinterpret_cast<uint4*>(sg_smem + slot * rowB + (side ? (W + 1) * 128 : 0) + ch * 16) = make_uint4(0, 0, 0, 0);
  }
  // cp.async plan of this thread: chunk i = tid + 256 j covers pixel i / 8, channels (i % 8) * 8 .. + 8
  const char* xsrc = reinterpret_cast<const char*>(x + (size_t)b * H * W * ldx + c_base) + (size_t)(tid >> 3) * ldx * 2 + (tid & 7) * 16;
  const size_t src_step = (size_t)32 * ldx * 2, row_step = (size_t)W * ldx * 2;
  const int nchunk = (W * 8 - tid + SG_T - 1) / SG_T;
  auto issue_row = [&](int r, int slot) {                     // r may lie outside the image (zero row) or the band (skipped)
    if (r <= h1) {
      const bool valid = r >= 0 && r < H;
      const char* src = xsrc + (valid ? r : 0) * row_step;
      const unsigned dst = sbase + slot * rowB + 128 + tid * 16;
      const int nb = valid ? 16 : 0;
      for (int j = 0; j < nchunk; j++)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst + j * 4096), "l"(src + j * src_step), "r"(nb) : "memory");
    }
    cp_async_commit();
  };
#pragma unroll
  for (int k = 0; k < SG_PF + 2; k++) issue_row(h0 - 1 + k, k);   // rows h0-1 .. h0+PF -> slots 0 .. PF+1
  const int cv = tid & 15, pg0 = tid >> 4;                     // 16 channel quads x 16 pixel groups
  const int c = c_base + cv * 4;
  const int npg = W / PW;
  f32x2 wv[9][2], bv[2], sc[2], sh[2];
#pragma unroll
  for (int t = 0; t < 9; t++) {
    const float4 w4 = *reinterpret_cast<const float4*>(w9c + (size_t)t * C + c);
    wv[t][0] = pk2(w4.x, w4.y); wv[t][1] = pk2(w4.z, w4.w);
  }
  {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f), o4 = make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 b4 = bias ? *reinterpret_cast<const float4*>(bias + c) : z4;
    const float4 s4 = scale ? *reinterpret_cast<const float4*>(scale + c) : o4;
    const float4 t4 = scale ? *reinterpret_cast<const float4*>(shift + c) : z4;
    bv[0] = pk2(b4.x, b4.y); bv[1] = pk2(b4.z, b4.w);
    sc[0] = pk2(s4.x, s4.y); sc[1] = pk2(s4.z, s4.w);
    sh[0] = pk2(t4.x, t4.y); sh[1] = pk2(t4.z, t4.w);
  }
  const bool affine = scale != nullptr;
  const size_t ystride = (size_t)ldy * 2, zstride = (size_t)C * 2;
  int slot_top = 0;                                            // ring slot of row h-1
  for (int h = h0; h < h1; h++) {
    cp_async_wait<SG_PF - 1>();                                // row h+1 has landed (this thread's part)
    __syncthreads();                                           // ... everybody's part; and row h-2's slot is free
    {
      int slot_new = slot_top + SG_NR - 1;                     // slot of row h-2 == slot of row h+1+PF
      if (slot_new >= SG_NR) slot_new -= SG_NR;
      issue_row(h + 1 + SG_PF, slot_new);
    }
    unsigned rows[3];
#pragma unroll
    for (int dh = 0; dh < 3; dh++) {
      int sl = slot_top + dh;
      if (sl >= SG_NR) sl -= SG_NR;
      rows[dh] = sbase + sl * rowB + cv * 8;
    }
    const size_t pix0 = (size_t)(b * H + h) * W;
    for (int pg = pg0; pg < npg; pg += 16) {
      f32x2 acc[PW][2];
#pragma unroll
      for (int p = 0; p < PW; p++) { acc[p][0] = bv[0]; acc[p][1] = bv[1]; }
#pragma unroll
      for (int dh = 0; dh < 3; dh++) {
        const unsigned ra = rows[dh] + pg * (PW * 128);
#pragma unroll
        for (int q = 0; q < PW + 2; q++) {
          unsigned w0, w1;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(ra + q * 128));   // base + immediate in SASS
          const f32x2 x01 = bf2_to_f2(w0), x23 = bf2_to_f2(w1);
#pragma unroll
          for (int p = 0; p < PW; p++) {
            const int t = q - p;
            if (t < 0 || t > 2) continue;
            acc[p][0] = ffma2(x01, wv[dh * 3 + t][0], acc[p][0]);
            acc[p][1] = ffma2(x23, wv[dh * 3 + t][1], acc[p][1]);
          }
        }
      }
      char* yp = reinterpret_cast<char*>(y + (pix0 + pg * PW) * ldy + c);
      char* zp = zout ? reinterpret_cast<char*>(zout + (pix0 + pg * PW) * C + c) : nullptr;
#pragma unroll
      for (int p = 0; p < PW; p++) {
        f32x2 o0 = acc[p][0], o1 = acc[p][1];
        if (affine) { o0 = ffma2(o0, sc[0], sh[0]); o1 = ffma2(o1, sc[1], sh[1]); }
        if (zp) *reinterpret_cast<uint2*>(zp + p * zstride) = make_uint2(f2_to_bf2(o0), f2_to_bf2(o1));   // pre-activation
        if (act == CENET_ACT_GELU) { o0 = gelu2(o0); o1 = gelu2(o1); }
        else if (act != CENET_ACT_NONE) {
          float a0, a1, a2, a3; upk2(o0, a0, a1); upk2(o1, a2, a3);
          o0 = pk2(apply_act(a0, act, slope), apply_act(a1, act, slope));
          o1 = pk2(apply_act(a2, act, slope), apply_act(a3, act, slope));
        }
        *reinterpret_cast<uint2*>(yp + p * ystride) = make_uint2(f2_to_bf2(o0), f2_to_bf2(o1));
      }
    }
    if (++slot_top == SG_NR) slot_top = 0;
  }
  cp_async_wait<0>();
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
  static const bool use_staged = getenv("CENET_B200_DW_STAGED") == nullptr || atoi(getenv("CENET_B200_DW_STAGED")) != 0;
  // pixels per thread: 16 pixel groups per CTA pass should cover the row (56 -> 4, 28 -> 2, 14 -> 1, 128 -> 4 in two passes)
  const int pw = (W % 4 == 0 && W >= 48) ? 4 : (W % 2 == 0 && W >= 24) ? 2 : (W >= 12 ? 1 : 0);
  if (use_staged && pw && x_dtype == CENET_BF16 && y_dtype == CENET_BF16 && dil == 1 && !up2 && C % SG_C == 0 && ldx % 8 == 0 &&
      ldy % 4 == 0 && (((uintptr_t)x & 15) == 0) && (((uintptr_t)y & 7) == 0) && (!zout || ((uintptr_t)zout & 7) == 0) &&
      (((uintptr_t)w9c | (uintptr_t)bias | (uintptr_t)scale | (uintptr_t)shift) & 15) == 0 &&
      (size_t)SG_NR * (W + 2) * 128 <= 160 * 1024 && B <= 65535) {
    // band height: whole image if that already fills the machine, otherwise halve until ~4 CTAs per SM exist
    int rpc = H;
    while ((long long)(C / SG_C) * cdiv(H, rpc) * B < 4LL * kNumSMs && rpc > 7) rpc = (rpc + 1) / 2;
    const size_t smem = (size_t)SG_NR * (W + 2) * 128;
    dim3 grid(C / SG_C, cdiv(H, rpc), B);
#define LAUNCH_SG(PWV)                                                                                                       \
    do {                                                                                                                       \
      static std::atomic<size_t> configured{0};                                                                                \
      if (smem > 48 * 1024 && configured.load() < smem) {                                                                      \
        cudaFuncSetAttribute(dwconv3x3_staged_kernel<PWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);          \
        configured.store(160 * 1024);                                                                                          \
      }                                                                                                                        \
      dwconv3x3_staged_kernel<PWV><<<grid, SG_T, smem, to_stream(s)>>>((const bf16*)x, (int)ldx, (bf16*)y, (int)ldy, w9c, bias, scale, \
                                                                      shift, H, W, C, act, slope, (bf16*)zout, rpc);           \
    } while (0)
    if (pw == 4) LAUNCH_SG(4); else if (pw == 2) LAUNCH_SG(2); else LAUNCH_SG(1);
#undef LAUNCH_SG
    CENET_LAUNCH_CHECK("dwconv3x3_staged");
    return 0;
  }
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
