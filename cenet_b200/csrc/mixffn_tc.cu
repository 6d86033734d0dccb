// Mix-FFN tail in ONE kernel (pvtv2.py:40-47, 364-370):  t += fc2( GELU( dwconv3x3(h) + b_dw ) ) + b2
//   h [B,H,W,Ch] bf16 is the fc1 output, t [B*H*W, C] the fp32 residual stream of the stage, Ch = mlp_ratio * C.
// The unfused plan writes the depthwise result (Ch = 8 C channels wide) to HBM and reads it back as the A operand of the
// fc2 GEMM: 2 * M * Ch * 2 bytes of the block's 4 * M * Ch * 2.  Here a CTA owns TR whole image rows (TR * W <= 128 pixels
// = the M of one tcgen05 MMA) and walks the hidden channels in chunks of 64:
//   * 8 compute warps stage the chunk's TR + 2 input rows with cp.async (one chunk ahead), run the depthwise 3x3 + GELU in
//     packed fp32 (the arithmetic of dwconv3x3_staged_kernel) and write the bf16 result straight into a SWIZZLE_128B K-major
//     shared-memory tile -- the A operand of the MMA; it never exists in global memory;
//   * one more warp streams the matching [C x 64] slice of W2 by TMA and issues tcgen05.mma (128 x C x 16, kind::f16) into
//     a TMEM accumulator that lives across all chunks (fc2's K loop = the chunk loop);
//   * after the last chunk the compute warps read the accumulator (tcgen05.ld), add bias and the fp32 residual and store.
// Algorithmic bytes per launch: M*Ch*2 (h, halo rows come from L2) + 2*M*C*4 (t) + Ch*C*2 + 10*Ch*4.
#include "tc_ptx.cuh"
#include <atomic>

namespace {
using namespace tcx;
constexpr int MF_T = 256;                 // compute threads
constexpr int MF_NTH = MF_T + 32;         // + the TMA / MMA warp
constexpr int MF_PW = 4;                  // pixels per thread and pass
constexpr int MF_A_BYTES = 128 * 128;     // one A tile: 128 pixels x 64 channels bf16

struct MfParams {
  const bf16* h;
  float* t;
  const float* w9c;    // [9][Ch]
  const float* dwb;    // [Ch]
  const float* b2;     // [C] or NULL
  int H, W, Ch, C, TR, nch;
};

// tanh-form GELU on two channels (see dwconv.cu: below the bf16 rounding of the value it produces)
__device__ __forceinline__ f32x2 mf_gelu2(f32x2 x) {
  const f32x2 k0 = pk2(0.7978845608f, 0.7978845608f), k1 = pk2(0.0356774081f, 0.0356774081f), hf = pk2(0.5f, 0.5f);
  const f32x2 u = fmul2(x, ffma2(fmul2(x, x), k1, k0));
  float ua, ub; upk2(u, ua, ub);
  float ta, tb;
  asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(ua));
  asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(ub));
  const f32x2 hx = fmul2(x, hf);
  return ffma2(hx, pk2(ta, tb), hx);
}

__global__ void __launch_bounds__(MF_NTH, 2) mixffn_tail_kernel(const __grid_constant__ CUtensorMap tmW, const MfParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int W = p.W, H = p.H, C = p.C, Ch = p.Ch, TR = p.TR, nch = p.nch;
  const int b = blockIdx.y, h0 = blockIdx.x * TR;
  const int rowB = (W + 2) * 128;                     // one staged row: pixels -1 .. W of a 64-channel chunk
  const int ringB = (TR + 2) * rowB;
  const uint32_t sA = sbase;                          // [2] A tiles
  const uint32_t sW = sA + 2 * MF_A_BYTES;            // [2] W2 slices, C rows x 128 bytes
  const uint32_t sR = sW + 2 * C * 128;               // [2] input rings
  const uint32_t bar = sR + 2 * ringB;
  const uint32_t a_full = bar, a_empty = bar + 16, w_full = bar + 32, d_full = bar + 48, tmem_slot = bar + 56;

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int i = 0; i < 2; i++) { mbar_init(a_full + 8 * i, MF_T); mbar_init(a_empty + 8 * i, 1); mbar_init(w_full + 8 * i, 1); }
    mbar_init(d_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tcols = C <= 32 ? 32u : (C <= 64 ? 64u : (C <= 128 ? 128u : 256u));
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tcols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // halo columns (w = -1 and w = W) of both rings stay zero
  for (int i = tid; i < 2 * (TR + 2) * 16; i += MF_NTH) {
    const int slot = i >> 4, side = (i >> 3) & 1, ch = i & 7;
    const uint32_t a = sR + slot * rowB + (side ? (W + 1) * 128 : 0) + ch * 16;
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0u) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 8) {
    // ===================== W2 slices by TMA + MMA issue (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t wbytes = (uint32_t)C * 128u;
      for (int j = 0; j < 2 && j < nch; j++) {
        mbar_arrive_expect_tx(w_full + 8 * j, wbytes);
        tma_load_3d(sW + j * wbytes, &tmW, w_full + 8 * j, j * 64, 0, 0);
      }
      for (int j = 0; j < nch; j++) {
        const int s = j & 1, ph = (j >> 1) & 1;
        mbar_wait(w_full + 8 * s, ph);
        mbar_wait(a_full + 8 * s, ph);
        tc_fence_after();
        const uint64_t ad = desc_k(sA + s * MF_A_BYTES), bd = desc_k(sW + s * wbytes);
#pragma unroll
        for (int k = 0; k < 4; k++) umma_f16(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (j | k) != 0);
        umma_commit(a_empty + 8 * s);                               // frees A tile s and W2 slice s
        if (j + 2 < nch) {
          mbar_wait(a_empty + 8 * s, ph);
          mbar_arrive_expect_tx(w_full + 8 * s, wbytes);
          tma_load_3d(sW + s * wbytes, &tmW, w_full + 8 * s, (j + 2) * 64, 0, 0);
        }
      }
      umma_commit(d_full);
    }
  } else {
    // ===================== depthwise 3x3 + GELU -> A tiles; epilogue =====================
    const bf16* hb = p.h + (size_t)b * H * W * Ch;
    auto issue_chunk = [&](int j) {                                  // TR + 2 rows x W pixels x 128 bytes, 16 bytes per cp.async
      const uint32_t ring = sR + (j & 1) * ringB;
      const int total = (TR + 2) * W * 8;
      int pw = tid >> 3, r = 0;
      const int c8 = tid & 7;
      while (pw >= W) { pw -= W; r++; }
      for (int i = tid; i < total; i += MF_T) {
        const int hr = h0 - 1 + r;
        const bool valid = hr >= 0 && hr < H;
        const bf16* src = hb + ((size_t)(valid ? hr : 0) * W + pw) * Ch + j * 64 + c8 * 8;
        const uint32_t dst = ring + r * rowB + (pw + 1) * 128 + c8 * 16;
        const int nb = valid ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(nb) : "memory");
        pw += MF_T / 8;
        while (pw >= W) { pw -= W; r++; }
      }
      cp_async_commit();
    };
    issue_chunk(0);
    const int cv = tid & 15, pg0 = tid >> 4;                         // 16 channel quads x 16 pixel groups
    const int npg = W / MF_PW, nunit = TR * npg;
    for (int j = 0; j < nch; j++) {
      const int s = j & 1;
      // filter taps / bias of this thread's 4 channels of chunk j: requested before the waits below so that they overlap them
      const int c = j * 64 + cv * 4;
      f32x2 wv[9][2], bv[2];
#pragma unroll
      for (int t = 0; t < 9; t++) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.w9c + (size_t)t * Ch + c));
        wv[t][0] = pk2(w4.x, w4.y); wv[t][1] = pk2(w4.z, w4.w);
      }
      {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.dwb + c));
        bv[0] = pk2(b4.x, b4.y); bv[1] = pk2(b4.z, b4.w);
      }
      cp_async_wait<0>();                                            // chunk j has landed (this thread's part)
      asm volatile("bar.sync 1, 256;" ::: "memory");                 // ... everybody's; and ring s^1 is no longer read
      if (j + 1 < nch) issue_chunk(j + 1);
      if (j >= 2) mbar_wait(a_empty + 8 * s, ((j >> 1) - 1) & 1);    // the MMAs of chunk j-2 have read A tile s
      const uint32_t ring = sR + s * ringB + cv * 8;
      const uint32_t at = sA + s * MF_A_BYTES + (cv & 1) * 8;
      for (int u = pg0; u < nunit; u += 16) {
        const int tr = u / npg, pg = u - tr * npg;
        f32x2 acc[MF_PW][2];
#pragma unroll
        for (int q = 0; q < MF_PW; q++) { acc[q][0] = bv[0]; acc[q][1] = bv[1]; }
#pragma unroll
        for (int dh = 0; dh < 3; dh++) {
          const uint32_t ra = ring + (tr + dh) * rowB + pg * (MF_PW * 128);
#pragma unroll
          for (int q = 0; q < MF_PW + 2; q++) {
            unsigned w0, w1;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(ra + q * 128));
            const f32x2 x01 = bf2_to_f2(w0), x23 = bf2_to_f2(w1);
#pragma unroll
            for (int pp = 0; pp < MF_PW; pp++) {
              const int t = q - pp;
              if (t < 0 || t > 2) continue;
              acc[pp][0] = ffma2(x01, wv[dh * 3 + t][0], acc[pp][0]);
              acc[pp][1] = ffma2(x23, wv[dh * 3 + t][1], acc[pp][1]);
            }
          }
        }
        const int m0 = tr * W + pg * MF_PW;
#pragma unroll
        for (int pp = 0; pp < MF_PW; pp++) {
          const int m = m0 + pp;
          const f32x2 o0 = mf_gelu2(acc[pp][0]), o1 = mf_gelu2(acc[pp][1]);
          const uint32_t dst = at + (m >> 3) * 1024 + (m & 7) * 128 + ((((uint32_t)cv >> 1) ^ (uint32_t)(m & 7)) << 4);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(f2_to_bf2(o0)), "r"(f2_to_bf2(o1)) : "memory");
        }
      }
      fence_proxy_async();                                           // generic-proxy writes -> visible to the tensor core
      mbar_arrive(a_full + 8 * s);
    }
    // ---- epilogue: accumulator + bias + residual, fp32, in place ----
    mbar_wait(d_full, 0);
    tc_fence_after();
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const int ncol = C >> 1;                                         // columns of this warp (32 or 64)
    const bool ok = row < TR * W && h0 + row / W < H;
    float* tp = p.t + ((size_t)(b * H + h0) * W + row) * C + half * ncol;
    for (int c0 = 0; c0 < ncol; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * ncol + c0), v);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 r = *reinterpret_cast<const float4*>(tp + c0 + i);
          float4 bb = p.b2 ? *reinterpret_cast<const float4*>(p.b2 + half * ncol + c0 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          r.x += __uint_as_float(v[i]) + bb.x; r.y += __uint_as_float(v[i + 1]) + bb.y;
          r.z += __uint_as_float(v[i + 2]) + bb.z; r.w += __uint_as_float(v[i + 3]) + bb.w;
          *reinterpret_cast<float4*>(tp + c0 + i) = r;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tcols) : "memory");
  }
}
}  // namespace

// 1 when the fused kernel takes this shape (the caller keeps dwconv3x3 + linear otherwise)
extern "C" int cenet_mixffn_tail_supported(int H, int W, int Ch, int C) {
  if (W % MF_PW != 0 || W > 128 || W < 8 || H < 1) return 0;
  if (!(C == 64 || C == 128) || Ch % 64 != 0 || Ch < 64) return 0;
  const int TR = 128 / W;
  const size_t smem = 2 * MF_A_BYTES + 2 * (size_t)C * 128 + 2 * (size_t)(TR + 2) * (W + 2) * 128 + 64 + 1024;
  return smem <= 220 * 1024 ? 1 : 0;
}

extern "C" int cenet_mixffn_tail(const void* h, void* t, const float* w9c, const float* dw_bias, const void* w2, const float* b2,
                                 int B, int H, int W, int Ch, int C, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(h && t && w9c && dw_bias && w2, "cenet_mixffn_tail: null pointer");
  CENET_REQUIRE(cenet_mixffn_tail_supported(H, W, Ch, C), "cenet_mixffn_tail: shape H=%d W=%d Ch=%d C=%d not supported", H, W, Ch, C);
  CENET_REQUIRE(((uintptr_t)h & 15) == 0 && ((uintptr_t)t & 15) == 0 && ((uintptr_t)w2 & 15) == 0 &&
                (((uintptr_t)w9c | (uintptr_t)dw_bias | (uintptr_t)b2) & 15) == 0, "cenet_mixffn_tail: operands must be 16-byte aligned");
  CENET_REQUIRE(B <= 65535, "cenet_mixffn_tail: batch too large");
  MfParams p;
  p.h = (const bf16*)h; p.t = (float*)t; p.w9c = w9c; p.dwb = dw_bias; p.b2 = b2;
  p.H = H; p.W = W; p.Ch = Ch; p.C = C; p.TR = 128 / W; p.nch = Ch / 64;
  CUtensorMap tmW;
  if (encode3(&tmW, w2, Ch, C, 1, Ch, (long long)C * Ch, 64, C, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
  const size_t smem = 2 * MF_A_BYTES + 2 * (size_t)C * 128 + 2 * (size_t)(p.TR + 2) * (W + 2) * 128 + 64 + 1024;
  static std::atomic<size_t> configured{0};
  if (configured.load() < smem) {
    cudaFuncSetAttribute(mixffn_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    configured.store(220 * 1024);
  }
  dim3 grid(cdiv(H, p.TR), B);
  mixffn_tail_kernel<<<grid, MF_NTH, smem, to_stream(s)>>>(tmW, p);
  CENET_LAUNCH_CHECK("mixffn_tail");
  return 0;
}
