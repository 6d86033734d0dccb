// Mix-FFN tail in ONE kernel (pvtv2.py:40-47, 364-370):  t += fc2( GELU( dwconv3x3(h) + b_dw ) ) + b2
//   h [B,H,W,Ch] bf16 is the fc1 output, t [B*H*W, C] the fp32 residual stream of the stage, Ch = mlp_ratio * C.
// The unfused plan writes the depthwise result (Ch = 8 C channels wide) to HBM and reads it back as the A operand of the
// fc2 GEMM: 2 * M * Ch * 2 bytes of the block's 4 * M * Ch * 2.  Here a persistent CTA (one per SM) owns tiles of TR whole
// image rows (TR * W <= 128 pixels = the M of one tcgen05 MMA) and walks the hidden channels in chunks of 64:
//   * a producer warp streams, NR - 1 chunks ahead and across tile boundaries, the chunk's (TR + 2) x (W + 2) x 64 input box
//     (ONE 4-D TMA box per chunk: out-of-image rows and the two halo columns are the TMA's zero fill) and the matching
//     [C x 64] slice of W2;
//   * 16 compute warps run the depthwise 3x3 + GELU in packed fp32 (the arithmetic of dwconv3x3_staged_kernel) and write the
//     bf16 result straight into a SWIZZLE_128B K-major shared-memory tile -- the A operand of the MMA; it never exists in
//     global memory;
//   * one thread issues tcgen05.mma (128 x C x 16, kind::f16) into a TMEM accumulator that lives across the chunks of a tile
//     (fc2's K loop = the chunk loop);
//   * after the last chunk the compute warps read the accumulator (tcgen05.ld), add bias and the fp32 residual and store,
//     while the producer is already fetching the next tile.
// First version (one tile per CTA, two CTAs per SM, cp.async one chunk ahead): 336 us against 201 us unfused on B=64 56x56
// 512 -> 64 -- ~60 KB in flight per SM and a block-wide barrier per chunk left it latency-bound at 0.9 TB/s.
// Algorithmic bytes per launch: M*Ch*2 (h, halo rows come from L2) + 2*M*C*4 (t) + Ch*C*2 + 10*Ch*4.
#include "tc_ptx.cuh"
#include <atomic>

namespace {
using namespace tcx;
constexpr int MF_CW = 16;                 // compute warps
constexpr int MF_T = MF_CW * 32;          // compute threads
constexpr int MF_NTH = MF_T + 64;         // + the TMA producer warp + the MMA warp
constexpr int MF_PW = 4;                  // pixels per thread and pass
constexpr int MF_A_BYTES = 128 * 128;     // one A tile: 128 pixels x 64 channels bf16
constexpr int MF_MAXNR = 4;
constexpr int MF_MAXNA = 2;               // A tiles between the depthwise warps and the MMAs (3-4 measured slower: they cost ring depth)

struct MfParams {
  float* t;
  const float* w9c;    // [9][Ch]
  const float* dwb;    // [Ch]
  const float* b2;     // [C] or NULL
  int H, W, Ch, C, TR, nch, NR, NA, bands, ntiles;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

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

__global__ void __launch_bounds__(MF_NTH, 1) mixffn_tail_kernel(const __grid_constant__ CUtensorMap tmH,
                                                                const __grid_constant__ CUtensorMap tmW, const MfParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int W = p.W, H = p.H, C = p.C, Ch = p.Ch, TR = p.TR, nch = p.nch, NR = p.NR, NA = p.NA;
  const int rowB = (W + 2) * 128;                     // one staged row: pixels -1 .. W of a 64-channel chunk
  const int ringB = (TR + 2) * rowB;                  // = the bytes of one TMA box
  const uint32_t wbytes = (uint32_t)C * 128u;
  const uint32_t sA = sbase;                          // [NA] A tiles
  const uint32_t sW = sA + NA * MF_A_BYTES;            // [NR] W2 slices, C rows x 128 bytes
  const uint32_t sR = sW + NR * wbytes;               // [NR] input boxes
  const uint32_t sF = sR + NR * ringB;                // depthwise filter + bias of ALL hidden channels: [10][Ch] fp32
  const uint32_t bar = sF + 10 * Ch * 4;
  const uint32_t h_full = bar, h_empty = bar + 8 * MF_MAXNR, w_full = bar + 16 * MF_MAXNR, w_empty = bar + 24 * MF_MAXNR;
  const uint32_t a_full = bar + 32 * MF_MAXNR, a_empty = a_full + 8 * MF_MAXNA, d_full = a_empty + 8 * MF_MAXNA, tmem_slot = d_full + 16;

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmH) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int i = 0; i < MF_MAXNR; i++) { mbar_init(h_full + 8 * i, 1); mbar_init(h_empty + 8 * i, MF_CW); mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
    for (int i = 0; i < MF_MAXNA; i++) { mbar_init(a_full + 8 * i, MF_CW); mbar_init(a_empty + 8 * i, 1); }
    mbar_init(d_full, 1); mbar_init(d_full + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tcols = C <= 64 ? 128u : 256u;          // two accumulators (tile parity) of C fp32 columns
  if (warp == MF_CW + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tcols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the persistent CTA keeps the depthwise filter in shared memory (the per-chunk reload from L2 stalled every warp at once)
  for (int i = tid; i < 10 * Ch / 4; i += MF_NTH) {
    const float4 v = i < 9 * Ch / 4 ? __ldg(reinterpret_cast<const float4*>(p.w9c) + i) : __ldg(reinterpret_cast<const float4*>(p.dwb) + (i - 9 * Ch / 4));
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sF + i * 16), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == MF_CW) {
    // ===================== TMA producer: input boxes and W2 slices, NR - 1 chunks ahead =====================
    if (lane == 0) {
      int g = 0;
      for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x) {
        const int b = ti / p.bands, h0 = (ti - b * p.bands) * TR;
        for (int j = 0; j < nch; j++, g++) {
          const int slot = g % NR, use = g / NR;
          if (g >= NR) {
            mbar_wait(h_empty + 8 * slot, (use - 1) & 1);                       // the compute warps have read the box of chunk g - NR
            mbar_wait(w_empty + 8 * slot, (use - 1) & 1);                       // ... and its MMAs have read its W2 slice
          }
          mbar_arrive_expect_tx(h_full + 8 * slot, (uint32_t)ringB);
          tma_load_4d(sR + slot * ringB, &tmH, h_full + 8 * slot, j * 64, -1, h0 - 1, b);
          mbar_arrive_expect_tx(w_full + 8 * slot, wbytes);
          tma_load_3d(sW + slot * wbytes, &tmW, w_full + 8 * slot, j * 64, 0, 0);
        }
      }
    }
  } else if (warp == MF_CW + 1) {
    // ===================== MMA issue (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int g = 0, it = 0;
      for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x, it++) {
        const uint32_t acc = tmem_base + (uint32_t)((it & 1) * C);
        for (int j = 0; j < nch; j++, g++) {
          const int s = g % NA, slot = g % NR;
          mbar_wait(w_full + 8 * slot, (g / NR) & 1);
          mbar_wait(a_full + 8 * s, (g / NA) & 1);
          tc_fence_after();
          const uint64_t ad = desc_k(sA + s * MF_A_BYTES), bd = desc_k(sW + slot * wbytes);
#pragma unroll
          for (int k = 0; k < 4; k++) umma_f16(acc, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (j | k) != 0);
          umma_commit(a_empty + 8 * s);                             // frees A tile s ...
          umma_commit(w_empty + 8 * slot);                          // ... and the W2 slice
        }
        umma_commit(d_full + 8 * (it & 1));
      }
    }
  } else {
    // ===================== depthwise 3x3 + GELU -> A tiles; epilogue =====================
    const int cv = tid & 15, pg0 = tid >> 4;                         // 16 channel quads x 32 pixel groups
    const int npg = W / MF_PW, nunit = TR * npg;
    const int quarter = warp & 3, part = warp >> 2;                  // epilogue: TMEM lane quarter, column quarter
    const int row = quarter * 32 + lane;
    const int ncol = C >> 2;                                         // accumulator columns of this warp (16 or 32)
    // this thread's units (pixel groups of 4 in a tile row): ring offset of the top-left input pixel and first A row
    constexpr int MAXU = 2;                                          // nunit = TR * W / 4 <= 32 * MAXU (TR * W <= 128)
    uint32_t uro[MAXU]; int um0[MAXU];
#pragma unroll
    for (int k = 0; k < MAXU; k++) {
      const int u = pg0 + k * (MF_T / 16);
      const int tr = u / npg, pg = u - tr * npg;
      uro[k] = (uint32_t)(tr * rowB + pg * (MF_PW * 128));
      um0[k] = u < nunit ? tr * W + pg * MF_PW : -1;
    }
    // ---- epilogue of a tile: accumulator + bias + residual, fp32, in place (16 columns at a time).  It runs one chunk into
    //      the NEXT tile (two TMEM accumulators alternate), so neither the last MMAs nor the residual loads are waited for ----
    auto epilogue = [&](int b, int h0, int e) {
      const bool ok = row < TR * W && h0 + row / W < H;
      float* tp = p.t + ((size_t)(b * H + h0) * W + row) * C + part * ncol;
      float4 r4[4];
#pragma unroll
      for (int i = 0; i < 4; i++) r4[i] = ok ? *reinterpret_cast<const float4*>(tp + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      mbar_wait(d_full + 8 * (e & 1), (e >> 1) & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < ncol; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((e & 1) * C + part * ncol + c0), v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            float4 r = c0 == 0 ? r4[i] : *reinterpret_cast<const float4*>(tp + c0 + 4 * i);
            const float4 bb = p.b2 ? __ldg(reinterpret_cast<const float4*>(p.b2 + part * ncol + c0 + 4 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
            r.x += __uint_as_float(v[4 * i]) + bb.x; r.y += __uint_as_float(v[4 * i + 1]) + bb.y;
            r.z += __uint_as_float(v[4 * i + 2]) + bb.z; r.w += __uint_as_float(v[4 * i + 3]) + bb.w;
            *reinterpret_cast<float4*>(tp + c0 + 4 * i) = r;
          }
        }
      }
      tc_fence_before();                  // the MMAs that next overwrite this accumulator are ordered after these loads through the
    };                                    // a_full arrivals that follow in program order
    int pb = 0, ph0 = 0;
    int g = 0, it = 0;
    for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x, it++) {
      const int b = ti / p.bands, h0 = (ti - b * p.bands) * TR;
      for (int j = 0; j < nch; j++, g++) {
        const int s = g % NA, slot = g % NR;
        // filter taps / bias of this thread's 4 channels of chunk j
        const uint32_t fa = sF + (uint32_t)(j * 64 + cv * 4) * 4;
        f32x2 wv[9][2], bv[2];
#pragma unroll
        for (int t = 0; t < 10; t++) {
          float4 w4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w4.x), "=f"(w4.y), "=f"(w4.z), "=f"(w4.w) : "r"(fa + (uint32_t)(t * Ch * 4)));
          if (t < 9) { wv[t][0] = pk2(w4.x, w4.y); wv[t][1] = pk2(w4.z, w4.w); }
          else { bv[0] = pk2(w4.x, w4.y); bv[1] = pk2(w4.z, w4.w); }
        }
        mbar_wait(h_full + 8 * slot, (g / NR) & 1);                  // the chunk's input box has landed
        if (g >= NA) mbar_wait(a_empty + 8 * s, (g / NA - 1) & 1);   // the MMAs of chunk g - NA have read A tile s
        const uint32_t ring = sR + slot * ringB + cv * 8;
        const uint32_t at = sA + s * MF_A_BYTES + (cv & 1) * 8;
#pragma unroll
        for (int k = 0; k < MAXU; k++) {
          if (um0[k] < 0) continue;
          f32x2 acc[MF_PW][2];
#pragma unroll
          for (int q = 0; q < MF_PW; q++) { acc[q][0] = bv[0]; acc[q][1] = bv[1]; }
#pragma unroll
          for (int dh = 0; dh < 3; dh++) {
            const uint32_t ra = ring + uro[k] + dh * rowB;
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
          const int m0 = um0[k];
#pragma unroll
          for (int pp = 0; pp < MF_PW; pp++) {
            const int m = m0 + pp;
            const f32x2 o0 = mf_gelu2(acc[pp][0]), o1 = mf_gelu2(acc[pp][1]);
            const uint32_t dst = at + (m >> 3) * 1024 + (m & 7) * 128 + ((((uint32_t)cv >> 1) ^ (uint32_t)(m & 7)) << 4);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(f2_to_bf2(o0)), "r"(f2_to_bf2(o1)) : "memory");
          }
        }
        fence_proxy_async();                                         // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) { mbar_arrive(a_full + 8 * s); mbar_arrive(h_empty + 8 * slot); }
        if (j == 0 && it > 0) epilogue(pb, ph0, it - 1);
      }
      pb = b; ph0 = h0;                                              // epilogue deferred: after chunk 0 of the next tile
    }
    if (it > 0) epilogue(pb, ph0, it - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MF_CW + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tcols) : "memory");
  }
}

// pipeline depths / shared-memory footprint of a shape: NR input boxes + W2 slices in flight, NA A tiles between the depthwise
// warps and the MMAs (how far a fast warp may run ahead of the slowest).  Returns NR (0 when nothing fits).
inline int mf_plan(int W, int C, int Ch, size_t* smem_out, int* na_out = nullptr) {
  const int TR = 128 / W;
  const size_t ringB = (size_t)(TR + 2) * (W + 2) * 128, wbytes = (size_t)C * 128;
  const size_t fixed = (size_t)10 * Ch * 4 + 512 + 1024, cap = 220 * 1024;
  for (int pass = 0; pass < 2; pass++)                          // first: at least 3 boxes in flight, as many A tiles as fit
    for (int na = MF_MAXNA; na >= 2; na--) {
      if (fixed + (size_t)na * MF_A_BYTES >= cap) continue;
      int nr = (int)((cap - fixed - (size_t)na * MF_A_BYTES) / (ringB + wbytes));
      if (nr > MF_MAXNR) nr = MF_MAXNR;
      if (nr >= (pass == 0 ? 3 : 2)) {
        if (smem_out) *smem_out = fixed + (size_t)na * MF_A_BYTES + (size_t)nr * (ringB + wbytes);
        if (na_out) *na_out = na;
        return nr;
      }
    }
  return 0;
}
}  // namespace

// 1 when the fused kernel takes this shape (the caller keeps dwconv3x3 + linear otherwise)
extern "C" int cenet_mixffn_tail_supported(int H, int W, int Ch, int C) {
  if (W % MF_PW != 0 || W > 128 || W < 8 || H < 1) return 0;
  if (!(C == 64 || C == 128) || Ch % 64 != 0 || Ch < 64) return 0;
  return mf_plan(W, C, Ch, nullptr) >= 2 ? 1 : 0;
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
  p.t = (float*)t; p.w9c = w9c; p.dwb = dw_bias; p.b2 = b2;
  p.H = H; p.W = W; p.Ch = Ch; p.C = C; p.TR = 128 / W; p.nch = Ch / 64;
  size_t smem = 0;
  p.NR = mf_plan(W, C, Ch, &smem, &p.NA);
  p.bands = cdiv(H, p.TR);
  p.ntiles = p.bands * B;
  CUtensorMap tmW, tmH;
  if (encode3(&tmW, w2, Ch, C, 1, Ch, (long long)C * Ch, 64, C, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
  {
    EncodeTiledFn enc = get_encode();
    CENET_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)Ch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)Ch * 2, (cuuint64_t)W * Ch * 2, (cuuint64_t)H * W * Ch * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(W + 2), (cuuint32_t)(p.TR + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmH, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(h), dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CENET_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (mixffn_tail input) failed with CUresult %d", (int)r);
  }
  static std::atomic<int> configured{0};
  if (!configured.exchange(1)) cudaFuncSetAttribute(mixffn_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
  mixffn_tail_kernel<<<grid, MF_NTH, smem, to_stream(s)>>>(tmH, tmW, p);
  CENET_LAUNCH_CHECK("mixffn_tail");
  return 0;
}
