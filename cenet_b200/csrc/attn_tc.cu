// tcgen05 / TMEM / TMA flash attention for sm_100a, head width D in {64, 128}:   O = softmax(Q K^T * scale) V
//
// Serves the two single-softmax attentions of the network whose contraction is deep enough for the 5th-gen tensor core:
//   * the non-local block of the CFAM decoder (nlb.py:116-137): one head, d = C (64 at 56x56 with N = 3136 keys, 128 at 28x28)
//   * the encoder's spatial-reduction attention (pvtv2.py:88-105): heads of 64, 49 reduced keys (256 at 512x512)
// One CTA = 128 query rows of one (image, head); 192 threads, FlashAttention-style online softmax, nothing N x N in HBM:
//   warp 0     : TMA producer.  Q once ([128 x D] as D/64 SWIZZLE_128B boxes), then a 3-stage ring of 64-key K and V tiles
//                (3-D tensor maps {columns, rows, image}: rows past the end of an image are zero-filled, never the next image).
//   warp 1     : single-thread tcgen05.mma issue.  S_j = Q K_j^T (M=128, N=64, K=D; both operands K-major) into one of two TMEM
//                score buffers, issued one tile AHEAD of the softmax; O += P_j V_j (M=128, N=D, K=64; A = P from shared memory,
//                B = the V tile as it lies in memory = MN-major) into the TMEM output accumulator.  tcgen05.commit releases
//                the K/V stage and the P buffer and tells the softmax warps that O is quiescent.
//   warps 2..5 : softmax, one thread per query row (TMEM lane = row): tcgen05.ld the 64 scores, running max in log2 units with
//                LAZY rescaling (the max -- and with it O in TMEM and the row sum -- is only moved when it grows by more than
//                2^8, which fp32 sums and bf16 probabilities absorb exactly; otherwise O is never touched), exp2 on the MUFU
//                pipe, P written as bf16 into the swizzled K-major layout the MMA reads, fence.proxy.async, arrive.
//                At the end: tcgen05.ld O, multiply by 1/rowsum, 16-byte stores.
// D = 64: 96 KB of shared memory and 256 TMEM columns per CTA -> two CTAs per SM, so one CTA's softmax overlaps the other's
// MMAs/loads.  The kernel is bound by the exponentials (N^2 per image on 16 MUFU lanes/clk/SM), not by the tensor pipe.
#include "tc_ptx.cuh"
#include <cstdlib>

namespace {
using namespace tcx;
constexpr int QT = 128;    // query rows per CTA = TMEM lanes
constexpr int KT = 64;     // keys per tile
constexpr int NTH = 192;
constexpr int STAGES = 3;

template <int D>
struct Cfg {
  static constexpr int KB = D / 64;                 // 64-column blocks of Q / K / V rows
  static constexpr int Q_BYTES = KB * QT * 128;
  static constexpr int K_BYTES = KB * KT * 128;
  static constexpr int V_BYTES = KB * KT * 128;
  static constexpr int STAGE = K_BYTES + V_BYTES;
  static constexpr int P_BYTES = QT * KT * 2;
  static constexpr int TMEM_COLS = 256;             // 2 x 64 score columns + D (<= 128) output columns
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM = Q_BYTES + STAGES * STAGE + 2 * P_BYTES + BAR_BYTES + 1024;
};

struct TcAttnParams {
  bf16* o;
  float* lse;                // optional [B, heads, Nq]: natural-log-sum-exp of the scaled scores (training forward)
  long long ldo, bo;
  int Nq, Nk, heads;
  float scale_log2, lse_mul;  // lse_mul: ln 2 (natural-log LSE) or 1 (log2 units)
};

struct SoftmaxBars {
  uint32_t s_full, s_empty, p_full, p_empty, o_ready, o_final;
};

// Softmax role of one warp (warps 2..5 of the CTA): one thread per query row, TMEM lane = row.  Scores in TMEM columns
// [0,64) / [64,128) (two buffers), output accumulator of DOUT columns at column 128; P tiles [2][128 x 64] bf16 at sP.
template <int DOUT>
__device__ __forceinline__ void softmax_role(uint32_t tmem_base, uint32_t sP, const SoftmaxBars& B_, int nt, const TcAttnParams& p, int warp,
                                             int lane, int q0, int b, int out_col0, int lse_map) {
  constexpr int P_BYTES = QT * KT * 2;
  const uint32_t s_full = B_.s_full, s_empty = B_.s_empty, p_full = B_.p_full, p_empty = B_.p_empty, o_ready = B_.o_ready,
                 o_final = B_.o_final;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t t_o = tmem_base + lane_off + 2 * KT;
    float m = -INFINITY, l = 0.f;
    const float sc = p.scale_log2;
    for (int j = 0; j < nt; j++) {
      mbar_wait(s_full + 8 * (j & 1), (j >> 1) & 1);
      tc_fence_after();
      uint32_t sv[64];
      {
        uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[0]);
        uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[32]);
        tmem_ld32(tmem_base + lane_off + (j & 1) * KT, lo);
        tmem_ld32(tmem_base + lane_off + (j & 1) * KT + 32, hi);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty + 8 * (j & 1));             // the MMA warp may overwrite this score buffer
      const int nvalid = p.Nk - j * KT;
      if (nvalid < KT) {                                             // ragged last tile (TMA zero-filled the missing keys)
#pragma unroll
        for (int c = 0; c < KT; c++)
          if (c >= nvalid) sv[c] = 0xff800000u;                       // -inf
      }
      float mx0 = __uint_as_float(sv[0]), mx1 = __uint_as_float(sv[1]), mx2 = __uint_as_float(sv[2]), mx3 = __uint_as_float(sv[3]);
#pragma unroll
      for (int c = 4; c < KT; c += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(sv[c])); mx1 = fmaxf(mx1, __uint_as_float(sv[c + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(sv[c + 2])); mx3 = fmaxf(mx3, __uint_as_float(sv[c + 3]));
      }
      const float m_new = fmaxf(m, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc);
      if (j == 0) {
        m = m_new;
      } else if (__any_sync(0xffffffffu, m_new > m + 8.f)) {
        // lazy rescale (warp-uniform branch: the TMEM accesses are warp-wide): move this warp's rows to their new maxima
        const float corr = ex2(m - m_new);
        m = m_new;
        l *= corr;
        mbar_wait(o_ready, (j - 1) & 1);                             // P(j-1) V(j-1) has landed in O
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < DOUT; c0 += 32) {
          uint32_t o[32];
          tmem_ld32(t_o + c0, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
          tmem_st32(t_o + c0, o);
        }
        tmem_st_wait();
      }
      mbar_wait(p_empty + 8 * (j & 1), ((j >> 1) & 1) ^ 1);          // P(j-2) V(j-2) no longer reads this P buffer
      const uint32_t prow = sP + (j & 1) * P_BYTES + row * 128;
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      const float nm = -m;
#pragma unroll
      for (int c = 0; c < KT / 8; c++) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = ex2(fmaf(__uint_as_float(sv[c * 8 + i]), sc, nm));
        l0 += e[0] + e[4]; l1 += e[1] + e[5]; l2 += e[2] + e[6]; l3 += e[3] + e[7];
        const uint32_t w0 = pack2(e[0], e[1]), w1 = pack2(e[2], e[3]), w2 = pack2(e[4], e[5]), w3 = pack2(e[6], e[7]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (uint32_t)((c ^ (row & 7)) << 4)), "r"(w0), "r"(w1),
                     "r"(w2), "r"(w3)
                     : "memory");
      }
      l += (l0 + l1) + (l2 + l3);
      fence_proxy_async();                                           // generic-proxy stores -> visible to the tensor core
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full + 8 * (j & 1));
    }
    // ---- epilogue: O / rowsum -> bf16 ----
    mbar_wait(o_final, 0);
    tc_fence_after();
    const float inv = 1.f / l;
    const int n = q0 + row;
    bf16* orow = p.o + (long long)b * p.bo + (long long)n * p.ldo + out_col0;
#pragma unroll
    for (int c0 = 0; c0 < DOUT; c0 += 32) {
      uint32_t o[32];
      tmem_ld32(t_o + c0, o);
      tmem_ld_wait();
      if (n < p.Nq) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] = __uint_as_float(o[g * 8 + i]) * inv;
          stv<8>(orow + c0 + g * 8, v);
        }
      }
    }
    if (p.lse != nullptr && n < p.Nq)
      p.lse[((long long)b * p.heads + lse_map) * p.Nq + n] = (m + log2f(l)) * p.lse_mul;
}

template <int D>
__global__ void __launch_bounds__(NTH, D == 64 ? 2 : 1) attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                      const __grid_constant__ CUtensorMap tmK,
                                                                      const __grid_constant__ CUtensorMap tmV,
                                                                      const TcAttnParams p) {
  using C = Cfg<D>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * QT;
  const uint32_t sQ = sbase;
  const uint32_t sKV = sQ + C::Q_BYTES;
  const uint32_t sP = sKV + STAGES * C::STAGE;
  const uint32_t bar = sP + 2 * C::P_BYTES;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8 + s * 8; };
  auto kv_empty = [&](int s) { return bar + 8 + (STAGES + s) * 8; };
  const uint32_t s_full = bar + 8 + 2 * STAGES * 8;      // [2]
  const uint32_t s_empty = s_full + 16;                  // [2]
  const uint32_t p_full = s_empty + 16;                  // [2]
  const uint32_t p_empty = p_full + 16;                  // [2]
  const uint32_t o_ready = p_empty + 16;                 // completes once per key tile (P_j V_j retired)
  const uint32_t o_final = o_ready + 8;                  // completes once, after the last tile (the per-tile barrier's parity
  const uint32_t tmem_slot = o_final + 8;                // could alias two tiles back at the epilogue)
  const int nt = (p.Nk + KT - 1) / KT;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; s++) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int i = 0; i < 2; i++) {
      mbar_init(s_full + 8 * i, 1); mbar_init(s_empty + 8 * i, 4);
      mbar_init(p_full + 8 * i, 4); mbar_init(p_empty + 8 * i, 1);
    }
    mbar_init(o_ready, 1);
    mbar_init(o_final, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int kb = 0; kb < C::KB; kb++) tma_load_3d(sQ + kb * QT * 128, &tmQ, q_full, head * D + kb * 64, q0, b);
      for (int j = 0; j < nt; j++) {
        const int s = j % STAGES;
        mbar_wait(kv_empty(s), ((j / STAGES) & 1) ^ 1);
        const uint32_t sk = sKV + s * C::STAGE, sv = sk + C::K_BYTES;
        mbar_arrive_expect_tx(kv_full(s), C::STAGE);
#pragma unroll
        for (int kb = 0; kb < C::KB; kb++) {
          tma_load_3d(sk + kb * KT * 128, &tmK, kv_full(s), head * D + kb * 64, j * KT, b);
          tma_load_3d(sv + kb * KT * 128, &tmV, kv_full(s), head * D + kb * 64, j * KT, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // D = f32 (1<<4), A = B = bf16 (1<<7, 1<<10), N>>3 at [17,23), M>>4 at [24,29); bit 16: B is MN-major (the V tile)
      const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KT >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
      const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
      const uint32_t t_o = tmem_base + 2 * KT;
      auto issue_qk = [&](int j) {
        const int s = j % STAGES;
        mbar_wait(kv_full(s), (j / STAGES) & 1);
        mbar_wait(s_empty + 8 * (j & 1), ((j >> 1) & 1) ^ 1);        // the softmax has read S(j-2) out of this buffer
        tc_fence_after();
        const uint32_t sk = sKV + s * C::STAGE;
#pragma unroll
        for (int kb = 0; kb < C::KB; kb++) {
          const uint64_t ad = desc_k(sQ + kb * QT * 128), bd = desc_k(sk + kb * KT * 128);
#pragma unroll
          for (int k = 0; k < 4; k++) umma_f16(tmem_base + (j & 1) * KT, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_qk, (kb | k) != 0);
        }
        umma_commit(s_full + 8 * (j & 1));
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < nt; j++) {
        if (j + 1 < nt) issue_qk(j + 1);                             // scores of the next tile while the softmax works on this one
        const int s = j % STAGES;
        mbar_wait(p_full + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        const uint64_t ad = desc_k(sP + (j & 1) * C::P_BYTES);
        const uint64_t bd = desc_mn(sKV + s * C::STAGE + C::K_BYTES, KT * 128);
#pragma unroll
        for (int k = 0; k < KT / 16; k++) umma_f16(t_o, ad + (uint64_t)(2 * k), bd + (uint64_t)(128 * k), idesc_pv, (j | k) != 0);
        umma_commit(kv_empty(s));
        umma_commit(p_empty + 8 * (j & 1));
        umma_commit(o_ready);
      }
      umma_commit(o_final);
    }
  } else {
    // ===================== softmax warps: one thread per query row =====================
    SoftmaxBars sb{s_full, s_empty, p_full, p_empty, o_ready, o_final};
    softmax_role<D>(tmem_base, sP, sb, nt, p, warp, lane, q0, b, head * D, head);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
template <int D>
int launch(const cenet_attn_tc_args& a, cudaStream_t s) {
  using C = Cfg<D>;
  CUtensorMap tmQ, tmK, tmV;
  if (encode3(&tmQ, a.q, (long long)a.heads * D, a.Nq, a.B, a.ldq, a.bq, 64, QT, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
  if (encode3(&tmK, a.k, (long long)a.heads * D, a.Nk, a.B, a.ldk, a.bk, 64, KT, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
  if (encode3(&tmV, a.v, (long long)a.heads * D, a.Nk, a.B, a.ldv, a.bv, 64, KT, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
  TcAttnParams p;
  p.o = (bf16*)a.o; p.lse = a.lse; p.ldo = a.ldo; p.bo = a.bo; p.Nq = a.Nq; p.Nk = a.Nk; p.heads = a.heads;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.lse_mul = a.lse_base2 ? 1.f : 0.6931471805599453f;
  auto kern = attn_tc_kernel<D>;
  static std::once_flag once;
  std::call_once(once, [&] { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM); });
  dim3 grid(cdiv(a.Nq, QT), a.heads, a.B);
  kern<<<grid, NTH, C::SMEM, s>>>(tmQ, tmK, tmV, p);
  CENET_LAUNCH_CHECK("attn_tc");
  return 0;
}

// ---- wide heads (D = 192, 256, 320, ... 512): the contraction of Q K^T is STREAMED in 64-column blocks --------------------------
// The non-local blocks of the coarse decoder levels have d = C = 320 (14x14) and 512 (7x7): neither a [128 x C] Q tile plus K / V
// stages nor a C-column output accumulator fits next to the scores.  Here one CTA owns 128 query rows and ONE 64-column chunk of
// the output (blockIdx.y): per 64-key tile the producer streams {Q[:, kb], K_j[:, kb]} blocks through a 4-stage ring while the MMA
// thread accumulates S_j over the kb blocks, then the 64 x 64 V chunk feeds O += P_j V_j as in the kernel above.  The scores are
// recomputed once per output chunk (C/64 times) -- these levels have N <= 196 keys at 224x224, so that is noise, and it keeps the
// contraction on the tensor pipe instead of a CUDA-core GEMM over a materialised N x N map.
constexpr int WSTAGES = 4;
constexpr int W_STAGE_BYTES = QT * 128 + KT * 128;        // Q block + K block

__global__ void __launch_bounds__(NTH, 1) attn_tc_wide_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                             const __grid_constant__ CUtensorMap tmK,
                                                             const __grid_constant__ CUtensorMap tmV, const TcAttnParams p,
                                                             int kqb) {
  constexpr int P_BYTES = QT * KT * 2, V_BYTES = KT * 128;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * QT;
  const uint32_t sRing = sbase;
  const uint32_t sV = sRing + WSTAGES * W_STAGE_BYTES;
  const uint32_t sP = sV + 2 * V_BYTES;
  const uint32_t bar = sP + 2 * P_BYTES;
  auto r_full = [&](int s) { return bar + s * 8; };
  auto r_empty = [&](int s) { return bar + (WSTAGES + s) * 8; };
  const uint32_t v_full = bar + 2 * WSTAGES * 8;         // [2]
  const uint32_t v_empty = v_full + 16;                  // [2]
  const uint32_t s_full = v_empty + 16, s_empty = s_full + 16, p_full = s_empty + 16, p_empty = p_full + 16;
  const uint32_t o_ready = p_empty + 16, o_final = o_ready + 8, tmem_slot = o_final + 8;
  const int nt = (p.Nk + KT - 1) / KT;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    for (int s = 0; s < WSTAGES; s++) { mbar_init(r_full(s), 1); mbar_init(r_empty(s), 1); }
    for (int i = 0; i < 2; i++) {
      mbar_init(v_full + 8 * i, 1); mbar_init(v_empty + 8 * i, 1);
      mbar_init(s_full + 8 * i, 1); mbar_init(s_empty + 8 * i, 4);
      mbar_init(p_full + 8 * i, 4); mbar_init(p_empty + 8 * i, 1);
    }
    mbar_init(o_ready, 1);
    mbar_init(o_final, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int j = 0; j < nt; j++) {
        for (int kb = 0; kb < kqb; kb++, it++) {
          const int s = it % WSTAGES;
          mbar_wait(r_empty(s), ((it / WSTAGES) & 1) ^ 1);
          const uint32_t sq = sRing + s * W_STAGE_BYTES;
          mbar_arrive_expect_tx(r_full(s), W_STAGE_BYTES);
          tma_load_3d(sq, &tmQ, r_full(s), kb * 64, q0, b);
          tma_load_3d(sq + QT * 128, &tmK, r_full(s), kb * 64, j * KT, b);
        }
        mbar_wait(v_empty + 8 * (j & 1), ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(v_full + 8 * (j & 1), V_BYTES);
        tma_load_3d(sV + (j & 1) * V_BYTES, &tmV, v_full + 8 * (j & 1), chunk * 64, j * KT, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KT >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
      const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
      const uint32_t t_o = tmem_base + 2 * KT;
      uint32_t it = 0;
      auto issue_qk = [&](int j) {
        mbar_wait(s_empty + 8 * (j & 1), ((j >> 1) & 1) ^ 1);
        for (int kb = 0; kb < kqb; kb++, it++) {
          const int s = it % WSTAGES;
          mbar_wait(r_full(s), (it / WSTAGES) & 1);
          tc_fence_after();
          const uint32_t sq = sRing + s * W_STAGE_BYTES;
          const uint64_t ad = desc_k(sq), bd = desc_k(sq + QT * 128);
#pragma unroll
          for (int k = 0; k < 4; k++) umma_f16(tmem_base + (j & 1) * KT, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_qk, (kb | k) != 0);
          umma_commit(r_empty(s));
        }
        umma_commit(s_full + 8 * (j & 1));
      };
      issue_qk(0);
      for (int j = 0; j < nt; j++) {
        if (j + 1 < nt) issue_qk(j + 1);
        mbar_wait(p_full + 8 * (j & 1), (j >> 1) & 1);
        mbar_wait(v_full + 8 * (j & 1), (j >> 1) & 1);
        tc_fence_after();
        const uint64_t ad = desc_k(sP + (j & 1) * P_BYTES);
        const uint64_t bd = desc_mn(sV + (j & 1) * V_BYTES, KT * 128);
#pragma unroll
        for (int k = 0; k < KT / 16; k++) umma_f16(t_o, ad + (uint64_t)(2 * k), bd + (uint64_t)(128 * k), idesc_pv, (j | k) != 0);
        umma_commit(v_empty + 8 * (j & 1));
        umma_commit(p_empty + 8 * (j & 1));
        umma_commit(o_ready);
      }
      umma_commit(o_final);
    }
  } else {
    SoftmaxBars sb{s_full, s_empty, p_full, p_empty, o_ready, o_final};
    softmax_role<64>(tmem_base, sP, sb, nt, p, warp, lane, q0, b, chunk * 64, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

int launch_wide(const cenet_attn_tc_args& a, cudaStream_t s) {
  constexpr int SMEM = WSTAGES * W_STAGE_BYTES + 2 * KT * 128 + 2 * QT * KT * 2 + 256 + 1024;
  CUtensorMap tmQ, tmK, tmV;
  if (encode3(&tmQ, a.q, a.D, a.Nq, a.B, a.ldq, a.bq, 64, QT, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
  if (encode3(&tmK, a.k, a.D, a.Nk, a.B, a.ldk, a.bk, 64, KT, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
  if (encode3(&tmV, a.v, a.D, a.Nk, a.B, a.ldv, a.bv, 64, KT, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
  TcAttnParams p;
  p.o = (bf16*)a.o; p.lse = nullptr; p.ldo = a.ldo; p.bo = a.bo; p.Nq = a.Nq; p.Nk = a.Nk; p.heads = 1;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.lse_mul = 1.f;
  static std::once_flag once;
  std::call_once(once, [&] { cudaFuncSetAttribute(attn_tc_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM); });
  dim3 grid(cdiv(a.Nq, QT), a.D / 64, a.B);
  attn_tc_wide_kernel<<<grid, NTH, SMEM, s>>>(tmQ, tmK, tmV, p, a.D / 64);
  CENET_LAUNCH_CHECK("attn_tc_wide");
  return 0;
}
}  // namespace

bool cenet_attn_tc_eligible(const cenet_attn_tc_args* a) {
  static const bool off = getenv("CENET_B200_ATTN_TC") && atoi(getenv("CENET_B200_ATTN_TC")) == 0;
  if (off) return false;
  if (a->D != 64 && a->D != 128 && !(a->D > 128 && a->D <= 1024 && a->D % 64 == 0 && a->heads == 1 && a->lse == nullptr)) return false;
  if (a->B < 1 || a->B > 65535 || a->heads < 1 || a->heads > 65535 || a->Nq < 1 || a->Nk < 1) return false;
  const uintptr_t al = (uintptr_t)a->q | (uintptr_t)a->k | (uintptr_t)a->v | (uintptr_t)a->o;
  if (al & 15) return false;
  if ((a->ldq | a->ldk | a->ldv | a->ldo | a->bq | a->bk | a->bv | a->bo) % 8) return false;
  return true;
}

extern "C" int cenet_attn_tc(const cenet_attn_tc_args* a, cenet_stream_t s) {
  CENET_REQUIRE(a && a->q && a->k && a->v && a->o, "cenet_attn_tc: null pointer");
  CENET_REQUIRE(cenet_attn_tc_eligible(a), "cenet_attn_tc: needs bf16 operands with head width 64 or 128 (or one head of 192..1024, multiple of 64), 16-byte aligned pointers and "
                "row / image strides that are multiples of 8 elements (D=%d)", a->D);
  if (a->D == 64) return launch<64>(*a, to_stream(s));
  if (a->D == 128) return launch<128>(*a, to_stream(s));
  return launch_wide(*a, to_stream(s));              // one head of width 192..1024 (multiple of 64): streamed contraction
}
