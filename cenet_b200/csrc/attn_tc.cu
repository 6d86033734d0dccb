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
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <mutex>

namespace {
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

// ---- PTX wrappers (same forms as gemm_tc.cu / wgrad_tc.cu) ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // bounded: a protocol error traps (the launch fails with an error) instead of hanging the GPU -- each try may suspend up to
  // 10 ms, so the bound is far beyond any legitimate wait
  for (uint32_t tries = 0;; tries++) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    if (done) return;
    if (tries > 4000u) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,"
      "%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// K-major SWIZZLE_128B operand (rows of 128 bytes, 8-row groups 1024 bytes apart): Q, K and P tiles
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B operand: the V tile as TMA stores it ([keys] x 64 columns per box; LBO = next 64-column box)
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

struct TcAttnParams {
  bf16* o;
  float* lse;                // optional [B, heads, Nq]: natural-log-sum-exp of the scaled scores (training forward)
  long long ldo, bo;
  int Nq, Nk, heads;
  float scale_log2;
};

template <int D>
__global__ void __launch_bounds__(NTH, D == 64 ? 2 : 1) attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                      const __grid_constant__ CUtensorMap tmK,
                                                                      const __grid_constant__ CUtensorMap tmV,
                                                                      const TcAttnParams p) {
  pdl_prologue();
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
        for (int c0 = 0; c0 < D; c0 += 32) {
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
      const uint32_t prow = sP + (j & 1) * C::P_BYTES + row * 128;
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
    bf16* orow = p.o + (long long)b * p.bo + (long long)n * p.ldo + head * D;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
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
      p.lse[((long long)b * p.heads + head) * p.Nq + n] = (m + log2f(l)) * 0.6931471805599453f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// {columns, rows, images} view of a [B][N][ld] bf16 operand starting at `base`; box = 64 columns x `box_rows` rows
int encode3(CUtensorMap* tm, const void* base, long long cols, long long rows, long long B, long long ld, long long bstride,
            int box_rows) {
  EncodeTiledFn enc = get_encode();
  CENET_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bstride * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, str, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CENET_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (attention operand) failed with CUresult %d", (int)r);
  return 0;
}

template <int D>
int launch(const cenet_attn_tc_args& a, cudaStream_t s) {
  using C = Cfg<D>;
  CUtensorMap tmQ, tmK, tmV;
  if (encode3(&tmQ, a.q, (long long)a.heads * D, a.Nq, a.B, a.ldq, a.bq, QT)) return -1;
  if (encode3(&tmK, a.k, (long long)a.heads * D, a.Nk, a.B, a.ldk, a.bk, KT)) return -1;
  if (encode3(&tmV, a.v, (long long)a.heads * D, a.Nk, a.B, a.ldv, a.bv, KT)) return -1;
  TcAttnParams p;
  p.o = (bf16*)a.o; p.lse = a.lse; p.ldo = a.ldo; p.bo = a.bo; p.Nq = a.Nq; p.Nk = a.Nk; p.heads = a.heads;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  auto kern = attn_tc_kernel<D>;
  static std::once_flag once;
  std::call_once(once, [&] { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM); });
  dim3 grid(cdiv(a.Nq, QT), a.heads, a.B);
  kern<<<grid, NTH, C::SMEM, s>>>(tmQ, tmK, tmV, p);
  CENET_LAUNCH_CHECK("attn_tc");
  return 0;
}
}  // namespace

bool cenet_attn_tc_eligible(const cenet_attn_tc_args* a) {
  static const bool off = getenv("CENET_B200_ATTN_TC") && atoi(getenv("CENET_B200_ATTN_TC")) == 0;
  if (off) return false;
  if (a->D != 64 && a->D != 128) return false;
  if (a->B < 1 || a->B > 65535 || a->heads < 1 || a->heads > 65535 || a->Nq < 1 || a->Nk < 1) return false;
  const uintptr_t al = (uintptr_t)a->q | (uintptr_t)a->k | (uintptr_t)a->v | (uintptr_t)a->o;
  if (al & 15) return false;
  if ((a->ldq | a->ldk | a->ldv | a->ldo | a->bq | a->bk | a->bv | a->bo) % 8) return false;
  return true;
}

extern "C" int cenet_attn_tc(const cenet_attn_tc_args* a, cenet_stream_t s) {
  CENET_REQUIRE(a && a->q && a->k && a->v && a->o, "cenet_attn_tc: null pointer");
  CENET_REQUIRE(cenet_attn_tc_eligible(a), "cenet_attn_tc: needs bf16 operands with head width 64 or 128, 16-byte aligned pointers and "
                "row / image strides that are multiples of 8 elements (D=%d)", a->D);
  if (a->D == 64) return launch<64>(*a, to_stream(s));
  return launch<128>(*a, to_stream(s));
}
