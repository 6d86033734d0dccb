// Flash-style attention kernels for the two N x N attentions of the decoder:
//   * differential attention of the DSE blocks (multihead_diffattn.py:92-124)
//   * non-local block core (nlb.py:116-137)
// Neither materialises the N x N maps (the reference writes 2h fp32 maps of N^2 per image and re-reads them ~5x).
//
// Why mma.sync and not tcgen05 here: with head_dim 8..32 the contraction depth of Q K^T is one or two k16 steps, so
// the kernel is bound by the softmax (one MUFU.EX2 per score: 2h*N^2 per image, 16/clk/SM) and by issue slots, not by
// the tensor pipe.  mma.sync leaves the scores in registers where the softmax needs them; a tcgen05 version would add
// a TMEM->register round trip per score for no gain.  The GEMM-shaped layers use tcgen05 (gemm_tc.cu).
#include "common.cuh"
#include <cstdlib>
#include <type_traits>

bool cenet_attn_tc_eligible(const cenet_attn_tc_args* a);   // attn_tc.cu
int cenet_diffattn_tc(const void* qkv, void* out, void* om, float* lse, int B, int N, int heads, int hd, float lambda, float eps,
                      float mult, const float* kmax, cudaStream_t s);        // diffattn_tc.cu: 0 done, 1 not applicable, -1 error

namespace {

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp2 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax of 2^f on [-0.5,0.5], rel. err < 1.1e-4, below the
// 2^-9 rounding of the bf16 probabilities it feeds).  The softmax of these kernels is bound by the MUFU (XU) pipe
// (ncu: sm__inst_executed_pipe_xu 67% vs fma 23% / alu 26%), so every POLY_EVERY-th exponential is moved over.
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;                  // 1.5*2^23: round-to-nearest integer lands in the low mantissa bits
  const float f = x - (t - 12582912.f);            // f in [-0.5, 0.5]
  float p = fmaf(0.0555041086f, f, 0.2402264923f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// Two exponentials at once on the FMA pipe with packed fp32 (FFMA2): same Cody-Waite split and cubic as poly_exp2, 3 packed
// adds + 3 FFMA2 + 2 clamps + 2 exponent inserts for two values = 5 issue slots per value against 8 XU cycles per warp-wide
// MUFU.EX2, and it runs on pipes the softmax leaves idle.  x must be finite (masked tiles keep the MUFU path).
__device__ __forceinline__ void poly_exp2_x2(float x0, float x1, float& e0, float& e1) {
  const f32x2 magic = pk2(12582912.f, 12582912.f), one = pk2(1.f, 1.f), neg = pk2(-1.f, -1.f);
  const f32x2 x = pk2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const f32x2 t = ffma2(x, one, magic);                        // round(x) in the low mantissa bits
  const f32x2 f = ffma2(ffma2(magic, neg, t), neg, x);         // x - (t - magic), in [-0.5, 0.5]
  f32x2 p = ffma2(pk2(0.0555041086f, 0.0555041086f), f, pk2(0.2402264923f, 0.2402264923f));
  p = ffma2(p, f, pk2(0.6931471806f, 0.6931471806f));
  p = ffma2(p, f, one);
  float p0, p1, t0, t1;
  upk2(p, p0, p1); upk2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int QT = 64;    // query rows per CTA (4 warps x 16)
constexpr int KT = 64;    // keys per pipeline stage
constexpr int NTHREADS = 128;

// =====================================================================================================================
// One attention "stream": 16 query rows of one warp against a KT-key tile; scores in registers.
//   HDP: padded q/k depth (multiple of 16), DV: value width (multiple of 16)
// S = Q K^T -> online softmax (base-2, scale folded) -> O += P V ; l accumulates row sums.
// =====================================================================================================================
template <int HDP>
__device__ __forceinline__ void qk_tile(float (&S)[KT / 8][4], const uint32_t (&qf)[HDP / 16][4], uint32_t k_smem,
                                        int k_stride, int lane) {
#pragma unroll
  for (int j = 0; j < KT / 8; j++) { S[j][0] = S[j][1] = S[j][2] = S[j][3] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < HDP / 16; ks++) {
#pragma unroll
    for (int jp = 0; jp < KT / 16; jp++) {
      uint32_t b[4];
      const int key = jp * 16 + (lane & 7) + ((lane >> 4) << 3);
      const int d = ks * 16 + (((lane >> 3) & 1) << 3);
      ldsm_x4(b, k_smem + key * k_stride + d * 2);
      mma_16816(S[2 * jp], qf[ks], b[0], b[1]);
      mma_16816(S[2 * jp + 1], qf[ks], b[2], b[3]);
    }
  }
}

// online softmax update for the 2 rows this thread owns (g and g+8); returns P packed as A fragments
template <int DV, int POLY>
__device__ __forceinline__ void softmax_tile(float (&S)[KT / 8][4], uint32_t (&P)[KT / 16][4], float (&m)[2],
                                             float (&l)[2], float (&O)[DV / 8][4], float scale_log2, int kbase, int N,
                                             int lane) {
  const int t = lane & 3;
  if (kbase + KT > N) {   // ragged last tile: mask keys >= N
#pragma unroll
    for (int j = 0; j < KT / 8; j++) {
      const int key = kbase + j * 8 + 2 * t;
      if (key >= N) { S[j][0] = -INFINITY; S[j][2] = -INFINITY; }
      if (key + 1 >= N) { S[j][1] = -INFINITY; S[j][3] = -INFINITY; }
    }
  }
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < KT / 8; j++) {
    mx0 = fmaxf(mx0, fmaxf(S[j][0], S[j][1]));
    mx1 = fmaxf(mx1, fmaxf(S[j][2], S[j][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const float mn0 = fmaxf(m[0], mx0 * scale_log2), mn1 = fmaxf(m[1], mx1 * scale_log2);
  const float c0 = fast_exp2(m[0] - mn0), c1 = fast_exp2(m[1] - mn1);
  m[0] = mn0; m[1] = mn1;
  l[0] *= c0; l[1] *= c1;
#pragma unroll
  for (int j = 0; j < DV / 8; j++) { O[j][0] *= c0; O[j][1] *= c0; O[j][2] *= c1; O[j][3] *= c1; }
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < KT / 8; j++) {
    const float p0 = fast_exp2(fmaf(S[j][0], scale_log2, -mn0));
    const float p1 = (POLY > 0 && (j % (POLY > 0 ? POLY : 1)) == 0) ? poly_exp2(fmaf(S[j][1], scale_log2, -mn0))
                                                   : fast_exp2(fmaf(S[j][1], scale_log2, -mn0));
    const float p2 = fast_exp2(fmaf(S[j][2], scale_log2, -mn1));
    const float p3 = (POLY > 0 && (j % (POLY > 0 ? POLY : 1)) == 1 % (POLY > 0 ? POLY : 1)) ? poly_exp2(fmaf(S[j][3], scale_log2, -mn1))
                                                          : fast_exp2(fmaf(S[j][3], scale_log2, -mn1));
    s0 += p0 + p1; s1 += p2 + p3;
    P[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
    P[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
  }
  l[0] += s0; l[1] += s1;
}

// =====================================================================================================================
// Differential attention.  grid = (ceil(N/64), heads, B), 128 threads.
//   HD: real head_dim (8,16,32,64); HDP = max(HD,16); DV = 2*HD
// smem rows are padded by 16 bytes so that ldmatrix's 8 row addresses fall in distinct 16-byte bank groups.
// =====================================================================================================================
// Pre-pass of the bounded-softmax mode: kmax[b, j] = max_n |k_{b,n,j}|_2 for every softmax map j (keys of width HD).
// With it every score of query row r obeys s <= |q_r| kmax (Cauchy-Schwarz), which serves as a FIXED softmax shift:
// no running max, no rescaling of the accumulators, and the exponentials no longer wait for a cross-lane reduction.
template <int HD>
__global__ void __launch_bounds__(256) kmax_kernel(const bf16* __restrict__ qkv, float* __restrict__ kmax, int N,
                                                   long long row, int koff) {
  __shared__ float red[8];
  const int j = blockIdx.x, b = blockIdx.y;
  const bf16* kb = qkv + (long long)b * N * row + koff + j * HD;
  float mx = 0.f;
  for (int n = threadIdx.x; n < N; n += 256) {
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < HD; c += 8) {
      float v[8];
      ldv<8>(kb + (long long)n * row + c, v);
#pragma unroll
      for (int i = 0; i < 8; i++) ss = fmaf(v[i], v[i], ss);
    }
    mx = fmaxf(mx, ss);
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; i++) mx = fmaxf(mx, red[i]);
    kmax[(long long)b * gridDim.x + j] = sqrtf(mx);
  }
}

template <int HD, int DVT>
struct DiffCfg {
  static constexpr int HDP = HD < 16 ? 16 : HD;
  static constexpr int DV = DVT;
  static constexpr int KSTR = HDP * 2 + 16;           // bytes per K/Q smem row
  static constexpr int VSTR = DV * 2 + 16;            // bytes per V smem row
  static constexpr int Q_BYTES = 2 * QT * KSTR;       // two maps
  static constexpr int K_BYTES = 2 * KT * KSTR;       // two maps, one stage
  static constexpr int V_BYTES = KT * VSTR;
  static constexpr int STAGE = K_BYTES + V_BYTES;
  static constexpr int SMEM = Q_BYTES + 2 * STAGE;
};

// Global layout of one token row: [ q: 2h heads x HD | k: 2h x HD | v: h x DV ]; output rows: [ h x DV ].
// For the reference's natural layout HD = hd, DV = 2 hd; head dims that are not MMA friendly (hd = 20) are zero-padded
// by the host (HD = 32, DV = 48): `dv_real` restores the RMSNorm mean and `scale_log2` carries the real 1/sqrt(hd).
template <int HD, int DVT, int POLY, int MINB>
__global__ void __launch_bounds__(NTHREADS, MINB) diffattn_flash_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                  int N, int heads, float dv_real, float scale_log2,
                                                                  float lambda, float eps, float mult,
                                                                  const float* __restrict__ kmax) {
  using Cfg = DiffCfg<HD, DVT>;
  const int E = 2 * heads * HD;            // width of the q block (= k block)
  const int EO = heads * DVT;              // output row width
  constexpr int HDP = Cfg::HDP, DV = Cfg::DV, KSTR = Cfg::KSTR, VSTR = Cfg::VSTR;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * QT;
  const long long row3E = 2LL * E + EO;
  const bf16* base = qkv + (long long)b * N * row3E;
  const uint32_t sQ = smem_u32(smem), sKV = sQ + Cfg::Q_BYTES;

  if (HD < 16) {   // zero the padding half of every Q/K row once (cp.async only ever writes the first 16 bytes)
    for (int i = tid; i < (Cfg::SMEM) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
  }
  constexpr int QCH = HD / 8;     // 16-byte chunks per real q/k row
  constexpr int VCH = DV / 8;
  // ---- Q tile (both maps) ----
  for (int i = tid; i < 2 * QT * QCH; i += NTHREADS) {
    const int ch = i % QCH, r = (i / QCH) % QT, mp = i / (QCH * QT);
    const int n = q0 + r;
    const bf16* src = base + (long long)(n < N ? n : N - 1) * row3E + (2 * head + mp) * HD + ch * 8;
    cp_async16(sQ + (mp * QT + r) * KSTR + ch * 16, src, n < N);
  }
  auto load_kv = [&](int tile, int stage) {
    const uint32_t sK = sKV + stage * Cfg::STAGE, sV = sK + Cfg::K_BYTES;
    const int k0 = tile * KT;
    for (int i = tid; i < 2 * KT * QCH; i += NTHREADS) {
      const int ch = i % QCH, r = (i / QCH) % KT, mp = i / (QCH * KT);
      const int n = k0 + r;
      const bf16* src = base + (long long)(n < N ? n : N - 1) * row3E + E + (2 * head + mp) * HD + ch * 8;
      cp_async16(sK + (mp * KT + r) * KSTR + ch * 16, src, n < N);
    }
    for (int i = tid; i < KT * VCH; i += NTHREADS) {
      const int ch = i % VCH, r = i / VCH;
      const int n = k0 + r;
      const bf16* src = base + (long long)(n < N ? n : N - 1) * row3E + 2 * E + head * DV + ch * 8;
      cp_async16(sV + r * VSTR + ch * 16, src, n < N);
    }
  };
  const int ntiles = (N + KT - 1) / KT;
  load_kv(0, 0);
  cp_async_commit();
  if (ntiles > 1) load_kv(1, 1);
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();

  // Q fragments for this warp's 16 rows, both maps
  uint32_t qf[2][HDP / 16][4];
#pragma unroll
  for (int mp = 0; mp < 2; mp++)
#pragma unroll
    for (int ks = 0; ks < HDP / 16; ks++) {
      const int r = warp * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
      const int d = ks * 16 + ((lane >> 4) << 3);
      ldsm_x4(qf[mp][ks], sQ + (mp * QT + r) * KSTR + d * 2);
    }

  // O accumulators; L[mp] is a 9th "value column" of ones: the tensor pipe accumulates the softmax row sums
  // (column 0 of that tile = sum_k P[row,k]) and they are rescaled together with O -- 4 HMMA per tile instead of 32 FADD.
  float O[2][DV / 8][4], L[2][4];
  float m[2][2];
#pragma unroll
  for (int mp = 0; mp < 2; mp++) {
    m[mp][0] = m[mp][1] = -INFINITY;
    L[mp][0] = L[mp][1] = L[mp][2] = L[mp][3] = 0.f;
#pragma unroll
    for (int j = 0; j < DV / 8; j++) O[mp][j][0] = O[mp][j][1] = O[mp][j][2] = O[mp][j][3] = 0.f;
  }
  const uint32_t ones_b = (lane < 4) ? 0x3F803F80u : 0u;     // B fragment of the ones column (n = 0 <=> lane/4 == 0)
  // ---- bounded-softmax mode: fixed shift |q_r| * max_n|k_n| (log2 units) instead of a running maximum.  The shift
  // may exceed the true row maximum by up to 2x its own size, so it is only used while 2*shift stays far inside the
  // fp32 / bf16 exponent range (< 60 => p >= 2^-120); otherwise this warp keeps the online-max path. ----
  bool bounded = false;
  if (kmax != nullptr) {
    float worst = 0.f;
#pragma unroll
    for (int mp = 0; mp < 2; mp++) {
      float q0s = 0.f, q1s = 0.f;
#pragma unroll
      for (int ks = 0; ks < HDP / 16; ks++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&qf[mp][ks][r]));
          const float ssq = f.x * f.x + f.y * f.y;
          if (r & 1) q1s += ssq; else q0s += ssq;          // a0,a2 -> row g ; a1,a3 -> row g+8
        }
      }
      q0s += __shfl_xor_sync(0xffffffffu, q0s, 1); q0s += __shfl_xor_sync(0xffffffffu, q0s, 2);
      q1s += __shfl_xor_sync(0xffffffffu, q1s, 1); q1s += __shfl_xor_sync(0xffffffffu, q1s, 2);
      const float km = kmax[(long long)b * (2 * heads) + 2 * head + mp] * scale_log2 * 1.0001f;
      m[mp][0] = sqrtf(q0s) * km + 1e-3f;
      m[mp][1] = sqrtf(q1s) * km + 1e-3f;
      worst = fmaxf(worst, fmaxf(m[mp][0], m[mp][1]));
    }
    bounded = __all_sync(0xffffffffu, worst < 60.f);
    if (!bounded) { m[0][0] = m[0][1] = m[1][0] = m[1][1] = -INFINITY; }
  }
  // lane-dependent ldmatrix offsets, hoisted out of the tile loop
  const uint32_t k_lane = ((lane & 7) + ((lane >> 4) << 3)) * KSTR + (((lane >> 3) & 1) << 4);
  const uint32_t v_lane = ((lane & 7) + (((lane >> 3) & 1) << 3)) * VSTR + ((lane >> 4) << 4);
  const int tq = lane & 3;

  auto process_tile = [&](auto masked_tag, auto bounded_tag, int tile) {
    constexpr bool MASKED = decltype(masked_tag)::value;
    constexpr bool BOUNDED = decltype(bounded_tag)::value;
    const int stage = tile & 1;
    const uint32_t sK = sKV + stage * Cfg::STAGE, sV = sK + Cfg::K_BYTES;
#pragma unroll
    for (int mp = 0; mp < 2; mp++) {
      float S[KT / 8][4];
#pragma unroll
      for (int j = 0; j < KT / 8; j++) { S[j][0] = S[j][1] = S[j][2] = S[j][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < HDP / 16; ks++) {
#pragma unroll
        for (int jp = 0; jp < KT / 16; jp++) {
          uint32_t bfr[4];
          ldsm_x4(bfr, sK + (mp * KT + jp * 16) * KSTR + ks * 32 + k_lane);
          mma_16816(S[2 * jp], qf[mp][ks], bfr[0], bfr[1]);
          mma_16816(S[2 * jp + 1], qf[mp][ks], bfr[2], bfr[3]);
        }
      }
      if (MASKED) {   // ragged last tile only: keys >= N
#pragma unroll
        for (int j = 0; j < KT / 8; j++) {
          const int key = tile * KT + j * 8 + 2 * tq;
          if (key >= N) { S[j][0] = -INFINITY; S[j][2] = -INFINITY; }
          if (key + 1 >= N) { S[j][1] = -INFINITY; S[j][3] = -INFINITY; }
        }
      }
      float mn0 = m[mp][0], mn1 = m[mp][1];
      if (!BOUNDED) {
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < KT / 8; j++) {
          mx0 = fmaxf(mx0, fmaxf(S[j][0], S[j][1]));
          mx1 = fmaxf(mx1, fmaxf(S[j][2], S[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        mn0 = fmaxf(m[mp][0], mx0 * scale_log2); mn1 = fmaxf(m[mp][1], mx1 * scale_log2);
        const float c0 = fast_exp2(m[mp][0] - mn0), c1 = fast_exp2(m[mp][1] - mn1);
        m[mp][0] = mn0; m[mp][1] = mn1;
        L[mp][0] *= c0; L[mp][2] *= c1;
#pragma unroll
        for (int j = 0; j < DV / 8; j++) { O[mp][j][0] *= c0; O[mp][j][1] *= c0; O[mp][j][2] *= c1; O[mp][j][3] *= c1; }
      }
      uint32_t P[KT / 16][4];
      // POLY > 0: of the four exponentials of score group j, the pair of one row goes to the FMA pipe (packed cubic) when
      // j % POLY == 0 (rows alternate with j); the rest stay on the MUFU pipe.  Masked tiles (-inf scores) use MUFU only.
      const f32x2 sc2 = pk2(scale_log2, scale_log2), nm0 = pk2(-mn0, -mn0), nm1 = pk2(-mn1, -mn1);
#pragma unroll
      for (int j = 0; j < KT / 8; j++) {
        float x0, x1, x2, x3, p0, p1, p2, p3;
        upk2(ffma2(pk2(S[j][0], S[j][1]), sc2, nm0), x0, x1);
        upk2(ffma2(pk2(S[j][2], S[j][3]), sc2, nm1), x2, x3);
        constexpr int PERIOD = POLY > 0 ? POLY : 1;
        const bool poly_j = POLY > 0 && !MASKED && (j % PERIOD) == 0;
        if (poly_j && ((j / PERIOD) & 1) == 0) poly_exp2_x2(x0, x1, p0, p1);
        else { p0 = fast_exp2(x0); p1 = fast_exp2(x1); }
        if (poly_j && ((j / PERIOD) & 1) == 1) poly_exp2_x2(x2, x3, p2, p3);
        else { p2 = fast_exp2(x2); p3 = fast_exp2(x3); }
        P[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
        P[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
#pragma unroll
      for (int kk = 0; kk < KT / 16; kk++) {
        mma_16816(L[mp], P[kk], ones_b, ones_b);
#pragma unroll
        for (int np = 0; np < DV / 16; np++) {
          uint32_t v[4];
          ldsm_x4_t(v, sV + kk * 16 * VSTR + np * 32 + v_lane);
          mma_16816(O[mp][2 * np], P[kk], v[0], v[1]);
          mma_16816(O[mp][2 * np + 1], P[kk], v[2], v[3]);
        }
      }
    }
  };

  const bool ragged = (N % KT) != 0;
  for (int tile = 0; tile < ntiles; tile++) {
    const bool last_ragged = ragged && tile == ntiles - 1;
    if (bounded) {
      if (last_ragged) process_tile(std::true_type{}, std::true_type{}, tile);
      else process_tile(std::false_type{}, std::true_type{}, tile);
    } else {
      if (last_ragged) process_tile(std::true_type{}, std::false_type{}, tile);
      else process_tile(std::false_type{}, std::false_type{}, tile);
    }
    __syncthreads();                       // everyone is done with this stage
    if (tile + 2 < ntiles) load_kv(tile + 2, tile & 1);
    cp_async_commit();
    cp_async_wait<1>();                    // tile+1 has landed
    __syncthreads();
  }

  // ---- epilogue: normalise both maps, difference, RMSNorm over DV, scale, store ----
  float inv[2][2];
#pragma unroll
  for (int mp = 0; mp < 2; mp++) {
    inv[mp][0] = 1.f / __shfl_sync(0xffffffffu, L[mp][0], lane & ~3);   // column 0 of the ones tile lives in lane 4g
    inv[mp][1] = 1.f / __shfl_sync(0xffffffffu, L[mp][2], lane & ~3);
  }
  float ss0 = 0.f, ss1 = 0.f;
#pragma unroll
  for (int j = 0; j < DV / 8; j++) {
    O[0][j][0] = O[0][j][0] * inv[0][0] - lambda * O[1][j][0] * inv[1][0];
    O[0][j][1] = O[0][j][1] * inv[0][0] - lambda * O[1][j][1] * inv[1][0];
    O[0][j][2] = O[0][j][2] * inv[0][1] - lambda * O[1][j][2] * inv[1][1];
    O[0][j][3] = O[0][j][3] * inv[0][1] - lambda * O[1][j][3] * inv[1][1];
    ss0 += O[0][j][0] * O[0][j][0] + O[0][j][1] * O[0][j][1];
    ss1 += O[0][j][2] * O[0][j][2] + O[0][j][3] * O[0][j][3];
  }
  ss0 += __shfl_xor_sync(0xffffffffu, ss0, 1); ss0 += __shfl_xor_sync(0xffffffffu, ss0, 2);
  ss1 += __shfl_xor_sync(0xffffffffu, ss1, 1); ss1 += __shfl_xor_sync(0xffffffffu, ss1, 2);
  const float r0 = rsqrtf(ss0 / dv_real + eps) * mult, r1 = rsqrtf(ss1 / dv_real + eps) * mult;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = q0 + warp * 16 + g, n1 = n0 + 8;
  bf16* ob = out + (long long)b * N * EO + head * DV;
#pragma unroll
  for (int j = 0; j < DV / 8; j++) {
    const int col = j * 8 + 2 * t;
    if (n0 < N) *reinterpret_cast<uint32_t*>(ob + (long long)n0 * EO + col) = pack_bf16(O[0][j][0] * r0, O[0][j][1] * r0);
    if (n1 < N) *reinterpret_cast<uint32_t*>(ob + (long long)n1 * EO + col) = pack_bf16(O[0][j][2] * r1, O[0][j][3] * r1);
  }
}

// =====================================================================================================================
// Non-local block core: single head, d = C in {64,128}.  tpg rows are [theta | phi | g].
// =====================================================================================================================
template <int D>
struct NlCfg {
  static constexpr int STR = D * 2 + 16;
  static constexpr int Q_BYTES = QT * STR;
  static constexpr int STAGE = 2 * KT * STR;   // phi + g
  static constexpr int SMEM = Q_BYTES + 2 * STAGE;
};

// Generic strided single-softmax attention, head width D: out = softmax(q k^T * scale) v.
//   non-local block: q|k|v are the three column blocks of one [B,N,3D] tensor, one head;
//   encoder SR attention: q [B,Nq,heads*64], k|v column blocks of kv [B,Nk,2*heads*64], Nk <= 64 reduced keys.
struct AttnPtrs {
  const bf16 *q, *k, *v;
  bf16* o;
  long long ldq, ldk, ldv, ldo;          // row pitches (elements)
  long long bq, bk, bv, bo;              // per-image strides (elements)
};

template <int D>
__global__ void __launch_bounds__(NTHREADS) nonlocal_flash_kernel(const AttnPtrs a, int Nq, int N, float scale_log2) {
  using Cfg = NlCfg<D>;
  constexpr int STR = Cfg::STR, CH = D / 8;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int head = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * QT;
  const bf16* qb = a.q + (long long)b * a.bq + head * D;
  const bf16* kb = a.k + (long long)b * a.bk + head * D;
  const bf16* vb = a.v + (long long)b * a.bv + head * D;
  const uint32_t sQ = smem_u32(smem), sKV = sQ + Cfg::Q_BYTES;
  for (int i = tid; i < QT * CH; i += NTHREADS) {
    const int ch = i % CH, r = i / CH, n = q0 + r;
    cp_async16(sQ + r * STR + ch * 16, qb + (long long)(n < Nq ? n : Nq - 1) * a.ldq + ch * 8, n < Nq);
  }
  auto load_kv = [&](int tile, int stage) {
    const uint32_t sK = sKV + stage * Cfg::STAGE, sV = sK + KT * STR;
    const int k0 = tile * KT;
    for (int i = tid; i < KT * CH; i += NTHREADS) {
      const int ch = i % CH, r = i / CH, n = k0 + r;
      const long long rr = n < N ? n : N - 1;
      cp_async16(sK + r * STR + ch * 16, kb + rr * a.ldk + ch * 8, n < N);
      cp_async16(sV + r * STR + ch * 16, vb + rr * a.ldv + ch * 8, n < N);
    }
  };
  const int ntiles = (N + KT - 1) / KT;
  load_kv(0, 0);
  cp_async_commit();
  if (ntiles > 1) load_kv(1, 1);
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  uint32_t qf[D / 16][4];
#pragma unroll
  for (int ks = 0; ks < D / 16; ks++) {
    const int r = warp * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
    const int d = ks * 16 + ((lane >> 4) << 3);
    ldsm_x4(qf[ks], sQ + r * STR + d * 2);
  }
  float O[D / 8][4];
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < D / 8; j++) O[j][0] = O[j][1] = O[j][2] = O[j][3] = 0.f;
  for (int tile = 0; tile < ntiles; tile++) {
    const int stage = tile & 1;
    const uint32_t sK = sKV + stage * Cfg::STAGE, sV = sK + KT * STR;
    float S[KT / 8][4];
    uint32_t P[KT / 16][4];
    qk_tile<D>(S, qf, sK, STR, lane);
    softmax_tile<D, 0>(S, P, m, l, O, scale_log2, tile * KT, N, lane);
#pragma unroll
    for (int kk = 0; kk < KT / 16; kk++) {
#pragma unroll
      for (int np = 0; np < D / 16; np++) {
        uint32_t v[4];
        const int key = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int dv = np * 16 + ((lane >> 4) << 3);
        ldsm_x4_t(v, sV + key * STR + dv * 2);
        mma_16816(O[2 * np], P[kk], v[0], v[1]);
        mma_16816(O[2 * np + 1], P[kk], v[2], v[3]);
      }
    }
    __syncthreads();
    if (tile + 2 < ntiles) load_kv(tile + 2, stage);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
  }
  float inv[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    float s = l[r];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    inv[r] = 1.f / s;
  }
  const int g = lane >> 2, t = lane & 3;
  const int n0 = q0 + warp * 16 + g, n1 = n0 + 8;
  bf16* ob = a.o + (long long)b * a.bo + head * D;
#pragma unroll
  for (int j = 0; j < D / 8; j++) {
    const int col = j * 8 + 2 * t;
    if (n0 < Nq) *reinterpret_cast<uint32_t*>(ob + (long long)n0 * a.ldo + col) = pack_bf16(O[j][0] * inv[0], O[j][1] * inv[0]);
    if (n1 < Nq) *reinterpret_cast<uint32_t*>(ob + (long long)n1 * a.ldo + col) = pack_bf16(O[j][2] * inv[1], O[j][3] * inv[1]);
  }
}

template <int HD, int DVT, int POLY, int MINB>
int launch_diff(const bf16* qkv, bf16* out, int B, int N, int heads, int hd_real, float lambda, float eps, float mult,
                float* kmax_ws, cudaStream_t s) {
  using Cfg = DiffCfg<HD, DVT>;
  if (kmax_ws) {
    const long long row = 4LL * heads * HD + (long long)heads * DVT;
    kmax_kernel<HD><<<dim3(2 * heads, B), 256, 0, s>>>(qkv, kmax_ws, N, row, 2 * heads * HD);
    CENET_LAUNCH_CHECK("diffattn_kmax");
  }
  auto kern = diffattn_flash_kernel<HD, DVT, POLY, MINB>;
  if (Cfg::SMEM > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  dim3 grid(cdiv(N, QT), heads, B);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)hd_real);
  kern<<<grid, NTHREADS, Cfg::SMEM, s>>>(qkv, out, N, heads, (float)(2 * hd_real), scale_log2, lambda, eps, mult, kmax_ws);
  CENET_LAUNCH_CHECK("diffattn_flash");
  return 0;
}
template <int D>
int launch_attn(const AttnPtrs& a, int B, int heads, int Nq, int Nk, float scale, cudaStream_t s, const char* name) {
  {   // tcgen05 / TMEM / TMA kernel (attn_tc.cu) whenever the operands are TMA-addressable; this mma.sync kernel otherwise
    cenet_attn_tc_args t;
    t.q = a.q; t.k = a.k; t.v = a.v; t.o = a.o; t.lse = nullptr;
    t.ldq = a.ldq; t.ldk = a.ldk; t.ldv = a.ldv; t.ldo = a.ldo; t.bq = a.bq; t.bk = a.bk; t.bv = a.bv; t.bo = a.bo;
    t.B = B; t.heads = heads; t.Nq = Nq; t.Nk = Nk; t.D = D; t.scale = scale; t.lse_base2 = 0;
    if (cenet_attn_tc_eligible(&t)) return cenet_attn_tc(&t, (cenet_stream_t)s);
  }
  using Cfg = NlCfg<D>;
  auto kern = nonlocal_flash_kernel<D>;
  if (Cfg::SMEM > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  dim3 grid(cdiv(Nq, QT), heads, B);
  kern<<<grid, NTHREADS, Cfg::SMEM, s>>>(a, Nq, Nk, scale * 1.4426950408889634f);
  CENET_LAUNCH_CHECK(name);
  return 0;
}
template <int D>
int launch_nl(const bf16* tpg, bf16* out, int B, int N, float scale, cudaStream_t s) {
  AttnPtrs a;
  a.q = tpg; a.k = tpg + D; a.v = tpg + 2 * D; a.o = out;
  a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
  a.bq = a.bk = a.bv = (long long)N * 3 * D; a.bo = (long long)N * D;
  return launch_attn<D>(a, B, 1, N, N, scale, s, "nonlocal_flash");
}
}  // namespace

static int diffattn_dispatch(const bf16* q, bf16* o, int B, int N, int heads, int hdp, int dvp, int hd_real, float lambda,
                             float eps, float mult, float* ws, cudaStream_t st) {
  static const int poly = getenv("CENET_DA_POLY") ? atoi(getenv("CENET_DA_POLY")) : 4;   // measured: 0: 2.98 ms, 1: 3.22, 2: 2.92, 4: 2.91 (S56, B=64)
  if (hdp == 8 && dvp == 16) {
    if (poly == 4) return launch_diff<8, 16, 4, 4>(q, o, B, N, heads, hd_real, lambda, eps, mult, ws, st);
    if (poly == 3) return launch_diff<8, 16, 3, 4>(q, o, B, N, heads, hd_real, lambda, eps, mult, ws, st);
    if (poly == 2) return launch_diff<8, 16, 2, 4>(q, o, B, N, heads, hd_real, lambda, eps, mult, ws, st);
    if (poly == 1) return launch_diff<8, 16, 1, 4>(q, o, B, N, heads, hd_real, lambda, eps, mult, ws, st);
    return launch_diff<8, 16, 0, 4>(q, o, B, N, heads, hd_real, lambda, eps, mult, ws, st);
  }
  if (hdp == 16 && dvp == 32) return launch_diff<16, 32, 0, 4>(q, o, B, N, heads, hd_real, lambda, eps, mult, ws, st);
  if (hdp == 32 && dvp == 48) return launch_diff<32, 48, 0, 2>(q, o, B, N, heads, hd_real, lambda, eps, mult, ws, st);
  if (hdp == 32 && dvp == 64) return launch_diff<32, 64, 0, 2>(q, o, B, N, heads, hd_real, lambda, eps, mult, ws, st);
  if (hdp == 64 && dvp == 128) return launch_diff<64, 128, 0, 1>(q, o, B, N, heads, hd_real, lambda, eps, mult, nullptr, st);
  CENET_FAIL("cenet_diffattn_flash: no kernel for padded head_dim %d / value width %d; use the materialised path", hdp, dvp);
}

static int launch_kmax(const void* qkv, float* kmax_ws, int B, int N, int E, int heads, int hd, cenet_stream_t s) {
  const long long row = 3LL * E;
  switch (hd) {
    case 8: kmax_kernel<8><<<dim3(2 * heads, B), 256, 0, to_stream(s)>>>((const bf16*)qkv, kmax_ws, N, row, E); break;
    case 16: kmax_kernel<16><<<dim3(2 * heads, B), 256, 0, to_stream(s)>>>((const bf16*)qkv, kmax_ws, N, row, E); break;
    case 32: kmax_kernel<32><<<dim3(2 * heads, B), 256, 0, to_stream(s)>>>((const bf16*)qkv, kmax_ws, N, row, E); break;
    default: kmax_kernel<64><<<dim3(2 * heads, B), 256, 0, to_stream(s)>>>((const bf16*)qkv, kmax_ws, N, row, E); break;
  }
  CENET_LAUNCH_CHECK("diffattn_kmax");
  return 0;
}

// Training forward of the differential attention on the tcgen05 kernel: Om[:, m*2hd : +2hd] = softmax(q_m k_m^T / sqrt(hd)) v_{m/2}
// for the 2*heads maps of every image and lse (log2 units, [B, 2*heads, N]) for cenet_flash_bwd; qkv rows are
// [q: 2h x hd | k: 2h x hd | v: h x 2hd] (multihead_diffattn.py:92-113).   head_dim in {8,16,32,64}.
extern "C" int cenet_diffattn_fwd_train(const void* qkv, void* om, float* lse, int B, int N, int E, int heads, float* kmax_ws,
                                        cenet_stream_t s) {
  if (B == 0 || N == 0) return 0;
  CENET_REQUIRE(qkv && om && lse, "cenet_diffattn_fwd_train: null pointer");
  CENET_REQUIRE(heads >= 1 && E % (2 * heads) == 0, "cenet_diffattn_fwd_train: E=%d not divisible by 2*heads=%d", E, 2 * heads);
  const int hd = E / (2 * heads);
  CENET_REQUIRE((hd == 8 || hd == 16 || hd == 32 || hd == 64) && B <= 65535 && heads <= 65535,
                "cenet_diffattn_fwd_train: head_dim %d has no tcgen05 instantiation (8/16/32/64); use cenet_flash_fwd", hd);
  if (kmax_ws && launch_kmax(qkv, kmax_ws, B, N, E, heads, hd, s)) return -1;
  const int rc = cenet_diffattn_tc(qkv, nullptr, om, lse, B, N, heads, hd, 0.f, 0.f, 1.f, kmax_ws, to_stream(s));
  CENET_REQUIRE(rc <= 0, "cenet_diffattn_fwd_train: operands must be 16-byte aligned");
  return rc;
}

extern "C" int cenet_diffattn_flash(const void* qkv, void* out, int B, int N, int E, int heads, float lambda, float eps,
                                    float mult, float* kmax_ws, cenet_stream_t s) {
  if (B == 0 || N == 0) return 0;
  CENET_REQUIRE(qkv && out, "cenet_diffattn_flash: null pointer");
  CENET_REQUIRE(heads >= 1 && E % (2 * heads) == 0, "cenet_diffattn_flash: E=%d not divisible by 2*heads=%d", E, 2 * heads);
  CENET_REQUIRE(B <= 65535 && heads <= 65535, "cenet_diffattn_flash: grid too large");
  const int hd = E / (2 * heads);
  if (hd == 8 || hd == 16 || hd == 32 || hd == 64) {
    // tcgen05 / TMEM / TMA kernel (diffattn_tc.cu); the kmax pre-pass feeds its fixed softmax shift
    if (kmax_ws && launch_kmax(qkv, kmax_ws, B, N, E, heads, hd, s)) return -1;
    const int rc = cenet_diffattn_tc(qkv, out, nullptr, nullptr, B, N, heads, hd, lambda, eps, mult, kmax_ws, to_stream(s));
    if (rc <= 0) return rc;
    return diffattn_dispatch((const bf16*)qkv, (bf16*)out, B, N, heads, hd, 2 * hd, hd, lambda, eps, mult, nullptr, to_stream(s));
  }
  return diffattn_dispatch((const bf16*)qkv, (bf16*)out, B, N, heads, hd, 2 * hd, hd, lambda, eps, mult, kmax_ws, to_stream(s));
}

extern "C" int cenet_diffattn_flash_padded(const void* qkv, void* out, int B, int N, int heads, int hd_pad, int dv_pad,
                                           int hd_real, float lambda, float eps, float mult, float* kmax_ws,
                                           cenet_stream_t s) {
  if (B == 0 || N == 0) return 0;
  CENET_REQUIRE(qkv && out, "cenet_diffattn_flash_padded: null pointer");
  CENET_REQUIRE(heads >= 1 && hd_real >= 1 && hd_real <= hd_pad && 2 * hd_real <= dv_pad,
                "cenet_diffattn_flash_padded: bad head geometry (hd %d pad %d, dv pad %d)", hd_real, hd_pad, dv_pad);
  CENET_REQUIRE(B <= 65535 && heads <= 65535, "cenet_diffattn_flash_padded: grid too large");
  return diffattn_dispatch((const bf16*)qkv, (bf16*)out, B, N, heads, hd_pad, dv_pad, hd_real, lambda, eps, mult, kmax_ws, to_stream(s));
}

// bf16 fast path of cenet_sr_attention (attn_sr.cu): q [B,N,C], kv [B,Nk,2C], head_dim 64
int cenet_sr_attention_mma(const void* q, const void* kv, void* out, int B, int N, int Nk, int C, int heads, float scale,
                           cudaStream_t s) {
  AttnPtrs a;
  a.q = (const bf16*)q; a.k = (const bf16*)kv; a.v = (const bf16*)kv + C; a.o = (bf16*)out;
  a.ldq = C; a.ldk = a.ldv = 2 * C; a.ldo = C;
  a.bq = (long long)N * C; a.bk = a.bv = (long long)Nk * 2 * C; a.bo = (long long)N * C;
  return launch_attn<64>(a, B, heads, N, Nk, scale, s, "sr_attention_mma");
}

extern "C" int cenet_nonlocal_flash(const void* tpg, void* out, int B, int N, int C, float scale, cenet_stream_t s) {
  if (B == 0 || N == 0) return 0;
  CENET_REQUIRE(tpg && out, "cenet_nonlocal_flash: null pointer");
  CENET_REQUIRE(B <= 65535, "cenet_nonlocal_flash: grid too large");
  switch (C) {
    case 64: return launch_nl<64>((const bf16*)tpg, (bf16*)out, B, N, scale, to_stream(s));
    case 128: return launch_nl<128>((const bf16*)tpg, (bf16*)out, B, N, scale, to_stream(s));
    default: {
      cenet_attn_tc_args t;
      const bf16* tp = (const bf16*)tpg;
      t.q = tp; t.k = tp + C; t.v = tp + 2 * C; t.o = out; t.lse = nullptr;
      t.ldq = t.ldk = t.ldv = 3 * C; t.ldo = C;
      t.bq = t.bk = t.bv = (long long)N * 3 * C; t.bo = (long long)N * C;
      t.B = B; t.heads = 1; t.Nq = N; t.Nk = N; t.D = C; t.scale = scale; t.lse_base2 = 0;
      if (cenet_attn_tc_eligible(&t)) return cenet_attn_tc(&t, s);        // wide heads: streamed-contraction tcgen05 kernel
      CENET_FAIL("cenet_nonlocal_flash: C=%d is neither 64 / 128 nor a multiple of 64 up to 1024; use the materialised path", C);
    }
  }
}
