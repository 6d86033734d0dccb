// Attention kernels of the training path (bf16, mma.sync m16n8k16, fp32 accumulate, scores never leave registers).
//
// One generic flash attention with saved log-sum-exp serves the three attention sites in train mode:
//   encoder SR attention   (pvtv2.py:88-105)             dqk = dv = 64, <= 64 reduced keys
//   non-local block        (nlb.py:116-137)              dqk = dv = C (64)
//   differential attention (multihead_diffattn.py:92-116) 2h softmax maps, dqk = hd, dv = 2hd, two maps share one V head
//     (the A1 - lambda*A2 combination + RMSNorm runs on the per-map outputs in diff_rmsnorm_*; so its backward hands each
//      map its own dO and the attention backward is the plain one).
// Layout: token matrices, row pitch ld*, map m reads columns [m*dqk, +dqk) of Q / K and [(m / vdiv)*dv, +dv) of V and
// writes columns [m*dv, +dv) of O; images are consecutive blocks of N rows.
// Backward = FlashAttention-2 scheme without atomics: delta = rowsum(dO * O); a query-parallel kernel for dQ and a
// key-parallel kernel for dK / dV (which loops over the vdiv maps that share its V head) -> deterministic.
#include "train_common.cuh"

bool cenet_attn_tc_eligible(const cenet_attn_tc_args* a);   // attn_tc.cu

namespace {
constexpr int FT = 128;          // threads per CTA (4 warps x 16 rows)
constexpr int BQ = 64, BKEY = 64;

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const bf16* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const bf16* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t& r0, uint32_t& r1, const bf16* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// A fragment (16 rows x 16 k) straight from a global row-major matrix; rows >= nrows and columns >= D read as zero
template <int D>
__device__ __forceinline__ void load_a_global(uint32_t (&a)[4], const bf16* __restrict__ base, long long ld, int row0, int nrows, int k0,
                                              int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = row0 + g + (i & 1) * 8, c = k0 + 2 * t + (i >> 1) * 8;
    a[i] = (r < nrows && c < D) ? *reinterpret_cast<const uint32_t*>(base + (long long)r * ld + c) : 0u;
  }
}

// [64][D] tile (row pitch D + 8) from global rows row0.. (rows >= nrows zero-filled); 128 threads, 16-byte chunks
// key / query tiles loaded per __syncthreads round: small head dims do little math per 64-row tile, so several tiles are staged at
// once and the (exposed) load latency is paid once per KT tiles
template <int DQK>
struct KTiles { static constexpr int value = DQK <= 16 ? 4 : (DQK <= 32 ? 2 : 1); };

template <int D, int ROWS = 64>
__device__ __forceinline__ void load_tile(bf16* __restrict__ sm, const bf16* __restrict__ base, long long ld, int row0, int nrows, int tid) {
  constexpr int CH = D / 8, P = D + 8;
  for (int i = tid; i < ROWS * CH; i += FT) {
    const int r = i / CH, c = (i - r * CH) * 8;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (row0 + r < nrows) u = *reinterpret_cast<const uint4*>(base + (long long)(row0 + r) * ld + c);
    *reinterpret_cast<uint4*>(sm + r * P + c) = u;
  }
}

// The same tile in two steps: fetch this thread's 16-byte chunks into registers (loads stay in flight while the previous
// tile is being computed), later store them into the shared tile.  ROWS * D / 8 must be a multiple of FT.
template <int D, int ROWS>
struct TileRegs { uint4 u[ROWS * (D / 8) / FT]; };
template <int D, int ROWS>
__device__ __forceinline__ void fetch_tile(TileRegs<D, ROWS>& t, const bf16* __restrict__ base, long long ld, int row0, int nrows, int tid) {
  constexpr int CH = D / 8;
  static_assert((ROWS * CH) % FT == 0, "tile chunks must divide evenly over the CTA");
#pragma unroll
  for (int k = 0; k < ROWS * CH / FT; k++) {
    const int i = tid + k * FT, r = i / CH, c = (i - r * CH) * 8;
    t.u[k] = make_uint4(0, 0, 0, 0);
    if (row0 + r < nrows) t.u[k] = *reinterpret_cast<const uint4*>(base + (long long)(row0 + r) * ld + c);
  }
}
template <int D, int ROWS>
__device__ __forceinline__ void store_tile(bf16* __restrict__ sm, const TileRegs<D, ROWS>& t, int tid) {
  constexpr int CH = D / 8, P = D + 8;
#pragma unroll
  for (int k = 0; k < ROWS * CH / FT; k++) {
    const int i = tid + k * FT, r = i / CH, c = (i - r * CH) * 8;
    *reinterpret_cast<uint4*>(sm + r * P + c) = t.u[k];
  }
}

// acc[j] (16 x 8 tiles over 64 "n" rows of the smem tile) += A(16 x D) * tile^T, tile stored [n][D]
template <int D>
__device__ __forceinline__ void mma_a_tileT(float (&acc)[8][4], const uint32_t (&af)[(D + 15) / 16][4], const bf16* sm, int lane) {
  constexpr int P = D + 8;
  const int jm = lane >> 3, r = lane & 7;
  if constexpr (D == 8) {
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(b0, b1, b2, b3, sm + ((j + jm) * 8 + r) * P);
      mma16816(acc[j], af[0], b0, 0u);
      mma16816(acc[j + 1], af[0], b1, 0u);
      mma16816(acc[j + 2], af[0], b2, 0u);
      mma16816(acc[j + 3], af[0], b3, 0u);
    }
  } else {
#pragma unroll
    for (int ks = 0; ks < D / 16; ks++)
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(b0, b1, b2, b3, sm + ((j + (jm >> 1)) * 8 + r) * P + ks * 16 + (jm & 1) * 8);
        mma16816(acc[j], af[ks], b0, b1);
        mma16816(acc[j + 1], af[ks], b2, b3);
      }
  }
}

// out[nj] (16 x 8 tiles over D columns) += A(16 x 64, four k16 fragments pf) * tile, tile stored [k = 64 rows][D]
template <int D>
__device__ __forceinline__ void mma_a_tile(float (&out)[D / 8][4], const uint32_t (&pf)[4][4], const bf16* sm, int lane) {
  constexpr int P = D + 8;
  const int jm = lane >> 3, r = lane & 7;
#pragma unroll
  for (int kk = 0; kk < 4; kk++) {
    if constexpr (D == 8) {
      uint32_t b0, b1;
      ldsm_x2_t(b0, b1, sm + (kk * 16 + (jm & 1) * 8 + r) * P);
      mma16816(out[0], pf[kk], b0, b1);
    } else {
#pragma unroll
      for (int nj = 0; nj < D / 8; nj += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(b0, b1, b2, b3, sm + (kk * 16 + (jm & 1) * 8 + r) * P + (nj + (jm >> 1)) * 8);
        mma16816(out[nj], pf[kk], b0, b1);
        mma16816(out[nj + 1], pf[kk], b2, b3);
      }
    }
  }
}

// C fragments (8 n-tiles of a 16 x 64 block) -> four k16 A fragments (bf16)
__device__ __forceinline__ void c_to_a(uint32_t (&pf)[4][4], const float (&c)[8][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; kk++) {
    pf[kk][0] = pack_bf16(c[2 * kk][0], c[2 * kk][1]);
    pf[kk][1] = pack_bf16(c[2 * kk][2], c[2 * kk][3]);
    pf[kk][2] = pack_bf16(c[2 * kk + 1][0], c[2 * kk + 1][1]);
    pf[kk][3] = pack_bf16(c[2 * kk + 1][2], c[2 * kk + 1][3]);
  }
}

struct FlashArgs {
  const bf16 *Q, *K, *V;
  bf16* O;
  const bf16* dO;
  bf16 *dQ, *dK, *dV;
  float *lse, *delta;
  long long ldq, ldk, ldv, ldo;
  int maps, Nq, Nk, vdiv;
  float scale, c;        // c = scale * log2(e)
  // dK / dV of a SHORT key set (<= 64 keys: the SR attention of the encoder, 49 reduced keys) -- the key-parallel kernel would
  // be B x heads CTAs walking all queries; instead blockIdx.x splits the QUERIES, every CTA leaves fp32 partials
  // part[split][b][map][64][D] and flash_kv_reduce_kernel adds them in split order (deterministic)
  int qsplit, qchunk;
  float *partK, *partV;
};

// ------------------------------------------------------------------------------------------------ forward
template <int DQK, int DV>
__global__ void __launch_bounds__(FT) flash_fwd_kernel(const FlashArgs p) {
  constexpr int KT = KTiles<DQK>::value;
  __shared__ __align__(16) bf16 sKb[KT * BKEY * (DQK + 8)];
  __shared__ __align__(16) bf16 sVb[KT * BKEY * (DV + 8)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, m = blockIdx.y, q0 = blockIdx.x * BQ + warp * 16;
  const bf16* Qp = p.Q + (long long)b * p.Nq * p.ldq + m * DQK;
  const bf16* Kp = p.K + (long long)b * p.Nk * p.ldk + m * DQK;
  const bf16* Vp = p.V + (long long)b * p.Nk * p.ldv + (m / p.vdiv) * DV;
  constexpr int KS = (DQK + 15) / 16;
  uint32_t qf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ks++) load_a_global<DQK>(qf[ks], Qp, p.ldq, q0, p.Nq, ks * 16, lane);
  float mi[2] = {-INFINITY, -INFINITY}, li[2] = {0.f, 0.f};
  float o[DV / 8][4];
#pragma unroll
  for (int j = 0; j < DV / 8; j++)
#pragma unroll
    for (int e = 0; e < 4; e++) o[j][e] = 0.f;
  // head dims >= 32: the next key / value tile is prefetched into registers while this one is computed (see fetch_tile);
  // small head dims keep the plain load (their register budget buys occupancy instead)
  constexpr bool PREF = DQK >= 32;
  TileRegs<DQK, KT * BKEY> rk;
  TileRegs<DV, KT * BKEY> rv;
  if constexpr (PREF) {
    fetch_tile<DQK, KT * BKEY>(rk, Kp, p.ldk, 0, p.Nk, tid);
    fetch_tile<DV, KT * BKEY>(rv, Vp, p.ldv, 0, p.Nk, tid);
  }
  for (int kt0 = 0; kt0 < p.Nk; kt0 += KT * BKEY) {
    __syncthreads();
    if constexpr (PREF) {
      store_tile<DQK, KT * BKEY>(sKb, rk, tid);
      store_tile<DV, KT * BKEY>(sVb, rv, tid);
    } else {
      load_tile<DQK, KT * BKEY>(sKb, Kp, p.ldk, kt0, p.Nk, tid);
      load_tile<DV, KT * BKEY>(sVb, Vp, p.ldv, kt0, p.Nk, tid);
    }
    __syncthreads();
    if constexpr (PREF) {
      if (kt0 + KT * BKEY < p.Nk) {
        fetch_tile<DQK, KT * BKEY>(rk, Kp, p.ldk, kt0 + KT * BKEY, p.Nk, tid);
        fetch_tile<DV, KT * BKEY>(rv, Vp, p.ldv, kt0 + KT * BKEY, p.Nk, tid);
      }
    }
   for (int sb = 0; sb < KT; sb++) {
    const int k0 = kt0 + sb * BKEY;
    if (k0 >= p.Nk) break;
    const bf16* sK = sKb + sb * BKEY * (DQK + 8);
    const bf16* sV = sVb + sb * BKEY * (DV + 8);
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) s[j][e] = 0.f;
    mma_a_tileT<DQK>(s, qf, sK, lane);
    // scores stay raw; the running max lives in the scaled (base-2) domain: one FMNMX + one FFMA + one EX2 + one FADD per score.
    // Only a ragged last tile pays for the key mask (warp-uniform branch).
    if (k0 + BKEY > p.Nk) {
#pragma unroll
      for (int j = 0; j < 8; j++)
#pragma unroll
        for (int e = 0; e < 4; e++)
          if (k0 + j * 8 + 2 * t + (e & 1) >= p.Nk) s[j][e] = -INFINITY;
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
    float al[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const float mn = fmaxf(mi[h], mx[h] * p.c);              // p.c > 0
      al[h] = ex2(mi[h] - mn);
      mi[h] = mn;
      li[h] *= al[h];
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) {
        s[j][e] = ex2(fmaf(s[j][e], p.c, -mi[e >> 1]));
        li[e >> 1] += s[j][e];
      }
#pragma unroll
    for (int j = 0; j < DV / 8; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) o[j][e] *= al[e >> 1];
    uint32_t pf[4][4];
    c_to_a(pf, s);
    mma_a_tile<DV>(o, pf, sV, lane);
   }
  }
#pragma unroll
  for (int h = 0; h < 2; h++) {
    li[h] += __shfl_xor_sync(0xffffffffu, li[h], 1);
    li[h] += __shfl_xor_sync(0xffffffffu, li[h], 2);
  }
  bf16* Op = p.O + (long long)b * p.Nq * p.ldo + m * DV;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int r = q0 + g + h * 8;
    if (r >= p.Nq) continue;
    const float inv = 1.f / li[h];
#pragma unroll
    for (int j = 0; j < DV / 8; j++)
      *reinterpret_cast<uint32_t*>(Op + (long long)r * p.ldo + j * 8 + 2 * t) = pack_bf16(o[j][2 * h] * inv, o[j][2 * h + 1] * inv);
    if (t == 0) p.lse[((long long)b * p.maps + m) * p.Nq + r] = mi[h] + log2f(li[h]);
  }
}

// ------------------------------------------------------------------------------------------------ delta = rowsum(dO * O)
template <int DV>
__global__ void __launch_bounds__(256) flash_delta_kernel(const FlashArgs p, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;        // (b, m, r)
  if (i >= total) return;
  const int r = (int)(i % p.Nq);
  const long long bm = i / p.Nq;
  const int m = (int)(bm % p.maps), b = (int)(bm / p.maps);
  const long long off = ((long long)b * p.Nq + r) * p.ldo + m * DV;
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < DV; c += 8) {
    float a[8], d[8];
    ldv<8>(p.O + off + c, a);
    ldv<8>(p.dO + off + c, d);
#pragma unroll
    for (int v = 0; v < 8; v++) s = fmaf(a[v], d[v], s);
  }
  p.delta[i] = s;
}

// ------------------------------------------------------------------------------------------------ dQ
template <int DQK, int DV>
__global__ void __launch_bounds__(FT) flash_bwd_dq_kernel(const FlashArgs p) {
  constexpr int KT = KTiles<DQK>::value;
  __shared__ __align__(16) bf16 sKb[KT * BKEY * (DQK + 8)];
  __shared__ __align__(16) bf16 sVb[KT * BKEY * (DV + 8)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, m = blockIdx.y, q0 = blockIdx.x * BQ + warp * 16;
  const bf16* Qp = p.Q + (long long)b * p.Nq * p.ldq + m * DQK;
  const bf16* Kp = p.K + (long long)b * p.Nk * p.ldk + m * DQK;
  const bf16* Vp = p.V + (long long)b * p.Nk * p.ldv + (m / p.vdiv) * DV;
  const bf16* dOp = p.dO + (long long)b * p.Nq * p.ldo + m * DV;
  constexpr int KS = (DQK + 15) / 16;
  uint32_t qf[KS][4], gf[DV / 16][4];
#pragma unroll
  for (int ks = 0; ks < KS; ks++) load_a_global<DQK>(qf[ks], Qp, p.ldq, q0, p.Nq, ks * 16, lane);
#pragma unroll
  for (int ks = 0; ks < DV / 16; ks++) load_a_global<DV>(gf[ks], dOp, p.ldo, q0, p.Nq, ks * 16, lane);
  float lse[2], dl[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int r = q0 + g + h * 8;
    const long long idx = ((long long)b * p.maps + m) * p.Nq + r;
    lse[h] = r < p.Nq ? p.lse[idx] : INFINITY;
    dl[h] = r < p.Nq ? p.delta[idx] : 0.f;
  }
  float dq[DQK / 8][4];
#pragma unroll
  for (int j = 0; j < DQK / 8; j++)
#pragma unroll
    for (int e = 0; e < 4; e++) dq[j][e] = 0.f;
  // head dims >= 32: the next key / value tile is prefetched into registers while this one is computed (see fetch_tile);
  // small head dims keep the plain load (their register budget buys occupancy instead)
  constexpr bool PREF = DQK >= 32;
  TileRegs<DQK, KT * BKEY> rk;
  TileRegs<DV, KT * BKEY> rv;
  if constexpr (PREF) {
    fetch_tile<DQK, KT * BKEY>(rk, Kp, p.ldk, 0, p.Nk, tid);
    fetch_tile<DV, KT * BKEY>(rv, Vp, p.ldv, 0, p.Nk, tid);
  }
  for (int kt0 = 0; kt0 < p.Nk; kt0 += KT * BKEY) {
    __syncthreads();
    if constexpr (PREF) {
      store_tile<DQK, KT * BKEY>(sKb, rk, tid);
      store_tile<DV, KT * BKEY>(sVb, rv, tid);
    } else {
      load_tile<DQK, KT * BKEY>(sKb, Kp, p.ldk, kt0, p.Nk, tid);
      load_tile<DV, KT * BKEY>(sVb, Vp, p.ldv, kt0, p.Nk, tid);
    }
    __syncthreads();
    if constexpr (PREF) {
      if (kt0 + KT * BKEY < p.Nk) {
        fetch_tile<DQK, KT * BKEY>(rk, Kp, p.ldk, kt0 + KT * BKEY, p.Nk, tid);
        fetch_tile<DV, KT * BKEY>(rv, Vp, p.ldv, kt0 + KT * BKEY, p.Nk, tid);
      }
    }
   for (int sb = 0; sb < KT; sb++) {
    const int k0 = kt0 + sb * BKEY;
    if (k0 >= p.Nk) break;
    const bf16* sK = sKb + sb * BKEY * (DQK + 8);
    const bf16* sV = sVb + sb * BKEY * (DV + 8);
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) s[j][e] = dp[j][e] = 0.f;
    mma_a_tileT<DQK>(s, qf, sK, lane);
    mma_a_tileT<DV>(dp, gf, sV, lane);
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int col = k0 + j * 8 + 2 * t + (e & 1);
        const float pr = col < p.Nk ? ex2(s[j][e] * p.c - lse[e >> 1]) : 0.f;
        s[j][e] = pr * (dp[j][e] - dl[e >> 1]);                    // dS (without the softmax scale)
      }
    uint32_t pf[4][4];
    c_to_a(pf, s);
    mma_a_tile<DQK>(dq, pf, sK, lane);
   }
  }
  bf16* dQp = p.dQ + (long long)b * p.Nq * p.ldq + m * DQK;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int r = q0 + g + h * 8;
    if (r >= p.Nq) continue;
#pragma unroll
    for (int j = 0; j < DQK / 8; j++)
      *reinterpret_cast<uint32_t*>(dQp + (long long)r * p.ldq + j * 8 + 2 * t) =
          pack_bf16(dq[j][2 * h] * p.scale, dq[j][2 * h + 1] * p.scale);
  }
}

// ------------------------------------------------------------------------------------------------ dK, dV
// MODE 0: dK and dV in one pass; MODE 1: dV only; MODE 2: dK only (large head dims: the two accumulators do not fit together)
template <int DQK, int DV, int MODE>
__global__ void __launch_bounds__(FT) flash_bwd_dkv_kernel(const FlashArgs p) {
  constexpr int KT = KTiles<DQK>::value;
  __shared__ __align__(16) bf16 sQb[KT * BQ * (DQK + 8)];
  __shared__ __align__(16) bf16 sGb[KT * BQ * (DV + 8)];
  __shared__ float sLb[KT * BQ], sDb[KT * BQ];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const bool split = p.qsplit > 1;
  const int b = blockIdx.z, vh = blockIdx.y, kr0 = (split ? 0 : blockIdx.x * BKEY) + warp * 16;
  const int q_lo = split ? blockIdx.x * p.qchunk : 0, q_hi = split ? min(p.Nq, q_lo + p.qchunk) : p.Nq;
  const bf16* Vp = p.V + (long long)b * p.Nk * p.ldv + vh * DV;
  constexpr int KS = (DQK + 15) / 16;
  constexpr bool DO_DV = MODE != 2, DO_DK = MODE != 1;
  uint32_t vf[DO_DK ? DV / 16 : 1][4];
  if constexpr (DO_DK) {
#pragma unroll
    for (int ks = 0; ks < DV / 16; ks++) load_a_global<DV>(vf[ks], Vp, p.ldv, kr0, p.Nk, ks * 16, lane);
  }
  float dv[DO_DV ? DV / 8 : 1][4];
#pragma unroll
  for (int j = 0; j < (DO_DV ? DV / 8 : 1); j++)
#pragma unroll
    for (int e = 0; e < 4; e++) dv[j][e] = 0.f;
  const bool kvalid[2] = {kr0 + g < p.Nk, kr0 + g + 8 < p.Nk};
  for (int jm = 0; jm < p.vdiv; jm++) {
    const int m = vh * p.vdiv + jm;
    const bf16* Qp = p.Q + (long long)b * p.Nq * p.ldq + m * DQK;
    const bf16* Kp = p.K + (long long)b * p.Nk * p.ldk + m * DQK;
    const bf16* dOp = p.dO + (long long)b * p.Nq * p.ldo + m * DV;
    const float* lsep = p.lse + ((long long)b * p.maps + m) * p.Nq;
    const float* dlp = p.delta + ((long long)b * p.maps + m) * p.Nq;
    uint32_t kf[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) load_a_global<DQK>(kf[ks], Kp, p.ldk, kr0, p.Nk, ks * 16, lane);
    float dk[DO_DK ? DQK / 8 : 1][4];
#pragma unroll
    for (int j = 0; j < (DO_DK ? DQK / 8 : 1); j++)
#pragma unroll
      for (int e = 0; e < 4; e++) dk[j][e] = 0.f;
    // query tiles are prefetched into registers one round ahead: the global loads of round i+1 are in flight while round i
    // is computed (ncu on the first version: 13 % issue utilisation, 8.5 long-scoreboard stalls per issued instruction)
    constexpr int QR = KT * BQ;
    static_assert(QR % FT == 0 || QR < FT, "per-row scalars: one or more rows per thread");
    constexpr int SR = (QR + FT - 1) / FT;
    TileRegs<DQK, QR> rq;
    TileRegs<DV, QR> rg;
    float rl[SR], rd[SR];
    auto fetch = [&](int qt0) {
      fetch_tile<DQK, QR>(rq, Qp, p.ldq, qt0, q_hi, tid);
      fetch_tile<DV, QR>(rg, dOp, p.ldo, qt0, q_hi, tid);
#pragma unroll
      for (int k = 0; k < SR; k++) {
        const int i = tid + k * FT;
        const bool in = i < QR && qt0 + i < q_hi;
        rl[k] = in ? lsep[qt0 + i] : INFINITY;
        rd[k] = in ? dlp[qt0 + i] : 0.f;
      }
    };
    constexpr bool PREF = DQK >= 32;                      // small head dims: plain load (registers buy occupancy there)
    if constexpr (PREF) fetch(q_lo);
    for (int qt0 = q_lo; qt0 < q_hi; qt0 += QR) {
      __syncthreads();
      if constexpr (!PREF) fetch(qt0);
      store_tile<DQK, QR>(sQb, rq, tid);
      store_tile<DV, QR>(sGb, rg, tid);
#pragma unroll
      for (int k = 0; k < SR; k++) {
        const int i = tid + k * FT;
        if (i < QR) { sLb[i] = rl[k]; sDb[i] = rd[k]; }
      }
      __syncthreads();
      if constexpr (PREF) {
        if (qt0 + QR < q_hi) fetch(qt0 + QR);
      }
     for (int sb = 0; sb < KT; sb++) {
      if (qt0 + sb * BQ >= q_hi) break;
      const bf16* sQ = sQb + sb * BQ * (DQK + 8);
      const bf16* sG = sGb + sb * BQ * (DV + 8);
      const float* sL = sLb + sb * BQ;
      const float* sD = sDb + sb * BQ;
      float st[8][4];                                  // S^T: rows = keys of this warp, columns = 64 queries
#pragma unroll
      for (int j = 0; j < 8; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) st[j][e] = 0.f;
      mma_a_tileT<DQK>(st, kf, sQ, lane);
#pragma unroll
      for (int j = 0; j < 8; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int qi = j * 8 + 2 * t + (e & 1);
          st[j][e] = kvalid[e >> 1] ? ex2(st[j][e] * p.c - sL[qi]) : 0.f;      // P^T
        }
      uint32_t pf[4][4];
      if constexpr (DO_DV) {
        c_to_a(pf, st);
        mma_a_tile<DV>(dv, pf, sG, lane);
      }
      if constexpr (DO_DK) {
        float dpt[8][4];                               // dP^T
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
          for (int e = 0; e < 4; e++) dpt[j][e] = 0.f;
        mma_a_tileT<DV>(dpt, vf, sG, lane);
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int qi = j * 8 + 2 * t + (e & 1);
            dpt[j][e] = st[j][e] * (dpt[j][e] - sD[qi]);     // dS^T
          }
        c_to_a(pf, dpt);
        mma_a_tile<DQK>(dk, pf, sQ, lane);
      }
     }
    }
    if constexpr (DO_DK) {
      if (split) {
        float* pk = p.partK + ((((long long)blockIdx.x * gridDim.z + b) * p.maps + m) * BKEY) * DQK;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int r = kr0 + g + h * 8;
#pragma unroll
          for (int j = 0; j < DQK / 8; j++)
            *reinterpret_cast<float2*>(pk + (long long)r * DQK + j * 8 + 2 * t) = make_float2(dk[j][2 * h], dk[j][2 * h + 1]);
        }
        continue;
      }
      bf16* dKp = p.dK + (long long)b * p.Nk * p.ldk + m * DQK;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int r = kr0 + g + h * 8;
        if (r >= p.Nk) continue;
#pragma unroll
        for (int j = 0; j < DQK / 8; j++)
          *reinterpret_cast<uint32_t*>(dKp + (long long)r * p.ldk + j * 8 + 2 * t) =
              pack_bf16(dk[j][2 * h] * p.scale, dk[j][2 * h + 1] * p.scale);
      }
    }
  }
  if constexpr (DO_DV) {
    if (split) {
      float* pv = p.partV + ((((long long)blockIdx.x * gridDim.z + b) * gridDim.y + vh) * BKEY) * DV;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int r = kr0 + g + h * 8;
#pragma unroll
        for (int j = 0; j < DV / 8; j++)
          *reinterpret_cast<float2*>(pv + (long long)r * DV + j * 8 + 2 * t) = make_float2(dv[j][2 * h], dv[j][2 * h + 1]);
      }
      return;
    }
    bf16* dVp = p.dV + (long long)b * p.Nk * p.ldv + vh * DV;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int r = kr0 + g + h * 8;
      if (r >= p.Nk) continue;
#pragma unroll
      for (int j = 0; j < DV / 8; j++)
        *reinterpret_cast<uint32_t*>(dVp + (long long)r * p.ldv + j * 8 + 2 * t) = pack_bf16(dv[j][2 * h], dv[j][2 * h + 1]);
    }
  }
}

// out[b, r, h*D + c] = mult * sum_split part[split][b][h][r][c]   (r < Nk; fixed summation order); two columns per thread
__global__ void __launch_bounds__(256) flash_kv_reduce_kernel(const float* __restrict__ part, int nsplit, int B, int H, int D, int Nk,
                                                              bf16* __restrict__ out, long long ld, float mult) {
  const long long total = (long long)B * H * Nk * (D / 2);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % (D / 2)) * 2;
  long long t = i / (D / 2);
  const int r = (int)(t % Nk); t /= Nk;
  const int h = (int)(t % H);
  const int b = (int)(t / H);
  const long long stride = (long long)B * H * BKEY * D;
  const float* src = part + (((long long)b * H + h) * BKEY + r) * D + c;
  float2 a = make_float2(0.f, 0.f);
  for (int z = 0; z < nsplit; z++) {
    const float2 v = *reinterpret_cast<const float2*>(src + z * stride);
    a.x += v.x; a.y += v.y;
  }
  *reinterpret_cast<uint32_t*>(out + ((long long)b * Nk + r) * ld + h * D + c) = pack_bf16(a.x * mult, a.y * mult);
}

// ------------------------------------------------------------------------------------------------ materialised path
// dS = P * (dP - rowsum(P * dP)), in place over dP; one warp per row
template <typename T>
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const T* __restrict__ P, T* __restrict__ dP, long long rows, int n) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const T* pr = P + r * n;
  T* dr = dP + r * n;
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s = fmaf(ldf(pr + i), ldf(dr + i), s);
  s = warp_sum(s);
  for (int i = lane; i < n; i += 32) stf(dr + i, ldf(pr + i) * (ldf(dr + i) - s));
}

// ------------------------------------------------------------------------------------------------ lambda, diff + RMSNorm
__global__ void lambda_fwd_kernel(const float* q1, const float* k1, const float* q2, const float* k2, int hd, float init, float* lam) {
  const int lane = threadIdx.x;
  float a = 0.f, b = 0.f;
  for (int i = lane; i < hd; i += 32) { a = fmaf(q1[i], k1[i], a); b = fmaf(q2[i], k2[i], b); }
  a = warp_sum(a); b = warp_sum(b);
  if (lane == 0) { lam[0] = expf(a) - expf(b) + init; lam[1] = expf(a); lam[2] = expf(b); }
}
__global__ void lambda_bwd_kernel(const float* dlam, const float* q1, const float* k1, const float* q2, const float* k2, int hd,
                                  float* g1, float* g2, float* g3, float* g4) {
  const int lane = threadIdx.x;
  float a = 0.f, b = 0.f;
  for (int i = lane; i < hd; i += 32) { a = fmaf(q1[i], k1[i], a); b = fmaf(q2[i], k2[i], b); }
  const float e1 = expf(warp_sum(a)), e2 = expf(warp_sum(b)), d = dlam[0];
  for (int i = lane; i < hd; i += 32) {
    g1[i] = d * e1 * k1[i]; g2[i] = d * e1 * q1[i];
    g3[i] = -d * e2 * k2[i]; g4[i] = -d * e2 * q2[i];
  }
}

// o[r, h*seg + i] = mult * a_i * rsqrt(mean_i a^2 + eps),  a = Om[r, (2h)*seg + i] - lam * Om[r, (2h+1)*seg + i]
// A group of LPS lanes (power of two <= 32) owns one (row, head) segment; each lane keeps up to MAXV 8-element vectors of
// both maps in registers (16-byte loads, the two maps of a head are adjacent so their sectors are fully used), the two
// reductions run over the group with shuffles.  One read of Om (+ dO), one write.  seg % 8 == 0, seg <= 8 * 32 * MAXV.
constexpr int DR_MAXV = 2;
template <typename T>
__global__ void __launch_bounds__(256) diff_rmsnorm_fwd_kernel(const T* __restrict__ Om, const float* __restrict__ lamp, T* __restrict__ o,
                                                               long long rows, int heads, int seg, int lps, float eps, float mult) {
  const int sub = threadIdx.x % lps;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / lps;
  const bool live = i < rows * heads;
  const float lam = lamp[0];
  const long long r = live ? i / heads : 0;
  const int h = live ? (int)(i % heads) : 0;
  const T* a1 = Om + r * (2LL * heads * seg) + (2 * h) * seg;
  const T* a2 = a1 + seg;
  const int nv = seg >> 3;
  float a[DR_MAXV][8];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < DR_MAXV; k++) {
    const int v = sub + k * lps;
    if (v < nv) {
      float x1[8], x2[8];
      ldv<8>(a1 + v * 8, x1);
      ldv<8>(a2 + v * 8, x2);
#pragma unroll
      for (int j = 0; j < 8; j++) { a[k][j] = x1[j] - lam * x2[j]; ss = fmaf(a[k][j], a[k][j], ss); }
    }
  }
  for (int off = lps >> 1; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if (!live) return;
  const float rs = rsqrtf(ss / seg + eps) * mult;
  T* op = o + r * ((long long)heads * seg) + h * seg;
#pragma unroll
  for (int k = 0; k < DR_MAXV; k++) {
    const int v = sub + k * lps;
    if (v < nv) {
      float out[8];
#pragma unroll
      for (int j = 0; j < 8; j++) out[j] = a[k][j] * rs;
      stv<8>(op + v * 8, out);
    }
  }
}

// backward: da = mult*rs*(do - a * rs^2 * mean(do*a));  dOm1 = da, dOm2 = -lam*da, dlam -= sum(da * Om2)
template <typename T>
__global__ void __launch_bounds__(256) diff_rmsnorm_bwd_kernel(const T* __restrict__ dO, const T* __restrict__ Om,
                                                               const float* __restrict__ lamp, T* __restrict__ dOm, long long rows,
                                                               int heads, int seg, int lps, float eps, float mult,
                                                               float* __restrict__ ws) {
  __shared__ float red[8];
  const int sub = threadIdx.x % lps;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / lps;
  const bool live = i < rows * heads;
  const float lam = lamp[0];
  const long long r = live ? i / heads : 0;
  const int h = live ? (int)(i % heads) : 0;
  const T* a1 = Om + r * (2LL * heads * seg) + (2 * h) * seg;
  const T* a2 = a1 + seg;
  const T* gp = dO + r * ((long long)heads * seg) + h * seg;
  const int nv = seg >> 3;
  float a[DR_MAXV][8], x2[DR_MAXV][8], g[DR_MAXV][8];
  float ss = 0.f, ga = 0.f;
#pragma unroll
  for (int k = 0; k < DR_MAXV; k++) {
    const int v = sub + k * lps;
    if (v < nv) {
      float x1[8];
      ldv<8>(a1 + v * 8, x1);
      ldv<8>(a2 + v * 8, x2[k]);
      ldv<8>(gp + v * 8, g[k]);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        a[k][j] = x1[j] - lam * x2[k][j];
        ss = fmaf(a[k][j], a[k][j], ss);
        ga = fmaf(a[k][j], g[k][j], ga);
      }
    }
  }
  for (int off = lps >> 1; off > 0; off >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, off);
    ga += __shfl_xor_sync(0xffffffffu, ga, off);
  }
  float dl = 0.f;
  if (live) {
    const float rs = rsqrtf(ss / seg + eps);
    const float kk = rs * rs * ga / seg;
    T* d1 = dOm + r * (2LL * heads * seg) + (2 * h) * seg;
    T* d2 = d1 + seg;
#pragma unroll
    for (int k = 0; k < DR_MAXV; k++) {
      const int v = sub + k * lps;
      if (v < nv) {
        float o1[8], o2[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float da = mult * rs * (g[k][j] - a[k][j] * kk);
          o1[j] = da;
          o2[j] = -lam * da;
          dl = fmaf(-da, x2[k][j], dl);
        }
        stv<8>(d1 + v * 8, o1);
        stv<8>(d2 + v * 8, o2);
      }
    }
  }
  dl = warp_sum(dl);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dl;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; w++) s += red[w];
    ws[blockIdx.x] = s;
  }
}

// sum of n partials in fixed order -> out[0]
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ ws, int n, float* out) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += ws[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

FlashArgs make_args(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, void* O, long long ldo,
                    int maps, int Nq, int Nk, int vdiv, float scale) {
  FlashArgs a = {};
  a.Q = (const bf16*)Q; a.K = (const bf16*)K; a.V = (const bf16*)V; a.O = (bf16*)O;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
  a.maps = maps; a.Nq = Nq; a.Nk = Nk; a.vdiv = vdiv; a.scale = scale; a.c = scale * 1.4426950408889634f;
  return a;
}
int check_flash(const FlashArgs& a, int dqk, int dv, int B) {
  CENET_REQUIRE(a.Q && a.K && a.V && a.O, "flash attention: null pointer");
  CENET_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0, "flash attention: pitches must be multiples of 8");
  CENET_REQUIRE((((uintptr_t)a.Q | (uintptr_t)a.K | (uintptr_t)a.V | (uintptr_t)a.O) & 15) == 0, "flash attention: 16-byte alignment");
  CENET_REQUIRE(a.vdiv >= 1 && a.maps % a.vdiv == 0, "flash attention: maps must be a multiple of vdiv");
  CENET_REQUIRE(a.maps <= 65535 && B <= 65535, "flash attention: grid too large");
  return 0;
}
}  // namespace

#define FLASH_DISPATCH(dqk, dv, CALL)                                              \
  do {                                                                             \
    if (dqk == 8 && dv == 16) { constexpr int DQK = 8, DV = 16; CALL; }            \
    else if (dqk == 16 && dv == 32) { constexpr int DQK = 16, DV = 32; CALL; }     \
    else if (dqk == 32 && dv == 64) { constexpr int DQK = 32, DV = 64; CALL; }     \
    else if (dqk == 64 && dv == 64) { constexpr int DQK = 64, DV = 64; CALL; }     \
    else if (dqk == 128 && dv == 128) { constexpr int DQK = 128, DV = 128; CALL; } \
    else if (dqk == 80 && dv == 160) { constexpr int DQK = 80, DV = 160; CALL; }   \
    else CENET_FAIL("flash attention: (dqk, dv) = (%d, %d) not instantiated", dqk, dv); \
  } while (0)

extern "C" int cenet_flash_fwd(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, void* O,
                               long long ldo, float* lse, int B, int maps, int Nq, int Nk, int dqk, int dv, int vdiv, float scale,
                               cenet_stream_t st) {
  FlashArgs a = make_args(Q, ldq, K, ldk, V, ldv, O, ldo, maps, Nq, Nk, vdiv, scale);
  a.lse = lse;
  CENET_REQUIRE(lse, "cenet_flash_fwd: null lse");
  if (check_flash(a, dqk, dv, B)) return -1;
  if (B == 0) return 0;
  if (dqk == dv && vdiv == 1 && (dqk == 64 || dqk == 128)) {
    // head widths the 5th-gen tensor core can tile: tcgen05 / TMEM / TMA forward (attn_tc.cu), LSE in log2 units for the backward
    cenet_attn_tc_args t;
    t.q = Q; t.k = K; t.v = V; t.o = O; t.lse = lse;
    t.ldq = ldq; t.ldk = ldk; t.ldv = ldv; t.ldo = ldo;
    t.bq = (long long)Nq * ldq; t.bk = (long long)Nk * ldk; t.bv = (long long)Nk * ldv; t.bo = (long long)Nq * ldo;
    t.B = B; t.heads = maps; t.Nq = Nq; t.Nk = Nk; t.D = dqk; t.scale = scale; t.lse_base2 = 1;
    if (cenet_attn_tc_eligible(&t)) return cenet_attn_tc(&t, st);
  }
  dim3 grid(cdiv(Nq, BQ), maps, B);
  FLASH_DISPATCH(dqk, dv, (flash_fwd_kernel<DQK, DV><<<grid, FT, 0, to_stream(st)>>>(a)));
  CENET_LAUNCH_CHECK("flash_fwd");
  return 0;
}

extern "C" int cenet_flash_bwd(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, const void* O,
                               const void* dO, long long ldo, const float* lse, float* delta, void* dQ, void* dK, void* dV, int B,
                               int maps, int Nq, int Nk, int dqk, int dv, int vdiv, float scale, float* ws, long long ws_elems,
                               cenet_stream_t st) {
  FlashArgs a = make_args(Q, ldq, K, ldk, V, ldv, const_cast<void*>(O), ldo, maps, Nq, Nk, vdiv, scale);
  a.dO = (const bf16*)dO; a.lse = const_cast<float*>(lse); a.delta = delta;
  a.dQ = (bf16*)dQ; a.dK = (bf16*)dK; a.dV = (bf16*)dV;
  CENET_REQUIRE(dO && lse && delta && dQ && dK && dV, "cenet_flash_bwd: null pointer");
  CENET_REQUIRE((((uintptr_t)dO | (uintptr_t)dQ | (uintptr_t)dK | (uintptr_t)dV) & 15) == 0, "cenet_flash_bwd: 16-byte alignment");
  if (check_flash(a, dqk, dv, B)) return -1;
  if (B == 0) return 0;
  cudaStream_t s = to_stream(st);
  const long long total = (long long)B * maps * Nq;
  FLASH_DISPATCH(dqk, dv, (flash_delta_kernel<DV><<<cdiv(total, 256), 256, 0, s>>>(a, total)));
  CENET_LAUNCH_CHECK("flash_delta");
  FLASH_DISPATCH(dqk, dv, (flash_bwd_dq_kernel<DQK, DV><<<dim3(cdiv(Nq, BQ), maps, B), FT, 0, s>>>(a)));
  CENET_LAUNCH_CHECK("flash_bwd_dq");
  if (dqk >= 80) {                      // large head dims: dV and dK in separate passes (register budget)
    FLASH_DISPATCH(dqk, dv, (flash_bwd_dkv_kernel<DQK, DV, 1><<<dim3(cdiv(Nk, BKEY), maps / vdiv, B), FT, 0, s>>>(a)));
    CENET_LAUNCH_CHECK("flash_bwd_dv");
    FLASH_DISPATCH(dqk, dv, (flash_bwd_dkv_kernel<DQK, DV, 2><<<dim3(cdiv(Nk, BKEY), maps / vdiv, B), FT, 0, s>>>(a)));
    CENET_LAUNCH_CHECK("flash_bwd_dk");
  } else {
    // few keys, many queries (SR attention: 49 keys, up to 3136 queries): split the queries over CTAs, reduce fp32 partials
    const int hv = maps / vdiv;
    int qsplit = 1;
    if (ws && Nk <= BKEY && Nq >= 4 * BQ) {
      const int qr = (dqk <= 16 ? 4 : (dqk <= 32 ? 2 : 1)) * BQ;      // queries staged per round (KTiles)
      qsplit = std::min(cdiv(Nq, 2 * qr), std::max(1, cdiv(2 * kNumSMs, B * hv)));
      a.qchunk = cdiv(cdiv(Nq, qsplit), qr) * qr;
      qsplit = cdiv(Nq, a.qchunk);
      const long long need = (long long)qsplit * B * BKEY * ((long long)maps * dqk + (long long)hv * dv);
      if (qsplit < 2 || need > ws_elems) qsplit = 1;
    }
    a.qsplit = qsplit;
    if (qsplit > 1) {
      a.partK = ws;
      a.partV = ws + (long long)qsplit * B * maps * BKEY * dqk;
      FLASH_DISPATCH(dqk, dv, (flash_bwd_dkv_kernel<DQK, DV, 0><<<dim3(qsplit, hv, B), FT, 0, s>>>(a)));
      CENET_LAUNCH_CHECK("flash_bwd_dkv(split)");
      flash_kv_reduce_kernel<<<cdiv((long long)B * maps * Nk * (dqk / 2), 256), 256, 0, s>>>(a.partK, qsplit, B, maps, dqk, Nk, a.dK, ldk,
                                                                                         scale);
      flash_kv_reduce_kernel<<<cdiv((long long)B * hv * Nk * (dv / 2), 256), 256, 0, s>>>(a.partV, qsplit, B, hv, dv, Nk, a.dV, ldv, 1.f);
      CENET_LAUNCH_CHECK("flash_kv_reduce");
    } else {
      FLASH_DISPATCH(dqk, dv, (flash_bwd_dkv_kernel<DQK, DV, 0><<<dim3(cdiv(Nk, BKEY), hv, B), FT, 0, s>>>(a)));
      CENET_LAUNCH_CHECK("flash_bwd_dkv");
    }
  }
  return 0;
}

extern "C" int cenet_softmax_bwd_rows(const void* P, void* dP, int dtype, long long rows, int n, cenet_stream_t st) {
  CENET_REQUIRE(P && dP, "cenet_softmax_bwd_rows: null pointer");
  if (rows == 0) return 0;
  CENET_DISPATCH(dtype, T, (softmax_bwd_rows_kernel<T><<<cdiv(rows, 8), 256, 0, to_stream(st)>>>((const T*)P, (T*)dP, rows, n)));
  CENET_LAUNCH_CHECK("softmax_bwd_rows");
  return 0;
}

extern "C" int cenet_lambda_fwd(const float* q1, const float* k1, const float* q2, const float* k2, int hd, float init, float* lam,
                                cenet_stream_t st) {
  CENET_REQUIRE(q1 && k1 && q2 && k2 && lam, "cenet_lambda_fwd: null pointer");
  lambda_fwd_kernel<<<1, 32, 0, to_stream(st)>>>(q1, k1, q2, k2, hd, init, lam);
  CENET_LAUNCH_CHECK("lambda_fwd");
  return 0;
}
extern "C" int cenet_lambda_bwd(const float* dlam, const float* q1, const float* k1, const float* q2, const float* k2, int hd, float* g1,
                                float* g2, float* g3, float* g4, cenet_stream_t st) {
  CENET_REQUIRE(dlam && q1 && k1 && q2 && k2 && g1 && g2 && g3 && g4, "cenet_lambda_bwd: null pointer");
  lambda_bwd_kernel<<<1, 32, 0, to_stream(st)>>>(dlam, q1, k1, q2, k2, hd, g1, g2, g3, g4);
  CENET_LAUNCH_CHECK("lambda_bwd");
  return 0;
}

extern "C" int cenet_diff_rmsnorm_fwd(const void* Om, int dtype, const float* lam, void* o, long long rows, int heads, int seg,
                                      float eps, float mult, cenet_stream_t st) {
  CENET_REQUIRE(Om && lam && o, "cenet_diff_rmsnorm_fwd: null pointer");
  if (rows == 0) return 0;
  CENET_REQUIRE(seg % 8 == 0 && seg <= 8 * 32 * DR_MAXV && ((((uintptr_t)Om | (uintptr_t)o) & 15) == 0),
                "cenet_diff_rmsnorm_fwd: segment of %d elements (needs a multiple of 8, <= %d, 16-byte aligned rows)", seg, 8 * 32 * DR_MAXV);
  int lps = 1;
  while (lps < 32 && lps * DR_MAXV < seg / 8) lps <<= 1;            // lanes per segment
  CENET_DISPATCH(dtype, T, (diff_rmsnorm_fwd_kernel<T><<<cdiv(rows * heads * lps, 256), 256, 0, to_stream(st)>>>(
                               (const T*)Om, lam, (T*)o, rows, heads, seg, lps, eps, mult)));
  CENET_LAUNCH_CHECK("diff_rmsnorm_fwd");
  return 0;
}
extern "C" int cenet_diff_rmsnorm_bwd(const void* dO, const void* Om, int dtype, const float* lam, void* dOm, float* dlam,
                                      long long rows, int heads, int seg, float eps, float mult, float* ws, long long ws_elems,
                                      cenet_stream_t st) {
  CENET_REQUIRE(dO && Om && lam && dOm && dlam && ws, "cenet_diff_rmsnorm_bwd: null pointer");
  CENET_REQUIRE(seg % 8 == 0 && seg <= 8 * 32 * DR_MAXV && ((((uintptr_t)Om | (uintptr_t)dO | (uintptr_t)dOm) & 15) == 0),
                "cenet_diff_rmsnorm_bwd: segment of %d elements (needs a multiple of 8, <= %d, 16-byte aligned rows)", seg, 8 * 32 * DR_MAXV);
  int lps = 1;
  while (lps < 32 && lps * DR_MAXV < seg / 8) lps <<= 1;            // lanes per segment
  const int nb = cdiv(rows * heads * lps, 256);
  CENET_REQUIRE(nb <= ws_elems, "cenet_diff_rmsnorm_bwd: workspace too small");
  cudaStream_t s = to_stream(st);
  CENET_DISPATCH(dtype, T, (diff_rmsnorm_bwd_kernel<T><<<nb, 256, 0, s>>>((const T*)dO, (const T*)Om, lam, (T*)dOm, rows, heads, seg,
                                                                          lps, eps, mult, ws)));
  CENET_LAUNCH_CHECK("diff_rmsnorm_bwd");
  sum_partials_kernel<<<1, 256, 0, s>>>(ws, nb, dlam);
  CENET_LAUNCH_CHECK("sum_partials");
  return 0;
}
