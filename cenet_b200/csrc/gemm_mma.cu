// Batched bf16 tensor-core GEMM (mma.sync m16n8k16, fp32 accumulate) for the operand layouts the tcgen05 kernel does not take:
// batched problems (batch / batch_inner strides), A stored [K, M] (a_mmajor) and W stored [K, N] (w_nmajor).
//
// It serves the MATERIALISED attention of the training path where no flash kernel is instantiated -- the non-local block at
// C = 320 / 512 (nlb.py:116-137: 196 / 49 tokens per image) and odd head dims of the differential attention: S = Q K^T,
// O = P V, dP = dO V^T, dV = P^T dO, dQ = dS K, dK = dS^T Q, i.e. all four transpose combinations, 24-64 small problems per
// launch -- and the few small GEMMs whose pitches are not TMA-able.  Before this kernel those contractions ran on the
// CUDA-core GEMM (gemm_simt.cu): 60-100 us per launch at batch 24.
//
// Tile 64 x 64 x 32, 128 threads (4 warps, 32 x 32 each).  An operand that is contiguous along the contraction is staged as
// [rows][k] and read with ldmatrix; one that is contiguous along M / N is staged as [k][rows] and read with ldmatrix.trans.
// Global loads are 16-byte vectors where base, pitch and extent allow, element-wise otherwise; the next tile is fetched into
// registers while the current one is multiplied.  Epilogue = the generic per-element epilogue of the C ABI.
#include "common.cuh"

namespace {
constexpr int TM = 64, TN = 64, TK = 32, NT = 128;
constexpr int PK = TK + 8;      // pitch of a [rows][k] tile (conflict-free ldmatrix)
constexpr int PR = TM + 8;      // pitch of a [k][rows] tile

struct MmaParams {
  int M, N, K, batch_inner;
  const bf16* A; long long lda, a_bso, a_bsi; int a_mmajor, a_vec;
  const bf16* W; long long ldw, w_bso, w_bsi; int w_nmajor, w_vec;
  long long c_bso, c_bsi;
  EpiParams epi;
};

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const bf16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const bf16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 8 consecutive elements of a row-major matrix starting at (r, c): rows >= R or columns >= C read as zero
__device__ __forceinline__ uint4 load8(const bf16* base, long long ld, int r, int R, int c, int C, int vec) {
  uint4 u = make_uint4(0, 0, 0, 0);
  if (r >= R || c >= C) return u;
  const bf16* p = base + (long long)r * ld + c;
  if (vec && c + 8 <= C) return *reinterpret_cast<const uint4*>(p);
  bf16 t[8];
#pragma unroll
  for (int i = 0; i < 8; i++) t[i] = c + i < C ? p[i] : __float2bfloat16_rn(0.f);
  return *reinterpret_cast<uint4*>(t);
}

template <bool AT, bool BT>
__global__ void __launch_bounds__(NT) gemm_mma_kernel(const MmaParams p) {
  // A tile: AT ? [TK][PR] : [TM][PK];  B tile: BT ? [TK][PR] : [TN][PK]   (two buffers each)
  constexpr int A_ELEMS = AT ? TK * PR : TM * PK, B_ELEMS = BT ? TK * PR : TN * PK;
  __shared__ __align__(16) bf16 sA[2][A_ELEMS];
  __shared__ __align__(16) bf16 sB[2][B_ELEMS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int z = blockIdx.z, zo = z / p.batch_inner, zi = z % p.batch_inner;
  const bf16* A = p.A + zo * p.a_bso + zi * p.a_bsi;
  const bf16* W = p.W + zo * p.w_bso + zi * p.w_bsi;
  const long long coff = zo * p.c_bso + zi * p.c_bsi;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) acc[i][j][q] = 0.f;

  // 256 chunks of 8 elements per tile, 2 per thread
  uint4 ra[2], rb[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int ch = tid + i * NT;
      if constexpr (AT) {      // stored [K][M]: chunk = (k row, 8 m)
        const int kr = ch >> 3, mc = (ch & 7) * 8;
        ra[i] = load8(A, p.lda, k0 + kr, p.K, m0 + mc, p.M, p.a_vec);
      } else {                 // stored [M][K]: chunk = (m row, 8 k)
        const int r = ch >> 2, kc = (ch & 3) * 8;
        ra[i] = load8(A, p.lda, m0 + r, p.M, k0 + kc, p.K, p.a_vec);
      }
      if constexpr (BT) {
        const int kr = ch >> 3, nc = (ch & 7) * 8;
        rb[i] = load8(W, p.ldw, k0 + kr, p.K, n0 + nc, p.N, p.w_vec);
      } else {
        const int r = ch >> 2, kc = (ch & 3) * 8;
        rb[i] = load8(W, p.ldw, n0 + r, p.N, k0 + kc, p.K, p.w_vec);
      }
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int ch = tid + i * NT;
      if constexpr (AT) *reinterpret_cast<uint4*>(&sA[buf][(ch >> 3) * PR + (ch & 7) * 8]) = ra[i];
      else *reinterpret_cast<uint4*>(&sA[buf][(ch >> 2) * PK + (ch & 3) * 8]) = ra[i];
      if constexpr (BT) *reinterpret_cast<uint4*>(&sB[buf][(ch >> 3) * PR + (ch & 7) * 8]) = rb[i];
      else *reinterpret_cast<uint4*>(&sB[buf][(ch >> 2) * PK + (ch & 3) * 8]) = rb[i];
    }
  };
  fetch(0);
  commit(0);
  __syncthreads();
  int buf = 0;
  const int j = lane >> 3, r = lane & 7;
  for (int k0 = 0; k0 < p.K; k0 += TK) {
    const bool more = k0 + TK < p.K;
    if (more) fetch(k0 + TK);
#pragma unroll
    for (int ks = 0; ks < TK; ks += 16) {
      uint32_t a[2][4], b[4][2];
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const int mb = wm + i * 16;
        if constexpr (AT)        // matrices (m 0-7, k 0-7), (m 8-15, k 0-7), (m 0-7, k 8-15), (m 8-15, k 8-15) from [k][m]
          ldsm_x4_t(a[i][0], a[i][1], a[i][2], a[i][3], &sA[buf][(ks + r + ((j & 2) ? 8 : 0)) * PR + mb + ((j & 1) ? 8 : 0)]);
        else
          ldsm_x4(a[i][0], a[i][1], a[i][2], a[i][3], &sA[buf][(mb + r + ((j & 1) ? 8 : 0)) * PK + ks + ((j & 2) ? 8 : 0)]);
      }
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const int nb = wn + i * 16;
        uint32_t t0, t1, t2, t3;
        if constexpr (BT)        // from [k][n]: (k 0-7, n 0-7), (k 8-15, n 0-7), (k 0-7, n 8-15), (k 8-15, n 8-15)
          ldsm_x4_t(t0, t1, t2, t3, &sB[buf][(ks + r + ((j & 1) ? 8 : 0)) * PR + nb + ((j & 2) ? 8 : 0)]);
        else                     // from [n][k]: (n 0-7, k 0-7), (n 0-7, k 8-15), (n 8-15, k 0-7), (n 8-15, k 8-15)
          ldsm_x4(t0, t1, t2, t3, &sB[buf][(nb + r + ((j & 2) ? 8 : 0)) * PK + ks + ((j & 1) ? 8 : 0)]);
        b[2 * i][0] = t0; b[2 * i][1] = t1; b[2 * i + 1][0] = t2; b[2 * i + 1][1] = t3;
      }
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int q = 0; q < 4; q++) mma_bf16(acc[i][q], a[i][0], a[i][1], a[i][2], a[i][3], b[q][0], b[q][1]);
    }
    if (more) commit(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const long long m = m0 + wm + i * 16 + g + h * 8;
        if (m >= p.M) continue;
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int n = n0 + wn + q * 8 + 2 * t + e;
          if (n < p.N) epi_store(p.epi, epi_value(p.epi, acc[i][q][2 * h + e], m, n, coff), m, n, coff);
        }
      }
}
}  // namespace

bool cenet_gemm_mma_eligible(const cenet_gemm_args* a) {
  return a->a_dtype == CENET_BF16 && a->w_dtype == CENET_BF16 && !a->conv && !a->k_scale;
}

int cenet_gemm_mma(const cenet_gemm_args* a, cudaStream_t s) {
  MmaParams p;
  p.M = a->M; p.N = a->N; p.K = a->K; p.batch_inner = a->batch_inner;
  p.A = (const bf16*)a->A; p.lda = a->lda; p.a_bso = a->a_bs_outer; p.a_bsi = a->a_bs_inner; p.a_mmajor = a->a_mmajor;
  p.W = (const bf16*)a->Wt; p.ldw = a->ldw; p.w_bso = a->w_bs_outer; p.w_bsi = a->w_bs_inner; p.w_nmajor = a->w_nmajor;
  p.c_bso = a->c_bs_outer; p.c_bsi = a->c_bs_inner;
  p.epi = make_epi(a);
  // 16-byte vector loads: base, pitch and every batch offset must keep rows 16-byte aligned
  auto vec = [](const void* ptr, long long ld, long long bso, long long bsi) {
    return (((uintptr_t)ptr & 15) == 0 && ld % 8 == 0 && bso % 8 == 0 && bsi % 8 == 0) ? 1 : 0;
  };
  p.a_vec = vec(a->A, a->lda, a->a_bs_outer, a->a_bs_inner);
  p.w_vec = vec(a->Wt, a->ldw, a->w_bs_outer, a->w_bs_inner);
  dim3 grid(cdiv(a->M, TM), cdiv(a->N, TN), a->batch);
  CENET_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "cenet_gemm_mma: grid too large (N=%d batch=%d)", a->N, a->batch);
  if (a->a_mmajor && a->w_nmajor) gemm_mma_kernel<true, true><<<grid, NT, 0, s>>>(p);
  else if (a->a_mmajor) gemm_mma_kernel<true, false><<<grid, NT, 0, s>>>(p);
  else if (a->w_nmajor) gemm_mma_kernel<false, true><<<grid, NT, 0, s>>>(p);
  else gemm_mma_kernel<false, false><<<grid, NT, 0, s>>>(p);
  CENET_LAUNCH_CHECK("gemm_mma");
  return 0;
}
