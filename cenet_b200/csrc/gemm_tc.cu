// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T )      bf16 operands, fp32 accumulation in TMEM
//
// Structure (one CTA = one 128 x BN output tile, 192 threads):
//   warp 0      : TMA producer.  A and W k-blocks -> shared memory (SWIZZLE_128B / 64B), mbarrier complete_tx.
//   warp 1      : TMEM allocation + single-thread tcgen05.mma issue (UMMA 128 x BN x 16, kind::f16), tcgen05.commit
//                 releases smem stages and finally signals the epilogue.
//   warps 2..5  : epilogue.  tcgen05.ld (32 lanes x 32 columns per instruction) -> bias / BN-folded scale /
//                 activation / gating / residuals (common.cuh::epi_value) -> global store.
// Several CTAs are resident per SM (smem and TMEM are sized per problem), so one CTA's epilogue overlaps another's
// loads and MMAs -- most GEMMs of this network have K = 64..512, i.e. one to eight k-blocks, and are HBM-bound.
//
// Convolution mode (stride 1, "same" padding): the M tile is an 8 x 16 pixel patch of one image; for each filter
// tap the A k-block is the SAME 4-D TMA box shifted by (kh-pad, kw-pad); out-of-image elements are zero-filled by
// the TMA unit, so there is no im2col buffer and no bounds logic in the kernel.
#include "common.cuh"
#include <cuda.h>
#include <mutex>

namespace {
constexpr int BM = 128;
constexpr int NTHREADS = 192;
constexpr int TILE_H = 8, TILE_W = 16;   // conv-mode M tile (pixels)

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major canonical layout (cute/arch/mma_sm100_desc.hpp field layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 | [46,48) version=1 |
//   [61,64) layout type (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
}

struct TcParams {
  int M, N, K;
  int bn;            // N tile (multiple of 16, <= 256)
  int bk;            // k-block in elements: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
  int stages;
  int tmem_cols;     // power of two >= 32
  int a_k0;          // K coordinate offset of A (alignment workaround for channel slices)
  int num_kb;        // k-blocks
  // conv mode
  int conv, H, W, Cin, KH, KW, pad, tiles_h, tiles_w, cblks;
  EpiParams epi;
  int epi_vec;       // every epilogue operand is 16-byte addressable -> smem-transposed, vectorised epilogue
};

__global__ void __launch_bounds__(NTHREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // dynamic smem base is only guaranteed 16-byte aligned: round up to 1024 for the swizzle atoms
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = BM * p.bk * 2, w_bytes = p.bn * p.bk * 2;
  const uint32_t stage_bytes = a_bytes + ((w_bytes + 1023u) & ~1023u);
  const uint32_t bar_base = sbase + p.stages * stage_bytes;          // full[stages], empty[stages], tmem_full
  const uint32_t tmem_slot = bar_base + (2 * p.stages + 1) * 8;
  const uint32_t slab_base = (tmem_slot + 4 + 15u) & ~15u;           // 4 epilogue warps x 32 x 36 floats
  auto full_bar = [&](int s) { return bar_base + s * 8; };
  auto empty_bar = [&](int s) { return bar_base + (p.stages + s) * 8; };
  const uint32_t tmem_full_bar = bar_base + 2 * p.stages * 8;

  // ---- tile coordinates ----
  const int n0 = blockIdx.y * p.bn;
  int m0 = 0, img = 0, h0 = 0, w0 = 0;
  if (p.conv) {
    int t = blockIdx.x;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h;
    img = t / p.tiles_h;
    h0 = th * TILE_H; w0 = tw * TILE_W;
  } else {
    m0 = blockIdx.x * BM;
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                 : "memory");
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
      for (int kb = 0; kb < p.num_kb; kb++) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t sa = sbase + s * stage_bytes, sw = sa + a_bytes;
        mbar_arrive_expect_tx(full_bar(s), a_bytes + w_bytes);
        if (p.conv) {
          const int tap = kb / p.cblks, cb = kb % p.cblks;
          const int kh = tap / p.KW, kw = tap % p.KW;
          tma_load_4d(sa, &tmA, full_bar(s), cb * p.bk, w0 + kw - p.pad, h0 + kh - p.pad, img);
          tma_load_2d(sw, &tmW, full_bar(s), tap * p.Cin + cb * p.bk, n0);
        } else {
          tma_load_2d(sa, &tmA, full_bar(s), p.a_k0 + kb * p.bk, m0);
          tma_load_2d(sw, &tmW, full_bar(s), kb * p.bk, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=f32 (1<<4), A=bf16 (1<<7), B=bf16 (1<<10), K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t layout_type = p.bk == 64 ? 2u : 4u;
      const uint32_t sbo = p.bk == 64 ? 1024u : 512u;
      for (int kb = 0; kb < p.num_kb; kb++) {
        const int s = kb % p.stages;
        const uint32_t ph = (kb / p.stages) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = sbase + s * stage_bytes, sw = sa + a_bytes;
        const uint64_t adesc = make_smem_desc(sa, sbo, layout_type), bdesc = make_smem_desc(sw, sbo, layout_type);
        for (int k = 0; k < p.bk / 16; k++) {
          // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the 16-byte start-address field
          umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
        }
        umma_commit(empty_bar(s));                       // frees the smem stage when these MMAs retire
        if (kb == p.num_kb - 1) umma_commit(tmem_full_bar);   // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
    // tcgen05.ld gives each thread 32 consecutive columns of ITS row.  Row-per-thread global accesses would touch 32
    // different lines per instruction, so the tile is transposed through a per-warp smem slab: phase A (row per
    // thread) applies alpha / row_scale / bias / act and writes the slab; phase B re-reads it with 4 lanes per row,
    // 8 columns (16 bytes of bf16) per lane, so residual / gate loads and the output stores are full 32-byte sectors.
    const int quarter = warp & 3;
    const int r_own = quarter * 32 + lane;      // accumulator row == TMEM lane
    float* slab = reinterpret_cast<float*>(smem_raw + (slab_base - smem_u32(smem_raw))) + (warp - 2) * (32 * 36);
    auto row_index = [&](int r, long long& m) -> bool {
      if (p.conv) {
        const int h = h0 + r / TILE_W, w = w0 + r % TILE_W;
        m = ((long long)img * p.H + h) * p.W + w;
        return (h < p.H) && (w < p.W);
      }
      m = (long long)m0 + r;
      return m < p.M;
    };
    long long m_own;
    const bool own_ok = row_index(r_own, m_own);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const EpiParams& e = p.epi;
    const int nlim = min(p.N, n0 + p.bn);       // columns owned by this tile (bn % 32 may be 16)
    for (int c0 = 0; c0 < p.bn; c0 += 32) {
      if (n0 + c0 >= p.N) break;                 // warp-uniform
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, acc);
      tmem_ld_wait();
      const int nbase = n0 + c0;
      if (p.epi_vec) {
        // ---- phase A: raw accumulators into the slab (row stride 36 floats: 16-byte aligned, conflict-free) ----
#pragma unroll
        for (int q = 0; q < 8; q++)
          *reinterpret_cast<float4*>(slab + lane * 36 + q * 4) =
              make_float4(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]), __uint_as_float(acc[4 * q + 2]),
                          __uint_as_float(acc[4 * q + 3]));
        __syncwarp();
        // ---- phase B: 4 lanes per row, 8 columns per lane; the whole epilogue on 8-wide vectors ----
        const int cg = (lane & 3) * 8;
        const int n = nbase + cg;
        float bcol[8];
#pragma unroll
        for (int j = 0; j < 8; j++) bcol[j] = (e.bias && !e.bias_per_row && n + j < p.N) ? e.bias[n + j] : 0.f;
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
          const int rr = it * 8 + (lane >> 2);
          long long m;
          const bool ok = row_index(quarter * 32 + rr, m);
          if (!ok || n >= nlim) continue;
          float v[8];
          {
            const float4 lo = *reinterpret_cast<const float4*>(slab + rr * 36 + cg);
            const float4 hi = *reinterpret_cast<const float4*>(slab + rr * 36 + cg + 4);
            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
          }
          const float rs = e.alpha * (e.row_scale ? e.row_scale[m] : 1.f);
          const float brow = (e.bias && e.bias_per_row) ? e.bias[m] : 0.f;
#pragma unroll
          for (int j = 0; j < 8; j++) v[j] = fmaf(v[j], rs, bcol[j] + brow);
          if (!e.act_after_res && e.act != CENET_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = apply_act(v[j], e.act, e.slope);
          }
          const bool full = n + 8 <= nlim;
          if (e.mul) {
            float t[8];
            if (full) ldv<8>(reinterpret_cast<const bf16*>(e.mul) + m * e.ldmul + n, t);
            else for (int j = 0; j < 8; j++) t[j] = n + j < nlim ? ld_any(e.mul, e.mul_dtype, m * e.ldmul + n + j) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] *= apply_act(t[j], e.mul_act, 0.f);
          }
          if (e.res1) {
            float t[8];
            if (full) ldv<8>(reinterpret_cast<const bf16*>(e.res1) + m * e.ldr1 + n, t);
            else for (int j = 0; j < 8; j++) t[j] = n + j < nlim ? ld_any(e.res1, e.res1_dtype, m * e.ldr1 + n + j) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] += t[j] * (e.res1_cscale ? (n + j < p.N ? e.res1_cscale[n + j] : 0.f) : e.res1_scale);
          }
          if (e.res2) {
            float t[8];
            if (full) ldv<8>(reinterpret_cast<const bf16*>(e.res2) + m * e.ldr2 + n, t);
            else for (int j = 0; j < 8; j++) t[j] = n + j < nlim ? ld_any(e.res2, e.res2_dtype, m * e.ldr2 + n + j) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] += t[j];
          }
          if (e.act_after_res && e.act != CENET_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = apply_act(v[j], e.act, e.slope);
          }
          if (full) {
            if (e.c_dtype == CENET_BF16) stv<8>(reinterpret_cast<bf16*>(e.C) + m * e.ldc + n, v);
            else stv<8>(reinterpret_cast<float*>(e.C) + m * e.ldc + n, v);
          } else {
            for (int j = 0; j < 8 && n + j < nlim; j++) epi_store(e, v[j], m, n + j, 0);
          }
        }
        __syncwarp();
      } else {
        // operands that are not 16-byte addressable (odd pitches / channel-slice outputs): element-wise epilogue
#pragma unroll
        for (int q = 0; q < 8; q++)
          *reinterpret_cast<float4*>(slab + lane * 36 + q * 4) =
              make_float4(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]), __uint_as_float(acc[4 * q + 2]),
                          __uint_as_float(acc[4 * q + 3]));
        __syncwarp();
        if (own_ok) {
#pragma unroll 1
          for (int c = 0; c < 32; c++) {
            const int n = nbase + c;
            if (n < nlim) epi_store(e, epi_value(e, slab[lane * 36 + c], m_own, n, 0), m_own, n, 0);
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---- host side: tensor-map encoding through the driver entry point (no link-time libcuda dependency) -------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
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

int encode_map(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  CENET_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CENET_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu)", (int)r,
                rank, (unsigned long long)dims[0], (unsigned long long)dims[1]);
  return 0;
}

int pick_bn(int N) {
  const int nt = (N + 255) / 256;
  int bn = ((N + nt - 1) / nt + 15) / 16 * 16;
  if (bn < 16) bn = 16;
  return bn;
}
}  // namespace

bool cenet_gemm_tc_eligible(const cenet_gemm_args* a) {
  if (a->a_dtype != CENET_BF16 || a->w_dtype != CENET_BF16) return false;
  if (a->w_nmajor || a->batch != 1) return false;
  if (a->ldw % 8 != 0 || ((uintptr_t)a->Wt & 15)) return false;
  if (a->conv) {
    // stride-1 "same" convolutions whose channel count fills a swizzle atom
    if (a->stride != 1 || a->KH != a->KW || a->pad != a->KH / 2 || a->Ho != a->H || a->Wo != a->W) return false;
    if (!(a->Cin == 32 || a->Cin % 64 == 0)) return false;
    if (a->lda != a->Cin || ((uintptr_t)a->A & 15)) return false;
    return true;
  }
  // TMA: 16-byte aligned base and row pitch; K padded to 8 elements by the caller (zero columns)
  if (a->lda % 8 != 0 || a->K % 8 != 0 || ((uintptr_t)a->A & 15)) return false;
  return true;
}

int cenet_gemm_tc(const cenet_gemm_args* a, cudaStream_t s) {
  TcParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bn = pick_bn(a->N);
  p.conv = a->conv;
  p.epi = make_epi(a);
  p.a_k0 = 0;
  CUtensorMap tmA, tmW;
  dim3 grid;
  if (a->conv) {
    p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.KH = a->KH; p.KW = a->KW; p.pad = a->pad;
    p.bk = a->Cin == 32 ? 32 : 64;
    p.cblks = a->Cin / p.bk;
    p.num_kb = a->KH * a->KW * p.cblks;
    p.tiles_h = cdiv(a->H, TILE_H); p.tiles_w = cdiv(a->W, TILE_W);
    const CUtensorMapSwizzle swz = p.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint64_t dims[4] = {(cuuint64_t)a->Cin, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->Bimg};
    cuuint64_t str[3] = {(cuuint64_t)a->Cin * 2, (cuuint64_t)a->W * a->Cin * 2, (cuuint64_t)a->H * a->W * a->Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.bk, TILE_W, TILE_H, 1};
    if (encode_map(&tmA, a->A, 4, dims, str, box, swz)) return -1;
    cuuint64_t wd[2] = {(cuuint64_t)a->K, (cuuint64_t)a->N};
    cuuint64_t ws[1] = {(cuuint64_t)a->ldw * 2};
    cuuint32_t wb[2] = {(cuuint32_t)p.bk, (cuuint32_t)p.bn};
    if (encode_map(&tmW, a->Wt, 2, wd, ws, wb, swz)) return -1;
    grid = dim3(a->Bimg * p.tiles_h * p.tiles_w, cdiv(a->N, p.bn), 1);
  } else {
    p.H = p.W = p.Cin = p.KH = p.KW = p.pad = p.tiles_h = p.tiles_w = p.cblks = 0;
    p.bk = 64;
    const void* abase = a->A;
    p.num_kb = cdiv(a->K, 64);
    cuuint64_t ad[2] = {(cuuint64_t)a->K, (cuuint64_t)a->M};
    cuuint64_t as[1] = {(cuuint64_t)a->lda * 2};
    cuuint32_t ab[2] = {64, BM};
    if (encode_map(&tmA, abase, 2, ad, as, ab, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    cuuint64_t wd[2] = {(cuuint64_t)a->K, (cuuint64_t)a->N};
    cuuint64_t ws[1] = {(cuuint64_t)a->ldw * 2};
    cuuint32_t wb[2] = {64, (cuuint32_t)p.bn};
    if (encode_map(&tmW, a->Wt, 2, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    grid = dim3(cdiv(a->M, BM), cdiv(a->N, p.bn), 1);
  }
  CENET_REQUIRE(grid.y <= 65535, "cenet_gemm_tc: too many N tiles");
  const int a_bytes = BM * p.bk * 2, w_bytes = p.bn * p.bk * 2;
  const int stage_bytes = a_bytes + ((w_bytes + 1023) & ~1023);
  int stages = p.num_kb < 4 ? p.num_kb : 4;
  while (stages > 1 && stages * stage_bytes > 96 * 1024) stages--;
  p.stages = stages;
  int cols = 32;
  while (cols < p.bn) cols <<= 1;
  p.tmem_cols = cols;
  auto vec_ok = [](const void* ptr, int dtype, long long ld) {
    return ptr == nullptr || (dtype == CENET_BF16 && ld % 8 == 0 && ((uintptr_t)ptr & 15) == 0);
  };
  const bool c_ok = (((uintptr_t)a->C & 15) == 0) && ((a->c_dtype == CENET_BF16 && a->ldc % 8 == 0) ||
                                                      (a->c_dtype == CENET_F32 && a->ldc % 4 == 0));
  p.epi_vec = c_ok && vec_ok(a->res1, a->res1_dtype, a->ldr1) && vec_ok(a->res2, a->res2_dtype, a->ldr2) &&
              vec_ok(a->mul, a->mul_dtype, a->ldmul);
  const size_t smem = (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 32 + 4 * 32 * 36 * 4 + 1024;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  gemm_tc_kernel<<<grid, NTHREADS, smem, s>>>(tmA, tmW, p);
  CENET_LAUNCH_CHECK("gemm_tc");
  return 0;
}
