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
// Halo mode (conv == 2; the OutHead 5x5 / 3x3 convs): the M tile is a 16 x 8 pixel patch and its (16+2p) x (8+2p) HALO is
// fetched ONCE per tile as Cin/8 TMA boxes of 8 channels -> [channel block][halo pixel][16 B] = the tensor core's no-swizzle
// K-major core-matrix layout.  A filter tap is then only a different START ADDRESS of the A descriptor ((kh*Wh + kw) * 16 B;
// 8 pixels of a tile row are one core matrix, tile rows are SBO = Wh*16 B apart, channel blocks LBO apart), so the 25 (9)
// taps re-read shared memory instead of re-fetching the tile from L2 25 (9) times.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <mutex>

namespace {
constexpr int BM = 128;
constexpr int TILE_H = 8, TILE_W = 16;   // conv-mode M tile (pixels); halo mode uses 16 x 8

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

// no-swizzle K-major descriptor: 8-row groups `sbo` bytes apart, 16-byte K chunks `lbo` bytes apart
__device__ __forceinline__ uint64_t make_smem_desc_ns(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&o)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) { const float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}

struct TcParams {
  int M, N, K;
  int bn;            // N tile (multiple of 16, <= 256)
  int bk;            // k-block in elements: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
  int stages;        // depth of the A (or A+W) ring
  int tmem_cols;     // power of two >= 2*acc_stride (double-buffered accumulator)
  int acc_stride;    // TMEM columns per accumulator stage = bn rounded up to 32 (tcgen05.ld reads 32-column chunks)
  int num_kb;        // k-blocks per tile
  int num_m_tiles;   // tiles along M (conv: images * tiles_h * tiles_w)
  int w_resident;    // the whole [bn x K] weight slice stays in shared memory for the CTA's lifetime
  int n_epi;         // epilogue warps: 4 (one per TMEM lane quarter) or 8 (two per quarter, alternating column chunks)
  // conv mode
  int conv, H, W, Cin, KH, KW, pad, tiles_h, tiles_w, cblks;   // H, W: OUTPUT extent (= input extent for the stride-1 'same' convs)
  int stride;        // conv stride: the tap boxes are TMA boxes with element strides (stride, stride) over the input image
  int tile_h, tile_w;  // M tile in pixels: 8 x 16 (shifted boxes) or 16 x 8 (halo mode)
  int tw_shift;        // log2(tile_w)
  int halo, Hh, Wh;    // halo mode: halo tile extent
  int ablk;            // bytes of one 8-channel block of a halo tile: Hh*Wh*16 rounded up to 128 (TMA destination alignment)
  EpiParams epi;
  int epi_vec;       // every epilogue operand is 16-byte addressable -> smem-transposed, vectorised epilogue
  int epi_fast;      // 1: C = bf16(acc + bias); 2: C = bf16(acc + bias + res1)  (most Linear layers) -- compact code path
  // split-K (plain GEMM only): blockIdx.z owns k-blocks [z * kb_per_split, +kb_per_split) and stores its raw fp32 accumulator
  // tile into split_ws + z * M * N; splitk_reduce_kernel sums the slices in fixed order and applies bias / cast
  int splits, kb_per_split;
  float* split_ws;
};

constexpr int MAX_EPI_WARPS = 8;
constexpr int NTHREADS_P = 64 + 32 * MAX_EPI_WARPS;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent, warp-specialised: grid = (CTAs along M, N tiles); CTA c walks M tiles c, c+G, c+2G, ...
//   warp 0   : TMA producer (weights once if resident, then the A ring)
//   warp 1   : TMEM alloc + tcgen05.mma issue into accumulator stage (tile & 1)
//   warps 2-9: epilogue of the previous tile out of the other accumulator stage, overlapping the next tile's loads+MMAs
// Vector epilogue of one 32x32 accumulator block for the 4 rows x 8 columns a lane owns, specialised at COMPILE time on which
// operands exist (FLAGS): the all-runtime version executed ~950 instructions per block, most of them flag tests, constant
// reloads and branches, at an IPC of 0.35 (two epilogue warps per scheduler cannot hide dependent-issue latency), which made a
// ReLU cost 2.7x the plain store.  Here the four rows are independent instruction chains the scheduler can interleave, every
// operand row is requested before the first use, and the activation code exists once (two-pass loop).
enum { VF_MUL = 1, VF_R1 = 2, VF_R2 = 4, VF_ROW = 8, VF_R1F32 = 16 };   // R1F32: res1 is an fp32 row (encoder residual stream)
__device__ __forceinline__ void act32(float (&v)[4][8], int act, float slope) {
  switch (act) {
    case CENET_ACT_RELU:
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int j = 0; j < 8; j++) v[r][j] = fmaxf(v[r][j], 0.0f);
      break;
    case CENET_ACT_LEAKY:
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int j = 0; j < 8; j++) v[r][j] = v[r][j] > 0.0f ? v[r][j] : v[r][j] * slope;
      break;
    case CENET_ACT_SILU:
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int j = 0; j < 8; j++) v[r][j] = v[r][j] * sigmoid_fast(v[r][j]);
      break;
    case CENET_ACT_SIGMOID:
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int j = 0; j < 8; j++) v[r][j] = sigmoid_fast(v[r][j]);
      break;
    case CENET_ACT_GELU:
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int j = 0; j < 8; j++) v[r][j] = gelu_erf(v[r][j]);
      break;
    case CENET_ACT_GELU_GRAD:
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int j = 0; j < 8; j++) v[r][j] = gelu_grad_fast(v[r][j]);
      break;
    default: break;
  }
}

template <int FLAGS>
__device__ __forceinline__ void epi_vec_block(const EpiParams& e, const float* slab, int lane, int n, const long long (&mrow)[4],
                                              const bool (&okrow)[4], const float (&rs)[4], const float (&prs)[4],
                                              const float (&brow)[4]) {
  const int cg = (lane & 3) * 8;
  uint4 um[4], u1[4], u2[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const long long m = mrow[r];                                     // 0 for rows outside the matrix: loads stay in bounds
    if (FLAGS & VF_MUL) um[r] = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(e.mul) + m * e.ldmul + n);
    if (FLAGS & VF_R1F32) {                          // (never together with R2: u2 holds the second half of the fp32 row)
      const float* rp = reinterpret_cast<const float*>(e.res1) + m * e.ldr1 + n;
      u1[r] = *reinterpret_cast<const uint4*>(rp);
      u2[r] = *reinterpret_cast<const uint4*>(rp + 4);
    } else {
      if (FLAGS & VF_R1) u1[r] = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(e.res1) + m * e.ldr1 + n);
      if (FLAGS & VF_R2) u2[r] = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(e.res2) + m * e.ldr2 + n);
    }
  }
  float bcol[8], cs1[8];
  if (e.bias && !e.bias_per_row) ldv<8>(e.bias + n, bcol);
  else {
#pragma unroll
    for (int j = 0; j < 8; j++) bcol[j] = 0.f;
  }
  if (FLAGS & (VF_R1 | VF_R1F32)) {
    if (e.res1_cscale) ldv<8>(e.res1_cscale + n, cs1);
    else {
#pragma unroll
      for (int j = 0; j < 8; j++) cs1[j] = e.res1_scale;
    }
  }
  float v[4][8];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int rr = r * 8 + (lane >> 2);
    const float4 lo = *reinterpret_cast<const float4*>(slab + rr * 36 + cg);
    const float4 hi = *reinterpret_cast<const float4*>(slab + rr * 36 + cg + 4);
    const float a[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int j = 0; j < 8; j++) v[r][j] = fmaf(a[j], rs[r], bcol[j] + brow[r]);
  }
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    if (e.act != CENET_ACT_NONE && pass == (e.act_after_res ? 1 : 0)) act32(v, e.act, e.slope);
    if (pass) break;
    if (FLAGS & VF_MUL) {
      const int ma = e.mul_act;                     // GELU' (Mix-FFN dgrad), SiLU (gated CFAM branch) or none on this path
#pragma unroll
      for (int r = 0; r < 4; r++) {
        float t[8];
        unpack8(um[r], t);
        if (ma == CENET_ACT_GELU_GRAD) {
#pragma unroll
          for (int j = 0; j < 8; j++) t[j] = gelu_grad_fast(t[j]);
        } else if (ma == CENET_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 8; j++) t[j] = t[j] * sigmoid_fast(t[j]);
        } else if (ma != CENET_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 8; j++) t[j] = apply_act(t[j], ma, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) v[r][j] *= t[j];
      }
    }
    if (FLAGS & VF_ROW) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int j = 0; j < 8; j++) v[r][j] *= prs[r];
    }
    if (FLAGS & VF_R1F32) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const float t[8] = {__uint_as_float(u1[r].x), __uint_as_float(u1[r].y), __uint_as_float(u1[r].z), __uint_as_float(u1[r].w),
                            __uint_as_float(u2[r].x), __uint_as_float(u2[r].y), __uint_as_float(u2[r].z), __uint_as_float(u2[r].w)};
#pragma unroll
        for (int j = 0; j < 8; j++) v[r][j] = fmaf(t[j], cs1[j], v[r][j]);
      }
    } else if (FLAGS & VF_R1) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
        float t[8];
        unpack8(u1[r], t);
#pragma unroll
        for (int j = 0; j < 8; j++) v[r][j] = fmaf(t[j], cs1[j], v[r][j]);
      }
    }
    if ((FLAGS & VF_R2) && !(FLAGS & VF_R1F32)) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
        float t[8];
        unpack8(u2[r], t);
#pragma unroll
        for (int j = 0; j < 8; j++) v[r][j] += t[j];
      }
    }
  }
  if (e.c_dtype == CENET_BF16) {
#pragma unroll
    for (int r = 0; r < 4; r++)
      if (okrow[r]) stv<8>(reinterpret_cast<bf16*>(e.C) + mrow[r] * e.ldc + n, v[r]);
  } else {
#pragma unroll
    for (int r = 0; r < 4; r++)
      if (okrow[r]) stv<8>(reinterpret_cast<float*>(e.C) + mrow[r] * e.ldc + n, v[r]);
  }
}

// EPI selects the ONE epilogue compiled into an instantiation (ncu on the all-in-one kernel: 9 k SASS instructions, the epilogue
// warps' top stall reason was `no_instruction` -- instruction-cache misses -- and a ReLU cost 2.7x the plain store):
enum { EPI_SPLITK = 0, EPI_FAST = 1, EPI_VEC = 3, EPI_SCALAR = 4 };
template <int EPI, int FLAGS>
__global__ void __launch_bounds__(NTHREADS_P, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;     // swizzle atoms need 1024-byte alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = p.halo ? (uint32_t)(((p.Cin / 8) * p.ablk + 1023) & ~1023) : BM * p.bk * 2, w_bytes = p.bn * p.bk * 2;
  const uint32_t wblk = (w_bytes + 1023u) & ~1023u;
  const uint32_t wres_bytes = p.w_resident ? (p.splits > 1 ? p.kb_per_split : p.num_kb) * wblk : 0;   // k-blocks one CTA walks
  const uint32_t stage_bytes = a_bytes + (p.w_resident ? 0 : wblk);
  const uint32_t ring_base = sbase + wres_bytes;
  const uint32_t bar_base = ring_base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + s * 8; };
  auto empty_bar = [&](int s) { return bar_base + (p.stages + s) * 8; };
  const uint32_t tfull_bar = bar_base + 2 * p.stages * 8;            // [2]
  const uint32_t tempty_bar = tfull_bar + 16;                        // [2]
  const uint32_t wfull_bar = tempty_bar + 16;
  const uint32_t tmem_slot = wfull_bar + 8;
  const uint32_t slab_base = (tmem_slot + 4 + 15u) & ~15u;           // n_epi x 32 x 36 floats
  const int n0 = blockIdx.y * p.bn;
  const int kb_begin = p.splits > 1 ? blockIdx.z * p.kb_per_split : 0;
  const int kb_end = p.splits > 1 ? min(p.num_kb, kb_begin + p.kb_per_split) : p.num_kb;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int i = 0; i < 2; i++) { mbar_init(tfull_bar + 8 * i, 1); mbar_init(tempty_bar + 8 * i, p.n_epi); }
    mbar_init(wfull_bar, 1);
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

  auto tile_coords = [&](int tile, int& m0, int& img, int& h0, int& w0) {
    m0 = img = h0 = w0 = 0;
    if (p.conv) {
      int t = tile;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h;
      img = t / p.tiles_h;
      h0 = th * p.tile_h; w0 = tw * p.tile_w;
    } else {
      m0 = tile * BM;
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (p.w_resident) {
        mbar_arrive_expect_tx(wfull_bar, (kb_end - kb_begin) * w_bytes);
        for (int kb = kb_begin; kb < kb_end; kb++) {
          const int kcoord = p.conv ? (kb / p.cblks) * p.Cin + (kb % p.cblks) * p.bk : kb * p.bk;
          tma_load_2d(sbase + (kb - kb_begin) * wblk, &tmW, wfull_bar, kcoord, n0);
        }
      }
      // ring position kept incrementally (no integer division on the issue path: one thread feeds the whole CTA)
      int s = 0;
      uint32_t ph = 0;
      auto advance = [&]() { if (++s == p.stages) { s = 0; ph ^= 1; } };
      for (int tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x) {
        int m0, img, h0, w0;
        tile_coords(tile, m0, img, h0, w0);
        if (p.halo) {                                      // one stage = the whole halo tile, Cin/8 boxes of 8 channels
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t sa = ring_base + s * stage_bytes;
          mbar_arrive_expect_tx(full_bar(s), (uint32_t)((p.Cin / 8) * p.Hh * p.Wh * 16));
          const uint32_t ablk = (uint32_t)p.ablk;
          for (int cb = 0; cb < p.Cin / 8; cb++) tma_load_4d(sa + cb * ablk, &tmA, full_bar(s), cb * 8, w0 - p.pad, h0 - p.pad, img);
          advance();
          continue;
        }
        int tap = 0, cb = 0, kh = 0, kw = 0;               // conv mode: (tap, channel block) of the current k-block
        if (p.conv && kb_begin) { tap = kb_begin / p.cblks; cb = kb_begin % p.cblks; kh = tap / p.KW; kw = tap % p.KW; }
        for (int kb = kb_begin; kb < kb_end; kb++) {
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t sa = ring_base + s * stage_bytes;
          mbar_arrive_expect_tx(full_bar(s), a_bytes + (p.w_resident ? 0 : w_bytes));
          if (p.conv) {
            tma_load_4d(sa, &tmA, full_bar(s), cb * p.bk, w0 * p.stride + kw - p.pad, h0 * p.stride + kh - p.pad, img);
            if (!p.w_resident) tma_load_2d(sa + a_bytes, &tmW, full_bar(s), tap * p.Cin + cb * p.bk, n0);
            if (++cb == p.cblks) { cb = 0; tap++; if (++kw == p.KW) { kw = 0; kh++; } }
          } else {
            tma_load_2d(sa, &tmA, full_bar(s), kb * p.bk, m0);
            if (!p.w_resident) tma_load_2d(sa + a_bytes, &tmW, full_bar(s), kb * p.bk, n0);
          }
          advance();
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
      if (p.w_resident) { mbar_wait(wfull_bar, 0); tc_fence_after(); }
      uint32_t i = 0;
      int s = 0;
      uint32_t ph = 0;
      auto advance = [&]() { if (++s == p.stages) { s = 0; ph ^= 1; } };
      const uint32_t kmma = (uint32_t)p.bk / 16;              // k16 steps per k-block (2 or 4)
      const uint64_t wstep = (uint64_t)(wblk >> 4);           // descriptor start-address units (16 B) per resident weight k-block
      for (int tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x, i++) {
        const uint32_t as = i & 1, use = i >> 1;
        mbar_wait(tempty_bar + 8 * as, (use & 1) ^ 1);       // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * p.acc_stride;
        if (p.halo) {
          // every tap = the same halo tile read from a different start address: descriptors advance by ADDITIONS only
          // (one thread issues all MMAs of the CTA; address arithmetic with divisions had made the issue loop the bottleneck)
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = ring_base + s * stage_bytes;
          const uint64_t kstepA = (uint64_t)((2u * (uint32_t)p.ablk) >> 4);          // two 8-channel blocks per k16 step
          uint64_t a_row = make_smem_desc_ns(sa, (uint32_t)p.ablk, (uint32_t)(p.Wh * 16));
          uint64_t b_tap = make_smem_desc(sbase, sbo, layout_type);
          const int ksteps = p.Cin / 16;
          uint32_t acc = 0;
          for (int kh = 0; kh < p.KH; kh++, a_row += (uint64_t)p.Wh) {
            uint64_t a_tap = a_row;
            for (int kw = 0; kw < p.KW; kw++, a_tap += 1, b_tap += wstep * (uint64_t)p.cblks) {
              uint64_t ad = a_tap, bd = b_tap;
              uint32_t kin = 0;
              for (int k = 0; k < ksteps; k++, ad += kstepA) {
                umma_f16(d_tmem, ad, bd + (uint64_t)(2 * kin), idesc, acc);
                acc = 1;
                if (++kin == kmma) { kin = 0; bd += wstep; }
              }
            }
          }
          umma_commit(empty_bar(s));
          advance();
          umma_commit(tfull_bar + 8 * as);
          continue;
        }
        uint64_t wdesc = make_smem_desc(sbase, sbo, layout_type);                     // resident weights: k-block kb_begin
        for (int kb = kb_begin; kb < kb_end; kb++, wdesc += wstep) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = ring_base + s * stage_bytes;
          const uint64_t adesc = make_smem_desc(sa, sbo, layout_type);
          const uint64_t bdesc = p.w_resident ? wdesc : make_smem_desc(sa + a_bytes, sbo, layout_type);
          for (uint32_t k = 0; k < kmma; k++) {
            // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the 16-byte start-address field
            umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, kb != kb_begin || k != 0);
          }
          umma_commit(empty_bar(s));                           // frees the ring slot when these MMAs retire
          advance();
        }
        umma_commit(tfull_bar + 8 * as);                       // accumulator of this tile complete
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // tcgen05.ld gives each thread 32 consecutive columns of ITS row.  Row-per-thread global accesses would touch 32
    // different lines per instruction, so each 32x32 block is transposed through a per-warp smem slab and re-read with
    // 4 lanes per row, 8 columns (16 bytes of bf16) per lane: residual / gate loads and the output stores are then
    // full 32-byte sectors.  Warps 2..9: TMEM lane quarter = warp % 4, column-chunk parity = (warp - 2) / 4.
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const int cstep = p.n_epi == 8 ? 64 : 32;
    float* slab = reinterpret_cast<float*>(smem_raw + (slab_base - smem_u32(smem_raw))) + ew * (32 * 36);
    const EpiParams& e = p.epi;
    const int nlim = min(p.N, n0 + p.bn);       // columns owned by this CTA (bn % 32 may be 16)
    uint32_t i = 0;
    for (int tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x, i++) {
      int m0, img, h0, w0;
      tile_coords(tile, m0, img, h0, w0);
      auto row_index = [&](int r, long long& m) -> bool {
        if (p.conv) {
          const int h = h0 + (r >> p.tw_shift), w = w0 + (r & (p.tile_w - 1));        // tile_w is 8 or 16
          m = ((long long)img * p.H + h) * p.W + w;
          return (h < p.H) && (w < p.W);
        }
        m = (long long)m0 + r;
        return m < p.M;
      };
      const uint32_t as = i & 1, use = i >> 1;
      long long m_own;
      const bool own_ok = row_index(quarter * 32 + lane, m_own);
      // fast path bookkeeping: the 4 rows this lane serves in phase B (4 lanes per row, 8 rows per pass)
      bf16* crow[4];
      long long mrow[4];
      bool okrow[4];
      float rsv[4], prsv[4], browv[4];
#pragma unroll
      for (int itr = 0; itr < 4; itr++) {
        long long m;
        const bool ok = row_index(quarter * 32 + itr * 8 + (lane >> 2), m);
        mrow[itr] = ok ? m : 0; okrow[itr] = ok;
        // per-row factors do not depend on the column chunk: once per tile (rows < 2^31: 32-bit divisions)
        rsv[itr] = e.alpha; prsv[itr] = 1.f; browv[itr] = 0.f;
        if constexpr (EPI == EPI_VEC && (FLAGS & VF_ROW) != 0) {
          const unsigned mu = (unsigned)mrow[itr];
          if (e.row_scale) rsv[itr] *= e.row_scale[e.rs_div > 1 ? mu / (unsigned)e.rs_div : mu];
          if (e.post_rs) prsv[itr] = e.post_rs[e.post_rs_div > 1 ? mu / (unsigned)e.post_rs_div : mu];
          if (e.bias && e.bias_per_row) browv[itr] = e.bias[mu];
        }
        crow[itr] = ok ? reinterpret_cast<bf16*>(e.C) + m * e.ldc : nullptr;
      }
      mbar_wait(tfull_bar + 8 * as, use & 1);
      tc_fence_after();
      for (int c0 = half * 32; c0 < p.bn; c0 += cstep) {
        if (n0 + c0 >= p.N) break;                 // warp-uniform
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * p.acc_stride + c0), acc);
        tmem_ld_wait();
        const int nbase = n0 + c0;
#pragma unroll
        for (int q = 0; q < 8; q++)
          *reinterpret_cast<float4*>(slab + lane * 36 + q * 4) =
              make_float4(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]), __uint_as_float(acc[4 * q + 2]),
                          __uint_as_float(acc[4 * q + 3]));
        __syncwarp();
        if constexpr (EPI == EPI_SPLITK) {
          // split-K: raw fp32 partial tile -> split_ws[z][m][n]  (N % 8 == 0; bias / cast happen in the reduce kernel)
          const int cg = (lane & 3) * 8;
          const int n = nbase + cg;
          if (n < nlim) {
#pragma unroll
            for (int itr = 0; itr < 4; itr++) {
              long long m;
              if (!row_index(quarter * 32 + itr * 8 + (lane >> 2), m)) continue;
              const int rr = itr * 8 + (lane >> 2);
              float* dst = p.split_ws + ((long long)blockIdx.z * p.M + m) * p.N + n;
              *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(slab + rr * 36 + cg);
              *reinterpret_cast<float4*>(dst + 4) = *reinterpret_cast<const float4*>(slab + rr * 36 + cg + 4);
            }
          }
        } else if constexpr (EPI == EPI_FAST) {
          // C = bf16(acc + bias); N % 8 == 0, so an 8-column group is entirely inside or outside the tile
          const int cg = (lane & 3) * 8;
          const int n = nbase + cg;
          if (n < nlim) {
            float bcol[8];
            if (e.bias) ldv<8>(e.bias + n, bcol);
            else {
#pragma unroll
              for (int j = 0; j < 8; j++) bcol[j] = 0.f;
            }
#pragma unroll
            for (int itr = 0; itr < 4; itr++) {
              if (crow[itr] == nullptr) continue;
              const int rr = itr * 8 + (lane >> 2);
              const float4 lo = *reinterpret_cast<const float4*>(slab + rr * 36 + cg);
              const float4 hi = *reinterpret_cast<const float4*>(slab + rr * 36 + cg + 4);
              float v[8] = {lo.x + bcol[0], lo.y + bcol[1], lo.z + bcol[2], lo.w + bcol[3],
                            hi.x + bcol[4], hi.y + bcol[5], hi.z + bcol[6], hi.w + bcol[7]};
              stv<8>(crow[itr] + n, v);
            }
          }
        } else if constexpr (EPI == EPI_VEC) {
          // general vector epilogue: N % 8 == 0 and every operand 16-byte addressable (the host routes the rest to EPI_SCALAR),
          // so an 8-column group is entirely inside or outside the tile
          const int n = nbase + (lane & 3) * 8;
          if (n < nlim) epi_vec_block<FLAGS>(e, slab, lane, n, mrow, okrow, rsv, prsv, browv);
        } else if (own_ok) {
          // EPI_SCALAR: operands that are not 16-byte addressable (odd pitches / channel-slice outputs): element-wise epilogue
#pragma unroll 1
          for (int c = 0; c < 32; c++) {
            const int n = nbase + c;
            if (n < nlim) epi_store(e, epi_value(e, slab[lane * 36 + c], m_own, n, 0), m_own, n, 0);
          }
        }
        __syncwarp();
      }
      // this warp no longer reads the accumulator stage: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * as);
    }
  }
  tc_fence_before();
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
               const cuuint32_t* box, CUtensorMapSwizzle swz, int pixel_stride = 1) {
  EncodeTiledFn enc = get_encode();
  CENET_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (pixel_stride > 1) estr[1] = estr[2] = (cuuint32_t)pixel_stride;     // strided conv: every stride-th pixel of the W / H box extent
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
  if (a->w_nmajor || a->a_mmajor || a->k_scale || a->batch != 1) return false;
  if (a->ldw % 8 != 0 || ((uintptr_t)a->Wt & 15)) return false;
  if (a->conv) {
    // square filters whose channel count fills a swizzle atom: stride-1 "same" convolutions, and strided ones (patch embeds 3x3 s2,
    // SR convs k = s; pvtv2.py:164-165, 68) through element-strided TMA boxes
    if (a->KH != a->KW || a->stride < 1) return false;
    if (a->stride == 1) {
      if (a->pad != a->KH / 2 || a->Ho != a->H || a->Wo != a->W) return false;
    } else {
      if (a->Ho != (a->H + 2 * a->pad - a->KH) / a->stride + 1 || a->Wo != (a->W + 2 * a->pad - a->KW) / a->stride + 1) return false;
      if ((TILE_W - 1) * a->stride + 1 > 256 || a->Cin % 64 != 0 || a->Ho < 1 || a->Wo < 1) return false;
    }
    if (!(a->Cin == 32 || a->Cin % 64 == 0)) return false;
    if (a->lda != a->Cin || ((uintptr_t)a->A & 15)) return false;
    return true;
  }
  // TMA: 16-byte aligned base and row pitch; K padded to 8 elements by the caller (zero columns)
  if (a->lda % 8 != 0 || a->K % 8 != 0 || ((uintptr_t)a->A & 15)) return false;
  return true;
}

// out[m, n] = bf16( sum_z ws[z][m][n] + bias[n] ), slices added in z order (deterministic); 8 columns per thread
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int S, long long M, int N, const float* __restrict__ bias,
                                                            bf16* __restrict__ C, long long ldc) {
  const int ng = N >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * ng) return;
  const long long m = i / ng;
  const int n = (int)(i % ng) * 8;
  float v[8];
  if (bias) ldv<8>(bias + n, v);
  else {
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = 0.f;
  }
  for (int z = 0; z < S; z++) {
    float t[8];
    ldv<8>(ws + ((long long)z * M + m) * N + n, t);
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] += t[j];
  }
  stv<8>(C + m * ldc + n, v);
}

int cenet_gemm_tc(const cenet_gemm_args* a, cudaStream_t s) {
  TcParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bn = pick_bn(a->N);
  p.conv = a->conv;
  p.epi = make_epi(a);
  CUtensorMap tmA, tmW;
  if (a->conv) {
    p.H = a->Ho; p.W = a->Wo; p.Cin = a->Cin; p.KH = a->KH; p.KW = a->KW; p.pad = a->pad; p.stride = a->stride;
    p.bk = a->Cin == 32 ? 32 : 64;
    p.cblks = a->Cin / p.bk;
    p.num_kb = a->KH * a->KW * p.cblks;
    // halo mode: the whole filter bank resident in shared memory + halo tiles of (16+2p) x (8+2p) pixels
    static const bool halo_on = !(getenv("CENET_B200_CONV_HALO") && atoi(getenv("CENET_B200_CONV_HALO")) == 0);
    const int wblk_h = ((pick_bn(a->N) * p.bk * 2) + 1023) & ~1023;
    // measured (tools/one_conv.py, B=64): 5x5 32->32 0.79 -> 0.67 ms, 3x3 64->32 0.186 -> 0.176 ms, but 3x3 64->64 0.307 -> 0.336 ms
    // (its 72 KB filter bank leaves room for one CTA per SM only) -> halo mode for narrow outputs (N <= 32)
    p.halo = halo_on && a->stride == 1 && a->KH > 1 && a->Cin % 16 == 0 && a->Cin <= 128 && pick_bn(a->N) <= 32 && a->N <= 32 &&
             (long long)p.num_kb * wblk_h <= 120 * 1024;
    p.tile_h = p.halo ? 16 : TILE_H; p.tile_w = p.halo ? 8 : TILE_W; p.tw_shift = p.halo ? 3 : 4;
    p.Hh = p.tile_h + 2 * a->pad; p.Wh = p.tile_w + 2 * a->pad;
    p.ablk = (p.Hh * p.Wh * 16 + 127) & ~127;
    p.tiles_h = cdiv(a->Ho, p.tile_h); p.tiles_w = cdiv(a->Wo, p.tile_w);
    p.num_m_tiles = a->Bimg * p.tiles_h * p.tiles_w;
    const CUtensorMapSwizzle swz = p.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint64_t dims[4] = {(cuuint64_t)a->Cin, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->Bimg};
    cuuint64_t str[3] = {(cuuint64_t)a->Cin * 2, (cuuint64_t)a->W * a->Cin * 2, (cuuint64_t)a->H * a->W * a->Cin * 2};
    if (p.halo) {
      cuuint32_t box[4] = {8, (cuuint32_t)p.Wh, (cuuint32_t)p.Hh, 1};
      if (encode_map(&tmA, a->A, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
    } else {
      cuuint32_t box[4] = {(cuuint32_t)p.bk, (cuuint32_t)((TILE_W - 1) * a->stride + 1), (cuuint32_t)((TILE_H - 1) * a->stride + 1), 1};
      if (encode_map(&tmA, a->A, 4, dims, str, box, swz, a->stride)) return -1;
    }
    cuuint64_t wd[2] = {(cuuint64_t)a->K, (cuuint64_t)a->N};
    cuuint64_t ws[1] = {(cuuint64_t)a->ldw * 2};
    cuuint32_t wb[2] = {(cuuint32_t)p.bk, (cuuint32_t)p.bn};
    if (encode_map(&tmW, a->Wt, 2, wd, ws, wb, swz)) return -1;
  } else {
    p.H = p.W = p.Cin = p.KH = p.KW = p.pad = p.tiles_h = p.tiles_w = p.cblks = 0; p.stride = 1;
    p.halo = 0; p.Hh = p.Wh = p.ablk = 0; p.tile_h = TILE_H; p.tile_w = TILE_W; p.tw_shift = 4;
    p.bk = 64;
    p.num_kb = cdiv(a->K, 64);
    p.num_m_tiles = cdiv(a->M, BM);
    cuuint64_t ad[2] = {(cuuint64_t)a->K, (cuuint64_t)a->M};
    cuuint64_t as[1] = {(cuuint64_t)a->lda * 2};
    cuuint32_t ab[2] = {64, BM};
    if (encode_map(&tmA, a->A, 2, ad, as, ab, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    cuuint64_t wd[2] = {(cuuint64_t)a->K, (cuuint64_t)a->N};
    cuuint64_t ws[1] = {(cuuint64_t)a->ldw * 2};
    cuuint32_t wb[2] = {64, (cuuint32_t)p.bn};
    if (encode_map(&tmW, a->Wt, 2, wd, ws, wb, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
  }
  const int n_tiles = cdiv(a->N, p.bn);
  CENET_REQUIRE(n_tiles <= 65535, "cenet_gemm_tc: too many N tiles");
  // ---- split-K: few output tiles, long contraction, plain bias epilogue, workspace offered by the caller ----
  p.splits = 1; p.kb_per_split = p.num_kb; p.split_ws = nullptr;
  {
    const long long ctas = (long long)p.num_m_tiles * n_tiles;
    const bool plain_epi = a->c_dtype == CENET_BF16 && a->alpha == 1.0f && !a->row_scale && !a->post_row_scale && !a->bias_per_row &&
                           a->act == CENET_ACT_NONE && !a->mul && !a->res1 && !a->res2 && a->N % 8 == 0 && a->ldc % 8 == 0 &&
                           (((uintptr_t)a->C & 15) == 0) && (!a->bias || (((uintptr_t)a->bias & 15) == 0));
    // (plain GEMMs and the strided convs: few output tiles -- 7x7 tokens per image for the SR convs -- and K = k*k*Cin up to 4096)
    if ((!a->conv || a->stride > 1) && a->split_ws && (((uintptr_t)a->split_ws & 15) == 0) && plain_epi && ctas * 2 <= kNumSMs && p.num_kb >= 8) {
      int want = (int)std::min<long long>(kNumSMs / ctas, p.num_kb / 4);            // >= 4 k-blocks (256 columns) per slice
      const long long room = a->split_ws_elems / ((long long)a->M * a->N);
      if (want > room) want = (int)room;
      if (want >= 2) {
        p.kb_per_split = cdiv(p.num_kb, want);
        p.splits = cdiv(p.num_kb, p.kb_per_split);
        p.split_ws = a->split_ws;
      }
    }
  }
  const int kb_cta = p.splits > 1 ? p.kb_per_split : p.num_kb;                      // k-blocks one CTA walks
  // ---- shared-memory plan (one persistent CTA per SM, <= ~200 KB) ----
  const int a_bytes = p.halo ? (((p.Cin / 8) * p.ablk + 1023) & ~1023) : BM * p.bk * 2, w_bytes = p.bn * p.bk * 2;
  const int wblk = (w_bytes + 1023) & ~1023;
  p.acc_stride = (p.bn + 31) & ~31;
  // narrow-N problems are epilogue/issue-bound per CTA: two CTAs per SM (TMEM 2 x 256 columns, ~100 KB smem each);
  // wide tiles get the whole SM and eight epilogue warps
  int ctas_per_sm = (2 * p.acc_stride <= 256) ? 2 : 1;
  if (p.halo && (long long)kb_cta * wblk + 2 * a_bytes > 88 * 1024) ctas_per_sm = 1;   // filter bank + 2 halo tiles must fit
  p.n_epi = (ctas_per_sm == 1) ? 8 : 4;
  const int slab_bytes = p.n_epi * 32 * 36 * 4;
  const int budget = (ctas_per_sm == 1 ? 200 : 100) * 1024 - slab_bytes - 1024 - 256;
  p.w_resident = p.halo || ((long long)kb_cta * wblk <= (ctas_per_sm == 1 ? 96 : 56) * 1024 && p.num_m_tiles > 1);
  const int wres = p.w_resident ? kb_cta * wblk : 0;
  const int stage_bytes = a_bytes + (p.w_resident ? 0 : wblk);
  int stages = (budget - wres) / stage_bytes;
  if (stages > (p.halo ? 4 : 8)) stages = p.halo ? 4 : 8;
  if (stages < 1) stages = 1;                              // the ring runs ahead across tiles: 8 x 16 KB in flight per SM
  p.stages = stages;
  int cols = 32;
  while (cols < 2 * p.acc_stride) cols <<= 1;
  p.tmem_cols = cols;
  auto vec_ok = [](const void* ptr, int dtype, long long ld) {
    return ptr == nullptr || (dtype == CENET_BF16 && ld % 8 == 0 && ((uintptr_t)ptr & 15) == 0);
  };
  const bool c_ok = (((uintptr_t)a->C & 15) == 0) && ((a->c_dtype == CENET_BF16 && a->ldc % 8 == 0) ||
                                                      (a->c_dtype == CENET_F32 && a->ldc % 4 == 0));
  p.epi_vec = c_ok && vec_ok(a->res1, a->res1_dtype, a->ldr1) && vec_ok(a->res2, a->res2_dtype, a->ldr2) &&
              vec_ok(a->mul, a->mul_dtype, a->ldmul);
  p.epi_fast = 0;
  if (p.epi_vec && a->c_dtype == CENET_BF16 && a->alpha == 1.0f && !a->row_scale && !a->post_row_scale && !a->bias_per_row && a->act == CENET_ACT_NONE &&
      !a->mul && !a->res2 && a->N % 8 == 0 && (!a->bias || (((uintptr_t)a->bias & 15) == 0)) &&
      (!a->res1 || (!a->res1_cscale && a->res1_scale == 1.0f)))
    p.epi_fast = a->res1 ? 2 : 1;
  if (c_ok && a->c_dtype == CENET_F32 && a->res1 && a->res1_dtype == CENET_F32 && a->ldr1 % 4 == 0 && (((uintptr_t)a->res1 & 15) == 0) &&
      !a->res1_cscale && a->res1_scale == 1.0f && a->alpha == 1.0f && !a->row_scale && !a->post_row_scale && !a->bias_per_row &&
      a->act == CENET_ACT_NONE && !a->mul && !a->res2 && a->N % 8 == 0 && (!a->bias || (((uintptr_t)a->bias & 15) == 0)))
    p.epi_fast = 3;                                          // fp32 residual stream of the encoder
  const size_t smem = (size_t)wres + (size_t)stages * stage_bytes + (2 * stages + 5) * 8 + 32 + slab_bytes + 1024;
  int gx = kNumSMs * ctas_per_sm / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > p.num_m_tiles) gx = p.num_m_tiles;
  dim3 grid(gx, n_tiles, p.splits);
  const bool vec16 = p.epi_vec && a->N % 8 == 0 && (!a->bias || a->bias_per_row || (((uintptr_t)a->bias & 15) == 0)) &&
                     (!a->res1_cscale || (((uintptr_t)a->res1_cscale & 15) == 0));
  const int nthr = 64 + 32 * p.n_epi;
  const int vflags = (a->mul ? VF_MUL : 0) | (a->res1 ? VF_R1 : 0) | (a->res2 ? VF_R2 : 0) |
                     ((a->row_scale || a->post_row_scale || (a->bias && a->bias_per_row)) ? VF_ROW : 0);
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const TcParams);
  static const KernelFn vec_kernels[16] = {
      gemm_tc_kernel<EPI_VEC, 0>,  gemm_tc_kernel<EPI_VEC, 1>,  gemm_tc_kernel<EPI_VEC, 2>,  gemm_tc_kernel<EPI_VEC, 3>,
      gemm_tc_kernel<EPI_VEC, 4>,  gemm_tc_kernel<EPI_VEC, 5>,  gemm_tc_kernel<EPI_VEC, 6>,  gemm_tc_kernel<EPI_VEC, 7>,
      gemm_tc_kernel<EPI_VEC, 8>,  gemm_tc_kernel<EPI_VEC, 9>,  gemm_tc_kernel<EPI_VEC, 10>, gemm_tc_kernel<EPI_VEC, 11>,
      gemm_tc_kernel<EPI_VEC, 12>, gemm_tc_kernel<EPI_VEC, 13>, gemm_tc_kernel<EPI_VEC, 14>, gemm_tc_kernel<EPI_VEC, 15>};
  // (the batched-load vector block also beats the older dedicated paths for bias + bf16 / fp32 residual: 60 vs 112 us on the
  // M=75264, N=512, K=64 GEMM; only the plain C = acc + bias store keeps its own minimal kernel)
  KernelFn kern = p.splits > 1 ? gemm_tc_kernel<EPI_SPLITK, 0>
                  : p.epi_fast == 3 ? gemm_tc_kernel<EPI_VEC, VF_R1F32>
                  : p.epi_fast == 1 ? gemm_tc_kernel<EPI_FAST, 0>
                  : vec16 ? vec_kernels[vflags] : gemm_tc_kernel<EPI_SCALAR, 0>;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(gemm_tc_kernel<EPI_SPLITK, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(gemm_tc_kernel<EPI_FAST, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(gemm_tc_kernel<EPI_VEC, VF_R1F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(gemm_tc_kernel<EPI_SCALAR, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int i = 0; i < 16; i++) cudaFuncSetAttribute(vec_kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  });
  kern<<<grid, nthr, smem, s>>>(tmA, tmW, p);
  CENET_LAUNCH_CHECK("gemm_tc");
  if (p.splits > 1) {
    const long long groups = (long long)a->M * (a->N / 8);
    splitk_reduce_kernel<<<(unsigned)cdiv(groups, 256), 256, 0, s>>>(p.split_ws, p.splits, a->M, a->N, a->bias, (bf16*)a->C, a->ldc);
    CENET_LAUNCH_CHECK("splitk_reduce");
  }
  return 0;
}
