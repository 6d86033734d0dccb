// tcgen05 weight-gradient GEMM for sm_100a:   P_z[n, k] = sum_{m in split z} dY[m, n] * X[m, k]      (bf16 in, fp32 out)
//
// Both operands are row-major over the contraction index m, i.e. "MN-major" for the 5th-gen tensor core: a TMA box
// {64 columns, 64 rows} with SWIZZLE_128B lands in shared memory exactly as the canonical MN-major SW128 layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) [16-byte units] with SBO = 1024 B (8-row groups) and LBO = 8192 B (next 64-column box), so
// no transpose pass is needed.  One CTA = one 128 (n) x BN (k) accumulator tile in TMEM for one split of the contraction:
//   warp 0     : TMA producer, 4-stage ring of [64 rows] x (128 + BN) columns
//   warp 1     : TMEM alloc + single-thread tcgen05.mma (M=128, N=BN, K=16; a_major = b_major = MN), commit frees ring slots
//   warps 2..5 : epilogue, tcgen05.ld -> (per-sample DropPath scale) -> fp32 partial tile in the workspace
// The splits are reduced in fixed order by wgrad_finalize (train_gemm.cu) -> deterministic.
// Rows are addressed through a 3-D tensor map {columns, rows per group, groups}: a group is one sample when a per-sample row
// scale applies (the scale then multiplies the partial in the epilogue) and the tail rows of a group are zero-filled by TMA.
#include "train_common.cuh"
#include <cuda.h>
#include <mutex>
#include <cstdio>
#include <cstdlib>

namespace {
constexpr int WT_THREADS = 192;
constexpr int CH = 64;            // contraction rows per pipeline stage
constexpr int BOX_BYTES = CH * 128;

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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
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

// MN-major SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp field layout):
//   [0,14) start>>4 | [16,30) LBO>>4 = stride between 64-element MN blocks | [32,46) SBO>>4 = stride between 8-row K groups |
//   [46,48) version = 1 | [61,64) layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// MN-major descriptor with an explicit swizzle mode: layout type 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t make_desc_mn_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
}

// ---- dense-conv weight gradient -----------------------------------------------------------------------------------------
// dW[(tap, ci), n] = sum over pixels of X[pixel + tap offset, ci] * dY[pixel, n], accumulated for ALL taps in TMEM:
// UMMA M = 128 rows = (128 / Cin) taps x Cin channels (one shifted 4-D TMA box per tap, OOB = zero padding), N = Cout,
// K = the 128 pixels of an 8 x 16 patch.  A persistent CTA streams patches; its whole [T*Cin, Cout] partial lives in TMEM
// (T*Cin/128 accumulator tiles) and is written once at the end.
constexpr int CW_TH = 8, CW_TW = 16;
struct CwParams {
  int Bimg, H, W, Cin, N, KS, pad, T;
  int taps_per_tile, mtiles;
  int tiles_h, tiles_w, npatches;
  int stages, tmem_cols;
  uint32_t xbox_bytes, ybox_bytes;
  float* ws;              // [gridDim.x][N][T*Cin]
};

__global__ void __launch_bounds__(WT_THREADS, 1) conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY,
                                                                     const __grid_constant__ CUtensorMap tmX, const CwParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t stage_bytes = p.ybox_bytes + p.taps_per_tile * p.xbox_bytes;
  const uint32_t bar_base = sbase + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + s * 8; };
  auto empty_bar = [&](int s) { return bar_base + (p.stages + s) * 8; };
  const uint32_t tfull_bar = bar_base + 2 * p.stages * 8;
  const uint32_t tmem_slot = tfull_bar + 8;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  int my_patches = 0;
  for (int pt = blockIdx.x; pt < p.npatches; pt += gridDim.x) my_patches++;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;                               // ring position kept incrementally: no integer division on the issue path
      for (int pt = blockIdx.x; pt < p.npatches; pt += gridDim.x) {
        int t = pt;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h;
        const int img = t / p.tiles_h;
        const int h0 = th * CW_TH, w0 = tw * CW_TW;
        int tkh = 0, tkw = 0;                          // (kh, kw) of the next tap
        for (int mt = 0; mt < p.mtiles; mt++) {
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t sa = sbase + s * stage_bytes;
          const int ntaps = min(p.taps_per_tile, p.T - mt * p.taps_per_tile);
          mbar_arrive_expect_tx(full_bar(s), p.ybox_bytes + ntaps * p.xbox_bytes);
          tma_load_4d(sa, &tmY, full_bar(s), 0, w0, h0, img);
          for (int j = 0; j < ntaps; j++) {
            tma_load_4d(sa + p.ybox_bytes + j * p.xbox_bytes, &tmX, full_bar(s), 0, w0 + tkw - p.pad, h0 + tkh - p.pad, img);
            if (++tkw == p.KS) { tkw = 0; tkh++; }
          }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.N >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      const uint32_t xrow = p.Cin * 2, yrow = p.N * 2;                 // bytes per pixel row of a box (64 or 128)
      const uint32_t xlt = xrow == 128 ? 2u : 4u, ylt = yrow == 128 ? 2u : 4u;
      int s = 0;
      uint32_t ph = 0;
      int pi = 0;
      const uint64_t xk = (uint64_t)((16 * xrow) >> 4), yk = (uint64_t)((16 * yrow) >> 4);   // descriptor step per 16 pixels
      for (int pt = blockIdx.x; pt < p.npatches; pt += gridDim.x, pi++) {
        for (int mt = 0; mt < p.mtiles; mt++) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = sbase + s * stage_bytes;
          // A = X boxes of this tile's taps: MN blocks (one per tap) LBO apart; B = dY box
          const uint64_t adesc = make_desc_mn_sw(sa + p.ybox_bytes, p.xbox_bytes, 8 * xrow, xlt);
          const uint64_t bdesc = make_desc_mn_sw(sa, p.ybox_bytes, 8 * yrow, ylt);
          const uint32_t d_tmem = tmem_base + mt * p.N;
#pragma unroll
          for (int k = 0; k < (CW_TH * CW_TW) / 16; k++)                // 16 pixels per UMMA: 16 rows of the boxes
            umma_f16(d_tmem, adesc + xk * (uint64_t)k, bdesc + yk * (uint64_t)k, idesc, (pi | k) != 0);
          umma_commit(empty_bar(s));
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int quarter = warp & 3;
    const int Ktot = p.T * p.Cin;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    float* out = p.ws + (size_t)blockIdx.x * p.N * Ktot;
    for (int mt = 0; mt < p.mtiles; mt++) {
      const int krow = mt * 128 + quarter * 32 + lane;                    // (tap, ci) index of this thread's accumulator row
      for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * p.N + c0), acc);
        tmem_ld_wait();
        if (krow < Ktot) {
#pragma unroll
          for (int j = 0; j < 32; j++) out[(size_t)(c0 + j) * Ktot + krow] = my_patches > 0 ? __uint_as_float(acc[j]) : 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

struct WtParams {
  int N, K;               // output rows (dY columns) / columns (X columns)
  int bn;                 // X-column tile: multiple of 64, <= 256
  int stages, tmem_cols;
  int rows_per_group;     // rows of one group (sample); M when there is no row scale
  int groups, cpg;        // groups, 64-row chunks per group
  int per_group;          // 1: a split stays inside one group (arbitrary per-sample scale, applied in the epilogue)
  int parts;              // per_group: splits per group
  int cps;                // chunks per split
  int total_chunks;       // flat mode: groups * cpg
  int binary;             // flat mode with a row scale: rs is {0, c} -- groups with rs == 0 are skipped, c scales the result
  const float* rs;        // per-group scale or NULL
  float* out;             // partial z -> out + z * out_stride, [N][K]   (the final gradient itself when there is one split)
  long long out_stride;
  float* bias;            // column sums of dY (bias gradient): partial z -> bias + z * bias_stride, [N]; NULL: not wanted
  long long bias_stride;
};

// One CTA = one 128 (n) x bn (k) tile of one split.  Flat mode walks the 64-row chunks [z*cps, (z+1)*cps) of the
// (group, chunk) enumeration and -- with a binary DropPath mask -- skips the chunks of dropped samples, so the number of
// splits no longer grows with the batch.  The bias gradient rides along as one more N=16 MMA per k-step against a constant
// all-ones B tile (columns [bn, bn+16) of the accumulator; only the CTAs of the first k tile do it).
__global__ void __launch_bounds__(WT_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY,
                                                                const __grid_constant__ CUtensorMap tmX, const WtParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xboxes = p.bn / 64;
  const uint32_t stage_bytes = (2 + xboxes) * BOX_BYTES;
  const uint32_t ones_base = sbase + p.stages * stage_bytes;          // 2 KB of bf16 1.0 (16 contraction rows x 64 columns)
  const uint32_t bar_base = ones_base + 2048;
  auto full_bar = [&](int s) { return bar_base + s * 8; };
  auto empty_bar = [&](int s) { return bar_base + (p.stages + s) * 8; };
  const uint32_t tfull_bar = bar_base + 2 * p.stages * 8;
  const uint32_t tmem_slot = tfull_bar + 8;
  const int n0 = blockIdx.x * 128, k0 = blockIdx.y * p.bn;
  const int z = blockIdx.z;
  const bool do_bias = p.bias != nullptr && blockIdx.y == 0;
  // this split's chunk range: [c_lo, c_hi) of group grp0 (per_group) or of the flat (group, chunk) enumeration
  int grp0, cig0, nwalk;
  if (p.per_group) {
    grp0 = z / p.parts;
    cig0 = (z % p.parts) * p.cps;
    nwalk = max(0, min(p.cpg, cig0 + p.cps) - cig0);
  } else {
    const int g0 = z * p.cps;
    grp0 = g0 / p.cpg;
    cig0 = g0 % p.cpg;
    nwalk = max(0, min(p.total_chunks, g0 + p.cps) - g0);
  }
  const bool mask = p.binary && p.rs != nullptr;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (do_bias && warp >= 2) {                      // 128 threads x 16 bytes = the 2 KB ones tile
    const uint32_t a = ones_base + (threadIdx.x - 64) * 16;
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0x3F803F80u) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;                               // ring position kept incrementally: no integer division on the issue path
      int grp = grp0, cig = cig0;
      bool live = !mask || p.rs[grp < p.groups ? grp : 0] != 0.f;
      for (int c = 0; c < nwalk; c++) {
        if (live) {
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t sa = sbase + s * stage_bytes;
          mbar_arrive_expect_tx(full_bar(s), stage_bytes);
          const int r0 = cig * CH;
          tma_load_3d(sa, &tmY, full_bar(s), n0, r0, grp);
          tma_load_3d(sa + BOX_BYTES, &tmY, full_bar(s), n0 + 64, r0, grp);
          for (int j = 0; j < xboxes; j++) tma_load_3d(sa + (2 + j) * BOX_BYTES, &tmX, full_bar(s), k0 + 64 * j, r0, grp);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (++cig == p.cpg) {
          cig = 0; grp++;
          live = !mask || (grp < p.groups && p.rs[grp] != 0.f);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D = f32 (1<<4), A = B = bf16 (1<<7, 1<<10), A and B MN-major (bits 15, 16), N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc = idesc0 | ((uint32_t)(p.bn >> 3) << 17);
      const uint32_t idesc_b = idesc0 | ((uint32_t)(16 >> 3) << 17);
      const uint64_t odesc = make_desc_mn(ones_base, BOX_BYTES, 1024);
      int s = 0;
      uint32_t ph = 0;
      uint32_t started = 0;
      int grp = grp0, cig = cig0;
      bool live = !mask || p.rs[grp < p.groups ? grp : 0] != 0.f;
      for (int c = 0; c < nwalk; c++) {
        if (live) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = sbase + s * stage_bytes;
          const uint64_t adesc = make_desc_mn(sa, BOX_BYTES, 1024), bdesc = make_desc_mn(sa + 2 * BOX_BYTES, BOX_BYTES, 1024);
#pragma unroll
          for (int k = 0; k < CH / 16; k++) {    // 16 contraction rows = two 8-row groups = 2048 bytes = +128 in the address field
            umma_f16(tmem_base, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc, started | k);
            if (do_bias) umma_f16(tmem_base + p.bn, adesc + (uint64_t)(128 * k), odesc, idesc_b, started | k);
          }
          started = 1;
          umma_commit(empty_bar(s));
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (++cig == p.cpg) {
          cig = 0; grp++;
          live = !mask || (grp < p.groups && p.rs[grp] != 0.f);
        }
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int quarter = warp & 3;                // TMEM lane quarter this warp may read (warps 2..5 -> 2,3,0,1)
    const int n = n0 + quarter * 32 + lane;
    // scale of this split and whether any chunk was accumulated at all (an untouched accumulator holds garbage)
    float scale = 1.f;
    bool any = nwalk > 0;
    if (p.rs) {
      if (p.per_group) {
        scale = p.rs[grp0];
      } else if (mask) {
        float c = 0.f;
        for (int g = 0; g < p.groups; g++) c = fmaxf(c, p.rs[g]);     // the common value of the kept samples
        scale = c;
        any = false;
        const int g_last = nwalk > 0 ? (z * p.cps + nwalk - 1) / p.cpg : -1;
        for (int g = grp0; g <= g_last && g < p.groups; g++) any |= p.rs[g] != 0.f;
      }
    }
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    float* out = p.out + (size_t)z * p.out_stride + (size_t)n * p.K;
    for (int c0 = 0; c0 < p.bn; c0 += 32) {
      if (k0 + c0 >= p.K) break;
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, acc);
      tmem_ld_wait();
      if (n < p.N) {
        if (!any) {
#pragma unroll
          for (int j = 0; j < 32; j++) acc[j] = 0u;
        }
        const int kk = k0 + c0;
        if (kk + 32 <= p.K && (p.K & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(out + kk + j) =
                make_float4(__uint_as_float(acc[j]) * scale, __uint_as_float(acc[j + 1]) * scale, __uint_as_float(acc[j + 2]) * scale,
                            __uint_as_float(acc[j + 3]) * scale);
        } else {
          for (int j = 0; j < 32 && kk + j < p.K; j++) out[kk + j] = __uint_as_float(acc[j]) * scale;
        }
      }
    }
    if (do_bias) {
      uint32_t acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)p.bn, acc);
      tmem_ld_wait();
      if (n < p.N) p.bias[(size_t)z * p.bias_stride + n] = any ? __uint_as_float(acc[0]) * scale : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

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

int encode_3d(CUtensorMap* tm, const void* base, long long cols, long long ld, long long rows_per_group, long long groups) {
  EncodeTiledFn enc = get_encode();
  CENET_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows_per_group, (cuuint64_t)groups};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows_per_group * ld * 2};
  cuuint32_t box[3] = {64, CH, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CENET_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (wgrad) failed with CUresult %d", (int)r);
  return 0;
}

int encode_4d(CUtensorMap* tm, const void* base, int C, int W, int H, int B, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  CENET_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, CW_TW, CW_TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CENET_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (conv wgrad) failed with CUresult %d", (int)r);
  return 0;
}
}  // namespace

bool cenet_conv_wgrad_tc_eligible(const void* dy, int dy_dtype, long long ldy, const void* x, int x_dtype, long long ldx, int Cin, int N,
                                  int ksize) {
  if (dy_dtype != CENET_BF16 || x_dtype != CENET_BF16) return false;
  if (!(Cin == 32 || Cin == 64) || !(N == 32 || N == 64)) return false;
  if (ldy != N || ldx != Cin || ((uintptr_t)dy & 15) || ((uintptr_t)x & 15)) return false;
  const int T = ksize * ksize, tpt = 128 / Cin, mtiles = (T + tpt - 1) / tpt;
  return mtiles * N <= 512;
}

// writes gridDim.x partials [N][T*Cin] into ws and returns their number (or -1 / -2)
int cenet_conv_wgrad_tc(const void* dy, const void* x, int B, int H, int W, int Cin, int ksize, int N, float* ws, long long ws_elems,
                        cudaStream_t s) {
  CwParams p = {};
  p.Bimg = B; p.H = H; p.W = W; p.Cin = Cin; p.N = N; p.KS = ksize; p.pad = ksize / 2; p.T = ksize * ksize;
  p.taps_per_tile = 128 / Cin;
  p.mtiles = (p.T + p.taps_per_tile - 1) / p.taps_per_tile;
  p.tiles_h = cdiv(H, CW_TH); p.tiles_w = cdiv(W, CW_TW);
  p.npatches = B * p.tiles_h * p.tiles_w;
  p.xbox_bytes = 128 * Cin * 2; p.ybox_bytes = 128 * N * 2;
  p.ws = ws;
  int cols = 32;
  while (cols < p.mtiles * N) cols <<= 1;
  p.tmem_cols = cols;
  const int stage_bytes = p.ybox_bytes + p.taps_per_tile * p.xbox_bytes;
  p.stages = std::min(6, (200 * 1024) / stage_bytes);
  int grid = std::min(kNumSMs, p.npatches);
  const long long nk = (long long)N * p.T * Cin;
  if ((long long)grid * nk > ws_elems) grid = (int)(ws_elems / nk);
  if (grid < 1) return -2;
  CUtensorMap tmY, tmX;
  if (encode_4d(&tmY, dy, N, W, H, B, N == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) return -1;
  if (encode_4d(&tmX, x, Cin, W, H, B, Cin == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) return -1;
  const size_t smem = (size_t)p.stages * stage_bytes + (2 * p.stages + 2) * 8 + 16 + 1024;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] { cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); });
  conv_wgrad_tc_kernel<<<grid, WT_THREADS, smem, s>>>(tmY, tmX, p);
  CENET_LAUNCH_CHECK("conv_wgrad_tc");
  return grid;
}

// eligibility: bf16 operands, 16-byte aligned bases and pitches, per-sample (or no) row scale
bool cenet_wgrad_tc_eligible(const void* dy, int dy_dtype, long long ldy, const void* x, int x_dtype, long long ldx, long long M, int N,
                             int K, const float* rs, int rs_div) {
  if (dy_dtype != CENET_BF16 || x_dtype != CENET_BF16) return false;
  if (ldy % 8 || ldx % 8 || ((uintptr_t)dy & 15) || ((uintptr_t)x & 15)) return false;
  if (M < 512 || N < 32 || K < 8) return false;
  if (rs && (rs_div < 16 || M % rs_div != 0)) return false;      // a group shorter than the 64-row TMA box is zero-filled (7x7 level: 49)
  return true;
}

// Tile width and split count.  A CTA's time is modelled as pipeline fill + chunks x max(MMA, shared-memory fill) + epilogue,
// the reduction of S partials as their write + read; the (bn, S) pair with the smallest estimate wins.  Few, long splits
// for large outputs over short contractions (stage 3/4: N*K up to 2 M, M = 1-5 k rows), many splits of a narrow tile for the
// small outputs over long contractions (stage 1, head: N*K = 4-16 k, M = 75 k-1.2 M rows).
static inline int wt_stages(int bn) { return bn <= 64 ? 4 : (bn <= 128 ? 5 : 4); }   // 96 / 160 / 160 / 192 KB rings
void cenet_wgrad_tc_plan(long long M, int N, int K, bool has_rs, int rs_div, bool binary, long long max_partials, cenet_wgrad_plan* pl) {
  const bool per_group = has_rs && !binary;
  pl->groups = (int)(has_rs ? M / rs_div : 1);
  pl->rows_per_group = (int)(has_rs ? rs_div : M);
  pl->cpg = cdiv(pl->rows_per_group, CH);
  pl->total_chunks = pl->groups * pl->cpg;
  pl->per_group = per_group;
  const long long nk = (long long)N * K;
  const int walk = per_group ? pl->cpg : pl->total_chunks;          // chunks one "unit" (group / whole problem) offers for splitting
  const int units = per_group ? pl->groups : 1;
  double best = 1e30;
  pl->bn = 64; pl->parts = 1; pl->cps = walk;
  const int kr = (K + 63) / 64 * 64;
  // Calibrated on the B200 (tools/sweep_wgrad.py): the best plans put about one CTA on every SM; their time is the bytes all
  // CTAs pull through L2 (~5 TB/s for these boxes: every dY tile is re-read once per k tile, every X tile once per n tile)
  // unless a CTA's own chain of chunks (tensor pipe / fill bandwidth / TMA round trip over the ring depth) is longer.
  for (int bn = 64; bn <= 256 && bn <= std::max(64, kr); bn += 64) {
    const int tiles = cdiv(N, 128) * cdiv(K, bn);
    const int occ = bn == 64 ? 2 : 1;
    const double eb = (std::min(128, N) + std::min(bn, K)) * 128.0;            // bytes a chunk really fetches (OOB box parts are free)
    const double t_mma = 128.0 * bn * 64 * 2 / 9.0e6, t_fill = eb / 0.12e6;
    const double t_lat = 2.4 / wt_stages(bn);                                  // loaded TMA round trip hidden by the ring depth
    const double t_epi = 128.0 * std::min(bn, kr) * 4 / 0.12e6;
    const double t_mem = (double)tiles * units * (double)walk * eb / 5.0e6;
    const int target = std::max(1, (int)((kNumSMs + tiles * units / 2) / ((long long)tiles * units)));
    const int cand[4] = {target, std::max(1, target / 2), target * 2, 1};
    for (int ci = 0; ci < 4; ci++) {
      const int parts = std::min(cand[ci], std::max(1, walk / 2));
      const int cps = cdiv(walk, parts);
      const int rparts = cdiv(walk, cps);
      const long long S = (long long)units * rparts;
      if (S > max_partials || S > 65535) continue;
      const long long ctas = S * tiles;
      const long long waves = (ctas + (long long)kNumSMs * occ - 1) / ((long long)kNumSMs * occ);
      const double share = ctas > kNumSMs ? std::min((double)occ, (double)ctas / kNumSMs) : 1.0;   // co-resident CTAs share an SM
      const double t_cta = 2.5 + cps * std::max(share * std::max(t_mma, t_fill), t_lat) + share * t_epi;
      double t = std::max(waves * t_cta, t_mem);
      if (S > 1) t += (double)S * nk * 8 / 10.0e6;                 // write + batched read of the partials
      if (t < best) { best = t; pl->bn = bn; pl->parts = rparts; pl->cps = cps; }
    }
  }
  if (const char* f = getenv("CENET_B200_WGRAD_PLAN")) {            // tuning override: "bn,parts"
    int fbn = 0, fparts = 0;
    if (sscanf(f, "%d,%d", &fbn, &fparts) == 2 && fbn >= 64 && fbn <= 256 && fbn % 64 == 0 && fparts >= 1) {
      fparts = std::min(fparts, walk);
      pl->bn = std::min(fbn, std::max(64, kr));
      pl->cps = cdiv(walk, fparts);
      pl->parts = cdiv(walk, pl->cps);
    }
  }
  pl->S = units * pl->parts;
}

// host-only introspection of the plan (tools/one_wgrad.py, tests): out = {bn, S, per_group, parts, cps, cpg, groups, rows_per_group,
// total_chunks}
extern "C" int cenet_wgrad_plan_query(long long M, int N, int K, int has_rs, int rs_div, int binary, long long max_partials, int* out) {
  cenet_wgrad_plan pl = {};
  cenet_wgrad_tc_plan(M, N, K, has_rs != 0, rs_div, binary != 0, max_partials, &pl);
  const int v[9] = {pl.bn, pl.S, pl.per_group, pl.parts, pl.cps, pl.cpg, pl.groups, pl.rows_per_group, pl.total_chunks};
  for (int i = 0; i < 9; i++) out[i] = v[i];
  return 0;
}

int cenet_wgrad_tc_launch(const cenet_wgrad_plan* pl, const void* dy, long long ldy, const void* x, long long ldx, int N, int K,
                          const float* rs, bool binary, float* out, long long out_stride, float* bias, long long bias_stride,
                          cudaStream_t s) {
  WtParams p = {};
  p.N = N; p.K = K; p.rs = rs; p.bn = pl->bn;
  p.rows_per_group = pl->rows_per_group; p.groups = pl->groups; p.cpg = pl->cpg; p.per_group = pl->per_group;
  p.parts = pl->parts; p.cps = pl->cps; p.total_chunks = pl->total_chunks; p.binary = binary ? 1 : 0;
  p.out = out; p.out_stride = out_stride; p.bias = bias; p.bias_stride = bias_stride;
  CUtensorMap tmY, tmX;
  if (encode_3d(&tmY, dy, N, ldy, p.rows_per_group, p.groups)) return -1;
  if (encode_3d(&tmX, x, K, ldx, p.rows_per_group, p.groups)) return -1;
  const int stage_bytes = (2 + p.bn / 64) * BOX_BYTES;
  p.stages = wt_stages(p.bn);
  int cols = 32;
  while (cols < p.bn + (bias ? 16 : 0)) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = (size_t)p.stages * stage_bytes + 2048 + (2 * p.stages + 2) * 8 + 16 + 1024;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] { cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); });
  dim3 grid(cdiv(N, 128), cdiv(K, p.bn), (unsigned)pl->S);
  wgrad_tc_kernel<<<grid, WT_THREADS, smem, s>>>(tmY, tmX, p);
  CENET_LAUNCH_CHECK("wgrad_tc");
  return 0;
}
