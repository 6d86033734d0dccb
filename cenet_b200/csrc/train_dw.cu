// Backward helpers of the depthwise / strided-conv family (HBM-bound):
//   dwconv3x3_wgrad : filter + bias gradient of every depthwise 3x3 (pvtv2.py:364-370, cfam.py:151-152, blocks.py:169-178,
//                     blocks.py:303-311) -- 10 per-channel reductions over all pixels, two-stage deterministic.
//   (the data gradient of a depthwise conv is the forward kernel with flipped taps, see train.py::dwconv)
//   sumpool2        : adjoint of the nearest x2 up-sampling of EUCB (blocks.py:304)
//   col2im          : adjoint of cenet_im2col (patch-embed / SR-conv data gradients, pvtv2.py:164-165, 68)
#include "train_common.cuh"
#include <atomic>
#include <cstdlib>

namespace {
// V consecutive elements as one raw vector register (kept packed until use)
template <typename T, int V>
struct alignas(sizeof(T) * V) RawVec { T v[V]; };
template <typename T, int V>
__device__ __forceinline__ RawVec<T, V> ld_raw(const T* p) { return *reinterpret_cast<const RawVec<T, V>*>(p); }
template <typename T, int V>
__device__ __forceinline__ void unpack_raw(const RawVec<T, V>& r, float (&o)[V]) {
#pragma unroll
  for (int i = 0; i < V; i++) o[i] = (float)r.v[i];
}

// Stage 1 of the filter gradient  dw[c, kh, kw] = sum_{b,h,w} dz[b,h,w,c] * x[b, h+(kh-1)d, w+(kw-1)d, c]  (+ dbias = sum dz).
// Work unit = one IMAGE ROW (b, h): a thread owns V channels and walks the row left to right, so the pixel index math
// (one 32-bit division) is paid per row, not per pixel.  For the common case (dilation 1, no fused up-sampling) the 3x3
// neighbourhood slides through registers: per pixel a thread issues 4 vector loads (dz and one new x column of three
// rows) instead of 10, and every x element is fetched three times overall (by the rows above / at / below), which L1/L2
// absorb -- HBM sees x and dz once.  A block is `ngrp` channel groups x `nrl` row lanes; adjacent lanes take adjacent rows.
template <typename T, int V, bool WINDOW>
__global__ void __launch_bounds__(kColThreads, 2) dw_wgrad_partial_kernel(const T* __restrict__ x, long long ldx, const T* __restrict__ dz,
                                                                       long long ldz, int B, int H, int W, int C, int dil, int up2,
                                                                       int ngrp, int nrl, int rows_per_block, int wsegs,
                                                                       float* __restrict__ ws) {
  __shared__ float smem[V * kColThreads];
  const int grp = threadIdx.x % ngrp, rl = threadIdx.x / ngrp;
  const int c0 = (blockIdx.y * ngrp + grp) * V;
  // generic branch only: a work unit is one of `wsegs` column segments of an image row, so that narrow channel slices at a
  // small batch (dec1: 20 channels, 24 x 56 rows = 42 blocks) still put a block on every SM
  const int R = B * H * wsegs;
  const int wseg = (W + wsegs - 1) / wsegs;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  const int Hi = up2 ? H >> 1 : H, Wi = up2 ? W >> 1 : W;
  float acc[10][V];
#pragma unroll
  for (int k = 0; k < 10; k++)
#pragma unroll
    for (int v = 0; v < V; v++) acc[k][v] = 0.f;
  if (c0 < C) {
    for (int rr = r0 + rl; rr < r1; rr += nrl) {
      const int r = WINDOW ? rr : rr / wsegs;
      const int w_lo = WINDOW ? 0 : (rr - r * wsegs) * wseg, w_hi = WINDOW ? W : min(W, w_lo + wseg);
      const int b = r / H, h = r - b * H;
      const T* zrow = dz + (long long)r * W * ldz + c0;
      if constexpr (WINDOW) {
        // Software pipeline: the loads of step t (dz[t] and the x column t+1 of the three rows) are issued DEPTH steps
        // ahead into a ring of raw registers, unconditionally (clamped addresses, zeroed on use), so that a warp keeps
        // DEPTH * 4 vector loads in flight instead of waiting one memory latency per pixel.
        constexpr int DEPTH = 2;
        static_assert(6 % DEPTH == 0, "ring depth must divide the unroll factor");
        const T* xr[3];
        bool ok[3];
#pragma unroll
        for (int dh = 0; dh < 3; dh++) {
          const int hh = h + dh - 1;
          ok[dh] = hh >= 0 && hh < H;
          xr[dh] = x + (long long)(b * H + (ok[dh] ? hh : h)) * W * ldx + c0;
        }
        RawVec<T, V> rz[DEPTH], rx[DEPTH][3];
#pragma unroll
        for (int t = 0; t < DEPTH; t++) {
          const int tz = min(t, W - 1), tx = min(t + 1, W - 1);
          rz[t] = ld_raw<T, V>(zrow + (long long)tz * ldz);
#pragma unroll
          for (int dh = 0; dh < 3; dh++) rx[t][dh] = ld_raw<T, V>(xr[dh] + (long long)tx * ldx);
        }
        float col[3][3][V];                       // [slot][dh][v]: slots rotate over the columns w-1, w, w+1
#pragma unroll
        for (int dh = 0; dh < 3; dh++) {
          unpack_raw<T, V>(ld_raw<T, V>(xr[dh]), col[0][dh]);
#pragma unroll
          for (int v = 0; v < V; v++) { col[2][dh][v] = 0.f; col[0][dh][v] = ok[dh] ? col[0][dh][v] : 0.f; }
        }
        for (int wb = 0; wb < W; wb += 6) {
#pragma unroll
          for (int u = 0; u < 6; u++) {
            const int w = wb + u;
            if (w < W) {
              const int sl = (u + 2) % 3, sc = u % 3, sr = (u + 1) % 3;       // window slots of columns w-1, w, w+1
              const int rs = u % DEPTH;                                        // ring slot of step w (DEPTH divides 6)
              float g[V];
              unpack_raw<T, V>(rz[rs], g);
              const bool last = w + 1 >= W;
#pragma unroll
              for (int dh = 0; dh < 3; dh++) {
                unpack_raw<T, V>(rx[rs][dh], col[sr][dh]);
#pragma unroll
                for (int v = 0; v < V; v++) col[sr][dh][v] = (ok[dh] && !last) ? col[sr][dh][v] : 0.f;
              }
              {                                                                // refill the slot with step w + DEPTH
                const int tz = min(w + DEPTH, W - 1), tx = min(w + DEPTH + 1, W - 1);
                rz[rs] = ld_raw<T, V>(zrow + (long long)tz * ldz);
#pragma unroll
                for (int dh = 0; dh < 3; dh++) rx[rs][dh] = ld_raw<T, V>(xr[dh] + (long long)tx * ldx);
              }
#pragma unroll
              for (int v = 0; v < V; v++) acc[9][v] += g[v];
#pragma unroll
              for (int dh = 0; dh < 3; dh++)
#pragma unroll
                for (int v = 0; v < V; v++) {
                  acc[dh * 3 + 0][v] = fmaf(g[v], col[sl][dh][v], acc[dh * 3 + 0][v]);
                  acc[dh * 3 + 1][v] = fmaf(g[v], col[sc][dh][v], acc[dh * 3 + 1][v]);
                  acc[dh * 3 + 2][v] = fmaf(g[v], col[sr][dh][v], acc[dh * 3 + 2][v]);
                }
            }
          }
        }
      } else {
        const T* xb = x + (long long)b * Hi * Wi * ldx + c0;
        for (int w = w_lo; w < w_hi; w++) {
          float g[V];
          ldv<V>(zrow + (long long)w * ldz, g);
#pragma unroll
          for (int v = 0; v < V; v++) acc[9][v] += g[v];
#pragma unroll
          for (int dh = -1; dh <= 1; dh++) {
            const int hh = h + dh * dil;
            if (hh < 0 || hh >= H) continue;
            const int hs = up2 ? hh >> 1 : hh;
#pragma unroll
            for (int dw = -1; dw <= 1; dw++) {
              const int ww = w + dw * dil;
              if (ww < 0 || ww >= W) continue;
              const int wsrc = up2 ? ww >> 1 : ww;
              float xv[V];
              ldv<V>(xb + ((long long)hs * Wi + wsrc) * ldx, xv);
#pragma unroll
              for (int v = 0; v < V; v++) acc[(dh + 1) * 3 + dw + 1][v] = fmaf(g[v], xv[v], acc[(dh + 1) * 3 + dw + 1][v]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 10; k++) {
    col_block_reduce<V>(acc[k], smem, grp, rl, ngrp, nrl);
    if (rl == 0 && c0 < C) {
#pragma unroll
      for (int v = 0; v < V; v++)
        if (c0 + v < C) ws[((size_t)blockIdx.x * 10 + k) * C + c0 + v] = acc[k][v];
    }
  }
}

// Shared-memory staged variant of stage 1 (bf16, dilation 1, no fused up-sampling, C % 64 == 0): same structure as
// dwconv3x3_staged_kernel (dwconv.cu).  A CTA owns 64 channels of a band of rows of one image and streams x rows (ring of
// 5, zero halo columns) and dz rows (ring of 3) through shared memory with cp.async, two rows ahead of the arithmetic.
// Thread = 4 channels x 4 consecutive pixels with its 10 x 4 accumulators in registers for the whole band; one shared
// memory reduction over the 16 pixel groups at the end; partial per (image, band): ws[((b * nbands + band) * 10 + k) * C + c].
constexpr int WS_C = 64, WS_T = 256, WS_PF = 2, WS_NRX = 3 + WS_PF, WS_NRZ = 1 + WS_PF;

__global__ void __launch_bounds__(WS_T, 2) dw_wgrad_staged_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ dz, int ldz,
                                                                  int H, int W, int C, int rows_per_cta, float* __restrict__ ws) {
  extern __shared__ __align__(16) unsigned char wsm[];
  const int tid = threadIdx.x;
  const int rowX = (W + 2) * 128, rowZ = W * 128;
  unsigned char* zring = wsm + WS_NRX * rowX;
  const int c_base = blockIdx.x * WS_C;
  const int h0 = blockIdx.y * rows_per_cta, h1 = min(H, h0 + rows_per_cta);
  const int b = blockIdx.z;
  const bf16* xb = x + (size_t)b * H * W * ldx + c_base;
  const bf16* zb = dz + (size_t)b * H * W * ldz + c_base;
  for (int i = tid; i < WS_NRX * 16; i += WS_T) {
    const int slot = i >> 4, side = (i >> 3) & 1, ch = i & 7;
    *reinterpret_cast<uint4*>(wsm + slot * rowX + (side ? (W + 1) * 128 : 0) + ch * 16) = make_uint4(0, 0, 0, 0);
  }
  // group k = { x row h0-1+k -> x slot k % NRX,  dz row h0-2+k -> z slot (k-2) % NRZ }
  auto issue = [&](int k) {
    const int rx = h0 - 1 + k, rz = h0 - 2 + k;
    if (rx <= h1) {
      const bool valid = rx >= 0 && rx < H;
      const bf16* src = xb + (size_t)(valid ? rx : 0) * W * ldx;
      unsigned char* dst = wsm + (k % WS_NRX) * rowX + 128;
      for (int i = tid; i < W * 8; i += WS_T) {
        const int w = i >> 3, ch = i & 7;
        cp_async16_zfill(dst + w * 128 + ch * 16, src + (size_t)w * ldx + ch * 8, valid ? 16 : 0);
      }
    }
    if (rz >= h0 && rz < h1) {
      const bf16* src = zb + (size_t)rz * W * ldz;
      unsigned char* dst = zring + ((k - 2) % WS_NRZ) * rowZ;
      for (int i = tid; i < W * 8; i += WS_T) {
        const int w = i >> 3, ch = i & 7;
        cp_async16_zfill(dst + w * 128 + ch * 16, src + (size_t)w * ldz + ch * 8, 16);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int k = 0; k < WS_PF + 2; k++) issue(k);
  const int cv = tid & 15, pg0 = tid >> 4;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(wsm);
  const unsigned zbase = sbase + WS_NRX * rowX;
  f32x2 acc[10][2];                                            // packed channel pairs (FFMA2)
#pragma unroll
  for (int k = 0; k < 10; k++) acc[k][0] = acc[k][1] = pk2(0.f, 0.f);
  const f32x2 one2 = pk2(1.f, 1.f);
  const int npg = W / 4;                                       // W % 4 == 0 (dispatch)
  for (int h = h0; h < h1; h++) {
    const int k0 = h - h0;                                     // groups 0 .. k0+2 hold x rows <= h+1 and dz rows <= h
    cp_async_wait<WS_PF - 1>();
    __syncthreads();
    issue(k0 + WS_PF + 2);
    unsigned rows[3];
#pragma unroll
    for (int dh = 0; dh < 3; dh++) rows[dh] = sbase + ((k0 + dh) % WS_NRX) * rowX + cv * 8;
    const unsigned zrow = zbase + (k0 % WS_NRZ) * rowZ + cv * 8;
    for (int pg = pg0; pg < npg; pg += 16) {
      f32x2 g[4][2];
#pragma unroll
      for (int p = 0; p < 4; p++) {
        unsigned w0, w1;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(zrow + pg * 512 + p * 128));
        g[p][0] = bf2_to_f2(w0); g[p][1] = bf2_to_f2(w1);
        acc[9][0] = ffma2(g[p][0], one2, acc[9][0]);
        acc[9][1] = ffma2(g[p][1], one2, acc[9][1]);
      }
#pragma unroll
      for (int dh = 0; dh < 3; dh++) {
        const unsigned ra = rows[dh] + pg * 512;
#pragma unroll
        for (int q = 0; q < 6; q++) {
          unsigned w0, w1;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(ra + q * 128));
          const f32x2 x01 = bf2_to_f2(w0), x23 = bf2_to_f2(w1);
#pragma unroll
          for (int p = 0; p < 4; p++) {
            const int t = q - p;
            if (t < 0 || t > 2) continue;
            acc[dh * 3 + t][0] = ffma2(g[p][0], x01, acc[dh * 3 + t][0]);
            acc[dh * 3 + t][1] = ffma2(g[p][1], x23, acc[dh * 3 + t][1]);
          }
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();                                             // the rings are free: reuse them for the block reduction
  float* red = reinterpret_cast<float*>(wsm);                  // [16 pixel groups][16 channel quads * 40]
#pragma unroll
  for (int k = 0; k < 10; k++) {
    float a0, a1, a2, a3;
    upk2(acc[k][0], a0, a1); upk2(acc[k][1], a2, a3);
    *reinterpret_cast<float4*>(red + pg0 * 640 + cv * 40 + k * 4) = make_float4(a0, a1, a2, a3);
  }
  __syncthreads();
  const int blk = blockIdx.z * gridDim.y + blockIdx.y;
  for (int o = tid; o < 640; o += WS_T) {
    float t = 0.f;
#pragma unroll
    for (int p = 0; p < 16; p++) t += red[p * 640 + o];
    const int q = o / 40, k = (o % 40) >> 2, v = o & 3;
    ws[((size_t)blk * 10 + k) * C + c_base + q * 4 + v] = t;
  }
}

__global__ void dw_wgrad_finalize_kernel(const float* __restrict__ ws, int nblk, int C, float* dw, float* dbias) {
  __shared__ float sm[kFinThreads];
  const int i = blockIdx.x * kFinOut + threadIdx.x % kFinOut;
  const int k = i / C, c = i % C;
  float t[1];
  fin_lane_sums<1>(nblk, i < 10 * C, t, sm, [&](int b, int) { return ws[((size_t)b * 10 + k) * C + c]; });
  if (threadIdx.x >= kFinOut || i >= 10 * C) return;
  const float s = t[0];
  if (k < 9) dw[c * 9 + k] = s;
  else if (dbias) dbias[c] = s;
}

template <typename T, int V>
__global__ void __launch_bounds__(256) sumpool2_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int Ho, int Wo, int C, int acc) {
  const int groups = C / V;
  const long long total = (long long)B * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % groups) * V;
    long long p = i / groups;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const int b = (int)(p / Ho);
    float o[V];
#pragma unroll
    for (int v = 0; v < V; v++) o[v] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
      for (int dx = 0; dx < 2; dx++) {
        float t[V];
        ldv<V>(x + (((long long)b * 2 * Ho + 2 * ho + dy) * 2 * Wo + 2 * wo + dx) * C + c0, t);
#pragma unroll
        for (int v = 0; v < V; v++) o[v] += t[v];
      }
    T* yp = y + (((long long)b * Ho + ho) * Wo + wo) * C + c0;
    if (acc) {
      float t[V];
      ldv<V>(yp, t);
#pragma unroll
      for (int v = 0; v < V; v++) o[v] += t[v];
    }
    stv<V>(yp, o);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) col2im_kernel(const T* __restrict__ dcol, T* __restrict__ dx, int B, int H, int W, int Cin, int k,
                                                     int stride, int pad, int Ho, int Wo, int Kp, int acc) {
  const int groups = Cin / V;
  const long long total = (long long)B * H * W * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % groups) * V;
    long long p = i / groups;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int b = (int)(p / H);
    float o[V];
#pragma unroll
    for (int v = 0; v < V; v++) o[v] = 0.f;
    for (int kh = (h + pad) % stride; kh < k; kh += stride) {
      const int ho = (h + pad - kh) / stride;
      if (ho < 0 || ho >= Ho) continue;
      for (int kw = (w + pad) % stride; kw < k; kw += stride) {
        const int wo = (w + pad - kw) / stride;
        if (wo < 0 || wo >= Wo) continue;
        float t[V];
        ldv<V>(dcol + (((long long)b * Ho + ho) * Wo + wo) * Kp + (kh * k + kw) * Cin + c0, t);
#pragma unroll
        for (int v = 0; v < V; v++) o[v] += t[v];
      }
    }
    T* xp = dx + (((long long)b * H + h) * W + w) * Cin + c0;
    if (acc) {
      float t[V];
      ldv<V>(xp, t);
#pragma unroll
      for (int v = 0; v < V; v++) o[v] += t[v];
    }
    stv<V>(xp, o);
  }
}
}  // namespace

#define DISPATCH_V(V_, ...)                                  \
  do {                                                       \
    if (V_ == 8) { constexpr int V = 8; __VA_ARGS__; }       \
    else if (V_ == 4) { constexpr int V = 4; __VA_ARGS__; }  \
    else if (V_ == 2) { constexpr int V = 2; __VA_ARGS__; }  \
    else { constexpr int V = 1; __VA_ARGS__; }               \
  } while (0)

static int vec_of(int es, std::initializer_list<const void*> ptrs, std::initializer_list<long long> qs) {
  int v = pick_vec(qs);
  for (const void* p : ptrs)
    if (p) { long long al = ptr_align_elems(p, es); while (v > al) v >>= 1; }
  if (es == 4 && v > 4) v = 4;
  return v;
}

extern "C" int cenet_dwconv3x3_wgrad(const void* x, int x_dtype, long long ldx, const void* dz, int dz_dtype, long long ldz, int B,
                                     int H, int W, int C, int dil, int up2, float* dw, float* dbias, float* ws, long long ws_elems,
                                     cenet_stream_t st) {
  CENET_REQUIRE(x && dz && dw && ws, "cenet_dwconv3x3_wgrad: null pointer");
  CENET_REQUIRE(x_dtype == dz_dtype, "cenet_dwconv3x3_wgrad: x and dz must share one dtype");
  cudaStream_t s = to_stream(st);
  const long long rows = (long long)B * H * W;
  CENET_DISPATCH(x_dtype, T, {
    int Vv = vec_of(sizeof(T), {x, dz}, {C, ldx, ldz});
    if (Vv > 4) Vv = 4;                                   // 10 accumulators per channel: keep the register footprint small
    static const bool use_staged = getenv("CENET_B200_DW_STAGED") == nullptr || atoi(getenv("CENET_B200_DW_STAGED")) != 0;
    // (a thread owns 4 consecutive pixels: rows narrower than 48 pixels leave most of the 16 pixel groups idle)
    if (use_staged && x_dtype == CENET_BF16 && dil == 1 && !up2 && W >= 48 && W % 4 == 0 && C % WS_C == 0 && ldx % 8 == 0 && ldz % 8 == 0 &&
        ((((uintptr_t)x | (uintptr_t)dz) & 15) == 0) && B <= 65535) {
      int rpc = H;
      while ((long long)(C / WS_C) * cdiv(H, rpc) * B < 4LL * kNumSMs && rpc > 4) rpc = (rpc + 1) / 2;
      const int nbands = cdiv(H, rpc);
      const size_t ring = (size_t)WS_NRX * (W + 2) * 128 + (size_t)WS_NRZ * W * 128;
      const size_t smem = std::max<size_t>(ring, 16 * 640 * sizeof(float));
      if (smem <= 200 * 1024 && (long long)B * nbands * 10 * C <= ws_elems) {
        static std::atomic<size_t> configured{0};
        if (smem > 48 * 1024 && configured.load() < smem) {
          cudaFuncSetAttribute(dw_wgrad_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
          configured.store(200 * 1024);
        }
        dw_wgrad_staged_kernel<<<dim3(C / WS_C, nbands, B), WS_T, smem, s>>>((const bf16*)x, (int)ldx, (const bf16*)dz, (int)ldz, H, W, C,
                                                                           rpc, ws);
        CENET_LAUNCH_CHECK("dw_wgrad_staged");
        dw_wgrad_finalize_kernel<<<cdiv(10 * C, kFinOut), kFinThreads, 0, s>>>(ws, B * nbands, C, dw, dbias);
        CENET_LAUNCH_CHECK("dw_wgrad_finalize");
        return 0;
      }
    }
    // rows of the plan are IMAGE rows (b, h); every thread takes whole rows, at least one
    ColPlan p = plan_cols(rows, C, Vv);
    const bool window = dil == 1 && !up2;
    int wsegs = 1;
    if (!window)
      while (wsegs < 8 && (long long)cdiv(B * H * wsegs, p.nrl) * p.gy < 2 * kNumSMs && W / (2 * wsegs) >= 7) wsegs *= 2;
    const int R = B * H * wsegs;
    {
      long long want = (R + p.nrl - 1) / p.nrl;
      long long cap = std::max(1, 8 * kNumSMs / p.gy);
      if (want > cap) want = cap;
      const long long per = (R + want - 1) / want;
      p.rows_per_block = (int)((per + p.nrl - 1) / p.nrl * p.nrl);
      p.nrb = (R + p.rows_per_block - 1) / p.rows_per_block;
    }
    CENET_REQUIRE((long long)p.nrb * 10 * C <= ws_elems, "cenet_dwconv3x3_wgrad: workspace too small");
    if (window) {
      DISPATCH_V(Vv, (dw_wgrad_partial_kernel<T, V, true><<<dim3(p.nrb, p.gy), kColThreads, 0, s>>>(
                          (const T*)x, ldx, (const T*)dz, ldz, B, H, W, C, dil, up2, p.ngrp, p.nrl, p.rows_per_block, 1, ws)));
    } else {
      DISPATCH_V(Vv, (dw_wgrad_partial_kernel<T, V, false><<<dim3(p.nrb, p.gy), kColThreads, 0, s>>>(
                          (const T*)x, ldx, (const T*)dz, ldz, B, H, W, C, dil, up2, p.ngrp, p.nrl, p.rows_per_block, wsegs, ws)));
    }
    CENET_LAUNCH_CHECK("dw_wgrad_partial");
    dw_wgrad_finalize_kernel<<<cdiv(10 * C, kFinOut), kFinThreads, 0, s>>>(ws, p.nrb, C, dw, dbias);
    CENET_LAUNCH_CHECK("dw_wgrad_finalize");
  });
  return 0;
}

extern "C" int cenet_sumpool2(const void* x, int x_dtype, void* y, int y_dtype, int B, int Ho, int Wo, int C, int acc,
                              cenet_stream_t st) {
  CENET_REQUIRE(x && y && x_dtype == y_dtype, "cenet_sumpool2: bad arguments");
  cudaStream_t s = to_stream(st);
  CENET_DISPATCH(x_dtype, T, {
    int Vv = vec_of(sizeof(T), {x, y}, {C});
    const long long total = (long long)B * Ho * Wo * (C / Vv);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 8LL * kNumSMs);
    DISPATCH_V(Vv, (sumpool2_kernel<T, V><<<blocks, 256, 0, s>>>((const T*)x, (T*)y, B, Ho, Wo, C, acc)));
    CENET_LAUNCH_CHECK("sumpool2");
  });
  return 0;
}

extern "C" int cenet_col2im(const void* dcol, int c_dtype, void* dx, int x_dtype, int B, int H, int W, int Cin, int k, int stride,
                            int pad, int Ho, int Wo, int Kp, int acc, cenet_stream_t st) {
  CENET_REQUIRE(dcol && dx && c_dtype == x_dtype, "cenet_col2im: bad arguments");
  CENET_REQUIRE(Kp >= k * k * Cin && stride >= 1, "cenet_col2im: bad shape");
  cudaStream_t s = to_stream(st);
  CENET_DISPATCH(x_dtype, T, {
    int Vv = vec_of(sizeof(T), {dcol, dx}, {Cin, Kp});
    const long long total = (long long)B * H * W * (Cin / Vv);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 8LL * kNumSMs);
    DISPATCH_V(Vv, (col2im_kernel<T, V><<<blocks, 256, 0, s>>>((const T*)dcol, (T*)dx, B, H, W, Cin, k, stride, pad, Ho, Wo, Kp, acc)));
    CENET_LAUNCH_CHECK("col2im");
  });
  return 0;
}
