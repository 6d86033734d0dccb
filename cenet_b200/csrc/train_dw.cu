// Backward helpers of the depthwise / strided-conv family (HBM-bound):
//   dwconv3x3_wgrad : filter + bias gradient of every depthwise 3x3 (pvtv2.py:364-370, cfam.py:151-152, blocks.py:169-178,
//                     blocks.py:303-311) -- 10 per-channel reductions over all pixels, two-stage deterministic.
//   (the data gradient of a depthwise conv is the forward kernel with flipped taps, see train.py::dwconv)
//   sumpool2        : adjoint of the nearest x2 up-sampling of EUCB (blocks.py:304)
//   col2im          : adjoint of cenet_im2col (patch-embed / SR-conv data gradients, pvtv2.py:164-165, 68)
#include "train_common.cuh"

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
                                                                       int ngrp, int nrl, int rows_per_block, float* __restrict__ ws) {
  __shared__ float smem[V * kColThreads];
  const int grp = threadIdx.x % ngrp, rl = threadIdx.x / ngrp;
  const int c0 = (blockIdx.y * ngrp + grp) * V;
  const int R = B * H;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  const int Hi = up2 ? H >> 1 : H, Wi = up2 ? W >> 1 : W;
  float acc[10][V];
#pragma unroll
  for (int k = 0; k < 10; k++)
#pragma unroll
    for (int v = 0; v < V; v++) acc[k][v] = 0.f;
  if (c0 < C) {
    for (int r = r0 + rl; r < r1; r += nrl) {
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
        for (int w = 0; w < W; w++) {
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
    // rows of the plan are IMAGE rows (b, h); every thread takes whole rows, at least one
    ColPlan p = plan_cols(rows, C, Vv);
    const int R = B * H;
    {
      long long want = (R + p.nrl - 1) / p.nrl;
      long long cap = std::max(1, 8 * kNumSMs / p.gy);
      if (want > cap) want = cap;
      const long long per = (R + want - 1) / want;
      p.rows_per_block = (int)((per + p.nrl - 1) / p.nrl * p.nrl);
      p.nrb = (R + p.rows_per_block - 1) / p.rows_per_block;
    }
    CENET_REQUIRE((long long)p.nrb * 10 * C <= ws_elems, "cenet_dwconv3x3_wgrad: workspace too small");
    if (dil == 1 && !up2) {
      DISPATCH_V(Vv, (dw_wgrad_partial_kernel<T, V, true><<<dim3(p.nrb, p.gy), kColThreads, 0, s>>>(
                          (const T*)x, ldx, (const T*)dz, ldz, B, H, W, C, dil, up2, p.ngrp, p.nrl, p.rows_per_block, ws)));
    } else {
      DISPATCH_V(Vv, (dw_wgrad_partial_kernel<T, V, false><<<dim3(p.nrb, p.gy), kColThreads, 0, s>>>(
                          (const T*)x, ldx, (const T*)dz, ldz, B, H, W, C, dil, up2, p.ngrp, p.nrl, p.rows_per_block, ws)));
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
