// Training-mode normalisation kernels: BatchNorm batch statistics / apply / backward (all 53 BN2d of the decoder and
// head, blocks.py:169-185, cfam.py:365-374, unet.py:174-214, nlb.py:140-143) and LayerNorm backward (pvtv2.py:146-147,
// 189, 320).  HBM-bound: every pass reads/writes each element once; per-channel reductions are two-stage and
// deterministic (train_common.cuh).
#include "train_common.cuh"

__global__ void __launch_bounds__(kFinThreads) finalize_partials_kernel(const float* __restrict__ ws, int nblk, int n, float* outA,
                                                                        int nA, float* outB, float scale) {
  __shared__ float sm[kFinThreads];
  const int i = blockIdx.x * kFinOut + threadIdx.x % kFinOut;
  float t[1];
  fin_lane_sums<1>(nblk, i < n, t, sm, [&](int b, int) { return ws[(size_t)b * n + i]; });
  if (threadIdx.x >= kFinOut || i >= n) return;
  const float s = t[0] * scale;
  if (i < nA) { if (outA) outA[i] = s; }
  else if (outB) outB[i - nA] = s;
}

int launch_finalize(const float* ws, int nblk, int n, float* outA, int nA, float* outB, float scale, cudaStream_t s) {
  finalize_partials_kernel<<<cdiv(n, kFinOut), kFinThreads, 0, s>>>(ws, nblk, n, outA, nA, outB, scale);
  CENET_LAUNCH_CHECK("finalize_partials");
  return 0;
}

namespace {

// ------------------------------------------------------------------------------------------------ BN statistics
template <typename T, int V>
__global__ void __launch_bounds__(kColThreads) bn_stats_partial_kernel(const T* __restrict__ x, long long ld, long long rows, int C,
                                                                       int ngrp, int nrl, int rows_per_block, float* __restrict__ ws) {
  __shared__ float smem[V * kColThreads];
  const int grp = threadIdx.x % ngrp, rl = threadIdx.x / ngrp;
  const int c0 = (blockIdx.y * ngrp + grp) * V;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s1[V], s2[V];
#pragma unroll
  for (int v = 0; v < V; v++) s1[v] = s2[v] = 0.f;
  if (c0 < C) {
#pragma unroll 4                     // independent row loads in flight (these passes are bound by bytes in flight)
    for (long long r = r0 + rl; r < r1; r += nrl) {
      float xv[V];
      ldv<V>(x + r * ld + c0, xv);
#pragma unroll
      for (int v = 0; v < V; v++) { s1[v] += xv[v]; s2[v] = fmaf(xv[v], xv[v], s2[v]); }
    }
  }
  col_block_reduce<V>(s1, smem, grp, rl, ngrp, nrl);
  col_block_reduce<V>(s2, smem, grp, rl, ngrp, nrl);
  if (rl == 0 && c0 < C) {
#pragma unroll
    for (int v = 0; v < V; v++)
      if (c0 + v < C) {
        ws[((size_t)blockIdx.x * 2 + 0) * C + c0 + v] = s1[v];
        ws[((size_t)blockIdx.x * 2 + 1) * C + c0 + v] = s2[v];
      }
  }
}

__global__ void bn_stats_finalize_kernel(const float* __restrict__ ws, int nblk, int C, long long rows, const float* gamma,
                                         const float* beta, float* rmean, float* rvar, long long* nbt, float momentum, float eps,
                                         float* scale, float* shift, float* mean, float* rstd) {
  __shared__ double sm[2 * kFinThreads];
  const int c = blockIdx.x * kFinOut + threadIdx.x % kFinOut;
  double t[2];
  fin_lane_sums<2>(nblk, c < C, t, sm, [&](int i, int k) { return ws[((size_t)i * 2 + k) * C + c]; });
  if (threadIdx.x >= kFinOut) return;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  const double a = t[0], b = t[1];
  const double mu = a / (double)rows;
  double var = b / (double)rows - mu * mu;
  if (var < 0.0) var = 0.0;
  const float r = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, bb = beta ? beta[c] : 0.f;
  mean[c] = (float)mu;
  rstd[c] = r;
  scale[c] = g * r;
  shift[c] = bb - (float)mu * g * r;
  if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mu;
  if (rvar) {
    const double unb = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
  }
}

// eval-mode ("frozen") BatchNorm: the running statistics ARE the statistics; nothing is updated
__global__ void bn_frozen_stats_kernel(int C, const float* gamma, const float* beta, const float* rmean, const float* rvar, float eps,
                                       float* scale, float* shift, float* mean, float* rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float r = rsqrtf(rvar[c] + eps), g = gamma ? gamma[c] : 1.f, bb = beta ? beta[c] : 0.f;
  mean[c] = rmean[c]; rstd[c] = r; scale[c] = g * r; shift[c] = bb - rmean[c] * g * r;
}

// ------------------------------------------------------------------------------------------------ affine + act
template <typename T, int V>
__global__ void __launch_bounds__(256) affine_act_kernel(const T* __restrict__ a, long long lda, const float* __restrict__ sa,
                                                         const float* __restrict__ ta, const T* __restrict__ b, long long ldb,
                                                         const float* __restrict__ sb, const float* __restrict__ tb,
                                                         T* __restrict__ out, long long ldo, long long rows, int C, int act,
                                                         float slope) {
  const int groups = C / V;
  const long long total = rows * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int c0 = (int)(i % groups) * V;
    float v[V];
    ldv<V>(a + r * lda + c0, v);
    if (sa) {
#pragma unroll
      for (int j = 0; j < V; j++) v[j] = fmaf(v[j], sa[c0 + j], ta[c0 + j]);
    }
    if (b) {
      float w[V];
      ldv<V>(b + r * ldb + c0, w);
#pragma unroll
      for (int j = 0; j < V; j++) v[j] += sb ? fmaf(w[j], sb[c0 + j], tb[c0 + j]) : w[j];
    }
    if (act != CENET_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < V; j++) v[j] = apply_act(v[j], act, slope);
    }
    stv<V>(out + r * ldo + c0, v);
  }
}

// ------------------------------------------------------------------------------------------------ BN backward
template <typename T, int V>
__global__ void __launch_bounds__(kColThreads) bn_bwd_partial_kernel(const T* __restrict__ dy, const T* __restrict__ y, long long ldy,
                                                                     const T* __restrict__ a, long long lda,
                                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                     long long rows, int C, int act, float slope, int ngrp, int nrl,
                                                                     int rows_per_block, float* __restrict__ ws) {
  __shared__ float smem[V * kColThreads];
  const int grp = threadIdx.x % ngrp, rl = threadIdx.x / ngrp;
  const int c0 = (blockIdx.y * ngrp + grp) * V;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s1[V], s2[V], mu[V], rs[V];
#pragma unroll
  for (int v = 0; v < V; v++) {
    s1[v] = s2[v] = 0.f;
    mu[v] = (c0 + v < C) ? mean[c0 + v] : 0.f;
    rs[v] = (c0 + v < C) ? rstd[c0 + v] : 0.f;
  }
  if (c0 < C) {
#pragma unroll 4                     // independent row loads in flight (these passes are bound by bytes in flight)
    for (long long r = r0 + rl; r < r1; r += nrl) {
      float g[V], av[V];
      ldv<V>(dy + r * ldy + c0, g);
      ldv<V>(a + r * lda + c0, av);
      if (y && act != CENET_ACT_NONE) {
        float yv[V];
        ldv<V>(y + r * ldy + c0, yv);
#pragma unroll
        for (int v = 0; v < V; v++) g[v] *= act_grad_from_out(yv[v], act, slope);
      }
#pragma unroll
      for (int v = 0; v < V; v++) { s1[v] += g[v]; s2[v] = fmaf(g[v], (av[v] - mu[v]) * rs[v], s2[v]); }
    }
  }
  col_block_reduce<V>(s1, smem, grp, rl, ngrp, nrl);
  col_block_reduce<V>(s2, smem, grp, rl, ngrp, nrl);
  if (rl == 0 && c0 < C) {
#pragma unroll
    for (int v = 0; v < V; v++)
      if (c0 + v < C) {
        ws[((size_t)blockIdx.x * 2 + 0) * C + c0 + v] = s1[v];
        ws[((size_t)blockIdx.x * 2 + 1) * C + c0 + v] = s2[v];
      }
  }
}

// sums[0..C) = d(beta), sums[C..2C) = d(gamma); also written to the parameter gradients
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ ws, int nblk, int C, float* sums, float* dgamma, float* dbeta, int frozen) {
  __shared__ float sm[2 * kFinThreads];
  const int c = blockIdx.x * kFinOut + threadIdx.x % kFinOut;
  float t[2];
  fin_lane_sums<2>(nblk, c < C, t, sm, [&](int i, int k) { return ws[((size_t)i * 2 + k) * C + c]; });
  if (threadIdx.x >= kFinOut || c >= C) return;
  const float a = t[0], b = t[1];
  sums[c] = frozen ? 0.f : a;                // frozen statistics: d(input) = gamma * rstd * g, no batch terms
  sums[C + c] = frozen ? 0.f : b;
  if (dbeta) dbeta[c] = a;
  if (dgamma) dgamma[c] = b;
}

template <typename T, int V>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ y, long long ldy,
                                                           const T* __restrict__ a, long long lda, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ sums, long long rows, int C, int act, float slope,
                                                           T* __restrict__ da, int acc_da, T* __restrict__ dres, long long lddres,
                                                           int acc_dres) {
  const int groups = C / V;
  const long long total = rows * groups;
  const float inv = 1.f / (float)rows;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int c0 = (int)(i % groups) * V;
    float g[V], av[V], o[V];
    ldv<V>(dy + r * ldy + c0, g);
    ldv<V>(a + r * lda + c0, av);
    if (y && act != CENET_ACT_NONE) {
      float yv[V];
      ldv<V>(y + r * ldy + c0, yv);
#pragma unroll
      for (int v = 0; v < V; v++) g[v] *= act_grad_from_out(yv[v], act, slope);
    }
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int c = c0 + v;
      const float xh = (av[v] - mean[c]) * rstd[c];
      o[v] = gamma[c] * rstd[c] * (g[v] - sums[c] * inv - xh * sums[C + c] * inv);
    }
    if (acc_da) {
      float old[V];
      ldv<V>(da + r * lda + c0, old);
#pragma unroll
      for (int v = 0; v < V; v++) o[v] += old[v];
    }
    stv<V>(da + r * lda + c0, o);
    if (dres) {
      if (acc_dres) {
        float old[V];
        ldv<V>(dres + r * lddres + c0, old);
#pragma unroll
        for (int v = 0; v < V; v++) g[v] += old[v];
      }
      stv<V>(dres + r * lddres + c0, g);
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// LPR lanes per row (8 for C = 64, 16 for C = 128, 32 otherwise) with 8-element (16-byte bf16) vectors, 32 / LPR rows per warp
// at a time: the four row reductions cost log2(LPR) shuffles for several rows at once and every load / store is a full
// vector (the first version used one warp per row with 4-byte accesses: 1 TB/s).  Every lane accumulates d(gamma) / d(beta)
// of ITS channels over the rows it sees; block partials go through shared memory in fixed order -> deterministic.
template <typename T, int LPR, int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            const float* __restrict__ gamma, float eps, long long rows, int C,
                                                            T* __restrict__ dx, int acc, float* __restrict__ ws) {
  constexpr int RPW = 32 / LPR;                           // rows per warp pass
  extern __shared__ float dyn[];                          // [8 warps * RPW][2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % LPR, grp = lane / LPR;
  const int nvec = C >> 3;
  const float invC = 1.f / (float)C;
  float gsum[NV][8], bsum[NV][8], gam[NV][8];
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const int vi = sub + i * LPR;
#pragma unroll
    for (int j = 0; j < 8; j++) { gsum[i][j] = bsum[i][j] = 0.f; gam[i][j] = vi < nvec ? gamma[vi * 8 + j] : 0.f; }
  }
  const long long rstride = (long long)gridDim.x * 8 * RPW;
  for (long long r0 = ((long long)blockIdx.x * 8 + warp) * RPW; r0 < rows; r0 += rstride) {
    const long long r = r0 + grp;
    const bool live = r < rows;
    float xv[NV][8], gv[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      const int vi = sub + i * LPR;
      if (live && vi < nvec) {
        ldv<8>(x + r * C + vi * 8, xv[i]);
        ldv<8>(dy + r * C + vi * 8, gv[i]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) xv[i][j] = gv[i][j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; j++) s += xv[i][j];
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      const bool in = sub + i * LPR < nvec;
#pragma unroll
      for (int j = 0; j < 8; j++) { xv[i][j] = in ? xv[i][j] - mu : 0.f; q = fmaf(xv[i][j], xv[i][j], q); }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = rsqrtf(q * invC + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) {
        xv[i][j] *= rs;                                   // x_hat
        gsum[i][j] = fmaf(gv[i][j], xv[i][j], gsum[i][j]);
        bsum[i][j] += gv[i][j];
        gv[i][j] *= gam[i][j];                            // dy * gamma
        s1 += gv[i][j];
        s2 = fmaf(gv[i][j], xv[i][j], s2);
      }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 *= invC; s2 *= invC;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      const int vi = sub + i * LPR;
      if (live && vi < nvec) {
        float o8[8];
#pragma unroll
        for (int j = 0; j < 8; j++) o8[j] = rs * (gv[i][j] - s1 - xv[i][j] * s2);
        if (acc) {
          float old[8];
          ldv<8>(dx + r * C + vi * 8, old);
#pragma unroll
          for (int j = 0; j < 8; j++) o8[j] += old[j];
        }
        stv<8>(dx + r * C + vi * 8, o8);
      }
    }
  }
  // block partials of d(gamma), d(beta): one slot per (warp, row group), summed in slot order
  const int slot = warp * RPW + grp;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const int vi = sub + i * LPR;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        dyn[(slot * 2 + 0) * C + vi * 8 + j] = gsum[i][j];
        dyn[(slot * 2 + 1) * C + vi * 8 + j] = bsum[i][j];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    const int k = i / C, c = i - k * C;
    float t = 0.f;
    for (int w = 0; w < 8 * RPW; w++) t += dyn[(w * 2 + k) * C + c];
    ws[(size_t)blockIdx.x * 2 * C + i] = t;
  }
}

template <typename T>
int vec_for(const void* p0, const void* p1, const void* p2, const void* p3, std::initializer_list<long long> qs) {
  int v = pick_vec(qs);
  for (const void* p : {p0, p1, p2, p3})
    if (p) { long long al = ptr_align_elems(p, sizeof(T)); while (v > al) v >>= 1; }
  return v;
}
}  // namespace

#define DISPATCH_V(V_, ...)                                  \
  do {                                                       \
    if (V_ == 8) { constexpr int V = 8; __VA_ARGS__; }       \
    else if (V_ == 4) { constexpr int V = 4; __VA_ARGS__; }  \
    else if (V_ == 2) { constexpr int V = 2; __VA_ARGS__; }  \
    else { constexpr int V = 1; __VA_ARGS__; }               \
  } while (0)

extern "C" int cenet_bn_stats(const void* x, int x_dtype, long long ldx, long long rows, int C, const float* gamma,
                              const float* beta, float* rmean, float* rvar, long long* nbt, float momentum, float eps,
                              int frozen, float* scale, float* shift, float* mean, float* rstd, float* ws, long long ws_elems,
                              cenet_stream_t st) {
  CENET_REQUIRE(x && scale && shift && mean && rstd && ws, "cenet_bn_stats: null pointer");
  CENET_REQUIRE(rows > 0 && C > 0, "cenet_bn_stats: bad shape");
  cudaStream_t s = to_stream(st);
  if (frozen) {                               // eval-mode BatchNorm inside a gradient-enabled pass
    CENET_REQUIRE(rmean && rvar, "cenet_bn_stats(frozen): running statistics required");
    bn_frozen_stats_kernel<<<cdiv(C, 128), 128, 0, s>>>(C, gamma, beta, rmean, rvar, eps, scale, shift, mean, rstd);
    CENET_LAUNCH_CHECK("bn_frozen_stats");
    return 0;
  }
  CENET_DISPATCH(x_dtype, T, {
    int Vv = vec_for<T>(x, nullptr, nullptr, nullptr, {C, ldx});
    if (sizeof(T) == 4 && Vv > 4) Vv = 4;
    ColPlan p = plan_cols(rows, C, Vv);
    CENET_REQUIRE((long long)p.nrb * 2 * C <= ws_elems, "cenet_bn_stats: workspace too small");
    DISPATCH_V(Vv, (bn_stats_partial_kernel<T, V><<<dim3(p.nrb, p.gy), kColThreads, 0, s>>>(
                        (const T*)x, ldx, rows, C, p.ngrp, p.nrl, p.rows_per_block, ws)));
    CENET_LAUNCH_CHECK("bn_stats_partial");
    bn_stats_finalize_kernel<<<cdiv(C, kFinOut), kFinThreads, 0, s>>>(ws, p.nrb, C, rows, gamma, beta, rmean, rvar, nbt, momentum, eps, scale,
                                                          shift, mean, rstd);
    CENET_LAUNCH_CHECK("bn_stats_finalize");
  });
  return 0;
}

extern "C" int cenet_affine_act(const void* a, int a_dtype, long long lda, const float* sa, const float* ta, const void* b,
                                int b_dtype, long long ldb, const float* sb, const float* tb, void* out, int o_dtype,
                                long long ldo, long long rows, int C, int act, float slope, cenet_stream_t st) {
  CENET_REQUIRE(a && out, "cenet_affine_act: null pointer");
  CENET_REQUIRE(a_dtype == o_dtype && (!b || b_dtype == a_dtype), "cenet_affine_act: operands must share one dtype");
  CENET_REQUIRE((sa == nullptr) == (ta == nullptr) && (sb == nullptr) == (tb == nullptr), "cenet_affine_act: scale/shift pairs");
  if (rows == 0) return 0;
  cudaStream_t s = to_stream(st);
  CENET_DISPATCH(a_dtype, T, {
    int Vv = vec_for<T>(a, b, out, nullptr, {C, lda, ldo, b ? ldb : 8});
    if (sizeof(T) == 4 && Vv > 4) Vv = 4;
    const long long total = rows * (C / Vv);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 8LL * kNumSMs);
    DISPATCH_V(Vv, (affine_act_kernel<T, V><<<blocks, 256, 0, s>>>((const T*)a, lda, sa, ta, (const T*)b, ldb, sb, tb, (T*)out, ldo,
                                                                  rows, C, act, slope)));
    CENET_LAUNCH_CHECK("affine_act");
  });
  return 0;
}

extern "C" int cenet_bn_bwd(const void* dy, int dy_dtype, const void* y, int y_dtype, long long ldy, const void* a, int a_dtype,
                            long long lda, const float* mean, const float* rstd, const float* gamma, long long rows, int C,
                            int act, float slope, void* da, int da_dtype, int acc_da, float* dgamma, float* dbeta, void* dres,
                            int dres_dtype, long long lddres, int acc_dres, int frozen, float* ws, long long ws_elems,
                            cenet_stream_t st) {
  CENET_REQUIRE(dy && a && da && mean && rstd && gamma && ws, "cenet_bn_bwd: null pointer");
  CENET_REQUIRE(dy_dtype == a_dtype && da_dtype == a_dtype && (!y || y_dtype == a_dtype) && (!dres || dres_dtype == a_dtype),
                "cenet_bn_bwd: operands must share one dtype");
  CENET_REQUIRE(act == CENET_ACT_NONE || act == CENET_ACT_RELU || act == CENET_ACT_LEAKY, "cenet_bn_bwd: unsupported activation");
  cudaStream_t s = to_stream(st);
  CENET_DISPATCH(a_dtype, T, {
    int Vv = vec_for<T>(dy, y, a, da, {C, ldy, lda, dres ? lddres : 8});
    if (dres) { long long al = ptr_align_elems(dres, sizeof(T)); while (Vv > al) Vv >>= 1; }
    if (sizeof(T) == 4 && Vv > 4) Vv = 4;
    ColPlan p = plan_cols(rows, C, Vv);
    CENET_REQUIRE((long long)p.nrb * 2 * C + 2 * C <= ws_elems, "cenet_bn_bwd: workspace too small");
    float* sums = ws + (size_t)p.nrb * 2 * C;
    DISPATCH_V(Vv, (bn_bwd_partial_kernel<T, V><<<dim3(p.nrb, p.gy), kColThreads, 0, s>>>(
                        (const T*)dy, (const T*)y, ldy, (const T*)a, lda, mean, rstd, rows, C, act, slope, p.ngrp, p.nrl,
                        p.rows_per_block, ws)));
    CENET_LAUNCH_CHECK("bn_bwd_partial");
    bn_bwd_finalize_kernel<<<cdiv(C, kFinOut), kFinThreads, 0, s>>>(ws, p.nrb, C, sums, dgamma, dbeta, frozen);
    CENET_LAUNCH_CHECK("bn_bwd_finalize");
    const long long total = rows * (C / Vv);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 8LL * kNumSMs);
    DISPATCH_V(Vv, (bn_bwd_apply_kernel<T, V><<<blocks, 256, 0, s>>>((const T*)dy, (const T*)y, ldy, (const T*)a, lda, mean, rstd,
                                                                    gamma, sums, rows, C, act, slope, (T*)da, acc_da, (T*)dres,
                                                                    lddres, acc_dres)));
    CENET_LAUNCH_CHECK("bn_bwd_apply");
  });
  return 0;
}

extern "C" int cenet_layernorm_bwd(const void* dy, const void* x, int dtype, const float* gamma, float eps, long long rows, int C,
                                   void* dx, int acc, float* dgamma, float* dbeta, float* ws, long long ws_elems,
                                   int* n_partials, cenet_stream_t st) {
  CENET_REQUIRE(dy && x && gamma && dx && dgamma && dbeta && ws, "cenet_layernorm_bwd: null pointer");
  CENET_REQUIRE(C == 64 || C == 128 || C == 320 || C == 512, "cenet_layernorm_bwd: C=%d not instantiated (64/128/320/512)", C);
  if (n_partials) *n_partials = 0;
  if (rows == 0) return 0;
  cudaStream_t s = to_stream(st);
  CENET_REQUIRE(((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dx) & 15) == 0), "cenet_layernorm_bwd: rows must be 16-byte aligned");
  const int lpr = C == 64 ? 8 : C == 128 ? 16 : 32, rpw = 32 / lpr;
  int nblk = (int)std::min<long long>((rows + 8 * rpw - 1) / (8 * rpw), 8LL * kNumSMs);
  CENET_REQUIRE((long long)nblk * 2 * C <= ws_elems, "cenet_layernorm_bwd: workspace too small");
  const size_t smem = (size_t)8 * rpw * 2 * C * sizeof(float);
#define LN_CASE(LPR, NV)                                                                                                    \
  layernorm_bwd_kernel<T, LPR, NV><<<nblk, 256, smem, s>>>((const T*)dy, (const T*)x, gamma, eps, rows, C, (T*)dx, acc, ws)
  CENET_DISPATCH(dtype, T, {
    if (C == 64) LN_CASE(8, 1);
    else if (C == 128) LN_CASE(16, 1);
    else LN_CASE(32, 2);                                   // 320 (40 vectors) and 512 (64 vectors)
    CENET_LAUNCH_CHECK("layernorm_bwd");
  });
#undef LN_CASE
  // ws layout per block: [2][C] = d(gamma) | d(beta)
  if (n_partials) {                 // deferred: the caller reduces the partials with cenet_wgrad_reduce_batch (two jobs, N = 1, K = C)
    *n_partials = nblk;
    return 0;
  }
  return launch_finalize(ws, nblk, 2 * C, dgamma, C, dbeta, 1.f, s);
}
