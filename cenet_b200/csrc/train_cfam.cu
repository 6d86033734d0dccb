// CFAM statistics gates in train mode and their backward (cfam.py): CCU (251-264), SRM (93-101), SiLU gating (303-304),
// layer-scale residuals and the non-local mix (365-374, nlb.py:145-148).  All HBM-bound or tiny; reductions are
// deterministic (fixed-order partials).
#include "train_common.cuh"

namespace {
// ---------------------------------------------------------------------------------------------- CCU
// per (b, c): max (+ first arg max), mean, biased std over HW;  block = (b, tile of channel groups)
template <typename T, int V>
__global__ void __launch_bounds__(kColThreads) ccu_stats_kernel(const T* __restrict__ x, float* __restrict__ u, int* __restrict__ arg,
                                                                int HW, int C, int ngrp, int nrl) {
  __shared__ float smem[V * kColThreads];
  __shared__ int sidx[V * kColThreads];
  const int grp = threadIdx.x % ngrp, rl = threadIdx.x / ngrp;
  const int c0 = (blockIdx.x * ngrp + grp) * V;
  const int b = blockIdx.y;
  const T* xb = x + (long long)b * HW * C;
  float mx[V], s1[V], s2[V];
  int ix[V];
#pragma unroll
  for (int v = 0; v < V; v++) { mx[v] = -INFINITY; s1[v] = s2[v] = 0.f; ix[v] = 0; }
  if (c0 < C) {
    for (int r = rl; r < HW; r += nrl) {
      float xv[V];
      ldv<V>(xb + (long long)r * C + c0, xv);
#pragma unroll
      for (int v = 0; v < V; v++) {
        if (xv[v] > mx[v]) { mx[v] = xv[v]; ix[v] = r; }
        s1[v] += xv[v];
        s2[v] = fmaf(xv[v], xv[v], s2[v]);
      }
    }
  }
  // max / arg max across row lanes
  __syncthreads();
#pragma unroll
  for (int v = 0; v < V; v++) { smem[(v * nrl + rl) * ngrp + grp] = mx[v]; sidx[(v * nrl + rl) * ngrp + grp] = ix[v]; }
  __syncthreads();
  if (rl == 0) {
#pragma unroll
    for (int v = 0; v < V; v++) {
      float m = -INFINITY; int mi = 0;
      for (int l = 0; l < nrl; l++) {
        const float t = smem[(v * nrl + l) * ngrp + grp];
        const int ti = sidx[(v * nrl + l) * ngrp + grp];
        if (t > m || (t == m && ti < mi)) { m = t; mi = ti; }
      }
      mx[v] = m; ix[v] = mi;
    }
  }
  col_block_reduce<V>(s1, smem, grp, rl, ngrp, nrl);
  col_block_reduce<V>(s2, smem, grp, rl, ngrp, nrl);
  if (rl == 0 && c0 < C) {
#pragma unroll
    for (int v = 0; v < V; v++) {
      const int c = c0 + v;
      if (c >= C) continue;
      const float mean = s1[v] / HW;
      const float var = fmaxf(s2[v] / HW - mean * mean, 0.f);
      float* up = u + ((long long)b * C + c) * 3;
      up[0] = mx[v]; up[1] = mean; up[2] = sqrtf(var);
      arg[(long long)b * C + c] = ix[v];
    }
  }
}

// dgate[b, c] = sum_hw dx1 * xb
template <typename T, int V>
__global__ void __launch_bounds__(kColThreads) ccu_dgate_kernel(const T* __restrict__ dx1, const T* __restrict__ xb, float* __restrict__ dgate,
                                                                int HW, int C, int ngrp, int nrl) {
  __shared__ float smem[V * kColThreads];
  const int grp = threadIdx.x % ngrp, rl = threadIdx.x / ngrp;
  const int c0 = (blockIdx.x * ngrp + grp) * V;
  const int b = blockIdx.y;
  float s[V];
#pragma unroll
  for (int v = 0; v < V; v++) s[v] = 0.f;
  if (c0 < C) {
    for (int r = rl; r < HW; r += nrl) {
      float a[V], d[V];
      ldv<V>(xb + ((long long)b * HW + r) * C + c0, a);
      ldv<V>(dx1 + ((long long)b * HW + r) * C + c0, d);
#pragma unroll
      for (int v = 0; v < V; v++) s[v] = fmaf(a[v], d[v], s[v]);
    }
  }
  col_block_reduce<V>(s, smem, grp, rl, ngrp, nrl);
  if (rl == 0 && c0 < C) {
#pragma unroll
    for (int v = 0; v < V; v++)
      if (c0 + v < C) dgate[(long long)b * C + c0 + v] = s[v];
  }
}

// thread = channel; save[b,c,:] = z1[0..2], z2, xhat, rstd
__global__ void ccu_mlp_fwd_kernel(const float* __restrict__ u, const float* __restrict__ fc1, const float* __restrict__ fc2,
                                   const float* gamma, const float* beta, float* rmean, float* rvar, long long* nbt, float momentum,
                                   float eps, float* __restrict__ gate, float* __restrict__ save, int B, int C, int frozen) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && gamma && nbt && !frozen) *nbt += 1;
  if (c >= C) return;
  float w1[9], w2[3];
#pragma unroll
  for (int i = 0; i < 9; i++) w1[i] = fc1[c * 9 + i];
#pragma unroll
  for (int i = 0; i < 3; i++) w2[i] = fc2[c * 3 + i];
  float s1 = 0.f, s2 = 0.f;
  for (int b = 0; b < B; b++) {
    const float* up = u + ((long long)b * C + c) * 3;
    float* sp = save + ((long long)b * C + c) * 8;
    float z2 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const float z1 = w1[j * 3] * up[0] + w1[j * 3 + 1] * up[1] + w1[j * 3 + 2] * up[2];
      sp[j] = z1;
      z2 = fmaf(w2[j], fmaxf(z1, 0.f), z2);
    }
    sp[3] = z2;
    s1 += z2; s2 = fmaf(z2, z2, s2);
  }
  float mu = 0.f, rs = 1.f, g = 1.f, bt = 0.f;
  if (gamma && frozen) {                      // eval-mode BatchNorm1d: running statistics, no update
    mu = rmean[c]; rs = rsqrtf(rvar[c] + eps);
    g = gamma[c]; bt = beta[c];
  } else if (gamma) {
    mu = s1 / B;
    float var = 0.f;
    for (int b = 0; b < B; b++) { const float d = save[((long long)b * C + c) * 8 + 3] - mu; var = fmaf(d, d, var); }
    var /= B;
    rs = rsqrtf(var + eps);
    g = gamma[c]; bt = beta[c];
    if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * mu;
    if (rvar) rvar[c] = (1.f - momentum) * rvar[c] + momentum * (B > 1 ? var * B / (B - 1) : var);
  }
  for (int b = 0; b < B; b++) {
    float* sp = save + ((long long)b * C + c) * 8;
    const float xh = gamma ? (sp[3] - mu) * rs : sp[3];
    sp[4] = xh; sp[5] = rs;
    gate[(long long)b * C + c] = sigmoidf_(gamma ? fmaf(g, xh, bt) : xh);
  }
}

__global__ void ccu_mlp_bwd_kernel(const float* __restrict__ dgate, const float* __restrict__ u, const float* __restrict__ fc1,
                                   const float* __restrict__ fc2, const float* gamma, const float* beta, const float* __restrict__ save,
                                   float* __restrict__ du, float* dfc1, float* dfc2, float* dgamma, float* dbeta, int B, int C,
                                   int frozen) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float w1[9], w2[3], g1[9], g2[3];
#pragma unroll
  for (int i = 0; i < 9; i++) { w1[i] = fc1[c * 9 + i]; g1[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < 3; i++) { w2[i] = fc2[c * 3 + i]; g2[i] = 0.f; }
  const float g = gamma ? gamma[c] : 1.f, bt = gamma ? beta[c] : 0.f;
  float sg = 0.f, sb = 0.f;
  if (gamma) {
    for (int b = 0; b < B; b++) {
      const float* sp = save + ((long long)b * C + c) * 8;
      const float gt = sigmoidf_(fmaf(g, sp[4], bt));
      const float dzn = dgate[(long long)b * C + c] * gt * (1.f - gt);
      sb += dzn; sg = fmaf(dzn, sp[4], sg);
    }
  }
  for (int b = 0; b < B; b++) {
    const float* sp = save + ((long long)b * C + c) * 8;
    const float* up = u + ((long long)b * C + c) * 3;
    const float gt = sigmoidf_(gamma ? fmaf(g, sp[4], bt) : sp[4]);
    const float dzn = dgate[(long long)b * C + c] * gt * (1.f - gt);
    const float dz2 = gamma ? (frozen ? g * sp[5] * dzn : g * sp[5] * (dzn - sb / B - sp[4] * sg / B)) : dzn;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const float z1 = sp[j];
      g2[j] = fmaf(dz2, fmaxf(z1, 0.f), g2[j]);
      const float dh = z1 > 0.f ? dz2 * w2[j] : 0.f;
      g1[j * 3] = fmaf(dh, up[0], g1[j * 3]);
      g1[j * 3 + 1] = fmaf(dh, up[1], g1[j * 3 + 1]);
      g1[j * 3 + 2] = fmaf(dh, up[2], g1[j * 3 + 2]);
      d0 = fmaf(dh, w1[j * 3], d0); d1 = fmaf(dh, w1[j * 3 + 1], d1); d2 = fmaf(dh, w1[j * 3 + 2], d2);
    }
    float* dp = du + ((long long)b * C + c) * 3;
    dp[0] = d0; dp[1] = d1; dp[2] = d2;
  }
#pragma unroll
  for (int i = 0; i < 9; i++) dfc1[c * 9 + i] = g1[i];
#pragma unroll
  for (int i = 0; i < 3; i++) dfc2[c * 3 + i] = g2[i];
  if (dgamma) dgamma[c] = gamma ? sg : 0.f;
  if (dbeta) dbeta[c] = gamma ? sb : 0.f;
}

template <typename T, int V>
__global__ void __launch_bounds__(256) ccu_apply_bwd_kernel(const T* __restrict__ dx1, const T* __restrict__ xb, const float* __restrict__ gate,
                                                            const float* __restrict__ u, const int* __restrict__ arg,
                                                            const float* __restrict__ du, T* __restrict__ dxb, int acc, int B, int HW,
                                                            int C) {
  const int groups = C / V;
  const long long total = (long long)B * HW * groups;
  const float inv = 1.f / HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % groups) * V;
    const long long p = i / groups;
    const int r = (int)(p % HW), b = (int)(p / HW);
    float d[V], x[V], o[V];
    ldv<V>(dx1 + p * C + c0, d);
    ldv<V>(xb + p * C + c0, x);
#pragma unroll
    for (int v = 0; v < V; v++) {
      const long long bc = (long long)b * C + c0 + v;
      const float* up = u + bc * 3;
      const float* dp = du + bc * 3;
      float t = d[v] * gate[bc] + dp[1] * inv;
      if (up[2] > 0.f) t += dp[2] * (x[v] - up[1]) * inv / up[2];
      if (arg[bc] == r) t += dp[0];
      o[v] = t;
    }
    if (acc) {
      float old[V];
      ldv<V>(dxb + p * C + c0, old);
#pragma unroll
      for (int v = 0; v < V; v++) o[v] += old[v];
    }
    stv<V>(dxb + p * C + c0, o);
  }
}

// ---------------------------------------------------------------------------------------------- SRM
// warp per row: max (+ first arg max), mean, unbiased std over C channels
template <typename T>
__global__ void __launch_bounds__(256) row_stats_arg_kernel(const T* __restrict__ x, float* __restrict__ u, int* __restrict__ arg,
                                                            long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const T* xr = x + r * C;
  float mx = -INFINITY, s1 = 0.f;
  int ix = 0;
  for (int c = lane * 2; c < C; c += 64) {
    float v[2];
    ldv<2>(xr + c, v);
    if (v[0] > mx) { mx = v[0]; ix = c; }
    if (v[1] > mx) { mx = v[1]; ix = c + 1; }
    s1 += v[0] + v[1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
    if (om > mx || (om == mx && oi < ix)) { mx = om; ix = oi; }
  }
  const float mean = warp_sum(s1) / C;
  float q = 0.f;
  for (int c = lane * 2; c < C; c += 64) {
    float v[2];
    ldv<2>(xr + c, v);
    q = fmaf(v[0] - mean, v[0] - mean, q);
    q = fmaf(v[1] - mean, v[1] - mean, q);
  }
  q = warp_sum(q);
  if (lane == 0) {
    u[r * 3] = mx; u[r * 3 + 1] = mean; u[r * 3 + 2] = sqrtf(q / (C - 1));
    arg[r] = ix;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) row_dot_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ out,
                                                      long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float s = 0.f;
  for (int c = lane * 2; c < C; c += 64) {
    float x[2], y[2];
    ldv<2>(a + r * C + c, x);
    ldv<2>(b + r * C + c, y);
    s = fmaf(x[0], y[0], s);
    s = fmaf(x[1], y[1], s);
  }
  s = warp_sum(s);
  if (lane == 0) out[r] = s;
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < 8; w++) s += red[w];
  return s;
}

// f_pre = pwc(u) + dwc3x3(u), fg = GELU(f_pre); save[m] = (f_pre, fg); block partial sums of fg, fg^2 -> ws[blk*2 + {0,1}]
__global__ void __launch_bounds__(256) srm_fwd_a_kernel(const float* __restrict__ u, const float* __restrict__ pw, const float* __restrict__ dw,
                                                        float* __restrict__ save, float* __restrict__ ws, int B, int H, int W) {
  __shared__ float red[8];
  const long long M = (long long)B * H * W;
  float s1 = 0.f, s2 = 0.f;
  for (long long m = (long long)blockIdx.x * 256 + threadIdx.x; m < M; m += (long long)gridDim.x * 256) {
    const int w = (int)(m % W);
    const long long t = m / W;
    const int h = (int)(t % H);
    float f = pw[0] * u[m * 3] + pw[1] * u[m * 3 + 1] + pw[2] * u[m * 3 + 2];
#pragma unroll
    for (int kh = 0; kh < 3; kh++) {
      const int hh = h + kh - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; kw++) {
        const int ww = w + kw - 1;
        if (ww < 0 || ww >= W) continue;
        const float* un = u + (m + (long long)(kh - 1) * W + (kw - 1)) * 3;
        f += dw[kh * 3 + kw] * un[0] + dw[9 + kh * 3 + kw] * un[1] + dw[18 + kh * 3 + kw] * un[2];
      }
    }
    const float fg = gelu_erf(f);
    save[m * 2] = f; save[m * 2 + 1] = fg;
    s1 += fg; s2 = fmaf(fg, fg, s2);
  }
  s1 = block_sum_256(s1, red);
  s2 = block_sum_256(s2, red);
  if (threadIdx.x == 0) { ws[blockIdx.x * 2] = s1; ws[blockIdx.x * 2 + 1] = s2; }
}

__global__ void srm_fwd_b_kernel(const float* __restrict__ ws, int nblk, long long M, float* rmean, float* rvar, long long* nbt,
                                 float momentum, float eps, float* st, int frozen) {
  if (threadIdx.x != 0) return;
  if (frozen) {                               // eval-mode BatchNorm2d(1): running statistics, no update
    st[0] = rmean[0];
    st[1] = rsqrtf(rvar[0] + eps);
    return;
  }
  double a = 0.0, b = 0.0;
  for (int i = 0; i < nblk; i++) { a += ws[i * 2]; b += ws[i * 2 + 1]; }
  const double mu = a / (double)M;
  double var = b / (double)M - mu * mu;
  if (var < 0.0) var = 0.0;
  st[0] = (float)mu;
  st[1] = (float)(1.0 / sqrt(var + (double)eps));
  if (rmean) rmean[0] = (1.f - momentum) * rmean[0] + momentum * (float)mu;
  if (rvar) rvar[0] = (1.f - momentum) * rvar[0] + momentum * (float)(M > 1 ? var * (double)M / (double)(M - 1) : var);
  if (nbt) *nbt += 1;
}

__global__ void __launch_bounds__(256) srm_fwd_c_kernel(const float* __restrict__ save, const float* __restrict__ st, const float* gamma,
                                                        const float* beta, float* __restrict__ gm, long long M) {
  const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
  if (m >= M) return;
  gm[m] = sigmoidf_(fmaf(gamma[0], (save[m * 2 + 1] - st[0]) * st[1], beta[0]));
}

// backward stage a: partial sums of dfn = dgm * gm (1 - gm) and dfn * xhat
__global__ void __launch_bounds__(256) srm_bwd_a_kernel(const float* __restrict__ dgm, const float* __restrict__ gm, const float* __restrict__ save,
                                                        const float* __restrict__ st, float* __restrict__ ws, long long M) {
  __shared__ float red[8];
  float s1 = 0.f, s2 = 0.f;
  for (long long m = (long long)blockIdx.x * 256 + threadIdx.x; m < M; m += (long long)gridDim.x * 256) {
    const float g = gm[m];
    const float d = dgm[m] * g * (1.f - g);
    s1 += d;
    s2 = fmaf(d, (save[m * 2 + 1] - st[0]) * st[1], s2);
  }
  s1 = block_sum_256(s1, red);
  s2 = block_sum_256(s2, red);
  if (threadIdx.x == 0) { ws[blockIdx.x * 2] = s1; ws[blockIdx.x * 2 + 1] = s2; }
}
__global__ void srm_bwd_b_kernel(const float* __restrict__ ws, int nblk, float* sums, float* dgamma, float* dbeta, int frozen) {
  if (threadIdx.x != 0) return;
  float a = 0.f, b = 0.f;
  for (int i = 0; i < nblk; i++) { a += ws[i * 2]; b += ws[i * 2 + 1]; }
  sums[0] = frozen ? 0.f : a; sums[1] = frozen ? 0.f : b;
  dbeta[0] = a; dgamma[0] = b;
}
// stage c: df_pre[m] -> save[m*2+1] (fg is dead from here on)
__global__ void __launch_bounds__(256) srm_bwd_c_kernel(const float* __restrict__ dgm, const float* __restrict__ gm, float* __restrict__ save,
                                                        const float* __restrict__ st, const float* __restrict__ sums, const float* gamma,
                                                        long long M) {
  const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
  if (m >= M) return;
  const float g = gm[m];
  const float d = dgm[m] * g * (1.f - g);
  const float xh = (save[m * 2 + 1] - st[0]) * st[1];
  const float dfg = gamma[0] * st[1] * (d - sums[0] / (float)M - xh * sums[1] / (float)M);
  save[m * 2 + 1] = dfg * gelu_grad(save[m * 2]);
}
// stage d: du = pwc^T(df) + dwc^T(df); block partials of the 3 + 27 filter gradients -> ws[blk*30 + i]
__global__ void __launch_bounds__(256) srm_bwd_d_kernel(const float* __restrict__ u, const float* __restrict__ save, const float* __restrict__ pw,
                                                        const float* __restrict__ dw, float* __restrict__ du, float* __restrict__ ws, int B,
                                                        int H, int W) {
  __shared__ float red[8];
  const long long M = (long long)B * H * W;
  float gp[3] = {0.f, 0.f, 0.f}, gd[27];
#pragma unroll
  for (int i = 0; i < 27; i++) gd[i] = 0.f;
  for (long long m = (long long)blockIdx.x * 256 + threadIdx.x; m < M; m += (long long)gridDim.x * 256) {
    const int w = (int)(m % W);
    const long long t = m / W;
    const int h = (int)(t % H);
    const float df = save[m * 2 + 1];
    float d0 = pw[0] * df, d1 = pw[1] * df, d2 = pw[2] * df;
#pragma unroll
    for (int k = 0; k < 3; k++) gp[k] = fmaf(df, u[m * 3 + k], gp[k]);
#pragma unroll
    for (int kh = 0; kh < 3; kh++) {
#pragma unroll
      for (int kw = 0; kw < 3; kw++) {
        const int hh = h + kh - 1, ww = w + kw - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
          // filter gradient: output m reads input mn through tap (kh, kw)
          const long long mn = m + (long long)(kh - 1) * W + (kw - 1);
#pragma unroll
          for (int k = 0; k < 3; k++) gd[k * 9 + kh * 3 + kw] = fmaf(df, u[mn * 3 + k], gd[k * 9 + kh * 3 + kw]);
        }
        // data gradient: input m is read by output mo = m - (kh-1, kw-1) through tap (kh, kw)
        const int ho = h - (kh - 1), wo = w - (kw - 1);
        if (ho >= 0 && ho < H && wo >= 0 && wo < W) {
          const float dfo = save[(m - (long long)(kh - 1) * W - (kw - 1)) * 2 + 1];
          d0 = fmaf(dw[kh * 3 + kw], dfo, d0);
          d1 = fmaf(dw[9 + kh * 3 + kw], dfo, d1);
          d2 = fmaf(dw[18 + kh * 3 + kw], dfo, d2);
        }
      }
    }
    du[m * 3] = d0; du[m * 3 + 1] = d1; du[m * 3 + 2] = d2;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float s = block_sum_256(gp[k], red);
    if (threadIdx.x == 0) ws[blockIdx.x * 30 + k] = s;
  }
#pragma unroll
  for (int k = 0; k < 27; k++) {
    const float s = block_sum_256(gd[k], red);
    if (threadIdx.x == 0) ws[blockIdx.x * 30 + 3 + k] = s;
  }
}

// d(z) = (dh3 * gm + du_mean / C + du_std * (x - mean) / ((C-1) std) + [c == arg] du_max) * gelu'(z)
template <typename T, int V>
__global__ void __launch_bounds__(256) srm_apply_bwd_kernel(const T* __restrict__ dh3, const T* __restrict__ h2, const T* __restrict__ z,
                                                            const float* __restrict__ gm, const float* __restrict__ u,
                                                            const int* __restrict__ arg, const float* __restrict__ du, T* __restrict__ dz,
                                                            long long rows, int C) {
  const int groups = C / V;
  const long long total = rows * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int c0 = (int)(i % groups) * V;
    float d[V], x[V], zz[V], o[V];
    ldv<V>(dh3 + r * C + c0, d);
    ldv<V>(h2 + r * C + c0, x);
    ldv<V>(z + r * C + c0, zz);
    const float g = gm[r], mean = u[r * 3 + 1], sd = u[r * 3 + 2];
    const float k1 = du[r * 3 + 1] / C, k2 = sd > 0.f ? du[r * 3 + 2] / ((C - 1) * sd) : 0.f, k0 = du[r * 3];
    const int am = arg[r];
#pragma unroll
    for (int v = 0; v < V; v++) {
      float t = fmaf(d[v], g, k1) + k2 * (x[v] - mean);
      if (c0 + v == am) t += k0;
      o[v] = t * gelu_grad(zz[v]);
    }
    stv<V>(dz + r * C + c0, o);
  }
}

// ---------------------------------------------------------------------------------------------- elementwise
template <typename T, int V>
__global__ void __launch_bounds__(256) silu_mul_fwd_kernel(const T* __restrict__ g, const T* __restrict__ v, T* __restrict__ out, long long n) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V; i < n; i += (long long)gridDim.x * blockDim.x * V) {
    float a[V], b[V], o[V];
    ldv<V>(g + i, a);
    ldv<V>(v + i, b);
#pragma unroll
    for (int j = 0; j < V; j++) o[j] = a[j] * sigmoidf_(a[j]) * b[j] * sigmoidf_(b[j]);
    stv<V>(out + i, o);
  }
}
__device__ __forceinline__ float silu_grad(float x) {
  const float s = sigmoidf_(x);
  return s * (1.f + x * (1.f - s));
}
template <typename T, int V>
__global__ void __launch_bounds__(256) silu_mul_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ g, const T* __restrict__ v,
                                                           T* __restrict__ dg, T* __restrict__ dv, long long n) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V; i < n; i += (long long)gridDim.x * blockDim.x * V) {
    float d[V], a[V], b[V], oa[V], ob[V];
    ldv<V>(dout + i, d);
    ldv<V>(g + i, a);
    ldv<V>(v + i, b);
#pragma unroll
    for (int j = 0; j < V; j++) {
      oa[j] = d[j] * silu_grad(a[j]) * b[j] * sigmoidf_(b[j]);
      ob[j] = d[j] * silu_grad(b[j]) * a[j] * sigmoidf_(a[j]);
    }
    stv<V>(dg + i, oa);
    stv<V>(dv + i, ob);
  }
}

// out = x + ls[c] * inner,  inner = y ? (1-w) y + w pz : pz,  pz = s ? p*s[c]+t[c] : p
template <typename T, int V>
__global__ void __launch_bounds__(256) ls_combine_fwd_kernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ p,
                                                             const float* __restrict__ s, const float* __restrict__ t,
                                                             const float* __restrict__ ls, const float* __restrict__ wp, T* __restrict__ out,
                                                             long long rows, int C) {
  const int groups = C / V;
  const long long total = rows * groups;
  const float w = wp ? wp[0] : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * V;
    const int c0 = (int)(i % groups) * V;
    float xv[V], pv[V], o[V];
    ldv<V>(x + e, xv);
    ldv<V>(p + e, pv);
    if (s) {
#pragma unroll
      for (int j = 0; j < V; j++) pv[j] = fmaf(pv[j], s[c0 + j], t[c0 + j]);
    }
    if (y) {
      float yv[V];
      ldv<V>(y + e, yv);
#pragma unroll
      for (int j = 0; j < V; j++) pv[j] = (1.f - w) * yv[j] + w * pv[j];
    }
#pragma unroll
    for (int j = 0; j < V; j++) o[j] = fmaf(ls[c0 + j], pv[j], xv[j]);
    stv<V>(out + e, o);
  }
}

// backward: d2 = ls*dout; dy (+)= (1-w) d2; dp = w d2 (gradient w.r.t. pz); partials: q0 = dout*inner (-> dls), q1 = d2*(pz-y) (-> dw)
template <typename T, int V>
__global__ void __launch_bounds__(kColThreads) ls_combine_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ y, const T* __restrict__ p,
                                                                     const float* __restrict__ s, const float* __restrict__ t,
                                                                     const float* __restrict__ ls, const float* __restrict__ wp,
                                                                     T* __restrict__ dy, int acc_dy, T* __restrict__ dp, long long rows,
                                                                     int C, int ngrp, int nrl, int rows_per_block, float* __restrict__ ws) {
  __shared__ float smem[V * kColThreads];
  const int grp = threadIdx.x % ngrp, rl = threadIdx.x / ngrp;
  const int c0 = (blockIdx.y * ngrp + grp) * V;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  const float w = wp ? wp[0] : 1.f;
  float q0[V], q1[V], lsv[V], sv[V], tv[V];
#pragma unroll
  for (int j = 0; j < V; j++) {
    q0[j] = q1[j] = 0.f;
    const bool ok = c0 + j < C;
    lsv[j] = ok ? ls[c0 + j] : 0.f;
    sv[j] = (ok && s) ? s[c0 + j] : 1.f;
    tv[j] = (ok && s) ? t[c0 + j] : 0.f;
  }
  if (c0 < C) {
    for (long long r = r0 + rl; r < r1; r += nrl) {
      const long long e = r * C + c0;
      float d[V], pv[V], yv[V], o[V];
      ldv<V>(dout + e, d);
      ldv<V>(p + e, pv);
#pragma unroll
      for (int j = 0; j < V; j++) { pv[j] = fmaf(pv[j], sv[j], tv[j]); yv[j] = 0.f; }
      if (y) ldv<V>(y + e, yv);
#pragma unroll
      for (int j = 0; j < V; j++) {
        const float inner = y ? (1.f - w) * yv[j] + w * pv[j] : pv[j];
        const float d2 = lsv[j] * d[j];
        q0[j] = fmaf(d[j], inner, q0[j]);
        q1[j] = fmaf(d2, pv[j] - yv[j], q1[j]);
        o[j] = (y ? w : 1.f) * d2;
        yv[j] = (1.f - w) * d2;
      }
      stv<V>(dp + e, o);
      if (y) {
        if (acc_dy) {
          float old[V];
          ldv<V>(dy + e, old);
#pragma unroll
          for (int j = 0; j < V; j++) yv[j] += old[j];
        }
        stv<V>(dy + e, yv);
      }
    }
  }
  col_block_reduce<V>(q0, smem, grp, rl, ngrp, nrl);
  col_block_reduce<V>(q1, smem, grp, rl, ngrp, nrl);
  if (rl == 0 && c0 < C) {
#pragma unroll
    for (int j = 0; j < V; j++)
      if (c0 + j < C) {
        ws[((size_t)blockIdx.x * 2 + 0) * C + c0 + j] = q0[j];
        ws[((size_t)blockIdx.x * 2 + 1) * C + c0 + j] = q1[j];
      }
  }
}
// dw = sum_c percol[c]   (one block; percol = the per-channel sums produced by launch_finalize)
__global__ void __launch_bounds__(256) ls_combine_finalize_kernel(const float* __restrict__ percol, int C, float* dw) {
  __shared__ float red[8];
  float tw = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) tw += percol[c];
  tw = block_sum_256(tw, red);
  if (threadIdx.x == 0) dw[0] = tw;
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
static inline int ew_blocks(long long total) { return (int)std::min<long long>((total + 255) / 256, 8LL * kNumSMs); }

extern "C" int cenet_ccu_stats(const void* xb, int dtype, float* u, int* arg, int B, int HW, int C, float* ws, long long ws_elems,
                               cenet_stream_t st) {
  CENET_REQUIRE(xb && u && arg, "cenet_ccu_stats: null pointer");
  if (B == 0) return 0;
  CENET_DISPATCH(dtype, T, {
    const int Vv = vec_of(sizeof(T), {xb}, {C});
    ColPlan p = plan_cols(HW, C, Vv);
    DISPATCH_V(Vv, (ccu_stats_kernel<T, V><<<dim3(p.gy, B), kColThreads, 0, to_stream(st)>>>((const T*)xb, u, arg, HW, C, p.ngrp, p.nrl)));
    CENET_LAUNCH_CHECK("ccu_stats");
  });
  return 0;
}
extern "C" int cenet_ccu_dgate(const void* dx1, const void* xb, int dtype, float* dgate, int B, int HW, int C, float* ws,
                               long long ws_elems, cenet_stream_t st) {
  CENET_REQUIRE(dx1 && xb && dgate, "cenet_ccu_dgate: null pointer");
  if (B == 0) return 0;
  CENET_DISPATCH(dtype, T, {
    const int Vv = vec_of(sizeof(T), {dx1, xb}, {C});
    ColPlan p = plan_cols(HW, C, Vv);
    DISPATCH_V(Vv, (ccu_dgate_kernel<T, V><<<dim3(p.gy, B), kColThreads, 0, to_stream(st)>>>((const T*)dx1, (const T*)xb, dgate, HW, C,
                                                                                        p.ngrp, p.nrl)));
    CENET_LAUNCH_CHECK("ccu_dgate");
  });
  return 0;
}
extern "C" int cenet_ccu_mlp_fwd(const float* u, const float* fc1, const float* fc2, const float* gamma, const float* beta, float* rmean,
                                 float* rvar, long long* nbt, float momentum, float eps, float* gate, float* save, int B, int C,
                                 int frozen, cenet_stream_t st) {
  CENET_REQUIRE(u && fc1 && fc2 && gate && save, "cenet_ccu_mlp_fwd: null pointer");
  CENET_REQUIRE(!(frozen && gamma) || (rmean && rvar), "cenet_ccu_mlp_fwd(frozen): running statistics required");
  CENET_REQUIRE((gamma == nullptr) == (beta == nullptr), "cenet_ccu_mlp_fwd: gamma and beta come together");
  ccu_mlp_fwd_kernel<<<cdiv(C, 64), 64, 0, to_stream(st)>>>(u, fc1, fc2, gamma, beta, rmean, rvar, nbt, momentum, eps, gate, save, B, C, frozen);
  CENET_LAUNCH_CHECK("ccu_mlp_fwd");
  return 0;
}
extern "C" int cenet_ccu_mlp_bwd(const float* dgate, const float* u, const float* fc1, const float* fc2, const float* gamma,
                                 const float* beta, const float* save, float* du, float* dfc1, float* dfc2, float* dgamma, float* dbeta,
                                 int B, int C, int frozen, cenet_stream_t st) {
  CENET_REQUIRE(dgate && u && fc1 && fc2 && save && du && dfc1 && dfc2, "cenet_ccu_mlp_bwd: null pointer");
  ccu_mlp_bwd_kernel<<<cdiv(C, 64), 64, 0, to_stream(st)>>>(dgate, u, fc1, fc2, gamma, beta, save, du, dfc1, dfc2, dgamma, dbeta, B, C, frozen);
  CENET_LAUNCH_CHECK("ccu_mlp_bwd");
  return 0;
}
extern "C" int cenet_ccu_apply_bwd(const void* dx1, const void* xb, int dtype, const float* gate, const float* u, const int* arg,
                                   const float* du, void* dxb, int acc, int B, int HW, int C, cenet_stream_t st) {
  CENET_REQUIRE(dx1 && xb && gate && u && arg && du && dxb, "cenet_ccu_apply_bwd: null pointer");
  if (B == 0) return 0;
  CENET_DISPATCH(dtype, T, {
    const int Vv = vec_of(sizeof(T), {dx1, xb, dxb}, {C});
    DISPATCH_V(Vv, (ccu_apply_bwd_kernel<T, V><<<ew_blocks((long long)B * HW * (C / Vv)), 256, 0, to_stream(st)>>>(
                        (const T*)dx1, (const T*)xb, gate, u, arg, du, (T*)dxb, acc, B, HW, C)));
    CENET_LAUNCH_CHECK("ccu_apply_bwd");
  });
  return 0;
}

extern "C" int cenet_row_stats_arg(const void* x, int dtype, float* u, int* arg, long long rows, int C, cenet_stream_t st) {
  CENET_REQUIRE(x && u && arg && C % 2 == 0 && C >= 2, "cenet_row_stats_arg: bad arguments");
  if (rows == 0) return 0;
  CENET_DISPATCH(dtype, T, (row_stats_arg_kernel<T><<<cdiv(rows, 8), 256, 0, to_stream(st)>>>((const T*)x, u, arg, rows, C)));
  CENET_LAUNCH_CHECK("row_stats_arg");
  return 0;
}
extern "C" int cenet_row_dot(const void* a, const void* b, int dtype, float* out, long long rows, int C, cenet_stream_t st) {
  CENET_REQUIRE(a && b && out && C % 2 == 0, "cenet_row_dot: bad arguments");
  if (rows == 0) return 0;
  CENET_DISPATCH(dtype, T, (row_dot_kernel<T><<<cdiv(rows, 8), 256, 0, to_stream(st)>>>((const T*)a, (const T*)b, out, rows, C)));
  CENET_LAUNCH_CHECK("row_dot");
  return 0;
}

extern "C" int cenet_srm_fwd(const float* u, const float* pw, const float* dw, const float* gamma, const float* beta, float* rmean,
                             float* rvar, long long* nbt, float momentum, float eps, float* gm, float* save, float* stt, int B, int H,
                             int W, int frozen, float* ws, long long ws_elems, cenet_stream_t st) {
  CENET_REQUIRE(u && pw && dw && gamma && beta && gm && save && stt && ws, "cenet_srm_fwd: null pointer");
  cudaStream_t s = to_stream(st);
  const long long M = (long long)B * H * W;
  const int nblk = (int)std::min<long long>((M + 255) / 256, 2LL * kNumSMs);
  CENET_REQUIRE(2LL * nblk <= ws_elems, "cenet_srm_fwd: workspace too small");
  srm_fwd_a_kernel<<<nblk, 256, 0, s>>>(u, pw, dw, save, ws, B, H, W);
  CENET_LAUNCH_CHECK("srm_fwd_a");
  srm_fwd_b_kernel<<<1, 32, 0, s>>>(ws, nblk, M, rmean, rvar, nbt, momentum, eps, stt, frozen);
  CENET_LAUNCH_CHECK("srm_fwd_b");
  srm_fwd_c_kernel<<<cdiv(M, 256), 256, 0, s>>>(save, stt, gamma, beta, gm, M);
  CENET_LAUNCH_CHECK("srm_fwd_c");
  return 0;
}
extern "C" int cenet_srm_bwd(const float* dgm, const float* u, const float* gm, float* save, const float* stt, const float* pw,
                             const float* dw, const float* gamma, const float* beta, float* du, float* dpw, float* ddw, float* dgamma,
                             float* dbeta, int B, int H, int W, int frozen, float* ws, long long ws_elems, cenet_stream_t st) {
  CENET_REQUIRE(dgm && u && gm && save && stt && pw && dw && gamma && du && dpw && ddw && dgamma && dbeta && ws, "cenet_srm_bwd: null pointer");
  cudaStream_t s = to_stream(st);
  const long long M = (long long)B * H * W;
  const int nblk = (int)std::min<long long>((M + 255) / 256, 2LL * kNumSMs);
  CENET_REQUIRE(30LL * nblk + 2 <= ws_elems, "cenet_srm_bwd: workspace too small");
  float* sums = ws + 30LL * nblk;
  srm_bwd_a_kernel<<<nblk, 256, 0, s>>>(dgm, gm, save, stt, ws, M);
  CENET_LAUNCH_CHECK("srm_bwd_a");
  srm_bwd_b_kernel<<<1, 32, 0, s>>>(ws, nblk, sums, dgamma, dbeta, frozen);
  CENET_LAUNCH_CHECK("srm_bwd_b");
  srm_bwd_c_kernel<<<cdiv(M, 256), 256, 0, s>>>(dgm, gm, save, stt, sums, gamma, M);
  CENET_LAUNCH_CHECK("srm_bwd_c");
  srm_bwd_d_kernel<<<nblk, 256, 0, s>>>(u, save, pw, dw, du, ws, B, H, W);
  CENET_LAUNCH_CHECK("srm_bwd_d");
  return launch_finalize(ws, nblk, 30, dpw, 3, ddw, 1.f, s);
}
extern "C" int cenet_srm_apply_bwd(const void* dh3, const void* h2, const void* z, int dtype, const float* gm, const float* u,
                                   const int* arg, const float* du, void* dz, long long rows, int C, cenet_stream_t st) {
  CENET_REQUIRE(dh3 && h2 && z && gm && u && arg && du && dz, "cenet_srm_apply_bwd: null pointer");
  if (rows == 0) return 0;
  CENET_DISPATCH(dtype, T, {
    const int Vv = vec_of(sizeof(T), {dh3, h2, z, dz}, {C});
    DISPATCH_V(Vv, (srm_apply_bwd_kernel<T, V><<<ew_blocks(rows * (C / Vv)), 256, 0, to_stream(st)>>>(
                        (const T*)dh3, (const T*)h2, (const T*)z, gm, u, arg, du, (T*)dz, rows, C)));
    CENET_LAUNCH_CHECK("srm_apply_bwd");
  });
  return 0;
}

extern "C" int cenet_silu_mul_fwd(const void* g, const void* v, void* out, int dtype, long long n, cenet_stream_t st) {
  CENET_REQUIRE(g && v && out, "cenet_silu_mul_fwd: null pointer");
  if (n == 0) return 0;
  CENET_DISPATCH(dtype, T, {
    const int Vv = vec_of(sizeof(T), {g, v, out}, {n});
    DISPATCH_V(Vv, (silu_mul_fwd_kernel<T, V><<<ew_blocks(n / Vv), 256, 0, to_stream(st)>>>((const T*)g, (const T*)v, (T*)out, n)));
    CENET_LAUNCH_CHECK("silu_mul_fwd");
  });
  return 0;
}
extern "C" int cenet_silu_mul_bwd(const void* dout, const void* g, const void* v, void* dg, void* dv, int dtype, long long n,
                                  cenet_stream_t st) {
  CENET_REQUIRE(dout && g && v && dg && dv, "cenet_silu_mul_bwd: null pointer");
  if (n == 0) return 0;
  CENET_DISPATCH(dtype, T, {
    const int Vv = vec_of(sizeof(T), {dout, g, v, dg, dv}, {n});
    DISPATCH_V(Vv, (silu_mul_bwd_kernel<T, V><<<ew_blocks(n / Vv), 256, 0, to_stream(st)>>>((const T*)dout, (const T*)g, (const T*)v,
                                                                                        (T*)dg, (T*)dv, n)));
    CENET_LAUNCH_CHECK("silu_mul_bwd");
  });
  return 0;
}

extern "C" int cenet_ls_combine_fwd(const void* x, const void* y, const void* p, int dtype, const float* s, const float* t, const float* ls,
                                    const float* w, void* out, long long rows, int C, cenet_stream_t st) {
  CENET_REQUIRE(x && p && ls && out, "cenet_ls_combine_fwd: null pointer");
  CENET_REQUIRE((y == nullptr) == (w == nullptr) && (s == nullptr) == (t == nullptr), "cenet_ls_combine_fwd: operand pairs");
  if (rows == 0) return 0;
  CENET_DISPATCH(dtype, T, {
    const int Vv = vec_of(sizeof(T), {x, y, p, out}, {C});
    DISPATCH_V(Vv, (ls_combine_fwd_kernel<T, V><<<ew_blocks(rows * (C / Vv)), 256, 0, to_stream(st)>>>(
                        (const T*)x, (const T*)y, (const T*)p, s, t, ls, w, (T*)out, rows, C)));
    CENET_LAUNCH_CHECK("ls_combine_fwd");
  });
  return 0;
}
extern "C" int cenet_ls_combine_bwd(const void* dout, const void* y, const void* p, int dtype, const float* s, const float* t,
                                    const float* ls, const float* w, void* dy, int acc_dy, void* dp, float* dls, float* dw, long long rows,
                                    int C, float* ws, long long ws_elems, cenet_stream_t st) {
  CENET_REQUIRE(dout && p && ls && dp && dls && ws, "cenet_ls_combine_bwd: null pointer");
  CENET_REQUIRE((y == nullptr) == (w == nullptr) && (y == nullptr) == (dy == nullptr), "cenet_ls_combine_bwd: operand pairs");
  cudaStream_t sm = to_stream(st);
  CENET_DISPATCH(dtype, T, {
    int Vv = vec_of(sizeof(T), {dout, y, p, dy, dp}, {C});
    if (Vv > 4) Vv = 4;
    ColPlan pl = plan_cols(rows, C, Vv);
    CENET_REQUIRE((long long)pl.nrb * 2 * C + C <= ws_elems, "cenet_ls_combine_bwd: workspace too small");
    DISPATCH_V(Vv, (ls_combine_bwd_kernel<T, V><<<dim3(pl.nrb, pl.gy), kColThreads, 0, sm>>>(
                        (const T*)dout, (const T*)y, (const T*)p, s, t, ls, w, (T*)dy, acc_dy, (T*)dp, rows, C, pl.ngrp, pl.nrl,
                        pl.rows_per_block, ws)));
    CENET_LAUNCH_CHECK("ls_combine_bwd");
    // dls[c] = sum_b ws[b][0][c];  dw = sum_c sum_b ws[b][1][c]
    float* percol = ws + (size_t)pl.nrb * 2 * C;
    if (launch_finalize(ws, pl.nrb, 2 * C, dls, C, y ? percol : nullptr, 1.f, sm)) return -1;
    if (y) {
      ls_combine_finalize_kernel<<<1, 256, 0, sm>>>(percol, C, dw);
      CENET_LAUNCH_CHECK("ls_combine_finalize");
    }
  });
  return 0;
}
