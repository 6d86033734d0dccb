// Remaining backward kernels of the training path:
//   fea_bwd              : backward of FEA + gate combine, z = 2y + w*edge(y) + gate*y (dseb.py:63-76,157,118,162)
//   nchw_to_nhwc_slice   : adjoint of the NCHW concat of the DSEB (dseb.py:156)
//   resample             : sparse separable resampling with host-built CSR tap tables -- adaptive average pooling, bilinear
//                          up-sampling and their adjoints (cfam.py:209-218,231-232; blocks.py:210)
//   maxpool2_scale_bwd   : out.py:43,70;  head_upsample_bwd: adjoint of the final bilinear x2 (out.py:74)
//   adamw                : torch.optim.AdamW step over the flat parameter buffer (utils/core.py:16-18)
#include "train_common.cuh"

namespace {
__device__ __forceinline__ float sgn(float x) { return (x > 0.f) - (x < 0.f); }

// one block per (b, c) plane; working set: Y, DZ, T, ACC, R[ns] (each HW floats).  It lives in dynamic shared memory when
// it fits (planes up to 56x56 with 3 scales); larger planes (128x128 at 512x512 inputs) use a per-block slab of the
// workspace instead (`scratch`, L2-resident) and the blocks stride over the planes.
template <typename T>
__global__ void __launch_bounds__(512) fea_bwd_kernel(const T* __restrict__ y, const T* __restrict__ gate, const T* __restrict__ dz,
                                                      const float* __restrict__ wch, T* __restrict__ dy, int acc_dy, T* __restrict__ dgate,
                                                      int C2, int H, int W, const float* __restrict__ mats,
                                                      const int* __restrict__ bands, int nmax, int ns, float* __restrict__ ws,
                                                      int nplanes, float* __restrict__ scratch, int ident_mask, int tab_off) {
  extern __shared__ float sm_dyn[];
  __shared__ float red[16];
  __shared__ int s_nb;
  const int nthr = blockDim.x;                               // 256, or 512 for planes >= 2048 pixels (more threads per plane:
                                                             // the separable passes are latency-bound, shared memory caps the CTAs per SM)
  const int HW = H * W;
  float* sm = scratch ? scratch + (size_t)blockIdx.x * (4 + ns) * HW : sm_dyn;
  float* Y = sm;
  float* DZ = Y + HW;
  float* Tm = DZ + HW;
  float* ACC = Tm + HW;
  float* R = ACC + HW;                 // [ns][HW]
  const int tid = threadIdx.x;
  // ---- compact band tables (shared memory, built once per CTA): the separable operators are banded (bilinear taps), at most
  //      NB = 4 or 8 non-zeros per row / column.  Table u of scale s: u = 0 rows of A_h (forward, vertical), 1 rows of A_w (forward,
  //      horizontal), 2 columns of A_w (backward, horizontal), 3 columns of A_h (backward, vertical); entry i keeps its first tap
  //      index lo (shifted down so that lo + NB stays inside the axis) and NB coefficients.  The passes become fixed-length
  //      unrolled dot products on shared memory (the first version walked [lo, hi) with one __ldg per tap: L1 latency inside every
  //      dependent FMA chain, ~0.5 warp instructions per cycle and SM).  nb == 0 -> a band is wider than 8: generic loops below.
  const int nax = max(H, W);
  float* tcoef = sm_dyn + tab_off;                      // [ns*4][nax][8]
  int* tlo = reinterpret_cast<int*>(tcoef + (size_t)ns * 4 * nax * 8);
  {
    int wmax = 0;
    for (int i = tid; i < ns * 4 * nax; i += nthr) {
      const int s = i / (4 * nax), u = (i / nax) & 3, e = i % nax;
      const int ax = (u == 0 || u == 3) ? 0 : 1, n = ax ? W : H;
      if (((ident_mask >> s) & 1) || e >= n) continue;
      const int* bd = bands + ((((size_t)s * 2 + ax) * 2 + (u >= 2)) * nmax + e) * 2;
      wmax = max(wmax, bd[1] - bd[0]);
    }
    if (tid == 0) s_nb = 0;
    __syncthreads();
    if (wmax > 0) atomicMax(&s_nb, wmax);
    __syncthreads();
    const int wm = s_nb;
    __syncthreads();
    const int nb = (wm <= 4 && H >= 4 && W >= 4) ? 4 : ((wm <= 8 && H >= 8 && W >= 8) ? 8 : 0);
    if (tid == 0) s_nb = nb;
    if (nb) {
      for (int i = tid; i < ns * 4 * nax; i += nthr) {
        const int s = i / (4 * nax), u = (i / nax) & 3, e = i % nax;
        const int ax = (u == 0 || u == 3) ? 0 : 1, n = ax ? W : H;
        float* cf = tcoef + (size_t)i * 8;
        if (((ident_mask >> s) & 1) || e >= n) {
          for (int t = 0; t < 8; t++) cf[t] = 0.f;
          tlo[i] = 0;
          continue;
        }
        const int* bd = bands + ((((size_t)s * 2 + ax) * 2 + (u >= 2)) * nmax + e) * 2;
        const float* A = mats + ((size_t)s * 2 + ax) * nmax * nmax;
        const int lo = min(bd[0], n - nb);
        for (int t = 0; t < 8; t++) {
          const int k = lo + t;
          cf[t] = (t < nb && k >= bd[0] && k < bd[1]) ? (u < 2 ? A[(size_t)e * nmax + k] : A[(size_t)k * nmax + e]) : 0.f;
        }
        tlo[i] = lo;
      }
    }
    __syncthreads();
  }
  const int nb = s_nb;
  const unsigned wmagic = 0xFFFFFFFFu / (unsigned)W + 1u;           // i / W for i < 2^16 (HW < 65536 is checked by the launcher)
  // dot product of table entry e of (s, u) with src[base + t * stride], t < nb
  auto band_dot = [&](int s, int u, int e, const float* src, int base0, int stride) -> float {
    const int ti = (s * 4 + u) * nax + e;
    const float4 c0 = *reinterpret_cast<const float4*>(tcoef + (size_t)ti * 8);
    const float* q = src + base0 + tlo[ti] * stride;
    float a = c0.x * q[0];
    a = fmaf(c0.y, q[stride], a); a = fmaf(c0.z, q[2 * stride], a); a = fmaf(c0.w, q[3 * stride], a);
    if (nb == 8) {
      const float4 c1 = *reinterpret_cast<const float4*>(tcoef + (size_t)ti * 8 + 4);
      a = fmaf(c1.x, q[4 * stride], a); a = fmaf(c1.y, q[5 * stride], a); a = fmaf(c1.z, q[6 * stride], a); a = fmaf(c1.w, q[7 * stride], a);
    }
    return a;
  };
  for (int plane = blockIdx.x; plane < nplanes; plane += gridDim.x) {
  const int c = plane % C2;
  const long long base = (long long)plane * HW;
  __syncthreads();                     // the previous plane of this block is done with the working set
  for (int i = tid; i < HW; i += nthr) { Y[i] = ldf(y + base + i); DZ[i] = ldf(dz + base + i); }
  __syncthreads();
  // forward residuals R_s = Y - A_h Y A_w^T   (scale factor 1.0: the operator is the identity, R_s = 0 exactly and so is its
  // gradient term -- bit `s` of ident_mask -- which removes four of the separable passes)
  for (int s = 0; s < ns; s++) {
    if ((ident_mask >> s) & 1) {
      for (int i = tid; i < HW; i += nthr) R[s * HW + i] = 0.f;
      continue;
    }
    const float* Ah = mats + ((size_t)s * 2 + 0) * nmax * nmax;
    const float* Aw = mats + ((size_t)s * 2 + 1) * nmax * nmax;
    // the operators are banded (bilinear taps): [lo, hi) of the non-zeros of every row / column comes from the host
    const int* rbh = bands + (((size_t)s * 2 + 0) * 2 + 0) * nmax * 2;     // row bands of A_h
    const int* rbw = bands + (((size_t)s * 2 + 1) * 2 + 0) * nmax * 2;     // row bands of A_w
    if (nb) {
      for (int i = tid; i < HW; i += nthr) {
        const int r = (int)__umulhi((unsigned)i, wmagic), w = i - r * W;
        Tm[i] = band_dot(s, 0, r, Y, w, W);
      }
      __syncthreads();
      for (int i = tid; i < HW; i += nthr) {
        const int r = (int)__umulhi((unsigned)i, wmagic), j = i - r * W;
        R[s * HW + i] = Y[i] - band_dot(s, 1, j, Tm, r * W, 1);
      }
      __syncthreads();
      continue;
    }
    for (int i = tid; i < HW; i += nthr) {
      const int r = i / W, w = i % W;
      float a = 0.f;
      for (int h = rbh[2 * r]; h < rbh[2 * r + 1]; h++) a = fmaf(__ldg(Ah + r * nmax + h), Y[h * W + w], a);
      Tm[i] = a;
    }
    __syncthreads();
    for (int i = tid; i < HW; i += nthr) {
      const int r = i / W, j = i % W;
      float a = 0.f;
      for (int w = rbw[2 * j]; w < rbw[2 * j + 1]; w++) a = fmaf(Tm[r * W + w], __ldg(Aw + j * nmax + w), a);
      R[s * HW + i] = Y[i] - a;
    }
    __syncthreads();
  }
  // element-wise: edge, d(w), and dR_s (stored over R_s)
  const int npairs = ns * (ns - 1) / 2;
  const float invm = npairs > 0 ? 1.f / npairs : 0.f;
  const float wc = wch[c];
  float dwp = 0.f;
  for (int i = tid; i < HW; i += nthr) {
    float r[3], e[3], de[3];
#pragma unroll
    for (int s = 0; s < 3; s++) { r[s] = s < ns ? R[s * HW + i] : 0.f; e[s] = fabsf(r[s]); de[s] = 0.f; }
    float edge = 0.f;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = a + 1; b < 3; b++)
        if (b < ns) {
          const float d = e[a] - e[b];
          edge += fabsf(d);
          de[a] += sgn(d);
          de[b] -= sgn(d);
        }
    edge *= invm;
    const float dzv = DZ[i];
    dwp = fmaf(dzv, edge, dwp);
    const float k = wc * dzv * invm;
#pragma unroll
    for (int s = 0; s < 3; s++)
      if (s < ns) R[s * HW + i] = k * de[s] * sgn(r[s]);
    const float g = ldf(gate + base + i);
    ACC[i] = (2.f + g) * dzv;
    stf(dgate + base + i, dzv * Y[i]);
  }
  __syncthreads();
  // d(y) += G_s - A_h^T (G_s A_w)
  for (int s = 0; s < ns; s++) {
    if ((ident_mask >> s) & 1) continue;                      // G_s == 0
    const float* Ah = mats + ((size_t)s * 2 + 0) * nmax * nmax;
    const float* Aw = mats + ((size_t)s * 2 + 1) * nmax * nmax;
    const float* G = R + s * HW;
    const int* cbh = bands + (((size_t)s * 2 + 0) * 2 + 1) * nmax * 2;     // column bands of A_h
    const int* cbw = bands + (((size_t)s * 2 + 1) * 2 + 1) * nmax * 2;     // column bands of A_w
    if (nb) {
      for (int i = tid; i < HW; i += nthr) {
        const int r = (int)__umulhi((unsigned)i, wmagic), w = i - r * W;
        Tm[i] = band_dot(s, 2, w, G, r * W, 1);
      }
      __syncthreads();
      for (int i = tid; i < HW; i += nthr) {
        const int h = (int)__umulhi((unsigned)i, wmagic), w = i - h * W;
        ACC[i] += G[i] - band_dot(s, 3, h, Tm, w, W);
      }
      __syncthreads();
      continue;
    }
    for (int i = tid; i < HW; i += nthr) {
      const int r = i / W, w = i % W;
      float a = 0.f;
      for (int j = cbw[2 * w]; j < cbw[2 * w + 1]; j++) a = fmaf(G[r * W + j], __ldg(Aw + j * nmax + w), a);
      Tm[i] = a;
    }
    __syncthreads();
    for (int i = tid; i < HW; i += nthr) {
      const int h = i / W, w = i % W;
      float a = 0.f;
      for (int r = cbh[2 * h]; r < cbh[2 * h + 1]; r++) a = fmaf(__ldg(Ah + r * nmax + h), Tm[r * W + w], a);
      ACC[i] += G[i] - a;
    }
    __syncthreads();
  }
  for (int i = tid; i < HW; i += nthr) {
    float v = ACC[i];
    if (acc_dy) v += ldf(dy + base + i);
    stf(dy + base + i, v);
  }
  dwp = warp_sum(dwp);
  if ((tid & 31) == 0) red[tid >> 5] = dwp;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < (nthr >> 5); w++) t += red[w];
    ws[plane] = t;
  }
  }
}
// dw[c] = sum_b ws[b*C2 + c]
__global__ void fea_dw_finalize_kernel(const float* __restrict__ ws, int B, int C2, float* dw) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C2) return;
  float s = 0.f;
  for (int b = 0; b < B; b++) s += ws[b * C2 + c];
  dw[c] = s;
}

// out[b, hw, c] (+)= x[b, coff + c, hw]  (32 x 32 smem tiles)
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_slice_kernel(const T* __restrict__ x, T* __restrict__ out, int HW, int C, int Ctot,
                                                                 int coff, int acc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, hw = hw0 + tx;
    tile[k][tx] = (c < C && hw < HW) ? ldf(x + ((long long)b * Ctot + coff + c) * HW + hw) : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int hw = hw0 + k, c = c0 + tx;
    if (hw < HW && c < C) {
      T* p = out + ((long long)b * HW + hw) * C + c;
      float v = tile[tx][k];
      if (acc) v += ldf(p);
      stf(p, v);
    }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) add_kernel(T* __restrict__ dst, const T* __restrict__ src, long long n, int acc) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V; i < n; i += (long long)gridDim.x * blockDim.x * V) {
    float a[V];
    ldv<V>(src + i, a);
    if (acc) {
      float b[V];
      ldv<V>(dst + i, b);
#pragma unroll
      for (int j = 0; j < V; j++) a[j] += b[j];
    }
    stv<V>(dst + i, a);
  }
}

// out[m, c] = x[m, c] * rs[m]   (the per-pixel SRM gate applied to d(fc2 output), so that its weight gradient is a plain GEMM)
template <typename T, int V>
__global__ void __launch_bounds__(256) row_scale_kernel(const T* __restrict__ x, const float* __restrict__ rs, T* __restrict__ out,
                                                        long long rows, int C) {
  const int groups = C / V;
  const long long total = rows * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int c0 = (int)(i % groups) * V;
    float a[V];
    ldv<V>(x + r * C + c0, a);
    const float sc = rs[r];
#pragma unroll
    for (int j = 0; j < V; j++) a[j] *= sc;
    stv<V>(out + r * C + c0, a);
  }
}

// dx[m, 0:N) (+)= sum_{k<K} dy[m, k] * w[k, n]    with a tiny contraction (K <= 16: the class dimension of the logits): every
// thread owns 8 output columns of one row; w lives in shared memory.  Replaces a CUDA-core GEMM launch whose tiles did K = 4.
template <typename T>
__global__ void __launch_bounds__(256) smallk_dgrad_kernel(const float* __restrict__ dy, int K, const float* __restrict__ w, long long ldw,
                                                           T* __restrict__ dx, long long ldx, long long rows, int N, int acc) {
  extern __shared__ float sw[];                 // [K][N]
  for (int i = threadIdx.x; i < K * N; i += blockDim.x) sw[i] = w[(long long)(i / N) * ldw + (i % N)];
  __syncthreads();
  const int groups = N / 8;
  const long long total = rows * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups;
    const int c0 = (int)(i % groups) * 8;
    float o[8];
    if (acc) ldv<8>(dx + r * ldx + c0, o);
    else {
#pragma unroll
      for (int j = 0; j < 8; j++) o[j] = 0.f;
    }
    for (int k = 0; k < K; k++) {
      const float g = dy[r * K + k];
#pragma unroll
      for (int j = 0; j < 8; j++) o[j] = fmaf(g, sw[k * N + c0 + j], o[j]);
    }
    stv<8>(dx + r * ldx + c0, o);
  }
}

// y[b,i,j,c] (+)= sum_{a in taps_h(i)} sum_{e in taps_w(j)} wh[a] ww[e] x[b, hi[a], wi[e], c]
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256) resample_kernel(const TI* __restrict__ x, long long ldx, TO* __restrict__ y, long long ldy, int B,
                                                       int Hi, int Wi, int Ho, int Wo, int C, const int* __restrict__ hs,
                                                       const int* __restrict__ hi, const float* __restrict__ hw, const int* __restrict__ wsx,
                                                       const int* __restrict__ wi, const float* __restrict__ ww, int acc) {
  const int groups = C / V;
  const long long total = (long long)B * Ho * Wo * groups;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(t % groups) * V;
    long long p = t / groups;
    const int j = (int)(p % Wo); p /= Wo;
    const int i = (int)(p % Ho);
    const int b = (int)(p / Ho);
    float o[V];
#pragma unroll
    for (int v = 0; v < V; v++) o[v] = 0.f;
    for (int a = hs[i]; a < hs[i + 1]; a++) {
      const float wa = hw[a];
      const TI* xr = x + ((long long)b * Hi + hi[a]) * Wi * ldx + c0;
      for (int e = wsx[j]; e < wsx[j + 1]; e++) {
        const float wt = wa * ww[e];
        float xv[V];
        ldv<V>(xr + (long long)wi[e] * ldx, xv);
#pragma unroll
        for (int v = 0; v < V; v++) o[v] = fmaf(wt, xv[v], o[v]);
      }
    }
    TO* yp = y + (((long long)b * Ho + i) * Wo + j) * ldy + c0;
    if (acc) {
      float old[V];
      ldv<V>(yp, old);
#pragma unroll
      for (int v = 0; v < V; v++) o[v] += old[v];
    }
    stv<V>(yp, o);
  }
}

// drb[b,h,w,c] = w[c] * dz[b,h/2,w/2,c] at the (first) arg max of each 2x2 window, 0 elsewhere; partials of d(w)
template <typename T, int V>
__global__ void __launch_bounds__(kColThreads) maxpool2_scale_bwd_kernel(const T* __restrict__ dz, long long lddz, const T* __restrict__ rb,
                                                                         const float* __restrict__ wch, T* __restrict__ drb, int B, int H,
                                                                         int W, int C, int ngrp, int nrl, int rows_per_block,
                                                                         float* __restrict__ ws) {
  __shared__ float smem[V * kColThreads];
  const int grp = threadIdx.x % ngrp, rl = threadIdx.x / ngrp;
  const int c0 = (blockIdx.y * ngrp + grp) * V;
  const int Hp = H / 2, Wp = W / 2;
  const long long rows = (long long)B * Hp * Wp;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float q[V], wv[V];
#pragma unroll
  for (int v = 0; v < V; v++) { q[v] = 0.f; wv[v] = c0 + v < C ? wch[c0 + v] : 0.f; }
  if (c0 < C) {
    for (long long r = r0 + rl; r < r1; r += nrl) {
      const int wp = (int)(r % Wp);
      const long long t = r / Wp;
      const int hp = (int)(t % Hp), b = (int)(t / Hp);
      float g[V], x[4][V];
      ldv<V>(dz + r * lddz + c0, g);
#pragma unroll
      for (int k = 0; k < 4; k++)
        ldv<V>(rb + (((long long)b * H + 2 * hp + (k >> 1)) * W + 2 * wp + (k & 1)) * C + c0, x[k]);
      float o[4][V];
#pragma unroll
      for (int v = 0; v < V; v++) {
        int am = 0;
        float m = x[0][v];
#pragma unroll
        for (int k = 1; k < 4; k++)
          if (x[k][v] > m) { m = x[k][v]; am = k; }
        q[v] = fmaf(g[v], m, q[v]);
#pragma unroll
        for (int k = 0; k < 4; k++) o[k][v] = k == am ? wv[v] * g[v] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; k++)
        stv<V>(drb + (((long long)b * H + 2 * hp + (k >> 1)) * W + 2 * wp + (k & 1)) * C + c0, o[k]);
    }
  }
  col_block_reduce<V>(q, smem, grp, rl, ngrp, nrl);
  if (rl == 0 && c0 < C) {
#pragma unroll
    for (int v = 0; v < V; v++)
      if (c0 + v < C) ws[(size_t)blockIdx.x * C + c0 + v] = q[v];
  }
}

// adjoint of bilinear x2 (align_corners=False): dlogits [B,ncls,2h,2w] fp32 -> dyh [B,h,w,ncls] fp32
__global__ void __launch_bounds__(256) head_upsample_bwd_kernel(const float* __restrict__ dl, float* __restrict__ dyh, int B, int h, int w,
                                                                int ncls) {
  const long long total = (long long)B * h * w * ncls;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % ncls);
  long long p = i / ncls;
  const int x = (int)(p % w); p /= w;
  const int y = (int)(p % h);
  const int b = (int)(p / h);
  const int H2 = 2 * h, W2 = 2 * w;
  float wy[5], wx[5];
#pragma unroll
  for (int d = 0; d < 5; d++) {
    wy[d] = wx[d] = 0.f;
    int i0, i1; float l1;
    const int oy = 2 * y - 2 + d;
    if (oy >= 0 && oy < H2) {
      bilin_src(oy, 0.5f, h, i0, i1, l1);
      if (i0 == y) wy[d] += 1.f - l1;
      if (i1 == y) wy[d] += l1;
    }
    const int ox = 2 * x - 2 + d;
    if (ox >= 0 && ox < W2) {
      bilin_src(ox, 0.5f, w, i0, i1, l1);
      if (i0 == x) wx[d] += 1.f - l1;
      if (i1 == x) wx[d] += l1;
    }
  }
  const float* src = dl + ((long long)b * ncls + k) * H2 * W2;
  float s = 0.f;
#pragma unroll
  for (int dy = 0; dy < 5; dy++) {
    if (wy[dy] == 0.f) continue;
    const int oy = 2 * y - 2 + dy;
#pragma unroll
    for (int dx = 0; dx < 5; dx++) {
      if (wx[dx] == 0.f) continue;
      s = fmaf(wy[dy] * wx[dx], src[(long long)oy * W2 + 2 * x - 2 + dx], s);
    }
  }
  dyh[i] = s;
}

// DropPath masks of one training step (timm drop_path, pvtv2.py:146-147: bernoulli(keep) / keep per sample and per branch):
// out[r, b] = u(seed, counter, r, b) < keep[r] ? 1 / keep[r] : 0 with a counter-based generator (splitmix64 finaliser); the
// device-resident counter advances by one per launch, so every CUDA-graph replay draws fresh masks.  One block.
__global__ void __launch_bounds__(256) droppath_mask_kernel(float* __restrict__ out, const float* __restrict__ keep, int n, int B,
                                                            unsigned long long seed, unsigned long long* counter) {
  const unsigned long long c = *counter;
  for (int i = threadIdx.x; i < n * B; i += blockDim.x) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (c * 0x100000001B3ull + (unsigned long long)i + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = (float)(z >> 40) * (1.0f / 16777216.0f);          // 24 random bits -> [0, 1)
    const float k = keep[i / B];
    out[i] = u < k ? 1.0f / k : 0.0f;
  }
  __syncthreads();
  if (threadIdx.x == 0) *counter = c + 1ull;
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, const float* __restrict__ hyper) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], step = hyper[5];
  const float bc1 = 1.f - powf(b1, step), bc2s = sqrtf(1.f - powf(b2, step));
  const float step_size = lr / bc1;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
    float4 pv = *reinterpret_cast<float4*>(p + i), mv = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
    const float4 gv = *reinterpret_cast<const float4*>(g + i);
    float* pp = &pv.x; float* mm = &mv.x; float* vp = &vv.x; const float* gg = &gv.x;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      pp[j] *= 1.f - lr * wd;
      mm[j] = b1 * mm[j] + (1.f - b1) * gg[j];
      vp[j] = b2 * vp[j] + (1.f - b2) * gg[j] * gg[j];
      pp[j] -= step_size * mm[j] / (sqrtf(vp[j]) / bc2s + eps);
    }
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
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
static inline int ew_blocks(long long total) { return (int)std::min<long long>((total + 255) / 256, 8LL * kNumSMs); }

extern "C" int cenet_fea_bwd(const void* y, const void* gate, const void* dz, int dtype, const float* w, void* dy, int acc_dy,
                             void* dgate, float* dw, int B, int C2, int H, int W, const float* mats, const int* bands, int nmax,
                             int nscales, int ident_mask, float* ws, long long ws_elems, cenet_stream_t st) {
  CENET_REQUIRE(y && gate && dz && w && dy && dgate && dw && mats && bands && ws, "cenet_fea_bwd: null pointer");
  CENET_REQUIRE(nscales >= 1 && nscales <= 3, "cenet_fea_bwd: 1..3 scales");
  CENET_REQUIRE(H <= nmax && W <= nmax, "cenet_fea_bwd: operator matrices smaller than the plane");
  CENET_REQUIRE((long long)B * C2 <= ws_elems, "cenet_fea_bwd: workspace too small");
  CENET_REQUIRE((long long)H * W < 65536, "cenet_fea_bwd: plane too large");
  const size_t slab = (size_t)(4 + nscales) * H * W;            // floats of working set per plane
  const int nax = H > W ? H : W;
  const size_t tab_bytes = (size_t)nscales * 4 * nax * (8 * sizeof(float) + sizeof(int));     // compact band tables
  size_t smem = slab * sizeof(float) + tab_bytes;
  const int nplanes = B * C2;
  int blocks = nplanes;
  int tab_off = (int)slab;
  float* scratch = nullptr;
  if (smem > 200 * 1024) {                                      // big planes: working set in the workspace, after the partials
    const long long room = (ws_elems - nplanes) / (long long)slab;
    CENET_REQUIRE(room >= 1, "cenet_fea_bwd: plane %dx%d needs %zu workspace floats per block", H, W, slab);
    blocks = (int)std::min<long long>(std::min<long long>(nplanes, room), 2LL * kNumSMs);
    scratch = ws + nplanes;
    smem = tab_bytes;
    tab_off = 0;
  } else {
    // persistent blocks striding over the planes: the band tables are built once per block
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(227 * 1024 / (smem + 1024), 2048 / (H * W >= 2048 ? 512 : 256)));
    blocks = std::min(nplanes, per_sm * kNumSMs);
  }
  cudaStream_t s = to_stream(st);
  CENET_DISPATCH(dtype, T, {
    static std::atomic<size_t> configured{0};
    if (smem > 48 * 1024 && configured.load() < smem) {
      cudaFuncSetAttribute(fea_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      configured.store(200 * 1024);
    }
    fea_bwd_kernel<T><<<blocks, H * W >= 2048 ? 512 : 256, smem, s>>>((const T*)y, (const T*)gate, (const T*)dz, w, (T*)dy, acc_dy, (T*)dgate, C2, H, W, mats,
                                                bands, nmax, nscales, ws, nplanes, scratch, ident_mask, tab_off);
    CENET_LAUNCH_CHECK("fea_bwd");
  });
  fea_dw_finalize_kernel<<<cdiv(C2, 128), 128, 0, s>>>(ws, B, C2, dw);
  CENET_LAUNCH_CHECK("fea_dw_finalize");
  return 0;
}

extern "C" int cenet_nchw_to_nhwc_slice(const void* x, int dtype, void* out, int B, int HW, int C, int Ctot, int coff, int acc,
                                        cenet_stream_t st) {
  CENET_REQUIRE(x && out, "cenet_nchw_to_nhwc_slice: null pointer");
  if (B == 0) return 0;
  CENET_REQUIRE(B <= 65535 && cdiv(C, 32) <= 65535, "cenet_nchw_to_nhwc_slice: grid too large");
  dim3 grid(cdiv(HW, 32), cdiv(C, 32), B);
  CENET_DISPATCH(dtype, T, (nchw_to_nhwc_slice_kernel<T><<<grid, 256, 0, to_stream(st)>>>((const T*)x, (T*)out, HW, C, Ctot, coff, acc)));
  CENET_LAUNCH_CHECK("nchw_to_nhwc_slice");
  return 0;
}

extern "C" int cenet_add(void* dst, const void* src, int dtype, long long n, int acc, cenet_stream_t st) {
  CENET_REQUIRE(dst && src, "cenet_add: null pointer");
  if (n == 0) return 0;
  CENET_DISPATCH(dtype, T, {
    const int Vv = vec_of(sizeof(T), {dst, src}, {n});
    DISPATCH_V(Vv, (add_kernel<T, V><<<ew_blocks(n / Vv), 256, 0, to_stream(st)>>>((T*)dst, (const T*)src, n, acc)));
    CENET_LAUNCH_CHECK("add");
  });
  return 0;
}

extern "C" int cenet_droppath_mask(float* out, const float* keep, int n, int B, unsigned long long seed, unsigned long long* counter,
                                   cenet_stream_t st) {
  CENET_REQUIRE(out && keep && counter && n >= 1 && B >= 1, "cenet_droppath_mask: bad arguments");
  droppath_mask_kernel<<<1, 256, 0, to_stream(st)>>>(out, keep, n, B, seed, counter);
  CENET_LAUNCH_CHECK("droppath_mask");
  return 0;
}

extern "C" int cenet_row_scale(const void* x, int dtype, const float* rs, void* out, long long rows, int C, cenet_stream_t st) {
  CENET_REQUIRE(x && rs && out, "cenet_row_scale: null pointer");
  if (rows == 0) return 0;
  CENET_DISPATCH(dtype, T, {
    const int Vv = vec_of(sizeof(T), {x, out}, {C});
    DISPATCH_V(Vv, (row_scale_kernel<T, V><<<ew_blocks(rows * (C / Vv)), 256, 0, to_stream(st)>>>((const T*)x, rs, (T*)out, rows, C)));
    CENET_LAUNCH_CHECK("row_scale");
  });
  return 0;
}

extern "C" int cenet_smallk_dgrad(const float* dy, int K, const float* w, long long ldw, void* dx, int dx_dtype, long long ldx,
                                  long long rows, int N, int acc, cenet_stream_t st) {
  CENET_REQUIRE(dy && w && dx, "cenet_smallk_dgrad: null pointer");
  CENET_REQUIRE(K >= 1 && K <= 16 && N % 8 == 0 && N <= 512 && ldx % 8 == 0 && (((uintptr_t)dx & 15) == 0),
                "cenet_smallk_dgrad: K=%d N=%d ldx=%lld not supported (K <= 16, N %% 8 == 0, 16-byte rows)", K, N, ldx);
  if (rows == 0) return 0;
  CENET_DISPATCH(dx_dtype, T, {
    smallk_dgrad_kernel<T><<<ew_blocks(rows * (N / 8)), 256, (size_t)K * N * sizeof(float), to_stream(st)>>>(dy, K, w, ldw, (T*)dx, ldx,
                                                                                                            rows, N, acc);
    CENET_LAUNCH_CHECK("smallk_dgrad");
  });
  return 0;
}

extern "C" int cenet_resample(const void* x, int x_dtype, long long ldx, void* y, int y_dtype, long long ldy, int B, int Hi, int Wi, int Ho,
                              int Wo, int C, const int* hs, const int* hi, const float* hw, const int* wsx, const int* wi, const float* ww,
                              int acc, cenet_stream_t st) {
  CENET_REQUIRE(x && y && hs && hi && hw && wsx && wi && ww, "cenet_resample: null pointer");
  CENET_REQUIRE(x_dtype == y_dtype, "cenet_resample: x and y must share one dtype");
  if (B == 0) return 0;
  CENET_DISPATCH(x_dtype, T, {
    const int Vv = vec_of(sizeof(T), {x, y}, {C, ldx, ldy});
    DISPATCH_V(Vv, (resample_kernel<T, T, V><<<ew_blocks((long long)B * Ho * Wo * (C / Vv)), 256, 0, to_stream(st)>>>(
                        (const T*)x, ldx, (T*)y, ldy, B, Hi, Wi, Ho, Wo, C, hs, hi, hw, wsx, wi, ww, acc)));
    CENET_LAUNCH_CHECK("resample");
  });
  return 0;
}

extern "C" int cenet_maxpool2_scale_bwd(const void* dz, int dtype, long long lddz, const void* rb, int rb_dtype, const float* w, void* drb,
                                        float* dw, int B, int H, int W, int C, float* ws, long long ws_elems, cenet_stream_t st) {
  CENET_REQUIRE(dz && rb && w && drb && dw && ws, "cenet_maxpool2_scale_bwd: null pointer");
  CENET_REQUIRE(dtype == rb_dtype && H % 2 == 0 && W % 2 == 0, "cenet_maxpool2_scale_bwd: bad arguments");
  cudaStream_t s = to_stream(st);
  const long long rows = (long long)B * (H / 2) * (W / 2);
  CENET_DISPATCH(dtype, T, {
    int Vv = vec_of(sizeof(T), {dz, rb, drb}, {C, lddz});
    if (Vv > 4) Vv = 4;
    ColPlan p = plan_cols(rows, C, Vv);
    CENET_REQUIRE((long long)p.nrb * C <= ws_elems, "cenet_maxpool2_scale_bwd: workspace too small");
    DISPATCH_V(Vv, (maxpool2_scale_bwd_kernel<T, V><<<dim3(p.nrb, p.gy), kColThreads, 0, s>>>(
                        (const T*)dz, lddz, (const T*)rb, w, (T*)drb, B, H, W, C, p.ngrp, p.nrl, p.rows_per_block, ws)));
    CENET_LAUNCH_CHECK("maxpool2_scale_bwd");
    return launch_finalize(ws, p.nrb, C, dw, C, nullptr, 1.f, s);
  });
  return 0;
}

extern "C" int cenet_head_upsample_bwd(const float* dlogits, float* dyh, int B, int h, int w, int ncls, cenet_stream_t st) {
  CENET_REQUIRE(dlogits && dyh, "cenet_head_upsample_bwd: null pointer");
  const long long total = (long long)B * h * w * ncls;
  if (total == 0) return 0;
  head_upsample_bwd_kernel<<<cdiv(total, 256), 256, 0, to_stream(st)>>>(dlogits, dyh, B, h, w, ncls);
  CENET_LAUNCH_CHECK("head_upsample_bwd");
  return 0;
}

// dst[i] = map[i] ? src[map[i] - 1] : 0 -- the per-step re-pack of the master weights into every compute layout at once
template <typename TO>
__global__ void __launch_bounds__(256) gather_cast_kernel(const float* __restrict__ src, const int* __restrict__ map, TO* __restrict__ dst,
                                                          long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int4 m = reinterpret_cast<const int4*>(map)[i];
    float v[4];
    v[0] = m.x ? __ldg(src + m.x - 1) : 0.f;
    v[1] = m.y ? __ldg(src + m.y - 1) : 0.f;
    v[2] = m.z ? __ldg(src + m.z - 1) : 0.f;
    v[3] = m.w ? __ldg(src + m.w - 1) : 0.f;
    stv<4>(dst + i * 4, v);
  }
}

extern "C" int cenet_gather_cast(const float* src, const int* map, void* dst, int dst_dtype, long long n, cenet_stream_t st) {
  CENET_REQUIRE(src && map && dst, "cenet_gather_cast: null pointer");
  CENET_REQUIRE(n % 4 == 0 && ((((uintptr_t)map | (uintptr_t)dst) & 15) == 0), "cenet_gather_cast: n %% 4 == 0 and 16-byte aligned buffers");
  if (n == 0) return 0;
  CENET_DISPATCH(dst_dtype, T, {
    gather_cast_kernel<T><<<ew_blocks(n / 4), 256, 0, to_stream(st)>>>(src, map, (T*)dst, n / 4);
    CENET_LAUNCH_CHECK("gather_cast");
  });
  return 0;
}

extern "C" int cenet_adamw(float* p, const float* g, float* m, float* v, long long n, const float* hyper, cenet_stream_t st) {
  CENET_REQUIRE(p && g && m && v && hyper, "cenet_adamw: null pointer");
  CENET_REQUIRE(n % 4 == 0 && ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0),
                "cenet_adamw: flat buffers must be 16-byte aligned with n %% 4 == 0");
  if (n == 0) return 0;
  adamw_kernel<<<ew_blocks(n / 4), 256, 0, to_stream(st)>>>(p, g, m, v, n, hyper);
  CENET_LAUNCH_CHECK("adamw");
  return 0;
}
