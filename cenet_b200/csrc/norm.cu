// Row-wise normalisations / reductions: LayerNorm, row softmax, SRM channel statistics, segment RMSNorm.
// All are HBM-bound single-pass (per warp) kernels: one warp owns one row, values stay in registers.
#include "common.cuh"

namespace {

// ---- LayerNorm: LPR lanes per row (8 for C=64, 16 for C=128, 32 otherwise), 16-byte vectors, 32/LPR rows per warp -------
// Each lane keeps its NV 8-element vectors in registers: one HBM read, one write per element.
template <typename TI, typename TO, int LPR, int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const TI* __restrict__ x, TO* __restrict__ y,
                                                        const float* __restrict__ g, const float* __restrict__ b,
                                                        long long rows, int C, float eps) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool live = row < rows;
  const TI* xr = x + (live ? row : 0) * C;
  const int nvec = C >> 3;
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const int vi = sub + i * LPR;
    if (vi < nvec) {
      ldv<8>(xr + vi * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; j++) s += v[i][j];
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    if (sub + i * LPR < nvec) {
#pragma unroll
      for (int j = 0; j < 8; j++) { const float d = v[i][j] - mean; q = fmaf(d, d, q); }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + eps);
  if (!live) return;
  TO* yr = y + row * C;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const int vi = sub + i * LPR;
    if (vi < nvec) {
      float gv[8], bv[8], o[8];
      ldv<8>(g + vi * 8, gv);
      ldv<8>(b + vi * 8, bv);
#pragma unroll
      for (int j = 0; j < 8; j++) o[j] = fmaf((v[i][j] - mean) * rstd, gv[j], bv[j]);
      stv<8>(yr + vi * 8, o);
    }
  }
}

// ---- in-place row softmax (materialised attention path only) --------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) softmax_rows_kernel(T* __restrict__ x, int n, long long ld) {
  __shared__ float red[4];
  T* r = x + (long long)blockIdx.x * ld;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float m = -INFINITY;
  for (int i = tid; i < n; i += 128) m = fmaxf(m, ldf(r + i));
  m = warp_max(m);
  if (lane == 0) red[wid] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float s = 0.f;
  for (int i = tid; i < n; i += 128) s += expf(ldf(r + i) - m);
  s = warp_sum(s);
  if (lane == 0) red[wid] = s;
  __syncthreads();
  const float inv = 1.0f / (red[0] + red[1] + red[2] + red[3]);
  for (int i = tid; i < n; i += 128) stf(r + i, expf(ldf(r + i) - m) * inv);
}

// ---- per-row [max, mean, std] over C channels (SRM) -------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) row_stats_kernel(const T* __restrict__ x, long long rows, int C, long long ld,
                                                        int unbiased, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * ld;
  float mx = -INFINITY, s = 0.f;
  for (int c = lane * 2; c < C; c += 64) {
    float t[2];
    ldv<2>(xr + c, t);
    mx = fmaxf(mx, fmaxf(t[0], t[1]));
    s += t[0] + t[1];
  }
  mx = warp_max(mx);
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane * 2; c < C; c += 64) {
    float t[2];
    ldv<2>(xr + c, t);
    q += (t[0] - mean) * (t[0] - mean) + (t[1] - mean) * (t[1] - mean);
  }
  q = warp_sum(q);
  if (lane == 0) {
    stats[row * 3 + 0] = mx;
    stats[row * 3 + 1] = mean;
    stats[row * 3 + 2] = sqrtf(q / (float)(unbiased ? C - 1 : C));
  }
}

// ---- segment RMSNorm: warp per (row, segment) ------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) rmsnorm_seg_kernel(const TI* __restrict__ x, TO* __restrict__ y,
                                                          long long nseg_total, int seg, float eps, float mult) {
  const int lane = threadIdx.x & 31;
  const long long sidx = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (sidx >= nseg_total) return;
  const TI* xs = x + sidx * seg;
  float q = 0.f;
  for (int c = lane; c < seg; c += 32) { float v = ldf(xs + c); q += v * v; }
  const float r = rsqrtf(warp_sum(q) / (float)seg + eps) * mult;
  TO* ys = y + sidx * seg;
  for (int c = lane; c < seg; c += 32) stf(ys + c, ldf(xs + c) * r);
}
}  // namespace

extern "C" int cenet_layernorm(const void* x, int x_dtype, void* y, int y_dtype, const float* gamma,
                               const float* beta, long long rows, int C, float eps, cenet_stream_t s) {
  if (rows == 0) return 0;
  CENET_REQUIRE(x && y && gamma && beta, "cenet_layernorm: null pointer");
  CENET_REQUIRE(C % 64 == 0 && C <= 512, "cenet_layernorm: C=%d must be a multiple of 64 and <= 512", C);
  CENET_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0,
                "cenet_layernorm: pointers must be 16-byte aligned");
  const int wpb = 8;
#define LN_LAUNCH(LPR, NV)                                                                                              \
  do {                                                                                                                    \
    dim3 grid(cdiv(rows, (long long)wpb * (32 / LPR)));                                                                   \
    CENET_DISPATCH(x_dtype, TI, CENET_DISPATCH(y_dtype, TO, (layernorm_kernel<TI, TO, LPR, NV><<<grid, wpb * 32, 0, to_stream(s)>>>( \
        (const TI*)x, (TO*)y, gamma, beta, rows, C, eps))));                                                              \
  } while (0)
  if (C == 64) LN_LAUNCH(8, 1);
  else if (C == 128) LN_LAUNCH(16, 1);
  else if (C <= 256) LN_LAUNCH(32, 1);
  else LN_LAUNCH(32, 2);
#undef LN_LAUNCH
  CENET_LAUNCH_CHECK("layernorm");
  return 0;
}

extern "C" int cenet_softmax_rows(void* x, int dtype, long long rows, int n, long long ld, cenet_stream_t s) {
  if (rows == 0) return 0;
  CENET_REQUIRE(x && n > 0 && ld >= n, "cenet_softmax_rows: bad arguments");
  CENET_DISPATCH(dtype, T, (softmax_rows_kernel<T><<<(unsigned)rows, 128, 0, to_stream(s)>>>((T*)x, n, ld)));
  CENET_LAUNCH_CHECK("softmax_rows");
  return 0;
}

extern "C" int cenet_row_stats(const void* x, int dtype, long long rows, int C, long long ld, int unbiased,
                               float* stats, cenet_stream_t s) {
  if (rows == 0) return 0;
  CENET_REQUIRE(x && stats, "cenet_row_stats: null pointer");
  CENET_REQUIRE(C % 2 == 0 && ld % 2 == 0, "cenet_row_stats: C and ld must be even");
  const int wpb = 8;
  CENET_DISPATCH(dtype, T, (row_stats_kernel<T><<<cdiv(rows, wpb), wpb * 32, 0, to_stream(s)>>>(
      (const T*)x, rows, C, ld, unbiased, stats)));
  CENET_LAUNCH_CHECK("row_stats");
  return 0;
}

extern "C" int cenet_rmsnorm_seg(const void* x, int x_dtype, void* y, int y_dtype, long long rows, int C, int seg,
                                 float eps, float mult, cenet_stream_t s) {
  if (rows == 0) return 0;
  CENET_REQUIRE(x && y && seg > 0 && C % seg == 0, "cenet_rmsnorm_seg: C=%d not a multiple of seg=%d", C, seg);
  const long long nseg = rows * (C / seg);
  const int wpb = 8;
  CENET_DISPATCH(x_dtype, TI, CENET_DISPATCH(y_dtype, TO, (rmsnorm_seg_kernel<TI, TO><<<cdiv(nseg, wpb), wpb * 32, 0, to_stream(s)>>>(
      (const TI*)x, (TO*)y, nseg, seg, eps, mult))));
  CENET_LAUNCH_CHECK("rmsnorm_seg");
  return 0;
}
