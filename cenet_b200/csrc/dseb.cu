// DSEB pieces that are not GEMMs (dseb.py): Feature-Edge-Amplifier fused with the final combine, and the two small
// helpers of the materialised (validation) differential-attention path.
#include "common.cuh"
#include <cmath>

namespace {
constexpr int kMaxScales = 3;
struct FeaScales {
  int n;
  int hd[kMaxScales], wd[kMaxScales];     // down-sampled sizes floor(H*s)
  float inv_s[kMaxScales];                // source scale of the down pass = 1/s   (scale_factor= semantics)
  float up_h[kMaxScales], up_w[kMaxScales];  // source scale of the up pass = hd/H  (size= semantics)
  int identity[kMaxScales];               // s == 1.0 -> edge term exactly 0
  int off[kMaxScales];                    // smem offset (floats) of each down buffer
};

// One CTA per (b,c) plane.  Plane and its down-sampled copies live in shared memory; HBM sees y, gate once (read)
// and z once (write): 3 * H*W * sizeof(T) algorithmic bytes per plane.
template <typename T>
__global__ void __launch_bounds__(256) fea_combine_kernel(const T* __restrict__ y, const T* __restrict__ gate,
                                                          T* __restrict__ z, const float* __restrict__ w_c, int C2,
                                                          int H, int W, const FeaScales sc) {
  extern __shared__ float sm[];
  float* plane = sm;
  const long long pl = blockIdx.x;
  const int c = (int)(pl % C2);
  const int HW = H * W;
  const T* yp = y + pl * HW;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) plane[i] = ldf(yp + i);
  __syncthreads();
  for (int k = 0; k < sc.n; k++) {
    if (sc.identity[k]) continue;
    float* d = sm + sc.off[k];
    const int hd = sc.hd[k], wd = sc.wd[k];
    for (int i = threadIdx.x; i < hd * wd; i += blockDim.x) {
      const int r = i / wd, q = i % wd;
      int h0, h1, w0, w1;
      float lh, lw;
      bilin_src(r, sc.inv_s[k], H, h0, h1, lh);
      bilin_src(q, sc.inv_s[k], W, w0, w1, lw);
      d[i] = (1.f - lh) * ((1.f - lw) * plane[h0 * W + w0] + lw * plane[h0 * W + w1]) +
             lh * ((1.f - lw) * plane[h1 * W + w0] + lw * plane[h1 * W + w1]);
    }
  }
  __syncthreads();
  const float wc = w_c[c];
  const int npair = sc.n * (sc.n - 1) / 2;
  const T* gp = gate + pl * HW;
  T* zp = z + pl * HW;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const int h = i / W, w = i % W;
    const float x = plane[i];
    float e[kMaxScales];
#pragma unroll
    for (int k = 0; k < kMaxScales; k++) {
      e[k] = 0.f;
      if (k < sc.n && !sc.identity[k]) {
        const float* d = sm + sc.off[k];
        const int hd = sc.hd[k], wd = sc.wd[k];
        int h0, h1, w0, w1;
        float lh, lw;
        bilin_src(h, sc.up_h[k], hd, h0, h1, lh);
        bilin_src(w, sc.up_w[k], wd, w0, w1, lw);
        const float up = (1.f - lh) * ((1.f - lw) * d[h0 * wd + w0] + lw * d[h0 * wd + w1]) +
                         lh * ((1.f - lw) * d[h1 * wd + w0] + lw * d[h1 * wd + w1]);
        e[k] = fabsf(x - up);
      }
    }
    float edge = 0.f;
#pragma unroll
    for (int a = 0; a < kMaxScales; a++)
#pragma unroll
      for (int b = a + 1; b < kMaxScales; b++)
        if (b < sc.n) edge += fabsf(e[a] - e[b]);
    edge = npair > 0 ? edge / (float)npair : 0.f;
    stf(zp + i, 2.f * x + wc * edge + ldf(gp + i) * x);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) diff_combine_kernel(T* __restrict__ P, long long npairs, long long map_elems,
                                                           float lambda) {
  const long long total = npairs * map_elems;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long pr = idx / map_elems, e = idx % map_elems;
    T* a = P + (2 * pr) * map_elems + e;
    stf(a, ldf(a) - lambda * ldf(a + map_elems));
  }
}
}  // namespace

extern "C" int cenet_fea_combine(const void* y, const void* gate, void* z, int dtype, const float* w_c, int B, int C2,
                                 int H, int W, const float* scales, int nscales, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(y && gate && z && w_c && scales, "cenet_fea_combine: null pointer");
  CENET_REQUIRE(nscales >= 1 && nscales <= kMaxScales, "cenet_fea_combine: 1..3 scale factors supported, got %d", nscales);
  FeaScales sc;
  sc.n = nscales;
  int off = H * W;
  for (int k = 0; k < kMaxScales; k++) { sc.hd[k] = sc.wd[k] = 1; sc.inv_s[k] = sc.up_h[k] = sc.up_w[k] = 1.f; sc.identity[k] = 1; sc.off[k] = 0; }
  for (int k = 0; k < nscales; k++) {
    const double sf = (double)scales[k];
    CENET_REQUIRE(sf > 0.0, "cenet_fea_combine: scale factor must be positive");
    sc.identity[k] = (scales[k] == 1.0f);
    sc.hd[k] = (int)std::floor((double)H * sf);
    sc.wd[k] = (int)std::floor((double)W * sf);
    CENET_REQUIRE(sc.hd[k] >= 1 && sc.wd[k] >= 1, "cenet_fea_combine: scale %f collapses a %dx%d plane", scales[k], H, W);
    sc.inv_s[k] = (float)(1.0 / sf);
    sc.up_h[k] = (float)sc.hd[k] / (float)H;
    sc.up_w[k] = (float)sc.wd[k] / (float)W;
    sc.off[k] = off;
    if (!sc.identity[k]) off += sc.hd[k] * sc.wd[k];
  }
  const size_t smem = (size_t)off * sizeof(float);
  CENET_REQUIRE(smem <= 227 * 1024, "cenet_fea_combine: plane %dx%d needs %zu bytes of shared memory", H, W, smem);
  const long long planes = (long long)B * C2;
#define LAUNCH_FEA(T)                                                                                         \
  do {                                                                                                        \
    if (smem > 48 * 1024) cudaFuncSetAttribute(fea_combine_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    fea_combine_kernel<T><<<(unsigned)planes, 256, smem, to_stream(s)>>>((const T*)y, (const T*)gate, (T*)z, w_c, C2, H, W, sc); \
  } while (0)
  CENET_DISPATCH(dtype, T, LAUNCH_FEA(T));
#undef LAUNCH_FEA
  CENET_LAUNCH_CHECK("fea_combine");
  return 0;
}

extern "C" int cenet_diff_combine(void* P, int dtype, long long npairs, long long map_elems, float lambda,
                                  cenet_stream_t s) {
  if (npairs == 0) return 0;
  CENET_REQUIRE(P && map_elems > 0, "cenet_diff_combine: bad arguments");
  const long long total = npairs * map_elems;
  const int grid = (int)std::min<long long>(cdiv(total, 256), (long long)kNumSMs * 32);
  CENET_DISPATCH(dtype, T, (diff_combine_kernel<T><<<grid, 256, 0, to_stream(s)>>>((T*)P, npairs, map_elems, lambda)));
  CENET_LAUNCH_CHECK("diff_combine");
  return 0;
}
