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

// A CTA handles PPB consecutive (b,c) planes, each by a team of 256/PPB threads (PPB = 1 for 56x56 planes, 4 for 28x28,
// 8 for 14x14, so small planes do not leave most of the block idle).  Each plane and its down-sampled copies live in
// shared memory; the bilinear source indices / weights of every row and column are tabulated once per CTA, so the
// per-pixel work is 4 smem reads + 3 FMAs per scale.  HBM sees y and gate once (read) and z once (write):
// 3 * H*W * sizeof(T) algorithmic bytes per plane.
struct LerpTab { short i0, i1; float l; };

template <typename T>
__global__ void __launch_bounds__(256) fea_combine_kernel(const T* __restrict__ y, const T* __restrict__ gate,
                                                          T* __restrict__ z, const float* __restrict__ w_c, int C2,
                                                          int H, int W, long long nplanes, int ppb, unsigned wmagic,
                                                          const FeaScales sc, int plane_floats, int mode) {
  // mode 0 (CENet, dseb.py:63-76,157): z = 2y + w * mean_{i<j} | |y - y_i| - |y - y_j| | + gate * y
  // mode 1 (CENetOrg DoGEdge, cenet_org/decoders.py:112-125): z = y + w * |y_0 - y_1|          (gate unused)
  // mode 2 (CENetOrg SkipEnhancer combine, decoders.py:140-143): z = y + gate * y               (no scales)
  // y_k = bilinear up(down_k(y)) of scale k
  extern __shared__ float sm[];
  // tables: for each scale: down rows (hd), down cols (wd), up rows (H), up cols (W)
  LerpTab* tab = reinterpret_cast<LerpTab*>(sm);
  int toff[kMaxScales][4];
  int tcount = 0;
#pragma unroll
  for (int k = 0; k < kMaxScales; k++) {
    toff[k][0] = tcount; tcount += (k < sc.n && !sc.identity[k]) ? sc.hd[k] : 0;
    toff[k][1] = tcount; tcount += (k < sc.n && !sc.identity[k]) ? sc.wd[k] : 0;
    toff[k][2] = tcount; tcount += (k < sc.n && !sc.identity[k]) ? H : 0;
    toff[k][3] = tcount; tcount += (k < sc.n && !sc.identity[k]) ? W : 0;
  }
  for (int k = 0; k < sc.n; k++) {
    if (sc.identity[k]) continue;
    for (int i = threadIdx.x; i < sc.hd[k] + sc.wd[k] + H + W; i += blockDim.x) {
      int i0, i1; float l; int dst;
      if (i < sc.hd[k]) { bilin_src(i, sc.inv_s[k], H, i0, i1, l); dst = toff[k][0] + i; }
      else if (i < sc.hd[k] + sc.wd[k]) { bilin_src(i - sc.hd[k], sc.inv_s[k], W, i0, i1, l); dst = toff[k][1] + i - sc.hd[k]; }
      else if (i < sc.hd[k] + sc.wd[k] + H) { bilin_src(i - sc.hd[k] - sc.wd[k], sc.up_h[k], sc.hd[k], i0, i1, l); dst = toff[k][2] + i - sc.hd[k] - sc.wd[k]; }
      else { bilin_src(i - sc.hd[k] - sc.wd[k] - H, sc.up_w[k], sc.wd[k], i0, i1, l); dst = toff[k][3] + i - sc.hd[k] - sc.wd[k] - H; }
      tab[dst].i0 = (short)i0; tab[dst].i1 = (short)i1; tab[dst].l = l;
    }
  }
  const int team = blockDim.x / ppb, tm = threadIdx.x / team, tt = threadIdx.x - tm * team;
  const long long pl = (long long)blockIdx.x * ppb + tm;
  const bool live = pl < nplanes;
  const int HW = H * W;
  float* plane = sm + 2 * tcount + (size_t)tm * plane_floats;     // LerpTab = 8 bytes = 2 floats
  if (live) {
    const T* yp = y + pl * HW;
    for (int i = tt; i < HW; i += team) plane[i] = ldf(yp + i);
  }
  __syncthreads();
  if (live) {
    for (int k = 0; k < sc.n; k++) {
      if (sc.identity[k]) continue;
      float* d = plane + sc.off[k];
      const int hd = sc.hd[k], wd = sc.wd[k];
      const unsigned dmagic = 0xFFFFFFFFu / (unsigned)wd + 1u;
      for (int i = tt; i < hd * wd; i += team) {
        const int r = (int)__umulhi((unsigned)i, dmagic), q = i - r * wd;
        const LerpTab a = tab[toff[k][0] + r], bcol = tab[toff[k][1] + q];
        const float top = fmaf(bcol.l, plane[a.i0 * W + bcol.i1] - plane[a.i0 * W + bcol.i0], plane[a.i0 * W + bcol.i0]);
        const float bot = fmaf(bcol.l, plane[a.i1 * W + bcol.i1] - plane[a.i1 * W + bcol.i0], plane[a.i1 * W + bcol.i0]);
        d[i] = fmaf(a.l, bot - top, top);
      }
    }
  }
  __syncthreads();
  if (!live) return;
  const int c = (int)(pl % C2);
  const float wc = w_c[c];
  const int npair = sc.n * (sc.n - 1) / 2;
  const float inv_pair = npair > 0 ? 1.f / (float)npair : 0.f;
  const T* gp = gate ? gate + pl * HW : nullptr;
  T* zp = z + pl * HW;
  for (int i = tt; i < HW; i += team) {
    const int h = (int)__umulhi((unsigned)i, wmagic), w = i - h * W;
    const float x = plane[i];
    float e[kMaxScales];
#pragma unroll
    for (int k = 0; k < kMaxScales; k++) {
      e[k] = mode == 1 ? x : 0.f;                      // identity scale: y_k == y
      if (k < sc.n && !sc.identity[k]) {
        const float* d = plane + sc.off[k];
        const int wd = sc.wd[k];
        const LerpTab a = tab[toff[k][2] + h], bcol = tab[toff[k][3] + w];
        const float top = fmaf(bcol.l, d[a.i0 * wd + bcol.i1] - d[a.i0 * wd + bcol.i0], d[a.i0 * wd + bcol.i0]);
        const float bot = fmaf(bcol.l, d[a.i1 * wd + bcol.i1] - d[a.i1 * wd + bcol.i0], d[a.i1 * wd + bcol.i0]);
        const float yk = fmaf(a.l, bot - top, top);
        e[k] = mode == 1 ? yk : fabsf(x - yk);
      }
    }
    if (mode == 1) {
      stf(zp + i, fmaf(wc, fabsf(e[0] - e[1]), x));
      continue;
    }
    if (mode == 2) {
      stf(zp + i, fmaf(ldf(gp + i), x, x));
      continue;
    }
    float edge = 0.f;
#pragma unroll
    for (int a = 0; a < kMaxScales; a++)
#pragma unroll
      for (int b = a + 1; b < kMaxScales; b++)
        if (b < sc.n) edge += fabsf(e[a] - e[b]);
    stf(zp + i, fmaf(wc * inv_pair, edge, fmaf(ldf(gp + i), x, 2.f * x)));
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

static int fea_launch(const void* y, const void* gate, void* z, int dtype, const float* w_c, int B, int C2, int H, int W,
                      const float* scales, int nscales, int mode, cenet_stream_t s) {
  if (B == 0) return 0;
  CENET_REQUIRE(y && z && w_c, "cenet_fea_combine: null pointer");
  CENET_REQUIRE(mode == 1 || gate, "cenet_fea_combine: null gate");
  CENET_REQUIRE(mode == 2 || scales, "cenet_fea_combine: null scale factors");
  if (mode == 2) nscales = 0;
  CENET_REQUIRE((mode == 2 || nscales >= 1) && nscales <= kMaxScales, "cenet_fea_combine: 1..3 scale factors supported, got %d", nscales);
  CENET_REQUIRE(mode != 1 || nscales == 2, "cenet_dog_combine: the difference of Gaussians takes exactly two scale factors, got %d", nscales);
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
  const int plane_floats = off;
  int tcount = 0;
  for (int k = 0; k < nscales; k++)
    if (!sc.identity[k]) tcount += sc.hd[k] + sc.wd[k] + H + W;
  const int HW = H * W;
  const int ppb = HW >= 2048 ? 1 : (HW >= 512 ? 4 : 8);
  const size_t smem = ((size_t)2 * tcount + (size_t)ppb * plane_floats) * sizeof(float);
  CENET_REQUIRE(smem <= 227 * 1024, "cenet_fea_combine: plane %dx%d needs %zu bytes of shared memory", H, W, smem);
  CENET_REQUIRE(H <= 32767 && W <= 32767 && HW < 65536, "cenet_fea_combine: plane too large");
  const long long planes = (long long)B * C2;
  const unsigned wmagic = 0xFFFFFFFFu / (unsigned)W + 1u;        // i / W for i < 2^16 via __umulhi
  const unsigned grid = (unsigned)((planes + ppb - 1) / ppb);
#define LAUNCH_FEA(T)                                                                                         \
  do {                                                                                                        \
    if (smem > 48 * 1024) cudaFuncSetAttribute(fea_combine_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    fea_combine_kernel<T><<<grid, 256, smem, to_stream(s)>>>((const T*)y, (const T*)gate, (T*)z, w_c, C2, H, W, planes, ppb, wmagic, sc, plane_floats, mode); \
  } while (0)
  CENET_DISPATCH(dtype, T, LAUNCH_FEA(T));
#undef LAUNCH_FEA
  CENET_LAUNCH_CHECK("fea_combine");
  return 0;
}

extern "C" int cenet_fea_combine(const void* y, const void* gate, void* z, int dtype, const float* w_c, int B, int C2,
                                 int H, int W, const float* scales, int nscales, cenet_stream_t s) {
  return fea_launch(y, gate, z, dtype, w_c, B, C2, H, W, scales, nscales, 0, s);
}

extern "C" int cenet_dog_combine(const void* y, const void* gate, void* z, int dtype, const float* w_c, int B, int C2, int H,
                                 int W, const float* scales, int nscales, int mode, cenet_stream_t s) {
  CENET_REQUIRE(mode == 1 || mode == 2, "cenet_dog_combine: mode %d (1 = DoG edge, 2 = y + gate*y)", mode);
  return fea_launch(y, gate, z, dtype, w_c, B, C2, H, W, scales, nscales, mode, s);
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
