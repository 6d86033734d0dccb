// DSEB pieces that are not GEMMs (dseb.py): Feature-Edge-Amplifier fused with the final combine, and the two small
// helpers of the materialised (validation) differential-attention path.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace {
constexpr int kMaxScales = 3;
struct FeaScales {
  int n;
  int hd[kMaxScales], wd[kMaxScales];     // down-sampled sizes floor(H*s)
  float inv_s[kMaxScales];                // source scale of the down pass = 1/s   (scale_factor= semantics)
  float up_h[kMaxScales], up_w[kMaxScales];  // source scale of the up pass = hd/H  (size= semantics)
  int identity[kMaxScales];               // s == 1.0 -> edge term exactly 0
  int off[kMaxScales];                    // smem offset (floats) of each down buffer
  int off2[kMaxScales];                   // smem offset of the horizontally up-sampled copy [hd][W] (separable up pass)
  int sep;                                // 1: the up pass runs separably (W even and the extra buffers fit)
  int goff;                               // >= 0: smem offset (floats) of the staged bf16 gate plane (HW / 2 floats); -1: gate from global
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
  float* plane = sm + ((2 * tcount + 3) & ~3) + (size_t)tm * plane_floats;     // LerpTab = 8 bytes = 2 floats; 16-byte aligned planes
  if (live) {
    const T* yp = y + pl * HW;
    if (sizeof(T) == 2 && (HW & 7) == 0 && ((reinterpret_cast<uintptr_t>(yp) & 15) == 0)) {
      for (int i = tt * 8; i < HW; i += team * 8) {                  // 16 bytes per load
        float v[8];
        ldv<8>(yp + i, v);
#pragma unroll
        for (int j = 0; j < 8; j++) plane[i + j] = v[j];
      }
    } else {
      for (int i = tt; i < HW; i += team) plane[i] = ldf(yp + i);
    }
    // the gate plane is staged with 16-byte loads as well: read per pixel inside the final loop it left ~4 bytes per thread in
    // flight (8 KB per SM against the ~40 KB the HBM latency needs) and the kernel ran at 0.7 TB/s
    if (sc.goff >= 0 && gate) {
      const uint4* gsrc = reinterpret_cast<const uint4*>(gate + pl * HW);
      uint4* gdst = reinterpret_cast<uint4*>(plane + sc.goff);
      for (int i = tt; i < (HW >> 3); i += team) gdst[i] = gsrc[i];
    }
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
  const int c = live ? (int)(pl % C2) : 0;
  const float wc = w_c[c];
  const int npair = sc.n * (sc.n - 1) / 2;
  const float inv_pair = npair > 0 ? 1.f / (float)npair : 0.f;
  const T* gp = gate ? gate + pl * HW : nullptr;
  T* zp = z + pl * HW;
  if (sc.sep) {
    // ---- separable up pass: T2_k[r][w] = horizontal lerp of down row r, then the final pass needs ONE vertical lerp per scale
    //      (2 shared-memory reads instead of 4 + two table entries) and handles two adjacent pixels per thread.  The arithmetic
    //      (lerp order) is that of the direct form below: results are bit-identical. ----
    if (live) {
      for (int k = 0; k < sc.n; k++) {
        if (sc.identity[k]) continue;
        const float* d = plane + sc.off[k];
        float* t2 = plane + sc.off2[k];
        const int wd = sc.wd[k], n2 = sc.hd[k] * W;
        for (int i = tt; i < n2; i += team) {
          const int r = (int)__umulhi((unsigned)i, wmagic), w = i - r * W;
          const LerpTab bcol = tab[toff[k][3] + w];
          const float d0 = d[r * wd + bcol.i0];
          t2[i] = fmaf(bcol.l, d[r * wd + bcol.i1] - d0, d0);
        }
      }
    }
    __syncthreads();
    if (!live) return;
    for (int i2 = tt; i2 < (HW >> 1); i2 += team) {
      const int i = 2 * i2;
      const int h = (int)__umulhi((unsigned)i, wmagic), w = i - h * W;
      const float2 x = *reinterpret_cast<const float2*>(plane + i);
      float e0[kMaxScales], e1[kMaxScales];
#pragma unroll
      for (int k = 0; k < kMaxScales; k++) {
        e0[k] = mode == 1 ? x.x : 0.f; e1[k] = mode == 1 ? x.y : 0.f;
        if (k < sc.n && !sc.identity[k]) {
          const float* t2 = plane + sc.off2[k];
          const LerpTab a = tab[toff[k][2] + h];
          const float2 top = *reinterpret_cast<const float2*>(t2 + a.i0 * W + w);
          const float2 bot = *reinterpret_cast<const float2*>(t2 + a.i1 * W + w);
          const float y0 = fmaf(a.l, bot.x - top.x, top.x), y1 = fmaf(a.l, bot.y - top.y, top.y);
          e0[k] = mode == 1 ? y0 : fabsf(x.x - y0);
          e1[k] = mode == 1 ? y1 : fabsf(x.y - y1);
        }
      }
      float z0, z1;
      if (mode == 1) {
        z0 = fmaf(wc, fabsf(e0[0] - e0[1]), x.x); z1 = fmaf(wc, fabsf(e1[0] - e1[1]), x.y);
      } else {
        float g0, g1;
        if (sc.goff >= 0) {
          const float2 g = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(plane + sc.goff)[i2]);
          g0 = g.x; g1 = g.y;
        } else if (sizeof(T) == 2) {
          const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(gp + i));
          g0 = g.x; g1 = g.y;
        } else {
          g0 = ldf(gp + i); g1 = ldf(gp + i + 1);
        }
        float ed0 = 0.f, ed1 = 0.f;
#pragma unroll
        for (int a = 0; a < kMaxScales; a++)
#pragma unroll
          for (int b = a + 1; b < kMaxScales; b++)
            if (b < sc.n) { ed0 += fabsf(e0[a] - e0[b]); ed1 += fabsf(e1[a] - e1[b]); }
        z0 = fmaf(wc * inv_pair, ed0, fmaf(g0, x.x, 2.f * x.x));
        z1 = fmaf(wc * inv_pair, ed1, fmaf(g1, x.y, 2.f * x.y));
      }
      if (sizeof(T) == 2) *reinterpret_cast<__nv_bfloat162*>(zp + i) = __floats2bfloat162_rn(z0, z1);
      else { stf(zp + i, z0); stf(zp + i + 1, z1); }
    }
    return;
  }
  if (!live) return;
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

// ---------------------------------------------------------------------------------------------------------------- v2
// Same arithmetic (bit-identical lerp order), organised for instruction count -- ncu on the first kernel: 172 thread instructions
// per pixel, issue-bound at 0.7 TB/s.  Persistent CTAs (tables built once), 16-byte table entries with pre-multiplied row offsets,
// and every pass gives a thread ONE column (or column pair) and a run of rows, so the column taps live in registers, the row taps
// are one broadcast LDS.128 and the addresses advance by additions:
//   down   D[r][q]   = lerp_rows(lerp_cols(plane))          thread = down column q x row chunk
//   up-h   T2[r][w]  = lerp_cols(D[r])                      thread = column w x row chunk
//   final  z[h][w..w+1] from x, lerp_rows(T2), gate         thread = column pair x row chunk
struct Tab4 { int a, b; float l; int pad; };

template <typename T>
__global__ void __launch_bounds__(256, 4) fea_combine_v2_kernel(const T* __restrict__ y, const T* __restrict__ gate, T* __restrict__ z,
                                                             const float* __restrict__ w_c, int C2, int H, int W, long long nplanes,
                                                             int ppb, const FeaScales sc, int plane_floats, int mode, int ngroups) {
  extern __shared__ __align__(16) float sm2[];
  Tab4* tab = reinterpret_cast<Tab4*>(sm2);
  // per active scale: DR (hd rows), DC (wd cols), UR (H rows), UC (W cols)
  int tb[kMaxScales][4], act[kMaxScales], hdv[kMaxScales], wdv[kMaxScales], offv[kMaxScales], off2v[kMaxScales];
  int tcount = 0, nact = 0;
#pragma unroll
  for (int k = 0; k < kMaxScales; k++) {
    act[k] = (k < sc.n && !sc.identity[k]) ? 1 : 0;
    hdv[k] = sc.hd[k]; wdv[k] = sc.wd[k]; offv[k] = sc.off[k]; off2v[k] = sc.off2[k];
    tb[k][0] = tcount; tcount += act[k] ? hdv[k] : 0;
    tb[k][1] = tcount; tcount += act[k] ? wdv[k] : 0;
    tb[k][2] = tcount; tcount += act[k] ? H : 0;
    tb[k][3] = tcount; tcount += act[k] ? W : 0;
    nact += act[k];
  }
#pragma unroll
  for (int k = 0; k < kMaxScales; k++) {
    if (!act[k]) continue;
    const int hd = hdv[k], wd = wdv[k];
    for (int i = threadIdx.x; i < hd + wd + H + W; i += blockDim.x) {
      int i0, i1; float l; Tab4 e;
      if (i < hd) { bilin_src(i, sc.inv_s[k], H, i0, i1, l); e.a = i0 * W; e.b = i1 * W; e.l = l; e.pad = 0; tab[tb[k][0] + i] = e; }
      else if (i < hd + wd) { bilin_src(i - hd, sc.inv_s[k], W, i0, i1, l); e.a = i0; e.b = i1; e.l = l; e.pad = 0; tab[tb[k][1] + i - hd] = e; }
      else if (i < hd + wd + H) { bilin_src(i - hd - wd, sc.up_h[k], hd, i0, i1, l); e.a = i0 * W; e.b = i1 * W; e.l = l; e.pad = 0; tab[tb[k][2] + i - hd - wd] = e; }
      else { bilin_src(i - hd - wd - H, sc.up_w[k], wd, i0, i1, l); e.a = i0; e.b = i1; e.l = l; e.pad = 0; tab[tb[k][3] + i - hd - wd - H] = e; }
    }
  }
  const int team = blockDim.x / ppb, tm = threadIdx.x / team, tt = threadIdx.x - tm * team;
  const int HW = H * W;
  float* plane = sm2 + 4 * tcount + (size_t)tm * plane_floats;
  const int npair = sc.n * (sc.n - 1) / 2;
  const float inv_pair = npair > 0 ? 1.f / (float)npair : 0.f;
  // thread -> (column, row chunk) of each pass
  auto split = [&](int ncols, int nrows, int& col, int& r0, int& r1) {
    const int cw = min(ncols, team), nchunk = max(1, team / cw), rpc = (nrows + nchunk - 1) / nchunk;
    const int ch = tt / cw;
    col = tt - ch * cw;
    r0 = min(nrows, ch * rpc); r1 = ch < nchunk ? min(nrows, r0 + rpc) : r0;
    return cw;
  };
  // the (column, row chunk) assignment of every pass depends on the plane geometry only: computed once (the integer divisions of
  // `split` inside the plane loop were 20 % of all instructions)
  int dq[kMaxScales], dr0[kMaxScales], dr1[kMaxScales], dcw[kMaxScales], uw[kMaxScales], ur0[kMaxScales], ur1[kMaxScales], ucw[kMaxScales];
#pragma unroll
  for (int k = 0; k < kMaxScales; k++) {
    dq[k] = dr0[k] = dr1[k] = uw[k] = ur0[k] = ur1[k] = 0; dcw[k] = ucw[k] = 1;
    if (act[k]) {
      dcw[k] = split(wdv[k], hdv[k], dq[k], dr0[k], dr1[k]);
      ucw[k] = split(W, hdv[k], uw[k], ur0[k], ur1[k]);
    }
  }
  int fcp, fr0, fr1;
  const int fcw = split(W >> 1, H, fcp, fr0, fr1);
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const long long pl = (long long)grp * ppb + tm;
    const bool live = pl < nplanes;
    __syncthreads();                                                   // tables ready / previous group done with the planes
    if (live) {
      const T* yp = y + pl * HW;
      if (sizeof(T) == 2 && (HW & 7) == 0 && ((reinterpret_cast<uintptr_t>(yp) & 15) == 0)) {
        for (int i = tt * 8; i < HW; i += team * 8) {
          float v[8];
          ldv<8>(yp + i, v);
          *reinterpret_cast<float4*>(plane + i) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(plane + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      } else {
        for (int i = tt; i < HW; i += team) plane[i] = ldf(yp + i);
      }
      if (sc.goff >= 0 && gate) {
        const uint4* gsrc = reinterpret_cast<const uint4*>(gate + pl * HW);
        uint4* gdst = reinterpret_cast<uint4*>(plane + sc.goff);
        for (int i = tt; i < (HW >> 3); i += team) gdst[i] = gsrc[i];
      }
    }
    __syncthreads();
    if (live) {
#pragma unroll
      for (int k = 0; k < kMaxScales; k++) {
        if (!act[k]) continue;
        float* d = plane + offv[k];
        const int hd = hdv[k], wd = wdv[k];
        const int r0 = dr0[k], r1 = dr1[k], cw = dcw[k];
        for (int q = dq[k]; q < wd; q += cw) {
          const Tab4 bc = tab[tb[k][1] + q];
          for (int r = r0; r < r1; r++) {
            const Tab4 rt = tab[tb[k][0] + r];
            const float* p0 = plane + rt.a;
            const float* p1 = plane + rt.b;
            const float top = fmaf(bc.l, p0[bc.b] - p0[bc.a], p0[bc.a]);
            const float bot = fmaf(bc.l, p1[bc.b] - p1[bc.a], p1[bc.a]);
            d[r * wd + q] = fmaf(rt.l, bot - top, top);
          }
        }
      }
    }
    __syncthreads();
    if (live) {
#pragma unroll
      for (int k = 0; k < kMaxScales; k++) {
        if (!act[k]) continue;
        const float* d = plane + offv[k];
        float* t2 = plane + off2v[k];
        const int hd = hdv[k], wd = wdv[k];
        const int r0 = ur0[k], r1 = ur1[k], cw = ucw[k];
        for (int w = uw[k]; w < W; w += cw) {
          const Tab4 bc = tab[tb[k][3] + w];
          for (int r = r0; r < r1; r++) {
            const float d0 = d[r * wd + bc.a];
            t2[r * W + w] = fmaf(bc.l, d[r * wd + bc.b] - d0, d0);
          }
        }
      }
    }
    __syncthreads();
    if (live) {
      const float wc = w_c[(int)(pl % C2)];
      const T* gp = gate ? gate + pl * HW : nullptr;
      T* zp = z + pl * HW;
      const int r0 = fr0, r1 = fr1, cw = fcw;
      for (int cp = fcp; cp < (W >> 1); cp += cw) {
        const int w = 2 * cp;
        for (int h = r0; h < r1; h++) {
          const int i = h * W + w;
          const float2 x = *reinterpret_cast<const float2*>(plane + i);
          float e0[kMaxScales], e1[kMaxScales];
#pragma unroll
          for (int k = 0; k < kMaxScales; k++) {
            e0[k] = mode == 1 ? x.x : 0.f; e1[k] = mode == 1 ? x.y : 0.f;
            if (act[k]) {
              const float* t2 = plane + off2v[k] + w;
              const Tab4 a = tab[tb[k][2] + h];
              const float2 top = *reinterpret_cast<const float2*>(t2 + a.a);
              const float2 bot = *reinterpret_cast<const float2*>(t2 + a.b);
              const float y0 = fmaf(a.l, bot.x - top.x, top.x), y1 = fmaf(a.l, bot.y - top.y, top.y);
              e0[k] = mode == 1 ? y0 : fabsf(x.x - y0);
              e1[k] = mode == 1 ? y1 : fabsf(x.y - y1);
            }
          }
          float z0, z1;
          if (mode == 1) {
            z0 = fmaf(wc, fabsf(e0[0] - e0[1]), x.x); z1 = fmaf(wc, fabsf(e1[0] - e1[1]), x.y);
          } else {
            float g0, g1;
            if (sc.goff >= 0) {
              const float2 g = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(plane + sc.goff)[i >> 1]);
              g0 = g.x; g1 = g.y;
            } else {
              g0 = ldf(gp + i); g1 = ldf(gp + i + 1);
            }
            float ed0 = 0.f, ed1 = 0.f;
#pragma unroll
            for (int a = 0; a < kMaxScales; a++)
#pragma unroll
              for (int b = a + 1; b < kMaxScales; b++)
                if (b < sc.n) { ed0 += fabsf(e0[a] - e0[b]); ed1 += fabsf(e1[a] - e1[b]); }
            z0 = fmaf(wc * inv_pair, ed0, fmaf(g0, x.x, 2.f * x.x));
            z1 = fmaf(wc * inv_pair, ed1, fmaf(g1, x.y, 2.f * x.y));
          }
          if (sizeof(T) == 2) *reinterpret_cast<__nv_bfloat162*>(zp + i) = __floats2bfloat162_rn(z0, z1);
          else { stf(zp + i, z0); stf(zp + i + 1, z1); }
        }
      }
    }
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
    if (!sc.identity[k]) off += (sc.hd[k] * sc.wd[k] + 1) & ~1;            // even offsets: the pair pass reads float2
  }
  const int HW_ = H * W;
  const int ppb_ = HW_ >= 2048 ? 1 : (HW_ >= 512 ? 4 : 8);
  int off2 = off, tc2 = 0;
  for (int k = 0; k < kMaxScales; k++) {
    sc.off2[k] = off2;
    if (k < nscales && !sc.identity[k]) { off2 += sc.hd[k] * W; tc2 += sc.hd[k] + sc.wd[k] + H + W; }
  }
  // separable up pass when the planes are even (pairs of pixels, 8-byte accesses) and the extra buffers fit; mode 2 has no scales
  static const bool sep_on = !(getenv("CENET_B200_FEA_SEP") && atoi(getenv("CENET_B200_FEA_SEP")) == 0);
  sc.sep = sep_on && mode != 2 && (W % 2 == 0) && (HW_ % 2 == 0) && ((size_t)2 * tc2 + (size_t)ppb_ * off2) * sizeof(float) <= 200 * 1024;
  if (sc.sep) off = off2;
  sc.goff = -1;
  if (sc.sep && dtype == CENET_BF16 && gate && mode == 0 && HW_ % 8 == 0 && (((uintptr_t)gate & 15) == 0)) {
    off = (off + 3) & ~3;                                                  // 16-byte aligned staging area
    sc.goff = off;
    off += HW_ / 2;
    off = (off + 1) & ~1;
  }
  const int plane_floats = (off + 3) & ~3;
  int tcount = 0;
  for (int k = 0; k < nscales; k++)
    if (!sc.identity[k]) tcount += sc.hd[k] + sc.wd[k] + H + W;
  const int HW = H * W;
  const int ppb = HW >= 2048 ? 1 : (HW >= 512 ? 4 : 8);
  const size_t smem = ((((size_t)2 * tcount + 3) & ~(size_t)3) + (size_t)ppb * plane_floats) * sizeof(float);
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
  static const bool v2_on = !(getenv("CENET_B200_FEA_V2") && atoi(getenv("CENET_B200_FEA_V2")) == 0);
  if (sc.sep && v2_on) {
    // persistent column-ownership kernel; 16-byte table entries
    const size_t smem2 = ((size_t)4 * tcount + (size_t)ppb * plane_floats) * sizeof(float);
    if (smem2 <= 200 * 1024) {
      const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, 227 * 1024 / (smem2 + 1024)));
      const int ngroups = (int)grid;
      const int blocks = std::min(ngroups, per_sm * kNumSMs);
#define LAUNCH_FEA2(T)                                                                                        \
      do {                                                                                                    \
        if (smem2 > 48 * 1024) cudaFuncSetAttribute(fea_combine_v2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2); \
        fea_combine_v2_kernel<T><<<blocks, 256, smem2, to_stream(s)>>>((const T*)y, (const T*)gate, (T*)z, w_c, C2, H, W, planes, ppb, sc, plane_floats, mode, ngroups); \
      } while (0)
      CENET_DISPATCH(dtype, T, LAUNCH_FEA2(T));
#undef LAUNCH_FEA2
      CENET_LAUNCH_CHECK("fea_combine_v2");
      return 0;
    }
  }
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
