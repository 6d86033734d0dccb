// Weight-gradient GEMM:  dW[n, k] = sum_m rs[m] * dY[m, n] * X[m, k]   (contraction over the token / pixel dimension).
//
// Both operands are row-major over m, i.e. "MN-major" for the tensor core, so the tiles are staged [m][n] / [m][k] in shared
// memory and read with ldmatrix.trans into mma.sync m16n8k16 bf16 fragments (fp32 accumulate).  The long contraction is
// split over CTAs (split-M); every split writes an fp32 partial tile into the workspace and a second kernel reduces the
// partials in fixed order (deterministic) while permuting k = (tap, cin) into the reference's [Cout, Cin, KH, KW] order.
// Replaces autograd's weight gradients of every nn.Linear / 1x1 / dense conv on the path (pvtv2.py:41-45,90-106;
// cfam.py:150-157,301-303; dseb.py:164; unet.py:201-214; blocks.py:209-214; nlb.py:107-143).
// fp32 x fp32 operands (validation precision) go through the CUDA-core GEMM with the same split / reduce.
#include "train_common.cuh"
#include <cstdlib>

namespace {
constexpr int TN = 64, TK = 64, TM = 64, PITCH = 72;      // +8 bf16 padding: conflict-free ldmatrix
constexpr int WG_THREADS = 128;

struct WgParams {
  const void* dy; int dy_dtype; long long ldy;
  const void* x; int x_dtype; long long ldx;
  long long M; int N, K;
  const float* rs; int rs_div;
  long long m_per_split;
  float* ws;                       // [S][N][K]
  int fast_y, fast_x;              // 16-byte vector loads allowed
  // implicit im2col of the X operand (stride-1 "same" conv): x is the NHWC image, k = (tap, ci)
  int conv, H, W, Cin, KS, pad;
};

__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const bf16* p) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// stage one [TM][64] tile chunk (8 columns) into registers as 8 bf16
__device__ __forceinline__ uint4 load_chunk(const void* base, int dtype, long long ld, long long m, long long M, int c, int Ccols,
                                            int fast, float scale) {
  uint4 u = make_uint4(0, 0, 0, 0);
  if (m >= M || c >= Ccols) return u;
  if (fast && c + 8 <= Ccols) {
    u = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + m * ld + c);
    if (scale != 1.f) {
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float2 f = __bfloat1622float2(h[i]);
        h[i] = __floats2bfloat162_rn(f.x * scale, f.y * scale);
      }
    }
    return u;
  }
  bf16 t[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float v = (c + i < Ccols) ? ld_any(base, dtype, m * ld + c + i) * scale : 0.f;
    t[i] = __float2bfloat16_rn(v);
  }
  return *reinterpret_cast<uint4*>(t);
}

// X chunk of an implicit-im2col operand: row m = output pixel (b, h, w), columns c..c+7 = tap t = c / Cin, channels c % Cin..
__device__ __forceinline__ uint4 load_chunk_conv(const WgParams& p, long long m, long long M, int c) {
  uint4 u = make_uint4(0, 0, 0, 0);
  if (m >= M || c >= p.K) return u;
  const int w = (int)(m % p.W);
  const long long t2 = m / p.W;
  const int h = (int)(t2 % p.H);
  const long long b = t2 / p.H;
  if (p.fast_x && (p.Cin & 7) == 0) {
    const int t = c / p.Cin, ci = c - t * p.Cin;
    const int hh = h + t / p.KS - p.pad, ww = w + t % p.KS - p.pad;
    if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W)
      u = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.x) + ((b * p.H + hh) * p.W + ww) * p.ldx + ci);
    return u;
  }
  bf16 tmp[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float v = 0.f;
    const int cc = c + i;
    if (cc < p.K) {
      const int t = cc / p.Cin, ci = cc - t * p.Cin;
      const int hh = h + t / p.KS - p.pad, ww = w + t % p.KS - p.pad;
      if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) v = ld_any(p.x, p.x_dtype, ((b * p.H + hh) * p.W + ww) * p.ldx + ci);
    }
    tmp[i] = __float2bfloat16_rn(v);
  }
  return *reinterpret_cast<uint4*>(tmp);
}

__global__ void __launch_bounds__(WG_THREADS) wgrad_mma_kernel(const WgParams p) {
  __shared__ __align__(16) bf16 sY[2][TM][PITCH];
  __shared__ __align__(16) bf16 sX[2][TM][PITCH];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * TN, k0 = blockIdx.y * TK;
  const long long mb = (long long)blockIdx.z * p.m_per_split;
  const long long me = min(p.M, mb + p.m_per_split);
  const int wn = (warp >> 1) * 32, wk = (warp & 1) * 32;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) acc[i][j][q] = 0.f;

  // staging map: 512 chunks per tile (64 rows x 8 column chunks); thread handles 4 of each tile
  uint4 ry[4], rx[4];
  auto stage = [&](long long m0) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int ch = tid + i * WG_THREADS;
      const int r = ch >> 3, c = (ch & 7) * 8;
      const long long m = m0 + r;
      float sc = 1.f;
      if (p.rs && m < me) sc = p.rs[m / p.rs_div];
      ry[i] = load_chunk(p.dy, p.dy_dtype, p.ldy, m, me, n0 + c, p.N, p.fast_y, sc);
      rx[i] = p.conv ? load_chunk_conv(p, m, me, k0 + c) : load_chunk(p.x, p.x_dtype, p.ldx, m, me, k0 + c, p.K, p.fast_x, 1.f);
    }
  };
  auto commit = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int ch = tid + i * WG_THREADS;
      const int r = ch >> 3, c = (ch & 7) * 8;
      *reinterpret_cast<uint4*>(&sY[buf][r][c]) = ry[i];
      *reinterpret_cast<uint4*>(&sX[buf][r][c]) = rx[i];
    }
  };
  if (mb < me) {
    stage(mb);
    commit(0);
  }
  __syncthreads();
  int buf = 0;
  for (long long m0 = mb; m0 < me; m0 += TM) {
    const bool more = m0 + TM < me;
    if (more) stage(m0 + TM);
#pragma unroll
    for (int ks = 0; ks < TM / 16; ks++) {
      const int mrow = ks * 16;
      uint32_t a[2][4], b[4][2];
      const int j = lane >> 3, r = lane & 7;
#pragma unroll
      for (int i = 0; i < 2; i++)           // A = dY^T: matrices (n, m), (n+8, m), (n, m+8), (n+8, m+8)
        ldsm_x4_t(a[i][0], a[i][1], a[i][2], a[i][3], &sY[buf][mrow + r + ((j & 2) ? 8 : 0)][wn + i * 16 + ((j & 1) ? 8 : 0)]);
#pragma unroll
      for (int i = 0; i < 2; i++) {         // B = X: (m, k), (m+8, k), (m, k+8), (m+8, k+8)
        uint32_t t0, t1, t2, t3;
        ldsm_x4_t(t0, t1, t2, t3, &sX[buf][mrow + r + ((j & 1) ? 8 : 0)][wk + i * 16 + ((j & 2) ? 8 : 0)]);
        b[2 * i][0] = t0; b[2 * i][1] = t1; b[2 * i + 1][0] = t2; b[2 * i + 1][1] = t3;
      }
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int q = 0; q < 4; q++) mma_bf16(acc[i][q], a[i][0], a[i][1], a[i][2], a[i][3], b[q][0], b[q][1]);
    }
    if (more) commit(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  // partial tile -> ws[z][n][k]
  float* out = p.ws + (size_t)blockIdx.z * p.N * p.K;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int n = n0 + wn + i * 16 + g, k = k0 + wk + q * 8 + 2 * t;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int nn = n + h * 8;
        if (nn < p.N) {
          if (k < p.K) out[(size_t)nn * p.K + k] = acc[i][q][2 * h];
          if (k + 1 < p.K) out[(size_t)nn * p.K + k + 1] = acc[i][q][2 * h + 1];
        }
      }
    }
}

// dw[n, ci, t] = sum_s ws[s][n][t*Cin + ci]
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float* __restrict__ ws, int S, int N, int K, int T, float* dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  float s = 0.f;
  for (int z = 0; z < S; z++) s += ws[(size_t)z * N * K + i];
  const int n = i / K, k = i % K, Cin = K / T;
  const int tt = k / Cin, ci = k % Cin;
  dw[(size_t)n * K + ci * T + tt] = s;
}

// T == 1 (plain GEMM weights, no tap permutation), N*K % 4 == 0: four outputs per thread with 16-byte loads / stores
__global__ void __launch_bounds__(256) wgrad_finalize4_kernel(const float4* __restrict__ ws, int S, long long nk4, float4* dw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nk4) return;
  float4 s = ws[i];
#pragma unroll 4
  for (int z = 1; z < S; z++) {
    const float4 v = ws[(size_t)z * nk4 + i];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  dw[i] = s;
}

// column sums with optional per-row scale (bias gradients)
template <typename T, int V>
__global__ void __launch_bounds__(kColThreads) colsum_partial_kernel(const T* __restrict__ x, long long ld, long long rows, int C,
                                                                     const float* __restrict__ rs, int rs_div, int ngrp, int nrl,
                                                                     int rows_per_block, float* __restrict__ ws) {
  __shared__ float smem[V * kColThreads];
  const int grp = threadIdx.x % ngrp, rl = threadIdx.x / ngrp;
  const int c0 = (blockIdx.y * ngrp + grp) * V;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s1[V];
#pragma unroll
  for (int v = 0; v < V; v++) s1[v] = 0.f;
  if (c0 < C) {
#pragma unroll 4                     // independent row loads in flight (these passes are bound by bytes in flight)
    for (long long r = r0 + rl; r < r1; r += nrl) {
      float xv[V];
      ldv<V>(x + r * ld + c0, xv);
      const float sc = rs ? rs[r / rs_div] : 1.f;
#pragma unroll
      for (int v = 0; v < V; v++) s1[v] = fmaf(xv[v], sc, s1[v]);
    }
  }
  col_block_reduce<V>(s1, smem, grp, rl, ngrp, nrl);
  if (rl == 0 && c0 < C) {
#pragma unroll
    for (int v = 0; v < V; v++)
      if (c0 + v < C) ws[(size_t)blockIdx.x * C + c0 + v] = s1[v];
  }
}

// ---- deferred, batched reduction of weight-gradient partials --------------------------------------------------------------
// One launch reduces the partials of MANY weight-gradient GEMMs (a whole gradient bucket): job j owns the blocks
// [blk0_j, blk0_{j+1}).  A block is 256 threads = `sl` split lanes x 256/sl outputs; lane l adds the partials z = l, l+sl, ...
// and the lanes are then added in lane order (fixed summation order -> bit-reproducible, no float atomics).
__host__ __device__ inline int wgrad_reduce_lanes(int S) { return S <= 8 ? 1 : (S <= 64 ? 8 : 32); }
__host__ __device__ inline bool wgrad_reduce_vec(const cenet_wgrad_job& j) {
  return j.T == 1 && (j.K & 3) == 0 && (j.src_ld & 3) == 0 && (j.stride & 3) == 0 && ((((uintptr_t)j.src) | ((uintptr_t)j.dst)) & 15) == 0;
}
__global__ void __launch_bounds__(256) wgrad_reduce_batch_kernel(const cenet_wgrad_job* __restrict__ jobs, int njobs) {
  __shared__ int sj;
  __shared__ float4 red[256];
  if (threadIdx.x == 0) {
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {                                  // last job whose first block is <= blockIdx.x
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].blk0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    sj = lo;
  }
  __syncthreads();
  const cenet_wgrad_job j = jobs[sj];
  const int lb = blockIdx.x - j.blk0;
  const int sl = wgrad_reduce_lanes(j.S), per = 256 / sl;
  const int o = threadIdx.x % per, lane = threadIdx.x / per;
  const long long nk = (long long)j.N * j.K;
  const int sld = j.src_ld > 0 ? j.src_ld : j.K;          // the N x K block may sit inside wider partial rows (block-diagonal GEMMs)
  if (wgrad_reduce_vec(j)) {
    const long long i = (long long)lb * per + o, nk4 = nk >> 2, st4 = j.stride >> 2;
    const float4* src = reinterpret_cast<const float4*>(j.src);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nk4) {
      const long long e0 = i * 4;
      const size_t so = sld == j.K ? (size_t)i : (size_t)(((e0 / j.K) * sld + (e0 % j.K)) >> 2);
#pragma unroll 4
      for (int z = lane; z < j.S; z += sl) {
        const float4 v = src[(size_t)z * st4 + so];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    }
    if (sl > 1) {
      red[threadIdx.x] = a;
      __syncthreads();
      if (lane == 0) {
        for (int l = 1; l < sl; l++) {
          const float4 v = red[l * per + o];
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
      }
    }
    if (lane == 0 && i < nk4) reinterpret_cast<float4*>(j.dst)[i] = a;
  } else {
    const long long i = (long long)lb * per + o;
    float a = 0.f;
    if (i < nk) {
      const size_t so = sld == j.K ? (size_t)i : (size_t)((i / j.K) * sld + (i % j.K));
#pragma unroll 4
      for (int z = lane; z < j.S; z += sl) a += j.src[(size_t)z * j.stride + so];
    }
    if (sl > 1) {
      float* r1 = reinterpret_cast<float*>(red);
      r1[threadIdx.x] = a;
      __syncthreads();
      if (lane == 0)
        for (int l = 1; l < sl; l++) a += r1[l * per + o];
    }
    if (lane == 0 && i < nk) {
      const int n = (int)(i / j.K), k = (int)(i % j.K), Cin = j.K / j.T;
      const int tt = k / Cin, ci = k % Cin;
      j.dst[(size_t)n * j.K + (size_t)ci * j.T + tt] = a;            // k = (tap, ci) -> the reference's [Cout, Cin, KH, KW]
    }
  }
}
}  // namespace

static inline void launch_wgrad_finalize(const float* ws, int S, int N, int K, int T, float* dw, cudaStream_t s) {
  const long long nk = (long long)N * K;
  if (T == 1 && nk % 4 == 0 && ((((uintptr_t)ws | (uintptr_t)dw) & 15) == 0))
    wgrad_finalize4_kernel<<<cdiv(nk / 4, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(ws), S, nk / 4, reinterpret_cast<float4*>(dw));
  else
    wgrad_finalize_kernel<<<cdiv(nk, 256), 256, 0, s>>>(ws, S, N, K, T, dw);
}

int launch_colsum(const void* x, int dtype, long long ld, long long rows, int C, const float* rs, int rs_div, float* out, float* ws,
                  long long ws_elems, cudaStream_t s) {
  CENET_DISPATCH(dtype, T, {
    int Vv = pick_vec({C, ld});
    long long al = ptr_align_elems(x, sizeof(T));
    while (Vv > al) Vv >>= 1;
    if (sizeof(T) == 4 && Vv > 4) Vv = 4;
    ColPlan p = plan_cols(rows, C, Vv);
    CENET_REQUIRE((long long)p.nrb * C <= ws_elems, "colsum: workspace too small");
    if (Vv == 8) colsum_partial_kernel<T, 8><<<dim3(p.nrb, p.gy), kColThreads, 0, s>>>((const T*)x, ld, rows, C, rs, rs_div, p.ngrp, p.nrl, p.rows_per_block, ws);
    else if (Vv == 4) colsum_partial_kernel<T, 4><<<dim3(p.nrb, p.gy), kColThreads, 0, s>>>((const T*)x, ld, rows, C, rs, rs_div, p.ngrp, p.nrl, p.rows_per_block, ws);
    else if (Vv == 2) colsum_partial_kernel<T, 2><<<dim3(p.nrb, p.gy), kColThreads, 0, s>>>((const T*)x, ld, rows, C, rs, rs_div, p.ngrp, p.nrl, p.rows_per_block, ws);
    else colsum_partial_kernel<T, 1><<<dim3(p.nrb, p.gy), kColThreads, 0, s>>>((const T*)x, ld, rows, C, rs, rs_div, p.ngrp, p.nrl, p.rows_per_block, ws);
    CENET_LAUNCH_CHECK("colsum_partial");
    return launch_finalize(ws, p.nrb, C, out, C, nullptr, 1.f, s);
  });
  return 0;
}

bool cenet_wgrad_tc_eligible(const void* dy, int dy_dtype, long long ldy, const void* x, int x_dtype, long long ldx, long long M, int N,
                             int K, const float* rs, int rs_div);
void cenet_wgrad_tc_plan(long long M, int N, int K, bool has_rs, int rs_div, bool binary, long long max_partials, cenet_wgrad_plan* pl);
int cenet_wgrad_tc_launch(const cenet_wgrad_plan* pl, const void* dy, long long ldy, const void* x, long long ldx, int N, int K,
                          const float* rs, bool binary, float* out, long long out_stride, float* bias, long long bias_stride,
                          cudaStream_t s);

bool cenet_conv_wgrad_tc_eligible(const void* dy, int dy_dtype, long long ldy, const void* x, int x_dtype, long long ldx, int Cin, int N,
                                  int ksize);
int cenet_conv_wgrad_tc(const void* dy, const void* x, int B, int H, int W, int Cin, int ksize, int N, float* ws, long long ws_elems,
                        cudaStream_t s);

// mma.sync path: writes S partials [N][K] into p.ws and returns S
static int launch_wgrad_mma(WgParams p, long long ws_elems, cudaStream_t s) {
  const long long nk = (long long)p.N * p.K;
  const int ntiles = cdiv(p.N, TN) * cdiv(p.K, TK);
  long long want = cdiv(3 * kNumSMs, ntiles);
  long long maxs = std::max<long long>(1, p.M / 256);
  if (want > maxs) want = maxs;
  if (want * nk > ws_elems) want = ws_elems / nk;
  if (want > 65535) want = 65535;
  if (want < 1) want = 1;
  p.m_per_split = ((p.M + want - 1) / want + TM - 1) / TM * TM;
  const int S = (int)((p.M + p.m_per_split - 1) / p.m_per_split);
  dim3 grid(cdiv(p.N, TN), cdiv(p.K, TK), S);
  CENET_REQUIRE(grid.y <= 65535, "wgrad: K too large");
  wgrad_mma_kernel<<<grid, WG_THREADS, 0, s>>>(p);
  CENET_LAUNCH_CHECK("wgrad_mma");
  return S;
}

// weight gradient of a dense stride-1 "same" conv without materialising im2col: x is the NHWC image [B,H,W,Cin] (pitch ldx)
extern "C" int cenet_conv_wgrad(const void* dy, int dy_dtype, long long ldy, const void* x, int x_dtype, long long ldx, int B, int H,
                                int W, int Cin, int ksize, int N, float* dw, float* ws, long long ws_elems, cenet_stream_t st) {
  CENET_REQUIRE(dy && x && dw && ws, "cenet_conv_wgrad: null pointer");
  CENET_REQUIRE(ksize % 2 == 1 && Cin > 0 && N > 0, "cenet_conv_wgrad: bad shape");
  static const bool use_tc = getenv("CENET_B200_WGRAD_TC") == nullptr || atoi(getenv("CENET_B200_WGRAD_TC")) != 0;
  if (use_tc && cenet_conv_wgrad_tc_eligible(dy, dy_dtype, ldy, x, x_dtype, ldx, Cin, N, ksize)) {
    const long long nk = (long long)N * ksize * ksize * Cin;
    const int St = cenet_conv_wgrad_tc(dy, x, B, H, W, Cin, ksize, N, ws, ws_elems, to_stream(st));
    if (St == -1) return -1;
    if (St > 0) {
      launch_wgrad_finalize(ws, St, N, ksize * ksize * Cin, ksize * ksize, dw, to_stream(st));
      CENET_LAUNCH_CHECK("wgrad_finalize");
      return 0;
    }
  }
  WgParams p = {};
  p.dy = dy; p.dy_dtype = dy_dtype; p.ldy = ldy; p.x = x; p.x_dtype = x_dtype; p.ldx = ldx;
  p.M = (long long)B * H * W; p.N = N; p.K = ksize * ksize * Cin; p.rs = nullptr; p.rs_div = 1; p.ws = ws;
  p.fast_y = dy_dtype == CENET_BF16 && ldy % 8 == 0 && ((uintptr_t)dy & 15) == 0;
  p.fast_x = x_dtype == CENET_BF16 && ldx % 8 == 0 && ((uintptr_t)x & 15) == 0;
  p.conv = 1; p.H = H; p.W = W; p.Cin = Cin; p.KS = ksize; p.pad = ksize / 2;
  CENET_REQUIRE((long long)N * p.K <= ws_elems, "cenet_conv_wgrad: workspace too small");
  const int S = launch_wgrad_mma(p, ws_elems, to_stream(st));
  if (S < 0) return -1;
  launch_wgrad_finalize(ws, S, N, p.K, ksize * ksize, dw, to_stream(st));
  CENET_LAUNCH_CHECK("wgrad_finalize");
  return 0;
}

// Partial products of one weight-gradient GEMM.  On return *n_partials == 0 means dw (and dbias) already hold the final
// result; otherwise ws holds S = *n_partials partial matrices [N][K] (S*N*K floats), followed -- when *bias_partials != 0 -- by
// S partial bias rows [N]; they are reduced by cenet_wgrad_reduce_batch (any number of GEMMs in one launch) or, in
// cenet_gemm_wgrad, immediately.
extern "C" int cenet_gemm_wgrad_partial(const void* dy, int dy_dtype, long long ldy, const void* x, int x_dtype, long long ldx,
                                        long long M, int N, int K, int T, const float* row_scale, int rs_div, int rs_binary,
                                        float* dw, float* dbias, int bias_unscaled, float* ws, long long ws_elems, int* n_partials,
                                        int* bias_partials, cenet_stream_t st) {
  CENET_REQUIRE(dy && x && ws && n_partials && bias_partials, "cenet_gemm_wgrad: null pointer");
  CENET_REQUIRE(M > 0 && N > 0 && K > 0 && T >= 1 && K % T == 0, "cenet_gemm_wgrad: bad shape M=%lld N=%d K=%d T=%d", M, N, K, T);
  CENET_REQUIRE(rs_div >= 1, "cenet_gemm_wgrad: rs_div must be >= 1");
  cudaStream_t s = to_stream(st);
  *n_partials = 0; *bias_partials = 0;
  const long long nk = (long long)N * K;
  CENET_REQUIRE(nk <= ws_elems, "cenet_gemm_wgrad: workspace too small (N*K = %lld)", nk);
  static const bool use_tc = getenv("CENET_B200_WGRAD_TC") == nullptr || atoi(getenv("CENET_B200_WGRAD_TC")) != 0;
  const bool tc = use_tc && cenet_wgrad_tc_eligible(dy, dy_dtype, ldy, x, x_dtype, ldx, M, N, K, row_scale, rs_div);
  cenet_wgrad_plan pl = {};
  const bool bias_in_tc = tc && dbias && !(bias_unscaled && row_scale);
  bool tc_ok = false;
  if (tc) {
    cenet_wgrad_tc_plan(M, N, K, row_scale != nullptr, rs_div, rs_binary != 0, ws_elems / (nk + (bias_in_tc ? N : 0)), &pl);
    tc_ok = (long long)pl.S * (nk + (bias_in_tc ? N : 0)) <= ws_elems;      // (a per-group plan cannot go below one split per group)
  }
  if (dbias && !(tc_ok && bias_in_tc)) {
    if (launch_colsum(dy, dy_dtype, ldy, M, N, bias_unscaled ? nullptr : row_scale, rs_div, dbias, ws, ws_elems, s)) return -1;
  }
  if (tc_ok) {
    const bool direct = pl.S == 1 && T == 1 && dw != nullptr;     // dw == NULL: the caller wants the partials (block extraction)
    float* out = direct ? dw : ws;
    float* bout = bias_in_tc ? (direct ? dbias : ws + (size_t)pl.S * nk) : nullptr;
    if (cenet_wgrad_tc_launch(&pl, dy, ldy, x, ldx, N, K, row_scale, rs_binary != 0, out, nk, bout, N, s)) return -1;
    if (!direct) { *n_partials = pl.S; *bias_partials = bias_in_tc ? 1 : 0; }
    return 0;
  }
  int S;
  if (dy_dtype == CENET_F32 && x_dtype == CENET_F32) {
    // validation precision: CUDA-core GEMM  C_z[N,K] = A_z^T W_z  over row chunks, z = split
    S = 1;
    for (int c = 2; c <= 64; c++)
      if (M % c == 0 && (long long)c * nk <= ws_elems && M / c >= 64) S = c;
    const long long chunk = M / S;
    cenet_gemm_args g = {};
    g.M = N; g.N = K; g.K = (int)chunk; g.batch = S; g.batch_inner = 1;
    g.A = dy; g.a_dtype = dy_dtype; g.lda = ldy; g.a_bs_outer = chunk * ldy; g.a_mmajor = 1;
    g.Wt = x; g.w_dtype = x_dtype; g.ldw = ldx; g.w_bs_outer = chunk * ldx; g.w_nmajor = 1;
    g.C = ws; g.c_dtype = CENET_F32; g.ldc = K; g.c_bs_outer = nk;
    g.alpha = 1.f; g.res1_scale = 1.f; g.rs_div = 1; g.post_rs_div = 1;
    g.k_scale = row_scale; g.k_scale_div = rs_div; g.k_scale_bs = chunk;
    if (cenet_gemm_simt(&g, s)) return -1;
  } else {
    WgParams p = {};
    p.dy = dy; p.dy_dtype = dy_dtype; p.ldy = ldy; p.x = x; p.x_dtype = x_dtype; p.ldx = ldx;
    p.M = M; p.N = N; p.K = K; p.rs = row_scale; p.rs_div = rs_div;
    p.ws = ws;
    p.fast_y = dy_dtype == CENET_BF16 && ldy % 8 == 0 && ((uintptr_t)dy & 15) == 0;
    p.fast_x = x_dtype == CENET_BF16 && ldx % 8 == 0 && ((uintptr_t)x & 15) == 0;
    S = launch_wgrad_mma(p, ws_elems, s);
    if (S < 0) return -1;
  }
  *n_partials = S;
  return 0;
}

extern "C" int cenet_gemm_wgrad(const void* dy, int dy_dtype, long long ldy, const void* x, int x_dtype, long long ldx, long long M,
                                int N, int K, int T, const float* row_scale, int rs_div, float* dw, float* dbias,
                                int bias_unscaled, float* ws, long long ws_elems, cenet_stream_t st) {
  int S = 0, bp = 0;
  CENET_REQUIRE(dw, "cenet_gemm_wgrad: null pointer");
  if (cenet_gemm_wgrad_partial(dy, dy_dtype, ldy, x, x_dtype, ldx, M, N, K, T, row_scale, rs_div, 0, dw, dbias, bias_unscaled, ws,
                               ws_elems, &S, &bp, st))
    return -1;
  if (S == 0) return 0;
  cudaStream_t s = to_stream(st);
  launch_wgrad_finalize(ws, S, N, K, T, dw, s);
  CENET_LAUNCH_CHECK("wgrad_finalize");
  if (bp) return launch_finalize(ws + (size_t)S * N * K, S, N, dbias, N, nullptr, 1.f, s);
  return 0;
}

// out[c] = sum_r rs[r / rs_div] * x[r, c]   (bias gradients; with rs = a one-channel input image it is the weight gradient of a
// 1x1 conv with Cin = 1, unet.py:205-207 at input_channels = 1)
extern "C" int cenet_colsum(const void* x, int dtype, long long ld, long long rows, int C, const float* row_scale, int rs_div, float* out,
                            float* ws, long long ws_elems, cenet_stream_t st) {
  CENET_REQUIRE(x && out && ws && rs_div >= 1, "cenet_colsum: bad arguments");
  return launch_colsum(x, dtype, ld, rows, C, row_scale, rs_div, out, ws, ws_elems, to_stream(st));
}

// number of 256-thread blocks job j needs in cenet_wgrad_reduce_batch (the caller lays out blk0 with it)
extern "C" int cenet_wgrad_reduce_blocks(const cenet_wgrad_job* j) {
  const long long nk = (long long)j->N * j->K;
  const int per = 256 / wgrad_reduce_lanes(j->S);
  return cdiv(wgrad_reduce_vec(*j) ? nk / 4 : nk, per);
}

// jobs: DEVICE array of njobs descriptors whose blk0 fields are the running sum of cenet_wgrad_reduce_blocks; nblocks = the total
extern "C" int cenet_wgrad_reduce_batch(const cenet_wgrad_job* jobs, int njobs, int nblocks, cenet_stream_t st) {
  CENET_REQUIRE(jobs && njobs > 0 && nblocks > 0, "cenet_wgrad_reduce_batch: empty job list");
  wgrad_reduce_batch_kernel<<<nblocks, 256, 0, to_stream(st)>>>(jobs, njobs);
  CENET_LAUNCH_CHECK("wgrad_reduce_batch");
  return 0;
}
