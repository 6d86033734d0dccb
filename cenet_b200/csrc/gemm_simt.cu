// CUDA-core GEMM / implicit-GEMM convolution (fp32 accumulate), any dtype / layout the C ABI allows.
// Used for (a) the fp32 validation precision, (b) shapes the tcgen05 kernel does not take (K % 8 != 0, e.g. the
// first 5x5 conv with Cin = 1, K = 25), (c) cross-checking the tcgen05 kernel in tests.
#include "common.cuh"

namespace {
constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct SimtParams {
  int M, N, K;
  int batch_inner;
  const void* A; int a_dtype; long long lda, a_bso, a_bsi; int a_mmajor;
  const float* kscale; int kscale_div; long long kscale_bs;
  int conv, H, W, Cin, KH, KW, stride, pad, Ho, Wo;
  const void* Wt; int w_dtype; long long ldw, w_bso, w_bsi; int w_nmajor;
  long long c_bso, c_bsi;
  EpiParams epi;
};

__global__ void __launch_bounds__(NT) gemm_simt_kernel(const SimtParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int z = blockIdx.z, zo = z / p.batch_inner, zi = z % p.batch_inner;
  const long long aoff = zo * p.a_bso + zi * p.a_bsi;
  const long long woff = zo * p.w_bso + zi * p.w_bsi;
  const long long coff = zo * p.c_bso + zi * p.c_bsi;

  // A loader: thread loads rows (tid/16 + 16*i), column tid%16 of the BMxBK tile
  const int lk = tid % 16, lr = tid / 16;
  long long arow_base[4];
  int a_h0[4], a_w0[4];
  bool arow_ok[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int m = m0 + lr + 16 * i;
    arow_ok[i] = m < p.M;
    if (p.conv) {
      int mm = arow_ok[i] ? m : 0;
      int wo = mm % p.Wo, t = mm / p.Wo, ho = t % p.Ho, b = t / p.Ho;
      a_h0[i] = ho * p.stride - p.pad;
      a_w0[i] = wo * p.stride - p.pad;
      arow_base[i] = (long long)b * p.H * p.W * p.lda;   // lda = channel pitch of the NHWC image
    } else {
      arow_base[i] = (long long)m * p.lda;
      a_h0[i] = a_w0[i] = 0;
    }
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    const int k = k0 + lk;
    int kh = 0, kw = 0, ci = k;
    if (p.conv && k < p.K) { ci = k % p.Cin; int t = k / p.Cin; kw = t % p.KW; kh = t / p.KW; }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float v = 0.f;
      if (arow_ok[i] && k < p.K) {
        if (p.conv) {
          int h = a_h0[i] + kh, w = a_w0[i] + kw;
          if (h >= 0 && h < p.H && w >= 0 && w < p.W)
            v = ld_any(p.A, p.a_dtype, aoff + arow_base[i] + ((long long)h * p.W + w) * p.lda + ci);
        } else if (p.a_mmajor) {
          v = ld_any(p.A, p.a_dtype, aoff + (long long)k * p.lda + (m0 + lr + 16 * i));
          if (p.kscale) v *= p.kscale[((long long)z * p.kscale_bs + k) / p.kscale_div];
        } else {
          v = ld_any(p.A, p.a_dtype, aoff + arow_base[i] + k);
        }
      }
      As[lk][lr + 16 * i] = v;
    }
    if (!p.w_nmajor) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int n = n0 + lr + 16 * i;
        float v = 0.f;
        if (n < p.N && k < p.K) v = ld_any(p.Wt, p.w_dtype, woff + (long long)n * p.ldw + k);
        Bs[lk][lr + 16 * i] = v;
      }
    } else {
      // W stored [K,N]: thread loads k = tid/64 + 4*i, n = tid%64 (coalesced along n)
      const int nn = tid % 64, kk0 = tid / 64;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int kk = kk0 + 4 * i, n = n0 + nn;
        float v = 0.f;
        if (n < p.N && k0 + kk < p.K) v = ld_any(p.Wt, p.w_dtype, woff + (long long)(k0 + kk) * p.ldw + n);
        Bs[kk][nn] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      epi_store(p.epi, epi_value(p.epi, acc[i][j], m, n, coff), m, n, coff);
    }
  }
}
}  // namespace

int cenet_gemm_simt(const cenet_gemm_args* a, cudaStream_t s) {
  SimtParams p;
  p.M = a->M; p.N = a->N; p.K = a->K; p.batch_inner = a->batch_inner;
  p.A = a->A; p.a_dtype = a->a_dtype; p.lda = a->lda; p.a_bso = a->a_bs_outer; p.a_bsi = a->a_bs_inner;
  p.a_mmajor = a->a_mmajor; p.kscale = a->k_scale; p.kscale_div = a->k_scale_div > 0 ? a->k_scale_div : 1; p.kscale_bs = a->k_scale_bs;
  CENET_REQUIRE(!(a->a_mmajor && a->conv), "cenet_gemm_simt: a_mmajor and conv are exclusive");
  p.conv = a->conv; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.KH = a->KH; p.KW = a->KW; p.stride = a->stride;
  p.pad = a->pad; p.Ho = a->Ho; p.Wo = a->Wo;
  p.Wt = a->Wt; p.w_dtype = a->w_dtype; p.ldw = a->ldw; p.w_bso = a->w_bs_outer; p.w_bsi = a->w_bs_inner;
  p.w_nmajor = a->w_nmajor;
  p.c_bso = a->c_bs_outer; p.c_bsi = a->c_bs_inner;
  p.epi = make_epi(a);
  dim3 grid(cdiv(a->M, BM), cdiv(a->N, BN), a->batch);   // M tiles on x (2^31 limit)
  CENET_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "cenet_gemm_simt: grid too large (N=%d batch=%d)", a->N, a->batch);
  gemm_simt_kernel<<<grid, NT, 0, s>>>(p);
  CENET_LAUNCH_CHECK("gemm_simt");
  return 0;
}
