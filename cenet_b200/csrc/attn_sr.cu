// Encoder spatial-reduction attention (pvtv2.py:88-105): every query attends to the reduced keys (49 at 224x224, 256 at
// 512x512), head_dim 64.  bf16 tensors take the mma.sync flash kernel of attn_flash.cu (64-key tiles, K/V of an (image,
// head) staged in smem); the CUDA-core kernel below serves the fp32 validation precision and walks the keys in tiles of 64.
// AI ~ 49 FLOP/B -> HBM/L2-bound: K and V of one (image, head) live in shared memory (25 KB fp32), one thread owns
// one query row and streams the keys with an online softmax, so q is read once and o written once.
#include "common.cuh"

int cenet_sr_attention_mma(const void* q, const void* kv, void* out, int B, int N, int Nk, int C, int heads, float scale,
                           cudaStream_t s);

namespace {
constexpr int HD = 64, MAXK = 64, QT = 128;

template <typename TQ, typename TKV, typename TO>
__global__ void __launch_bounds__(QT) sr_attention_kernel(const TQ* __restrict__ q, const TKV* __restrict__ kv,
                                                          TO* __restrict__ out, int N, int Nk, int C, float scale) {
  __shared__ __align__(16) float Ks[MAXK][HD];
  __shared__ __align__(16) float Vs[MAXK][HD];
  const int head = blockIdx.y, b = blockIdx.z;
  const TKV* kvb = kv + (long long)b * Nk * 2 * C + head * HD;
  const int n = blockIdx.x * QT + threadIdx.x;
  const bool live = n < N;                       // tail threads still help staging the key tiles
  const TQ* qp = q + ((long long)b * N + (live ? n : 0)) * C + head * HD;
  float qr[HD], o[HD];
#pragma unroll
  for (int d = 0; d < HD; d += 8) {
    float t[8];
    ldv<8>(qp + d, t);
#pragma unroll
    for (int i = 0; i < 8; i++) { qr[d + i] = t[i] * scale; o[d + i] = 0.f; }
  }
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < Nk; j0 += MAXK) {
    const int nj = min(MAXK, Nk - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < nj * (HD / 2); i += QT) {
      const int j = i / (HD / 2), d = (i % (HD / 2)) * 2;
      float t[2];
      ldv<2>(kvb + (long long)(j0 + j) * 2 * C + d, t);
      Ks[j][d] = t[0]; Ks[j][d + 1] = t[1];
      ldv<2>(kvb + (long long)(j0 + j) * 2 * C + C + d, t);
      Vs[j][d] = t[0]; Vs[j][d + 1] = t[1];
    }
    __syncthreads();
    for (int j = 0; j < nj; j++) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(&Ks[j][d]);
        s = fmaf(qr[d], k4.x, s); s = fmaf(qr[d + 1], k4.y, s); s = fmaf(qr[d + 2], k4.z, s); s = fmaf(qr[d + 3], k4.w, s);
      }
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn), p = __expf(s - mn);
      l = l * corr + p;
      m = mn;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(&Vs[j][d]);
        o[d] = fmaf(p, v4.x, o[d] * corr); o[d + 1] = fmaf(p, v4.y, o[d + 1] * corr);
        o[d + 2] = fmaf(p, v4.z, o[d + 2] * corr); o[d + 3] = fmaf(p, v4.w, o[d + 3] * corr);
      }
    }
  }
  if (!live) return;
  const float inv = 1.f / l;
  TO* op = out + ((long long)b * N + n) * C + head * HD;
#pragma unroll
  for (int d = 0; d < HD; d += 8) {
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = o[d + i] * inv;
    stv<8>(op + d, t);
  }
}
}  // namespace

extern "C" int cenet_sr_attention(const void* q, int q_dtype, const void* kv, int kv_dtype, void* out, int o_dtype,
                                  int B, int N, int Nk, int C, int heads, float scale, cenet_stream_t s) {
  if (B == 0 || N == 0) return 0;
  CENET_REQUIRE(q && kv && out, "cenet_sr_attention: null pointer");
  CENET_REQUIRE(C == heads * HD, "cenet_sr_attention: head_dim must be 64 (C=%d heads=%d)", C, heads);
  CENET_REQUIRE(Nk >= 1, "cenet_sr_attention: needs at least one key, got %d", Nk);
  CENET_REQUIRE(q_dtype == kv_dtype && q_dtype == o_dtype, "cenet_sr_attention: q/kv/out must share one dtype");
  CENET_REQUIRE(B <= 65535 && heads <= 65535, "cenet_sr_attention: grid too large");
  if (q_dtype == CENET_BF16 && (((uintptr_t)q | (uintptr_t)kv | (uintptr_t)out) & 15) == 0)
    return cenet_sr_attention_mma(q, kv, out, B, N, Nk, C, heads, scale, to_stream(s));   // tensor-core path (attn_flash.cu)
  dim3 grid(cdiv(N, QT), heads, B);
  CENET_DISPATCH(q_dtype, T, (sr_attention_kernel<T, T, T><<<grid, QT, 0, to_stream(s)>>>((const T*)q, (const T*)kv, (T*)out, N, Nk, C, scale)));
  CENET_LAUNCH_CHECK("sr_attention");
  return 0;
}
