// Shared device/host helpers for libcenet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>
#include "../../include/cenet_b200.h"

typedef __nv_bfloat16 bf16;

// ---- host-side error plumbing -----------------------------------------------------------------------------
void cenet_set_error(const char* fmt, ...);
extern std::atomic<long long> g_cenet_launches;

#define CENET_FAIL(...)              \
  do {                               \
    cenet_set_error(__VA_ARGS__);    \
    return -1;                       \
  } while (0)

#define CENET_REQUIRE(cond, ...)     \
  do {                               \
    if (!(cond)) CENET_FAIL(__VA_ARGS__); \
  } while (0)

// Called after every launch: counts it and surfaces launch-configuration errors (never synchronises).
#define CENET_LAUNCH_CHECK(name)                                              \
  do {                                                                        \
    g_cenet_launches.fetch_add(1, std::memory_order_relaxed);                 \
    cudaError_t e__ = cudaPeekAtLastError();                                  \
    if (e__ != cudaSuccess) {                                                 \
      cudaGetLastError();                                                     \
      CENET_FAIL("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
    }                                                                         \
  } while (0)

// dtype dispatch: binds T to float / bf16
#define CENET_DISPATCH(dtype, T, ...)                                  \
  do {                                                                 \
    if ((dtype) == CENET_F32) { typedef float T; __VA_ARGS__; }        \
    else if ((dtype) == CENET_BF16) { typedef bf16 T; __VA_ARGS__; }   \
    else CENET_FAIL("bad dtype code %d", (int)(dtype));                \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline cudaStream_t to_stream(cenet_stream_t s) { return (cudaStream_t)s; }
constexpr int kNumSMs = 148;

// ---- device helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// V-wide vector load/store (V in {1,2,4,8}); pointers must be aligned to V*sizeof(T)
template <int V>
__device__ __forceinline__ void ldv(const float* p, float (&o)[V]) {
  if constexpr (V == 8) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  } else if constexpr (V == 4) {
    float4 a = *reinterpret_cast<const float4*>(p);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  } else if constexpr (V == 2) {
    float2 a = *reinterpret_cast<const float2*>(p);
    o[0] = a.x; o[1] = a.y;
  } else {
    o[0] = *p;
  }
}
template <int V>
__device__ __forceinline__ void ldv(const bf16* p, float (&o)[V]) {
  if constexpr (V == 8) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  } else if constexpr (V == 4) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 2; i++) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  } else if constexpr (V == 2) {
    float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
    o[0] = f.x; o[1] = f.y;
  } else {
    o[0] = __bfloat162float(*p);
  }
}
template <int V>
__device__ __forceinline__ void stv(float* p, const float (&o)[V]) {
  if constexpr (V == 8) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
  } else if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  } else if constexpr (V == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(o[0], o[1]);
  } else {
    *p = o[0];
  }
}
template <int V>
__device__ __forceinline__ void stv(bf16* p, const float (&o)[V]) {
  if constexpr (V == 8) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  } else if constexpr (V == 4) {
    uint2 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 2; i++) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint2*>(p) = u;
  } else if constexpr (V == 2) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(o[0], o[1]);
  } else {
    *p = __float2bfloat16_rn(o[0]);
  }
}

// largest V in {8,4,2,1} such that every listed quantity is a multiple of V (element counts / pitches / offsets)
static inline int pick_vec(std::initializer_list<long long> qs) {
  for (int v : {8, 4, 2}) {
    bool ok = true;
    for (long long q : qs) ok = ok && (q % v == 0);
    if (ok) return v;
  }
  return 1;
}
// byte alignment of a pointer expressed in elements of size es
static inline long long ptr_align_elems(const void* p, int es) {
  uintptr_t u = (uintptr_t)p;
  long long a = 16;
  while (a > 1 && (u % a)) a >>= 1;
  long long e = a / es;
  return e < 1 ? 1 : e;
}
static inline int dtype_size(int dt) { return dt == CENET_F32 ? 4 : 2; }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case CENET_ACT_GELU: return gelu_erf(v);
    case CENET_ACT_RELU: return fmaxf(v, 0.0f);
    case CENET_ACT_LEAKY: return v > 0.0f ? v : v * slope;
    case CENET_ACT_SILU: return v * sigmoidf_(v);
    case CENET_ACT_SIGMOID: return sigmoidf_(v);
    case CENET_ACT_GELU_GRAD:   // d gelu(v) / dv (training: fused into the dgrad GEMM that produces d(pre-activation))
      return 0.5f * (1.0f + erff(v * 0.70710678118654752440f)) + v * __expf(-0.5f * v * v) * 0.39894228040143267794f;
    default: return v;
  }
}

// Activation of 8 values with the dispatch OUTSIDE the element loop: a per-element `switch` serialises the eight dependent
// chains (MUFU + division latency each) -- measured on the GEMM epilogue: SiLU cost 60 us per 12.8 M elements that way.
// bf16 / tensor-core paths only: sigmoid uses the approximate division (2 ulp), far below the output rounding.
// GELU'(x) = Phi(x) + x phi(x) with erf from Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7): its exp(-(x/sqrt2)^2) IS the
// exp(-x^2/2) of the density term, so the whole derivative costs one EX2, one RCP and ~12 FMA-pipe instructions instead of
// erff (~25 instructions with a branch) plus expf.  The dgrad epilogue of the Mix-FFN fc2 applies it to 38 M elements per call.
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float ax = fabsf(x);
  const float E = __expf(-0.5f * x * x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.0f));
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  const float erfa = fmaf(-q * t, E, 1.0f);                 // erf(|x| / sqrt 2)
  return fmaf(0.5f, copysignf(erfa, x), 0.5f) + x * E * 0.39894228040143267794f;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ void apply_act8(float (&v)[8], int act, float slope) {
  switch (act) {
    case CENET_ACT_RELU:
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j], 0.0f);
      break;
    case CENET_ACT_LEAKY:
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = v[j] > 0.0f ? v[j] : v[j] * slope;
      break;
    case CENET_ACT_SILU:
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = v[j] * sigmoid_fast(v[j]);
      break;
    case CENET_ACT_SIGMOID:
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = sigmoid_fast(v[j]);
      break;
    case CENET_ACT_GELU:
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = gelu_erf(v[j]);
      break;
    case CENET_ACT_GELU_GRAD:
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = gelu_grad_fast(v[j]);
      break;
    default: break;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// PyTorch's bilinear source index (aten UpSample.h area_pixel_compute_source_index, align_corners=False)
__device__ __forceinline__ void bilin_src(int d, float scale, int in_size, int& i0, int& i1, float& l1) {
  float src = scale * (d + 0.5f) - 0.5f;
  src = src < 0.0f ? 0.0f : src;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
}
// align_corners=True variant: src = d * (in-1)/(out-1)
__device__ __forceinline__ void bilin_src_ac(int d, float scale, int in_size, int& i0, int& i1, float& l1) {
  float src = scale * d;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
}

// ---- shared epilogue of the SIMT and tcgen05 GEMMs ----------------------------------------------------------
struct EpiParams {
  float alpha;
  const float* bias;
  int bias_per_row;
  const float* row_scale;
  int rs_div;                         // row_scale index = m / rs_div (per-sample scales)
  const float* post_rs; int post_rs_div;   // applied after bias/act/mul, before the residuals (DropPath)
  int act;
  float slope;
  int act_after_res;
  const void* res1; int res1_dtype; long long ldr1; const float* res1_cscale; float res1_scale;
  const void* res2; int res2_dtype; long long ldr2;
  const void* mul; int mul_dtype; long long ldmul; int mul_act;
  void* C; int c_dtype; long long ldc;
};

__device__ __forceinline__ float ld_any(const void* p, int dtype, long long idx) {
  return dtype == CENET_F32 ? reinterpret_cast<const float*>(p)[idx]
                            : __bfloat162float(reinterpret_cast<const bf16*>(p)[idx]);
}

// one element; `coff` = batch offset (elements) applied to C / res / mul
__device__ __forceinline__ float epi_value(const EpiParams& e, float acc, long long m, int n, long long coff) {
  float v = e.alpha * acc;
  if (e.row_scale) v *= e.row_scale[m / e.rs_div];
  if (e.bias) v += e.bias_per_row ? e.bias[m] : e.bias[n];
  if (!e.act_after_res) v = apply_act(v, e.act, e.slope);
  if (e.mul) v *= apply_act(ld_any(e.mul, e.mul_dtype, coff + m * e.ldmul + n), e.mul_act, 0.f);
  if (e.post_rs) v *= e.post_rs[m / e.post_rs_div];
  if (e.res1) v += ld_any(e.res1, e.res1_dtype, coff + m * e.ldr1 + n) * (e.res1_cscale ? e.res1_cscale[n] : e.res1_scale);
  if (e.res2) v += ld_any(e.res2, e.res2_dtype, coff + m * e.ldr2 + n);
  if (e.act_after_res) v = apply_act(v, e.act, e.slope);
  return v;
}
__device__ __forceinline__ void epi_store(const EpiParams& e, float v, long long m, int n, long long coff) {
  long long idx = coff + m * e.ldc + n;
  if (e.c_dtype == CENET_F32) reinterpret_cast<float*>(e.C)[idx] = v;
  else reinterpret_cast<bf16*>(e.C)[idx] = __float2bfloat16_rn(v);
}

static inline EpiParams make_epi(const cenet_gemm_args* a) {
  EpiParams e;
  e.alpha = a->alpha; e.bias = a->bias; e.bias_per_row = a->bias_per_row; e.row_scale = a->row_scale;
  e.rs_div = a->rs_div > 0 ? a->rs_div : 1; e.post_rs = a->post_row_scale; e.post_rs_div = a->post_rs_div > 0 ? a->post_rs_div : 1;
  e.act = a->act; e.slope = a->slope; e.act_after_res = a->act_after_res;
  e.res1 = a->res1; e.res1_dtype = a->res1_dtype; e.ldr1 = a->ldr1; e.res1_cscale = a->res1_cscale;
  e.res1_scale = a->res1_scale;
  e.res2 = a->res2; e.res2_dtype = a->res2_dtype; e.ldr2 = a->ldr2;
  e.mul = a->mul; e.mul_dtype = a->mul_dtype; e.ldmul = a->ldmul; e.mul_act = a->mul_act;
  e.C = a->C; e.c_dtype = a->c_dtype; e.ldc = a->ldc;
  return e;
}

// internal entry points (one per translation unit)
int cenet_gemm_simt(const cenet_gemm_args* a, cudaStream_t s);
int cenet_gemm_tc(const cenet_gemm_args* a, cudaStream_t s);
bool cenet_gemm_tc_eligible(const cenet_gemm_args* a);
int cenet_gemm_mma(const cenet_gemm_args* a, cudaStream_t s);
bool cenet_gemm_mma_eligible(const cenet_gemm_args* a);

// ---- cp.async (LDGSTS): 16-byte global -> shared copies that cost no registers while in flight; src_bytes = 0 zero-fills
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(a), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2: two channels per instruction)
typedef unsigned long long f32x2;                              // two packed floats (lo = even channel)

__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// bf16x2 word -> packed floats (exact): low half << 16, high half masked
__device__ __forceinline__ f32x2 bf2_to_f2(unsigned w) { return pk2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ unsigned f2_to_bf2(f32x2 v) {
  float a, b; upk2(v, a, b);
  unsigned r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));    // first source -> upper half
  return r;
}

