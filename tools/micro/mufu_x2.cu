// MUFU.EX2 throughput of the PACKED half-precision forms (ex2.approx.ftz.bf16x2 / f16x2: two exponentials per instruction?) against
// the fp32 form.  SASS (cuobjdump, nvcc 12.9, sm_100a): ex2.approx.ftz.bf16x2 / ex2.approx.f16x2 compile to TWO scalar
// `MUFU.EX2.BF16 Rd, Rs` / `MUFU.EX2.BF16 Rd, Rs.H1` (resp. .F16) plus a PRMT -- there is no packed MUFU, so the packed forms cannot
// raise the 16 exponentials per clock and SM; measured on the B200: f32 15.9, bf16x2 15.7, f16x2 15.7 exponentials per clock and SM.  It decides whether a packed-bf16 softmax exponent could beat the 16 ex2/clk/SM of MUFU.EX2 (tools/micro/mufu.cu)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned ex2_bf16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned ex2_f16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float ex2_f32(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int ILP, int MODE>
__global__ void k(unsigned* out, int iters) {
  unsigned v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) v[i] = 0xbf00bf00u + threadIdx.x + i;      // about -0.5 in both halves (bf16 / f16 bit patterns differ; only timing matters)
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (MODE == 0) v[i] = ex2_bf16x2(v[i]) ^ 0x80008000u;                // keep the argument negative: result in (0.5, 1]
      else if (MODE == 1) v[i] = ex2_f16x2(v[i]) ^ 0x80008000u;
      else v[i] = __float_as_uint(ex2_f32(__uint_as_float(v[i]))) ^ 0x80000000u;
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int MODE>
void run(int blocks_per_sm, int threads, const char* name) {
  unsigned* out; cudaMalloc(&out, 148 * 16 * 1024 * 4);
  int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP, MODE><<<148 * blocks_per_sm, threads>>>(out, 16);
  cudaEventRecord(e0);
  k<ILP, MODE><<<148 * blocks_per_sm, threads>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)148 * blocks_per_sm * threads * iters * ILP * (MODE == 2 ? 1 : 2);
  printf("%s ILP %d warps/SM %d: %.2f G exponentials/s = %.2f per clk per SM @1.965GHz (%s)\n", name, ILP, blocks_per_sm * threads / 32,
         n / ms / 1e6, n / ms / 1e6 / 148 / 1.965, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  run<8, 2>(16, 128, "f32    "); run<8, 0>(16, 128, "bf16x2 "); run<8, 1>(16, 128, "f16x2  ");
  run<8, 0>(8, 128, "bf16x2 "); run<16, 0>(8, 128, "bf16x2 ");
  return 0;
}
