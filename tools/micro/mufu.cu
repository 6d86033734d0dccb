// MUFU.EX2 throughput microbenchmark: per-thread chains of independent ex2 (with an FFMA feeding each, like softmax)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int ILP, bool FMA>
__global__ void k(float* out, int iters, float a, float b) {
  float v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) v[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = ex2(FMA ? fmaf(v[i], a, b) : v[i]);
  }
  float s = 0; 
#pragma unroll
  for (int i = 0; i < ILP; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, bool FMA>
void run(int blocks_per_sm, int threads) {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP, FMA><<<148 * blocks_per_sm, threads>>>(out, 16, 0.5f, -1.f);
  cudaEventRecord(e0);
  k<ILP, FMA><<<148 * blocks_per_sm, threads>>>(out, iters, 0.5f, -1.f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double n = (double)148 * blocks_per_sm * threads * iters * ILP;
  printf("ILP %d fma %d warps/SM %d: %.2f Gex2/s  = %.2f ex2/clk/SM @1.965GHz\n", ILP, (int)FMA, blocks_per_sm * threads / 32,
         n / ms / 1e6, n / ms / 1e6 / 148 / 1.965);
  cudaFree(out);
}
int main() {
  run<8, true>(4, 128); run<8, true>(8, 128); run<8, true>(16, 128);
  run<8, false>(4, 128); run<8, false>(16, 128);
  run<16, true>(4, 128); run<4, true>(16, 128);
  return 0;
}
