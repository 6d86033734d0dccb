// Shared helpers of the training kernels: two-stage (deterministic) column reductions over [rows, C] matrices.
//
// Stage 1: grid (row blocks, channel-group tiles); a block is 256 threads = `ngrp` channel groups x `nrl` row lanes,
// each thread owns V consecutive channels and walks its rows with stride nrl, accumulating K quantities in registers;
// the row lanes are then summed through shared memory and the block writes ONE partial per (k, channel) into the
// workspace: ws[(rb * K + k) * C + c].  Stage 2 (`finalize`) sums the partials in fixed order -> bit-reproducible,
// no float atomics.
#pragma once
#include "common.cuh"

constexpr int kColThreads = 256;

// tile / split plan of the tcgen05 weight-gradient kernel (wgrad_tc.cu)
struct cenet_wgrad_plan {
  int bn, S, per_group, parts, cps, cpg, groups, rows_per_group, total_chunks;
};

struct ColPlan {
  int V, ngrp, nrl, nrb, rows_per_block, gy;
};

static inline ColPlan plan_cols(long long rows, int C, int V, int max_blocks = 4 * kNumSMs) {
  ColPlan p;
  p.V = V;
  const int groups = (C + V - 1) / V;
  int ngrp = 1;
  while (ngrp < groups && ngrp < 32) ngrp <<= 1;
  p.ngrp = ngrp;
  p.nrl = kColThreads / ngrp;
  p.gy = (groups + ngrp - 1) / ngrp;
  long long want = (rows + 4LL * p.nrl - 1) / (4LL * p.nrl);       // >= 4 rows per thread
  long long cap = max_blocks / p.gy;
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  p.rows_per_block = (int)((rows + want - 1) / want);
  p.nrb = (int)((rows + p.rows_per_block - 1) / p.rows_per_block);
  return p;
}

// sum `acc[V]` over the row lanes of the block; lanes with rl == 0 get the total.  smem: V * 256 floats.
template <int V>
__device__ __forceinline__ void col_block_reduce(float (&acc)[V], float* smem, int grp, int rl, int ngrp, int nrl) {
  __syncthreads();
#pragma unroll
  for (int v = 0; v < V; v++) smem[(v * nrl + rl) * ngrp + grp] = acc[v];
  __syncthreads();
  if (rl == 0) {
#pragma unroll
    for (int v = 0; v < V; v++) {
      float s = 0.f;
      for (int l = 0; l < nrl; l++) s += smem[(v * nrl + l) * ngrp + grp];
      acc[v] = s;
    }
  }
}

// Stage 2 helper.  A finalize block is 256 threads = kFinOut outputs x kFinLanes lanes (o = tid % kFinOut, lane = tid /
// kFinOut): lane l adds the partials b = l, l + kFinLanes, ... (independent loads, 4 in flight), the lanes are then added
// in lane order through shared memory.  Fixed summation order -> bit-reproducible; the serial chain per output is
// nblk / kFinLanes long instead of nblk.  K quantities at once; the result is valid in the lane-0 threads.
constexpr int kFinOut = 8, kFinLanes = 32, kFinThreads = kFinOut * kFinLanes;
template <int K, typename Acc, typename F>
__device__ __forceinline__ void fin_lane_sums(int nblk, bool valid, Acc (&tot)[K], Acc* smem /* [K][kFinLanes][kFinOut] */, F&& load) {
  const int o = threadIdx.x % kFinOut, lane = threadIdx.x / kFinOut;
  Acc s[K];
#pragma unroll
  for (int k = 0; k < K; k++) s[k] = (Acc)0;
  if (valid) {
#pragma unroll 4
    for (int b = lane; b < nblk; b += kFinLanes) {
#pragma unroll
      for (int k = 0; k < K; k++) s[k] += (Acc)load(b, k);
    }
  }
#pragma unroll
  for (int k = 0; k < K; k++) smem[(k * kFinLanes + lane) * kFinOut + o] = s[k];
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      Acc t = (Acc)0;
      for (int l = 0; l < kFinLanes; l++) t += smem[(k * kFinLanes + l) * kFinOut + o];
      tot[k] = t;
    }
  }
}

// out[i] = scale * sum_b ws[b * n + i]   (i < n); the first nA go to outA, the rest to outB
__global__ void __launch_bounds__(kFinThreads) finalize_partials_kernel(const float* __restrict__ ws, int nblk, int n, float* outA,
                                                                        int nA, float* outB, float scale);

int launch_finalize(const float* ws, int nblk, int n, float* outA, int nA, float* outB, float scale, cudaStream_t s);

__device__ __forceinline__ float act_grad_from_out(float y, int act, float slope) {
  if (act == CENET_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == CENET_ACT_LEAKY) return y > 0.f ? 1.f : slope;
  return 1.f;
}
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * __expf(-0.5f * x * x) * 0.39894228040143267794f;
}
