// Per-volume evaluation tail (utils/metrics_eval.py:53-71, utils_synapse.py:69-84): the reference copies every slice's argmax
// map to the host, resizes it back with scipy `zoom(order=0)` and counts per-class overlaps with medpy on the CPU.  Here the
// whole volume stays on the device: one pass gathers the nearest-neighbour source pixel of every voxel, writes the full-size
// prediction and accumulates the three INTEGER counts medpy's `dc` is made of, per class:
//     |pred == c  AND  label == c|,   |pred == c|,   |label == c|          (dc = 2*I / (P + L))
// Counts are exact integers (block-local shared-memory histograms, then integer atomics: order-independent, deterministic).
#include "common.cuh"

namespace {
constexpr int MAXC = 32;

template <typename LT>
__global__ void __launch_bounds__(256) volume_labels_counts_kernel(const long long* __restrict__ pred_patch, int ph, int pw,
                                                                   const int* __restrict__ iy, const int* __restrict__ ix,
                                                                   const LT* __restrict__ label, long long* __restrict__ pred_out,
                                                                   unsigned long long* __restrict__ counts, int D, int H, int W,
                                                                   int ncls) {
  __shared__ unsigned int h[3 * MAXC];
  for (int i = threadIdx.x; i < 3 * ncls; i += blockDim.x) h[i] = 0u;
  __syncthreads();
  const long long total = (long long)D * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long t = i / W;
    const int y = (int)(t % H);
    const int d = (int)(t / H);
    const int sy = iy[y], sx = ix[x];            // -1: scipy's coordinate fell outside the patch -> constant 0
    const long long p = (sy < 0 || sx < 0) ? 0ll : pred_patch[((long long)d * ph + sy) * pw + sx];
    if (pred_out) pred_out[i] = p;
    if (label) {
      const long long l = (long long)label[i];
      if (p >= 0 && p < ncls) atomicAdd(&h[ncls + (int)p], 1u);
      if (l >= 0 && l < ncls) {
        atomicAdd(&h[2 * ncls + (int)l], 1u);
        if (l == p) atomicAdd(&h[(int)l], 1u);
      }
    }
  }
  __syncthreads();
  if (label)
    for (int i = threadIdx.x; i < 3 * ncls; i += blockDim.x)
      if (h[i]) atomicAdd(&counts[i], (unsigned long long)h[i]);
}
}  // namespace

extern "C" int cenet_volume_labels_counts(const long long* pred_patch, int ph, int pw, const int* iy, const int* ix,
                                          const void* label, int label_kind, long long* pred_out, long long* counts, int D, int H,
                                          int W, int ncls, cenet_stream_t s) {
  CENET_REQUIRE(ncls >= 1 && ncls <= MAXC, "cenet_volume_labels_counts: ncls %d outside [1,%d]", ncls, MAXC);
  CENET_REQUIRE(D > 0 && H > 0 && W > 0 && ph > 0 && pw > 0, "cenet_volume_labels_counts: empty volume");
  CENET_REQUIRE(label == nullptr || counts != nullptr, "cenet_volume_labels_counts: label given without a counts buffer");
  cudaStream_t st = to_stream(s);
  if (label) cudaMemsetAsync(counts, 0, sizeof(long long) * 3 * ncls, st);
  const long long total = (long long)D * H * W;
  const int grid = (int)std::min<long long>(cdiv(total, 256 * 4), kNumSMs * 8);
  unsigned long long* c = reinterpret_cast<unsigned long long*>(counts);
  switch (label_kind) {
    case 0: volume_labels_counts_kernel<float><<<grid, 256, 0, st>>>(pred_patch, ph, pw, iy, ix, (const float*)label, pred_out, c, D, H, W, ncls); break;
    case 1: volume_labels_counts_kernel<long long><<<grid, 256, 0, st>>>(pred_patch, ph, pw, iy, ix, (const long long*)label, pred_out, c, D, H, W, ncls); break;
    case 2: volume_labels_counts_kernel<unsigned char><<<grid, 256, 0, st>>>(pred_patch, ph, pw, iy, ix, (const unsigned char*)label, pred_out, c, D, H, W, ncls); break;
    default: CENET_FAIL("cenet_volume_labels_counts: label_kind %d (0 float32, 1 int64, 2 uint8)", label_kind);
  }
  CENET_LAUNCH_CHECK("volume_labels_counts");
  return 0;
}
