// Batch assembly from the region-feature table (the reference does this per example on host workers:
// FlickrDataset.__getitem__, cliora/data/dataloader.py:205-222, then collate + .cuda(),
// cliora/data/batch_iterator.py:116-168).  The table is ragged: image i owns rows [pos[i,0], pos[i,1]) of
// `features` [total_rows, F] / `bboxes` [total_rows, 4] / `classes` [total_rows]; a batch entry takes the first
// min(rows, R) of them and pads the rest (features 0, boxes -1, classes -1).
// The table pointer may be device memory (table resident in HBM) or pinned host memory (zero-copy reads over
// the host link) - same kernel.  One block per output row (b, r); 16-byte vector loads/stores.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace cliora {

template <typename T>
__global__ __launch_bounds__(128) void gather_regions_kernel(int B, int R, int F, const T* __restrict__ features,
                                                             const float* __restrict__ bboxes,
                                                             const int32_t* __restrict__ classes,
                                                             const int64_t* __restrict__ pos,
                                                             const int64_t* __restrict__ img_index,
                                                             float* __restrict__ obj_feats, float* __restrict__ boxes,
                                                             int64_t* __restrict__ obj_cates) {
  pdl_prologue();
  const int row = blockIdx.x;          // b * R + r
  const int b = row / R, r = row - b * R;
  const int64_t img = img_index[b];
  const int64_t s = pos[2 * img], e = pos[2 * img + 1];
  const bool live = (int64_t)r < e - s;
  const int64_t src = s + r;
  float* dst = obj_feats + (int64_t)row * F;
  if (sizeof(T) == 4) {
    const float* f = reinterpret_cast<const float*>(features) + src * F;
    for (int j = threadIdx.x * 4; j < F; j += blockDim.x * 4)
      st4(dst + j, live ? ld4(f + j) : make_float4(0.f, 0.f, 0.f, 0.f));
  } else {
    const __half* f = reinterpret_cast<const __half*>(features) + src * F;
    for (int j = threadIdx.x * 8; j < F; j += blockDim.x * 8) {
      float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
      if (live) {
        const uint4 raw = *reinterpret_cast<const uint4*>(f + j);
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
        const float2 a = __half22float2(h[0]), bq = __half22float2(h[1]), c = __half22float2(h[2]),
                     d = __half22float2(h[3]);
        lo = make_float4(a.x, a.y, bq.x, bq.y);
        hi = make_float4(c.x, c.y, d.x, d.y);
      }
      st4(dst + j, lo);
      st4(dst + j + 4, hi);
    }
  }
  if (threadIdx.x < 4 && boxes != nullptr)
    boxes[(int64_t)row * 4 + threadIdx.x] = live ? bboxes[src * 4 + threadIdx.x] : -1.f;
  if (threadIdx.x == 4 && obj_cates != nullptr)
    obj_cates[row] = (live && classes != nullptr) ? (int64_t)classes[src] : (int64_t)-1;
}

}  // namespace cliora
