// Fused optimiser step (cliora/net/trainer.py:450-455: clip_grad_norm_(5.0) then Adam.step) as two
// multi-tensor kernels over a device-side table of (param, grad, exp_avg, exp_avg_sq, numel) entries:
//   1. adam_gradnorm_kernel: per-block partial sums of grad^2 over all tensors (deterministic two-stage reduce)
//   2. adam_update_kernel  : clip coefficient = min(1, max_norm / (norm + 1e-6)), Adam with bias correction
// The step count lives on the device (incremented by kernel 2) so the pair can sit inside a CUDA graph.
#pragma once
#include "common.cuh"

namespace cliora {

struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  int64_t n;
  int64_t block0;   // first block of this tensor in the flattened block list
};

constexpr int kAdamBlock = 256;
constexpr int kAdamPerThread = 4;
constexpr int kAdamChunk = kAdamBlock * kAdamPerThread;

CL_D int find_tensor(const AdamTensor* __restrict__ t, int nt, int64_t blk) {
  int lo = 0, hi = nt - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (t[mid].block0 <= blk) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

__global__ __launch_bounds__(kAdamBlock) void adam_gradnorm_kernel(const AdamTensor* __restrict__ tab, int nt,
                                                                   float* __restrict__ partial) {
  pdl_prologue();
  __shared__ float red[64];
  const int ti = find_tensor(tab, nt, blockIdx.x);
  const AdamTensor t = tab[ti];
  const int64_t base = ((int64_t)blockIdx.x - t.block0) * kAdamChunk;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kAdamPerThread; ++i) {
    const int64_t idx = base + (int64_t)i * kAdamBlock + threadIdx.x;
    if (idx < t.n) {
      const float g = t.g[idx];
      s = fmaf(g, g, s);
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// one block: total = sum(partial); writes state[0] = ||g||, state[1] = clip coefficient, state[2] += 1 (step)
__global__ void adam_norm_finish_kernel(const float* __restrict__ partial, int nblocks, float max_norm,
                                        float* __restrict__ state) {
  pdl_prologue();
  __shared__ float red[64];
  float s = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) s += partial[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float norm = sqrtf(s);
    state[0] = norm;
    state[1] = fminf(1.f, max_norm / (norm + 1e-6f));   // torch.nn.utils.clip_grad_norm_
    state[2] += 1.f;
  }
}

__global__ __launch_bounds__(kAdamBlock) void adam_update_kernel(const AdamTensor* __restrict__ tab, int nt, float lr,
                                                                 float beta1, float beta2, float eps,
                                                                 const float* __restrict__ state) {
  pdl_prologue();
  const int ti = find_tensor(tab, nt, blockIdx.x);
  const AdamTensor t = tab[ti];
  const int64_t base = ((int64_t)blockIdx.x - t.block0) * kAdamChunk;
  const float clip = state[1], step = state[2];
  const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
#pragma unroll
  for (int i = 0; i < kAdamPerThread; ++i) {
    const int64_t idx = base + (int64_t)i * kAdamBlock + threadIdx.x;
    if (idx < t.n) {
      const float g = t.g[idx] * clip;
      const float m = fmaf(beta1, t.m[idx], (1.f - beta1) * g);
      const float v = fmaf(beta2, t.v[idx], (1.f - beta2) * g * g);
      t.m[idx] = m;
      t.v[idx] = v;
      t.p[idx] -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
    }
  }
}

}  // namespace cliora
