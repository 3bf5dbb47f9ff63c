// Reconstruction loss (cliora/net/trainer.py:46-78): for every word (b,i) the scores
//   s_0 = pos[b,i] . cell[b,i],   s_{1+e} = neg[e] . cell[b,i]   (e < K negatives)
// and CE(s, target 0), fused: one warp per word, the K negative vectors staged through shared memory in
// chunks of 16, scores owned by lane e % 32 (K + 1 <= 128).  Lane l owns columns 4l + 128t (D <= 512).
#pragma once
#include "cell_warp_kernels.cuh"

namespace cliora {

constexpr int kNegChunk = 16;

// rowloss[row] = logsumexp(s) - s_0;  P[row, 0..K] = softmax(s)
__global__ __launch_bounds__(256) void recon_ce_fwd_kernel(int rows, int D, int K, const float* __restrict__ cell,
                                                           const float* __restrict__ pos,
                                                           const float* __restrict__ neg, float* __restrict__ rowloss,
                                                           float* __restrict__ P) {
  pdl_prologue();
  extern __shared__ __align__(16) float s_neg[];   // [kNegChunk][D]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = blockIdx.x * 8 + warp;
  const bool active = row < rows;
  float4 c[kColT];
  float sc[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // scores e = lane + 32*i
  if (active) {
    float d = 0.f;
#pragma unroll
    for (int t = 0; t < kColT; ++t) {
      const int j = lane * 4 + t * 128;
      c[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < D) {
        c[t] = ld4(cell + (int64_t)row * D + j);
        d += dot4(c[t], ld4(pos + (int64_t)row * D + j));
      }
    }
    d = warp_sum(d);
    if (lane == 0) sc[0] = d;
  }
  for (int e0 = 0; e0 < K; e0 += kNegChunk) {
    const int ne = min(kNegChunk, K - e0);
    __syncthreads();
    for (int i = tid * 4; i < ne * D; i += 1024) st4(s_neg + i, ld4(neg + (int64_t)e0 * D + i));
    __syncthreads();
    if (active) {
      for (int e = 0; e < ne; ++e) {
        float d = 0.f;
#pragma unroll
        for (int t = 0; t < kColT; ++t) {
          const int j = lane * 4 + t * 128;
          if (j < D) d += dot4(c[t], ld4(s_neg + e * D + j));
        }
        d = warp_sum(d);
        const int idx = 1 + e0 + e;
        if ((idx & 31) == lane) sc[idx >> 5] = d;
      }
    }
  }
  if (!active) return;
  float mx = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
  mx = warp_max(mx);
  float ex[4], sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ex[i] = (lane + 32 * i <= K) ? expf(sc[i] - mx) : 0.f;
    sum += ex[i];
  }
  sum = warp_sum(sum);
  const float s0 = __shfl_sync(0xffffffffu, sc[0], 0);
  if (lane == 0) rowloss[row] = mx + logf(sum) - s0;
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (lane + 32 * i <= K) P[(int64_t)row * (K + 1) + lane + 32 * i] = ex[i] * inv;
}

// G[row, e] = (P[row, e] - [e == 0]) * scale;  g_cell[row] = G[row,0] pos[row] + sum_e G[row,1+e] neg[e];
// g_pos[row] = G[row,0] cell[row]
__global__ __launch_bounds__(256) void recon_ce_bwd_kernel(int rows, int D, int K, const float* __restrict__ cell,
                                                           const float* __restrict__ pos,
                                                           const float* __restrict__ neg, const float* __restrict__ P,
                                                           const float* __restrict__ gloss, float inv_rows,
                                                           float* __restrict__ G, float* __restrict__ g_cell,
                                                           float* __restrict__ g_pos) {
  pdl_prologue();
  extern __shared__ __align__(16) float s_neg[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = blockIdx.x * 8 + warp;
  const bool active = row < rows;
  const float scale = gloss[0] * inv_rows;
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  float4 acc[kColT];
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = lane + 32 * i;
      if (e <= K) {
        g[i] = (P[(int64_t)row * (K + 1) + e] - (e == 0 ? 1.f : 0.f)) * scale;
        G[(int64_t)row * (K + 1) + e] = g[i];
      }
    }
    const float g0 = __shfl_sync(0xffffffffu, g[0], 0);
#pragma unroll
    for (int t = 0; t < kColT; ++t) {
      const int j = lane * 4 + t * 128;
      acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < D) {
        const float4 pv = ld4(pos + (int64_t)row * D + j), cv = ld4(cell + (int64_t)row * D + j);
        acc[t] = make_float4(g0 * pv.x, g0 * pv.y, g0 * pv.z, g0 * pv.w);
        st4(g_pos + (int64_t)row * D + j, make_float4(g0 * cv.x, g0 * cv.y, g0 * cv.z, g0 * cv.w));
      }
    }
  }
  for (int e0 = 0; e0 < K; e0 += kNegChunk) {
    const int ne = min(kNegChunk, K - e0);
    __syncthreads();
    for (int i = tid * 4; i < ne * D; i += 1024) st4(s_neg + i, ld4(neg + (int64_t)e0 * D + i));
    __syncthreads();
    if (active) {
      for (int e = 0; e < ne; ++e) {
        const int idx = 1 + e0 + e;
        const int q = idx >> 5;
        const float own = q == 0 ? g[0] : q == 1 ? g[1] : q == 2 ? g[2] : g[3];
        const float w = __shfl_sync(0xffffffffu, own, idx & 31);
#pragma unroll
        for (int t = 0; t < kColT; ++t) {
          const int j = lane * 4 + t * 128;
          if (j < D) fma4(acc[t], w, ld4(s_neg + e * D + j));
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int t = 0; t < kColT; ++t) {
    const int j = lane * 4 + t * 128;
    if (j < D) st4(g_cell + (int64_t)row * D + j, acc[t]);
  }
}

}  // namespace cliora
