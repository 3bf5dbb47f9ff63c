// CKY argmax decode over the inside split scores (cliora/analysis/cky.py:31-99 fed by the hook of
// cliora/analysis/utils.py:78-95).  One block per sentence; the Viterbi chart lives in shared memory;
// the level loop runs inside the kernel (the reference does one device->host sync per cell).
#pragma once
#include "common.cuh"

namespace cliora {

// E: raw inside split scores, level blocks [B, L, N] at row offset B * inside_rows_before(n, level).
// best[l,p] = max_k best[k,p] + best[l-1-k,p+k+1] + (e_k - max_k e_k); leaves = 1; first max wins.
// Dynamic smem: cells floats.
__global__ __launch_bounds__(64) void cky_kernel(int B, int n, const float* __restrict__ E,
                                                 int32_t* __restrict__ backptr, float* __restrict__ best_out) {
  pdl_prologue();
  extern __shared__ float s_best[];
  const int b = blockIdx.x;
  const int C = (int)num_cells(n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s_best[i] = 1.f;
    backptr[(int64_t)b * C + i] = -1;
  }
  __syncthreads();
  for (int level = 1; level < n; ++level) {
    const int L = n - level, N = level;
    const int64_t lvl_rows = (int64_t)B * inside_rows_before(n, level);
    for (int p = threadIdx.x; p < L; p += blockDim.x) {
      const float* e = E + lvl_rows + ((int64_t)b * L + p) * N;
      float mx = e[0];
      for (int k = 1; k < N; ++k) mx = fmaxf(mx, e[k]);
      float bv = -INFINITY;
      int bk = 0;
      for (int k = 0; k < N; ++k) {
        int l, r;
        inside_children(n, level, p, k, l, r);
        const float lr = __fadd_rn(s_best[l], s_best[r]);
        const float cand = __fadd_rn(lr, __fsub_rn(e[k], mx));
        if (cand > bv) { bv = cand; bk = k; }
      }
      const int cell = lvl_off(n, level) + p;
      s_best[cell] = bv;
      backptr[(int64_t)b * C + cell] = bk;
    }
    __syncthreads();
  }
  if (best_out != nullptr)
    for (int i = threadIdx.x; i < C; i += blockDim.x) best_out[(int64_t)b * C + i] = s_best[i];
}

}  // namespace cliora
