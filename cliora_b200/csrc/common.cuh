// Shared helpers for the cliora_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <vector>

#include "../../include/cliora_b200.h"

#define CL_HD __host__ __device__ __forceinline__
#define CL_D __device__ __forceinline__

namespace cliora {

constexpr float kTiny = 1e-8f;       // cliora/net/utils.py:10 (UnitNorm clamp)
constexpr float kKeepScale = 1.0f / 0.9f;  // nn.Dropout(0.1) in AttentionHead, cliora/net/cliora.py:32

// ---- chart geometry (closed forms of cliora/net/offset_cache.py, inside_index.py, outside_index.py) ----
CL_HD int64_t num_cells(int n) { return (int64_t)n * (n + 1) / 2; }
CL_HD int lvl_off(int n, int l) { return l * n - l * (l - 1) / 2; }

// Inside split (level, pos p, split k): left child (k, p), right child (level-1-k, p+k+1).
CL_HD void inside_children(int n, int level, int p, int k, int& left, int& right) {
  left = lvl_off(n, k) + p;
  right = lvl_off(n, level - 1 - k) + p + k + 1;
}

// Outside entry k of cell (level, p): first Rn = n-1-level-p entries take the sibling on the right,
// the remaining p entries take it on the left (order of cliora/net/outside_index.py:39-62).
CL_HD void outside_parent_sibling(int n, int level, int p, int k, int& parent, int& sibling) {
  const int Rn = n - 1 - level - p;
  if (k < Rn) {
    sibling = lvl_off(n, Rn - 1 - k) + p + level + 1;
    parent = lvl_off(n, level + Rn - k) + p;
  } else {
    const int j = k - Rn;
    sibling = lvl_off(n, j) + p - 1 - j;
    parent = lvl_off(n, level + j + 1) + p - 1 - j;
  }
}

// number of inside split rows per sentence in levels [1, level)
CL_HD int64_t inside_rows_before(int n, int level) {
  // sum_{j=1}^{m} (n-j) j with m = level-1
  const int64_t m = level - 1;
  return (int64_t)n * m * (m + 1) / 2 - m * (m + 1) * (2 * m + 1) / 6;
}
// number of outside split rows per sentence in levels [0, level): sum_{j<level} (n-j)(n-j-1)
CL_HD int64_t outside_rows_before(int n, int level) {
  int64_t s = 0;
  for (int j = 0; j < level; ++j) s += (int64_t)(n - j) * (n - j - 1);
  return s;
}

// Row r of a "virtual" [M, ld] matrix lives at physical row (r / L) * bstride + base + (r % L).
// Lets a GEMM read / write one chart level ([B, L] cells inside [B, cells]) in place.
struct RowMap {
  int L;
  int64_t bstride;
  int64_t base;
};
CL_HD RowMap dense_rows() { return RowMap{1 << 30, 0, 0}; }
CL_HD RowMap level_rows(int n, int level) { return RowMap{n - level, num_cells(n), lvl_off(n, level)}; }
CL_HD int64_t map_row(const RowMap& m, int r) { return (int64_t)(r / m.L) * m.bstride + m.base + (r % m.L); }

// ---- warp / block reductions ----
CL_D float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
CL_D float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Sum over the whole block; every thread gets the result.  `red` is >= 33 floats of shared memory.
CL_D float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// x = hi + lo exactly, hi = x rounded to tf32 (low 13 mantissa bits zero)
CL_D void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}

CL_D float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
CL_D void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// 16-byte vector reduction to global memory (red.global.add.v4.f32, sm_90+)
CL_D void red_add4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// ---- asynchronous global -> shared copies (16 bytes, L1-bypassing) ----
CL_D void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
CL_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
CL_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N>
CL_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- programmatic dependent launch (PDL) ----
// First statement of every kernel: let the next kernel in the stream get scheduled while this one runs, then
// wait until the previous kernel has completed and flushed its memory.  All global accesses come after the
// wait, so semantics equal plain stream order; what overlaps is launch latency / block scheduling.
CL_D void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- host-side error plumbing ----
struct LaunchCtx {
  cudaStream_t stream;
  int status = CLIORA_OK;
};
extern thread_local char g_last_cuda_error[256];
extern std::atomic<long long> g_launch_count;
extern int g_pdl;   // 1: launch kernels with the programmatic-stream-serialization attribute
extern int g_carveout;   // >= 0: preferred shared-memory carveout (percent) applied to every kernel once
void apply_carveout(const void* kern);
// cudaFuncSetAttribute is per device: raise an attribute of `kern` to at least `value` on the CURRENT device, once
// (the cache is keyed by device, kernel and attribute, so a process that drives several GPUs configures each).
cudaError_t func_attr_at_least(const void* kern, cudaFuncAttribute attr, int value);

inline int record_cuda_error(cudaError_t e, const char* what) {
  snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s", what, cudaGetErrorString(e));
  return CLIORA_ERR_CUDA;
}

#define CL_CHECK_LAUNCH(name)                                       \
  do {                                                              \
    ++::cliora::g_launch_count;                                     \
    cudaError_t e__ = cudaPeekAtLastError();                        \
    if (e__ != cudaSuccess) {                                       \
      cudaGetLastError();                                           \
      return ::cliora::record_cuda_error(e__, name);                \
    }                                                               \
  } while (0)

#define CL_CUDA(call)                                               \
  do {                                                              \
    cudaError_t e__ = (call);                                       \
    if (e__ != cudaSuccess) return ::cliora::record_cuda_error(e__, #call); \
  } while (0)

#define CL_TRY(expr)                 \
  do {                               \
    int s__ = (expr);                \
    if (s__ != CLIORA_OK) return s__; \
  } while (0)

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// kernel launch with the optional PDL attribute (replaces the triple-chevron syntax everywhere)
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  if (g_carveout >= 0) apply_carveout(reinterpret_cast<const void*>(kern));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);   // errors are picked up by CL_CHECK_LAUNCH
}

// ---- optional per-launch CUDA-event profiler (bench.py's roofline pass) ----
struct ProfEntry {
  const char* name;
  cudaEvent_t a, b;
  double flops, bytes;
};
struct Profiler {
  bool on = false;
  std::vector<ProfEntry> entries;
};
extern Profiler g_prof;
// Records an event before and after whatever is launched on `st` inside its lifetime.
struct ProfScope {
  cudaStream_t st;
  int idx = -1;
  ProfScope(cudaStream_t s, const char* name, double flops, double bytes) : st(s) {
    if (!g_prof.on) return;
    ProfEntry e{name, nullptr, nullptr, flops, bytes};
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
    cudaEventRecord(e.a, st);
    g_prof.entries.push_back(e);
    idx = (int)g_prof.entries.size() - 1;
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(g_prof.entries[idx].b, st);
  }
};

}  // namespace cliora
