// PROTOTYPE for round 2 -- NOT part of libcliora_b200.so, never run on hardware yet (written after the round's
// GPU budget was spent; it is compile-checked only:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC \
//        -o dev/proto/libproto_f32a.so dev/proto/tc_gemm_f32a.cu
// and dev/proto/test_f32a.py is the harness to run first thing next round).
//
// Idea: the 3xTF32 compose GEMM is bound by the bytes an SM ingests per k-block
// (52 KB: hi+lo of a 128x32 A tile and of an 80x32 W tile; measured 0.57-0.64 us per k-block whether 1 or 125
// CTAs run).  The activations (A) do not have to arrive as a split pair: TMA can bring the plain fp32 tile
// (16 KB instead of 32 KB) and the four epilogue warps, idle during the main loop, can split it in shared
// memory.  The split is element-wise, so it is layout-agnostic under the 128-byte swizzle: read 16 bytes at
// offset o of the stage's A buffer, write hi back to offset o and lo to offset o of the A_lo buffer.
// Weights stay pre-split (they are small and shared by every CTA).
//
// Pipeline per stage s:  producer  --TMA-->  full[s]  --converters (128 thr)-->  conv[s]  --MMA-->  empty[s]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../cliora_b200/csrc/tc_gemm.cuh"

namespace cliora {
// symbols the headers expect from api.cu
thread_local char g_last_cuda_error[256] = "";
long long g_launch_count = 0;
Profiler g_prof;
int g_pdl = 0;
int g_carveout = -1;
int g_splitk_target = 4 * 148;
void apply_carveout(const void*) {}
namespace tc {
int g_tc_small_tmem = 0, g_tc_narrow_stages = 3, g_tc_xnarrow = 0;

CL_D void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
CL_D void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

template <int BLOCK_N, int STAGES>
struct F32aSmem {
  static constexpr int A_BYTES = kBlockM * 128;    // one 128 x 32 fp32 tile
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // A_hi (raw lands here), A_lo, W_hi, W_lo
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (3 * STAGES + 1) * 8 + 16 + 1024;
};

// C[M,N] = A[M,K] (plain fp32) @ W[N,K]^T (split pair), fp32-grade accuracy (cross terms in a second accumulator)
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_nt_f32a_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       float* __restrict__ C, int64_t ldc, int M, int N, int K) {
  using S = F32aSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* conv = full + STAGES;
  uint64_t* empty = conv + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBlockM, n0 = blockIdx.x * BLOCK_N;
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  const int n_cur = min(BLOCK_N, ((N - n0 + 15) / 16) * 16);
  constexpr uint32_t TMEM_COLS = tmem_cols(2 * BLOCK_N);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&conv[i], 128);     // every converter thread arrives once per use of the stage
        mbar_init(&empty[i], 1);
      }
      mbar_init(tmem_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], S::A_BYTES + 2 * S::B_BYTES);
        uint8_t* s = smem + stage * S::STAGE_BYTES;
        tma_load_2d(s, &tmA, &full[stage], kb * kBlockK, m0);                            // plain fp32 A tile
        tma_load_3d(s + 2 * S::A_BYTES, &tmB, &full[stage], kb * kBlockK, n0, 0);        // W hi + lo in one box
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(n_cur);
      uint32_t acc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&conv[stage], phase);       // implies full[stage]: the converters waited on it
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
        const uint64_t a_hi = umma_desc_k_sw128(sa), a_lo = umma_desc_k_sw128(sa + S::A_BYTES);
        const uint64_t b_hi = umma_desc_k_sw128(sa + 2 * S::A_BYTES);
        const uint64_t b_lo = umma_desc_k_sw128(sa + 2 * S::A_BYTES + S::B_BYTES);
#pragma unroll
        for (int k = 0; k < kBlockK / 8; ++k) {
          umma_tf32(tmem_base + BLOCK_N, a_lo + 2 * k, b_hi + 2 * k, idesc, acc);
          umma_tf32(tmem_base + BLOCK_N, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
          umma_tf32(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);
          acc = 1;
        }
        umma_commit(&empty[stage]);
      }
      umma_commit(tmem_full);
    }
  } else {
    // ---- warps 2-5: first the in-place operand split of every k-block, then the epilogue ----
    const int ct = threadIdx.x - 64;          // 0 .. 127
    for (int kb = 0; kb < num_kb; ++kb) {
      const int stage = kb % STAGES;
      const uint32_t phase = (kb / STAGES) & 1;
      mbar_wait(&full[stage], phase);         // TMA data (generic-proxy visible after the barrier completes)
      uint8_t* a_hi = smem + stage * S::STAGE_BYTES;
      uint8_t* a_lo = a_hi + S::A_BYTES;
#pragma unroll
      for (int i = 0; i < S::A_BYTES / 16 / 128; ++i) {
        const int off = (ct + i * 128) * 16;
        const float4 x = *reinterpret_cast<const float4*>(a_hi + off);
        float4 hi, lo;
        split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(a_hi + off) = hi;
        *reinterpret_cast<float4*>(a_lo + off) = lo;
      }
      fence_proxy_async();                    // generic-proxy writes -> visible to the UMMA (async proxy) reads
      mbar_arrive(&conv[stage]);
    }
    const int q = warp & 3;
    mbar_wait(tmem_full, 0);
    tcgen05_fence_after();
    const int r = m0 + q * 32 + lane;
    const bool row_ok = r < M;
    float* crow = row_ok ? C + (int64_t)r * ldc : nullptr;
#pragma unroll 1
    for (int c = 0; c < n_cur; c += 16) {
      float v[16], x[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BLOCK_N + c), x);
      if (!row_ok) continue;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const int col = n0 + c + j;
        if (col >= N) break;     // N % 4 == 0
        st4(crow + col, make_float4(v[j] + x[j], v[j + 1] + x[j + 1], v[j + 2] + x[j + 2], v[j + 3] + x[j + 3]));
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// 2-D tensor map over plain fp32 [rows, K] (row pitch ld floats), box {32, 128}, 128-byte swizzle
inline int make_plain_map(CUtensorMap* tm, const float* base, int64_t rows, int K, int64_t ld) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return CLIORA_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kBlockM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CLIORA_OK : CLIORA_ERR_CUDA;
}
}  // namespace tc
}  // namespace cliora

using namespace cliora;

// C[M,N] = A[M,K] @ W[N,K]^T with A plain fp32 and W a split pair [2, N, K]
extern "C" int proto_tc_linear_f32a(int M, int N, int K, const float* A, const float* W_pair, float* C, void* stream) {
  constexpr int BN = 80, ST = 4;
  using S = tc::F32aSmem<BN, ST>;
  if (M < 1 || N < 4 || N % 4 || K < 4 || K % 4) return CLIORA_ERR_BAD_SHAPE;
  CUtensorMap tmA, tmB;
  CL_TRY(tc::make_plain_map(&tmA, A, M, K, K));
  CL_TRY(tc::make_pair_map(&tmB, W_pair, N, K, K, (int64_t)N * K, BN, CU_TENSOR_MAP_SWIZZLE_128B, 2));
  CL_CUDA(cudaFuncSetAttribute(tc::tc_gemm_nt_f32a_kernel<BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               S::TOTAL));
  dim3 grid(ceil_div(N, BN), ceil_div(M, tc::kBlockM));
  tc::tc_gemm_nt_f32a_kernel<BN, ST><<<grid, tc::kThreads, S::TOTAL, (cudaStream_t)stream>>>(tmA, tmB, C, N, M, N, K);
  return cudaPeekAtLastError() == cudaSuccess ? CLIORA_OK : CLIORA_ERR_CUDA;
}
