// tcgen05 / TMEM / TMA GEMM for sm_100a with fp32-grade accuracy ("3xTF32").
//
// Every operand is a SPLIT PAIR: two fp32 arrays hi, lo with hi = tf32-rounded x (low 13 mantissa bits
// zero) and lo = x - hi (exact), stored back to back as a [2, rows, K] tensor.  The kernel accumulates
//     lo_a*hi_b + hi_a*lo_b + hi_a*hi_b
// into one fp32 TMEM accumulator with kind::tf32 UMMAs; the dropped lo*lo term and the truncation of lo
// are ~2^-22 relative, so results agree with an fp32 FMA chain to ~1e-6 (single-pass TF32 is 3.5e-3 of
// max on the chart, 35x over the 1e-4 budget -- BASELINE.md section 2).
//
// Structure (one 128 x BLOCK_N output tile per CTA, 192 threads):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 3-D boxes {32 k, rows, part} with 128-byte swizzle
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit frees smem stages
//   warps 2-5: epilogue      -- tcgen05.ld 32x32b from the accumulator, bias / ReLU / mask, global stores
// Layouts: NT (A[m,k], W[n,k] both K-major).  NN is served by passing a pre-transposed weight pair.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_simt.cuh"

namespace cliora {
namespace tc {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;          // fp32 elements = one 128-byte swizzle row
constexpr int kThreads = 192;

CL_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

CL_D void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
CL_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
CL_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
CL_D void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
CL_D void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
CL_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
CL_D void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
CL_D void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

CL_D void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
CL_D void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// K-major operand tile, 128-byte swizzle, rows packed at a 128-byte pitch: SBO = 8 rows * 128 B.
CL_D uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address, 16-byte units
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major; CUTLASS writes 1)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

// instruction descriptor: D=F32, A=B=TF32, both K-major, M=128, N=n
CL_HD uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}

CL_D void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
CL_D void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
CL_D void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct TcEpilogue {
  float* C;
  int64_t ldc;
  RowMap cmap;
  const float* bias;    // [N] or null
  const float* mask;    // dense [M, ldm]: result zeroed where mask <= 0; or null
  int64_t ldm;
  int64_t mask_lo_off;  // mask stored as a split pair: value = mask[i] + mask[i + mask_lo_off] (0: plain)
  const uint32_t* maskbits;  // [M, 16] ReLU bitmask written by split_build (word t*4+comp, bit l <-> col 128t+4l+comp); or null
  float* C_lo;          // optional: also emit the split pair of the result (C gets hi, C_lo gets lo); or null
  int act;              // 0 none, 1 relu, 2 tanh
  int accumulate;       // C += result
  // --- group-max mode (span-region alignment, cliora.py:457 + trainer.py:101): columns are B_img images x R
  // regions; instead of storing C the epilogue writes, per row (a, cell) and image c, the max over that
  // image's R columns and its argmax (first max).  Tiles advance by whole images (n_stride = imgs * R).
  float* gmax;          // [B_sent, B_img, ncell] or null
  int32_t* gargmax;
  int R, ncell, B_img;
  int n_stride;         // columns advanced per blockIdx.x (0: BLOCK_N)
};

CL_HD constexpr int tmem_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

template <int BLOCK_N, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = kBlockM * 128;
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
};

// C[m0.., n0..] = epi( sum_k A[a_row0 + m, k] * W[n, k] ),  A pair [2, a_rows_total, K], W pair [2, N, K]
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const TcEpilogue ep, int a_row0, int M, int N, int K, int mode_in) {
  pdl_prologue();
  const int mode = mode_in & 0xff;              // bit 8: allocate only the TMEM columns the mode needs
  const bool small_tmem = (mode_in & 0x100) != 0;
  using S = TcSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBlockM, n0 = blockIdx.x * (ep.n_stride > 0 ? ep.n_stride : BLOCK_N);
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  // columns this CTA really owns, rounded up to the UMMA granularity (N % 16 == 0 for M = 128)
  const int n_cur = min(BLOCK_N, ((N - n0 + 15) / 16) * 16);
  // accumulator sets: kSets x (main, cross); consecutive k-steps rotate over the sets so back-to-back UMMAs
  // never depend on each other's TMEM write-back, and each accumulator sees 1/kSets of the truncating adds
  constexpr int kSets = (6 * BLOCK_N <= 512) ? 3 : ((4 * BLOCK_N <= 512) ? 2 : 1);
  // TMEM columns: (main, cross) accumulators; the rotating-set mode 3 needs kSets of them.  Allocating only what the
  // mode uses lets two CTAs (e.g. one per sentence chain) share an SM's 512 columns.
  const uint32_t TMEM_COLS = (mode == 3 || !small_tmem) ? tmem_cols(2 * kSets * BLOCK_N) : tmem_cols(2 * BLOCK_N);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full[i], 1);
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
        mbar_expect_tx(&full[stage], mode == 1 ? S::A_BYTES + S::B_BYTES : S::STAGE_BYTES);
        uint8_t* s = smem + stage * S::STAGE_BYTES;
        // each box holds the hi tile followed by the lo tile (1-part boxes in single-pass mode)
        tma_load_3d(s, &tmA, &full[stage], kb * kBlockK, a_row0 + m0, 0);
        tma_load_3d(s + 2 * S::A_BYTES, &tmB, &full[stage], kb * kBlockK, n0, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(n_cur);
      // (tried: A_hi x [W_hi; W_lo] as one UMMA of N = 2 * BLOCK_N, i.e. 2 instead of 3 UMMAs per k-step --
      // measured 10 % slower per k-block than three N = 80 UMMAs, so not used)
      uint32_t acc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&full[stage], phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
        const uint64_t a_hi = umma_desc_k_sw128(sa), a_lo = umma_desc_k_sw128(sa + S::A_BYTES);
        const uint64_t b_hi = umma_desc_k_sw128(sa + 2 * S::A_BYTES);
        const uint64_t b_lo = umma_desc_k_sw128(sa + 2 * S::A_BYTES + S::B_BYTES);
        // small terms first, then the main product; 4 k-steps of 8 tf32 (32 bytes -> +2 in 16-byte units)
        // mode 0: all three products into one accumulator; mode 1: hi*hi only (plain TF32, for reference);
        // mode 2: cross terms (lo*hi + hi*lo) into a second accumulator, added in fp32 by the epilogue.
#pragma unroll
        for (int k = 0; k < kBlockK / 8; ++k) {
          if (mode == 0) {
            umma_tf32(tmem_base, a_lo + 2 * k, b_hi + 2 * k, idesc, acc);
            umma_tf32(tmem_base, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
            umma_tf32(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, 1);
          } else if (mode == 1) {
            umma_tf32(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);
          } else if (mode == 2) {
            umma_tf32(tmem_base + BLOCK_N, a_lo + 2 * k, b_hi + 2 * k, idesc, acc);
            umma_tf32(tmem_base + BLOCK_N, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
            umma_tf32(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);
          } else {
            // mode 3: rotate over kSets independent (main, cross) accumulator pairs
            const int step = kb * (kBlockK / 8) + k;
            const int set = step % kSets;
            const uint32_t first = step < kSets ? 0u : 1u;
            const uint32_t tm = tmem_base + (uint32_t)(2 * set * BLOCK_N);
            umma_tf32(tm, a_hi + 2 * k, b_hi + 2 * k, idesc, first);
            umma_tf32(tm + BLOCK_N, a_lo + 2 * k, b_hi + 2 * k, idesc, first);
            umma_tf32(tm + BLOCK_N, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
          }
          acc = 1;
        }
        umma_commit(&empty[stage]);   // arrives when the MMAs above have finished reading this stage
      }
      umma_commit(tmem_full);         // accumulator complete
    }
  } else {
    // ---- epilogue warps: TMEM lane quarter q = warp % 4 ----
    const int q = warp & 3;
    mbar_wait(tmem_full, 0);
    tcgen05_fence_after();
    const int r = m0 + q * 32 + lane;
    const bool row_ok = r < M;
    float* crow = row_ok ? ep.C + map_row(ep.cmap, r) * ep.ldc : nullptr;
    float* clo = (row_ok && ep.C_lo) ? ep.C_lo + map_row(ep.cmap, r) * ep.ldc : nullptr;
    const float* mrow = (row_ok && ep.mask) ? ep.mask + (int64_t)r * ep.ldm : nullptr;
    const uint32_t* brow = (row_ok && ep.maskbits) ? ep.maskbits + (int64_t)r * 16 : nullptr;
    if (ep.gmax != nullptr) {
      // running max / argmax over each image's R columns; this thread owns the whole row of the tile
      const int imgs = ep.n_stride / ep.R;
      const int img0 = n0 / ep.R;
      float best = -INFINITY;
      int bi = 0, gi = 0;   // argmax within the group, current group index
      const int a = row_ok ? r / ep.ncell : 0, cell = row_ok ? r % ep.ncell : 0;
#pragma unroll 1
      for (int c = 0; c < n_cur; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        if (mode == 2) {
          float x[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BLOCK_N + c), x);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += x[i];
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int col = c + i;              // column within the tile
          const int g = col / ep.R;           // image within the tile
          if (g >= imgs) break;
          if (g != gi) { best = -INFINITY; bi = 0; gi = g; }
          const int rr = col - g * ep.R;
          if (v[i] > best) { best = v[i]; bi = rr; }
          if (rr == ep.R - 1 && row_ok && img0 + g < ep.B_img) {
            const int64_t o = ((int64_t)a * ep.B_img + img0 + g) * ep.ncell + cell;
            ep.gmax[o] = best;
            ep.gargmax[o] = bi;
          }
        }
      }
      tcgen05_fence_before();
    } else
#pragma unroll 1
    for (int c = 0; c < n_cur; c += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);   // warp-collective: no early exit
      if (mode == 2) {
        float x[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BLOCK_N + c), x);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += x[i];
      } else if (mode == 3) {
        const int nsteps = num_kb * (kBlockK / 8);
        float cross[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BLOCK_N + c), cross);
#pragma unroll
        for (int st = 1; st < kSets; ++st) {
          if (st < nsteps) {
            float x[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(2 * st * BLOCK_N + c), x);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += x[i];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((2 * st + 1) * BLOCK_N + c), x);
#pragma unroll
            for (int i = 0; i < 16; ++i) cross[i] += x[i];
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += cross[i];
      }
      const int col0 = n0 + c;
      if (!row_ok || col0 >= N) continue;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const int col = col0 + j;
        if (col >= N) break;     // N % 4 == 0
        float o[4] = {v[j], v[j + 1], v[j + 2], v[j + 3]};
        if (ep.bias) {
          const float4 bv = ld4(ep.bias + col);
          o[0] += bv.x; o[1] += bv.y; o[2] += bv.z; o[3] += bv.w;
        }
        if (ep.act == 1) {
#pragma unroll
          for (int t = 0; t < 4; ++t) o[t] = fmaxf(o[t], 0.f);
        } else if (ep.act == 2) {
#pragma unroll
          for (int t = 0; t < 4; ++t) o[t] = tanhf(o[t]);
        }
        if (brow) {
          const uint4 w = *reinterpret_cast<const uint4*>(brow + (col >> 7) * 4);
          const int l = (col & 127) >> 2;
          if (!((w.x >> l) & 1u)) o[0] = 0.f;
          if (!((w.y >> l) & 1u)) o[1] = 0.f;
          if (!((w.z >> l) & 1u)) o[2] = 0.f;
          if (!((w.w >> l) & 1u)) o[3] = 0.f;
        } else if (mrow) {
          float4 mv = ld4(mrow + col);
          if (ep.mask_lo_off != 0) {
            const float4 ml = ld4(mrow + col + ep.mask_lo_off);
            mv.x += ml.x; mv.y += ml.y; mv.z += ml.z; mv.w += ml.w;
          }
          if (!(mv.x > 0.f)) o[0] = 0.f;
          if (!(mv.y > 0.f)) o[1] = 0.f;
          if (!(mv.z > 0.f)) o[2] = 0.f;
          if (!(mv.w > 0.f)) o[3] = 0.f;
        }
        if (ep.accumulate) {
          const float4 old = ld4(crow + col);
          o[0] += old.x; o[1] += old.y; o[2] += old.z; o[3] += old.w;
        }
        if (clo) {
          float hi[4], lo[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) split_tf32(o[t], hi[t], lo[t]);
          st4(crow + col, make_float4(hi[0], hi[1], hi[2], hi[3]));
          st4(clo + col, make_float4(lo[0], lo[1], lo[2], lo[3]));
        } else {
          st4(crow + col, make_float4(o[0], o[1], o[2], o[3]));
        }
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

// ------------------------------------------------------------------------------------------
// TN (weight gradient):  C[i, j] = sum_r A[r, i] * B[r, j]   with A pair [2, rows, Ka], B pair [2, rows, Kb]
// Both operands are MN-major for the UMMA (the reduction index r is the slow memory dimension).
// TMA boxes are {32 columns, kTnBlockK rows}: one 128-byte-swizzled MN chunk each; chunk g of an operand sits
// at g * kTnBlockK * 128 B (LBO); 32-bit MN-major operands need the 32-byte-atom swizzle (K atoms of 4 rows,
// SBO = 512 B; one UMMA k-step of 8 rows spans two of them).  Split-K over blockIdx.z; each CTA
// writes a dense fp32 partial [Ka, Kb] tile to scratch, reduced in order by splitk_reduce_kernel.
// ------------------------------------------------------------------------------------------
constexpr int kTnBlockK = 32;   // rows of the reduction per stage (4 UMMA k-steps)

CL_D uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // stride between 32-element MN chunks
  d |= (uint64_t)(512 >> 4) << 32;                    // stride between K atoms (4 rows x 128 B)
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                             // SWIZZLE_128B_BASE32B: 32-bit MN-major operands swizzle
  return d;                                           // 32-byte atoms (TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
}
CL_HD uint32_t umma_idesc_tf32_mn(int n) { return umma_idesc_tf32(n) | (1u << 15) | (1u << 16); }

template <int BLOCK_N, int STAGES>
struct TnSmem {
  static constexpr int CHUNK = kTnBlockK * 128;            // one {32 x kTnBlockK} box
  static constexpr int A_BYTES = (kBlockM / 32) * CHUNK;   // one part
  static constexpr int B_BYTES = (BLOCK_N / 32) * CHUNK;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  float* __restrict__ partial, int rows, int Ka, int Kb, int rows_per_split, int mode) {
  pdl_prologue();
  using S = TnSmem<BLOCK_N, STAGES>;
  static_assert(BLOCK_N % 32 == 0 && 2 * BLOCK_N <= 512, "BLOCK_N");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.y * kBlockM, j0 = blockIdx.x * BLOCK_N;
  const int r_begin = blockIdx.z * rows_per_split;
  const int r_end = min(rows, r_begin + rows_per_split);
  const int num_kb = r_end > r_begin ? (r_end - r_begin + kTnBlockK - 1) / kTnBlockK : 0;
  constexpr uint32_t TMEM_COLS = tmem_cols(2 * BLOCK_N);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full[i], 1);
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
        mbar_expect_tx(&full[stage], mode == 1 ? S::A_BYTES + S::B_BYTES : S::STAGE_BYTES);
        uint8_t* s = smem + stage * S::STAGE_BYTES;
        const int r = r_begin + kb * kTnBlockK;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          if (part == 1 && mode == 1) break;
#pragma unroll
          for (int g = 0; g < kBlockM / 32; ++g)
            tma_load_3d(s + part * S::A_BYTES + g * S::CHUNK, &tmA, &full[stage], i0 + g * 32, r, part);
#pragma unroll
          for (int g = 0; g < BLOCK_N / 32; ++g)
            tma_load_3d(s + 2 * S::A_BYTES + part * S::B_BYTES + g * S::CHUNK, &tmB, &full[stage], j0 + g * 32, r,
                        part);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32_mn(BLOCK_N);
      uint32_t acc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&full[stage], phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
        const uint64_t a_hi = umma_desc_mn_sw128(sa, S::CHUNK), a_lo = umma_desc_mn_sw128(sa + S::A_BYTES, S::CHUNK);
        const uint64_t b_hi = umma_desc_mn_sw128(sa + 2 * S::A_BYTES, S::CHUNK);
        const uint64_t b_lo = umma_desc_mn_sw128(sa + 2 * S::A_BYTES + S::B_BYTES, S::CHUNK);
#pragma unroll
        for (int k = 0; k < kTnBlockK / 8; ++k) {   // one 8-row K group = 1024 B = 64 x 16 B
          if (mode != 1) {
            umma_tf32(tmem_base + BLOCK_N, a_lo + 64 * k, b_hi + 64 * k, idesc, acc);
            umma_tf32(tmem_base + BLOCK_N, a_hi + 64 * k, b_lo + 64 * k, idesc, 1);
          }
          umma_tf32(tmem_base, a_hi + 64 * k, b_hi + 64 * k, idesc, acc);
          acc = 1;
        }
        umma_commit(&empty[stage]);
      }
      umma_commit(tmem_full);
    }
  } else {
    const int q = warp & 3;
    const int i = i0 + q * 32 + lane;
    float* prow = partial + (int64_t)blockIdx.z * Ka * Kb + (int64_t)i * Kb;
    if (num_kb > 0) {
      mbar_wait(tmem_full, 0);
      tcgen05_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 16) {
      float v[16];
      if (num_kb > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        if (mode != 1) {
          float x[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BLOCK_N + c), x);
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] += x[t];
        }
      } else {
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = 0.f;
      }
      if (i < Ka) {
#pragma unroll
        for (int t = 0; t < 16; t += 4) {
          const int col = j0 + c + t;
          if (col < Kb) st4(prow + col, make_float4(v[t], v[t + 1], v[t + 2], v[t + 3]));   // Kb % 4 == 0
        }
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

// x -> split pair (hi at out, lo at out + n)
__global__ void split_tf32_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  pdl_prologue();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    const float hi = __uint_as_float(h);
    out[i] = hi;
    out[n + i] = v - hi;
  }
}
// src[rows, cols] (row pitch ld) -> dense split pair [2, rows, cols]
__global__ void split_tf32_rows_kernel(const float* __restrict__ src, int64_t rows, int cols, int64_t ld,
                                      float* __restrict__ out) {
  pdl_prologue();
  const int64_t n4 = rows * (cols / 4);
  const int64_t part = rows * (int64_t)cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (cols / 4);
    const int c = (int)(i % (cols / 4)) * 4;
    const float4 v = ld4(src + r * ld + c);
    float4 hi, lo;
    split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
    st4(out + r * cols + c, hi);
    st4(out + part + r * cols + c, lo);
  }
}

// W[rows, cols] -> split pair of W^T ([2, cols, rows])
__global__ void split_tf32_transpose_kernel(const float* __restrict__ W, int rows, int cols, int64_t ldw,
                                            float* __restrict__ out) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? W[(int64_t)r * ldw + c] : 0.f;
  }
  __syncthreads();
  const int64_t n = (int64_t)rows * cols;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;   // out[c, r]
    if (c < cols && r < rows) {
      const float v = tile[threadIdx.x][i];
      uint32_t h;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
      const float hi = __uint_as_float(h);
      out[(int64_t)c * rows + r] = hi;
      out[n + (int64_t)c * rows + r] = v - hi;
    }
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tensor map over a split pair [2, rows, K] (fp32, K contiguous, row pitch ld floats, part pitch part_stride
// floats), box {32, box_rows, 1}, 128-byte swizzle, zero fill out of bounds.
inline int make_pair_map(CUtensorMap* tm, const float* base, int64_t rows, int K, int64_t ld, int64_t part_stride,
                         int box_rows, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, int box_parts = 1) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "cuTensorMapEncodeTiled entry point not available");
    return CLIORA_ERR_CUDA;
  }
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, 2};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 4, (cuuint64_t)part_stride * 4};
  cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows, (cuuint32_t)box_parts};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CLIORA_ERR_CUDA;
  }
  return CLIORA_OK;
}

struct PairRef {            // a split-pair operand in global memory
  const float* base;        // hi part; lo part at base + part_stride
  int64_t rows;             // rows of the whole tensor (TMA bound)
  int64_t ld;               // floats between rows
  int64_t part_stride;      // floats between hi and lo
};

// Two tile configurations.  Wide (256 columns, 2 stages): 21.8 MAC per shared-memory byte read by the UMMAs, main +
// cross accumulators fill the 512 TMEM columns -- used whenever the narrow grid would exceed one wave.  Narrow (80
// columns, 4 stages): 5 CTAs per 128 rows at N = 400, more SMs busy on the small chart levels.
extern int g_tc_xnarrow;         // 1: allow the 128x48 tile for one-wave launches
constexpr int kTcXNarrowN = 48;
extern int g_tc_narrow_stages;   // debug knob: pipeline depth of the narrow tile (2 default, 3, 4)
extern int g_tc_small_tmem;   // debug knob: narrow-tile CTAs allocate 256 instead of 512 TMEM columns (two per SM possible)
constexpr int kTcNarrowN = 80, kTcNarrowStages = 3;   // 2 / 3 / 4 stages measured 5780 / 5829 / 5827 sent/s; 3 leaves 70 KB of smem to co-resident kernels
constexpr int kTcWideN = 256, kTcWideStages = 2;
constexpr int kTcMidN = 160, kTcMidStages = 3;       // 216 KB: A tile reused over 2x the columns of narrow, 3 stages

inline bool tc_supported(int N, int K, const PairRef& A, const PairRef& W) {
  return (N % 4 == 0) && (K % 4 == 0) && (A.ld % 4 == 0) && (W.ld % 4 == 0) && (A.part_stride % 4 == 0) &&
         (W.part_stride % 4 == 0) && aligned16(A.base) && aligned16(W.base);
}

template <int BLOCK_N, int STAGES>
inline int launch_tc_gemm_nt_cfg(cudaStream_t st, const PairRef& A, int a_row0, const PairRef& W, int M, int N, int K,
                                 const TcEpilogue& ep, const char* tag, int mode) {
  using S = TcSmem<BLOCK_N, STAGES>;
  CUtensorMap tmA, tmB;
  // one box carries the hi and the lo tile of an operand (they are adjacent in the stage buffer): 2 TMA
  // operations per k-block instead of 4; the single-pass mode only ever needs the hi part
  const int parts = mode == 1 ? 1 : 2;
  CL_TRY(make_pair_map(&tmA, A.base, A.rows, K, A.ld, A.part_stride, kBlockM, CU_TENSOR_MAP_SWIZZLE_128B, parts));
  CL_TRY(make_pair_map(&tmB, W.base, W.rows, K, W.ld, W.part_stride, BLOCK_N, CU_TENSOR_MAP_SWIZZLE_128B, parts));
  CL_CUDA(func_attr_at_least(reinterpret_cast<const void*>(tc_gemm_nt_kernel<BLOCK_N, STAGES>),
                             cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  dim3 grid(ceil_div(N, ep.n_stride > 0 ? ep.n_stride : BLOCK_N), ceil_div(M, kBlockM));
  ProfScope prof(st, tag, 2.0 * M * N * K, 4.0 * ((double)M * K * 2 + (double)N * K * 2 + (double)M * N));
  launch_k(tc_gemm_nt_kernel<BLOCK_N, STAGES>, grid, kThreads, S::TOTAL, st, tmA, tmB, ep, a_row0, M, N, K,
           mode | (g_tc_small_tmem ? 0x100 : 0));
  CL_CHECK_LAUNCH("tc_gemm_nt_kernel");
  return CLIORA_OK;
}

// C[M,N] = epi(A[a_row0 : a_row0+M, :K] @ W[:N, :K]^T)
inline int launch_tc_gemm_nt(cudaStream_t st, const PairRef& A, int a_row0, const PairRef& W, int M, int N, int K,
                             const TcEpilogue& ep, const char* tag, int mode = 2, int force_cfg = 0) {
  if (M <= 0 || N <= 0) return CLIORA_OK;
  // pick the tile by estimated time = waves x (measured time of one CTA of that shape at K = 400, scaled by K)
  int cfg = force_cfg;
  if (cfg == 0) {
    const int mt = ceil_div(M, kBlockM);
    const double kscale = 0.5 + 0.5 * (double)K / 400.0;
    auto cost = [&](int bn, double t_cta) { return ceil_div((int64_t)mt * ceil_div(N, bn), 148) * t_cta * kscale; };
    const double cn = cost(kTcNarrowN, 16.5), cm = cost(kTcMidN, 22.0), cw = cost(kTcWideN, 32.0);
    cfg = (cn <= cm && cn <= cw) ? 1 : (cm <= cw ? 3 : 2);
    // extra-narrow 128x48 tiles for launches that still fit one wave with them (small chart levels): less W per
    // k-block and shorter UMMAs per CTA, i.e. a shorter critical path when SMs would otherwise idle
    if (cfg == 1 && g_tc_xnarrow && ep.n_stride == 0 && ep.gmax == nullptr &&
        (int64_t)mt * ceil_div(N, kTcXNarrowN) <= 148)
      cfg = 4;
  }
  if (cfg == 4) return launch_tc_gemm_nt_cfg<kTcXNarrowN, 4>(st, A, a_row0, W, M, N, K, ep, tag, mode);
  if (cfg != 1 && mode == 3) mode = 2;   // only the narrow tile has room for rotating accumulator sets
  if (cfg == 3) return launch_tc_gemm_nt_cfg<kTcMidN, kTcMidStages>(st, A, a_row0, W, M, N, K, ep, tag, mode);
  if (cfg == 2) return launch_tc_gemm_nt_cfg<kTcWideN, kTcWideStages>(st, A, a_row0, W, M, N, K, ep, tag, mode);
  if (g_tc_narrow_stages == 4) return launch_tc_gemm_nt_cfg<kTcNarrowN, 4>(st, A, a_row0, W, M, N, K, ep, tag, mode);
  if (g_tc_narrow_stages == 3) return launch_tc_gemm_nt_cfg<kTcNarrowN, 3>(st, A, a_row0, W, M, N, K, ep, tag, mode);
  return launch_tc_gemm_nt_cfg<kTcNarrowN, kTcNarrowStages>(st, A, a_row0, W, M, N, K, ep, tag, mode);
}

constexpr int kTnBlockN = 224;
constexpr int kTnStages = 2;

// box {32 columns, box_rows rows, 1 part} over a pair whose contiguous dimension is the MN dimension
inline int make_pair_map_mn(CUtensorMap* tm, const float* base, int64_t rows, int cols, int64_t ld,
                            int64_t part_stride) {
  return make_pair_map(tm, base, rows, cols, ld, part_stride, kTnBlockK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

inline int tn_tc_splits(int rows, int Ka, int Kb) {
  const int tiles = ceil_div(Ka, kBlockM) * ceil_div(Kb, kTnBlockN);
  int s = (148 + tiles - 1) / tiles;
  const int kb = ceil_div(rows, kTnBlockK);
  if (s > kb) s = kb;
  if (s < 1) s = 1;
  return s;
}
inline int64_t tn_tc_scratch_floats(int rows, int Ka, int Kb) { return (int64_t)tn_tc_splits(rows, Ka, Kb) * Ka * Kb; }

// C[Ka,Kb] (+)= A[rows,Ka]^T B[rows,Kb], operands as split pairs
inline int launch_tc_gemm_tn(cudaStream_t st, const PairRef& A, const PairRef& B, int rows, int Ka, int Kb, float* C,
                             int64_t ldc, int accumulate, float* scratch, const char* tag, int mode = 2) {
  if (Ka <= 0 || Kb <= 0) return CLIORA_OK;
  if (rows <= 0) {   // empty reduction (a chart of single-word sentences has no splits): the sum is zero
    if (!accumulate) CL_CUDA(cudaMemset2DAsync(C, ldc * sizeof(float), 0, Kb * sizeof(float), Ka, st));
    return CLIORA_OK;
  }
  using S = TnSmem<kTnBlockN, kTnStages>;
  CUtensorMap tmA, tmB;
  CL_TRY(make_pair_map_mn(&tmA, A.base, A.rows, Ka, A.ld, A.part_stride));
  CL_TRY(make_pair_map_mn(&tmB, B.base, B.rows, Kb, B.ld, B.part_stride));
  CL_CUDA(func_attr_at_least(reinterpret_cast<const void*>(tc_gemm_tn_kernel<kTnBlockN, kTnStages>),
                             cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  const int splits = tn_tc_splits(rows, Ka, Kb);
  int per = ceil_div(ceil_div(rows, kTnBlockK), splits) * kTnBlockK;
  if (per < kTnBlockK) per = kTnBlockK;
  dim3 grid(ceil_div(Kb, kTnBlockN), ceil_div(Ka, kBlockM), splits);
  ProfScope prof(st, tag, 2.0 * rows * Ka * Kb, 4.0 * ((double)rows * (Ka + Kb) * 2 + (double)Ka * Kb));
  launch_k(tc_gemm_tn_kernel<kTnBlockN, kTnStages>, grid, kThreads, S::TOTAL, st, tmA, tmB, scratch, rows, Ka, Kb, per, mode);
  CL_CHECK_LAUNCH("tc_gemm_tn_kernel");
  const int64_t total = (int64_t)Ka * Kb;
  launch_k(splitk_reduce_kernel, ceil_div(total, 256), 256, 0, st, scratch, splits, total, Ka, Kb, C, ldc, accumulate);
  CL_CHECK_LAUNCH("splitk_reduce_kernel");
  return CLIORA_OK;
}

}  // namespace tc
}  // namespace cliora
