// Fused chart-level kernels (sm_100a): one launch does what split_build + compose GEMM + cell_aggregate did.
//
// A chart level is a set of cells, each with N splits (inside: N = level; outside: N = n-level-1).  The reference
// evaluates, per split row, y = ReLU(W2 ReLU(W1 [l;r] + b1) + b2) and e = l^T Wb r + s_l + s_r, then per cell
// p = softmax_k(e), a = sum_k p_k y_k, h = unit(a) (+ region attention for CLIORA)
// (cliora/net/diora.py:295-331,358-398; cliora/net/cliora.py:128-157,304-341).
//
// level_fwd_kernel: grid (nc column slices, tiles), launched as thread-block clusters of nc CTAs.  A cluster owns
// a tile of G whole cells (G*N <= 128 split rows); CTA `rank` owns output columns [rank*ncols, (rank+1)*ncols).
//   warp 0      TMA producer for the W2 slice (split pair, 128-byte swizzle)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (3xTF32: main + cross accumulators)
//   warps 2-5   thread = split row: (a) while the MMAs run, the partial bilinear score over this CTA's columns,
//               all-gathered across the cluster through distributed shared memory -> e, softmax p;
//               (b) epilogue: TMEM -> bias, ReLU -> Y row (saved for backward) -> p-weighted rows staged in smem
//               -> per-cell sums -> cluster-wide norms -> (CLIORA: region attention, second norm) -> chart
//   warps 6-13  A-operand producers: gather the two per-cell projection rows of every split from the chart's
//               projection buffer, add b1, ReLU, split into tf32 hi/lo, store swizzled into the smem stage
//               (Z never makes a round trip through HBM before the GEMM; its pair is streamed out once for the
//               weight-gradient GEMM of the backward pass)
#pragma once
#include "chart_kernels.cuh"
#include "cell_warp_kernels.cuh"
#include "tc_gemm.cuh"

namespace cliora {
extern int g_debug[16];
namespace lvl {

using tc::fence_barrier_init;
using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;
using tc::tcgen05_fence_after;
using tc::tcgen05_fence_before;
using tc::tma_load_3d;
using tc::tma_prefetch_desc;
using tc::tmem_ld16;
using tc::umma_commit;
using tc::umma_desc_k_sw128;
using tc::umma_idesc_tf32;
using tc::umma_tf32;

constexpr int kRows = 128;            // split rows per tile = UMMA M
constexpr int kAStages = 4;           // A-operand ring (gathered + transformed by the producer warps), 32 KB per stage
constexpr int kProdWarps = 8;
constexpr int kProdWarp0 = 6;
constexpr int kCopyWarp0 = kProdWarp0 + kProdWarps;      // warps 14-17: cp.async gathers of the raw operand rows
constexpr int kCopyWarps = 4;
constexpr int kCopyRowsPerPass = kCopyWarps * 4;         // tile rows one pass of the copy threads covers (8 lanes per row)
constexpr int kCopyIters = kRows / kCopyRowsPerPass;      // copies per thread, operand and k-block
constexpr int kThreads = (kCopyWarp0 + kCopyWarps) * 32;   // 576
constexpr int kMaxCluster = 8;
constexpr int kMaxUmmaN = 208;          // wide tiles: two column slices of a 400-wide level (single accumulator)
constexpr int kNarrowUmmaN = 112;       // up to here: main + cross accumulator, four raw stages
constexpr int kMaxD = 896;
constexpr int kABytes = kRows * 128;  // one 128 x 32 fp32 operand tile
// extras after the operand rings: barriers (256 B), b2 (128 f), e, p, nrm, nrm2 (4 x 128 f), three [8][128] exchange
// buffers, the global row of every tile row (int64)
constexpr int kExtraFloats = 256 + 4 * 128 + 3 * kMaxCluster * 128;
constexpr int kExtraBytes = 256 + 4 * kExtraFloats + 8 * 128;
// W2-slice ring (TMA): three stages when the slice is narrow, two otherwise (shared-memory budget)
CL_HD int b_stages(int n_umma) { return n_umma <= 80 ? 3 : 2; }
// raw-operand stages of the A pipeline (32 KB each): the wide W2 slice leaves room for three
CL_HD int raw_stages(int n_umma) { return n_umma > kNarrowUmmaN ? 3 : kAStages; }
CL_HD int ring_bytes(int n_umma) { return raw_stages(n_umma) * 2 * kABytes + b_stages(n_umma) * 2 * n_umma * 128; }

CL_D uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
CL_D void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
CL_D uint32_t map_to_cta(uint32_t saddr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
  return r;
}
CL_D void st_cluster_f32(uint32_t caddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(caddr), "f"(v) : "memory");
}
CL_D void mbar_arrive_cluster(uint32_t caddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
CL_D void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
CL_D void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
CL_D void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the four epilogue warps
CL_D float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
CL_D void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Cluster all-gather of small per-row / per-cell values through distributed shared memory.
// put: lanes 0..nc-1 of the calling warp each store `v` (warp-uniform) into slot[idx] of one CTA.
CL_D void xchg_put_uniform(float* slot, int idx, float v, int nc, int lane) {
  if (lane < nc) st_cluster_f32(map_to_cta(smem_u32(slot + idx), (uint32_t)lane), v);
}
// put: the calling thread stores its own value into slot[idx] of every CTA of the cluster.
CL_D void xchg_put(float* slot, int idx, float v, int nc) {
  const uint32_t a = smem_u32(slot + idx);
  for (int c = 0; c < nc; ++c) st_cluster_f32(map_to_cta(a, (uint32_t)c), v);
}
// One arrival per warp and destination CTA (barrier count = 4 epilogue warps x nc): every lane fences its remote
// stores at cluster scope, the warp synchronises, lane c releases on CTA c's barrier.
CL_D void xchg_arrive_warp(uint64_t* bar, int nc, int lane) {
  asm volatile("fence.acq_rel.cluster;" ::: "memory");
  __syncwarp();
  if (lane < nc) mbar_arrive_cluster(map_to_cta(smem_u32(bar), (uint32_t)lane));
}
// Only lane 0 spins with cluster-scope acquire (a cluster-scope acquire invalidates L1); __syncwarp orders the rest.
CL_D void xchg_wait_warp(uint64_t* bar, int lane) {
  if (lane == 0) mbar_wait_cluster(bar, 0);
  __syncwarp();
}
CL_D long long clock_now() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
  return t;
}
CL_D long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct LevelFwdArgs {
  int B, n, level, L, N, D, R;
  int G;         // cells per tile (G * N <= 128)
  int cells;     // B * L
  int ncols;     // output columns per CTA (multiple of 4)
  int n_umma;    // ncols rounded up to 16 (<= kMaxUmmaN)
  int nc;        // cluster size = column slices
  int mode;      // 2: fp32-accurate 3xTF32 (main + cross accumulator), 1: single TF32 pass, 3: bf16 operands (fp32 accumulate)
  int store_lo;    // the streamed-out operand pair (Z forward, GY backward) includes its lo part (0: the single-pass modes
                   // never read it, the plain fp32 value is written -- the tensor core truncates it to tf32 by itself)
  int single_acc;  // mode 2 with the cross terms accumulated into the main accumulator (wide tiles: 2 x n_umma > 256 columns)
  int outside;
  int64_t C;
  // first / second operand of a split: inside (left, right) both from the inside chart; outside (sibling from the
  // inside chart, parent from the outside chart) -- always in that argument order (diora.py:366-371)
  const float* P1; int ld1; int off_a1;               // projection row of `first`:  A-part at off_a1
  const float* P2; int ld2; int off_a2; int off_v2;   // projection row of `second`: A-part at off_a2, V = Wb h at off_v2
  const float* h1;                                    // chart vectors of `first` [B,C,D]
  const float* s1; const float* s2;                   // chart scores of first / second [B,C]
  const float* b1; const float* b2;
  float* Z; int64_t z_lo_off; uint32_t* zmask;        // level blocks: Z pair (hi, lo at +z_lo_off), ReLU bits [rows,16] or null
  uint16_t* zbits;                                    // level block [rows,32]: bit j of halfword c = [z[16 c + j] > 0], or null
  float* Y; float* E; float* Pr;                      // level blocks [rows,D], [rows], [rows]
  float* chart_h; float* chart_s;                     // [B,C,D], [B,C]
  float* q; float* nrm; float* nrm2; float* att;      // saved per cell (q, nrm2, att: R > 0 only)
  const float* obj; const uint8_t* keep;
  int max_sent;                                       // CLIORA: sentences a tile may span (their region slices are staged in smem)
  int no_norm;                                        // --normalize none: norms are the identity, saved norms hold -1
  int exp_flags;                                      // timing experiments only (results become wrong): 1 no proxy fence, 2 no Z stores, 4 no transform
  long long* dbg;                                     // optional in-kernel timeline [ctas][32] (clock64 stamps), or null
};

struct RowInfo {
  bool ok;
  int g, k, b;
  int64_t g1, g2, m, cell;
};

CL_D void cell_of(const LevelFwdArgs& a, int cg, int& b, int& p, int64_t& cell) {
  b = cg / a.L;
  p = cg - b * a.L;
  cell = (int64_t)b * a.C + lvl_off(a.n, a.level) + p;
}

CL_D RowInfo decode_row(const LevelFwdArgs& a, int tile, int cells_here, int r) {
  RowInfo ri;
  ri.g = r / a.N;
  ri.k = r - ri.g * a.N;
  ri.ok = ri.g < cells_here;
  const int g = ri.ok ? ri.g : 0, k = ri.ok ? ri.k : 0;
  const int cg = tile * a.G + g;
  int p;
  cell_of(a, cg, ri.b, p, ri.cell);
  int c1, c2;
  if (!a.outside) {
    inside_children(a.n, a.level, p, k, c1, c2);
    ri.m = (int64_t)cg * a.N + k;                       // reference row order (b, pos, split)
  } else {
    outside_parent_sibling(a.n, a.level, p, k, c2, c1); // first = sibling, second = parent
    ri.m = ((int64_t)ri.b * a.N + k) * a.L + p;         // reference row order (b, split, pos)
  }
  ri.g1 = (int64_t)ri.b * a.C + c1;
  ri.g2 = (int64_t)ri.b * a.C + c2;
  return ri;
}

// x = hi + lo exactly with hi = x truncated to tf32 (what the tensor core keeps of an fp32 operand anyway): two
// instructions per element instead of the four of a round-to-nearest split; lo <= 2^-10 |x|, so the pair still
// carries 21+ mantissa bits into the 3xTF32 product.
CL_D void split_trunc(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}

// 16 consecutive 32-bit columns of this thread's TMEM lane
CL_D void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
        "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
CL_D void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows x 8 tf32) is read from tensor memory -- lane = row, one
// 32-bit column per k element -- so the tensor core spends no shared-memory bandwidth on it
CL_D void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// bf16 mode (mode 3): D[tmem] (+)= A[tmem, bf16 pairs] * B[smem, bf16], fp32 accumulate (kind::f16, K = 16 per MMA)
CL_D void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = n
CL_HD uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
}
// two floats -> one packed bf16x2 word, element `even` in the low half (the order bf16 arrays have in memory)
CL_D uint32_t pack_bf16x2(float even, float odd) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(odd), "f"(even));
  return d;
}
CL_D void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
constexpr uint32_t kTmemA0 = 256;      // first column of the A-operand stages in tensor memory (64 columns per stage)

// zero-filling 16-byte asynchronous copy (src_bytes = 0 writes zeros)
CL_D void cp_async16_zfill(uint32_t dst_saddr, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_saddr), "l"(src), "r"(src_bytes) : "memory");
}

// ------------------------------------------------------------------------------------------
// A-operand pipeline shared by the forward and backward level kernels.
//   copy warps (kCopyWarp0 .. +4)   thread owns chunk c of rows rbase + 16 i (i < 8): cp.async lands the two raw
//       128-byte row slices of every tile row in the raw stage (first 16 KB: operand a, second 16 KB: operand b,
//       128-byte swizzled) and makes the stage's raw_full barrier track their completion.  They run ahead of the
//       transform by up to kAStages k-blocks and do nothing else, so gather latency never sits in the transform chain.
//   transform warps (kProdWarp0 .. +8)  thread = tile row (TMEM lane) x 16-column half of the k-block: reads its 16
//       floats of each raw row, applies xform four elements at a time, splits into tf32 (hi, lo) and writes both into
//       the tensor-memory stage (tcgen05.st); the raw stage goes back to the copy warps, the TMEM stage to the MMA.
// offs(r, oa, ob, ok): element offsets of row r's two raw rows relative to base_a / base_b;  xform(xa, xb, kc) -> float4;
// stream(kb, k0, hi, lo): optional streaming of the 16-float pair to global memory.
// Barriers: raw_full[s] (count 128 copy threads, cp.async completion), raw_empty[s] (8 transform warps),
//           fullA[s] (8 transform warps, TMEM stage written), emptyA[s] (tcgen05.commit: TMEM stage consumed).
// ------------------------------------------------------------------------------------------
CL_D void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <class OffFn, class StoreFn, class EndFn>
CL_D void copy_a_raw(uint8_t* smem, uint64_t* raw_full, uint64_t* raw_empty, int num_kb, int D, const float* base_a,
                     const float* base_b, OffFn offs, int store_turn0, int store_every, StoreFn store, EndFn store_end,
                     int nraw, long long* dbg = nullptr) {
  const int gt = threadIdx.x - kCopyWarp0 * 32;          // 0 .. 32 kCopyWarps - 1
  const int c = gt & 7, rbase = gt >> 3;
  uint32_t oa[kCopyIters], ob[kCopyIters];
  uint32_t okmask = 0;
#pragma unroll
  for (int i = 0; i < kCopyIters; ++i) {
    bool ok;
    offs(rbase + kCopyRowsPerPass * i, oa[i], ob[i], ok);
    oa[i] += c * 4;
    ob[i] += c * 4;
    okmask |= ok ? (1u << i) : 0u;
  }
  const uint32_t soff0 = (uint32_t)(rbase * 128 + ((c ^ (rbase & 7)) << 4));     // + 128 kCopyRowsPerPass i
  const uint32_t smem_base = smem_u32(smem);
  // The pair of k-block kp (every store_every-th one is this CTA's to stream out) is written from the raw stage by
  // the thread that copied it, right before the stage is refilled -- coalesced (eight lanes per 128-byte row slice).
  auto flush = [&](int kp) {
    if (store_every <= 0 || (kp % store_every) != store_turn0) return;
    const uint8_t* sA = smem + (kp % nraw) * 2 * kABytes + soff0;
    const int kc = kp * 32 + c * 4;
    const bool live = kc < D;
#pragma unroll
    for (int i = 0; i < kCopyIters; ++i) {
      if (live && ((okmask >> i) & 1u)) {
        const float4 xa = *reinterpret_cast<const float4*>(sA + i * (128 * kCopyRowsPerPass));
        const float4 xb = *reinterpret_cast<const float4*>(sA + kABytes + i * (128 * kCopyRowsPerPass));
        store(rbase + kCopyRowsPerPass * i, kc, xa, xb);
      }
    }
    store_end(kc, live);          // reached by every lane of the warp
  };
  for (int kb = 0; kb < num_kb; ++kb) {
    const int stage = kb % nraw;
    mbar_wait(&raw_empty[stage], ((kb / nraw) & 1) ^ 1);
    if (dbg && gt == 0 && kb >= 4 && kb < 8) dbg[56 + (kb - 4) * 2] = clock_now();
    if (kb >= nraw) {
      // raw_empty[stage] implies raw_full[stage] of k-block kb - 4 completed, i.e. every copy of it (this thread's
      // included) has landed and is visible: no cp.async.wait_group needed before reading the stage back
      if (dbg && gt == 0 && kb >= 4 && kb < 8) dbg[64 + (kb - 4) * 4] = clock_now();
      flush(kb - nraw);
      if (dbg && gt == 0 && kb >= 4 && kb < 8) dbg[65 + (kb - 4) * 4] = clock_now();
    }
    const int kcol = kb * 32 + c * 4;
    const uint32_t sA = smem_base + (uint32_t)(stage * 2 * kABytes) + soff0;
#pragma unroll
    for (int i = 0; i < kCopyIters; ++i) {
      const uint32_t nbytes = (((okmask >> i) & 1u) && kcol < D) ? 16u : 0u;
      cp_async16_zfill(sA + i * (128 * kCopyRowsPerPass), base_a + oa[i] + kb * 32, nbytes);
      cp_async16_zfill(sA + kABytes + i * (128 * kCopyRowsPerPass), base_b + ob[i] + kb * 32, nbytes);
    }
    if (dbg && gt == 0 && kb >= 4 && kb < 8) dbg[66 + (kb - 4) * 4] = clock_now();
    cp_async_arrive_noinc(&raw_full[stage]);              // arrives once this thread's copies above have landed
    cp_async_commit();
    if (dbg && gt == 0 && kb >= 4 && kb < 8) dbg[57 + (kb - 4) * 2] = clock_now();
  }
  cp_async_wait_all();
  for (int kp = (num_kb > nraw ? num_kb - nraw : 0); kp < num_kb; ++kp) flush(kp);
}

template <class XformFn>
CL_D void transform_a_tmem(uint8_t* smem, uint64_t* raw_full, uint64_t* raw_empty, uint64_t* fullA, uint64_t* emptyA,
                           uint32_t tmem_base, int num_kb, int mode, int nraw, XformFn xform, long long* dbg = nullptr,
                           uint16_t* zb_row = nullptr, int zb_turn0 = 0, int zb_every = 1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qd = warp & 3, half = (warp - kProdWarp0) >> 2;
  const int row = qd * 32 + lane;
  const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
  for (int kb = 0; kb < num_kb; ++kb) {
    const int stage = kb % kAStages;                        // tensor-memory stage
    const uint32_t phase = (kb / kAStages) & 1;
    const int rstage = kb % nraw;                           // raw shared-memory stage
    const bool st_ = dbg != nullptr && threadIdx.x == kProdWarp0 * 32 && kb >= 4 && kb < 8;
    if (st_) dbg[32 + (kb - 4) * 6] = clock_now();
    mbar_wait(&raw_full[rstage], (kb / nraw) & 1);          // the raw rows of k-block kb have landed
    if (st_) dbg[33 + (kb - 4) * 6] = clock_now();
    const uint8_t* sA = smem + rstage * 2 * kABytes + row * 128;
    const int k0 = kb * 32 + half * 16;
    float4 xa[4], xb[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int ch = ((half * 4 + t) ^ (row & 7)) << 4;
      xa[t] = *reinterpret_cast<const float4*>(sA + ch);
      xb[t] = *reinterpret_cast<const float4*>(sA + kABytes + ch);
    }
    float hi[16], lo[16];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float4 o = xform(xa[t], xb[t], k0 + t * 4);
      split_trunc(o.x, hi[4 * t], lo[4 * t]); split_trunc(o.y, hi[4 * t + 1], lo[4 * t + 1]);
      split_trunc(o.z, hi[4 * t + 2], lo[4 * t + 2]); split_trunc(o.w, hi[4 * t + 3], lo[4 * t + 3]);
    }
    if (zb_row != nullptr && (kb % zb_every) == zb_turn0) {   // sign bits of this thread's 16 values (CTA kb % nc writes)
      uint32_t bits = 0u;
#pragma unroll
      for (int j = 0; j < 16; ++j) bits |= (hi[j] > 0.f ? 1u : 0u) << j;
      zb_row[kb * 2 + half] = (uint16_t)bits;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_local(&raw_empty[rstage]);  // raw stage read: the copy warps may refill it
    if (st_) dbg[34 + (kb - 4) * 6] = clock_now();
    mbar_wait(&emptyA[stage], phase ^ 1);                   // the MMAs that read this TMEM stage have retired
    tcgen05_fence_after();
    if (st_) dbg[35 + (kb - 4) * 6] = clock_now();
    if (mode == 3) {       // bf16: the 16 values of this half become 8 packed columns (hi + lo restores the value)
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(hi[2 * j] + lo[2 * j], hi[2 * j + 1] + lo[2 * j + 1]);
      tmem_st8(tmem_base + lane_addr + kTmemA0 + (uint32_t)(stage * 64 + half * 8), pk);
    } else {
      const uint32_t ta = tmem_base + lane_addr + kTmemA0 + (uint32_t)(stage * 64 + half * 16);
      tmem_st16(ta, hi);
      if (mode != 1) tmem_st16(ta + 32, lo);
    }
    tmem_st_wait();
    if (st_) dbg[36 + (kb - 4) * 6] = clock_now();
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_local(&fullA[stage]);
    if (st_) dbg[37 + (kb - 4) * 6] = clock_now();
  }
}

template <bool A_TMEM>
__global__ void __launch_bounds__(kThreads, 1)
level_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const LevelFwdArgs a) {
  pdl_prologue();
  // The dynamic shared-memory window starts 1024-byte aligned (no static shared memory in this kernel); deriving every
  // pointer from the array itself keeps the shared address space visible to the compiler (LDS/STS instead of generic
  // LD/ST in the transform and finalize loops).
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int tile = blockIdx.y;
  const int n0 = rank * a.ncols;
  const int D = a.D;
  const int num_kb = (D + 31) / 32;
  const int b_bytes = a.n_umma * 128;
  const int nbs = b_stages(a.n_umma);
  const int a_stage_bytes = 2 * kABytes, b_stage_bytes = 2 * b_bytes;
  const int nraw = A_TMEM ? raw_stages(a.n_umma) : kAStages;
  uint8_t* ringB = smem + nraw * a_stage_bytes;
  const int cells_here = min(a.G, a.cells - tile * a.G);
  const int nc = a.nc, ncols = a.ncols, N = a.N;
  long long* dbg_row = a.dbg ? a.dbg + ((int64_t)blockIdx.y * a.nc + rank) * 128 : nullptr;
#define LV_STAMP(slot) do { if (dbg_row != nullptr && tid == 128) dbg_row[slot] = clock_now(); } while (0)
  if (dbg_row && tid == 0) { dbg_row[0] = clock_now(); dbg_row[30] = global_ns(); }

  uint8_t* ex = smem + ring_bytes(a.n_umma);
  uint64_t* fullA = reinterpret_cast<uint64_t*>(ex);
  uint64_t* emptyA = fullA + kAStages;
  uint64_t* fullB = emptyA + kAStages;
  uint64_t* emptyB = fullB + 3;
  uint64_t* tmem_full = emptyB + 3;
  uint64_t* xbar = tmem_full + 1;                 // 4 single-use cluster exchange barriers
  uint64_t* raw_full = xbar + 4;                  // raw operand rows of a stage landed (cp.async completion)
  uint64_t* raw_empty = raw_full + kAStages;      // raw stage read by the transform warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + kAStages);
  float* s_b2 = reinterpret_cast<float*>(ex + 256);
  float* s_e = s_b2 + 256;
  float* s_p = s_e + 128;
  float* s_nrm = s_p + 128;
  float* s_nrm2 = s_nrm + 128;
  float* s_xe = s_nrm2 + 128;                     // [nc][128] partial scores
  float* s_xs = s_xe + kMaxCluster * 128;         // [nc][128] partial |a|^2
  float* s_xs2 = s_xs + kMaxCluster * 128;        // [nc][128] partial |a2|^2
  long long* s_m = reinterpret_cast<long long*>(s_xs2 + kMaxCluster * 128);   // [128] global row of every tile row, -1 = none
  // A_TMEM: accumulators in columns [0, 2 n_umma), four A-operand stages (hi 32 | lo 32 columns each) from column 256
  const uint32_t tmem_cols_alloc = A_TMEM ? 512u : (uint32_t)tc::tmem_cols(2 * a.n_umma);

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmW);
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kAStages; ++i) {
        mbar_init(&fullA[i], kProdWarps);         // one arrival per producer / transform warp
        mbar_init(&emptyA[i], 1);
        mbar_init(&raw_full[i], kCopyWarps * 32); // cp.async completion of every copy thread
        mbar_init(&raw_empty[i], kProdWarps);
      }
      for (int i = 0; i < 3; ++i) {
        mbar_init(&fullB[i], 1);                  // the TMA thread's expect_tx arrival
        mbar_init(&emptyB[i], 1);
      }
      mbar_init(tmem_full, 1);
      mbar_init(&xbar[0], (uint32_t)nc * 4);                  // scores: the four epilogue warps of every CTA
      for (int i = 1; i < 4; ++i) mbar_init(&xbar[i], (uint32_t)nc * (kThreads / 32));   // cell phases: all warps
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols_alloc)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 256; i += kThreads) s_b2[i] = (i < ncols && n0 + i < D) ? a.b2[n0 + i] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  cluster_sync_all();          // every CTA's barriers exist before a peer may arrive on them
  const uint32_t tmem_base = *tmem_slot;
  if (dbg_row && tid == 0) dbg_row[1] = clock_now();

  if (warp == 0) {
    // ------------------------------------------------------------ W2 slice: TMA
    if (lane == 0) {
      if (a.mode == 3) {
        // bf16 W2 slice: one 128-byte-swizzled box of 64 k-elements feeds two k-blocks of the A pipeline
        for (int kq = 0; kq < (num_kb + 1) / 2; ++kq) {
          const int sb = kq % nbs;
          mbar_wait(&emptyB[sb], ((kq / nbs) & 1) ^ 1);
          mbar_expect_tx(&fullB[sb], (uint32_t)b_bytes);
          tma_load_3d(ringB + sb * b_stage_bytes, &tmW, &fullB[sb], kq * 64, n0, 0);
        }
      } else {
        const uint32_t tx = (uint32_t)(a.mode == 1 ? b_bytes : 2 * b_bytes);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int sb = kb % nbs;
          const uint32_t phase = (kb / nbs) & 1;
          mbar_wait(&emptyB[sb], phase ^ 1);
          mbar_expect_tx(&fullB[sb], tx);
          tma_load_3d(ringB + sb * b_stage_bytes, &tmW, &fullB[sb], kb * 32, n0, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issue
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(a.n_umma);
      uint32_t acc = 0;
      if (A_TMEM && a.mode == 3) {
        const uint32_t idesc16 = umma_idesc_bf16(a.n_umma);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int sA_i = kb % kAStages, kq = kb >> 1, sb = kq % nbs;
          mbar_wait(&fullB[sb], (kq / nbs) & 1);
          mbar_wait(&fullA[sA_i], (kb / kAStages) & 1);
          tcgen05_fence_after();
          const uint64_t bd = umma_desc_k_sw128(smem_u32(ringB + sb * b_stage_bytes)) + (uint64_t)((kb & 1) * 4);
          const uint32_t ta = tmem_base + kTmemA0 + (uint32_t)(sA_i * 64);
          umma_bf16_ts(tmem_base, ta, bd, idesc16, acc);              // k elements 0..15 of the k-block
          umma_bf16_ts(tmem_base, ta + 8, bd + 2, idesc16, 1);        // 16..31
          acc = 1;
          umma_commit(&emptyA[sA_i]);
          if ((kb & 1) || kb == num_kb - 1) umma_commit(&emptyB[sb]);
        }
        umma_commit(tmem_full);
      } else {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int sA_i = kb % kAStages, sb = kb % nbs;
        mbar_wait(&fullB[sb], (kb / nbs) & 1);
        mbar_wait(&fullA[sA_i], (kb / kAStages) & 1);
        tcgen05_fence_after();
        if (dbg_row && kb == 0) dbg_row[20] = clock_now();
        if (dbg_row && kb >= 3 && kb < 7) dbg_row[48 + (kb - 3) * 2] = clock_now();
        const uint32_t sa = smem_u32(smem + sA_i * a_stage_bytes);
        const uint32_t sbb = smem_u32(ringB + sb * b_stage_bytes);
        const uint64_t a_hi = umma_desc_k_sw128(sa), a_lo = umma_desc_k_sw128(sa + kABytes);
        const uint64_t b_hi = umma_desc_k_sw128(sbb);
        const uint64_t b_lo = umma_desc_k_sw128(sbb + b_bytes);
        const uint32_t ta_hi = tmem_base + kTmemA0 + (uint32_t)(sA_i * 64), ta_lo = ta_hi + 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (A_TMEM) {
            if (a.mode == 1) {
              umma_tf32_ts(tmem_base, ta_hi + 8 * k, b_hi + 2 * k, idesc, acc);
            } else if (a.single_acc) {
              umma_tf32_ts(tmem_base, ta_lo + 8 * k, b_hi + 2 * k, idesc, acc);
              umma_tf32_ts(tmem_base, ta_hi + 8 * k, b_lo + 2 * k, idesc, 1);
              umma_tf32_ts(tmem_base, ta_hi + 8 * k, b_hi + 2 * k, idesc, 1);
            } else {
              umma_tf32_ts(tmem_base + a.n_umma, ta_lo + 8 * k, b_hi + 2 * k, idesc, acc);
              umma_tf32_ts(tmem_base + a.n_umma, ta_hi + 8 * k, b_lo + 2 * k, idesc, 1);
              umma_tf32_ts(tmem_base, ta_hi + 8 * k, b_hi + 2 * k, idesc, acc);
            }
          } else if (a.mode == 1) {
            umma_tf32(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);
          } else {
            umma_tf32(tmem_base + a.n_umma, a_lo + 2 * k, b_hi + 2 * k, idesc, acc);
            umma_tf32(tmem_base + a.n_umma, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
            umma_tf32(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);
          }
          acc = 1;
        }
        umma_commit(&emptyA[sA_i]);
        umma_commit(&emptyB[sb]);
        if (dbg_row && kb >= 3 && kb < 7) dbg_row[49 + (kb - 3) * 2] = clock_now();
      }
      umma_commit(tmem_full);
      }
      if (dbg_row) dbg_row[21] = clock_now();
    }
  } else if (warp >= kCopyWarp0) {
    // ------------------------------------------------------------ raw operand rows: cp.async gathers (TMEM variant)
    if (A_TMEM) {
      uint32_t zo[kCopyIters];                     // element offset of the Z row of this thread's i-th tile row
#pragma unroll
      for (int i = 0; i < kCopyIters; ++i)
        zo[i] = (uint32_t)(decode_row(a, tile, cells_here, ((tid - kCopyWarp0 * 32) >> 3) + kCopyRowsPerPass * i).m * D);
      copy_a_raw(
          smem, raw_full, raw_empty, num_kb, D, a.P1, a.P2,
          [&](int r, uint32_t& oa, uint32_t& ob, bool& ok) {
            const RowInfo ri = decode_row(a, tile, cells_here, r);
            ok = ri.ok;
            oa = (uint32_t)(ri.g1 * a.ld1 + a.off_a1);
            ob = (uint32_t)(ri.g2 * a.ld2 + a.off_a2);
          },
          // the Z pair of k-block kb (for the dW2 GEMM of the backward pass) is streamed out by CTA kb % nc
          rank, (a.Z != nullptr && !(a.exp_flags & 2)) ? nc : 0,
          [&](int r, int kc, const float4& xa, const float4& xb) {
            const float4 o = make_float4(fmaxf(xa.x + xb.x, 0.f), fmaxf(xa.y + xb.y, 0.f), fmaxf(xa.z + xb.z, 0.f),
                                         fmaxf(xa.w + xb.w, 0.f));
            float* dst = a.Z + zo[r / kCopyRowsPerPass] + kc;
            if (!a.store_lo) {
              __stcs(reinterpret_cast<float4*>(dst), o);                 // streamed out once, read by the dW2 GEMM much later
            } else {
              float4 hi, lo;
              split_trunc(o.x, hi.x, lo.x); split_trunc(o.y, hi.y, lo.y); split_trunc(o.z, hi.z, lo.z); split_trunc(o.w, hi.w, lo.w);
              __stcs(reinterpret_cast<float4*>(dst), hi);
              __stcs(reinterpret_cast<float4*>(dst + a.z_lo_off), lo);
            }
          },
          [](int, bool) {}, nraw, dbg_row);
    }
  } else if (warp >= kProdWarp0 && A_TMEM) {
    // ------------------------------------------------------------ A operand -> tensor memory (see transform_a_tmem)
    // z = relu(Al[first] + Ar[second]); b1 rides on the second operand's projection, padding was zero-filled
    if (dbg_row && tid == kProdWarp0 * 32) dbg_row[22] = clock_now();
    uint16_t* zb_row = nullptr;
    if (a.zbits != nullptr && !(a.exp_flags & 8)) {
      const RowInfo ri = decode_row(a, tile, cells_here, (warp & 3) * 32 + lane);
      if (ri.ok) zb_row = a.zbits + ri.m * 32;
    }
    transform_a_tmem(
        smem, raw_full, raw_empty, fullA, emptyA, tmem_base, num_kb, a.mode, nraw,
        [&](const float4& xa, const float4& xb, int kc) {
          (void)kc;
          return make_float4(fmaxf(xa.x + xb.x, 0.f), fmaxf(xa.y + xb.y, 0.f), fmaxf(xa.z + xb.z, 0.f),
                             fmaxf(xa.w + xb.w, 0.f));
        },
        dbg_row, zb_row, rank, nc);
    if (dbg_row && tid == kProdWarp0 * 32) dbg_row[23] = clock_now();
  } else if (warp >= kProdWarp0) {
    // ------------------------------------------------------------ A operand: gather + ReLU + tf32 split
    // Each thread owns chunk c of rows rbase + 32 i.  The two projection rows of a split are copied asynchronously
    // (cp.async, L1-bypassing) straight into the stage's hi / lo slots -- kLookahead k-blocks ahead of their use --
    // and later transformed IN PLACE by the same thread: z = relu(al + ar + b1) -> (hi, lo).
    // Four A stages, lookahead 2: the copies of k-blocks it and it-1 fly while k-block it-2 is transformed, and the MMAs
    // of k-block it-2 overlap the next iteration (the stage that iteration refills was consumed one k-block earlier).
    constexpr int kLookahead = 2;
    const int pt = tid - kProdWarp0 * 32;      // 0..255
    const int c = pt & 7;                      // 16-byte chunk of the 128-byte k-block row
    const int rw = lane >> 3;                  // row within the warp's group of four
    const int rbase = (pt >> 3);               // rows rbase + 32 i, i < 4
    const float* pa[4];
    const float* pb[4];
    int64_t mrow[4];
    bool ok[4];
    uint32_t soff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = rbase + 32 * i;
      const RowInfo ri = decode_row(a, tile, cells_here, r);
      ok[i] = ri.ok;
      mrow[i] = ri.m;
      pa[i] = a.P1 + ri.g1 * a.ld1 + a.off_a1 + c * 4;
      pb[i] = a.P2 + ri.g2 * a.ld2 + a.off_a2 + c * 4;
      soff[i] = (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
    }
    const bool write_mask = a.zmask != nullptr && rank == 0;
    uint32_t mw[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) mw[i][j] = 0u;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dbg_row && pt == 0) dbg_row[22] = clock_now();
    const uint32_t smem_base = smem_u32(smem);
    int zturn = 0;                             // k-block kb's Z pair is streamed out by CTA kb % nc
    for (int it = 0; it < num_kb + kLookahead; ++it) {
      if (it < num_kb) {
        // ---- issue the copies of k-block `it`
        const int stage = it % kAStages;
        const uint32_t phase = (it / kAStages) & 1;
        const int kcol = it * 32 + c * 4;
        mbar_wait(&emptyA[stage], phase ^ 1);
        if (dbg_row && pt == 0 && it >= 4 && it < 8) dbg_row[32 + (it - 4) * 4] = clock_now();
        const uint32_t sA = smem_base + (uint32_t)(stage * a_stage_bytes);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t nbytes = (ok[i] && kcol < D) ? 16u : 0u;
          cp_async16_zfill(sA + soff[i], pa[i] + it * 32, nbytes);
          cp_async16_zfill(sA + kABytes + soff[i], pb[i] + it * 32, nbytes);
        }
      }
      cp_async_commit();                        // (empty groups in the tail keep the group count uniform)
      if (dbg_row && pt == 0 && it >= 4 && it < 8) dbg_row[33 + (it - 4) * 4] = clock_now();
      const int kb = it - kLookahead;
      if (kb < 0) continue;
      cp_async_wait<kLookahead>();              // this thread's copies of k-block kb have landed
      if (dbg_row && pt == 0 && it >= 4 && it < 8) dbg_row[34 + (it - 4) * 4] = clock_now();
      // ---- transform k-block kb in place
      const int stage = kb % kAStages;
      const int kcol = kb * 32 + c * 4;
      uint8_t* sA = smem + stage * a_stage_bytes;
      const float4 bv = zero4;      // b1 is folded into the second operand's projection (init_proj_kernel)
      const bool store_z = a.Z != nullptr && zturn == rank && kcol < D && !(a.exp_flags & 2);
      zturn = (zturn + 1 == nc) ? 0 : zturn + 1;
      if (!(a.exp_flags & 4))
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 xa = *reinterpret_cast<const float4*>(sA + soff[i]);
        const float4 xb = *reinterpret_cast<const float4*>(sA + kABytes + soff[i]);
        float4 o;
        o.x = fmaxf(xa.x + xb.x + bv.x, 0.f);
        o.y = fmaxf(xa.y + xb.y + bv.y, 0.f);
        o.z = fmaxf(xa.z + xb.z + bv.z, 0.f);
        o.w = fmaxf(xa.w + xb.w + bv.w, 0.f);
        if (!ok[i] || kcol >= D) o = zero4;
        float4 hi, lo;
        split_trunc(o.x, hi.x, lo.x); split_trunc(o.y, hi.y, lo.y); split_trunc(o.z, hi.z, lo.z); split_trunc(o.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(sA + soff[i]) = hi;
        *reinterpret_cast<float4*>(sA + kABytes + soff[i]) = lo;
        if (store_z && ok[i]) {
          st4(a.Z + mrow[i] * D + kcol, hi);
          st4(a.Z + a.z_lo_off + mrow[i] * D + kcol, lo);
        }
        if (write_mask) {      // warp-uniform
          const unsigned q0 = __ballot_sync(0xffffffffu, o.x > 0.f), q1 = __ballot_sync(0xffffffffu, o.y > 0.f);
          const unsigned q2 = __ballot_sync(0xffffffffu, o.z > 0.f), q3 = __ballot_sync(0xffffffffu, o.w > 0.f);
          const int sh = 8 * rw, up = 8 * (kb & 3);
          mw[i][0] |= ((q0 >> sh) & 0xffu) << up;
          mw[i][1] |= ((q1 >> sh) & 0xffu) << up;
          mw[i][2] |= ((q2 >> sh) & 0xffu) << up;
          mw[i][3] |= ((q3 >> sh) & 0xffu) << up;
        }
      }
      if (!(a.exp_flags & 1)) fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive_local(&fullA[stage]);
      if (dbg_row && pt == 0 && it >= 4 && it < 8) dbg_row[35 + (it - 4) * 4] = clock_now();
      if (write_mask && ((kb & 3) == 3 || kb == num_kb - 1)) {
        const int t = kb >> 2;                  // 128-column group
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c == 0 && ok[i] && t < 4)
            *reinterpret_cast<uint4*>(a.zmask + mrow[i] * 16 + t * 4) = make_uint4(mw[i][0], mw[i][1], mw[i][2], mw[i][3]);
          mw[i][0] = mw[i][1] = mw[i][2] = mw[i][3] = 0u;
        }
      }
    }
    if (dbg_row && pt == 0) dbg_row[23] = clock_now();
  } else {
    // ------------------------------------------------------------ warps 2-5 (thread = split row): scores, softmax
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const RowInfo ri = decode_row(a, tile, cells_here, r);
    s_m[r] = ri.ok ? (long long)ri.m : -1ll;
    LV_STAMP(2);
    // partial bilinear score over this CTA's columns, all-gathered over the cluster.  Coalesced: eight lanes walk one
    // row (16 bytes each, chunks c, c+8, ...), four rows per load instruction, two row groups (16 loads) in flight.
    {
      const int rr = lane >> 3, c = lane & 7;
      const int nch = ncols >> 2;                         // 16-byte chunks per row slice
      const int ok_i = ri.ok ? 1 : 0;
      const int g1_i = (int)ri.g1, g2_i = (int)ri.g2;     // chart rows fit 32 bits
#pragma unroll 1
      for (int grp = 0; grp < 8; grp += 2) {
        float parts[2] = {0.f, 0.f};
        bool okr[2];
        const float* hp[2];
        const float* vp[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int src_lane = (grp + u) * 4 + rr;
          okr[u] = __shfl_sync(0xffffffffu, ok_i, src_lane) != 0;
          const int r1 = __shfl_sync(0xffffffffu, g1_i, src_lane), r2 = __shfl_sync(0xffffffffu, g2_i, src_lane);
          hp[u] = a.h1 + (int64_t)r1 * D + n0;
          vp[u] = a.P2 + (int64_t)r2 * a.ld2 + a.off_v2 + n0;
        }
#pragma unroll 1
        for (int cb0 = 0; cb0 < nch; cb0 += 32) {         // 32 chunks (128 columns) per batch; wide slices take two
          float4 hv[2][4], vv[2][4];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int ch = cb0 + c + 8 * t;
              const bool ld = okr[u] && ch < nch && n0 + ch * 4 < D && !(a.exp_flags & 16);
              hv[u][t] = ld ? ldcg4(hp[u] + ch * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
              vv[u][t] = ld ? ldcg4(vp[u] + ch * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              parts[u] = fmaf(hv[u][t].x, vv[u][t].x, parts[u]); parts[u] = fmaf(hv[u][t].y, vv[u][t].y, parts[u]);
              parts[u] = fmaf(hv[u][t].z, vv[u][t].z, parts[u]); parts[u] = fmaf(hv[u][t].w, vv[u][t].w, parts[u]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float part = parts[u];
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          part += __shfl_xor_sync(0xffffffffu, part, 4);
          if (c == 0) xchg_put(s_xe + rank * 128, qd * 32 + (grp + u) * 4 + rr, part, nc);
        }
      }
    }
    float e = 0.f;
    if (ri.ok) e = a.s1[ri.g1] + a.s2[ri.g2];
    LV_STAMP(3);
    xchg_arrive_warp(&xbar[0], nc, lane);
    xchg_wait_warp(&xbar[0], lane);
    LV_STAMP(4);
    float dot = 0.f;
    for (int cc = 0; cc < nc; ++cc) dot += s_xe[cc * 128 + r];
    e += dot;
    s_e[r] = e;
    epi_bar_sync();
    float p = 0.f;
    if (ri.ok) {
      const int base = ri.g * N;
      float mx = -INFINITY;
      for (int kk = 0; kk < N; ++kk) mx = fmaxf(mx, s_e[base + kk]);
      float sum = 0.f;
      for (int kk = 0; kk < N; ++kk) sum += expf(s_e[base + kk] - mx);
      const float inv = 1.f / sum;
      p = expf(e - mx) * inv;
      if (rank == 0) {
        a.E[ri.m] = e;
        a.Pr[ri.m] = p;
        if (ri.k == 0) {
          float sbar = 0.f;
          for (int kk = 0; kk < N; ++kk) sbar = fmaf(expf(s_e[base + kk] - mx) * inv, s_e[base + kk], sbar);
          a.chart_s[ri.cell] = sbar;
        }
      }
    }
    s_p[r] = p;
    LV_STAMP(5);
  }

  // ================================================================ every warp: epilogue + cell finalize
  mbar_wait(tmem_full, 0);
  tcgen05_fence_after();
  __syncthreads();                               // softmax probabilities published; pipeline stages are free
  LV_STAMP(6);
  const int nc4 = ncols >> 2;
  const int stride = (nc4 & 1) ? ncols : ncols + 4;            // stride / 4 odd: conflict-free 16-byte accesses by row
  float* s_stage = reinterpret_cast<float*>(smem);            // [128][stride] compose outputs y
  float* s_a = s_stage + kRows * stride;                       // [G][ncols]
  const bool vl = a.R > 0;
  const int R = a.R;
  const int GRs = a.G * R, GRp = (GRs + 3) & ~3;
  float* s_att = s_a + a.G * ncols;              // [G][R] logits -> attention weights
  float* s_patt = s_att + GRp;                   // [G][R] dropout-scaled weights
  float* s_keep = s_patt + GRp;                  // [G][R] dropout scale of every (cell, region): 0 or 1 / (1 - p)
  float* s_obj = s_keep + GRp;                   // [max_sent][R][stride]: this CTA's column slice of the tile's images
  // [nc][G*R] partial region logits, written by the peers after the norm exchange (every CTA's rings are free by then);
  // same offset in every CTA of the cluster
  float* s_xl = s_obj + a.max_sent * R * stride;
  const int b_first = (tile * a.G) / a.L;
  if (vl) {
    // the region features of the sentences this tile spans: asynchronous copies, consumed after the first normalisation
    const int b_last = (tile * a.G + cells_here - 1) / a.L;
    const int total = (b_last - b_first + 1) * R * nc4;
    const float* src = a.obj + (int64_t)b_first * R * D + n0;
    for (int idx = tid; idx < total; idx += kThreads) {
      const int row = idx / nc4, j4 = idx - row * nc4;        // row = sentence * R + region
      float* dst = s_obj + row * stride + j4 * 4;
      if (n0 + j4 * 4 < D) cp_async16(dst, src + (int64_t)row * D + j4 * 4);
      else st4(dst, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    cp_async_commit();
  }
  if (warp >= 2) {
    // y = relu(acc + b2), staged row-major.  TMEM lane quarter = warp % 4, the 16-column chunks of a quarter are dealt
    // to its warps.
    constexpr int kSub = (kThreads / 32 - 2) / 4;          // warps per TMEM lane quarter
    const int qd = warp & 3, r = qd * 32 + lane, sub = (warp - 2) >> 2;
#pragma unroll 1
    for (int c0 = sub * 16; c0 < a.n_umma; c0 += 16 * kSub) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, v);    // warp-collective: no early exit
      if (a.mode == 2 && !a.single_acc) {
        float x[16];
        tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(a.n_umma + c0), x);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += x[i];
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const int col = c0 + j;
        if (col < ncols) {
          const float4 bv = *reinterpret_cast<const float4*>(s_b2 + col);
          float4 o;
          o.x = fmaxf(v[j] + bv.x, 0.f); o.y = fmaxf(v[j + 1] + bv.y, 0.f);
          o.z = fmaxf(v[j + 2] + bv.z, 0.f); o.w = fmaxf(v[j + 3] + bv.w, 0.f);
          if (n0 + col >= D) o = make_float4(0.f, 0.f, 0.f, 0.f);
          st4(s_stage + r * stride + col, o);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  LV_STAMP(7);
  // per-cell softmax-weighted sums over the N splits (fixed order: deterministic); the Y rows (saved for the backward
  // pass) leave from here, one 16-byte chunk per thread and a row slice per group of consecutive threads
  for (int item = tid; item < cells_here * nc4; item += kThreads) {
    const int g = item / nc4, j4 = item - g * nc4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = s_stage + (g * N) * stride + j4 * 4;
    const bool live = n0 + j4 * 4 < D;
    for (int kk = 0; kk < N; ++kk) {
      const float4 t = ld4(src + kk * stride);
      const float pk = s_p[g * N + kk];
      if (live) st4(a.Y + s_m[g * N + kk] * D + n0 + j4 * 4, t);
      acc.x = fmaf(pk, t.x, acc.x); acc.y = fmaf(pk, t.y, acc.y); acc.z = fmaf(pk, t.z, acc.z); acc.w = fmaf(pk, t.w, acc.w);
    }
    st4(s_a + g * ncols + j4 * 4, acc);
  }
  __syncthreads();
  // |a|^2 over this CTA's columns: warp per cell, lanes over columns; all-gathered over the cluster
  for (int g = warp; g < cells_here; g += kThreads / 32) {
    float ss = 0.f;
    for (int j = lane; j < ncols; j += 32) {
      const float t = s_a[g * ncols + j];
      ss = fmaf(t, t, ss);
    }
    ss = warp_sum(ss);
    xchg_put_uniform(s_xs + rank * 128, g, ss, nc, lane);
  }
  LV_STAMP(8);
  xchg_arrive_warp(&xbar[1], nc, lane);
  if (vl) {
    // dropout scale of every (cell, region) of the tile, fetched while the norms travel
    for (int item = tid; item < cells_here * R; item += kThreads) {
      const int g = item / R, rr = item - g * R;
      int cb, cp;
      int64_t ccell;
      cell_of(a, tile * a.G + g, cb, cp, ccell);
      s_keep[item] = a.keep != nullptr ? (a.keep[ccell * R + rr] ? kKeepScale : 0.f) : 1.f;
    }
  }
  xchg_wait_warp(&xbar[1], lane);
  LV_STAMP(9);
  if (tid < cells_here) {
    int cb, cp;
    int64_t ccell;
    cell_of(a, tile * a.G + tid, cb, cp, ccell);
    float tot = 0.f;
    for (int cc = 0; cc < nc; ++cc) tot += s_xs[cc * 128 + tid];
    const float nrm = a.no_norm ? 1.f : fmaxf(sqrtf(tot), kTiny);
    s_nrm[tid] = nrm;
    if (rank == 0) a.nrm[ccell] = a.no_norm ? -1.f : nrm;
  }
  if (vl) cp_async_wait_all();                   // this thread's region-feature copies have landed
  __syncthreads();
  for (int item = tid; item < cells_here * nc4; item += kThreads) {
    const int g = item / nc4, j4 = item - g * nc4;
    const float inv = 1.f / s_nrm[g];
    float4 t = ld4(s_a + g * ncols + j4 * 4);
    t.x *= inv; t.y *= inv; t.z *= inv; t.w *= inv;
    st4(s_a + g * ncols + j4 * 4, t);
    if (n0 + j4 * 4 < D) {
      int b2_, p2_;
      int64_t cell2;
      cell_of(a, tile * a.G + g, b2_, p2_, cell2);
      st4((vl ? a.q : a.chart_h) + cell2 * D + n0 + j4 * 4, t);
    }
  }
  LV_STAMP(10);
  if (vl) {
    // region attention of every cell against ITS OWN image only (the reference computes all B x B pairs and
    // keeps the diagonal, cliora.py:35-42)
    __syncthreads();
    for (int item = tid; item < cells_here * R; item += kThreads) {     // partial logits q . obj_r
      const int g = item / R, rr = item - g * R;
      const int bs = (tile * a.G + g) / a.L - b_first;
      const float* op = s_obj + (bs * R + rr) * stride;
      const float* qp = s_a + g * ncols;
      float d0 = 0.f, d1 = 0.f;
#pragma unroll 2
      for (int j4 = 0; j4 < nc4; ++j4) {
        const float4 qv = ld4(qp + j4 * 4), ov = ld4(op + j4 * 4);
        d0 = fmaf(qv.x, ov.x, d0); d1 = fmaf(qv.y, ov.y, d1); d0 = fmaf(qv.z, ov.z, d0); d1 = fmaf(qv.w, ov.w, d1);
      }
      xchg_put(s_xl + rank * GRs, item, d0 + d1, nc);
    }
    LV_STAMP(12);
    xchg_arrive_warp(&xbar[2], nc, lane);
    xchg_wait_warp(&xbar[2], lane);
    LV_STAMP(13);
    for (int g = warp; g < cells_here; g += kThreads / 32) {            // softmax over regions: warp per cell
      int cb, cp;
      int64_t ccell;
      cell_of(a, tile * a.G + g, cb, cp, ccell);
      float lg0 = -INFINITY, lg1 = -INFINITY;
      if (lane < R) {
        lg0 = 0.f;
        for (int cc = 0; cc < nc; ++cc) lg0 += s_xl[cc * GRs + g * R + lane];
      }
      if (lane + 32 < R) {
        lg1 = 0.f;
        for (int cc = 0; cc < nc; ++cc) lg1 += s_xl[cc * GRs + g * R + lane + 32];
      }
      const float mx = warp_max(fmaxf(lg0, lg1));
      const float e0 = lane < R ? expf(lg0 - mx) : 0.f, e1 = lane + 32 < R ? expf(lg1 - mx) : 0.f;
      const float inv = 1.f / warp_sum(e0 + e1);
      if (lane < R) {
        const float at = e0 * inv;
        if (rank == 0) a.att[ccell * R + lane] = at;
        s_patt[g * R + lane] = at * s_keep[g * R + lane];
      }
      if (lane + 32 < R) {
        const float at = e1 * inv;
        if (rank == 0) a.att[ccell * R + lane + 32] = at;
        s_patt[g * R + lane + 32] = at * s_keep[g * R + lane + 32];
      }
    }
    __syncthreads();
    LV_STAMP(14);
    for (int item = tid; item < cells_here * nc4; item += kThreads) {    // a2 = q + sum_r patt_r obj_r
      const int g = item / nc4, j4 = item - g * nc4;
      const int bs = (tile * a.G + g) / a.L - b_first;
      const float* op = s_obj + (bs * R) * stride + j4 * 4;
      const float* wp = s_patt + g * R;
      float4 t = ld4(s_a + g * ncols + j4 * 4);
#pragma unroll 4
      for (int rr = 0; rr < R; ++rr) {
        const float w = wp[rr];
        const float4 ov = ld4(op + rr * stride);
        t.x = fmaf(w, ov.x, t.x); t.y = fmaf(w, ov.y, t.y); t.z = fmaf(w, ov.z, t.z); t.w = fmaf(w, ov.w, t.w);
      }
      st4(s_a + g * ncols + j4 * 4, t);
    }
    __syncthreads();
    for (int g = warp; g < cells_here; g += kThreads / 32) {
      float ss = 0.f;
      for (int j = lane; j < ncols; j += 32) {
        const float t = s_a[g * ncols + j];
        ss = fmaf(t, t, ss);
      }
      ss = warp_sum(ss);
      xchg_put_uniform(s_xs2 + rank * 128, g, ss, nc, lane);
    }
    LV_STAMP(15);
    xchg_arrive_warp(&xbar[3], nc, lane);
    xchg_wait_warp(&xbar[3], lane);
    LV_STAMP(16);
    if (tid < cells_here) {
      int cb, cp;
      int64_t ccell;
      cell_of(a, tile * a.G + tid, cb, cp, ccell);
      float tot = 0.f;
      for (int cc = 0; cc < nc; ++cc) tot += s_xs2[cc * 128 + tid];
      const float nrm2 = a.no_norm ? 1.f : fmaxf(sqrtf(tot), kTiny);
      s_nrm2[tid] = nrm2;
      if (rank == 0) a.nrm2[ccell] = a.no_norm ? -1.f : nrm2;
    }
    __syncthreads();
    for (int item = tid; item < cells_here * nc4; item += kThreads) {
      const int g = item / nc4, j4 = item - g * nc4;
      if (n0 + j4 * 4 < D) {
        const float inv = 1.f / s_nrm2[g];
        float4 t = ld4(s_a + g * ncols + j4 * 4);
        t.x *= inv; t.y *= inv; t.z *= inv; t.w *= inv;
        int b2_, p2_;
        int64_t cell2;
        cell_of(a, tile * a.G + g, b2_, p2_, cell2);
        st4(a.chart_h + cell2 * D + n0 + j4 * 4, t);
      }
    }
  }
  LV_STAMP(17);
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols_alloc) : "memory");
  }
  cluster_sync_all();          // no CTA leaves while a peer may still store into its shared memory
  if (dbg_row && tid == 0) { dbg_row[18] = clock_now(); dbg_row[31] = global_ns(); }
#undef LV_STAMP
}

// ==========================================================================================
// level_bwd_kernel: the per-split part of a level's backward in one launch (replaces the per-split loop of cell_bwd,
// the GZ GEMM and split_scatter).  The per-cell part (normalise / attention backward) runs before it in cells-only
// mode of the cell kernels and leaves GA = d loss / d a (pre-normalisation cell sum) and the softmax constant CM.
//   per split row (cell c, split k):  gy = p_k * GA_c * [y_k > 0]         (A operand, built by the producer warps)
//                                     d_k = y_k . GA_c;  ge_k = p_k (gs_c + d_k + e_k gs_c - cm_c)
//   GZ = (GY W2) * [z > 0]            (tcgen05, W2^T slice by TMA; epilogue masks with the sign of the stored Z)
//   scatter: GP[first].A += GZ, GP[second].A += GZ, Gh[first] += ge V[second], GP[second].V += ge h[first],
//            Gs[first] += ge, Gs[second] += ge            (red.global.add, as split_scatter did)
//   db2 += column sums of GY (warps 2-5 while the MMAs run);  the GY pair is streamed out for the dW2 GEMM.
// Grid (nc column slices, tiles); no cluster: a CTA needs nothing from its column neighbours.
// ==========================================================================================
struct LevelBwdArgs {
  LevelFwdArgs geo;          // geometry only: B, n, level, L, N, D, G, cells, ncols, n_umma, nc, mode, outside, C
  const float* Y;            // level block [rows, D]: forward compose outputs
  const float* Zhi;          // level block [rows, D]: hi part of the hidden activations (its sign is the ReLU mask)
  const uint16_t* zbits;     // level block [rows, 32]: the same signs as bits (written by the forward level kernel), or null
  const float* Pr; const float* E;       // level blocks [rows]
  const float* GA; const float* Gs; const float* CM;   // [B,C,D], [B,C], [B,C] of this pass's chart
  const float* h1;           // chart vectors of `first` [B,C,D]
  const float* P2; int ld2; int off_a2; int off_v2;    // forward projection row of `second` (V part read here)
  float* GP1; int ld1; int off_a1;       // projection-gradient row of `first`
  float* GP2;                            // projection-gradient row of `second` (same pitch / offsets as P2)
  float* Gh1; float* Gs1; float* Gs2;    // vector / score gradient accumulators of first, score of second
  float* GYp; int64_t gy_lo_off;         // level block pair out [2][rows, D]
  float* db2;                            // [D] accumulator
  // text cells (no region attention): the per-cell normalise backward runs in this kernel's prologue instead of a
  // separate cells-only launch; null cellGh = GA / CM were prepared by the cell kernel
  const float* cellGh; const float* cellH; const float* cellNrm; const float* cellS;   // [B,C,D], [B,C,D], [B,C], [B,C]
  float* GAw; float* CMw;                // writable aliases of GA / CM
  // CLIORA cells (region attention): the same prologue also runs the attention / second-normalise backward against the
  // tile's images, staged in the still idle operand rings; null vl_obj = text cells.  Needs D <= 512, R <= 64.
  const float* vl_obj; const uint8_t* vl_keep; const float* vl_att; const float* vl_q; const float* vl_nrm2;
  float* vl_GA2; float* vl_coef;         // saved for the region-feature gradient: ga2 [B,C,D], (patt, g_logit) [B,C,2,R]
  int vl_R;
};

constexpr int kBwdDb2Floats = kMaxD;
// barriers, row ids, p/cell/d0/d1/ge, chart rows of first/second, ReLU bit masks [128][8], db2 partial sums [D]
constexpr int kBwdExtraBytes = 256 + 128 * 8 + 4 * (128 * 5 + 128) + 4 * 256 + 4 * 1024 + 4 * kBwdDb2Floats;

template <int DUMMY>
__global__ void __launch_bounds__(kThreads, 1)
level_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const LevelBwdArgs g) {
  pdl_prologue();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
  const LevelFwdArgs& a = g.geo;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = blockIdx.x;
  const int tile = blockIdx.y;
  const int n0 = rank * a.ncols;
  const int D = a.D;
  const int num_kb = (D + 31) / 32;
  const int b_bytes = a.n_umma * 128;
  const int nbs = b_stages(a.n_umma);
  const int a_stage_bytes = 2 * kABytes, b_stage_bytes = 2 * b_bytes;
  const int nraw = raw_stages(a.n_umma);
  uint8_t* ringB = smem + nraw * a_stage_bytes;
  const int cells_here = min(a.G, a.cells - tile * a.G);
  const int nc = a.nc, ncols = a.ncols;
  long long* dbg_row = a.dbg ? a.dbg + ((int64_t)blockIdx.y * a.nc + rank) * 128 : nullptr;
#define LB_STAMP(slot) do { if (dbg_row != nullptr && tid == 128) dbg_row[slot] = clock_now(); } while (0)
  if (dbg_row && tid == 0) { dbg_row[0] = clock_now(); dbg_row[30] = global_ns(); }

  uint8_t* ex = smem + ring_bytes(a.n_umma);
  uint64_t* fullA = reinterpret_cast<uint64_t*>(ex);
  uint64_t* emptyA = fullA + kAStages;
  uint64_t* fullB = emptyA + kAStages;
  uint64_t* emptyB = fullB + 3;
  uint64_t* tmem_full = emptyB + 3;
  uint64_t* raw_full = tmem_full + 1;
  uint64_t* raw_empty = raw_full + kAStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + kAStages);
  long long* s_m = reinterpret_cast<long long*>(ex + 256);          // [128] global row, -1 = none
  float* s_p = reinterpret_cast<float*>(s_m + 128);                  // [128] softmax probability of the row
  int* s_cell = reinterpret_cast<int*>(s_p + 128);                   // [128] chart cell (b*C + c) of the row
  float* s_d = reinterpret_cast<float*>(s_cell + 128);               // [2][128] y . ga, one half of the columns each
  float* s_ge = s_d + 256;                                           // [128]
  int* s_g1 = reinterpret_cast<int*>(s_ge + 256);                    // [128] chart row (b*C + c) of `first`
  int* s_g2 = s_g1 + 128;                                            // [128] chart row of `second`
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_g2 + 128);        // [128][8] ReLU bits of this CTA's columns of z
  float* s_db2 = reinterpret_cast<float*>(s_mask + 1024);             // [D] column sums of the GY k-blocks this CTA streams

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmW);
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kAStages; ++i) {
        mbar_init(&fullA[i], kProdWarps);
        mbar_init(&emptyA[i], 1);
        mbar_init(&raw_full[i], kCopyWarps * 32);
        mbar_init(&raw_empty[i], kProdWarps);
      }
      for (int i = 0; i < 3; ++i) {
        mbar_init(&fullB[i], 1);
        mbar_init(&emptyB[i], 1);
      }
      mbar_init(tmem_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < 128) {
    const RowInfo ri = decode_row(a, tile, cells_here, tid);
    s_m[tid] = ri.ok ? (long long)ri.m : -1ll;
    s_p[tid] = ri.ok ? g.Pr[ri.m] : 0.f;
    s_cell[tid] = (int)ri.cell;
    s_g1[tid] = (int)ri.g1;
    s_g2[tid] = (int)ri.g2;
  }
  for (int j = tid; j < D; j += kThreads) s_db2[j] = 0.f;
  if (g.cellGh != nullptr && g.vl_obj != nullptr) {
    // CLIORA cells: second normalise -> region attention -> first normalise, backward (cliora.py:128-157), warp per
    // cell against the image's regions staged in shared memory; every column-slice CTA computes the same values.
    const int R = g.vl_R;
    const int b_first = (tile * a.G) / a.L, b_last = (tile * a.G + cells_here - 1) / a.L;
    float* s_obj = reinterpret_cast<float*>(smem);             // [sentences of the tile][R][D]
    const int total4 = (b_last - b_first + 1) * R * D / 4;
    const float* src = g.vl_obj + (int64_t)b_first * R * D;
    for (int i = tid; i < total4; i += kThreads) cp_async16(s_obj + i * 4, src + i * 4);
    cp_async_commit();
    LB_STAMP(6);
    for (int base = 0; base < cells_here; base += kThreads / 32) {      // uniform trip count: the barrier below is safe
      const int gi = base + warp;
      const bool active = gi < cells_here;
      int cb = 0, cp = 0;
      int64_t cell = 0;
      if (active) cell_of(a, tile * a.G + gi, cb, cp, cell);
      float4 gv[kColT], qv[kColT];
      float nrm_raw = 1.f, at0 = 0.f, at1 = 0.f, sc0 = 1.f, sc1 = 1.f;
      if (active) {
        // everything that does not need the regions first: ga2 = unit_bwd(g, h, nrm2), the attention weights and masks
        nrm_raw = g.cellNrm[cell];
        const float nrm2_raw = g.vl_nrm2[cell];
        if (lane < R) {
          at0 = g.vl_att[cell * R + lane];
          if (g.vl_keep != nullptr) sc0 = g.vl_keep[cell * R + lane] ? kKeepScale : 0.f;
        }
        if (lane + 32 < R) {
          at1 = g.vl_att[cell * R + lane + 32];
          if (g.vl_keep != nullptr) sc1 = g.vl_keep[cell * R + lane + 32] ? kKeepScale : 0.f;
        }
        float4 hv[kColT];
        float hd = 0.f;
#pragma unroll
        for (int t = 0; t < kColT; ++t) {
          const int j = lane * 4 + t * 128;
          gv[t] = hv[t] = qv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < D) {
            gv[t] = ldcg4(g.cellGh + cell * D + j);
            hv[t] = ldcg4(g.cellH + cell * D + j);
            qv[t] = ldcg4(g.vl_q + cell * D + j);
            hd += dot4(hv[t], gv[t]);
          }
        }
        hd = warp_sum(hd);
        const float coef2 = unit_bwd_coef(nrm2_raw, hd), inv2 = 1.f / fabsf(nrm2_raw);
#pragma unroll
        for (int t = 0; t < kColT; ++t) {
          const int j = lane * 4 + t * 128;
          gv[t] = make_float4((gv[t].x - hv[t].x * coef2) * inv2, (gv[t].y - hv[t].y * coef2) * inv2,
                              (gv[t].z - hv[t].z * coef2) * inv2, (gv[t].w - hv[t].w * coef2) * inv2);   // ga2
          if (j < D) st4(g.vl_GA2 + cell * D + j, gv[t]);
        }
      }
      LB_STAMP(7);
      if (base == 0) {
        cp_async_wait_all();
        __syncthreads();                         // the regions of the tile's images are staged
      }
      LB_STAMP(8);
      if (!active) continue;
      const float* ob = s_obj + (int64_t)(cb - b_first) * R * D;
      // g_att_r = (ga2 . obj_r) * scale_r, owned by lane r % 32
      float ga0, ga1;
      region_dots(gv, ob, R, D, lane, ga0, ga1);
      if (lane >= R) ga0 = 0.f;
      if (lane + 32 >= R) ga1 = 0.f;
      LB_STAMP(9);
      ga0 *= sc0;
      ga1 *= sc1;
      const float dsum = warp_sum(at0 * ga0 + at1 * ga1);
      const float gl0 = at0 * (ga0 - dsum), gl1 = at1 * (ga1 - dsum);
      if (lane < R) {
        g.vl_coef[(cell * 2) * R + lane] = at0 * sc0;
        g.vl_coef[(cell * 2 + 1) * R + lane] = gl0;
      }
      if (lane + 32 < R) {
        g.vl_coef[(cell * 2) * R + lane + 32] = at1 * sc1;
        g.vl_coef[(cell * 2 + 1) * R + lane + 32] = gl1;
      }
#pragma unroll 4
      for (int r = 0; r < R; ++r) {   // gq = ga2 + sum_r g_logit_r obj_r
        const float w = __shfl_sync(0xffffffffu, r < 32 ? gl0 : gl1, r & 31);
#pragma unroll
        for (int t = 0; t < kColT; ++t) {
          const int j = lane * 4 + t * 128;
          if (j < D) fma4(gv[t], w, ld4(ob + r * D + j));
        }
      }
      float hd = 0.f;
#pragma unroll
      for (int t = 0; t < kColT; ++t)
        if (lane * 4 + t * 128 < D) hd += dot4(qv[t], gv[t]);
      hd = warp_sum(hd);
      // ga = unit_bwd(gq, q, nrm)
      const float coef = unit_bwd_coef(nrm_raw, hd), inv = 1.f / fabsf(nrm_raw);
      float ad = 0.f;
#pragma unroll
      for (int t = 0; t < kColT; ++t) {
        const int j = lane * 4 + t * 128;
        gv[t] = make_float4((gv[t].x - qv[t].x * coef) * inv, (gv[t].y - qv[t].y * coef) * inv,
                            (gv[t].z - qv[t].z * coef) * inv, (gv[t].w - qv[t].w * coef) * inv);
        if (j < D) {
          ad += dot4(qv[t], gv[t]);
          st4(g.GAw + cell * D + j, gv[t]);
        }
      }
      ad = warp_sum(ad);
      if (lane == 0) g.CMw[cell] = fabsf(nrm_raw) * ad + g.cellS[cell] * g.Gs[cell];
      LB_STAMP(10);
    }
    __threadfence_block();
  } else if (g.cellGh != nullptr) {
    // per-cell normalise backward (text cells): ga = (g - h (h.g)) / nrm on the live branch, g / eps on the clamped one;
    // cm = sum_m p_m gp_m = nrm (h . ga) + s gs.  Warp per cell; every column-slice CTA computes the same values.
    for (int gi = warp; gi < cells_here; gi += kThreads / 32) {
      int cb, cp;
      int64_t cell;
      cell_of(a, tile * a.G + gi, cb, cp, cell);
      const float* gh = g.cellGh + cell * D;
      const float* hv = g.cellH + cell * D;
      const float nrm_raw = g.cellNrm[cell];
      const float nrm = fabsf(nrm_raw);
      float hd = 0.f;
      for (int j = lane * 4; j < D; j += 128) {
        const float4 x = ldcg4(gh + j), h4 = ldcg4(hv + j);
        hd = fmaf(h4.x, x.x, hd); hd = fmaf(h4.y, x.y, hd); hd = fmaf(h4.z, x.z, hd); hd = fmaf(h4.w, x.w, hd);
      }
      hd = warp_sum(hd);
      const float coef = unit_bwd_coef(nrm_raw, hd), inv = 1.f / nrm;
      float ad = 0.f;
      for (int j = lane * 4; j < D; j += 128) {
        const float4 x = ldcg4(gh + j), h4 = ldcg4(hv + j);
        const float4 v = make_float4((x.x - h4.x * coef) * inv, (x.y - h4.y * coef) * inv, (x.z - h4.z * coef) * inv,
                                     (x.w - h4.w * coef) * inv);
        ad = fmaf(h4.x, v.x, ad); ad = fmaf(h4.y, v.y, ad); ad = fmaf(h4.z, v.z, ad); ad = fmaf(h4.w, v.w, ad);
        st4(g.GAw + cell * D + j, v);
      }
      ad = warp_sum(ad);
      if (lane == 0) g.CMw[cell] = nrm * ad + g.cellS[cell] * g.Gs[cell];
    }
    __threadfence_block();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  LB_STAMP(1);

  if (warp == 0) {
    if (lane == 0) {
      if (a.mode == 3) {
        // bf16 W2 slice: one 128-byte-swizzled box of 64 k-elements feeds two k-blocks of the A pipeline
        for (int kq = 0; kq < (num_kb + 1) / 2; ++kq) {
          const int sb = kq % nbs;
          mbar_wait(&emptyB[sb], ((kq / nbs) & 1) ^ 1);
          mbar_expect_tx(&fullB[sb], (uint32_t)b_bytes);
          tma_load_3d(ringB + sb * b_stage_bytes, &tmW, &fullB[sb], kq * 64, n0, 0);
        }
      } else {
        const uint32_t tx = (uint32_t)(a.mode == 1 ? b_bytes : 2 * b_bytes);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int sb = kb % nbs;
          mbar_wait(&emptyB[sb], ((kb / nbs) & 1) ^ 1);
          mbar_expect_tx(&fullB[sb], tx);
          tma_load_3d(ringB + sb * b_stage_bytes, &tmW, &fullB[sb], kb * 32, n0, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(a.n_umma);
      uint32_t acc = 0;
      if (a.mode == 3) {
        const uint32_t idesc16 = umma_idesc_bf16(a.n_umma);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int sA_i = kb % kAStages, kq = kb >> 1, sb = kq % nbs;
          mbar_wait(&fullB[sb], (kq / nbs) & 1);
          mbar_wait(&fullA[sA_i], (kb / kAStages) & 1);
          tcgen05_fence_after();
          const uint64_t bd = umma_desc_k_sw128(smem_u32(ringB + sb * b_stage_bytes)) + (uint64_t)((kb & 1) * 4);
          const uint32_t ta = tmem_base + kTmemA0 + (uint32_t)(sA_i * 64);
          umma_bf16_ts(tmem_base, ta, bd, idesc16, acc);
          umma_bf16_ts(tmem_base, ta + 8, bd + 2, idesc16, 1);
          acc = 1;
          umma_commit(&emptyA[sA_i]);
          if ((kb & 1) || kb == num_kb - 1) umma_commit(&emptyB[sb]);
        }
      } else
      for (int kb = 0; kb < num_kb; ++kb) {
        const int sA_i = kb % kAStages, sb = kb % nbs;
        mbar_wait(&fullB[sb], (kb / nbs) & 1);
        mbar_wait(&fullA[sA_i], (kb / kAStages) & 1);
        tcgen05_fence_after();
        if (dbg_row && kb == 0) dbg_row[20] = clock_now();
        if (dbg_row && kb == num_kb - 1) dbg_row[21] = clock_now();
        const uint32_t sbb = smem_u32(ringB + sb * b_stage_bytes);
        const uint64_t b_hi = umma_desc_k_sw128(sbb), b_lo = umma_desc_k_sw128(sbb + b_bytes);
        const uint32_t ta_hi = tmem_base + kTmemA0 + (uint32_t)(sA_i * 64), ta_lo = ta_hi + 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (a.mode == 1) {
            umma_tf32_ts(tmem_base, ta_hi + 8 * k, b_hi + 2 * k, idesc, acc);
          } else if (a.single_acc) {
            umma_tf32_ts(tmem_base, ta_lo + 8 * k, b_hi + 2 * k, idesc, acc);
            umma_tf32_ts(tmem_base, ta_hi + 8 * k, b_lo + 2 * k, idesc, 1);
            umma_tf32_ts(tmem_base, ta_hi + 8 * k, b_hi + 2 * k, idesc, 1);
          } else {
            umma_tf32_ts(tmem_base + a.n_umma, ta_lo + 8 * k, b_hi + 2 * k, idesc, acc);
            umma_tf32_ts(tmem_base + a.n_umma, ta_hi + 8 * k, b_lo + 2 * k, idesc, 1);
            umma_tf32_ts(tmem_base, ta_hi + 8 * k, b_hi + 2 * k, idesc, acc);
          }
          acc = 1;
        }
        umma_commit(&emptyA[sA_i]);
        umma_commit(&emptyB[sb]);
      }
      umma_commit(tmem_full);
    }
  } else if (warp >= kCopyWarp0) {
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
    copy_a_raw(
        smem, raw_full, raw_empty, num_kb, D, g.Y, g.GA,
        [&](int r, uint32_t& oa, uint32_t& ob, bool& ok) {
          const long long m = s_m[r];
          ok = m >= 0;
          oa = (uint32_t)((ok ? m : 0) * D);
          ob = (uint32_t)((int64_t)s_cell[r] * D);
        },
        // the GY pair of k-block kb (the A operand of the dW2 GEMM) is streamed out by CTA kb % nc
        rank, g.GYp != nullptr ? nc : 0,
        [&](int r, int kc, const float4& y, const float4& ga) {
          const float pr = s_p[r];
          const float4 o = make_float4(y.x > 0.f ? pr * ga.x : 0.f, y.y > 0.f ? pr * ga.y : 0.f,
                                       y.z > 0.f ? pr * ga.z : 0.f, y.w > 0.f ? pr * ga.w : 0.f);
          csum.x += o.x; csum.y += o.y; csum.z += o.z; csum.w += o.w;
          float* dst = g.GYp + s_m[r] * D + kc;
          if (!a.store_lo) {
            __stcs(reinterpret_cast<float4*>(dst), o);
          } else {
            float4 hi, lo;
            split_trunc(o.x, hi.x, lo.x); split_trunc(o.y, hi.y, lo.y);
            split_trunc(o.z, hi.z, lo.z); split_trunc(o.w, hi.w, lo.w);
            __stcs(reinterpret_cast<float4*>(dst), hi);
            __stcs(reinterpret_cast<float4*>(dst + g.gy_lo_off), lo);
          }
        },
        // db2 += column sums of GY: this thread's rows of the four columns it streamed, combined in shared memory
        [&](int kc, bool live) {
          // the four lanes of a warp that share a chunk (lane, lane ^ 8, lane ^ 16, lane ^ 24) are summed first
#pragma unroll
          for (int off = 8; off < 32; off <<= 1) {
            csum.x += __shfl_xor_sync(0xffffffffu, csum.x, off); csum.y += __shfl_xor_sync(0xffffffffu, csum.y, off);
            csum.z += __shfl_xor_sync(0xffffffffu, csum.z, off); csum.w += __shfl_xor_sync(0xffffffffu, csum.w, off);
          }
          if (live && (threadIdx.x & 31) < 8) {
            atomicAdd(s_db2 + kc, csum.x); atomicAdd(s_db2 + kc + 1, csum.y);
            atomicAdd(s_db2 + kc + 2, csum.z); atomicAdd(s_db2 + kc + 3, csum.w);
          }
          csum = make_float4(0.f, 0.f, 0.f, 0.f);
        },
        nraw);
  } else if (warp >= kProdWarp0) {
    // ---- A operand: gy = p * ga * [y > 0] -> tensor memory (see transform_a_tmem); d = y . ga on the side
    const int row = (warp & 3) * 32 + lane, half = (warp - kProdWarp0) >> 2;
    const float prow = s_p[row];
    float dpart = 0.f;
    transform_a_tmem(
        smem, raw_full, raw_empty, fullA, emptyA, tmem_base, num_kb, a.mode, nraw,
        [&](const float4& y, const float4& ga, int kc) {
          dpart = fmaf(y.x, ga.x, dpart); dpart = fmaf(y.y, ga.y, dpart);
          dpart = fmaf(y.z, ga.z, dpart); dpart = fmaf(y.w, ga.w, dpart);
          (void)kc;
          return make_float4(y.x > 0.f ? prow * ga.x : 0.f, y.y > 0.f ? prow * ga.y : 0.f,
                             y.z > 0.f ? prow * ga.z : 0.f, y.w > 0.f ? prow * ga.w : 0.f);
        });
    s_d[half * 128 + row] = dpart;
  } else {
    // ---- warps 2-5 while the MMAs run (thread = tile row): the ReLU mask of this CTA's columns of the row's z,
    // four bits per 16-byte chunk, so that the epilogue neither waits on global memory nor re-reads Z
    const int r = tid - 64;            // 0..127
    const long long m = s_m[r];
    const int nch = ncols >> 2;
    uint32_t mk[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (m >= 0 && g.zbits != nullptr) {
      // 64 bytes of sign bits per row: the bits of columns n0 .. n0 + 255 are funnel-shifted out of nine words
      const uint32_t* wrow = reinterpret_cast<const uint32_t*>(g.zbits + m * 32);
      const int w0 = n0 >> 5, sh = n0 & 31;
      uint32_t wv[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) wv[q] = (w0 + q < 16 && q * 32 < ncols + 32) ? __ldcg(wrow + w0 + q) : 0u;
#pragma unroll
      for (int q = 0; q < 8; ++q) mk[q] = __funnelshift_r(wv[q], wv[q + 1], sh);
    } else if (m >= 0) {
      const float* zrow = g.Zhi + m * D + n0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q * 8 < nch) {
          float4 z[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = q * 8 + u;
            z[u] = (j < nch && n0 + 4 * j < D) ? ldcg4(zrow + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t b = (z[u].x > 0.f ? 1u : 0u) | (z[u].y > 0.f ? 2u : 0u) | (z[u].z > 0.f ? 4u : 0u) |
                               (z[u].w > 0.f ? 8u : 0u);
            mk[q] |= b << (4 * u);
          }
        }
      }
    }
    *reinterpret_cast<uint4*>(s_mask + r * 8) = make_uint4(mk[0], mk[1], mk[2], mk[3]);
    *reinterpret_cast<uint4*>(s_mask + r * 8 + 4) = make_uint4(mk[4], mk[5], mk[6], mk[7]);
    LB_STAMP(2);
  }

  // ================================================================ every warp: epilogue + scatter
  mbar_wait(tmem_full, 0);
  tcgen05_fence_after();
  __syncthreads();
  LB_STAMP(3);
  // The operand rings are free now: this CTA's slices of h[first] and V[second] (score path) are copied into them
  // asynchronously while the accumulators are drained, masked and transposed (thread = row -> row-major in shared
  // memory), so that the scatter below issues coalesced reductions (one row slice per warp instruction) and never waits
  // on a global load.
  const int nch = ncols >> 2;
  const int rows_here = cells_here * a.N;
  if (tid < 128) {
    // ge = p (gs + gp - cm), gp = y . ga + e gs      (softmax-weighted-sum backward)
    const long long m = s_m[tid];
    float ge = 0.f;
    if (m >= 0) {
      const int cell = s_cell[tid];
      const float gs = g.Gs[cell];
      const float gp = s_d[tid] + s_d[128 + tid] + g.E[m] * gs;
      ge = s_p[tid] * (gs + (gp - g.CM[cell]));   // the difference first: gs is tiny next to gp when the chart is not normalised
      if (a.no_norm) ge = gp;
    }
    s_ge[tid] = ge;
    if (a.no_norm) {
      // Without normalisation the chart grows geometrically and the softmax saturates: cm must cancel gp of the chosen
      // split exactly, so it is re-summed from the very gp values it is subtracted from (as autograd does).
      asm volatile("bar.sync 2, 128;" ::: "memory");
      const int base = (tid / a.N) * a.N;
      float cm2 = 0.f;
      if (m >= 0)
        for (int kk = 0; kk < a.N; ++kk) cm2 = fmaf(s_p[base + kk], s_ge[base + kk], cm2);
      asm volatile("bar.sync 2, 128;" ::: "memory");
      s_ge[tid] = m >= 0 ? s_p[tid] * (g.Gs[s_cell[tid]] + (ge - cm2)) : 0.f;
    }
  }
  // Column passes: a narrow slice (<= 112 columns) fits the rings in one pass, a wide one takes two of 112 + the rest.
  int pass_cols = ncols > kNarrowUmmaN ? kNarrowUmmaN : ncols;              // multiple of 16 when it is not everything
  if (ncols > kNarrowUmmaN)                                                 // three [128][pitch] stages must fit the rings
    while (pass_cols > 16 && 3 * kRows * (((pass_cols >> 2) & 1) ? pass_cols : pass_cols + 4) * 4 > ring_bytes(a.n_umma))
      pass_cols -= 16;
  const int pch = pass_cols >> 2;
  const int pitch = (pch & 1) ? pass_cols : pass_cols + 4;    // pitch / 4 odd: conflict-free 16-byte row accesses
  float* s_gz = reinterpret_cast<float*>(smem);              // [128][pitch] masked GZ
  float* s_h = s_gz + kRows * pitch;                         // [128][pitch] h[first] slice
  float* s_v = s_h + kRows * pitch;                          // [128][pitch] V[second] slice
  for (int cbeg = 0; cbeg < ncols; cbeg += pass_cols) {
    const int cw = min(pass_cols, ncols - cbeg), cwch = cw >> 2;
    if (cbeg > 0) __syncthreads();                           // the previous pass has been scattered
    for (int idx = tid; idx < rows_here * cwch; idx += kThreads) {
      const int r = idx / cwch, ch = idx - r * cwch;
      const int col = cbeg + ch * 4;
      if (n0 + col < D) {
        cp_async16(s_h + r * pitch + ch * 4, g.h1 + (int64_t)s_g1[r] * D + n0 + col);
        cp_async16(s_v + r * pitch + ch * 4, g.P2 + (int64_t)s_g2[r] * g.ld2 + g.off_v2 + n0 + col);
      }
    }
    cp_async_commit();
    if (warp >= 2) {
      // GZ = acc * [z > 0] -> shared memory, row-major
      constexpr int kSub = (kThreads / 32 - 2) / 4;
      const int qd = warp & 3, r = qd * 32 + lane, sub = (warp - 2) >> 2;
      const uint32_t* mrow = s_mask + r * 8;
#pragma unroll 1
      for (int c0 = cbeg + sub * 16; c0 < min(cbeg + cw, a.n_umma); c0 += 16 * kSub) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, v);
        if (a.mode == 2 && !a.single_acc) {
          float x[16];
          tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(a.n_umma + c0), x);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += x[i];
        }
        const uint32_t bits16 = (mrow[c0 >> 5] >> (c0 & 31)) & 0xffffu;      // c0 is a multiple of 16
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int col = c0 + j;
          if (col < ncols) {
            const uint32_t b = bits16 >> j;
            st4(s_gz + r * pitch + col - cbeg, make_float4((b & 1u) ? v[j] : 0.f, (b & 2u) ? v[j + 1] : 0.f,
                                                           (b & 4u) ? v[j + 2] : 0.f, (b & 8u) ? v[j + 3] : 0.f));
          }
        }
      }
    }
    cp_async_wait_all();
    tcgen05_fence_before();
    __syncthreads();
    if (cbeg == 0) LB_STAMP(4);
    // scatter (red.global.add, as split_scatter did): GP[first].A += GZ, GP[second].A += GZ, Gh[first] += ge V[second],
    // GP[second].V += ge h[first], Gs[first] += ge, Gs[second] += ge.  Warp per row, lanes over the 16-byte chunks.
    for (int r = warp; r < rows_here; r += kThreads / 32) {
      const int64_t g1 = s_g1[r], g2 = s_g2[r];
      const float ge = s_ge[r];
      float* d1 = g.GP1 + g1 * g.ld1 + g.off_a1 + n0 + cbeg;
      float* d2 = g.GP2 + g2 * g.ld2 + g.off_a2 + n0 + cbeg;
      float* gh = g.Gh1 + g1 * D + n0 + cbeg;
      float* gv = g.GP2 + g2 * g.ld2 + g.off_v2 + n0 + cbeg;
      for (int ch = lane; ch < cwch; ch += 32) {
        if (n0 + cbeg + ch * 4 < D) {
          const float4 gz = ld4(s_gz + r * pitch + ch * 4);
          const float4 hv = ld4(s_h + r * pitch + ch * 4), vv = ld4(s_v + r * pitch + ch * 4);
          red_add4(d1 + ch * 4, gz);
          red_add4(d2 + ch * 4, gz);
          red_add4(gh + ch * 4, make_float4(ge * vv.x, ge * vv.y, ge * vv.z, ge * vv.w));
          red_add4(gv + ch * 4, make_float4(ge * hv.x, ge * hv.y, ge * hv.z, ge * hv.w));
        }
      }
      if (lane == 0 && rank == 0 && cbeg == 0) {
        atomicAdd(g.Gs1 + g1, ge);
        atomicAdd(g.Gs2 + g2, ge);
      }
    }
  }
  // db2 += the column sums of the GY k-blocks this CTA streamed out
  if (g.db2 != nullptr && g.GYp != nullptr)
    for (int j = tid; j < D; j += kThreads)
      if (((j >> 5) % nc) == rank) atomicAdd(g.db2 + j, s_db2[j]);
  __syncthreads();
  LB_STAMP(5);
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (dbg_row && tid == 0) { dbg_row[18] = clock_now(); dbg_row[31] = global_ns(); }
#undef LB_STAMP
}

// ---------------------------------------------------------------- host side
struct LevelGeom {
  int nc, ncols, n_umma;
};
// wide = false: column slices of <= 112 (latency-optimal: four CTAs share a tile of a 400-wide level);
// wide = true: slices of <= 208 (half the CTAs and half the replicated operand gathers per tile: for levels that do not
// fit one wave anyway).  Wide tiles accumulate the 3xTF32 cross terms into the main accumulator (tensor memory:
// n_umma + 256 operand columns <= 512).
inline bool level_geom(int D, LevelGeom& g, bool wide = false) {
  if (D < 32 || (D % 4) != 0) return false;
  const int cap = wide ? kMaxUmmaN : kNarrowUmmaN;
  // power-of-two clusters pack the GPCs best (measured: 26 clusters of 5 are co-resident on a B200, 33 of 4)
  int nc = 1;
  while (nc <= kMaxCluster && ceil_div(D, nc) > cap) nc *= 2;
  if (nc > kMaxCluster) return false;
  if (!wide && g_debug[11] > 0 && g_debug[11] <= kMaxCluster && ceil_div(D, g_debug[11]) <= cap) nc = g_debug[11];
  int ncols = ((ceil_div(D, nc) + 3) / 4) * 4;
  g.nc = nc;
  g.ncols = ncols;
  g.n_umma = ((ncols + 15) / 16) * 16;
  return g.n_umma <= cap;
}
// clusters of this shape that can be resident at once (cached per device and shape)
int max_active_clusters(int nc, size_t smem);
// cells per tile: whole cells only, G*N <= 128.  CLIORA adds one bound: the shared-memory staging of the region slices
// of every image a tile can span ((G-1)/L + 2 sentences) plus the logit exchange buffer (G*R*nc floats).  Otherwise the
// level's cells are spread over about one wave of clusters.
inline int level_cells_per_tile(int cells, int N, int L, int R, const LevelGeom& g, int target_tiles, int& max_sent) {
  int gmax = kRows / N;
  max_sent = 0;
  if (gmax < 1) return 0;
  if (target_tiles < 1) target_tiles = 1;
  int G = ceil_div(cells, target_tiles);
  if (G < 1) G = 1;
  if (G > gmax) G = gmax;
  if (R > 0) {
    const int64_t area = (int64_t)ring_bytes(g.n_umma) / 4;     // floats in the operand rings
    for (; G >= 1; --G) {
      max_sent = (G - 1) / L + 2;
      const int64_t need = (int64_t)kRows * (g.ncols + 4) + (int64_t)G * g.ncols + 3 * ((G * R + 3) & ~3) +
                           (int64_t)max_sent * R * (g.ncols + 4) + (int64_t)g.nc * G * R;
      if (need <= area) break;
    }
  }
  if ((int64_t)kRows * (g.ncols + 4) + (int64_t)G * g.ncols > (int64_t)ring_bytes(g.n_umma) / 4) return 0;
  return G;
}
inline size_t level_fwd_smem(int n_umma);
inline int max_active_clusters_query(int nc, size_t smem) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nc, 148, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (func_attr_at_least(reinterpret_cast<const void*>(level_fwd_kernel<true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaOccupancyMaxActiveClusters(&n, level_fwd_kernel<true>, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 132 / nc;
  }
  return n;
}
inline size_t level_fwd_smem(int n_umma) { return (size_t)ring_bytes(n_umma) + kExtraBytes; }

extern long long* g_level_dbg;   // debug: timeline buffer handed to the launch selected by g_debug[8] (level + 1) / g_debug[9] (outside)

// bf16 matrix [rows, K] (K contiguous) as a 3-D map {K, rows, 1}, box {64, box_rows, 1}, 128-byte swizzle
inline int make_bf16_map(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (fn == nullptr) return CLIORA_ERR_CUDA;
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, 1};
  cuuint64_t gstride[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * 2 * (cuuint64_t)rows};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "cuTensorMapEncodeTiled (bf16) failed (%d)", (int)r);
    return CLIORA_ERR_CUDA;
  }
  return CLIORA_OK;
}

inline int launch_level_fwd(cudaStream_t st, const LevelFwdArgs& a_in, const float* W2pair, const char* tag) {
  LevelFwdArgs a = a_in;
  a.exp_flags = g_debug[12];
  a.dbg = (g_level_dbg != nullptr && g_debug[8] == a.level + 1 && g_debug[9] == a.outside) ? g_level_dbg : nullptr;
  if (g_level_dbg != nullptr && g_debug[8] == 99 && g_debug[9] == a.outside)      // every level: 1024 rows of 128 per level
    a.dbg = g_level_dbg + (int64_t)a.level * 1024 * 128;
  CUtensorMap tmW;
  if (a.mode == 3) CL_TRY(make_bf16_map(&tmW, W2pair, a.D, a.D, a.n_umma));      // W2pair = the bf16 copy in this mode
  else
    CL_TRY(tc::make_pair_map(&tmW, W2pair, a.D, a.D, a.D, (int64_t)a.D * a.D, a.n_umma, CU_TENSOR_MAP_SWIZZLE_128B,
                             a.mode == 1 ? 1 : 2));
  const size_t smem = level_fwd_smem(a.n_umma);
  const bool a_tmem = (a.zmask == nullptr && g_debug[13] == 0) || a.mode == 3;      // the bit-mask variant keeps the A pair in shared memory
  const void* kern = a_tmem ? reinterpret_cast<const void*>(level_fwd_kernel<true>)
                            : reinterpret_cast<const void*>(level_fwd_kernel<false>);
  CL_CUDA(func_attr_at_least(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (g_carveout >= 0) apply_carveout(kern);
  const int tiles = ceil_div(a.cells, a.G);
  const double rows = (double)a.cells * a.N;
  ProfScope prof(st, tag, 2.0 * rows * a.D * a.D, 4.0 * rows * (5.0 * a.D + 3));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.nc, tiles, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.nc;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 2 : 1;
  if (a_tmem) cudaLaunchKernelEx(&cfg, level_fwd_kernel<true>, tmW, a);
  else cudaLaunchKernelEx(&cfg, level_fwd_kernel<false>, tmW, a);
  CL_CHECK_LAUNCH("level_fwd_kernel");
  return CLIORA_OK;
}

inline size_t level_bwd_smem(int n_umma) { return (size_t)ring_bytes(n_umma) + kBwdExtraBytes; }

inline int launch_level_bwd(cudaStream_t st, const LevelBwdArgs& g_in, const float* W2Tpair, const char* tag) {
  LevelBwdArgs g = g_in;
  g.geo.dbg = (g_level_dbg != nullptr && g_debug[8] == g.geo.level + 1 && g_debug[9] == 2 + g.geo.outside) ? g_level_dbg : nullptr;
  if (g_level_dbg != nullptr && g_debug[8] == 99 && g_debug[9] == 2 + g.geo.outside)
    g.geo.dbg = g_level_dbg + (int64_t)g.geo.level * 1024 * 128;
  const LevelFwdArgs& a = g.geo;
  CUtensorMap tmW;
  if (a.mode == 3) CL_TRY(make_bf16_map(&tmW, W2Tpair, a.D, a.D, a.n_umma));
  else
    CL_TRY(tc::make_pair_map(&tmW, W2Tpair, a.D, a.D, a.D, (int64_t)a.D * a.D, a.n_umma, CU_TENSOR_MAP_SWIZZLE_128B,
                             a.mode == 1 ? 1 : 2));
  const size_t smem = level_bwd_smem(a.n_umma);
  const void* kern = reinterpret_cast<const void*>(level_bwd_kernel<0>);
  CL_CUDA(func_attr_at_least(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (g_carveout >= 0) apply_carveout(kern);
  const int tiles = ceil_div(a.cells, a.G);
  const double rows = (double)a.cells * a.N;
  ProfScope prof(st, tag, 2.0 * rows * a.D * a.D, 4.0 * rows * (9.0 * a.D + 3));
  launch_k(level_bwd_kernel<0>, dim3(a.nc, tiles, 1), dim3(kThreads, 1, 1), smem, st, tmW, g);
  CL_CHECK_LAUNCH("level_bwd_kernel");
  return CLIORA_OK;
}

}  // namespace lvl
}  // namespace cliora
