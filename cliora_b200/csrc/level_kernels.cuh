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
#include "tc_gemm.cuh"

namespace cliora {
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
constexpr int kStages = 3;
constexpr int kProdWarps = 8;
constexpr int kProdWarp0 = 6;
constexpr int kThreads = (kProdWarp0 + kProdWarps) * 32;   // 448
constexpr int kMaxCluster = 8;
constexpr int kMaxUmmaN = 112;
constexpr int kABytes = kRows * 128;  // one 128 x 32 fp32 operand tile
constexpr int kXlFloats = 4096;       // region-logit exchange buffer (nc * G * R floats)
// extras after the pipeline stages: barriers (128 B), b1 (1024 f), b2 (128 f), e, p, nrm, nrm2 (4 x 128 f),
// three [8][128] exchange buffers, the logit exchange buffer
constexpr int kExtraFloats = 1024 + 128 + 4 * 128 + 3 * kMaxCluster * 128 + kXlFloats;
constexpr int kExtraBytes = 128 + 4 * kExtraFloats;

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

// value `v` into slot[idx] of every CTA of the cluster, then one arrival on each CTA's barrier
CL_D void xchg_put(float* slot, int idx, float v, int nc) {
  const uint32_t a = smem_u32(slot + idx);
  for (int c = 0; c < nc; ++c) st_cluster_f32(map_to_cta(a, (uint32_t)c), v);
}
CL_D void xchg_arrive(uint64_t* bar, int nc) {
  const uint32_t a = smem_u32(bar);
  for (int c = 0; c < nc; ++c) mbar_arrive_cluster(map_to_cta(a, (uint32_t)c));
}

struct LevelFwdArgs {
  int B, n, level, L, N, D, R;
  int G;         // cells per tile (G * N <= 128)
  int cells;     // B * L
  int ncols;     // output columns per CTA (multiple of 4)
  int n_umma;    // ncols rounded up to 16 (<= kMaxUmmaN)
  int nc;        // cluster size = column slices
  int mode;      // 2: fp32-accurate 3xTF32 (main + cross accumulator), 1: single TF32 pass
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
  float* Y; float* E; float* Pr;                      // level blocks [rows,D], [rows], [rows]
  float* chart_h; float* chart_s;                     // [B,C,D], [B,C]
  float* q; float* nrm; float* nrm2; float* att;      // saved per cell (q, nrm2, att: R > 0 only)
  const float* obj; const uint8_t* keep;
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

__global__ void __launch_bounds__(kThreads, 1)
level_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const LevelFwdArgs a) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int tile = blockIdx.y;
  const int n0 = rank * a.ncols;
  const int D = a.D;
  const int num_kb = (D + 31) / 32;
  const int b_bytes = a.n_umma * 128;
  const int stage_bytes = 2 * kABytes + 2 * b_bytes;
  const int cells_here = min(a.G, a.cells - tile * a.G);

  uint8_t* ex = smem + kStages * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ex);
  uint64_t* empty = full + kStages;
  uint64_t* tmem_full = empty + kStages;
  uint64_t* xbar = tmem_full + 1;                 // 4 single-use cluster exchange barriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 4);
  float* s_b1 = reinterpret_cast<float*>(ex + 128);
  float* s_b2 = s_b1 + 1024;
  float* s_e = s_b2 + 128;
  float* s_p = s_e + 128;
  float* s_nrm = s_p + 128;
  float* s_nrm2 = s_nrm + 128;
  float* s_xe = s_nrm2 + 128;                     // [nc][128] partial scores
  float* s_xs = s_xe + kMaxCluster * 128;         // [nc][128] partial |a|^2
  float* s_xs2 = s_xs + kMaxCluster * 128;        // [nc][128] partial |a2|^2
  float* s_xl = s_xs2 + kMaxCluster * 128;        // [nc][G*R] partial region logits
  const uint32_t tmem_cols_alloc = (uint32_t)tc::tmem_cols(2 * a.n_umma);

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmW);
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kStages; ++i) {
        mbar_init(&full[i], 1 + kProdWarps);      // the TMA thread's expect_tx arrival + one per producer warp
        mbar_init(&empty[i], 1);
      }
      mbar_init(tmem_full, 1);
      for (int i = 0; i < 4; ++i) mbar_init(&xbar[i], (uint32_t)a.nc * 128);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols_alloc)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 1024; i += kThreads) s_b1[i] = i < D ? a.b1[i] : 0.f;
  for (int i = tid; i < 128; i += kThreads) s_b2[i] = (i < a.ncols && n0 + i < D) ? a.b2[n0 + i] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  cluster_sync_all();          // every CTA's barriers exist before a peer may arrive on them
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ W2 slice: TMA
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(a.mode == 1 ? b_bytes : 2 * b_bytes);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % kStages;
        const uint32_t phase = (kb / kStages) & 1;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], tx);
        tma_load_3d(smem + stage * stage_bytes + 2 * kABytes, &tmW, &full[stage], kb * 32, n0, 0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issue
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(a.n_umma);
      uint32_t acc = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % kStages;
        const uint32_t phase = (kb / kStages) & 1;
        mbar_wait(&full[stage], phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint64_t a_hi = umma_desc_k_sw128(sa), a_lo = umma_desc_k_sw128(sa + kABytes);
        const uint64_t b_hi = umma_desc_k_sw128(sa + 2 * kABytes);
        const uint64_t b_lo = umma_desc_k_sw128(sa + 2 * kABytes + b_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (a.mode == 1) {
            umma_tf32(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);
          } else {
            umma_tf32(tmem_base + a.n_umma, a_lo + 2 * k, b_hi + 2 * k, idesc, acc);
            umma_tf32(tmem_base + a.n_umma, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
            umma_tf32(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, acc);
          }
          acc = 1;
        }
        umma_commit(&empty[stage]);
      }
      umma_commit(tmem_full);
    }
  } else if (warp >= kProdWarp0) {
    // ------------------------------------------------------------ A operand: gather + ReLU + tf32 split
    const int pt = tid - kProdWarp0 * 32;      // 0..255
    const int c = pt & 7;                      // 16-byte chunk of the 128-byte k-block row
    const int rw = lane >> 3;                  // row within the warp's group of four
    const int rbase = (pt >> 3);               // rows rbase + 32 i, i < 4
    const float* pa[4];
    const float* pb[4];
    int64_t mrow[4];
    bool ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const RowInfo ri = decode_row(a, tile, cells_here, rbase + 32 * i);
      ok[i] = ri.ok;
      mrow[i] = ri.m;
      pa[i] = a.P1 + ri.g1 * a.ld1 + a.off_a1 + c * 4;
      pb[i] = a.P2 + ri.g2 * a.ld2 + a.off_a2 + c * 4;
    }
    const bool write_mask = a.zmask != nullptr && rank == 0;
    uint32_t mw[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) mw[i][j] = 0u;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 xa[4], xb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool ld = ok[i] && c * 4 < D;
      xa[i] = ld ? ldcg4(pa[i]) : zero4;
      xb[i] = ld ? ldcg4(pb[i]) : zero4;
    }
    for (int kb = 0; kb < num_kb; ++kb) {
      const int stage = kb % kStages;
      const uint32_t phase = (kb / kStages) & 1;
      const int kcol = kb * 32 + c * 4;
      float4 na[4], nb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {             // next k-block's rows are in flight while this one is processed
        const bool ld = ok[i] && (kb + 1 < num_kb) && (kcol + 32 < D);
        na[i] = ld ? ldcg4(pa[i] + (kb + 1) * 32) : zero4;
        nb[i] = ld ? ldcg4(pb[i] + (kb + 1) * 32) : zero4;
      }
      const float4 bv = *reinterpret_cast<const float4*>(s_b1 + kcol);
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* sA = smem + stage * stage_bytes;
      const bool store_z = a.Z != nullptr && (kb % a.nc) == rank && kcol < D;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rbase + 32 * i;
        float4 o;
        o.x = fmaxf(xa[i].x + xb[i].x + bv.x, 0.f);
        o.y = fmaxf(xa[i].y + xb[i].y + bv.y, 0.f);
        o.z = fmaxf(xa[i].z + xb[i].z + bv.z, 0.f);
        o.w = fmaxf(xa[i].w + xb[i].w + bv.w, 0.f);
        if (!ok[i] || kcol >= D) o = zero4;
        float4 hi, lo;
        split_tf32(o.x, hi.x, lo.x); split_tf32(o.y, hi.y, lo.y); split_tf32(o.z, hi.z, lo.z); split_tf32(o.w, hi.w, lo.w);
        const int off = r * 128 + ((c ^ (r & 7)) << 4);
        *reinterpret_cast<float4*>(sA + off) = hi;
        if (a.mode != 1) *reinterpret_cast<float4*>(sA + kABytes + off) = lo;
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
      fence_proxy_async_smem();                 // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive_local(&full[stage]);
      if (write_mask && ((kb & 3) == 3 || kb == num_kb - 1)) {
        const int t = kb >> 2;                  // 128-column group
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c == 0 && ok[i] && t < 4)
            *reinterpret_cast<uint4*>(a.zmask + mrow[i] * 16 + t * 4) = make_uint4(mw[i][0], mw[i][1], mw[i][2], mw[i][3]);
          mw[i][0] = mw[i][1] = mw[i][2] = mw[i][3] = 0u;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { xa[i] = na[i]; xb[i] = nb[i]; }
    }
  } else {
    // ------------------------------------------------------------ warps 2-5: scores, softmax, epilogue, cell finalize
    const int qd = warp & 3;                    // TMEM lane quarter this warp may read
    const int r = qd * 32 + lane;               // split row of the tile = TMEM lane
    const RowInfo ri = decode_row(a, tile, cells_here, r);
    const int nc = a.nc, ncols = a.ncols, N = a.N;
    // (a) partial bilinear score over this CTA's columns, all-gathered over the cluster
    float part = 0.f;
    if (ri.ok) {
      const float* hp = a.h1 + ri.g1 * D + n0;
      const float* vp = a.P2 + ri.g2 * a.ld2 + a.off_v2 + n0;
#pragma unroll 5
      for (int j = 0; j < ncols; j += 4) {
        if (n0 + j < D) {
          const float4 hv = ldcg4(hp + j), vv = ldcg4(vp + j);
          part = fmaf(hv.x, vv.x, part); part = fmaf(hv.y, vv.y, part);
          part = fmaf(hv.z, vv.z, part); part = fmaf(hv.w, vv.w, part);
        }
      }
    }
    xchg_put(s_xe + rank * 128, r, part, nc);
    xchg_arrive(&xbar[0], nc);
    float e = 0.f;
    if (ri.ok) e = a.s1[ri.g1] + a.s2[ri.g2];
    mbar_wait_cluster(&xbar[0], 0);
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
    // (b) epilogue: y = relu(acc + b2) -> Y row; p * y staged for the per-cell sums
    mbar_wait(tmem_full, 0);
    tcgen05_fence_after();
    const int stride = ncols + 4;
    float* s_stage = reinterpret_cast<float*>(smem);           // pipeline stages are free now
    float* s_a = s_stage + kRows * stride;                      // [G][ncols]
    float* yrow = ri.ok ? a.Y + ri.m * D + n0 : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < a.n_umma; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, v);    // warp-collective: no early exit
      if (a.mode != 1) {
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
          else if (yrow) st4(yrow + col, o);
          st4(s_stage + r * stride + col, make_float4(p * o.x, p * o.y, p * o.z, p * o.w));
        }
      }
    }
    tcgen05_fence_before();
    epi_bar_sync();
    // per-cell sums over the N splits (fixed order: deterministic)
    const int nc4 = ncols >> 2;
    for (int item = r; item < cells_here * nc4; item += 128) {
      const int g = item / nc4, j4 = item - g * nc4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* src = s_stage + (g * N) * stride + j4 * 4;
      for (int kk = 0; kk < N; ++kk) {
        const float4 t = ld4(src + kk * stride);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      st4(s_a + g * ncols + j4 * 4, acc);
    }
    epi_bar_sync();
    int cb = 0, cp = 0;
    int64_t ccell = 0;
    if (r < cells_here) {
      cell_of(a, tile * a.G + r, cb, cp, ccell);
      float ss = 0.f;
      for (int j = 0; j < ncols; ++j) {
        const float t = s_a[r * ncols + j];
        ss = fmaf(t, t, ss);
      }
      xchg_put(s_xs + rank * 128, r, ss, nc);
    }
    xchg_arrive(&xbar[1], nc);
    mbar_wait_cluster(&xbar[1], 0);
    if (r < cells_here) {
      float tot = 0.f;
      for (int cc = 0; cc < nc; ++cc) tot += s_xs[cc * 128 + r];
      const float nrm = fmaxf(sqrtf(tot), kTiny);
      s_nrm[r] = nrm;
      if (rank == 0) a.nrm[ccell] = nrm;
    }
    epi_bar_sync();
    const bool vl = a.R > 0;
    for (int item = r; item < cells_here * nc4; item += 128) {
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
    if (vl) {
      // region attention of every cell against ITS OWN image only (the reference computes all B x B pairs and
      // keeps the diagonal, cliora.py:35-42)
      const int R = a.R, GRs = a.G * R;
      float* s_att = s_a + kRows * ncols;       // [G][R]
      float* s_patt = s_att + GRs;              // [G][R]
      epi_bar_sync();
      for (int item = r; item < cells_here * R; item += 128) {
        const int g = item / R, rr = item - g * R;
        const int bg = (tile * a.G + g) / a.L;
        const float* op = a.obj + ((int64_t)bg * R + rr) * D + n0;
        const float* qp = s_a + g * ncols;
        float d = 0.f;
#pragma unroll 5
        for (int j = 0; j < ncols; j += 4) {
          if (n0 + j < D) {
            const float4 qv = ld4(qp + j), ov = ldcg4(op + j);
            d = fmaf(qv.x, ov.x, d); d = fmaf(qv.y, ov.y, d); d = fmaf(qv.z, ov.z, d); d = fmaf(qv.w, ov.w, d);
          }
        }
        xchg_put(s_xl + rank * GRs, item, d, nc);
      }
      xchg_arrive(&xbar[2], nc);
      mbar_wait_cluster(&xbar[2], 0);
      if (r < cells_here) {
        float mx = -INFINITY;
        for (int rr = 0; rr < R; ++rr) {
          float lg = 0.f;
          for (int cc = 0; cc < nc; ++cc) lg += s_xl[cc * GRs + r * R + rr];
          s_att[r * R + rr] = lg;
          mx = fmaxf(mx, lg);
        }
        float sum = 0.f;
        for (int rr = 0; rr < R; ++rr) {
          const float ex2 = expf(s_att[r * R + rr] - mx);
          s_att[r * R + rr] = ex2;
          sum += ex2;
        }
        const float inv = 1.f / sum;
        for (int rr = 0; rr < R; ++rr) {
          const float at = s_att[r * R + rr] * inv;
          if (rank == 0) a.att[ccell * R + rr] = at;
          float sc = 1.f;
          if (a.keep != nullptr) sc = a.keep[ccell * R + rr] ? kKeepScale : 0.f;
          s_patt[r * R + rr] = at * sc;
        }
      }
      epi_bar_sync();
      for (int item = r; item < cells_here * nc4; item += 128) {
        const int g = item / nc4, j4 = item - g * nc4;
        float4 t = ld4(s_a + g * ncols + j4 * 4);
        if (n0 + j4 * 4 < D) {
          const int bg = (tile * a.G + g) / a.L;
          const float* op = a.obj + (int64_t)bg * R * D + n0 + j4 * 4;
          const float* wp = s_patt + g * R;
          for (int rr = 0; rr < R; ++rr) {
            const float w = wp[rr];
            const float4 ov = ldcg4(op + (int64_t)rr * D);
            t.x = fmaf(w, ov.x, t.x); t.y = fmaf(w, ov.y, t.y); t.z = fmaf(w, ov.z, t.z); t.w = fmaf(w, ov.w, t.w);
          }
        }
        st4(s_a + g * ncols + j4 * 4, t);
      }
      epi_bar_sync();
      if (r < cells_here) {
        float ss = 0.f;
        for (int j = 0; j < ncols; ++j) {
          const float t = s_a[r * ncols + j];
          ss = fmaf(t, t, ss);
        }
        xchg_put(s_xs2 + rank * 128, r, ss, nc);
      }
      xchg_arrive(&xbar[3], nc);
      mbar_wait_cluster(&xbar[3], 0);
      if (r < cells_here) {
        float tot = 0.f;
        for (int cc = 0; cc < nc; ++cc) tot += s_xs2[cc * 128 + r];
        const float nrm2 = fmaxf(sqrtf(tot), kTiny);
        s_nrm2[r] = nrm2;
        if (rank == 0) a.nrm2[ccell] = nrm2;
      }
      epi_bar_sync();
      for (int item = r; item < cells_here * nc4; item += 128) {
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
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols_alloc) : "memory");
  }
  cluster_sync_all();          // no CTA leaves while a peer may still store into its shared memory
}

// ---------------------------------------------------------------- host side
struct LevelGeom {
  int nc, ncols, n_umma;
};
inline bool level_geom(int D, LevelGeom& g) {
  if (D < 32 || (D % 4) != 0) return false;
  int nc = ceil_div(D, 80);
  if (nc > kMaxCluster) nc = ceil_div(D, kMaxUmmaN);
  if (nc > kMaxCluster) return false;
  int ncols = ((ceil_div(D, nc) + 3) / 4) * 4;
  g.nc = nc;
  g.ncols = ncols;
  g.n_umma = ((ncols + 15) / 16) * 16;
  return g.n_umma <= kMaxUmmaN;
}
// cells per tile: whole cells only, G*N <= 128; for CLIORA the logit exchange buffer bounds G*R*nc; otherwise spread the
// level's cells over about one wave of clusters
inline int level_cells_per_tile(int cells, int N, int R, int nc) {
  int gmax = kRows / N;
  if (R > 0) {
    const int gv = kXlFloats / (R * nc);
    if (gv < gmax) gmax = gv;
  }
  if (gmax < 1) return 0;
  const int target_tiles = 132 / nc > 0 ? 132 / nc : 1;
  int G = ceil_div(cells, target_tiles);
  if (G < 1) G = 1;
  if (G > gmax) G = gmax;
  return G;
}
inline size_t level_fwd_smem(int n_umma) { return (size_t)kStages * (2 * kABytes + 2 * n_umma * 128) + kExtraBytes + 1024; }

inline int launch_level_fwd(cudaStream_t st, const LevelFwdArgs& a, const float* W2pair, const char* tag) {
  CUtensorMap tmW;
  CL_TRY(tc::make_pair_map(&tmW, W2pair, a.D, a.D, a.D, (int64_t)a.D * a.D, a.n_umma, CU_TENSOR_MAP_SWIZZLE_128B,
                           a.mode == 1 ? 1 : 2));
  const size_t smem = level_fwd_smem(a.n_umma);
  const void* kern = reinterpret_cast<const void*>(level_fwd_kernel);
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
  cudaLaunchKernelEx(&cfg, level_fwd_kernel, tmW, a);
  CL_CHECK_LAUNCH("level_fwd_kernel");
  return CLIORA_OK;
}

}  // namespace lvl
}  // namespace cliora
