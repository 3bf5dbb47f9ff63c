// Vision-language cell kernels, warp-per-cell variant.
//
// The block-per-cell kernels in chart_kernels.cuh make every cell re-read its image's R region vectors
// (R*D*4 = 57.6 KB at R=36, D=400) from L2 twice; at 640 cells per level that is ~74 MB of L2 traffic per
// level and ~22 us.  Here one CTA owns up to 8 cells of the SAME sentence: the image's regions are staged
// once in shared memory and each warp then does everything for one cell with warp shuffles only (no block
// barrier after the staging).  Lane l owns columns j = 4*l + 128*t, t < 4, so D <= 512.
#pragma once
#include "chart_kernels.cuh"

namespace cliora {

#ifndef CLIORA_CELLS_PER_CTA
#define CLIORA_CELLS_PER_CTA 8
#endif
constexpr int kCellsPerCta = CLIORA_CELLS_PER_CTA;
constexpr int kCellThreads = kCellsPerCta * 32;
constexpr int kColT = 4;   // float4 column groups per lane

CL_D float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
CL_D void fma4(float4& acc, float w, const float4& v) {
  acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
}

// Dot products of one D-vector (register fragments v[t]) with the R staged region vectors.  Returns the logits
// in the "lane r % 32 owns region r" layout (lg0: r < 32, lg1: r >= 32).  The first 32 regions are reduced with a
// transposing butterfly (31 shuffles in total instead of 5 per region); the few beyond 32 with plain reductions.
CL_D void region_dots(const float4 (&v)[kColT], const float* __restrict__ s_obj, int R, int D, int lane, float& lg0,
                      float& lg1) {
  float d[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    float acc = 0.f;
    if (r < R) {
#pragma unroll
      for (int t = 0; t < kColT; ++t) {
        const int j = lane * 4 + t * 128;
        if (j < D) acc += dot4(v[t], ld4(s_obj + r * D + j));
      }
    }
    d[r] = acc;
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? d[i] : d[i + s];
      const float keep = upper ? d[i + s] : d[i];
      d[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  lg0 = d[0];            // lane l now holds the full sum of region l
  lg1 = 0.f;
  for (int r = 32; r < R; ++r) {
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < kColT; ++t) {
      const int j = lane * 4 + t * 128;
      if (j < D) acc += dot4(v[t], ld4(s_obj + r * D + j));
    }
    acc = warp_sum(acc);
    if ((r & 31) == lane) lg1 = acc;
  }
}

// forward: softmax over splits, weighted sum, normalise, region attention, second normalise
// dynamic smem: (VL ? R*D : 0) + 8*N floats
template <bool VL>
__global__ __launch_bounds__(kCellThreads) void cell_fwd_warp_kernel(const CellArgs a) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  float* s_obj = sm;
  float* s_pbase = sm + (VL ? a.R * a.D : 0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunks = (a.L + kCellsPerCta - 1) / kCellsPerCta;
  const int b = blockIdx.x / chunks;
  const int p = (blockIdx.x % chunks) * kCellsPerCta + warp;
  if (VL) {   // stage the image's regions asynchronously; they are first needed in step 3
    const float* obj = a.obj + (int64_t)b * a.R * a.D;
    for (int i = tid * 4; i < a.R * a.D; i += kCellThreads * 4) cp_async16(s_obj + i, obj + i);
    cp_async_commit();
  }
  const bool active = p < a.L;
  float* s_p = s_pbase + warp * a.N;
  const int64_t cell = (int64_t)b * a.C + lvl_off(a.n, a.level) + (active ? p : 0);
  const int64_t row0 = (int64_t)b * a.L * a.N + (int64_t)(active ? p : 0) * a.sp;
  float4 q[kColT];
  if (active) {
    // 1. softmax over the N splits (lanes over k)
    float sbar = 0.f;
    if (a.E == nullptr) {
      if (lane == 0) s_p[0] = 1.f;
    } else {
      float mx = -INFINITY;
      for (int k = lane; k < a.N; k += 32) mx = fmaxf(mx, a.E[row0 + (int64_t)k * a.sk]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int k = lane; k < a.N; k += 32) {
        const float ex = expf(a.E[row0 + (int64_t)k * a.sk] - mx);
        s_p[k] = ex;
        sum += ex;
      }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      for (int k = lane; k < a.N; k += 32) {
        const float pk = s_p[k] * inv;
        s_p[k] = pk;
        a.Pr[row0 + (int64_t)k * a.sk] = pk;
        sbar = fmaf(pk, a.E[row0 + (int64_t)k * a.sk], sbar);
      }
      sbar = warp_sum(sbar);
    }
    if (lane == 0) a.chart_s[cell] = sbar;
    __syncwarp();

    // 2. a = sum_k p_k y_k   (four rows = 16 independent 16-byte loads per lane in flight)
    float4 acc[kColT];
#pragma unroll
    for (int t = 0; t < kColT; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < a.N; k += 4) {
      float pk[4];
      const float* yk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = k + u < a.N;
        pk[u] = ok ? s_p[k + u] : 0.f;
        yk[u] = a.Y + (row0 + (int64_t)(ok ? k + u : k) * a.sk) * a.D;
      }
      float4 v[4][kColT];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < kColT; ++t) {
          const int j = lane * 4 + t * 128;
          if (j < a.D) v[u][t] = ld4(yk[u] + j);
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < kColT; ++t)
          if (lane * 4 + t * 128 < a.D) fma4(acc[t], pk[u], v[u][t]);
    }
    float ss = 0.f;
#pragma unroll
    for (int t = 0; t < kColT; ++t)
      if (lane * 4 + t * 128 < a.D) ss += dot4(acc[t], acc[t]);
    ss = warp_sum(ss);
    const float nrm = a.no_norm ? 1.f : fmaxf(sqrtf(ss), kTiny);
    const float inv_nrm = 1.f / nrm;
#pragma unroll
    for (int t = 0; t < kColT; ++t) {
      const int j = lane * 4 + t * 128;
      q[t] = make_float4(acc[t].x * inv_nrm, acc[t].y * inv_nrm, acc[t].z * inv_nrm, acc[t].w * inv_nrm);
      if (j < a.D) st4((VL ? a.q : a.chart_h) + cell * a.D + j, q[t]);
    }
    if (lane == 0) a.nrm[cell] = a.no_norm ? -1.f : nrm;
  }
  if (!VL) return;
  cp_async_wait_all();
  __syncthreads();
  if (!active) return;

  // 3. attention against the staged regions: logits owned by lane r % 32 (R <= 64)
  float lg0, lg1;
  region_dots(q, s_obj, a.R, a.D, lane, lg0, lg1);
  if (lane >= a.R) lg0 = -INFINITY;
  if (lane + 32 >= a.R) lg1 = -INFINITY;
  const float mx = warp_max(fmaxf(lg0, lg1));
  const float e0 = (lane < a.R) ? expf(lg0 - mx) : 0.f;
  const float e1 = (lane + 32 < a.R) ? expf(lg1 - mx) : 0.f;
  const float inv = 1.f / warp_sum(e0 + e1);
  const float at0 = e0 * inv, at1 = e1 * inv;
  float pa0 = at0, pa1 = at1;
  if (lane < a.R) {
    a.att[cell * a.R + lane] = at0;
    if (a.keep != nullptr) pa0 = a.keep[cell * a.R + lane] ? at0 * kKeepScale : 0.f;
  }
  if (lane + 32 < a.R) {
    a.att[cell * a.R + lane + 32] = at1;
    if (a.keep != nullptr) pa1 = a.keep[cell * a.R + lane + 32] ? at1 * kKeepScale : 0.f;
  }
  // a2 = q + sum_r patt_r obj_r
  for (int r = 0; r < a.R; ++r) {
    const float w = __shfl_sync(0xffffffffu, r < 32 ? pa0 : pa1, r & 31);
#pragma unroll
    for (int t = 0; t < kColT; ++t) {
      const int j = lane * 4 + t * 128;
      if (j < a.D) fma4(q[t], w, ld4(s_obj + r * a.D + j));
    }
  }
  float ss2 = 0.f;
#pragma unroll
  for (int t = 0; t < kColT; ++t)
    if (lane * 4 + t * 128 < a.D) ss2 += dot4(q[t], q[t]);
  ss2 = warp_sum(ss2);
  const float nrm2 = a.no_norm ? 1.f : fmaxf(sqrtf(ss2), kTiny);
  const float inv2 = 1.f / nrm2;
#pragma unroll
  for (int t = 0; t < kColT; ++t) {
    const int j = lane * 4 + t * 128;
    if (j < a.D)
      st4(a.chart_h + cell * a.D + j, make_float4(q[t].x * inv2, q[t].y * inv2, q[t].z * inv2, q[t].w * inv2));
  }
  if (lane == 0) a.nrm2[cell] = a.no_norm ? -1.f : nrm2;
}

// backward of the above + per-split gradients (same maths as cell_bwd_kernel)
// phase 1: warp per cell (attention / normalise backward against the staged regions) -> ga in shared memory
// phase 2: all (cell, split) items of the CTA spread over the 8 warps
// dynamic smem: (VL ? R*D : 0) + 8*D + 32 floats
template <bool VL>
__global__ __launch_bounds__(kCellThreads) void cell_bwd_warp_kernel(const CellBwdArgs g) {
  pdl_prologue();
  const CellArgs& a = g.c;
  extern __shared__ __align__(16) float sm[];
  float* s_obj = sm;
  float* s_ga = sm + (VL ? a.R * a.D : 0);            // [8][D]
  float* s_sc = s_ga + kCellsPerCta * a.D;             // [8][2]: gs, cm
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunks = (a.L + kCellsPerCta - 1) / kCellsPerCta;
  const int b = blockIdx.x / chunks;
  const int p0 = (blockIdx.x % chunks) * kCellsPerCta;
  const int p = p0 + warp;
  const bool active = p < a.L;
  if (VL) {
    const float* obj = a.obj + (int64_t)b * a.R * a.D;
    for (int i = tid * 4; i < a.R * a.D; i += kCellThreads * 4) cp_async16(s_obj + i, obj + i);
    cp_async_commit();
    cp_async_wait_all();   // the cell's own vectors are tiny: nothing worth overlapping before the first use
    __syncthreads();
  }
  if (active) {
    const int64_t cell = (int64_t)b * a.C + lvl_off(a.n, a.level) + p;
    const float nrm_raw = a.nrm[cell];
    const float nrm = fabsf(nrm_raw);
    float4 gv[kColT], hv[kColT], qv[kColT];
    float hd = 0.f;
#pragma unroll
    for (int t = 0; t < kColT; ++t) {
      const int j = lane * 4 + t * 128;
      gv[t] = hv[t] = qv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < a.D) {
        gv[t] = ld4(g.Gh + cell * a.D + j);
        hv[t] = ld4(a.chart_h + cell * a.D + j);
        qv[t] = VL ? ld4(a.q + cell * a.D + j) : hv[t];
        hd += dot4(hv[t], gv[t]);
      }
    }
    hd = warp_sum(hd);
    if (VL) {
      const float nrm2_raw = a.nrm2[cell];
      const float coef = unit_bwd_coef(nrm2_raw, hd);
      const float inv2 = 1.f / fabsf(nrm2_raw);
#pragma unroll
      for (int t = 0; t < kColT; ++t) {
        const int j = lane * 4 + t * 128;
        gv[t] = make_float4((gv[t].x - hv[t].x * coef) * inv2, (gv[t].y - hv[t].y * coef) * inv2,
                            (gv[t].z - hv[t].z * coef) * inv2, (gv[t].w - hv[t].w * coef) * inv2);   // ga2
        if (j < a.D) st4(g.GA2 + cell * a.D + j, gv[t]);
      }
      float ga0, ga1;   // g_att_r = (ga2 . obj_r) * scale_r, owned by lane r % 32
      region_dots(gv, s_obj, a.R, a.D, lane, ga0, ga1);
      if (lane >= a.R) ga0 = 0.f;
      if (lane + 32 >= a.R) ga1 = 0.f;
      float at0 = 0.f, at1 = 0.f, sc0 = 1.f, sc1 = 1.f;
      if (lane < a.R) {
        at0 = a.att[cell * a.R + lane];
        if (a.keep != nullptr) sc0 = a.keep[cell * a.R + lane] ? kKeepScale : 0.f;
      }
      if (lane + 32 < a.R) {
        at1 = a.att[cell * a.R + lane + 32];
        if (a.keep != nullptr) sc1 = a.keep[cell * a.R + lane + 32] ? kKeepScale : 0.f;
      }
      ga0 *= sc0;
      ga1 *= sc1;
      const float dsum = warp_sum(at0 * ga0 + at1 * ga1);
      const float gl0 = at0 * (ga0 - dsum), gl1 = at1 * (ga1 - dsum);
      if (lane < a.R) {
        g.coef[(cell * 2) * a.R + lane] = at0 * sc0;
        g.coef[(cell * 2 + 1) * a.R + lane] = gl0;
      }
      if (lane + 32 < a.R) {
        g.coef[(cell * 2) * a.R + lane + 32] = at1 * sc1;
        g.coef[(cell * 2 + 1) * a.R + lane + 32] = gl1;
      }
      for (int r = 0; r < a.R; ++r) {   // gq = ga2 + sum_r g_logit_r obj_r
        const float w = __shfl_sync(0xffffffffu, r < 32 ? gl0 : gl1, r & 31);
#pragma unroll
        for (int t = 0; t < kColT; ++t) {
          const int j = lane * 4 + t * 128;
          if (j < a.D) fma4(gv[t], w, ld4(s_obj + r * a.D + j));
        }
      }
      hd = 0.f;
#pragma unroll
      for (int t = 0; t < kColT; ++t)
        if (lane * 4 + t * 128 < a.D) hd += dot4(qv[t], gv[t]);
      hd = warp_sum(hd);
    }
    // ga = unit_bwd(gq, q, nrm)
    const float coef = unit_bwd_coef(nrm_raw, hd);
    const float inv = 1.f / nrm;
    float ad = 0.f;
#pragma unroll
    for (int t = 0; t < kColT; ++t) {
      gv[t] = make_float4((gv[t].x - qv[t].x * coef) * inv, (gv[t].y - qv[t].y * coef) * inv,
                          (gv[t].z - qv[t].z * coef) * inv, (gv[t].w - qv[t].w * coef) * inv);
      if (lane * 4 + t * 128 < a.D) ad += dot4(qv[t], gv[t]);
    }
    ad = warp_sum(ad);
    if (a.E == nullptr) {   // leaf: gu = ga * (1 - t^2)
      const int64_t row = (int64_t)b * a.L + p;
#pragma unroll
      for (int t = 0; t < kColT; ++t) {
        const int j = lane * 4 + t * 128;
        if (j < a.D) {
          const float4 tv = ld4(g.leaf_t + row * a.D + j);
          st4(g.gu + row * a.D + j, make_float4(gv[t].x * (1.f - tv.x * tv.x), gv[t].y * (1.f - tv.y * tv.y),
                                                 gv[t].z * (1.f - tv.z * tv.z), gv[t].w * (1.f - tv.w * tv.w)));
        }
      }
    } else {
#pragma unroll
      for (int t = 0; t < kColT; ++t) {
        const int j = lane * 4 + t * 128;
        if (j < a.D) st4(s_ga + warp * a.D + j, gv[t]);
      }
      if (lane == 0) {
        const float gs = g.Gs[cell];
        s_sc[warp * 2] = gs;
        s_sc[warp * 2 + 1] = nrm * ad + a.chart_s[cell] * gs;   // sum_m p_m gp_m
        if (g.ga_out != nullptr) g.cm_out[cell] = s_sc[warp * 2 + 1];
      }
      if (g.ga_out != nullptr) {     // cells-only mode: the fused level kernel does the per-split part
#pragma unroll
        for (int t = 0; t < kColT; ++t) {
          const int j = lane * 4 + t * 128;
          if (j < a.D) st4(g.ga_out + cell * a.D + j, gv[t]);
        }
      }
    }
  }
  if (a.E == nullptr || g.ga_out != nullptr) return;   // uniform over the CTA
  __syncthreads();
  const int ncell = min(kCellsPerCta, a.L - p0);
  const int items = ncell * a.N;
  // (cell, split) items over the 8 warps, four items (16 independent 16-byte loads per lane) in flight per warp
  for (int it0 = warp; it0 < items; it0 += 4 * kCellsPerCta) {
    float4 yv[4][kColT];
    int64_t row[4];
    int cw[4];
    float pk[4], ev[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int it = it0 + u * kCellsPerCta;
      const bool ok = it < items;            // warp-uniform
      cw[u] = ok ? it / a.N : 0;
      const int k = ok ? it % a.N : 0;
      row[u] = (int64_t)b * a.L * a.N + (int64_t)(p0 + cw[u]) * a.sp + (int64_t)k * a.sk;
      if (ok) {
        const float* y = a.Y + row[u] * a.D;
#pragma unroll
        for (int t = 0; t < kColT; ++t) {
          const int j = lane * 4 + t * 128;
          if (j < a.D) yv[u][t] = ld4(y + j);
        }
        pk[u] = a.Pr[row[u]];
        ev[u] = a.E[row[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (it0 + u * kCellsPerCta >= items) break;
      float* y = a.Y + row[u] * a.D;
      const float gs = s_sc[cw[u] * 2], cm = s_sc[cw[u] * 2 + 1];
      float d = 0.f;
#pragma unroll
      for (int t = 0; t < kColT; ++t) {
        const int j = lane * 4 + t * 128;
        if (j < a.D) {
          const float4 gg = ld4(s_ga + cw[u] * a.D + j);
          d += dot4(yv[u][t], gg);
          float4 o;
          o.x = yv[u][t].x > 0.f ? pk[u] * gg.x : 0.f;
          o.y = yv[u][t].y > 0.f ? pk[u] * gg.y : 0.f;
          o.z = yv[u][t].z > 0.f ? pk[u] * gg.z : 0.f;
          o.w = yv[u][t].w > 0.f ? pk[u] * gg.w : 0.f;
          if (a.y_lo_off != 0) {
            float4 hi, lo;
            split_tf32(o.x, hi.x, lo.x); split_tf32(o.y, hi.y, lo.y); split_tf32(o.z, hi.z, lo.z); split_tf32(o.w, hi.w, lo.w);
            st4(y + j, hi);
            st4(y + a.y_lo_off + j, lo);
          } else {
            st4(y + j, o);
          }
        }
      }
      d = warp_sum(d);
      if (lane == 0) {
        const float gp = d + ev[u] * gs;
        g.GE[row[u]] = a.no_norm ? gp : pk[u] * (gs + (gp - cm));
      }
    }
  }
  if (a.no_norm) {     // exact cancellation of the chosen split (see cell_bwd_kernel)
    __syncthreads();
    if (warp < ncell) {
      const int64_t r0 = (int64_t)b * a.L * a.N + (int64_t)(p0 + warp) * a.sp;
      const float gs = s_sc[warp * 2];
      float cm2 = 0.f;
      for (int k = 0; k < a.N; ++k) cm2 = fmaf(a.Pr[r0 + (int64_t)k * a.sk], g.GE[r0 + (int64_t)k * a.sk], cm2);
      __syncwarp();
      for (int k = lane; k < a.N; k += 32) {
        const int64_t row = r0 + (int64_t)k * a.sk;
        g.GE[row] = a.Pr[row] * (gs + (g.GE[row] - cm2));
      }
    }
  }
}

}  // namespace cliora
