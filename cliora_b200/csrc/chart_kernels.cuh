// Chart kernels: everything of the inside/outside recursion that is not a dense GEMM.
// Each kernel implements one phase of oracle/factored.py (the CPU emulator the maths was
// validated with); names match the phase comments there.
#pragma once
#include "common.cuh"

namespace cliora {

// ------------------------------------------------------------------------------------------
// pack_weights: Wcat_in [PI*D, D] = [W1[:, :D]; W1[:, D:]; Wb; (oW1[:, :D])],  Wcat_out [2D, D] = [oW1[:, D:]; oWb]
// ------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(int D, int PI, const float* __restrict__ W1, const float* __restrict__ Wb,
                                    const float* __restrict__ oW1, const float* __restrict__ oWb,
                                    float* __restrict__ Wcat_in, float* __restrict__ Wcat_out) {
  pdl_prologue();
  const int64_t total = (int64_t)(PI + 2) * D * D;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int blk = (int)(idx / ((int64_t)D * D));
    const int rem = (int)(idx % ((int64_t)D * D));
    const int i = rem / D, j = rem % D;
    float v;
    if (blk == 0) v = W1[(int64_t)i * 2 * D + j];
    else if (blk == 1) v = W1[(int64_t)i * 2 * D + D + j];
    else if (blk == 2) v = Wb[(int64_t)i * D + j];
    else if (blk == 3 && PI == 4) v = oW1[(int64_t)i * 2 * D + j];
    else if (blk == PI) v = oW1[(int64_t)i * 2 * D + D + j];
    else v = oWb[(int64_t)i * D + j];
    if (blk < PI) Wcat_in[idx] = v;
    else Wcat_out[idx - (int64_t)PI * D * D] = v;
  }
}

// Projection buffers start from their bias row instead of zero: P[row, :] = 0 except P[row, boff : boff + D] = bias.
// Folding b1 into the second operand's projection (Ar = h W1r^T + b1) makes the hidden activation of a split
// z = relu(Al[first] + Ar[second]) -- one add less per element in the split kernels, and rows that are zero-filled
// (tile padding) come out as exact zeros without a select.
__global__ void init_proj_kernel(float* __restrict__ P, int64_t rows, int ld, int boff, int D,
                                 const float* __restrict__ bias) {
  pdl_prologue();
  const int64_t n4 = rows * (ld / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % (ld / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j >= boff && j < boff + D) v = ld4(bias + (j - boff));
    st4(P + i * 4, v);
  }
}

// ------------------------------------------------------------------------------------------
// split_build: per split row, z = relu(Al[first] + Ar[second]) (b1 is folded into Ar, see init_proj_kernel), e = h[first].V[second] + s[first] + s[second]
// inside : first = left child (inside chart), second = right child (inside chart)
// outside: first = sibling (inside chart),    second = parent (outside chart)
// One warp per row; rows ordered (b,p,k) inside, (b,k,p) outside (the reference's order).
// ------------------------------------------------------------------------------------------
struct SplitArgs {
  int B, n, level, L, N, D;
  int64_t C;
  const float* ih;    // inside_h [B,C,D]
  const float* is_;   // inside_s [B,C]
  const float* os_;   // outside_s [B,C] (outside only)
  const float* Pin;   // [B,C,ldPin]
  const float* Pout;  // [B,C,2D]
  int ldPin;
  int iAl;            // column offset of the "first argument" projection in Pin (0, or 3D when !share && outside)
  const float* b1;
  float* Z;           // level block [B*L*N, D] (hi part of the split pair when z_lo_off != 0)
  float* E;           // level block [B*L*N]
  int64_t z_lo_off;   // floats from Z to the lo part of the pair (0: store plain fp32)
  uint32_t* zmask;    // level block [B*L*N, 16]: ReLU bits, word t*4+comp bit l <-> column 128t + 4l + comp; or null
};

template <bool OUTSIDE>
CL_D void decode_row(const SplitArgs& a, int64_t m, int& b, int& first, int& second) {
  const int per = a.L * a.N;
  b = (int)(m / per);
  const int rem = (int)(m % per);
  if (!OUTSIDE) {
    const int p = rem / a.N, k = rem % a.N;
    inside_children(a.n, a.level, p, k, first, second);
  } else {
    const int k = rem / a.L, p = rem % a.L;
    outside_parent_sibling(a.n, a.level, p, k, second, first);
  }
}

template <bool OUTSIDE>
__global__ __launch_bounds__(256) void split_build_kernel(const SplitArgs a) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t rows = (int64_t)a.B * a.L * a.N;
  if (m >= rows) return;
  int b, c1, c2;
  decode_row<OUTSIDE>(a, m, b, c1, c2);
  const int64_t g1 = (int64_t)b * a.C + c1, g2 = (int64_t)b * a.C + c2;
  const float* Al = a.Pin + g1 * a.ldPin + a.iAl;
  const float* Ar = OUTSIDE ? a.Pout + g2 * 2 * a.D : a.Pin + g2 * a.ldPin + a.D;
  const float* V = OUTSIDE ? a.Pout + g2 * 2 * a.D + a.D : a.Pin + g2 * a.ldPin + 2 * a.D;
  const float* h1 = a.ih + g1 * a.D;
  float* z = a.Z + m * a.D;
  float dot = 0.f;
  // warp-uniform trip counts (the ballots need every lane); four 128-column chunks per round so that all of a
  // row's gathers (up to 16 independent 16-byte loads per lane) are in flight before the first use
  for (int t0 = 0; t0 * 128 < a.D; t0 += 4) {
    float4 x[4], y[4], hv[4], vv[4], bb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = lane * 4 + (t0 + u) * 128;
      if (j < a.D) {
        x[u] = ld4(Al + j); y[u] = ld4(Ar + j); bb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        hv[u] = ld4(h1 + j); vv[u] = ld4(V + j);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u;
      if (t * 128 >= a.D) break;                   // uniform
      const int j = lane * 4 + t * 128;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < a.D) {
        o.x = fmaxf(x[u].x + y[u].x + bb[u].x, 0.f);
        o.y = fmaxf(x[u].y + y[u].y + bb[u].y, 0.f);
        o.z = fmaxf(x[u].z + y[u].z + bb[u].z, 0.f);
        o.w = fmaxf(x[u].w + y[u].w + bb[u].w, 0.f);
        if (a.z_lo_off != 0) {
          float4 hi, lo;
          split_tf32(o.x, hi.x, lo.x); split_tf32(o.y, hi.y, lo.y); split_tf32(o.z, hi.z, lo.z); split_tf32(o.w, hi.w, lo.w);
          st4(z + j, hi);
          st4(z + a.z_lo_off + j, lo);
        } else {
          st4(z + j, o);
        }
        dot = fmaf(hv[u].x, vv[u].x, dot);
        dot = fmaf(hv[u].y, vv[u].y, dot);
        dot = fmaf(hv[u].z, vv[u].z, dot);
        dot = fmaf(hv[u].w, vv[u].w, dot);
      }
      if (a.zmask != nullptr) {   // ReLU bits for the backward GEMM epilogue (invalid lanes contribute zeros)
        const unsigned b0 = __ballot_sync(0xffffffffu, o.x > 0.f), b1 = __ballot_sync(0xffffffffu, o.y > 0.f);
        const unsigned b2 = __ballot_sync(0xffffffffu, o.z > 0.f), b3 = __ballot_sync(0xffffffffu, o.w > 0.f);
        if (lane == 0 && t < 4) *reinterpret_cast<uint4*>(a.zmask + m * 16 + t * 4) = make_uint4(b0, b1, b2, b3);
      }
    }
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    const float s2 = OUTSIDE ? a.os_[g2] : a.is_[g2];
    a.E[m] = dot + a.is_[g1] + s2;
  }
}

// ------------------------------------------------------------------------------------------
// cell_aggregate (+ cell_finalize): one block per chart cell of a level.
//   p = softmax_k(e_k);  a = sum_k p_k y_k;  sbar = sum_k p_k e_k;  q = a / max(|a|, eps)
//   VL: att = softmax_r(q . obj_r); patt = att * keep / 0.9; a2 = q + sum_r patt_r obj_r; h = a2 / max(|a2|, eps)
// Leaf mode (E == nullptr): N == 1, p = 1, sbar = 0, y = tanh(W_leaf x + b).
// Dynamic shared memory: (D + N + 2R + 64) floats.
// ------------------------------------------------------------------------------------------
struct CellArgs {
  int B, n, level, L, N, D, R;
  int64_t C;
  int sp, sk;          // row(b,p,k) = b*L*N + p*sp + k*sk   (inside: sp=N, sk=1; outside: sp=1, sk=L)
  float* Y;            // level block [B*L*N, D]   (bwd: overwritten with its gradient)
  int64_t y_lo_off;    // bwd: floats from Y to the lo part of the gradient pair (0: plain fp32)
  const float* E;      // level block [B*L*N] or nullptr (leaf)
  float* Pr;           // level block [B*L*N] softmax probabilities (fwd out, bwd in)
  float* chart_h;      // [B,C,D]
  float* chart_s;      // [B,C]
  float* q;            // [B,C,D] or nullptr (then q == chart_h)
  float* nrm;          // [B,C]
  float* nrm2;         // [B,C]   (R > 0)
  float* att;          // [B,C,R] (R > 0)
  const float* obj;    // [B,R,D] (R > 0)
  const uint8_t* keep; // [B,C,R] or nullptr
  int no_norm;         // --normalize none: norms are the identity; the saved norm is the sentinel -1
};

template <bool VL>
__global__ __launch_bounds__(128) void cell_aggregate_kernel(const CellArgs a) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  float* s_a = sm;                 // [D]
  float* s_p = s_a + a.D;          // [N]
  float* s_att = s_p + a.N;        // [R]
  float* s_patt = s_att + a.R;     // [R]
  float* s_red = s_patt + a.R;     // [64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int b = blockIdx.x / a.L, p = blockIdx.x % a.L;
  const int64_t cell = (int64_t)b * a.C + lvl_off(a.n, a.level) + p;
  const int64_t row0 = (int64_t)b * a.L * a.N + (int64_t)p * a.sp;

  // 1. softmax over the N splits (warp 0)
  if (warp == 0) {
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
  }
  __syncthreads();

  // 2. a = sum_k p_k y_k
  float ss = 0.f;
  for (int j = tid * 4; j < a.D; j += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int k = 0; k < a.N; ++k) {
      const float pk = s_p[k];
      const float4 y = ld4(a.Y + (row0 + (int64_t)k * a.sk) * a.D + j);
      acc.x = fmaf(pk, y.x, acc.x);
      acc.y = fmaf(pk, y.y, acc.y);
      acc.z = fmaf(pk, y.z, acc.z);
      acc.w = fmaf(pk, y.w, acc.w);
    }
    st4(s_a + j, acc);
    ss += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
  }
  ss = block_sum(ss, s_red);
  const float nrm = a.no_norm ? 1.f : fmaxf(sqrtf(ss), kTiny);
  const float inv_nrm = 1.f / nrm;
  // q = a / nrm
  for (int j = tid * 4; j < a.D; j += blockDim.x * 4) {
    float4 v = ld4(s_a + j);
    v.x *= inv_nrm; v.y *= inv_nrm; v.z *= inv_nrm; v.w *= inv_nrm;
    st4(s_a + j, v);
    if (VL) st4(a.q + cell * a.D + j, v);
    else st4(a.chart_h + cell * a.D + j, v);
  }
  if (tid == 0) a.nrm[cell] = a.no_norm ? -1.f : nrm;
  if (!VL) return;

  __syncthreads();
  // 3. region attention of this cell against its own image's R regions
  const float* obj = a.obj + (int64_t)b * a.R * a.D;
  for (int r = warp; r < a.R; r += nwarps) {
    float d = 0.f;
    for (int j = lane * 4; j < a.D; j += 128) {
      const float4 qv = ld4(s_a + j), ov = ld4(obj + (int64_t)r * a.D + j);
      d = fmaf(qv.x, ov.x, d); d = fmaf(qv.y, ov.y, d); d = fmaf(qv.z, ov.z, d); d = fmaf(qv.w, ov.w, d);
    }
    d = warp_sum(d);
    if (lane == 0) s_att[r] = d;
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int r = lane; r < a.R; r += 32) mx = fmaxf(mx, s_att[r]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int r = lane; r < a.R; r += 32) {
      const float ex = expf(s_att[r] - mx);
      s_att[r] = ex;
      sum += ex;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int r = lane; r < a.R; r += 32) {
      const float pr = s_att[r] * inv;
      s_att[r] = pr;
      a.att[cell * a.R + r] = pr;
      float sc = 1.f;
      if (a.keep != nullptr) sc = a.keep[cell * a.R + r] ? kKeepScale : 0.f;
      s_patt[r] = pr * sc;
    }
  }
  __syncthreads();
  float ss2 = 0.f;
  for (int j = tid * 4; j < a.D; j += blockDim.x * 4) {
    float4 acc = ld4(s_a + j);
    for (int r = 0; r < a.R; ++r) {
      const float w = s_patt[r];
      const float4 ov = ld4(obj + (int64_t)r * a.D + j);
      acc.x = fmaf(w, ov.x, acc.x); acc.y = fmaf(w, ov.y, acc.y);
      acc.z = fmaf(w, ov.z, acc.z); acc.w = fmaf(w, ov.w, acc.w);
    }
    st4(s_a + j, acc);
    ss2 += acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
  }
  ss2 = block_sum(ss2, s_red);
  const float nrm2 = a.no_norm ? 1.f : fmaxf(sqrtf(ss2), kTiny);
  const float inv2 = 1.f / nrm2;
  for (int j = tid * 4; j < a.D; j += blockDim.x * 4) {
    float4 v = ld4(s_a + j);
    v.x *= inv2; v.y *= inv2; v.z *= inv2; v.w *= inv2;
    st4(a.chart_h + cell * a.D + j, v);
  }
  if (tid == 0) a.nrm2[cell] = a.no_norm ? -1.f : nrm2;
}

// outside root: outside_h[:, root] = unit(root_vector), outside_s[:, root] = 0   (diora.py:337-356)
__global__ void outside_root_kernel(int B, int D, int64_t C, const float* __restrict__ root,
                                    float* __restrict__ oh, float* __restrict__ os_, float* __restrict__ nrm_out,
                                    int no_norm) {
  pdl_prologue();
  __shared__ float red[64];
  float ss = 0.f;
  for (int j = threadIdx.x; j < D; j += blockDim.x) ss += root[j] * root[j];
  ss = block_sum(ss, red);
  const float nrm = no_norm ? 1.f : fmaxf(sqrtf(ss), kTiny);
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < D; j += blockDim.x) oh[((int64_t)b * C + C - 1) * D + j] = root[j] / nrm;
  if (threadIdx.x == 0) {
    os_[(int64_t)b * C + C - 1] = 0.f;
    nrm_out[(int64_t)b * C + C - 1] = no_norm ? -1.f : nrm;
  }
}

// ==========================================================================================
// backward
// ==========================================================================================

// d/da of a / clamp(|a|, eps): live branch (g - h (h.g)) / nrm, clamped branch g / eps.  With --normalize none the saved
// norm is the sentinel -1: the coefficient is 0 and 1 / |nrm| = 1, i.e. the gradient passes through unchanged.
CL_D float unit_bwd_coef(float nrm, float hdotg) { return (nrm > kTiny) ? hdotg : 0.f; }

struct CellBwdArgs {
  CellArgs c;            // same geometry; Y is overwritten with GY, Pr is read
  const float* Gh;       // [B,C,D] total gradient wrt the cell vector
  const float* Gs;       // [B,C]   total gradient wrt the cell score
  float* GE;             // level block [B*L*N] out: gradient wrt split scores
  float* GA2;            // [B,C,D] out (VL): gradient wrt the pre-second-normalisation vector
  float* coef;           // [B,C,2R] out (VL): patt_r, g_logit_r
  const float* leaf_t;   // leaf mode: tanh outputs [B*n, D]
  float* gu;             // leaf mode out: gradient wrt the pre-tanh activations [B*n, D]
  float* cellsum;        // [B,C,D] out or null: sum over the cell's splits of the GY rows (colsum of these = db2)
  // cells-only mode (the fused level backward kernel does the per-split part): when ga_out != null the kernel stores
  // the gradient wrt the cell's pre-normalisation sum and the softmax constant sum_m p_m gp_m, and stops there
  float* ga_out;         // [B,C,D]
  float* cm_out;         // [B,C]
};

// One block per cell.  Dynamic shared memory: (2D + 3R + 64) floats, + 8D when cellsum != null.
template <bool VL>
__global__ __launch_bounds__(256) void cell_bwd_kernel(const CellBwdArgs g) {
  pdl_prologue();
  const CellArgs& a = g.c;
  extern __shared__ __align__(16) float sm[];
  float* s_g = sm;                  // [D] working gradient
  float* s_q = s_g + a.D;           // [D] q (VL)
  float* s_r0 = s_q + a.D;          // [R] g_att -> g_logit
  float* s_r1 = s_r0 + a.R;         // [R] att
  float* s_r2 = s_r1 + a.R;         // [R] scale
  float* s_red = s_r2 + a.R;        // [64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int b = blockIdx.x / a.L, p = blockIdx.x % a.L;
  const int64_t cell = (int64_t)b * a.C + lvl_off(a.n, a.level) + p;
  const int64_t row0 = (int64_t)b * a.L * a.N + (int64_t)p * a.sp;
  const float* hvec = a.chart_h + cell * a.D;
  const float* qvec = VL ? a.q + cell * a.D : hvec;
  const float nrm_raw = a.nrm[cell];
  const float nrm = fabsf(nrm_raw);

  // ---- gradient wrt q (the first-normalised vector) ----
  float hd = 0.f;
  for (int j = tid; j < a.D; j += blockDim.x) {
    const float gv = g.Gh[cell * a.D + j];
    s_g[j] = gv;
    hd = fmaf(hvec[j], gv, hd);
    if (VL) s_q[j] = qvec[j];
  }
  hd = block_sum(hd, s_red);
  if (VL) {
    const float nrm2_raw = a.nrm2[cell];
    const float coef = unit_bwd_coef(nrm2_raw, hd);
    const float inv2 = 1.f / fabsf(nrm2_raw);
    for (int j = tid; j < a.D; j += blockDim.x) {
      const float v = (s_g[j] - hvec[j] * coef) * inv2;   // ga2
      s_g[j] = v;
      g.GA2[cell * a.D + j] = v;
    }
    __syncthreads();
    const float* obj = a.obj + (int64_t)b * a.R * a.D;
    for (int r = warp; r < a.R; r += nwarps) {
      float d = 0.f;
      for (int j = lane * 4; j < a.D; j += 128) {     // 16-byte loads (D % 4 == 0)
        const float4 gv = ld4(s_g + j), ov = ld4(obj + (int64_t)r * a.D + j);
        d = fmaf(gv.x, ov.x, d); d = fmaf(gv.y, ov.y, d); d = fmaf(gv.z, ov.z, d); d = fmaf(gv.w, ov.w, d);
      }
      d = warp_sum(d);
      if (lane == 0) {
        float sc = 1.f;
        if (a.keep != nullptr) sc = a.keep[cell * a.R + r] ? kKeepScale : 0.f;
        const float at = a.att[cell * a.R + r];
        s_r0[r] = d * sc;   // g_att
        s_r1[r] = at;
        s_r2[r] = sc;
      }
    }
    __syncthreads();
    float part = 0.f;
    for (int r = tid; r < a.R; r += blockDim.x) part = fmaf(s_r1[r], s_r0[r], part);
    const float dsum = block_sum(part, s_red);
    for (int r = tid; r < a.R; r += blockDim.x) {
      const float gl = s_r1[r] * (s_r0[r] - dsum);   // g_logit
      s_r0[r] = gl;
      g.coef[(cell * 2) * a.R + r] = s_r1[r] * s_r2[r];   // patt
      g.coef[(cell * 2 + 1) * a.R + r] = gl;
    }
    __syncthreads();
    // gq = ga2 + sum_r g_logit_r obj_r
    float qd = 0.f;
    for (int j = tid * 4; j < a.D; j += blockDim.x * 4) {
      float4 v = ld4(s_g + j);
#pragma unroll 6
      for (int r = 0; r < a.R; ++r) {
        const float w = s_r0[r];
        const float4 ov = ld4(obj + (int64_t)r * a.D + j);
        v.x = fmaf(w, ov.x, v.x); v.y = fmaf(w, ov.y, v.y); v.z = fmaf(w, ov.z, v.z); v.w = fmaf(w, ov.w, v.w);
      }
      st4(s_g + j, v);
      const float4 qv = ld4(s_q + j);
      qd = fmaf(qv.x, v.x, qd); qd = fmaf(qv.y, v.y, qd); qd = fmaf(qv.z, v.z, qd); qd = fmaf(qv.w, v.w, qd);
    }
    hd = block_sum(qd, s_red);
  }
  // ---- ga = unit_bwd(gq, q, nrm) ----
  {
    const float coef = unit_bwd_coef(nrm_raw, hd);
    const float inv = 1.f / nrm;
    float ad = 0.f;
    for (int j = tid; j < a.D; j += blockDim.x) {
      const float qv = VL ? s_q[j] : hvec[j];
      const float v = (s_g[j] - qv * coef) * inv;
      s_g[j] = v;
      ad = fmaf(qv, v, ad);
    }
    ad = block_sum(ad, s_red);   // q . ga   (also makes s_g visible)
    if (a.E == nullptr) {
      // leaf: gu = ga * (1 - t^2)
      const int64_t row = (int64_t)b * a.L + p;
      for (int j = tid; j < a.D; j += blockDim.x) {
        const float t = g.leaf_t[row * a.D + j];
        g.gu[row * a.D + j] = s_g[j] * (1.f - t * t);
      }
      return;
    }
    const float gs = g.Gs[cell];
    const float cm = nrm * ad + a.chart_s[cell] * gs;   // sum_m p_m gp_m
    if (g.ga_out != nullptr) {     // cells-only mode
      for (int j = tid * 4; j < a.D; j += blockDim.x * 4) st4(g.ga_out + cell * a.D + j, ld4(s_g + j));
      if (tid == 0) g.cm_out[cell] = cm;
      return;
    }
    // ---- per split: ge, gy ----
    const bool want_sum = g.cellsum != nullptr;     // only offered for D <= 512 (one 4-chunk round per row)
    float4 bacc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) bacc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = warp; k < a.N; k += nwarps) {
      const int64_t row = row0 + (int64_t)k * a.sk;
      float* y = a.Y + row * a.D;
      const float pk = a.Pr[row];
      float d = 0.f;
      for (int j0 = lane * 4; j0 < a.D; j0 += 512) {   // the row's loads are issued together
        float4 yq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u * 128 < a.D) yq[u] = ld4(y + j0 + u * 128);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u * 128;
          if (j >= a.D) break;
          const float4 yv = yq[u], gv = ld4(s_g + j);
          d = fmaf(yv.x, gv.x, d); d = fmaf(yv.y, gv.y, d); d = fmaf(yv.z, gv.z, d); d = fmaf(yv.w, gv.w, d);
          float4 o;
          o.x = yv.x > 0.f ? pk * gv.x : 0.f;
          o.y = yv.y > 0.f ? pk * gv.y : 0.f;
          o.z = yv.z > 0.f ? pk * gv.z : 0.f;
          o.w = yv.w > 0.f ? pk * gv.w : 0.f;
          if (want_sum) { bacc[u].x += o.x; bacc[u].y += o.y; bacc[u].z += o.z; bacc[u].w += o.w; }
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
        const float gp = d + a.E[row] * gs;
        g.GE[row] = a.no_norm ? gp : pk * (gs + (gp - cm));
      }
    }
    if (a.no_norm) {
      // Without normalisation the chart grows geometrically and the softmax saturates: cm must cancel gp of the chosen
      // split exactly, so it is re-summed from the very gp values it is subtracted from (as autograd does).
      __syncthreads();
      float cm2 = 0.f;
      for (int k = 0; k < a.N; ++k) {
        const int64_t row = row0 + (int64_t)k * a.sk;
        cm2 = fmaf(a.Pr[row], g.GE[row], cm2);
      }
      __syncthreads();
      for (int k = tid; k < a.N; k += blockDim.x) {
        const int64_t row = row0 + (int64_t)k * a.sk;
        g.GE[row] = a.Pr[row] * (gs + (g.GE[row] - cm2));
      }
    }
    if (want_sum) {   // warps -> block: the cell's sum of GY rows
      float* s_part = sm + ((2 * a.D + 3 * a.R + 64 + 3) & ~3);   // [nwarps][D], 16-byte aligned
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = lane * 4 + u * 128;
        if (j < a.D) st4(s_part + warp * a.D + j, bacc[u]);
      }
      __syncthreads();
      for (int j = tid * 4; j < a.D; j += blockDim.x * 4) {
        float4 t = ld4(s_part + j);
        for (int w = 1; w < nwarps; ++w) {
          const float4 x = ld4(s_part + w * a.D + j);
          t.x += x.x; t.y += x.y; t.z += x.z; t.w += x.w;
        }
        st4(g.cellsum + cell * a.D + j, t);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// split_scatter: push the per-split gradients into the accumulators of the two cells a split read.
// ------------------------------------------------------------------------------------------
struct ScatterArgs {
  SplitArgs s;           // geometry + forward tensors (ih, Pin, Pout)
  const float* GZ;       // level-local [rows, D]
  const float* GE;       // level block [rows]
  float* Gh_in;          // [B,C,D]
  float* Gs_in;          // [B,C]
  float* GP_in;          // [B,C,ldPin]
  float* Gs_out;         // [B,C]
  float* GP_out;         // [B,C,2D]
};

template <bool OUTSIDE>
__global__ __launch_bounds__(256) void split_scatter_kernel(const ScatterArgs g) {
  pdl_prologue();
  const SplitArgs& a = g.s;
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t rows = (int64_t)a.B * a.L * a.N;
  if (m >= rows) return;
  int b, c1, c2;
  decode_row<OUTSIDE>(a, m, b, c1, c2);
  const int64_t g1 = (int64_t)b * a.C + c1, g2 = (int64_t)b * a.C + c2;
  const float ge = g.GE[m];
  const float* gz = g.GZ + m * a.D;
  const float* h1 = a.ih + g1 * a.D;
  const float* V = OUTSIDE ? a.Pout + g2 * 2 * a.D + a.D : a.Pin + g2 * a.ldPin + 2 * a.D;
  float* dAl = g.GP_in + g1 * a.ldPin + a.iAl;
  float* dAr = OUTSIDE ? g.GP_out + g2 * 2 * a.D : g.GP_in + g2 * a.ldPin + a.D;
  float* dV = OUTSIDE ? g.GP_out + g2 * 2 * a.D + a.D : g.GP_in + g2 * a.ldPin + 2 * a.D;
  float* dh1 = g.Gh_in + g1 * a.D;
  for (int j0 = lane * 4; j0 < a.D; j0 += 512) {   // four chunks per round: 12 independent loads before the reds
    float4 z[4], vv[4], hv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * 128;
      if (j < a.D) { z[u] = ld4(gz + j); vv[u] = ld4(V + j); hv[u] = ld4(h1 + j); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * 128;
      if (j < a.D) {
        red_add4(dAl + j, z[u]);
        red_add4(dAr + j, z[u]);
        red_add4(dh1 + j, make_float4(ge * vv[u].x, ge * vv[u].y, ge * vv[u].z, ge * vv[u].w));
        red_add4(dV + j, make_float4(ge * hv[u].x, ge * hv[u].y, ge * hv[u].z, ge * hv[u].w));
      }
    }
  }
  if (lane == 0) {
    atomicAdd(g.Gs_in + g1, ge);
    atomicAdd((OUTSIDE ? g.Gs_out : g.Gs_in) + g2, ge);
  }
}

// g_root += unit_bwd(Gh_out[b, root], h_root, nrm_root)   one block per sentence b; g_root zero-filled by the caller
__global__ void outside_root_bwd_kernel(int B, int D, int64_t C, const float* __restrict__ Gh_out,
                                        const float* __restrict__ oh, const float* __restrict__ nrm_out,
                                        float* __restrict__ g_root) {
  pdl_prologue();
  __shared__ float red[64];
  const int b = blockIdx.x;
  const float nrm_raw = nrm_out[(int64_t)b * C + C - 1];
  const float nrm = fabsf(nrm_raw);
  const float* gh = Gh_out + ((int64_t)b * C + C - 1) * D;
  const float* h = oh + ((int64_t)b * C + C - 1) * D;
  float d = 0.f;
  for (int j = threadIdx.x; j < D; j += blockDim.x) d = fmaf(h[j], gh[j], d);
  d = block_sum(d, red);
  const float coef = unit_bwd_coef(nrm_raw, d);
  for (int j = threadIdx.x; j < D; j += blockDim.x) atomicAdd(g_root + j, (gh[j] - h[j] * coef) / nrm);
}

// g_obj[b,r,:] = sum_c patt[b,c,r] ga2[b,c,:] + g_logit[b,c,r] q[b,c,:]     grid (ceil(D/32), B), block (32, 4)
// threadIdx.x = column within a 32-wide slab, threadIdx.y = quarter of the region list; cells are staged
// 16 at a time (coefficients in smem, the two vectors prefetched in registers) so loads overlap.
template <int RMAX>
__global__ __launch_bounds__(128) void obj_grad_kernel(int D, int R, int64_t C, const float* __restrict__ GA2,
                                                       const float* __restrict__ q, const float* __restrict__ coef,
                                                       float* __restrict__ g_obj, int accumulate) {
  pdl_prologue();
  constexpr int RQ = RMAX / 4;         // upper bound of regions per threadIdx.y
  constexpr int CH = 16;               // cells per chunk
  const int rq = ((R + 15) / 16) * 4;  // regions per threadIdx.y, a multiple of 4 so coefficients are read 16 bytes at a time
  const int b = blockIdx.y;
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  __shared__ __align__(16) float s_c[CH][2 * RMAX];   // [cell][patt_r (r < RMAX) | g_logit_r], zero padded
  float acc[RQ];
#pragma unroll
  for (int i = 0; i < RQ; ++i) acc[i] = 0.f;
  const int r0 = threadIdx.y * rq;
  // cells are split over blockIdx.z (partials are red.add'ed into a zero-filled g_obj when gridDim.z > 1)
  const int64_t per = (C + gridDim.z - 1) / gridDim.z;
  const int64_t c_begin = blockIdx.z * per, c_end = min(C, c_begin + per);
  for (int64_t c0 = c_begin; c0 < c_end; c0 += CH) {
    const int nc = (int)min((int64_t)CH, c_end - c0);
    __syncthreads();
    for (int t = tid; t < CH * 2 * RMAX; t += 128) {
      const int cc = t / (2 * RMAX), k = t % (2 * RMAX);
      const int half = k / RMAX, r = k % RMAX;
      s_c[cc][k] = (cc < nc && r < R) ? coef[((int64_t)b * C + c0 + cc) * 2 * R + half * R + r] : 0.f;
    }
    float gv[CH], qv[CH];
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) {
      const bool ok = cc < nc && j < D;
      const int64_t cell = (int64_t)b * C + c0 + cc;
      gv[cc] = ok ? GA2[cell * D + j] : 0.f;
      qv[cc] = ok ? q[cell * D + j] : 0.f;
    }
    __syncthreads();
    if (r0 < R) {
#pragma unroll
      for (int cc = 0; cc < CH; ++cc) {
#pragma unroll
        for (int i = 0; i < RQ; i += 4) {
          if (i < rq) {
            const float4 pa = ld4(&s_c[cc][r0 + i]), gl = ld4(&s_c[cc][RMAX + r0 + i]);
            acc[i + 0] = fmaf(pa.x, gv[cc], fmaf(gl.x, qv[cc], acc[i + 0]));
            acc[i + 1] = fmaf(pa.y, gv[cc], fmaf(gl.y, qv[cc], acc[i + 1]));
            acc[i + 2] = fmaf(pa.z, gv[cc], fmaf(gl.z, qv[cc], acc[i + 2]));
            acc[i + 3] = fmaf(pa.w, gv[cc], fmaf(gl.w, qv[cc], acc[i + 3]));
          }
        }
      }
    }
  }
  if (j < D) {
#pragma unroll
    for (int i = 0; i < RQ; ++i) {
      const int r = r0 + i;
      if (i < rq && r < R) {
        float* dst = g_obj + ((int64_t)b * R + r) * D + j;
        if (gridDim.z > 1) atomicAdd(dst, acc[i]);
        else *dst = accumulate ? *dst + acc[i] : acc[i];
      }
    }
  }
}

// column sums: dst[j] (+)= sum_r src[r*ld + j].  Two deterministic stages through `part` [S, cols].
__global__ void colsum_stage1_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int cols,
                                     float* __restrict__ part) {
  pdl_prologue();
  __shared__ float s[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int64_t chunk = (rows + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * chunk, r1 = min(rows, r0 + chunk);
  float acc = 0.f;
  if (j < cols)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) acc += src[r * ld + j];
  s[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && j < cols) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += s[y][threadIdx.x];
    part[(int64_t)blockIdx.y * cols + j] = t;
  }
}
__global__ void colsum_stage2_kernel(const float* __restrict__ part, int S, int cols, float* __restrict__ dst,
                                     int accumulate) {
  pdl_prologue();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  float t = 0.f;
  for (int s = 0; s < S; ++s) t += part[(int64_t)s * cols + j];
  dst[j] = accumulate ? dst[j] + t : t;
}

}  // namespace cliora
