// Span-region alignment: scores = h . obj^T with a max-over-regions epilogue, its backward, and the
// contrastive / visual-grounding losses (cliora/net/cliora.py:457-466, cliora/net/trainer.py:81-171).
#pragma once
#include "common.cuh"

namespace cliora {

// ------------------------------------------------------------------------------------------
// atten_max: smax[a, c, cell] = max_r h[a,cell] . obj[c,r]   (argmax = first max)
// Block = 64 (a,cell) rows x one image c (R <= 64 region columns); never materialises [B,B,cells,R].
// grid (B images, ceil(B*ncell / 64)), 256 threads, 4x4 micro-tile like gemm_simt_kernel.
// ------------------------------------------------------------------------------------------
__global__ __launch_bounds__(256) void atten_max_kernel(int B, int ncell, int D, int R, const float* __restrict__ h,
                                                        int64_t h_bstride, const float* __restrict__ obj,
                                                        float* __restrict__ smax, int32_t* __restrict__ amax) {
  pdl_prologue();
  __shared__ __align__(16) float As[2][16][68];
  __shared__ __align__(16) float Bs[2][16][68];
  const int tid = threadIdx.x;
  const int c = blockIdx.x;
  const int m0 = blockIdx.y * 64;
  const int M = B * ncell;
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int am = m0 + lr;
  const bool a_ok = am < M;
  const float* a_ptr = a_ok ? h + ((int64_t)(am / ncell) * h_bstride + (am % ncell)) * D : h;
  const bool b_ok = lr < R;
  const float* b_ptr = b_ok ? obj + ((int64_t)c * R + lr) * D : obj;
  float4 ra, rb;
  auto load = [&](int k0) {
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = ra;
    const int k = k0 + lk;
    if (k < D) {   // D % 4 == 0
      if (a_ok) ra = ld4(a_ptr + k);
      if (b_ok) rb = ld4(b_ptr + k);
    }
  };
  auto store = [&](int buf) {
    As[buf][lk + 0][lr] = ra.x; As[buf][lk + 1][lr] = ra.y; As[buf][lk + 2][lr] = ra.z; As[buf][lk + 3][lr] = ra.w;
    Bs[buf][lk + 0][lr] = rb.x; Bs[buf][lk + 1][lr] = rb.y; Bs[buf][lk + 2][lr] = rb.z; Bs[buf][lk + 3][lr] = rb.w;
  };
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nk = (D + 15) / 16;
  load(0);
  store(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) load((kt + 1) * 16);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = ld4(&As[cur][kk][ty * 4]);
      const float4 b = ld4(&Bs[cur][kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
      }
    }
    if (kt + 1 < nk) store(cur ^ 1);
    __syncthreads();
  }
  // max over the R columns of each row: 4 in-thread, then across the 16 tx lanes of the row.
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = tx * 4 + j;
      if (r < R && acc[i][j] > best) { best = acc[i][j]; bi = r; }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    const int m = m0 + ty * 4 + i;
    if (tx == 0 && m < M) {
      const int a = m / ncell, cell = m % ncell;
      const int64_t o = ((int64_t)a * B + c) * ncell + cell;
      smax[o] = best;
      amax[o] = min(bi, R - 1);   // bi stays INT_MAX only if every score was NaN
    }
  }
}

// g_h[a,cell,:] += sum_c g[a,c,cell] obj[c, amax[a,c,cell], :]       one block (128 thr) per (a,cell)
__global__ __launch_bounds__(128) void atten_max_bwd_h_kernel(int B, int ncell, int D, int R,
                                                              const float* __restrict__ obj,
                                                              const float* __restrict__ g,
                                                              const int32_t* __restrict__ amax,
                                                              float* __restrict__ g_h, int64_t gh_bstride) {
  pdl_prologue();
  const int a = blockIdx.x / ncell, cell = blockIdx.x % ncell;
  float* dst = g_h + ((int64_t)a * gh_bstride + cell) * D;
  for (int j = threadIdx.x * 4; j < D; j += blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < B; ++c) {
      const int64_t o = ((int64_t)a * B + c) * ncell + cell;
      const float gv = g[o];
      if (gv != 0.f) {
        const float4 ov = ld4(obj + ((int64_t)c * R + amax[o]) * D + j);
        acc.x = fmaf(gv, ov.x, acc.x); acc.y = fmaf(gv, ov.y, acc.y);
        acc.z = fmaf(gv, ov.z, acc.z); acc.w = fmaf(gv, ov.w, acc.w);
      }
    }
    float4 old = ld4(dst + j);
    old.x += acc.x; old.y += acc.y; old.z += acc.z; old.w += acc.w;
    st4(dst + j, old);
  }
}

// g_obj[c,r,:] += sum_{a,cell : amax == r} g[a,c,cell] h[a,cell,:]    one block (4 warps) per (c,r).
// The (a,cell) entries of image c are scanned 128 at a time; matches are compacted in index order with
// warp ballots (deterministic), then warp w accumulates matches w, w+4, ... four rows at a time (one dominant
// region can own most cells of an image, so the row loads must overlap), partials summed in fixed order.
// Dynamic smem: 4 * D floats.
__global__ __launch_bounds__(128) void atten_max_bwd_obj_kernel(int B, int ncell, int D, int R,
                                                                const float* __restrict__ h, int64_t h_bstride,
                                                                const float* __restrict__ g,
                                                                const int32_t* __restrict__ amax,
                                                                float* __restrict__ g_obj) {
  pdl_prologue();
  extern __shared__ __align__(16) float s_part[];   // [4][D]
  __shared__ int s_row[128];
  __shared__ float s_gv[128];
  __shared__ int s_wcount[4];
  const int c = blockIdx.x / R, r = blockIdx.x % R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int E = B * ncell;
  // blockIdx.y splits the (a,cell) scan: a dominant region can own most cells of an image, so one block per
  // (c,r) would walk thousands of rows serially; partial sums are red.add'ed into g_obj (caller zero-fills)
  const int e_chunk = ((E + gridDim.y - 1) / gridDim.y + 127) / 128 * 128;
  const int e_lo = blockIdx.y * e_chunk, e_hi = min(E, e_lo + e_chunk);
  // lane owns columns j = lane*4 + t*128, t < 4 (D <= 512), looping beyond for larger D
  for (int jbase = 0; jbase < D; jbase += 512) {
    float4 acc[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e0 = e_lo; e0 < e_hi; e0 += 128) {
      const int e = e0 + tid;
      bool hit = false;
      float gv = 0.f;
      int a = 0, cell = 0;
      if (e < e_hi) {
        a = e / ncell;
        cell = e % ncell;
        const int64_t o = ((int64_t)a * B + c) * ncell + cell;
        if (amax[o] == r) {
          gv = g[o];
          hit = gv != 0.f;
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      __syncthreads();   // previous chunk's list fully consumed
      if (lane == 0) s_wcount[warp] = __popc(bal);
      __syncthreads();
      int base = 0;
      for (int w = 0; w < warp; ++w) base += s_wcount[w];
      const int total = s_wcount[0] + s_wcount[1] + s_wcount[2] + s_wcount[3];
      if (hit) {
        const int pos = base + __popc(bal & ((1u << lane) - 1u));
        s_row[pos] = (int)((int64_t)a * h_bstride + cell);
        s_gv[pos] = gv;
      }
      __syncthreads();
      for (int i = warp; i < total; i += 16) {
        float w[4];
        const float* hr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ii = i + 4 * u;
          w[u] = ii < total ? s_gv[ii] : 0.f;
          hr[u] = h + (int64_t)s_row[ii < total ? ii : i] * D + jbase;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int j = lane * 4 + t * 128;
          if (jbase + j < D) {
            float4 hv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) hv[u] = ld4(hr[u] + j);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              acc[t].x = fmaf(w[u], hv[u].x, acc[t].x); acc[t].y = fmaf(w[u], hv[u].y, acc[t].y);
              acc[t].z = fmaf(w[u], hv[u].z, acc[t].z); acc[t].w = fmaf(w[u], hv[u].w, acc[t].w);
            }
          }
        }
      }
    }
    const int Dc = min(512, D - jbase);
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = lane * 4 + t * 128;
      if (j < Dc) st4(s_part + warp * 512 + j, acc[t]);
    }
    __syncthreads();
    float* dst = g_obj + ((int64_t)c * R + r) * D + jbase;
    for (int j = tid * 4; j < Dc; j += 512) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int w2 = 0; w2 < 4; ++w2) {
        const float4 pv = ld4(s_part + w2 * 512 + j);
        o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
      }
      if (o.x != 0.f || o.y != 0.f || o.z != 0.f || o.w != 0.f) red_add4(dst + j, o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// ContrastiveLoss (trainer.py:91-128).  One block per cell (< ncell); smax is [B, B, ncell].
//   txt[b] = 1/B sum_{c != b} max(m + S[b,c] - S[b,b], 1e-8)
//   img[b] = 1/B sum_{a != b} max(m + S[a,b] - S[b,b], 1e-8)
//   w[b]   = exp(is[b,cell] + os[b,cell] - is[b,root]);   partial[cell] = sum_b w[b] (txt[b] + img[b])
// Gradients (when g_smax != nullptr) are written for this cell's slice; g_is/g_os get g_w * w and the
// root term is accumulated per cell into g_root_part[cell, b] (summed by contrastive_finish_kernel).
// Dynamic smem: 3B + 64 floats.
// ------------------------------------------------------------------------------------------
__global__ __launch_bounds__(128) void contrastive_cell_kernel(int B, int64_t C, int ncell, const float* __restrict__ S,
                                                               const float* __restrict__ is_,
                                                               const float* __restrict__ os_, float margin,
                                                               float scale /* alpha / B */, float* __restrict__ partial,
                                                               float* __restrict__ g_S, float* __restrict__ g_is,
                                                               float* __restrict__ g_os,
                                                               float* __restrict__ g_root_part) {
  pdl_prologue();
  extern __shared__ float sm[];
  float* s_diag = sm;          // [B]
  float* s_gvl = sm + B;       // [B]  d loss / d vl[b]  (already divided by B for the mean over negatives)
  float* s_red = sm + 2 * B;   // [64]
  const int cell = blockIdx.x;
  const float invB = 1.f / (float)B;
  for (int b = threadIdx.x; b < B; b += blockDim.x) s_diag[b] = S[((int64_t)b * B + b) * ncell + cell];
  __syncthreads();
  float local = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float db = s_diag[b];
    float txt = 0.f, img = 0.f;
    for (int c = 0; c < B; ++c) {
      if (c == b) continue;
      txt += fmaxf(margin + S[((int64_t)b * B + c) * ncell + cell] - db, kTiny);
      img += fmaxf(margin + S[((int64_t)c * B + b) * ncell + cell] - db, kTiny);
    }
    const float vl = (txt + img) * invB;
    const int64_t ci = (int64_t)b * C + cell;
    const float w = expf(is_[ci] + os_[ci] - is_[(int64_t)b * C + C - 1]);
    local += w * vl;
    s_gvl[b] = scale * w * invB;
    if (g_S != nullptr) {
      const float gw = scale * vl * w;   // d loss / d (is + os - is_root)
      g_is[ci] = gw;
      g_os[ci] = gw;
      g_root_part[(int64_t)cell * B + b] = gw;
    }
  }
  local = block_sum(local, s_red);
  if (threadIdx.x == 0) partial[cell] = local;
  if (g_S == nullptr) return;
  __syncthreads();
  // g_S[a,c] = act_txt(a,c) gvl[a] + act_img(a,c) gvl[c]  (a != c);  g_S[a,a] = -(sum of its row txt + its column img)
  for (int a = threadIdx.x; a < B; a += blockDim.x) {
    const float da = s_diag[a];
    float gd = 0.f;
    for (int c = 0; c < B; ++c) {
      if (c == a) continue;
      const int64_t o = ((int64_t)a * B + c) * ncell + cell;
      const float sac = S[o];
      float gv = 0.f;
      if (margin + sac - da > kTiny) { gv += s_gvl[a]; gd -= s_gvl[a]; }
      if (margin + sac - s_diag[c] > kTiny) gv += s_gvl[c];
      g_S[o] = gv;
      // column term of the diagonal: img(a', a) for a' = c uses S[c, a] - S[a, a]
      const float sca = S[((int64_t)c * B + a) * ncell + cell];
      if (margin + sca - da > kTiny) gd -= s_gvl[a];
    }
    g_S[((int64_t)a * B + a) * ncell + cell] = gd;
  }
}

// loss = scale-free sum of partial[cell]; g_is[b, root] -= sum_cell g_root_part[cell, b]
__global__ void contrastive_finish_kernel(int B, int64_t C, int ncell, const float* __restrict__ partial, float scale,
                                          float* __restrict__ loss, const float* __restrict__ g_root_part,
                                          float* __restrict__ g_is) {
  pdl_prologue();
  __shared__ float red[64];
  float t = 0.f;
  for (int i = threadIdx.x; i < ncell; i += blockDim.x) t += partial[i];
  t = block_sum(t, red);
  if (threadIdx.x == 0) *loss = t * scale;
  if (g_root_part == nullptr) return;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float s = 0.f;
    for (int cell = 0; cell < ncell; ++cell) s += g_root_part[(int64_t)cell * B + b];
    // the root cell itself may be < ncell only when n == 1 (ncell == 0 then), so this never aliases a written slot
    g_is[(int64_t)b * C + C - 1] -= s;
  }
}

// ------------------------------------------------------------------------------------------
// VGLoss (trainer.py:139-171): logits[a,c] = mean_w wmax[a,c,w]; loss = alpha * mean_a CE(logits[a,:], a)
// One block per sentence a.  rowloss[a] out; g_wmax[a,c,w] = alpha/B (softmax[a,c] - [a==c]) / n
// Dynamic smem: B + 64 floats.
// ------------------------------------------------------------------------------------------
__global__ __launch_bounds__(128) void vg_loss_kernel(int B, int n, const float* __restrict__ wmax, float alpha,
                                                      float* __restrict__ rowloss, float* __restrict__ g_wmax) {
  pdl_prologue();
  extern __shared__ float sm[];
  float* s_logit = sm;
  float* s_red = sm + B;
  const int a = blockIdx.x;
  for (int c = threadIdx.x; c < B; c += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < n; ++w) s += wmax[((int64_t)a * B + c) * n + w];
    s_logit[c] = s / (float)n;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < B; c += blockDim.x) mx = fmaxf(mx, s_logit[c]);
  // block max via sum trick is not available; reduce with shuffles + smem
  mx = warp_max(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) mx = fmaxf(mx, s_red[i]);
  float se = 0.f;
  for (int c = threadIdx.x; c < B; c += blockDim.x) se += expf(s_logit[c] - mx);
  se = block_sum(se, s_red);
  const float lse = mx + logf(se);
  if (threadIdx.x == 0) rowloss[a] = alpha * (lse - s_logit[a]) / (float)B;
  if (g_wmax == nullptr) return;
  const float k = alpha / ((float)B * (float)n);
  for (int c = threadIdx.x; c < B; c += blockDim.x) {
    const float gl = (expf(s_logit[c] - lse) - (c == a ? 1.f : 0.f)) * k;
    for (int w = 0; w < n; ++w) g_wmax[((int64_t)a * B + c) * n + w] = gl;
  }
}

__global__ void sum_small_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
  pdl_prologue();
  __shared__ float red[64];
  float t = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t += v[i];
  t = block_sum(t, red);
  if (threadIdx.x == 0) *out = t;
}

}  // namespace cliora
