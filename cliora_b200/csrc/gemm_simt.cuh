// Generic fp32 SIMT GEMM (exact fp32 FMA accumulation) used for every dense contraction of the
// fp32-accurate path: compose W2, per-cell projections, their backward, weight gradients.
// One kernel template covers the three operand layouts:
//   NT  C[m,j] = sum_k A[m,k] W[j,k]      (nn.Linear forward)         A_KMAJOR=0 B_KMAJOR=0
//   NN  C[m,j] = sum_k A[m,k] W[k,j]      (grad wrt activations)      A_KMAJOR=0 B_KMAJOR=1
//   TN  C[i,j] = sum_r A[r,i] B[r,j]      (grad wrt weights, split-K) A_KMAJOR=1 B_KMAJOR=1
// Rows of A and C can be remapped (RowMap) so a chart level is read / written in place.
#pragma once
#include "common.cuh"

namespace cliora {

struct GemmParams {
  const float* A;
  int64_t lda;
  RowMap amap;
  const float* W;
  int64_t ldw;
  float* C;
  int64_t ldc;
  RowMap cmap;
  const float* bias;  // [N] or null
  const float* mask;  // same virtual rows as C (dense, ldm); result *= (mask > 0); or null
  int64_t ldm;
  int M, N, K;
  int act;         // 0 none, 1 relu, 2 tanh
  int accumulate;  // C += result
  int k_chunk;     // split-K: K range per blockIdx.z (multiple of 16); 0 = no split
  int64_t split_stride;  // floats between split-K partial outputs
  int vec_a, vec_w, vec_c;  // 16-byte vector access allowed (alignment checked on the host)
  const char* tag;          // profiler label (host only)
  int64_t a_lo_off, w_lo_off;  // operands stored as split pairs: value = X[i] + X[i + lo_off] (0: plain)
  int atomic_splitk;           // k_chunk > 0 and partials are red.add'ed straight into C (bias by split 0)
  int mma_ok;                  // caller allows the 3xTF32 mma.sync kernel (fp32-grade, not bit-exact fp32 FMA order)
};

extern int g_splitk_target;   // CTAs a small GEMM is split up to along K (default 4 x 148: 296 / 444 / 592 measured 5814 / 5834 / 5837 sent/s)
constexpr int kBN = 64;
constexpr int kBK = 16;

template <int TM, bool A_KMAJOR, bool B_KMAJOR>
__global__ __launch_bounds__(256) void gemm_simt_kernel(const GemmParams p) {
  pdl_prologue();
  constexpr int BM = 16 * TM;
  constexpr int A_LD = BM + 4;
  constexpr int B_LD = kBN + 4;
  constexpr int A_PER_THREAD = BM / 64;  // float4 loads of A per thread per k-tile
  __shared__ __align__(16) float As[2][kBK][A_LD];
  __shared__ __align__(16) float Bs[2][kBK][B_LD];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * kBN;
  int k_begin = 0, k_end = p.K;
  if (p.k_chunk > 0) {
    k_begin = blockIdx.z * p.k_chunk;
    k_end = min(p.K, k_begin + p.k_chunk);
  }

  // ---- per-thread load coordinates ----
  int a_r[A_PER_THREAD], a_c[A_PER_THREAD];      // tile-local (row in BM or k-row, column quad)
  const float* a_base[A_PER_THREAD];
  bool a_valid[A_PER_THREAD];
#pragma unroll
  for (int i = 0; i < A_PER_THREAD; ++i) {
    const int idx = tid + i * 256;
    if (!A_KMAJOR) {
      a_r[i] = idx >> 2;         // m row in tile
      a_c[i] = (idx & 3) * 4;    // k quad
      const int m = m0 + a_r[i];
      a_valid[i] = m < p.M;
      a_base[i] = a_valid[i] ? p.A + map_row(p.amap, m) * p.lda : p.A;
    } else {
      a_r[i] = idx / (BM / 4);         // k row in tile
      a_c[i] = (idx % (BM / 4)) * 4;   // output-row quad
      a_valid[i] = true;
      a_base[i] = p.A + (m0 + a_c[i]);
    }
  }
  int b_r, b_c;
  const float* b_base;
  bool b_valid;
  if (!B_KMAJOR) {
    b_r = tid >> 2;        // j row in tile
    b_c = (tid & 3) * 4;   // k quad
    b_valid = (n0 + b_r) < p.N;
    b_base = b_valid ? p.W + (int64_t)(n0 + b_r) * p.ldw : p.W;
  } else {
    b_r = tid >> 4;         // k row in tile
    b_c = (tid & 15) * 4;   // j quad
    b_valid = true;
    b_base = p.W + (n0 + b_c);
  }

  float4 ra[A_PER_THREAD], rb;

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_PER_THREAD; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!A_KMAJOR) {
        const int k = k0 + a_c[i];
        if (a_valid[i]) {
          if (p.vec_a && k + 3 < k_end) {
            v = ld4(a_base[i] + k);
          } else {
            if (k + 0 < k_end) v.x = a_base[i][k + 0];
            if (k + 1 < k_end) v.y = a_base[i][k + 1];
            if (k + 2 < k_end) v.z = a_base[i][k + 2];
            if (k + 3 < k_end) v.w = a_base[i][k + 3];
          }
        }
      } else {
        const int k = k0 + a_r[i];
        const int c = m0 + a_c[i];
        if (k < k_end) {
          const float* src = a_base[i] + (int64_t)k * p.lda;
          if (p.vec_a && c + 3 < p.M) {
            v = ld4(src);
          } else {
            if (c + 0 < p.M) v.x = src[0];
            if (c + 1 < p.M) v.y = src[1];
            if (c + 2 < p.M) v.z = src[2];
            if (c + 3 < p.M) v.w = src[3];
          }
        }
      }
      if (p.a_lo_off != 0) {
        // split-pair operand (vector path only; pairs are always 16-byte aligned)
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!A_KMAJOR) {
          const int k = k0 + a_c[i];
          if (a_valid[i] && k + 3 < k_end) w = ld4(a_base[i] + p.a_lo_off + k);
        } else {
          const int k = k0 + a_r[i];
          const int c = m0 + a_c[i];
          if (k < k_end && c + 3 < p.M) w = ld4(a_base[i] + (int64_t)k * p.lda + p.a_lo_off);
        }
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
      ra[i] = v;
    }
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!B_KMAJOR) {
        const int k = k0 + b_c;
        if (b_valid) {
          if (p.vec_w && k + 3 < k_end) {
            v = ld4(b_base + k);
          } else {
            if (k + 0 < k_end) v.x = b_base[k + 0];
            if (k + 1 < k_end) v.y = b_base[k + 1];
            if (k + 2 < k_end) v.z = b_base[k + 2];
            if (k + 3 < k_end) v.w = b_base[k + 3];
          }
        }
      } else {
        const int k = k0 + b_r;
        const int c = n0 + b_c;
        if (k < k_end) {
          const float* src = b_base + (int64_t)k * p.ldw;
          if (p.vec_w && c + 3 < p.N) {
            v = ld4(src);
          } else {
            if (c + 0 < p.N) v.x = src[0];
            if (c + 1 < p.N) v.y = src[1];
            if (c + 2 < p.N) v.z = src[2];
            if (c + 3 < p.N) v.w = src[3];
          }
        }
      }
      if (p.w_lo_off != 0) {
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!B_KMAJOR) {
          const int k = k0 + b_c;
          if (b_valid && k + 3 < k_end) w = ld4(b_base + p.w_lo_off + k);
        } else {
          const int k = k0 + b_r;
          const int c = n0 + b_c;
          if (k < k_end && c + 3 < p.N) w = ld4(b_base + (int64_t)k * p.ldw + p.w_lo_off);
        }
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
      rb = v;
    }
  };

  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER_THREAD; ++i) {
      if (!A_KMAJOR) {
        As[buf][a_c[i] + 0][a_r[i]] = ra[i].x;
        As[buf][a_c[i] + 1][a_r[i]] = ra[i].y;
        As[buf][a_c[i] + 2][a_r[i]] = ra[i].z;
        As[buf][a_c[i] + 3][a_r[i]] = ra[i].w;
      } else {
        st4(&As[buf][a_r[i]][a_c[i]], ra[i]);
      }
    }
    if (!B_KMAJOR) {
      Bs[buf][b_c + 0][b_r] = rb.x;
      Bs[buf][b_c + 1][b_r] = rb.y;
      Bs[buf][b_c + 2][b_r] = rb.z;
      Bs[buf][b_c + 3][b_r] = rb.w;
    } else {
      st4(&Bs[buf][b_r][b_c], rb);
    }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (k_end - k_begin + kBK - 1) / kBK;
  if (nk > 0) {
    load_tiles(k_begin);
    store_tiles(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) load_tiles(k_begin + (kt + 1) * kBK);
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[TM];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = ld4(&As[cur][kk][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
      const float4 b = ld4(&Bs[cur][kk][tx * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
      }
    }
    if (kt + 1 < nk) store_tiles(cur ^ 1);
    __syncthreads();
  }

  // ---- epilogue ----
  const int c0 = n0 + tx * 4;
  if (c0 >= p.N) return;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias != nullptr && (!p.atomic_splitk || blockIdx.z == 0)) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c0 + j < p.N) bias[j] = p.bias[c0 + j];
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = m0 + ty * TM + i;
    if (r >= p.M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = acc[i][j] + bias[j];
      if (p.act == 1) t = fmaxf(t, 0.f);
      else if (p.act == 2) t = tanhf(t);
      v[j] = t;
    }
    float* dst;
    if (p.atomic_splitk) {
      dst = p.C + map_row(p.cmap, r) * p.ldc + c0;
      const bool fullv = (c0 + 3 < p.N) && p.vec_c;
      if (fullv) {
        red_add4(dst, make_float4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c0 + j < p.N) atomicAdd(dst + j, v[j]);
      }
      continue;
    }
    if (p.k_chunk > 0) {
      dst = p.C + (int64_t)blockIdx.z * p.split_stride + (int64_t)r * p.N + c0;  // dense partial
    } else {
      dst = p.C + map_row(p.cmap, r) * p.ldc + c0;
    }
    if (p.mask != nullptr) {
      const float* mk = p.mask + (int64_t)r * p.ldm + c0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + j < p.N && !(mk[j] > 0.f)) v[j] = 0.f;
    }
    const bool full = (c0 + 3 < p.N);
    if (full && p.vec_c) {
      float4 o = make_float4(v[0], v[1], v[2], v[3]);
      if (p.accumulate && p.k_chunk == 0) {
        const float4 old = ld4(dst);
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      st4(dst, o);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + j < p.N) dst[j] = (p.accumulate && p.k_chunk == 0) ? dst[j] + v[j] : v[j];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Small-M GEMM on the warp-level tensor-core path (mma.sync m16n8k8 TF32, 3xTF32 error compensation done in
// registers: x = hi + lo, acc += lo*hi + hi*lo + hi*hi in fp32).  Used for the per-cell projection and
// cell-gradient GEMMs (M = B * cells-of-a-level <= a few hundred rows): those launches are one wave, so the
// tcgen05 kernel's fixed cost (TMEM allocation, descriptor fetch, first TMA round trip, ~6 us) and its operand
// pair format buy nothing there, while the fp32 FMA pipe caps the SIMT kernel.  A row-major [M,K] with RowMap;
// W is [N,K] (B_KMAJOR = false, "NT") or [K,N] (B_KMAJOR = true, "NN").  No activation / mask.
// 64x64x16 tiles, 4 warps as 2 (m) x 2 (n), warp tile 32x32.  Shared tiles are laid out so that every
// fragment load is bank-conflict free: [row][k] with stride 20, or [k][n] with stride 72.
// ------------------------------------------------------------------------------------------
CL_D void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
CL_D void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
CL_D void red_add2(float* p, float x, float y) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(y) : "memory");
}

template <bool B_KMAJOR>
__global__ __launch_bounds__(128) void gemm_mma_kernel(const GemmParams p) {
  pdl_prologue();
  // 4 warps as 2 (m) x 2 (n), warp tile 32x32: every fragment value is split (cvt, sub, cvt) by the warp that
  // loads it, so the split cost per MMA falls with the warp tile (ncu on the 8-warp / 32x16 version: issue slots
  // 58 % busy, legacy tensor pipe 36 % -- instruction-issue bound by the splitting, not by the MMAs)
  constexpr int BM = 64, LDK = kBK + 4, LDN = kBN + 8, ST = 4;   // 4-stage cp.async pipeline
  __shared__ __align__(16) float As[ST][BM][LDK];
  __shared__ __align__(16) float Bs[ST][B_KMAJOR ? kBK * LDN : kBN * LDK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * kBN;
  int k_begin = 0, k_end = p.K;
  if (p.k_chunk > 0) {
    k_begin = blockIdx.z * p.k_chunk;
    k_end = min(p.K, k_begin + p.k_chunk);
  }
  // two 16-byte asynchronous copies of A and two of W per thread per k-tile (zero fill outside the matrix)
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto issue = [&](int kt) {
    const int buf = kt % ST, k0 = k_begin + kt * kBK;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = tid + h * 128;
      const int a_r = idx >> 2, a_q = (idx & 3) * 4;
      float* da = &As[buf][a_r][a_q];
      if (m0 + a_r < p.M && k0 + a_q < k_end)       // K % 4 == 0, chunks % 16 == 0
        cp_async16(da, p.A + map_row(p.amap, m0 + a_r) * p.lda + k0 + a_q);
      else
        st4(da, zero4);
      if (!B_KMAJOR) {
        const int b_r = idx >> 2, b_q = (idx & 3) * 4;       // n row, k quad
        float* db = &Bs[buf][b_r * LDK + b_q];
        if (n0 + b_r < p.N && k0 + b_q < k_end) cp_async16(db, p.W + (int64_t)(n0 + b_r) * p.ldw + k0 + b_q);
        else st4(db, zero4);
      } else {
        const int b_r = idx >> 4, b_q = (idx & 15) * 4;      // k row, n quad (N % 4 == 0 on this path)
        float* db = &Bs[buf][b_r * LDN + b_q];
        if (n0 + b_q < p.N && k0 + b_r < k_end) cp_async16(db, p.W + (int64_t)(k0 + b_r) * p.ldw + n0 + b_q);
        else st4(db, zero4);
      }
    }
  };
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

  const int nk = (k_end - k_begin + kBK - 1) / kBK;
#pragma unroll
  for (int s0 = 0; s0 < ST - 1; ++s0) {
    if (s0 < nk) issue(s0);
    cp_async_commit();              // one group per k-tile, empty ones included, so the wait count is uniform
  }
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt % ST;
    cp_async_wait<ST - 2>();        // k-tile kt has landed (for this thread's copies) ...
    __syncthreads();                // ... and for everyone's; also: everyone is done reading tile kt-1
    if (kt + ST - 1 < nk) issue(kt + ST - 1);   // overwrites the buffer of tile kt-1
    cp_async_commit();
#pragma unroll
    for (int kb = 0; kb < kBK; kb += 8) {
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int r = wm * 32 + mi * 16 + g;
        tf32_split(As[cur][r][kb + t], ah[mi][0], al[mi][0]);
        tf32_split(As[cur][r + 8][kb + t], ah[mi][1], al[mi][1]);
        tf32_split(As[cur][r][kb + t + 4], ah[mi][2], al[mi][2]);
        tf32_split(As[cur][r + 8][kb + t + 4], ah[mi][3], al[mi][3]);
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int cidx = wn * 32 + ni * 8 + g;
        const float b0 = B_KMAJOR ? Bs[cur][(kb + t) * LDN + cidx] : Bs[cur][cidx * LDK + kb + t];
        const float b1 = B_KMAJOR ? Bs[cur][(kb + t + 4) * LDN + cidx] : Bs[cur][cidx * LDK + kb + t + 4];
        tf32_split(b0, bh[ni][0], bl[ni][0]);
        tf32_split(b1, bh[ni][1], bl[ni][1]);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          mma_tf32(acc[mi][ni], al[mi], bh[ni]);   // small terms first
          mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
          mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
        }
    }
  }

  // ---- epilogue: c0,c1 -> (row g, cols 2t, 2t+1); c2,c3 -> (row g+8, same cols) ----
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = m0 + wm * 32 + mi * 16 + g + half * 8;
      if (r >= p.M) continue;
      float* crow = p.C + map_row(p.cmap, r) * p.ldc;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int col = n0 + wn * 32 + ni * 8 + 2 * t;
        if (col >= p.N) continue;            // N even on this path
        float x = acc[mi][ni][half * 2], y = acc[mi][ni][half * 2 + 1];
        if (p.bias != nullptr && (!p.atomic_splitk || blockIdx.z == 0)) {
          x += p.bias[col];
          y += p.bias[col + 1];
        }
        if (p.atomic_splitk) {
          red_add2(crow + col, x, y);
        } else {
          float2* dst = reinterpret_cast<float2*>(crow + col);
          if (p.accumulate) {
            const float2 old = *dst;
            x += old.x;
            y += old.y;
          }
          *dst = make_float2(x, y);
        }
      }
    }
}

// C[i*ldc + j] (+)= sum_z part[z][i][j]
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t split_stride, int Mo,
                                     int No, float* __restrict__ C, int64_t ldc, int accumulate) {
  pdl_prologue();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Mo * No) return;
  const int i = (int)(idx / No), j = (int)(idx % No);
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(int64_t)z * split_stride + idx];
  float* dst = C + (int64_t)i * ldc + j;
  *dst = accumulate ? (*dst + s) : s;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// NT / NN launcher.  W is [N,K] (nt) or [K,N] (!nt).
// `atomic_ok`: C already holds the value to accumulate onto (zeros or a running sum) and there is no
// activation / mask, so a small-M problem may be split along K over blockIdx.z with red.add epilogues.
inline int launch_gemm(cudaStream_t st, bool nt, GemmParams p, bool atomic_ok = false) {
  if (p.M <= 0 || p.N <= 0) return CLIORA_OK;
  p.k_chunk = 0;
  p.split_stride = 0;
  p.atomic_splitk = 0;
  p.vec_a = aligned16(p.A) && (p.lda % 4 == 0);
  p.vec_w = aligned16(p.W) && (p.ldw % 4 == 0);
  p.vec_c = aligned16(p.C) && (p.ldc % 4 == 0);
  const int64_t ctas128 = (int64_t)ceil_div(p.M, 128) * ceil_div(p.N, kBN);
  const bool mma = p.mma_ok && p.act == 0 && p.mask == nullptr && p.a_lo_off == 0 && p.w_lo_off == 0 &&
                   p.vec_a && p.vec_w && p.K % 4 == 0 && p.N % 4 == 0 && p.ldc % 2 == 0 &&
                   (reinterpret_cast<uintptr_t>(p.C) & 7) == 0;
  const bool big = !mma && ctas128 >= 2 * 148;     // the mma kernel always works on 64-row tiles
  ProfScope prof(st, p.tag ? p.tag : "gemm", 2.0 * p.M * p.N * p.K,
                 4.0 * ((double)p.M * p.K + (double)p.N * p.K + (double)p.M * p.N * (p.mask ? 2 : 1)));
  dim3 grid(ceil_div(p.N, kBN), ceil_div(p.M, big ? 128 : 64), 1);
  if (atomic_ok && !big && p.act == 0 && p.mask == nullptr) {
    const int ctas = grid.x * grid.y;
    const int ktiles = ceil_div(p.K, kBK);
    int splits = (g_splitk_target + ctas - 1) / ctas;
    if (splits > ktiles / 4) splits = ktiles / 4;   // at least 4 k-tiles (64 k) per split
    if (splits > 1) {
      p.k_chunk = ceil_div(ktiles, splits) * kBK;
      grid.z = ceil_div(p.K, p.k_chunk);
      p.atomic_splitk = 1;
    }
  }
  if (mma) {
    if (nt) launch_k(gemm_mma_kernel<false>, grid, 128, 0, st, p);
    else launch_k(gemm_mma_kernel<true>, grid, 128, 0, st, p);
    CL_CHECK_LAUNCH("gemm_mma_kernel");
    return CLIORA_OK;
  }
  if (nt) {
    if (big) launch_k(gemm_simt_kernel<8, false, false>, grid, 256, 0, st, p);
    else launch_k(gemm_simt_kernel<4, false, false>, grid, 256, 0, st, p);
  } else {
    if (big) launch_k(gemm_simt_kernel<8, false, true>, grid, 256, 0, st, p);
    else launch_k(gemm_simt_kernel<4, false, true>, grid, 256, 0, st, p);
  }
  CL_CHECK_LAUNCH("gemm_simt_kernel");
  return CLIORA_OK;
}

inline int tn_splits(int rows, int Ka, int Kb) {
  const int64_t tiles = (int64_t)ceil_div(Ka, 64) * ceil_div(Kb, kBN);
  int s = (int)((2 * 148 + tiles - 1) / tiles);
  const int max_by_rows = (rows + 255) / 256;  // at least 256 rows per split
  if (s > max_by_rows) s = max_by_rows;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return s;
}
inline int64_t tn_scratch_floats(int rows, int Ka, int Kb) {
  return (int64_t)tn_splits(rows, Ka, Kb) * Ka * Kb;
}

// C[Ka,Kb] (+)= A[rows,Ka]^T B[rows,Kb]   (deterministic split-K: partials to scratch, then ordered sum)
inline int launch_gemm_tn(cudaStream_t st, int rows, int Ka, int Kb, const float* A, int64_t lda, const float* B,
                          int64_t ldb, float* C, int64_t ldc, int accumulate, float* scratch,
                          const char* tag = "gemm_wgrad", int64_t a_lo_off = 0, int64_t b_lo_off = 0) {
  if (Ka <= 0 || Kb <= 0) return CLIORA_OK;
  ProfScope prof(st, tag, 2.0 * rows * Ka * Kb, 4.0 * ((double)rows * (Ka + Kb) + (double)Ka * Kb));
  const int splits = tn_splits(rows, Ka, Kb);
  GemmParams p{};
  p.A = A; p.lda = lda; p.amap = dense_rows();
  p.W = B; p.ldw = ldb;
  p.C = scratch; p.ldc = Kb; p.cmap = dense_rows();
  p.bias = nullptr; p.mask = nullptr; p.ldm = 0;
  p.M = Ka; p.N = Kb; p.K = rows;
  p.act = 0; p.accumulate = 0;
  p.a_lo_off = a_lo_off; p.w_lo_off = b_lo_off;
  int chunk = (rows + splits - 1) / splits;
  chunk = ((chunk + kBK - 1) / kBK) * kBK;
  if (chunk < kBK) chunk = kBK;
  p.k_chunk = chunk;
  p.split_stride = (int64_t)Ka * Kb;
  p.vec_a = aligned16(A) && (lda % 4 == 0);
  p.vec_w = aligned16(B) && (ldb % 4 == 0);
  p.vec_c = (Kb % 4 == 0) && aligned16(scratch);
  dim3 grid(ceil_div(Kb, kBN), ceil_div(Ka, 64), splits);
  launch_k(gemm_simt_kernel<4, true, true>, grid, 256, 0, st, p);
  CL_CHECK_LAUNCH("gemm_simt_kernel<tn>");
  const int64_t total = (int64_t)Ka * Kb;
  launch_k(splitk_reduce_kernel, ceil_div(total, 256), 256, 0, st, scratch, splits, p.split_stride, Ka, Kb, C, ldc, accumulate);
  CL_CHECK_LAUNCH("splitk_reduce_kernel");
  return CLIORA_OK;
}

}  // namespace cliora
