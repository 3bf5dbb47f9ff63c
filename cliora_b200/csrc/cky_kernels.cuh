// CKY argmax decode over the inside split scores (cliora/analysis/cky.py:31-99 fed by the hook of
// cliora/analysis/utils.py:78-95).  One block per sentence; the Viterbi chart lives in shared memory;
// the level loop runs inside the kernel (the reference does one device->host sync per cell).
#pragma once
#include "common.cuh"

namespace cliora {

// E: raw inside split scores, level blocks [B, L, N] at row offset B * inside_rows_before(n, level).
// best[l,p] = max_k best[k,p] + best[l-1-k,p+k+1] + (e_k - max_k e_k); leaves = 1; first max wins (torch.argmax,
// cky.py:86).  Warp-parallel: one block per sentence, one warp per cell of the current level, lanes over the cell's
// splits; the per-cell maximum and the first-max argmax are warp-shuffle reductions on (value, split) pairs.  Every
// candidate is computed with the oracle's own operation order, so scores and backpointers are bit-exact.
// The Viterbi chart lives in shared memory (GLOBAL_CHART = false: cells floats of dynamic smem) or, for sentences
// too long for that, in the caller's best_out rows (GLOBAL_CHART = true).
constexpr int kCkyThreads = 256;

template <bool GLOBAL_CHART>
__global__ __launch_bounds__(kCkyThreads) void cky_kernel(int B, int n, const float* __restrict__ E,
                                                          int32_t* __restrict__ backptr, float* __restrict__ best_out) {
  pdl_prologue();
  extern __shared__ float s_chart[];
  const int b = blockIdx.x;
  const int C = (int)num_cells(n);
  float* chart = GLOBAL_CHART ? best_out + (int64_t)b * C : s_chart;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    chart[i] = 1.f;
    backptr[(int64_t)b * C + i] = -1;
  }
  __syncthreads();
  for (int level = 1; level < n; ++level) {
    const int L = n - level, N = level;
    const int64_t lvl_rows = (int64_t)B * inside_rows_before(n, level);
    for (int p = warp; p < L; p += nwarps) {
      const float* e = E + lvl_rows + ((int64_t)b * L + p) * N;
      float mx = -INFINITY;
      for (int k = lane; k < N; k += 32) mx = fmaxf(mx, e[k]);
      mx = warp_max(mx);
      float bv = -INFINITY;
      int bk = 0x7fffffff;
      for (int k = lane; k < N; k += 32) {          // ascending k per lane: a strict > keeps the lane's first maximum
        int l, r;
        inside_children(n, level, p, k, l, r);
        const float lr = __fadd_rn(chart[l], chart[r]);
        const float cand = __fadd_rn(lr, __fsub_rn(e[k], mx));
        if (cand > bv) { bv = cand; bk = k; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {             // first maximum over the warp: larger value, then smaller split
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
        if (ov > bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
      }
      if (lane == 0) {
        const int cell = lvl_off(n, level) + p;
        chart[cell] = bv;
        backptr[(int64_t)b * C + cell] = bk;
      }
    }
    __syncthreads();
  }
  if (!GLOBAL_CHART && best_out != nullptr)
    for (int i = threadIdx.x; i < C; i += blockDim.x) best_out[(int64_t)b * C + i] = s_chart[i];
}

// Spans of the CKY tree straight from the backpointer table (replaces the host round trip
// tree -> str -> get_actions -> get_spans of cliora/analysis/utils.py:3-48 used by scripts/parse.py:215-219).
// One thread per sentence, explicit stack, post-order like get_spans: spans[b, i] = (start, end) inclusive word
// positions of the i-th reduced constituent, i < n-1 (the last one is the whole sentence).
__global__ void tree_spans_kernel(int B, int n, const int32_t* __restrict__ backptr, int32_t* __restrict__ spans,
                                  int32_t* __restrict__ stack_mem) {
  pdl_prologue();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int C = (int)num_cells(n);
  const int32_t* bp = backptr + (int64_t)b * C;
  int32_t* out = spans + (int64_t)b * (n - 1) * 2;
  int32_t* st = stack_mem + (int64_t)b * 3 * n;   // (level, pos, state) triples, depth <= n
  int sp = 0, emitted = 0;
  st[0] = n - 1; st[1] = 0; st[2] = 0; sp = 1;
  while (sp > 0) {
    int32_t* top = st + 3 * (sp - 1);
    const int level = top[0], pos = top[1], state = top[2];
    if (level == 0) { --sp; continue; }
    const int k = bp[lvl_off(n, level) + pos];
    if (state == 0) {          // visit left child (k, pos)
      top[2] = 1;
      int32_t* nx = st + 3 * sp++;
      nx[0] = k; nx[1] = pos; nx[2] = 0;
    } else if (state == 1) {   // visit right child (level-1-k, pos+k+1)
      top[2] = 2;
      int32_t* nx = st + 3 * sp++;
      nx[0] = level - 1 - k; nx[1] = pos + k + 1; nx[2] = 0;
    } else {                   // both children done: emit this constituent
      out[2 * emitted] = pos;
      out[2 * emitted + 1] = pos + level;
      ++emitted;
      --sp;
    }
  }
}

// Bracketing scores of predicted vs gold spans with the reference's set semantics (scripts/parse.py:216-233:
// gold = set(GT[:-1]), pred = set(get_spans(...)[:-1]), get_stats -> tp/fp/fn, sentence F1 with the 1e-8 guards).
// pred [B, n-1, 2] from tree_spans_kernel (last entry = whole sentence, dropped); gold [B, G, 2] padded,
// gold_len[b] entries valid of which the LAST is dropped like the reference's [:-1].  One thread per sentence.
// out [B, 4]: tp, fp, fn, sentence F1.
__global__ void span_f1_kernel(int B, int n, int G, const int32_t* __restrict__ pred, const int32_t* __restrict__ gold,
                               const int32_t* __restrict__ gold_len, float* __restrict__ out) {
  pdl_prologue();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int32_t* p = pred + (int64_t)b * (n - 1) * 2;
  const int32_t* g = gold + (int64_t)b * G * 2;
  const int np = n - 2;                              // predicted spans without the whole-sentence one
  const int ng = max(0, min(G, gold_len[b]) - 1);    // gold[:-1]
  int tp = 0, fn = 0, ngold = 0;
  for (int i = 0; i < np; ++i) {
    bool hit = false;
    for (int j = 0; j < ng && !hit; ++j) hit = (g[2 * j] == p[2 * i]) && (g[2 * j + 1] == p[2 * i + 1]);
    tp += hit;
  }
  for (int j = 0; j < ng; ++j) {
    bool dup = false;
    for (int k = 0; k < j && !dup; ++k) dup = (g[2 * k] == g[2 * j]) && (g[2 * k + 1] == g[2 * j + 1]);
    if (dup) continue;                               // set(): duplicates in the gold list count once
    ++ngold;
    bool hit = false;
    for (int i = 0; i < np && !hit; ++i) hit = (g[2 * j] == p[2 * i]) && (g[2 * j + 1] == p[2 * i + 1]);
    fn += !hit;
  }
  const int fp = np - tp;
  float prec = (float)tp / ((float)np + 1e-8f);
  float reca = (float)tp / ((float)ngold + 1e-8f);
  if (ngold == 0) {
    reca = 1.f;
    if (np == 0) prec = 1.f;
  }
  float* o = out + (int64_t)b * 4;
  o[0] = (float)tp; o[1] = (float)fp; o[2] = (float)fn;
  o[3] = 2.f * prec * reca / (prec + reca + 1e-8f);
}

// Phrase grounding recall (scripts/parse.py:174-212, scripts/train.py:158-179): for a phrase covering words
// [start, end) of sentence b, pick the word whose best region score is highest (first max), take that word's
// best region (first max), and compare the region's box with the annotated box: hit iff IoU > thresh.
// IoU follows torchvision.ops.box_iou (areas (x2-x1)(y2-y1), intersection clamped at 0).
// One warp per phrase; lanes stride over the R regions of a word.  phrases [P,3] = (b, start, end).
__global__ __launch_bounds__(128) void grounding_eval_kernel(int B, int n, int R, int P,
                                                             const float* __restrict__ atten,   // [B,n,R]
                                                             const float* __restrict__ boxes,   // [B,R,4]
                                                             const int32_t* __restrict__ phrases,
                                                             const float* __restrict__ gt_boxes,  // [P,4]
                                                             float thresh, int32_t* __restrict__ sel,
                                                             float* __restrict__ iou_out, int32_t* __restrict__ hit) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int ph = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ph >= P) return;
  const int b = phrases[3 * ph], start = max(phrases[3 * ph + 1], 0), end = min(phrases[3 * ph + 2], n);
  float best = -INFINITY;
  int bw = -1, br = -1;
  for (int w = start; w < end; ++w) {
    const float* row = atten + ((int64_t)b * n + w) * R;
    float v = -INFINITY;
    int vi = 0x7fffffff;
    for (int r = lane; r < R; r += 32) {
      const float x = row[r];
      if (x > v) { v = x; vi = r; }      // ascending r per lane: first max within the lane
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
      if (ov > v || (ov == v && oi < vi)) { v = ov; vi = oi; }
    }
    if (v > best) { best = v; bw = w; br = vi; }   // strict: the first word wins ties
  }
  if (lane != 0) return;
  float iou = 0.f;
  int h = 0;
  if (bw >= 0 && br < R) {
    const float* pb = boxes + ((int64_t)b * R + br) * 4;
    const float* gb = gt_boxes + (int64_t)ph * 4;
    const float a1 = __fmul_rn(pb[2] - pb[0], pb[3] - pb[1]);
    const float a2 = __fmul_rn(gb[2] - gb[0], gb[3] - gb[1]);
    const float iw = fmaxf(fminf(pb[2], gb[2]) - fmaxf(pb[0], gb[0]), 0.f);
    const float ih = fmaxf(fminf(pb[3], gb[3]) - fmaxf(pb[1], gb[1]), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fsub_rn(__fadd_rn(a1, a2), inter);
    iou = __fdiv_rn(inter, uni);
    h = iou > thresh ? 1 : 0;
  } else {
    br = -1;
  }
  sel[2 * ph] = bw;
  sel[2 * ph + 1] = br;
  iou_out[ph] = iou;
  hit[ph] = h;
}

}  // namespace cliora
