/*
 * cliora_b200 -- C ABI of the B200-native CLIORA chart hot path.
 *
 * This is the drop-in boundary: a plain-C shared library (libcliora_b200.so)
 * whose entry points take device pointers into caller-owned buffers, sizes and
 * a CUDA stream.  No torch types.  The library allocates nothing persistent and
 * keeps no pointer after a call returns.  The Python host side
 * (cliora_b200/net/*.py) binds these with ctypes and mirrors the reference's
 * cliora/net module API on top (INTEGRATION.md shows the binding).
 *
 * All tensors are fp32, row-major, contiguous.  Chart tensors are laid out
 * exactly like the reference's Chart (cliora/net/diora.py:7-23): [B, cells, D]
 * with cells level-major, cell (level l, pos p) at l*n - l(l-1)/2 + p
 * (cliora/net/offset_cache.py:1-7).
 *
 * Every function returns 0 on success or a negative cliora_status; the Python
 * wrapper raises RuntimeError (the reference lets exceptions propagate out of
 * Trainer.step, cliora/net/trainer.py:469-481).  There is no CPU fallback.
 */
#ifndef CLIORA_B200_H_
#define CLIORA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cliora_status {
  CLIORA_OK = 0,
  CLIORA_ERR_BAD_SHAPE = -1,    /* B,n,D,R out of range, D % 4 != 0, ... */
  CLIORA_ERR_NULL_POINTER = -2, /* a required buffer is NULL */
  CLIORA_ERR_CUDA = -3,         /* a CUDA runtime call or kernel launch failed */
  CLIORA_ERR_NO_DEVICE = -4,    /* no sm_100 device: the library never falls back to the CPU */
  CLIORA_ERR_UNSUPPORTED = -5
} cliora_status;

/* Opaque to C callers that only pass it through; identical to cudaStream_t. */
typedef void* cliora_stream_t;

const char* cliora_status_string(int status);
/* Last CUDA error text recorded by this thread's most recent failing call. */
const char* cliora_last_cuda_error(void);
int cliora_abi_version(void);

/* ------------------------------------------------------------------------
 * Chart geometry (replaces cliora/net/offset_cache.py:1-7,
 * cliora/net/inside_index.py:182-197, cliora/net/outside_index.py:93-127).
 * Host-side, no GPU needed.  The kernels use the same closed forms inline
 * instead of index tensors.
 * ---------------------------------------------------------------------- */
int64_t cliora_num_cells(int n);
int64_t cliora_level_offset(int n, int level);
/* left/right child chart indices of every (pos, split) of an inside level; out arrays hold (n-level)*level entries. */
int cliora_inside_index(int n, int level, int64_t* left, int64_t* right);
/* parent/sibling chart indices of an outside level in the reference's (split, pos) order; (n-level-1)*(n-level) entries. */
int cliora_outside_index(int n, int level, int64_t* parent, int64_t* sibling);
/* Row offset (in split rows, all B sentences) of a level's block inside the per-split buffers. */
int64_t cliora_split_row_offset(int B, int n, int level, int outside);

/* ------------------------------------------------------------------------
 * Model weights = the reference DioraMLP state_dict tensors
 * (cliora/net/diora.py:453-471; keys listed in SURVEY.md section 8b).
 * nn.Linear layout [out, in].  When share != 0 the o* pointers are ignored.
 * ---------------------------------------------------------------------- */
typedef struct cliora_weights {
  const float* W_leaf; /* inside_compose_func.leaf_fc.weight   [D, D]  */
  const float* b_leaf; /* inside_compose_func.leaf_fc.bias     [D]     */
  const float* W1;     /* inside_compose_func.h_fcs.0.weight   [D, 2D] */
  const float* b1;     /* inside_compose_func.h_fcs.0.bias     [D]     */
  const float* W2;     /* inside_compose_func.h_fcs.2.weight   [D, D]  */
  const float* b2;     /* inside_compose_func.h_fcs.2.bias     [D]     */
  const float* Wb;     /* inside_score_func.mat                [D, D]  */
  const float* root;   /* root_vector_out_h                    [D]     */
  const float* oW1;    /* outside_compose_func.h_fcs.0.weight  (share == 0) */
  const float* ob1;
  const float* oW2;
  const float* ob2;
  const float* oWb;    /* outside_score_func.mat */
} cliora_weights;

/* Same fields, writable: gradients.  Every non-NULL field is fully overwritten. */
typedef struct cliora_weight_grads {
  float *W_leaf, *b_leaf, *W1, *b1, *W2, *b2, *Wb, *root, *oW1, *ob1, *oW2, *ob2, *oWb;
} cliora_weight_grads;

typedef struct cliora_dims {
  int B;       /* sentences in the batch */
  int n;       /* words per sentence (all sentences in a batch share it, dataloader.py:52-91) */
  int D;       /* hidden size, D % 4 == 0 */
  int R;       /* regions per image; 0 = text-only DIORA (cliora/net/diora.py) */
  int share;   /* 1: outside pass uses the inside weights (--share default) */
  int flags;   /* CLIORA_FLAG_* */
} cliora_dims;

#define CLIORA_PHASE_LEVELS 1
#define CLIORA_PHASE_WEIGHTS 2
#define CLIORA_FLAG_DETERMINISTIC 1 /* reserved */
/* Reduced-precision mode: the tensor-core GEMMs issue one TF32 pass instead of the fp32-accurate three.
 * Stated tolerance 1e-2 of max on chart vectors (measured ~3e-3); CKY trees are NOT guaranteed identical. */
#define CLIORA_FLAG_TF32_1PASS 2
/* --normalize none (cliora/net/utils.py:17-27): every unit-normalisation of the chart (leaves, cells, the CLIORA
 * attention residual, the outside root) becomes the identity.  The saved norms hold the sentinel -1. */
#define CLIORA_FLAG_NO_NORMALIZE 16
/* bf16-GEMM mode: the compose GEMMs of the fused level kernels take bf16 operands (fp32 accumulate, kind::f16 UMMAs);
 * every other tensor-core GEMM runs single-pass TF32.  Stated tolerance 3e-2 of max on chart vectors, 1e-2 on scores
 * (SURVEY.md section 7); CKY trees are NOT guaranteed identical.  Implies the fused level kernels. */
#define CLIORA_FLAG_BF16 8
/* Run the unfused per-level kernel chain (split_build -> tcgen05 GEMM -> cell kernels -> scatter) instead of the fused
 * level kernels.  Same results; it is the faster of the two once a level no longer fits one wave of clusters (batches
 * above ~32 sentences at length 20), where per-tile SM time rather than per-level latency decides. */
#define CLIORA_FLAG_UNFUSED 4
/* Bits 8-11: how many sentence chains (independent sub-batches on separate streams) the caller runs concurrently.
 * A sizing hint only: each chain's fused level kernels then aim at 1/k of the co-resident clusters. */
#define CLIORA_FLAG_CHAINS(k) (((k) & 15) << 8)

/* Float offsets of every sub-buffer inside the single forward workspace `ws`
 * (saved for backward) and the backward scratch `bws`.  -1 = not present. */
typedef struct cliora_layout {
  int64_t ws_floats;   /* size of the forward workspace, in floats */
  int64_t bws_floats;  /* size of the backward scratch, in floats  */
  int64_t rows_in;     /* inside split rows over the whole batch  = B (n-1)n(n+1)/6 */
  int64_t rows_out;    /* outside split rows                      = 2 rows_in       */
  int64_t PI;          /* projections per inside cell (3, or 4 when share == 0) */
  /* forward workspace */
  int64_t Pin, Pout;           /* [B,C,PI*D], [B,C,2D] per-cell projections */
  int64_t q_in;                /* [B,C,D] pre-attention unit vectors (R > 0 only) */
  int64_t nrm_in, nrm2_in;     /* [B,C] */
  int64_t att_in;              /* [B,C,R] softmax over regions (R > 0 only) */
  int64_t nrm_out;             /* [B,C] */
  int64_t leaf_t;              /* [B*n, D] tanh(W_leaf x + b) */
  int64_t Zin, Yin, Ein, Prin;     /* [2, rows_in, D] x2 (split pairs hi|lo), [rows_in] x2 */
  int64_t Zout, Yout, Eout, Prout; /* [2, rows_out, D] x2, [rows_out] x2 */
  int64_t Wcat_in, Wcat_out;   /* packed projection weights [PI*D, D], [2D, D] */
  /* backward scratch */
  int64_t Gh_in, Gs_in, GP_in, Gh_out, Gs_out, GP_out;
  int64_t GA2, coef;          /* [B,C,D], [B,C,2R] attention backward intermediates (R > 0 only) */
  int64_t GE, GZ, splitk, gu; /* per-level split scratch, split-K partials, leaf pre-activation grads */
  /* forward workspace, tensor-core operands: split pairs [2, D, D] of W2 and W2^T (o* alias when share) */
  int64_t W2p, W2Tp, oW2p, oW2Tp;
  /* backward scratch, tensor-core weight gradients: split pairs of GP [2,B*C,PI*D] and of the chart vectors [2,B*C,D] */
  int64_t GPp, Hp;
  /* forward workspace: ReLU bitmasks of the hidden activations, 16 x uint32 per split row (D <= 512), else -1 */
  int64_t Mbin, Mbout;
  int64_t CSin, CSout;        /* bws: per-cell sums over splits of the compose-output gradients [B,C,D] (-> db2) */
  /* backward scratch of the fused level kernels: split pairs of the compose-output gradients (the A operand of the
   * weight-gradient GEMM), per-cell gradients wrt the pre-normalisation sums, per-cell softmax constants, and the
   * bias-gradient accumulators of both passes */
  int64_t GYp_in, GYp_out;    /* [2, rows_in, D], [2, rows_out, D] */
  int64_t GA, CM;             /* [B,C,D], [B,C] */
  int64_t db2acc;             /* [2, D]: inside, outside */
  /* forward workspace, bf16 mode: bf16 copies of W2 and W2^T (o* alias when share), D*D/2 floats each */
  int64_t W2h, W2Th, oW2h, oW2Th;
} cliora_layout;

int cliora_chart_layout(const cliora_dims* dims, cliora_layout* out);

/* ------------------------------------------------------------------------
 * Forward.  Replaces DioraBase.forward's leaf_transform + inside_pass
 * (cliora/net/diora.py:283-331, cliora/net/cliora.py:290-341) and
 * outside_pass (diora.py:337-398).
 *
 *  x          [B,n,D]     projected word embeddings (x_span)
 *  obj        [B,R,D]     projected region features, NULL when R == 0
 *  keep       [B,cells,R] uint8 dropout keep-mask of AttentionHead (cliora.py:32,40);
 *                         NULL = dropout off (eval).  Kept probabilities are scaled 1/0.9.
 *  inside_h   [B,cells,D] out;  inside_s [B,cells] out
 *  ws         forward workspace of cliora_layout.ws_floats floats
 *
 * After the call, per-level views of ws give what the reference passes to
 * inside_hook(level, h, c, s) (diora.py:333): h = Yin rows of the level
 * [B*L*N, D] in (b,pos,split) order, s = Ein rows [B,L,N,1]; c is all zeros.
 * ---------------------------------------------------------------------- */
int cliora_inside_fwd(const cliora_dims* dims, const cliora_weights* w, const float* x, const float* obj,
                      const uint8_t* keep, float* inside_h, float* inside_s, float* ws, cliora_stream_t stream);

int cliora_outside_fwd(const cliora_dims* dims, const cliora_weights* w, const float* inside_h,
                       const float* inside_s, float* outside_h, float* outside_s, float* ws,
                       cliora_stream_t stream);

/* ------------------------------------------------------------------------
 * Backward (replaces torch autograd over the reference graph; maths in
 * SURVEY.md section 8a last row, derivation checked in oracle/factored.py).
 *
 * cliora_chart_bwd_begin  seeds the accumulators in bws from the incoming
 *                         chart cotangents (any may be NULL = zero).
 * cliora_outside_bwd      outside pass backward, levels 0 -> n-2, then the root vector.
 * cliora_inside_bwd       inside pass backward, levels n-1 -> 1, leaves, weight grads.
 *                         `had_outside` says whether cliora_outside_bwd ran on this bws.
 * Destroys the Y buffers in ws (overwritten by their gradients).
 *
 * `phase` (CLIORA_PHASE_LEVELS | CLIORA_PHASE_WEIGHTS, or either alone in that order): the level chain and
 * the weight-gradient GEMMs of a pass.  The outside pass's weight phase only reads what the outside level phase
 * produced, so a caller may run it on a second stream concurrently with the inside level phase, and join before
 * the inside weight phase (which accumulates into the same gradient tensors when share != 0).
 * ---------------------------------------------------------------------- */
int cliora_chart_bwd_begin(const cliora_dims* dims, const float* g_inside_h, const float* g_inside_s,
                           const float* g_outside_h, const float* g_outside_s, float* bws,
                           cliora_stream_t stream);

int cliora_outside_bwd(const cliora_dims* dims, const cliora_weights* w, const float* inside_h,
                       const float* inside_s, const float* outside_h, const float* outside_s, float* ws,
                       float* bws, cliora_weight_grads* grads, int phase, cliora_stream_t stream);

int cliora_inside_bwd(const cliora_dims* dims, const cliora_weights* w, const float* x, const float* obj,
                      const uint8_t* keep, const float* inside_h, const float* inside_s,
                      const float* outside_h, float* ws, float* bws, int had_outside, float* grad_x,
                      float* grad_obj, cliora_weight_grads* grads, int phase, cliora_stream_t stream);

/* ------------------------------------------------------------------------
 * Span-region alignment (replaces the einsums of cliora/net/cliora.py:457-466
 * and the losses of cliora/net/trainer.py:81-171).
 * ---------------------------------------------------------------------- */

/* scores[a, c, cell, r] = h[a, cell] . obj[c, r]   -> [B, B, ncell, R]   (cliora.py:457 with
 * h = inside_h + outside_h summed by the caller; cliora.py:459 with h = x_word, ncell = n).
 * h rows of sentence a start at row a * h_batch_stride.  Materialises the full tensor: only for
 * callers that read diora.all_atten_score / vg_atten_score. */
int cliora_atten_scores(int B, int ncell, int D, int R, const float* h, int64_t h_batch_stride,
                        const float* obj, float* scores, cliora_stream_t stream);

/* smax[a, c, cell] = max_r h[a,cell] . obj[c,r], amax = argmax (first max), WITHOUT materialising
 * the scores.  Only the first `ncell` cells of each sentence are scored (R <= 64). */
int cliora_atten_max_fwd(int B, int ncell, int D, int R, const float* h, int64_t h_batch_stride,
                         const float* obj, float* smax, int32_t* amax, cliora_stream_t stream);

/* Backward of cliora_atten_max_fwd: g_h[a,cell] += sum_c g_smax[a,c,cell] obj[c, amax];
 * g_obj[c, r] += sum_{a,cell: amax == r} g_smax[a,c,cell] h[a,cell].  Both are ACCUMULATED into
 * (caller zero-fills); either may be NULL. */
int cliora_atten_max_bwd(int B, int ncell, int D, int R, const float* h, int64_t h_batch_stride,
                         const float* obj, const float* g_smax, const int32_t* amax, float* g_h,
                         int64_t gh_batch_stride, float* g_obj, cliora_stream_t stream);

/* ContrastiveLoss.forward (trainer.py:91-128) on smax [B,B,ncell] (ncell = cells//2 is what the
 * loss reads): loss (1 float on device) and, when g_smax != NULL, gradients wrt smax [B,B,ncell],
 * inside_s and outside_s ([B,cells], caller zero-fills).  scratch: ncell*(B+1)+4 floats. */
int cliora_contrastive_loss(int B, int cells, int ncell, const float* smax, const float* inside_s,
                            const float* outside_s, float margin, float alpha, float* loss_out,
                            float* g_smax, float* g_inside_s, float* g_outside_s, float* scratch,
                            cliora_stream_t stream);

/* VGLoss.forward (trainer.py:139-171) on wmax [B,B,n] = max_r word-region scores:
 * logits[a,c] = mean_w wmax[a,c,w]; loss = alpha * CE(logits, arange(B)).  g_wmax may be NULL.
 * scratch: B floats. */
int cliora_vg_loss(int B, int n, const float* wmax, float alpha, float* loss_out, float* g_wmax,
                   float* scratch, cliora_stream_t stream);

/* ReconstructionSoftmaxLoss.forward (trainer.py:46-78), the 1+K-way scoring and cross-entropy fused:
 *   rows = B*n words; cell [rows, D] = outside_h leaves; pos [rows, D], neg [K, D] = projected embeddings
 *   rowloss[row] = logsumexp(s) - s_0 with s_0 = pos.cell, s_{1+e} = neg[e].cell; probs [rows, K+1] = softmax(s)
 * The loss is mean(rowloss).  Backward: g_scores [rows, K+1] = (probs - onehot_0) g_loss / rows,
 * g_cell / g_pos [rows, D]; the gradient wrt neg is g_scores[:, 1:]^T cell (cliora_matmul_tn).
 * D <= 512, K <= 127. */
int cliora_recon_ce_fwd(int rows, int D, int K, const float* cell, const float* pos, const float* neg, float* rowloss,
                        float* probs, cliora_stream_t stream);
int cliora_recon_ce_bwd(int rows, int D, int K, const float* cell, const float* pos, const float* neg,
                        const float* probs, const float* g_loss, float* g_scores, float* g_cell, float* g_pos,
                        cliora_stream_t stream);

/* ------------------------------------------------------------------------
 * CKY decode (replaces ParsePredictor.batched_cky, cliora/analysis/cky.py:31-99,
 * fed by the hook of cliora/analysis/utils.py:78-95).
 *  split_scores = the Ein region of ws (raw inside split scores, all levels)
 *  backptr [B, cells] int32: best split k per cell (first max wins), -1 at leaves
 *  best    [B, cells] fp32 : Viterbi scores (may be NULL while the chart fits shared memory, n <= 319; longer
 *                            sentences keep their chart in these rows: NULL then returns CLIORA_ERR_UNSUPPORTED)
 * One block per sentence, one warp per cell, lanes over the splits (warp-shuffle max / first-max argmax).
 * ---------------------------------------------------------------------- */
int cliora_cky(int B, int n, const float* split_scores, int32_t* backptr, float* best, cliora_stream_t stream);

/* ------------------------------------------------------------------------
 * Fused optimiser step (replaces Trainer.gradient_update's clip_grad_norm_(params, 5.0) + Adam.step,
 * cliora/net/trainer.py:450-455): global-norm clip and Adam over a table of tensors in three launches.
 *   cliora_adam_table_fill  fills a HOST staging table (cliora_adam_table_bytes(n) bytes) from pointer arrays;
 *                           the caller copies it to the device once.  Pointers must stay valid (graph-safe:
 *                           gradients are written in place every step).
 *   cliora_adam_step        state [3] floats on the device: out ||g||, out clip coefficient, in/out step count
 *                           (incremented on the device); scratch: total_blocks floats.
 * ---------------------------------------------------------------------- */
int64_t cliora_adam_table_bytes(int ntensors);
int cliora_adam_table_fill(int ntensors, void* const* params, const void* const* grads, void* const* exp_avg,
                           void* const* exp_avg_sq, const int64_t* numel, void* host_table, int64_t* total_blocks);
int cliora_adam_step(const void* device_table, int ntensors, int64_t total_blocks, float lr, float beta1, float beta2,
                     float eps, float max_norm, float* state, float* scratch, cliora_stream_t stream);

/* Constituent spans of the decoded trees, on the device (replaces tree -> str -> get_actions -> get_spans,
 * cliora/analysis/utils.py:3-48, scripts/parse.py:215-219).  spans [B, n-1, 2] int32: (start, end) inclusive
 * word positions in post-order (the last entry is the whole sentence); scratch: 3*B*n int32. */
int cliora_tree_spans(int B, int n, const int32_t* backptr, int32_t* spans, int32_t* scratch, cliora_stream_t stream);

/* Bracketing scores on the device (replaces get_stats + the sentence-F1 arithmetic of scripts/parse.py:216-233):
 * spans from cliora_tree_spans, gold [B, G, 2] padded with gold_len[b] valid entries (the last one is dropped like
 * the reference's GT[:-1]); out [B, 4] = tp, fp, fn, sentence F1. */
int cliora_span_f1(int B, int n, int G, const int32_t* spans, const int32_t* gold, const int32_t* gold_len, float* out,
                   cliora_stream_t stream);

/* Phrase-grounding recall on the device (replaces the per-phrase host loop of scripts/parse.py:174-212 and
 * scripts/train.py:158-179 over `diora.atten_score.cpu()`): phrases [P, 3] int32 = (sentence b, first word,
 * one-past-last word); for each, the word with the highest best-region score (first max) selects its best region
 * (first max), whose box [B, R, 4] (x1, y1, x2, y2) is compared with gt_boxes [P, 4] by the IoU of
 * torchvision.ops.box_iou.  sel [P, 2] = (word, region), iou [P], hit [P] = iou > iou_thresh (reference: 0.5).
 * An empty word range gives sel = (-1, -1), iou 0, hit 0. */
int cliora_grounding_eval(int B, int n, int R, int P, const float* atten_score, const float* boxes,
                          const int32_t* phrases, const float* gt_boxes, float iou_thresh, int32_t* sel, float* iou,
                          int32_t* hit, cliora_stream_t stream);

/* Batch assembly from the ragged region-feature table, on the device (replaces FlickrDataset.__getitem__,
 * cliora/data/dataloader.py:205-222, + collate + .cuda(), cliora/data/batch_iterator.py:116-168).  Image i owns
 * table rows [pos_bboxes[i,0], pos_bboxes[i,1]); batch entry b takes the first min(rows, R) rows of image
 * img_index[b] and pads the rest (features 0, boxes -1, classes -1).  features: [rows, F] fp32 (feat_dtype 0) or
 * fp16 (feat_dtype 1, widened to fp32 on the fly); bboxes [rows, 4]; classes [rows] int32 or NULL.  Table
 * pointers may be device memory or pinned host memory (zero-copy).  Outputs: obj_feats [B, R, F] fp32,
 * boxes [B, R, 4] (nullable), obj_cates [B, R] int64 (nullable).  img_index (device, int64) must be in range. */
int cliora_gather_regions(int B, int R, int F, int feat_dtype, const void* features, const float* bboxes,
                          const int32_t* classes, const int64_t* pos_bboxes, const int64_t* img_index, float* obj_feats,
                          float* boxes, int64_t* obj_cates, cliora_stream_t stream);

/* ------------------------------------------------------------------------
 * Dense helper used on both sides of the chart (Embed, ImageEncoder,
 * reconstruction loss; trainer.py:219-224, utils.py:52-55):
 *   C[M,N] = act(A[M,K] W[N,K]^T + bias)      act: 0 none, 1 relu, 2 tanh
 * ---------------------------------------------------------------------- */
int cliora_linear(int M, int N, int K, const float* A, const float* W, const float* bias, int act, float* C,
                  cliora_stream_t stream);
/* C[M,N] = A[M,K] B[K,N] (accumulate != 0: C += ...) */
int cliora_matmul_nn(int M, int N, int K, const float* A, const float* B, float* C, int accumulate,
                     cliora_stream_t stream);
/* C[Ka,Kb] = A[M,Ka]^T B[M,Kb]; scratch >= cliora_matmul_tn_scratch_floats(M,Ka,Kb) floats */
int64_t cliora_matmul_tn_scratch_floats(int M, int Ka, int Kb);
int cliora_matmul_tn(int M, int Ka, int Kb, const float* A, const float* B, float* C, int accumulate,
                     float* scratch, cliora_stream_t stream);

/* ------------------------------------------------------------------------
 * Tensor-core (tcgen05, 3xTF32) primitives.  Operands are SPLIT PAIRS: [2, rows, K] fp32 with
 * part 0 = tf32-rounded value, part 1 = exact remainder (cliora_split_tf32 builds one).
 *   C[M,N] = act(A[M,K] W[N,K]^T + bias)   with fp32-grade accuracy (~1e-6 of max)
 * ---------------------------------------------------------------------- */
int cliora_split_tf32(const float* x, int64_t n, float* out_pair, cliora_stream_t stream);
/* Development / measurement knobs (process-global; defaults are what ships):
 *   0  accumulation mode of the tcgen05 GEMM (2 = cross terms in a second TMEM accumulator [default], 0 = single
 *      accumulator, 1 = plain TF32, 3 = rotating accumulator sets)
 *   1  1 = every GEMM on the exact-fp32 SIMT kernel (no tensor cores)
 *   2  tcgen05 tile: 0 = cost model, 1 = 128x80, 2 = 128x256, 3 = 128x160
 *   3  vision-language cell kernels: bit 0 = block-per-cell forward, bit 1 = block-per-cell backward (default 2)
 *   4  1 = bias gradient db2 from the full GY rows instead of the per-cell sums
 *   5  1 = per-cell GEMMs on the fp32 SIMT kernel instead of the mma.sync 3xTF32 kernel
 * 100  programmatic dependent launch on/off (default off)      101  shared-memory carveout percent (default 100)
 * 102  1 = narrow tcgen05 tile allocates 256 TMEM columns       103  narrow tile pipeline depth (2, 3 [default], 4)
 * 104  CTA target of the small-GEMM split-K (default 4 x 148)   105  1 = allow the 128x48 tcgen05 tile */
void cliora_debug_set(int key, int value);
/* How one chart level is tiled by the fused level kernels (csrc/level_kernels.cuh) for the given problem -- pure host
 * arithmetic (the occupancy query falls back to a constant without a device), exposed so that the tiling rules can be
 * checked without a GPU (tests/test_level_plan_cpu.py) and inspected by callers.  `level`: inside 1..n-1, outside
 * 0..n-2; `backward`: the plan of level_bwd_kernel instead of level_fwd_kernel.  fused == 0: the level runs the unfused
 * kernel chain (split_build -> GEMM -> cell_aggregate) and the other fields are 0 except splits / cells. */
typedef struct cliora_level_plan {
  int32_t fused;
  int32_t splits;                 /* N: splits per cell at this level */
  int32_t cells;                  /* B * (n - level) */
  int32_t column_slices;          /* CTAs that share a tile (forward: the cluster size) */
  int32_t slice_cols;             /* output columns per CTA */
  int32_t umma_n;                 /* slice_cols rounded up to 16: the UMMA N */
  int32_t cells_per_tile;         /* G: whole cells per tile, G * N <= 128 split rows */
  int32_t tiles;
  int32_t max_sentences_per_tile;
  int32_t ring_bytes;             /* operand rings (raw A stages + W2 slice stages), reused by the epilogue */
  int64_t smem_bytes;             /* dynamic shared memory of the launch */
} cliora_level_plan;
int cliora_level_plan_query(const cliora_dims* dims, int level, int outside, int backward, cliora_level_plan* plan);

/* Development only: hands a device buffer to an instrumented kernel (key 0: int64 [ctas][32] timeline of the fused
 * level kernel selected by debug keys 8 (level + 1) and 9 (outside)). */
void cliora_debug_ptr(int key, void* p);
int cliora_tc_linear(int M, int N, int K, const float* A_pair, const float* W_pair, const float* bias, int act,
                     float* C, cliora_stream_t stream);

/* cliora_atten_max_fwd on tensor cores: the [B*ncell, D] x [B*R, D]^T alignment GEMM with the max-over-regions
 * (+ argmax) epilogue, operands as split pairs [2, B*ncell, D] and [2, B*R, D].  D >= 32, R <= 64. */
int cliora_tc_atten_max_fwd(int B, int ncell, int D, int R, const float* h_pair, const float* obj_pair, float* smax,
                            int32_t* amax, cliora_stream_t stream);

/* C[Ka,Kb] (+)= A[M,Ka]^T B[M,Kb] on tensor cores (MN-major UMMA, split-K); Ka, Kb multiples of 4. */
int64_t cliora_tc_matmul_tn_scratch_floats(int M, int Ka, int Kb);
int cliora_tc_matmul_tn(int M, int Ka, int Kb, const float* A_pair, const float* B_pair, float* C, int accumulate,
                        float* scratch, cliora_stream_t stream);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t cliora_launch_count(void);

/* Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline
 * pass; off by default, adds two event records per launch when on).  flops / bytes are the
 * ALGORITHMIC counts of the launches (DESIGN.md states the per-unit figures). */
typedef struct cliora_profile_row {
  char name[48];
  int64_t launches;
  double ms;     /* summed device time between the two events of every launch of this class */
  double flops;
  double bytes;
} cliora_profile_row;
void cliora_profile_start(void);
/* Synchronises the device, aggregates by kernel class into rows[0..ret), returns the row count. */
int cliora_profile_stop(cliora_profile_row* rows, int max_rows);

#ifdef __cplusplus
}
#endif
#endif /* CLIORA_B200_H_ */
