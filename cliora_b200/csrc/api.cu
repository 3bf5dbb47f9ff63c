// extern "C" entry points of libcliora_b200.so (declared in include/cliora_b200.h).
// Host-side orchestration only: level loops, buffer carving, kernel launches on the caller's stream.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <utility>

#include "align_kernels.cuh"
#include "chart_kernels.cuh"
#include "cell_warp_kernels.cuh"
#include "cky_kernels.cuh"
#include "data_kernels.cuh"
#include "recon_kernels.cuh"
#include "optim_kernels.cuh"
#include "common.cuh"
#include "gemm_simt.cuh"
#include "tc_gemm.cuh"
#include "level_kernels.cuh"

namespace cliora {
thread_local char g_last_cuda_error[256] = "";
std::atomic<long long> g_launch_count{0};
Profiler g_prof;
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
int g_pdl = env_int("CLIORA_PDL", 0);
int g_splitk_target = 4 * 148;
namespace tc { int g_tc_xnarrow = env_int("CLIORA_TC_XNARROW", 0); int g_tc_small_tmem = 0; int g_tc_narrow_stages = kTcNarrowStages; }
// One shared-memory carveout for every kernel of the library: the tcgen05 GEMMs need the maximum carveout, and an
// SM has to drain before it can switch configuration, so mixed carveouts serialise neighbouring kernels.
int g_carveout = env_int("CLIORA_CARVEOUT", 100);
namespace {
struct AttrKey {
  int dev;
  const void* kern;
  int attr;
  bool operator<(const AttrKey& o) const {
    if (dev != o.dev) return dev < o.dev;
    if (kern != o.kern) return kern < o.kern;
    return attr < o.attr;
  }
};
std::mutex g_attr_mu;
std::map<AttrKey, int> g_attr_done;
}  // namespace
cudaError_t func_attr_at_least(const void* kern, cudaFuncAttribute attr, int value) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_attr_mu);
  auto it = g_attr_done.find(AttrKey{dev, kern, (int)attr});
  if (it != g_attr_done.end() && it->second >= value) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, attr, value);
  if (e == cudaSuccess) g_attr_done[AttrKey{dev, kern, (int)attr}] = value;
  return e;
}
void apply_carveout(const void* kern) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_attr_mu);
  auto key = AttrKey{dev, kern, (int)cudaFuncAttributePreferredSharedMemoryCarveout};
  auto it = g_attr_done.find(key);
  if (it != g_attr_done.end() && it->second == g_carveout) return;
  g_attr_done[key] = g_carveout;
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, g_carveout);
}
namespace lvl {
long long* g_level_dbg = nullptr;
int max_active_clusters(int nc, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<int, std::pair<int, size_t>>, int> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(dev, std::make_pair(nc, smem));
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  const int n = max_active_clusters_query(nc, smem);
  cache[key] = n;
  return n;
}
}  // namespace lvl
int g_debug[16] = {2, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [6] = 1: unfused per-level forward (split_build + GEMM + cell_aggregate), [7] = 1: unfused backward
// [0] = tc accumulate mode, [1] = 1: force the SIMT GEMMs, [2] = tc tile (0 auto, 1 narrow, 2 wide), [3] = bit0: block-per-cell VL forward kernel, bit1: block-per-cell VL backward kernel, [4] = 1: db2 from the full GY rows instead of the per-cell sums, [5] = 1: per-cell GEMMs on the fp32 SIMT kernel instead of mma.sync 3xTF32

static int validate(const cliora_dims* d) {
  if (d == nullptr) return CLIORA_ERR_NULL_POINTER;
  if (d->B < 1 || d->n < 1 || d->n > 512 || d->D < 4 || (d->D % 4) != 0 || d->R < 0 || d->R > 64)
    return CLIORA_ERR_BAD_SHAPE;
  return CLIORA_OK;
}

static int64_t align4(int64_t x) { return (x + 3) & ~(int64_t)3; }

static int compute_layout(const cliora_dims& d, cliora_layout& L) {
  CL_TRY(validate(&d));
  const int64_t B = d.B, n = d.n, D = d.D, R = d.R, C = num_cells(d.n);
  const int64_t PI = d.share ? 3 : 4;
  memset(&L, 0, sizeof(L));
  L.PI = PI;
  L.rows_in = B * inside_rows_before(d.n, d.n);
  L.rows_out = B * outside_rows_before(d.n, d.n - 1);
  int64_t o = 0;
  auto take = [&](int64_t nf) { int64_t r = o; o += align4(nf > 0 ? nf : 4); return r; };
  L.Pin = take(B * C * PI * D);
  L.Pout = take(B * C * 2 * D);
  L.q_in = R > 0 ? take(B * C * D) : -1;
  L.nrm_in = take(B * C);
  L.nrm2_in = R > 0 ? take(B * C) : -1;
  L.att_in = R > 0 ? take(B * C * R) : -1;
  L.nrm_out = take(B * C);
  L.leaf_t = take(B * n * D);
  L.Zin = take(2 * L.rows_in * D);   // split pairs: hi part, then lo part at + rows * D
  L.Yin = take(2 * L.rows_in * D);
  L.Ein = take(L.rows_in);
  L.Prin = take(L.rows_in);
  L.Zout = take(2 * L.rows_out * D);
  L.Yout = take(2 * L.rows_out * D);
  L.Eout = take(L.rows_out);
  L.Prout = take(L.rows_out);
  L.Wcat_in = take(PI * D * D);
  L.Wcat_out = take(2 * D * D);
  L.W2p = take(2 * D * D);
  L.W2Tp = take(2 * D * D);
  L.oW2p = d.share ? L.W2p : take(2 * D * D);
  L.oW2Tp = d.share ? L.W2Tp : take(2 * D * D);
  L.Mbin = D <= 512 ? take(L.rows_in * 16) : -1;    // ReLU bitmasks of Z (16 x uint32 per split row)
  L.Mbout = D <= 512 ? take(L.rows_out * 16) : -1;
  L.W2h = take(D * D / 2 + 4);
  L.W2Th = take(D * D / 2 + 4);
  L.oW2h = d.share ? L.W2h : take(D * D / 2 + 4);
  L.oW2Th = d.share ? L.W2Th : take(D * D / 2 + 4);
  L.ws_floats = o;

  const int64_t max_rows = B * n * (n - 1) > 0 ? B * n * (n - 1) : 4;
  o = 0;
  L.Gh_in = take(B * C * D);
  L.Gs_in = take(B * C);
  L.GP_in = take(B * C * PI * D);
  L.Gh_out = take(B * C * D);
  L.Gs_out = take(B * C);
  L.GP_out = take(B * C * 2 * D);
  L.GA2 = R > 0 ? take(B * C * D) : -1;
  L.coef = R > 0 ? take(B * C * 2 * R) : -1;
  L.GE = take(max_rows);
  L.GZ = take(max_rows * D);
  int64_t sk = tn_scratch_floats((int)(B * C), (int)D, (int)D);
  int64_t t;
  if ((t = tn_scratch_floats((int)L.rows_in, (int)D, (int)D)) > sk) sk = t;
  if ((t = tn_scratch_floats((int)L.rows_out, (int)D, (int)D)) > sk) sk = t;
  if ((t = tn_scratch_floats((int)(B * n), (int)D, (int)D)) > sk) sk = t;
  if ((t = 64 * PI * D) > sk) sk = t;
  if ((t = tc::tn_tc_scratch_floats((int)L.rows_in, (int)D, (int)D)) > sk) sk = t;
  if ((t = tc::tn_tc_scratch_floats((int)L.rows_out, (int)D, (int)D)) > sk) sk = t;
  if ((t = tc::tn_tc_scratch_floats((int)(B * C), (int)D, (int)D)) > sk) sk = t;
  L.splitk = take(sk);
  L.gu = take(B * n * D);
  L.GPp = take(2 * B * C * PI * D);   // split pairs of the projection-gradient accumulators (tensor-core wgrad)
  L.Hp = take(2 * B * C * D);         // split pair of the chart vectors
  L.CSin = take(B * C * D);           // per-cell sums of the GY rows (zero for cells without splits)
  L.CSout = take(B * C * D);
  L.GYp_in = take(2 * L.rows_in * D);
  L.GYp_out = take(2 * L.rows_out * D);
  L.GA = take(B * C * D);
  L.CM = take(B * C);
  L.db2acc = take(2 * D);
  L.bws_floats = o;
  return CLIORA_OK;
}

struct Ctx {
  cliora_dims d;
  cliora_layout L;
  int64_t C;
  cudaStream_t st;
  bool use_tc;   // compose GEMMs on tcgen05 (3xTF32 split pairs) instead of the SIMT fp32 kernel
  int tc_mode;   // UMMA accumulation scheme: 2 = fp32-accurate 3xTF32 (default), 1 = single-pass TF32 (CLIORA_FLAG_TF32_1PASS)
  int lvl_mode;  // mode of the fused level kernels: tc_mode, or 3 = bf16 operands (CLIORA_FLAG_BF16)
};

static int make_ctx(const cliora_dims* dims, cliora_stream_t stream, Ctx& c) {
  CL_TRY(validate(dims));
  c.d = *dims;
  CL_TRY(compute_layout(c.d, c.L));
  c.C = num_cells(dims->n);
  c.st = (cudaStream_t)stream;
  c.use_tc = (dims->D >= 32) && (g_debug[1] == 0);
  c.tc_mode = (dims->flags & (CLIORA_FLAG_TF32_1PASS | CLIORA_FLAG_BF16)) ? 1 : g_debug[0];
  c.lvl_mode = (dims->flags & CLIORA_FLAG_BF16) ? 3 : (c.tc_mode == 1 ? 1 : 2);
  return CLIORA_OK;
}

// C[rows of level] = act(A[rows of level] W^T + bias): chart-level projection
static int project_level(const Ctx& c, int level, const float* chart_h, const float* Wcat, int ncols, float* P) {
  GemmParams p{};
  p.A = chart_h; p.lda = c.d.D; p.amap = level_rows(c.d.n, level);
  p.W = Wcat; p.ldw = c.d.D;
  p.C = P; p.ldc = ncols; p.cmap = level_rows(c.d.n, level);
  p.M = c.d.B * (c.d.n - level); p.N = ncols; p.K = c.d.D;
  p.tag = "gemm_cell_project";
  p.mma_ok = c.use_tc && g_debug[5] == 0;
  p.accumulate = 1;   // P was zero-filled at the start of the pass: lets small levels split K with red.add
  return launch_gemm(c.st, /*nt=*/true, p, /*atomic_ok=*/true);
}

// Gh[rows of level] += GP[rows of level] @ Wcat
static int cellgrad_level(const Ctx& c, int level, const float* GP, int ncols, const float* Wcat, float* Gh) {
  GemmParams p{};
  p.A = GP; p.lda = ncols; p.amap = level_rows(c.d.n, level);
  p.W = Wcat; p.ldw = c.d.D;
  p.C = Gh; p.ldc = c.d.D; p.cmap = level_rows(c.d.n, level);
  p.M = c.d.B * (c.d.n - level); p.N = c.d.D; p.K = ncols;
  p.accumulate = 1;
  p.mma_ok = c.use_tc && g_debug[5] == 0;
  p.tag = "gemm_cell_grad";
  return launch_gemm(c.st, /*nt=*/false, p, /*atomic_ok=*/true);
}

static int dense_linear(cudaStream_t st, int M, int N, int K, const float* A, const float* W, const float* bias,
                        int act, float* Cout, const char* tag = "gemm_linear") {
  GemmParams p{};
  p.A = A; p.lda = K; p.amap = dense_rows();
  p.W = W; p.ldw = K;
  p.C = Cout; p.ldc = N; p.cmap = dense_rows();
  p.bias = bias; p.act = act;
  p.M = M; p.N = N; p.K = K;
  p.tag = tag;
  if (act == 0 && K >= 512 && (int64_t)ceil_div(M, 64) * ceil_div(N, kBN) < 148) {
    CL_CUDA(cudaMemsetAsync(Cout, 0, (size_t)M * N * sizeof(float), st));
    p.accumulate = 1;
    return launch_gemm(st, true, p, /*atomic_ok=*/true);
  }
  return launch_gemm(st, true, p);
}

static int colsum(cudaStream_t st, const float* src, int64_t ld, int64_t rows, int cols, float* dst, int accumulate,
                  float* scratch) {
  int S = (int)((rows + 511) / 512);
  if (S < 1) S = 1;
  if (S > 64) S = 64;
  dim3 grid(ceil_div(cols, 32), S);
  launch_k(colsum_stage1_kernel, grid, dim3(32, 8), 0, st, src, ld, rows, cols, scratch);
  CL_CHECK_LAUNCH("colsum_stage1_kernel");
  launch_k(colsum_stage2_kernel, ceil_div(cols, 128), 128, 0, st, scratch, S, cols, dst, accumulate);
  CL_CHECK_LAUNCH("colsum_stage2_kernel");
  return CLIORA_OK;
}

static CellArgs cell_args(const Ctx& c, int level, bool outside, float* ws, float* chart_h, float* chart_s) {
  CellArgs a{};
  const int n = c.d.n;
  a.B = c.d.B; a.n = n; a.level = level; a.D = c.d.D; a.R = outside ? 0 : c.d.R; a.C = c.C;
  a.L = n - level;
  a.no_norm = (c.d.flags & CLIORA_FLAG_NO_NORMALIZE) ? 1 : 0;
  a.chart_h = chart_h; a.chart_s = chart_s;
  if (!outside) {
    a.N = level == 0 ? 1 : level;
    a.sp = a.N; a.sk = 1;
    if (level == 0) {
      a.Y = ws + c.L.leaf_t; a.E = nullptr; a.Pr = nullptr;
    } else {
      const int64_t r0 = c.d.B * inside_rows_before(n, level);
      a.Y = ws + c.L.Yin + r0 * c.d.D; a.E = ws + c.L.Ein + r0; a.Pr = ws + c.L.Prin + r0;
      a.y_lo_off = c.use_tc ? c.L.rows_in * c.d.D : 0;
    }
    a.q = c.d.R > 0 ? ws + c.L.q_in : nullptr;
    a.nrm = ws + c.L.nrm_in;
    a.nrm2 = c.d.R > 0 ? ws + c.L.nrm2_in : nullptr;
    a.att = c.d.R > 0 ? ws + c.L.att_in : nullptr;
  } else {
    a.N = n - level - 1;
    a.sp = 1; a.sk = a.L;
    const int64_t r0 = c.d.B * outside_rows_before(n, level);
    a.Y = ws + c.L.Yout + r0 * c.d.D; a.E = ws + c.L.Eout + r0; a.Pr = ws + c.L.Prout + r0;
    a.y_lo_off = c.use_tc ? c.L.rows_out * c.d.D : 0;
    a.q = nullptr;
    a.nrm = ws + c.L.nrm_out;
  }
  return a;
}

static SplitArgs split_args(const Ctx& c, int level, bool outside, const float* ih, const float* is_,
                            const float* os_, float* ws, const float* b1) {
  SplitArgs s{};
  const int n = c.d.n;
  s.B = c.d.B; s.n = n; s.level = level; s.D = c.d.D; s.C = c.C;
  s.L = n - level;
  s.N = outside ? n - level - 1 : level;
  s.ih = ih; s.is_ = is_; s.os_ = os_;
  s.Pin = ws + c.L.Pin; s.Pout = ws + c.L.Pout;
  s.ldPin = (int)(c.L.PI * c.d.D);
  s.iAl = (outside && !c.d.share) ? 3 * c.d.D : 0;
  s.b1 = b1;
  const int64_t r0 = outside ? c.d.B * outside_rows_before(n, level) : c.d.B * inside_rows_before(n, level);
  s.Z = ws + (outside ? c.L.Zout : c.L.Zin) + r0 * c.d.D;
  s.E = ws + (outside ? c.L.Eout : c.L.Ein) + r0;
  s.z_lo_off = c.use_tc ? (outside ? c.L.rows_out : c.L.rows_in) * c.d.D : 0;
  const int64_t mb = outside ? c.L.Mbout : c.L.Mbin;
  s.zmask = (c.use_tc && mb >= 0) ? reinterpret_cast<uint32_t*>(ws + mb) + r0 * 16 : nullptr;
  return s;
}

// Y[level rows] = relu(Z[level rows] W2^T + b2): the dense contraction of the compose MLP (diora.py:65-72)
static int compose_gemm(const Ctx& c, bool outside, int64_t r0, int64_t rows, const float* W2, const float* b2,
                        float* ws) {
  const int D = c.d.D;
  const int64_t total = outside ? c.L.rows_out : c.L.rows_in;
  float* Zb = ws + (outside ? c.L.Zout : c.L.Zin);
  float* Yb = ws + (outside ? c.L.Yout : c.L.Yin);
  if (c.use_tc) {
    tc::PairRef A{Zb, total, D, total * D};
    tc::PairRef W{ws + (outside ? c.L.oW2p : c.L.W2p), D, D, (int64_t)D * D};
    tc::TcEpilogue ep{};
    ep.C = Yb + r0 * D; ep.ldc = D; ep.cmap = dense_rows();
    ep.bias = b2; ep.act = 1;
    return tc::launch_tc_gemm_nt(c.st, A, (int)r0, W, (int)rows, D, D, ep, "tc_gemm_compose_w2", c.tc_mode, g_debug[2]);
  }
  return dense_linear(c.st, (int)rows, D, D, Zb + r0 * D, W2, b2, 1, Yb + r0 * D, "gemm_compose_w2");
}

// GZ[level rows] = (GY[level rows] W2) * (Z > 0)
static bool fused_level_ok(const Ctx& c, int N, lvl::LevelGeom& g);
static int compose_gemm_bwd(const Ctx& c, bool outside, int level, int64_t r0, int64_t rows, const float* W2, float* ws,
                            float* GZ) {
  const int D = c.d.D;
  const int64_t total = outside ? c.L.rows_out : c.L.rows_in;
  float* Zb = ws + (outside ? c.L.Zout : c.L.Zin);
  float* Yb = ws + (outside ? c.L.Yout : c.L.Yin);
  if (c.use_tc) {
    tc::PairRef A{Yb, total, D, total * D};
    tc::PairRef W{ws + (outside ? c.L.oW2Tp : c.L.W2Tp), D, D, (int64_t)D * D};
    tc::TcEpilogue ep{};
    ep.C = GZ; ep.ldc = D; ep.cmap = dense_rows();
    ep.mask = Zb + r0 * D; ep.ldm = D; ep.mask_lo_off = total * D;
    const int64_t mb = outside ? c.L.Mbout : c.L.Mbin;
    lvl::LevelGeom geom;
    const int nsplit = outside ? c.d.n - level - 1 : level;
    const bool bits_written = !fused_level_ok(c, nsplit, geom) || g_debug[10] != 0;   // the fused forward skips them
    if (mb >= 0 && bits_written) ep.maskbits = reinterpret_cast<const uint32_t*>(ws + mb) + r0 * 16;
    return tc::launch_tc_gemm_nt(c.st, A, (int)r0, W, (int)rows, D, D, ep, "tc_gemm_compose_w2_bwd", c.tc_mode, g_debug[2]);
  }
  GemmParams p{};
  p.A = Yb + r0 * D; p.lda = D; p.amap = dense_rows();
  p.W = W2; p.ldw = D;
  p.C = GZ; p.ldc = D; p.cmap = dense_rows();
  p.mask = Zb + r0 * D; p.ldm = D;
  p.M = (int)rows; p.N = D; p.K = D;
  p.tag = "gemm_compose_w2_bwd";
  return launch_gemm(c.st, /*nt=*/false, p);
}

// W [D, D] fp32 -> bf16 copy and bf16 transposed copy (bf16 mode of the fused level kernels)
__global__ void to_bf16_kernel(const float* __restrict__ W, int D, __nv_bfloat16* __restrict__ out,
                               __nv_bfloat16* __restrict__ outT) {
  pdl_prologue();
  const int64_t n = (int64_t)D * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), cidx = (int)(i % D);
    const __nv_bfloat16 v = __float2bfloat16_rn(W[i]);
    out[i] = v;
    outT[(int64_t)cidx * D + r] = v;
  }
}

static int prepare_w2_bf16(const Ctx& c, const float* W2, float* h, float* hT) {
  launch_k(to_bf16_kernel, 296, 256, 0, c.st, W2, c.d.D, reinterpret_cast<__nv_bfloat16*>(h),
           reinterpret_cast<__nv_bfloat16*>(hT));
  CL_CHECK_LAUNCH("to_bf16_kernel");
  return CLIORA_OK;
}

static int prepare_w2_pairs(const Ctx& c, const float* W2, float* pair, float* pairT) {
  const int D = c.d.D;
  const int64_t n = (int64_t)D * D;
  launch_k(tc::split_tf32_kernel, ceil_div(n, 256), 256, 0, c.st, W2, n, pair);
  CL_CHECK_LAUNCH("split_tf32_kernel");
  dim3 grid(ceil_div(D, 32), ceil_div(D, 32));
  launch_k(tc::split_tf32_transpose_kernel, grid, dim3(32, 8), 0, c.st, W2, D, D, D, pairT);
  CL_CHECK_LAUNCH("split_tf32_transpose_kernel");
  return CLIORA_OK;
}

// dW_k[D, D] (+)= GP[:, k*D:(k+1)*D]^T H  for the nblk column blocks of a projection-gradient accumulator GP [B*C, nblk*D]
// (dst[k], ldc[k], acc[k] say where block k goes).  Tensor-core path: GP and H are first re-written as split pairs.
static int cell_wgrad(const Ctx& c, const float* GP, int nblk, const float* H, float* const* dst, const int64_t* ldc,
                      const int* acc, float* bws) {
  const int D = c.d.D;
  const int64_t BC = (int64_t)c.d.B * c.C;
  float* scratch = bws + c.L.splitk;
  if (c.use_tc) {
    float* GPp = bws + c.L.GPp;
    float* Hp = bws + c.L.Hp;
    const int64_t n4 = BC * nblk * D / 4;
    launch_k(tc::split_tf32_rows_kernel, ceil_div(n4, 256) < 2368 ? ceil_div(n4, 256) : 2368, 256, 0, c.st, GP, BC, nblk * D, nblk * D, GPp);
    CL_CHECK_LAUNCH("split_tf32_rows_kernel");
    launch_k(tc::split_tf32_rows_kernel, ceil_div(BC * D / 4, 256) < 2368 ? ceil_div(BC * D / 4, 256) : 2368, 256, 0, c.st, H, BC, D, D, Hp);
    CL_CHECK_LAUNCH("split_tf32_rows_kernel");
    tc::PairRef Bp{Hp, BC, D, BC * D};
    for (int k = 0; k < nblk; ++k) {
      if (!dst[k]) continue;
      tc::PairRef Ap{GPp + (int64_t)k * D, BC, (int64_t)nblk * D, BC * nblk * D};
      CL_TRY(tc::launch_tc_gemm_tn(c.st, Ap, Bp, (int)BC, D, D, dst[k], ldc[k], acc[k], scratch, "tc_gemm_wgrad_cell", c.tc_mode));
    }
    return CLIORA_OK;
  }
  for (int k = 0; k < nblk; ++k) {
    if (!dst[k]) continue;
    CL_TRY(launch_gemm_tn(c.st, (int)BC, D, D, GP + (int64_t)k * D, (int64_t)nblk * D, H, D, dst[k], ldc[k], acc[k], scratch));
  }
  return CLIORA_OK;
}

template <bool VL>
static int launch_cell_aggregate(const Ctx& c, const CellArgs& a) {
  const double rows = (double)a.B * a.L * a.N;
  ProfScope prof(c.st, "cell_aggregate", 2.0 * rows * a.D, 4.0 * (rows * (a.D + 2) + 2.0 * a.B * a.L * a.D));
  if (VL && a.D <= 128 * kColT && (g_debug[3] & 1) == 0) {
    // warp per cell, 8 cells of one sentence per CTA, the image's regions staged once in shared memory
    const size_t smem = ((size_t)a.R * a.D + (size_t)kCellsPerCta * a.N) * sizeof(float);
    if (smem > 48 * 1024)
      CL_CUDA(func_attr_at_least(reinterpret_cast<const void*>(cell_fwd_warp_kernel<true>),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int chunks = ceil_div(a.L, kCellsPerCta);
    launch_k(cell_fwd_warp_kernel<true>, a.B * chunks, kCellThreads, smem, c.st, a);
    CL_CHECK_LAUNCH("cell_fwd_warp_kernel");
    return CLIORA_OK;
  }
  const size_t smem = (size_t)(a.D + a.N + 2 * a.R + 64) * sizeof(float);
  launch_k(cell_aggregate_kernel<VL>, a.B * a.L, 128, smem, c.st, a);
  CL_CHECK_LAUNCH("cell_aggregate_kernel");
  return CLIORA_OK;
}

// db2 = column sum of the GY rows.  The block-per-cell backward kernel leaves per-cell sums of those rows in
// CSin / CSout, so the column sum runs over B*C rows instead of all split rows (19x fewer at n = 20).
static bool cellsum_enabled(const Ctx& c, bool outside) {
  if (c.d.D > 512 || g_debug[4] != 0) return false;
  const bool warp_bwd = !outside && c.d.R > 0 && c.d.D <= 128 * kColT && (g_debug[3] & 2) == 0;
  return !warp_bwd;
}

template <bool VL>
static int launch_cell_bwd(const Ctx& c, const CellBwdArgs& g, float* cellsum) {
  const double rows = (double)g.c.B * g.c.L * g.c.N;
  ProfScope prof(c.st, "cell_bwd", 4.0 * rows * g.c.D, 4.0 * (2.0 * rows * (g.c.D + 2) + 3.0 * g.c.B * g.c.L * g.c.D));
  if (VL && g.c.D <= 128 * kColT && (g_debug[3] & 2) == 0) {
    const size_t smem = ((size_t)g.c.R * g.c.D + (size_t)kCellsPerCta * g.c.D + 32) * sizeof(float);
    if (smem > 48 * 1024)
      CL_CUDA(func_attr_at_least(reinterpret_cast<const void*>(cell_bwd_warp_kernel<true>),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int chunks = ceil_div(g.c.L, kCellsPerCta);
    launch_k(cell_bwd_warp_kernel<true>, g.c.B * chunks, kCellThreads, smem, c.st, g);
    CL_CHECK_LAUNCH("cell_bwd_warp_kernel");
    return CLIORA_OK;
  }
  CellBwdArgs gb = g;
  if (g.c.D <= 512 && g.c.E != nullptr && g_debug[4] == 0 && g.ga_out == nullptr) gb.cellsum = cellsum;
  const size_t smem = (size_t)(2 * g.c.D + 3 * g.c.R + 64 + (gb.cellsum ? 8 * g.c.D + 4 : 0)) * sizeof(float);
  launch_k(cell_bwd_kernel<VL>, g.c.B * g.c.L, 256, smem, c.st, gb);
  CL_CHECK_LAUNCH("cell_bwd_kernel");
  return CLIORA_OK;
}

// Concurrent sentence chains the caller runs (bits 8-11 of cliora_dims.flags; 0 = 1): a hint for tile sizing only
static int chain_count(const Ctx& c) {
  const int k = (c.d.flags >> 8) & 15;
  return k < 1 ? 1 : k;
}

// One launch for a whole forward level (gather + compose GEMM + softmax-weighted sums + cell finalize): lvl::level_fwd_kernel.
// Returns false when the shape is outside what the fused kernel covers (the unfused chain then runs).
static bool fused_level_ok(const Ctx& c, int N, lvl::LevelGeom& g) {
  const bool unfused = (c.d.flags & CLIORA_FLAG_UNFUSED) && !(c.d.flags & CLIORA_FLAG_BF16);   // bf16 lives in the fused kernels
  if (!c.use_tc || g_debug[6] != 0 || unfused || N < 1 || N > lvl::kRows) return false;
  if (!lvl::level_geom(c.d.D, g)) return false;
  if (c.d.D > 1024) return false;
  return true;
}

// A level whose tiles (whole cells, at most 128 split rows each) cannot all be resident at once with the narrow
// geometry anyway is run with wide column slices instead: half as many CTAs per tile, so half the waves.
// `slots`: CTAs of the narrow kernel that can be resident for this chain.
static void widen_if_crowded(const Ctx& c, int cells, int N, int slots, lvl::LevelGeom& g) {
  if (g_debug[15] == 1 || (g_debug[13] != 0) || g_debug[10] != 0) return;      // 15 = 1: never wide
  const int gmax = lvl::kRows / N;
  const int min_tiles = ceil_div(cells, gmax < 1 ? 1 : gmax);
  if (g_debug[15] != 2 && min_tiles * g.nc <= slots) return;                    // 15 = 2: always wide
  lvl::LevelGeom w;
  if (lvl::level_geom(c.d.D, w, /*wide=*/true) && w.nc < g.nc) g = w;
}

// The fused backward needs every level of a pass to qualify (its compose-output gradients live in one buffer per pass)
static bool fused_bwd_ok(const Ctx& c) {
  lvl::LevelGeom g;
  return g_debug[7] == 0 && c.d.n >= 2 && fused_level_ok(c, c.d.n - 1, g);
}

// Tile plan of one level (shared by the launches and by cliora_level_plan_query): column slices (narrow, or wide when
// the level cannot be resident in one wave anyway), cells per tile, sentences a tile may span.
static bool plan_level_fwd(const Ctx& c, int level, bool outside, lvl::LevelGeom& geom, int& G, int& max_sent) {
  const int n = c.d.n, L = n - level, N = outside ? n - level - 1 : level, R = outside ? 0 : c.d.R;
  const int cells = c.d.B * L;
  lvl::LevelGeom narrow;
  if (!fused_level_ok(c, N, narrow)) return false;
  geom = narrow;
  // sentence chains run their level kernels side by side: each aims at its share of the co-resident clusters
  const int slots = lvl::max_active_clusters(geom.nc, lvl::level_fwd_smem(geom.n_umma)) * geom.nc / chain_count(c);
  widen_if_crowded(c, cells, N, slots, geom);
  G = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    G = lvl::level_cells_per_tile(cells, N, L, R, geom,
                                  lvl::max_active_clusters(geom.nc, lvl::level_fwd_smem(geom.n_umma)) / chain_count(c),
                                  max_sent);
    if (G >= 1 || geom.nc == narrow.nc) break;
    geom = narrow;                                   // the wide tile does not fit this level's staging: narrow again
  }
  return G >= 1;
}
static bool plan_level_bwd(const Ctx& c, int level, bool outside, lvl::LevelGeom& geom, int& G, int& max_sent) {
  const int n = c.d.n, L = n - level, N = outside ? n - level - 1 : level;
  if (N < 1 || !fused_bwd_ok(c) || !fused_level_ok(c, N, geom)) return false;
  widen_if_crowded(c, c.d.B * L, N, 148 / chain_count(c), geom);
  int ms = 0;
  G = lvl::level_cells_per_tile(c.d.B * L, N, L, 0, geom, (148 / geom.nc) / chain_count(c), ms);
  max_sent = G > 0 ? (G - 1) / L + 2 : 0;
  return G >= 1;
}

static int fused_level_fwd(const Ctx& c, int level, bool outside, const lvl::LevelGeom& geom_in, const cliora_weights* w,
                           const float* ih, const float* is_, const float* os_, float* chart_h, float* chart_s,
                           const float* obj, const uint8_t* keep, float* ws) {
  const int n = c.d.n, D = c.d.D, B = c.d.B;
  const bool sh = c.d.share != 0;
  lvl::LevelFwdArgs a{};
  a.B = B; a.n = n; a.level = level; a.L = n - level; a.N = outside ? n - level - 1 : level; a.D = D;
  a.R = outside ? 0 : c.d.R;
  a.cells = B * a.L;
  lvl::LevelGeom geom = geom_in;
  if (!plan_level_fwd(c, level, outside, geom, a.G, a.max_sent)) return CLIORA_ERR_UNSUPPORTED;
  a.nc = geom.nc; a.ncols = geom.ncols; a.n_umma = geom.n_umma;
  a.mode = c.lvl_mode;
  a.store_lo = c.lvl_mode == 2 ? 1 : 0;      // modes 1 and 3 never read the lo parts
  a.single_acc = (a.n_umma > lvl::kNarrowUmmaN) ? 1 : 0;
  a.no_norm = (c.d.flags & CLIORA_FLAG_NO_NORMALIZE) ? 1 : 0;
  a.outside = outside ? 1 : 0;
  a.C = c.C;
  const int ldPin = (int)(c.L.PI * D);
  a.P1 = ws + c.L.Pin; a.ld1 = ldPin; a.off_a1 = (outside && !sh) ? 3 * D : 0;
  if (!outside) { a.P2 = ws + c.L.Pin; a.ld2 = ldPin; a.off_a2 = D; a.off_v2 = 2 * D; }
  else { a.P2 = ws + c.L.Pout; a.ld2 = 2 * D; a.off_a2 = 0; a.off_v2 = D; }
  a.h1 = ih; a.s1 = is_; a.s2 = outside ? os_ : is_;
  a.b1 = (outside && !sh) ? w->ob1 : w->b1;
  a.b2 = (outside && !sh) ? w->ob2 : w->b2;
  const int64_t r0 = outside ? B * outside_rows_before(n, level) : B * inside_rows_before(n, level);
  const int64_t total = outside ? c.L.rows_out : c.L.rows_in;
  a.Z = ws + (outside ? c.L.Zout : c.L.Zin) + r0 * D;
  a.z_lo_off = total * D;
  const int64_t mb = outside ? c.L.Mbout : c.L.Mbin;
  // ReLU bit masks are only produced on request: the backward reads the sign of the stored Z pair instead
  a.zmask = (mb >= 0 && g_debug[10] != 0) ? reinterpret_cast<uint32_t*>(ws + mb) + r0 * 16 : nullptr;
  // the tensor-memory variant leaves the signs as plain bits in the same buffer (64 bytes per row, D <= 512)
  a.zbits = (mb >= 0 && g_debug[10] == 0 && g_debug[13] == 0) ? reinterpret_cast<uint16_t*>(ws + mb) + r0 * 32 : nullptr;
  a.Y = ws + (outside ? c.L.Yout : c.L.Yin) + r0 * D;
  a.E = ws + (outside ? c.L.Eout : c.L.Ein) + r0;
  a.Pr = ws + (outside ? c.L.Prout : c.L.Prin) + r0;
  a.chart_h = chart_h; a.chart_s = chart_s;
  a.q = (!outside && c.d.R > 0) ? ws + c.L.q_in : nullptr;
  a.nrm = ws + (outside ? c.L.nrm_out : c.L.nrm_in);
  a.nrm2 = (!outside && c.d.R > 0) ? ws + c.L.nrm2_in : nullptr;
  a.att = (!outside && c.d.R > 0) ? ws + c.L.att_in : nullptr;
  a.obj = obj; a.keep = keep;
  const float* W2pair = c.lvl_mode == 3 ? ws + (outside ? c.L.oW2h : c.L.W2h) : ws + (outside ? c.L.oW2p : c.L.W2p);
  return lvl::launch_level_fwd(c.st, a, W2pair, outside ? "level_fwd_outside" : "level_fwd_inside");
}

// shared by both passes: cell backward, GZ GEMM, scatter for one level
template <bool OUTSIDE, bool VL>
static int level_bwd(const Ctx& c, int level, const cliora_weights* w, const float* ih, const float* is_,
                     const float* os_, float* chart_h, float* chart_s, const float* obj, const uint8_t* keep,
                     float* ws, float* bws) {
  const int B = c.d.B, n = c.d.n, D = c.d.D;
  CellBwdArgs g{};
  g.c = cell_args(c, level, OUTSIDE, ws, chart_h, chart_s);
  g.c.obj = obj; g.c.keep = keep;
  g.Gh = bws + (OUTSIDE ? c.L.Gh_out : c.L.Gh_in);
  g.Gs = bws + (OUTSIDE ? c.L.Gs_out : c.L.Gs_in);
  g.GE = bws + c.L.GE;
  g.GA2 = VL ? bws + c.L.GA2 : nullptr;
  g.coef = VL ? bws + c.L.coef : nullptr;
  g.leaf_t = ws + c.L.leaf_t;
  g.gu = bws + c.L.gu;
  // GE is level-local here: point the kernel at a level block starting at GE[0]
  // (cell_bwd indexes GE with the same row ids as E, relative to the level block).
  const bool fused = g.c.E != nullptr && fused_bwd_ok(c);
  // the fused kernel's prologue does the cell part: text cells always, CLIORA cells when a tile's images fit the rings
  lvl::LevelGeom geom0;
  int G0 = 0, max_sent0 = 0;
  if (fused && !plan_level_bwd(c, level, OUTSIDE, geom0, G0, max_sent0)) return CLIORA_ERR_UNSUPPORTED;
  const bool vl_fits = VL && D <= 128 * kColT && c.d.R <= 64 && (D % 4) == 0 && G0 > 0 &&
                       (size_t)max_sent0 * c.d.R * D * sizeof(float) <= (size_t)lvl::ring_bytes(geom0.n_umma);
  const bool cells_inline = fused && (!VL || vl_fits) && g_debug[14] == 0;
  if (fused) {
    g.ga_out = bws + c.L.GA;
    g.cm_out = bws + c.L.CM;
  }
  if (!cells_inline) CL_TRY(launch_cell_bwd<VL>(c, g, fused ? nullptr : bws + (OUTSIDE ? c.L.CSout : c.L.CSin)));
  if (g.c.E == nullptr) return CLIORA_OK;  // leaf level: no splits
  if (fused) {
    const lvl::LevelGeom geom = geom0;               // narrow, or wide when the level is crowded (see above)
    const bool sh = c.d.share != 0;
    lvl::LevelBwdArgs b{};
    lvl::LevelFwdArgs& a = b.geo;
    a.B = B; a.n = n; a.level = level; a.L = n - level; a.N = g.c.N; a.D = D; a.R = 0;
    a.cells = B * a.L;
    a.nc = geom.nc; a.ncols = geom.ncols; a.n_umma = geom.n_umma;
    a.single_acc = (a.n_umma > lvl::kNarrowUmmaN) ? 1 : 0;
    a.store_lo = c.lvl_mode == 2 ? 1 : 0;
    a.G = G0;
    a.mode = c.lvl_mode;
    a.no_norm = (c.d.flags & CLIORA_FLAG_NO_NORMALIZE) ? 1 : 0;
    a.outside = OUTSIDE ? 1 : 0;
    a.C = c.C;
    const int64_t r0 = OUTSIDE ? B * outside_rows_before(n, level) : B * inside_rows_before(n, level);
    const int ldPin = (int)(c.L.PI * D);
    b.Y = ws + (OUTSIDE ? c.L.Yout : c.L.Yin) + r0 * D;
    b.Zhi = ws + (OUTSIDE ? c.L.Zout : c.L.Zin) + r0 * D;
    {
      const int64_t mb = OUTSIDE ? c.L.Mbout : c.L.Mbin;
      b.zbits = (mb >= 0 && g_debug[10] == 0 && g_debug[13] == 0) ? reinterpret_cast<const uint16_t*>(ws + mb) + r0 * 32 : nullptr;
    }
    b.Pr = ws + (OUTSIDE ? c.L.Prout : c.L.Prin) + r0;
    b.E = ws + (OUTSIDE ? c.L.Eout : c.L.Ein) + r0;
    b.GA = bws + c.L.GA; b.CM = bws + c.L.CM;
    b.Gs = bws + (OUTSIDE ? c.L.Gs_out : c.L.Gs_in);
    b.h1 = ih;
    if (!OUTSIDE) { b.P2 = ws + c.L.Pin; b.ld2 = ldPin; b.off_a2 = D; b.off_v2 = 2 * D; b.GP2 = bws + c.L.GP_in; }
    else { b.P2 = ws + c.L.Pout; b.ld2 = 2 * D; b.off_a2 = 0; b.off_v2 = D; b.GP2 = bws + c.L.GP_out; }
    b.GP1 = bws + c.L.GP_in; b.ld1 = ldPin; b.off_a1 = (OUTSIDE && !sh) ? 3 * D : 0;
    b.Gh1 = bws + c.L.Gh_in; b.Gs1 = bws + c.L.Gs_in;
    b.Gs2 = bws + (OUTSIDE ? c.L.Gs_out : c.L.Gs_in);
    b.GYp = bws + (OUTSIDE ? c.L.GYp_out : c.L.GYp_in) + r0 * D;
    b.gy_lo_off = (OUTSIDE ? c.L.rows_out : c.L.rows_in) * D;
    b.db2 = bws + c.L.db2acc + (OUTSIDE ? D : 0);
    b.GAw = bws + c.L.GA; b.CMw = bws + c.L.CM;
    if (cells_inline) {
      b.cellGh = g.Gh; b.cellH = chart_h; b.cellNrm = g.c.nrm; b.cellS = chart_s;
      if (VL) {
        b.vl_obj = obj; b.vl_keep = keep; b.vl_att = g.c.att; b.vl_q = g.c.q; b.vl_nrm2 = g.c.nrm2;
        b.vl_GA2 = g.GA2; b.vl_coef = g.coef; b.vl_R = c.d.R;
      }
    }
    const float* W2T = c.lvl_mode == 3 ? ws + (OUTSIDE ? c.L.oW2Th : c.L.W2Th) : ws + (OUTSIDE ? c.L.oW2Tp : c.L.W2Tp);
    return lvl::launch_level_bwd(c.st, b, W2T, OUTSIDE ? "level_bwd_outside" : "level_bwd_inside");
  }

  const float* W2 = (OUTSIDE && !c.d.share) ? w->oW2 : w->W2;
  const float* b1 = (OUTSIDE && !c.d.share) ? w->ob1 : w->b1;
  SplitArgs s = split_args(c, level, OUTSIDE, ih, is_, os_, ws, b1);
  const int64_t rows = (int64_t)B * s.L * s.N;
  const int64_t r0 = OUTSIDE ? B * outside_rows_before(n, level) : B * inside_rows_before(n, level);
  CL_TRY(compose_gemm_bwd(c, OUTSIDE, level, r0, rows, W2, ws, bws + c.L.GZ));

  ScatterArgs sc{};
  sc.s = s;
  sc.GZ = bws + c.L.GZ; sc.GE = bws + c.L.GE;
  sc.Gh_in = bws + c.L.Gh_in; sc.Gs_in = bws + c.L.Gs_in; sc.GP_in = bws + c.L.GP_in;
  sc.Gs_out = bws + c.L.Gs_out; sc.GP_out = bws + c.L.GP_out;
  {
    ProfScope prof(c.st, "split_scatter", 2.0 * rows * D, 4.0 * rows * (7.0 * D + 3));
    launch_k(split_scatter_kernel<OUTSIDE>, ceil_div(rows, 8), 256, 0, c.st, sc);
  }
  CL_CHECK_LAUNCH("split_scatter_kernel");
  (void)n;
  return CLIORA_OK;
}

}  // namespace cliora

using namespace cliora;

extern "C" {

const char* cliora_status_string(int s) {
  switch (s) {
    case CLIORA_OK: return "ok";
    case CLIORA_ERR_BAD_SHAPE: return "bad shape (need B>=1, 1<=n<=512, D%4==0, 0<=R<=64)";
    case CLIORA_ERR_NULL_POINTER: return "required pointer is NULL";
    case CLIORA_ERR_CUDA: return "CUDA error";
    case CLIORA_ERR_NO_DEVICE: return "no CUDA device";
    case CLIORA_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}
const char* cliora_last_cuda_error(void) { return g_last_cuda_error; }
int cliora_abi_version(void) { return 1; }
int64_t cliora_launch_count(void) { return g_launch_count; }

void cliora_profile_start(void) {
  for (auto& e : g_prof.entries) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  g_prof.entries.clear();
  g_prof.on = true;
}

int cliora_profile_stop(cliora_profile_row* rows, int max_rows) {
  g_prof.on = false;
  cudaDeviceSynchronize();
  int nrows = 0;
  for (auto& e : g_prof.entries) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e.a, e.b) != cudaSuccess) ms = 0.f;
    int r = -1;
    for (int i = 0; i < nrows; ++i)
      if (strncmp(rows[i].name, e.name, sizeof(rows[i].name)) == 0) { r = i; break; }
    if (r < 0) {
      if (nrows >= max_rows) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); continue; }
      r = nrows++;
      memset(&rows[r], 0, sizeof(rows[r]));
      strncpy(rows[r].name, e.name, sizeof(rows[r].name) - 1);
    }
    rows[r].launches += 1;
    rows[r].ms += ms;
    rows[r].flops += e.flops;
    rows[r].bytes += e.bytes;
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  g_prof.entries.clear();
  cudaGetLastError();
  return nrows;
}

int cliora_level_plan_query(const cliora_dims* dims, int level, int outside, int backward, cliora_level_plan* plan) {
  if (plan == nullptr) return CLIORA_ERR_NULL_POINTER;
  Ctx c;
  CL_TRY(make_ctx(dims, nullptr, c));
  const int n = c.d.n;
  if (level < (outside ? 0 : 1) || level > (outside ? n - 2 : n - 1)) return CLIORA_ERR_BAD_SHAPE;
  memset(plan, 0, sizeof(*plan));
  plan->splits = outside ? n - level - 1 : level;
  plan->cells = c.d.B * (n - level);
  lvl::LevelGeom geom{};
  int G = 0, max_sent = 0;
  const bool ok = backward ? plan_level_bwd(c, level, outside != 0, geom, G, max_sent)
                           : plan_level_fwd(c, level, outside != 0, geom, G, max_sent);
  if (!ok) return CLIORA_OK;          // plan->fused == 0: this level runs the unfused kernel chain
  plan->fused = 1;
  plan->column_slices = geom.nc;
  plan->slice_cols = geom.ncols;
  plan->umma_n = geom.n_umma;
  plan->cells_per_tile = G;
  plan->tiles = ceil_div(plan->cells, G);
  plan->max_sentences_per_tile = max_sent;
  plan->ring_bytes = lvl::ring_bytes(geom.n_umma);
  plan->smem_bytes = (int64_t)(backward ? lvl::level_bwd_smem(geom.n_umma) : lvl::level_fwd_smem(geom.n_umma));
  return CLIORA_OK;
}

int64_t cliora_num_cells(int n) { return num_cells(n); }
int64_t cliora_level_offset(int n, int level) { return lvl_off(n, level); }

int cliora_inside_index(int n, int level, int64_t* left, int64_t* right) {
  if (left == nullptr || right == nullptr) return CLIORA_ERR_NULL_POINTER;
  if (n < 1 || level < 1 || level >= n) return CLIORA_ERR_BAD_SHAPE;
  int64_t i = 0;
  for (int p = 0; p < n - level; ++p)
    for (int k = 0; k < level; ++k, ++i) {
      int l, r;
      inside_children(n, level, p, k, l, r);
      left[i] = l;
      right[i] = r;
    }
  return CLIORA_OK;
}

int cliora_outside_index(int n, int level, int64_t* parent, int64_t* sibling) {
  if (parent == nullptr || sibling == nullptr) return CLIORA_ERR_NULL_POINTER;
  if (n < 1 || level < 0 || level >= n - 1) return CLIORA_ERR_BAD_SHAPE;
  const int L = n - level, N = L - 1;
  for (int k = 0; k < N; ++k)
    for (int p = 0; p < L; ++p) {
      int pa, si;
      outside_parent_sibling(n, level, p, k, pa, si);
      parent[(int64_t)k * L + p] = pa;
      sibling[(int64_t)k * L + p] = si;
    }
  return CLIORA_OK;
}

int64_t cliora_split_row_offset(int B, int n, int level, int outside) {
  return (int64_t)B * (outside ? outside_rows_before(n, level) : inside_rows_before(n, level));
}

int cliora_chart_layout(const cliora_dims* dims, cliora_layout* out) {
  if (dims == nullptr || out == nullptr) return CLIORA_ERR_NULL_POINTER;
  return compute_layout(*dims, *out);
}

// ---------------------------------------------------------------------------------------------
int cliora_inside_fwd(const cliora_dims* dims, const cliora_weights* w, const float* x, const float* obj,
                      const uint8_t* keep, float* inside_h, float* inside_s, float* ws, cliora_stream_t stream) {
  Ctx c;
  CL_TRY(make_ctx(dims, stream, c));
  if (!w || !x || !inside_h || !inside_s || !ws) return CLIORA_ERR_NULL_POINTER;
  if (c.d.R > 0 && obj == nullptr) return CLIORA_ERR_NULL_POINTER;
  if (!c.d.share && (!w->oW1 || !w->ob1 || !w->oW2 || !w->ob2 || !w->oWb)) return CLIORA_ERR_NULL_POINTER;
  const int B = c.d.B, n = c.d.n, D = c.d.D, PI = (int)c.L.PI;
  const bool vl = c.d.R > 0;
  const float* oW1 = c.d.share ? w->W1 : w->oW1;
  const float* oWb = c.d.share ? w->Wb : w->oWb;
  float* Wcat_in = ws + c.L.Wcat_in;

  // Pin rows start as [0 | b1 | 0 (| 0)]: the bias of the compose MLP's first layer rides on the Ar projection
  launch_k(init_proj_kernel, 296, 256, 0, c.st, ws + c.L.Pin, (int64_t)B * c.C, PI * D, D, D, w->b1);
  CL_CHECK_LAUNCH("init_proj_kernel");
  launch_k(pack_weights_kernel, 296, 256, 0, c.st, D, PI, w->W1, w->Wb, oW1, oWb, Wcat_in, ws + c.L.Wcat_out);
  CL_CHECK_LAUNCH("pack_weights_kernel");
  if (c.use_tc) {
    CL_TRY(prepare_w2_pairs(c, w->W2, ws + c.L.W2p, ws + c.L.W2Tp));
    if (!c.d.share) CL_TRY(prepare_w2_pairs(c, w->oW2, ws + c.L.oW2p, ws + c.L.oW2Tp));
    if (c.lvl_mode == 3) {
      CL_TRY(prepare_w2_bf16(c, w->W2, ws + c.L.W2h, ws + c.L.W2Th));
      if (!c.d.share) CL_TRY(prepare_w2_bf16(c, w->oW2, ws + c.L.oW2h, ws + c.L.oW2Th));
    }
  }

  // leaves: t = tanh(W_leaf x + b); h = finalize(t)
  CL_TRY(dense_linear(c.st, B * n, D, D, x, w->W_leaf, w->b_leaf, 2, ws + c.L.leaf_t));
  for (int level = 0; level < n; ++level) {
    lvl::LevelGeom geom;
    if (level > 0 && fused_level_ok(c, level, geom)) {
      CL_TRY(fused_level_fwd(c, level, false, geom, w, inside_h, inside_s, nullptr, inside_h, inside_s, obj, keep, ws));
      if (level < n - 1) CL_TRY(project_level(c, level, inside_h, Wcat_in, PI * D, ws + c.L.Pin));
      continue;
    }
    if (level > 0) {
      SplitArgs s = split_args(c, level, false, inside_h, inside_s, nullptr, ws, w->b1);
      const int64_t rows = (int64_t)B * s.L * s.N;
      {
        ProfScope prof(c.st, "split_build", 2.0 * rows * D, 4.0 * rows * (5.0 * D + 3));
        launch_k(split_build_kernel<false>, ceil_div(rows, 8), 256, 0, c.st, s);
      }
      CL_CHECK_LAUNCH("split_build_kernel<inside>");
      const int64_t r0 = B * inside_rows_before(n, level);
      CL_TRY(compose_gemm(c, false, r0, rows, w->W2, w->b2, ws));
    }
    CellArgs a = cell_args(c, level, false, ws, inside_h, inside_s);
    a.obj = obj; a.keep = keep;
    if (vl) CL_TRY(launch_cell_aggregate<true>(c, a));
    else CL_TRY(launch_cell_aggregate<false>(c, a));
    if (level < n - 1) CL_TRY(project_level(c, level, inside_h, Wcat_in, PI * D, ws + c.L.Pin));
  }
  return CLIORA_OK;
}

int cliora_outside_fwd(const cliora_dims* dims, const cliora_weights* w, const float* inside_h,
                       const float* inside_s, float* outside_h, float* outside_s, float* ws,
                       cliora_stream_t stream) {
  Ctx c;
  CL_TRY(make_ctx(dims, stream, c));
  if (!w || !inside_h || !inside_s || !outside_h || !outside_s || !ws) return CLIORA_ERR_NULL_POINTER;
  const int B = c.d.B, n = c.d.n, D = c.d.D;
  const float* oW2 = c.d.share ? w->W2 : w->oW2;
  const float* ob1 = c.d.share ? w->b1 : w->ob1;
  const float* ob2 = c.d.share ? w->b2 : w->ob2;
  float* Wcat_out = ws + c.L.Wcat_out;

  launch_k(init_proj_kernel, 296, 256, 0, c.st, ws + c.L.Pout, (int64_t)B * c.C, 2 * D, 0, D, ob1);
  CL_CHECK_LAUNCH("init_proj_kernel");
  launch_k(outside_root_kernel, B, 128, 0, c.st, B, D, c.C, w->root, outside_h, outside_s, ws + c.L.nrm_out,
           (c.d.flags & CLIORA_FLAG_NO_NORMALIZE) ? 1 : 0);
  CL_CHECK_LAUNCH("outside_root_kernel");
  if (n > 1) CL_TRY(project_level(c, n - 1, outside_h, Wcat_out, 2 * D, ws + c.L.Pout));
  for (int level = n - 2; level >= 0; --level) {
    lvl::LevelGeom geom;
    if (fused_level_ok(c, n - level - 1, geom)) {
      CL_TRY(fused_level_fwd(c, level, true, geom, w, inside_h, inside_s, outside_s, outside_h, outside_s, nullptr,
                             nullptr, ws));
      if (level > 0) CL_TRY(project_level(c, level, outside_h, Wcat_out, 2 * D, ws + c.L.Pout));
      continue;
    }
    SplitArgs s = split_args(c, level, true, inside_h, inside_s, outside_s, ws, ob1);
    const int64_t rows = (int64_t)B * s.L * s.N;
    {
      ProfScope prof(c.st, "split_build", 2.0 * rows * D, 4.0 * rows * (5.0 * D + 3));
      launch_k(split_build_kernel<true>, ceil_div(rows, 8), 256, 0, c.st, s);
    }
    CL_CHECK_LAUNCH("split_build_kernel<outside>");
    const int64_t r0 = B * outside_rows_before(n, level);
    CL_TRY(compose_gemm(c, true, r0, rows, oW2, ob2, ws));
    CellArgs a = cell_args(c, level, true, ws, outside_h, outside_s);
    CL_TRY(launch_cell_aggregate<false>(c, a));
    if (level > 0) CL_TRY(project_level(c, level, outside_h, Wcat_out, 2 * D, ws + c.L.Pout));
  }
  return CLIORA_OK;
}

// ---------------------------------------------------------------------------------------------
int cliora_chart_bwd_begin(const cliora_dims* dims, const float* g_inside_h, const float* g_inside_s,
                           const float* g_outside_h, const float* g_outside_s, float* bws,
                           cliora_stream_t stream) {
  Ctx c;
  CL_TRY(make_ctx(dims, stream, c));
  if (!bws) return CLIORA_ERR_NULL_POINTER;
  const int64_t BC = (int64_t)c.d.B * c.C, D = c.d.D;
  auto seed = [&](float* dst, const float* src, int64_t nfl) -> int {
    if (src) CL_CUDA(cudaMemcpyAsync(dst, src, nfl * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
    else CL_CUDA(cudaMemsetAsync(dst, 0, nfl * sizeof(float), c.st));
    return CLIORA_OK;
  };
  CL_TRY(seed(bws + c.L.Gh_in, g_inside_h, BC * D));
  CL_TRY(seed(bws + c.L.Gs_in, g_inside_s, BC));
  CL_TRY(seed(bws + c.L.Gh_out, g_outside_h, BC * D));
  CL_TRY(seed(bws + c.L.Gs_out, g_outside_s, BC));
  CL_CUDA(cudaMemsetAsync(bws + c.L.GP_in, 0, BC * c.L.PI * D * sizeof(float), c.st));
  CL_CUDA(cudaMemsetAsync(bws + c.L.GP_out, 0, BC * 2 * D * sizeof(float), c.st));
  CL_CUDA(cudaMemsetAsync(bws + c.L.CSin, 0, 2 * BC * D * sizeof(float), c.st));   // CSin and CSout are adjacent
  CL_CUDA(cudaMemsetAsync(bws + c.L.db2acc, 0, 2 * D * sizeof(float), c.st));
  return CLIORA_OK;
}

int cliora_outside_bwd(const cliora_dims* dims, const cliora_weights* w, const float* inside_h,
                       const float* inside_s, const float* outside_h, const float* outside_s, float* ws,
                       float* bws, cliora_weight_grads* grads, int phase, cliora_stream_t stream) {
  Ctx c;
  CL_TRY(make_ctx(dims, stream, c));
  if (!w || !inside_h || !inside_s || !outside_h || !outside_s || !ws || !bws || !grads)
    return CLIORA_ERR_NULL_POINTER;
  if ((phase & 3) == 0) return CLIORA_ERR_BAD_SHAPE;
  const int B = c.d.B, n = c.d.n, D = c.d.D;
  const int64_t BC = (int64_t)B * c.C;
  float* Wcat_out = ws + c.L.Wcat_out;
  float* scratch = bws + c.L.splitk;
  if (phase & CLIORA_PHASE_LEVELS) {
  for (int level = 0; level <= n - 2; ++level) {
    if (level > 0) CL_TRY(cellgrad_level(c, level, bws + c.L.GP_out, 2 * D, Wcat_out, bws + c.L.Gh_out));
    CL_TRY((level_bwd<true, false>(c, level, w, inside_h, inside_s, outside_s, const_cast<float*>(outside_h),
                                   const_cast<float*>(outside_s), nullptr, nullptr, ws, bws)));
  }
  if (n > 1) CL_TRY(cellgrad_level(c, n - 1, bws + c.L.GP_out, 2 * D, Wcat_out, bws + c.L.Gh_out));
  if (grads->root) {
    CL_CUDA(cudaMemsetAsync(grads->root, 0, D * sizeof(float), c.st));
    launch_k(outside_root_bwd_kernel, B, 128, 0, c.st, B, D, c.C, bws + c.L.Gh_out, outside_h, ws + c.L.nrm_out, grads->root);
    CL_CHECK_LAUNCH("outside_root_bwd_kernel");
  }
  }   // CLIORA_PHASE_LEVELS
  if (!(phase & CLIORA_PHASE_WEIGHTS)) return CLIORA_OK;
  // weight gradients contributed by the outside pass (independent of the inside backward's level chain)
  const bool sh = c.d.share != 0;
  float* dW1 = sh ? grads->W1 : grads->oW1;
  float* dW2 = sh ? grads->W2 : grads->oW2;
  float* db2 = sh ? grads->b2 : grads->ob2;
  float* dWb = sh ? grads->Wb : grads->oWb;
  const bool fusedb = fused_bwd_ok(c);
  const float* GY = fusedb ? bws + c.L.GYp_out : ws + c.L.Yout;
  const float* Z = ws + c.L.Zout;
  const float* GPo = bws + c.L.GP_out;
  const int64_t lo_out = c.use_tc ? c.L.rows_out * D : 0;
  if (dW2) {
    if (c.use_tc) {
      tc::PairRef Ap{GY, c.L.rows_out, D, lo_out}, Bp{Z, c.L.rows_out, D, lo_out};
      CL_TRY(tc::launch_tc_gemm_tn(c.st, Ap, Bp, (int)c.L.rows_out, D, D, dW2, D, 0, scratch, "tc_gemm_wgrad_w2", c.tc_mode));
    } else {
      CL_TRY(launch_gemm_tn(c.st, (int)c.L.rows_out, D, D, GY, D, Z, D, dW2, D, 0, scratch, "gemm_wgrad", 0, 0));
    }
  }
  if (db2 && fusedb) {
    CL_TRY(colsum(c.st, bws + c.L.db2acc + D, D, 1, D, db2, 0, scratch));
  } else if (db2) {
    if (cellsum_enabled(c, true)) {
      CL_TRY(colsum(c.st, bws + c.L.CSout, D, BC, D, db2, 0, scratch));
    } else {
      CL_TRY(colsum(c.st, GY, D, c.L.rows_out, D, db2, 0, scratch));
      if (lo_out) CL_TRY(colsum(c.st, GY + lo_out, D, c.L.rows_out, D, db2, 1, scratch));
    }
  }
  {
    float* dst[2] = {dW1 ? dW1 + D : nullptr, dWb};
    const int64_t ldc[2] = {2 * D, D};
    const int accs[2] = {0, 0};
    CL_TRY(cell_wgrad(c, GPo, 2, outside_h, dst, ldc, accs, bws));
  }
  if (!sh) {
    const float* GPi = bws + c.L.GP_in;
    if (grads->oW1)   // first-argument projection of the outside compose (column block 3 of GP_in)
      CL_TRY(launch_gemm_tn(c.st, (int)BC, D, D, GPi + 3 * D, 4 * D, inside_h, D, grads->oW1, 2 * D, 0, scratch));
    if (grads->ob1) CL_TRY(colsum(c.st, GPi + 3 * D, 4 * D, BC, D, grads->ob1, 0, scratch));
  }
  return CLIORA_OK;
}

int cliora_inside_bwd(const cliora_dims* dims, const cliora_weights* w, const float* x, const float* obj,
                      const uint8_t* keep, const float* inside_h, const float* inside_s,
                      const float* outside_h, float* ws, float* bws, int had_outside, float* grad_x,
                      float* grad_obj, cliora_weight_grads* grads, int phase, cliora_stream_t stream) {
  Ctx c;
  CL_TRY(make_ctx(dims, stream, c));
  if (!w || !x || !inside_h || !inside_s || !ws || !bws || !grads) return CLIORA_ERR_NULL_POINTER;
  if ((phase & 3) == 0) return CLIORA_ERR_BAD_SHAPE;
  (void)outside_h;
  const int B = c.d.B, n = c.d.n, D = c.d.D, R = c.d.R, PI = (int)c.L.PI;
  const int64_t BC = (int64_t)B * c.C;
  const bool vl = R > 0;
  if (vl && obj == nullptr) return CLIORA_ERR_NULL_POINTER;
  float* Wcat_in = ws + c.L.Wcat_in;
  float* scratch = bws + c.L.splitk;
  float* ih = const_cast<float*>(inside_h);
  float* is_ = const_cast<float*>(inside_s);
  if (phase & CLIORA_PHASE_LEVELS) {
  for (int level = n - 1; level >= 0; --level) {
    if (level < n - 1) CL_TRY(cellgrad_level(c, level, bws + c.L.GP_in, PI * D, Wcat_in, bws + c.L.Gh_in));
    if (vl) CL_TRY((level_bwd<false, true>(c, level, w, inside_h, inside_s, nullptr, ih, is_, obj, keep, ws, bws)));
    else CL_TRY((level_bwd<false, false>(c, level, w, inside_h, inside_s, nullptr, ih, is_, obj, keep, ws, bws)));
  }
  }   // CLIORA_PHASE_LEVELS (uses no split-K scratch, so the outside weight phase may run concurrently)
  if (!(phase & CLIORA_PHASE_WEIGHTS)) return CLIORA_OK;
  // leaf linear layer
  const float* gu = bws + c.L.gu;
  if (grad_x) {
    GemmParams p{};
    p.A = gu; p.lda = D; p.amap = dense_rows();
    p.W = w->W_leaf; p.ldw = D;
    p.C = grad_x; p.ldc = D; p.cmap = dense_rows();
    p.M = B * n; p.N = D; p.K = D;
    p.mma_ok = c.use_tc && g_debug[5] == 0;
    CL_TRY(launch_gemm(c.st, /*nt=*/false, p));
  }
  if (grads->W_leaf) CL_TRY(launch_gemm_tn(c.st, B * n, D, D, gu, D, x, D, grads->W_leaf, D, 0, scratch));
  if (grads->b_leaf) CL_TRY(colsum(c.st, gu, D, (int64_t)B * n, D, grads->b_leaf, 0, scratch));

  const int acc = (c.d.share && had_outside) ? 1 : 0;
  const bool fusedb = fused_bwd_ok(c);
  const float* GY = fusedb ? bws + c.L.GYp_in : ws + c.L.Yin;
  const float* Z = ws + c.L.Zin;
  const float* GPi = bws + c.L.GP_in;
  const int ldp = PI * D;
  const int64_t lo_in = c.use_tc ? c.L.rows_in * D : 0;
  if (grads->W2) {
    if (c.use_tc) {
      tc::PairRef Ap{GY, c.L.rows_in, D, lo_in}, Bp{Z, c.L.rows_in, D, lo_in};
      CL_TRY(tc::launch_tc_gemm_tn(c.st, Ap, Bp, (int)c.L.rows_in, D, D, grads->W2, D, acc, scratch, "tc_gemm_wgrad_w2", c.tc_mode));
    } else {
      CL_TRY(launch_gemm_tn(c.st, (int)c.L.rows_in, D, D, GY, D, Z, D, grads->W2, D, acc, scratch, "gemm_wgrad", 0, 0));
    }
  }
  if (grads->b2 && fusedb) {
    CL_TRY(colsum(c.st, bws + c.L.db2acc, D, 1, D, grads->b2, acc, scratch));
  } else if (grads->b2) {
    if (cellsum_enabled(c, false)) {
      CL_TRY(colsum(c.st, bws + c.L.CSin, D, BC, D, grads->b2, acc, scratch));
    } else {
      CL_TRY(colsum(c.st, GY, D, c.L.rows_in, D, grads->b2, acc, scratch));
      if (lo_in) CL_TRY(colsum(c.st, GY + lo_in, D, c.L.rows_in, D, grads->b2, 1, scratch));
    }
  }
  {
    float* dst[4] = {grads->W1, grads->W1 ? grads->W1 + D : nullptr, grads->Wb, nullptr};
    const int64_t ldc[4] = {2 * D, 2 * D, D, D};
    const int accs[4] = {0, acc, acc, 0};
    CL_TRY(cell_wgrad(c, GPi, PI, inside_h, dst, ldc, accs, bws));
  }
  if (grads->b1) CL_TRY(colsum(c.st, GPi, ldp, BC, D, grads->b1, 0, scratch));
  if (vl && grad_obj) {
    // cells split over grid.z so that the launch fills the GPU at small batch sizes
    int zs = (int)((8 * 148 + (int64_t)ceil_div(D, 32) * B - 1) / ((int64_t)ceil_div(D, 32) * B));
    if (zs > (int)((c.C + 15) / 16)) zs = (int)((c.C + 15) / 16);
    if (zs < 1) zs = 1;
    dim3 grid(ceil_div(D, 32), B, zs);
    if (zs > 1) CL_CUDA(cudaMemsetAsync(grad_obj, 0, (size_t)B * R * D * sizeof(float), c.st));
    launch_k(obj_grad_kernel<64>, grid, dim3(32, 4), 0, c.st, D, R, c.C, bws + c.L.GA2, ws + c.L.q_in, bws + c.L.coef, grad_obj, 0);
    CL_CHECK_LAUNCH("obj_grad_kernel");
  }
  if (!had_outside) {
    if (grads->root) CL_CUDA(cudaMemsetAsync(grads->root, 0, D * sizeof(float), c.st));
    if (!c.d.share) {
      if (grads->oW1) CL_CUDA(cudaMemsetAsync(grads->oW1, 0, (size_t)2 * D * D * sizeof(float), c.st));
      if (grads->ob1) CL_CUDA(cudaMemsetAsync(grads->ob1, 0, D * sizeof(float), c.st));
      if (grads->oW2) CL_CUDA(cudaMemsetAsync(grads->oW2, 0, (size_t)D * D * sizeof(float), c.st));
      if (grads->ob2) CL_CUDA(cudaMemsetAsync(grads->ob2, 0, D * sizeof(float), c.st));
      if (grads->oWb) CL_CUDA(cudaMemsetAsync(grads->oWb, 0, (size_t)D * D * sizeof(float), c.st));
    }
  }
  return CLIORA_OK;
}

// ---------------------------------------------------------------------------------------------
int cliora_atten_scores(int B, int ncell, int D, int R, const float* h, int64_t h_batch_stride, const float* obj,
                        float* scores, cliora_stream_t stream) {
  if (!h || !obj || !scores) return CLIORA_ERR_NULL_POINTER;
  if (B < 1 || ncell < 1 || D < 4 || D % 4 || R < 1) return CLIORA_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  // one GEMM per image c: scores[a, c, cell, :] = h[a, cell] . obj[c]^T
  for (int cimg = 0; cimg < B; ++cimg) {
    GemmParams p{};
    p.A = h; p.lda = D; p.amap = RowMap{ncell, h_batch_stride, 0};
    p.W = obj + (int64_t)cimg * R * D; p.ldw = D;
    p.C = scores; p.ldc = R; p.cmap = RowMap{ncell, (int64_t)B * ncell, (int64_t)cimg * ncell};
    p.M = B * ncell; p.N = R; p.K = D;
    CL_TRY(launch_gemm(st, true, p));
  }
  return CLIORA_OK;
}

int cliora_atten_max_fwd(int B, int ncell, int D, int R, const float* h, int64_t h_batch_stride, const float* obj,
                         float* smax, int32_t* amax, cliora_stream_t stream) {
  if (B < 1 || ncell < 0 || D < 4 || D % 4 || R < 1 || R > 64) return CLIORA_ERR_BAD_SHAPE;
  if (ncell == 0) return CLIORA_OK;   // single-word sentences: no cell takes part in the loss, outputs are empty
  if (!h || !obj || !smax || !amax) return CLIORA_ERR_NULL_POINTER;
  dim3 grid(B, ceil_div((int64_t)B * ncell, 64));
  {
    ProfScope prof((cudaStream_t)stream, "atten_max", 2.0 * B * ncell * (double)B * R * D,
                   4.0 * ((double)B * ncell * D + (double)B * R * D + 2.0 * B * B * ncell));
    launch_k(atten_max_kernel, grid, 256, 0, (cudaStream_t)stream, B, ncell, D, R, h, h_batch_stride, obj, smax, amax);
  }
  CL_CHECK_LAUNCH("atten_max_kernel");
  return CLIORA_OK;
}

int cliora_atten_max_bwd(int B, int ncell, int D, int R, const float* h, int64_t h_batch_stride, const float* obj,
                         const float* g_smax, const int32_t* amax, float* g_h, int64_t gh_batch_stride,
                         float* g_obj, cliora_stream_t stream) {
  if (B < 1 || ncell < 0 || D < 4 || D % 4 || R < 1 || R > 64) return CLIORA_ERR_BAD_SHAPE;
  if (ncell == 0) return CLIORA_OK;
  if (!h || !obj || !g_smax || !amax) return CLIORA_ERR_NULL_POINTER;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(st, "atten_max_bwd", 4.0 * B * ncell * (double)B * D, 4.0 * 2.0 * B * ncell * (double)B * D);
  if (g_h) {
    launch_k(atten_max_bwd_h_kernel, B * ncell, 128, 0, st, B, ncell, D, R, obj, g_smax, amax, g_h, gh_batch_stride);
    CL_CHECK_LAUNCH("atten_max_bwd_h_kernel");
  }
  if (g_obj) {
    launch_k(atten_max_bwd_obj_kernel, dim3(B * R, 8), 128, 4 * 512 * sizeof(float), st, B, ncell, D, R, h, h_batch_stride, g_smax, amax, g_obj);
    CL_CHECK_LAUNCH("atten_max_bwd_obj_kernel");
  }
  return CLIORA_OK;
}

int cliora_contrastive_loss(int B, int cells, int ncell, const float* smax, const float* inside_s,
                            const float* outside_s, float margin, float alpha, float* loss_out, float* g_smax,
                            float* g_inside_s, float* g_outside_s, float* scratch, cliora_stream_t stream) {
  if (!inside_s || !outside_s || !loss_out || !scratch) return CLIORA_ERR_NULL_POINTER;
  if (B < 1 || cells < 1 || ncell < 0 || ncell > cells) return CLIORA_ERR_BAD_SHAPE;
  if (ncell > 0 && !smax) return CLIORA_ERR_NULL_POINTER;
  if (g_smax && (!g_inside_s || !g_outside_s)) return CLIORA_ERR_NULL_POINTER;
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = scratch;                        // [ncell]
  float* root_part = scratch + ((ncell + 3) & ~3); // [ncell, B]
  const float scale = alpha / (float)B;
  if (ncell > 0) {
    const size_t smem = (size_t)(2 * B + 64) * sizeof(float);
    launch_k(contrastive_cell_kernel, ncell, 128, smem, st, B, cells, ncell, smax, inside_s, outside_s, margin, scale, partial, g_smax, g_inside_s, g_outside_s, g_smax ? root_part : nullptr);
    CL_CHECK_LAUNCH("contrastive_cell_kernel");
  }
  launch_k(contrastive_finish_kernel, 1, 128, 0, st, B, cells, ncell, partial, scale, loss_out, g_smax ? root_part : nullptr, g_inside_s);
  CL_CHECK_LAUNCH("contrastive_finish_kernel");
  return CLIORA_OK;
}

int cliora_vg_loss(int B, int n, const float* wmax, float alpha, float* loss_out, float* g_wmax, float* scratch,
                   cliora_stream_t stream) {
  if (!wmax || !loss_out || !scratch) return CLIORA_ERR_NULL_POINTER;
  if (B < 1 || n < 1) return CLIORA_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)(B + 64) * sizeof(float);
  launch_k(vg_loss_kernel, B, 128, smem, st, B, n, wmax, alpha, scratch, g_wmax);
  CL_CHECK_LAUNCH("vg_loss_kernel");
  launch_k(sum_small_kernel, 1, 128, 0, st, scratch, B, loss_out);
  CL_CHECK_LAUNCH("sum_small_kernel");
  return CLIORA_OK;
}

int64_t cliora_adam_table_bytes(int ntensors) { return (int64_t)ntensors * sizeof(AdamTensor); }

int cliora_adam_table_fill(int ntensors, void* const* params, const void* const* grads, void* const* exp_avg,
                           void* const* exp_avg_sq, const int64_t* numel, void* host_table, int64_t* total_blocks) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !numel || !host_table || !total_blocks)
    return CLIORA_ERR_NULL_POINTER;
  AdamTensor* t = static_cast<AdamTensor*>(host_table);
  int64_t blk = 0;
  for (int i = 0; i < ntensors; ++i) {
    t[i].p = static_cast<float*>(params[i]);
    t[i].g = static_cast<const float*>(grads[i]);
    t[i].m = static_cast<float*>(exp_avg[i]);
    t[i].v = static_cast<float*>(exp_avg_sq[i]);
    t[i].n = numel[i];
    t[i].block0 = blk;
    blk += (numel[i] + kAdamChunk - 1) / kAdamChunk;
  }
  *total_blocks = blk;
  return CLIORA_OK;
}

int cliora_adam_step(const void* device_table, int ntensors, int64_t total_blocks, float lr, float beta1, float beta2,
                     float eps, float max_norm, float* state, float* scratch, cliora_stream_t stream) {
  if (!device_table || !state || !scratch) return CLIORA_ERR_NULL_POINTER;
  if (ntensors < 1 || total_blocks < 1) return CLIORA_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const AdamTensor* tab = static_cast<const AdamTensor*>(device_table);
  launch_k(adam_gradnorm_kernel, (unsigned)total_blocks, kAdamBlock, 0, st, tab, ntensors, scratch);
  CL_CHECK_LAUNCH("adam_gradnorm_kernel");
  launch_k(adam_norm_finish_kernel, 1, 256, 0, st, (const float*)scratch, (int)total_blocks, max_norm, state);
  CL_CHECK_LAUNCH("adam_norm_finish_kernel");
  launch_k(adam_update_kernel, (unsigned)total_blocks, kAdamBlock, 0, st, tab, ntensors, lr, beta1, beta2, eps,
           (const float*)state);
  CL_CHECK_LAUNCH("adam_update_kernel");
  return CLIORA_OK;
}

int cliora_tree_spans(int B, int n, const int32_t* backptr, int32_t* spans, int32_t* scratch, cliora_stream_t stream) {
  if (!backptr || !spans || !scratch) return CLIORA_ERR_NULL_POINTER;
  if (B < 1 || n < 2 || n > 512) return CLIORA_ERR_BAD_SHAPE;
  launch_k(tree_spans_kernel, ceil_div(B, 64), 64, 0, (cudaStream_t)stream, B, n, backptr, spans, scratch);
  CL_CHECK_LAUNCH("tree_spans_kernel");
  return CLIORA_OK;
}

int cliora_span_f1(int B, int n, int G, const int32_t* spans, const int32_t* gold, const int32_t* gold_len, float* out,
                   cliora_stream_t stream) {
  if (!spans || !gold || !gold_len || !out) return CLIORA_ERR_NULL_POINTER;
  if (B < 1 || n < 2 || G < 0) return CLIORA_ERR_BAD_SHAPE;
  launch_k(span_f1_kernel, ceil_div(B, 64), 64, 0, (cudaStream_t)stream, B, n, G, spans, gold, gold_len, out);
  CL_CHECK_LAUNCH("span_f1_kernel");
  return CLIORA_OK;
}

int cliora_grounding_eval(int B, int n, int R, int P, const float* atten_score, const float* boxes,
                          const int32_t* phrases, const float* gt_boxes, float iou_thresh, int32_t* sel, float* iou,
                          int32_t* hit, cliora_stream_t stream) {
  if (B < 1 || n < 1 || R < 1 || P < 0) return CLIORA_ERR_BAD_SHAPE;
  if (P == 0) return CLIORA_OK;
  if (!atten_score || !boxes || !phrases || !gt_boxes || !sel || !iou || !hit) return CLIORA_ERR_NULL_POINTER;
  launch_k(grounding_eval_kernel, ceil_div(P, 4), 128, 0, (cudaStream_t)stream, B, n, R, P, atten_score, boxes, phrases,
           gt_boxes, iou_thresh, sel, iou, hit);
  CL_CHECK_LAUNCH("grounding_eval_kernel");
  return CLIORA_OK;
}

int cliora_gather_regions(int B, int R, int F, int feat_dtype, const void* features, const float* bboxes,
                          const int32_t* classes, const int64_t* pos_bboxes, const int64_t* img_index, float* obj_feats,
                          float* boxes, int64_t* obj_cates, cliora_stream_t stream) {
  if (!features || !pos_bboxes || !img_index || !obj_feats) return CLIORA_ERR_NULL_POINTER;
  if (boxes && !bboxes) return CLIORA_ERR_NULL_POINTER;
  if (B < 1 || R < 1 || F < 8 || F % 8 || (feat_dtype != 0 && feat_dtype != 1)) return CLIORA_ERR_BAD_SHAPE;
  ProfScope prof((cudaStream_t)stream, "gather_regions", 0.0,
                 (double)B * R * F * (feat_dtype ? 6.0 : 8.0) + (double)B * R * 44.0);
  if (feat_dtype == 0)
    launch_k(gather_regions_kernel<float>, B * R, 128, 0, (cudaStream_t)stream, B, R, F, (const float*)features, bboxes,
             classes, pos_bboxes, img_index, obj_feats, boxes, obj_cates);
  else
    launch_k(gather_regions_kernel<__half>, B * R, 128, 0, (cudaStream_t)stream, B, R, F, (const __half*)features,
             bboxes, classes, pos_bboxes, img_index, obj_feats, boxes, obj_cates);
  CL_CHECK_LAUNCH("gather_regions_kernel");
  return CLIORA_OK;
}

int cliora_recon_ce_fwd(int rows, int D, int K, const float* cell, const float* pos, const float* neg, float* rowloss,
                        float* probs, cliora_stream_t stream) {
  if (!cell || !pos || !neg || !rowloss || !probs) return CLIORA_ERR_NULL_POINTER;
  if (rows < 1 || D < 4 || D % 4 || D > 128 * kColT || K < 1 || K > 127) return CLIORA_ERR_BAD_SHAPE;
  launch_k(recon_ce_fwd_kernel, ceil_div(rows, 8), 256, (size_t)kNegChunk * D * sizeof(float), (cudaStream_t)stream, rows,
           D, K, cell, pos, neg, rowloss, probs);
  CL_CHECK_LAUNCH("recon_ce_fwd_kernel");
  return CLIORA_OK;
}

int cliora_recon_ce_bwd(int rows, int D, int K, const float* cell, const float* pos, const float* neg,
                        const float* probs, const float* g_loss, float* g_scores, float* g_cell, float* g_pos,
                        cliora_stream_t stream) {
  if (!cell || !pos || !neg || !probs || !g_loss || !g_scores || !g_cell || !g_pos) return CLIORA_ERR_NULL_POINTER;
  if (rows < 1 || D < 4 || D % 4 || D > 128 * kColT || K < 1 || K > 127) return CLIORA_ERR_BAD_SHAPE;
  launch_k(recon_ce_bwd_kernel, ceil_div(rows, 8), 256, (size_t)kNegChunk * D * sizeof(float), (cudaStream_t)stream, rows,
           D, K, cell, pos, neg, probs, g_loss, 1.f / (float)rows, g_scores, g_cell, g_pos);
  CL_CHECK_LAUNCH("recon_ce_bwd_kernel");
  return CLIORA_OK;
}

int cliora_cky(int B, int n, const float* split_scores, int32_t* backptr, float* best, cliora_stream_t stream) {
  if (!backptr) return CLIORA_ERR_NULL_POINTER;
  if (B < 1 || n < 1 || n > 512) return CLIORA_ERR_BAD_SHAPE;
  if (n > 1 && !split_scores) return CLIORA_ERR_NULL_POINTER;
  const size_t smem = (size_t)num_cells(n) * sizeof(float);
  if (smem <= 200 * 1024) {
    if (smem > 48 * 1024)
      CL_CUDA(func_attr_at_least(reinterpret_cast<const void*>(cky_kernel<false>),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_k(cky_kernel<false>, B, kCkyThreads, smem, (cudaStream_t)stream, B, n, split_scores, backptr, best);
  } else {
    // sentences whose Viterbi chart does not fit shared memory (n > 319): the chart lives in the caller's `best` rows
    if (!best) return CLIORA_ERR_UNSUPPORTED;
    launch_k(cky_kernel<true>, B, kCkyThreads, 0, (cudaStream_t)stream, B, n, split_scores, backptr, best);
  }
  CL_CHECK_LAUNCH("cky_kernel");
  return CLIORA_OK;
}

// ---------------------------------------------------------------------------------------------
int cliora_linear(int M, int N, int K, const float* A, const float* W, const float* bias, int act, float* C,
                  cliora_stream_t stream) {
  if (!A || !W || !C) return CLIORA_ERR_NULL_POINTER;
  if (M < 0 || N < 1 || K < 1) return CLIORA_ERR_BAD_SHAPE;
  return dense_linear((cudaStream_t)stream, M, N, K, A, W, bias, act, C);
}

int cliora_matmul_nn(int M, int N, int K, const float* A, const float* Bm, float* C, int accumulate,
                     cliora_stream_t stream) {
  if (!A || !Bm || !C) return CLIORA_ERR_NULL_POINTER;
  if (M < 0 || N < 1 || K < 1) return CLIORA_ERR_BAD_SHAPE;
  GemmParams p{};
  p.A = A; p.lda = K; p.amap = dense_rows();
  p.W = Bm; p.ldw = N;
  p.C = C; p.ldc = N; p.cmap = dense_rows();
  p.M = M; p.N = N; p.K = K;
  p.accumulate = accumulate;
  return launch_gemm((cudaStream_t)stream, false, p);
}

void cliora_debug_ptr(int key, void* p) {
  if (key == 0) lvl::g_level_dbg = static_cast<long long*>(p);
}

void cliora_debug_set(int key, int value) {
  if (key == 100) { g_pdl = value ? 1 : 0; return; }
  if (key == 101) { g_carveout = value; return; }
  if (key == 102) { tc::g_tc_small_tmem = value; return; }
  if (key == 103) { tc::g_tc_narrow_stages = value; return; }
  if (key == 104) { g_splitk_target = value; return; }
  if (key == 105) { tc::g_tc_xnarrow = value; return; }      // preferred shared-memory carveout, percent (-1: leave)   // programmatic dependent launch on/off
  if (key >= 0 && key < 16) g_debug[key] = value;
}

int cliora_split_tf32(const float* x, int64_t n, float* out_pair, cliora_stream_t stream) {
  if (!x || !out_pair) return CLIORA_ERR_NULL_POINTER;
  if (n <= 0) return CLIORA_OK;
  launch_k(tc::split_tf32_kernel, ceil_div(n, 256 * 4) < 1184 ? ceil_div(n, 256 * 4) : 1184, 256, 0, (cudaStream_t)stream, x, n, out_pair);
  CL_CHECK_LAUNCH("split_tf32_kernel");
  return CLIORA_OK;
}

int cliora_tc_linear(int M, int N, int K, const float* A_pair, const float* W_pair, const float* bias, int act,
                     float* C, cliora_stream_t stream) {
  if (!A_pair || !W_pair || !C) return CLIORA_ERR_NULL_POINTER;
  if (M < 0 || N < 1 || K < 1) return CLIORA_ERR_BAD_SHAPE;
  tc::PairRef A{A_pair, M, K, (int64_t)M * K};
  tc::PairRef W{W_pair, N, K, (int64_t)N * K};
  if (!tc::tc_supported(N, K, A, W)) return CLIORA_ERR_UNSUPPORTED;
  tc::TcEpilogue ep{};
  ep.C = C; ep.ldc = N; ep.cmap = dense_rows();
  ep.bias = bias; ep.act = act;
  return tc::launch_tc_gemm_nt((cudaStream_t)stream, A, 0, W, M, N, K, ep, "tc_gemm_linear", g_debug[0], g_debug[2]);
}

int cliora_tc_atten_max_fwd(int B, int ncell, int D, int R, const float* h_pair, const float* obj_pair, float* smax,
                            int32_t* amax, cliora_stream_t stream) {
  if (B < 1 || ncell < 0 || D < 32 || D % 4 || R < 1 || R > 64) return CLIORA_ERR_BAD_SHAPE;
  if (ncell == 0) return CLIORA_OK;
  if (!h_pair || !obj_pair || !smax || !amax) return CLIORA_ERR_NULL_POINTER;
  const int64_t M = (int64_t)B * ncell, N = (int64_t)B * R;
  tc::PairRef A{h_pair, M, D, M * D};
  tc::PairRef W{obj_pair, N, D, N * D};
  tc::TcEpilogue ep{};
  ep.cmap = dense_rows();
  ep.gmax = smax; ep.gargmax = amax;
  ep.R = R; ep.ncell = ncell; ep.B_img = B;
  ep.n_stride = (tc::kTcNarrowN / R) * R;     // whole images per 80-column tile
  return tc::launch_tc_gemm_nt_cfg<tc::kTcNarrowN, tc::kTcNarrowStages>((cudaStream_t)stream, A, 0, W, (int)M, (int)N, D,
                                                                        ep, "tc_atten_max", 2);
}

int64_t cliora_tc_matmul_tn_scratch_floats(int M, int Ka, int Kb) { return tc::tn_tc_scratch_floats(M, Ka, Kb); }

int cliora_tc_matmul_tn(int M, int Ka, int Kb, const float* A_pair, const float* B_pair, float* C, int accumulate,
                        float* scratch, cliora_stream_t stream) {
  if (!A_pair || !B_pair || !C || !scratch) return CLIORA_ERR_NULL_POINTER;
  if (M < 0 || Ka < 1 || Kb < 1 || (Ka % 4) || (Kb % 4)) return CLIORA_ERR_BAD_SHAPE;
  tc::PairRef A{A_pair, M, Ka, (int64_t)M * Ka};
  tc::PairRef B{B_pair, M, Kb, (int64_t)M * Kb};
  return tc::launch_tc_gemm_tn((cudaStream_t)stream, A, B, M, Ka, Kb, C, Kb, accumulate, scratch, "tc_gemm_wgrad");
}

int64_t cliora_matmul_tn_scratch_floats(int M, int Ka, int Kb) { return tn_scratch_floats(M, Ka, Kb); }

int cliora_matmul_tn(int M, int Ka, int Kb, const float* A, const float* Bm, float* C, int accumulate,
                     float* scratch, cliora_stream_t stream) {
  if (!A || !Bm || !C || !scratch) return CLIORA_ERR_NULL_POINTER;
  if (M < 0 || Ka < 1 || Kb < 1) return CLIORA_ERR_BAD_SHAPE;
  return launch_gemm_tn((cudaStream_t)stream, M, Ka, Kb, A, Ka, Bm, Kb, C, Kb, accumulate, scratch);
}

}  // extern "C"
