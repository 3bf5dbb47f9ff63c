// Micro-benchmark (dev tool, not part of the library): how fast can one SM gather the A-operand rows of a fused chart
// level?  Every CTA plays the producer side of level_fwd_kernel's main loop: per k-block it lands 2 x 128 row slices of
// 128 bytes (random rows of a [6720, 1200] fp32 projection buffer that sits in L2) in a 4-stage shared-memory ring, eight
// consumer warps read the stage (checksum) and hand it back.  Variants of the copy:
//   0  cp.async 16 B, 128 threads x 16 copies per k-block (what level_fwd_kernel ships today)
//   1  cp.async.bulk 1-D, one 128-byte copy per row slice (256 per k-block), one issuing warp
//   2  TMA tile::gather4, 4 row slices per instruction (64 per k-block), one issuing warp
//   3  TMA tile::gather4 + multicast over a cluster of 4: every CTA issues a quarter (16) for all four
//   4  cp.async.bulk 1-D + multicast over a cluster of 4 (64 per CTA and k-block)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o dev/proto/gather_bench.bin dev/proto/gather_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int ROWS = 6720, LD = 1200, D = 400;
constexpr int kStages = 4, kStageBytes = 2 * 128 * 128;
#ifndef COPY_THREADS
#define COPY_THREADS 128
#endif
constexpr int kConsWarps = 8, kCopyThreads = COPY_THREADS;
constexpr int kThreads = (1 + kConsWarps) * 32 + kCopyThreads;   // warp 0 = TMA issuer, 1..8 consumers, then copy threads

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t caddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ uint32_t ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_1d_mc(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int col, int r0, int r1, int r2,
                                        int r3) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gather4_mc(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int col, int r0, int r1,
                                           int r2, int r3, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;"
               ::"r"(dst), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Args {
  const float* P;
  const int* rows;     // [tiles][2][128] chart rows of the two operands of every tile row
  float* sums;         // [tiles * nc][128] per-row checksum (as seen by every CTA)
  long long* cycles;   // [tiles * nc]
  int num_kb, mode, nc, swz;
};

__global__ void __launch_bounds__(kThreads, 1) gather_kernel(const __grid_constant__ CUtensorMap tm, const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  __shared__ int s_rows[256];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nc = a.nc;
  const int rank = nc > 1 ? (int)ctarank() : 0;
  const int tile = blockIdx.x / nc;
  const bool mc = a.mode == 3 || a.mode == 4;
  if (tid < 256) s_rows[tid] = a.rows[tile * 256 + tid];
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], a.mode == 0 ? kCopyThreads : 1);
      mbar_init(&empty[i], mc ? kConsWarps * nc : kConsWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (nc > 1) cluster_sync();
  const long long t0 = clock64();
  const uint32_t sbase = smem_u32(smem);
  if (warp == 0) {
    if (a.mode >= 1) {
      for (int kb = 0; kb < a.num_kb; ++kb) {
        const int st = kb % kStages;
        if (mc) { if (lane == 0) mbar_wait_cluster(&empty[st], ((kb / kStages) & 1) ^ 1); __syncwarp(); }
        else mbar_wait(&empty[st], ((kb / kStages) & 1) ^ 1);
        if (lane == 0) mbar_expect_tx(&full[st], kStageBytes);
        __syncwarp();
        const uint32_t sdst = sbase + st * kStageBytes;
        if (a.mode == 1) {
          for (int i = lane; i < 256; i += 32) {         // i = operand * 128 + row
            const int r = i & 127;
            bulk_1d(sdst + i * 128, a.P + (int64_t)s_rows[i] * LD + (i >> 7) * D + kb * 32, 128, &full[st]);
            (void)r;
          }
        } else if (a.mode == 2) {
          for (int i = lane; i < 64; i += 32) {          // i = operand * 32 + row group
            const int op = i >> 5, r0 = (i & 31) * 4;
            const int* rr = s_rows + op * 128 + r0;
            gather4(sdst + op * 16384 + r0 * 128, &tm, &full[st], op * D + kb * 32, rr[0], rr[1], rr[2], rr[3]);
          }
        } else if (a.mode == 3) {
          if (lane < 16) {                               // this CTA's quarter: row groups rank*8 .. +8 of both operands
            const int op = lane >> 3, r0 = (rank * 8 + (lane & 7)) * 4;
            const int* rr = s_rows + op * 128 + r0;
            gather4_mc(sdst + op * 16384 + r0 * 128, &tm, &full[st], op * D + kb * 32, rr[0], rr[1], rr[2], rr[3],
                       (uint16_t)((1u << nc) - 1));
          }
        } else {
          for (int j = lane; j < 64; j += 32) {          // this CTA's quarter of the 256 row slices
            const int i = (j >> 5) * 128 + rank * 32 + (j & 31);
            bulk_1d_mc(sdst + i * 128, a.P + (int64_t)s_rows[i] * LD + (i >> 7) * D + kb * 32, 128, &full[st],
                       (uint16_t)((1u << nc) - 1));
          }
        }
      }
    }
  } else if (warp <= kConsWarps) {
    const int cw = warp - 1, qd = cw & 3, half = cw >> 2;
    const int row = qd * 32 + lane;
    float acc = 0.f;
    for (int kb = 0; kb < a.num_kb; ++kb) {
      const int st = kb % kStages;
      mbar_wait(&full[st], (kb / kStages) & 1);
      const uint8_t* sA = smem + st * kStageBytes + row * 128;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int chunk = half * 4 + t;
        const int ch = (a.swz ? (chunk ^ (row & 7)) : chunk) << 4;
        const float4 xa = *reinterpret_cast<const float4*>(sA + ch);
        const float4 xb = *reinterpret_cast<const float4*>(sA + 16384 + ch);
        const float w = (float)(chunk + 1);
        acc += w * (xa.x + 2.f * xa.y + 3.f * xa.z + 4.f * xa.w) + (xb.x + xb.y + xb.z + xb.w);
      }
      __syncwarp();
      if (mc) {
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
        if (lane < nc) mbar_arrive_cluster(mapa(smem_u32(&empty[st]), lane));
      } else if (lane == 0) mbar_arrive(&empty[st]);
    }
    atomicAdd(a.sums + (int64_t)blockIdx.x * 128 + row, acc);
  } else if (a.mode == 0) {
    const int gt = tid - (1 + kConsWarps) * 32;
    const int c = gt & 7, rbase = gt >> 3;
    for (int kb = 0; kb < a.num_kb; ++kb) {
      const int st = kb % kStages;
      mbar_wait(&empty[st], ((kb / kStages) & 1) ^ 1);
      const uint32_t sdst = sbase + st * kStageBytes;
#pragma unroll
      for (int i = 0; i < 1024 / kCopyThreads; ++i) {
        const int r = rbase + (kCopyThreads / 8) * i;
        const uint32_t so = (uint32_t)(r * 128 + (((a.swz ? (c ^ (r & 7)) : c)) << 4));
        cp_async16(sdst + so, a.P + (int64_t)s_rows[r] * LD + kb * 32 + c * 4);
        cp_async16(sdst + 16384 + so, a.P + (int64_t)s_rows[128 + r] * LD + D + kb * 32 + c * 4);
      }
      cp_async_arrive_noinc(&full[st]);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  if (nc > 1) cluster_sync();
  if (tid == 0) a.cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int tiles = argc > 2 ? atoi(argv[2]) : 25;
  const int box_rows = argc > 3 ? atoi(argv[3]) : 1;
  const int swz = argc > 4 ? atoi(argv[4]) : 1;
  const int num_kb = argc > 5 ? atoi(argv[5]) : 52;
  const int nc = argc > 6 ? atoi(argv[6]) : 4;      // CTAs that want the same tile (cluster size for modes 3, 4)
  std::vector<float> hP((size_t)ROWS * LD);
  for (size_t i = 0; i < hP.size(); ++i) hP[i] = (float)((i * 2654435761u) % 1000) * 1e-3f;
  std::vector<int> hrows((size_t)tiles * 256);
  uint32_t s = 12345;
  for (auto& r : hrows) { s = s * 1664525u + 1013904223u; r = (int)((s >> 8) % ROWS); }
  float* P; int* rows; float* sums; long long* cyc;
  cudaMalloc(&P, hP.size() * 4); cudaMalloc(&rows, hrows.size() * 4);
  cudaMalloc(&sums, (size_t)tiles * nc * 128 * 4); cudaMalloc(&cyc, (size_t)tiles * nc * 8);
  cudaMemcpy(P, hP.data(), hP.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(rows, hrows.data(), hrows.size() * 4, cudaMemcpyHostToDevice);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)LD, (cuuint64_t)ROWS};
  cuuint64_t gstride[1] = {(cuuint64_t)LD * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = reinterpret_cast<EncodeTiledFn>(fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, P, gdim, gstride, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                    swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", (int)cr); return 1; }
  const size_t smem = kStages * kStageBytes + 256;
  cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  Args a{P, rows, sums, cyc, num_kb, mode, nc, swz};
  const bool cluster = mode == 3 || mode == 4;
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaMemset(sums, 0, (size_t)tiles * nc * 128 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tiles * nc, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster ? nc : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, gather_kernel, tm, a);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(err)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  std::vector<float> hs((size_t)tiles * nc * 128);
  std::vector<long long> hc((size_t)tiles * nc);
  cudaMemcpy(hs.data(), sums, hs.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hc.data(), cyc, hc.size() * 8, cudaMemcpyDeviceToHost);
  // expected checksums (both halves of the consumer warps add into the same row)
  double maxerr = 0;
  for (int b = 0; b < tiles * nc; ++b) {
    const int tile = b / nc;
    for (int r = 0; r < 128; ++r) {
      double acc = 0;
      for (int kb = 0; kb < num_kb; ++kb)
        for (int chunk = 0; chunk < 8; ++chunk) {
          const float* xa = &hP[(size_t)hrows[tile * 256 + r] * LD + kb * 32 + chunk * 4];
          const float* xb = &hP[(size_t)hrows[tile * 256 + 128 + r] * LD + D + kb * 32 + chunk * 4];
          acc += (chunk + 1) * (xa[0] + 2.0 * xa[1] + 3.0 * xa[2] + 4.0 * xa[3]) + (xb[0] + xb[1] + xb[2] + xb[3]);
        }
      const double e = fabs(acc - hs[(size_t)b * 128 + r]) / (fabs(acc) + 1e-9);
      if (e > maxerr) maxerr = e;
    }
  }
  long long cmax = 0; double cmean = 0;
  for (auto c : hc) { cmax = c > cmax ? c : cmax; cmean += (double)c / hc.size(); }
  printf("mode %d tiles %d nc %d box_rows %d swz %d num_kb %d: kernel %.1f us, cycles/k-block mean %.0f max %.0f (%.2f us @1.965GHz), checksum rel err %.2e %s\n",
         mode, tiles, nc, box_rows, swz, num_kb, best * 1e3, cmean / num_kb, (double)cmax / num_kb,
         cmean / num_kb / 1965.0, maxerr, maxerr < 1e-4 ? "OK" : "WRONG DATA");
  return 0;
}
