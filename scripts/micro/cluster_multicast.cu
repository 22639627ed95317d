// Operand staging of the generic kernel's cluster-per-tile mode, two ways (round-2 groundwork):
//   A. today: every CTA of the cluster fetches the WHOLE tile from L2 with cp.async (cs x the traffic);
//   B. each rank fetches 1/cs of the tile with ONE cp.async.bulk and the hardware multicasts it into the same
//      shared-memory offset of every CTA of the cluster; completion is tracked per CTA by an mbarrier (complete_tx bytes).
// The program checks B's result against the source and prints the time per staged tile of both, for the tile sizes of the
// Euler net (activations 3 x 200 x 32 floats = 76.8 KB; weight-gradient operands of one stream 2 x 32 x 200 floats = 51.2 KB).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o cluster_multicast cluster_multicast.cu && ./cluster_multicast
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));                    \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a lost completion must not hang the GPU box (returns false after ~1 s)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (long long spin = 0; spin < (1ll << 26); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
// one contiguous slice global -> the same shared-memory offset in every CTA named by `mask`, bytes reported to `bar` there
__device__ __forceinline__ void bulk_multicast(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// mode 0: every CTA stages the whole tile itself (cp.async, 16 B per thread and instruction)
// mode 1: rank r stages slice r and multicasts it to the cluster
template <int MODE>
__global__ void __launch_bounds__(256, 1) stage_kernel(const float* __restrict__ src, int tile_floats, int iters, float* __restrict__ out,
                                                         int* __restrict__ bad) {
  extern __shared__ __align__(128) float tile[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t cs;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(cs));
  const uint32_t rank = cluster_rank();
  const int cluster_id = blockIdx.x / cs;
  const float* my = src + (size_t)cluster_id * tile_floats;  // one tile per cluster, shared by its CTAs
  if (MODE == 1 && threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();
  float acc = 0.f;
  const uint32_t tile_bytes = (uint32_t)tile_floats * 4u;
  const uint32_t slice = tile_bytes / cs;  // the host picks sizes that divide evenly into 16 B multiples
  for (int it = 0; it < iters; ++it) {
    if (MODE == 1 && (*reinterpret_cast<volatile int*>(bad) >> 20)) break;  // a completion was lost: every CTA stops at a round boundary
    if (MODE == 0) {
      for (int i = threadIdx.x * 4; i < tile_floats; i += blockDim.x * 4) cp_async16(tile + i, my + i);
      asm volatile("cp.async.wait_all;\n" ::: "memory");
      __syncthreads();
    } else {
      if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, tile_bytes);  // all cs slices land here
        bulk_multicast(reinterpret_cast<char*>(tile) + rank * slice, reinterpret_cast<const char*>(my) + rank * slice, slice, &bar,
                       (uint16_t)((1u << cs) - 1u));
      }
      if (!mbar_wait(&bar, it & 1) && threadIdx.x == 0) atomicOr(bad, 1 << 20);  // keep going to the barrier below, stop next round
    }
    // consume: something that depends on every word (and, on the last round, a correctness check)
    for (int i = threadIdx.x; i < tile_floats; i += blockDim.x) acc += tile[i];
    if (it == iters - 1) {
      int wrong = 0;
      for (int i = threadIdx.x; i < tile_floats; i += blockDim.x) wrong += (tile[i] != my[i]);
      if (wrong) atomicAdd(bad, wrong > 1000 ? 1000 : wrong);
    }
    // nobody may refill a buffer that a peer is still reading: the multicast writes into the PEERS' shared memory
    if (MODE == 1) cluster_sync();
    else __syncthreads();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
static float run(const float* src, int tile_floats, int clusters, int cs, int iters, float* out, int* bad) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * cs);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = (size_t)tile_floats * 4;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaFuncSetAttribute(stage_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_floats * 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaLaunchKernelEx(&cfg, stage_kernel<MODE>, src, tile_floats, 2, out, bad));  // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  CK(cudaLaunchKernelEx(&cfg, stage_kernel<MODE>, src, tile_floats, iters, out, bad));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms * 1e3f / iters;  // us per staged tile (all clusters in parallel)
}

int main() {
  const int cases[][2] = {{3 * 200 * 32, 3}, {3 * 200 * 32, 4}, {2 * 32 * 200, 4}, {3 * 200 * 32, 6}};  // {tile floats, cluster size}
  for (auto& c : cases) {
    const int tile_floats = c[0], cs = c[1];
    if ((tile_floats * 4 / cs) % 16 != 0) {
      printf("tile %d floats does not split into 16 B multiples over %d ranks, skipped\n", tile_floats, cs);
      continue;
    }
    const int clusters = (cs == 3) ? 39 : (cs == 4 ? 32 : 20);
    float *src, *out;
    int* bad;
    CK(cudaMalloc(&src, (size_t)clusters * tile_floats * 4));
    CK(cudaMalloc(&out, (size_t)clusters * cs * 256 * 4));
    CK(cudaMalloc(&bad, 4));
    CK(cudaMemset(bad, 0, 4));
    float* h = (float*)malloc((size_t)clusters * tile_floats * 4);
    for (size_t i = 0; i < (size_t)clusters * tile_floats; ++i) h[i] = (float)((i * 2654435761u) >> 8) * 1e-7f;
    CK(cudaMemcpy(src, h, (size_t)clusters * tile_floats * 4, cudaMemcpyHostToDevice));
    const float a = run<0>(src, tile_floats, clusters, cs, 200, out, bad);
    const float b = run<1>(src, tile_floats, clusters, cs, 200, out, bad);
    int nbad = 0;
    CK(cudaMemcpy(&nbad, bad, 4, cudaMemcpyDeviceToHost));
    printf("tile %6.1f KB, %2d clusters x %d CTAs: every CTA fetches all %7.2f us/tile | 1/%d each + multicast %7.2f us/tile | mismatches %d%s\n",
           tile_floats * 4 / 1024.0, clusters, cs, a, cs, b, nbad & 0xFFFFF, (nbad >> 20) ? "  (mbarrier wait TIMED OUT)" : "");
    free(h);
    CK(cudaFree(src));
    CK(cudaFree(out));
    CK(cudaFree(bad));
  }
  return 0;
}
