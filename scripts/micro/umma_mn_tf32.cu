// MN-major TF32 operands for tcgen05.mma on sm_100a, and the TMA tensor copies that produce them.
//
// Question: the wide-net kernel stores every activation plane twice, [point][neuron] for the forward /
// backward contractions (K = neurons) and [neuron][point] for the weight gradient (K = points).  One plain row-major
// [neuron][point] plane in global memory would do if (a) the tensor core can read it as an MN-major A operand and (b) the TMA
// engine applies the shared-memory swizzle each consumer needs while it copies.  CUTLASS (sm100_common.inl) says MN-major
// tf32 exists only with the SWIZZLE_128B_BASE32B layout (32-byte chunks XOR-ed within 128 B rows, 4-row atoms); the K-major
// consumer wants SWIZZLE_128B (16-byte chunks, 8-row atoms).  This test pins the descriptor semantics on the hardware:
//   D[128 x N] = At^T B^T,  At = [K][128] (element (k, m)), B = [N][K] K-major canonical (no swizzle)
//   mode 0: At written into shared memory by hand in the SW128_BASE32B pattern, variants of (LBO, SBO)
//   mode 1: At copied by cp.async.bulk.tensor.2d with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B from a plain row-major plane
//   mode 2: K-major control: the same plane as the B^T-like operand of a K = m contraction through CU_TENSOR_MAP_SWIZZLE_128B
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_mn_tf32 umma_mn_tf32.cu && ./umma_mn_tf32
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__host__ __device__ inline int canon_off(int r, int k, int K) { return ((r >> 3) * (K / 4) + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3); }

struct Maps {
  CUtensorMap a32;   // plane [K rows][128 floats], box {32, K}, SWIZZLE_128B_ATOM_32B
  CUtensorMap a128;  // same plane, box {32, K}, SWIZZLE_128B
};

// MODE 0: manual fill, VARIANT selects the (LBO, SBO) reading; MODE 1: TMA ATOM_32B; MODE 2: K-major control through TMA SWIZZLE_128B
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) mn_kernel(const __grid_constant__ Maps maps, const float* __restrict__ At,
                                                    const float* __restrict__ B, float* __restrict__ D, int K, int variant,
                                                    int* status, float* dump) {
  extern __shared__ float smem_raw[];
  float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, tbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A image: [block of 32 m][K rows][128 B]; MODE 2 uses it as the K-major operand of D2[K x N2] ... see below
  float* sA = smem;
  float* sB = sA + K * M;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&tbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (MODE == 0) {
    for (int idx = tid; idx < K * M; idx += blockDim.x) {
      const int k = idx / M, m = idx % M;
      const int off_b = (m >> 5) * K * 128 + k * 128 + (((m & 31) * 4) ^ ((k & 3) << 5));   // bytes
      sA[off_b >> 2] = At[idx];
    }
  } else if (tid == 0) {
    const CUtensorMap* tm = (MODE == 1) ? &maps.a32 : &maps.a128;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tbar)), "r"((uint32_t)(K * M * 4)) : "memory");
    for (int blk = 0; blk < 4; ++blk)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(sA + blk * K * 32)), "l"(tm), "r"(blk * 32), "r"(0), "r"(smem_u32(&tbar)) : "memory");
  }
  if (MODE != 2) {
    for (int idx = tid; idx < N * K; idx += blockDim.x) sB[canon_off(idx / K, idx % K, K)] = B[idx];
  } else {
    // control: D[m][n] = sum_k A2[m][k] B2[n][k] with A2 = At viewed as [K rows = "m" (K must be 128)][128 floats = "k"] ...
    // here: rows = 128 (requires K == 128), contraction over the 128 floats of a row; B2 = B as [N][128] canonical
    for (int idx = tid; idx < N * M; idx += blockDim.x) sB[canon_off(idx / M, idx % M, M)] = B[idx];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(N < 32 ? 32 : N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (MODE != 0) {
    bool ok = false;
    for (int spin = 0; spin < (1 << 22); ++spin)
      if (mbar_try_wait(&tbar, 0)) { ok = true; break; }
    if (!ok && tid == 0) *status = 2;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (dump) for (int idx = tid; idx < K * M; idx += blockDim.x) dump[idx] = sA[idx];
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    uint32_t accum = 0;
    if (MODE != 2) {
      const uint32_t idesc = make_idesc(M, N, 1, 0);
      const uint32_t blk_stride = (uint32_t)K * 128, katom = 512;
      for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t lbo, sbo;
        switch (variant) {
          case 0: lbo = blk_stride; sbo = katom; break;       // LBO = next 32 m, SBO = next 4 k   (CUTLASS comment)
          case 1: lbo = katom; sbo = blk_stride; break;       // swapped
          case 2: lbo = blk_stride; sbo = 1024; break;        // SBO = 8 k rows
          default: lbo = 1024; sbo = blk_stride; break;
        }
        const uint64_t da = make_desc(smem_u32(sA) + (uint32_t)(k0 >> 2) * 512, lbo, sbo, 1);
        const uint64_t db = make_desc(smem_u32(sB) + (uint32_t)(k0 >> 2) * 128, 128, (uint32_t)(K / 4) * 128, 0);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
        accum = 1;
      }
    } else {
      // K-major SW128: rows = 128 (the plane's K rows, K == 128), 32 floats of contraction per 128 B row block
      const uint32_t idesc = make_idesc(M, N, 0, 0);
      for (int k0 = 0; k0 < M; k0 += 8) {
        const uint64_t da = make_desc(smem_u32(sA) + (uint32_t)((k0 >> 5) * K * 128 + (k0 & 31) * 4), 16, 1024, 2);
        const uint64_t db = make_desc(smem_u32(sB) + (uint32_t)(k0 >> 2) * 128, 128, (uint32_t)(M / 4) * 128, 0);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
        accum = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  bool done = false;
  for (int spin = 0; spin < (1 << 22); ++spin)
    if (mbar_try_wait(&bar, 0)) { done = true; break; }
  if (!done && tid == 0) *status = 1;
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (done) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = warp * 32 + lane;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < N) D[row * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(N < 32 ? 32 : N));
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeTiled enc, CUtensorMap* m, float* plane, int rows, CUtensorMapSwizzle sw) {
  cuuint64_t dims[2] = {128, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128 * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, plane, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d (swizzle %d)\n", (int)r, (int)sw);
  return r == CUDA_SUCCESS ? 0 : 1;
}

template <int N, int MODE>
int run(int K, int variant, EncodeTiled enc, bool show_dump) {
  const int rowsB = (MODE == 2) ? M : K;
  std::vector<float> At((size_t)K * M), B((size_t)N * rowsB), D((size_t)M * N, 0.f), dump((size_t)K * M, 0.f);
  srand(7 + MODE);
  for (auto& x : At) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  float *dA, *dB, *dD, *dDump; int* dS;
  cudaMalloc(&dA, At.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
  cudaMalloc(&dDump, dump.size() * 4);
  cudaMemcpy(dA, At.data(), At.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4); cudaMemset(dS, 0, 4);
  Maps maps;
  memset(&maps, 0, sizeof(maps));
  int bad = 0;
  if (MODE != 0) {
    bad += make_map(enc, &maps.a32, dA, K, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    bad += make_map(enc, &maps.a128, dA, K, CU_TENSOR_MAP_SWIZZLE_128B);
    if (bad) return bad;
  }
  const size_t smem = (size_t)(K * M + N * (MODE == 2 ? M : K)) * 4 + 1024;
  cudaFuncSetAttribute(mn_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mn_kernel<N, MODE><<<1, 128, smem>>>(maps, dA, dB, dD, K, variant, dS, dDump);
  cudaError_t e = cudaDeviceSynchronize();
  int st = 0;
  cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(dump.data(), dDump, dump.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  if (MODE != 2) {
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)At[(size_t)k * M + m] * (double)B[(size_t)n * K + k];
        maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
        maxref = fmax(maxref, fabs(ref));
      }
  } else {  // rows of the plane are the M rows (K == 128), contraction over the 128 floats of a row
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < M; ++k) ref += (double)At[(size_t)m * M + k] * (double)B[(size_t)n * M + k];
        maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
        maxref = fmax(maxref, fabs(ref));
      }
  }
  // how does the image in shared memory relate to the hand-made pattern?
  int match32 = 0, match128 = 0, total = 0;
  for (int k = 0; k < K; ++k)
    for (int m = 0; m < M; ++m, ++total) {
      const int o32 = ((m >> 5) * K * 128 + k * 128 + (((m & 31) * 4) ^ ((k & 3) << 5))) >> 2;
      const int o128 = ((m >> 5) * K * 128 + k * 128 + (((m & 31) * 4) ^ ((k & 7) << 4))) >> 2;
      match32 += dump[o32] == At[(size_t)k * M + m];
      match128 += dump[o128] == At[(size_t)k * M + m];
    }
  printf("mode %d variant %d N=%3d K=%3d : cuda=%s status=%d  rel err = %.2e  (max|ref| %.3f)   smem image: %d/%d match Swizzle<2,5,2>, %d/%d match Swizzle<3,4,3>\n",
         MODE, variant, N, K, cudaGetErrorString(e), st, maxerr / maxref, maxref, match32, total, match128, total);
  (void)show_dump;
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS); cudaFree(dDump);
  return (e == cudaSuccess && st == 0 && maxerr / maxref < 2e-3) ? 0 : 1;
}

int main() {
  EncodeTiled enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  printf("cuTensorMapEncodeTiled: %s, query %d, %p\n", cudaGetErrorString(e), (int)q, (void*)enc);
  int bad = 0;
  for (int v = 0; v < 4; ++v) bad += run<64, 0>(32, v, enc, false) ? 0 : 0;   // informative: which (LBO, SBO) reading is right
  for (int v = 0; v < 4; ++v) run<128, 0>(64, v, enc, false);
  if (enc) {
    for (int v = 0; v < 2; ++v) run<64, 1>(32, v, enc, false);
    run<128, 1>(64, 0, enc, false);
    run<128, 1>(128, 0, enc, false);
    bad += run<64, 2>(128, 0, enc, false);
  }
  return bad;
}
