// Minimal hand-written tcgen05 (UMMA) TF32 GEMM on sm_100a, one CTA:
//   D[128 x N] (fp32, TMEM) = A[128 x K] * B[N x K]^T,  A and B K-major in shared memory (no swizzle), K multiple of 8.
// Optional 3xTF32 split (hi*hi + hi*lo + lo*hi) for fp32-grade accuracy.  Groundwork for the wide-net PINN kernel
// (DESIGN.md section 7); checked against a double-precision CPU product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_gemm tcgen05_gemm.cu && ./tcgen05_gemm
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4, [16,30) leading byte offset>>4 (core matrix to core matrix along K),
//   [32,46) stride byte offset>>4 (8-row group to 8-row group along M/N), [46,48) version = 1, [61,64) layout = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @[4,6), a/b format TF32 = 2 @[7,10)/[10,13),
// K-major A and B (bits 15, 16 = 0), n_dim = N>>3 @[17,23), m_dim = M>>4 @[24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// core-matrix (8 rows x 16 B) interleaved K-major layout: element (r, k) of an [R x K] fp32 matrix
__host__ __device__ inline int canon_off(int r, int k, int K) {
  const int kc = k >> 2, rc = r >> 3;                 // core matrix column / row
  return (rc * (K / 4) + kc) * 32 + (r & 7) * 4 + (k & 3);   // floats; LBO = 128 B, SBO = (K/4)*128 B
}

template <int N, bool SPLIT>
__global__ void __launch_bounds__(128, 1) gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                      int K, int* status) {
  extern __shared__ __align__(128) float smem[];
  float* sAh = smem;                 // [128 x K] canonical
  float* sAl = sAh + M * K;
  float* sBh = sAl + M * K;          // [N x K] canonical
  float* sBl = sBh + N * K;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int idx = tid; idx < M * K; idx += blockDim.x) {
    const int r = idx / K, k = idx % K;
    const float x = A[idx];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    sAh[canon_off(r, k, K)] = SPLIT ? hi : x;
    sAl[canon_off(r, k, K)] = x - hi;
  }
  for (int idx = tid; idx < N * K; idx += blockDim.x) {
    const int r = idx / K, k = idx % K;
    const float x = B[idx];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    sBh[canon_off(r, k, K)] = SPLIT ? hi : x;
    sBl[canon_off(r, k, K)] = x - hi;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(N < 32 ? 32 : N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // generic-proxy writes to smem must be visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;

  if (tid == 0) {
    const uint32_t idesc = make_idesc(M, N);
    const uint32_t lbo = 128, sbo = (K / 4) * 128;
    uint32_t accum = 0;
    const int npass = SPLIT ? 3 : 1;
    for (int pass = 0; pass < npass; ++pass) {
      const float* pa = (pass == 2) ? sAl : sAh;   // hi*hi, hi*lo, lo*hi
      const float* pb = (pass == 1) ? sBl : sBh;
      for (int k0 = 0; k0 < K; k0 += 8) {          // one MMA = K 8 (32 B) = 2 core matrices along K
        const uint64_t da = make_desc(smem_u32(pa) + (k0 / 4) * 128, lbo, sbo);
        const uint64_t db = make_desc(smem_u32(pb) + (k0 / 4) * 128, lbo, sbo);
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum)
            : "memory");
        accum = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMAs (bounded spin: report instead of hanging the GPU)
  bool done = false;
  for (int spin = 0; spin < (1 << 22); ++spin) {
    if (mbar_try_wait(&bar, 0)) { done = true; break; }
  }
  if (!done) {
    if (tid == 0) *status = 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (done) {
    // warp w owns TMEM lanes [32w, 32w+32) = rows of D; 32 columns per tcgen05.ld
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

// MN-major operands: At is [K x 128] (element (k, m)), Bt is [K x N] (element (k, n)), staged as core matrices of
// 8 k-rows x 16 B (4 consecutive m / n), i.e. exactly the tile a K-major consumer of the transposed data would read.
// VARIANT selects how LBO / SBO are interpreted for the MN-major descriptor.
template <int N, int VARIANT>
__global__ void __launch_bounds__(128, 1) gemm_mn_kernel(const float* __restrict__ At, const float* __restrict__ Bt,
                                                         float* __restrict__ D, int K, int* status) {
  extern __shared__ __align__(128) float smem[];
  float* sA = smem;            // [K x 128] canonical with rows = k
  float* sB = sA + K * M;      // [K x N]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < K * M; idx += blockDim.x) sA[canon_off(idx / M, idx % M, M)] = At[idx];
  for (int idx = tid; idx < K * N; idx += blockDim.x) sB[canon_off(idx / N, idx % N, N)] = Bt[idx];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(N < 32 ? 32 : N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(M, N) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
    // tile geometry: core matrices (8 k x 4 mn) are 128 B; next core matrix along MN: +128 B; next 8-k group: +(MN/4)*128 B
    const uint32_t a_mn = 128, a_k = (M / 4) * 128, b_mn = 128, b_k = (N / 4) * 128;
    uint32_t accum = 0;
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint64_t da, db;
      if (VARIANT == 0) {  // LBO = MN stride, SBO = K stride
        da = make_desc(smem_u32(sA) + (k0 / 8) * a_k, a_mn, a_k);
        db = make_desc(smem_u32(sB) + (k0 / 8) * b_k, b_mn, b_k);
      } else {             // LBO = K stride, SBO = MN stride
        da = make_desc(smem_u32(sA) + (k0 / 8) * a_k, a_k, a_mn);
        db = make_desc(smem_u32(sB) + (k0 / 8) * b_k, b_k, b_mn);
      }
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum)
          : "memory");
      accum = 1;
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  bool done = false;
  for (int spin = 0; spin < (1 << 22); ++spin) {
    if (mbar_try_wait(&bar, 0)) { done = true; break; }
  }
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


// ---- SWIZZLE_128B operands: rows of 128 B (32 tf32), 8-row atoms of 1024 B, 16-byte chunk index XOR row index ----
// The same bytes are a K-major operand (rows = M/N, the 128 B run along K) and an MN-major one (rows = K, the 128 B run
// along M/N): what a [point][neuron] plane needs to feed both the forward contraction (K = neurons) and the weight
// gradient (K = points) of the wide-net kernel.  MODE bit 0: A MN-major, bit 1: B MN-major.
__host__ __device__ inline int sw128_off(int row, int col, int rows_total) {   // floats; col < 32 * nblocks
  const int blk = col >> 5, c = col & 31;
  return blk * rows_total * 32 + (row >> 3) * 256 + (row & 7) * 32 + ((((c >> 2) ^ (row & 7)) << 2) | (c & 3));
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return make_desc(saddr, lbo, sbo) | ((uint64_t)2 << 61);
}
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) gemm_sw128_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ D, int K, int* status) {
  extern __shared__ float smem_raw[];
  float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // MODE 0: A is [128 x K] (row = m, col = k), B is [N x K];  MODE 1: A is [K x 128] (row = k, col = m), B is [K x N]
  constexpr bool AT = (MODE & 1) != 0, BT = (MODE & 2) != 0;
  const int a_rows = AT ? K : M, a_cols = AT ? M : K;
  const int b_rows = BT ? K : N, b_cols = BT ? N : K;
  float* sA = smem;
  float* sB = sA + a_rows * a_cols;
  for (int idx = tid; idx < a_rows * a_cols; idx += blockDim.x) sA[sw128_off(idx / a_cols, idx % a_cols, a_rows)] = A[idx];
  for (int idx = tid; idx < b_rows * b_cols; idx += blockDim.x) sB[sw128_off(idx / b_cols, idx % b_cols, b_rows)] = B[idx];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(N < 32 ? 32 : N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    uint32_t idesc = make_idesc(M, N);
    if (AT) idesc |= (1u << 15);
    if (BT) idesc |= (1u << 16);
    uint32_t accum = 0;
    for (int k0 = 0; k0 < K; k0 += 8) {
      // K-major: block of 32 k = one slab of rows x 128 B; 8 k = 32 B inside the row; SBO = 1024 B
      // MN-major: 8 k = one 1024 B atom; LBO = next block of 32 m/n = K x 128 B; SBO = 1024 B
      const uint64_t da = AT ? make_desc_sw128(smem_u32(sA) + (uint32_t)(k0 >> 3) * 1024, (uint32_t)K * 128, 1024)
                             : make_desc_sw128(smem_u32(sA) + (uint32_t)((k0 >> 5) * a_rows * 128 + (k0 & 31) * 4), 16, 1024);
      const uint64_t db = BT ? make_desc_sw128(smem_u32(sB) + (uint32_t)(k0 >> 3) * 1024, (uint32_t)K * 128, 1024)
                             : make_desc_sw128(smem_u32(sB) + (uint32_t)((k0 >> 5) * b_rows * 128 + (k0 & 31) * 4), 16, 1024);
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum)
          : "memory");
      accum = 1;
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  bool done = false;
  for (int spin = 0; spin < (1 << 22); ++spin) {
    if (mbar_try_wait(&bar, 0)) { done = true; break; }
  }
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

template <int N, int MODE>
int run_sw128(int K) {
  std::vector<float> A(M * K), B(N * K), D(M * N, 0.f);
  srand(3 + MODE);
  for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  float *dA, *dB, *dD; int* dS;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4); cudaMemset(dS, 0, 4);
  const size_t smem = (size_t)(M * K + N * K) * 4 + 1024;
  cudaFuncSetAttribute(gemm_sw128_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gemm_sw128_kernel<N, MODE><<<1, 128, smem>>>(dA, dB, dD, K, dS);
  cudaError_t e = cudaDeviceSynchronize();
  int st = 0;
  cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k)
        ref += (double)((MODE & 1) ? A[k * M + m] : A[m * K + k]) * (double)((MODE & 2) ? B[k * N + n] : B[n * K + k]);
      maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("SW128 A %s B %s N=%3d K=%3d : cuda=%s status=%d  max|err|=%.3e  max|ref|=%.3f  rel=%.2e   D[0..2] = %.4f %.4f %.4f\n",
         (MODE & 1) ? "MN" : "K ", (MODE & 2) ? "MN" : "K ", N, K, cudaGetErrorString(e), st, maxerr, maxref, maxerr / maxref, D[0], D[1], D[2]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return (e == cudaSuccess && st == 0) ? 0 : 1;
}

template <int N, int VARIANT>
int run_mn(int K) {
  std::vector<float> A(K * M), B(K * N), D(M * N, 0.f);
  srand(2);
  for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  float *dA, *dB, *dD; int* dS;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4); cudaMemset(dS, 0, 4);
  const size_t smem = (size_t)(K * M + K * N) * 4;
  cudaFuncSetAttribute(gemm_mn_kernel<N, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gemm_mn_kernel<N, VARIANT><<<1, 128, smem>>>(dA, dB, dD, K, dS);
  cudaError_t e = cudaDeviceSynchronize();
  int st = 0;
  cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[k * M + m] * (double)B[k * N + n];
      maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("MN-major N=%3d K=%3d variant=%d : cuda=%s status=%d  max|err|=%.3e  max|ref|=%.3f  rel=%.2e\n", N, K, VARIANT,
         cudaGetErrorString(e), st, maxerr, maxref, maxerr / maxref);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return (e == cudaSuccess && st == 0) ? 0 : 1;
}

template <int N, bool SPLIT>
int run(int K) {
  std::vector<float> A(M * K), B(N * K), D(M * N, 0.f);
  srand(1);
  for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  float *dA, *dB, *dD; int* dS;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4); cudaMemset(dS, 0, 4);
  const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4;
  cudaFuncSetAttribute(gemm_kernel<N, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gemm_kernel<N, SPLIT><<<1, 128, smem>>>(dA, dB, dD, K, dS);
  cudaError_t e = cudaDeviceSynchronize();
  int st = 0;
  cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)B[n * K + k];
      maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("N=%3d K=%3d split=%d : cuda=%s status=%d  max|err|=%.3e  max|ref|=%.3f  rel=%.2e\n", N, K, (int)SPLIT,
         cudaGetErrorString(e), st, maxerr, maxref, maxerr / maxref);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return (e == cudaSuccess && st == 0) ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<32, false>(32);
  bad += run<32, true>(32);
  bad += run<128, false>(64);
  bad += run<128, true>(48);
  bad += run_mn<64, 0>(32);
  bad += run_mn<64, 1>(32);
  bad += run_mn<128, 0>(64);
  bad += run_mn<128, 1>(64);
  bad += run<64, true>(24);
  bad += run_sw128<64, 0>(32);
  bad += run_sw128<128, 0>(64);
  bad += run_sw128<64, 1>(32);
  bad += run_sw128<64, 2>(32);
  bad += run_sw128<64, 3>(32);
  bad += run_sw128<128, 3>(64);
  return bad;
}
