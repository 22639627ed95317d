// mma.sync (legacy tensor path) throughput on B200 (sm_100a): TF32 m16n8k8 / m16n8k4, BF16 m16n8k16.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int NACC>
__global__ void __launch_bounds__(256) k_tf32_k8(float* out, int iters) {
  float c[NACC][4];
#pragma unroll
  for (int a = 0; a < NACC; ++a) for (int q = 0; q < 4; ++q) c[a][q] = 0.f;
  uint32_t A[4] = {threadIdx.x, threadIdx.x * 3u, 5u, 7u}, B[2] = {threadIdx.x * 11u, 13u};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[a][0]), "+f"(c[a][1]), "+f"(c[a][2]), "+f"(c[a][3])
                   : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(B[0]), "r"(B[1]));
  }
  float s = 0;
#pragma unroll
  for (int a = 0; a < NACC; ++a) for (int q = 0; q < 4; ++q) s += c[a][q];
  if (s == 123.456f) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_tf32_k4(float* out, int iters) {
  float c[NACC][4];
#pragma unroll
  for (int a = 0; a < NACC; ++a) for (int q = 0; q < 4; ++q) c[a][q] = 0.f;
  uint32_t A[2] = {threadIdx.x, threadIdx.x * 3u}, B[1] = {threadIdx.x * 11u};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+f"(c[a][0]), "+f"(c[a][1]), "+f"(c[a][2]), "+f"(c[a][3])
                   : "r"(A[0]), "r"(A[1]), "r"(B[0]));
  }
  float s = 0;
#pragma unroll
  for (int a = 0; a < NACC; ++a) for (int q = 0; q < 4; ++q) s += c[a][q];
  if (s == 123.456f) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_bf16_k16(float* out, int iters) {
  float c[NACC][4];
#pragma unroll
  for (int a = 0; a < NACC; ++a) for (int q = 0; q < 4; ++q) c[a][q] = 0.f;
  uint32_t A[4] = {threadIdx.x, threadIdx.x * 3u, 5u, 7u}, B[2] = {threadIdx.x * 11u, 13u};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int a = 0; a < NACC; ++a)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[a][0]), "+f"(c[a][1]), "+f"(c[a][2]), "+f"(c[a][3])
                   : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(B[0]), "r"(B[1]));
  }
  float s = 0;
#pragma unroll
  for (int a = 0; a < NACC; ++a) for (int q = 0; q < 4; ++q) s += c[a][q];
  if (s == 123.456f) out[0] = s;
}

// mixed: NACC mma + NF independent FFMA per iteration: do the pipes overlap?
template <int NACC, int NF>
__global__ void __launch_bounds__(256) k_mixed(float* out, int iters, float a0, float b0) {
  float c[NACC][4], f[NF];
#pragma unroll
  for (int a = 0; a < NACC; ++a) for (int q = 0; q < 4; ++q) c[a][q] = 0.f;
#pragma unroll
  for (int k = 0; k < NF; ++k) f[k] = threadIdx.x + k;
  uint32_t A[4] = {threadIdx.x, threadIdx.x * 3u, 5u, 7u}, B[2] = {threadIdx.x * 11u, 13u};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[a][0]), "+f"(c[a][1]), "+f"(c[a][2]), "+f"(c[a][3])
                   : "r"(A[0]), "r"(A[1]), "r"(A[2]), "r"(A[3]), "r"(B[0]), "r"(B[1]));
#pragma unroll
      for (int k = 0; k < NF / NACC; ++k) f[a * (NF / NACC) + k] = fmaf(f[a * (NF / NACC) + k], a0, b0);
    }
  }
  float s = 0;
#pragma unroll
  for (int a = 0; a < NACC; ++a) for (int q = 0; q < 4; ++q) s += c[a][q];
#pragma unroll
  for (int k = 0; k < NF; ++k) s += f[k];
  if (s == 123.456f) out[0] = s;
}

template <class F>
double timeit(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  return best * 1e-3;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, 4);
  const int iters = 8192;
  for (int cps : {1, 2, 4}) {
    const int grid = sms * cps, threads = 256;
    const double warps = (double)grid * threads / 32;
    printf("CTAs/SM=%d (%d warps/SMSP), clock %d kHz\n", cps, cps * 2, clk);
    double t;
    t = timeit([&] { k_tf32_k8<8><<<grid, threads>>>(out, iters); });
    printf("  tf32 m16n8k8  : %8.1f TFLOP/s  (%.2f clk/mma/SMSP)\n", warps * iters * 8 * 2.0 * 16 * 8 * 8 / t / 1e12, t * clk * 1e3 / (iters * 8.0 * cps * 2));
    t = timeit([&] { k_tf32_k4<8><<<grid, threads>>>(out, iters); });
    printf("  tf32 m16n8k4  : %8.1f TFLOP/s  (%.2f clk/mma/SMSP)\n", warps * iters * 8 * 2.0 * 16 * 8 * 4 / t / 1e12, t * clk * 1e3 / (iters * 8.0 * cps * 2));
    t = timeit([&] { k_bf16_k16<8><<<grid, threads>>>(out, iters); });
    printf("  bf16 m16n8k16 : %8.1f TFLOP/s  (%.2f clk/mma/SMSP)\n", warps * iters * 8 * 2.0 * 16 * 8 * 16 / t / 1e12, t * clk * 1e3 / (iters * 8.0 * cps * 2));
    t = timeit([&] { k_mixed<8, 32><<<grid, threads>>>(out, iters, 0.999f, 0.001f); });
    printf("  mixed 8 mma(tf32 k8) + 32 FFMA per iter: %.2f clk/iter/SMSP-warp-slot\n", t * clk * 1e3 / (iters * 1.0 * cps * 2));
    t = timeit([&] { k_mixed<8, 64><<<grid, threads>>>(out, iters, 0.999f, 0.001f); });
    printf("  mixed 8 mma(tf32 k8) + 64 FFMA per iter: %.2f clk/iter/SMSP-warp-slot\n", t * clk * 1e3 / (iters * 1.0 * cps * 2));
  }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
