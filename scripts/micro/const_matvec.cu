// Does the F/B matvec run faster when the weight rows come through the constant/uniform path
// (LDCU -> uniform register operand of FFMA2) instead of broadcast LDS.128?  Working set = NLAY x 400 floats,
// warps de-phased over the layers like the real kernel.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__constant__ float cW[16384];

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  uint64_t dd = *reinterpret_cast<uint64_t*>(&d);
  const uint64_t aa = *reinterpret_cast<const uint64_t*>(&a), bb = *reinterpret_cast<const uint64_t*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}

template <int MODE>  // 0 = LDS weights, 1 = constant weights via uniform registers (LDCU), 2 = constant weights via per-lane LDC
__global__ void __launch_bounds__(256, 1) k_mv(float* out, const float* __restrict__ in, int iters, int nlay, int dephase) {
  extern __shared__ __align__(16) float sm[];
  float* sw = sm;               // [nlay][400]
  float* sx = sm + 16 * 400;
  for (int k = threadIdx.x; k < nlay * 400; k += 256) sw[k] = in[k];
  for (int k = threadIdx.x; k < 256 * 84; k += 256) sx[k] = in[6400 + k];
  __syncthreads();
  float2 acc[4][10];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int j = 0; j < 10; ++j) acc[s][j] = make_float2(0.f, 0.f);
  const float* xrow = sx + threadIdx.x * 84;
  int lay = (MODE == 2) ? ((threadIdx.x >> 5) * 3) % nlay : 0;
  if (dephase) {  // same layer sequence, shifted in time per warp
    const long long t0 = clock64();
    while (clock64() - t0 < (threadIdx.x >> 5) * 777) {}
  }
  for (int it = 0; it < iters; ++it) {
    const float* W = (MODE == 0 ? sw : cW) + lay * 400;
    lay = (lay + 1 == nlay) ? 0 : lay + 1;
#pragma unroll 2
    for (int i = 0; i < 20; ++i) {
      const float4 xv = *reinterpret_cast<const float4*>(xrow + 4 * i);
      const float2 x0 = make_float2(xv.x, xv.x), x1 = make_float2(xv.y, xv.y), x2 = make_float2(xv.z, xv.z), x3 = make_float2(xv.w, xv.w);
      float2 w[10];
      if (MODE == 0) {
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const float4 t = *reinterpret_cast<const float4*>(W + i * 20 + 4 * q);
          w[2 * q] = make_float2(t.x, t.y);
          w[2 * q + 1] = make_float2(t.z, t.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 10; ++j) w[j] = make_float2(W[i * 20 + 2 * j], W[i * 20 + 2 * j + 1]);
      }
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        ffma2(acc[0][j], x0, w[j]);
        ffma2(acc[1][j], x1, w[j]);
        ffma2(acc[2][j], x2, w[j]);
        ffma2(acc[3][j], x3, w[j]);
      }
    }
    *reinterpret_cast<float4*>(sx + threadIdx.x * 84 + 4 * (it % 20)) = make_float4(acc[0][1].x * 1e-30f, acc[1][2].y * 1e-30f, acc[2][3].x * 1e-30f, acc[3][4].y * 1e-30f);
  }
  float sum = 0;
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int j = 0; j < 10; ++j) sum += acc[s][j].x + acc[s][j].y;
  if (sum == 123.456f) out[0] = sum;
}

template <class F>
double timeit(F launch, double flops) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  return best;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float *out, *in; cudaMalloc(&out, 4); cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20);
  const int threads = 256, iters = 2048;
  const size_t smem = (6400 + 256 * 84) * 4;
  cudaFuncSetAttribute(k_mv<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_mv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_mv<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = sms;
  const double fl = 2.0 * grid * threads * (double)iters * 1600;
  for (int nlay : {1, 4, 8, 16})
    for (int dp : {0, 1}) {
      printf("layers in flight %2d (%5.1f KB) dephase=%d :  LDS weights %6.2f   LDCU/uniform %6.2f   per-lane LDC %6.2f TFLOP/s\n", nlay, nlay * 1.6, dp,
             timeit([&] { k_mv<0><<<grid, threads, smem>>>(out, in, iters, nlay, dp); }, fl),
             timeit([&] { k_mv<1><<<grid, threads, smem>>>(out, in, iters, nlay, dp); }, fl),
             timeit([&] { k_mv<2><<<grid, threads, smem>>>(out, in, iters, nlay, dp); }, fl));
    }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
