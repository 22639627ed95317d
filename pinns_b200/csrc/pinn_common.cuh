// Shared device/host definitions for libpinn_b200 (sm_100a only).
//
// Arithmetic follows SURVEY.md appendix A (Taylor-forward + one reverse sweep of the
// reference graph INF-L2:96-120 / AB-ADMM:170-180 / EUL:176-198), all in float32 like
// the reference's tf.float32 placeholders and variables (INF-L2:58-63,:85,:94).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pinn_b200.h"

#define PINN_TILE 32            // collocation points per tile in the generic kernel (= warp width)
#define PINN_NSUMS 8            // scalar partial sums appended to the packed vector
#define PINN_SUM_DATA 0         // data-term loss (already weighted / normalised)
#define PINN_SUM_RES 1          // separable residual-term loss (already normalised)
#define PINN_SUM_ABSF 2         // sum |f|            (V3 needs the global value before seeding)
#define PINN_SUM_MISFIT 3       // sum |f - z|        (admm_misfit, AB-ADMM:60)
#define PINN_SUM_F2 4           // sum f^2

struct NetDesc {
  int L;                         // number of weight layers = len(layers)-1
  int n[PINN_MAX_LAYERS];        // layer widths n[0..L]
  int np[PINN_MAX_LAYERS];       // widths rounded up to a multiple of 8
  int w_off[PINN_MAX_LAYERS];    // offset of W_l in the flat parameter vector
  int b_off[PINN_MAX_LAYERS];    // offset of b_l
  int wp_off[PINN_MAX_LAYERS];   // offset of the zero-padded copy  Wp_l [n_in][np_out]
  int wt_off[PINN_MAX_LAYERS];   // offset of the padded transpose  WT_l [n_out][np_in]
  int P;                         // weights + biases
  int npmax;                     // max padded hidden/output width
  int n_out;                     // n[L]
  int n_res;                     // residuals per point: 1 Burgers, 3 Euler
  int pde;
  float lbx, lbt, spanx, spant;  // float32(lb), float32(ub - lb)  (INF-L2:99: the span is formed in float64 numpy first)
};

// How the residual adjoint f_bar and the per-point loss are formed (appendix A.3):
//   f_bar = cA*f + cB*sign(f) + cC*(f - z) + cD*gamma
struct LossCoef {
  int loss;       // PINN_LOSS_*
  float cA, cB, cC, cD;
  float rho;      // ADMM penalty
  float inv_nf;   // 1 / N_f of the whole job
};

// Accurate float32 tanh.  tanh.approx (2^-11) is far too coarse for 1e-5 parity on
// third-derivative chains; this form has ~1.2e-7 absolute error:
//   tanh(x) = 1 - 2/(exp(2x)+1), exp via ex2.approx (2 ulp), reciprocal refined by one Newton step.
__device__ __forceinline__ float pinn_tanh_fast(float x) {
  float e = exp2f(x * 2.885390081777927f);  // 2*log2(e); exp2f -> ex2.approx.ftz under -use_fast_math, else accurate
  float d = e + 1.0f;
  float r = __frcp_rn(d);
  return fmaf(-2.0f, r, 1.0f);
}

__device__ __forceinline__ float pinn_tanh(float x) { return tanhf(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox4x32-10 (Salmon et al. 2011), counter-based: sample i depends only on (seed, i).
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
