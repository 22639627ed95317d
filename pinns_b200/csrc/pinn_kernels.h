// Internal kernel-launch interface of libpinn_b200 (not part of the C ABI).
#pragma once
#include "pinn_common.cuh"

#define GEN_MODE_TRAIN 0    // forward + residual + reverse sweep
#define GEN_MODE_FORWARD 1  // forward + residual only (predict, loss value, ADMM updates, V3 pass 1)

struct ScratchDesc {
  int in0;                    // S x 8 x T  Taylor seeds of the input layer
  int hid[PINN_MAX_LAYERS];   // per hidden layer: (2S-1) x np x T   a | Z_x Z_t Z_xx | H_x H_t H_xx
  int Y;                      // S x 8 x T  head outputs per stream
  int zb[2];                  // S x npmax x T  adjoint ping-pong
  // point-major copies ([stream][point][neuron]) of the weight-gradient operands
  int in0T;                   // S x T x 8
  int hidT[PINN_MAX_LAYERS];  // per hidden layer: S x T x np   a | H_x H_t H_xx
  int zbT[2];                 // S x T x npmax
  int total;                  // floats per CTA
};

struct GenParams {
  NetDesc net;
  LossCoef lc;
  ScratchDesc sd;
  const float* theta;   // [P+2]: flat parameters, then lambda1, lambda2
  const float* wp;      // padded weights
  const float* wt;      // padded transposed weights
  const float* X;       // [N,2]
  int64_t N;
  int64_t nf_global;
  int mode;
  const float* seed;    // S == 1: dL/du^ per point [N, n_out] (or null)
  const float* u_data;  // S == 1: measurements [N, n_out]; the kernel forms data_c * (u^ - u)^2 and its adjoint (or null)
  float data_c;         // data_weight / N_u
  const float* l1_sum;  // V3: device pointer to the job-wide sum |f| (or null)
  float* u_out;         // [N, n_out] or null
  float* f_out;         // [N, n_res] or null
  float* z;             // ADMM state [N, n_res]
  float* gamma;
  int admm_op;          // 0 none, 1 z <- f, 2 z/gamma update, 3 update with the INF-ADMM quirk, 4 update BEFORE the seed (training pass), 5 the same with the quirk
  float* scratch;
  float* part;          // [grid][rvlen] per-CTA partial packed vectors
  int rvlen;
  int cluster;          // CTAs per 32-point tile (1, 2, 4 or 8); grid = clusters * cluster
  int kch;              // K rows staged per shared-memory chunk (set by pinn_generic_launch)
  int wg_nbuf;          // weight-gradient staging buffers (1 or 2; set by pinn_generic_launch)
};

size_t pinn_generic_smem_bytes(const NetDesc& net, int S, bool want_occupancy, int* kch_out, int* nbuf_out, int threads = 256);
cudaError_t pinn_generic_launch(const GenParams& g, int S, int grid, cudaStream_t stream);
int pinn_generic_cluster_capacity(int cs);  // clusters of cs CTAs resident at one CTA per SM (0: unknown)
cudaError_t pinn_generic_dual_launch(const GenParams& g, int S, int grid_res, const GenParams& gd, int grid_data,
                                     cudaStream_t stream);

// small kernels (pinn_aux.cu)
cudaError_t pinn_repack_launch(const NetDesc& net, const float* theta, float* wp, float* wt, cudaStream_t stream);
cudaError_t pinn_finalize_launch(const float* part, int nrows, int rvlen, float* packed, int accumulate, const float* extra,
                                 int extra_idx, cudaStream_t stream, int stride = 0);
cudaError_t pinn_data_seed_launch(const float* u_pred, const float* u_data, int64_t n_u, int n_out, int loss, float weight,
                                  float* seed, float* loss_out, cudaStream_t stream);
struct AdamState {
  float* m;
  float* v;
};
cudaError_t pinn_adam_launch(float* theta, const float* packed, AdamState st, int n, float alpha, float beta1, float beta2,
                             float eps, cudaStream_t stream);
cudaError_t pinn_sample_launch(float* X, int64_t n, uint64_t seed, uint64_t first_index, float lbx, float lbt, float spanx,
                               float spant, cudaStream_t stream);
cudaError_t pinn_lhs_launch(float* X, int64_t n, uint64_t seed, uint64_t first_index, uint64_t n_total, double lbx, double lbt,
                            double wx, double wt, cudaStream_t stream);
cudaError_t pinn_fill_launch(float* p, int64_t n, float v, cudaStream_t stream);
cudaError_t pinn_fma_peak(double* tflops);
