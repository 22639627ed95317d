// tcgen05 / TMEM / TMA kernel for wide Burgers and Euler nets (pinn_tensor.cu): host-side state and entry points.
#pragma once
#include <string>
#include "pinn_kernels.h"

struct TensorState {
  bool enabled = false;
  bool forced = false;       // PINN_PATH_TENSOR: use it whatever the batch size
  int n = 0, NL = 0, grid_max = 0, rvlen = 0;
  int S = 4, NO = 1;         // Taylor streams (4 Burgers, 3 Euler), outputs (1 / 3)
  int shape[11] = {};        // the kernel's TcShape (padded width, column blocks, chunk sizes)
  size_t scratch_stride = 0;
  float* d_scratch = nullptr;
  float* d_wcan = nullptr;   // canonical hi/lo TF32 weight planes
  float* d_part = nullptr;   // [grid][part_stride]: a CTA's partial packed vector, 16-byte aligned (part_stride = rvlen rounded up to 4)
  int part_stride = 0;
  int* d_hang = nullptr;     // set by the kernel if an mbarrier wait times out
  void* d_units = nullptr;   // the tile schedule (contraction units / worker items), built once per shape
  void* d_items = nullptr;
  int nu_f = 0, nu = 0, ni_f = 0, ni = 0;
  // TMA tensor maps over the scratch slab (three CUtensorMap objects: rows of 128 B, boxes of 32 / 32 / 128 rows, written
  // to shared memory with the 32-byte-atom swizzle the MN-major operands need or the 128-byte swizzle of the K-major ones)
  alignas(64) unsigned char tmaps[3][128] = {};
};

// AUTO uses the tensor kernel from this many points on: below it the job has fewer 128-point tiles than the GPU has SMs
// and the generic kernel (clusters of CTAs per 32-point tile) is faster
constexpr int64_t TENSOR_MIN_POINTS = 8192;

int tensor_init(TensorState& ts, const NetDesc& net, const pinn_config_t& cfg, int num_sms, int rvlen, std::string& err);
void tensor_destroy(TensorState& ts);
int tensor_prep(TensorState& ts, const float* theta, cudaStream_t stream, std::string& err);
int tensor_run(TensorState& ts, const NetDesc& net, const LossCoef& lc, const float* theta, const float* X, int64_t n_pts,
               int64_t nf_global, int mode, const float* l1_sum, float* z, float* gamma, int admm_op, float* u_out,
               float* f_out, int* grid_out, cudaStream_t stream, std::string& err);
int tensor_check_hang(TensorState& ts, cudaStream_t stream, std::string& err);
