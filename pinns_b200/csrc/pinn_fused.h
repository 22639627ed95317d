// Fused thread-per-point kernel family (pinn_fused.cu): host-side state and entry points.
#pragma once
#include <string>
#include "pinn_kernels.h"

// optional Adam update fused into the finalize kernel (n = 0: off); alpha = lr_t of TF-1 Adam, computed on the host
struct AdamFused {
  int n = 0;
  float* theta = nullptr;
  float* m = nullptr;
  float* v = nullptr;
  float alpha = 0.f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f;
};

// Peer-memory exchange of the packed vector inside the reduction kernel (one process per GPU, buffers shared through
// CUDA IPC over NVLink / NVSwitch): every rank stores its partial elements straight into every peer's receive slot,
// raises a per-chunk flag there, waits for the peers' flags and sums the slots in rank order -- the sum-allreduce of the
// data-parallel step without a separate collective launch (pinn_capi.cu: pinn_comm_*).
constexpr int PINN_MAX_RANKS = 8;
struct FusedComm {
  int world = 1, rank = 0;
  float* slot[PINN_MAX_RANKS] = {};      // receive buffer of rank r as mapped into THIS process: [2 parities][world][rvlen_pad]
  unsigned* flag[PINN_MAX_RANKS] = {};   // flags of rank r: [2][world][nchunks]
  unsigned seq = 0;                      // number of this exchange (1, 2, ...): flag value and parity
  int rvlen_pad = 0, nchunks = 0;
  int* hang = nullptr;                   // set when a peer's flag never arrives (bounded spin)
};

struct FusedState {
  bool enabled = false;
  int hidden = 0;        // hidden width (all hidden layers equal)
  int n_hidden = 0;      // number of hidden layers
  int grid = 0;
  int threads = 0;
  size_t smem = 0;
  float* d_stash = nullptr;   // per-thread activation stash (L2 resident)
  float* d_part = nullptr;    // [grid*warps][region] warp-private gradient / loss accumulators
  float* d_zeros = nullptr;   // one accumulator tile of zeros (read by the first batch of a launch)
  // Low-traffic mode, opt-in (PINN_FUSED_TMEM=1 PINN_FUSED_DISCARD=1): measured on B200 at 16 Mi points it cuts the
  // kernel's DRAM traffic from 4377 to 363 B/point (discard alone: 1893) but costs 2.3 % of the step (discard alone
  // 0.9 %): the kernel is FMA / shared-memory bound, not DRAM bound, so the default is the faster form.
  int tmem_acc = 0;           // W-bar tiles accumulate in tensor memory instead of a global read-modify-write per layer and batch
  int discard = 0;            // discard.global.L2 on stash lines after their last read: dead data is not written back
  int rvlen = 0;
  int region = 0;            // floats per warp-private accumulator region
  // passes of up to this many 8-point batches per warp run on pinn_fused_small_kernel (four lanes per point): the
  // reference's own batch sizes (N_f = 1000 ... 10 771); PINN_FUSED_SMALL_ROUNDS overrides, 0 disables
  int small_rounds = 1;
  int small_extra = 1;   // warps per CTA that may take one batch more than small_rounds
  int small_compact = 1; // the small-batch kernel merges the two k-group copies of a tile slot before the store (half the lines to reduce)
  int pdl = 1;           // programmatic dependent launch of the small-batch kernel and its reduction
  int small_nine = 1;    // nine warps per CTA when eight do not give every batch a warp of its own
};

// decides whether the net qualifies (Burgers, [2, H x k, 1] with a supported H) and allocates
int fused_init(FusedState& fs, const NetDesc& net, const pinn_config_t& cfg, int num_sms, int rvlen, std::string& err);
void fused_destroy(FusedState& fs);
// residual term on the collocation points (+ the squared data term on Xu/ud as extra batches when Xu != null):
// loss sums (+ gradient in mode TRAIN) -> packed (overwritten); optional fused Adam update of theta.
// ev_before / ev_after (optional) bracket the main kernel only.
// accumulate = 1 adds to the warp-private accumulators of the previous launch instead of zeroing them and
// packed = null skips the reduction: a batch that arrives from the host in chunks is one launch per chunk
// (grid_fixed > 0 pins the grid so that every chunk owns the same regions) and one reduction at the end.
int fused_run(FusedState& fs, const NetDesc& net, const LossCoef& lc, const float* theta, const float* X, int64_t n,
              int64_t nf_global, int mode, const float* l1_sum, float* z, float* gamma, int admm_op, float* u_out,
              float* f_out, const float* Xu, const float* ud, int64_t n_u, float data_c, float* packed, const AdamFused& ad,
              cudaEvent_t ev_before, cudaEvent_t ev_after, cudaStream_t stream, std::string& err, int accumulate = 0,
              int grid_fixed = 0, float v1_data_weight = 0.f, const FusedComm* comm = nullptr);
// v1_data_weight != 0: the data batches carry INF-L2's UN-squared norm (pass data_c = 0.5): they must own their warps
// (n/32 + n_u/32 batches <= grid x warps) and the reduction scales their gradient by v1_data_weight / ||r||.
bool fused_v1_fits(const FusedState& fs, int64_t n, int64_t n_u);
