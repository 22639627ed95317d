// C ABI of libpinn_b200.so (see include/pinn_b200.h for the contract and the reference
// lines each entry point replaces).  Host-side orchestration only: every FLOP of the hot
// path runs in the sm_100a kernels of pinn_fused.cu / pinn_generic.cu; there is no CPU path.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "pinn_kernels.h"
#include <cstdlib>

#include "pinn_fused.h"
#include "pinn_tensor.h"

struct pinn_handle_s {
  pinn_config_t cfg;
  NetDesc net;
  int S_res = 4;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  std::string err;
  int64_t launches = 0;
  int path_used = PINN_PATH_GENERIC;

  float* d_theta = nullptr;  // [P+2]
  float* d_wp = nullptr;
  float* d_wt = nullptr;
  bool weights_dirty = true;
  int rvlen = 0;
  float* d_packed = nullptr;

  float* d_Xu = nullptr;
  float* d_u = nullptr;
  float* d_upred = nullptr;
  float* d_seed = nullptr;
  float* d_data_loss = nullptr;
  int64_t n_u = 0;
  float data_weight = 1.0f;

  float* d_Xf = nullptr;
  float* d_Xf_owned = nullptr;
  int64_t xf_cap = 0;
  int64_t n_f = 0, nf_global = 0;
  float* d_z = nullptr;
  float* d_gamma = nullptr;
  int64_t admm_cap = 0;

  AdamState adam = {nullptr, nullptr};
  float lr = 1e-3f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f;
  int64_t adam_t = 0;               // steps applied so far (TF's beta1_power / beta2_power, kept in double on the host)
  double b1pow = 1.0, b2pow = 1.0;

  float* d_scratch = nullptr;
  int gen_grid_max = 0;
  int cluster_cap[9] = {0};  // clusters of c CTAs resident at one CTA per SM (pinn_generic_cluster_capacity)
  float* d_part = nullptr;
  float* d_part_data = nullptr;
  ScratchDesc sd_res, sd_data;
  float* d_l1sum = nullptr;
  bool l1_ready = false;

  FusedState fused;
  TensorState tensor;
  bool tensor_dirty = true;

  bool timing = false;
  std::vector<cudaEvent_t> ev;   // pairs (before, after) of the dominant kernel
  size_t ev_used = 0;

  // host feed in flight (pinn_feed_collocation): chunk c = points [feed_first[c], feed_first[c+1]) has landed in
  // d_Xf_owned once feed_ev[c] fires on copy_stream
  cudaStream_t copy_stream = nullptr;
  // pinn_set_collocation from host memory: small batches (the reference's per-epoch resampling: 8 KB) go through a ring of pinned
  // staging buffers, so the copy is asynchronous and the host prepares batch k+1 while the GPU works on batch k (a copy from
  // pageable memory makes the driver wait for the stream first)
  static constexpr int STAGE_SLOTS = 4;
  static constexpr size_t STAGE_BYTES = 256 * 1024;
  float* stage[STAGE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t stage_ev[STAGE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  unsigned stage_next = 0;
  cudaEvent_t feed_start = nullptr;
  std::vector<cudaEvent_t> feed_ev;
  std::vector<int64_t> feed_first;
  int feed_chunks = 0;

  // peer-memory exchange of the packed vector (pinn_comm_*): own receive buffer + the peers' buffers opened through CUDA IPC
  struct {
    bool attached = false;
    int rank = 0, world = 1;
    void* own = nullptr;                 // [2][PINN_MAX_RANKS][rvlen_pad] floats | [2][PINN_MAX_RANKS][nchunks] flags
    void* opened[PINN_MAX_RANKS] = {};
    size_t slot_bytes = 0;
    int rvlen_pad = 0, nchunks = 0;
    unsigned seq = 0;
    int* h_hang = nullptr;               // pinned, mapped: the reduction kernel raises it, every synchronising call reads it
    int* d_hang = nullptr;               // device alias of h_hang
  } comm;

  // staging buffers of pinn_predict for host callers (grown on demand, kept for the handle's lifetime)
  double last_misfit = 0.0;  // mean |f - z| of the last pinn_loss_value pass (admm_misfit, AB-ADMM:60)
  float* d_pred_X = nullptr;
  float* d_pred_u = nullptr;
  float* d_pred_f = nullptr;
  int64_t pred_cap = 0;
};

static std::string g_create_err;

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      h->err = std::string(#call) + ": " + cudaGetErrorString(_e);                                 \
      return PINN_E_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

#define REQUIRE(cond, code, msg) \
  do {                           \
    if (!(cond)) {               \
      h->err = (msg);            \
      return (code);             \
    }                            \
  } while (0)

static int round8(int v) { return (v + 7) / 8 * 8; }

static ScratchDesc make_scratch_desc(const NetDesc& net, int S) {
  ScratchDesc sd;
  memset(&sd, 0, sizeof(sd));
  int off = 0;
  sd.in0 = off;
  off += S * 8 * PINN_TILE;
  for (int hl = 0; hl < net.L - 1; ++hl) {
    sd.hid[hl] = off;
    off += (2 * S - 1) * net.np[hl + 1] * PINN_TILE;
  }
  sd.Y = off;
  off += S * 8 * PINN_TILE;
  for (int k = 0; k < 2; ++k) {
    sd.zb[k] = off;
    off += S * net.npmax * PINN_TILE;
  }
  sd.in0T = off;
  off += S * PINN_TILE * 8;
  for (int hl = 0; hl < net.L - 1; ++hl) {
    sd.hidT[hl] = off;
    off += S * PINN_TILE * net.np[hl + 1];
  }
  for (int k = 0; k < 2; ++k) {
    sd.zbT[k] = off;
    off += S * PINN_TILE * net.npmax;
  }
  sd.total = off;
  return sd;
}

static LossCoef make_loss_coef(const pinn_handle_s* h, int loss) {
  LossCoef lc;
  memset(&lc, 0, sizeof(lc));
  lc.loss = loss;
  lc.rho = h->cfg.rho;
  const double nf = (double)(h->nf_global > 0 ? h->nf_global : (h->n_f > 0 ? h->n_f : 1));
  lc.inv_nf = (float)(1.0 / nf);
  switch (loss) {
    case PINN_LOSS_V1_INF_L2:
    case PINN_LOSS_V4_MSE:
      lc.cA = (float)(2.0 / nf);
      break;
    case PINN_LOSS_V3_L1SQ:
      break;  // cB formed on the device from the job-wide sum |f|
    case PINN_LOSS_V5_ADMM:
      lc.cC = h->cfg.rho;
      lc.cD = 1.0f;
      break;
    case PINN_LOSS_V2_INF_ADMM:
      lc.cC = h->cfg.rho;
      lc.cD = 2.0f;
      break;
  }
  return lc;
}

static int ensure_weights(pinn_handle_t h) {
  if (h->weights_dirty) {
    CK(pinn_repack_launch(h->net, h->d_theta, h->d_wp, h->d_wt, h->stream));
    h->launches += 1;
    h->weights_dirty = false;
  }
  return PINN_OK;
}

// Grid and cluster size of the generic kernel.  With fewer tiles than SMs a cluster of cs CTAs shares a tile (the CTAs
// split the output neurons / weight-gradient tiles and meet at cluster barriers).  Measured on B200
// (scripts/cluster_size_probe.py, [2,200x5,3] and [2,200x8,1], logs under profiles/): a tile's time falls with cs
// (1 : 2 : 3 : 4 CTAs = 3.2-4.2 : 2.1-2.8 : 1.3-1.4 : 1) as long as every CTA has an SM to itself; clusters beyond what the GPU
// holds at one CTA per SM (a cluster lives inside one GPC: 148 / 74 / 45 / 33 / 26 / 22 / 15 / 15 clusters of 1..8 CTAs) share SMs with the others and the
// kernel slows down by about (clusters / capacity + 0.4).  So: the cs that minimises  t(cs) * that penalty,  where
// `extra_tiles` counts the data-term tiles that ride in the same launch (39 clusters for the reference's Euler batch:
// cs = 3 fits, cs = 4 does not: 337 -> 294 us per step).
static int gen_grid_for(const pinn_handle_s* h, int64_t n, int64_t extra_tiles, int* cluster) {
  int64_t tiles = (n + PINN_TILE - 1) / PINN_TILE;
  if (tiles < 1) tiles = 1;
  static const int cs_max = [] {
    const char* e = getenv("PINN_GEN_CLUSTER_MAX");  // measurement knob: cap on the CTAs that share a tile
    const int v = e ? atoi(e) : 8;
    return v >= 1 && v <= 8 ? v : 8;
  }();
  // narrow layers have too few 8-wide output groups to share
  int groups = 1;
  for (int l = 1; l <= h->net.L; ++l) groups = h->net.np[l] / 8 > groups ? h->net.np[l] / 8 : groups;
  static const double t_rel[9] = {0, 3.7, 2.5, 1.37, 1.0, 0.95, 0.9, 0.88, 0.85};
  int cs = 1;
  double best = 1e300;
  for (int c = 1; c <= cs_max; ++c) {
    if (c > 1 && c * 8 > groups * 2) break;
    const int cap = h->cluster_cap[c] > 0 ? h->cluster_cap[c] : h->num_sms / c;
    const double load = (double)(tiles + extra_tiles) / (double)cap;
    const double cost = t_rel[c] * (load > 1.0 ? load + 0.4 : 1.0);  // measured: 1.15 -> 1.5x, 1.6 -> 1.9x, 2.1 -> 2.4x
    if (cost < best * 0.999) {
      best = cost;
      cs = c;
    }
  }
  static const int cs_force = [] {
    const char* e = getenv("PINN_GEN_CLUSTER_FORCE");  // measurement knob
    const int v = e ? atoi(e) : 0;
    return v >= 1 && v <= 8 ? v : 0;
  }();
  if (cs_force) cs = cs_force;
  *cluster = cs;
  const int64_t clusters = tiles < h->gen_grid_max / cs ? tiles : h->gen_grid_max / cs;
  return (int)(clusters * cs);
}

static void fill_gen_params(const pinn_handle_s* h, int S, int mode, int loss, const float* X, int64_t n, const float* seed,
                            float* u_out, float* f_out, int admm_op, bool use_state, float* part, GenParams& g) {
  memset(&g, 0, sizeof(g));
  g.net = h->net;
  g.lc = make_loss_coef(h, loss);
  g.sd = (S == 1) ? h->sd_data : h->sd_res;
  g.theta = h->d_theta;
  g.wp = h->d_wp;
  g.wt = h->d_wt;
  g.X = X;
  g.N = n;
  g.nf_global = h->nf_global > 0 ? h->nf_global : h->n_f;
  g.mode = mode;
  g.seed = seed;
  g.l1_sum = (loss == PINN_LOSS_V3_L1SQ && mode == GEN_MODE_TRAIN) ? h->d_l1sum : nullptr;
  g.u_out = u_out;
  g.f_out = f_out;
  g.z = use_state ? h->d_z : nullptr;
  g.gamma = use_state ? h->d_gamma : nullptr;
  g.admm_op = admm_op;
  g.scratch = h->d_scratch;
  g.part = part;
  g.rvlen = h->rvlen;
}

// one launch of the generic kernel; returns the number of partial-sum rows (= clusters) through *grid_out.  u_data != null (S == 1): the squared
// data misfit and its adjoint are formed inside the kernel
static int run_generic(pinn_handle_t h, int S, int mode, int loss, const float* X, int64_t n, const float* seed,
                       float* u_out, float* f_out, int admm_op, bool use_state, float* part, int* grid_out,
                       const float* u_data = nullptr) {
  {
    int rcw = ensure_weights(h);  // padded / transposed weight copies are only needed by this kernel
    if (rcw) return rcw;
  }
  GenParams g;
  fill_gen_params(h, S, mode, loss, X, n, seed, u_out, f_out, admm_op, use_state, part, g);
  g.u_data = u_data;
  g.data_c = h->n_u > 0 ? h->data_weight / (float)h->n_u : 0.f;
  const int grid = gen_grid_for(h, n, 0, &g.cluster);
  CK(pinn_generic_launch(g, S, grid, h->stream));
  h->launches += 1;
  if (grid_out) *grid_out = grid / g.cluster;
  return PINN_OK;
}

// residual tiles and the (squared) data term in ONE launch of the generic kernel: the data CTAs follow the residual
// CTAs in the grid, in the partial-sum rows and in the scratch slabs
static int run_generic_dual(pinn_handle_t h, int loss, int admm_op, bool use_state, int* rows_out) {
  int rcw = ensure_weights(h);
  if (rcw) return rcw;
  GenParams g, gd;
  fill_gen_params(h, h->S_res, GEN_MODE_TRAIN, loss, h->d_Xf, h->n_f, nullptr, nullptr, nullptr, admm_op, use_state, h->d_part, g);
  const int grid_res = gen_grid_for(h, h->n_f, (h->n_u + PINN_TILE - 1) / PINN_TILE, &g.cluster);
  const int cs = g.cluster;
  // data clusters: one per tile (when the residual job leaves too few free CTA slots the last ones simply queue)
  int64_t dclusters = (h->n_u + PINN_TILE - 1) / PINN_TILE;
  if (dclusters > h->gen_grid_max / cs) dclusters = h->gen_grid_max / cs;
  const int grid_data = (int)dclusters * cs;
  fill_gen_params(h, 1, GEN_MODE_TRAIN, PINN_LOSS_V4_MSE, h->d_Xu, h->n_u, nullptr, nullptr, nullptr, 0, false,
                  h->d_part + (size_t)(grid_res / cs) * h->rvlen, gd);
  gd.u_data = h->d_u;
  gd.data_c = h->data_weight / (float)h->n_u;
  gd.cluster = cs;
  gd.scratch = h->d_scratch + (size_t)(grid_res / cs) * h->sd_res.total;
  CK(pinn_generic_dual_launch(g, h->S_res, grid_res, gd, grid_data, h->stream));
  h->launches += 1;
  *rows_out = (grid_res + grid_data) / cs;
  return PINN_OK;
}

static bool loss_uses_state(int loss) { return loss == PINN_LOSS_V2_INF_ADMM || loss == PINN_LOSS_V5_ADMM; }

static int ensure_admm(pinn_handle_t h) {
  const int64_t need = h->n_f * h->net.n_res;
  if (need > h->admm_cap) {
    if (h->d_z) cudaFree(h->d_z);
    if (h->d_gamma) cudaFree(h->d_gamma);
    h->d_z = h->d_gamma = nullptr;
    CK(cudaMalloc(&h->d_z, need * sizeof(float)));
    CK(cudaMalloc(&h->d_gamma, need * sizeof(float)));
    h->admm_cap = need;
    CK(pinn_fill_launch(h->d_z, need, 1.0f, h->stream));  // tf.ones (AB-ADMM:121-122)
    CK(pinn_fill_launch(h->d_gamma, need, 1.0f, h->stream));
    h->launches += 2;
  }
  return PINN_OK;
}

extern "C" {

const char* pinn_last_error(pinn_handle_t h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int pinn_create(const pinn_config_t* cfg, pinn_handle_t* out) {
  if (!cfg || !out) {
    g_create_err = "pinn_create: null argument";
    return PINN_E_INVALID;
  }
  *out = nullptr;
  if (cfg->abi_version != PINN_B200_ABI_VERSION) {
    g_create_err = "pinn_create: abi_version mismatch";
    return PINN_E_INVALID;
  }
  if (cfg->n_layers < 3 || cfg->n_layers > PINN_MAX_LAYERS || cfg->layers[0] != 2) {
    g_create_err = "pinn_create: need 3..16 layers with layers[0] == 2 (x,t)";
    return PINN_E_INVALID;
  }
  const int n_out = cfg->layers[cfg->n_layers - 1];
  if ((cfg->pde == PINN_PDE_BURGERS && n_out != 1) || (cfg->pde == PINN_PDE_EULER && n_out != 3) ||
      (cfg->pde != PINN_PDE_BURGERS && cfg->pde != PINN_PDE_EULER)) {
    g_create_err = "pinn_create: Burgers needs 1 output, Euler needs 3 (rho,u,E)";
    return PINN_E_INVALID;
  }
  if (cfg->loss < PINN_LOSS_V1_INF_L2 || cfg->loss > PINN_LOSS_V5_ADMM) {
    g_create_err = "pinn_create: unknown loss variant";
    return PINN_E_INVALID;
  }
  for (int l = 0; l < cfg->n_layers; ++l)
    if (cfg->layers[l] < 1 || cfg->layers[l] > 1024) {
      g_create_err = "pinn_create: layer width out of range [1,1024]";
      return PINN_E_INVALID;
    }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("pinn_create: no CUDA device (") + cudaGetErrorString(e) + "); libpinn_b200 has no CPU path";
    return PINN_E_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) {
    g_create_err = "pinn_create: device ordinal out of range";
    return PINN_E_INVALID;
  }
  pinn_handle_s* h = new (std::nothrow) pinn_handle_s();
  if (!h) {
    g_create_err = "pinn_create: out of host memory";
    return PINN_E_NOMEM;
  }
  h->cfg = *cfg;
  NetDesc& net = h->net;
  memset(&net, 0, sizeof(net));
  net.L = cfg->n_layers - 1;
  int off = 0, wp = 0, wt = 0, npmax = 8;
  for (int l = 0; l <= net.L; ++l) {
    net.n[l] = cfg->layers[l];
    net.np[l] = round8(cfg->layers[l]);
    if (l > 0 && net.np[l] > npmax) npmax = net.np[l];
  }
  for (int l = 0; l < net.L; ++l) {
    net.w_off[l] = off;
    off += net.n[l] * net.n[l + 1];
    net.b_off[l] = off;
    off += net.n[l + 1];
    net.wp_off[l] = wp;
    wp += net.n[l] * net.np[l + 1];
    net.wt_off[l] = wt;
    wt += net.n[l + 1] * net.np[l];
  }
  net.P = off;
  net.npmax = npmax;
  net.n_out = n_out;
  net.n_res = cfg->pde == PINN_PDE_BURGERS ? 1 : 3;
  net.pde = cfg->pde;
  net.lbx = (float)cfg->lb[0];
  net.lbt = (float)cfg->lb[1];
  net.spanx = (float)(cfg->ub[0] - cfg->lb[0]);
  net.spant = (float)(cfg->ub[1] - cfg->lb[1]);
  h->S_res = cfg->pde == PINN_PDE_BURGERS ? 4 : 3;
  h->rvlen = net.P + 2 + PINN_NSUMS;

  auto fail = [&](const char* what, cudaError_t ce) {
    g_create_err = std::string("pinn_create: ") + what + ": " + cudaGetErrorString(ce);
    pinn_destroy(h);
    return PINN_E_CUDA;
  };
  if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return fail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
  if (prop.major != 10) {
    g_create_err = "pinn_create: libpinn_b200 is built for sm_100a only (found sm_" + std::to_string(prop.major) +
                   std::to_string(prop.minor) + ")";
    pinn_destroy(h);
    return PINN_E_INVALID;
  }
  h->num_sms = prop.multiProcessorCount;
  h->gen_grid_max = h->num_sms * 2;
  for (int c = 1; c <= 8; ++c) h->cluster_cap[c] = pinn_generic_cluster_capacity(c);
  if (getenv("PINN_GEN_DEBUG"))
    fprintf(stderr, "pinn_b200: cluster capacity at one CTA per SM: %d %d %d %d %d %d %d %d\n", h->cluster_cap[1], h->cluster_cap[2],
            h->cluster_cap[3], h->cluster_cap[4], h->cluster_cap[5], h->cluster_cap[6], h->cluster_cap[7], h->cluster_cap[8]);
  h->sd_res = make_scratch_desc(net, h->S_res);
  h->sd_data = make_scratch_desc(net, 1);

  const size_t px = (size_t)net.P + 2;
  if ((e = cudaMalloc(&h->d_theta, px * sizeof(float))) != cudaSuccess) return fail("cudaMalloc theta", e);
  if ((e = cudaMemset(h->d_theta, 0, px * sizeof(float))) != cudaSuccess) return fail("cudaMemset", e);
  if ((e = cudaMalloc(&h->d_wp, (size_t)wp * sizeof(float))) != cudaSuccess) return fail("cudaMalloc wp", e);
  if ((e = cudaMalloc(&h->d_wt, (size_t)wt * sizeof(float))) != cudaSuccess) return fail("cudaMalloc wt", e);
  if ((e = cudaMalloc(&h->d_packed, (size_t)h->rvlen * sizeof(float))) != cudaSuccess) return fail("cudaMalloc packed", e);
  if ((e = cudaMemset(h->d_packed, 0, (size_t)h->rvlen * sizeof(float))) != cudaSuccess) return fail("cudaMemset", e);
  if ((e = cudaMalloc(&h->d_scratch, (size_t)h->sd_res.total * h->gen_grid_max * 2 * sizeof(float))) != cudaSuccess)
    return fail("cudaMalloc scratch", e);
  if ((e = cudaMalloc(&h->d_part, (size_t)h->rvlen * h->gen_grid_max * 2 * sizeof(float))) != cudaSuccess)
    return fail("cudaMalloc partials", e);
  if ((e = cudaMalloc(&h->d_part_data, (size_t)h->rvlen * h->gen_grid_max * sizeof(float))) != cudaSuccess)
    return fail("cudaMalloc data partials", e);
  if ((e = cudaMalloc(&h->d_l1sum, sizeof(float))) != cudaSuccess) return fail("cudaMalloc l1", e);
  if ((e = cudaMalloc(&h->d_data_loss, sizeof(float))) != cudaSuccess) return fail("cudaMalloc dl", e);
  if ((e = cudaMemset(h->d_data_loss, 0, sizeof(float))) != cudaSuccess) return fail("cudaMemset", e);
  if ((e = cudaMalloc(&h->adam.m, px * sizeof(float))) != cudaSuccess) return fail("cudaMalloc m", e);
  if ((e = cudaMalloc(&h->adam.v, px * sizeof(float))) != cudaSuccess) return fail("cudaMalloc v", e);
  *out = h;
  int rc = pinn_adam_reset(h);
  if (rc != PINN_OK) {
    g_create_err = h->err;
    pinn_destroy(h);
    *out = nullptr;
    return rc;
  }
  rc = pinn_set_lambda(h, cfg->lambda1, cfg->lambda2);
  if (rc == PINN_OK) rc = fused_init(h->fused, h->net, h->cfg, h->num_sms, h->rvlen, h->err);
  if (rc == PINN_OK) rc = tensor_init(h->tensor, h->net, h->cfg, h->num_sms, h->rvlen, h->err);
  if (rc != PINN_OK) {
    g_create_err = h->err;
    pinn_destroy(h);
    *out = nullptr;
    return rc;
  }
  h->path_used = h->tensor.enabled ? PINN_PATH_TENSOR : (h->fused.enabled ? PINN_PATH_FUSED : PINN_PATH_GENERIC);
  return PINN_OK;
}

int pinn_destroy(pinn_handle_t h) {
  if (!h) return PINN_OK;
  cudaSetDevice(h->cfg.device);
  fused_destroy(h->fused);
  tensor_destroy(h->tensor);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  if (h->copy_stream) {
    cudaStreamSynchronize(h->copy_stream);
    cudaStreamDestroy(h->copy_stream);
    cudaEventDestroy(h->feed_start);
  }
  for (cudaEvent_t e : h->feed_ev) cudaEventDestroy(e);
  for (int k = 0; k < pinn_handle_s::STAGE_SLOTS; ++k) {
    if (h->stage_ev[k]) cudaEventDestroy(h->stage_ev[k]);
    if (h->stage[k]) cudaFreeHost(h->stage[k]);
  }
  pinn_comm_detach(h);
  if (h->comm.own) cudaFree(h->comm.own);
  if (h->comm.h_hang) cudaFreeHost(h->comm.h_hang);
  if (h->d_pred_X) cudaFree(h->d_pred_X);
  if (h->d_pred_u) cudaFree(h->d_pred_u);
  if (h->d_pred_f) cudaFree(h->d_pred_f);
  float* bufs[] = {h->d_theta, h->d_wp,   h->d_wt,   h->d_packed, h->d_Xu,      h->d_u,   h->d_upred,     h->d_seed,
                   h->d_Xf_owned, h->d_z, h->d_gamma, h->adam.m,  h->adam.v,    h->d_scratch, h->d_part, h->d_part_data,
                   h->d_l1sum, h->d_data_loss};
  for (float* b : bufs)
    if (b) cudaFree(b);
  delete h;
  return PINN_OK;
}

int pinn_set_stream(pinn_handle_t h, void* s) {
  if (!h) return PINN_E_INVALID;
  h->stream = (cudaStream_t)s;
  return PINN_OK;
}

// after a stream synchronisation: did a kernel give up waiting (a lost data-parallel peer, a tcgen05 mbarrier that never
// completed)?  Then the replicated state is no longer trustworthy and every synchronising entry point says so.
static int check_device_flags(pinn_handle_t h) {
  if (h->comm.h_hang && *(volatile int*)h->comm.h_hang) {
    h->err = "peer-memory exchange: a peer's flag did not arrive within 120 s; the step was dropped (no Adam update), the job is out of sync";
    return PINN_E_STATE;
  }
  if (h->tensor.enabled) return tensor_check_hang(h->tensor, h->stream, h->err);
  return PINN_OK;
}

int pinn_synchronize(pinn_handle_t h) {
  if (!h) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  return check_device_flags(h);
}

int pinn_num_params(pinn_handle_t h, int64_t* n) {
  if (!h || !n) return PINN_E_INVALID;
  *n = h->net.P;
  return PINN_OK;
}
int pinn_packed_len(pinn_handle_t h, int64_t* n) {
  if (!h || !n) return PINN_E_INVALID;
  *n = h->rvlen;
  return PINN_OK;
}
static bool tensor_takes(const pinn_handle_s* h, int64_t n) {
  return !h->fused.enabled && h->tensor.enabled && (h->tensor.forced || n >= TENSOR_MIN_POINTS);
}

int pinn_kernel_path(pinn_handle_t h, int32_t* p) {
  if (!h || !p) return PINN_E_INVALID;
  // the path the CURRENT collocation batch takes (before any batch is set: the one a large batch would take)
  if (h->fused.enabled) *p = PINN_PATH_FUSED;
  else if (h->tensor.enabled && (h->n_f == 0 || tensor_takes(h, h->n_f))) *p = PINN_PATH_TENSOR;
  else *p = PINN_PATH_GENERIC;
  return PINN_OK;
}
int pinn_launch_count(pinn_handle_t h, int64_t* n) {
  if (!h || !n) return PINN_E_INVALID;
  *n = h->launches;
  return PINN_OK;
}

int pinn_set_params(pinn_handle_t h, const float* theta, int on_device) {
  if (!h || !theta) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemcpyAsync(h->d_theta, theta, (size_t)h->net.P * sizeof(float),
                     on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  if (!on_device) CK(cudaStreamSynchronize(h->stream));
  h->weights_dirty = h->tensor_dirty = true;
  return PINN_OK;
}

int pinn_get_params(pinn_handle_t h, float* theta, int on_device) {
  if (!h || !theta) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemcpyAsync(theta, h->d_theta, (size_t)h->net.P * sizeof(float),
                     on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  if (!on_device) {
    CK(cudaStreamSynchronize(h->stream));
    return check_device_flags(h);
  }
  return PINN_OK;
}

int pinn_set_lambda(pinn_handle_t h, float l1, float l2) {
  if (!h) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  const float v[2] = {l1, l2};
  CK(cudaMemcpyAsync(h->d_theta + h->net.P, v, sizeof(v), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PINN_OK;
}

int pinn_get_lambda(pinn_handle_t h, float* l1, float* l2) {
  if (!h || !l1 || !l2) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  float v[2];
  CK(cudaMemcpyAsync(v, h->d_theta + h->net.P, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *l1 = v[0];
  *l2 = v[1];
  return PINN_OK;
}

int pinn_set_data(pinn_handle_t h, const float* X_u, const float* u, int64_t n_u, int on_device) {
  if (!h || n_u < 0 || (n_u > 0 && (!X_u || !u))) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  // the new buffers are allocated and filled first: a failure leaves the previous data term in place
  float* nb[4] = {nullptr, nullptr, nullptr, nullptr};
  if (n_u > 0) {
    const size_t no = (size_t)h->net.n_out;
    const size_t bytes[4] = {(size_t)n_u * 2 * sizeof(float), (size_t)n_u * no * sizeof(float), (size_t)n_u * no * sizeof(float),
                             (size_t)n_u * no * sizeof(float)};
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 4 && e == cudaSuccess; ++k) e = cudaMalloc(&nb[k], bytes[k]);
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (e == cudaSuccess) e = cudaMemcpyAsync(nb[0], X_u, bytes[0], kind, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(nb[1], u, bytes[1], kind, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
      for (float* b : nb)
        if (b) cudaFree(b);
      h->err = std::string("pinn_set_data: ") + cudaGetErrorString(e);
      return PINN_E_CUDA;
    }
  } else {
    CK(cudaStreamSynchronize(h->stream));  // nothing in flight reads the buffers about to be freed
  }
  float** bufs[] = {&h->d_Xu, &h->d_u, &h->d_upred, &h->d_seed};
  for (int k = 0; k < 4; ++k) {
    if (*bufs[k]) cudaFree(*bufs[k]);
    *bufs[k] = nb[k];
  }
  h->n_u = n_u;
  return PINN_OK;
}

static int feed_join(pinn_handle_t h);

static int ensure_xf_owned(pinn_handle_t h, int64_t n_f) {
  if (n_f > h->xf_cap) {
    if (h->d_Xf_owned) cudaFree(h->d_Xf_owned);
    h->d_Xf_owned = nullptr;
    CK(cudaMalloc(&h->d_Xf_owned, (size_t)n_f * 2 * sizeof(float)));
    h->xf_cap = n_f;
  }
  return PINN_OK;
}

int pinn_set_collocation(pinn_handle_t h, const float* X_f, int64_t n_f, int64_t nf_global, int on_device) {
  if (!h || !X_f || n_f <= 0) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  {
    int rc = feed_join(h);
    if (rc) return rc;
  }
  if (on_device) {
    h->d_Xf = const_cast<float*>(X_f);  // borrowed
  } else {
    int rc = ensure_xf_owned(h, n_f);
    if (rc) return rc;
    const size_t bytes = (size_t)n_f * 2 * sizeof(float);
    if (bytes <= pinn_handle_s::STAGE_BYTES) {
      const int slot = (int)(h->stage_next++ % pinn_handle_s::STAGE_SLOTS);
      if (!h->stage[slot]) {
        CK(cudaHostAlloc(&h->stage[slot], pinn_handle_s::STAGE_BYTES, cudaHostAllocDefault));
        CK(cudaEventCreateWithFlags(&h->stage_ev[slot], cudaEventDisableTiming));
      } else {
        CK(cudaEventSynchronize(h->stage_ev[slot]));  // the copy that last used this slot (four batches ago) has left it
      }
      memcpy(h->stage[slot], X_f, bytes);               // the caller's buffer is free again when this call returns
      CK(cudaMemcpyAsync(h->d_Xf_owned, h->stage[slot], bytes, cudaMemcpyHostToDevice, h->stream));
      CK(cudaEventRecord(h->stage_ev[slot], h->stream));
    } else {
      // pageable source: the runtime stages the data before it returns (and waits for the stream to do so)
      CK(cudaMemcpyAsync(h->d_Xf_owned, X_f, bytes, cudaMemcpyHostToDevice, h->stream));
    }
    h->d_Xf = h->d_Xf_owned;
  }
  h->n_f = n_f;
  h->nf_global = nf_global > 0 ? nf_global : n_f;
  h->l1_ready = false;
  return PINN_OK;
}

// chunk boundaries of a host feed: a small first chunk (its copy is the only one nobody hides), then 4x growth in
// whole rounds of the persistent grid.  The kernel consumes ~4.4 GB/s of points, PCIe delivers 25-55 GB/s, so the
// copy of chunk k+1 is done before the kernel has finished chunk k as long as the growth factor stays below that ratio.
static void feed_plan(const pinn_handle_s* h, int64_t n_f, std::vector<int64_t>& first) {
  first.clear();
  first.push_back(0);
  const int64_t round = (int64_t)h->fused.grid * (h->fused.threads / 32) * 32;  // points per round of all warps
  const bool chunked = h->fused.enabled && h->cfg.loss != PINN_LOSS_V3_L1SQ && round > 0 && n_f >= 16 * round;
  if (chunked) {
    int64_t at = 0, len = 4 * round;
    while (at + len < n_f) {
      at += len;
      first.push_back(at);
      len *= 4;
    }
  }
  first.push_back(n_f);
}

int pinn_feed_collocation(pinn_handle_t h, const float* X_f_host, int64_t n_f, int64_t nf_global) {
  if (!h || !X_f_host || n_f <= 0) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  int rc = ensure_xf_owned(h, n_f);
  if (rc) return rc;
  if (!h->copy_stream) {
    CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->feed_start, cudaEventDisableTiming));
  }
  feed_plan(h, n_f, h->feed_first);
  const int nc = (int)h->feed_first.size() - 1;
  while ((int)h->feed_ev.size() < nc) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->feed_ev.push_back(e);
  }
  // the previous step's kernels may still be reading the buffer
  CK(cudaEventRecord(h->feed_start, h->stream));
  CK(cudaStreamWaitEvent(h->copy_stream, h->feed_start, 0));
  for (int c = 0; c < nc; ++c) {
    const int64_t a = h->feed_first[c], b = h->feed_first[c + 1];
    CK(cudaMemcpyAsync(h->d_Xf_owned + 2 * a, X_f_host + 2 * a, (size_t)(b - a) * 2 * sizeof(float), cudaMemcpyHostToDevice,
                       h->copy_stream));
    CK(cudaEventRecord(h->feed_ev[c], h->copy_stream));
  }
  h->feed_chunks = nc;
  h->d_Xf = h->d_Xf_owned;
  h->n_f = n_f;
  h->nf_global = nf_global > 0 ? nf_global : n_f;
  h->l1_ready = false;
  return PINN_OK;
}

// everything that is not the chunk-aware fused training pass waits for the whole feed
static int feed_join(pinn_handle_t h) {
  if (h->feed_chunks > 0) {
    CK(cudaStreamWaitEvent(h->stream, h->feed_ev[h->feed_chunks - 1], 0));
    h->feed_chunks = 0;
  }
  return PINN_OK;
}

int pinn_sample_collocation(pinn_handle_t h, uint64_t seed, uint64_t first_index, int64_t n_f, int64_t nf_global) {
  if (!h || n_f <= 0) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  int rc = feed_join(h);
  if (rc) return rc;
  rc = ensure_xf_owned(h, n_f);
  if (rc) return rc;
  CK(pinn_sample_launch(h->d_Xf_owned, n_f, seed, first_index, h->net.lbx, h->net.lbt, h->net.spanx, h->net.spant,
                        h->stream));
  h->launches += 1;
  h->d_Xf = h->d_Xf_owned;
  h->n_f = n_f;
  h->nf_global = nf_global > 0 ? nf_global : n_f;
  h->l1_ready = false;
  return PINN_OK;
}

int pinn_sample_lhs(pinn_handle_t h, uint64_t seed, uint64_t first_index, int64_t n_f, int64_t n_design, int64_t nf_global) {
  if (!h || n_f <= 0) return PINN_E_INVALID;
  if (n_design <= 0) n_design = n_f;
  REQUIRE(first_index + (uint64_t)n_f <= (uint64_t)n_design, PINN_E_INVALID,
          "pinn_sample_lhs: [first_index, first_index + n_f) must lie inside the design of n_design points");
  REQUIRE((uint64_t)n_design <= (1ull << 48), PINN_E_INVALID, "pinn_sample_lhs: designs beyond 2^48 points are not supported");
  CK(cudaSetDevice(h->cfg.device));
  int rc = feed_join(h);
  if (rc) return rc;
  rc = ensure_xf_owned(h, n_f);
  if (rc) return rc;
  CK(pinn_lhs_launch(h->d_Xf_owned, n_f, seed, first_index, (uint64_t)n_design, h->cfg.lb[0], h->cfg.lb[1],
                     h->cfg.ub[0] - h->cfg.lb[0], h->cfg.ub[1] - h->cfg.lb[1], h->stream));
  h->launches += 1;
  h->d_Xf = h->d_Xf_owned;
  h->n_f = n_f;
  h->nf_global = nf_global > 0 ? nf_global : n_f;
  h->l1_ready = false;
  return PINN_OK;
}

int pinn_get_collocation(pinn_handle_t h, float* X_f, int on_device) {
  if (!h || !X_f) return PINN_E_INVALID;
  REQUIRE(h->d_Xf && h->n_f > 0, PINN_E_STATE, "pinn_get_collocation: no collocation points set");
  CK(cudaSetDevice(h->cfg.device));
  {
    int rc = feed_join(h);
    if (rc) return rc;
  }
  CK(cudaMemcpyAsync(X_f, h->d_Xf, (size_t)h->n_f * 2 * sizeof(float),
                     on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  if (!on_device) CK(cudaStreamSynchronize(h->stream));
  return PINN_OK;
}

int pinn_set_data_weight(pinn_handle_t h, float w) {
  if (!h) return PINN_E_INVALID;
  h->data_weight = w;
  return PINN_OK;
}

// data term: forward on X_u, misfit + adjoints, seeded reverse sweep (primal stream only)
static int data_term(pinn_handle_t h, bool want_grad) {
  if (h->n_u == 0 || h->data_weight == 0.0f) {
    CK(cudaMemsetAsync(h->d_data_loss, 0, sizeof(float), h->stream));
    return PINN_OK;
  }
  if (want_grad && h->cfg.loss != PINN_LOSS_V1_INF_L2) {  // squared misfit: seed and loss inside the kernel, one pass
    int grid = 0;
    int rc1 = run_generic(h, 1, GEN_MODE_TRAIN, PINN_LOSS_V4_MSE, h->d_Xu, h->n_u, nullptr, nullptr, nullptr, 0, false,
                          h->d_part_data, &grid, h->d_u);
    if (rc1) return rc1;
    CK(pinn_finalize_launch(h->d_part_data, grid, h->rvlen, h->d_packed, 1, nullptr, -1, h->stream));
    h->launches += 1;
    return PINN_OK;
  }
  int rc = run_generic(h, 1, GEN_MODE_FORWARD, PINN_LOSS_V4_MSE, h->d_Xu, h->n_u, nullptr, h->d_upred, nullptr, 0, false,
                       h->d_part_data, nullptr);
  if (rc) return rc;
  CK(pinn_data_seed_launch(h->d_upred, h->d_u, h->n_u, h->net.n_out, h->cfg.loss, h->data_weight, h->d_seed,
                           h->d_data_loss, h->stream));
  h->launches += 1;
  if (want_grad) {
    int grid = 0;
    rc = run_generic(h, 1, GEN_MODE_TRAIN, PINN_LOSS_V4_MSE, h->d_Xu, h->n_u, h->d_seed, nullptr, nullptr, 0, false,
                     h->d_part_data, &grid);
    if (rc) return rc;
    CK(pinn_finalize_launch(h->d_part_data, grid, h->rvlen, h->d_packed, 1, h->d_data_loss, h->net.P + 2 + PINN_SUM_DATA,
                            h->stream));
    h->launches += 1;
  }
  return PINN_OK;
}

static int timing_event(pinn_handle_t h, cudaEvent_t* out) {
  *out = nullptr;
  if (!h->timing) return PINN_OK;
  if (h->ev_used == h->ev.size()) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    h->ev.push_back(e);
  }
  *out = h->ev[h->ev_used++];
  return PINN_OK;
}

// the squared data term rides inside the fused kernel (extra batches); V1's un-squared norm needs ||r|| first
static bool fused_handles_data(const pinn_handle_s* h) {
  if (!h->fused.enabled || h->n_u <= 0 || h->data_weight == 0.0f) return false;
  // V1's un-squared norm rides along when every warp gets one batch at most (the reference's sizes: 10 456 + 100 points)
  if (h->cfg.loss == PINN_LOSS_V1_INF_L2) return h->feed_chunks <= 1 && fused_v1_fits(h->fused, h->n_f, h->n_u);
  return true;
}

// same for the generic kernel's training pass (dual launch) when neither the fused nor the tensor kernel takes the batch
static bool generic_handles_data(const pinn_handle_s* h) {
  if (h->fused.enabled) return false;
  if (tensor_takes(h, h->n_f)) return false;
  return h->cfg.loss != PINN_LOSS_V1_INF_L2 && h->n_u > 0 && h->data_weight != 0.0f;
}

static float adam_next_alpha(pinn_handle_s* h) {
  h->adam_t += 1;
  h->b1pow *= (double)h->beta1;
  h->b2pow *= (double)h->beta2;
  return (float)((double)h->lr * std::sqrt(1.0 - h->b2pow) / (1.0 - h->b1pow));
}

// the exchange descriptor of the next reduction (null when this handle is not part of a peer-memory group)
static const FusedComm* next_comm(pinn_handle_s* h, int mode, FusedComm& cm) {
  if (!h->comm.attached || h->comm.world <= 1 || mode != GEN_MODE_TRAIN) return nullptr;
  cm.world = h->comm.world;
  cm.rank = h->comm.rank;
  for (int r = 0; r < cm.world; ++r) {
    char* base = static_cast<char*>(r == cm.rank ? h->comm.own : h->comm.opened[r]);
    cm.slot[r] = reinterpret_cast<float*>(base);
    cm.flag[r] = reinterpret_cast<unsigned*>(base + h->comm.slot_bytes);
  }
  cm.seq = ++h->comm.seq;
  cm.rvlen_pad = h->comm.rvlen_pad;
  cm.nchunks = h->comm.nchunks;
  cm.hang = h->comm.d_hang;
  return &cm;
}

static int residual_pass(pinn_handle_t h, int mode, int admm_op, bool fuse_adam = false) {
  const bool state = loss_uses_state(h->cfg.loss) || admm_op != 0;
  if (state) {
    int rc = ensure_admm(h);
    if (rc) return rc;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = timing_event(h, &e0);
  if (rc == PINN_OK) rc = timing_event(h, &e1);
  if (rc) return rc;
  if (h->fused.enabled) {
    const bool with_data = fused_handles_data(h);
    AdamFused ad;
    if (fuse_adam) {
      ad.n = h->net.P + (h->cfg.trainable_lambda ? 2 : 0);
      ad.theta = h->d_theta;
      ad.m = h->adam.m;
      ad.v = h->adam.v;
      ad.alpha = adam_next_alpha(h);
      ad.beta1 = h->beta1;
      ad.beta2 = h->beta2;
      ad.eps = h->eps;
    }
    const int64_t nfg = h->nf_global > 0 ? h->nf_global : h->n_f;
    const float* l1 = (h->cfg.loss == PINN_LOSS_V3_L1SQ && mode == GEN_MODE_TRAIN) ? h->d_l1sum : nullptr;
    FusedComm cm;
    const bool v1 = with_data && h->cfg.loss == PINN_LOSS_V1_INF_L2;
    const float data_c = !with_data ? 0.f : (v1 ? 0.5f : h->data_weight / (float)h->n_u);
    if (h->feed_chunks > 1 && mode == GEN_MODE_TRAIN && admm_op == 0) {
      // host-fed batch: one launch per chunk as soon as its copy has landed, accumulators carried over, the data
      // term rides with the last chunk, one reduction (+ Adam) at the end
      const int nc = h->feed_chunks;
      h->feed_chunks = 0;
      for (int c = 0; c < nc && rc == PINN_OK; ++c) {
        const int64_t a = h->feed_first[c], n = h->feed_first[c + 1] - a;
        const bool last = (c == nc - 1);
        CK(cudaStreamWaitEvent(h->stream, h->feed_ev[c], 0));
        rc = fused_run(h->fused, h->net, make_loss_coef(h, h->cfg.loss), h->d_theta, h->d_Xf + 2 * a, n, nfg, mode, l1,
                       state ? h->d_z + a : nullptr, state ? h->d_gamma + a : nullptr, 0, nullptr, nullptr,
                       (with_data && last) ? h->d_Xu : nullptr, (with_data && last) ? h->d_u : nullptr, h->n_u, data_c,
                       last ? h->d_packed : nullptr, ad, c == 0 ? e0 : nullptr, last ? e1 : nullptr, h->stream, h->err,
                       /*accumulate=*/c > 0, /*grid_fixed=*/h->fused.grid, 0.f, last ? next_comm(h, mode, cm) : nullptr);
        h->launches += 1;
      }
      if (rc) return rc;
      h->launches -= 1;  // the common tail below counts the last kernel + the reduction
    } else {
      rc = feed_join(h);
      if (rc) return rc;
      rc = fused_run(h->fused, h->net, make_loss_coef(h, h->cfg.loss), h->d_theta, h->d_Xf, h->n_f, nfg, mode, l1,
                     state ? h->d_z : nullptr, state ? h->d_gamma : nullptr, admm_op, nullptr, nullptr,
                     with_data ? h->d_Xu : nullptr, with_data ? h->d_u : nullptr, h->n_u, data_c, h->d_packed, ad, e0, e1,
                     h->stream, h->err, 0, 0, (v1 && mode == GEN_MODE_TRAIN) ? h->data_weight : 0.f, next_comm(h, mode, cm));
    }
    if (rc) return rc;
    h->launches += 2;
    if (fuse_adam) h->weights_dirty = h->tensor_dirty = true;
    return PINN_OK;
  }
  rc = feed_join(h);
  if (rc) return rc;
  int grid = 0;
  if (tensor_takes(h, h->n_f)) {
    if (h->tensor_dirty) {
      rc = tensor_prep(h->tensor, h->d_theta, h->stream, h->err);
      if (rc) return rc;
      h->tensor_dirty = false;
      h->launches += 1;
    }
    if (e0) CK(cudaEventRecord(e0, h->stream));
    rc = tensor_run(h->tensor, h->net, make_loss_coef(h, h->cfg.loss), h->d_theta, h->d_Xf, h->n_f,
                    h->nf_global > 0 ? h->nf_global : h->n_f, mode,
                    (h->cfg.loss == PINN_LOSS_V3_L1SQ && mode == GEN_MODE_TRAIN) ? h->d_l1sum : nullptr,
                    state ? h->d_z : nullptr, state ? h->d_gamma : nullptr, admm_op, nullptr, nullptr, &grid, h->stream, h->err);
    if (rc) return rc;
    if (e1) CK(cudaEventRecord(e1, h->stream));
    CK(pinn_finalize_launch(h->tensor.d_part, grid, h->rvlen, h->d_packed, 0, nullptr, -1, h->stream, h->tensor.part_stride));
    h->launches += 2;
    return PINN_OK;
  }
  if (e0) CK(cudaEventRecord(e0, h->stream));
  if (mode == GEN_MODE_TRAIN && generic_handles_data(h))
    rc = run_generic_dual(h, h->cfg.loss, admm_op, state, &grid);
  else
    rc = run_generic(h, h->S_res, mode, h->cfg.loss, h->d_Xf, h->n_f, nullptr, nullptr, nullptr, admm_op, state, h->d_part,
                     &grid);
  if (rc) return rc;
  if (e1) CK(cudaEventRecord(e1, h->stream));
  CK(pinn_finalize_launch(h->d_part, grid, h->rvlen, h->d_packed, 0, nullptr, -1, h->stream));
  h->launches += 1;
  return PINN_OK;
}

int pinn_l1_pass1(pinn_handle_t h, float** dev_sum) {
  if (!h) return PINN_E_INVALID;
  REQUIRE(h->d_Xf && h->n_f > 0, PINN_E_STATE, "pinn_l1_pass1: no collocation points set");
  CK(cudaSetDevice(h->cfg.device));
  int rc = residual_pass(h, GEN_MODE_FORWARD, 0);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->d_l1sum, h->d_packed + h->net.P + 2 + PINN_SUM_ABSF, sizeof(float), cudaMemcpyDeviceToDevice,
                     h->stream));
  h->l1_ready = true;
  if (dev_sum) *dev_sum = h->d_l1sum;
  return PINN_OK;
}

int pinn_loss_grad_device(pinn_handle_t h) {
  if (!h) return PINN_E_INVALID;
  REQUIRE(h->d_Xf && h->n_f > 0, PINN_E_STATE, "pinn_loss_grad: no collocation points set");
  CK(cudaSetDevice(h->cfg.device));
  int rc = PINN_OK;
  if (h->cfg.loss == PINN_LOSS_V3_L1SQ && !h->l1_ready) {
    rc = pinn_l1_pass1(h, nullptr);
    if (rc) return rc;
  }
  const bool data_inside = fused_handles_data(h) || generic_handles_data(h);
  REQUIRE(!(h->comm.attached && h->comm.world > 1) || data_inside || h->n_u == 0 || h->data_weight == 0.0f, PINN_E_STATE,
          "pinn_loss_grad_device: peer-memory exchange attached but the data term is not part of the fused pass (V1 loss on "
          "a large shard): detach and combine the ranks with an allreduce of the packed vector instead");
  rc = residual_pass(h, GEN_MODE_TRAIN, 0);
  if (rc) return rc;
  h->l1_ready = false;
  if (data_inside) return PINN_OK;
  return data_term(h, true);
}

int pinn_packed_ptr(pinn_handle_t h, float** p) {
  if (!h || !p) return PINN_E_INVALID;
  *p = h->d_packed;
  return PINN_OK;
}

static double assemble_loss(const pinn_handle_s* h, const float* sums) {
  const double nf = (double)(h->nf_global > 0 ? h->nf_global : h->n_f);
  double loss = (double)sums[PINN_SUM_DATA];
  if (h->cfg.loss == PINN_LOSS_V3_L1SQ)
    loss += (double)sums[PINN_SUM_ABSF] * (double)sums[PINN_SUM_ABSF] / nf;
  else
    loss += (double)sums[PINN_SUM_RES];
  return loss;
}

int pinn_loss_grad(pinn_handle_t h, double* loss, float* grad_host) {
  if (!h) return PINN_E_INVALID;
  int rc = pinn_loss_grad_device(h);
  if (rc) return rc;
  float sums[PINN_NSUMS];
  CK(cudaMemcpyAsync(sums, h->d_packed + h->net.P + 2, sizeof(sums), cudaMemcpyDeviceToHost, h->stream));
  if (grad_host) {
    const size_t n = (size_t)h->net.P + (h->cfg.trainable_lambda ? 2 : 0);
    CK(cudaMemcpyAsync(grad_host, h->d_packed, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  rc = check_device_flags(h);
  if (rc) return rc;
  if (loss) *loss = assemble_loss(h, sums);
  return PINN_OK;
}

int pinn_loss_value(pinn_handle_t h, double* loss) {
  if (!h || !loss) return PINN_E_INVALID;
  REQUIRE(h->d_Xf && h->n_f > 0, PINN_E_STATE, "pinn_loss_value: no collocation points set");
  CK(cudaSetDevice(h->cfg.device));
  int rc = residual_pass(h, GEN_MODE_FORWARD, 0);
  if (rc) return rc;
  rc = data_term(h, false);
  if (rc) return rc;
  float sums[PINN_NSUMS];
  float dl = 0.f;
  CK(cudaMemcpyAsync(sums, h->d_packed + h->net.P + 2, sizeof(sums), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&dl, h->d_data_loss, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  rc = check_device_flags(h);
  if (rc) return rc;
  sums[PINN_SUM_DATA] = dl;
  h->last_misfit = (double)sums[PINN_SUM_MISFIT] / ((double)h->n_f * h->net.n_res);
  *loss = assemble_loss(h, sums);
  return PINN_OK;
}

int pinn_adam_config(pinn_handle_t h, float lr, float b1, float b2, float eps) {
  if (!h) return PINN_E_INVALID;
  h->lr = lr;
  h->beta1 = b1;
  h->beta2 = b2;
  h->eps = eps;
  return PINN_OK;
}

int pinn_adam_reset(pinn_handle_t h) {
  if (!h) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  const size_t px = (size_t)h->net.P + 2;
  CK(cudaMemsetAsync(h->adam.m, 0, px * sizeof(float), h->stream));
  CK(cudaMemsetAsync(h->adam.v, 0, px * sizeof(float), h->stream));
  h->adam_t = 0;
  h->b1pow = h->b2pow = 1.0;
  return PINN_OK;
}

int pinn_adam_apply(pinn_handle_t h) {
  if (!h) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  const int n = h->net.P + (h->cfg.trainable_lambda ? 2 : 0);
  CK(pinn_adam_launch(h->d_theta, h->d_packed, h->adam, n, adam_next_alpha(h), h->beta1, h->beta2, h->eps, h->stream));
  h->launches += 1;
  h->weights_dirty = h->tensor_dirty = true;
  return PINN_OK;
}

int pinn_adam_steps(pinn_handle_t h, int64_t n_steps) {
  if (!h || n_steps < 0) return PINN_E_INVALID;
  REQUIRE(h->d_Xf && h->n_f > 0, PINN_E_STATE, "pinn_adam_steps: no collocation points set");
  CK(cudaSetDevice(h->cfg.device));
  // single-GPU fast lane: when the fused kernel also carries the data term, a step is two launches
  // (residual+grad kernel, finalize+Adam) with no host round trip
  const bool lane2 = h->fused.enabled && (fused_handles_data(h) || h->n_u == 0 || h->data_weight == 0.0f);
  // with a peer-memory group attached the exchange happens inside the reduction of the residual pass: a data term that
  // is added afterwards (INF-L2's un-squared norm on a batch too large to carry it) would miss the sum over ranks
  REQUIRE(!(h->comm.attached && h->comm.world > 1) || lane2, PINN_E_STATE,
          "pinn_adam_steps: peer-memory exchange attached but the data term is not part of the fused pass (V1 loss on a "
          "large shard): detach and combine the ranks with an allreduce of the packed vector instead");
  for (int64_t it = 0; it < n_steps; ++it) {
    int rc;
    if (lane2) {
      if (h->cfg.loss == PINN_LOSS_V3_L1SQ) {
        rc = pinn_l1_pass1(h, nullptr);
        if (rc) return rc;
      }
      rc = residual_pass(h, GEN_MODE_TRAIN, 0, /*fuse_adam=*/true);
      h->l1_ready = false;
      if (rc) return rc;
      continue;
    }
    rc = pinn_loss_grad_device(h);
    if (rc) return rc;
    rc = pinn_adam_apply(h);
    if (rc) return rc;
  }
  return PINN_OK;
}

// One epoch boundary of the batch-ADMM loops (AB-ADMM:213-226, EUL:229-242): the z/gamma update that closes epoch k and
// the Adam step that opens epoch k+1 evaluate the same residuals f(theta, batch) -- one training pass does both
// (admm_op 4 / 5: update first, then seed the reverse sweep with the new state).  Bit-identical to pinn_admm_update(h, quirk)
// followed by pinn_adam_steps(h, 1) -- INF-ADMM:189-193 with its double dual update when quirk is set; the tcgen05 path and the multi-pass L1^2 loss take exactly that route.
int pinn_admm_adam_step(pinn_handle_t h, int quirk) {
  if (!h) return PINN_E_INVALID;
  REQUIRE(h->d_Xf && h->n_f > 0, PINN_E_STATE, "pinn_admm_adam_step: no collocation points set");
  REQUIRE(h->cfg.loss == PINN_LOSS_V5_ADMM || h->cfg.loss == PINN_LOSS_V2_INF_ADMM, PINN_E_STATE,
          "pinn_admm_adam_step: the loss is not an ADMM loss");
  const int op = quirk ? 5 : 4;
  CK(cudaSetDevice(h->cfg.device));
  int rc = feed_join(h);
  if (rc) return rc;
  if (tensor_takes(h, h->n_f)) {
    rc = pinn_admm_update(h, quirk);
    return rc ? rc : pinn_adam_steps(h, 1);
  }
  const bool lane2 = h->fused.enabled && (fused_handles_data(h) || h->n_u == 0 || h->data_weight == 0.0f);
  if (lane2) return residual_pass(h, GEN_MODE_TRAIN, op, /*fuse_adam=*/true);
  REQUIRE(!(h->comm.attached && h->comm.world > 1), PINN_E_STATE,
          "pinn_admm_adam_step: peer-memory exchange attached but the data term is not part of the fused pass");
  rc = residual_pass(h, GEN_MODE_TRAIN, op);
  if (rc) return rc;
  if (!(fused_handles_data(h) || generic_handles_data(h))) {
    rc = data_term(h, true);
    if (rc) return rc;
  }
  return pinn_adam_apply(h);
}

int pinn_resampled_epochs(pinn_handle_t h, int64_t n_epochs, int admm, int pending, uint64_t seed, uint64_t first_batch,
                          int64_t n_f, int64_t nf_global) {
  if (!h || n_epochs < 0 || n_f <= 0) return PINN_E_INVALID;
  REQUIRE(!admm || h->cfg.loss == PINN_LOSS_V5_ADMM || h->cfg.loss == PINN_LOSS_V2_INF_ADMM, PINN_E_STATE,
          "pinn_resampled_epochs: admm = 1 but the loss is not an ADMM loss");
  // the batches are drawn for the whole job: a data-parallel rank samples its own counter range and sums over ranks between
  // the pass and the update (distributed.DataParallelStepper), which this single-handle loop does not do
  REQUIRE(!(h->comm.attached && h->comm.world > 1), PINN_E_STATE, "pinn_resampled_epochs: not for handles of a data-parallel group");
  for (int64_t k = 0; k < n_epochs; ++k) {
    int rc = pending ? pinn_admm_adam_step(h, 0) : pinn_adam_steps(h, 1);
    if (rc) return rc;
    rc = pinn_sample_collocation(h, seed, (first_batch + (uint64_t)k) * (uint64_t)n_f, n_f, nf_global);
    if (rc) return rc;
    pending = admm;
  }
  return PINN_OK;
}

int pinn_predict(pinn_handle_t h, const float* X, int64_t n, float* u_out, float* f_out, int on_device) {
  if (!h || !X || n <= 0) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  int rc = PINN_OK;
  const float* dX = X;
  float *du = u_out, *df = f_out;
  const size_t no = h->net.n_out, nr = h->net.n_res;
  if (!on_device) {
    // staging buffers live in the handle: record_data calls this every 100 / 1000 epochs (AB-ADMM:238-246)
    if (n > h->pred_cap) {
      CK(cudaStreamSynchronize(h->stream));
      float** bufs[] = {&h->d_pred_X, &h->d_pred_u, &h->d_pred_f};
      for (float** b : bufs) {
        if (*b) cudaFree(*b);
        *b = nullptr;
      }
      h->pred_cap = 0;
      CK(cudaMalloc(&h->d_pred_X, (size_t)n * 2 * sizeof(float)));
      CK(cudaMalloc(&h->d_pred_u, (size_t)n * no * sizeof(float)));
      CK(cudaMalloc(&h->d_pred_f, (size_t)n * nr * sizeof(float)));
      h->pred_cap = n;
    }
    CK(cudaMemcpyAsync(h->d_pred_X, X, (size_t)n * 2 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    dX = h->d_pred_X;
    du = u_out ? h->d_pred_u : nullptr;
    df = f_out ? h->d_pred_f : nullptr;
  }
  if (h->fused.enabled) {
    // forward-only mode of the fused kernel: u and f of every point in one pass (INF-L2:143-148 does two sess.runs)
    LossCoef lc = make_loss_coef(h, PINN_LOSS_V4_MSE);
    AdamFused none;
    rc = fused_run(h->fused, h->net, lc, h->d_theta, dX, n, n, GEN_MODE_FORWARD, nullptr, nullptr, nullptr, 0, du, df, nullptr,
                   nullptr, 0, 0.f, nullptr, none, nullptr, nullptr, h->stream, h->err);
    h->launches += 1;
  } else if (f_out && tensor_takes(h, n)) {
    if (h->tensor_dirty) {
      rc = tensor_prep(h->tensor, h->d_theta, h->stream, h->err);
      h->tensor_dirty = false;
    }
    if (rc == PINN_OK)
      rc = tensor_run(h->tensor, h->net, make_loss_coef(h, PINN_LOSS_V4_MSE), h->d_theta, dX, n, n, GEN_MODE_FORWARD, nullptr,
                      nullptr, nullptr, 0, du, df, nullptr, h->stream, h->err);
    h->launches += 1;
  } else if (f_out) {
    rc = run_generic(h, h->S_res, GEN_MODE_FORWARD, PINN_LOSS_V4_MSE, dX, n, nullptr, du, df, 0, false, h->d_part, nullptr);
  } else {
    rc = run_generic(h, 1, GEN_MODE_FORWARD, PINN_LOSS_V4_MSE, dX, n, nullptr, du, nullptr, 0, false, h->d_part_data, nullptr);
  }
  if (rc == PINN_OK && !on_device) {
    if (u_out) CK(cudaMemcpyAsync(u_out, du, (size_t)n * no * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (f_out) CK(cudaMemcpyAsync(f_out, df, (size_t)n * nr * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    rc = check_device_flags(h);
  }
  return rc;
}

int pinn_admm_misfit(pinn_handle_t h, double* misfit) {
  if (!h || !misfit) return PINN_E_INVALID;
  *misfit = h->last_misfit;
  return PINN_OK;
}

int pinn_admm_init(pinn_handle_t h) {
  if (!h) return PINN_E_INVALID;
  REQUIRE(h->d_Xf && h->n_f > 0, PINN_E_STATE, "pinn_admm_init: no collocation points set");
  CK(cudaSetDevice(h->cfg.device));
  int rc = ensure_admm(h);
  if (rc) return rc;
  const int64_t need = h->n_f * h->net.n_res;
  CK(pinn_fill_launch(h->d_z, need, 1.0f, h->stream));
  CK(pinn_fill_launch(h->d_gamma, need, 1.0f, h->stream));
  h->launches += 2;
  return residual_pass(h, GEN_MODE_FORWARD, 1);
}

int pinn_admm_update(pinn_handle_t h, int quirk) {
  if (!h) return PINN_E_INVALID;
  REQUIRE(h->d_Xf && h->n_f > 0, PINN_E_STATE, "pinn_admm_update: no collocation points set");
  CK(cudaSetDevice(h->cfg.device));
  return residual_pass(h, GEN_MODE_FORWARD, quirk ? 3 : 2);
}

int pinn_admm_get_state(pinn_handle_t h, float* z, float* gamma, int on_device) {
  if (!h) return PINN_E_INVALID;
  REQUIRE(h->d_z && h->n_f > 0, PINN_E_STATE, "pinn_admm_get_state: ADMM state not initialised");
  CK(cudaSetDevice(h->cfg.device));
  const size_t bytes = (size_t)h->n_f * h->net.n_res * sizeof(float);
  const cudaMemcpyKind k = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (z) CK(cudaMemcpyAsync(z, h->d_z, bytes, k, h->stream));
  if (gamma) CK(cudaMemcpyAsync(gamma, h->d_gamma, bytes, k, h->stream));
  if (!on_device) CK(cudaStreamSynchronize(h->stream));
  return PINN_OK;
}

int pinn_admm_set_state(pinn_handle_t h, const float* z, const float* gamma, int on_device) {
  if (!h) return PINN_E_INVALID;
  REQUIRE(h->n_f > 0, PINN_E_STATE, "pinn_admm_set_state: no collocation points set");
  CK(cudaSetDevice(h->cfg.device));
  int rc = ensure_admm(h);
  if (rc) return rc;
  const size_t bytes = (size_t)h->n_f * h->net.n_res * sizeof(float);
  const cudaMemcpyKind k = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (z) CK(cudaMemcpyAsync(h->d_z, z, bytes, k, h->stream));
  if (gamma) CK(cudaMemcpyAsync(h->d_gamma, gamma, bytes, k, h->stream));
  if (!on_device) CK(cudaStreamSynchronize(h->stream));
  return PINN_OK;
}

/* ---- peer-memory exchange group (one process per GPU; buffers shared through CUDA IPC) ---- */
int pinn_comm_export(pinn_handle_t h, void* handle_out) {
  if (!h || !handle_out) return PINN_E_INVALID;
  REQUIRE(h->fused.enabled, PINN_E_STATE, "pinn_comm_export: the peer-memory exchange lives in the fused path's reduction");
  CK(cudaSetDevice(h->cfg.device));
  if (!h->comm.own) {
    h->comm.rvlen_pad = (h->rvlen + 31) / 32 * 32;
    h->comm.nchunks = h->fused.region / 32;
    h->comm.slot_bytes = (size_t)2 * PINN_MAX_RANKS * h->comm.rvlen_pad * sizeof(float);
    const size_t bytes = h->comm.slot_bytes + (size_t)2 * PINN_MAX_RANKS * h->comm.nchunks * sizeof(unsigned);
    CK(cudaMalloc(&h->comm.own, bytes));
    CK(cudaMemset(h->comm.own, 0, bytes));
    CK(cudaHostAlloc(&h->comm.h_hang, sizeof(int), cudaHostAllocMapped));
    *h->comm.h_hang = 0;
    CK(cudaHostGetDevicePointer(&h->comm.d_hang, h->comm.h_hang, 0));
    CK(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t ipc;
  CK(cudaIpcGetMemHandle(&ipc, h->comm.own));
  static_assert(sizeof(ipc) == PINN_COMM_HANDLE_BYTES, "CUDA IPC handle size");
  memcpy(handle_out, &ipc, sizeof(ipc));
  return PINN_OK;
}

int pinn_comm_attach(pinn_handle_t h, int rank, int world, const void* handles) {
  if (!h || !handles) return PINN_E_INVALID;
  REQUIRE(h->comm.own, PINN_E_STATE, "pinn_comm_attach: call pinn_comm_export first");
  REQUIRE(world >= 1 && world <= PINN_MAX_RANKS && rank >= 0 && rank < world, PINN_E_INVALID, "pinn_comm_attach: bad rank / world");
  CK(cudaSetDevice(h->cfg.device));
  const char* hs = static_cast<const char*>(handles);
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, hs + (size_t)r * PINN_COMM_HANDLE_BYTES, sizeof(ipc));
    const cudaError_t e = cudaIpcOpenMemHandle(&h->comm.opened[r], ipc, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {  // leave nothing mapped behind: the caller falls back to the allreduce
      h->comm.opened[r] = nullptr;
      pinn_comm_detach(h);
      h->err = std::string("pinn_comm_attach: cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e);
      return PINN_E_CUDA;
    }
  }
  // a fresh numbering of the exchanges needs fresh flags (the peers store into this buffer only after the caller's
  // barrier that follows the attach)
  CK(cudaMemset(h->comm.own, 0, h->comm.slot_bytes + (size_t)2 * PINN_MAX_RANKS * h->comm.nchunks * sizeof(unsigned)));
  CK(cudaDeviceSynchronize());
  *h->comm.h_hang = 0;
  h->comm.rank = rank;
  h->comm.world = world;
  h->comm.seq = 0;
  h->comm.attached = true;
  return PINN_OK;
}

int pinn_comm_detach(pinn_handle_t h) {
  if (!h) return PINN_E_INVALID;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);  // (the legacy default stream is a valid argument) no exchange kernel may still be running
  for (int r = 0; r < PINN_MAX_RANKS; ++r)
    if (h->comm.opened[r]) {
      cudaIpcCloseMemHandle(h->comm.opened[r]);
      h->comm.opened[r] = nullptr;
    }
  h->comm.attached = false;
  return PINN_OK;
}

int pinn_comm_status(pinn_handle_t h, int32_t* attached, int32_t* hang) {
  if (!h) return PINN_E_INVALID;
  if (attached) *attached = h->comm.attached ? 1 : 0;
  if (hang) {
    *hang = 0;
    if (h->comm.h_hang) {
      CK(cudaSetDevice(h->cfg.device));
      CK(cudaStreamSynchronize(h->stream));
      *hang = *(volatile int*)h->comm.h_hang;
    }
  }
  return PINN_OK;
}

int pinn_kernel_timing(pinn_handle_t h, int enable) {
  if (!h) return PINN_E_INVALID;
  h->timing = enable != 0;
  h->ev_used = 0;
  return PINN_OK;
}

int pinn_kernel_time(pinn_handle_t h, double* total_ms, int64_t* n_launches) {
  if (!h || !total_ms) return PINN_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaStreamSynchronize(h->stream));
  double tot = 0.0;
  for (size_t k = 0; k + 1 < h->ev_used; k += 2) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev[k], h->ev[k + 1]));
    tot += ms;
  }
  *total_ms = tot;
  if (n_launches) *n_launches = (int64_t)(h->ev_used / 2);
  h->ev_used = 0;
  return PINN_OK;
}

int pinn_measure_fma_peak(int device, double* tflops) {
  if (!tflops) return PINN_E_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return PINN_E_CUDA;
  return pinn_fma_peak(tflops) == cudaSuccess ? PINN_OK : PINN_E_CUDA;
}

}  // extern "C"
