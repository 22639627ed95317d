/*
 * pinn_b200.h -- C ABI of libpinn_b200.so, the B200 (sm_100a) replacement for the
 * TensorFlow-1 graph + session under the reference's `PhysicsInformedNN` classes.
 *
 * The reference (jonwittmer/PINNs) has no FFI of its own: its boundary is the
 * Python class, whose methods build a TF graph and call `sess.run`.  Each entry
 * point below replaces one group of those graph/session calls; the file:line
 * citations are relative to /root/reference and use SURVEY.md's abbreviations
 * (INF-L2 = Burgers/continuous_inference/Hwan_L2Regularization_Burgers.py,
 *  INF-ADMM = .../Hwan_L1Regularization_ADMM_Burgers.py,
 *  AB-ADMM / AB-L2 / AB-L1 = Burgers/continuous_identification/Abgrall_{ADMM,L2,L1}.py,
 *  ID-L2b / ID-ADMMb = .../Burgers_batch_L2.py / Burgers_ADMM_batch.py,
 *  EUL = Eulers/continuous_inference/Euler_ADMM.py).
 *
 * Conventions
 *  - every function returns 0 on success, a negative PINN_E_* code on failure and
 *    never throws; pinn_last_error() returns the message of the last failure on
 *    that handle (or of the last failed pinn_create when handle is NULL);
 *  - plain pointers and sizes only; `on_device` says whether a caller buffer is a
 *    host pointer (copied with cudaMemcpyAsync on the handle's stream) or a device
 *    pointer on the handle's GPU; the library never frees caller memory;
 *  - one host thread per handle; all device work is issued on the handle's stream
 *    (pinn_set_stream; default: the legacy default stream of the handle's device);
 *  - parameters travel as ONE flat float32 vector in the reference's variable
 *    creation order W1,b1,...,WL,bL with W_l [in,out] row-major (INF-L2:79-88),
 *    which is also the order ScipyOptimizerInterface packs them in (AB-ADMM:66-72);
 *  - there is no CPU fallback: every entry point that computes needs the GPU.
 */
#ifndef PINN_B200_H
#define PINN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PINN_B200_ABI_VERSION 1

typedef struct pinn_handle_s* pinn_handle_t;

/* error codes */
#define PINN_OK 0
#define PINN_E_INVALID (-1)  /* bad argument / unsupported configuration */
#define PINN_E_CUDA (-2)     /* a CUDA runtime call failed */
#define PINN_E_STATE (-3)    /* call made before the data it needs was set */
#define PINN_E_NOMEM (-4)

/* PDE residual (net_f) */
#define PINN_PDE_BURGERS 0 /* f = u_t + lam1*u*u_x - lam2*u_xx  INF-L2:113-120 (lam1=1, lam2=nu), AB-ADMM:170-180 */
#define PINN_PDE_EULER 1   /* (f1,f2,f3) of EUL:176-198, outputs (rho,u,E), gamma = 1.4 */

/* loss variants */
#define PINN_LOSS_V1_INF_L2 1     /* ||u-u^||_2 + mean(f^2)                                    INF-L2:68-69   */
#define PINN_LOSS_V2_INF_ADMM 2   /* (1/Nu)||r||^2 + g^T f + (rho/2)||f - z + g/rho||^2         INF-ADMM:98-100 */
#define PINN_LOSS_V3_L1SQ 3       /* (1/Nu)||r||^2 + (1/Nf)(sum|f|)^2                           ID-L2b:57-58, AB-L1:59-60 */
#define PINN_LOSS_V4_MSE 4        /* (1/Nu)||r||^2 + (1/Nf)||f||^2  (per residual for Euler)   AB-L2:59-60    */
#define PINN_LOSS_V5_ADMM 5       /* (1/Nu)||r||^2 + (rho/2)||f - z + g/rho||^2 (per residual) AB-ADMM:129-130, EUL:128-133 */

/* kernel selection */
#define PINN_PATH_AUTO 0    /* fused thread-per-point kernels for [2,20xk,1]; tcgen05 kernel for wide Burgers / Euler nets at N_f >= 8192; else generic */
#define PINN_PATH_GENERIC 1 /* force the generic tiled FP32 kernel */
#define PINN_PATH_FUSED 2   /* force the fused kernel (error if the net does not qualify) */
#define PINN_PATH_TENSOR 3  /* tcgen05 / TMEM / TMA 3xTF32 kernel: Burgers [2, n x k, 1] and Euler [2, n x k, 3], equal hidden widths 32 <= n <= 256 */

#define PINN_MAX_LAYERS 16

typedef struct pinn_config {
  int32_t abi_version;              /* PINN_B200_ABI_VERSION */
  int32_t n_layers;                 /* len(layers), e.g. 10 for [2,20x8,1]           INF-L2:158 */
  int32_t layers[PINN_MAX_LAYERS];  /* layers[0] must be 2 (x,t)                               */
  int32_t pde;                      /* PINN_PDE_*                                             */
  int32_t loss;                     /* PINN_LOSS_*                                            */
  double lb[2], ub[2];              /* domain bounds (x,t), float64 as in INF-L2:173-174      */
  float lambda1, lambda2;           /* Burgers coefficients; lambda2 = nu for INF-*           */
  float rho;                        /* ADMM penalty (rho / pen), ignored by V1,V3,V4          */
  int32_t trainable_lambda;         /* 1: Adam also updates (lambda1, lambda2) (BASELINE config 2) */
  int32_t device;                   /* CUDA device ordinal                                    */
  int32_t path;                     /* PINN_PATH_*                                            */
  int32_t reserved[8];
} pinn_config_t;

/* ---- lifetime: replaces graph construction + tf.Session (INF-L2:26-77) ---- */
int pinn_create(const pinn_config_t* cfg, pinn_handle_t* out);
int pinn_destroy(pinn_handle_t h);
const char* pinn_last_error(pinn_handle_t h);
int pinn_set_stream(pinn_handle_t h, void* cuda_stream); /* cudaStream_t */
int pinn_synchronize(pinn_handle_t h);

/* ---- introspection ---- */
int pinn_num_params(pinn_handle_t h, int64_t* n_params);   /* P: weights + biases             */
int pinn_packed_len(pinn_handle_t h, int64_t* n);          /* length of the packed vector below */
int pinn_kernel_path(pinn_handle_t h, int32_t* path);      /* the kernel family the CURRENT collocation batch runs on */
int pinn_launch_count(pinn_handle_t h, int64_t* n);        /* kernels launched by this handle so far */

/* ---- variables: tf.Variable init / assign / read (INF-L2:79-94, AB-ADMM:105-106) ---- */
int pinn_set_params(pinn_handle_t h, const float* theta, int on_device);
int pinn_get_params(pinn_handle_t h, float* theta, int on_device);
int pinn_set_lambda(pinn_handle_t h, float lambda1, float lambda2);
int pinn_get_lambda(pinn_handle_t h, float* lambda1, float* lambda2);

/* ---- feeds: the feed_dict of every sess.run (INF-L2:127-128, AB-ADMM:201-202,:220-223) ----
 * X_* are [N,2] row-major (x,t) float32 (the feed-time float64->float32 cast is the
 * caller's, as in TF); u is [N_u, n_out].  Device buffers passed with on_device=1 are
 * BORROWED for collocation points (no copy; must stay valid until replaced) and copied
 * for the data term.  nf_global is the number of collocation points of the whole job
 * (sum over ranks) used in the 1/N_f factors; pass 0 for "same as n_f".              */
int pinn_set_data(pinn_handle_t h, const float* X_u, const float* u, int64_t n_u, int on_device);
int pinn_set_collocation(pinn_handle_t h, const float* X_f, int64_t n_f, int64_t nf_global, int on_device);
/* the per-step feed of the training loop, `sess.run(train_op_Adam, feed_dict={x_f, t_f, ...})` (INF-L2:127-135,
 * AB-ADMM:220-228): X_f_host is HOST memory (pinned for a truly asynchronous copy) that must stay untouched
 * until the next synchronising call.  Returns at once; the points travel in growing chunks on a copy stream
 * and the next pinn_loss_grad_device / pinn_adam_steps starts on chunk k while chunk k+1 is still on the
 * bus (fused path; the other paths and the forward-only passes wait for the whole batch).              */
int pinn_feed_collocation(pinn_handle_t h, const float* X_f_host, int64_t n_f, int64_t nf_global);
/* device-side replacement of np.random.uniform(lb, ub, [N_f,1]) x2 (AB-ADMM:220-221, EUL:232-233):
 * counter-based Philox4x32-10, point i of the job uses counter (first_index + i), so the
 * stream is independent of how the job is sharded over GPUs.                            */
int pinn_sample_collocation(pinn_handle_t h, uint64_t seed, uint64_t first_index, int64_t n_f, int64_t nf_global);
/* device-side replacement of `lb + (ub - lb) * lhs(2, N_f)` (pyDOE, INF-L2:183, INF-ADMM:270): a Latin hypercube design
 * of n_design points over [lb, ub) -- one point in each of the n_design strata of either axis -- of which this handle
 * takes points [first_index, first_index + n_f).  The stratum of point i along axis d is a keyed bijection of i
 * (Feistel network, cycle-walked), the offset inside it a Philox4x32-10 uniform: point i depends on (seed, i, n_design)
 * only, so the design is the same however it is sharded over GPUs.  Float64 arithmetic, one rounding to float32 (the
 * reference's feed-time cast).  Same distribution as pyDOE's default, not NumPy's stream.  n_design = 0: n_f.        */
int pinn_sample_lhs(pinn_handle_t h, uint64_t seed, uint64_t first_index, int64_t n_f, int64_t n_design, int64_t nf_global);
int pinn_get_collocation(pinn_handle_t h, float* X_f, int on_device);
/* scale applied to the data term of loss and gradient (1 on the rank that owns it, 0 elsewhere) */
int pinn_set_data_weight(pinn_handle_t h, float w);

/* ---- data-parallel exchange through peer memory (new capability; the reference has no data parallelism) ----
 * One process per GPU on one NVLink / NVSwitch box.  Each rank exports a CUDA IPC handle of its receive buffer
 * (pinn_comm_export, PINN_COMM_HANDLE_BYTES bytes), the host side gathers the handles of all ranks (any transport:
 * torch.distributed, MPI, files) and hands the table to pinn_comm_attach.  From then on the reduction kernel of every
 * TRAINING pass (pinn_loss_grad_device, pinn_adam_steps) stores its partial vector into every peer's slot, waits for
 * the peers' flags and leaves the SUM over ranks in the packed vector -- the one sum-allreduce of the step without a
 * separate collective -- and the fused Adam update applies to the summed gradient.  All ranks must run the same
 * sequence of training passes.  Fused path only (pinn_kernel_path == PINN_PATH_FUSED); world <= 8.  The data term
 * must ride inside the fused pass (every loss but INF-L2's un-squared norm on shards above ~37 k points; the training
 * entry points return PINN_E_STATE otherwise).  A peer that does not answer within 120 s makes the kernel drop the
 * step (no Adam update) and raise a flag that every synchronising entry point reports as PINN_E_STATE.            */
#define PINN_COMM_HANDLE_BYTES 64
int pinn_comm_export(pinn_handle_t h, void* handle_out);
int pinn_comm_attach(pinn_handle_t h, int rank, int world, const void* handles /* world x PINN_COMM_HANDLE_BYTES */);
int pinn_comm_detach(pinn_handle_t h);
int pinn_comm_status(pinn_handle_t h, int32_t* attached, int32_t* peer_timed_out);

/* ---- the hot path: sess.run([loss, grads]) (INF-L2:135; L-BFGS callback AB-ADMM:216) ----
 * pinn_loss_grad_device leaves the PACKED vector in a device buffer:
 *   [0,P)      d loss / d theta (local sum over this handle's points)
 *   [P,P+2)    d loss / d (lambda1, lambda2)
 *   [P+2,P+10) partial sums: 0 data-term loss, 1 residual-term loss, 2 sum|f|, 3 sum|f-z|,
 *              4 sum f^2, 5..7 reserved
 * so that ONE sum-allreduce of pinn_packed_len() floats combines ranks.  For the L1^2
 * loss (V3) the caller may supply the global sum|f| between the two passes
 * (pinn_l1_pass1 / set_l1_sum); single-GPU callers just use pinn_loss_grad.            */
int pinn_loss_grad_device(pinn_handle_t h);
int pinn_packed_ptr(pinn_handle_t h, float** dev_ptr);
int pinn_l1_pass1(pinn_handle_t h, float** dev_sum_abs_f); /* V3 only: forward pass, local sum|f| */
int pinn_loss_grad(pinn_handle_t h, double* loss, float* grad_host); /* grad_host: P (+2 if trainable_lambda) floats or NULL */
int pinn_loss_value(pinn_handle_t h, double* loss);                   /* sess.run(self.loss)  INF-L2:138 */
/* self.admm_misfit = mean|f_pred - z| (AB-ADMM:60, printed as "r(w) - z" at :232-233): the value the last
 * pinn_loss_value pass accumulated on this handle's points (0 for the non-ADMM losses).                    */
int pinn_admm_misfit(pinn_handle_t h, double* misfit);

/* ---- train_op_Adam: tf.train.AdamOptimizer.minimize (INF-L2:72-73,:135) ----
 * pinn_adam_apply consumes the packed vector (after the caller's allreduce, if any).
 * pinn_adam_steps = n x (loss_grad_device + adam_apply) with no host round trip.       */
int pinn_adam_config(pinn_handle_t h, float lr, float beta1, float beta2, float eps);
int pinn_adam_apply(pinn_handle_t h);
int pinn_adam_steps(pinn_handle_t h, int64_t n_steps);
int pinn_adam_reset(pinn_handle_t h);

/* ---- predict / net_u / net_f callbacks (INF-L2:143-148, AB-ADMM:254-262, EUL:260-272) ----
 * u_out: [N, n_out] or NULL; f_out: [N, n_res] or NULL (n_res = 1 Burgers, 3 Euler).     */
int pinn_predict(pinn_handle_t h, const float* X, int64_t n, float* u_out, float* f_out, int on_device);

/* ---- ADMM state (AB-ADMM:119-134,:185-198,:225-226; INF-ADMM:88-107; EUL:114-141,:237-242) ---- */
int pinn_admm_init(pinn_handle_t h);          /* z = gamma = 1, then z <- f(theta) on the current points (AB-ADMM:121-122,:96-97) */
int pinn_admm_update(pinn_handle_t h, int inf_admm_quirk); /* z_update then gamma_update on the current points */
/* The z/gamma update that closes one epoch of the batch-ADMM loops and the Adam step that opens the next
 * (AB-ADMM:225-226 then :213; EUL:237-242 then :229) evaluate the same residuals: one training pass does both.
 * Equivalent, bit for bit, to pinn_admm_update(h, inf_admm_quirk) followed by pinn_adam_steps(h, 1)
 * (INF-ADMM:189-193 with inf_admm_quirk = 1).                                                            */
int pinn_admm_adam_step(pinn_handle_t h, int inf_admm_quirk);
/* n_epochs iterations of the batch loops of the Dialect-B scripts with the device sampler, without returning to the host
 * (AB-ADMM:211-226, EUL:227-242, AB-L2:205-210: Adam step on the current batch, new batch, z/gamma update on it):
 *   epoch k:  pending ? pinn_admm_adam_step(h, 0) : pinn_adam_steps(h, 1);
 *             pinn_sample_collocation(h, seed, (first_batch + k) * n_f, n_f, nf_global);
 *             pending = admm
 * `pending` says whether a z/gamma update is still owed on entry (it is folded into the first Adam step's pass) and, with
 * admm = 1, one is owed again on return (the caller folds it into its next step or flushes it with pinn_admm_update).
 * Same launches, same bits as the calls above made one by one: at the reference's batch sizes an epoch is ~25 us of GPU work
 * and the per-call host overhead of an interpreted loop was as much again.  Single-GPU loops only (a data-parallel rank
 * samples its own counter range and the ranks are summed every step: distributed.DataParallelStepper).                */
int pinn_resampled_epochs(pinn_handle_t h, int64_t n_epochs, int admm, int pending, uint64_t seed, uint64_t first_batch,
                          int64_t n_f, int64_t nf_global);
int pinn_admm_get_state(pinn_handle_t h, float* z, float* gamma, int on_device); /* [N_f, n_res] each */
int pinn_admm_set_state(pinn_handle_t h, const float* z, const float* gamma, int on_device);

/* ---- measurement hooks (bench.py; no reference counterpart) ----
 * pinn_kernel_timing(1) brackets every launch of the dominant residual(+grad) kernel with CUDA
 * events on the handle's stream; pinn_kernel_time synchronizes and returns their summed duration.
 * pinn_measure_fma_peak runs an FFMA-only micro-kernel: the measured FP32 roofline denominator. */
int pinn_kernel_timing(pinn_handle_t h, int enable);
int pinn_kernel_time(pinn_handle_t h, double* total_ms, int64_t* n_launches);
int pinn_measure_fma_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* PINN_B200_H */
