"""Drop-in `PhysicsInformedNN` classes: the reference's Python class API over libpinn_b200.

Same constructor arguments, method names, argument meaning and stopping rules as the
reference classes; the TF graph + session underneath is replaced by `Engine` (one C-ABI
handle per GPU).  No math happens in this file.

Dialect A (explicit arrays; Raissi / Goh):
  PhysicsInformedNN            INF-L2   Burgers/continuous_inference/Hwan_L2Regularization_Burgers.py:24-148
  PhysicsInformedNN_ADMM       INF-ADMM .../Hwan_L1Regularization_ADMM_Burgers.py:31-225
Dialect B (parameter object, trains inside the constructor; Wittmer):
  BurgersIdentification        AB-ADMM / AB-L2 / AB-L1 / ID-L2b / ID-ADMMb  Burgers/continuous_identification/*.py
  EulerInference               EUL      Eulers/continuous_inference/Euler_ADMM.py:36-437
"""
from __future__ import annotations

import os
import time
from typing import Optional, Sequence

import numpy as np

from .engine import Engine

# options passed verbatim to SciPy by the reference
LBFGS_OPTIONS_AB_ADMM = {'maxiter': 5000, 'maxfun': 50000, 'maxcor': 50, 'maxls': 50, 'ftol': 1e-7}  # AB-ADMM:68-72


def xavier_init_flat(layers: Sequence[int], rng: np.random.Generator) -> np.ndarray:
    """initialize_NN + xavier_init (INF-L2:79-94): truncated normal (re-draw beyond 2 sigma),
    stddev sqrt(2/(in+out)), zero biases; flat float32 vector in creation order W1,b1,..."""
    parts = []
    for l in range(len(layers) - 1):
        n_in, n_out = int(layers[l]), int(layers[l + 1])
        std = np.sqrt(2 / (n_in + n_out))
        w = rng.standard_normal(n_in * n_out)
        bad = np.abs(w) > 2.0
        while bad.any():
            w[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(w) > 2.0
        parts.append((w * std).astype(np.float32))
        parts.append(np.zeros(n_out, np.float32))
    return np.concatenate(parts)


def _first_gpu(GPU_number) -> int:
    """`visible_device_list` string of the reference ('3', '0,1') -> ordinal used by this process."""
    if GPU_number is None or GPU_number == '':
        return 0
    first = int(str(GPU_number).split(',')[0])
    try:
        import torch
        n = torch.cuda.device_count()
        return first if first < max(n, 1) else 0
    except Exception:
        return first


class _Base:
    """Shared plumbing: variables as numpy views, L-BFGS-B through SciPy, callbacks."""
    engine: Engine
    layers: Sequence[int]

    def _split(self, theta):
        ws, bs, off = [], [], 0
        for l in range(len(self.layers) - 1):
            n_in, n_out = self.layers[l], self.layers[l + 1]
            ws.append(theta[off:off + n_in * n_out].reshape(n_in, n_out)); off += n_in * n_out
            bs.append(theta[off:off + n_out].reshape(1, n_out)); off += n_out
        return ws, bs

    @property
    def weights(self):
        return self._split(self.engine.get_params())[0]

    @property
    def biases(self):
        return self._split(self.engine.get_params())[1]

    def get_flat_params(self) -> np.ndarray:
        return self.engine.get_params()

    def set_flat_params(self, theta):
        self.engine.set_params(theta)

    # graph-level callbacks of the reference (INF-L2:109-120): evaluated, not symbolic
    def net_u(self, x, t):
        X = np.hstack([np.asarray(x).reshape(-1, 1), np.asarray(t).reshape(-1, 1)])
        return self.engine.predict(X, want_f=False)[0]

    def net_f(self, x, t):
        X = np.hstack([np.asarray(x).reshape(-1, 1), np.asarray(t).reshape(-1, 1)])
        return self.engine.predict(X, want_f=True)[1]

    def neural_net(self, X, weights=None, biases=None):
        return self.engine.predict(np.asarray(X), want_f=False)[0]

    def lhs_collocation_on_device(self, N_f: int, seed: int = 1234, rank: int = 0, world: int = 1) -> np.ndarray:
        """`X_f_train = lb + (ub - lb) * lhs(2, N_f)` (INF-L2:183, INF-ADMM:270) drawn ON THE DEVICE: a Latin hypercube
        design of N_f points (pinn_sample_lhs: Philox offsets, Feistel stratum permutations); with world > 1 this rank
        takes its contiguous slice of the same design.  Becomes the engine's collocation batch; the host copy is returned
        for the class attributes x_f / t_f.  (The reference appends its IC/BC points to the design, INF-L2:184: callers
        that want them keep passing X_f through the constructor.)"""
        from .distributed import shard_range
        first, cnt = shard_range(int(N_f), rank, world)
        self.engine.sample_lhs(seed, cnt, first_index=first, n_design=int(N_f), nf_global=int(N_f))
        X = self.engine.get_collocation()
        self.x_f, self.t_f = X[:, 0:1], X[:, 1:2]
        return X

    def train_step_from_host(self, X_f_host, nf_global: int = 0, stepper=None) -> float:
        """One `sess.run(train_op_Adam, feed_dict)` with the collocation points fed from HOST memory
        (INF-L2:127-135 re-feeds them every step): H2D copy of X_f (float32 [N,2] torch tensor, ideally
        pinned), loss + gradient (+ allreduce when a DataParallelStepper is given) + Adam, and a D2H read of
        the step's residual loss.  Returns that loss term."""
        import torch
        eng = self.engine
        if getattr(self, "_loss_host", None) is None:
            self._loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
            self._packed = eng.packed_tensor()
        eng.feed_collocation(X_f_host, nf_global)   # asynchronous, chunked: the kernel starts on the first chunk
        if stepper is not None:
            stepper.adam_step()
        else:
            eng.loss_grad_device()
            eng.adam_apply()
        P = eng.num_params
        self._loss_host.copy_(self._packed[P + 3:P + 4], non_blocking=True)
        torch.cuda.current_stream(eng.device).synchronize()
        return float(self._loss_host[0])

    def loss_and_grad(self, theta=None):
        """The L-BFGS-B `loss_grad(x)` callback of ScipyOptimizerInterface (AB-ADMM:216)."""
        if theta is not None:
            self.engine.set_params(np.asarray(theta, np.float64))
        loss, grad = self.engine.loss_grad()
        return loss, grad.astype(np.float64)

    def lbfgs_minimize(self, options=None, loss_callback=None):
        """tf.contrib.opt.ScipyOptimizerInterface(loss, method='L-BFGS-B', options).minimize
        (AB-ADMM:66-72,:216): float64 iterate on the host, loss+grad from the GPU."""
        import scipy.optimize
        opts = dict(LBFGS_OPTIONS_AB_ADMM if options is None else options)
        n_theta = self.engine.num_params

        def fun(x):
            self.engine.set_params(x[:n_theta])
            if self.engine.trainable_lambda:
                self.engine.set_lambda(x[n_theta], x[n_theta + 1])
            loss, grad = self.engine.loss_grad()
            if loss_callback is not None:
                loss_callback(loss)
            return float(loss), grad.astype(np.float64)

        x0 = self.engine.get_params().astype(np.float64)
        if self.engine.trainable_lambda:
            x0 = np.concatenate([x0, np.asarray(self.engine.get_lambda(), np.float64)])
        res = scipy.optimize.minimize(fun, x0, jac=True, method='L-BFGS-B', options=opts)
        self.engine.set_params(res.x[:n_theta])
        if self.engine.trainable_lambda:
            self.engine.set_lambda(res.x[n_theta], res.x[n_theta + 1])
        return res


class PhysicsInformedNN(_Base):
    """INF-L2:24-148.  loss = ||u - u_pred||_2 + mean(f^2)  (:68-69), Adam 1e-3 (:72-73)."""

    def __init__(self, X_u, u, X_f, layers, lb, ub, nu, GPU_number='0', seed: int = 1234, theta0=None,
                 loss: str = "v1", verbose: bool = True):
        self.lb = lb
        self.ub = ub
        self.x_u = X_u[:, 0:1]
        self.t_u = X_u[:, 1:2]
        self.x_f = X_f[:, 0:1]
        self.t_f = X_f[:, 1:2]
        self.u = u
        self.layers = list(layers)
        self.nu = nu
        self.GPU_number = GPU_number
        self.verbose = verbose
        self.tol = 0.0001                                                   # INF-L2:74
        self.engine = Engine(self.layers, lb, ub, pde="burgers", loss=loss, lambda1=1.0, lambda2=nu,
                             device=_first_gpu(GPU_number))
        theta = xavier_init_flat(self.layers, np.random.default_rng(seed)) if theta0 is None else theta0
        self.engine.set_params(theta)                                       # global_variables_initializer (:76-77)
        self.engine.set_data(X_u, u)
        self.engine.set_collocation(X_f)
        self.engine.adam_config(lr=0.001)

    def callback(self, loss):                                               # INF-L2:122-123
        print('Loss: %.3e, GPU Number: %s\n' % (loss, self.GPU_number))

    def train(self, number_of_epochs, filename='', GPU_number=None):        # INF-L2:126-141
        GPU_number = self.GPU_number if GPU_number is None else GPU_number
        start_time = time.time()
        iter_counter = 0
        loss_value = 1000
        while iter_counter < number_of_epochs and abs(loss_value) > self.tol:
            if iter_counter % 100 == 0:
                self.engine.adam_steps(1)
                time_elapsed = time.time() - start_time
                loss_value = self.engine.loss_value()
                if self.verbose:
                    print('%s: \nIteration Number: %d, Loss: %.3e, Time Elapsed: %.2f, GPU Number: %s\n'
                          % (filename, iter_counter, loss_value, time_elapsed, GPU_number))
                start_time = time.time()
                iter_counter += 1
            else:  # the loss is only refreshed every 100 iterations: run the stretch on the device
                k = int(min(100 - iter_counter % 100, number_of_epochs - iter_counter))
                self.engine.adam_steps(k)
                iter_counter += k
        self.loss_value = loss_value
        return iter_counter

    def predict(self, X_star):                                              # INF-L2:143-148
        u_star, f_star = self.engine.predict(X_star, want_f=True)
        return u_star, f_star


class PhysicsInformedNN_ADMM(_Base):
    """INF-ADMM:31-225: L1 residual via ADMM; loss :98-100; z / multiplier updates :94-95,:106-107."""

    def __init__(self, X_u, u, X_f, layers, lb, ub, nu, lagrange_initial_guess, penalty_parameter, filename='',
                 GPU_number='0', seed: int = 1234, theta0=None, verbose: bool = True):
        self.filename = filename
        self.lb, self.ub = lb, ub
        self.x_u, self.t_u = X_u[:, 0:1], X_u[:, 1:2]
        self.x_f, self.t_f = X_f[:, 0:1], X_f[:, 1:2]
        self.u = u
        self.layers = list(layers)
        self.nu = nu
        self.N_u = self.x_u.shape[0]
        self.N_r = self.x_f.shape[0]
        self.verbose = verbose
        self.tol = 0.0001
        self.GPU_number = GPU_number
        # lagrange_initial_guess is accepted and ignored, as in the reference (:35,:88)
        self.engine = Engine(self.layers, lb, ub, pde="burgers", loss="v2", lambda1=1.0, lambda2=nu,
                             rho=penalty_parameter, device=_first_gpu(GPU_number))
        theta = xavier_init_flat(self.layers, np.random.default_rng(seed)) if theta0 is None else theta0
        self.engine.set_params(theta)
        self.engine.set_data(X_u, u)
        self.engine.set_collocation(X_f)
        self.engine.adam_config(lr=0.001)
        self.engine.admm_init()                                             # z <- r(w)  (:114-115)

    def callback(self, loss):
        print('Loss:', loss)

    def train(self, number_of_ADMM_iterations, number_of_w_optimization_steps, filename='', GPU_number=None):
        GPU_number = self.GPU_number if GPU_number is None else GPU_number  # INF-ADMM:180-200
        start_time = time.time()
        iter_counter = 0
        loss_value = 1000
        pending = False  # a z / multiplier update held back to ride in the next Adam step's pass (same residuals, same bits)
        while iter_counter < number_of_ADMM_iterations and abs(loss_value) > self.tol:
            if pending:
                self.engine.admm_adam_step(inf_admm_quirk=True)
                pending = False
            else:
                self.engine.adam_steps(1)
            if iter_counter % number_of_w_optimization_steps == 0:
                # z_update then lagrange_update; the graph quirk (:106-107) advances the multiplier twice
                pending = self._fold_admm
                if not pending:
                    self.engine.admm_update(inf_admm_quirk=True)
            if pending and iter_counter % 100 == 0:
                self.engine.admm_update(inf_admm_quirk=True)
                pending = False
            if iter_counter % 100 == 0:
                time_elapsed = time.time() - start_time
                loss_value = self.engine.loss_value()
                if self.verbose:
                    print('%s: \nIteration Number: %d, Loss: %.3e, Time Elapsed: %.2f, GPU Number: %s\n'
                          % (filename, iter_counter, loss_value, time_elapsed, GPU_number))
                start_time = time.time()
            iter_counter += 1
        if pending:
            self.engine.admm_update(inf_admm_quirk=True)
        self.loss_value = loss_value
        return iter_counter

    _fold_admm = True

    def soft_thresholding(self):
        return self.engine.admm_state()[0]

    def predict(self, X_star):
        return self.engine.predict(X_star, want_f=True)


class Parameters:
    """AB-ADMM:29-34 (class attributes overridden by positional argv in the drivers)."""
    N_u = 100
    N_f = 1000
    rho = 10.0
    epochs = 1e5
    gpu = '0'


class EulerParameters:
    """EUL:29-34."""
    N_data = 200
    N_f = 1000
    pen = 40.0
    epochs = 1e5
    gpu = '0'


_VARIANTS = {
    # name: (mat file, layers, loss, resample each step, ADMM, L-BFGS switch epoch)
    "AB-ADMM": ("TwoSin_burgers_shock.mat", [2] + [20] * 8 + [1], "v5", True, True, 50000),   # AB-ADMM:269-271,:213-216
    "AB-L2": ("Abgrall_burgers_shock.mat", [2] + [200] * 8 + [1], "v4", True, False, None),    # AB-L2:247-249
    "AB-L1": ("Abgrall_burgers_shock.mat", [2] + [200] * 8 + [1], "v3", True, False, None),    # AB-L1:237-239
    "ID-L2b": ("burgers_shock.mat", [2] + [20] * 8 + [1], "v3", False, False, None),           # ID-L2b:202-204,:166-169
    "ID-ADMMb": ("burgers_shock.mat", [2] + [20] * 8 + [1], "v5", True, True, None),           # ID-ADMMb:244-246
}


def _load_solution(data):
    if isinstance(data, dict):
        return data
    if str(data).endswith(".npz"):
        return dict(np.load(data))
    import scipy.io
    return scipy.io.loadmat(data)


class BurgersIdentification(_Base):
    """The five `continuous_identification` classes (they differ only in data file, width,
    loss and whether collocation points are re-drawn).  As in the
    reference the constructor loads data, builds and TRAINS (`run_NN`) unless run=False."""

    def __init__(self, params, variant: str = "AB-ADMM", data=None, lambda_1: Optional[float] = None,
                 lambda_2: Optional[float] = None, trainable_lambda: bool = False, run: bool = True, seed: int = 1234,
                 resample: str = "host", layers=None, verbose: bool = True, filename: Optional[str] = None,
                 theta0=None):
        self.params = params
        self.variant = variant
        mat, vlayers, loss, self._resample_each_step, self._admm, self._lbfgs_after = _VARIANTS[variant]
        self.layers = list(vlayers if layers is None else layers)
        self._resample_mode = resample
        self.verbose = verbose
        np.random.seed(seed)                                                 # AB-ADMM:25
        self.load_data(mat if data is None else data)
        if filename is not None:
            self.filename = filename
        self.tol = 1e-4
        self.N_u = self.params.N_u
        self.N_f = self.params.N_f
        if lambda_1 is None:
            lambda_1 = 1.0                                                   # AB-ADMM:105
        if lambda_2 is None:
            lambda_2 = 0.0031831 if variant.startswith("ID-") else 0.0       # ID-L2b:90 / AB-ADMM:106
        rho = float(getattr(self.params, "rho", 1.0))
        self.engine = Engine(self.layers, self.lb, self.ub, pde="burgers", loss=loss, lambda1=lambda_1,
                             lambda2=lambda_2, rho=rho, trainable_lambda=trainable_lambda,
                             device=_first_gpu(getattr(params, "gpu", '0')))
        self.engine.set_params(xavier_init_flat(self.layers, np.random.default_rng(seed)) if theta0 is None else theta0)
        self.engine.set_data(self.X_u_train, self.u_train)
        self.engine.adam_config(lr=0.001)
        self._step_counter = 0
        # randomly choose collocation points (AB-ADMM:91-93)
        self._new_batch()
        if self._admm:
            self.engine.admm_init()                                          # z = gamma = 1, z <- f(theta0) (:96-97)
        self.df = None
        if run:
            self.run_NN()

    # ---- data (AB-ADMM:264-309) ----
    def load_data(self, data):
        p = self.params
        self.filename = 'figures/%s_Nu%d_Nf%d_e%d.png' % (self.variant, p.N_u, p.N_f, int(p.epochs))
        sol = _load_solution(data)
        self.t = sol['t'].flatten()[:, None]
        self.x = sol['x'].flatten()[:, None]
        self.Exact = np.real(sol['usol']).T
        self.X, self.T = np.meshgrid(self.x, self.t)
        self.X_star = np.hstack((self.X.flatten()[:, None], self.T.flatten()[:, None]))
        self.u_star = self.Exact.flatten()[:, None]
        self.lb = self.X_star.min(0)
        self.ub = self.X_star.max(0)
        xx1 = np.hstack((self.X[0:1, :].T, self.T[0:1, :].T)); uu1 = self.Exact[0:1, :].T
        xx2 = np.hstack((self.X[:, 0:1], self.T[:, 0:1])); uu2 = self.Exact[:, 0:1]
        xx3 = np.hstack((self.X[:, -1:], self.T[:, -1:])); uu3 = self.Exact[:, -1:]
        self.X_u_train = np.vstack([xx1, xx2, xx3])
        self.u_train = np.vstack([uu1, uu2, uu3])
        idx = np.random.choice(self.X_u_train.shape[0], self.params.N_u, replace=False)
        self.X_u_train = self.X_u_train[idx, :]
        self.u_train = self.u_train[idx, :]
        self.x_data = self.X_u_train[:, 0:1]
        self.t_data = self.X_u_train[:, 1:2]
        self.u = self.u_train

    def _new_batch(self):
        """np.random.uniform(lb, ub, [N_f,1]) twice (AB-ADMM:220-221), or the device Philox sampler."""
        if self._resample_mode == "device":
            self.engine.sample_collocation(1234, self._step_counter * self.N_f, self.N_f)
        else:
            self.x_phys = np.random.uniform(self.lb[0], self.ub[0], [self.N_f, 1])
            self.t_phys = np.random.uniform(self.lb[1], self.ub[1], [self.N_f, 1])
            self.engine.set_collocation(np.hstack([self.x_phys, self.t_phys]))
        self._step_counter += 1

    def _plain_epochs(self, epoch, nEpochs):
        """How many epochs from `epoch` on end with nothing but the next batch (see train)."""
        if getattr(self, "_per_epoch_calls", False):
            return 0
        if self._resample_each_step and self._resample_mode != "device":
            return 0
        if self._admm and not (self._fold_admm and self._resample_each_step):
            return 0
        k = 0
        while epoch + k < nEpochs:
            e = epoch + k
            if self._lbfgs_after is not None and e > self._lbfgs_after:
                break
            every = 1000 if (self._lbfgs_after is None or e < self._lbfgs_after) else 100
            if e % 1000 == 0 or (self._record and e % every == 0):
                break
            k += 1
        return k

    def callback(self, loss, lambda_1=None, lambda_2=None):                 # AB-ADMM:182-183
        l1, l2 = self.engine.get_lambda()
        print('Loss: %e, l1: %.5f, l2: %.5f' % (loss, l1 if lambda_1 is None else lambda_1, l2 if lambda_2 is None else lambda_2))

    def compute_z(self):
        return self.engine.admm_state()[0]

    def train(self, nEpochs):                                                # AB-ADMM:200-252
        start_time = time.time()
        # the Abgrall scripts count `epoch = 1 .. nEpochs-1` (AB-ADMM:206-210), the two ID scripts
        # `it = 0 .. nIter-1` (ID-L2b:156-159, ID-ADMMb:195-198): one Adam step more for the same argument
        epoch = 0 if self.variant.startswith("ID-") else 1
        # The z/gamma update that closes an epoch (:225-226) and the Adam step that opens the next (:213) evaluate the
        # same residuals: the update is held back (`pending`) and folded into the next step's training pass
        # (engine.admm_adam_step, same bits), and flushed before anything else looks at the state.
        pending = False
        while epoch < nEpochs:
            # epochs at whose end nothing but the next batch happens (no print, no record, no L-BFGS) run as ONE call when the
            # batches come from the device sampler: same launches, same bits, no interpreter between them (the per-epoch host
            # overhead was as long as the ~25 us of GPU work)
            k = self._plain_epochs(epoch, nEpochs)
            if k > 0:
                if self._resample_each_step:
                    self.engine.resampled_epochs(k, self._admm, pending, 1234, self._step_counter, self.N_f)
                    self._step_counter += k
                    pending = self._admm
                else:
                    self.engine.adam_steps(k)
                epoch += k
                continue
            if self._lbfgs_after is None or epoch <= self._lbfgs_after:
                if pending:
                    self.engine.admm_adam_step()
                    pending = False
                else:
                    self.engine.adam_steps(1)
            else:
                if pending:
                    self.engine.admm_update()
                    pending = False
                self.lbfgs_minimize(LBFGS_OPTIONS_AB_ADMM)
            if self._resample_each_step:
                self._new_batch()
            if self._admm:
                pending = self._fold_admm                                     # z_update, then gamma_update (:225-226)
                if not pending:
                    self.engine.admm_update()
            every = 1000 if (self._lbfgs_after is None or epoch < self._lbfgs_after) else 100
            if pending and (epoch % 1000 == 0 or (self._record and epoch % every == 0)):
                self.engine.admm_update()
                pending = False
            if epoch % 1000 == 0:
                elapsed = time.time() - start_time
                loss_value = self.engine.loss_value()
                if self.verbose:
                    if self._admm:                                            # AB-ADMM:231-234 prints admm_misfit too
                        self.r_z = self.engine.admm_misfit()
                        print('It: %d, Loss: %.3e, r(w) - z: %.3f ,Time: %.2f' % (epoch, loss_value, self.r_z, elapsed))
                    else:
                        print('It: %d, Loss: %.3e, Time: %.2f' % (epoch, loss_value, elapsed))
                start_time = time.time()
            if self._record and epoch % every == 0:
                self.record_data(epoch)
                self.save_data()
            epoch += 1
        if pending:
            self.engine.admm_update()

    _record = False  # CSV dumps during training are opt-in here (the reference always writes them)
    _fold_admm = True  # fold each epoch's z/gamma update into the next Adam step's pass (False: two passes, same bits)

    def predict(self, X_star):                                               # AB-ADMM:254-262
        return self.engine.predict(X_star, want_f=True)

    def run_NN(self):                                                        # AB-ADMM:311-319
        self.train(self.params.epochs)
        self.record_data(self.params.epochs)
        if self._record:
            self.save_data()
        self.error_u = np.linalg.norm(self.u_star - self.u_pred_val, 2) / np.linalg.norm(self.u_star, 2)
        if self.verbose:
            print('Error u: %e %%' % (self.error_u * 100))

    def record_data(self, epoch_num):                                        # AB-ADMM:400-406
        import pandas as pd
        self.u_pred_val, self.f_pred_val = self.predict(self.X_star)
        x = self.X_star[:, 0]
        t = self.X_star[:, 1]
        epoch = np.ones(len(x)) * epoch_num
        self.df = pd.DataFrame({'x': x, 't': t, 'u_pred': self.u_pred_val[:, 0], 'epoch': epoch})

    def save_data(self):                                                     # AB-ADMM:408-409 (header re-emitted per append)
        os.makedirs(os.path.dirname(self.filename) or '.', exist_ok=True)
        self.df.to_csv(self.filename[:-3] + 'csv', mode='a', index=False)


class EulerInference(_Base):
    """EUL:36-437: 1-D compressible Euler, outputs (rho,u,E), three ADMM residual blocks."""

    def __init__(self, params, data=None, run: bool = True, seed: int = 1234, resample: str = "host", layers=None,
                 loss: str = "v5", verbose: bool = True, filename: Optional[str] = None, theta0=None):
        self.params = params
        self.verbose = verbose
        self._resample_mode = resample
        self.layers = list([2, 200, 200, 200, 200, 200, 3] if layers is None else layers)   # EUL:279
        np.random.seed(seed)
        self.load_data('Abgrall_eulers.mat' if data is None else data)
        if filename is not None:
            self.filename = filename
        self.tol = 1e-4
        self.N_data = self.params.N_data
        self.N_f = self.params.N_f
        self._admm = (loss == "v5")
        self.engine = Engine(self.layers, self.lb, self.ub, pde="euler", loss=loss, rho=float(self.params.pen),
                             device=_first_gpu(getattr(params, "gpu", '0')))
        self.engine.set_params(xavier_init_flat(self.layers, np.random.default_rng(seed)) if theta0 is None else theta0)
        self.engine.set_data(self.X_data_train, np.hstack([self.rho, self.u, self.E]))
        self.engine.adam_config(lr=0.001)
        self._step_counter = 0
        self._new_batch()                                                    # EUL:85-87
        if self._admm:
            self.engine.admm_init()                                          # EUL:89-92
        self.df = None
        if run:
            self.run_NN()

    def load_data(self, data):                                               # EUL:274-333
        params = self.params
        self.filename = 'figures/Euler_Nu%d_Nf%d_pen%d_e%d.png' % (params.N_data, params.N_f, int(params.pen), int(params.epochs))
        sol = _load_solution(data)
        self.t = sol['t'].flatten()[:, None]
        self.x = sol['x'].flatten()[:, None]
        self.Exact_rho = np.real(sol['rhosol']).T
        self.Exact_u = np.real(sol['usol']).T
        self.Exact_E = np.real(sol['Enersol']).T
        self.X, self.T = np.meshgrid(self.x, self.t)
        self.X_star = np.hstack((self.X.flatten()[:, None], self.T.flatten()[:, None]))
        self.rho_star = self.Exact_rho.flatten()[:, None]
        self.u_star = self.Exact_u.flatten()[:, None]
        self.E_star = self.Exact_E.flatten()[:, None]
        self.lb = self.X_star.min(0)
        self.ub = self.X_star.max(0)
        dom = np.vstack([np.hstack((self.X[0:1, :].T, self.T[0:1, :].T)), np.hstack((self.X[:, 0:1], self.T[:, 0:1])),
                         np.hstack((self.X[:, -1:], self.T[:, -1:]))])

        def stack(E):
            return np.vstack([E[0:1, :].T, E[:, 0:1], E[:, -1:]])
        idx = np.random.choice(dom.shape[0], self.params.N_data, replace=False)
        self.X_data_train = dom[idx, :]
        self.rho_train = stack(self.Exact_rho)[idx, :]
        self.u_train = stack(self.Exact_u)[idx, :]
        self.E_train = stack(self.Exact_E)[idx, :]
        self.x_data = self.X_data_train[:, 0:1]
        self.t_data = self.X_data_train[:, 1:2]
        self.rho, self.u, self.E = self.rho_train, self.u_train, self.E_train

    def _new_batch(self):                                                    # EUL:232-233
        if self._resample_mode == "device":
            self.engine.sample_collocation(1234, self._step_counter * self.N_f, self.N_f)
        else:
            self.x_phys = np.random.uniform(self.lb[0], self.ub[0], [self.N_f, 1])
            self.t_phys = np.random.uniform(self.lb[1], self.ub[1], [self.N_f, 1])
            self.engine.set_collocation(np.hstack([self.x_phys, self.t_phys]))
        self._step_counter += 1

    def net_rho_u_E(self, x, t):                                             # EUL:172-174
        return self.net_u(x, t)

    def callback(self, loss):                                                # EUL:200-201 (format string fixed)
        print('Loss: %e' % (loss,))

    def train(self, nEpochs):                                                # EUL:217-258
        start_time = time.time()
        epoch = 1
        pending = False   # a z/lagrange update waiting to be folded into the next Adam step's pass (same residuals, same bits)
        while epoch < nEpochs:
            # plain epochs (no print, no record at their end) with the device sampler: one call, same launches, same bits
            k = 0
            if self._resample_mode == "device" and (not self._admm or self._fold_admm) and not getattr(self, "_per_epoch_calls", False):
                while epoch + k < nEpochs and (epoch + k) % 1000 != 0 and not (self._record and (epoch + k) % 10000 == 0):
                    k += 1
            if k > 0:
                self.engine.resampled_epochs(k, self._admm, pending, 1234, self._step_counter, self.N_f)
                self._step_counter += k
                pending = self._admm
                epoch += k
                continue
            if pending:
                self.engine.admm_adam_step()
                pending = False
            else:
                self.engine.adam_steps(1)
            self._new_batch()
            if self._admm:
                pending = self._fold_admm                                     # z1..3 then lagrange1..3 (:237-242)
                if not pending:
                    self.engine.admm_update()
            if pending and (epoch % 1000 == 0 or (self._record and epoch % 10000 == 0)):
                self.engine.admm_update()
                pending = False
            if epoch % 1000 == 0:
                elapsed = time.time() - start_time
                loss_value = self.engine.loss_value()
                if self.verbose:
                    print('It: %d, Loss: %.3e, Time: %.2f' % (epoch, loss_value, elapsed))
                start_time = time.time()
            if self._record and epoch % 10000 == 0:
                self.record_data(epoch)
                self.save_data()
            epoch += 1
        if pending:
            self.engine.admm_update()

    _record = False
    _fold_admm = True

    def predict(self, X_star):                                               # EUL:260-272: six arrays
        y, f = self.engine.predict(X_star, want_f=True)
        return y[:, 0:1], y[:, 1:2], y[:, 2:3], f[:, 0:1], f[:, 1:2], f[:, 2:3]

    def run_NN(self):                                                        # EUL:335-347
        self.train(self.params.epochs)
        self.record_data(self.params.epochs)
        if self._record:
            self.save_data()
        self.error_rho = np.linalg.norm(self.rho_star - self.rho_pred_val, 2) / np.linalg.norm(self.rho_star, 2)
        self.error_u = np.linalg.norm(self.u_star - self.u_pred_val, 2) / np.linalg.norm(self.u_star, 2)
        self.error_E = np.linalg.norm(self.E_star - self.E_pred_val, 2) / np.linalg.norm(self.E_star, 2)
        if self.verbose:
            print('Error rho: %e %%' % (self.error_rho * 100))
            print('Error u: %e %%' % (self.error_u * 100))
            print('Error E: %e %%' % (self.error_E * 100))

    def record_data(self, epoch_num):                                        # EUL:428-434
        import pandas as pd
        (self.rho_pred_val, self.u_pred_val, self.E_pred_val,
         self.f1_pred_val, self.f2_pred_val, self.f3_pred_val) = self.predict(self.X_star)
        x = self.X_star[:, 0]
        t = self.X_star[:, 1]
        epoch = np.ones(len(x)) * epoch_num
        self.df = pd.DataFrame({'x': x, 't': t, 'rho_pred': self.rho_pred_val[:, 0], 'u_pred': self.u_pred_val[:, 0],
                                'E_pred': self.E_pred_val[:, 0], 'epoch': epoch})

    def save_data(self):                                                     # EUL:436-437
        os.makedirs(os.path.dirname(self.filename) or '.', exist_ok=True)
        self.df.to_csv(self.filename[:-3] + 'csv', mode='a', index=False)
