"""Op-for-op CPU restatement of the reference's TensorFlow-1 graph (torch autograd).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned against runs of the
reference's own unmodified scripts over a TensorFlow-1 API stand-in
(tests/golden/ref_*.npz, tests/test_reference_pin.py); TensorFlow itself is not
installed, so the semantics of the TF ops remain restated from documentation.

``tf.gradients(y, x)[0]`` is restated as
``torch.autograd.grad(y, x, grad_outputs=ones_like(y), create_graph=True)[0]``
(the same vector-Jacobian product with an all-ones seed), so ``net_f`` below has
the reference's reverse-over-reverse structure, not the Taylor-forward schedule
the CUDA kernels use.  Every function cites the reference lines it follows;
paths are relative to /root/reference and abbreviated as in SURVEY.md:

INF-L2    Burgers/continuous_inference/Hwan_L2Regularization_Burgers.py
INF-ADMM  Burgers/continuous_inference/Hwan_L1Regularization_ADMM_Burgers.py
ID-L2b    Burgers/continuous_identification/Burgers_batch_L2.py
ID-ADMMb  Burgers/continuous_identification/Burgers_ADMM_batch.py
AB-ADMM   Burgers/continuous_identification/Abgrall_ADMM.py
AB-L2     Burgers/continuous_identification/Abgrall_L2.py
AB-L1     Burgers/continuous_identification/Abgrall_L1.py
EUL       Eulers/continuous_inference/Euler_ADMM.py
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

# loss variants, SURVEY.md appendix A.3
LOSS_V1 = "v1_inf_l2"       # INF-L2:68-69   ||r||_2 + mean(f^2)
LOSS_V2 = "v2_inf_admm"     # INF-ADMM:98-100
LOSS_V3 = "v3_l1sq"         # ID-L2b:57-58, AB-L1:59-60
LOSS_V4 = "v4_mse"          # AB-L2:59-60
LOSS_V5 = "v5_admm"         # ID-ADMMb:118-119, AB-ADMM:129-130
LOSS_V6 = "v6_euler_admm"   # EUL:128-133
LOSS_EULER_MSE = "euler_mse"  # BASELINE config 3 "plain MSE" (AB-L2:59-60 applied to EUL's three residuals)

PDE_BURGERS = "burgers"
PDE_EULER = "euler"


@dataclass
class Problem:
    layers: Sequence[int]
    lb: np.ndarray                 # float64 [2]
    ub: np.ndarray                 # float64 [2]
    pde: str = PDE_BURGERS
    loss: str = LOSS_V4
    lam1: float = 1.0              # AB-ADMM:105
    lam2: float = 0.0              # = nu in the inference scripts (INF-L2:118)
    rho: float = 1.0               # ADMM penalty (rho / pen)
    n_out: int = field(init=False)

    def __post_init__(self):
        self.n_out = int(self.layers[-1])
        self.lb = np.asarray(self.lb, dtype=np.float64)
        self.ub = np.asarray(self.ub, dtype=np.float64)


def num_params(layers: Sequence[int]) -> int:
    return sum(layers[l] * layers[l + 1] + layers[l + 1] for l in range(len(layers) - 1))


def feed(a: np.ndarray, dtype=torch.float64) -> torch.Tensor:
    """Feed-time cast of a float64 host array to a float32 placeholder
    (INF-L2:58-63, :127-128), then widened to the oracle's working dtype."""
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).astype(np.float32))).to(dtype)


def unpack(theta: torch.Tensor, layers: Sequence[int]):
    """Flat parameter vector -> (weights, biases) in creation order W1,b1,...,WL,bL
    with W_l [in,out] row-major and b_l [1,out] (INF-L2:79-88)."""
    weights, biases = [], []
    off = 0
    for l in range(len(layers) - 1):
        n_in, n_out = layers[l], layers[l + 1]
        weights.append(theta[off:off + n_in * n_out].reshape(n_in, n_out))
        off += n_in * n_out
        biases.append(theta[off:off + n_out].reshape(1, n_out))
        off += n_out
    assert off == theta.numel()
    return weights, biases


def neural_net(X, weights, biases, lb, ub):
    """INF-L2:96-107 (EUL:160-170).  lb/ub are float64 numpy constants that TF
    converts to float32 graph constants; (ub - lb) is evaluated in numpy first."""
    dtype = X.dtype
    lb_c = torch.from_numpy(np.asarray(lb, np.float64).astype(np.float32)).to(dtype)
    span_c = torch.from_numpy((np.asarray(ub, np.float64) - np.asarray(lb, np.float64)).astype(np.float32)).to(dtype)
    num_layers = len(weights) + 1
    H = 2.0 * (X - lb_c) / span_c - 1.0
    for l in range(0, num_layers - 2):
        H = torch.tanh(torch.add(torch.matmul(H, weights[l]), biases[l]))
    Y = torch.add(torch.matmul(H, weights[-1]), biases[-1])
    return Y


def net_u(x, t, weights, biases, lb, ub):
    """INF-L2:109-111 / EUL:172-174."""
    return neural_net(torch.cat([x, t], 1), weights, biases, lb, ub)


def tf_gradients(y, x):
    """tf.gradients(y, x)[0]"""
    return torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), create_graph=True)[0]


def net_f_burgers(x, t, weights, biases, lb, ub, lam1, lam2):
    """INF-L2:113-120 (lam1 = 1, lam2 = nu) and AB-ADMM:170-180."""
    u = net_u(x, t, weights, biases, lb, ub)
    u_t = tf_gradients(u, t)
    u_x = tf_gradients(u, x)
    u_xx = tf_gradients(u_x, x)
    f = u_t + lam1 * u * u_x - lam2 * u_xx
    return f


def net_f_euler(x, t, weights, biases, lb, ub):
    """EUL:176-198."""
    rho_u_E = net_u(x, t, weights, biases, lb, ub)
    rho = rho_u_E[:, 0:1]
    u = rho_u_E[:, 1:2]
    E = rho_u_E[:, 2:3]
    gamma = 1.4
    p = (gamma - 1) * (E - (1 / 2) * rho * (u ** 2))

    rho_t = tf_gradients(rho, t)
    rhou_t = tf_gradients(rho * u, t)
    E_t = tf_gradients(E, t)

    rhou_x = tf_gradients(rho * u, x)
    rhouu2_x = tf_gradients(rho * (u ** 2), x)
    p_x = tf_gradients(p, x)
    uE_x = tf_gradients(u * E, x)
    up_x = tf_gradients(u * p, x)

    f1 = rho_t + rhou_x
    f2 = rhou_t + rhouu2_x + p_x
    f3 = E_t + uE_x + up_x
    return f1, f2, f3


def soft_threshold(f, gamma, rho, n_f):
    """compute_z / soft_thresholding: AB-ADMM:185-198, INF-ADMM:205-215, EUL:203-215."""
    kappa = 1.0 / (rho * n_f)
    val = f + gamma / rho
    cond1 = torch.where(val > kappa, torch.ones_like(val), torch.zeros_like(val))
    cond3 = torch.where(val < -1.0 * kappa, torch.ones_like(val), torch.zeros_like(val))
    return cond1 * (val - kappa) + cond3 * (val + kappa)


def _norm2(v):
    return torch.sqrt(torch.sum(v * v))


@dataclass
class Evaluation:
    loss: float
    grad: np.ndarray                 # flat d loss / d theta, W1,b1,... order
    dlam: np.ndarray                 # [2] d loss / d(lam1, lam2) (Burgers only)
    u_pred: np.ndarray               # [N_u, n_out]
    f: np.ndarray                    # [N_f, n_res]


def evaluate(theta: np.ndarray, prob: Problem, X_u: np.ndarray, u_data: np.ndarray, X_f: np.ndarray,
             z: Optional[np.ndarray] = None, gamma: Optional[np.ndarray] = None,
             dtype=torch.float64, want_grad: bool = True) -> Evaluation:
    """loss, residuals and the gradient the optimizers see, for every loss variant.

    theta: flat float parameters (cast through float32 like a tf.Variable).
    X_u [N_u,2], u_data [N_u,n_out], X_f [N_f,2]: float64 host arrays as the
    reference drivers build them; z/gamma: ADMM state [N_f, n_res] (float32 values).
    """
    th = torch.from_numpy(np.asarray(theta, np.float64).astype(np.float32)).to(dtype).requires_grad_(want_grad)
    weights, biases = unpack(th, prob.layers)
    lam = torch.tensor([np.float32(prob.lam1), np.float32(prob.lam2)], dtype=dtype, requires_grad=want_grad)

    x_u = feed(X_u[:, 0:1], dtype)
    t_u = feed(X_u[:, 1:2], dtype)
    u_d = feed(u_data, dtype)
    x_f = feed(X_f[:, 0:1], dtype).requires_grad_(True)
    t_f = feed(X_f[:, 1:2], dtype).requires_grad_(True)
    n_u = X_u.shape[0]
    n_f = X_f.shape[0]

    u_pred = net_u(x_u, t_u, weights, biases, prob.lb, prob.ub)
    if prob.pde == PDE_BURGERS:
        f = net_f_burgers(x_f, t_f, weights, biases, prob.lb, prob.ub, lam[0], lam[1])
        fs = [f]
    else:
        fs = list(net_f_euler(x_f, t_f, weights, biases, prob.lb, prob.ub))
    n_res = len(fs)

    if z is not None:
        z_t = [torch.from_numpy(np.asarray(z, np.float32)[:, k:k + 1].astype(np.float64)).to(dtype) for k in range(n_res)]
        g_t = [torch.from_numpy(np.asarray(gamma, np.float32)[:, k:k + 1].astype(np.float64)).to(dtype) for k in range(n_res)]
    rho = float(np.float32(prob.rho))

    r = u_d - u_pred
    L = prob.loss
    if L == LOSS_V1:
        loss = _norm2(r) + torch.mean(torch.square(fs[0]))
    elif L == LOSS_V2:
        loss = (1 / n_u) * _norm2(r) ** 2 + torch.sum(g_t[0] * fs[0]) + \
            (rho / 2) * _norm2(fs[0] - z_t[0] + g_t[0] / rho) ** 2
    elif L == LOSS_V3:
        loss = 1 / n_u * _norm2(r) ** 2 + 1 / n_f * torch.sum(torch.abs(fs[0])) ** 2
    elif L == LOSS_V4:
        loss = 1 / n_u * _norm2(r) ** 2 + 1 / n_f * _norm2(fs[0]) ** 2
    elif L == LOSS_V5:
        loss = 1 / n_u * _norm2(r) ** 2 + rho / 2 * _norm2(fs[0] - z_t[0] + g_t[0] / rho) ** 2
    elif L == LOSS_V6:
        loss = sum(1 / n_u * _norm2(r[:, q:q + 1]) ** 2 for q in range(3)) + \
            sum(rho / 2 * _norm2(fs[k] - z_t[k] + g_t[k] / rho) ** 2 for k in range(3))
    elif L == LOSS_EULER_MSE:
        loss = sum(1 / n_u * _norm2(r[:, q:q + 1]) ** 2 for q in range(3)) + \
            sum(1 / n_f * _norm2(fs[k]) ** 2 for k in range(3))
    else:
        raise ValueError(L)

    if want_grad:
        g_th, g_lam = torch.autograd.grad(loss, [th, lam], allow_unused=True)
        grad = g_th.detach().to(torch.float64).numpy()
        dlam = np.zeros(2) if g_lam is None else g_lam.detach().to(torch.float64).numpy()
    else:
        grad = np.zeros(th.numel())
        dlam = np.zeros(2)
    return Evaluation(
        loss=float(loss.detach()),
        grad=grad,
        dlam=dlam,
        u_pred=u_pred.detach().to(torch.float64).numpy(),
        f=torch.cat(fs, 1).detach().to(torch.float64).numpy(),
    )


def predict(theta: np.ndarray, prob: Problem, X_star: np.ndarray, dtype=torch.float64):
    """INF-L2:143-148 / AB-ADMM:254-262 / EUL:260-272: forward outputs and residuals on X_star."""
    th = torch.from_numpy(np.asarray(theta, np.float64).astype(np.float32)).to(dtype)
    weights, biases = unpack(th, prob.layers)
    x = feed(X_star[:, 0:1], dtype).requires_grad_(True)
    t = feed(X_star[:, 1:2], dtype).requires_grad_(True)
    y = net_u(x, t, weights, biases, prob.lb, prob.ub)
    if prob.pde == PDE_BURGERS:
        lam1 = float(np.float32(prob.lam1))
        lam2 = float(np.float32(prob.lam2))
        fs = [net_f_burgers(x, t, weights, biases, prob.lb, prob.ub, lam1, lam2)]
    else:
        fs = list(net_f_euler(x, t, weights, biases, prob.lb, prob.ub))
    return y.detach().numpy(), torch.cat(fs, 1).detach().numpy()


def admm_update(f: np.ndarray, z: np.ndarray, gamma: np.ndarray, rho: float, n_f: int, inf_admm_quirk: bool = False):
    """One z-update followed by one dual update on residuals f (all [N_f, n_res]).

    AB-ADMM:225-226 (z_update then gamma_update, both on the same f), EUL:237-242.
    inf_admm_quirk=True reproduces INF-ADMM:106-107,:191-193:
    running z_update first advances the multiplier with the old z, thresholds with
    the advanced multiplier, and lagrange_update then advances it again with the new z.
    """
    f_t = torch.from_numpy(np.asarray(f, np.float64))
    z_t = torch.from_numpy(np.asarray(z, np.float64))
    g_t = torch.from_numpy(np.asarray(gamma, np.float64))
    if inf_admm_quirk:
        g_t = g_t + rho * (f_t - z_t)
    z_new = soft_threshold(f_t, g_t, rho, n_f)
    g_new = g_t + rho * (f_t - z_new)
    return z_new.numpy(), g_new.numpy()


def xavier_init(layers: Sequence[int], rng: np.random.Generator) -> np.ndarray:
    """initialize_NN / xavier_init (INF-L2:79-94): W ~ truncated_normal(stddev =
    sqrt(2/(in+out))) with draws beyond two stddev re-drawn, b = 0; returned as the
    flat float32 vector in creation order.  TF's RNG stream is not reproducible
    without TF, so only the distribution is restated."""
    parts = []
    for l in range(len(layers) - 1):
        n_in, n_out = layers[l], layers[l + 1]
        std = np.sqrt(2 / (n_in + n_out))
        w = rng.standard_normal(n_in * n_out)
        bad = np.abs(w) > 2.0
        while bad.any():
            w[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(w) > 2.0
        parts.append((w * std).astype(np.float32))
        parts.append(np.zeros(n_out, np.float32))
    return np.concatenate(parts)


def relative_l2(exact: np.ndarray, pred: np.ndarray) -> float:
    """INF-L2:206, AB-ADMM:318, EUL:342-344."""
    return float(np.linalg.norm(exact - pred, 2) / np.linalg.norm(exact, 2))
