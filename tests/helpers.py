"""Shared seeded problem builders for the oracle and parity tests."""
import numpy as np

from oracle import tf_graph as tg


def rand_theta(layers, rng, bias_scale=0.1):
    """Xavier weights plus NON-zero biases so every tanh'/''/''' path is exercised."""
    th = tg.xavier_init(layers, rng).astype(np.float64)
    off = 0
    for l in range(len(layers) - 1):
        off += layers[l] * layers[l + 1]
        th[off:off + layers[l + 1]] = bias_scale * rng.standard_normal(layers[l + 1])
        off += layers[l + 1]
    return th.astype(np.float32)


def make_case(pde, layers, loss, n_u, n_f, seed, lam1=1.0, lam2=0.01 / np.pi, rho=3.0):
    rng = np.random.default_rng(seed)
    lb = np.array([-1.0, 0.0]); ub = np.array([1.0, 0.99])
    th = rand_theta(layers, rng)
    X_u = lb + (ub - lb) * rng.random((n_u, 2))
    u = rng.standard_normal((n_u, layers[-1]))
    X_f = lb + (ub - lb) * rng.random((n_f, 2))
    n_res = 1 if pde == tg.PDE_BURGERS else 3
    z = (0.5 * rng.standard_normal((n_f, n_res))).astype(np.float32)
    g = (0.5 * rng.standard_normal((n_f, n_res))).astype(np.float32)
    prob = tg.Problem(layers, lb, ub, pde=pde, loss=loss, lam1=lam1, lam2=lam2, rho=rho)
    return dict(prob=prob, theta=th, X_u=X_u, u=u, X_f=X_f, z=z, gamma=g)


ENGINE_LOSS = {tg.LOSS_V1: "v1", tg.LOSS_V2: "v2", tg.LOSS_V3: "v3", tg.LOSS_V4: "v4", tg.LOSS_V5: "v5",
               tg.LOSS_V6: "v5", tg.LOSS_EULER_MSE: "v4"}


def make_engine(case, path="auto", trainable_lambda=False):
    from pinns_b200 import Engine
    p = case["prob"]
    eng = Engine(p.layers, p.lb, p.ub, pde=p.pde, loss=ENGINE_LOSS[p.loss], lambda1=p.lam1, lambda2=p.lam2, rho=p.rho,
                 trainable_lambda=trainable_lambda, path=path)
    eng.set_params(case["theta"])
    eng.set_data(case["X_u"], case["u"])
    eng.set_collocation(case["X_f"])
    if p.loss in (tg.LOSS_V2, tg.LOSS_V5, tg.LOSS_V6):
        eng.admm_set_state(case["z"], case["gamma"])
    return eng


def rel_err(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def max_rel_err(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
