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


# ---------------------------------------------------------------------------------------------------------------
# Fixtures made by running the reference's own scripts (tests/golden/make_ref_fixtures.py -> ref_<SCRIPT>.npz)
# ---------------------------------------------------------------------------------------------------------------
import json
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# script -> (solution file, pde, loss, first loop index, re-draw collocation points every step, ADMM)
# loop bounds: ID-L2b:159-185 and ID-ADMMb run `it = 0 .. nIter-1`, the Abgrall/Euler scripts `epoch = 1 .. nEpochs-1`
REF_RUNS = {
    "INF-L2": ("Abgrall_burgers_shock", tg.PDE_BURGERS, tg.LOSS_V1, 0, False, False),
    "INF-ADMM": ("burgers_shock", tg.PDE_BURGERS, tg.LOSS_V2, 0, False, True),
    "ID-L2b": ("burgers_shock", tg.PDE_BURGERS, tg.LOSS_V3, 0, False, False),
    "ID-ADMMb": ("burgers_shock", tg.PDE_BURGERS, tg.LOSS_V5, 0, True, True),
    "AB-ADMM": ("TwoSin_burgers_shock", tg.PDE_BURGERS, tg.LOSS_V5, 1, True, True),
    "AB-L2": ("Abgrall_burgers_shock", tg.PDE_BURGERS, tg.LOSS_V4, 1, True, False),
    "AB-L1": ("Abgrall_burgers_shock", tg.PDE_BURGERS, tg.LOSS_V3, 1, True, False),
    "EUL": ("Abgrall_eulers", tg.PDE_EULER, tg.LOSS_V6, 1, True, True),
}


def load_ref_fixture(name):
    fx = dict(np.load(os.path.join(GOLD, "ref_%s.npz" % name), allow_pickle=False))
    fx["meta"] = json.loads(str(fx["meta"]))
    if "theta0" not in fx:          # width-200 nets: regenerated from the shim's seeded initialiser
        from oracle.run_reference import shim_initial_theta
        fx["theta0"] = shim_initial_theta([int(n) for n in fx["layers"]])
    return fx


def ref_problem(name, fx):
    _, pde, loss, _, _, _ = REF_RUNS[name]
    meta = fx["meta"]
    if meta["dialect"] == "A":
        lam1, lam2 = 1.0, float(fx["nu"])
        rho = float(fx["penalty_parameter"]) if "penalty_parameter" in fx else 1.0
    else:
        lam1, lam2 = (float(fx["lambda"][0]), float(fx["lambda"][1])) if "lambda" in fx else (1.0, 0.0)
        rho = float(meta["params"].get("rho", meta["params"].get("pen", 1.0)))
    return tg.Problem([int(n) for n in fx["layers"]], fx["lb"], fx["ub"], pde=pde, loss=loss, lam1=lam1, lam2=lam2, rho=rho)


def oracle_replay(name, fx):
    """Replays the run the reference script made (see make_ref_fixtures.RUNS) with the oracle's restatements only:
    oracle.data for the RNG-ordered data preparation, tf_graph.evaluate for loss/gradient/residuals, optim.TF1Adam,
    tf_graph.admm_update.  Returns {stage: dict(theta, z, gamma, pred)}; `pred` on every pred_stride-th grid point."""
    from oracle import data as odata
    from oracle.optim import TF1Adam
    sol_name, pde, loss, first, resample, admm = REF_RUNS[name]
    sol = dict(np.load(os.path.join(GOLD, "data", sol_name + ".npz")))
    meta = fx["meta"]
    prob = ref_problem(name, fx)
    stride = meta["pred_stride"]
    theta = np.asarray(fx["theta0"], np.float64)
    adam = TF1Adam(theta.size)
    out = {}
    if meta["dialect"] == "A":
        g = odata.burgers_inference_inputs(sol, N_u=100, N_f=10000)
        X_u, u_data, X_f = g["X_u"], g["u"], g["X_f"]
        n_f = X_f.shape[0]
        if admm:
            gamma = np.ones((n_f, 1))
            z = tg.evaluate(theta, prob, X_u, u_data, X_f, z=gamma, gamma=gamma, want_grad=False).f   # INF-ADMM:114-115
        else:
            z = gamma = None
        for k, args in enumerate(meta["stages"], start=1):
            epochs, wsteps = (args + [None])[:2]
            it, loss_value = 0, 1000
            while it < epochs and abs(loss_value) > 1e-4:
                ev = tg.evaluate(theta, prob, X_u, u_data, X_f, z=z, gamma=gamma)
                theta = adam.step(theta, ev.grad)
                if admm and it % wsteps == 0:
                    f = tg.evaluate(theta, prob, X_u, u_data, X_f, z=z, gamma=gamma, want_grad=False).f
                    z, gamma = tg.admm_update(f, z, gamma, prob.rho, n_f, inf_admm_quirk=True)
                if it % 100 == 0:
                    loss_value = tg.evaluate(theta, prob, X_u, u_data, X_f, z=z, gamma=gamma, want_grad=False).loss
                it += 1
            y, f = tg.predict(theta, prob, g["X_star"][::stride])
            out[k] = dict(theta=theta.copy(), z=z, gamma=gamma, pred=np.hstack([y, f]), X_u=X_u, X_f=X_f, u_data=u_data)
        return out
    params = meta["params"]
    if pde == tg.PDE_EULER:
        g = odata.euler_inputs(sol, N_data=params["N_data"], N_f=params["N_f"])
    else:
        g = odata.burgers_identification_inputs(sol, N_u=params["N_u"], N_f=params["N_f"])
    X_u, u_data, X_f = g["X_u"], g["u"], g["X_f"]
    n_f = params["N_f"]
    n_res = 3 if pde == tg.PDE_EULER else 1
    if admm:
        gamma = np.ones((n_f, n_res))
        z = tg.evaluate(theta, prob, X_u, u_data, X_f, z=gamma, gamma=gamma, want_grad=False).f          # AB-ADMM:96-97
    else:
        z = gamma = None
    for k, n in enumerate(meta["stages"], start=1):
        it = first
        while it < n:
            ev = tg.evaluate(theta, prob, X_u, u_data, X_f, z=z, gamma=gamma)
            theta = adam.step(theta, ev.grad)
            if resample:
                x_phys = np.random.uniform(g["lb"][0], g["ub"][0], [n_f, 1])
                t_phys = np.random.uniform(g["lb"][1], g["ub"][1], [n_f, 1])
                X_f = np.hstack([x_phys, t_phys])
            if admm:
                f = tg.evaluate(theta, prob, X_u, u_data, X_f, z=z, gamma=gamma, want_grad=False).f
                z, gamma = tg.admm_update(f, z, gamma, prob.rho, n_f)
            it += 1
        y, f = tg.predict(theta, prob, g["X_star"][::stride])
        out[k] = dict(theta=theta.copy(), z=z, gamma=gamma, pred=np.hstack([y, f]), X_u=X_u, X_f=X_f, u_data=u_data)
    return out
