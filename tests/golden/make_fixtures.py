"""Regenerates everything under tests/golden/ (run in the build container, where /root/reference exists):

  data/*.npz          the reference's .mat solution fixtures (x, t, usol [, rhosol, Enersol]) re-saved as
                      compressed npz -- they are read-only INPUTS of the path, needed on
                      the GPU box where /root/reference does not exist
  vectors_*.npz       fp64 oracle outputs (loss, residuals, gradient, dlambda, 5-step TF-1 Adam trajectory) for
                      seeded inputs, one file per PDE x loss variant x width: they pin the oracle against silent
                      regressions and give the parity tests reference values that do not depend on torch
  e2e_*.json          oracle-trained final relative L2 errors for small end-to-end schedules

PARITY UNPINNED by the reference itself (it ships no golden vectors); these are produced by oracle/tf_graph.py,
which is cross-checked against oracle/taylor.py and finite differences in tests/test_oracle.py.
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import tf_graph as tg  # noqa: E402
from oracle.optim import TF1Adam  # noqa: E402
from tests.helpers import make_case  # noqa: E402

REF = "/root/reference"
MATS = {
    "burgers_shock": "Burgers/Data/burgers_shock.mat",
    "Abgrall_burgers_shock": "Burgers/Data/Abgrall_burgers_shock.mat",
    "TwoSin_burgers_shock": "Burgers/Data/TwoSin_burgers_shock.mat",
    "Abgrall_eulers": "Eulers/Data/Abgrall_eulers.mat",
}

VECTOR_CASES = [
    # name, pde, layers, loss, n_u, n_f
    ("burgers20_v1", tg.PDE_BURGERS, [2] + [20] * 8 + [1], tg.LOSS_V1, 100, 512),
    ("burgers20_v2", tg.PDE_BURGERS, [2] + [20] * 8 + [1], tg.LOSS_V2, 100, 512),
    ("burgers20_v3", tg.PDE_BURGERS, [2] + [20] * 8 + [1], tg.LOSS_V3, 100, 512),
    ("burgers20_v4", tg.PDE_BURGERS, [2] + [20] * 8 + [1], tg.LOSS_V4, 100, 512),
    ("burgers20_v5", tg.PDE_BURGERS, [2] + [20] * 8 + [1], tg.LOSS_V5, 100, 512),
    ("burgers128_v4", tg.PDE_BURGERS, [2] + [128] * 8 + [1], tg.LOSS_V4, 50, 128),
    ("burgers200_v4", tg.PDE_BURGERS, [2] + [200] * 8 + [1], tg.LOSS_V4, 50, 96),
    ("euler200_v6", tg.PDE_EULER, [2] + [200] * 5 + [3], tg.LOSS_V6, 64, 128),
    ("euler200_mse", tg.PDE_EULER, [2] + [200] * 5 + [3], tg.LOSS_EULER_MSE, 64, 128),
]


def save_data():
    import scipy.io
    for name, rel in MATS.items():
        d = scipy.io.loadmat(os.path.join(REF, rel))
        keep = {k: np.asarray(v) for k, v in d.items() if not k.startswith("__")}
        np.savez_compressed(os.path.join(HERE, "data", name + ".npz"), **keep)
        print("data", name, {k: v.shape for k, v in keep.items()})


def case_seed(name):
    return zlib.crc32(name.encode()) % 100000


def save_vectors():
    for name, pde, layers, loss, n_u, n_f in VECTOR_CASES:
        case = make_case(pde, layers, loss, n_u, n_f, seed=case_seed(name))
        ev = tg.evaluate(case["theta"], case["prob"], case["X_u"], case["u"], case["X_f"], case["z"], case["gamma"])
        # 5 TF-1 Adam steps in fp64 (theta round-tripped through float32 like a tf.Variable)
        theta = case["theta"].astype(np.float64)
        opt = TF1Adam(theta.size)
        losses = []
        for _ in range(5):
            e = tg.evaluate(theta, case["prob"], case["X_u"], case["u"], case["X_f"], case["z"], case["gamma"])
            losses.append(e.loss)
            theta = opt.step(theta, e.grad)
        big = theta.size > 20000
        rng = np.random.default_rng(1)
        idx = np.sort(rng.choice(theta.size, 4096, replace=False)) if big else np.arange(theta.size)
        np.savez_compressed(
            os.path.join(HERE, "vectors_%s.npz" % name),
            seed=case_seed(name), n_u=n_u, n_f=n_f, layers=np.asarray(layers), loss=ev.loss, f=ev.f, u_pred=ev.u_pred,
            dlam=ev.dlam, grad_idx=idx, grad=ev.grad[idx], grad_norm=np.linalg.norm(ev.grad),
            adam_losses=np.asarray(losses), adam_theta5=theta[idx], adam_theta5_norm=np.linalg.norm(theta))
        print("vectors", name, "loss", ev.loss, "|grad|", np.linalg.norm(ev.grad))


def e2e_schedule(full=False):
    """End-to-end schedules shared by the oracle (here) and the GPU tests: Burgers inference on burgers_shock, nu = 0.01/pi,
    MSE loss, TF-1 Adam then L-BFGS-B.  full=False: N_f = 2000+456, 1500 Adam steps, 400 L-BFGS iterations (seconds on a
    GPU, 2 minutes for the oracle).  full=True: BASELINE config 1 (N_u = 100, N_f = 10 000 + 456), 2000 Adam steps, then
    L-BFGS-B with the reference's options (AB-L2:68-72: maxiter 50000, ftol = eps), capped at 15000 iterations."""
    from oracle import data as odata
    sol = dict(np.load(os.path.join(HERE, "data", "burgers_shock.npz")))
    if full:
        g = odata.burgers_inference_inputs(sol, N_u=100, N_f=10000, seed=1234)
        layers = [2] + [20] * 8 + [1]
        theta0 = tg.xavier_init(layers, np.random.default_rng(1234))
        prob = tg.Problem(layers, g["lb"], g["ub"], pde=tg.PDE_BURGERS, loss=tg.LOSS_V4, lam1=1.0, lam2=0.01 / np.pi)
        return g, layers, theta0, prob, dict(adam_steps=2000, lbfgs={'maxiter': 15000, 'maxfun': 50000, 'maxcor': 50, 'maxls': 50,
                                                                    'ftol': 1.0 * np.finfo(float).eps})
    g = odata.burgers_inference_inputs(sol, N_u=100, N_f=2000, seed=1234)
    layers = [2] + [20] * 8 + [1]
    theta0 = tg.xavier_init(layers, np.random.default_rng(1234))
    prob = tg.Problem(layers, g["lb"], g["ub"], pde=tg.PDE_BURGERS, loss=tg.LOSS_V4, lam1=1.0, lam2=0.01 / np.pi)
    return g, layers, theta0, prob, dict(adam_steps=1500, lbfgs={'maxiter': 400, 'maxfun': 600, 'maxcor': 50, 'maxls': 50, 'ftol': 1e-9})


def save_e2e(full=False, fp32=False):
    import torch
    from oracle.optim import lbfgs_minimize
    torch.set_num_threads(6)
    g, layers, theta0, prob, sched = e2e_schedule(full)
    dt = torch.float32 if fp32 else torch.float64   # fp32 = the arithmetic of the reference's TF graph itself
    theta = theta0.astype(np.float64)
    opt = TF1Adam(theta.size)
    for it in range(sched["adam_steps"]):
        ev = tg.evaluate(theta, prob, g["X_u"], g["u"], g["X_f"], dtype=dt)
        theta = opt.step(theta, ev.grad)
    u_adam, _ = tg.predict(theta, prob, g["X_star"])
    err_adam = tg.relative_l2(g["u_star"], u_adam)
    loss_adam = tg.evaluate(theta, prob, g["X_u"], g["u"], g["X_f"], want_grad=False).loss

    def fun(x):
        e = tg.evaluate(x, prob, g["X_u"], g["u"], g["X_f"], dtype=dt)
        return e.loss, e.grad
    theta, res = lbfgs_minimize(fun, theta, sched["lbfgs"])
    u_fin, _ = tg.predict(theta, prob, g["X_star"])
    out = {"error_u_after_adam": err_adam, "loss_after_adam": loss_adam, "error_u_final": tg.relative_l2(g["u_star"], u_fin),
           "loss_final": float(res.fun), "lbfgs_nit": int(res.nit), "lbfgs_nfev": int(res.nfev)}
    out["oracle_dtype"] = "float32" if fp32 else "float64"
    out["lbfgs_message"] = str(res.message)
    json.dump(out, open(os.path.join(HERE, "e2e_burgers_inference%s%s.json" % ("_full" if full else "", "_fp32" if fp32 else "")), "w"), indent=1)
    print("e2e", out)


def trajectory_schedule(which):
    """Fixed-batch Adam schedules shared by the oracle (here) and tests/test_models_gpu.py for BASELINE configs 2 and 3.
    identification: burgers_shock, N_u = 2000 INTERIOR grid samples, N_f = 2000 uniform points,
                    loss AB-L2:59-60, lambda trainable from (0, ID-L2b:90's 0.0031831), 1500 TF-1 Adam steps.
    euler:          Abgrall_eulers, [2,200x5,3], N_data = 200 of the IC/BC points (EUL:316-326), N_f = 1000 uniform
                    points, pen = 40, plain-MSE residual loss, 300 Adam steps.
    The batch is held fixed (no per-epoch resampling) so that the trajectory is a deterministic function of the inputs."""
    from oracle import data as odata
    if which == "identification":
        sol = dict(np.load(os.path.join(HERE, "data", "burgers_shock.npz")))
        g = odata.burgers_identification_inputs(sol, N_u=100, N_f=2000, seed=1234)
        rng = np.random.default_rng(1234)
        idx = rng.choice(g["X_star"].shape[0], 2000, replace=False)
        g.update(X_u=g["X_star"][idx, :], u=g["u_star"][idx, :])
        layers = [2] + [20] * 8 + [1]
        theta0 = tg.xavier_init(layers, np.random.default_rng(4321))
        prob = tg.Problem(layers, g["lb"], g["ub"], pde=tg.PDE_BURGERS, loss=tg.LOSS_V4, lam1=0.0, lam2=0.0031831)
        return g, layers, theta0, prob, dict(adam_steps=1500, record=[0, 10, 100, 500, 1000, 1500])
    sol = dict(np.load(os.path.join(HERE, "data", "Abgrall_eulers.npz")))
    g = odata.euler_inputs(sol, N_data=200, N_f=1000, seed=1234)
    layers = [2] + [200] * 5 + [3]
    theta0 = tg.xavier_init(layers, np.random.default_rng(4321))
    prob = tg.Problem(layers, g["lb"], g["ub"], pde=tg.PDE_EULER, loss=tg.LOSS_EULER_MSE, rho=40.0)
    return g, layers, theta0, prob, dict(adam_steps=300, record=[0, 10, 50, 100, 200, 300])


def save_trajectory(which):
    import torch
    torch.set_num_threads(6)
    g, layers, theta0, prob, sched = trajectory_schedule(which)
    trainable = which == "identification"
    theta = theta0.astype(np.float64)
    lam = np.array([np.float32(prob.lam1), np.float32(prob.lam2)], np.float64)
    opt = TF1Adam(theta.size + (2 if trainable else 0))
    rec = {"steps": [], "loss": [], "lambda1": [], "lambda2": []}
    for it in range(sched["adam_steps"] + 1):
        pr = tg.Problem(layers, prob.lb, prob.ub, pde=prob.pde, loss=prob.loss, lam1=lam[0], lam2=lam[1], rho=prob.rho)
        ev = tg.evaluate(theta, pr, g["X_u"], g["u"], g["X_f"])
        if it in sched["record"]:
            rec["steps"].append(it); rec["loss"].append(float(ev.loss)); rec["lambda1"].append(float(lam[0])); rec["lambda2"].append(float(lam[1]))
        if it == sched["adam_steps"]:
            break
        if trainable:
            new = opt.step(np.concatenate([theta, lam]), np.concatenate([ev.grad, ev.dlam]))
            theta, lam = new[:-2], new[-2:]
        else:
            theta = opt.step(theta, ev.grad)
    pr = tg.Problem(layers, prob.lb, prob.ub, pde=prob.pde, loss=prob.loss, lam1=lam[0], lam2=lam[1], rho=prob.rho)
    pred, _ = tg.predict(theta, pr, g["X_star"])
    if which == "identification":
        rec["error_u"] = tg.relative_l2(g["u_star"], pred)
    else:
        rec["error_rho"] = tg.relative_l2(g["rho_star"], pred[:, 0:1])
        rec["error_u"] = tg.relative_l2(g["u_star"], pred[:, 1:2])
        rec["error_E"] = tg.relative_l2(g["E_star"], pred[:, 2:3])
    json.dump(rec, open(os.path.join(HERE, "trajectory_%s.json" % which), "w"), indent=1)
    print("trajectory", which, rec)


if __name__ == "__main__":
    what = sys.argv[1:] or ["data", "vectors"]
    if "trajectory_identification" in what:
        save_trajectory("identification")
    if "trajectory_euler" in what:
        save_trajectory("euler")
    if "data" in what:
        save_data()
    if "vectors" in what:
        save_vectors()
    if "e2e" in what:
        save_e2e()
    if "e2e_full" in what:
        save_e2e(True)
    if "e2e_full_fp32" in what:
        save_e2e(True, fp32=True)
