"""Oracle-trained FINAL accuracies of BASELINE configs 1-3 -- what "final relative L2 error within 10 % of the reference's"
(north_star) is measured against.  TEST INFRASTRUCTURE (imports oracle/ only).

The end point of an Adam / L-BFGS-B run is a chaotic function of rounding: the float64 and the float32 evaluation of the
SAME reference graph end config 1 at 4.3e-4 and 4.4e-3.  "The reference's final error" is therefore recorded as an
ENSEMBLE: the oracle's float32 graph (the arithmetic TensorFlow itself uses: tf.float32 variables and placeholders,
INF-L2:58-63,:85,:94) started from theta0 * (1 + 1e-7 * N(0,1)) for a few seeds (seed 0 = unperturbed) -- a perturbation
of one float32 rounding -- plus one float64 run.  tests/test_converged_gpu.py runs the CUDA path on the same schedules and
seeds and compares ensemble against ensemble.

    python tests/golden/make_converged.py identification|inference|euler_admm|euler_mse [dtype] [seed ...]

Schedules (shared with the GPU test through converged_schedule):
  inference       config 1: burgers_shock, N_u = 100, N_f = 10 000 LHS + 456 IC/BC points, nu = 0.01/pi, MSE loss, 2000 TF-1
                  Adam steps then L-BFGS-B with the reference's options (AB-L2:68-72), at most 15 000 iterations.
  identification  config 2: N_u = 2000 interior samples, N_f = 2000 fixed uniform points, loss AB-L2:59-60, lambda TRAINABLE
                  from (0, 0.0031831 -- ID-L2b:90), 2000 Adam steps then L-BFGS-B over (theta, lambda) with the options of
                  AB-ADMM:68-72 (at most 5000 iterations) and ftol = 1e-9.
  euler_admm      config 3: Abgrall_eulers, [2,200x5,3], N_data = 200, N_f = 1000 re-drawn every epoch (EUL:232-235), pen = 40,
                  the reference's ADMM loss and z/lagrange updates (EUL:128-141,:237-242), train(3000) = 2999 Adam epochs.
  euler_mse       the same with the plain-MSE residual loss.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import tf_graph as tg          # noqa: E402
from oracle.optim import TF1Adam, lbfgs_minimize   # noqa: E402

PERTURB = 1e-7
LBFGS_AB_L2 = {'maxfun': 50000, 'maxcor': 50, 'maxls': 50, 'ftol': 1.0 * np.finfo(float).eps}   # AB-L2:68-72 (+ maxiter below)
LBFGS_AB_ADMM = {'maxiter': 5000, 'maxfun': 50000, 'maxcor': 50, 'maxls': 50, 'ftol': 1e-7}       # AB-ADMM:68-72


def perturbed(theta0, seed):
    """seed 0: theta0 itself; else theta0 (1 + 1e-7 N(0,1)) rounded to float32 -- a one-rounding perturbation."""
    theta0 = np.asarray(theta0, np.float32)
    if seed == 0:
        return theta0
    rng = np.random.default_rng(1000 + seed)
    return (theta0.astype(np.float64) * (1.0 + PERTURB * rng.standard_normal(theta0.size))).astype(np.float32)


def converged_schedule(which):
    from oracle import data as odata
    from tests.golden.make_fixtures import e2e_schedule, trajectory_schedule
    if which == "inference":
        g, layers, theta0, prob, sched = e2e_schedule(full=True)
        return g, layers, theta0, prob, sched
    if which == "identification":
        g, layers, theta0, prob, _ = trajectory_schedule("identification")
        # L-BFGS-B with the options of Abgrall_ADMM.py:68-72 but ftol 1e-9: the run stops when the loss stalls, before
        # float32 rounding noise defeats the line search -- with Abgrall_L2.py's ftol = eps the float32 runs stop anywhere
        # between 2400 and 3700 iterations (error_u 0.032 ... 0.039, lambda1 0.957 ... 0.969; float64: 10 000 iterations,
        # 0.015, 0.999) and no two float32 evaluation orders end in the same place
        # (AB-ADMM's own ftol = 1e-7 is relative to max(|loss|, 1): with a loss of 1e-2 it stops L-BFGS-B at its first
        # iterations; 1e-9 sits between that and the float32 noise of the loss, ~1e-10)
        return g, layers, theta0, prob, dict(adam_steps=2000, lbfgs=dict(LBFGS_AB_ADMM, ftol=1e-9))
    sol = dict(np.load(os.path.join(HERE, "data", "Abgrall_eulers.npz")))
    g = odata.euler_inputs(sol, N_data=200, N_f=1000, seed=1234)   # seeds numpy's legacy RNG; later batches continue its stream
    layers = [2] + [200] * 5 + [3]
    theta0 = tg.xavier_init(layers, np.random.default_rng(4321))
    loss = tg.LOSS_V6 if which == "euler_admm" else tg.LOSS_EULER_MSE
    prob = tg.Problem(layers, g["lb"], g["ub"], pde=tg.PDE_EULER, loss=loss, rho=40.0)
    return g, layers, theta0, prob, dict(epochs=3000)


def run(which, dtype_name, seed, threads=3):
    import torch
    torch.set_num_threads(threads)
    dt = torch.float32 if dtype_name == "float32" else torch.float64
    ndt = np.float32 if dtype_name == "float32" else np.float64
    g, layers, theta0, prob, sched = converged_schedule(which)
    if os.environ.get("CONVERGED_SMOKE"):   # plumbing check only
        sched = dict(adam_steps=3, lbfgs=dict(LBFGS_AB_L2, maxiter=3), epochs=4)
    theta = perturbed(theta0, seed).astype(np.float64)
    out = {"which": which, "oracle_dtype": dtype_name, "perturb_seed": seed}
    t0 = time.time()
    if which == "inference":
        opt = TF1Adam(theta.size, dtype=ndt)
        for _ in range(sched["adam_steps"]):
            theta = opt.step(theta, tg.evaluate(theta, prob, g["X_u"], g["u"], g["X_f"], dtype=dt).grad).astype(np.float64)
        out["error_u_after_adam"] = tg.relative_l2(g["u_star"], tg.predict(theta, prob, g["X_star"])[0])

        def fun(x):
            e = tg.evaluate(x, prob, g["X_u"], g["u"], g["X_f"], dtype=dt)
            return e.loss, e.grad
        theta, res = lbfgs_minimize(fun, theta, sched["lbfgs"])
        out.update(error_u=tg.relative_l2(g["u_star"], tg.predict(theta, prob, g["X_star"])[0]), loss_final=float(res.fun),
                   lbfgs_nit=int(res.nit), lbfgs_nfev=int(res.nfev))
    elif which == "identification":
        lam = np.array([np.float32(prob.lam1), np.float32(prob.lam2)], np.float64)
        opt = TF1Adam(theta.size + 2, dtype=ndt)

        def problem(lam):
            return tg.Problem(layers, prob.lb, prob.ub, pde=prob.pde, loss=prob.loss, lam1=lam[0], lam2=lam[1], rho=prob.rho)
        for _ in range(sched["adam_steps"]):
            ev = tg.evaluate(theta, problem(lam), g["X_u"], g["u"], g["X_f"], dtype=dt)
            new = opt.step(np.concatenate([theta, lam]), np.concatenate([ev.grad, ev.dlam])).astype(np.float64)
            theta, lam = new[:-2], new[-2:]
        out.update(error_u_after_adam=tg.relative_l2(g["u_star"], tg.predict(theta, problem(lam), g["X_star"])[0]),
                   lambda1_after_adam=float(lam[0]), lambda2_after_adam=float(lam[1]))

        def fun(x):
            e = tg.evaluate(x[:-2], problem(x[-2:]), g["X_u"], g["u"], g["X_f"], dtype=dt)
            return e.loss, np.concatenate([e.grad, e.dlam])
        x, res = lbfgs_minimize(fun, np.concatenate([theta, lam]), sched["lbfgs"])
        theta, lam = x[:-2], x[-2:]
        out.update(error_u=tg.relative_l2(g["u_star"], tg.predict(theta, problem(lam), g["X_star"])[0]),
                   lambda1=float(np.float32(lam[0])), lambda2=float(np.float32(lam[1])), loss_final=float(res.fun),
                   lbfgs_nit=int(res.nit), lbfgs_nfev=int(res.nfev))
    else:
        admm = which == "euler_admm"
        X_u, u_data, X_f = g["X_u"], g["u"], g["X_f"]
        n_f = X_f.shape[0]
        z = gamma = None
        if admm:                                                   # EUL:114-141,:89-92
            gamma = np.ones((n_f, 3))
            z = tg.evaluate(theta, prob, X_u, u_data, X_f, z=gamma, gamma=gamma, want_grad=False, dtype=dt).f
        opt = TF1Adam(theta.size, dtype=ndt)
        it = 1
        while it < sched["epochs"]:                                # EUL:217-258
            ev = tg.evaluate(theta, prob, X_u, u_data, X_f, z=z, gamma=gamma, dtype=dt)
            theta = opt.step(theta, ev.grad).astype(np.float64)
            X_f = np.hstack([np.random.uniform(g["lb"][0], g["ub"][0], [n_f, 1]), np.random.uniform(g["lb"][1], g["ub"][1], [n_f, 1])])
            if admm:
                f = tg.evaluate(theta, prob, X_u, u_data, X_f, z=z, gamma=gamma, want_grad=False, dtype=dt).f
                z, gamma = tg.admm_update(f, z, gamma, prob.rho, n_f)
            it += 1
        pred = tg.predict(theta, prob, g["X_star"])[0]
        out.update(error_rho=tg.relative_l2(g["rho_star"], pred[:, 0:1]), error_u=tg.relative_l2(g["u_star"], pred[:, 1:2]),
                   error_E=tg.relative_l2(g["E_star"], pred[:, 2:3]),
                   loss_final=float(tg.evaluate(theta, prob, X_u, u_data, X_f, z=z, gamma=gamma, want_grad=False).loss))
    out["cpu_seconds"] = round(time.time() - t0, 1)
    return out


def record(entry):
    """appends to tests/golden/converged_<which>.json (one list of runs per config)"""
    path = os.path.join(HERE, "converged_%s.json" % entry["which"])
    runs = json.load(open(path)) if os.path.exists(path) else []
    runs = [r for r in runs if not (r["oracle_dtype"] == entry["oracle_dtype"] and r["perturb_seed"] == entry["perturb_seed"])]
    runs.append(entry)
    runs.sort(key=lambda r: (r["oracle_dtype"], r["perturb_seed"]))
    json.dump(runs, open(path, "w"), indent=1)


if __name__ == "__main__":
    which = sys.argv[1]
    dtype_name = sys.argv[2] if len(sys.argv) > 2 else "float32"
    seeds = [int(s) for s in sys.argv[3:]] or [0]
    for s in seeds:
        e = run(which, dtype_name, s)
        record(e)
        print(json.dumps(e), flush=True)
