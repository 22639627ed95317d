"""Fixtures produced by RUNNING THE REFERENCE'S OWN, UNMODIFIED MODEL SCRIPTS (build container only).

    python tests/golden/make_ref_fixtures.py            -> tests/golden/ref_<SCRIPT>.npz for all eight scripts
    python tests/golden/make_ref_fixtures.py EUL AB-L1  -> only those

Each of the eight `PhysicsInformedNN` scripts under /root/reference is executed as written through
oracle/run_reference.py (TensorFlow-1 API, pyDOE and matplotlib replaced by oracle/refshim/; NumPy's legacy RNG,
SciPy and pandas are the real ones): the driver block / constructor loads the .mat data, draws the training sets,
builds the graph, trains a few iterations and predicts on the evaluation grid.  The graph is evaluated in float64
with TF's float32 rounding of every feed, constant and initial value, so the fixture carries no fp32 evaluation
noise.  Recorded per script:

  theta0                 the initial parameters (flat W1,b1,...; float32-exact) the run started from
  X_u, u_data            the training data the script selected (np.random.choice stream of seed 1234)
  stageK_theta/_z/_gamma parameters and ADMM state after each stage of the run (stage 1 = what the unmodified driver /
  stageK_pred, _error    constructor does; later stages = further calls of the reference's own train()), its predict()
                         outputs on the evaluation grid (every PRED_STRIDE-th point) and relative L2 errors
  vec_*                  at the final state snapped to float32 (= float32(stage<last>_theta)) and the last batch: loss, d loss / d theta
                         (tf.gradients on the reference's loss tensor), residuals f_pred, u_pred -- the inputs of the
                         loss+grad parity tests
  csv_header, csv_rows   what record_data/save_data wrote (Dialect B)
  lbfgs_*                AB-ADMM only: the reference's ScipyOptimizerInterface object run for LBFGS_ITERS iterations from vec_*

/root/reference does not exist on the GPU box: tests read these files, never the reference.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import run_reference as rr  # noqa: E402

PRED_STRIDE = 7
LBFGS_ITERS = 25

# how each script is driven.  Dialect A drivers hard-code their arguments (1 epoch); stage 2 calls train() again.
# Dialect B: positional argv of the driver block where it parses one, else the launch_NN_L2.py way (import the
# module, override Parameters attributes, construct).
RUNS = {
    "INF-L2": dict(dialect="A", more=[(4,)]),
    "INF-ADMM": dict(dialect="A", more=[(3, 1), (4, 2)]),
    "ID-L2b": dict(dialect="B", params=dict(N_u=100, N_f=1000, epochs=4), more=[3]),
    "ID-ADMMb": dict(dialect="B", params=dict(N_u=100, N_f=1000, rho=40.0, epochs=5), more=[3]),
    "AB-ADMM": dict(dialect="B", argv=[100, 1000, 10.0, 5, "0"], more=[3]),
    "AB-L2": dict(dialect="B", argv=[60, 256, 4, "0"], more=[]),
    "AB-L1": dict(dialect="B", argv=[60, 256, 4, "0"], more=[]),
    "EUL": dict(dialect="B", argv=[200, 1000, 40.0, 4, "0"], more=[3]),
}


def _f32(a):
    return np.asarray(a, np.float64).astype(np.float32)


def _admm_vars(m, tf):
    """(z variables, multiplier variables) of a reference model object, [] when the script has none."""
    if hasattr(m, "z1"):
        return [m.z1, m.z2, m.z3], [m.lagrange1, m.lagrange2, m.lagrange3]
    if hasattr(m, "z") and hasattr(m, "gamma"):
        return [m.z], [m.gamma]
    if hasattr(m, "z"):
        # INF-ADMM re-binds self.lagrange to its assign op (:106); the variable is the op's target
        lag = [v for v in tf.global_variables() if not v.trainable][0]
        return [m.z], [lag]
    return [], []


def _state(m, tf, out, tag):
    out[tag + "_theta"] = rr.flat_params(m.sess, m.weights, m.biases)
    zs, gs = _admm_vars(m, tf)
    if zs:
        out[tag + "_z"] = np.hstack([np.asarray(v._value.double().numpy()) for v in zs])
        out[tag + "_gamma"] = np.hstack([np.asarray(v._value.double().numpy()) for v in gs])


def _feed(m):
    if hasattr(m, "x_u_tf"):        # Dialect A
        return {m.x_u_tf: m.x_u, m.t_u_tf: m.t_u, m.u_tf: m.u, m.x_f_tf: m.x_f, m.t_f_tf: m.t_f}
    d = {m.x_data_tf: m.x_data, m.t_data_tf: m.t_data, m.x_phys_tf: m.x_phys, m.t_phys_tf: m.t_phys}
    if hasattr(m, "rho_tf"):
        d.update({m.rho_tf: m.rho, m.u_tf: m.u, m.E_tf: m.E})
    else:
        d[m.u_tf] = m.u
    return d


def _vectors(m, tf, out):
    """Snap the state to float32 (what TF variables hold), then evaluate the reference's own tensors."""
    variables = [v for pair in zip(m.weights, m.biases) for v in pair]
    zs, gs = _admm_vars(m, tf)
    for v in variables + zs + gs:
        v.load(_f32(v._value.double().numpy()))
    fd = _feed(m)
    # (the snapped parameters are float32(stage<last>_theta); not stored twice)
    if zs:
        out["vec_z"] = np.hstack([_f32(v._value.double().numpy()) for v in zs])
        out["vec_gamma"] = np.hstack([_f32(v._value.double().numpy()) for v in gs])
    if hasattr(m, "x_u_tf"):
        out["vec_X_f"] = np.hstack([m.x_f, m.t_f])
    else:
        out["vec_X_f"] = np.hstack([m.x_phys, m.t_phys])
    out["vec_loss"] = np.float64(np.asarray(m.sess.run(m.loss, fd)).reshape(-1)[0])
    grads = m.sess.run(tf.gradients(m.loss, variables), fd)
    out["vec_grad"] = np.concatenate([np.asarray(g, np.float64).ravel() for g in grads])
    if hasattr(m, "f1_pred"):
        out["vec_f"] = np.hstack(m.sess.run([m.f1_pred, m.f2_pred, m.f3_pred], fd))
        out["vec_u_pred"] = np.hstack(m.sess.run([m.rho_pred, m.u_pred, m.E_pred], fd))
    else:
        out["vec_f"] = m.sess.run(m.f_pred, fd)
        out["vec_u_pred"] = m.sess.run(m.u_pred, fd)
    if hasattr(m, "admm_misfit"):
        out["vec_admm_misfit"] = np.float64(m.sess.run(m.admm_misfit, fd))


def _pred(arrays):
    return np.hstack([np.asarray(a, np.float64) for a in arrays])[::PRED_STRIDE]


def run_dialect_a(name, spec):
    g, tf = rr.run_script(name, compute="float64")
    m = g["model"]
    out = dict(theta0=rr.initial_flat_params(m.weights, m.biases).astype(np.float32),
               X_u=g["X_u_train"], u_data=g["u_train"], X_f=g["X_f_train"], lb=g["lb"], ub=g["ub"],
               nu=np.float64(g["nu"]), layers=np.asarray(g["layers"]))
    if name == "INF-ADMM":
        out["penalty_parameter"] = np.float64(g["penalty_parameter"])
        first = (g["number_of_ADMM_iterations"], g["number_of_w_optimization_steps"])
    else:
        first = (g["number_of_epochs"],)
    stages = [list(first)]
    _state(m, tf, out, "stage1")
    out["stage1_pred"] = _pred([g["u_pred"], g["f_pred"]])
    out["stage1_error"] = np.float64(g["error_u"])
    for k, args in enumerate(spec["more"], start=2):
        m.train(*args, g["filename"], g["GPU_number"])
        u_pred, f_pred = m.predict(g["X_star"])
        _state(m, tf, out, "stage%d" % k)
        out["stage%d_pred" % k] = _pred([u_pred, f_pred])
        out["stage%d_error" % k] = np.float64(np.linalg.norm(g["u_star"] - u_pred, 2) / np.linalg.norm(g["u_star"], 2))
        stages.append(list(args))
    _vectors(m, tf, out)
    out["meta"] = json.dumps(dict(script=rr.SCRIPTS[name], dialect="A", stages=stages, pred_stride=PRED_STRIDE))
    return out


def run_dialect_b(name, spec):
    import io
    import contextlib
    if "argv" in spec:
        g, tf = rr.run_script(name, argv=spec["argv"], compute="float64")
        m = g["A"]
    else:  # launch_NN_L2.py:4-12
        g, tf = rr.run_script(name, run_name="reference_module", compute="float64")
        p = g["Parameters"]()
        for k, v in spec["params"].items():
            setattr(p, k, v)
        cwd = os.getcwd()
        os.chdir(g["__scratch__"])
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                m = g["PhysicsInformedNN"](p)
        finally:
            os.chdir(cwd)
    p = m.params
    euler = hasattr(m, "rho_tf")
    out = dict(theta0=rr.initial_flat_params(m.weights, m.biases).astype(np.float32), lb=m.lb, ub=m.ub,
               layers=np.asarray(m.layers), X_u=np.hstack([m.x_data, m.t_data]),
               u_data=np.hstack([m.rho, m.u, m.E]) if euler else m.u)
    params = {k: getattr(p, k) for k in ("N_u", "N_data", "N_f", "rho", "pen", "epochs") if hasattr(p, k)}
    stages = [int(p.epochs)]

    def record(tag):
        _state(m, tf, out, tag)
        if euler:
            out[tag + "_pred"] = _pred([m.rho_pred_val, m.u_pred_val, m.E_pred_val, m.f1_pred_val, m.f2_pred_val, m.f3_pred_val])
            out[tag + "_error"] = np.array([m.error_rho, m.error_u, m.error_E])
        else:
            out[tag + "_pred"] = _pred([m.u_pred_val, m.f_pred_val])
            out[tag + "_error"] = np.float64(m.error_u)

    record("stage1")
    csv = os.path.join(g["__scratch__"], m.filename[:-3] + "csv")
    with open(csv) as fh:
        lines = fh.read().splitlines()
    out["csv_header"] = lines[0]
    out["csv_rows"] = np.int64(len(lines))
    out["csv_first_row"] = lines[1]
    cwd = os.getcwd()
    os.chdir(g["__scratch__"])
    try:
        for k, n in enumerate(spec["more"], start=2):
            with contextlib.redirect_stdout(io.StringIO()):
                m.params.epochs = n
                m.run_NN()            # train(n) + record_data + save_data + errors, as the constructor does
            record("stage%d" % k)
            stages.append(int(n))
    finally:
        os.chdir(cwd)
    if hasattr(m, "lambda_1"):
        out["lambda"] = np.array([m.sess.run(m.lambda_1)[0], m.sess.run(m.lambda_2)[0]], np.float64)
    _vectors(m, tf, out)
    if hasattr(m, "lbfgs") and name == "AB-ADMM":
        # the L-BFGS-B branch of AB-ADMM:213-216 (epoch > 50000) from the vec_* state, cut to LBFGS_ITERS iterations;
        # every other option is the one the reference constructor passed (AB-ADMM:66-72)
        m.lbfgs.options["maxiter"] = LBFGS_ITERS
        res = m.lbfgs.minimize(m.sess, feed_dict=_feed(m))
        out["lbfgs_options"] = json.dumps({k: float(v) for k, v in m.lbfgs.options.items()})
        out["lbfgs_theta"] = rr.flat_params(m.sess, m.weights, m.biases)
        out["lbfgs_loss"] = np.float64(res.fun)
        out["lbfgs_nit"] = np.int64(res.nit)
    out["meta"] = json.dumps(dict(script=rr.SCRIPTS[name], dialect="B", params=params, stages=stages,
                                  pred_stride=PRED_STRIDE))
    return out


def generate(name):
    spec = RUNS[name]
    out = run_dialect_a(name, spec) if spec["dialect"] == "A" else run_dialect_b(name, spec)
    big = out["theta0"].size > 100000
    if big:  # 162 003 / 282 201 parameters -> keep the file small: float32 storage, theta0 regenerated on load
        for k in list(out):
            if k.endswith("_theta") or k == "vec_grad":
                out[k] = np.asarray(out[k], np.float32)
        assert np.array_equal(rr.shim_initial_theta([int(n) for n in out["layers"]]), out.pop("theta0"))
    np.savez_compressed(os.path.join(HERE, "ref_%s.npz" % name), **out)
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or list(RUNS)
    for name in names:
        o = generate(name)
        print(name, "P =", o["theta0"].size, "loss =", float(o["vec_loss"]), "errors =",
              [np.round(o[k], 6).tolist() for k in sorted(o) if k.endswith("_error")], flush=True)
