"""The oracle against the reference's OWN code (CPU).

tests/golden/ref_<SCRIPT>.npz were produced by executing each of the reference's eight unmodified model scripts
(oracle/run_reference.py: TensorFlow-1 API shim over torch autograd, everything else real; generator
tests/golden/make_ref_fixtures.py).  Here the oracle's restatements -- tf_graph (loss, residuals, gradient, predict,
ADMM updates), taylor, optim.TF1Adam and data (RNG-ordered data preparation) -- must reproduce what the reference's
code computed: the vectors at the final state, and the whole run (training sets, per-stage parameters, ADMM state,
grid predictions) replayed from the same initial parameters.

Tolerances: the reference graph rounds python scalars such as 1/N_u to float32 constants, the oracle keeps them in
float64 (5e-8 relative in the loss); variables at rest are float32 in the oracle and float64 in the fixture's run
(1e-7 per step).  Residuals and network outputs are bit-identical.
"""
import os

import numpy as np
import pytest
import torch

from oracle import run_reference as rr
from oracle import taylor
from oracle import tf_graph as tg
from tests.helpers import GOLD, REF_RUNS, load_ref_fixture, max_rel_err, oracle_replay, ref_problem, rel_err

SCRIPTS = list(REF_RUNS)


def _last_stage(fx):
    return max(int(k[5:].split("_")[0]) for k in fx if k.startswith("stage") and k.endswith("_theta"))


@pytest.mark.parametrize("name", SCRIPTS)
def test_oracle_reproduces_reference_loss_gradient_residuals(name):
    fx = load_ref_fixture(name)
    prob = ref_problem(name, fx)
    theta = np.float32(fx["stage%d_theta" % _last_stage(fx)])
    args = (theta, prob, fx["X_u"], fx["u_data"], fx["vec_X_f"])
    kw = dict(z=fx.get("vec_z"), gamma=fx.get("vec_gamma"))
    ev = tg.evaluate(*args, **kw)
    assert abs(ev.loss - fx["vec_loss"]) <= 2e-7 * abs(fx["vec_loss"])
    # the Euler ADMM gradient is a sum over f - z + gamma/pen with z ~ f: cancellation amplifies the constant rounding
    assert rel_err(ev.grad, fx["vec_grad"]) <= (5e-6 if name == "EUL" else 2e-7)
    assert max_rel_err(ev.f, fx["vec_f"]) <= 1e-9
    assert max_rel_err(ev.u_pred, fx["vec_u_pred"]) <= 1e-12
    ty = taylor.evaluate(*args, **kw)                      # the hand-derived Taylor/reverse restatement as well
    assert abs(ty.loss - fx["vec_loss"]) <= 2e-7 * abs(fx["vec_loss"])
    assert rel_err(ty.grad, fx["vec_grad"]) <= (5e-6 if name == "EUL" else 2e-7)
    assert max_rel_err(ty.f, fx["vec_f"]) <= 1e-9
    if "vec_admm_misfit" in fx:                            # AB-ADMM:59 admm_misfit = mean |f - z|
        assert abs(np.abs(ev.f - fx["vec_z"]).mean() - fx["vec_admm_misfit"]) <= 1e-9


@pytest.mark.parametrize("name", SCRIPTS)
def test_oracle_replays_the_reference_run(name):
    """Same NumPy RNG stream -> identical training sets and collocation batches; TF-1 Adam, the train-loop bounds of
    each script, z/gamma updates (INF-ADMM's double dual update included) -> the same state after every stage."""
    fx = load_ref_fixture(name)
    rep = oracle_replay(name, fx)
    for k, st in rep.items():
        tag = "stage%d" % k
        assert np.array_equal(st["X_u"], fx["X_u"]) and np.array_equal(st["u_data"], fx["u_data"])
        assert np.abs(st["theta"] - fx[tag + "_theta"]).max() <= 1e-6        # Adam moves every parameter by ~1e-3 per step
        if st["z"] is not None:
            assert np.abs(st["z"] - fx[tag + "_z"]).max() <= 5e-6
            assert np.abs(st["gamma"] - fx[tag + "_gamma"]).max() <= 5e-6
        assert np.abs(st["pred"] - fx[tag + "_pred"]).max() <= 5e-6
    assert np.array_equal(st["X_f"], fx["vec_X_f"])                          # the last batch the reference drew
    if fx["meta"]["dialect"] == "A":
        assert np.array_equal(st["X_f"], fx["X_f"])


def test_reference_csv_schema():
    """record_data / save_data (AB-ADMM:400-409, EUL:428-437): header re-emitted on every append."""
    fx = load_ref_fixture("AB-ADMM")
    assert str(fx["csv_header"]) == "x,t,u_pred,epoch"
    sol = np.load(os.path.join(GOLD, "data", "TwoSin_burgers_shock.npz"))
    assert int(fx["csv_rows"]) == sol["usol"].size + 1                     # one header + one row per grid point
    assert str(load_ref_fixture("EUL")["csv_header"]) == "x,t,rho_pred,u_pred,E_pred,epoch"


@pytest.mark.skipif(not rr.available(), reason="/root/reference is only present in the build container")
def test_live_reference_run_matches_the_oracle():
    """Runs the unmodified AB-ADMM script now (other sizes than the committed fixture) -- in float32 like TensorFlow
    would -- and replays it with the fp64 oracle: parameters, ADMM state and loss agree within fp32 evaluation noise."""
    from tests.golden import make_ref_fixtures as mk
    argv = [40, 128, 5.0, 4, "0"]
    g, tf = rr.run_script("AB-ADMM", argv=argv, compute="float32")
    m = g["A"]
    fx = {"theta0": rr.initial_flat_params(m.weights, m.biases).astype(np.float32), "lb": m.lb, "ub": m.ub,
          "layers": np.asarray(m.layers), "lambda": np.array([1.0, 0.0]),
          "meta": dict(dialect="B", params=dict(N_u=40, N_f=128, rho=5.0), stages=[4], pred_stride=mk.PRED_STRIDE)}
    st = oracle_replay("AB-ADMM", fx)[1]
    out = {}
    mk._state(m, tf, out, "s")
    assert np.array_equal(st["X_u"], np.hstack([m.x_data, m.t_data]))
    assert np.array_equal(st["X_f"], np.hstack([m.x_phys, m.t_phys]))
    assert np.abs(st["theta"] - out["s_theta"]).max() <= 2e-5
    assert np.abs(st["z"] - out["s_z"]).max() <= 1e-4 * max(1.0, np.abs(st["z"]).max())
    pred = np.hstack([m.u_pred_val, m.f_pred_val])[::mk.PRED_STRIDE]
    assert np.abs(st["pred"] - pred).max() <= 1e-4 * max(1.0, np.abs(pred).max())


def test_shim_initialiser_is_reproducible_without_the_reference():
    fx = load_ref_fixture("AB-ADMM")
    assert np.array_equal(rr.shim_initial_theta([int(n) for n in fx["layers"]]), fx["theta0"])


def test_oracle_lbfgs_driver_follows_the_reference_scipy_interface():
    """AB-ADMM:66-72,:216: the reference's ScipyOptimizerInterface object (options as its constructor passed them, cut
    to 25 iterations) against oracle.optim.lbfgs_minimize from the same state: same packing order, float64 hand-off,
    jac=True.  The oracle rounds theta through float32 at every evaluation (a tf.Variable's storage), the fixture's run
    kept float64, so the two line searches drift apart slowly."""
    import json
    from oracle.optim import lbfgs_minimize
    fx = load_ref_fixture("AB-ADMM")
    prob = ref_problem("AB-ADMM", fx)
    opts = {k: (v if k == "ftol" else int(v)) for k, v in json.loads(str(fx["lbfgs_options"])).items()}
    assert {k: opts[k] for k in ("maxfun", "maxcor", "maxls", "ftol")} == dict(maxfun=50000, maxcor=50, maxls=50, ftol=1e-7)

    def loss_grad(x):
        ev = tg.evaluate(x, prob, fx["X_u"], fx["u_data"], fx["vec_X_f"], z=fx["vec_z"], gamma=fx["vec_gamma"])
        return ev.loss, ev.grad

    x, res = lbfgs_minimize(loss_grad, np.float32(fx["stage%d_theta" % _last_stage(fx)]).astype(np.float64), opts)
    assert res.nit == int(fx["lbfgs_nit"])
    assert abs(res.fun - fx["lbfgs_loss"]) <= 1e-4 * fx["lbfgs_loss"] and res.fun < 0.98 * fx["vec_loss"]
    assert np.abs(x - fx["lbfgs_theta"]).max() <= 5e-4


@pytest.mark.parametrize("name", SCRIPTS)
def test_float32_evaluation_noise_of_the_reference_graph_at_the_fixture_states(name):
    """What a float32 evaluation of the reference graph ITSELF (the oracle's op-for-op restatement run in torch float32,
    correctly rounded ops) makes of the fixture vectors.  Five scripts: loss, residuals and gradient to ~1e-7.  The three
    batch-ADMM scripts evaluate their gradient right after the z/gamma update on the same batch, where the adjoint seed
    rho (f - z) + gamma collapses to +-1/N_f and three to four digits of f cancel: there float32 itself is at 1e-5 .. 1e-3.
    This is the yardstick behind the scale-aware gradient bound of tests/test_reference_parity_gpu.py."""
    fx = load_ref_fixture(name)
    prob = ref_problem(name, fx)
    theta = np.float32(fx["stage%d_theta" % _last_stage(fx)])
    ev = tg.evaluate(theta, prob, fx["X_u"], fx["u_data"], fx["vec_X_f"], z=fx.get("vec_z"), gamma=fx.get("vec_gamma"),
                     dtype=torch.float32)
    noise = rel_err(ev.grad, fx["vec_grad"])
    assert abs(ev.loss - fx["vec_loss"]) <= 1e-6 * abs(fx["vec_loss"])
    assert max_rel_err(ev.f, fx["vec_f"]) <= 1e-5
    if name in ("ID-ADMMb", "AB-ADMM", "EUL"):
        zero = np.zeros_like(fx["vec_z"])
        bare = tg.evaluate(theta, prob, fx["X_u"], fx["u_data"], fx["vec_X_f"], z=zero, gamma=zero)
        cancel = np.linalg.norm(bare.grad) / np.linalg.norm(fx["vec_grad"])
        assert cancel > 500                                     # the gradient is what is left of a >500x larger one
        assert 1e-6 < noise < 5e-3, noise
        assert noise * np.linalg.norm(fx["vec_grad"]) <= 1e-6 * np.linalg.norm(bare.grad)
    else:
        assert noise <= 1e-6, noise
