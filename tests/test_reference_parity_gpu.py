"""GPU parity against the reference's OWN code: the CUDA path, through the C ABI and the drop-in classes, against
fixtures produced by executing the reference's eight unmodified model scripts (tests/golden/ref_<SCRIPT>.npz,
see tests/golden/make_ref_fixtures.py and oracle/run_reference.py).

Bound (north_star, fp32 path): loss, residuals and gradient within 1e-5 relative.  Where a looser bound is used the
reason is stated next to it."""
import numpy as np
import pytest

from oracle import tf_graph as tg
from tests.helpers import ENGINE_LOSS, GOLD, REF_RUNS, load_ref_fixture, max_rel_err, ref_problem, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
SCRIPTS = list(REF_RUNS)


def _last_stage(fx):
    return max(int(k[5:].split("_")[0]) for k in fx if k.startswith("stage") and k.endswith("_theta"))


@pytest.mark.parametrize("name", SCRIPTS)
def test_loss_gradient_residuals_against_the_reference_graph(name):
    from pinns_b200 import Engine
    fx = load_ref_fixture(name)
    p = ref_problem(name, fx)
    eng = Engine(p.layers, p.lb, p.ub, pde=p.pde, loss=ENGINE_LOSS[p.loss], lambda1=p.lam1, lambda2=p.lam2, rho=p.rho)
    eng.set_params(np.float32(fx["stage%d_theta" % _last_stage(fx)]))
    eng.set_data(fx["X_u"], fx["u_data"])
    eng.set_collocation(fx["vec_X_f"])
    if "vec_z" in fx:
        eng.admm_set_state(fx["vec_z"], fx["vec_gamma"])
    loss, grad = eng.loss_grad()
    P = eng.num_params
    assert abs(loss - fx["vec_loss"]) <= TOL * abs(fx["vec_loss"]), (loss, fx["vec_loss"])
    err = float(np.linalg.norm(grad[:P] - fx["vec_grad"])) / float(np.linalg.norm(fx["vec_grad"]))
    bound = TOL
    if "vec_z" in fx and name != "INF-ADMM":
        # The scripts evaluate this gradient right after the z/gamma update on the same batch (AB-ADMM:225-226,
        # Burgers_ADMM_batch.py:204-210), where the adjoint seed rho*(f - z) + gamma collapses to +-1/N_f: three to four
        # digits of f cancel and ANY float32 evaluation misses the float64 vector by 1e-5 ... 1e-3 of |g| -- the reference's
        # own graph evaluated in float32 included.  That evaluation is made here, on the same state, and the CUDA path is
        # held to three times its error (test_admm_cancellation_states_* below looks at a distribution of such states).
        import torch
        g32 = tg.evaluate(np.float32(fx["stage%d_theta" % _last_stage(fx)]), p, fx["X_u"], fx["u_data"], fx["vec_X_f"],
                          z=fx["vec_z"], gamma=fx["vec_gamma"], dtype=torch.float32).grad
        err32 = float(np.linalg.norm(g32 - fx["vec_grad"]) / np.linalg.norm(fx["vec_grad"]))
        bound = max(TOL, 3.0 * err32)
        print("%s: float32 evaluation of the reference graph: gradient error %.2e of |g|" % (name, err32))
    print("%s: CUDA path: gradient error %.2e of |g| (bound %.2e)" % (name, err, bound))
    assert err <= bound, (err, bound)
    y, f = eng.predict(fx["vec_X_f"])
    assert max_rel_err(f, fx["vec_f"]) <= TOL
    y_u, _ = eng.predict(fx["X_u"], want_f=False)
    assert max_rel_err(y_u, fx["vec_u_pred"]) <= TOL


@pytest.mark.parametrize("name", ["ID-ADMMb", "AB-ADMM"])
def test_admm_cancellation_states_error_distribution(name):
    """The cancellation regime as a distribution instead of one draw: eight fresh batches from the fixture's parameters,
    z/gamma updated from float64 residuals and rounded to float32 like the reference's variables, float64 gradient =
    truth.  The fused kernel's median error stays within 3x of the median error of a float32 evaluation of the reference
    graph on the same states (measured: 5.0e-5 against 2.6e-5 on ID-ADMMb, 1.2e-5 against 5.5e-6 on AB-ADMM; round 1's
    exponential-only tanh: 2.5e-4 / 4.0e-5)."""
    import torch
    from pinns_b200 import Engine
    fx = load_ref_fixture(name)
    p = ref_problem(name, fx)
    theta = np.float32(fx["stage%d_theta" % _last_stage(fx)])
    n_f = fx["vec_X_f"].shape[0]
    rng = np.random.default_rng(77)
    z, gamma = fx["vec_z"].astype(np.float64), fx["vec_gamma"].astype(np.float64)
    eng = Engine(p.layers, p.lb, p.ub, pde=p.pde, loss=ENGINE_LOSS[p.loss], lambda1=p.lam1, lambda2=p.lam2, rho=p.rho)
    eng.set_params(theta)
    eng.set_data(fx["X_u"], fx["u_data"])
    e_gpu, e_32 = [], []
    for _ in range(8):
        X_f = p.lb + (p.ub - p.lb) * rng.random((n_f, 2))
        f = tg.evaluate(theta, p, fx["X_u"], fx["u_data"], X_f, z=z, gamma=gamma, want_grad=False).f
        z2, g2 = tg.admm_update(f, z, gamma, p.rho, n_f)
        z2, g2 = z2.astype(np.float32), g2.astype(np.float32)
        ref = tg.evaluate(theta, p, fx["X_u"], fx["u_data"], X_f, z=z2, gamma=g2).grad
        g32 = tg.evaluate(theta, p, fx["X_u"], fx["u_data"], X_f, z=z2, gamma=g2, dtype=torch.float32).grad
        eng.set_collocation(X_f)
        eng.admm_set_state(z2, g2)
        _, g = eng.loss_grad()
        e_gpu.append(rel_err(g[:eng.num_params], ref))
        e_32.append(rel_err(g32, ref))
    print("%s: gradient error / |g| over 8 cancellation states: CUDA median %.2e (max %.2e), float32 reference graph median %.2e (max %.2e)"
          % (name, np.median(e_gpu), max(e_gpu), np.median(e_32), max(e_32)))
    assert np.median(e_gpu) <= 3.0 * np.median(e_32), (e_gpu, e_32)
    assert max(e_gpu) <= 2e-4


def _check_stage(fx, k, theta, z, gamma, pred, errors):
    tag = "stage%d" % k
    # k Adam steps of ~1e-3 each from identical parameters; where |g| is tiny m/sqrt(v) amplifies fp32 gradient noise
    assert np.abs(theta - fx[tag + "_theta"]).max() <= 2e-5, np.abs(theta - fx[tag + "_theta"]).max()
    assert rel_err(theta, fx[tag + "_theta"]) <= 1e-5
    if z is not None:
        scale = max(1.0, float(np.abs(fx[tag + "_z"]).max()))
        assert np.abs(z - fx[tag + "_z"]).max() <= 1e-4 * scale
        assert np.abs(gamma - fx[tag + "_gamma"]).max() <= 1e-4 * max(1.0, float(np.abs(fx[tag + "_gamma"]).max()))
    ref = fx[tag + "_pred"]
    n_out = pred.shape[1] // 2
    assert max_rel_err(pred[:, :n_out], ref[:, :n_out]) <= 1e-4          # outputs after the perturbed parameters
    assert max_rel_err(pred[:, n_out:], ref[:, n_out:]) <= 1e-3          # residuals: derivatives of the same
    assert np.allclose(np.atleast_1d(errors), np.atleast_1d(fx[tag + "_error"]), rtol=1e-4)


@pytest.mark.parametrize("name", ["INF-L2", "INF-ADMM"])
def test_dialect_a_classes_follow_the_reference_run(name):
    """Same constructor arguments, same initial parameters, same train() calls as the reference driver made."""
    from pinns_b200.models import PhysicsInformedNN, PhysicsInformedNN_ADMM
    fx = load_ref_fixture(name)
    sol = dict(np.load("%s/data/%s.npz" % (GOLD, REF_RUNS[name][0])))
    from oracle import data as odata
    grid = odata.burgers_grid(sol)
    layers = [int(n) for n in fx["layers"]]
    if name == "INF-L2":
        m = PhysicsInformedNN(fx["X_u"], fx["u_data"], fx["X_f"], layers, fx["lb"], fx["ub"], float(fx["nu"]), '0',
                              theta0=fx["theta0"], verbose=False)
    else:
        m = PhysicsInformedNN_ADMM(fx["X_u"], fx["u_data"], fx["X_f"], layers, fx["lb"], fx["ub"], float(fx["nu"]), 1,
                                   float(fx["penalty_parameter"]), 'fixture', '0', theta0=fx["theta0"], verbose=False)
    stride = fx["meta"]["pred_stride"]
    for k, args in enumerate(fx["meta"]["stages"], start=1):
        m.train(*args, 'fixture', '0')
        u_pred, f_pred = m.predict(grid["X_star"])
        assert u_pred.dtype == np.float32 and u_pred.shape == grid["u_star"].shape
        err = np.linalg.norm(grid["u_star"] - u_pred, 2) / np.linalg.norm(grid["u_star"], 2)
        z, gamma = m.engine.admm_state() if name == "INF-ADMM" else (None, None)
        _check_stage(fx, k, m.get_flat_params(), z, gamma, np.hstack([u_pred, f_pred])[::stride], err)


@pytest.mark.parametrize("name", ["ID-L2b", "ID-ADMMb", "AB-ADMM", "AB-L2", "AB-L1", "EUL"])
def test_dialect_b_classes_follow_the_reference_run(name):
    """Parameters object as the reference driver fills it; the constructor loads the data, draws the same training set
    and collocation batches from NumPy's legacy RNG, trains and records -- then run_NN() again for the later stages."""
    from pinns_b200.models import BurgersIdentification, EulerInference, EulerParameters, Parameters
    fx = load_ref_fixture(name)
    data = "%s/data/%s.npz" % (GOLD, REF_RUNS[name][0])
    meta = fx["meta"]
    p = EulerParameters() if name == "EUL" else Parameters()
    for key, val in meta["params"].items():
        setattr(p, key, val)
    p.epochs = meta["stages"][0]
    p.gpu = '0'
    if name == "EUL":
        m = EulerInference(p, data=data, theta0=fx["theta0"], verbose=False)
    else:
        m = BurgersIdentification(p, variant=name, data=data, theta0=fx["theta0"], verbose=False)
    assert np.array_equal(m.x_data, fx["X_u"][:, 0:1]) and np.array_equal(m.t_data, fx["X_u"][:, 1:2])
    stride = meta["pred_stride"]
    for k, n in enumerate(meta["stages"], start=1):
        if k > 1:
            m.params.epochs = n
            m.run_NN()
        if name == "EUL":
            pred = np.hstack([m.rho_pred_val, m.u_pred_val, m.E_pred_val, m.f1_pred_val, m.f2_pred_val, m.f3_pred_val])
            errors = np.array([m.error_rho, m.error_u, m.error_E])
        else:
            pred = np.hstack([m.u_pred_val, m.f_pred_val])
            errors = m.error_u
        z, gamma = m.engine.admm_state() if REF_RUNS[name][5] else (None, None)
        _check_stage(fx, k, m.get_flat_params(), z, gamma, pred[::stride], errors)
    assert np.array_equal(np.hstack([m.x_phys, m.t_phys]), fx["vec_X_f"])
    assert list(m.df.columns) == str(fx["csv_header"]).split(",")


def test_lbfgs_through_scipy_follows_the_reference_interface():
    """The L-BFGS-B branch (AB-ADMM:66-72,:213-216): the reference's own ScipyOptimizerInterface object ran 25 iterations
    from the fixture's state; `lbfgs_minimize` (host SciPy in float64, loss+grad from the GPU) does the same from the same
    state.  fp32 loss/gradient evaluations steer the line search slightly differently, hence 1e-3."""
    import json
    from pinns_b200.models import BurgersIdentification, Parameters
    fx = load_ref_fixture("AB-ADMM")
    p = Parameters()
    for key, val in fx["meta"]["params"].items():
        setattr(p, key, val)
    m = BurgersIdentification(p, variant="AB-ADMM", data="%s/data/%s.npz" % (GOLD, REF_RUNS["AB-ADMM"][0]), run=False,
                              theta0=np.float32(fx["stage%d_theta" % _last_stage(fx)]), verbose=False)
    m.engine.set_collocation(fx["vec_X_f"])
    m.engine.admm_set_state(fx["vec_z"], fx["vec_gamma"])
    opts = {k: (v if k == "ftol" else int(v)) for k, v in json.loads(str(fx["lbfgs_options"])).items()}
    res = m.lbfgs_minimize(opts)
    assert abs(res.fun - fx["lbfgs_loss"]) <= 1e-3 * fx["lbfgs_loss"] and res.fun < 0.98 * fx["vec_loss"], (res.fun, fx["lbfgs_loss"])
    assert np.abs(m.get_flat_params() - fx["lbfgs_theta"]).max() <= 5e-3
