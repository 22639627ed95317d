"""CPU tests of the oracle itself: two independent restatements + finite differences + committed fixtures."""
import glob
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import data as odata
from oracle import taylor as ty
from oracle import tf_graph as tg
from oracle.optim import TF1Adam, lbfgs_minimize
from oracle.philox import lhs_perm, philox4x32_10, sample_collocation, sample_lhs
from tests.helpers import make_case

GOLD = os.path.join(os.path.dirname(__file__), "golden")
B20 = [2] + [20] * 8 + [1]

CASES = [
    (tg.PDE_BURGERS, B20, tg.LOSS_V1), (tg.PDE_BURGERS, B20, tg.LOSS_V2), (tg.PDE_BURGERS, B20, tg.LOSS_V3),
    (tg.PDE_BURGERS, B20, tg.LOSS_V4), (tg.PDE_BURGERS, B20, tg.LOSS_V5),
    (tg.PDE_BURGERS, [2, 7, 13, 1], tg.LOSS_V4),
    (tg.PDE_EULER, [2, 24, 24, 24, 3], tg.LOSS_V6), (tg.PDE_EULER, [2, 24, 24, 24, 3], tg.LOSS_EULER_MSE),
]


@pytest.mark.parametrize("pde,layers,loss", CASES)
def test_autograd_vs_taylor_restatement(pde, layers, loss):
    """reverse-over-reverse (the reference graph, INF-L2:113-120 / EUL:176-198) == Taylor-forward + one reverse sweep"""
    c = make_case(pde, layers, loss, 23, 77, seed=5)
    a = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"], c["z"], c["gamma"])
    b = ty.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"], c["z"], c["gamma"])
    assert abs(a.loss - b.loss) <= 1e-12 * abs(a.loss)
    assert np.abs(a.grad - b.grad).max() <= 1e-11 * np.abs(a.grad).max()
    assert np.abs(a.f - b.f).max() <= 1e-12 * max(1.0, np.abs(a.f).max())
    assert np.allclose(a.dlam, b.dlam, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("pde,layers,loss", [(tg.PDE_BURGERS, [2, 6, 6, 1], tg.LOSS_V4), (tg.PDE_BURGERS, [2, 6, 6, 1], tg.LOSS_V5),
                                             (tg.PDE_EULER, [2, 6, 6, 3], tg.LOSS_V6)])
def test_gradient_vs_central_differences(pde, layers, loss):
    c = make_case(pde, layers, loss, 9, 21, seed=11)
    ev = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"], c["z"], c["gamma"])
    rng = np.random.default_rng(0)
    th = c["theta"].astype(np.float64)

    def loss_at(t):
        # bypass the float32 round trip of theta: evaluate with exact float64 perturbations through the Taylor restatement
        t32 = t.copy()
        return _loss_fp64(t32, c)

    for k in rng.choice(th.size, 12, replace=False):
        h = 1e-5
        tp, tm = th.copy(), th.copy()
        tp[k] += h
        tm[k] -= h
        fd = (loss_at(tp) - loss_at(tm)) / (2 * h)
        assert abs(fd - ev.grad[k]) <= 2e-5 * max(1.0, abs(ev.grad[k])), (k, fd, ev.grad[k])


def _loss_fp64(theta, c):
    """loss with theta NOT rounded to float32 (finite differences need exact perturbations)"""
    prob = c["prob"]
    th = torch.from_numpy(theta)
    weights, biases = tg.unpack(th, prob.layers)
    x_u = tg.feed(c["X_u"][:, 0:1]); t_u = tg.feed(c["X_u"][:, 1:2]); u_d = tg.feed(c["u"])
    x_f = tg.feed(c["X_f"][:, 0:1]).requires_grad_(True); t_f = tg.feed(c["X_f"][:, 1:2]).requires_grad_(True)
    n_u, n_f = c["X_u"].shape[0], c["X_f"].shape[0]
    u_pred = tg.net_u(x_u, t_u, weights, biases, prob.lb, prob.ub)
    rho = float(np.float32(prob.rho))
    if prob.pde == tg.PDE_BURGERS:
        fs = [tg.net_f_burgers(x_f, t_f, weights, biases, prob.lb, prob.ub, float(np.float32(prob.lam1)), float(np.float32(prob.lam2)))]
    else:
        fs = list(tg.net_f_euler(x_f, t_f, weights, biases, prob.lb, prob.ub))
    z = [torch.from_numpy(c["z"][:, k:k + 1].astype(np.float64)) for k in range(len(fs))]
    g = [torch.from_numpy(c["gamma"][:, k:k + 1].astype(np.float64)) for k in range(len(fs))]
    r = u_d - u_pred
    if prob.loss == tg.LOSS_V4:
        L = (r * r).sum() / n_u + (fs[0] ** 2).sum() / n_f
    elif prob.loss == tg.LOSS_V5:
        L = (r * r).sum() / n_u + rho / 2 * ((fs[0] - z[0] + g[0] / rho) ** 2).sum()
    else:
        L = (r * r).sum() / n_u + sum(rho / 2 * ((fs[k] - z[k] + g[k] / rho) ** 2).sum() for k in range(3))
    return float(L.detach())


def test_reverse_step_in_output_streams_equals_the_z_stream_form():
    """the kernels' reverse step works on the stashed H streams (oracle.taylor.reverse_step_hstream): same numbers as
    the Z-stream formulas of appendix A.2, also where tanh saturates"""
    from oracle.taylor import reverse_step_hstream
    rng = np.random.default_rng(5)
    z, zx, zt, zxx, hb, hxb, htb, hxxb = rng.standard_normal((8, 20000)) * np.array([[3.0]] + [[1.5]] * 7)
    a = np.tanh(z)
    d1 = 1 - a * a; d2 = -2 * a * d1; d3 = -2 * d1 * (1 - 3 * a * a)
    ref = (d1 * hb + d2 * (zx * hxb + zt * htb + zxx * hxxb) + d3 * zx * zx * hxxb, d1 * hxb + 2 * d2 * zx * hxxb, d1 * htb, d1 * hxxb)
    got = reverse_step_hstream(a, d1 * zx, d1 * zt, d1 * zxx + d2 * zx * zx, hb, hxb, htb, hxxb)
    for r, g in zip(ref, got):
        assert np.abs(r - g).max() <= 1e-13 * max(np.abs(r).max(), 1.0)


def test_fp32_graph_is_within_parity_budget_of_fp64():
    """the noise floor of an fp32 evaluation of the reference graph sits well under the 1e-5 parity target"""
    c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V4, 100, 2000, seed=3)
    a = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"])
    b = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"], dtype=torch.float32)
    assert abs(a.loss - b.loss) <= 2e-6 * abs(a.loss)
    assert np.linalg.norm(a.grad - b.grad) <= 1e-5 * np.linalg.norm(a.grad)


def test_committed_golden_vectors_match_oracle():
    files = sorted(glob.glob(os.path.join(GOLD, "vectors_*.npz")))
    assert len(files) >= 9
    from tests.golden.make_fixtures import VECTOR_CASES, case_seed
    cases = {c[0]: c for c in VECTOR_CASES}
    for f in files:
        name = os.path.basename(f)[len("vectors_"):-4]
        if "200" in name or "128" in name:
            continue  # wide nets: checked on the GPU box and in the slow marker below
        g = np.load(f)
        _, pde, layers, loss, n_u, n_f = cases[name]
        c = make_case(pde, layers, loss, n_u, n_f, seed=case_seed(name))
        ev = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"], c["z"], c["gamma"])
        assert abs(ev.loss - float(g["loss"])) <= 1e-12 * abs(ev.loss)
        assert np.abs(ev.grad[g["grad_idx"]] - g["grad"]).max() <= 1e-11 * np.abs(g["grad"]).max()
        assert np.abs(ev.f - g["f"]).max() <= 1e-12 * max(1.0, np.abs(g["f"]).max())


def test_tf1_adam_formula_and_trajectory_fixture():
    """TF-1 ApplyAdam: epsilon outside the bias correction"""
    rng = np.random.default_rng(0)
    th = rng.standard_normal(7)
    opt = TF1Adam(7)
    m = np.zeros(7); v = np.zeros(7)
    for t in range(1, 4):
        g = rng.standard_normal(7)
        new = opt.step(th, g)
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        lr_t = 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        assert np.allclose(new, th - lr_t * m / (np.sqrt(v) + 1e-8), rtol=1e-13, atol=0)
        th = new
    g = np.load(os.path.join(GOLD, "vectors_burgers20_v4.npz"))
    assert g["adam_losses"].shape == (5,) and g["adam_losses"][4] < g["adam_losses"][0]


def test_lbfgs_driver_minimises_a_small_pinn():
    c = make_case(tg.PDE_BURGERS, [2, 8, 8, 1], tg.LOSS_V4, 20, 100, seed=2)

    def fun(x):
        ev = tg.evaluate(x, c["prob"], c["X_u"], c["u"], c["X_f"])
        return ev.loss, ev.grad

    l0, _ = fun(c["theta"])
    x, res = lbfgs_minimize(fun, c["theta"], {'maxiter': 30, 'maxfun': 100, 'maxcor': 50, 'maxls': 50, 'ftol': 1e-12})
    assert res.fun < 0.9 * l0


def test_soft_threshold_and_admm_update():
    f = torch.tensor([[0.5], [-0.5], [0.001], [0.0]], dtype=torch.float64)
    gam = torch.zeros_like(f)
    z = tg.soft_threshold(f, gam, rho=10.0, n_f=4).numpy().ravel()
    kappa = 1 / 40
    assert np.allclose(z, [0.5 - kappa, -0.5 + kappa, 0.0, 0.0])
    zn, gn = tg.admm_update(f.numpy(), np.ones((4, 1)), np.ones((4, 1)), 10.0, 4)
    assert np.allclose(gn, 1.0 + 10.0 * (f.numpy() - zn))
    zq, gq = tg.admm_update(f.numpy(), np.ones((4, 1)), np.ones((4, 1)), 10.0, 4, inf_admm_quirk=True)
    g1 = 1.0 + 10.0 * (f.numpy() - 1.0)
    assert np.allclose(gq, g1 + 10.0 * (f.numpy() - zq))  # two dual updates per step (INF-ADMM:106-107)


def test_data_preparation_on_the_committed_fixtures():
    sol = dict(np.load(os.path.join(GOLD, "data", "burgers_shock.npz")))
    g = odata.burgers_inference_inputs(sol, N_u=100, N_f=10000)
    assert g["X_star"].shape == (25600, 2) and g["X_u_all"].shape == (456, 2)
    assert g["X_f"].shape == (10456, 2) and g["X_u"].shape == (100, 2)          # 10 000 LHS + 456 IC/BC
    assert np.allclose(g["lb"], [-1.0, 0.0]) and np.allclose(g["ub"], [1.0, 0.99])
    lhs = (g["X_f"][:10000] - g["lb"]) / (g["ub"] - g["lb"])
    for d in range(2):  # one sample per stratum in every dimension
        assert np.array_equal(np.sort(np.floor(lhs[:, d] * 10000).astype(int)), np.arange(10000))
    e = odata.euler_inputs(dict(np.load(os.path.join(GOLD, "data", "Abgrall_eulers.npz"))))
    assert e["X_u"].shape == (200, 2) and e["u"].shape == (200, 3) and e["X_star"].shape == (47100, 2)


def test_xavier_init_distribution():
    th = tg.xavier_init(B20, np.random.default_rng(1))
    assert th.size == tg.num_params(B20) == 3021
    W2 = th[60 + 20:60 + 20 + 400]
    std = np.sqrt(2 / 40)
    assert np.abs(W2).max() <= 2 * std + 1e-7 and abs(W2.std() - 0.88 * std) < 0.1 * std  # truncated normal
    assert np.all(th[40:60] == 0)


def test_philox_known_answer_and_sharding_invariance():
    # Random123 known-answer test: philox4x32-10, counter = key = 0
    o = philox4x32_10(np.array([0], np.uint64), 0)
    assert [int(v[0]) for v in o] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    lb, ub = [-1.0, 0.0], [1.0, 0.99]
    full = sample_collocation(1234, 0, 1000, lb, ub)
    parts = np.vstack([sample_collocation(1234, 0, 300, lb, ub), sample_collocation(1234, 300, 700, lb, ub)])
    assert np.array_equal(full, parts)
    assert full[:, 0].min() >= -1 and full[:, 0].max() < 1 and full[:, 1].min() >= 0 and full[:, 1].max() < 0.99
    assert abs(full[:, 0].mean()) < 0.1


@pytest.mark.parametrize("n", [1, 2, 5, 100, 1000, 10000])
def test_device_lhs_restatement_is_a_latin_hypercube(n):
    """pinn_sample_lhs's restatement (oracle.philox.sample_lhs) has the defining property of pyDOE.lhs (INF-L2:183,
    oracle.data.lhs): exactly one point in each of the n strata of either axis; any slice of the design can be produced
    on its own; the two axes use independent permutations."""
    lb, ub = np.array([-1.0, 0.0]), np.array([1.0, 0.99])
    for d in (0, 1):
        assert sorted(lhs_perm(np.arange(n), n, d, 1234).tolist()) == list(range(n))     # a bijection of [0, n)
    X = sample_lhs(1234, 0, n, n, lb, ub)
    assert X.dtype == np.float32 and X.shape == (n, 2)
    strata = np.floor((X.astype(np.float64) - lb) / (ub - lb) * n + 1e-6 * (n <= 1000)).astype(np.int64)
    strata = np.minimum(strata, n - 1)            # float32 rounding of the last stratum's upper end
    if n <= 1000:                                 # (beyond that the float32 cast may move a point across a stratum edge)
        for d in (0, 1):
            assert np.array_equal(np.sort(strata[:, d]), np.arange(n))
    else:
        for d in (0, 1):
            assert np.abs(np.sort(strata[:, d]) - np.arange(n)).max() <= 1
    if n >= 100:
        a, b = n // 3, n - n // 3
        assert np.array_equal(np.vstack([sample_lhs(1234, 0, a, n, lb, ub), sample_lhs(1234, a, b, n, lb, ub)]), X)
        p0, p1 = lhs_perm(np.arange(n), n, 0, 1234).astype(float), lhs_perm(np.arange(n), n, 1, 1234).astype(float)
        assert abs(np.corrcoef(p0, p1)[0, 1]) < 4.0 / np.sqrt(n)                         # axes permuted independently
        assert abs(np.corrcoef(np.arange(n), p0)[0, 1]) < 4.0 / np.sqrt(n)               # ... and not in index order
        assert not np.array_equal(sample_lhs(1235, 0, n, n, lb, ub), X)
        # same marginals as the reference's host construction: both are stratified uniforms
        from oracle.data import lhs
        H = lb + (ub - lb) * lhs(2, n, rng=np.random.RandomState(0))
        assert np.abs(np.sort(H[:, 0]) - np.sort(X[:, 0].astype(np.float64))).max() <= 2.0 / n + 1e-6


def test_autograd_vs_taylor_on_random_ragged_nets():
    """Property test (hypothesis): any depth / ragged widths / batch sizes, every PDE x loss variant -- the two restatements
    agree to rounding, so a bug would have to be made twice, independently, to slip through."""
    from hypothesis import given, settings, strategies as st

    burgers_losses = [tg.LOSS_V1, tg.LOSS_V2, tg.LOSS_V3, tg.LOSS_V4, tg.LOSS_V5]
    euler_losses = [tg.LOSS_V6, tg.LOSS_EULER_MSE]

    @settings(max_examples=30, deadline=None, derandomize=True)
    @given(st.lists(st.integers(1, 19), min_size=1, max_size=5), st.booleans(), st.integers(0, 4), st.integers(1, 40),
           st.integers(1, 70), st.integers(0, 10 ** 6))
    def check(hidden, euler, which, n_u, n_f, seed):
        pde = tg.PDE_EULER if euler else tg.PDE_BURGERS
        loss = euler_losses[which % 2] if euler else burgers_losses[which]
        layers = [2] + hidden + [3 if euler else 1]
        c = make_case(pde, layers, loss, n_u, n_f, seed=seed)
        a = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"], c["z"], c["gamma"])
        b = ty.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"], c["z"], c["gamma"])
        assert abs(a.loss - b.loss) <= 1e-11 * abs(a.loss)
        assert np.abs(a.grad - b.grad).max() <= 1e-10 * max(np.abs(a.grad).max(), 1e-30)
        assert np.abs(a.f - b.f).max() <= 1e-11 * max(1.0, np.abs(a.f).max())

    check()
