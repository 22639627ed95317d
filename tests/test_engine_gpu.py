"""GPU tests of the engine through the C ABI beyond plain loss/grad parity: committed golden vectors, Adam
trajectories, ADMM updates, device sampler, determinism, ragged sizes, size-independent properties at large N."""
import glob
import os

import numpy as np
import pytest

from oracle import tf_graph as tg
from oracle.optim import TF1Adam
from oracle.philox import sample_collocation, sample_lhs
from tests.helpers import make_case, make_engine, rel_err, max_rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
B20 = [2] + [20] * 8 + [1]
TOL = 1e-5


def _golden_cases():
    from tests.golden.make_fixtures import VECTOR_CASES
    return VECTOR_CASES


@pytest.mark.parametrize("case", _golden_cases(), ids=[c[0] for c in _golden_cases()])
def test_committed_golden_vectors(case):
    from tests.golden.make_fixtures import case_seed
    name, pde, layers, loss, n_u, n_f = case
    g = np.load(os.path.join(GOLD, "vectors_%s.npz" % name))
    c = make_case(pde, layers, loss, n_u, n_f, seed=case_seed(name))
    eng = make_engine(c, trainable_lambda=(pde == tg.PDE_BURGERS))
    loss_gpu, grad = eng.loss_grad()
    P = eng.num_params
    assert abs(loss_gpu - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    sub = grad[:P][g["grad_idx"]]
    assert np.linalg.norm(sub - g["grad"]) <= TOL * np.linalg.norm(g["grad"])
    assert abs(np.linalg.norm(grad[:P]) - float(g["grad_norm"])) <= TOL * float(g["grad_norm"])
    if pde == tg.PDE_BURGERS:
        assert np.allclose(grad[P:], g["dlam"], rtol=3e-5, atol=1e-6 * max(1.0, np.abs(g["dlam"]).max()))
    _, f = eng.predict(c["X_f"])
    assert max_rel_err(f, g["f"]) <= TOL
    # five TF-1 Adam steps on the device vs the fp64 trajectory
    eng.adam_reset()
    losses = []
    for _ in range(5):
        losses.append(eng.loss_grad(want_grad=False)[0])
        eng.adam_apply()
    # the L1^2 loss (V3) has sign(f) kinks: Adam normalises the step, so points whose residual changes sign between the
    # fp32 and fp64 trajectories move the iterate at first order; every other variant stays at rounding level
    # (Adam also amplifies rounding noise on components whose gradient is itself at noise level: |step| ~ lr whatever |g|)
    rt, tt = (2e-3, 2e-3) if loss == tg.LOSS_V3 else (5e-5, 1e-4)
    assert np.allclose(losses, g["adam_losses"], rtol=rt)
    th5 = eng.get_params()
    assert np.linalg.norm(th5[g["grad_idx"]] - g["adam_theta5"]) <= tt * np.linalg.norm(g["adam_theta5"])


@pytest.mark.parametrize("path", ["generic", "auto"])
@pytest.mark.parametrize("quirk", [False, True])
def test_admm_init_and_update_burgers(path, quirk):
    c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V5, 50, 777, seed=21, rho=10.0)
    eng = make_engine(c, path=path)
    _, f_ref = tg.predict(c["theta"], c["prob"], c["X_f"])
    eng.admm_init()                                   # z = gamma = 1 then z <- f(theta0)  (AB-ADMM:96-97,:121-122)
    z, gam = eng.admm_state()
    assert max_rel_err(z, f_ref) <= TOL and np.all(gam == 1.0)
    eng.admm_set_state(c["z"], c["gamma"])
    eng.admm_update(inf_admm_quirk=quirk)
    z_ref, g_ref = tg.admm_update(f_ref, c["z"], c["gamma"], 10.0, 777, inf_admm_quirk=quirk)
    z, gam = eng.admm_state()
    kappa = 1.0 / (10.0 * 777)
    val = f_ref + (c["gamma"] + (10.0 * (f_ref - c["z"]) if quirk else 0.0)) / 10.0
    safe = (np.abs(np.abs(val) - kappa) > 1e-5).ravel()   # away from the threshold kink fp32 and fp64 agree
    assert safe.mean() > 0.95
    assert np.abs(z - z_ref)[safe].max() <= 2e-5 * max(1.0, np.abs(z_ref).max())
    assert np.abs(gam - g_ref)[safe].max() <= 1e-4 * max(1.0, np.abs(g_ref).max())


def test_admm_update_euler_three_residual_blocks():
    c = make_case(tg.PDE_EULER, [2, 40, 40, 40, 3], tg.LOSS_V6, 30, 301, seed=4, rho=40.0)
    eng = make_engine(c)
    _, f_ref = tg.predict(c["theta"], c["prob"], c["X_f"])
    eng.admm_update()
    z_ref, g_ref = tg.admm_update(f_ref, c["z"], c["gamma"], 40.0, 301)
    z, gam = eng.admm_state()
    kappa = 1.0 / (40.0 * 301)
    safe = np.abs(np.abs(f_ref + c["gamma"] / 40.0) - kappa) > 1e-5
    assert np.abs(z - z_ref)[safe].max() <= 2e-5 * max(1.0, np.abs(z_ref).max())
    assert np.abs(gam - g_ref)[safe].max() <= 2e-4 * max(1.0, np.abs(g_ref).max())


def test_device_sampler_is_bit_exact_and_shard_invariant():
    c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V4, 10, 64, seed=1)
    eng = make_engine(c)
    eng.sample_collocation(1234, 0, 5000)
    full = eng.get_collocation()
    ref = sample_collocation(1234, 0, 5000, c["prob"].lb, c["prob"].ub)
    assert np.array_equal(full, ref)                           # integer pipeline + one fma: bit exact
    eng.sample_collocation(1234, 3000, 2000)
    assert np.array_equal(eng.get_collocation(), full[3000:])  # point i does not depend on the sharding
    big = (1 << 40) + 7
    eng.sample_collocation(99, big, 33)
    assert np.array_equal(eng.get_collocation(), sample_collocation(99, big, 33, c["prob"].lb, c["prob"].ub))


def test_device_lhs_is_bit_exact_shard_invariant_and_latin():
    """pinn_sample_lhs (INF-L2:183 `lb + (ub - lb) * lhs(2, N_f)` on the device) against its numpy restatement, bit for bit
    (integer permutation + Philox + float64 arithmetic with explicit roundings); slices of a design; and the Latin
    property at the size of BASELINE config 4's largest batch, counted on the device."""
    import torch
    c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V4, 10, 64, seed=1)
    eng = make_engine(c)
    lb, ub = c["prob"].lb, c["prob"].ub
    for n in (1, 2, 7, 1000, 10000, 65537):
        eng.sample_lhs(1234, n)
        assert np.array_equal(eng.get_collocation(), sample_lhs(1234, 0, n, n, lb, ub)), n
    eng.sample_lhs(1234, 10000)
    full = eng.get_collocation()
    eng.sample_lhs(1234, 2500, first_index=7000, n_design=10000)           # a rank's slice of the same design
    assert np.array_equal(eng.get_collocation(), full[7000:9500])
    eng.sample_lhs(77, 3000, first_index=(1 << 33) + 5, n_design=(1 << 33) + 4000)  # 64-bit indices, 34-bit Feistel domain
    assert np.array_equal(eng.get_collocation(), sample_lhs(77, (1 << 33) + 5, 3000, (1 << 33) + 4000, lb, ub))
    with pytest.raises(RuntimeError):
        eng.sample_lhs(1, 100, first_index=950, n_design=1000)             # slice beyond the design
    lo = torch.tensor(np.asarray(lb, np.float64), device="cuda")
    w = torch.tensor(np.asarray(ub, np.float64) - np.asarray(lb, np.float64), device="cuda")
    for n in (1 << 18, 1 << 24):
        eng.sample_lhs(5, n)
        X = torch.empty((n, 2), dtype=torch.float32, device="cuda")
        eng.get_collocation_device(X)
        Xd = X.double()
        for d in (0, 1):
            # sorted along an axis, point i sits in stratum i: within one stratum width of its centre (+ the float32 cast)
            centre = lo[d] + w[d] * (torch.arange(n, device="cuda").double() + 0.5) / n
            assert float((Xd[:, d].sort().values - centre).abs().max()) <= 0.5 * float(w[d]) / n + 1.3e-7
            if n == 1 << 18:   # strata 64 float32 ulps wide: the cast moves few points across an edge
                cnt = torch.bincount(torch.clamp(torch.floor((Xd[:, d] - lo[d]) / w[d] * n).long(), 0, n - 1), minlength=n)
                assert int(cnt.max()) <= 3 and int((cnt == 0).sum()) <= n // 20, (d, int(cnt.max()), int((cnt == 0).sum()))
        del X, Xd


@pytest.mark.parametrize("path", ["generic", "auto"])
def test_run_to_run_determinism(path):
    c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V4, 100, 50000, seed=6)
    eng = make_engine(c, path=path)
    l1, g1 = eng.loss_grad()
    l2, g2 = eng.loss_grad()
    assert l1 == l2 and np.array_equal(g1, g2)                # fixed-order reductions, no atomics


@pytest.mark.parametrize("n_f", [1000, 148 * 8 * 32 * 16 + 77])
def test_low_traffic_mode_is_bit_identical(n_f, monkeypatch):
    """PINN_FUSED_TMEM=1 keeps the warp-private W-bar tiles in tensor memory (tcgen05.ld/st, one dump per launch) and
    PINN_FUSED_DISCARD=1 drops dead stash lines from the L2: same additions in the same order -> the same bits, for a
    single-batch-per-warp job, a many-round job, the chunk-accumulating host feed and an Adam trajectory."""
    import torch
    c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V5, 100, n_f, seed=11)
    base = make_engine(c, path="fused")
    l0, g0 = base.loss_grad()
    monkeypatch.setenv("PINN_FUSED_TMEM", "1")
    monkeypatch.setenv("PINN_FUSED_DISCARD", "1")
    low = make_engine(c, path="fused")
    l1, g1 = low.loss_grad()
    assert l0 == l1 and np.array_equal(g0, g1)
    l2, g2 = low.loss_grad()                                  # tensor memory is re-initialised by every launch
    assert l2 == l1 and np.array_equal(g2, g1)
    for eng in (base, low):
        eng.adam_config(lr=1e-3)
        eng.adam_steps(5)
    assert np.array_equal(base.get_params(), low.get_params())
    if n_f > 600000:                                          # the host feed accumulates over several launches
        host = torch.from_numpy(c["X_f"].astype(np.float32)).pin_memory()
        for eng in (base, low):
            eng.set_params(c["theta"])
            eng.feed_collocation(host)
        lb, gb = base.loss_grad()
        ll, gl = low.loss_grad()
        assert lb == ll and np.array_equal(gb, gl)


@pytest.mark.parametrize("n_f", [1, 31, 32, 33, 255, 257, 1000])
@pytest.mark.parametrize("path", ["generic", "auto"])
def test_ragged_collocation_sizes(n_f, path):
    c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V4, 7, n_f, seed=n_f)
    ref = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"])
    eng = make_engine(c, path=path)
    loss, grad = eng.loss_grad()
    assert abs(loss - ref.loss) <= TOL * abs(ref.loss)
    assert rel_err(grad, ref.grad) <= TOL


def test_fused_and_generic_kernels_agree_at_scale():
    c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V4, 100, 64, seed=8)
    n = 1 << 18
    outs = []
    for path in ("fused", "generic"):
        eng = make_engine(c, path=path)
        eng.sample_collocation(7, 0, n)
        assert eng.kernel_path == path
        outs.append(eng.loss_grad())
    assert abs(outs[0][0] - outs[1][0]) <= TOL * abs(outs[1][0])
    assert rel_err(outs[0][1], outs[1][1]) <= TOL


@pytest.mark.parametrize("n", [1 << 22, 1 << 26], ids=["4Mi", "64Mi-BASELINE-config-4"])
def test_linearity_over_shards_at_full_size(n):
    """size-independent property at bench scale: grad(all points) = grad(shard A) + grad(shard B) with job-wide 1/N_f"""
    c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V4, 100, 64, seed=12)
    eng = make_engine(c)
    eng.set_data_weight(0.0)
    eng.sample_collocation(1234, 0, n, n)
    l_all, g_all = eng.loss_grad()
    na = n // 3 + 5
    eng.sample_collocation(1234, 0, na, n)
    l_a, g_a = eng.loss_grad()
    eng.sample_collocation(1234, na, n - na, n)
    l_b, g_b = eng.loss_grad()
    assert abs((l_a + l_b) - l_all) <= 2e-6 * abs(l_all)
    assert rel_err(g_a + g_b, g_all) <= 5e-6
    # and against the oracle on a sample the CPU finishes in seconds: same points, loss scaled by the sample size
    ns = 20000
    Xs = sample_collocation(1234, 0, ns, c["prob"].lb, c["prob"].ub).astype(np.float64)
    eng.set_data_weight(1.0)
    eng.sample_collocation(1234, 0, ns, ns)
    l_s, g_s = eng.loss_grad()
    ref = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], Xs)
    assert abs(l_s - ref.loss) <= TOL * abs(ref.loss) and rel_err(g_s, ref.grad) <= TOL


@pytest.mark.parametrize("pde,layers,loss,n", [(tg.PDE_BURGERS, [2] + [128] * 8 + [1], tg.LOSS_V4, 1 << 24),
                                               (tg.PDE_EULER, [2] + [200] * 5 + [3], tg.LOSS_EULER_MSE, 1 << 20)],
                         ids=["config5-128x8-16Mi", "euler-200x5-1Mi"])
def test_tensor_path_linearity_over_shards_at_config5_size(pde, layers, loss, n):
    """BASELINE config 5 at its full size ([2,128x8,1], 16 Mi points, tcgen05 path) and the reference's Euler net at 1 Mi points: the
    same shard-additivity, within the tensor path's stated 5e-5; plus the oracle on a sample of the same Philox stream."""
    c = make_case(pde, layers, loss, 100, 64, seed=13)
    eng = make_engine(c, path="tensor")
    eng.set_data_weight(0.0)
    eng.sample_collocation(1234, 0, n, n)
    assert eng.kernel_path == "tensor"
    l_all, g_all = eng.loss_grad()
    na = n // 3 + 5
    eng.sample_collocation(1234, 0, na, n)
    l_a, g_a = eng.loss_grad()
    eng.sample_collocation(1234, na, n - na, n)
    l_b, g_b = eng.loss_grad()
    assert abs((l_a + l_b) - l_all) <= 5e-6 * abs(l_all)
    assert rel_err(g_a + g_b, g_all) <= 2e-5
    ns = 8192
    Xs = sample_collocation(1234, 0, ns, c["prob"].lb, c["prob"].ub).astype(np.float64)
    eng.set_data_weight(1.0)
    eng.sample_collocation(1234, 0, ns, ns)
    l_s, g_s = eng.loss_grad()
    ref = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], Xs)
    assert abs(l_s - ref.loss) <= 5e-5 * abs(ref.loss) and rel_err(g_s, ref.grad) <= 5e-5


def test_trainable_lambda_adam_moves_lambda_like_the_oracle():
    c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V4, 100, 2000, seed=14, lam1=0.5, lam2=0.02)
    eng = make_engine(c, trainable_lambda=True)
    theta = c["theta"].astype(np.float64)
    lam = np.array([np.float32(0.5), np.float32(0.02)], np.float64)
    opt = TF1Adam(theta.size + 2)
    for _ in range(3):
        prob = tg.Problem(B20, c["prob"].lb, c["prob"].ub, loss=tg.LOSS_V4, lam1=lam[0], lam2=lam[1])
        ev = tg.evaluate(theta, prob, c["X_u"], c["u"], c["X_f"])
        new = opt.step(np.concatenate([theta, lam]), np.concatenate([ev.grad, ev.dlam]))
        theta, lam = new[:-2], new[-2:]
    eng.adam_steps(3)
    l1, l2 = eng.get_lambda()
    assert abs(l1 - lam[0]) <= 1e-5 and abs(l2 - lam[1]) <= 1e-5
    assert rel_err(eng.get_params(), theta) <= 1e-5


def test_v1_unsquared_data_norm_gradient_is_nan_at_zero_misfit_like_tf():
    """tf.norm's gradient is NaN at exactly zero misfit: preserved, not hidden"""
    c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V1, 16, 64, seed=3)
    eng = make_engine(c, path="generic")  # predict and the data term must come from the same kernel for an exact zero
    u_self, _ = eng.predict(c["X_u"], want_f=False)
    eng.set_data(c["X_u"], u_self.astype(np.float64))
    loss, grad = eng.loss_grad()
    assert np.isfinite(loss) and np.isnan(grad).any()


@pytest.mark.parametrize("loss", [tg.LOSS_V4, tg.LOSS_V5])
@pytest.mark.parametrize("n_f", [1000, 148 * 8 * 32 * 16 + 77, (1 << 21) + 5])
def test_host_feed_in_chunks_equals_resident_batch(loss, n_f):
    """pinn_feed_collocation (the per-step feed_dict, INF-L2:127-135): the batch arrives from pinned host memory in
    chunks and the fused kernel accumulates over one launch per chunk -- same loss and gradient as the resident batch
    (only the summation order over warps differs), z/gamma offsets included, and a fed Adam step moves theta alike."""
    import torch
    c = make_case(tg.PDE_BURGERS, B20, loss, 100, 64, seed=21)
    lb, ub = c["prob"].lb, c["prob"].ub
    X = sample_collocation(5, 0, n_f, lb, ub)
    rng = np.random.default_rng(3)
    z = (0.3 * rng.standard_normal((n_f, 1))).astype(np.float32)
    g = (0.3 * rng.standard_normal((n_f, 1))).astype(np.float32)
    outs, thetas = [], []
    for fed in (False, True):
        eng = make_engine(c, trainable_lambda=True)
        host = torch.from_numpy(X.copy()).pin_memory()
        if fed:
            eng.feed_collocation(host)
        else:
            eng.set_collocation(X)
        if loss == tg.LOSS_V5:
            eng.admm_set_state(z, g)
        outs.append(eng.loss_grad())
        if fed:
            eng.feed_collocation(host)
        eng.adam_steps(1)
        thetas.append(eng.get_params())
        assert np.array_equal(eng.get_collocation(), X)
    assert abs(outs[0][0] - outs[1][0]) <= 2e-6 * abs(outs[0][0])
    assert rel_err(outs[1][1], outs[0][1]) <= 5e-6
    assert rel_err(thetas[1] - c["theta"], thetas[0] - c["theta"]) <= 1e-3   # Adam's first step is +-lr per parameter


def test_v1_norm_inside_the_fused_pass_and_its_fallback():
    """INF-L2's un-squared data norm (appendix A.3 V1) rides inside the fused pass when every warp has one batch at most
    (data batches own their accumulator regions, the reduction divides by ||r||); larger jobs take the separate data
    pass.  Both agree with the oracle, an Adam step follows the oracle, and zero misfit gives tf.norm's NaN gradient."""
    c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V1, 100, 10456, seed=31)
    ref = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"])
    eng = make_engine(c)                                  # 327 + 4 batches: inline
    assert eng.kernel_path == "fused"
    n0 = eng.launch_count
    loss, grad = eng.loss_grad()
    assert eng.launch_count - n0 == 2                     # one kernel + one reduction: no separate data pass
    assert abs(loss - ref.loss) <= TOL * abs(ref.loss) and rel_err(grad, ref.grad) <= TOL
    opt = TF1Adam(c["theta"].size)
    th1 = opt.step(c["theta"].astype(np.float64), ref.grad)
    eng.adam_steps(1)
    assert rel_err(eng.get_params() - c["theta"], th1 - c["theta"]) <= 1e-4
    # fallback: more batches than warps
    big = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V1, 100, 148 * 8 * 32 + 5, seed=32)
    refb = tg.evaluate(big["theta"], big["prob"], big["X_u"], big["u"], big["X_f"])
    engb = make_engine(big)
    lossb, gradb = engb.loss_grad()
    assert abs(lossb - refb.loss) <= TOL * abs(refb.loss) and rel_err(gradb, refb.grad) <= TOL
    # zero misfit on the inline path
    small = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V1, 16, 64, seed=3)
    e0 = make_engine(small)
    u_self, _ = e0.predict(small["X_u"], want_f=False)    # same kernel family as the data batches -> exact zero
    e0.set_data(small["X_u"], u_self.astype(np.float64))
    l0, g0 = e0.loss_grad()
    assert np.isfinite(l0) and np.isnan(g0).any()


@pytest.mark.parametrize("nine", ["1", "0"], ids=["nine-warps", "second-batch-on-one-warp-per-SM"])
@pytest.mark.parametrize("loss", [tg.LOSS_V4, tg.LOSS_V5], ids=["v4", "v5-admm"])
def test_config1_sized_batch_stays_on_the_small_batch_kernel(nine, loss, monkeypatch):
    """INF-L2 / INF-ADMM's 10 456 + 100 points are 1320 batches of 8 points for 1184 warps: the small-batch kernel runs them with
    nine warps per CTA (default) or, PINN_FUSED_SMALL_NINE=0, with a second batch on one warp per SM.  Both against the oracle,
    bit-reproducible, and the ADMM state they leave is the same."""
    monkeypatch.setenv("PINN_FUSED_SMALL_NINE", nine)
    c = make_case(tg.PDE_BURGERS, B20, loss, 100, 10456, seed=41)
    ref = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"], z=c.get("z"), gamma=c.get("gamma"))
    eng = make_engine(c)
    assert eng.kernel_path == "fused"
    n0 = eng.launch_count
    l1, g1 = eng.loss_grad()
    assert eng.launch_count - n0 == 2
    l2, g2 = eng.loss_grad()
    assert l1 == l2 and np.array_equal(g1, g2)
    assert abs(l1 - ref.loss) <= TOL * abs(ref.loss) and rel_err(g1, ref.grad) <= TOL


def test_tensor_path_run_to_run_determinism():
    """tcgen05 path: per-CTA partial gradients are updated with red.global.add by a fixed thread per element and summed
    over CTAs in fixed order -> bit-identical loss and gradient from run to run"""
    layers = [2] + [64] * 4 + [1]
    c = make_case(tg.PDE_BURGERS, layers, tg.LOSS_V4, 50, 40000, seed=9)
    eng = make_engine(c, path="tensor")
    assert eng.kernel_path == "tensor"
    l1, g1 = eng.loss_grad()
    l2, g2 = eng.loss_grad()
    assert l1 == l2 and np.array_equal(g1, g2)


def test_state_and_argument_errors_are_reported_not_crashed():
    """Error behaviour at the boundary: the reference lets TF raise (e.g. "You must feed a value for
    placeholder"); here every entry point returns a negative PINN_E_* code with a message and the handle stays usable."""
    from pinns_b200 import Engine
    from pinns_b200._capi import PinnError
    eng = Engine(B20, [-1.0, 0.0], [1.0, 0.99], loss="v4")
    for call in (eng.loss_grad, lambda: eng.adam_steps(1), eng.admm_update, eng.loss_value):
        with pytest.raises(PinnError) as e:                     # no collocation points fed yet
            call()
        assert "collocation" in str(e.value)
    with pytest.raises(PinnError):
        eng.set_collocation(np.zeros((0, 2)))                   # an empty feed is rejected, not launched
    eng.set_collocation(np.random.default_rng(0).random((65, 2)))
    with pytest.raises(PinnError) as e:
        eng.admm_adam_step()                                    # folded ADMM step on a non-ADMM loss
    assert "ADMM" in str(e.value)
    loss, grad = eng.loss_grad()                                # the handle is still good
    assert np.isfinite(loss) and np.isfinite(grad).all() and grad.shape == (eng.num_params,)
