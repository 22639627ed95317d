"""CPU tests of the host-side mirror of the reference classes, driven through a fake engine
(the real engine needs a B200): loop structure, stopping rules, update order, CSV schema."""
import os

import numpy as np
import pytest

from pinns_b200 import models
from pinns_b200.distributed import shard_range, data_weight

GOLD = os.path.join(os.path.dirname(__file__), "golden", "data")


class FakeEngine:
    def __init__(self, *a, **k):
        self.calls = []
        self.num_params = 5
        self.trainable_lambda = False
        self.n_f = 0
        self._loss = 10.0
        self.layers = a[0] if a else None

    def set_params(self, th): self.calls.append(("set_params", len(np.asarray(th).ravel())))
    def get_params(self): return np.zeros(self.num_params, np.float32)
    def set_data(self, X, u): self.calls.append(("set_data", X.shape, u.shape))
    def set_collocation(self, X, nf_global=0): self.calls.append(("set_collocation", X.shape)); self.n_f = X.shape[0]
    def sample_collocation(self, seed, first, n, nf_global=0): self.calls.append(("sample", seed, first, n)); self.n_f = n
    def adam_config(self, **k): pass
    def adam_steps(self, n): self.calls.append(("adam", n))
    def loss_value(self): self.calls.append(("loss",)); return self._loss
    def admm_init(self): self.calls.append(("admm_init",))
    def admm_update(self, inf_admm_quirk=False): self.calls.append(("admm_update", inf_admm_quirk))
    def admm_adam_step(self, inf_admm_quirk=False):  # one pass on the device; in the reference's order: update, then the Adam step
        self.calls.append(("admm_update", inf_admm_quirk)); self.calls.append(("adam", 1)); self.calls.append(("folded",))
    def resampled_epochs(self, n_epochs, admm, pending, seed, first_batch, n_f, nf_global=0):
        # pinn_resampled_epochs: the same calls the per-epoch loop makes, issued inside the library
        for k in range(n_epochs):
            if pending:
                self.admm_adam_step()
            else:
                self.adam_steps(1)
            self.sample_collocation(seed, (first_batch + k) * n_f, n_f, nf_global)
            pending = bool(admm)
    def predict(self, X, want_f=True):
        n = np.asarray(X).shape[0]
        no = self.layers[-1] if self.layers else 1
        return np.zeros((n, no), np.float32), (np.zeros((n, 1 if no == 1 else 3), np.float32) if want_f else None)


@pytest.fixture
def fake(monkeypatch):
    monkeypatch.setattr(models, "Engine", FakeEngine)
    return FakeEngine


def test_inference_train_loop_matches_reference_stop_rule(fake):
    X_u = np.random.rand(10, 2); u = np.random.rand(10, 1); X_f = np.random.rand(50, 2)
    m = models.PhysicsInformedNN(X_u, u, X_f, [2, 20, 20, 1], np.zeros(2), np.ones(2), 0.0, '0', verbose=False)
    n = m.train(250, 'f', '0')
    adam = [c for c in m.engine.calls if c[0] == "adam"]
    # INF-L2:134-141: 1 step + loss at iteration 0, 100, 200; stretches of 99 / 99 / 49 in between
    assert [c[1] for c in adam] == [1, 99, 1, 99, 1, 49] and n == 250
    assert sum(1 for c in m.engine.calls if c[0] == "loss") == 3
    # early stop: |loss| <= tol is only seen at a multiple of 100
    m2 = models.PhysicsInformedNN(X_u, u, X_f, [2, 20, 20, 1], np.zeros(2), np.ones(2), 0.0, '0', verbose=False)
    m2.engine._loss = 1e-5
    assert m2.train(1000) == 1   # one step, loss refreshed at iteration 0, loop exits


def test_inf_admm_updates_every_k_steps_with_the_graph_quirk(fake):
    X_u = np.random.rand(10, 2); u = np.random.rand(10, 1); X_f = np.random.rand(50, 2)
    m = models.PhysicsInformedNN_ADMM(X_u, u, X_f, [2, 20, 20, 1], np.zeros(2), np.ones(2), 0.0, None, 0.5, 'f', '0', verbose=False)
    assert ("admm_init",) in m.engine.calls
    m.engine.calls.clear()
    m.engine._loss = 1e9
    m.train(7, 3)
    seq = [c[0] if c[0] != "admm_update" else ("admm_update", c[1]) for c in m.engine.calls if c[0] in ("adam", "admm_update")]
    # z/lagrange update at iterations 0, 3, 6 (INF-ADMM:191-193), always with the double dual update of :106-107
    assert seq.count(("admm_update", True)) == 3 and seq[1] == ("admm_update", True)


def test_inference_admm_loop_order_with_and_without_the_folded_update(fake):
    """INF-ADMM:189-193: Adam step, then (every w steps) z_update + lagrange_update, loss every 100 iterations.  The folded
    form issues the same operations in the same order; an update due at a loss print is flushed before the loss is read."""
    X_u = np.random.rand(10, 2); u = np.random.rand(10, 1); X_f = np.random.rand(50, 2)
    seqs = []
    for fold in (False, True):
        m = models.PhysicsInformedNN_ADMM(X_u, u, X_f, [2, 20, 20, 1], np.zeros(2), np.ones(2), 0.0, 1, 0.5, 'f', '0', verbose=False)
        m._fold_admm = fold
        m.train(7, 2, 'f', '0')
        seqs.append([c for c in m.engine.calls if c[0] in ("adam", "admm_update", "loss", "folded")])
    plain = [c for c in seqs[0] if c[0] != "folded"]
    assert [c[0] for c in plain] == ["adam", "admm_update", "loss", "adam", "adam", "admm_update", "adam", "adam", "admm_update",
                                     "adam", "adam", "admm_update"]
    assert all(c == ("admm_update", True) for c in plain if c[0] == "admm_update")          # the double dual update
    assert [c for c in seqs[1] if c[0] != "folded"] == plain and sum(c[0] == "folded" for c in seqs[1]) == 2


def test_identification_class_runs_inside_constructor_and_records_csv(fake, tmp_path):
    class P(models.Parameters):
        N_u = 50; N_f = 64; rho = 10.0; epochs = 6; gpu = '0'
    out = str(tmp_path / "fig.png")
    m = models.BurgersIdentification(P(), variant="AB-ADMM", data=os.path.join(GOLD, "TwoSin_burgers_shock.npz"),
                                     verbose=False, filename=out)
    calls = m.engine.calls
    assert calls[1][0] == "set_data" and calls[1][1] == (50, 2)
    # AB-ADMM:206-226: epochs 1..5 -> Adam step, NEW batch, then z/gamma update on the new batch
    body = [c[0] for c in calls if c[0] in ("adam", "set_collocation", "admm_update")]
    assert body == ["set_collocation"] + ["adam", "set_collocation", "admm_update"] * 5
    assert sum(c[0] == "folded" for c in calls) == 4         # updates 1..4 ride in the next Adam step's pass, the last one alone
    m2 = models.BurgersIdentification(P(), variant="AB-ADMM", data=os.path.join(GOLD, "TwoSin_burgers_shock.npz"),
                                      verbose=False, filename=out, run=False)
    m2._fold_admm = False
    m2.run_NN()
    assert [c[0] for c in m2.engine.calls if c[0] in ("adam", "set_collocation", "admm_update", "folded")] == body
    assert m.X_star.shape == (513 * 101, 2) and np.isfinite(m.error_u)
    m.save_data(); m.save_data()
    lines = open(out[:-3] + "csv").read().splitlines()
    assert lines[0] == "x,t,u_pred,epoch" and lines.count("x,t,u_pred,epoch") == 2   # header re-emitted per append


def test_euler_class_schema_and_device_resampling(fake, tmp_path):
    class P(models.EulerParameters):
        N_data = 40; N_f = 32; pen = 40.0; epochs = 3; gpu = '0'
    m = models.EulerInference(P(), data=os.path.join(GOLD, "Abgrall_eulers.npz"), verbose=False, resample="device",
                              filename=str(tmp_path / "e.png"))
    samples = [c for c in m.engine.calls if c[0] == "sample"]
    assert [c[2] for c in samples] == [0, 32, 64]            # disjoint counter ranges of the job-wide Philox stream
    assert len(m.predict(m.X_star[:5])) == 6                 # EUL:260-272 returns six arrays
    m.save_data()
    assert open(str(tmp_path / "e.csv")).readline().strip() == "x,t,rho_pred,u_pred,E_pred,epoch"


def test_device_resampled_stretches_issue_the_per_epoch_calls(fake):
    """With the device sampler the plain epochs of a Dialect-B loop go to the library as one stretch (pinn_resampled_epochs):
    the operations and their order are those of the epoch-by-epoch loop, prints and flushes at the multiples of 1000 included."""
    class P(models.Parameters):
        N_u = 50; N_f = 64; rho = 10.0; epochs = 1; gpu = '0'
    seqs = []
    for per_epoch in (True, False):
        m = models.BurgersIdentification(P(), variant="AB-ADMM", data=os.path.join(GOLD, "TwoSin_burgers_shock.npz"),
                                         verbose=False, run=False, resample="device")
        m._per_epoch_calls = per_epoch
        m.engine.calls.clear()
        m.train(2005)
        seqs.append([c for c in m.engine.calls if c[0] in ("adam", "sample", "admm_update", "folded", "loss")])
    assert seqs[0] == seqs[1] and sum(c[0] == "loss" for c in seqs[0]) == 2 and sum(c[0] == "sample" for c in seqs[0]) == 2004


def test_xavier_init_flat_layout():
    th = models.xavier_init_flat([2, 20, 20, 1], np.random.default_rng(0))
    assert th.dtype == np.float32 and th.size == 2 * 20 + 20 + 400 + 20 + 20 + 1
    assert np.all(th[40:60] == 0) and np.all(th[460:480] == 0) and th[-1] == 0


def test_shard_range_partitions_exactly():
    for n, w in [(10, 3), (64 * 2 ** 20, 8), (7, 8), (1, 1)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f0 + c0 == f1
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert data_weight(0) == 1.0 and data_weight(3) == 0.0


def test_sweep_farm_replacement_runs_every_scenario_once_per_free_gpu():
    from pinns_b200.sweep import get_combinations, schedule_runs
    combos = get_combinations({"N_u": [100, 200], "N_f": [1000, 5000], "rho": [10.0]})
    assert len(combos) == 4 and combos[0] == {"N_u": 100, "N_f": 1000, "rho": 10.0} and combos[-1]["N_f"] == 5000
    launched = []

    class FakeProc:
        def __init__(self, argv):
            launched.append(argv)
            self.polls = 0

        def poll(self):
            self.polls += 1
            return 0 if self.polls >= 2 else None

    res = schedule_runs(["./Abgrall_ADMM.py"], combos, gpus=[0, 1], arg_order=["N_u", "N_f", "rho"], poll_s=0.0, popen=FakeProc)
    assert len(launched) == 4 and launched[0] == ["./Abgrall_ADMM.py", "100", "1000", "10.0", "0"]
    assert sorted(r["gpu"] for r in res) == [0, 0, 1, 1] and all(r["returncode"] == 0 for r in res)
