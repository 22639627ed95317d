"""Step time and residual-kernel time of the fused path at the reference's batch sizes (small-batch kernel on / off)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine
from tests.helpers import rand_theta
B20 = [2] + [20] * 8 + [1]
for n_f, loss, n_u in ((1000, 'v5', 100), (5000, 'v5', 100), (9372, 'v1', 100), (10456, 'v1', 100), (10771, 'v1', 100), (2000, 'v4', 2000), (20000, 'v4', 100), (37000, 'v4', 100)):
    eng = Engine(B20, [-1, 0], [1, 0.99], loss=loss, lambda2=0.01 / np.pi, rho=10.0)
    eng.use_torch_stream()
    eng.set_params(rand_theta(B20, np.random.default_rng(0)))
    rng = np.random.default_rng(1)
    eng.set_data(rng.random((n_u, 2)), rng.random((n_u, 1)))
    eng.sample_collocation(1234, 0, n_f)
    if loss == 'v5': eng.admm_init()
    eng.adam_steps(20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.adam_steps(500); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 500
    eng.kernel_timing(True); eng.adam_steps(50); k_ms, k_n = eng.kernel_time(); eng.kernel_timing(False)
    print('N_f=%6d N_u=%5d %s: %7.2f us/step, residual kernel %7.2f us (%d launches timed)' % (n_f, n_u, loss, 1e3 * ms, 1e3 * k_ms / max(k_n, 1), k_n), flush=True)
