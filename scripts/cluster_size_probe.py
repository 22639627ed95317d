import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine
from tests.helpers import rand_theta
"""Step time of the generic kernel's small-batch (cluster-per-tile) mode; PINN_GEN_CLUSTER_MAX caps the cluster size."""
sizes = [int(a) for a in sys.argv[1:]] or [1000]
import os
ND = int(os.environ.get("PROBE_N_DATA", "200"))
cases = [("euler200x5", [2]+[200]*5+[3], "euler", "v5", ND, n) for n in sizes] + [("burgers200x8", [2]+[200]*8+[1], "burgers", "v4", 100, n) for n in sizes]
for name, layers, pde, loss, n_u, n_f in cases:
    eng = Engine(layers, [-1, 0], [1, 0.99], pde=pde, loss=loss, lambda2=0.01/np.pi, rho=40.0)
    eng.use_torch_stream()
    eng.set_params(rand_theta(layers, np.random.default_rng(0)))
    rng = np.random.default_rng(1)
    eng.set_data(rng.random((n_u, 2)), rng.random((n_u, layers[-1])))
    eng.sample_collocation(1234, 0, n_f)
    if loss == "v5": eng.admm_init()
    eng.adam_steps(20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.adam_steps(200); e1.record(); torch.cuda.synchronize()
    print("%-14s N_f=%5d  %.1f us/step" % (name, n_f, e0.elapsed_time(e1) / 200 * 1e3))
