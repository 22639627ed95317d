"""One Adam step of Euler [2,200x5,3] at the reference's batch (N_f = 1000, N_data = 200): run under
`ncu --metrics gpu__time_duration.sum` to list the launches of a step."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine
from tests.helpers import rand_theta
layers = [2] + [200] * 5 + [3]
eng = Engine(layers, [-1, 0], [1, 0.99], pde="euler", loss="v5", rho=40.0)
eng.set_params(rand_theta(layers, np.random.default_rng(0)))
rng = np.random.default_rng(1)
eng.set_data(rng.random((200, 2)), rng.random((200, 3)))
eng.sample_collocation(1234, 0, 1000)
eng.admm_init()
eng.adam_steps(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
eng.synchronize()
