"""Small-batch probe: one reference-sized configuration, a few Adam steps (for ncu launch lists / captures)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine
from tests.helpers import rand_theta

which = sys.argv[1] if len(sys.argv) > 1 else 'euler'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfgs = {
    'euler': ([2] + [200] * 5 + [3], 'euler', 'v5', 200, 1000),
    'b200': ([2] + [200] * 8 + [1], 'burgers', 'v4', 100, 1000),
    'b20': ([2] + [20] * 8 + [1], 'burgers', 'v5', 100, 1000),
    'c1': ([2] + [20] * 8 + [1], 'burgers', 'v1', 100, 10456),
}
layers, pde, loss, n_u, n_f = cfgs[which]
eng = Engine(layers, [-1, 0], [1, 0.99], pde=pde, loss=loss, lambda2=0.01 / np.pi, rho=40.0)
eng.use_torch_stream()
eng.set_params(rand_theta(layers, np.random.default_rng(0)))
rng = np.random.default_rng(1)
eng.set_data(rng.random((n_u, 2)), rng.random((n_u, layers[-1])))
eng.sample_collocation(1234, 0, n_f)
if loss == 'v5':
    eng.admm_init()
eng.adam_steps(3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.adam_steps(steps); e1.record(); torch.cuda.synchronize()
print('%s path=%s %.3f ms/step launches=%d' % (which, eng.kernel_path, e0.elapsed_time(e1) / steps, eng.launch_count))
