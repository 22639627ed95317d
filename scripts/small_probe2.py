"""Step-time anatomy at the reference's batch sizes: fused kernel alone vs whole Adam step."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine
from tests.helpers import rand_theta
layers = [2] + [20] * 8 + [1]
for N in (1000, 2456, 10456, 37888, 151552):
    eng = Engine(layers, [-1, 0], [1, 0.99], loss='v4', lambda2=0.01 / np.pi)
    eng.use_torch_stream()
    eng.set_params(rand_theta(layers, np.random.default_rng(0)))
    eng.set_data(np.random.rand(100, 2), np.random.rand(100, 1))
    eng.sample_collocation(1234, 0, N)
    eng.adam_steps(200); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.adam_steps(500); e1.record(); torch.cuda.synchronize()
    step = e0.elapsed_time(e1) / 500
    eng.kernel_timing(True)
    eng.adam_steps(50); torch.cuda.synchronize()
    ms, k = eng.kernel_time()
    eng.kernel_timing(False)
    print('N=%6d  step %.1f us   fused kernel %.1f us   rest %.1f us' % (N, 1e3 * step, 1e3 * ms / k, 1e3 * (step - ms / k)))
