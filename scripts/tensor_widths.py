import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine
from tests.helpers import rand_theta
for n in (32, 64, 96, 128):
    for path in ('tensor', 'generic'):
        layers = [2] + [n] * 8 + [1]
        N = 148 * 128 * 8
        eng = Engine(layers, [-1, 0], [1, 0.99], loss='v4', lambda2=0.01 / np.pi, path=path)
        eng.use_torch_stream()
        eng.set_params(rand_theta(layers, np.random.default_rng(0)))
        eng.set_data(np.random.rand(100, 2), np.random.rand(100, 1))
        eng.sample_collocation(1234, 0, N)
        for _ in range(2): eng.loss_grad_device()
        torch.cuda.synchronize()
        eng.kernel_timing(True)
        for _ in range(3): eng.loss_grad_device()
        ms, k = eng.kernel_time()
        print('n=%3d %-8s %.3f ms -> %.2f Mpts/s' % (n, path, ms / k, N / (ms / k) / 1e3))
