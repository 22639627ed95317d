import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine
from tests.helpers import rand_theta
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 4
layers = [2] + [n] * 8 + [1]
eng = Engine(layers, [-1, 0], [1, 0.99], loss='v4', lambda2=0.01 / np.pi, path='tensor')
eng.use_torch_stream()
eng.set_params(rand_theta(layers, np.random.default_rng(0)))
eng.set_data(np.random.rand(100, 2), np.random.rand(100, 1))
eng.sample_collocation(1234, 0, N)
for _ in range(2): eng.loss_grad_device()
torch.cuda.synchronize()
eng.kernel_timing(True)
for _ in range(3): eng.loss_grad_device()
ms, k = eng.kernel_time()
print('tensor kernel n=%d N=%d: %.3f ms -> %.2f Mpts/s' % (n, N, ms / k, N / (ms / k) / 1e3))
