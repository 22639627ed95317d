"""Kernel-only throughput of the tcgen05 path (CUDA events around the kernel), after ~1 s of warm-up so that the SM
clock has ramped up; prints the SM clock sampled right after the timed launches."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine
from tests.helpers import rand_theta
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 4
path = sys.argv[3] if len(sys.argv) > 3 else 'tensor'
layers = [2] + [n] * 8 + [1]
eng = Engine(layers, [-1, 0], [1, 0.99], loss='v4', lambda2=0.01 / np.pi, path=path)
eng.use_torch_stream()
eng.set_params(rand_theta(layers, np.random.default_rng(0)))
eng.set_data(np.random.rand(100, 2), np.random.rand(100, 1))
eng.sample_collocation(1234, 0, N)
eng.loss_grad_device(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.loss_grad_device(); e1.record(); torch.cuda.synchronize()
warm = max(3, int(1000.0 / max(e0.elapsed_time(e1), 0.01)))
for _ in range(min(warm, 2000)): eng.loss_grad_device()
torch.cuda.synchronize()
eng.kernel_timing(True)
for _ in range(5): eng.loss_grad_device()
ms, k = eng.kernel_time()
mhz = None
try:
    import pynvml
    pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
    eng.loss_grad_device(); mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    torch.cuda.synchronize()
except Exception:
    pass
print('%s kernel n=%d N=%d: %.3f ms -> %.2f Mpts/s  (sm clock %s MHz)' % (eng.kernel_path, n, N, ms / k, N / (ms / k) / 1e3, mhz))
