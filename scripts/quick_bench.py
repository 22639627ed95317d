import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine
from tests.helpers import rand_theta
layers=[2]+[20]*8+[1]
N = int(sys.argv[1]) if len(sys.argv)>1 else 1<<22
path = sys.argv[2] if len(sys.argv)>2 else 'auto'
eng = Engine(layers, [-1,0],[1,0.99], loss='v4', lambda2=0.01/np.pi, path=path)
eng.use_torch_stream()
eng.set_params(rand_theta(layers, np.random.default_rng(0)))
eng.set_data(np.random.rand(100,2), np.random.rand(100,1))
eng.sample_collocation(1234, 0, N)
print('path', eng.kernel_path)
for _ in range(3): eng.loss_grad_device()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
K=10
eng.kernel_timing(True)
ev[0].record()
for _ in range(K): eng.loss_grad_device()
ev[1].record(); torch.cuda.synchronize()
kms, kn = eng.kernel_time()
ms = ev[0].elapsed_time(ev[1])/K
print(f'kernel-only ms={kms/kn:.3f} -> {N/(kms/kn)/1e3:.1f} Mpts/s, frac_of_74.5={N*68320/(kms/kn)/1e9/74.5:.3f}')
print(f'N={N} ms/step={ms:.3f} Mpts/s={N/ms/1e3:.1f} TFLOPs={N*68320/ms/1e9:.2f} frac_of_74.5={N*68320/ms/1e9/74.5:.3f}')
