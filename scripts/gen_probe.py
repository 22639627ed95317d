import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine
from tests.helpers import rand_theta
for layers, n_f in (([2]+[20]*8+[1], 65536), ([2]+[20]*8+[1], 1000), ([2]+[64]*4+[1], 65536), ([2]+[64]*4+[1], 2000), ([2]+[128]*8+[1], 4096), ([2]+[128]*8+[1], 65536), ([2]+[50]*4+[1], 30000)):
    eng = Engine(layers, [-1, 0], [1, 0.99], loss='v4', lambda2=0.01/np.pi, path='generic')
    eng.use_torch_stream()
    eng.set_params(rand_theta(layers, np.random.default_rng(0)))
    rng = np.random.default_rng(1)
    eng.set_data(rng.random((100, 2)), rng.random((100, 1)))
    eng.sample_collocation(1234, 0, n_f)
    eng.adam_steps(3); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 20 if n_f <= 4096 else 5
    e0.record(); eng.adam_steps(k); e1.record(); torch.cuda.synchronize()
    print('%-28s N=%6d  %8.3f ms/step' % (str(layers[1])+'x'+str(len(layers)-2), n_f, e0.elapsed_time(e1)/k), flush=True)
