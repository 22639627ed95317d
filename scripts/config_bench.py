"""Step time / throughput of the BASELINE configs other than the headline (parity-test cases, not bench lines)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine
from tests.helpers import rand_theta

def run(name, layers, pde, loss, n_u, n_f, steps=20, fpp=None, path='auto'):
    eng = Engine(layers, [-1, 0], [1, 0.99], pde=pde, loss=loss, lambda2=0.01 / np.pi, rho=40.0, path=path)
    eng.use_torch_stream()
    eng.set_params(rand_theta(layers, np.random.default_rng(0)))
    rng = np.random.default_rng(1)
    eng.set_data(rng.random((n_u, 2)), rng.random((n_u, layers[-1])))
    eng.sample_collocation(1234, 0, n_f)
    if loss == 'v5': eng.admm_init()
    for _ in range(3): eng.adam_steps(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.adam_steps(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    tf = '' if fpp is None else ' %.2f TFLOP/s (%.1f%% of 74.5)' % (n_f * fpp / ms / 1e9, 100 * n_f * fpp / ms / 1e9 / 74.5)
    print('%-34s path=%-7s N_f=%-9d %9.3f ms/step %10.2f Mpts/s%s' % (name, eng.kernel_path, n_f, ms, n_f / ms / 1e3, tf), flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
B20, B128, B200, EUL = [2] + [20] * 8 + [1], [2] + [128] * 8 + [1], [2] + [200] * 8 + [1], [2] + [200] * 5 + [3]
if which in ('all', 'small'):
    run('config1 burgers20 N=10456', B20, 'burgers', 'v1', 100, 10456, 200, 68320)
    run('burgers20 N=1000 (AB batch)', B20, 'burgers', 'v5', 100, 1000, 200, 68320)
    run('burgers20 N=1M', B20, 'burgers', 'v4', 100, 1 << 20, 20, 68320)
    run('config3 euler200x5 N=1000', EUL, 'euler', 'v5', 200, 1000, 50, 2895600)
    run('burgers200x8 N=1000 (AB-L2)', B200, 'burgers', 'v4', 100, 1000, 50, 6731200)
if which in ('all', 'tensor'):
    run('euler200x5 N=65536 generic', EUL, 'euler', 'v5', 200, 65536, 3, 2895600, path='generic')
    run('euler200x5 N=65536 TENSOR', EUL, 'euler', 'v5', 200, 65536, 5, 2895600, path='tensor')
    run('euler200x5 N=1M TENSOR', EUL, 'euler', 'v5', 200, 1 << 20, 2, 2895600, path='tensor')
    run('euler200x5 N=8192 TENSOR', EUL, 'euler', 'v5', 200, 8192, 10, 2895600, path='tensor')
    run('euler200x5 N=1000 TENSOR', EUL, 'euler', 'v5', 200, 1000, 20, 2895600, path='tensor')
    run('config5 burgers128 N=262144 TENSOR', B128, 'burgers', 'v4', 100, 262144, 5, 2759680, path='tensor')
    run('burgers128 N=2M TENSOR', B128, 'burgers', 'v4', 100, 1 << 21, 2, 2759680, path='tensor')
    run('burgers128 N=1000 TENSOR', B128, 'burgers', 'v4', 100, 1000, 20, 2759680, path='tensor')
    run('burgers200x8 N=262144 TENSOR', B200, 'burgers', 'v4', 100, 262144, 3, 6731200, path='tensor')
    run('burgers200x8 N=65536 generic', B200, 'burgers', 'v4', 100, 65536, 2, 6731200, path='generic')
    run('burgers64x8 N=262144 TENSOR', [2] + [64] * 8 + [1], 'burgers', 'v4', 100, 262144, 5, None, path='tensor')
    run('burgers32x8 N=262144 TENSOR', [2] + [32] * 8 + [1], 'burgers', 'v4', 100, 262144, 5, None, path='tensor')
if which == 'tensor_short':
    run('euler200x5 N=65536 TENSOR', EUL, 'euler', 'v5', 200, 65536, 5, 2895600, path='tensor')
    run('config5 burgers128 N=262144 TENSOR', B128, 'burgers', 'v4', 100, 262144, 5, 2759680, path='tensor')
    run('burgers200x8 N=262144 TENSOR', B200, 'burgers', 'v4', 100, 262144, 3, 6731200, path='tensor')
    run('burgers64x8 N=262144 TENSOR', [2] + [64] * 8 + [1], 'burgers', 'v4', 100, 262144, 5, None, path='tensor')
