"""Smallest end-to-end exercise of every kernel family, for compute-sanitizer --tool memcheck."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine
from tests.helpers import rand_theta
def run(layers, pde, loss, path, n_f, n_u=40):
    eng = Engine(layers, [-1, 0], [1, 0.99], pde=pde, loss=loss, lambda2=0.01 / np.pi, rho=3.0, path=path, trainable_lambda=(pde == 'burgers'))
    eng.set_params(rand_theta(layers, np.random.default_rng(0)))
    rng = np.random.default_rng(1)
    eng.set_data(rng.random((n_u, 2)), rng.random((n_u, layers[-1])))
    eng.sample_collocation(7, 0, n_f)
    if loss in ('v2', 'v5'): eng.admm_init()
    l, g = eng.loss_grad()
    eng.adam_steps(2)
    if loss in ('v2', 'v5'): eng.admm_update()
    u, f = eng.predict(rng.random((77, 2)))
    print(path, eng.kernel_path, layers[1], loss, n_f, 'loss %.4e' % l, 'finite', bool(np.isfinite(g).all()))
run([2] + [20] * 8 + [1], 'burgers', 'v4', 'auto', 2000)
run([2] + [20] * 8 + [1], 'burgers', 'v1', 'auto', 333)
run([2] + [20] * 3 + [1], 'burgers', 'v5', 'auto', 100)
run([2] + [64] * 3 + [1], 'burgers', 'v4', 'tensor', 300)
run([2] + [128] * 3 + [1], 'burgers', 'v5', 'tensor', 129)
run([2] + [50] * 3 + [1], 'burgers', 'v3', 'generic', 200)
run([2] + [40] * 2 + [3], 'euler', 'v5', 'auto', 150)
