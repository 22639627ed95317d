"""Phase timeline of CTA 0 of the generic kernel (library built with `make EXTRA=-DPINN_TRACE`)."""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine, _capi
from tests.helpers import rand_theta

which = sys.argv[1] if len(sys.argv) > 1 else 'euler'
cfgs = {'euler': ([2] + [200] * 5 + [3], 'euler', 'v5', 200, 1000), 'b200': ([2] + [200] * 8 + [1], 'burgers', 'v4', 100, 1000)}
layers, pde, loss, n_u, n_f = cfgs[which]
eng = Engine(layers, [-1, 0], [1, 0.99], pde=pde, loss=loss, lambda2=0.01 / np.pi, rho=40.0)
eng.set_params(rand_theta(layers, np.random.default_rng(0)))
rng = np.random.default_rng(1)
eng.set_data(rng.random((n_u, 2)), rng.random((n_u, layers[-1])))
eng.sample_collocation(1234, 0, n_f)
if loss == 'v5':
    eng.admm_init()
eng.adam_steps(3)
lib = _capi.load_library()
buf = (C.c_longlong * 2048)(); n = C.c_int(0)
lib.pinn_debug_trace(buf, C.byref(n))      # drop the warm-up entries
eng.adam_steps(1)
lib.pinn_debug_trace(buf, C.byref(n))
ev = [(buf[2 * i], buf[2 * i + 1]) for i in range(n.value)]
t0 = ev[0][1]
prev = t0
for tag, t in ev:
    print('tag %3d  t=%8d clk  (+%6d)' % (tag, t - t0, t - prev))
    prev = t
