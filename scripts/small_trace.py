"""Phase timeline of one warp of the small-batch fused kernel (library built with -DPINN_FUSED_SMALL_TRACE)."""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine, _capi
from tests.helpers import rand_theta
B20 = [2] + [20] * 8 + [1]
n_f = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
eng = Engine(B20, [-1, 0], [1, 0.99], loss='v4', lambda2=0.01 / np.pi)
eng.set_params(rand_theta(B20, np.random.default_rng(0)))
eng.set_data(np.random.rand(100, 2), np.random.rand(100, 1))
eng.sample_collocation(1234, 0, n_f)
lib = C.CDLL(_capi.LIB_PATH)
buf = (C.c_longlong * 512)(); cnt = C.c_int(0)
for _ in range(3): eng.adam_steps(1)
lib.pinn_sk_debug_trace(buf, C.byref(cnt))
eng.adam_steps(1)
lib.pinn_sk_debug_trace(buf, C.byref(cnt))
ev = [(buf[2 * i], buf[2 * i + 1]) for i in range(cnt.value)]
names = {0: 'start', 1: 'theta in smem', 2: 'weight copies built', 3: 'forward done', 4: 'head + seeds + head reverse', 5: 'reverse layers done', 6: 'layer 0 reverse'}
t0 = ev[0][1]; p = t0
for tag, t in ev:
    nm = names.get(tag, ('reverse l=%d: stash->tile, G' % (tag - 10)) if tag < 30 else 'reverse l=%d: B + zbar' % (tag - 30))
    print('%-34s t=%8d (+%6d clk)' % (nm, t - t0, t - p)); p = t
