"""Phase timeline of CTA 0 of the tcgen05 kernel (library built with `make EXTRA=-DPINN_TC_TRACE`)."""
import sys, ctypes as C, numpy as np, torch, collections
sys.path.insert(0, '/root/repo')
from pinns_b200 import Engine, _capi
from tests.helpers import rand_theta
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 4
layers = [2] + [n] * 8 + [1]
eng = Engine(layers, [-1, 0], [1, 0.99], loss='v4', lambda2=0.01 / np.pi, path='tensor')
eng.set_params(rand_theta(layers, np.random.default_rng(0)))
eng.set_data(np.random.rand(100, 2), np.random.rand(100, 1))
eng.sample_collocation(1234, 0, N)
lib = C.CDLL(_capi.LIB_PATH) if hasattr(_capi, 'LIB_PATH') else _capi.lib
buf = (C.c_longlong * 4096)(); cnt = C.c_int(0)
for _ in range(2): eng.loss_grad_device()
lib.pinn_tc_debug_trace(buf, C.byref(cnt))
eng.loss_grad_device()
lib.pinn_tc_debug_trace(buf, C.byref(cnt))
ev = [(buf[2 * i], buf[2 * i + 1]) for i in range(cnt.value)]
names = {1: 'tile start', 2: 'layer 0', 3: 'head+residual', 4: 'head reverse'}
def name(t):
    if t in names: return names[t]
    if t >= 100: return 'fine %d' % t
    k, l = divmod(t, 10)
    return {1: 'F contract', 2: 'F epilogue', 3: 'bias sums', 4: 'G contract', 5: 'G flush', 6: 'B contract', 7: 'B epilogue'}[k] + ' l=%d' % l
agg = collections.OrderedDict(); prev = None; tiles = 0
for tag, t in ev:
    if tag == 1: tiles += 1
    if tag >= 100: continue
    if prev is not None and tag != 1:
        k = name(tag).split(' l=')[0]
        agg[k] = agg.get(k, 0) + (t - prev)
    prev = t
tot = sum(agg.values())
print('tiles traced %d, cycles per tile %.0f (%.1f us at 1.9 GHz)' % (tiles, tot / max(tiles, 1), tot / max(tiles, 1) / 1900))
for k, v in agg.items(): print('  %-16s %9.0f clk/tile  %5.1f%%' % (k, v / tiles, 100 * v / tot))
if len(sys.argv) > 3:
    t0 = ev[0][1]; p = t0
    for tag, t in ev[:80]:
        print('%-18s t=%9d (+%7d)' % (name(tag), t - t0, t - p)); p = t
# fine trace (-DPINN_TC_TRACE_FINE): raw event list of the first layers
if any(tag >= 100 for tag, _ in ev):
    t0 = ev[0][1]; p = t0
    for tag, t in ev[:int(sys.argv[4]) if len(sys.argv) > 4 else 400]:
        print('%4d t=%9d (+%7d)' % (tag, t - t0, t - p)); p = t
