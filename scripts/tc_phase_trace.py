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
names = {1: "tile start (prev tile layer-0 reverse)", 2: "layer 0", 3: "head+residual", 4: "head reverse"}
def name(t):
    if t in names: return names[t]
    if t >= 100: return 'fine %d' % t
    k, l = divmod(t, 10)
    return {1: 'F wait for MMAs', 2: 'F epilogue', 3: 'bias sums', 4: 'G wait for MMAs', 5: 'G flush', 6: 'B wait for MMAs', 7: 'B epilogue'}[k] + ' l=%d' % l
agg = collections.OrderedDict(); prev = None; tiles = 0
vals = collections.OrderedDict()
for tag, t in ev:
    if tag >= 200:
        vals.setdefault(tag, []).append(t)
ev = [e for e in ev if e[0] < 200]
ev.sort(key=lambda e: e[1])
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

vnames = {}
for k, nm in ((0, 'F'), (1, 'B'), (2, 'G')):
    vnames[200 + 10 * k] = nm + ' unit: issuer waited for the unit to begin'
    vnames[201 + 10 * k] = nm + ' unit: issuer waited for staged operands'
    vnames[202 + 10 * k] = nm + ' unit: issue span'
vnames[230] = 'G unit: splitter waited for the TMA copies'
vnames[231] = 'G unit: splitter spent splitting'
vnames[240] = 'F unit: splitter waited for TMA (stages 1..)'
vnames[241] = 'F unit: splitter spent splitting'
vnames[232] = 'G unit: splitter in fence.proxy.async'
for tag in sorted(vals):
    v = vals[tag]
    print('  %-48s mean %8.0f clk per unit (%d units)' % (vnames.get(tag, str(tag)), sum(v) / len(v), len(v)))
