"""Measured parity of the fused path against the fp64 oracle (how much of the 1e-5 budget is used)."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import tf_graph as tg
from tests.helpers import make_case, make_engine, rel_err, max_rel_err
B20 = [2] + [20] * 8 + [1]
for loss in (tg.LOSS_V4, tg.LOSS_V1, tg.LOSS_V5, tg.LOSS_V3):
    for n_f, seed in ((1000, 1), (10456, 2), (40000, 3)):
        c = make_case(tg.PDE_BURGERS, B20, loss, 100, n_f, seed=seed)
        ref = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"], c["z"], c["gamma"])
        eng = make_engine(c, trainable_lambda=True)
        l, g = eng.loss_grad()
        P = eng.num_params
        u, f = eng.predict(c["X_f"])
        _, f_ref = tg.predict(c["theta"], c["prob"], c["X_f"])
        print('%-8s N_f=%6d path=%s  loss rel %.2e   grad L2-rel %.2e  max-rel %.2e   f max-rel %.2e' % (
            loss, n_f, eng.kernel_path, abs(l - ref.loss) / abs(ref.loss), rel_err(g[:P], ref.grad), max_rel_err(g[:P], ref.grad), max_rel_err(f, f_ref)))
