import sys, numpy as np, time
sys.path.insert(0, '/root/repo')
from oracle import tf_graph as tg
from tests.helpers import make_case, make_engine, rel_err, max_rel_err
for n, nl, nf in [(32, 3, 300), (64, 4, 777), (128, 8, 1000), (96, 5, 129)]:
    layers = [2] + [n] * nl + [1]
    for loss in (tg.LOSS_V4, tg.LOSS_V5):
        c = make_case(tg.PDE_BURGERS, layers, loss, 50, nf, seed=n + nl)
        ref = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"], c["z"], c["gamma"])
        eng = make_engine(c, path="tensor", trainable_lambda=True)
        assert eng.kernel_path == "tensor"
        loss_gpu, grad = eng.loss_grad()
        P = eng.num_params
        u, f = eng.predict(c["X_f"])
        _, f_ref = tg.predict(c["theta"], c["prob"], c["X_f"])
        print('n=%d NL=%d nf=%d %s: loss rel %.2e  grad L2 rel %.2e  max rel %.2e  dlam %s vs %s  f max rel %.2e' % (
            n, nl, nf, loss, abs(loss_gpu - ref.loss) / abs(ref.loss), rel_err(grad[:P], ref.grad), max_rel_err(grad[:P], ref.grad),
            grad[P:], ref.dlam, max_rel_err(f, f_ref)))
