"""GPU parity: CUDA loss / residuals / gradient vs the fp64 oracle on identical
weights and inputs, through the C ABI.  Tolerance (BASELINE.json north_star, fp32
path): 1e-5 relative."""
import zlib

import numpy as np
import pytest

from oracle import tf_graph as tg
from tests.helpers import make_case, make_engine, rel_err, max_rel_err

TOL = 1e-5

BURGERS20 = [2] + [20] * 8 + [1]

CASES = [
    ("burgers-v4-20", tg.PDE_BURGERS, BURGERS20, tg.LOSS_V4, 100, 1000),
    ("burgers-v1-20", tg.PDE_BURGERS, BURGERS20, tg.LOSS_V1, 100, 1000),
    ("burgers-v3-20", tg.PDE_BURGERS, BURGERS20, tg.LOSS_V3, 100, 1000),
    ("burgers-v5-20", tg.PDE_BURGERS, BURGERS20, tg.LOSS_V5, 100, 1000),
    ("burgers-v2-20", tg.PDE_BURGERS, BURGERS20, tg.LOSS_V2, 100, 1000),
    ("burgers-v4-ragged", tg.PDE_BURGERS, [2, 13, 27, 9, 1], tg.LOSS_V4, 37, 333),
    ("burgers-v4-128", tg.PDE_BURGERS, [2] + [128] * 8 + [1], tg.LOSS_V4, 50, 300),
    ("burgers-v4-200", tg.PDE_BURGERS, [2] + [200] * 8 + [1], tg.LOSS_V4, 50, 200),
    ("euler-mse-200", tg.PDE_EULER, [2] + [200] * 5 + [3], tg.LOSS_EULER_MSE, 200, 1000),
    ("euler-admm-200", tg.PDE_EULER, [2] + [200] * 5 + [3], tg.LOSS_V6, 200, 1000),
    ("euler-admm-ragged", tg.PDE_EULER, [2, 30, 17, 3], tg.LOSS_V6, 11, 97),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,pde,layers,loss,n_u,n_f", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("path", ["generic", "auto"])
def test_loss_grad_parity(name, pde, layers, loss, n_u, n_f, path):
    case = make_case(pde, layers, loss, n_u, n_f, seed=zlib.crc32(name.encode()) % 1000)
    ref = tg.evaluate(case["theta"], case["prob"], case["X_u"], case["u"], case["X_f"], case["z"], case["gamma"])
    eng = make_engine(case, path=path, trainable_lambda=(pde == tg.PDE_BURGERS))
    loss_gpu, grad_gpu = eng.loss_grad()
    P = eng.num_params
    assert abs(loss_gpu - ref.loss) <= TOL * abs(ref.loss), (loss_gpu, ref.loss)
    assert rel_err(grad_gpu[:P], ref.grad) <= TOL, rel_err(grad_gpu[:P], ref.grad)
    assert max_rel_err(grad_gpu[:P], ref.grad) <= 2 * TOL
    if pde == tg.PDE_BURGERS:
        assert np.allclose(grad_gpu[P:], ref.dlam, rtol=2e-5, atol=1e-6 * max(1.0, np.abs(ref.dlam).max()))
    # residuals and outputs through predict()
    u_gpu, f_gpu = eng.predict(case["X_f"])
    _, f_ref = tg.predict(case["theta"], case["prob"], case["X_f"])
    assert max_rel_err(f_gpu, f_ref) <= TOL
    u_gpu2, _ = eng.predict(case["X_u"], want_f=False)
    assert max_rel_err(u_gpu2, ref.u_pred) <= TOL
    # loss value without gradient
    assert abs(eng.loss_value() - ref.loss) <= TOL * abs(ref.loss)
