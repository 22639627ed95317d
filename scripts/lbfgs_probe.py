"""Wall time per L-BFGS-B function evaluation on config 1 (SciPy on the host, loss+grad from the GPU)."""
import sys, time, json, numpy as np
sys.path.insert(0, '/root/repo')
from tests.golden.make_fixtures import e2e_schedule
from pinns_b200.models import PhysicsInformedNN
g, layers, theta0, prob, sched = e2e_schedule(True)
m = PhysicsInformedNN(g["X_u"], g["u"], g["X_f"], layers, g["lb"], g["ub"], 0.01 / np.pi, '0', theta0=theta0, loss="v1", verbose=False)
t0 = time.perf_counter(); m.engine.adam_steps(2000); m.engine.synchronize(); t1 = time.perf_counter()
print('2000 Adam steps: %.3f s (%.1f us/step)' % (t1 - t0, 1e6 * (t1 - t0) / 2000))
# raw callback cost
x = m.engine.get_params().astype(np.float64)
t0 = time.perf_counter()
for _ in range(500):
    m.engine.set_params(x); l, gr = m.engine.loss_grad(); gr = gr.astype(np.float64)
t1 = time.perf_counter()
print('callback alone: %.1f us per evaluation' % (1e6 * (t1 - t0) / 500))
t0 = time.perf_counter()
res = m.lbfgs_minimize({'maxiter': 2000, 'maxfun': 50000, 'maxcor': 50, 'maxls': 50, 'ftol': 1.0 * np.finfo(float).eps})
t1 = time.perf_counter()
print('L-BFGS-B: %d iterations, %d evaluations, %.3f s -> %.1f us per evaluation; loss %.3e' % (res.nit, res.nfev, t1 - t0, 1e6 * (t1 - t0) / res.nfev, res.fun))
