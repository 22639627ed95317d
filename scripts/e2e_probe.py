import sys, json, os, numpy as np
sys.path.insert(0, '/root/repo')
from tests.golden.make_fixtures import e2e_schedule
from pinns_b200.models import PhysicsInformedNN
from oracle import tf_graph as tg
g, layers, theta0, prob, sched = e2e_schedule()
m = PhysicsInformedNN(g["X_u"], g["u"], g["X_f"], layers, g["lb"], g["ub"], 0.01/np.pi, '0', theta0=theta0, loss="v4", verbose=False)
m.engine.adam_steps(sched["adam_steps"])
u,_ = m.predict(g["X_star"]); print('after adam err', tg.relative_l2(g["u_star"], u), 'loss', m.engine.loss_value())
res = m.lbfgs_minimize(sched["lbfgs"])
u,_ = m.predict(g["X_star"]); print('after lbfgs err', tg.relative_l2(g["u_star"], u), 'loss', res.fun, res.nit, res.nfev, res.message)
res = m.lbfgs_minimize({'maxiter': 3000, 'maxfun': 5000, 'maxcor': 50, 'maxls': 50, 'ftol': 1e-12})
u,_ = m.predict(g["X_star"]); print('after more lbfgs err', tg.relative_l2(g["u_star"], u), 'loss', res.fun, res.nit, res.nfev, res.message)
print('--- full config 1')
g, layers, theta0, prob, sched = e2e_schedule(full=True)
m = PhysicsInformedNN(g["X_u"], g["u"], g["X_f"], layers, g["lb"], g["ub"], 0.01/np.pi, '0', theta0=theta0, loss="v4", verbose=False)
import time; t0=time.time()
m.engine.adam_steps(sched["adam_steps"]); m.engine.synchronize(); t1=time.time()
u,_ = m.predict(g["X_star"]); print('after adam err', tg.relative_l2(g["u_star"], u), 'loss', m.engine.loss_value(), 'adam time %.2fs'%(t1-t0))
res = m.lbfgs_minimize(sched["lbfgs"]); t2=time.time()
u,_ = m.predict(g["X_star"]); print('after lbfgs err', tg.relative_l2(g["u_star"], u), 'loss', res.fun, res.nit, res.nfev, res.message, 'lbfgs time %.2fs'%(t2-t1))
