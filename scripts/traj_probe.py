import sys, json, os, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import tf_graph as tg
from tests.golden.make_fixtures import trajectory_schedule
from pinns_b200 import Engine
for which in ("identification", "euler"):
    gold = json.load(open('/root/repo/tests/golden/trajectory_%s.json' % which))
    g, layers, theta0, prob, sched = trajectory_schedule(which)
    tr = which == "identification"
    eng = Engine(layers, prob.lb, prob.ub, pde=prob.pde, loss="v4", lambda1=prob.lam1, lambda2=prob.lam2, rho=prob.rho, trainable_lambda=tr)
    eng.set_params(theta0); eng.set_data(g["X_u"], g["u"]); eng.set_collocation(g["X_f"])
    done = 0
    for k, step in enumerate(gold["steps"]):
        eng.adam_steps(step - done); done = step
        print(which, step, eng.loss_value(), gold["loss"][k], eng.get_lambda(), gold["lambda1"][k], gold["lambda2"][k])
    pred, _ = eng.predict(g["X_star"], want_f=False)
    if tr: print('error_u', tg.relative_l2(g["u_star"], pred), gold["error_u"])
    else: print('errors', tg.relative_l2(g["rho_star"], pred[:, 0:1]), tg.relative_l2(g["u_star"], pred[:, 1:2]), tg.relative_l2(g["E_star"], pred[:, 2:3]), gold["error_rho"], gold["error_u"], gold["error_E"])
