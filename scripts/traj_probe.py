import sys, json, os, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import tf_graph as tg
from tests.golden.make_fixtures import trajectory_schedule
from pinns_b200 import Engine
gold = json.load(open('/root/repo/tests/golden/trajectory_identification.json'))
g, layers, theta0, prob, sched = trajectory_schedule('identification')
eng = Engine(layers, prob.lb, prob.ub, pde=prob.pde, loss="v4", lambda1=prob.lam1, lambda2=prob.lam2, rho=prob.rho, trainable_lambda=True)
eng.set_params(theta0); eng.set_data(g["X_u"], g["u"]); eng.set_collocation(g["X_f"])
done=0
for k, step in enumerate(gold["steps"]):
    eng.adam_steps(step-done); done=step
    l1,l2=eng.get_lambda()
    print(step, "loss %.5e (oracle %.5e)  lam1 %.5f (%.5f) lam2 %.6f (%.6f)"%(eng.loss_value(), gold["loss"][k], l1, gold["lambda1"][k], l2, gold["lambda2"][k]))
# loss around the end: every 10 steps
for _ in range(10):
    eng.adam_steps(10); print("   +10: %.5e"%eng.loss_value())
pred,_=eng.predict(g["X_star"], want_f=False)
print("error_u %.4f (oracle %.4f)"%(tg.relative_l2(g["u_star"],pred), gold["error_u"]))
