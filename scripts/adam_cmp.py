import sys, numpy as np
sys.path.insert(0, '/root/repo')
from tests.helpers import make_case, make_engine
from oracle import tf_graph as tg
from oracle.optim import TF1Adam
B20=[2]+[20]*8+[1]
c = make_case(tg.PDE_BURGERS, B20, tg.LOSS_V4, 100, 2456, seed=5)
c['u'] = np.sin(np.pi*c['X_u'][:,0:1])
K=int(sys.argv[1]) if len(sys.argv)>1 else 100
e1 = make_engine(c, path='fused'); e1.adam_steps(K); t1 = e1.get_params()
e2 = make_engine(c, path='fused')
for _ in range(K): e2.loss_grad_device(); e2.adam_apply()
t2 = e2.get_params()
e3 = make_engine(c, path='generic'); e3.adam_steps(K); t3 = e3.get_params()
theta = c['theta'].astype(np.float64); opt = TF1Adam(theta.size)
for _ in range(K):
    ev = tg.evaluate(theta, c['prob'], c['X_u'], c['u'], c['X_f']); theta = opt.step(theta, ev.grad)
rel = lambda a,b: np.linalg.norm(a-b)/np.linalg.norm(b)
print('K',K,'fused-lane2 vs fused-separate', rel(t1,t2), 'fused-lane2 vs generic', rel(t1,t3), 'lane2 vs oracle', rel(t1,theta), 'generic vs oracle', rel(t3,theta))
print('losses', e1.loss_value(), e2.loss_value(), e3.loss_value(), tg.evaluate(theta, c['prob'], c['X_u'], c['u'], c['X_f'], want_grad=False).loss)
