import sys, zlib, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import tf_graph as tg
from tests.helpers import make_case, make_engine, rel_err, max_rel_err
B, E = tg.PDE_BURGERS, tg.PDE_EULER
for name,pde,n,nl,loss,n_f in [("tc-32",B,32,3,tg.LOSS_V4,300),("tc-128",B,128,3,tg.LOSS_V4,300),("tc-euler-64",E,64,3,tg.LOSS_EULER_MSE,333),("tc-200",B,200,3,tg.LOSS_V4,200),("tc-euler-200",E,200,3,tg.LOSS_V6,300)]:
    layers=[2]+[n]*nl+[1 if pde==B else 3]
    case=make_case(pde,layers,loss,50,n_f,seed=5)
    ref=tg.evaluate(case["theta"],case["prob"],case["X_u"],case["u"],case["X_f"],case["z"],case["gamma"])
    try:
        eng=make_engine(case,path="tensor",trainable_lambda=(pde==B))
        y,f=eng.predict(case["X_f"]); yr,fr=tg.predict(case["theta"],case["prob"],case["X_f"])
        print(name,"predict: y max-rel %.2e f max-rel %.2e"%(max_rel_err(y,yr),max_rel_err(f,fr)),flush=True)
        l,g=eng.loss_grad(); P=eng.num_params
        print(name,"loss rel %.2e grad rel %.2e"%(abs(l-ref.loss)/abs(ref.loss),rel_err(g[:P],ref.grad)),flush=True)
        # per-layer gradient errors
        off=0
        for li in range(len(layers)-1):
            nw=layers[li]*layers[li+1]; nb=layers[li+1]
            print("   layer %d W err %.2e  b err %.2e"%(li,rel_err(g[off:off+nw],ref.grad[off:off+nw]),rel_err(g[off+nw:off+nw+nb],ref.grad[off+nw:off+nw+nb])))
            off+=nw+nb
    except Exception as ex:
        print(name,"EXC",ex,flush=True)
