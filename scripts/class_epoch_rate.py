"""Wall-clock epochs/s of the Dialect-B drop-in classes at the reference's own sizes (host loop included)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pinns_b200.models import BurgersIdentification, EulerInference, EulerParameters, Parameters
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "data")

def rate(make, n):
    m = make()
    m.train(50)                      # warm-up
    m.engine.synchronize()
    t = time.perf_counter()
    m.train(n + 1)
    m.engine.synchronize()
    dt = time.perf_counter() - t
    return n / dt, dt / n * 1e6

class P(Parameters):
    N_u = 100; N_f = 1000; rho = 10.0; epochs = 1; gpu = '0'
class E(EulerParameters):
    N_data = 200; N_f = 1000; pen = 40.0; epochs = 1; gpu = '0'
for resample in ("host", "device"):
    r, us = rate(lambda: BurgersIdentification(P(), variant="AB-ADMM", data=os.path.join(GOLD, "TwoSin_burgers_shock.npz"), run=False, verbose=False, resample=resample), 3000)
    print("AB-ADMM  [2,20x8,1]  N_f=1000 resample=%-6s %8.0f epochs/s  %6.1f us/epoch" % (resample, r, us))
    r, us = rate(lambda: EulerInference(E(), data=os.path.join(GOLD, "Abgrall_eulers.npz"), run=False, verbose=False, resample=resample), 1000)
    print("EUL      [2,200x5,3] N_f=1000 resample=%-6s %8.0f epochs/s  %6.1f us/epoch" % (resample, r, us))
