"""TF-1 Adam and the SciPy L-BFGS-B driver, restated.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  UNPINNED BY TENSORFLOW: TF 1.x
(tf.train.AdamOptimizer, tf.contrib.opt.ScipyOptimizerInterface) is an
un-vendored, un-pinned dependency of the reference; the update rules below are
its documented algorithm, anchored on the call sites INF-L2:72-73 and
AB-ADMM:66-72,:216 (the TF-1 stand-in oracle/refshim restates the same rule, so
the reference-run fixtures pin only how the scripts drive it).
"""
from __future__ import annotations

import numpy as np
import scipy.optimize


class TF1Adam:
    """tf.train.AdamOptimizer(learning_rate=0.001) with TF-1 defaults.
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps): epsilon sits
    outside the bias correction, unlike torch.optim.Adam."""

    def __init__(self, n, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, dtype=np.float64):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.m = np.zeros(n, dtype)
        self.v = np.zeros(n, dtype)
        self.t = 0
        self.dtype = dtype

    def step(self, theta, grad):
        self.t += 1
        g = np.asarray(grad, self.dtype)
        lr_t = self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        self.m = self.b1 * self.m + (1.0 - self.b1) * g
        self.v = self.b2 * self.v + (1.0 - self.b2) * g * g
        return (np.asarray(theta, self.dtype) - self.dtype(lr_t) * self.m / (np.sqrt(self.v) + self.dtype(self.eps))).astype(self.dtype)


# options passed verbatim by the reference
LBFGS_OPTIONS_AB_ADMM = {'maxiter': 5000, 'maxfun': 50000, 'maxcor': 50, 'maxls': 50, 'ftol': 1e-7}          # AB-ADMM:68-72
LBFGS_OPTIONS_AB_L2 = {'maxiter': 50000, 'maxfun': 50000, 'maxcor': 50, 'maxls': 50,
                       'ftol': 1.0 * np.finfo(float).eps}                                                     # AB-L2:68-72


def lbfgs_minimize(loss_grad, theta0, options=None):
    """ScipyOptimizerInterface.minimize: variables packed in creation order into one
    float64 vector, scipy.optimize.minimize(method='L-BFGS-B', jac=True)."""
    options = dict(LBFGS_OPTIONS_AB_ADMM if options is None else options)

    def fun(x):
        loss, grad = loss_grad(x)
        return float(loss), np.asarray(grad, np.float64)

    res = scipy.optimize.minimize(fun, np.asarray(theta0, np.float64), jac=True, method='L-BFGS-B', options=options)
    return res.x, res
