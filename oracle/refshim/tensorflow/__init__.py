"""Minimal TensorFlow-1 graph-mode API over torch (CPU) autograd -- a STAND-IN so that the
reference's own, unmodified model scripts under /root/reference can be imported and run
in the build container, where TensorFlow is not installed.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by oracle/run_reference.py and
tests/golden/make_ref_fixtures.py to produce fixtures from the reference's graph-building
code; it never travels on the product path.

What this pins and what it does not: every line of the reference classes (graph
construction, loss formulas, ADMM assign ops incl. their data-flow quirks, train loops,
feed dictionaries, NumPy legacy RNG draws, predict) executes as written.  The semantics
of the ~30 TensorFlow symbols the scripts use (listed in __all__) are restated here from
the TF-1 documentation: they are the un-pinned remainder.

Graph model: a `Tensor` is a symbolic node; `Session.run` evaluates nodes with a per-run
memo, feeding placeholders and variables as autograd leaves so that `tf.gradients` is
`torch.autograd.grad(y, x, ones_like(y), create_graph=True)` (the same all-ones VJP) and
nests like the reference's reverse-over-reverse graph.  float32 placeholders / variables /
constants are rounded to float32 exactly like TF's feed-time and constant conversion and
then computed in the shim's compute dtype: float32 (default, what TF does) or float64
(`set_compute_dtype`) for comparisons that should not carry fp32 evaluation noise.
"""
from __future__ import annotations

import numpy as _np
import torch as _torch

__version__ = "1.15-shim"
__all__ = ["float32", "float64", "placeholder", "Variable", "constant", "zeros", "ones", "truncated_normal",
           "add", "matmul", "tanh", "concat", "gradients", "norm", "pow", "square", "abs", "exp", "div",
           "transpose", "where", "greater", "less", "reduce_mean", "reduce_sum", "set_random_seed",
           "global_variables_initializer", "trainable_variables", "Session", "GPUOptions", "ConfigProto",
           "train", "contrib"]

float32 = "float32"
float64 = "float64"


class _State:
    compute = _torch.float32
    round_scalars = True          # python scalars / numpy constants -> float32 first (TF's conversion)
    rng = _np.random.RandomState(0)
    variables = []                # creation order == tf.global_variables()


def set_compute_dtype(dtype, round_scalars=True):
    """Shim control (not TF API): dtype the graph is evaluated in, and whether python/numpy
    constants are rounded to float32 first like TF's constant conversion does."""
    _State.compute = dtype
    _State.round_scalars = round_scalars


def reset_default_graph():
    _State.variables = []


def set_random_seed(seed):
    _State.rng = _np.random.RandomState(seed)


def _to_compute(a):
    a = _np.asarray(a, dtype=_np.float64)
    if _State.round_scalars:
        a = a.astype(_np.float32).astype(_np.float64)
    return _torch.from_numpy(_np.ascontiguousarray(a)).to(_State.compute)


class Tensor:
    """Symbolic graph node."""
    __array_ufunc__ = None        # numpy_array (op) Tensor -> Tensor.__r(op)__

    def __init__(self, fn, inputs=(), name="op"):
        self._fn, self._inputs, self.name = fn, tuple(inputs), name

    def _eval(self, env):
        key = id(self)
        if key not in env:
            env[key] = self._fn(*[i._eval(env) for i in self._inputs])
        return env[key]

    # operator overloads (tf.Tensor.OVERLOADABLE_OPERATORS)
    def __add__(self, o): return _binary(_torch.add, self, o, "add")
    def __radd__(self, o): return _binary(_torch.add, o, self, "add")
    def __sub__(self, o): return _binary(_torch.sub, self, o, "sub")
    def __rsub__(self, o): return _binary(_torch.sub, o, self, "sub")
    def __mul__(self, o): return _binary(_torch.mul, self, o, "mul")
    def __rmul__(self, o): return _binary(_torch.mul, o, self, "mul")
    def __truediv__(self, o): return _binary(_torch.div, self, o, "truediv")
    def __rtruediv__(self, o): return _binary(_torch.div, o, self, "truediv")
    def __pow__(self, o): return _binary(_torch.pow, self, o, "pow")
    def __neg__(self): return Tensor(_torch.neg, [self], "neg")

    def __getitem__(self, idx):
        return Tensor(lambda v: v[idx], [self], "strided_slice")

    def __bool__(self):
        raise TypeError("using a tf.Tensor as a Python bool is not allowed")

    __hash__ = object.__hash__

    def __eq__(self, o):
        return self is o


def convert_to_tensor(v):
    if isinstance(v, Tensor):
        return v
    src = _np.array(v, dtype=_np.float64)      # captured at graph-construction time, like a tf constant
    return Tensor(lambda: _to_compute(src), [], "Const")


def _binary(fn, a, b, name):
    return Tensor(fn, [convert_to_tensor(a), convert_to_tensor(b)], name)


class _Placeholder(Tensor):
    def __init__(self, dtype, shape):
        super().__init__(None, [], "Placeholder")
        self.dtype, self.shape = dtype, shape

    def _eval(self, env):
        try:
            return env[id(self)]
        except KeyError:
            raise ValueError("You must feed a value for placeholder tensor") from None


def placeholder(dtype, shape=None, name=None):
    return _Placeholder(dtype, shape)


class Operation(Tensor):
    """Node run for its side effect (assign, train op, initializer)."""


class Variable(Tensor):
    def __init__(self, initial_value, dtype=None, trainable=True, name=None):
        super().__init__(None, [], "Variable")
        if isinstance(initial_value, Tensor):
            init = initial_value._eval({})
        else:
            init = _to_compute(initial_value)
        self._initial = init.detach().clone()
        self._value = None            # set by the initializer, as in TF
        self.trainable = trainable
        self.dtype = dtype or float32
        _State.variables.append(self)

    def _eval(self, env):
        key = id(self)
        if key not in env:
            if self._value is None:
                raise RuntimeError("Attempting to use uninitialized value " + self.name)
            env[key] = self._value.to(_State.compute).detach().requires_grad_(True)
        return env[key]

    def assign(self, value):
        value = convert_to_tensor(value)

        def do(v):
            self._value = v.detach().clone()
            return self._value

        return Operation(do, [value], "Assign")

    def load(self, value, session=None):
        self._value = _to_compute(value).reshape(self._initial.shape)

    def get_shape(self):
        return tuple(self._initial.shape)


def global_variables():
    return list(_State.variables)


def trainable_variables():
    return [v for v in _State.variables if v.trainable]


def global_variables_initializer():
    todo = list(_State.variables)

    def do():
        for v in todo:
            v._value = v._initial.to(_State.compute).clone()
        return None

    return Operation(do, [], "init")


def constant(value, dtype=None, shape=None, name=None):
    return convert_to_tensor(value)


def zeros(shape, dtype=float32):
    shape = tuple(int(s) for s in shape)
    return Tensor(lambda: _torch.zeros(shape, dtype=_State.compute), [], "zeros")


def ones(shape, dtype=float32):
    shape = tuple(int(s) for s in shape)
    return Tensor(lambda: _torch.ones(shape, dtype=_State.compute), [], "ones")


def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None):
    """N(mean, stddev^2) with draws beyond two standard deviations re-drawn (TF-1 docs).
    TF's own random stream cannot be reproduced without TF; the draw comes from the shim's
    generator (tf.set_random_seed), one draw per node at graph-construction time."""
    shape = tuple(int(s) for s in shape)
    w = _State.rng.standard_normal(shape)
    bad = _np.abs(w) > 2.0
    while bad.any():
        w[bad] = _State.rng.standard_normal(int(bad.sum()))
        bad = _np.abs(w) > 2.0
    sample = (mean + stddev * w).astype(_np.float32)
    return Tensor(lambda: _to_compute(sample), [], "truncated_normal")


def add(x, y, name=None): return _binary(_torch.add, x, y, "Add")
def div(x, y, name=None): return _binary(_torch.div, x, y, "Div")
def pow(x, y, name=None): return _binary(_torch.pow, x, y, "Pow")  # noqa: A001
def matmul(a, b, name=None): return _binary(_torch.matmul, a, b, "MatMul")
def tanh(x, name=None): return Tensor(_torch.tanh, [convert_to_tensor(x)], "Tanh")
def exp(x, name=None): return Tensor(_torch.exp, [convert_to_tensor(x)], "Exp")
def abs(x, name=None): return Tensor(_torch.abs, [convert_to_tensor(x)], "Abs")  # noqa: A001
def square(x, name=None): return Tensor(lambda v: v * v, [convert_to_tensor(x)], "Square")
def transpose(a, perm=None, name=None): return Tensor(lambda v: v.t(), [convert_to_tensor(a)], "Transpose")
def greater(x, y, name=None): return _binary(_torch.gt, x, y, "Greater")
def less(x, y, name=None): return _binary(_torch.lt, x, y, "Less")
def reduce_mean(x, axis=None): return Tensor(_torch.mean, [convert_to_tensor(x)], "Mean")
def reduce_sum(x, axis=None): return Tensor(_torch.sum, [convert_to_tensor(x)], "Sum")


def where(condition, x=None, y=None, name=None):
    return Tensor(_torch.where, [condition, convert_to_tensor(x), convert_to_tensor(y)], "Select")


def concat(values, axis, name=None):
    values = [convert_to_tensor(v) for v in values]
    return Tensor(lambda *vs: _torch.cat(vs, dim=axis), values, "ConcatV2")


def norm(tensor, ord="euclidean", axis=None, name=None):  # noqa: A002
    """tf.norm with axis=None: the vector norm of the flattened tensor; ord 2 / 'euclidean'
    = sqrt(sum(x*x)) (gradient x/norm, NaN at x == 0 like TF's Sqrt gradient), ord 1 = sum|x|."""
    t = convert_to_tensor(tensor)
    if ord in (2, "euclidean", 2.0):
        return Tensor(lambda v: _torch.sqrt(_torch.sum(v * v)), [t], "norm2")
    if ord in (1, 1.0):
        return Tensor(lambda v: _torch.sum(_torch.abs(v)), [t], "norm1")
    raise NotImplementedError("tf.norm ord=%r" % (ord,))


def gradients(ys, xs, grad_ys=None, name=None):
    """tf.gradients: d sum(ys) / d x for each x (symbolic; evaluates lazily inside Session.run)."""
    single = not isinstance(xs, (list, tuple))
    xs = [xs] if single else list(xs)
    ys = convert_to_tensor(ys)

    def make(x):
        def do(y, xv):
            g = _torch.autograd.grad(y, xv, grad_outputs=_torch.ones_like(y), create_graph=True,
                                     allow_unused=True)[0]
            return g
        return Tensor(do, [ys, x], "gradients")

    return [make(x) for x in xs]


class Session:
    def __init__(self, target="", graph=None, config=None):
        self.config = config

    def run(self, fetches, feed_dict=None):
        env = {}
        for ph, val in (feed_dict or {}).items():
            if isinstance(ph, _Placeholder):
                # feed-time cast to the placeholder dtype (float32), then the compute dtype
                a = _np.asarray(val, dtype=_np.float64).astype(_np.float32).astype(_np.float64)
                env[id(ph)] = _torch.from_numpy(_np.ascontiguousarray(a)).to(_State.compute).requires_grad_(True)
            else:
                raise TypeError("only placeholders can be fed in the shim")

        def out(v):
            if v is None:
                return None
            return v.detach().cpu().numpy().astype(_np.float32 if _State.compute == _torch.float32 else _np.float64)

        def fetch(f):
            if isinstance(f, (list, tuple)):
                return [fetch(g) for g in f]
            r = f._eval(env)
            if isinstance(f, Operation) and f.name != "Assign":
                return None
            return out(r)

        return fetch(fetches)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class GPUOptions:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class ConfigProto:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _AdamOptimizer:
    """tf.train.AdamOptimizer (TF-1 ApplyAdam, restated from the TF-1 docs):
    lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
    var -= lr_t * m / (sqrt(v) + eps)   -- eps OUTSIDE the bias correction."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, use_locking=False, name="Adam"):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon

    def minimize(self, loss, global_step=None, var_list=None):
        var_list = list(var_list) if var_list is not None else trainable_variables()
        grads = gradients(loss, var_list)
        slots = {}
        state = {"b1p": None, "b2p": None}

        def do(*gs):
            dt = _State.compute
            if state["b1p"] is None:
                state["b1p"] = _torch.tensor(self.b1, dtype=dt)
                state["b2p"] = _torch.tensor(self.b2, dtype=dt)
            lr = _torch.tensor(self.lr, dtype=dt)
            b1 = _torch.tensor(self.b1, dtype=dt)
            b2 = _torch.tensor(self.b2, dtype=dt)
            eps = _torch.tensor(self.eps, dtype=dt)
            lr_t = lr * _torch.sqrt(1 - state["b2p"]) / (1 - state["b1p"])
            for v, g in zip(var_list, gs):
                if g is None:
                    continue
                g = g.detach()
                if id(v) not in slots:
                    slots[id(v)] = [_torch.zeros_like(g), _torch.zeros_like(g)]
                m, s = slots[id(v)]
                m = m + (g - m) * (1 - b1)
                s = s + (g * g - s) * (1 - b2)
                slots[id(v)] = [m, s]
                v._value = (v._value.to(dt) - lr_t * m / (_torch.sqrt(s) + eps)).detach()
            state["b1p"] = state["b1p"] * b1
            state["b2p"] = state["b2p"] * b2
            return None

        return Operation(do, grads, "Adam")


class _ScipyOptimizerInterface:
    """tf.contrib.opt.ScipyOptimizerInterface: trainable variables packed in creation order into
    one float64 vector, scipy.optimize.minimize(jac=True, method, options), result written back."""

    def __init__(self, loss, var_list=None, method="L-BFGS-B", options=None, **kw):
        self.loss = loss
        self.vars = list(var_list) if var_list is not None else trainable_variables()
        self.grads = gradients(loss, self.vars)
        self.method, self.options = method, dict(options or {})

    def minimize(self, session, feed_dict=None, fetches=None, step_callback=None, loss_callback=None, **kw):
        import scipy.optimize

        shapes = [tuple(v._initial.shape) for v in self.vars]
        sizes = [int(_np.prod(s)) for s in shapes]

        def unpack(x):
            off = 0
            for v, s, n in zip(self.vars, shapes, sizes):
                v.load(x[off:off + n].reshape(s))
                off += n

        def fun(x):
            unpack(x)
            vals = session.run([self.loss] + list(self.grads), feed_dict)
            g = _np.concatenate([_np.asarray(a, _np.float64).ravel() for a in vals[1:]])
            if loss_callback is not None and fetches:
                loss_callback(*session.run(fetches, feed_dict))
            return float(vals[0]), g

        x0 = _np.concatenate([v._value.detach().double().numpy().ravel() for v in self.vars])
        res = scipy.optimize.minimize(fun, x0, jac=True, method=self.method, options=self.options,
                                      callback=step_callback)
        unpack(res.x)
        return res


class _Namespace:
    pass


train = _Namespace()
train.AdamOptimizer = _AdamOptimizer
contrib = _Namespace()
contrib.opt = _Namespace()
contrib.opt.ScipyOptimizerInterface = _ScipyOptimizerInterface
