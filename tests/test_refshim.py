"""Known-answer checks of the TensorFlow-1 stand-in (oracle/refshim/tensorflow) that the reference-run fixtures rest on:
the ~30 symbols the reference scripts use, against closed forms.  CPU only; no reference needed."""
import numpy as np
import pytest
import torch

from oracle import run_reference as rr


@pytest.fixture
def tf():
    with rr.shimmed(compute="float64") as mod:
        yield mod


def test_nested_gradients_are_pointwise_derivatives(tf):
    """tf.gradients(y, x)[0] with y_i = g(x_i): the all-ones VJP is the pointwise derivative, and it nests
    (INF-L2:113-117 u_t, u_x, u_xx)."""
    x = tf.placeholder(tf.float32, shape=[None, 1])
    t = tf.placeholder(tf.float32, shape=[None, 1])
    u = tf.tanh(2.0 * x) * t + x ** 2
    u_x = tf.gradients(u, x)[0]
    u_xx = tf.gradients(u_x, x)[0]
    u_t = tf.gradients(u, t)[0]
    xs = np.float32(np.linspace(-1, 1, 7))[:, None].astype(np.float64)
    ts = np.float32(np.linspace(0.1, 0.9, 7))[:, None].astype(np.float64)
    sess = tf.Session()
    gx, gxx, gt = sess.run([u_x, u_xx, u_t], {x: xs, t: ts})
    th = np.tanh(2 * xs)
    assert np.allclose(gx, 2 * (1 - th ** 2) * ts + 2 * xs, rtol=1e-12)
    assert np.allclose(gxx, -8 * th * (1 - th ** 2) * ts + 2, rtol=1e-12)
    assert np.allclose(gt, th, rtol=1e-12)


def test_feeds_constants_and_python_scalars_round_to_float32_like_tf(tf):
    x = tf.placeholder(tf.float32, shape=[None, 1])
    y = (1 / 3) * x + np.array([0.1])                      # python scalar and float64 numpy constant -> float32 constants
    out = tf.Session().run(y, {x: np.array([[0.1]])})      # float64 feed -> float32 placeholder
    want = float(np.float32(1 / 3)) * float(np.float32(0.1)) + float(np.float32(0.1))
    assert abs(out[0, 0] - want) < 1e-15


def test_norms_where_and_reductions(tf):
    v = tf.constant(np.array([[3.0], [-4.0]]))
    sess = tf.Session()
    assert sess.run(tf.norm(v, 2)) == 5.0 and sess.run(tf.norm(v, 1)) == 7.0
    assert sess.run(tf.pow(tf.norm(v, 2), 2)) == 25.0 and sess.run(tf.reduce_mean(tf.square(v))) == 12.5
    w = tf.where(tf.greater(v, 0.0), tf.ones([2, 1]), tf.zeros([2, 1])) - tf.where(tf.less(v, 0.0), tf.ones([2, 1]), tf.zeros([2, 1]))
    assert np.array_equal(sess.run(w), [[1.0], [-1.0]])
    x = tf.Variable([0.0], dtype=tf.float32)
    sess.run(tf.global_variables_initializer())
    g = sess.run(tf.gradients(tf.norm(x, 2), x)[0])
    assert np.isnan(g).all()                               # tf.norm's gradient at zero


def test_variables_assign_and_the_rebinding_quirk(tf):
    """INF-ADMM:95,:106-107: `lagrange` re-bound to its own assign op; an op built from it advances the variable as a
    side effect of being evaluated, once per Session.run."""
    lag = tf.Variable(tf.ones([2, 1]), dtype=tf.float32, trainable=False)
    z = tf.Variable(tf.zeros([2, 1]), dtype=tf.float32, trainable=False)
    sess = tf.Session()
    sess.run(tf.global_variables_initializer())
    update_built_first = lag.assign(lag + 1.0)
    rebound = lag.assign(lag + 10.0)
    z_update = z.assign(rebound * 2.0 + rebound)            # uses the op twice: still ONE evaluation per run
    assert np.array_equal(sess.run(z_update), [[33.0], [33.0]])
    assert np.array_equal(sess.run(lag), [[11.0], [11.0]])
    sess.run(update_built_first)
    assert np.array_equal(sess.run(lag), [[12.0], [12.0]])
    assert tf.trainable_variables() == []


def test_adam_is_the_tf1_formula_with_epsilon_outside_the_bias_correction(tf):
    w = tf.Variable([1.0, -2.0], dtype=tf.float32)
    loss = tf.reduce_sum(tf.square(w) * np.array([1.0, 3.0]))
    train = tf.train.AdamOptimizer(learning_rate=0.001).minimize(loss)
    sess = tf.Session()
    sess.run(tf.global_variables_initializer())
    th = np.array([1.0, -2.0]); m = np.zeros(2); v = np.zeros(2)
    for t in range(1, 4):
        g = 2 * th * np.array([1.0, 3.0])
        m = 0.9 * m + 0.1 * g; v = 0.999 * v + 0.001 * g * g
        lr_t = 0.001 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)   # hyper-parameters stay unrounded in the float64 mode
        th = th - lr_t * m / (np.sqrt(v) + 1e-8)
        sess.run(train)
        assert np.allclose(sess.run(w), th, rtol=1e-9, atol=0)


def test_scipy_interface_packs_trainable_variables_in_creation_order(tf):
    a = tf.Variable([[1.0, 2.0]], dtype=tf.float32)
    frozen = tf.Variable([5.0], dtype=tf.float32, trainable=False)
    b = tf.Variable([3.0], dtype=tf.float32)
    loss = tf.reduce_sum(tf.square(a - 0.5)) + tf.reduce_sum(tf.square(b + frozen))
    opt = tf.contrib.opt.ScipyOptimizerInterface(loss, method='L-BFGS-B', options={'maxiter': 50})
    sess = tf.Session()
    sess.run(tf.global_variables_initializer())
    res = opt.minimize(sess)
    assert res.x.shape == (3,)                               # a (2) then b (1); the non-trainable variable is not packed
    assert np.allclose(sess.run(a), 0.5, atol=1e-6) and np.allclose(sess.run(b), -5.0, atol=1e-6)
    assert np.array_equal(sess.run(frozen), [5.0])


def test_float32_compute_mode_is_what_tf_does():
    with rr.shimmed(compute="float32") as tf:
        x = tf.placeholder(tf.float32, shape=[None, 1])
        y = tf.matmul(x, tf.constant(np.array([[1.0, 2.0]]))) + 1e-9
        out = tf.Session().run(y, {x: np.array([[1.0]])})
        assert out.dtype == np.float32 and np.array_equal(out, np.float32([[1.0, 2.0]]))   # 1e-9 is below float32 resolution


def test_truncated_normal_is_seeded_and_truncated(tf):
    tf.set_random_seed(1234)
    a = tf.Session().run(tf.truncated_normal([200, 50], stddev=0.3))
    tf.set_random_seed(1234)
    b = tf.Session().run(tf.truncated_normal([200, 50], stddev=0.3))
    assert np.array_equal(a, b) and np.abs(a).max() <= 0.6 + 1e-7 and abs(a.std() - 0.3 * 0.8796) < 0.01
