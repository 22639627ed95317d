"""Runs the reference's own, UNMODIFIED model scripts from /root/reference in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  TensorFlow, pyDOE and matplotlib are not installed, so the
scripts are executed with `oracle/refshim/` first on sys.path: a TF-1 graph-mode API over torch autograd, pyDOE's
`lhs`, and inert plotting modules.  Everything else -- the class bodies, the drivers' data preparation, NumPy's
legacy RNG stream, SciPy, pandas -- is the real thing.  /root/reference is read-only and the scripts use paths
relative to their own directory ('../Data/...', 'figures/...'), so each run happens in a scratch mirror: the
script's directory re-created under a temp dir, `Data` symlinked, the `figures/` tree re-created empty.

Only tests/golden/make_ref_fixtures.py and the live check in tests/test_reference_pin.py call this; the GPU box
has no /root/reference and uses the committed fixtures instead.
"""
from __future__ import annotations

import contextlib
import io
import os
import re
import runpy
import sys
import tempfile

REF = "/root/reference"
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPTS = {
    "INF-L2": "Burgers/continuous_inference/Hwan_L2Regularization_Burgers.py",
    "INF-ADMM": "Burgers/continuous_inference/Hwan_L1Regularization_ADMM_Burgers.py",
    "ID-L2b": "Burgers/continuous_identification/Burgers_batch_L2.py",
    "ID-ADMMb": "Burgers/continuous_identification/Burgers_ADMM_batch.py",
    "AB-ADMM": "Burgers/continuous_identification/Abgrall_ADMM.py",
    "AB-L2": "Burgers/continuous_identification/Abgrall_L2.py",
    "AB-L1": "Burgers/continuous_identification/Abgrall_L1.py",
    "EUL": "Eulers/continuous_inference/Euler_ADMM.py",
}


def available() -> bool:
    return os.path.isdir(REF)


@contextlib.contextmanager
def shimmed(compute="float32", round_scalars=True):
    """sys.path / sys.modules set up so that `import tensorflow` resolves to the shim."""
    import torch
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules)
                  if k.split(".")[0] in ("tensorflow", "matplotlib", "mpl_toolkits", "pyDOE", "_null")}
    sys.path[:0] = [SHIM, ROOT]
    try:
        import tensorflow as tf
        tf.reset_default_graph()
        tf.set_compute_dtype(torch.float64 if compute == "float64" else torch.float32, round_scalars)
        yield tf
    finally:
        for k in list(sys.modules):
            if k.split(".")[0] in ("tensorflow", "matplotlib", "mpl_toolkits", "pyDOE", "_null"):
                del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path


def run_script(name, argv=(), run_name="__main__", compute="float32", round_scalars=True, quiet=True, hook=None):
    """Execute reference script `name` (key of SCRIPTS) under the shim.

    run_name="__main__" runs the driver block as the reference would (argv = its positional arguments);
    any other run_name only defines the module (classes) without running the driver.
    hook(tf): called after the shim is importable and before the script runs (e.g. to wrap nothing -- kept for
    fixtures that need the tf module object).  Returns (globals dict of the executed script, tf module).
    The scratch directory is left in place for the lifetime of the returned objects (CSV dumps land there).
    """
    rel = SCRIPTS[name]
    src = os.path.join(REF, rel)
    pkg, sub = rel.split("/")[0], rel.split("/")[1]
    scratch = tempfile.mkdtemp(prefix="refrun_")
    cwd = os.path.join(scratch, pkg, sub)
    os.makedirs(cwd)
    os.symlink(os.path.join(REF, pkg, "Data"), os.path.join(scratch, pkg, "Data"))
    figs = os.path.join(REF, pkg, sub, "figures")
    for d, _, _ in os.walk(figs):
        os.makedirs(os.path.join(cwd, os.path.relpath(d, os.path.join(REF, pkg, sub))), exist_ok=True)
    os.makedirs(os.path.join(cwd, "figures"), exist_ok=True)
    # output directories the script names but the repository does not contain (a user would mkdir them)
    with open(src) as fh:
        for d in set(re.findall(r"figures/(?:[A-Za-z0-9_]+/)+", fh.read())):
            os.makedirs(os.path.join(cwd, d), exist_ok=True)
    old_cwd, old_argv = os.getcwd(), list(sys.argv)
    with shimmed(compute, round_scalars) as tf:
        if hook is not None:
            hook(tf)
        os.chdir(cwd)
        sys.argv = [src] + [str(a) for a in argv]
        try:
            with contextlib.redirect_stdout(io.StringIO() if quiet else sys.stdout):
                g = runpy.run_path(src, run_name=run_name)
        finally:
            os.chdir(old_cwd)
            sys.argv = old_argv
        g["__scratch__"] = cwd
        return g, tf


def flat_params(sess, weights, biases):
    """W1,b1,...,WL,bL flattened row-major: the layout of SURVEY.md section 8(a1)."""
    import numpy as np
    vals = sess.run([v for pair in zip(weights, biases) for v in pair])
    return np.concatenate([np.asarray(v, np.float64).ravel() for v in vals])


def initial_flat_params(weights, biases):
    """The values the variables were initialised with (before any training inside the constructor)."""
    import numpy as np
    return np.concatenate([v._initial.double().numpy().ravel() for pair in zip(weights, biases) for v in pair])


def shim_initial_theta(layers, seed=1234):
    """The parameters a reference script starts from under the shim: tf.set_random_seed(1234) followed by one
    tf.truncated_normal per layer in creation order (INF-L2:79-94), biases zero -- regenerated without the reference."""
    import numpy as np
    sys.path.insert(0, SHIM)
    try:
        saved = sys.modules.pop("tensorflow", None)
        import tensorflow as tf
        tf.set_random_seed(seed)
        parts = []
        for l in range(len(layers) - 1):
            std = np.sqrt(2 / (layers[l] + layers[l + 1]))
            w = tf.truncated_normal([layers[l], layers[l + 1]], stddev=std)._eval({})
            parts += [w.double().numpy().ravel(), np.zeros(layers[l + 1])]
        return np.concatenate(parts).astype(np.float32)
    finally:
        sys.modules.pop("tensorflow", None)
        if saved is not None:
            sys.modules["tensorflow"] = saved
        sys.path.remove(SHIM)
