"""CPU oracle for the PINN training hot path of jonwittmer/PINNs.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pinns_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU arm -- never as the product path.

PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or
logged numbers for this path, and its arithmetic lives in
TensorFlow 1.x, which is neither vendored under /root/reference nor installed
here.  The oracle is therefore pinned only (a) against itself through two
independent restatements (``tf_graph``: reverse-over-reverse autograd that
mirrors the reference graph op for op; ``taylor``: hand-derived Taylor-forward
plus one reverse sweep in numpy) and central finite differences, and (b) by the
fixtures under ``tests/golden/`` generated from it (script committed there).

Modules
-------
tf_graph   torch (CPU) restatement of the reference TF-1 graph, fp64 or fp32
taylor     numpy fp64 restatement of SURVEY.md appendix A.2 (independent check)
optim      TF-1 Adam and the SciPy L-BFGS-B driver of ScipyOptimizerInterface
data       the reference drivers' data preparation (load_data / __main__ blocks)
"""
