"""CPU oracle for the PINN training hot path of jonwittmer/PINNs.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pinns_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU arm -- never as the product path.

PARITY PINNED TO THE REFERENCE'S OWN CODE, NOT TO TENSORFLOW: the reference ships
no golden vectors, known-answer tests or logged numbers for this path (SURVEY.md
section 8c) and TensorFlow 1.x is neither vendored nor installed.  The reference's
eight model scripts do run here, unmodified, over a TensorFlow-1 API stand-in
(``refshim``, driven by ``run_reference``); the fixtures tests/golden/ref_*.npz are
what those runs computed, and ``tf_graph`` / ``taylor`` / ``optim`` / ``data`` must
reproduce them (tests/test_reference_pin.py).  Unpinned remainder: the semantics of
the ~30 TensorFlow symbols restated in the stand-in.  The restatements are also
pinned against each other and central finite differences, and by the oracle-made
fixtures under ``tests/golden/`` (script committed there).

Modules
-------
tf_graph   torch (CPU) restatement of the reference TF-1 graph, fp64 or fp32
taylor     numpy fp64 restatement of SURVEY.md appendix A.2 (independent check)
optim      TF-1 Adam and the SciPy L-BFGS-B driver of ScipyOptimizerInterface
data       the reference drivers' data preparation (load_data / __main__ blocks)
philox     Philox4x32-10 counter-based sampler (Random123 known-answer vector)
refshim/   TensorFlow-1 / pyDOE / matplotlib stand-ins for running the reference scripts
run_reference  runs an unmodified reference script from /root/reference over refshim
"""
