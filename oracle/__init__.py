"""CPU oracle for the DMVAE / VaDE / MoE training step.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement of the reference
algorithm (ffs97/deep-mixture-vae, TensorFlow 1.x graph) written op-for-op in
PyTorch-CPU / NumPy.  It exists so that the CUDA path can be checked against
it; it is never imported by the product package ``deep-mixture-vae_b200``.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.

PARITY UNPINNED: the reference ships no tests, golden vectors, fixtures or
seeds and TensorFlow 1.x cannot be imported in this
image (Python 3.12), so the oracle cannot be checked against outputs of the
reference itself.  It is pinned instead by (a) the analytic known-answer tests
derived from the reference formulas (tests/test_oracle_kat.py), (b) fp64
autograd versus the independently hand-derived closed-form backward
(oracle/closed_form.py), and (c) central finite differences.
"""
