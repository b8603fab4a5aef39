"""CPU oracle for the BOGP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

float64 / complex128 restatement (NumPy + torch-CPU) of the arithmetic that the
reference delegates to TritonOA, BoTorch and GPyTorch.

PARITY UNPINNED: the reference repository holds no tests, golden vectors or
known-answer fixtures for this path, and none of the third-party packages that
own the arithmetic (TritonOA HEAD, OAOptimization HEAD, botorch/gpytorch via
ax-platform -- all un-pinned in ref/gp_dev.yml:19-26) can be imported in the
build container.  The oracle therefore follows (i) the formulas fixed by
BASELINE.json::north_star, (ii) the calling conventions evidenced at the
reference's call sites (cited per function), (iii) the in-repo formula
restatements (ref/projects/swellex96_inv/visualization/bo_examples.py:61-124),
and is cross-checked against independent implementations that *are* installed
(scikit-learn Matern-5/2 GP, SciPy normal cdf/pdf) plus the simulation-mode
invariants of SURVEY 8(c).  Pinned by executing the reference's own source: the
EI / UCB formulas (tools/make_golden_acq.py -> tests/golden/acq_reference.npz)
and, for the host logic, BaxusState / update_state (tools/make_golden_baxus.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  Nothing under
``bogp_b200/`` imports it; the product path fails loudly without its CUDA
extension.
"""
