"""Golden values for the analytic acquisition formulas, generated from the reference's own code.

The reference keeps numpy restatements of EI and UCB next to its BoTorch calls
(projects/swellex96_inv/visualization/bo_examples.py: ``expected_improvement`` :61-84, ``upper_confidence_bound``
:107-124).  The module cannot be imported here (botorch, tritonoa, matplotlib styles are absent) but those two functions
only need numpy and scipy.stats.norm: this script cuts them out of the reference file with ``ast``, executes them as
they are on a seeded set of (mu, sigma, best_f), and records the outputs in tests/golden/acq_reference.npz.
Run in the build container (needs /root/reference):   python tools/make_golden_acq.py
"""
import ast
from pathlib import Path

import numpy as np
from scipy.stats import norm

REF = Path("/root/reference/projects/swellex96_inv/visualization/bo_examples.py")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "acq_reference.npz"


def main():
    tree = ast.parse(REF.read_text())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("expected_improvement", "upper_confidence_bound")]
    assert len(wanted) == 2
    ns = {"np": np, "norm": norm, "print": lambda *a, **k: None}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), str(REF), "exec"), ns)
    rng = np.random.default_rng(0)
    mu = rng.normal(0.0, 1.0, 512)
    sigma = np.exp(rng.uniform(np.log(1e-4), np.log(2.0), 512))
    best_f = 0.7
    ei = ns["expected_improvement"](mu.copy(), sigma.copy(), best_f)
    ei_xi = ns["expected_improvement"](mu.copy(), sigma.copy(), best_f, 0.05)
    ucb = {f"ucb_kappa_{k}": ns["upper_confidence_bound"](mu.copy(), sigma.copy(), k) for k in (0.5, 1.0, 2.0)}
    np.savez_compressed(OUT, mu=mu, sigma=sigma, best_f=best_f, ei=ei, ei_xi005=ei_xi, **ucb)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
