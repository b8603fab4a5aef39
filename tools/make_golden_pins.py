"""Golden values produced by executing the reference's OWN source for three in-repo pieces of the hot path's set-up:

* ``simulate_covariance`` (optimization/utils.py:23-32) driven through ``functools.partial(run_kraken)`` -- the runner the
  unchanged configs build with ``pbuilds(run_kraken)`` (projects/swellex96_localization/conf/optimization/run.py:89-99);
* ``compute_covariance_worker`` (projects/swellex96_localization/data/acoustics.py:116-124);
* BAxUS ``embedding_matrix`` / ``increase_embedding_and_observations`` (projects/swellex96_inv/data/bo/baxus.py:156-235).

The modules cannot be imported here (hydra, tritonoa, oao, botorch are absent); the function definitions are cut out of
the reference files with ``ast`` and executed unchanged, with ``tritonoa``'s two names supplied by the shims of this
repo.  Run in the build container (needs /root/reference):   python tools/make_golden_pins.py
"""
import ast
import functools
import sys
import tempfile
from pathlib import Path
from typing import Optional, Union

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden" / "pins_r2.npz"


def cut(path: Path, names, ns):
    tree = ast.parse(path.read_text())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(wanted) == len(names), (path, names)
    exec(compile(ast.Module(body=wanted, type_ignores=[]), str(path), "exec"), ns)
    return ns


def main():
    from bogp_b200 import mfp
    from bogp_b200.modes import INV_TONALS_3, REC_Z_21, pekeris_modes
    out = {}
    # ---- simulate_covariance, the reference's loop over runner(parameters | {"freq": f})
    ns = cut(REF / "optimization" / "utils.py", ["simulate_covariance"],
             {"np": np, "Union": Union, "Optional": Optional, "covariance": mfp.covariance})
    mfp.set_mode_provider(pekeris_modes(INV_TONALS_3, REC_Z_21))
    runner = functools.partial(mfp.run_kraken)          # what instantiate(pbuilds(run_kraken)) yields
    for tag, params in (("vla", {"src_z": 71.11, "rec_r": 1.07}), ("tilt", {"src_z": 60.0, "rec_r": 3.2, "tilt": 1.5})):
        out[f"simcov_{tag}"] = ns["simulate_covariance"](runner, dict(params), list(INV_TONALS_3))
    mfp.set_mode_provider(None)
    # ---- compute_covariance_worker on a seeded data.npy
    import os
    ns = cut(REF / "projects" / "swellex96_localization" / "data" / "acoustics.py", ["compute_covariance_worker"],
             {"np": np, "os": os, "covariance": mfp.covariance})
    rng = np.random.default_rng(7)
    data = rng.standard_normal((6, 21)) + 1j * rng.standard_normal((6, 21))
    with tempfile.TemporaryDirectory() as td:
        np.save(Path(td) / "data.npy", data)
        ns["compute_covariance_worker"](Path(td), 6)
        out["snap_data"] = data
        out["snap_K"] = np.load(Path(td) / "covariance.npy")
    # ---- BAxUS embedding helpers on fixed torch seeds
    ns = cut(REF / "projects" / "swellex96_inv" / "data" / "bo" / "baxus.py",
             ["embedding_matrix", "increase_embedding_and_observations"], {"torch": torch})
    cpu = torch.device("cpu")
    for i, (seed, D, d) in enumerate([(0, 9, 2), (1, 9, 3), (2, 12, 5), (3, 7, 7), (4, 20, 4)]):
        torch.manual_seed(seed)
        S = ns["embedding_matrix"](D, d, torch.float64, cpu)
        out[f"emb_{i}_args"] = np.array([seed, D, d])
        out[f"emb_{i}_S"] = S.numpy()
        if d < D and all(int((S[r] != 0).sum()) > 1 for r in range(S.shape[0])):
            X = torch.rand(5, d, dtype=torch.float64) * 2 - 1
            S2, X2 = ns["increase_embedding_and_observations"](S, X, 3, torch.float64, cpu)
            out[f"emb_{i}_X"] = X.numpy()
            out[f"emb_{i}_S2"] = S2.numpy()
            out[f"emb_{i}_X2"] = X2.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, sorted(out))


if __name__ == "__main__":
    main()
