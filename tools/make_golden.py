"""Generates tests/golden/*.npz.

The reference (NeptuneProjects/BOGP) cannot be imported here (TritonOA, BoTorch, GPyTorch, Ax are absent and
un-pinned) and holds no golden vectors, so the fixtures come from (a) INDEPENDENT installed
implementations -- scikit-learn's Matern-5/2 / RBF GaussianProcessRegressor and SciPy's normal cdf/pdf for the
in-repo EI/UCB formulas (ref/projects/swellex96_inv/visualization/bo_examples.py:78-84, :124) -- and (b) the float64
oracle itself for the Bartlett surfaces (regression pins; "parity unpinned" against TritonOA).

    python tools/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from bogp_b200.modes import INV_TONALS_3, REC_Z_21, SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance  # noqa: E402
from oracle import bartlett as ob  # noqa: E402

OUT = ROOT / "tests" / "golden"


def bartlett_fixture():
    modes = pekeris_modes(INV_TONALS_3, REC_Z_21)
    K = synthetic_covariance(modes, 71.11, 1.07)
    r = np.linspace(0.1, 8.0, 100)
    z = np.linspace(1.0, 200.0, 50)
    surf = ob.ambiguity_surface(modes, K, r, z)
    surf_ml = ob.ambiguity_surface(modes, K, r, z, atype="cbf_ml", multifreq_method="product")
    np.savez_compressed(OUT / "bartlett_cfg1.npz", r=r, z=z, cbf_mean=surf, cbf_ml_product=surf_ml,
                        argmax=np.array(np.unravel_index(np.argmax(surf), surf.shape)))
    # 13-tonal, 64-element, tilted, rank-8 covariance: scattered points
    modes13 = pekeris_modes(SWELLEX_TONALS_13)
    K8 = synthetic_covariance(modes13, 60.0, 5.0, tilt_deg=1.5, snapshots=8, seed=3)
    rng = np.random.default_rng(7)
    cand = np.stack([rng.uniform(0.5, 10.0, 96), rng.uniform(1.0, 200.0, 96), rng.uniform(-4.0, 4.0, 96)], axis=1)
    cand[:32, 2] = 0.0
    vals = ob.bartlett_points(modes13, K8, cand)
    np.savez_compressed(OUT / "bartlett_points13.npz", cand=cand, cbf_mean=vals)


def gp_fixture():
    from scipy.stats import norm
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern
    rng = np.random.default_rng(11)
    out = {}
    for name, d, n in (("m52_d2", 2, 48), ("m52_d7", 7, 96), ("rbf_d3", 3, 64)):
        X = rng.uniform(-1.0, 1.0, (n, d))
        y = np.sin(3.0 * X[:, 0]) + 0.5 * np.cos(2.0 * X.sum(1)) + 0.05 * rng.standard_normal(n)
        y = (y - y.mean()) / y.std()
        ls = rng.uniform(0.3, 0.9, d)
        s, noise = 1.3, 1e-4
        base = Matern(length_scale=ls, nu=2.5) if name.startswith("m52") else RBF(length_scale=ls)
        gpr = GaussianProcessRegressor(ConstantKernel(s, "fixed") * base, alpha=noise, optimizer=None).fit(X, y)
        Xs = rng.uniform(-1.0, 1.0, (200, d))
        mu, sd = gpr.predict(Xs, return_std=True)
        best = float(y.max())
        u = (mu - best) / sd
        ei = sd * (norm.pdf(u) + u * norm.cdf(u))          # bo_examples.py:78-84
        ucb = mu + np.sqrt(4.0) * sd                        # bo_examples.py:124 (beta = 4)
        out.update({f"{name}_X": X, f"{name}_y": y, f"{name}_ls": ls, f"{name}_s": s, f"{name}_noise": noise,
                    f"{name}_Xs": Xs, f"{name}_mu": mu, f"{name}_sd": sd, f"{name}_ei": ei, f"{name}_ucb": ucb,
                    f"{name}_pi": norm.cdf(u), f"{name}_best": best})
    np.savez_compressed(OUT / "gp_sklearn.npz", **out)


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    bartlett_fixture()
    gp_fixture()
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))
