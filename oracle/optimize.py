"""Oracle: ``optimize_acqf`` orchestration and the hand-rolled BO loop.

Test infrastructure only (see oracle/__init__.py).  Restates SURVEY 8a a8 --
``botorch.optim.optimize_acqf(acq, bounds, q, num_restarts, raw_samples)`` as called at
ref/projects/swellex96_inv/data/bo/ei.py:81-92 -- and the loop around it (ei.py:31-115).
BoTorch's internals are RECALLED: scrambled-Sobol raw samples seeded from
the torch RNG, Boltzmann restart selection (``initialize_q_batch`` /
``initialize_q_batch_nonneg``, ``torch.multinomial``), one joint SciPy L-BFGS-B over all
restarts on ``-sum(acq)``, best restart returned.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np
import torch
from scipy.optimize import minimize

from . import gp as ogp

DT = torch.float64
NONNEG_ACQ = ("ei", "pi", "qei")


def draw_sobol_samples(bounds: torch.Tensor, n: int, q: int, seed: int) -> torch.Tensor:
    d = bounds.shape[-1]
    eng = torch.quasirandom.SobolEngine(dimension=q * d, scramble=True, seed=seed)
    u = eng.draw(n, dtype=DT).reshape(n, q, d)
    return bounds[0] + (bounds[1] - bounds[0]) * u


def initialize_q_batch(X: torch.Tensor, Y: torch.Tensor, n: int, eta: float = 2.0) -> torch.Tensor:
    if n >= X.shape[0]:
        return X
    std = Y.std()
    if float(std) == 0.0:
        return X[torch.randperm(X.shape[0])[:n]]
    max_idx = int(torch.argmax(Y))
    etaZ = eta * (Y - Y.mean()) / std
    w = torch.exp(etaZ)
    while bool(torch.isinf(w).any()):
        etaZ = etaZ * 0.5
        w = torch.exp(etaZ)
    idcs = torch.multinomial(w, n)
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs]


def initialize_q_batch_nonneg(X: torch.Tensor, Y: torch.Tensor, n: int, eta: float = 1.0,
                              alpha: float = 1e-4) -> torch.Tensor:
    if n >= X.shape[0]:
        return X
    max_val, max_idx = float(Y.max()), int(torch.argmax(Y))
    if max_val <= 0.0:
        return X[torch.randperm(X.shape[0])[:n]]
    pos = Y >= alpha * max_val
    while int(pos.sum()) < n:
        alpha *= 0.1
        pos = Y >= alpha * max_val
    pos_idcs = torch.nonzero(pos).squeeze(-1)
    w = torch.exp(eta * (Y[pos] / max_val - 1.0))
    idcs = pos_idcs[torch.multinomial(w, n)]
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs]


def optimize_acqf(value_fn: Callable[[torch.Tensor], torch.Tensor],
                  value_and_grad_fn: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                  bounds: torch.Tensor, q: int, num_restarts: int, raw_samples: int, *,
                  nonneg: bool, maxiter: int = 200) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns ``(candidates[q, d], acq_value)``."""
    bounds = torch.as_tensor(bounds, dtype=DT)
    d = bounds.shape[-1]
    seed = int(torch.randint(0, 10000, (1,)).item())
    X_raw = draw_sobol_samples(bounds, raw_samples, q, seed)
    Y_raw = value_fn(X_raw)
    init = initialize_q_batch_nonneg if nonneg else initialize_q_batch
    X0 = init(X_raw, Y_raw, num_restarts)
    shape = X0.shape
    lo = bounds[0].repeat(shape[0] * q).numpy()
    hi = bounds[1].repeat(shape[0] * q).numpy()

    def f_and_g(x: np.ndarray):
        X = torch.from_numpy(x.copy()).reshape(shape)
        v, g = value_and_grad_fn(X)
        return -float(v.sum()), -g.reshape(-1).numpy().astype(np.float64)

    x0 = np.clip(X0.reshape(-1).numpy().astype(np.float64), lo, hi)
    res = minimize(f_and_g, x0, jac=True, method="L-BFGS-B", bounds=list(zip(lo, hi)),
                   options={"maxiter": maxiter})
    Xc = torch.from_numpy(np.clip(res.x, lo, hi)).reshape(shape)
    vals = value_fn(Xc)
    best = int(torch.argmax(vals))
    return Xc[best], vals[best]


def bo_loop(objective: Callable[[np.ndarray], np.ndarray], dim: int, budget: int, n_init: int,
            num_restarts: int, raw_samples: int, seed: int = 0, acq: str = "ei", beta: float = 4.0,
            fit_cfg: ogp.FitConfig = ogp.FitConfig(), q: int = 1, mc_samples: int = 512):
    """ei.py:31-115 restated: Sobol init in [-1,1]^d, maximise ``Y = -objective``."""
    torch.manual_seed(seed)
    sobol = torch.quasirandom.SobolEngine(dimension=dim, scramble=True, seed=seed)
    X = 2.0 * sobol.draw(n_init).to(DT) - 1.0                       # helpers.py:18-23
    Y = -torch.as_tensor(np.asarray(objective(X.numpy())), dtype=DT)
    bounds = torch.stack([-torch.ones(dim, dtype=DT), torch.ones(dim, dtype=DT)])
    while len(Y) < budget:
        train_Y = (Y - Y.mean()) / Y.std()
        gp = ogp.fit_adamw(X, train_Y, fit_cfg)
        best_f = float(train_Y.max())
        base = ogp.sobol_normal_samples(q, mc_samples, seed) if acq == "qei" else None
        vf = lambda Xb: ogp.acq_value(gp, acq, Xb, best_f, beta, base)
        vg = lambda Xb: ogp.acq_value_and_grad(gp, acq, Xb, best_f, beta, base)
        cand, _ = optimize_acqf(vf, vg, bounds, q, num_restarts, raw_samples, nonneg=acq in NONNEG_ACQ)
        Y_next = -torch.as_tensor(np.asarray(objective(cand.numpy())), dtype=DT)
        X = torch.cat((X, cand), dim=0)
        Y = torch.cat((Y, Y_next.reshape(-1)), dim=0)
    return X, -Y
