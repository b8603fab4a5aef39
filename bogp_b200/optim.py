"""Boundary B: ``optimize_acqf``.

Mirrors ``botorch.optim.optimize_acqf(acq_function, bounds, q, num_restarts, raw_samples)`` as called at
ref/projects/swellex96_inv/data/bo/ei.py:81-92 (q=1, 40 restarts, 1024 raw samples, bounds [-1, 1]^d).  BoTorch's
internals are RECALLED: scrambled-Sobol raw samples seeded from the torch RNG, Boltzmann restart
selection (``initialize_q_batch[_nonneg]``), one joint SciPy L-BFGS-B over all restarts on ``-sum(acq)``, best
restart returned.  The orchestration is host Python/SciPy as in BoTorch; every acquisition evaluation (value and
gradient, all restarts at once) is one fused CUDA launch.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
from scipy.optimize import minimize

DT = torch.float64


def draw_sobol_samples(bounds: torch.Tensor, n: int, q: int, seed: Optional[int] = None) -> torch.Tensor:
    d = bounds.shape[-1]
    eng = torch.quasirandom.SobolEngine(dimension=q * d, scramble=True, seed=seed)
    u = eng.draw(n, dtype=DT).reshape(n, q, d)
    return bounds[0] + (bounds[1] - bounds[0]) * u


def _boltzmann(X: torch.Tensor, Y: torch.Tensor, n: int, eta: float = 2.0, generator=None) -> torch.Tensor:
    if n >= X.shape[0]:
        return X
    std = Y.std()
    if float(std) == 0.0:
        return X[torch.randperm(X.shape[0], generator=generator)[:n]]
    max_idx = int(torch.argmax(Y))
    etaZ = eta * (Y - Y.mean()) / std
    w = torch.exp(etaZ)
    while bool(torch.isinf(w).any()):
        etaZ = etaZ * 0.5
        w = torch.exp(etaZ)
    idcs = torch.multinomial(w, n, generator=generator)
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs]


def _boltzmann_nonneg(X: torch.Tensor, Y: torch.Tensor, n: int, eta: float = 1.0, alpha: float = 1e-4, generator=None) -> torch.Tensor:
    if n >= X.shape[0]:
        return X
    max_val, max_idx = float(Y.max()), int(torch.argmax(Y))
    if max_val <= 0.0:
        return X[torch.randperm(X.shape[0], generator=generator)[:n]]
    pos = Y >= alpha * max_val
    while int(pos.sum()) < n:
        alpha *= 0.1
        pos = Y >= alpha * max_val
    pos_idcs = torch.nonzero(pos).squeeze(-1)
    w = torch.exp(eta * (Y[pos] / max_val - 1.0))
    idcs = pos_idcs[torch.multinomial(w, n, generator=generator)]
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs]


def _all_gather_rows(x: torch.Tensor, group) -> torch.Tensor:
    """Concatenation over ranks of a float64 CPU tensor whose leading dimension may differ per rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    n = torch.tensor([x.shape[0]], dtype=torch.int64, device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    nmax = max(int(v) for v in ns)
    pad = torch.zeros((nmax,) + tuple(x.shape[1:]), dtype=x.dtype, device=dev)
    pad[: x.shape[0]] = x.to(dev)
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p_[: int(k)] for p_, k in zip(parts, ns)]).cpu()


def optimize_acqf(acq_function, bounds, q: int, num_restarts: int, raw_samples: int, options: Optional[dict] = None,
                  return_best_only: bool = True, generator: Optional[torch.Generator] = None,
                  group=None, shard: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns ``(candidates[q, d], acq_value)`` as float64 CPU tensors.  Random draws (Sobol seed, restart selection)
    come from the torch global generator as in BoTorch, or from ``generator`` when one is given (callers that must
    not disturb the user's random stream).

    ``shard=True`` (inside an initialised ``torch.distributed`` group, one process per GPU, the GP replicated): the raw
    samples are split over the ranks (``raw_samples / P`` evaluations each), their values all-gathered before the restart
    selection -- which rank 0 draws and broadcasts -- and every rank then optimises ``num_restarts / P`` of the restarts;
    the final candidates and values are all-gathered and every rank returns the same best one.  The joint
    L-BFGS-B problem is split into P smaller ones: the same local maxima within the optimiser's tolerance, not bit-equal
    iterates."""
    options = options or {}
    bounds = torch.as_tensor(bounds, dtype=DT).cpu()
    d = bounds.shape[-1]
    rank, world = 0, 1
    if shard:
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("optimize_acqf(shard=True) needs an initialised torch.distributed process group")
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    seed_t = torch.randint(0, 10000, (1,), generator=generator)
    if world > 1:
        cdev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        seed_t = seed_t.to(cdev)
        dist.broadcast(seed_t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    seed = int(seed_t.item())
    X_raw = draw_sobol_samples(bounds, raw_samples, q, seed)
    with torch.no_grad():
        if world > 1:
            lo, hi = rank * raw_samples // world, (rank + 1) * raw_samples // world
            Y_raw = _all_gather_rows(acq_function(X_raw[lo:hi]).detach().cpu().reshape(-1), group)
        else:
            Y_raw = acq_function(X_raw).detach().cpu()
    init = _boltzmann_nonneg if getattr(acq_function, "nonneg", False) else _boltzmann
    X0 = init(X_raw, Y_raw, num_restarts, generator=generator)
    if world > 1:
        X0 = X0.contiguous().to(cdev)
        dist.broadcast(X0, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        X0 = X0.cpu()
        n0 = X0.shape[0]
        X0 = X0[rank * n0 // world:(rank + 1) * n0 // world]
    shape = X0.shape
    lo = bounds[0].repeat(shape[0] * q).numpy()
    hi = bounds[1].repeat(shape[0] * q).numpy()

    fast = getattr(acq_function, "value_and_grad_host", None)

    def f_and_g(x: np.ndarray):
        if fast is not None:          # one C call: pinned staging, value + analytic gradient, one synchronisation
            v, g = fast(x.reshape(tuple(shape)))
            return -float(v.sum()), -g.reshape(-1)
        X = torch.from_numpy(x.copy()).reshape(shape).requires_grad_(True)
        v = acq_function(X)
        loss = -v.sum()
        (g,) = torch.autograd.grad(loss, X)
        return float(loss.detach()), g.reshape(-1).cpu().numpy().astype(np.float64)

    x0 = np.clip(X0.reshape(-1).numpy().astype(np.float64), lo, hi)
    res = minimize(f_and_g, x0, jac=True, method="L-BFGS-B", bounds=list(zip(lo, hi)),
                   options={"maxiter": int(options.get("maxiter", 200))})
    Xc = torch.from_numpy(np.clip(res.x, lo, hi)).reshape(shape)
    if fast is not None:
        vals = torch.from_numpy(fast(Xc.numpy(), need_grad=False)[0])
    else:
        with torch.no_grad():
            vals = acq_function(Xc).detach().cpu()
    if world > 1:
        Xc = _all_gather_rows(Xc.reshape(shape[0], -1), group).reshape((-1,) + tuple(shape[1:]))
        vals = _all_gather_rows(vals.reshape(-1).to(DT), group)
    if not return_best_only:
        return Xc, vals
    best = int(torch.argmax(vals))
    return Xc[best], vals[best]
