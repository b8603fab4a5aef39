"""BAxUS (trust region + growing sparse embedding) with Thompson sampling on the device -- SURVEY 8(f) row 4.

Host-side mirror of ref/projects/swellex96_inv/data/bo/baxus.py (same names, arguments and state machine):

* :class:`BAxUSLoopArgs` (:31-46), :class:`BaxusState` (:49-94), :func:`update_state` (:238-259)
* :func:`embedding_matrix` (:156-185) and :func:`increase_embedding_and_observations` (:188-235): the sparse
  ``target_dim x input_dim`` sign embedding and its splitting when the trust region collapses
* :func:`create_candidate` (:97-153): trust region scaled by the length scales, scrambled-Sobol perturbations, then
  ``acqf="ts"`` -> :class:`bogp_b200.generation.MaxPosteriorSampling` (one joint posterior draw over
  ``n_candidates <= 5000`` points: ``bogp_gp_thompson``, csrc/kernels_ts.cu) or ``acqf="ei"`` ->
  ``LogExpectedImprovement`` + ``optimize_acqf`` inside the trust region
* :func:`loop` (:262-425): Sobol warm-up in the target space, then per trial: standardise Y, fit the explicit
  Matern-5/2 ARD GP with ``Interval`` constraints (:304-319), create one candidate, evaluate, update the state, and
  split the embedding when the trust region shrinks below ``length_min``.

Difference from the reference, stated: its primary fit is ``fit_gpytorch_mll`` (scipy L-BFGS-B) with 100 AdamW steps
as the fallback (:333-345); here the AdamW loop is the fit (one persistent kernel on the device for d <= 16, the
cuSOLVER loop above), maximising the same marginal likelihood.  Random streams (Sobol scrambling, masks, embedding
signs) come from the torch global generator as in the reference, without a claim of stream identity.
"""
from __future__ import annotations

import logging
import math
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
from torch.quasirandom import SobolEngine

from .acquisition import LogExpectedImprovement
from .generation import MaxPosteriorSampling
from .gp import DT, FitConfig, SingleTaskGP, fit_gp_adamw
from .optim import optimize_acqf

log = logging.getLogger(__name__)


class EmbeddingExhausted(TypeError):
    """A row of the embedding has a single contributing input dimension left: it cannot be split.

    The reference reaches a ``TypeError`` in this situation (``len()`` of the squeezed 0-d index tensor, :199-207) and
    its loop ends the optimisation on it (:409-415); subclassing ``TypeError`` keeps ``except TypeError`` callers working.
    """


@dataclass
class BAxUSLoopArgs:
    """Arguments of :func:`loop` as ``run.py`` builds them (ref .../bo/run.py:63-67)."""

    true_dim: int
    budget: int = 500
    n_init: int = 100
    dtype: torch.dtype = torch.double
    device: str = "cuda"
    dummy_dim: Optional[int] = 200
    num_restarts: int = 40
    raw_samples: int = 1024
    dim: int = field(init=False)
    n_candidates: int = field(init=False)
    seed: int = 0

    def __post_init__(self):
        self.dim = self.true_dim + (self.dummy_dim or 0)
        self.n_candidates = min(5000, max(2000, 200 * self.dim))


@dataclass
class BaxusState:
    """Trust-region and embedding bookkeeping (ref baxus.py:49-94)."""

    dim: int
    eval_budget: int
    new_bins_on_split: int = 3
    d_init: int = field(init=False)
    target_dim: int = field(init=False)
    n_splits: int = field(init=False)
    length: float = 0.8
    length_init: float = 0.8
    length_min: float = 0.5 ** 7
    length_max: float = 1.6
    failure_counter: int = 0
    success_counter: int = 0
    success_tolerance: int = 3
    best_value: float = -float("inf")
    restart_triggered: bool = False

    def __post_init__(self):
        growth = self.new_bins_on_split + 1
        self.n_splits = round(math.log(self.dim, growth))
        # initial target dimension: the multiple i * growth^n_splits (i = 1 .. new_bins_on_split) closest to dim
        reach = (1 + np.arange(self.new_bins_on_split)) * growth ** self.n_splits
        self.d_init = int(1 + np.argmin(np.abs(reach - self.dim)))
        self.target_dim = self.d_init

    @property
    def split_budget(self) -> int:
        """Evaluations granted to the current target dimension (geometric share of ``eval_budget``)."""
        growth = self.new_bins_on_split + 1
        return round(-(self.new_bins_on_split * self.eval_budget * self.target_dim)
                     / (self.d_init * (1 - growth ** (self.n_splits + 1))))

    @property
    def failure_tolerance(self) -> int:
        if self.target_dim == self.dim:
            return self.target_dim
        halvings = math.floor(math.log(self.length_min / self.length_init, 0.5))
        return min(self.target_dim, max(1, math.floor(self.split_budget / halvings)))


def update_state(state: BaxusState, Y_next) -> BaxusState:
    """Success / failure counters -> trust-region length; flags a restart below ``length_min`` (ref :238-259)."""
    y_best = float(torch.as_tensor(Y_next).max())
    if y_best > state.best_value + 1e-3 * math.fabs(state.best_value):
        state.success_counter += 1
        state.failure_counter = 0
    else:
        state.success_counter = 0
        state.failure_counter += 1
    if state.success_counter == state.success_tolerance:
        state.length = min(2.0 * state.length, state.length_max)
        state.success_counter = 0
    elif state.failure_counter == state.failure_tolerance:
        state.length /= 2.0
        state.failure_counter = 0
    state.best_value = max(state.best_value, y_best)
    if state.length < state.length_min:
        state.restart_triggered = True
    return state


def embedding_matrix(input_dim: int, target_dim: int, dtype: torch.dtype = DT, device=None) -> torch.Tensor:
    """Sparse sign embedding ``S[target_dim, input_dim]``: every input dimension belongs to exactly one target
    dimension (a random permutation cut into ``target_dim`` near-equal bins) with a random sign; the identity when
    ``target_dim >= input_dim`` (ref :156-185).  ``x_input = x_target @ S``."""
    if target_dim >= input_dim:
        return torch.eye(input_dim, dtype=dtype)
    # same draws from the global generator as the reference (one permutation, one [target_dim, input_dim] sign table of
    # which row i uses its first len(bin i) entries): identical embeddings on a fixed torch seed
    perm = torch.randperm(input_dim)
    signs = 2 * torch.randint(2, (target_dim, input_dim), dtype=dtype) - 1
    S = torch.zeros(target_dim, input_dim, dtype=dtype)
    for row, members in enumerate(torch.tensor_split(perm, target_dim)):
        S[row, members] = signs[row, :len(members)]
    return S


def increase_embedding_and_observations(S: torch.Tensor, X: torch.Tensor, n_new_bins: int, dtype: torch.dtype = DT,
                                        device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Split every target dimension into up to ``n_new_bins`` bins (ref :188-235): the first bin keeps the row, each
    further bin becomes a new row appended to ``S``; the observations get one copy of the parent's column per new row,
    so that ``X_new @ S_new == X @ S`` (the embedded points do not move)."""
    if X.shape[1] != S.shape[0]:
        raise AssertionError("Observations don't lie in row space of S")
    S_new = S.clone()
    X_new = X.clone()
    extra_rows: List[torch.Tensor] = []
    extra_cols: List[torch.Tensor] = []
    for row in range(S.shape[0]):
        members = torch.nonzero(S[row]).reshape(-1)
        if len(members) == 1:
            raise EmbeddingExhausted(f"target dimension {row} has a single input dimension left")
        members = members[torch.randperm(len(members))]
        bins = torch.tensor_split(members, min(n_new_bins, len(members)))[1:]
        longest = max((len(b) for b in bins), default=0)
        for moved in bins:
            new_row = torch.zeros(S.shape[1], dtype=S.dtype)
            new_row[moved] = S[row, moved]
            # Reference behaviour kept for identical trajectories (ref :215-224): the moved bins of a row are zero-padded
            # to the longest one before the scatter, and the padding entries (index 0, value 0) are written after the
            # real ones -- a bin that is shorter than the longest AND contains input dimension 0 loses that dimension.
            if len(moved) < longest and bool((moved == 0).any()):
                new_row[0] = 0
            S_new[row, moved] = 0
            extra_rows.append(new_row)
            extra_cols.append(X[:, row])
    if extra_rows:
        S_new = torch.vstack([S_new, torch.stack(extra_rows)])
        X_new = torch.hstack([X_new, torch.stack(extra_cols, dim=1)])
    return S_new, X_new


def _trust_region(state: BaxusState, model: SingleTaskGP, X: torch.Tensor, Y: torch.Tensor):
    center = X[int(torch.argmax(Y.reshape(-1)))].clone()
    w = model.lengthscale.detach().reshape(-1).to(DT)
    w = w / w.mean()
    w = w / torch.prod(w.pow(1.0 / len(w)))          # geometric mean 1: the box keeps its volume
    lb = torch.clamp(center - w * state.length, -1.0, 1.0)
    ub = torch.clamp(center + w * state.length, -1.0, 1.0)
    return center, lb, ub


def thompson_candidates(center: torch.Tensor, lb: torch.Tensor, ub: torch.Tensor, n_candidates: int) -> torch.Tensor:
    """The discrete candidate set of the Thompson step (ref :121-134): scrambled-Sobol points of the trust region, of
    which every candidate keeps only a random subset of coordinates (each with probability ``min(20 / dim, 1)``, at
    least one) and stays at the incumbent ``center`` in the others."""
    dim = center.shape[-1]
    pert = SobolEngine(dim, scramble=True).draw(n_candidates).to(DT)
    pert = lb + (ub - lb) * pert
    prob = min(20.0 / dim, 1.0)
    mask = torch.rand(n_candidates, dim, dtype=DT) <= prob
    stuck = torch.where(mask.sum(dim=1) == 0)[0]
    mask[stuck, torch.randint(0, dim, size=(len(stuck),))] = True
    X_cand = center.expand(n_candidates, dim).clone()
    X_cand[mask] = pert[mask]
    return X_cand


def create_candidate(state: BaxusState, model: SingleTaskGP, X: torch.Tensor, Y: torch.Tensor,
                     dtype: torch.dtype = DT, device=None, n_candidates: Optional[int] = None, num_restarts: int = 10,
                     raw_samples: int = 512, acqf: str = "ts") -> torch.Tensor:
    """One new point ``[1, d]`` in the target space (ref :97-153)."""
    if acqf not in ("ts", "ei"):
        raise AssertionError("acqf must be 'ts' or 'ei'")
    X = torch.as_tensor(X, dtype=DT)
    Y = torch.as_tensor(Y, dtype=DT)
    if not (X.min() >= -1.0 and X.max() <= 1.0 and bool(torch.all(torch.isfinite(Y)))):
        raise AssertionError("X must lie in [-1, 1]^d and Y must be finite")
    dim = X.shape[-1]
    if n_candidates is None:
        n_candidates = min(5000, max(2000, 200 * dim))
    center, lb, ub = _trust_region(state, model, X, Y)
    if acqf == "ts":
        X_cand = thompson_candidates(center, lb, ub, n_candidates)
        return MaxPosteriorSampling(model=model, replacement=False)(X_cand, num_samples=1)
    logei = LogExpectedImprovement(model, float(Y.max()))
    X_next, _ = optimize_acqf(logei, bounds=torch.stack([lb, ub]), q=1, num_restarts=num_restarts,
                              raw_samples=raw_samples)
    return torch.as_tensor(X_next, dtype=DT).reshape(1, dim).cpu()


def get_initial_points(dim: int, n_pts: int, seed: int = 0) -> torch.Tensor:
    """Scrambled Sobol points in ``[-1, 1]^dim`` (ref .../bo/helpers.py:18-23)."""
    return 2.0 * SobolEngine(dimension=dim, scramble=True, seed=seed).draw(n=n_pts).to(DT) - 1.0


def baxus_fit_config(steps: int = 100, lr: float = 0.1) -> FitConfig:
    """The explicit model of ref :304-319: no priors, ``Interval`` constraints on noise / length scales / output scale."""
    return FitConfig(steps=steps, lr=lr, noise_bounds=(1e-8, 1e-3), lengthscale_prior=None, outputscale_prior=None,
                     lengthscale_bounds=(0.005, 10.0), outputscale_bounds=(0.05, 10.0))


def loop(objective: Callable, dim: int, budget: int, n_init: int, true_dim: int, num_restarts: int = 40,
         raw_samples: int = 1024, n_candidates: Optional[int] = None, seed: int = 0, dtype: torch.dtype = DT,
         device=None, acqf: str = "ts", fit_cfg: Optional[FitConfig] = None, *args, **kwargs):
    """Run BAxUS for ``budget`` evaluations (ref :262-425).

    ``objective(X_input[m, dim] numpy, true_dim=true_dim) -> m values`` is MINIMISED (the loop maximises its negative).
    Returns ``(X[:, :true_dim], Y, times)``: evaluated points in the input space, objective values, seconds per trial.
    """
    state = BaxusState(dim=dim, eval_budget=budget - n_init)
    S = embedding_matrix(input_dim=state.dim, target_dim=state.d_init)
    cfg = fit_cfg if fit_cfg is not None else baxus_fit_config()
    if n_candidates is None:
        n_candidates = min(5000, max(2000, 200 * dim))

    start = time.time()
    X_target = get_initial_points(state.d_init, n_init, seed)
    X_input = X_target @ S
    Y = -torch.as_tensor(np.asarray(objective(X_input.numpy(), true_dim=true_dim), dtype=np.float64), dtype=DT).reshape(-1)
    times = [(time.time() - start) / n_init] * n_init
    log.info("%d warmup trials complete.", n_init)

    for _ in range(budget - n_init):
        start = time.time()
        train_Y = (Y - Y.mean()) / Y.std()
        model = SingleTaskGP(X_target, train_Y, covar_module="matern52", noise_bounds=cfg.noise_bounds)
        fit_gp_adamw(model, cfg)
        X_next_target = create_candidate(state=state, model=model, X=X_target, Y=train_Y, n_candidates=n_candidates,
                                         num_restarts=num_restarts, raw_samples=raw_samples, acqf=acqf)
        X_next_input = X_next_target @ S
        Y_next = -torch.as_tensor(np.asarray(objective(X_next_input.numpy(), true_dim=true_dim), dtype=np.float64),
                                  dtype=DT).reshape(-1)
        state = update_state(state=state, Y_next=Y_next)
        X_input = torch.cat((X_input, X_next_input), dim=0)
        X_target = torch.cat((X_target, X_next_target), dim=0)
        Y = torch.cat((Y, Y_next), dim=0)
        times.append(time.time() - start)
        log.info("BAxUS | Trial %d | dim %d | TR length %.3g | best %.6g", len(X_input), X_target.shape[1], state.length,
                 float(-Y.max()))
        if state.restart_triggered:
            state.restart_triggered = False
            try:
                S, X_target = increase_embedding_and_observations(S, X_target, state.new_bins_on_split)
            except TypeError:
                log.info("Unable to increase target space dimensionality; optimization terminating.")
                return X_input[:, :true_dim], -Y, times
            log.info("New dimensionality: %d", len(S))
            state.target_dim = len(S)
            state.length = state.length_init
            state.failure_counter = 0
            state.success_counter = 0
    return X_input[:, :true_dim], -Y, times
