"""an ``oao``-compatible optimiser front-end over the CUDA hot path (no Ax, no BoTorch).

The reference's localisation / tilt jobs drive the objective through OAOptimization (``oao``, un-vendored and
un-pinned: ref/gp_dev.yml:26), a thin wrapper around Ax.  The interfaces mirrored here are the ones the reference's
call sites use (everything about oao/Ax internals is RECALLED):

* ``oao.space.{SearchParameter, SearchSpace, SearchParameterBounds, SearchSpaceBounds}``
  (ref/optimization/utils.py:12,72-98; ref/projects/swellex96_localization/conf/optimization/run.py:158-186);
* ``oao.objective.{NoiselessFormattedObjective, NoisyFormattedObjective}(objective, name, properties_kw,
  return_type)`` with ``.properties.minimize`` (conf/optimization/run.py:148-153, data/run.py:38-41);
* ``ax`` ``GenerationStep / GenerationStrategy / Models.{SOBOL, GPEI}`` as used by
  ref/optimization/strategy.py:15-69 (mirrored in :mod:`bogp_b200.strategy`) and ``oao.strategy.GridStrategy``;
* ``oao.optimizer.{BayesianOptimization, GridSearch, QuasiRandom}(objective, search_space, strategy)`` with
  ``.run(name)``, ``.client``, ``.batch_execution_times`` (data/run.py:30-41, conf/optimization/run.py:212-232);
* ``oao.results.get_results(client, times, minimize) -> DataFrame`` in the schema ref/optimization/aggregate.py:60-91
  consumes (index = trial index, one column per parameter, the metric, ``best_<param>``, ``best_value``, ``time``).

What runs where: trial generation for the Sobol step is host Python (``torch.quasirandom.SobolEngine``); the GPEI step
fits :class:`bogp_b200.gp.SingleTaskGP` on UnitX / StandardizeY-transformed data and optimises
``qExpectedImprovement`` with :func:`bogp_b200.optim.optimize_acqf` -- posterior, acquisition and gradients are the
float64 CUDA kernels of libbogp_sm100.so; a batch of trials is scored by ONE call of the objective with a list of
parameter dicts (one fused Bartlett launch) instead of Ax's one-trial-at-a-time loop.  A grid search over a
:class:`~bogp_b200.mfp.MatchedFieldProcessor` objective is one ``surface`` launch.
"""
from __future__ import annotations

import json
import time
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------ oao.space
@dataclass
class SearchParameter:
    name: str
    type: str = "range"
    bounds: Sequence[float] = (0.0, 1.0)
    value_type: str = "float"

    def __post_init__(self):
        if self.type != "range":
            raise ValueError("only range parameters are supported (the reference uses no other kind)")
        lo, hi = (float(b) for b in self.bounds)
        if not hi > lo:
            raise ValueError(f"parameter {self.name!r}: upper bound must exceed lower bound")
        self.bounds = [lo, hi]


class SearchSpace:
    def __init__(self, parameters: Sequence[SearchParameter]):
        self.parameters = list(parameters)
        names = [p.name for p in self.parameters]
        if len(set(names)) != len(names):
            raise ValueError("duplicate parameter names")

    @property
    def names(self) -> List[str]:
        return [p.name for p in self.parameters]

    @property
    def bounds(self) -> np.ndarray:
        """``[2, d]``"""
        return np.array([p.bounds for p in self.parameters], dtype=np.float64).T

    def __len__(self):
        return len(self.parameters)

    def to_unit(self, X: np.ndarray) -> np.ndarray:
        lo, hi = self.bounds
        return (np.asarray(X, dtype=np.float64) - lo) / (hi - lo)

    def from_unit(self, U: np.ndarray) -> np.ndarray:
        lo, hi = self.bounds
        return lo + (hi - lo) * np.asarray(U, dtype=np.float64)


@dataclass
class SearchParameterBounds:
    """conf/optimization/run.py:160-183: bounds relative to the scenario value, clipped to absolute limits."""

    name: str
    lower_bound: float
    upper_bound: float
    relative: bool = False
    min_lower_bound: Optional[float] = None
    max_upper_bound: Optional[float] = None


@dataclass
class SearchSpaceBounds:
    bounds: List[SearchParameterBounds] = field(default_factory=list)


def format_search_space(search_space: SearchSpaceBounds, scenario: dict) -> SearchSpace:
    """``optimization.utils.format_search_space`` (ref/optimization/utils.py:72-98), same clipping logic."""
    parameters = []
    for b in search_space.bounds:
        lower, upper = b.lower_bound, b.upper_bound
        if b.relative and b.name in scenario:
            lower = scenario[b.name] + b.lower_bound
            upper = scenario[b.name] + b.upper_bound
            if b.min_lower_bound is not None and lower < b.min_lower_bound:
                lower = b.min_lower_bound
                upper = lower + abs(b.lower_bound) + abs(b.upper_bound)
            elif b.max_upper_bound is not None and upper > b.max_upper_bound:
                upper = b.max_upper_bound
                lower = upper - abs(b.upper_bound) - abs(b.lower_bound)
        parameters.append(SearchParameter(name=b.name, type="range", bounds=[float(lower), float(upper)]))
    return SearchSpace(parameters)


# ------------------------------------------------------------------------------------------------ oao.objective
@dataclass
class Properties:
    minimize: bool = False


class Objective:
    """``oao.objective.Objective``: the base the configs annotate with (``format_objective(cfg, objective: Objective)``,
    ref/projects/swellex96_localization/conf/optimization/run.py:18,34)."""


class NoiselessFormattedObjective(Objective):
    """Wraps an objective callable ``f(params: dict) -> float`` (boundary A) for the optimiser loop."""

    noiseless = True

    def __init__(self, objective: Callable, name: str, properties_kw: Optional[dict] = None, return_type=float):
        self.objective = objective
        self.name = name
        self.properties = Properties(**(properties_kw or {}))
        self.return_type = return_type

    def __call__(self, parameters: dict):
        return {self.name: (self.return_type(self.objective(parameters)), 0.0 if self.noiseless else None)}

    def evaluate_batch(self, plist: List[dict]) -> np.ndarray:
        """One objective call for the whole batch when the objective accepts a list of dicts (the fused CUDA
        objective does: bogp_b200.mfp.MatchedFieldProcessor.__call__), else one call per trial."""
        if getattr(self.objective, "_call_batch", None) is not None:
            return np.asarray(self.objective(list(plist)), dtype=np.float64).reshape(-1)
        return np.array([float(self.objective(p)) for p in plist], dtype=np.float64)


class NoisyFormattedObjective(NoiselessFormattedObjective):
    noiseless = False


# ------------------------------------------------------------------------------------------------ ax stand-ins
class Models(Enum):
    SOBOL = "Sobol"
    GPEI = "GPEI"
    UNIFORM = "Uniform"


@dataclass
class GenerationStep:
    model: Models
    num_trials: int
    max_parallelism: Optional[int] = None
    model_kwargs: Dict[str, Any] = field(default_factory=dict)
    min_trials_observed: Optional[int] = None


class GenerationStrategy:
    def __init__(self, steps: Sequence[GenerationStep], name: Optional[str] = None):
        self._steps = list(steps)
        self.name = name or "+".join(s.model.value for s in self._steps)

    @property
    def num_trials(self) -> int:
        return sum(s.num_trials for s in self._steps)


class GridStrategy:
    """``oao.strategy.GridStrategy(num_trials, max_parallelism)``: ``num_trials`` points per search dimension
    [recalled]; subclassed by ref/optimization/strategy.py:15-23."""

    def __init__(self, num_trials: int = 12, max_parallelism: int = 1):
        self.num_trials = int(num_trials)
        self.max_parallelism = int(max_parallelism)

    def __call__(self, *args, **kwargs):
        return self


# ------------------------------------------------------------------------------------------------ client / results
class Client:
    """The slice of ``AxClient`` the reference touches: trial records and ``save_to_json_file`` (data/run.py:42)."""

    def __init__(self, name: str, search_space: SearchSpace, metric_name: str, minimize: bool):
        self.name, self.search_space, self.metric_name, self.minimize = name, search_space, metric_name, minimize
        self.trials: List[dict] = []

    def attach(self, parameters: dict, value: float, method: str, batch: int) -> int:
        idx = len(self.trials)
        self.trials.append({"trial_index": idx, "parameters": dict(parameters), "value": float(value),
                            "generation_method": method, "batch": int(batch)})
        return idx

    def get_best_parameters(self):
        if not self.trials:
            return None
        vals = np.array([t["value"] for t in self.trials])
        i = int(np.argmin(vals) if self.minimize else np.argmax(vals))
        return self.trials[i]["parameters"], {self.metric_name: self.trials[i]["value"]}

    def to_json_snapshot(self) -> dict:
        return {"name": self.name, "metric_name": self.metric_name, "minimize": self.minimize,
                "search_space": [{"name": p.name, "type": p.type, "bounds": list(p.bounds)}
                                 for p in self.search_space.parameters],
                "trials": self.trials}

    def save_to_json_file(self, filepath) -> None:
        with open(filepath, "w") as fh:
            json.dump(self.to_json_snapshot(), fh, indent=1)


def get_results(client: Client, times: Sequence[float], minimize: bool):
    """``oao.results.get_results``: one row per trial -- parameters, metric, running best, the wall time of the batch
    the trial belonged to (ref/optimization/aggregate.py:79-91 divides equal ``time`` values among their trials)."""
    import pandas as pd
    names = client.search_space.names
    rows, best, best_val = [], None, None
    for t in client.trials:
        v = t["value"]
        if best_val is None or (v < best_val if minimize else v > best_val):
            best_val, best = v, t["parameters"]
        row = {"trial_index": t["trial_index"], "generation_method": t["generation_method"]}
        row.update({n: t["parameters"][n] for n in names})
        row[client.metric_name] = v
        row.update({f"best_{n}": best[n] for n in names})
        row["best_value"] = best_val
        row["time"] = float(times[t["batch"]]) if t["batch"] < len(times) else float("nan")
        rows.append(row)
    return pd.DataFrame(rows).set_index("trial_index")


# ------------------------------------------------------------------------------------------------ optimisers
class _Optimizer:
    def __init__(self, objective: NoiselessFormattedObjective, search_space: SearchSpace, strategy,
                 monitor: Optional[Callable] = None):
        self.objective, self.search_space, self.strategy, self.monitor = objective, search_space, strategy, monitor
        self.client: Optional[Client] = None
        self.batch_execution_times: List[float] = []

    def _new_client(self, name: str) -> Client:
        self.client = Client(name, self.search_space, self.objective.name, self.objective.properties.minimize)
        self.batch_execution_times = []
        return self.client

    def _complete(self, X: np.ndarray, method: str, t0: float) -> None:
        names = self.search_space.names
        plist = [{n: float(x[i]) for i, n in enumerate(names)} for x in np.atleast_2d(X)]
        vals = self.objective.evaluate_batch(plist)
        batch = len(self.batch_execution_times)
        for p, v in zip(plist, vals):
            self.client.attach(p, v, method, batch)
        self.batch_execution_times.append(time.perf_counter() - t0)
        if self.monitor is not None:
            self.monitor(self.client)


class BayesianOptimization(_Optimizer):
    """Sobol warm-up then GP + (q)EI, as the ``GenerationStrategy`` says (ref/optimization/strategy.py:26-69).

    GPEI step [Ax legacy ``Models.GPEI`` -- recalled]: UnitX + StandardizeY transforms, Matern-5/2 ARD
    ``SingleTaskGP`` (noise fixed at ``noiseless_noise`` for a noiseless objective), hyper-parameters by
    ``fit(FitConfig)``, ``qExpectedImprovement`` over ``q = min(max_parallelism, remaining)`` points optimised jointly
    with ``optimize_acqf(num_restarts, raw_samples)`` on [0, 1]^d.
    """

    def __init__(self, objective, search_space, strategy: GenerationStrategy, monitor=None, *, num_restarts: int = 20,
                 raw_samples: int = 1024, mc_samples: int = 512, fit_steps: int = 100, covar_module: str = "matern52",
                 noiseless_noise: float = 1e-6, backend=None):
        super().__init__(objective, search_space, strategy, monitor)
        self.num_restarts, self.raw_samples, self.mc_samples = num_restarts, raw_samples, mc_samples
        self.fit_steps, self.covar_module, self.noiseless_noise = fit_steps, covar_module, noiseless_noise
        self.backend = backend          # test hook: object with gp_ei_candidates(U, y, q, seed) -> U_next [q, d]

    # -- one GPEI batch: returns q points of the unit cube
    def _gpei_candidates(self, U: np.ndarray, y: np.ndarray, q: int, seed: int) -> np.ndarray:
        if self.backend is not None:
            return self.backend.gp_ei_candidates(U, y, q, seed)
        from .acquisition import ExpectedImprovement, qExpectedImprovement
        from .gp import FitConfig, SingleTaskGP
        from .optim import optimize_acqf
        gen = torch.Generator().manual_seed(int(seed))      # local stream: the caller's global torch RNG is left alone
        d = U.shape[1]
        model = SingleTaskGP(torch.as_tensor(U), torch.as_tensor(y), covar_module=self.covar_module)
        cfg = FitConfig(steps=self.fit_steps)
        if self.objective.noiseless:
            cfg = FitConfig(steps=self.fit_steps, noise_bounds=(0.5 * self.noiseless_noise, 2.0 * self.noiseless_noise))
        model.fit(cfg)
        best_f = float(np.max(y))
        acq = (ExpectedImprovement(model, best_f) if q == 1 else
               qExpectedImprovement(model, best_f, q=q, mc_samples=self.mc_samples, seed=seed))
        bounds = torch.stack([torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)])
        cand, _ = optimize_acqf(acq, bounds=bounds, q=q, num_restarts=self.num_restarts, raw_samples=self.raw_samples, generator=gen)
        return cand.reshape(q, d).numpy()

    def run(self, name: str = "bo") -> Client:
        client = self._new_client(name)
        d = len(self.search_space)
        sign = -1.0 if self.objective.properties.minimize else 1.0
        for si, step in enumerate(self.strategy._steps):
            par = int(step.max_parallelism or 1)
            done = 0
            sobol = None
            if step.model in (Models.SOBOL, Models.UNIFORM):
                seed = step.model_kwargs.get("seed")
                sobol = (torch.quasirandom.SobolEngine(dimension=d, scramble=True, seed=seed)
                         if step.model is Models.SOBOL else np.random.default_rng(seed))
            while done < step.num_trials:
                t0 = time.perf_counter()
                q = min(par, step.num_trials - done)
                if step.model is Models.SOBOL:
                    U = sobol.draw(q, dtype=torch.float64).numpy()
                elif step.model is Models.UNIFORM:
                    U = sobol.random((q, d))
                else:
                    if not client.trials:
                        raise RuntimeError("the GPEI step needs completed trials (put a Sobol step first)")
                    X = np.array([[t["parameters"][n] for n in self.search_space.names] for t in client.trials])
                    y = sign * np.array([t["value"] for t in client.trials])
                    ys = (y - y.mean()) / (y.std(ddof=1) if len(y) > 1 and y.std(ddof=1) > 0 else 1.0)   # StandardizeY
                    U = np.clip(self._gpei_candidates(self.search_space.to_unit(X), ys, q,
                                                      seed=1000 * si + len(client.trials)), 0.0, 1.0)
                self._complete(self.search_space.from_unit(U), step.model.value, t0)
                done += q
        return client


class QuasiRandom(BayesianOptimization):
    """``oao.optimizer.QuasiRandom``: the same loop; the strategy holds only Sobol / uniform steps."""


class GridSearch(_Optimizer):
    """Full-factorial grid, ``strategy.num_trials`` points per dimension (end points included).  When the objective
    is the fused matched-field processor over (``rec_r``, ``src_z``[, ``tilt``]) the whole grid is ONE surface launch
    (what ref/projects/swellex96_localization/data/mfp.py:213-214 loops over)."""

    def run(self, name: str = "grid") -> Client:
        client = self._new_client(name)
        names = self.search_space.names
        n = int(self.strategy.num_trials)
        lo, hi = self.search_space.bounds
        axes = [np.linspace(lo[i], hi[i], n) for i in range(len(names))]
        t0 = time.perf_counter()
        obj = self.objective.objective
        if hasattr(obj, "surface") and set(names) in ({"rec_r", "src_z"}, {"rec_r", "src_z", "tilt"}):
            ax = dict(zip(names, axes))
            surf = np.asarray(obj.surface(ax["rec_r"], ax["src_z"], ax.get("tilt")))      # [(Nt,) Nz, Nr]
            order = (["tilt"] if "tilt" in ax else []) + ["src_z", "rec_r"]
            mesh = np.meshgrid(*[ax[k] for k in order], indexing="ij")
            batch = len(self.batch_execution_times)
            for idx in np.ndindex(*surf.shape):
                client.attach({k: float(m[idx]) for k, m in zip(order, mesh)}, float(surf[idx]), "Grid", batch)
            self.batch_execution_times.append(time.perf_counter() - t0)
        else:
            mesh = np.meshgrid(*axes, indexing="ij")
            X = np.stack([m.reshape(-1) for m in mesh], axis=1)
            self._complete(X, "Grid", t0)
        return client
