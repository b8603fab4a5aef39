"""Generation strategies with the reference's names and signatures (ref/optimization/strategy.py:15-69).

``SobolStrategy`` / ``GPEIStrategy`` / ``GridSearchStrategy`` describe the trial schedule the reference hands to Ax --
a quasi-random warm-up, optionally followed by model-based batches -- in terms of the stand-ins of
:mod:`bogp_b200.oao`, which :class:`bogp_b200.oao.BayesianOptimization` executes on the CUDA GP path.  Calling an
instance (``strategy(seed=...)()``) hands out a fresh strategy over the same schedule, which is what
ref/projects/swellex96_localization/data/run.py:28-34 relies on.
"""
from __future__ import annotations

from .oao import GenerationStep, GenerationStrategy, GridStrategy, Models


def _warmup(trials: int, parallelism: int, seed: int) -> GenerationStep:
    """Scrambled-Sobol block: ``trials`` points, ``parallelism`` of them in flight, reproducible through ``seed``."""
    return GenerationStep(model=Models.SOBOL, num_trials=trials, max_parallelism=parallelism, model_kwargs={"seed": seed})


def _model_based(trials: int, batch: int) -> GenerationStep:
    """GP + expected improvement: ``batch`` candidates per iteration (q-EI when ``batch > 1``)."""
    return GenerationStep(model=Models.GPEI, num_trials=trials, max_parallelism=batch)


class _Schedule(GenerationStrategy):
    """A strategy that re-issues its own schedule when called (the reference calls the configured object once per job)."""

    def __init__(self, *blocks: GenerationStep):
        super().__init__(steps=list(blocks))

    def __call__(self):
        return GenerationStrategy(steps=self._steps)


class SobolStrategy(_Schedule):
    """strategy.py:26-40 -- quasi-random search only."""

    def __init__(self, num_trials=144, max_parallelism=16, seed=2009):
        super().__init__(_warmup(num_trials, max_parallelism, seed))


class GPEIStrategy(_Schedule):
    """strategy.py:43-69 -- Sobol warm-up, then GP + (q)EI with ``max_parallelism`` trials per batch."""

    def __init__(self, warmup_trials=128, warmup_parallelism=16, num_trials=16, max_parallelism=1, seed=292288111):
        super().__init__(_warmup(warmup_trials, warmup_parallelism, seed), _model_based(num_trials, max_parallelism))


class GridSearchStrategy(GridStrategy):
    """strategy.py:15-23 -- the exhaustive grid; extra arguments of the config are accepted and ignored as there."""

    def __init__(self, num_trials=12, max_parallelism=1, *args, **kwargs):
        GridStrategy.__init__(self, num_trials=num_trials, max_parallelism=max_parallelism)

    def __call__(self, *args, **kwargs):
        return self
