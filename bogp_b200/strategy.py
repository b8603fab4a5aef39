"""Generation strategies with the reference's names and signatures (ref/optimization/strategy.py:15-69).

``SobolStrategy`` / ``GPEIStrategy`` / ``GridSearchStrategy`` build the step lists the reference builds from Ax's
``GenerationStep`` and ``Models`` -- here from the stand-ins in :mod:`bogp_b200.oao`, which
:class:`bogp_b200.oao.BayesianOptimization` executes on the CUDA GP path.  ``strategy(seed=...)()`` returns a fresh
strategy exactly as ref/projects/swellex96_localization/data/run.py:28-34 expects.
"""
from __future__ import annotations

from .oao import GenerationStep, GenerationStrategy, GridStrategy, Models


class GridSearchStrategy(GridStrategy):
    """strategy.py:15-23"""

    def __init__(self, num_trials=12, max_parallelism=1, *args, **kwargs):
        super().__init__(num_trials=num_trials, max_parallelism=max_parallelism)

    def __call__(self, *args, **kwargs):
        return self


class SobolStrategy(GenerationStrategy):
    """strategy.py:26-40"""

    def __init__(self, num_trials=144, max_parallelism=16, seed=2009):
        super().__init__(steps=[GenerationStep(model=Models.SOBOL, num_trials=num_trials,
                                               max_parallelism=max_parallelism, model_kwargs={"seed": seed})])

    def __call__(self):
        return GenerationStrategy(steps=self._steps)


class GPEIStrategy(GenerationStrategy):
    """strategy.py:43-69: Sobol warm-up, then GP + (q)EI with ``max_parallelism`` trials per batch."""

    def __init__(self, warmup_trials=128, warmup_parallelism=16, num_trials=16, max_parallelism=1, seed=292288111):
        super().__init__(steps=[
            GenerationStep(model=Models.SOBOL, num_trials=warmup_trials, max_parallelism=warmup_parallelism,
                           model_kwargs={"seed": seed}),
            GenerationStep(model=Models.GPEI, num_trials=num_trials, max_parallelism=max_parallelism),
        ])

    def __call__(self):
        return GenerationStrategy(steps=self._steps)
