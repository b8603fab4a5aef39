"""Dotted-path shims so the reference's unchanged configs and loops import this package.

The Hydra / hydra-zen configs and the hand-rolled BO loops name their callables by dotted path
(``tritonoa.sp.mfp.MatchedFieldProcessor``, ``tritonoa.sp.beamforming.beamformer``,
``tritonoa.at.models.kraken.runner.run_kraken``, ``oao.optimizer.BayesianOptimization``,
``botorch.models.SingleTaskGP``, ``gpytorch.mlls.ExactMarginalLogLikelihood``, ``botorch.fit.fit_gpytorch_mll`` ... -- ref/projects/swellex96_localization/conf/optimization/run.py:15-21,89-99;
ref/projects/swellex96_inv/data/bo/ei.py:8-18; ref/optimization/strategy.py:5-7).  ``install()`` registers module
objects under those names that resolve to the CUDA-backed mirrors of this package.  A real installation of one of
the packages is never shadowed unless ``force=True``.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from typing import Dict, List

_INSTALLED: List[str] = []


def _present(top: str) -> bool:
    if top in sys.modules and getattr(sys.modules[top], "__bogp_shim__", False):
        return False
    try:
        return importlib.util.find_spec(top) is not None
    except (ImportError, ValueError):
        return False


def _module(name: str, **attrs) -> types.ModuleType:
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        sub = ".".join(parts[:i])
        if sub not in sys.modules or not getattr(sys.modules[sub], "__bogp_shim__", False):
            m = types.ModuleType(sub)
            m.__bogp_shim__ = True
            m.__path__ = []            # behaves as a package
            sys.modules[sub] = m
            _INSTALLED.append(sub)
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)
    mod = sys.modules[name]
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def clean_up_kraken_files(path=None) -> None:
    """``tritonoa.at.models.kraken.kraken.clean_up_kraken_files`` (data/run.py:49): the fused path writes no
    ``.env/.mod/.prt`` files, so there is nothing to remove."""


class _NoOpSetting:
    """``botorch.settings.validate_input_scaling(False)`` (ei.py:61) / ``gpytorch.settings.max_cholesky_size(n)``
    (baxus.py:329): context managers that switch library checks this implementation does not have."""

    def __init__(self, *args, **kwargs):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def install(force: bool = False) -> Dict[str, bool]:
    """Registers the shims; returns ``{top-level package: shimmed?}``."""
    from . import acquisition, generation, gp, mfp, oao, optim, strategy

    done = {}
    if force or not _present("tritonoa"):
        _module("tritonoa.sp.mfp", MatchedFieldProcessor=mfp.MatchedFieldProcessor)
        _module("tritonoa.sp.beamforming", beamformer=mfp.beamformer, covariance=mfp.covariance)
        _module("tritonoa.sp.processing", simulate_covariance=mfp.simulate_covariance)
        _module("tritonoa.at.models.kraken.runner", run_kraken=mfp.run_kraken)
        _module("tritonoa.at.models.kraken.kraken", clean_up_kraken_files=clean_up_kraken_files)
        done["tritonoa"] = True
    else:
        done["tritonoa"] = False
    if force or not _present("oao"):
        _module("oao.space", SearchParameter=oao.SearchParameter, SearchSpace=oao.SearchSpace,
                SearchParameterBounds=oao.SearchParameterBounds, SearchSpaceBounds=oao.SearchSpaceBounds)
        _module("oao.objective", NoiselessFormattedObjective=oao.NoiselessFormattedObjective,
                NoisyFormattedObjective=oao.NoisyFormattedObjective, Properties=oao.Properties, Objective=oao.Objective)
        _module("oao.optimizer", BayesianOptimization=oao.BayesianOptimization, GridSearch=oao.GridSearch,
                QuasiRandom=oao.QuasiRandom)
        _module("oao.results", get_results=oao.get_results)
        _module("oao.strategy", GridStrategy=oao.GridStrategy)
        done["oao"] = True
    else:
        done["oao"] = False
    if force or not _present("ax"):
        _module("ax.modelbridge.generation_strategy", GenerationStep=oao.GenerationStep,
                GenerationStrategy=oao.GenerationStrategy)
        _module("ax.modelbridge.registry", Models=oao.Models)
        done["ax"] = True
    else:
        done["ax"] = False
    if force or not _present("botorch"):
        _module("botorch.models", SingleTaskGP=gp.SingleTaskGP)
        _module("botorch.acquisition", ExpectedImprovement=acquisition.ExpectedImprovement,
                LogExpectedImprovement=acquisition.LogExpectedImprovement,
                ProbabilityOfImprovement=acquisition.ProbabilityOfImprovement,
                UpperConfidenceBound=acquisition.UpperConfidenceBound,
                qExpectedImprovement=acquisition.qExpectedImprovement)
        _module("botorch.acquisition.analytic", ExpectedImprovement=acquisition.ExpectedImprovement,
                LogExpectedImprovement=acquisition.LogExpectedImprovement,
                ProbabilityOfImprovement=acquisition.ProbabilityOfImprovement,
                UpperConfidenceBound=acquisition.UpperConfidenceBound)
        _module("botorch.optim", optimize_acqf=optim.optimize_acqf)
        _module("botorch.generation", MaxPosteriorSampling=generation.MaxPosteriorSampling)
        _module("botorch.fit", fit_gpytorch_mll=gp.fit_gpytorch_mll)
        _module("botorch", fit_gpytorch_mll=gp.fit_gpytorch_mll)                     # `from botorch import fit_gpytorch_mll` (demo_1d.py:6)
        _module("botorch.exceptions", ModelFittingError=gp.ModelFittingError, InputDataWarning=gp.InputDataWarning)
        _module("botorch.settings", validate_input_scaling=_NoOpSetting, debug=_NoOpSetting)
        done["botorch"] = True
    else:
        done["botorch"] = False
    if force or not _present("gpytorch"):
        _module("gpytorch.constraints", Interval=gp.Interval)
        _module("gpytorch.likelihoods", GaussianLikelihood=gp.GaussianLikelihood)
        _module("gpytorch.mlls", ExactMarginalLogLikelihood=gp.ExactMarginalLogLikelihood)
        _module("gpytorch.kernels", MaternKernel=gp.MaternKernel, RBFKernel=gp.RBFKernel, ScaleKernel=gp.ScaleKernel)
        _module("gpytorch.settings", max_cholesky_size=_NoOpSetting, cholesky_jitter=_NoOpSetting)
        done["gpytorch"] = True
    else:
        done["gpytorch"] = False
    return done


def uninstall() -> None:
    for name in reversed(_INSTALLED):
        sys.modules.pop(name, None)
    _INSTALLED.clear()
