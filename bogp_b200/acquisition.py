"""Boundary B: acquisition functions with the BoTorch ``AcquisitionFunction.forward(X[b, q, d]) -> [b]`` interface.

Mirrors ``ExpectedImprovement(model, best_f)`` (ref/projects/swellex96_inv/data/bo/ei.py:80,
ref/projects/swellex96_localization/data/demo_1d.py:149), ``LogExpectedImprovement`` (logei.py:80, baxus.py:144),
``ProbabilityOfImprovement`` (pi.py:81), ``UpperConfidenceBound(model, beta)`` (ucb.py:82) and the MC
``qExpectedImprovement`` reached through Ax ``Models.GPEI`` (ref/optimization/strategy.py:61).  In-repo formula
pins: ref/projects/swellex96_inv/visualization/bo_examples.py:78-84 (EI), :124 (UCB).

Each call is ONE fused kernel launch (cross-covariances, ``L^-1 k*``, mean/variance, acquisition epilogue and --
when ``X.requires_grad`` -- the analytic ``d acq / d X``) in float64.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib
from .gp import DT, SingleTaskGP


class _AnalyticAcq(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, model, kind, best_f, beta):
        b, q, d = X.shape
        Xc = X.detach().contiguous()
        out = torch.empty(b, dtype=DT, device=X.device)
        need_grad = bool(ctx.needs_input_grad[0])
        grad = torch.empty(b, d, dtype=DT, device=X.device) if need_grad else None
        _lib.call("bogp_gp_acq", model.handle(), _lib.ACQ[kind], float(best_f), float(beta),
                  C.c_void_p(Xc.data_ptr()), b, C.c_void_p(out.data_ptr()),
                  C.c_void_p(grad.data_ptr()) if need_grad else None, C.c_void_p(_lib.current_stream_ptr()))
        ctx.grad = grad
        return out

    @staticmethod
    def backward(ctx, gout):
        if ctx.grad is None:
            return None, None, None, None, None
        return (gout.unsqueeze(-1) * ctx.grad).unsqueeze(1), None, None, None, None


class _QEI(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, model, base, best_f):
        b, q, d = X.shape
        Xc = X.detach().contiguous()
        out = torch.empty(b, dtype=DT, device=X.device)
        need_grad = bool(ctx.needs_input_grad[0])
        grad = torch.empty(b, q, d, dtype=DT, device=X.device) if need_grad else None
        _lib.call("bogp_gp_qei", model.handle(), C.c_void_p(Xc.data_ptr()), b, q, C.c_void_p(base.data_ptr()),
                  base.shape[0], float(best_f), C.c_void_p(out.data_ptr()),
                  C.c_void_p(grad.data_ptr()) if need_grad else None, C.c_void_p(_lib.current_stream_ptr()))
        ctx.grad = grad
        return out

    @staticmethod
    def backward(ctx, gout):
        if ctx.grad is None:
            return None, None, None, None
        return gout.reshape(-1, 1, 1) * ctx.grad, None, None, None


class AcquisitionFunction:
    """``acq(X[b, q, d]) -> [b]`` on the model's CUDA device (inputs on other devices are moved there)."""

    kind = ""
    nonneg = False       # optimize_acqf picks initialize_q_batch_nonneg for these [recalled]

    def __init__(self, model: SingleTaskGP):
        self.model = model

    def _prep(self, X: torch.Tensor) -> torch.Tensor:
        self.model.handle()
        X = torch.as_tensor(X)
        if X.ndim == 2:
            X = X.unsqueeze(1)
        if X.device != self.model.device or X.dtype != DT:
            X = X.to(device=self.model.device, dtype=DT)
        return X

    def __call__(self, X):
        return self.forward(X)

    def forward(self, X):
        raise NotImplementedError


class _Analytic(AcquisitionFunction):
    def __init__(self, model: SingleTaskGP, best_f: float = 0.0, beta: float = 0.0, maximize: bool = True):
        super().__init__(model)
        if not maximize:
            raise NotImplementedError("the reference maximises (it negates the objective, ei.py:49); maximize=False "
                                      "is not implemented")
        self.best_f = float(best_f)
        self.beta = float(beta)

    def forward(self, X):
        X = self._prep(X)
        if X.shape[-2] != 1:
            raise ValueError("analytic acquisition functions expect q = 1 (X[b, 1, d])")
        return _AnalyticAcq.apply(X, self.model, self.kind, self.best_f, self.beta)


    def value_and_grad_host(self, X: "np.ndarray", need_grad: bool = True):
        """``X[b, 1, d]`` float64 host array -> ``(values[b], grad[b, 1, d] or None)`` as host arrays through ONE call of
        ``bogp_gp_acq_host`` (what :func:`bogp_b200.optim.optimize_acqf` uses inside L-BFGS-B)."""
        import numpy as np
        X = np.ascontiguousarray(X, dtype=np.float64)
        b = X.shape[0]
        if X.ndim != 3 or X.shape[1] != 1:
            raise ValueError("analytic acquisition functions expect q = 1 (X[b, 1, d])")
        out = np.empty(b)
        grad = np.empty_like(X) if need_grad else None
        _lib.call("bogp_gp_acq_host", self.model.handle(), _lib.ACQ[self.kind], self.best_f, self.beta, _lib.dptr(X), b,
                  _lib.dptr(out), _lib.dptr(grad), C.c_void_p(_lib.current_stream_ptr()))
        return out, grad


class ExpectedImprovement(_Analytic):
    kind, nonneg = "ei", True

    def __init__(self, model, best_f, maximize: bool = True):
        super().__init__(model, best_f=best_f, maximize=maximize)


class LogExpectedImprovement(_Analytic):
    kind = "logei"

    def __init__(self, model, best_f, maximize: bool = True):
        super().__init__(model, best_f=best_f, maximize=maximize)


class ProbabilityOfImprovement(_Analytic):
    kind, nonneg = "pi", True

    def __init__(self, model, best_f, maximize: bool = True):
        super().__init__(model, best_f=best_f, maximize=maximize)


class UpperConfidenceBound(_Analytic):
    kind = "ucb"

    def __init__(self, model, beta, maximize: bool = True):
        super().__init__(model, beta=beta, maximize=maximize)


def sobol_normal_samples(q: int, n: int, seed: int) -> torch.Tensor:
    """Base samples of botorch ``SobolQMCNormalSampler`` [recalled]: scrambled Sobol through the inverse cdf."""
    eng = torch.quasirandom.SobolEngine(dimension=q, scramble=True, seed=seed)
    u = eng.draw(n, dtype=DT)
    v = 0.5 + (1.0 - 1e-10) * (u - 0.5)
    return math.sqrt(2.0) * torch.erfinv(2.0 * v - 1.0)


class qExpectedImprovement(AcquisitionFunction):
    """MC qEI: ``mean_s max_j relu(mu_j + (L z_s)_j - best_f)`` with fixed QMC base samples ``z[S, q]``."""

    kind, nonneg = "qei", True

    def __init__(self, model: SingleTaskGP, best_f: float, q: int = 1, mc_samples: int = 512, seed: int = 0,
                 base_samples: Optional[torch.Tensor] = None):
        super().__init__(model)
        self.best_f = float(best_f)
        self.q = int(q)
        model.handle()
        base = sobol_normal_samples(q, mc_samples, seed) if base_samples is None else torch.as_tensor(base_samples, dtype=DT)
        self.base_samples = base.reshape(-1, q).to(model.device).contiguous()

    def forward(self, X):
        X = self._prep(X)
        if X.shape[-2] != self.q:
            raise ValueError(f"expected X[b, q={self.q}, d]")
        return _QEI.apply(X, self.model, self.base_samples, self.best_f)

    def value_and_grad_host(self, X: "np.ndarray", need_grad: bool = True):
        """Host-array variant (``bogp_gp_qei_host``), see :meth:`_Analytic.value_and_grad_host`."""
        import numpy as np
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 3 or X.shape[1] != self.q:
            raise ValueError(f"expected X[b, q={self.q}, d]")
        b = X.shape[0]
        out = np.empty(b)
        grad = np.empty_like(X) if need_grad else None
        _lib.call("bogp_gp_qei_host", self.model.handle(), _lib.dptr(X), b, self.q, C.c_void_p(self.base_samples.data_ptr()),
                  self.base_samples.shape[0], self.best_f, _lib.dptr(out), _lib.dptr(grad),
                  C.c_void_p(_lib.current_stream_ptr()))
        return out, grad
