"""Boundary B: exact-GP surrogate with the BoTorch ``Model`` interface, evaluated by libbogp_sm100.so.

Mirrors what the reference uses from BoTorch/GPyTorch (all external to ref/):

* ``SingleTaskGP(X, train_Y, likelihood=GaussianLikelihood(noise_constraint=Interval(1e-8, 1e-3)))`` and the
  hand-rolled fit ``100 x AdamW(lr=0.1)`` on ``-mll`` (ref/projects/swellex96_inv/data/bo/ei.py:66-78);
* the explicit Matern-5/2 ARD ``ScaleKernel`` (ref/projects/swellex96_inv/data/bo/baxus.py:310-319) and the RBF-ARD
  surrogate ``LocalizationGP`` (ref/optimization/gpmodels.py:16-30): ``covar_module`` selects ``"matern52"``/``"rbf"``;
* ``model.posterior(X)`` -> ``.mean [b,q,1]``, ``.variance [b,q,1]``, ``.covariance``, ``.rsample(...)``
  (ref/projects/swellex96_localization/data/demo_1d.py:141; gpmodels.py:27-30).

All posterior arithmetic is float64 on the GPU (``bogp_gp_posterior*``); the n x n factorisation and ``L^-1`` are
computed once per hyper-parameter setting inside ``bogp_gp_create`` ON THE HOST (one thread, float64, jitter retries
1e-8 ... 1e-6: set-up cost, csrc/hostmath.cu).  The hyper-parameter fit is one launch of the
persistent fit kernel ``bogp_gp_fit_adamw`` (all AdamW steps on the device); the GPyTorch-side protocol of the reference's
hand-rolled loops (``model.parameters()``, ``ExactMarginalLogLikelihood``, ``fit_gpytorch_mll``, the explicit
``ScaleKernel(MaternKernel)``) is mirrored below on the same kernels (``bogp_gp_mll_grad``: one evaluation of the fit
objective and its gradient).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib

DT = torch.float64


def _cuda_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("bogp_b200.gp needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


@dataclass
class FitConfig:
    """ei.py:66-78 on top of BoTorch<=0.11 ``SingleTaskGP`` priors [recalled, SURVEY App. A]."""

    steps: int = 100
    lr: float = 0.1
    noise_bounds: Tuple[float, float] = (1e-8, 1e-3)
    lengthscale_prior: Optional[Tuple[float, float]] = (3.0, 6.0)    # Gamma(concentration, rate)
    outputscale_prior: Optional[Tuple[float, float]] = (2.0, 0.15)
    weight_decay: float = 1e-2                                       # torch.optim.AdamW default
    # Interval constraints (value = lo + (hi - lo) sigmoid(raw)) instead of softplus: the explicit kernel of
    # baxus.py:310-319 uses Interval(0.005, 10) on the length scales and Interval(0.05, 10) on the output scale
    lengthscale_bounds: Optional[Tuple[float, float]] = None
    outputscale_bounds: Optional[Tuple[float, float]] = None


class GPPosterior:
    """What the reference reads from ``model.posterior(X)``."""

    def __init__(self, mean: torch.Tensor, variance: torch.Tensor, covariance: Optional[torch.Tensor] = None):
        self.mean = mean.unsqueeze(-1)            # [b, q, 1]
        self.variance = variance.unsqueeze(-1)    # [b, q, 1]
        self._cov = covariance                    # [b, q, q] or None (q = 1)

    @property
    def covariance(self) -> torch.Tensor:
        if self._cov is None:
            return self.variance.squeeze(-1).unsqueeze(-1)   # q = 1: [b, 1, 1]
        return self._cov

    @property
    def stddev(self) -> torch.Tensor:
        return self.variance.sqrt()

    def confidence_region(self):
        """``(mean - 2 std, mean + 2 std)`` as GPyTorch's ``MultivariateNormal.confidence_region`` (demo_1d.py:143)."""
        m, s2 = self.mean.squeeze(-1), 2.0 * self.stddev.squeeze(-1)
        return m - s2, m + s2

    def rsample(self, sample_shape=torch.Size([1]), base_samples: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``mean + L z`` with ``L = chol(cov)``; ``base_samples[*sample_shape, b, q, 1]`` standard normal."""
        mean = self.mean.squeeze(-1)                                  # [..., q]
        if base_samples is None:
            z = torch.randn(*sample_shape, *mean.shape, dtype=mean.dtype, device=mean.device)
        else:
            z = base_samples.to(mean)
            if z.shape[-1] == 1 and z.ndim > mean.ndim:
                z = z.squeeze(-1)
        L = torch.linalg.cholesky(self.covariance)                    # [..., q, q]
        return (mean + (L @ z.unsqueeze(-1)).squeeze(-1)).unsqueeze(-1)


class _KernelView:
    def __init__(self, kind: str, lengthscale: torch.Tensor):
        self.kind, self.lengthscale = kind, lengthscale


class _CovarView:
    """Read-only view of a model's kernel in GPyTorch's attribute layout (``ScaleKernel(base_kernel)``)."""

    def __init__(self, kind: str, lengthscale: torch.Tensor, outputscale: torch.Tensor):
        self.base_kernel, self.outputscale = _KernelView(kind, lengthscale), outputscale


class SingleTaskGP:
    """Exact GP: constant mean, ``outputscale * {Matern-5/2 | RBF}`` ARD kernel, homoskedastic noise."""

    num_outputs = 1
    _num_outputs = 1       # gpmodels.py:17

    def __init__(self, train_X, train_Y, covar_module: str = "matern52", *, likelihood=None, lengthscale=None,
                 outputscale: Optional[float] = None, mean: float = 0.0, noise: Optional[float] = None,
                 noise_bounds: Tuple[float, float] = (1e-8, 1e-3)):
        # ``likelihood=GaussianLikelihood(noise_constraint=Interval(lo, hi))`` (ei.py:66-67): only the interval matters
        nc = getattr(likelihood, "noise_constraint", None)
        if nc is not None and hasattr(nc, "lower_bound"):
            noise_bounds = (float(nc.lower_bound), float(nc.upper_bound))
        # ``covar_module=ScaleKernel(MaternKernel(nu=2.5, ard_num_dims=d, lengthscale_constraint=Interval(..)),
        # outputscale_constraint=Interval(..))`` (baxus.py:310-319): an explicit GPyTorch kernel carries no priors, and
        # its Interval constraints replace the softplus transforms of the raw parameters
        self.fit_defaults = FitConfig(noise_bounds=tuple(noise_bounds))
        if isinstance(covar_module, ScaleKernel):
            base = covar_module.base_kernel
            self.fit_defaults = FitConfig(noise_bounds=tuple(noise_bounds), lengthscale_prior=None, outputscale_prior=None,
                                          lengthscale_bounds=_interval(base.lengthscale_constraint),
                                          outputscale_bounds=_interval(covar_module.outputscale_constraint))
            covar_module = base.kind
        if covar_module not in _lib.KERNEL:
            raise ValueError(f"covar_module must be one of {sorted(_lib.KERNEL)} or a ScaleKernel(MaternKernel | RBFKernel)")
        self.likelihood = likelihood if likelihood is not None else GaussianLikelihood()
        self._raw: Optional[torch.nn.Parameter] = None       # raw parameters of a hand-rolled fit (see parameters())
        self._raw_version = -1
        X = torch.as_tensor(train_X, dtype=DT)
        Y = torch.as_tensor(train_Y, dtype=DT).reshape(-1)
        if X.ndim != 2 or X.shape[0] != Y.shape[0]:
            raise ValueError("train_X must be [n, d] and train_Y [n] or [n, 1]")
        self.train_X, self.train_Y = X, Y
        self.kind = covar_module
        d = X.shape[1]
        sp0 = math.log(2.0)                     # softplus(0): GPyTorch raw parameters start at 0 [recalled]
        ls = torch.full((d,), sp0, dtype=DT) if lengthscale is None else torch.as_tensor(lengthscale, dtype=DT).reshape(-1)
        self.lengthscale = ls.expand(d).clone() if ls.numel() == 1 else ls.clone()
        self.outputscale = sp0 if outputscale is None else float(outputscale)
        self.mean_const = float(mean)
        self.noise_bounds = tuple(noise_bounds)
        self.noise = 0.5 * (noise_bounds[0] + noise_bounds[1]) if noise is None else float(noise)
        self._h: Optional[C.c_void_p] = None

    # ------------------------------------------------------------------ device handle
    def _invalidate(self):
        if self._h is not None and self._h.value:
            _lib.lib().bogp_gp_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    def handle(self) -> C.c_void_p:
        self._sync_raw()
        if self._h is None:
            self.device = _cuda_device()
            X = np.ascontiguousarray(self.train_X.cpu().numpy(), dtype=np.float64)
            y = np.ascontiguousarray(self.train_Y.cpu().numpy(), dtype=np.float64)
            ls = np.ascontiguousarray(self.lengthscale.cpu().numpy(), dtype=np.float64)
            h = C.c_void_p()
            _lib.call("bogp_gp_create", C.byref(h), X.shape[0], X.shape[1], _lib.dptr(X), _lib.dptr(y),
                      _lib.KERNEL[self.kind], _lib.dptr(ls), float(self.outputscale), float(self.mean_const),
                      float(self.noise))
            self._h = h
        return self._h

    @property
    def jitter(self) -> float:
        j = C.c_double()
        _lib.call("bogp_gp_jitter", self.handle(), C.byref(j))
        return j.value

    def set_hyperparameters(self, *, lengthscale=None, outputscale=None, mean=None, noise=None) -> "SingleTaskGP":
        if lengthscale is not None:
            ls = torch.as_tensor(lengthscale, dtype=DT).reshape(-1)
            self.lengthscale = ls.expand(self.train_X.shape[1]).clone() if ls.numel() == 1 else ls.clone()
        if outputscale is not None:
            self.outputscale = float(outputscale)
        if mean is not None:
            self.mean_const = float(mean)
        if noise is not None:
            self.noise = float(noise)
        self._invalidate()
        return self

    def fit(self, cfg: Optional["FitConfig"] = None) -> "SingleTaskGP":
        """The reference's hyper-parameter fit (100 x AdamW on -mll, ei.py:69-78); see :func:`fit_gp_adamw`."""
        self._raw = None
        if cfg is None:
            cfg = self.fit_defaults
            cfg = FitConfig(**{**cfg.__dict__, "noise_bounds": tuple(self.noise_bounds)})
        return fit_gp_adamw(self, cfg)

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        self.training = bool(mode)
        return self

    # ------------------------------------------------------------------ hand-rolled fit loops (ei.py:65-78)
    def parameters(self):
        """The raw (unconstrained) hyper-parameters as ONE float64 ``torch.nn.Parameter`` ``[d + 3]`` -- raw
        lengthscale[d], outputscale, noise, mean, initialised at 0 as GPyTorch's raw parameters are -- for
        ``torch.optim.AdamW([{"params": model.parameters()}], lr=0.1)`` (ei.py:69).  Its gradient comes from
        :class:`ExactMarginalLogLikelihood`; the constrained values are read back from it whenever the posterior is
        needed (``_sync_raw``)."""
        if self._raw is None:
            init = getattr(self, "raw_parameters", None)
            t = torch.zeros(self.train_X.shape[1] + 3, dtype=DT) if init is None else torch.as_tensor(init, dtype=DT).clone()
            self._raw = torch.nn.Parameter(t)
            self._raw_version = -1
        return [self._raw]

    @property
    def covar_module(self):
        """What the reference reads from a fitted model's kernel: ``model.covar_module.base_kernel.lengthscale`` -- the
        trust region of ``create_candidate`` is scaled by it (ref/projects/swellex96_inv/data/bo/baxus.py:116) -- and
        ``model.covar_module.outputscale``, as tensors on the training data's device (GPyTorch shapes: ``[1, d]`` and
        ``[]``)."""
        self._sync_raw()
        dev = self.train_X.device
        return _CovarView(self.kind, self.lengthscale.to(dev).reshape(1, -1), torch.tensor(self.outputscale, dtype=DT, device=dev))

    def state_dict(self) -> dict:
        """The raw hyper-parameters, for the warm start of ``initialize_model(x, y, model.state_dict())``
        (ref/projects/swellex96_localization/data/demo_1d.py:92-100,162): ``{"raw_parameters": tensor[d + 3]}``."""
        if self._raw is not None:
            raw = self._raw.detach().clone()
        else:
            init = getattr(self, "raw_parameters", None)
            raw = torch.zeros(self.train_X.shape[1] + 3, dtype=DT) if init is None else torch.as_tensor(init, dtype=DT).clone()
        return {"raw_parameters": raw}

    def load_state_dict(self, state_dict: dict, strict: bool = True) -> "SingleTaskGP":
        """Hyper-parameters of another model of the same input dimension (demo_1d.py:98-99): the next fit starts from
        them, and the posterior uses them until then."""
        raw = state_dict.get("raw_parameters") if isinstance(state_dict, dict) else None
        if raw is None or tuple(raw.shape) != (self.train_X.shape[1] + 3,):
            if strict:
                raise KeyError("load_state_dict: expected {'raw_parameters': tensor[d + 3]} of a model with the same input dimension")
            return self
        with torch.no_grad():
            self.parameters()[0].copy_(torch.as_tensor(raw, dtype=DT))
        self._sync_raw()
        return self

    def _sync_raw(self) -> None:
        r = self._raw
        if r is None or r._version == self._raw_version:
            return
        self._raw_version = r._version
        cfg = self.fit_defaults
        d = self.train_X.shape[1]
        v = r.detach().cpu().numpy().astype(np.float64)
        ls, os_, noise = _constrained(v, d, cfg.lengthscale_bounds, cfg.outputscale_bounds, self.noise_bounds)
        self.raw_parameters = v.copy()
        self.set_hyperparameters(lengthscale=torch.as_tensor(ls), outputscale=float(os_), mean=float(v[d + 2]), noise=float(noise))

    # ------------------------------------------------------------------ posterior
    def _prep(self, X) -> Tuple[torch.Tensor, Tuple[int, ...]]:
        h = self.handle()  # noqa: F841  (sets self.device)
        X = torch.as_tensor(X).to(device=self.device, dtype=DT)
        d = self.train_X.shape[1]
        if X.shape[-1] != d:
            raise ValueError(f"X must have last dimension d={d}")
        if X.ndim == 1:
            X = X.unsqueeze(0)
        if X.ndim == 2:
            X = X.unsqueeze(1)           # [b, d] -> [b, 1, d]
        batch_shape = tuple(X.shape[:-2])
        return X.reshape(-1, X.shape[-2], d).contiguous(), batch_shape

    def posterior(self, X, observation_noise: bool = False, posterior_transform=None) -> GPPosterior:
        """``X[b, q, d]`` (or ``[b, d]``): marginal mean/variance, plus the joint covariance when q > 1."""
        if posterior_transform is not None:
            raise NotImplementedError("posterior_transform is not supported")
        Xf, batch_shape = self._prep(X)
        b, q, d = Xf.shape
        stream = C.c_void_p(_lib.current_stream_ptr())
        if q == 1:
            mean = torch.empty(b, dtype=DT, device=self.device)
            var = torch.empty(b, dtype=DT, device=self.device)
            _lib.call("bogp_gp_posterior", self.handle(), C.c_void_p(Xf.data_ptr()), b, int(observation_noise),
                      C.c_void_p(mean.data_ptr()), C.c_void_p(var.data_ptr()), stream)
            return GPPosterior(mean.reshape(*batch_shape, 1), var.reshape(*batch_shape, 1))
        mean, cov = joint_posterior(self, Xf, observation_noise)
        var = torch.diagonal(cov, dim1=-2, dim2=-1).clamp_min(1e-10)
        return GPPosterior(mean.reshape(*batch_shape, q), var.reshape(*batch_shape, q),
                           cov.reshape(*batch_shape, q, q))

    # gpmodels.py:27-30 ``forward(x) -> MultivariateNormal``: the prior is not on the hot path; expose the posterior.
    # In training mode (``model.train()``, ei.py:70) the output of ``model(X)`` only travels into ``mll(output, y)``,
    # which evaluates the training data itself: nothing is computed here.
    def __call__(self, X):
        if getattr(self, "training", False):
            return None
        return self.posterior(X)


class _JointPosterior(torch.autograd.Function):
    """``(mean[b,q], cov[b,q,q])`` with the analytic backward of ``bogp_gp_posterior_joint_backward``."""

    @staticmethod
    def forward(ctx, X, model, observation_noise):
        b, q, d = X.shape
        mean = torch.empty(b, q, dtype=DT, device=X.device)
        cov = torch.empty(b, q, q, dtype=DT, device=X.device)
        Xc = X.detach().contiguous()
        _lib.call("bogp_gp_posterior_joint", model.handle(), C.c_void_p(Xc.data_ptr()), b, q,
                  C.c_void_p(mean.data_ptr()), C.c_void_p(cov.data_ptr()), C.c_void_p(_lib.current_stream_ptr()))
        if observation_noise:
            cov = cov + model.noise * torch.eye(q, dtype=DT, device=X.device)
        ctx.model = model
        ctx.save_for_backward(Xc)
        return mean, cov

    @staticmethod
    def backward(ctx, gmean, gcov):
        (Xc,) = ctx.saved_tensors
        b, q, d = Xc.shape
        gmean = torch.zeros(b, q, dtype=DT, device=Xc.device) if gmean is None else gmean.contiguous()
        gcov = torch.zeros(b, q, q, dtype=DT, device=Xc.device) if gcov is None else gcov.contiguous()
        grad = torch.empty_like(Xc)
        _lib.call("bogp_gp_posterior_joint_backward", ctx.model.handle(), C.c_void_p(Xc.data_ptr()), b, q,
                  C.c_void_p(gmean.data_ptr()), C.c_void_p(gcov.data_ptr()), C.c_void_p(grad.data_ptr()),
                  C.c_void_p(_lib.current_stream_ptr()))
        return grad, None, None


def joint_posterior(model: SingleTaskGP, X: torch.Tensor, observation_noise: bool = False):
    """Differentiable joint posterior over ``X[b, q, d]`` (CUDA float64)."""
    return _JointPosterior.apply(X, model, observation_noise)


# ---------------------------------------------------------------------------------------------- GPyTorch-side protocol
class Interval:
    """``gpytorch.constraints.Interval(lower, upper)`` (ei.py:66, baxus.py:306-317): the bounds of a constrained
    parameter, value = lower + (upper - lower) sigmoid(raw)."""

    def __init__(self, lower_bound, upper_bound):
        self.lower_bound, self.upper_bound = float(lower_bound), float(upper_bound)


def _interval(c) -> Optional[Tuple[float, float]]:
    return None if c is None else (float(c.lower_bound), float(c.upper_bound))


class GaussianLikelihood:
    """``gpytorch.likelihoods.GaussianLikelihood(noise_constraint=Interval(lo, hi))`` (ei.py:66)."""

    def __init__(self, noise_constraint=None):
        self.noise_constraint = noise_constraint

    def train(self, mode: bool = True):
        return self

    def eval(self):
        return self


class _Kernel:
    def __init__(self, kind: str, ard_num_dims: Optional[int] = None, lengthscale_constraint=None, **_ignored):
        self.kind, self.ard_num_dims, self.lengthscale_constraint = kind, ard_num_dims, lengthscale_constraint


class MaternKernel(_Kernel):
    """``gpytorch.kernels.MaternKernel(nu=2.5, ard_num_dims=d, lengthscale_constraint=Interval(..))`` (baxus.py:308-313);
    only nu = 2.5 has a kernel on the device."""

    def __init__(self, nu: float = 2.5, **kw):
        if float(nu) != 2.5:
            raise NotImplementedError("only the Matern-5/2 kernel (nu=2.5) is implemented")
        super().__init__("matern52", **kw)


class RBFKernel(_Kernel):
    """``gpytorch.kernels.RBFKernel(ard_num_dims=d)`` (ref/optimization/gpmodels.py:22-25)."""

    def __init__(self, **kw):
        super().__init__("rbf", **kw)


class ScaleKernel:
    """``gpytorch.kernels.ScaleKernel(base_kernel, outputscale_constraint=Interval(..))`` (baxus.py:307-318)."""

    def __init__(self, base_kernel: _Kernel, outputscale_constraint=None, **_ignored):
        if not isinstance(base_kernel, _Kernel):
            raise TypeError("ScaleKernel needs a MaternKernel or an RBFKernel")
        self.base_kernel, self.outputscale_constraint = base_kernel, outputscale_constraint


class ModelFittingError(RuntimeError):
    """``botorch.exceptions.ModelFittingError`` (baxus.py:12,344)."""


class InputDataWarning(UserWarning):
    """``botorch.exceptions.InputDataWarning`` (filtered at demo_1d.py:8,187); never raised here."""


def _constrained(raw: np.ndarray, d: int, lsb, osb, noise_bounds):
    """raw -> (lengthscale[d], outputscale, noise): softplus, or the Interval transform where bounds are given."""
    sp = lambda v: np.where(v > 20.0, v, np.log1p(np.exp(np.minimum(v, 20.0))))     # noqa: E731
    iv = lambda v, b: b[0] + (b[1] - b[0]) / (1.0 + np.exp(-v))                      # noqa: E731
    ls = sp(raw[:d]) if lsb is None else iv(raw[:d], lsb)
    os_ = sp(raw[d]) if osb is None else iv(raw[d], osb)
    lo, hi = noise_bounds
    return ls, float(os_), float(lo + (hi - lo) / (1.0 + math.exp(-raw[d + 1])))


def mll_value_and_grad(model: "SingleTaskGP", raw: np.ndarray, cfg: Optional[FitConfig] = None):
    """``(loss, grad[d + 3])`` of the fit objective ``-(log N(y | m, K + s2 I) + log priors) / n`` at the raw
    parameters, by one step of the fit kernels with the update switched off (``bogp_gp_mll_grad``)."""
    cfg = model.fit_defaults if cfg is None else cfg
    X = np.ascontiguousarray(model.train_X.cpu().numpy(), dtype=np.float64)
    y = np.ascontiguousarray(model.train_Y.cpu().numpy(), dtype=np.float64)
    n, d = X.shape
    raw = np.ascontiguousarray(np.asarray(raw, dtype=np.float64).reshape(1, d + 3))
    loss, grad = np.zeros(1), np.zeros((1, d + 3))
    lsp = None if cfg.lengthscale_prior is None else np.array(cfg.lengthscale_prior, dtype=np.float64)
    osp = None if cfg.outputscale_prior is None else np.array(cfg.outputscale_prior, dtype=np.float64)
    lsb = None if cfg.lengthscale_bounds is None else np.array(cfg.lengthscale_bounds, dtype=np.float64)
    osb = None if cfg.outputscale_bounds is None else np.array(cfg.outputscale_bounds, dtype=np.float64)
    lo, hi = model.noise_bounds
    _cuda_device()
    _lib.call("bogp_gp_mll_grad", n, d, _lib.dptr(X), _lib.dptr(y), _lib.KERNEL[model.kind], float(lo), float(hi),
              _lib.dptr(lsp), _lib.dptr(osp), _lib.dptr(lsb), _lib.dptr(osb), 1, _lib.dptr(raw), _lib.dptr(loss),
              _lib.dptr(grad), C.c_void_p(_lib.current_stream_ptr()))
    return float(loss[0]), grad[0]


class _MLL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, model):
        loss, grad = mll_value_and_grad(model, raw.detach().cpu().numpy())
        ctx.grad = torch.as_tensor(grad, dtype=raw.dtype, device=raw.device)
        return torch.tensor(-loss, dtype=raw.dtype, device=raw.device)         # the mll itself: GPyTorch's sign

    @staticmethod
    def backward(ctx, gout):
        return -gout * ctx.grad, None


class ExactMarginalLogLikelihood:
    """``gpytorch.mlls.ExactMarginalLogLikelihood(model.likelihood, model)`` (ei.py:68, baxus.py:326): calling it returns
    the marginal log likelihood per observation plus the log priors, ``(log N(y | m, K + s2 I) + sum log p) / n``, as a
    scalar tensor whose backward fills ``model.parameters()[0].grad`` -- value and gradient both from the fit kernels
    (``bogp_gp_mll_grad``), so the reference's loop ``loss = -mll(model(X), y); loss.backward(); optimizer.step()``
    runs unchanged, one kernel launch per step.  (``model.fit()`` runs the same 100 steps in one launch.)"""

    def __init__(self, likelihood, model: "SingleTaskGP"):
        self.likelihood, self.model = likelihood, model

    def __call__(self, output=None, target=None):
        return _MLL.apply(self.model.parameters()[0], self.model)

    def train(self, mode: bool = True):
        self.model.train(mode)
        return self

    def eval(self):
        self.model.eval()
        return self


def fit_gpytorch_mll(mll: ExactMarginalLogLikelihood, maxiter: int = 200, **_ignored) -> ExactMarginalLogLikelihood:
    """``botorch.fit.fit_gpytorch_mll(mll)`` (demo_1d.py:137, baxus.py:343): scipy L-BFGS-B over the raw parameters from
    their current values, objective and gradient from the fit kernels.  BoTorch's own routine is external to the
    reference and absent here, so its defaults are recalled, not pinned: L-BFGS-B on the unconstrained raw parameters,
    ``maxiter`` 200 [recalled]; raises :class:`ModelFittingError` when the kernel matrix is not positive definite."""
    from scipy.optimize import minimize
    model = mll.model
    raw = model.parameters()[0]

    def fun(v):
        try:
            return mll_value_and_grad(model, v)
        except _lib.BogpError as exc:
            raise ModelFittingError(str(exc)) from exc

    res = minimize(fun, raw.detach().cpu().numpy().astype(np.float64), jac=True, method="L-BFGS-B", options={"maxiter": int(maxiter)})
    with torch.no_grad():
        raw.copy_(torch.as_tensor(res.x, dtype=raw.dtype))
    model._sync_raw()
    return mll


# ---------------------------------------------------------------------------------------------- hyper-parameter fit
def _kernel_matrix(kind: str, X: torch.Tensor, ls: torch.Tensor, os_: torch.Tensor) -> torch.Tensor:
    a = X / ls
    d2 = ((a.unsqueeze(-2) - a.unsqueeze(-3)) ** 2).sum(-1)
    if kind == "rbf":
        return os_ * torch.exp(-0.5 * d2)
    r = torch.sqrt(d2.clamp_min(1e-30))
    return os_ * (1.0 + math.sqrt(5.0) * r + (5.0 / 3.0) * d2) * torch.exp(-math.sqrt(5.0) * r)


def _gamma_logpdf(x, conc, rate):
    return conc * math.log(rate) - math.lgamma(conc) + (conc - 1.0) * torch.log(x) - rate * x


FUSED_FIT_MAX_D = 16       # input dimensions the fused kernels take (kFitMaxD)
FUSED_FIT_MAX_N = 640      # measured (profiles/paths_r2.jsonl): the cooperative kernel beats the cuSOLVER loop up to n ~ 700 (n = 512: 1.3x,
                           # n = 1000: 0.7x); the reference's budgets end at 500 trials (ei.py:22-23)


def fit_gp_adamw(model: SingleTaskGP, cfg: FitConfig = FitConfig(), raw_init: Optional[np.ndarray] = None,
                 return_loss: bool = False, engine: str = "auto"):
    """The fit loop of ei.py:69-78 -- ``steps`` x AdamW over the raw (unconstrained) parameters, initialised at 0 --
    as ONE launch of the persistent fit kernel (``bogp_gp_fit_adamw``: kernel matrix, blocked Cholesky, K^-1,
    analytic gradient of ``-ExactMarginalLogLikelihood = -(log N(y | m, K + s2 I) + sum log-priors)/n`` and the
    AdamW update all on the device, float64).

    ``raw_init[B, d+3]`` (raw lengthscale[d], outputscale, noise, mean) runs B independent fits concurrently and keeps
    the one with the lowest final loss.  ``return_loss`` also returns the loss history ``[B, steps]``.

    ``engine``: ``"fused"`` / ``"auto"`` (the kernels: one CTA with the matrix in shared memory up to n ~ 150, a
    cooperative grid with the matrix in L2 above -- csrc/kernels_gpfit_coop.cu) or ``"torch"``
    (:func:`fit_gp_adamw_torch`, cuSOLVER per step: the timing comparison).  ``auto`` only leaves the kernels for more
    than 16 input dimensions or more than 640 observations (beyond the reference's 500-trial budgets, where the serial
    column chain of the blocked factorisation makes the cooperative kernel slower than cuSOLVER).
    """
    if engine not in ("auto", "fused", "torch"):
        raise ValueError("engine must be 'auto', 'fused' or 'torch'")
    _cuda_device()
    if engine == "torch" or (engine == "auto" and raw_init is None and not return_loss
                             and (model.train_X.shape[0] > FUSED_FIT_MAX_N or model.train_X.shape[1] > FUSED_FIT_MAX_D)):
        return fit_gp_adamw_torch(model, cfg)
    X = np.ascontiguousarray(model.train_X.cpu().numpy(), dtype=np.float64)
    y = np.ascontiguousarray(model.train_Y.cpu().numpy(), dtype=np.float64)
    n, d = X.shape
    raw = np.zeros((1, d + 3)) if raw_init is None else np.array(raw_init, dtype=np.float64).reshape(-1, d + 3)
    raw = np.ascontiguousarray(raw)
    B = raw.shape[0]
    loss = np.zeros((B, cfg.steps))
    lsp = None if cfg.lengthscale_prior is None else np.array(cfg.lengthscale_prior, dtype=np.float64)
    osp = None if cfg.outputscale_prior is None else np.array(cfg.outputscale_prior, dtype=np.float64)
    lo, hi = cfg.noise_bounds
    lsb = None if cfg.lengthscale_bounds is None else np.array(cfg.lengthscale_bounds, dtype=np.float64)
    osb = None if cfg.outputscale_bounds is None else np.array(cfg.outputscale_bounds, dtype=np.float64)
    _lib.call("bogp_gp_fit_adamw_bounded", n, d, _lib.dptr(X), _lib.dptr(y), _lib.KERNEL[model.kind], int(cfg.steps),
              float(cfg.lr), float(cfg.weight_decay), float(lo), float(hi), _lib.dptr(lsp), _lib.dptr(osp),
              _lib.dptr(lsb), _lib.dptr(osb), B, _lib.dptr(raw), _lib.dptr(loss), C.c_void_p(_lib.current_stream_ptr()))
    best = int(np.argmin(loss[:, -1])) if B > 1 else 0
    r = raw[best]
    ls_v, os_v, noise_v = _constrained(r, d, cfg.lengthscale_bounds, cfg.outputscale_bounds, (lo, hi))
    model.noise_bounds = (lo, hi)
    model.set_hyperparameters(lengthscale=torch.as_tensor(ls_v), outputscale=float(os_v), mean=float(r[d + 2]), noise=noise_v)
    model.raw_parameters = r.copy()
    return (model, loss) if return_loss else model


def fit_gp_adamw_torch(model: SingleTaskGP, cfg: FitConfig = FitConfig(), device=None) -> SingleTaskGP:
    """The same loop as torch float64 ops with autograd (cuSOLVER Cholesky per step): kept as the timing comparison
    for the fused kernels (tools/bench_paths.py fit); :func:`fit_gp_adamw` only routes here for d > 16 or n > 640."""
    dev = _cuda_device() if device is None else torch.device(device)
    X = model.train_X.to(dev)
    y = model.train_Y.to(dev)
    n, d = X.shape
    lo, hi = cfg.noise_bounds
    raw = {k: torch.zeros(s, dtype=DT, device=dev, requires_grad=True)
           for k, s in (("lengthscale", (d,)), ("outputscale", ()), ("noise", ()), ("mean", ()))}
    opt = torch.optim.AdamW(list(raw.values()), lr=cfg.lr, weight_decay=cfg.weight_decay)
    eye = torch.eye(n, dtype=DT, device=dev)
    sp = torch.nn.functional.softplus
    lsb, osb = cfg.lengthscale_bounds, cfg.outputscale_bounds
    f_ls = (lambda v: sp(v)) if lsb is None else (lambda v: lsb[0] + (lsb[1] - lsb[0]) * torch.sigmoid(v))
    f_os = (lambda v: sp(v)) if osb is None else (lambda v: osb[0] + (osb[1] - osb[0]) * torch.sigmoid(v))
    for _ in range(cfg.steps):
        opt.zero_grad()
        ls, os_ = f_ls(raw["lengthscale"]), f_os(raw["outputscale"])
        noise = lo + (hi - lo) * torch.sigmoid(raw["noise"])
        Kxx = _kernel_matrix(model.kind, X, ls, os_) + noise * eye
        L, info = torch.linalg.cholesky_ex(Kxx)
        jit = 1e-8
        while bool(info.any()) and jit <= 1e-6:      # gpytorch psd_safe_cholesky retries [recalled]
            L, info = torch.linalg.cholesky_ex(Kxx + jit * eye)
            jit *= 10.0
        r = (y - raw["mean"]).unsqueeze(-1)
        a = torch.cholesky_solve(r, L)
        ll = -0.5 * (r * a).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * n * math.log(2.0 * math.pi)
        if cfg.lengthscale_prior is not None:
            ll = ll + _gamma_logpdf(ls, *cfg.lengthscale_prior).sum()
        if cfg.outputscale_prior is not None:
            ll = ll + _gamma_logpdf(os_, *cfg.outputscale_prior).sum()
        (-ll / n).backward()
        opt.step()
    with torch.no_grad():
        model.noise_bounds = (lo, hi)
        model.set_hyperparameters(lengthscale=f_ls(raw["lengthscale"]).cpu(), outputscale=float(f_os(raw["outputscale"])),
                                  mean=float(raw["mean"]), noise=float(lo + (hi - lo) * torch.sigmoid(raw["noise"])))
    return model
