"""Oracle: exact-GP posterior, marginal likelihood fit and acquisition functions.

Test infrastructure only (see oracle/__init__.py).  torch-CPU float64 restatement of
SURVEY 8a a5-a7, following the recipe at ref/projects/swellex96_inv/data/bo/ei.py:60-92
(standardised Y, ``GaussianLikelihood(noise_constraint=Interval(1e-8, 1e-3))``,
``SingleTaskGP``, 100 AdamW(lr=0.1) steps on ``-mll``, ``ExpectedImprovement(model,
train_Y.max())``), the explicit Matern-5/2 ARD ``ScaleKernel`` of
ref/projects/swellex96_inv/data/bo/baxus.py:310-319, the RBF-ARD surrogates of
ref/optimization/gpmodels.py:16-57 and the in-repo formula pins
ref/projects/swellex96_inv/visualization/bo_examples.py:78-84 (EI), :124 (UCB).
BoTorch/GPyTorch internals (priors, constraints, clamps) are RECALLED, not verified
(SURVEY App. A); each is named where it is used.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

DT = torch.float64
SQRT5 = math.sqrt(5.0)
MIN_VARIANCE = 1e-10        # gpytorch.settings.min_variance (double) [recalled]
ACQ_MIN_VARIANCE = 1e-12    # botorch analytic acquisitions: sigma = sqrt(clamp_min(var, 1e-12)) [recalled]


def _t(x) -> torch.Tensor:
    return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=DT)


def kernel_matrix(kind: str, X1: torch.Tensor, X2: torch.Tensor, lengthscale: torch.Tensor,
                  outputscale) -> torch.Tensor:
    """``s * k(|| (x - x') / l ||)``; Matern-5/2: ``(1 + sqrt5 r + 5 r^2/3) exp(-sqrt5 r)``, RBF: ``exp(-r^2/2)``."""
    a = X1 / lengthscale
    b = X2 / lengthscale
    diff = a.unsqueeze(-2) - b.unsqueeze(-3)
    d2 = (diff * diff).sum(-1)
    if kind == "rbf":
        return outputscale * torch.exp(-0.5 * d2)
    if kind == "matern52":
        r = torch.sqrt(d2.clamp_min(1e-30))
        return outputscale * (1.0 + SQRT5 * r + (5.0 / 3.0) * d2) * torch.exp(-SQRT5 * r)
    raise ValueError(f"unknown kernel {kind!r}")


def safe_cholesky(A: torch.Tensor) -> torch.Tensor:
    """Cholesky with GPyTorch's jitter retries: +1e-8, 1e-7, 1e-6 on the diagonal [recalled]."""
    L, info = torch.linalg.cholesky_ex(A)
    if not bool(info.any()):
        return L
    eye = torch.eye(A.shape[-1], dtype=A.dtype)
    prev = 0.0
    for i in range(3):
        jit = 1e-8 * 10 ** i
        A = A + (jit - prev) * eye
        prev = jit
        L, info = torch.linalg.cholesky_ex(A)
        if not bool(info.any()):
            return L
    raise np.linalg.LinAlgError("matrix not positive definite after jitter 1e-6")


@dataclass
class GPState:
    """Fitted exact GP (hyper-parameters in constrained space) with cached factors."""

    X: torch.Tensor            # [n, d]
    y: torch.Tensor            # [n]
    kind: str = "matern52"
    lengthscale: torch.Tensor = None   # [d]
    outputscale: float = 1.0
    mean: float = 0.0
    noise: float = 1e-6
    L: torch.Tensor = field(default=None, repr=False)
    Linv: torch.Tensor = field(default=None, repr=False)
    alpha: torch.Tensor = field(default=None, repr=False)

    def __post_init__(self):
        self.X = _t(self.X)
        self.y = _t(self.y).reshape(-1)
        d = self.X.shape[1]
        self.lengthscale = (_t(self.lengthscale).reshape(-1) if self.lengthscale is not None
                            else torch.full((d,), math.log(2.0), dtype=DT))
        if self.lengthscale.numel() == 1:
            self.lengthscale = self.lengthscale.expand(d).clone()
        self.refactor()

    def refactor(self):
        n = self.X.shape[0]
        Kxx = kernel_matrix(self.kind, self.X, self.X, self.lengthscale, self.outputscale)
        self.L = safe_cholesky(Kxx + self.noise * torch.eye(n, dtype=DT))
        self.Linv = torch.linalg.solve_triangular(self.L, torch.eye(n, dtype=DT), upper=False)
        self.alpha = torch.cholesky_solve((self.y - self.mean).unsqueeze(-1), self.L).squeeze(-1)


def posterior(gp: GPState, Xs: torch.Tensor, observation_noise: bool = False):
    """``Xs[..., q, d]`` -> ``(mean[..., q], cov[..., q, q])``: ``mu = m + k*^T alpha``, ``Sigma = K** - V^T V``, ``V = L^-1 k*``."""
    Xs = _t(Xs)
    Ks = kernel_matrix(gp.kind, gp.X, Xs, gp.lengthscale, gp.outputscale)   # [..., n, q]
    mu = gp.mean + (gp.alpha.unsqueeze(-1) * Ks).sum(-2)
    V = gp.Linv @ Ks
    Kss = kernel_matrix(gp.kind, Xs, Xs, gp.lengthscale, gp.outputscale)
    cov = Kss - V.transpose(-1, -2) @ V
    if observation_noise:
        cov = cov + gp.noise * torch.eye(cov.shape[-1], dtype=DT)
    return mu, cov


def thompson_sample(gp: GPState, X_cand, z):
    """Joint posterior draw(s) over a candidate set and their arg-max without replacement.

    BAxUS ``create_candidate(acqf="ts")`` (ref/projects/swellex96_inv/data/bo/baxus.py:121-141): botorch
    ``MaxPosteriorSampling(model, replacement=False)(X_cand, num_samples)`` = ``posterior.rsample`` = ``mu + L z`` with
    ``L = psd_safe_cholesky(Sigma)`` of the noise-free posterior [recalled], then per sample the arg-max over the
    candidates not yet picked.  ``z[S, b]`` standard-normal base samples.  Returns ``(idx[S], f[S, b], jitter)``.
    """
    X_cand = _t(X_cand)
    z = _t(z).reshape(-1, X_cand.shape[0])
    mu, cov = posterior(gp, X_cand)
    cov = 0.5 * (cov + cov.T)
    L, info = torch.linalg.cholesky_ex(cov)
    jitter = 0.0
    if bool(info.any()):
        for i in range(3):
            jitter = 1e-8 * 10 ** i
            L, info = torch.linalg.cholesky_ex(cov + jitter * torch.eye(cov.shape[0], dtype=DT))
            if not bool(info.any()):
                break
        else:
            raise np.linalg.LinAlgError("posterior covariance not positive definite after jitter 1e-6")
    f = mu.unsqueeze(0) + z @ L.T
    picked = []
    for s in range(z.shape[0]):
        fs = f[s].clone()
        if picked:
            fs[torch.tensor(picked)] = -math.inf
        picked.append(int(torch.argmax(fs)))
    return np.asarray(picked, dtype=np.int64), f, jitter


def posterior_mean_var(gp: GPState, Xs: torch.Tensor, observation_noise: bool = False):
    """Marginals: ``(mean[..., q], var[..., q])`` with ``var = clamp_min(k** - ||v||^2, 1e-10)``."""
    Xs = _t(Xs)
    Ks = kernel_matrix(gp.kind, gp.X, Xs, gp.lengthscale, gp.outputscale)
    mu = gp.mean + (gp.alpha.unsqueeze(-1) * Ks).sum(-2)
    V = gp.Linv @ Ks
    var = gp.outputscale - (V * V).sum(-2)
    if observation_noise:
        var = var + gp.noise
    return mu, var.clamp_min(MIN_VARIANCE)


# ----------------------------------------------------------------------------- acquisitions
def _phi(u):
    return torch.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)


def _Phi(u):
    return 0.5 * torch.erfc(-u / math.sqrt(2.0))


def _log1mexp(x):
    """log(1 - exp(x)) for x < 0, stable on both ends."""
    return torch.where(x > -math.log(2.0), torch.log(-torch.expm1(x)), torch.log1p(-torch.exp(x)))


def _erfcx(x):
    return torch.special.erfcx(x)


def log_ei_helper(u: torch.Tensor) -> torch.Tensor:
    """``log(phi(u) + u Phi(u))`` (botorch ``_log_ei_helper`` [recalled])."""
    bound = -1.0
    neg_inv_sqrt_eps = -1e6
    u_upper = torch.where(u < bound, torch.full_like(u, bound), u)
    log_ei_upper = torch.log(_phi(u_upper) + u_upper * _Phi(u_upper))
    u_lower = torch.where(u > bound, torch.full_like(u, bound), u)
    u_eps = torch.where(u_lower < neg_inv_sqrt_eps, torch.full_like(u, neg_inv_sqrt_eps), u_lower)
    w = torch.log(_erfcx(-u_eps / math.sqrt(2.0)) * u_eps.abs()) + 0.5 * math.log(math.pi / 2.0)
    log_phi = -0.5 * u * u - 0.5 * math.log(2.0 * math.pi)
    log_ei_lower = log_phi + torch.where(u > neg_inv_sqrt_eps, _log1mexp(w), -2.0 * torch.log(u_lower.abs()))
    return torch.where(u > bound, log_ei_upper, log_ei_lower)


def analytic_acq(kind: str, mu: torch.Tensor, var: torch.Tensor, best_f: float = 0.0,
                 beta: float = 0.0, maximize: bool = True) -> torch.Tensor:
    """EI / LogEI / PI / UCB from marginal mean and variance (q = 1)."""
    sigma = torch.sqrt(var.clamp_min(ACQ_MIN_VARIANCE))
    if kind == "ucb":
        return (mu if maximize else -mu) + math.sqrt(beta) * sigma
    u = (mu - best_f) / sigma
    if not maximize:
        u = -u
    if kind == "ei":
        return sigma * (_phi(u) + u * _Phi(u))
    if kind == "logei":
        return log_ei_helper(u) + torch.log(sigma)
    if kind == "pi":
        return _Phi(u)
    raise ValueError(f"unknown acquisition {kind!r}")


def acq_value(gp: GPState, kind: str, X: torch.Tensor, best_f: float = 0.0, beta: float = 0.0,
              base_samples: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``X[b, q, d] -> [b]``.  ``kind`` in ei|logei|pi|ucb (q = 1) or qei (any q, needs base_samples[S, q])."""
    X = _t(X)
    if kind == "qei":
        mu, cov = posterior(gp, X)
        Ls = torch.stack([safe_cholesky(c) for c in cov.reshape(-1, *cov.shape[-2:])]).reshape(cov.shape)
        z = _t(base_samples)                                         # [S, q]
        samples = mu.unsqueeze(0) + torch.einsum("...ij,sj->s...i", Ls, z)   # [S, b, q]
        return (samples - best_f).clamp_min(0.0).max(dim=-1).values.mean(dim=0)
    if X.shape[-2] != 1:
        raise ValueError("analytic acquisitions expect q = 1")
    mu, var = posterior_mean_var(gp, X)
    return analytic_acq(kind, mu.squeeze(-1), var.squeeze(-1), best_f, beta)


def acq_value_and_grad(gp: GPState, kind: str, X, best_f: float = 0.0, beta: float = 0.0,
                       base_samples=None):
    """Autograd gradient oracle: returns ``(values[b], dvalues/dX [b, q, d])``."""
    X = _t(X).clone().requires_grad_(True)
    v = acq_value(gp, kind, X, best_f, beta, base_samples)
    (g,) = torch.autograd.grad(v.sum(), X)
    return v.detach(), g


def sobol_normal_samples(q: int, n: int, seed: int) -> torch.Tensor:
    """Base samples of botorch ``SobolQMCNormalSampler`` [recalled]: scrambled Sobol through the inverse cdf."""
    eng = torch.quasirandom.SobolEngine(dimension=q, scramble=True, seed=seed)
    u = eng.draw(n, dtype=DT)
    v = 0.5 + (1.0 - 1e-10) * (u - 0.5)
    return math.sqrt(2.0) * torch.erfinv(2.0 * v - 1.0)


# ----------------------------------------------------------------------------- fit
def _softplus_inv(x: float) -> float:
    return math.log(math.expm1(x))


def _gamma_logpdf(x: torch.Tensor, conc: float, rate: float) -> torch.Tensor:
    return conc * math.log(rate) - math.lgamma(conc) + (conc - 1.0) * torch.log(x) - rate * x


@dataclass
class FitConfig:
    """Defaults restate ei.py:66-78 on top of BoTorch<=0.11 ``SingleTaskGP`` priors [recalled]."""

    kind: str = "matern52"
    steps: int = 100
    lr: float = 0.1
    noise_bounds: tuple = (1e-8, 1e-3)
    lengthscale_prior: Optional[tuple] = (3.0, 6.0)     # Gamma(concentration, rate)
    outputscale_prior: Optional[tuple] = (2.0, 0.15)
    weight_decay: float = 1e-2                           # torch.optim.AdamW default
    # Interval(lo, hi) constraints instead of Positive/softplus: baxus.py:310-319 (0.005, 10) / (0.05, 10)
    lengthscale_bounds: Optional[tuple] = None
    outputscale_bounds: Optional[tuple] = None


def _constrained(raw: dict, cfg: "FitConfig"):
    """raw -> (lengthscale, outputscale): softplus (GPyTorch ``Positive``) or ``lo + (hi - lo) sigmoid`` (``Interval``)."""
    def tr(v, b):
        return torch.nn.functional.softplus(v) if b is None else b[0] + (b[1] - b[0]) * torch.sigmoid(v)
    return tr(raw["lengthscale"], cfg.lengthscale_bounds), tr(raw["outputscale"], cfg.outputscale_bounds)


def neg_mll(raw: dict, X: torch.Tensor, y: torch.Tensor, cfg: FitConfig) -> torch.Tensor:
    """``-ExactMarginalLogLikelihood`` = ``-(log N(y | m, K + s2 I) + sum log-priors) / n``."""
    n = X.shape[0]
    ls, os_ = _constrained(raw, cfg)
    lo, hi = cfg.noise_bounds
    noise = lo + (hi - lo) * torch.sigmoid(raw["noise"])
    Kxx = kernel_matrix(cfg.kind, X, X, ls, os_) + noise * torch.eye(n, dtype=DT)
    L = safe_cholesky(Kxx)
    r = (y - raw["mean"]).unsqueeze(-1)
    a = torch.cholesky_solve(r, L)
    ll = -0.5 * (r * a).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * n * math.log(2.0 * math.pi)
    if cfg.lengthscale_prior is not None:
        ll = ll + _gamma_logpdf(ls, *cfg.lengthscale_prior).sum()
    if cfg.outputscale_prior is not None:
        ll = ll + _gamma_logpdf(os_, *cfg.outputscale_prior).sum()
    return -ll / n


def fit_adamw(X, y, cfg: FitConfig = FitConfig()) -> GPState:
    """The hand-rolled fit loop of ei.py:69-78: AdamW over raw parameters initialised at 0."""
    X = _t(X)
    y = _t(y).reshape(-1)
    d = X.shape[1]
    raw = {"lengthscale": torch.zeros(d, dtype=DT, requires_grad=True),
           "outputscale": torch.zeros((), dtype=DT, requires_grad=True),
           "noise": torch.zeros((), dtype=DT, requires_grad=True),
           "mean": torch.zeros((), dtype=DT, requires_grad=True)}
    opt = torch.optim.AdamW(list(raw.values()), lr=cfg.lr, weight_decay=cfg.weight_decay)
    for _ in range(cfg.steps):
        opt.zero_grad()
        loss = neg_mll(raw, X, y, cfg)
        loss.backward()
        opt.step()
    lo, hi = cfg.noise_bounds
    with torch.no_grad():
        ls, os_ = _constrained(raw, cfg)
        return GPState(X, y, cfg.kind,
                       ls.clone(),
                       float(os_),
                       float(raw["mean"]),
                       float(lo + (hi - lo) * torch.sigmoid(raw["noise"])))
