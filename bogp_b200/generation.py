"""Boundary B: sampling strategies with the BoTorch ``botorch.generation`` interface.

``MaxPosteriorSampling(model, replacement=False)(X_cand, num_samples=1)`` is what BAxUS's ``create_candidate`` calls
(ref/projects/swellex96_inv/data/bo/baxus.py:136-138) on ``n_candidates = min(5000, max(2000, 200 d))`` points
(:45): draw ONE joint sample of the GP posterior over all candidates and return the candidate where it is largest.
The joint draw needs the Cholesky factor of the full ``b x b`` posterior covariance; here the covariance, its blocked
factorisation, ``f = mu + L z`` and the arg-max all run on the device in float64 (``bogp_gp_thompson``,
csrc/kernels_ts.cu) -- only the chosen indices come back.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .gp import DT, SingleTaskGP, _cuda_device


class MaxPosteriorSampling:
    """Thompson sampling over a discrete candidate set.

    ``forward(X_cand[b, d], num_samples=S) -> X_next[S, d]``.  ``replacement=False`` (the reference's setting): sample
    ``s`` picks its arg-max among the candidates no earlier sample picked.  The standard-normal base samples come from
    ``torch.randn`` on the global generator unless ``base_samples[S, b]`` is given (tests pin them).
    After a call, ``last_indices`` / ``last_values`` / ``last_jitter`` describe what was picked.
    """

    def __init__(self, model: SingleTaskGP, replacement: bool = False):
        if replacement:
            raise NotImplementedError("MaxPosteriorSampling(replacement=True) is not used by the reference (baxus.py:136)")
        self.model = model
        self.replacement = replacement
        self.last_indices: Optional[torch.Tensor] = None
        self.last_values: Optional[torch.Tensor] = None
        self.last_jitter: float = 0.0
        self.last_samples: Optional[torch.Tensor] = None

    def __call__(self, X_cand, num_samples: int = 1, base_samples: Optional[torch.Tensor] = None,
                 return_samples: bool = False) -> torch.Tensor:
        return self.forward(X_cand, num_samples, base_samples, return_samples)

    def forward(self, X_cand, num_samples: int = 1, base_samples: Optional[torch.Tensor] = None,
                return_samples: bool = False) -> torch.Tensor:
        dev = _cuda_device()
        Xh = torch.as_tensor(X_cand, dtype=DT)
        if Xh.dim() != 2 or Xh.shape[1] != self.model.train_X.shape[1]:
            raise ValueError(f"X_cand must be [b, {self.model.train_X.shape[1]}], got {tuple(Xh.shape)}")
        b = Xh.shape[0]
        S = int(num_samples)
        if S < 1 or S > b:
            raise ValueError(f"num_samples={S} outside [1, b={b}]")
        Xc = Xh.to(dev).contiguous()
        if base_samples is None:
            z = torch.randn(S, b, dtype=DT).to(dev)
        else:
            z = torch.as_tensor(base_samples, dtype=DT).reshape(S, b).to(dev).contiguous()
        idx = torch.empty(S, dtype=torch.int64, device=dev)
        val = torch.empty(S, dtype=DT, device=dev)
        samples = torch.empty(S, b, dtype=DT, device=dev) if return_samples else None
        jit = C.c_double(0.0)
        _lib.call("bogp_gp_thompson", self.model.handle(), C.c_void_p(Xc.data_ptr()), b, C.c_void_p(z.data_ptr()), S,
                  C.c_void_p(idx.data_ptr()), C.c_void_p(val.data_ptr()),
                  C.c_void_p(samples.data_ptr()) if samples is not None else None, None, C.byref(jit),
                  C.c_void_p(_lib.current_stream_ptr()))
        self.last_indices = idx.cpu()
        self.last_values = val.cpu()
        self.last_jitter = float(jit.value)
        self.last_samples = samples
        return Xh[self.last_indices].to(X_cand.device if torch.is_tensor(X_cand) else "cpu")
