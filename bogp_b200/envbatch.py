"""Geoacoustic-inversion scoring: every candidate environment brings its own host-computed mode set.

Replaces the process pool of ``MatchedFieldProcessor`` calls in ref/projects/swellex96_inv/data/bo/obj.py:28-30,
73-83 (each worker re-runs KRAKEN, then modal sum + Bartlett) by one batched launch: the host mode provider still
runs per candidate (KRAKEN stays external), the modal sum + Bartlett for all candidates runs in
``bogp_envbatch_bartlett`` (HBM-bound: ``Phi_rec`` is streamed once per candidate-tonal).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .modes import ModeSet


def pack_mode_sets(mode_sets: Sequence[ModeSet], src_z: np.ndarray):
    """Mode-major packing padded to ``Mmax`` (layout of ``bogp_envbatch_create``, include/bogp.h).

    ``amp[c, f, m] = phi_m(z_s)`` uses each ModeSet's own depth table and interpolation rule (data contract,
    modes.py) -- this is the mode provider's job, exactly what KRAKEN does when asked for a source depth.
    """
    Cn = len(mode_sets)
    F, N = mode_sets[0].n_freq, mode_sets[0].n_rec
    for ms in mode_sets:
        if ms.n_freq != F or ms.n_rec != N:
            raise ValueError("all candidate environments must share the tonals and the array")
        if any(np.iscomplexobj(p) for p in ms.phi_rec):
            raise NotImplementedError("complex receiver mode shapes are not supported in the env-batch path")
    Mmax = max(max(ms.n_modes) for ms in mode_sets)
    Mmax = (Mmax + 3) // 4 * 4
    k = np.zeros((Cn, F, Mmax), dtype=np.complex128)
    k[...] = 1.0
    amp = np.zeros((Cn, F, Mmax), dtype=np.complex128)
    phiT = np.zeros((Cn, F, Mmax, N), dtype=np.float32)
    for c, ms in enumerate(mode_sets):
        for f in range(F):
            M = ms.n_modes[f]
            k[c, f, :M] = ms.k[f]
            amp[c, f, :M] = ms.phi_src(f, float(src_z[c]))[0]
            phiT[c, f, :M, :] = np.asarray(ms.phi_rec[f], dtype=np.float64).T
    return k, amp, phiT, Mmax


class EnvBatch:
    """``bogp_envbatch_t`` handle over C candidate environments."""

    def __init__(self, mode_sets: Sequence[ModeSet], rec_r_km: np.ndarray, src_z: np.ndarray,
                 covariance_matrix: Optional[np.ndarray] = None, tilt_deg: Optional[np.ndarray] = None):
        rec_r_km = np.atleast_1d(np.asarray(rec_r_km, dtype=np.float64))
        src_z = np.atleast_1d(np.asarray(src_z, dtype=np.float64))
        if not (len(mode_sets) == len(rec_r_km) == len(src_z)):
            raise ValueError("one range and one source depth per candidate environment")
        k, amp, phiT, Mmax = pack_mode_sets(mode_sets, src_z)
        self.C, self.F, self.N, self.Mmax = len(mode_sets), mode_sets[0].n_freq, mode_sets[0].n_rec, Mmax
        k_re, k_im = np.ascontiguousarray(k.real), np.ascontiguousarray(k.imag)
        a_re = np.ascontiguousarray(amp.real)
        a_im = np.ascontiguousarray(amp.imag) if np.any(amp.imag) else None
        rng = np.ascontiguousarray(1000.0 * rec_r_km)
        self.bytes_streamed = int(phiT.nbytes + 2 * k_re.nbytes + a_re.nbytes // 2)
        self._h = C.c_void_p()
        _lib.call("bogp_envbatch_create", C.byref(self._h), self.C, self.F, self.N, Mmax, _lib.dptr(k_re),
                  _lib.dptr(k_im), _lib.dptr(a_re), _lib.dptr(a_im), _lib.fptr(phiT), _lib.dptr(rng))
        self.device = torch.device("cuda", torch.cuda.current_device())
        if tilt_deg is not None and np.any(np.asarray(tilt_deg, dtype=np.float64) != 0.0):
            # per-candidate array tilt (search key `tilt`, swellex96_inv/data/bo/common.py:63): levers of every candidate's own array
            tilt = np.ascontiguousarray(np.broadcast_to(np.asarray(tilt_deg, dtype=np.float64), (self.C,)))
            lever = np.ascontiguousarray(np.stack([ms.z_pivot - ms.rec_z for ms in mode_sets]), dtype=np.float64)
            _lib.call("bogp_envbatch_set_tilt", self._h, _lib.dptr(tilt), _lib.dptr(lever))
        if covariance_matrix is not None:
            self.set_covariance(covariance_matrix)

    def set_covariance(self, K: np.ndarray) -> None:
        K = np.asarray(K)
        if K.shape != (self.F, self.N, self.N):
            raise ValueError(f"covariance_matrix must be [{self.F}, {self.N}, {self.N}], got {K.shape}")
        K_re = np.ascontiguousarray(K.real, dtype=np.float64)
        K_im = np.ascontiguousarray(K.imag if np.iscomplexobj(K) else np.zeros_like(K_re), dtype=np.float64)
        _lib.call("bogp_envbatch_set_covariance", self._h, _lib.dptr(K_re), _lib.dptr(K_im))

    def bartlett(self, *, atype: str = "cbf", multifreq_method: str = "mean",
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty(self.C, dtype=torch.float32, device=self.device)
        _lib.call("bogp_envbatch_bartlett", self._h, _lib.ATYPE[atype], _lib.MF_METHOD[multifreq_method],
                  C.c_void_p(out.data_ptr()), C.c_void_p(_lib.current_stream_ptr()))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().bogp_envbatch_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
