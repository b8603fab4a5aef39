"""Oracle: covariance, Bartlett processors and the MatchedFieldProcessor loop.

Test infrastructure only (see oracle/__init__.py).  Restates SURVEY 8a a1, a3, a4:

* ``covariance`` / ``simulate_covariance``: ``K = d d^H`` with ``d <- d/||d||``
  (ref/optimization/utils.py:23-32; ref/projects/swellex96_localization/data/acoustics.py:116-124).
* ``beamformer(K, p, atype)``: ``cbf`` = ``w^H K w`` with ``w = p/||p||`` per range
  column (= ``|w^H d|^2 / (|w|^2 |d|^2)`` of the north_star when ``K = d d^H``,
  ``tr K = 1``); ``cbf_ml`` = ``tr(K) - w^H K w`` (minimised:
  ref/projects/swellex96_loc_tilt/conf/configure.py:71-76, plotted as ``1 - value``
  ref/projects/swellex96_inv/data/bo/sensitivity.py:118).
* ``MatchedFieldProcessor``: per tonal ``runner -> beamformer`` then ``mean`` or
  ``product`` over tonals (call sites ref/projects/swellex96_localization/data/mfp.py:198-214,
  ref/projects/swellex96_inv/data/bo/obj.py:62-70).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np

from . import field as _field


def covariance(d: np.ndarray) -> np.ndarray:
    """``K = d d^H`` for a data vector ``d [N]`` or ``[N, 1]``."""
    d = np.asarray(d, dtype=np.complex128).reshape(-1, 1)
    return d @ d.conj().T


def simulate_covariance(runner: Callable[[dict], np.ndarray], parameters: dict,
                        freq: Sequence[float]) -> np.ndarray:
    """``K[F, N, N]`` from the forward model at known truth (utils.py:23-32)."""
    K = []
    for f in freq:
        p = np.array(runner(parameters | {"freq": f}), dtype=np.complex128)
        p = p / np.linalg.norm(p)
        K.append(covariance(p))
    return np.array(K)


def beamformer(K: np.ndarray, replica: np.ndarray, atype: str = "cbf") -> np.ndarray:
    """Bartlett power per range column.  K [N, N], replica [N, Nr] -> [Nr] float64."""
    K = np.asarray(K, dtype=np.complex128)
    p = np.asarray(replica, dtype=np.complex128)
    if p.ndim == 1:
        p = p[:, None]
    w = p / np.linalg.norm(p, axis=0, keepdims=True)
    B = np.real(np.sum(w.conj() * (K @ w), axis=0))
    if atype == "cbf":
        return B
    if atype == "cbf_ml":
        return np.real(np.trace(K)) - B
    raise ValueError(f"unknown beamformer atype {atype!r}")


class MatchedFieldProcessor:
    """Literal restatement of ``tritonoa.sp.mfp.MatchedFieldProcessor`` [ext] semantics."""

    def __init__(self, runner: Callable[[dict], np.ndarray], covariance_matrix, freq,
                 parameters: dict, parameter_formatter: Optional[Callable] = None,
                 beamformer: Callable = beamformer, multifreq_method: str = "mean"):
        self.runner = runner
        self.covariance_matrix = covariance_matrix
        self.freq = list(freq)
        self.parameters = dict(parameters)
        self.parameter_formatter = parameter_formatter
        self.beamformer = beamformer
        if multifreq_method not in ("mean", "product"):
            raise ValueError(f"unknown multifreq_method {multifreq_method!r}")
        self.multifreq_method = multifreq_method

    def __call__(self, parameters: dict):
        per_f = []
        for f, K in zip(self.freq, self.covariance_matrix):
            title = f"{f:.1f}Hz"
            if self.parameter_formatter is not None:
                merged = self.parameter_formatter(freq=f, title=title,
                                                  fixed_parameters=self.parameters,
                                                  search_parameters=parameters)
            else:
                merged = self.parameters | {"freq": f, "title": title} | parameters
            per_f.append(self.beamformer(K, self.runner(merged)))
        per_f = np.array(per_f)
        out = per_f.mean(axis=0) if self.multifreq_method == "mean" else per_f.prod(axis=0)
        return float(out[0]) if out.shape == (1,) else out


def make_runner(modes) -> Callable[[dict], np.ndarray]:
    """``run_kraken`` stand-in over a ModeSet: dict(freq, src_z, rec_r[km][, tilt]) -> p[N, Nr]."""
    freq_index = {float(f): i for i, f in enumerate(modes.freq)}

    def runner(parameters: dict) -> np.ndarray:
        f = freq_index[float(parameters["freq"])]
        return _field.replica(modes, f, float(parameters["src_z"]), parameters["rec_r"],
                              float(parameters.get("tilt", 0.0) or 0.0))

    return runner


def ambiguity_surface(modes, K: np.ndarray, r_km: np.ndarray, z: np.ndarray, *,
                      tilt_deg: float = 0.0, atype: str = "cbf",
                      multifreq_method: str = "mean") -> np.ndarray:
    """``[Nz, Nr]`` surface with the reference's loop order (mfp.py:206-214): z rows outer,
    tonals inner, one vector-``rec_r`` call per row."""
    from functools import partial
    mfp = MatchedFieldProcessor(make_runner(modes), K, modes.freq, {"tilt": tilt_deg},
                                beamformer=partial(beamformer, atype=atype),
                                multifreq_method=multifreq_method)
    amb = np.zeros((len(z), len(r_km)))
    for zz, zs in enumerate(z):
        amb[zz, :] = mfp({"src_z": float(zs), "rec_r": np.asarray(r_km)})
    return amb


def bartlett_points(modes, K: np.ndarray, cand: np.ndarray, *, atype: str = "cbf",
                    multifreq_method: str = "mean") -> np.ndarray:
    """Scattered candidates ``cand[C, 3] = (r_km, z, tilt_deg)`` -> ``[C]``."""
    runner = make_runner(modes)
    out = np.empty(len(cand))
    for c, (r, zs, tau) in enumerate(np.asarray(cand, dtype=np.float64)):
        per_f = [beamformer(K[i], runner({"freq": f, "src_z": zs, "rec_r": r, "tilt": tau}), atype)[0]
                 for i, f in enumerate(modes.freq)]
        out[c] = np.mean(per_f) if multifreq_method == "mean" else np.prod(per_f)
    return out


def bartlett_envbatch(mode_sets, K: np.ndarray, cand: np.ndarray, *, atype: str = "cbf",
                      multifreq_method: str = "mean") -> np.ndarray:
    """Config 5: candidate ``c`` has its own ModeSet ``mode_sets[c]``; ``cand[C,3]`` as above."""
    out = np.empty(len(cand))
    for c, ms in enumerate(mode_sets):
        out[c] = bartlett_points(ms, K, np.asarray(cand)[c:c + 1], atype=atype,
                                 multifreq_method=multifreq_method)[0]
    return out


def _surface_rows(args):
    modes, K, r_km, z, kw = args
    return ambiguity_surface(modes, K, r_km, z, **kw)


def ambiguity_surface_pool(modes, K: np.ndarray, r_km: np.ndarray, z: np.ndarray, workers: Optional[int] = None,
                           **kw) -> np.ndarray:
    """:func:`ambiguity_surface` with the depth rows spread over a process pool, the way the reference spreads its time
    steps (ref/projects/swellex96_localization/data/mfp.py:166-177): the full 1000 x 1000 x 13 surface of BASELINE
    configs[1] takes a few seconds on 16 cores.  One BLAS thread per worker."""
    import os
    from concurrent.futures import ProcessPoolExecutor
    workers = workers or os.cpu_count() or 1
    z = np.asarray(z, dtype=np.float64)
    chunks = [c for c in np.array_split(np.arange(len(z)), min(len(z), 4 * workers)) if len(c)]
    old = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS")}
    for k in old:
        os.environ[k] = "1"
    try:
        import multiprocessing as mp
        with ProcessPoolExecutor(workers, mp_context=mp.get_context("spawn")) as ex:
            parts = list(ex.map(_surface_rows, [(modes, K, np.asarray(r_km), z[c], kw) for c in chunks]))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return np.concatenate(parts, axis=0)


def parity_stats(got: np.ndarray, ref: np.ndarray, floor: float = 0.05) -> dict:
    """Relative-error distribution of a surface against the float64 reference: un-floored (|got - ref| / |ref|, the
    literal reading of "within 1e-5 relative") and floored at `floor` * max|ref| (sidelobes far below the peak carry
    the peak's absolute FP32 noise)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    d = np.abs(got - ref)
    un = d / np.maximum(np.abs(ref), 1e-300)
    fl = d / np.maximum(np.abs(ref), floor * np.abs(ref).max())
    return {"max_unfloored": float(un.max()), "p99_unfloored": float(np.quantile(un, 0.99)),
            "p999_unfloored": float(np.quantile(un, 0.999)), "frac_unfloored_gt_1e-5": float((un > 1e-5).mean()),
            "max_floored": float(fl.max()), "p99_floored": float(np.quantile(fl, 0.99)), "max_abs": float(d.max()),
            "min_ref": float(np.abs(ref).min()), "same_argmax": bool(int(got.argmax()) == int(ref.argmax()))}
