"""covariance front-end, on-disk formats and the time-step sweep of ambiguity surfaces.

What the reference does around the hot path for a whole towed-source event:

* ``compute_covariance_worker`` (ref/projects/swellex96_localization/data/acoustics.py:116-124): per tonal, per time
  step ``K[t] = d d^H`` with ``d = p[t] / ||p[t]||`` from the FFT-bin data ``data.npy [T, N]``, saved as
  ``<f:.1f>Hz/covariance.npy [T, N, N]`` complex128; the tilt / inversion pipelines average ``covariance_averaging``
  snapshots and normalise (ref/projects/swellex96_loc_tilt/conf/data/acoustics.yaml:38-56);
* ``load_covariance`` (ref/projects/swellex96_localization/data/mfp.py:111-121): ``K[F, T, N, N]``;
* ``run_parameterizations`` / ``worker`` (mfp.py:131-216): a ``ProcessPoolExecutor`` over the T time steps, each worker
  building a ``MatchedFieldProcessor`` from ``K[:, t]`` and looping over the source depths, ``surface_<t:03d>.npy``
  ``[Nz, Nr]`` float64 and ``grid_parameters.pkl`` on disk.

Here the pool over time steps becomes the multi-GPU axis (time steps are independent: rank ``r`` of ``world`` takes a
contiguous block, no collective), and every surface is one fused launch on a mode set that stays resident; only the
covariance changes between steps (``bogp_covariance_set``: host factorisation of ``Phi^H K_t Phi``, set-up cost).
The covariance front-end is host NumPy: it is a one-off preprocessing product, not part of the scored path.
"""
from __future__ import annotations

import pickle
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .sharding import slab


# ------------------------------------------------------------------------------------------------ covariance front-end
def snapshot_covariance(p: np.ndarray, averaging: int = 1, normalize: bool = True) -> np.ndarray:
    """``data[T, N]`` (one FFT bin per time step and element) -> ``K[T', N, N]`` complex128.

    ``averaging == 1`` is acoustics.py:116-124 exactly (``d /= ||d||; K[t] = d d^H``).  ``averaging = a > 1`` averages
    the outer products of ``a`` consecutive snapshots per output step (``T' = T // a``) and, with ``normalize``,
    scales to unit trace (``covariance_averaging`` / ``normalize_covariance`` of loc_tilt acoustics.yaml:55-56).
    """
    p = np.asarray(p, dtype=np.complex128)
    if p.ndim != 2:
        raise ValueError("data must be [T, N]")
    a = int(averaging)
    if a < 1:
        raise ValueError("averaging must be >= 1")
    T = p.shape[0] // a
    K = np.zeros((T, p.shape[1], p.shape[1]), dtype=np.complex128)
    for t in range(T):
        for s in range(a):
            d = p[t * a + s].reshape(-1, 1)
            if a == 1 or normalize:
                nrm = np.linalg.norm(d)
                d = d / nrm if nrm > 0 else d
            K[t] += d @ d.conj().T
        if a > 1:
            K[t] /= a
            if normalize:
                tr = np.real(np.trace(K[t]))
                K[t] = K[t] / tr if tr > 0 else K[t]
    return K


def save_covariance(K: np.ndarray, path, freq: float) -> Path:
    """``<path>/<f:.1f>Hz/covariance.npy`` (acoustics.py:124, utils.py:35-39)."""
    d = Path(path) / f"{freq:.1f}Hz"
    d.mkdir(parents=True, exist_ok=True)
    np.save(d / "covariance.npy", np.asarray(K, dtype=np.complex128))
    return d / "covariance.npy"


def load_covariance(path, frequencies: Sequence[float], num_segments: Optional[int] = None,
                    num_elements: Optional[int] = None) -> np.ndarray:
    """mfp.py:111-121: ``K[F, T, N, N]`` from ``<path>/<f:.1f>Hz/covariance.npy``."""
    mats = [np.load(Path(path) / f"{f:.1f}Hz" / "covariance.npy") for f in frequencies]
    T = mats[0].shape[0] if num_segments is None else int(num_segments)
    N = mats[0].shape[1] if num_elements is None else int(num_elements)
    K = np.zeros((len(frequencies), T, N, N), dtype=complex)
    for i, m in enumerate(mats):
        if m.shape != (T, N, N):
            raise ValueError(f"{frequencies[i]:.1f}Hz/covariance.npy has shape {m.shape}, expected {(T, N, N)}")
        K[i] = m
    return K


def save_grid_parameters(grid: Dict[str, Iterable[float]], savepath) -> Path:
    """mfp.py:124-128."""
    out = Path(savepath) / "grid_parameters.pkl"
    with open(out, "wb") as fh:
        pickle.dump({k: list(map(float, v)) for k, v in grid.items()}, fh)
    return out


# ------------------------------------------------------------------------------------------------ the sweep
def time_steps_of(rank: int, world: int, n_steps: int) -> range:
    lo, hi = slab(n_steps, rank, world)
    return range(lo, hi)


def ambiguity_sweep(mfp, K: np.ndarray, rec_r: Sequence[float], src_z: Sequence[float], *, savepath=None,
                    rank: int = 0, world: int = 1, keep_surfaces: bool = True, engine: str = "auto",
                    timesteps: Optional[Sequence[int]] = None) -> Dict[str, object]:
    """All time steps of an event: ``K[F, T, N, N]`` -> one ``[Nz, Nr]`` surface per step (mfp.py:131-216).

    ``mfp``: a :class:`bogp_b200.mfp.MatchedFieldProcessor` (its covariance is replaced step by step; mode tables stay
    on the device).  Rank ``rank`` of ``world`` evaluates its contiguous block of time steps; nothing is exchanged
    (the reference's pool over time steps, mfp.py:166-177).  With ``savepath`` each surface is written as
    ``surface_<t:03d>.npy`` float64 like mfp.py:190 and ``grid_parameters.pkl`` once (rank 0).

    Returns ``{"timesteps": [...], "surfaces": [T_local, Nz, Nr] float32 or None, "peak": [(value, (jz, jr)), ...]}``;
    ``peak`` is ``np.unravel_index(np.argmax(surface))`` per step (collate.py:50), found on the device.
    """
    K = np.asarray(K)
    if K.ndim != 4 or K.shape[0] != len(mfp.freq):
        raise ValueError("K must be [F, T, N, N] with one block per frequency of the processor")
    T = K.shape[1]
    steps = list(time_steps_of(rank, world, T)) if timesteps is None else [int(t) for t in timesteps]
    rec_r = np.asarray(rec_r, dtype=np.float64)
    src_z = np.asarray(src_z, dtype=np.float64)
    if savepath is not None:
        savepath = Path(savepath)
        savepath.mkdir(parents=True, exist_ok=True)
        if rank == 0:
            save_grid_parameters({"rec_r": rec_r, "src_z": src_z}, savepath)
    dms = mfp._device_modes()
    surfaces = np.empty((len(steps), len(src_z), len(rec_r)), dtype=np.float32) if keep_surfaces else None
    buf = np.empty((len(src_z), len(rec_r)), dtype=np.float32)
    peaks: List[Tuple[float, Tuple[int, int]]] = []
    for k, t in enumerate(steps):
        dms.set_covariance(np.ascontiguousarray(K[:, t]))
        out = surfaces[k] if keep_surfaces else buf
        _, peak = dms.bartlett_grid_host(rec_r, src_z, atype=mfp.atype, multifreq_method=mfp.multifreq_method,
                                         engine=engine, out=out, argext="min" if mfp.atype == "cbf_ml" else "max")
        peaks.append(peak)
        if savepath is not None:
            np.save(savepath / f"surface_{t:03d}.npy", out.astype(np.float64))
    mfp.covariance_matrix = np.ascontiguousarray(K[:, steps[-1]]) if steps else mfp.covariance_matrix
    return {"timesteps": steps, "surfaces": surfaces, "peak": peaks}


def gather_peaks(result: Dict[str, object], group=None) -> Dict[int, Tuple[float, Tuple[int, int]]]:
    """Merge the per-rank ``ambiguity_sweep`` results into ``{time step: (value, (jz, jr))}`` on every rank: the one
    exchange of the multi-GPU sweep (a few bytes per time step; ``torch.distributed.all_gather_object`` over NCCL or
    gloo).  Single-process runs return the local dictionary."""
    import torch.distributed as dist
    local = list(zip(result["timesteps"], result["peak"]))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local)
    parts: List[object] = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    merged: Dict[int, Tuple[float, Tuple[int, int]]] = {}
    for part in parts:
        merged.update(dict(part))
    return dict(sorted(merged.items()))
