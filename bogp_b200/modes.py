"""Data contract for the matched-field hot path: precomputed normal modes.

The reference obtains modes by running KRAKEN once per (environment, tonal,
source-depth row) inside ``tritonoa.at.models.kraken.runner.run_kraken`` [ext]
(call sites: ref/projects/swellex96_localization/data/mfp.py:198-214,
ref/projects/swellex96_inv/data/bo/obj.py:62-70).  Eigenvalues ``k_m`` and
mode shapes ``phi_m(z)`` depend only on (environment, tonal), so this build
takes them *precomputed on the host* (north_star: "KRAKEN ... kept external on
the host") in a :class:`ModeSet` and evaluates every candidate on the GPU.

Contract (shared verbatim by ``oracle/`` and the CUDA kernels):

* per tonal ``f``: ``k[f]`` complex128 ``[M_f]``, ``phi_rec[f]`` ``[N, M_f]``
  (real for KRAKEN, complex for KRAKENC), and a source-depth table
  ``phi_src_tab[f]`` ``[nz_tab, M_f]`` sampled on the uniform mesh
  ``z0 + i*dz``; ``phi_m(z_s)`` is *defined* as the linear interpolation of
  that table (KRAKEN tabulates at the requested depth; a
  table + rule makes the number reproducible on both sides);
* receiver geometry ``rec_z [N]`` and ``z_pivot`` for the tilted array;
* ranges enter in **km** at the objective boundary (mfp.py:214 passes
  ``rec_r`` in km) and in metres below it.

The synthetic generator is the Pekeris-like waveguide of SURVEY 8(d): it
stands in for KRAKEN output so that every config is reproducible offline.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

SWELLEX_TONALS_13 = [49.0, 64.0, 79.0, 94.0, 112.0, 130.0, 148.0, 166.0, 201.0,
                     235.0, 283.0, 338.0, 388.0]   # ref/projects/swellex96_loc_tilt/conf/common.py:7-21
INV_TONALS_3 = [148.0, 235.0, 388.0]               # ref/projects/swellex96_inv/data/bo/common.py:7-11
# 21-element VLA of the inversion project (ref/projects/swellex96_inv/data/env.py:51-73)
REC_Z_21 = [94.125, 99.755, 105.38, 111.00, 116.62, 122.25, 127.88, 139.12, 144.74,
            150.38, 155.99, 161.62, 167.26, 172.88, 178.49, 184.12, 189.76, 195.38,
            200.99, 206.62, 212.25]


def rec_z_64() -> np.ndarray:
    """64-element VLA (ref/projects/swellex96_localization/data/env.py:49)."""
    return np.linspace(94.125, 212.25, 64)


@dataclass
class ModeSet:
    """Host-side container of precomputed modes for one environment."""

    freq: List[float]
    k: List[np.ndarray]              # per tonal [M_f] complex128
    phi_rec: List[np.ndarray]        # per tonal [N, M_f] float64 | complex128
    phi_src_tab: List[np.ndarray]    # per tonal [nz_tab, M_f] float64 | complex128
    z0: float
    dz: float
    rec_z: np.ndarray                # [N]
    z_pivot: float = 217.0
    meta: dict = field(default_factory=dict)

    def __post_init__(self):
        self.rec_z = np.asarray(self.rec_z, dtype=np.float64)
        F = len(self.freq)
        if not (len(self.k) == len(self.phi_rec) == len(self.phi_src_tab) == F):
            raise ValueError("ModeSet: per-tonal lists must all have len(freq) entries")
        for f in range(F):
            M = self.k[f].shape[0]
            if self.phi_rec[f].shape != (self.n_rec, M):
                raise ValueError(f"ModeSet: phi_rec[{f}] must be [N={self.n_rec}, M={M}]")
            if self.phi_src_tab[f].ndim != 2 or self.phi_src_tab[f].shape[1] != M:
                raise ValueError(f"ModeSet: phi_src_tab[{f}] must be [nz_tab, M={M}]")
            if self.phi_src_tab[f].shape[0] != self.phi_src_tab[0].shape[0]:
                raise ValueError("ModeSet: all tonals share one depth mesh")

    @property
    def n_freq(self) -> int:
        return len(self.freq)

    @property
    def n_rec(self) -> int:
        return int(self.rec_z.shape[0])

    @property
    def n_modes(self) -> List[int]:
        return [int(k.shape[0]) for k in self.k]

    @property
    def nz_tab(self) -> int:
        return int(self.phi_src_tab[0].shape[0])

    @property
    def is_complex(self) -> bool:
        return any(np.iscomplexobj(p) for p in self.phi_rec) or any(
            np.iscomplexobj(p) for p in self.phi_src_tab)

    def phi_src(self, f: int, z: np.ndarray) -> np.ndarray:
        """``phi_m(z)`` for tonal index ``f``: linear interpolation of the table.

        The rule (shared with the kernels): ``t = (z - z0)/dz`` clamped to
        ``[0, nz_tab-1]``, ``i0 = min(floor(t), nz_tab-2)``, ``w = t - i0``,
        ``phi = (1-w)*tab[i0] + w*tab[i0+1]``.
        """
        z = np.atleast_1d(np.asarray(z, dtype=np.float64))
        tab = self.phi_src_tab[f]
        nz = tab.shape[0]
        t = np.clip((z - self.z0) / self.dz, 0.0, float(nz - 1))
        i0 = np.minimum(np.floor(t).astype(np.int64), nz - 2)
        w = (t - i0)[:, None]
        return (1.0 - w) * tab[i0] + w * tab[i0 + 1]

    def select(self, idx: Sequence[int]) -> "ModeSet":
        """Sub-ModeSet over a subset of tonals."""
        return ModeSet([self.freq[i] for i in idx], [self.k[i] for i in idx],
                       [self.phi_rec[i] for i in idx], [self.phi_src_tab[i] for i in idx],
                       self.z0, self.dz, self.rec_z, self.z_pivot, dict(self.meta))


def pekeris_modes(freq: Sequence[float], rec_z: Optional[Sequence[float]] = None, *,
                  depth: float = 217.0, c_w: float = 1500.0, c_b: float = 1650.0,
                  kappa_max: float = 1e-5, dz_tab: float = 0.25, z_pivot: Optional[float] = None,
                  complex_modes: bool = False) -> ModeSet:
    """Pekeris-like synthetic waveguide; stands in for KRAKEN.

    ``phi_m(z) = sqrt(2/D) sin(kz_m z)``, ``kz_m = (m-1/2) pi / D``,
    ``k_m = sqrt(k0^2 - kz_m^2) + i kappa_m`` for the modes with
    ``Re k_m > omega/c_b``; ``kappa_m = -kappa_max * m / M`` (decaying field
    under the ``e^{-ikr}`` convention).  ``complex_modes=True`` adds a small
    imaginary part to the mode shapes to exercise the KRAKENC layout.
    """
    rec_z = rec_z_64() if rec_z is None else np.asarray(rec_z, dtype=np.float64)
    nz_tab = int(round(depth / dz_tab)) + 1
    z_tab = dz_tab * np.arange(nz_tab)
    ks, prs, pss = [], [], []
    for f in freq:
        omega = 2.0 * np.pi * float(f)
        k0 = omega / c_w
        m_all = np.arange(1, int(k0 * depth / np.pi) + 2)
        kz = (m_all - 0.5) * np.pi / depth
        keep = kz < k0
        kz = kz[keep]
        kr = np.sqrt(k0 * k0 - kz * kz)
        kz = kz[kr > omega / c_b]
        kr = kr[kr > omega / c_b]
        M = kr.shape[0]
        if M == 0:
            raise ValueError(f"no propagating modes at {f} Hz")
        kappa = -kappa_max * np.arange(1, M + 1) / M
        amp = np.sqrt(2.0 / depth)
        pr = amp * np.sin(np.outer(rec_z, kz))
        ps = amp * np.sin(np.outer(z_tab, kz))
        if complex_modes:
            eps = 0.02 * np.arange(1, M + 1) / M
            pr = pr * np.exp(1j * eps)[None, :]
            ps = ps * np.exp(1j * eps)[None, :]
        ks.append((kr + 1j * kappa).astype(np.complex128))
        prs.append(pr)
        pss.append(ps)
    return ModeSet([float(f) for f in freq], ks, prs, pss, 0.0, float(dz_tab), rec_z,
                   float(depth if z_pivot is None else z_pivot),
                   {"kind": "pekeris", "depth": depth, "c_w": c_w, "c_b": c_b})


def perturbed_pekeris_batch(freq: Sequence[float], rec_z: Sequence[float], depths: np.ndarray,
                            c_bs: np.ndarray, **kw) -> List[ModeSet]:
    """One ModeSet per candidate environment (config 5: geoacoustic inversion)."""
    return [pekeris_modes(freq, rec_z, depth=float(d), c_b=float(c), **kw)
            for d, c in zip(depths, c_bs)]


def replica_field(modes: ModeSet, f: int, src_z: float, rec_r_km, tilt_deg: float = 0.0) -> np.ndarray:
    """Replica pressure ``p[N]`` (scalar range) or ``p[N, Nr]`` of tonal index ``f`` for ONE source position: what a
    single ``run_kraken(parameters | {"freq": f})`` call returns to ``simulate_covariance``
    (ref/optimization/utils.py:29-31).  Host float64, set-up only -- the scored path never materialises a field."""
    r = np.asarray(rec_r_km, dtype=np.float64)
    lever = modes.z_pivot - modes.rec_z
    dr = -lever * np.sin(np.deg2rad(float(tilt_deg)))
    phi_s = modes.phi_src(f, src_z)[0]
    rn = 1000.0 * np.atleast_1d(r)[None, :] + dr[:, None]                       # [N, Nr]
    kr = modes.k[f][None, :, None] * rn[:, None, :]                             # [N, M, Nr]
    p_ = np.einsum("nm,m,nmr->nr", modes.phi_rec[f], phi_s, np.exp(-1j * kr) / np.sqrt(kr))
    return p_[:, 0] if r.ndim == 0 else p_


def modeset_from_tables(freq: Sequence[float], k: Sequence[np.ndarray], z: Sequence[np.ndarray],
                        phi: Sequence[np.ndarray], rec_z: Sequence[float], *, z_pivot: Optional[float] = None,
                        dz_tab: float = 0.25, meta: Optional[dict] = None) -> ModeSet:
    """ModeSet from the tables a host normal-mode program hands back: per tonal the eigenvalues ``k[f] [M_f]`` and the
    mode shapes ``phi[f] [nz_f, M_f]`` sampled at the depths ``z[f] [nz_f]`` (the content of a KRAKEN ``.mod`` file as
    TritonOA's reader exposes it: ``modes.k``, ``modes.z``, ``modes.phi``).  Receiver rows are the linear interpolation
    at ``rec_z``; the source table is the same interpolation onto the uniform mesh ``0, dz_tab, ...`` of the data
    contract (module docstring), shared by all tonals."""
    rec_z = np.asarray(rec_z, dtype=np.float64)
    zmax = max(float(np.max(zf)) for zf in z)
    nz_tab = int(np.floor(zmax / dz_tab)) + 1
    z_tab = dz_tab * np.arange(nz_tab)
    ks, prs, pss = [], [], []
    for kf, zf, pf in zip(k, z, phi):
        zf = np.asarray(zf, dtype=np.float64)
        pf = np.asarray(pf)
        if pf.shape != (zf.shape[0], np.asarray(kf).shape[0]):
            raise ValueError("modeset_from_tables: phi[f] must be [len(z[f]), len(k[f])]")
        interp = lambda zq: np.stack([np.interp(zq, zf, pf[:, m].real) + (1j * np.interp(zq, zf, pf[:, m].imag) if np.iscomplexobj(pf) else 0.0)
                                      for m in range(pf.shape[1])], axis=1)
        ks.append(np.asarray(kf, dtype=np.complex128))
        prs.append(interp(rec_z))
        pss.append(interp(z_tab))
    return ModeSet([float(f) for f in freq], ks, prs, pss, 0.0, float(dz_tab), rec_z,
                   float(np.max(rec_z) if z_pivot is None else z_pivot), dict(meta or {}))


def _outer(d: np.ndarray) -> np.ndarray:
    d = np.asarray(d, dtype=np.complex128).reshape(-1, 1)
    return d @ d.conj().T


def synthetic_covariance(modes: ModeSet, src_z: float, rec_r_km: float, tilt_deg: float = 0.0,
                         snapshots: int = 0, snr_db: float = 10.0, seed: int = 0) -> np.ndarray:
    """Synthetic INPUT data: the covariance ``simulate_covariance`` (ref/optimization/utils.py:23-32) would hand to
    the objective -- ``K_f = d d^H``, ``d = p/||p||`` at a known truth -- generated from a ModeSet.

    This is a data generator (like :func:`pekeris_modes`), float64 on the host, one modal sum per tonal; it is
    not part of the scored path.  ``snapshots > 0`` averages that many noisy snapshots
    (rank > 1 covariance, the ``covariance_averaging: 8`` case of loc_tilt/conf/data/acoustics.yaml:56) and
    renormalises to unit trace.
    """
    rng = np.random.default_rng(seed)
    Ks = []
    lever = modes.z_pivot - modes.rec_z
    dr = -lever * np.sin(np.deg2rad(tilt_deg))
    for f in range(modes.n_freq):
        phi_s = modes.phi_src(f, src_z)[0]
        rn = 1000.0 * float(rec_r_km) + dr                       # [N]
        kr = modes.k[f][None, :] * rn[:, None]                   # [N, M]
        p = np.sum(modes.phi_rec[f] * phi_s[None, :] * np.exp(-1j * kr) / np.sqrt(kr), axis=1)
        d = p / np.linalg.norm(p)
        if snapshots <= 0:
            Ks.append(_outer(d))
            continue
        sigma = 10.0 ** (-snr_db / 20.0) / np.sqrt(modes.n_rec)
        K = np.zeros((modes.n_rec, modes.n_rec), dtype=np.complex128)
        for _ in range(snapshots):
            noise = sigma * (rng.standard_normal(modes.n_rec) + 1j * rng.standard_normal(modes.n_rec)) / np.sqrt(2.0)
            phase = np.exp(2j * np.pi * rng.random())
            K += _outer(phase * d + noise)
        K /= np.real(np.trace(K))
        Ks.append(K)
    return np.array(Ks)
