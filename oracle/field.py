"""Oracle: normal-mode replica field.  Test infrastructure only.

Restates the GPU-resident half of ``tritonoa.at.models.kraken.runner.run_kraken``
[ext]: given modes, the pressure at receiver ``n`` from a source at range ``r``
and depth ``z_s`` (BASELINE.json north_star):

    p_n(r) = sum_m phi_m(z_s) phi_m(z_n) exp(-i k_m r_n') / sqrt(k_m r_n')

with ``r_n' = r`` for a vertical array and ``r + dr_n(tilt)`` for a tilted one
(keys ``tilt`` [deg], ``z_pivot``: ref/projects/swellex96_loc_tilt/data/env.py:73-74;
geometry helper ``compute_range_offsets`` used at
ref/projects/swellex96_loc_tilt/data/source_tow.py:105).  Output shape
``p[N, Nr]`` is evidenced by ref/optimization/utils.py:29-31 and
ref/projects/swellex96_localization/data/mfp.py:214.

Conventions chosen where the external package is unverifiable:
the rigid array pivots at ``z_pivot``; element ``n`` moves horizontally by
``dr_n = -(z_pivot - z_n) sin(tilt)``; element depths (hence ``phi_rec``) are
kept at their nominal values.  Global constants of the AT field
(``i e^{-i pi/4}/sqrt(8 pi)``) cancel in the normalised processors and are
omitted, as in the north_star formula.
"""
from __future__ import annotations

import numpy as np


def interp_phi_src(tab: np.ndarray, z0: float, dz: float, z) -> np.ndarray:
    """Linear interpolation of the source-depth mode table (data contract).

    ``t = clip((z - z0)/dz, 0, nz-1)``; ``i0 = min(floor(t), nz-2)``;
    ``phi = (1-w) tab[i0] + w tab[i0+1]``, ``w = t - i0``.  Returns ``[len(z), M]``.
    """
    z = np.atleast_1d(np.asarray(z, dtype=np.float64))
    nz = tab.shape[0]
    t = np.clip((z - z0) / dz, 0.0, float(nz - 1))
    i0 = np.minimum(np.floor(t).astype(np.int64), nz - 2)
    w = (t - i0)[:, None]
    return (1.0 - w) * tab[i0] + w * tab[i0 + 1]


def range_offsets(rec_z: np.ndarray, tilt_deg: float, z_pivot: float) -> np.ndarray:
    """Horizontal offsets ``dr_n`` [m] of a rigid array tilted by ``tilt_deg``."""
    rec_z = np.asarray(rec_z, dtype=np.float64)
    return -(z_pivot - rec_z) * np.sin(np.deg2rad(float(tilt_deg)))


def pressure_field(k: np.ndarray, phi_rec: np.ndarray, phi_src: np.ndarray,
                   r_m, dr: np.ndarray | None = None) -> np.ndarray:
    """Replica field ``p[N, Nr]`` (complex128) for one tonal and one source depth.

    k [M] complex, phi_rec [N, M], phi_src [M] (already evaluated at the source
    depth), r_m [Nr] ranges in metres, dr [N] per-receiver range offsets or None.
    """
    k = np.asarray(k, dtype=np.complex128)
    r = np.atleast_1d(np.asarray(r_m, dtype=np.float64))
    phi_src = np.asarray(phi_src)
    if dr is None or not np.any(dr):
        kr = k[:, None] * r[None, :]                         # [M, Nr]
        a = phi_src[:, None] * np.exp(-1j * kr) / np.sqrt(kr)
        return phi_rec @ a
    rn = r[None, :] + np.asarray(dr, dtype=np.float64)[:, None]   # [N, Nr]
    kr = k[None, :, None] * rn[:, None, :]                  # [N, M, Nr]
    h = np.exp(-1j * kr) / np.sqrt(kr)
    return np.einsum("nm,m,nmr->nr", phi_rec, phi_src, h)


def replica(modes, f: int, src_z: float, r_km, tilt_deg: float = 0.0) -> np.ndarray:
    """``run_kraken``-shaped call on a ModeSet-like object: ``p[N, Nr]`` for tonal ``f``.

    ``r_km`` in km as at the reference boundary (mfp.py:214).
    """
    phi_s = interp_phi_src(modes.phi_src_tab[f], modes.z0, modes.dz, src_z)[0]
    dr = range_offsets(modes.rec_z, tilt_deg, modes.z_pivot) if tilt_deg else None
    r_m = 1000.0 * np.atleast_1d(np.asarray(r_km, dtype=np.float64))
    return pressure_field(modes.k[f], modes.phi_rec[f], phi_s, r_m, dr)
