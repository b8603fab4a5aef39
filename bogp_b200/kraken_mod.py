"""KRAKEN ``.mod`` files -> :class:`bogp_b200.modes.ModeSet` (host side of the mode provider).

KRAKEN itself stays external (north_star): the reference runs it through TritonOA's ``run_kraken``
(ref/projects/swellex96_localization/conf/optimization/run.py:99), which writes an ``.env`` file, starts
``kraken.exe`` / ``krakenc.exe`` and reads the binary mode file back.  This module is that last step for callers who
run KRAKEN themselves (or keep mode files of a fixed environment on disk): ``read_mod`` returns the eigenvalues, the
sample depths and the mode shapes of one frequency, ``modeset_from_mod`` assembles the per-tonal files (TritonOA names
a run after its tonal, ``f"{freq:.1f}Hz"``, ref SURVEY App. A) or one broadband file into the ``ModeSet`` the kernels
take (receiver rows and the uniform source-depth table by linear interpolation, ``modes.modeset_from_tables``).

PARITY UNPINNED.  The layout below is the direct-access format of the Acoustics Toolbox's mode files as its MATLAB
reader ``read_modes_bin.m`` (2014-2020 releases, broadband-capable) walks it, restated from memory -- neither the
toolbox nor a file written by KRAKEN is available in this environment, so the reader is checked only against
``write_mod`` of this module (round trip) and against the data contract of ``ModeSet``.  Records of ``lrecl`` bytes,
little-endian; the first word of the file is ``lrecl / 4``:

====================  =====================================================================================
record                content
====================  =====================================================================================
0                     int32 lrecl/4 | char[80] title | int32 Nfreq | int32 Nmedia | int32 Ntot | int32 NMat
1                     per medium: int32 N | char[8] material
2                     per medium: float32 depth of its top interface | float32 density
3                     float64 freqVec[Nfreq]
4                     float32 z[Ntot]              (depths at which the mode shapes are tabulated)
p (first: p = 5)      int32 M                       (number of modes of this frequency)
p + 1                 top half-space: char BC | float32 cp(re, im) | float32 cs(re, im) | float32 rho | float32 depth;
                      bottom half-space: the same seven fields
p + 1 + m, m = 1..M   float32 (re, im) x NMat      (mode shape m)
p + 2 + M ...         float32 (re, im) x M          (eigenvalues k_m; continues over as many records as it needs)
next frequency        p += 3 + M + floor(4 (2 M - 1) / lrecl)
====================  =====================================================================================
"""
from __future__ import annotations

import struct
from pathlib import Path
from typing import Dict, Optional, Sequence, Union

import numpy as np

from .modes import ModeSet, modeset_from_tables

_HEADER_RECORDS = 5


def _halfspace(buf: bytes, off: int):
    bc = buf[off:off + 1].decode("latin-1")
    cp_re, cp_im, cs_re, cs_im, rho, depth = struct.unpack_from("<6f", buf, off + 1)
    return {"bc": bc, "cp": complex(cp_re, cp_im), "cs": complex(cs_re, cs_im), "rho": rho, "depth": depth}, off + 25


def read_mod(path: Union[str, Path], freq: float = 0.0, modes: Optional[Sequence[int]] = None) -> Dict:
    """Modes of the frequency of the file nearest to ``freq`` (``freq=0``: the first).  Returns a dict with ``k [M]``
    complex128, ``z [Ntot]`` float64, ``phi [Ntot, M]`` (float64, or complex128 if any imaginary part is non-zero:
    KRAKENC), ``freq``, ``freqs``, ``M``, ``title``, ``n_media``, ``depth``, ``rho``, ``top``, ``bottom``.
    ``modes``: optional 1-based mode indices to keep (as the toolbox reader's third argument)."""
    raw = Path(path).read_bytes()
    if len(raw) < 4:
        raise ValueError(f"{path}: not a mode file (too short)")
    lrecl = 4 * struct.unpack_from("<i", raw, 0)[0]
    if lrecl < 100 or len(raw) < _HEADER_RECORDS * lrecl:
        raise ValueError(f"{path}: not a mode file (record length {lrecl} bytes, file {len(raw)} bytes)")
    title = raw[4:84].decode("latin-1").rstrip("\x00 ").strip()
    n_freq, n_media, n_tot, n_mat = struct.unpack_from("<4i", raw, 84)
    if n_tot <= 0 or n_freq <= 0 or n_media <= 0:
        raise ValueError(f"{path}: empty mode file (Nfreq={n_freq}, Nmedia={n_media}, Ntot={n_tot})")
    if n_mat != n_tot:
        raise NotImplementedError(f"{path}: elastic media (NMat={n_mat} != Ntot={n_tot}) are not supported")
    med = np.frombuffer(raw, dtype="<f4", count=2 * n_media, offset=2 * lrecl).reshape(n_media, 2)
    freqs = np.frombuffer(raw, dtype="<f8", count=n_freq, offset=3 * lrecl).astype(np.float64)
    z = np.frombuffer(raw, dtype="<f4", count=n_tot, offset=4 * lrecl).astype(np.float64)
    i_freq = int(np.argmin(np.abs(freqs - freq))) if freq else 0
    p = _HEADER_RECORDS
    M = 0
    for i in range(i_freq + 1):
        if (p + 1) * lrecl > len(raw):
            raise ValueError(f"{path}: truncated before frequency {i}")
        M = struct.unpack_from("<i", raw, p * lrecl)[0]
        if M < 0:
            raise ValueError(f"{path}: negative mode count at record {p}")
        if i < i_freq:
            p += 3 + M + (4 * (2 * M - 1)) // lrecl
    top, off = _halfspace(raw, (p + 1) * lrecl)
    bottom, _ = _halfspace(raw, off)
    keep = np.arange(1, M + 1) if modes is None else np.asarray([m for m in modes if 1 <= m <= M], dtype=np.int64)
    end = (p + 2 + M) * lrecl + 8 * M
    if end > len(raw):
        raise ValueError(f"{path}: truncated ({len(raw)} bytes, frequency {i_freq} needs {end})")
    phi = np.empty((n_tot, len(keep)), dtype=np.complex128)
    for j, m in enumerate(keep):
        v = np.frombuffer(raw, dtype="<f4", count=2 * n_mat, offset=(p + 1 + int(m)) * lrecl).astype(np.float64)
        phi[:, j] = v[0::2] + 1j * v[1::2]
    kv = np.frombuffer(raw, dtype="<f4", count=2 * M, offset=(p + 2 + M) * lrecl).astype(np.float64)
    k = (kv[0::2] + 1j * kv[1::2])[keep - 1] if M else np.zeros(0, dtype=np.complex128)
    return {"k": k.astype(np.complex128), "z": z, "phi": phi if np.any(phi.imag) else np.ascontiguousarray(phi.real),
            "freq": float(freqs[i_freq]), "freqs": freqs, "M": int(len(keep)), "title": title, "n_media": int(n_media),
            "depth": med[:, 0].astype(np.float64), "rho": med[:, 1].astype(np.float64), "top": top, "bottom": bottom}


def write_mod(path: Union[str, Path], freqs: Sequence[float], z: Sequence[float], k: Sequence[np.ndarray],
              phi: Sequence[np.ndarray], *, title: str = "bogp_b200", depth: Sequence[float] = (0.0,),
              rho: Sequence[float] = (1.0,), bottom_cp: complex = 1650.0, bottom_rho: float = 1.8) -> None:
    """Writes the layout of the module docstring: one profile per frequency, ``k[f] [M_f]`` complex, ``phi[f] [Ntot, M_f]``
    tabulated at the common depths ``z``.  Single precision on disk, as KRAKEN writes it.  For tests and for keeping the
    modes of a fixed environment next to the data."""
    z = np.asarray(z, dtype=np.float64)
    n_tot, n_media, n_freq = len(z), len(depth), len(freqs)
    words = max(2 * n_tot, 32, 3 * n_media, 2 * n_freq, n_tot, 14)
    lrecl = 4 * words
    recs = []

    def rec(payload: bytes) -> None:
        if len(payload) > lrecl:
            raise ValueError("record overflow")
        recs.append(payload + b"\x00" * (lrecl - len(payload)))

    rec(struct.pack("<i", words) + title.encode("latin-1")[:80].ljust(80) + struct.pack("<4i", n_freq, n_media, n_tot, n_tot))
    rec(b"".join(struct.pack("<i", n_tot if i == 0 else 0) + b"ACOUSTIC" for i in range(n_media)))
    rec(np.stack([np.asarray(depth, dtype="<f4"), np.asarray(rho, dtype="<f4")], axis=1).tobytes())
    rec(np.asarray(freqs, dtype="<f8").tobytes())
    rec(z.astype("<f4").tobytes())
    for kf, pf in zip(k, phi):
        kf = np.asarray(kf, dtype=np.complex128)
        pf = np.asarray(pf, dtype=np.complex128)
        M = len(kf)
        if pf.shape != (n_tot, M):
            raise ValueError("write_mod: phi[f] must be [len(z), len(k[f])]")
        rec(struct.pack("<i", M))
        half = lambda bc, cp, rho_, d: bc.encode("latin-1") + struct.pack("<6f", cp.real, cp.imag, 0.0, 0.0, rho_, d)   # noqa: E731
        rec(half("V", 0j, 0.0, float(depth[0])) + half("A", complex(bottom_cp), float(bottom_rho), float(z[-1])))
        for m in range(M):
            rec(np.stack([pf[:, m].real, pf[:, m].imag], axis=1).astype("<f4").tobytes())
        kb = np.stack([kf.real, kf.imag], axis=1).astype("<f4").tobytes()
        n_k = 1 + (4 * (2 * M - 1)) // lrecl if M else 1
        kb = kb.ljust(n_k * lrecl, b"\x00")
        for i in range(n_k):
            recs.append(kb[i * lrecl:(i + 1) * lrecl])
    Path(path).write_bytes(b"".join(recs))


def modeset_from_mod(paths: Union[str, Path, Sequence[Union[str, Path]]], freq: Sequence[float], rec_z: Sequence[float], *,
                     z_pivot: Optional[float] = None, dz_tab: float = 0.25, meta: Optional[dict] = None) -> ModeSet:
    """ModeSet of the tonals ``freq`` from KRAKEN mode files: ``paths`` is one broadband file (every tonal is looked up in
    it by frequency) or one file per tonal in the order of ``freq`` (what a per-tonal ``run_kraken`` leaves behind).
    A tonal whose nearest tabulated frequency is more than 0.5 Hz away is an error."""
    files = [paths] * len(freq) if isinstance(paths, (str, Path)) else list(paths)
    if len(files) != len(freq):
        raise ValueError("modeset_from_mod: one file for all tonals, or one file per tonal")
    ks, zs, phis = [], [], []
    for f, path in zip(freq, files):
        m = read_mod(path, float(f))
        if abs(m["freq"] - float(f)) > 0.5:
            raise ValueError(f"{path}: nearest frequency {m['freq']} Hz is not the tonal {f} Hz")
        if m["M"] == 0:
            raise ValueError(f"{path}: no modes at {f} Hz")
        ks.append(m["k"]); zs.append(m["z"]); phis.append(m["phi"])
    info = {"kind": "kraken .mod", "files": [str(p) for p in files]}
    info.update(meta or {})
    return modeset_from_tables(freq, ks, zs, phis, rec_z, z_pivot=z_pivot, dz_tab=dz_tab, meta=info)
