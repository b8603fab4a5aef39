"""Boundary A: the matched-field objective, host side (mirrors the TritonOA interface the reference calls).

Reference interfaces mirrored here (all external to ref/, pinned by their call sites -- SURVEY 8b):

* ``tritonoa.sp.mfp.MatchedFieldProcessor(runner, covariance_matrix, freq, parameters,
  parameter_formatter=None, beamformer=beamformer, multifreq_method="mean")`` and its ``__call__(dict)``
  (ref/projects/swellex96_localization/data/mfp.py:198-214, ref/projects/swellex96_inv/data/bo/obj.py:62-70,
  ref/projects/swellex96_localization/conf/optimization/run.py:89-96);
* ``tritonoa.at.models.kraken.runner.run_kraken(parameters) -> p[N, Nr]`` -- here :class:`ModeRunner`, a *mode
  provider*: KRAKEN stays external on the host (north_star) and hands over a :class:`~bogp_b200.modes.ModeSet`;
* ``tritonoa.sp.beamforming.beamformer(K, replica, atype)`` / ``covariance(d)``
  (obj.py:68, ref/optimization/utils.py:23-32).

The arithmetic runs in libbogp_sm100.so (fused modal sum + Bartlett + multi-frequency combine, no replica field in
HBM).  There is no CPU fallback: an objective built from callables the library cannot fuse raises.
"""
from __future__ import annotations

import ctypes as C
import copy
from functools import partial
from typing import Callable, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .modes import ModeSet

ArrayLike = Union[np.ndarray, Sequence[float]]


def _split_planes(arrs: List[np.ndarray]):
    """Concatenate per-tonal arrays into contiguous float64 re (and im, if any is complex) planes."""
    cplx = any(np.iscomplexobj(a) for a in arrs)
    re = np.ascontiguousarray(np.concatenate([np.real(a).ravel() for a in arrs]), dtype=np.float64)
    im = (np.ascontiguousarray(np.concatenate([np.imag(a).ravel() for a in arrs]), dtype=np.float64)
          if cplx else None)
    return re, im


class DeviceModeSet:
    """A :class:`ModeSet` uploaded to the current CUDA device (``bogp_modeset_t`` handle)."""

    def __init__(self, modes: ModeSet, covariance_matrix: Optional[np.ndarray] = None):
        self.modes = modes
        self._h = C.c_void_p()
        self._vec_cache: dict = {}
        self._out_cache = None
        M = np.asarray(modes.n_modes, dtype=np.int32)
        k_re, k_im = _split_planes([np.asarray(k, dtype=np.complex128) for k in modes.k])
        pr_re, pr_im = _split_planes(modes.phi_rec)
        ps_re, ps_im = _split_planes(modes.phi_src_tab)
        rec_z = np.ascontiguousarray(modes.rec_z, dtype=np.float64)
        _lib.call("bogp_modeset_create", C.byref(self._h), modes.n_freq, modes.n_rec,
                  M.ctypes.data_as(C.POINTER(C.c_int)), _lib.dptr(k_re), _lib.dptr(k_im), _lib.dptr(pr_re),
                  _lib.dptr(pr_im), modes.nz_tab, float(modes.z0), float(modes.dz), _lib.dptr(ps_re),
                  _lib.dptr(ps_im), _lib.dptr(rec_z), float(modes.z_pivot))
        self.device = torch.device("cuda", torch.cuda.current_device())
        if covariance_matrix is not None:
            self.set_covariance(covariance_matrix)

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().bogp_modeset_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ per-call host work
    def _vec(self, key: str, x, scale: float = 1.0):
        """float64 C-contiguous copy of a grid vector (scaled: km -> m) and its ``double*``, kept while the caller passes
        the same values -- a sweep calls with ONE grid for all its time steps (mfp.py:166-216), and the conversion plus the
        ctypes pointer cost more than the comparison of 8 kB.  Returns ``(array, pointer)``."""
        a = np.atleast_1d(np.asarray(x, dtype=np.float64))
        raw = a.tobytes()
        hit = self._vec_cache.get(key)
        if hit is not None and hit[0] == raw:
            return hit[1], hit[2]
        v = np.ascontiguousarray(scale * a if scale != 1.0 else a)
        ptr = _lib.dptr(v)
        self._vec_cache[key] = (raw, v, ptr)
        return v, ptr

    def _out_ptr(self, out: np.ndarray):
        """``float*`` of a host output array, kept per array object (the reference is held, so the id cannot be reused)."""
        hit = self._out_cache
        if hit is not None and hit[0] is out:
            return hit[1]
        ptr = _lib.fptr(out)
        self._out_cache = (out, ptr)
        return ptr

    # ------------------------------------------------------------------ inputs
    def set_covariance(self, K: np.ndarray) -> None:
        """``covariance_matrix`` of MatchedFieldProcessor: ``K[F, N, N]`` complex (utils.py:23-32, 42-56)."""
        K = np.asarray(K)
        F, N = self.modes.n_freq, self.modes.n_rec
        if K.shape != (F, N, N):
            raise ValueError(f"covariance_matrix must be [F={F}, N={N}, N={N}], got {K.shape}")
        K_re = np.ascontiguousarray(K.real, dtype=np.float64)
        K_im = np.ascontiguousarray(K.imag if np.iscomplexobj(K) else np.zeros_like(K_re), dtype=np.float64)
        _lib.call("bogp_covariance_set", self._h, _lib.dptr(K_re), _lib.dptr(K_im))

    def tonal_info(self):
        """Per tonal ``(M_f, rows_f, J_f)``: modes, rows of the mode-space operator, covariance rank."""
        out = []
        for f in range(self.modes.n_freq):
            M, rows, J = C.c_int(), C.c_int(), C.c_int()
            _lib.call("bogp_modeset_info", self._h, f, C.byref(M), C.byref(rows), C.byref(J))
            out.append((M.value, rows.value, J.value))
        return out

    def tonal_info_umma(self):
        """Per tonal ``(M_f, rows_f, rows_padded_f, folded_f)`` of the operator the tcgen05 engine multiplies by."""
        out = []
        for f, (M, _, _) in enumerate(self.tonal_info()):
            rows, rp, fold = C.c_int(), C.c_int(), C.c_int()
            _lib.call("bogp_modeset_info_umma", self._h, f, C.byref(rows), C.byref(rp), C.byref(fold))
            out.append((M, rows.value, rp.value, bool(fold.value)))
        return out

    def umma_kind(self) -> int:
        """Arithmetic of the tcgen05 engine a vertical-array grid call runs: 16 (three FP16-split products), 32 (three
        TF32-split products) or 0 (mode set / device not covered by the tensor-core engines)."""
        return int(_lib.lib().bogp_umma_kind(self._h))

    def flops_per_candidate(self, engine: str = "simt") -> float:
        """Algorithmic flop of one candidate over all tonals on the vertical-array path: sum_f 4 rows_f M_f, with the
        row count of the operator the engine actually needs (the tcgen05 engine folds a rank-1 numerator into the norm
        rows: rank(G) rows instead of rank(G) + 2)."""
        if engine == "umma":
            return float(sum(4 * rows * M for M, rows, _, _ in self.tonal_info_umma()))
        return float(sum(4 * rows * M for M, rows, _ in self.tonal_info()))

    # ------------------------------------------------------------------ evaluation
    def bartlett_grid(self, r_km: ArrayLike, z: ArrayLike, tilt_deg: Optional[ArrayLike] = None, *,
                      atype: str = "cbf", multifreq_method: str = "mean", engine: str = "auto",
                      out: Optional[torch.Tensor] = None, argext: Optional[tuple] = None,
                      peer: Optional[tuple] = None) -> torch.Tensor:
        """Ambiguity surface ``[Nz, Nr]`` (or ``[Nt, Nz, Nr]`` with a tilt axis), float32 on the device.

        Replaces the row loop of ref/projects/swellex96_localization/data/mfp.py:213-214.  ``argext=(pair, scratch,
        maximize)`` also writes the extremum record of :func:`argext_device` (``pair`` int64[2], ``scratch`` >= 8192
        bytes) -- inside the surface kernel when the tcgen05 engine runs, so the final reduction of collate.py:50 costs
        no extra launch.  With ``peer=(PeerExchange, index_offset, pairs)`` the record ``(index_offset + index, value)``
        is also exchanged with the other GPUs of the box (``pairs`` int64[world, 2] gets every rank's record) -- from
        inside the same kernel on the FP16-split engine.
        """
        r_m, p_r = self._vec("r", r_km, 1000.0)
        zz, p_z = self._vec("z", z)
        tt, p_t = (None, None) if tilt_deg is None else self._vec("t", tilt_deg)
        Nt = 1 if tt is None else len(tt)
        shape = (len(zz), len(r_m)) if tt is None else (Nt, len(zz), len(r_m))
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 CUDA tensor of shape {shape}")
        if argext is not None:
            pair, scratch, maximize = argext
            if pair.dtype != torch.int64 or pair.numel() < 2 or not pair.is_cuda:
                raise ValueError("argext pair must be an int64[2] CUDA tensor")
            if peer is not None:
                px, offset, pairs = peer
                if not (pairs.is_cuda and pairs.dtype == torch.int64 and pairs.numel() == 2 * px.world and pairs.is_contiguous()):
                    raise ValueError("peer pairs must be a contiguous CUDA int64[world, 2] tensor")
                _lib.call("bogp_bartlett_grid_argext_peer", self._h, p_r, len(r_m), p_z, len(zz),
                          p_t, Nt, _lib.ATYPE[atype], _lib.MF_METHOD[multifreq_method], _lib.ENGINE[engine],
                          C.c_void_p(out.data_ptr()), 1 if maximize else 0, px._h, int(offset), C.c_void_p(pairs.data_ptr()),
                          C.c_void_p(pair.data_ptr() + 8), C.c_void_p(pair.data_ptr()), C.c_void_p(scratch.data_ptr()),
                          scratch.numel() * scratch.element_size(), C.c_void_p(_lib.current_stream_ptr()))
                return out
            _lib.call("bogp_bartlett_grid_argext", self._h, p_r, len(r_m), p_z, len(zz),
                      p_t, Nt, _lib.ATYPE[atype], _lib.MF_METHOD[multifreq_method], _lib.ENGINE[engine],
                      C.c_void_p(out.data_ptr()), 1 if maximize else 0, C.c_void_p(pair.data_ptr() + 8),
                      C.c_void_p(pair.data_ptr()), C.c_void_p(scratch.data_ptr()),
                      scratch.numel() * scratch.element_size(), C.c_void_p(_lib.current_stream_ptr()))
            return out
        _lib.call("bogp_bartlett_grid", self._h, p_r, len(r_m), p_z, len(zz), p_t, Nt,
                  _lib.ATYPE[atype], _lib.MF_METHOD[multifreq_method], _lib.ENGINE[engine],
                  C.c_void_p(out.data_ptr()), C.c_void_p(_lib.current_stream_ptr()))
        return out

    def bartlett_grid_host(self, r_km: ArrayLike, z: ArrayLike, tilt_deg: Optional[ArrayLike] = None, *,
                           atype: str = "cbf", multifreq_method: str = "mean", engine: str = "auto",
                           out: Optional[np.ndarray] = None, argext: Optional[str] = None, peer: Optional[tuple] = None):
        """Same through the host-buffer entry point (device->host copy inside the call).  ``out``: optional
        C-contiguous float32 host array of the result shape (pass page-locked memory for a DMA copy).
        ``argext="max"|"min"`` also returns ``(value, unravelled index)`` of the extremum (collate.py:50), computed
        on the device while the surface is still there: ``(surface, (value, index))``.  With ``peer=(PeerExchange,
        index_offset)`` the call returns ``(surface, records)``: the int64[world, 2] ``(index_offset + index, value
        bits)`` records of every rank of the box, exchanged over peer memory inside the same call (the records of the
        PREVIOUS call when the exchange is pipelined, ``PeerExchange.set_pipelined``).

        The per-call host work is kept to a few microseconds (a step is ~0.19 ms end to end): converted grid vectors and
        their pointers are reused while the caller passes the same values, see :meth:`_vec`."""
        r_m, p_r = self._vec("r", r_km, 1000.0)
        zz, p_z = self._vec("z", z)
        if tilt_deg is None:
            Nt, p_t, shape = 1, None, (len(zz), len(r_m))
        else:
            tt, p_t = self._vec("t", tilt_deg)
            Nt = len(tt)
            shape = (Nt, len(zz), len(r_m))
        if out is None:
            out = np.empty(shape, dtype=np.float32)
        elif out.shape != shape or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous float32 array of shape {shape}")
        p_out = self._out_ptr(out)
        stream = _lib.current_stream_ptr()
        if argext is not None:
            if argext not in ("max", "min"):
                raise ValueError("argext must be 'max', 'min' or None")
            if peer is not None:
                px, offset = peer
                recs = np.zeros((px.world, 2), dtype=np.int64)
                _lib.call("bogp_bartlett_grid_host_argmax_peer", self._h, p_r, len(r_m), p_z, len(zz),
                          p_t, Nt, _lib.ATYPE[atype], _lib.MF_METHOD[multifreq_method], _lib.ENGINE[engine],
                          p_out, 1 if argext == "max" else 0, px._h, int(offset),
                          recs.ctypes.data_as(C.POINTER(C.c_longlong)), stream)
                return out, recs
            val, idx = C.c_float(), C.c_longlong()
            _lib.call("bogp_bartlett_grid_host_argmax", self._h, p_r, len(r_m), p_z, len(zz),
                      p_t, Nt, _lib.ATYPE[atype], _lib.MF_METHOD[multifreq_method], _lib.ENGINE[engine],
                      p_out, 1 if argext == "max" else 0, C.byref(val), C.byref(idx), stream)
            i = idx.value
            if i < 0:
                pos = tuple(-1 for _ in shape)
            elif len(shape) == 2:
                pos = divmod(i, shape[1])
            else:
                pos = tuple(int(k) for k in np.unravel_index(i, shape))
            return out, (val.value, pos)
        _lib.call("bogp_bartlett_grid_host", self._h, p_r, len(r_m), p_z, len(zz), p_t,
                  Nt, _lib.ATYPE[atype], _lib.MF_METHOD[multifreq_method], _lib.ENGINE[engine], p_out, stream)
        return out

    def bartlett_points(self, cand: Union[np.ndarray, torch.Tensor], *, atype: str = "cbf",
                        multifreq_method: str = "mean") -> torch.Tensor:
        """Scattered candidates ``cand[C, 3] = (range km, depth m, tilt deg)`` -> ``[C]`` float32 on the device."""
        if not torch.is_tensor(cand):
            cand = torch.as_tensor(np.asarray(cand, dtype=np.float64))
        cand = cand.to(device=self.device, dtype=torch.float64).reshape(-1, 3).clone()
        cand[:, 0] *= 1000.0
        cand = cand.contiguous()
        out = torch.empty(cand.shape[0], dtype=torch.float32, device=self.device)
        _lib.call("bogp_bartlett_points", self._h, C.c_void_p(cand.data_ptr()), cand.shape[0], _lib.ATYPE[atype],
                  _lib.MF_METHOD[multifreq_method], C.c_void_p(out.data_ptr()), C.c_void_p(_lib.current_stream_ptr()))
        return out


def argmax(x: torch.Tensor):
    """``np.unravel_index(np.argmax(surface))`` (collate.py:50) on the device: ``(value, unravelled index)``."""
    return _argext("bogp_argmax_f32", x)


def argmin(x: torch.Tensor):
    return _argext("bogp_argmin_f32", x)


def _argext(fn: str, x: torch.Tensor):
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        raise ValueError("expected a contiguous float32 CUDA tensor")
    val, idx = C.c_float(), C.c_longlong()
    _lib.call(fn, C.c_void_p(x.data_ptr()), x.numel(), C.byref(val), C.byref(idx),
              C.c_void_p(_lib.current_stream_ptr()))
    return val.value, tuple(int(i) for i in np.unravel_index(idx.value, tuple(x.shape)))


def argext_device(x: torch.Tensor, pair: torch.Tensor, scratch: torch.Tensor, maximize: bool = True) -> None:
    """Arg-max/min without a host round trip: writes ``pair[0] = flat index`` (int64, -1 if all NaN) and the float32
    extremum into the low 4 bytes of ``pair[1]``.  ``pair``: int64[2] on the device (the 16-byte record the multi-GPU
    final reduction all-gathers); ``scratch``: >= 8192 bytes of device memory."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        raise ValueError("expected a contiguous float32 CUDA tensor")
    if pair.dtype != torch.int64 or pair.numel() < 2 or not pair.is_cuda:
        raise ValueError("pair must be an int64[2] CUDA tensor")
    _lib.call("bogp_argext_f32_dev", C.c_void_p(x.data_ptr()), x.numel(), 1 if maximize else 0,
              C.c_void_p(pair.data_ptr() + 8), C.c_void_p(pair.data_ptr()), C.c_void_p(scratch.data_ptr()),
              scratch.numel() * scratch.element_size(), C.c_void_p(_lib.current_stream_ptr()))


def unpack_pairs(pairs: torch.Tensor):
    """Host view of all-gathered ``[world, 2]`` int64 records -> (values float32[world], indices int64[world])."""
    p = pairs.reshape(-1, 2).cpu().numpy()
    return p[:, 1].astype(np.int64).view(np.float32)[::2].copy(), p[:, 0].copy()


# ---------------------------------------------------------------------------------------------- tritonoa mirror
def covariance(d: np.ndarray) -> np.ndarray:
    """``tritonoa.sp.beamforming.covariance``: ``K = d d^H`` (utils.py:31; acoustics.py:116-124).  Host, set-up only."""
    d = np.asarray(d, dtype=np.complex128).reshape(-1, 1)
    return d @ d.conj().T


def simulate_covariance(runner, parameters: dict, freq: Sequence[float]) -> np.ndarray:
    """``optimization.utils.simulate_covariance(runner, parameters, freq)`` (ref/optimization/utils.py:23-32):
    ``K[F, N, N]`` with ``K_f = d d^H``, ``d = p/||p||`` the replica at the truth in ``parameters`` (``src_z``,
    ``rec_r`` [km], optional ``tilt``).  ``runner`` is :func:`run_kraken` (or a partial of it, as the configs build it)
    or a :class:`ModeRunner`.  Host float64, set-up only (one modal sum per tonal)."""
    from .modes import synthetic_covariance
    modes = _as_mode_runner(runner).modes_for(parameters)
    idx = []
    for f in freq:
        hits = [i for i, g in enumerate(modes.freq) if abs(g - float(f)) < 1e-9]
        if not hits:
            raise ValueError(f"mode set has no tonal at {f} Hz")
        idx.append(hits[0])
    sel = modes if idx == list(range(modes.n_freq)) else modes.select(idx)
    return synthetic_covariance(sel, float(parameters["src_z"]), float(np.asarray(parameters["rec_r"])),
                                float(parameters.get("tilt", 0.0) or 0.0))


def beamformer(K: np.ndarray, replica: np.ndarray, atype: str = "cbf") -> np.ndarray:
    """Marker for ``tritonoa.sp.beamforming.beamformer``.

    Inside :class:`MatchedFieldProcessor` the processor named by ``atype`` is fused into the CUDA kernel and this
    function is never called with data; it exists so that ``partial(beamformer, atype="cbf_ml")`` (obj.py:68,
    loc_tilt/conf/run.py:146) keeps working as the selector.
    """
    raise NotImplementedError(
        "bogp_b200.mfp.beamformer is fused into the Bartlett kernels; pass it (or functools.partial of it) to "
        "MatchedFieldProcessor instead of calling it on a replica field")


class ModeRunner:
    """``run_kraken`` stand-in: the mode provider for the fused kernels.

    ``modes`` is either a :class:`ModeSet` (one fixed environment: localisation / tilt) or a callable
    ``parameters -> ModeSet`` that runs the host normal-mode program for the environment in ``parameters``
    (geoacoustic inversion, obj.py:73-83).  The modal sum itself happens on the GPU.
    """

    def __init__(self, modes: Union[ModeSet, Callable[[dict], ModeSet]]):
        self.modes = modes

    @property
    def is_static(self) -> bool:
        return isinstance(self.modes, ModeSet)

    def modes_for(self, parameters: dict) -> ModeSet:
        return self.modes if self.is_static else self.modes(parameters)

    def __call__(self, parameters: dict):
        raise NotImplementedError(
            "ModeRunner does not materialise replica fields; use it as MatchedFieldProcessor(runner=...)")


# ---- run_kraken: the function the unchanged configs name (conf/optimization/run.py:99 `pbuilds(run_kraken)`)
_MODE_PROVIDER: Union[None, ModeSet, Callable[[dict], ModeSet]] = None


def set_mode_provider(provider: Union[None, ModeSet, Callable[[dict], ModeSet]]) -> None:
    """Registers what :func:`run_kraken` gets its modes from: a :class:`ModeSet` (one fixed environment) or a callable
    ``parameters -> ModeSet`` that runs the host normal-mode program (KRAKEN stays external, north_star) for the
    environment in ``parameters`` -- e.g. ``lambda env: modeset_from_tables(...)`` around TritonOA's KRAKEN wrapper."""
    global _MODE_PROVIDER
    _MODE_PROVIDER = provider


def get_mode_provider():
    return _MODE_PROVIDER


def run_kraken(parameters: dict) -> np.ndarray:
    """``tritonoa.at.models.kraken.runner.run_kraken`` as the reference calls it.

    Two uses.  (1) As the ``runner`` of :class:`MatchedFieldProcessor` -- directly or as the ``functools.partial`` that
    ``pbuilds(run_kraken)`` / ``instantiate`` produce (ref/projects/swellex96_localization/conf/optimization/run.py:89-99)
    -- it is recognised by identity and never called: the modal sum is fused into the CUDA kernels and the modes come
    from the registered provider (:func:`set_mode_provider`).  (2) Called with an environment dict carrying ``freq``
    (``runner(parameters | {"freq": f})``, ref/optimization/utils.py:29) it returns the replica field ``p[N]``
    (``p[N, Nr]`` for a vector ``rec_r``) of that single source position in float64 on the host: the set-up step that
    builds the simulated covariance, one modal sum per tonal."""
    from .modes import replica_field
    modes = _as_mode_runner(run_kraken).modes_for(parameters)
    if "freq" not in parameters:
        raise KeyError("run_kraken(parameters): parameters['freq'] is required")
    hits = [i for i, g in enumerate(modes.freq) if abs(g - float(parameters["freq"])) < 1e-9]
    if not hits:
        raise ValueError(f"mode set has no tonal at {parameters['freq']} Hz")
    return replica_field(modes, hits[0], float(parameters["src_z"]), parameters["rec_r"],
                         float(parameters.get("tilt", 0.0) or 0.0))


def _as_mode_runner(runner) -> "ModeRunner":
    """ModeRunner behind what the caller passed as ``runner``: a ModeRunner, :func:`run_kraken`, or partials of either
    without bound arguments (hydra-zen ``zen_partial`` builds)."""
    while isinstance(runner, partial) and not runner.args and not runner.keywords:
        runner = runner.func
    if isinstance(runner, ModeRunner):
        return runner
    if runner is run_kraken:
        if _MODE_PROVIDER is None:
            raise RuntimeError("run_kraken needs a mode provider: call bogp_b200.mfp.set_mode_provider(ModeSet | "
                               "callable(parameters) -> ModeSet) first (KRAKEN itself stays on the host, outside this package)")
        return ModeRunner(_MODE_PROVIDER)
    raise TypeError("runner must be bogp_b200.mfp.run_kraken (or functools.partial of it) or a bogp_b200.mfp.ModeRunner; "
                    "the modal sum is fused into the CUDA kernel and there is no CPU fallback for arbitrary runner callables")


def _atype_of(bf: Callable) -> str:
    """Recover ``atype`` from ``beamformer`` or ``partial(beamformer, atype=...)``."""
    if bf is beamformer:
        return "cbf"
    if isinstance(bf, partial) and bf.func is beamformer and not bf.args:
        extra = set(bf.keywords) - {"atype"}
        if extra:
            raise TypeError(f"unsupported beamformer keywords {sorted(extra)}")
        return bf.keywords.get("atype", "cbf")
    raise TypeError("MatchedFieldProcessor needs bogp_b200.mfp.beamformer (or functools.partial of it); "
                    "arbitrary beamformer callables cannot be fused and there is no CPU fallback")


class MatchedFieldProcessor:
    """Drop-in for ``tritonoa.sp.mfp.MatchedFieldProcessor`` backed by the fused CUDA path.

    ``__call__(parameters)`` accepts what the reference passes: a dict with scalar ``src_z`` and scalar or vector
    ``rec_r`` [km] (mfp.py:214 -> ``ndarray[Nr]``; scalars -> ``float``), optionally ``tilt`` [deg].  It also accepts
    a *list* of such dicts (the batch that obj.py:28-30 spreads over a process pool) and returns ``ndarray[C]``.
    """

    def __init__(self, runner, covariance_matrix, freq: Sequence[float], parameters: dict,
                 parameter_formatter: Optional[Callable] = None, beamformer: Callable = beamformer,
                 multifreq_method: str = "mean"):
        runner = _as_mode_runner(runner)
        if multifreq_method not in _lib.MF_METHOD:
            raise ValueError(f"unknown multifreq_method {multifreq_method!r}")
        self.runner = runner
        self.covariance_matrix = np.asarray(covariance_matrix)
        self.freq = [float(f) for f in freq]
        self.parameters = dict(parameters)
        self.parameter_formatter = parameter_formatter
        self.beamformer = beamformer
        self.atype = _atype_of(beamformer)
        if self.atype not in _lib.ATYPE:
            raise ValueError(f"unknown beamformer atype {self.atype!r}")
        self.multifreq_method = multifreq_method
        if self.covariance_matrix.shape[0] != len(self.freq):
            raise ValueError("covariance_matrix needs one [N, N] matrix per frequency")
        self._dms: Optional[DeviceModeSet] = None

    # pickling (hydra-zen partials, process pools): device handles are rebuilt lazily on the other side
    def __getstate__(self):
        st = dict(self.__dict__)
        st["_dms"] = None
        return st

    # ------------------------------------------------------------------ internals
    def _merged(self, parameters: dict) -> dict:
        """Per-tonal merge exactly as the reference does it, then check the geometry is tonal-independent."""
        merged = None
        for f in self.freq:
            title = f"{f:.1f}Hz"
            if self.parameter_formatter is not None:
                # the reference's formatter (swellex96_inv/data/bo/param_map.py) pops keys from the search dict and
                # edits fixed_parameters['layerdata'] in place, and what it returns shares those nested lists: every
                # tonal and every candidate gets its own copies, the caller's dicts are never handed over
                m = self.parameter_formatter(freq=f, title=title, fixed_parameters=copy.deepcopy(self.parameters),
                                             search_parameters=copy.deepcopy(parameters))
            else:
                m = self.parameters | {"freq": f, "title": title} | parameters
            if merged is None:
                merged = m
            else:
                for key in ("src_z", "rec_r", "tilt"):
                    if not np.array_equal(np.asarray(m.get(key, 0.0)), np.asarray(merged.get(key, 0.0))):
                        raise ValueError(f"parameter {key!r} differs between tonals; cannot fuse")
        return merged

    def _select(self, modes: ModeSet) -> ModeSet:
        idx = []
        for f in self.freq:
            hits = [i for i, g in enumerate(modes.freq) if abs(g - f) < 1e-9]
            if not hits:
                raise ValueError(f"mode set has no tonal at {f} Hz")
            idx.append(hits[0])
        return modes if idx == list(range(modes.n_freq)) else modes.select(idx)

    def _device_modes(self) -> DeviceModeSet:
        if self._dms is None:
            self._dms = DeviceModeSet(self._select(self.runner.modes), self.covariance_matrix)
        return self._dms

    # ------------------------------------------------------------------ objective
    def __call__(self, parameters: Union[dict, Sequence[dict]]):
        if isinstance(parameters, dict):
            return self._call_one(parameters)
        return self._call_batch(list(parameters))

    def _call_one(self, parameters: dict):
        merged = self._merged(parameters)
        if not self.runner.is_static:
            return float(self._call_envs([merged])[0])
        dms = self._device_modes()
        rec_r = np.asarray(merged["rec_r"], dtype=np.float64)
        src_z = np.asarray(merged["src_z"], dtype=np.float64)
        tilt = float(merged.get("tilt", 0.0) or 0.0)
        scalar = rec_r.ndim == 0 and src_z.ndim == 0
        if src_z.ndim != 0:
            raise ValueError("src_z must be a scalar (the reference evaluates one source depth per call)")
        if scalar:
            out = dms.bartlett_points(np.array([[float(rec_r), float(src_z), tilt]]), atype=self.atype,
                                      multifreq_method=self.multifreq_method)
            return float(out.cpu()[0])
        out = dms.bartlett_grid(rec_r, [float(src_z)], [tilt] if tilt else None, atype=self.atype,
                                multifreq_method=self.multifreq_method)
        return out.reshape(-1).cpu().numpy().astype(np.float64)

    def surface(self, rec_r: ArrayLike, src_z: ArrayLike, tilt: Optional[ArrayLike] = None, *,
                engine: str = "auto") -> np.ndarray:
        """The whole ambiguity surface ``[Nz, Nr]`` (``[Nt, Nz, Nr]`` with a tilt axis) in ONE launch: what the
        reference assembles with ``for z in src_z: amb[zz, :] = MFP({"src_z": z, "rec_r": rvec})`` (mfp.py:206-214).
        ``rec_r`` in km as at the reference boundary; float64 host array like the reference's ``amb``."""
        if not self.runner.is_static:
            raise ValueError("surface() needs one shared mode set (static ModeRunner)")
        out = self._device_modes().bartlett_grid_host(rec_r, src_z, tilt, atype=self.atype,
                                                      multifreq_method=self.multifreq_method, engine=engine)
        return out.astype(np.float64)

    def _call_batch(self, plist: List[dict]) -> np.ndarray:
        merged = [self._merged(p) for p in plist]
        if not self.runner.is_static:
            return self._call_envs(merged)
        cand = np.array([[float(np.asarray(m["rec_r"])), float(np.asarray(m["src_z"])),
                          float(m.get("tilt", 0.0) or 0.0)] for m in merged], dtype=np.float64).reshape(-1, 3)
        out = self._device_modes().bartlett_points(cand, atype=self.atype, multifreq_method=self.multifreq_method)
        return out.cpu().numpy().astype(np.float64)

    def _call_envs(self, merged: List[dict]) -> np.ndarray:
        """Per-candidate environments: host mode provider per candidate, then one batched launch."""
        from .envbatch import EnvBatch
        mode_sets = [self._select(self.runner.modes_for(m)) for m in merged]
        cand = np.array([[float(np.asarray(m["rec_r"])), float(np.asarray(m["src_z"])),
                          float(m.get("tilt", 0.0) or 0.0)] for m in merged], dtype=np.float64).reshape(-1, 3)
        with EnvBatch(mode_sets, cand[:, 0], cand[:, 1], self.covariance_matrix, tilt_deg=cand[:, 2]) as eb:
            out = eb.bartlett(atype=self.atype, multifreq_method=self.multifreq_method)
        return out.cpu().numpy().astype(np.float64)
