"""ctypes binding of libbogp_sm100.so (the C ABI declared in include/bogp.h).

The product path has no CPU fallback: if the shared library is missing, :func:`lib` raises, and every
compute entry point of the library itself fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libbogp_sm100.so"

OK = 0
ATYPE = {"cbf": 0, "cbf_ml": 1}
MF_METHOD = {"mean": 0, "product": 1}
KERNEL = {"matern52": 0, "rbf": 1}
ACQ = {"ei": 0, "logei": 1, "pi": 2, "ucb": 3}
ENGINE = {"auto": 0, "simt": 1, "umma": 2}

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p
_ll = C.c_longlong

# name -> (restype, argtypes); this table is also what tests/test_cabi.py checks against include/bogp.h
SIGNATURES = {
    "bogp_last_error": (C.c_char_p, []),
    "bogp_version": (C.c_int, []),
    "bogp_kernel_launches": (_ll, []),
    "bogp_h2d_bytes": (_ll, []),
    "bogp_timing_enable": (C.c_int, [C.c_int]),
    "bogp_timing_read": (C.c_int, [_dp, C.POINTER(_ll)]),
    "bogp_device_info": (C.c_int, [_ip, _ip, _ip]),
    "bogp_modeset_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _ip, _dp, _dp, _dp, _dp, C.c_int, C.c_double,
                                      C.c_double, _dp, _dp, _dp, C.c_double]),
    "bogp_modeset_destroy": (C.c_int, [_vp]),
    "bogp_modeset_info": (C.c_int, [_vp, C.c_int, _ip, _ip, _ip]),
    "bogp_modeset_info_umma": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bogp_umma_kind": (C.c_int, [_vp]),
    "bogp_covariance_set": (C.c_int, [_vp, _dp, _dp]),
    "bogp_bartlett_grid": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "bogp_bartlett_grid_argext": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _vp,
                                           C.c_int, _vp, _vp, _vp, _ll, _vp]),
    "bogp_bartlett_grid_host": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _vp]),
    "bogp_bartlett_grid_host_argmax": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                                _fp, C.c_int, _fp, C.POINTER(_ll), _vp]),
    "bogp_bartlett_points": (C.c_int, [_vp, _vp, _ll, C.c_int, C.c_int, _vp, _vp]),
    "bogp_bartlett_points_host": (C.c_int, [_vp, _dp, _ll, C.c_int, C.c_int, _fp, _vp]),
    "bogp_envbatch_create": (C.c_int, [C.POINTER(_vp), _ll, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _fp, _dp]),
    "bogp_envbatch_set_covariance": (C.c_int, [_vp, _dp, _dp]),
    "bogp_envbatch_set_tilt": (C.c_int, [_vp, _dp, _dp]),
    "bogp_envbatch_bartlett": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "bogp_envbatch_destroy": (C.c_int, [_vp]),
    "bogp_argmax_f32": (C.c_int, [_vp, _ll, _fp, C.POINTER(_ll), _vp]),
    "bogp_argmin_f32": (C.c_int, [_vp, _ll, _fp, C.POINTER(_ll), _vp]),
    "bogp_argext_f32_dev": (C.c_int, [_vp, _ll, C.c_int, _vp, _vp, _vp, _ll, _vp]),
    "bogp_gp_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _dp, _dp, C.c_int, _dp, C.c_double, C.c_double, C.c_double]),
    "bogp_gp_destroy": (C.c_int, [_vp]),
    "bogp_gp_jitter": (C.c_int, [_vp, _dp]),
    "bogp_gp_acq_host": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, _dp, _ll, _dp, _dp, _vp]),
    "bogp_gp_qei_host": (C.c_int, [_vp, _dp, _ll, C.c_int, _vp, C.c_int, C.c_double, _dp, _dp, _vp]),
    "bogp_gp_fit_adamw": (C.c_int, [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                    C.c_double, _dp, _dp, C.c_int, _dp, _dp, _vp]),
    "bogp_gp_fit_adamw_bounded": (C.c_int, [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                            C.c_double, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp, _vp]),
    "bogp_gp_mll_grad": (C.c_int, [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, _dp, C.c_int,
                                   _dp, _dp, _dp, _vp]),
    "bogp_peer_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _vp]),
    "bogp_peer_open": (C.c_int, [_vp, _vp]),
    "bogp_peer_exchange": (C.c_int, [_vp, _vp, _vp, _vp]),
    "bogp_peer_destroy": (C.c_int, [_vp]),
    "bogp_peer_set_pipelined": (C.c_int, [_vp, C.c_int]),
    "bogp_peer_collect": (C.c_int, [_vp, _vp, _vp]),
    "bogp_bartlett_grid_argext_peer": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _vp,
                                                C.c_int, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _vp]),
    "bogp_bartlett_grid_host_argmax_peer": (C.c_int, [_vp, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                                     _fp, C.c_int, _vp, _ll, C.POINTER(_ll), _vp]),
    "bogp_gp_thompson": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _dp, _vp]),
    "bogp_gp_posterior": (C.c_int, [_vp, _vp, _ll, C.c_int, _vp, _vp, _vp]),
    "bogp_gp_acq": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, _vp, _ll, _vp, _vp, _vp]),
    "bogp_gp_posterior_joint": (C.c_int, [_vp, _vp, _ll, C.c_int, _vp, _vp, _vp]),
    "bogp_gp_posterior_joint_backward": (C.c_int, [_vp, _vp, _ll, C.c_int, _vp, _vp, _vp, _vp]),
    "bogp_gp_qei": (C.c_int, [_vp, _vp, _ll, C.c_int, _vp, C.c_int, C.c_double, _vp, _vp, _vp]),
}

_lib: Optional[C.CDLL] = None


class BogpError(RuntimeError):
    """An entry point of libbogp_sm100.so returned a negative status."""

    def __init__(self, fn: str, code: int, text: str):
        super().__init__(f"{fn} failed ({code}): {text}")
        self.code = code


def lib() -> C.CDLL:
    """Load (once) the shared library; raises if it has not been built (``python -m bogp_b200.build``)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m bogp_b200.build` "
                "(bogp_b200 has no CPU fallback; the CUDA extension is required)")
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(fn: str, rc: int) -> None:
    if rc != OK:
        raise BogpError(fn, rc, (lib().bogp_last_error() or b"").decode(errors="replace"))


_FN: dict = {}


def call(fn: str, *args) -> None:
    f = _FN.get(fn)
    if f is None:
        f = _FN[fn] = getattr(lib(), fn)
    rc = f(*args)
    if rc != OK:
        check(fn, rc)


def dptr(a):
    """numpy float64 array -> double* (or NULL for None)."""
    return None if a is None else a.ctypes.data_as(_dp)


def fptr(a):
    return None if a is None else a.ctypes.data_as(_fp)


def current_stream_ptr() -> int:
    """Raw ``cudaStream_t`` of torch's current stream on the current device (the C-level query when this torch has it: the
    Stream object costs ~3 us per call, which the host-buffer entry points pay on every step)."""
    import torch
    try:
        return int(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
    except AttributeError:
        return int(torch.cuda.current_stream().cuda_stream)


def device_info():
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    call("bogp_device_info", C.byref(sm), C.byref(maj), C.byref(mnr))
    return sm.value, maj.value, mnr.value


def kernel_launches() -> int:
    return int(lib().bogp_kernel_launches())


def h2d_bytes() -> int:
    """Input bytes the Bartlett entry points have copied host -> device since the library was loaded."""
    return int(lib().bogp_h2d_bytes())


def timing_enable(on: bool) -> None:
    call("bogp_timing_enable", 1 if on else 0)


def timing_read():
    """(summed dominant-kernel milliseconds, launches) since the last read; synchronises the recorded events."""
    ms, n = C.c_double(), _ll()
    call("bogp_timing_read", C.byref(ms), C.byref(n))
    return ms.value, n.value
