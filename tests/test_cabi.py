"""CPU tier: the C-ABI library loads, exports every symbol include/bogp.h declares, and refuses to compute on CPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from bogp_b200 import _lib

HEADER = Path(__file__).resolve().parent.parent / "include" / "bogp.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(bogp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bogp.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in bogp_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.bogp_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point fails loudly (no CPU fallback)."""
    sm = C.c_int()
    with pytest.raises(_lib.BogpError) as e:
        _lib.call("bogp_device_info", C.byref(sm), None, None)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    from bogp_b200.modes import INV_TONALS_3, REC_Z_21, pekeris_modes
    from bogp_b200.mfp import DeviceModeSet
    with pytest.raises(_lib.BogpError):
        DeviceModeSet(pekeris_modes(INV_TONALS_3, REC_Z_21))
    from bogp_b200.gp import SingleTaskGP
    with pytest.raises(RuntimeError):
        SingleTaskGP(np.zeros((4, 2)), np.zeros(4)).handle()


def test_argument_validation_before_device():
    h = C.c_void_p()
    M = (C.c_int * 1)(4)
    with pytest.raises(_lib.BogpError) as e:
        _lib.call("bogp_modeset_create", C.byref(h), 0, 4, M, None, None, None, None, 2, 0.0, 1.0, None, None, None, 0.0)
    assert e.value.code == -1
    with pytest.raises(_lib.BogpError) as e:
        _lib.call("bogp_gp_create", C.byref(h), 0, 2, None, None, 0, None, 1.0, 0.0, 0.0)
    assert e.value.code == -1
