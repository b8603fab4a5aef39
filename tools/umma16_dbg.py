"""Debug build of the FP16-split tcgen05 kernel on config 2: per-tonal issuer / epilogue time stamps of CTA 0, then the
bracketing runs (0 everything, 1 no MMAs issued, 2 epilogue releases without reading).
  BOGP_UMMA_DBG_BUILD=1 python -m bogp_b200.build --force && gpurun -- python tools/umma16_dbg.py"""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) > 1:
    sys.path.insert(0, str(ROOT))
    import numpy as np, torch
    from bogp_b200 import _lib, mfp
    from bogp_b200.modes import SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
    modes = pekeris_modes(SWELLEX_TONALS_13)
    r, z = np.linspace(0.1, 10.0, 1000), np.linspace(1.0, 200.0, 1000)
    dms = mfp.DeviceModeSet(modes, synthetic_covariance(modes, z[296], r[495]))
    for _ in range(3):
        dms.bartlett_grid(r, z, engine="umma")
    torch.cuda.synchronize()
    if sys.argv[1] == "stamps":
        os.environ["BOGP_UMMA_DBG"] = "1"
        dms.bartlett_grid(r, z, engine="umma")
        torch.cuda.synchronize()
        sys.exit(0)
    _lib.timing_read(); _lib.timing_enable(True)
    for _ in range(10):
        dms.bartlett_grid(r, z, engine="umma")
    torch.cuda.synchronize()
    ms, n = _lib.timing_read()
    print(f"skip={os.environ.get('BOGP_UMMA_SKIP','0')}: kernel {ms / n:.4f} ms")
else:
    subprocess.run([sys.executable, __file__, "stamps"], timeout=120)
    for sk in (sys.argv[2:] if False else ("0", "1", "2", "4", "6", "7")):
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, BOGP_UMMA_SKIP=sk), timeout=120)
