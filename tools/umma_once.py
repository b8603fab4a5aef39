"""One warm surface of config 2 through the tcgen05 engine (the command ncu wraps for a source-level capture)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bogp_b200 import mfp
from bogp_b200.modes import SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
modes = pekeris_modes(SWELLEX_TONALS_13)
r, z = np.linspace(0.1, 10.0, 1000), np.linspace(1.0, 200.0, 1000)
dms = mfp.DeviceModeSet(modes, synthetic_covariance(modes, z[296], r[495]))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    dms.bartlett_grid(r, z, engine="umma")
torch.cuda.synchronize()
print("ok")
