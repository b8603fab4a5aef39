"""One tilted-array volume (1000 r x 100 z x 10 tilt, 13 tonals) for an ncu capture of umma_grid_kernel<true>."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bogp_b200 import mfp
from bogp_b200.modes import SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
modes = pekeris_modes(SWELLEX_TONALS_13)
dms = mfp.DeviceModeSet(modes, synthetic_covariance(modes, 60.0, 5.0, tilt_deg=1.0))
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 10
r, z, tilt = np.linspace(0.1, 10.0, 1000), np.linspace(1.0, 200.0, 100), np.linspace(-4.0, 4.0, nt)
out = dms.bartlett_grid(r, z, tilt, engine="umma")
torch.cuda.synchronize()
print(float(out.max()))
