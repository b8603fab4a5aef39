"""Config-2 surface calls for a launch list (ncu --metrics gpu__time_duration.sum): table kernels and the surface kernel."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bogp_b200 import mfp
from bogp_b200.modes import SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
modes = pekeris_modes(SWELLEX_TONALS_13)
r, z = np.linspace(0.1, 10.0, 1000), np.linspace(1.0, 200.0, 1000)
dms = mfp.DeviceModeSet(modes, synthetic_covariance(modes, z[296], r[495]))
out = torch.empty((1000, 1000), dtype=torch.float32, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    dms.bartlett_grid(r, z, engine="umma", out=out)
torch.cuda.synchronize()
print(float(out.max()))
