"""Prints the per-role wait counters of the tcgen05 grid kernel on the full 1000x1000x13 surface (BOGP_UMMA_DBG=1)."""
import os, sys
from pathlib import Path
os.environ["BOGP_UMMA_DBG"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bogp_b200 import mfp
from bogp_b200.modes import SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
modes = pekeris_modes(SWELLEX_TONALS_13)
r, z = np.linspace(0.1, 10.0, 1000), np.linspace(1.0, 200.0, 1000)
K = synthetic_covariance(modes, z[296], r[495])
dms = mfp.DeviceModeSet(modes, K)
for i in range(3):
    print("--- call", i, file=sys.stderr, flush=True)
    dms.bartlett_grid(r, z, engine="umma")
    torch.cuda.synchronize()
