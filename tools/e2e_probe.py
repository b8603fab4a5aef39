import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bogp_b200 import mfp, _lib
from bogp_b200.modes import SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
modes = pekeris_modes(SWELLEX_TONALS_13)
K = synthetic_covariance(modes, 60.0, 5.0)
dms = mfp.DeviceModeSet(modes, K)
r, z = np.linspace(0.1, 10.0, 1000), np.linspace(1.0, 200.0, 1000)
out = torch.empty((1000, 1000), dtype=torch.float32).pin_memory().numpy()
def t(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("host argext pinned out :", t(lambda: dms.bartlett_grid_host(r, z, out=out, argext="max")))
print("host no argext pinned  :", t(lambda: dms.bartlett_grid_host(r, z, out=out)))
outp = np.empty((1000, 1000), dtype=np.float32)
print("host argext pageable   :", t(lambda: dms.bartlett_grid_host(r, z, out=outp, argext="max")))
dout = torch.empty((1000, 1000), dtype=torch.float32, device="cuda")
pair = torch.empty(4, dtype=torch.int64, device="cuda")
print("device grid            :", t(lambda: dms.bartlett_grid(r, z, out=dout)))
os.environ["BOGP_NO_ZEROCOPY"] = "1"
print("host argext, no zerocopy:", t(lambda: dms.bartlett_grid_host(r, z, out=out, argext="max")))
