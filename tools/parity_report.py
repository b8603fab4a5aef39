"""Parity margins of the Bartlett engines against the float64 oracle on the test configurations (max of
|got - ref| / max(|ref|, 0.05 max|ref|); bar 1e-5)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from bogp_b200.mfp import DeviceModeSet
from bogp_b200.modes import INV_TONALS_3, REC_Z_21, SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
from oracle import bartlett as ob

def err(got, ref):
    e = np.abs(got - ref) / np.maximum(np.abs(ref), 0.05 * np.abs(ref).max())
    return e.max(), np.unravel_index(e.argmax(), e.shape)

cases = {
    "cfg1 (3 tonals, 21 rec, 100x50)": (pekeris_modes(INV_TONALS_3, REC_Z_21), (71.11, 1.07), np.linspace(0.1, 8.0, 100), np.linspace(1.0, 200.0, 50), 0),
    "cfg2 small (13 tonals, 64 rec, 160x24)": (pekeris_modes(SWELLEX_TONALS_13), (60.0, 5.0), np.linspace(0.1, 10.0, 160), np.linspace(1.0, 200.0, 24), 0),
    "cfg2 rank 8 (300x70)": (pekeris_modes(SWELLEX_TONALS_13), (60.0, 5.0), np.linspace(0.3, 10.0, 300), np.linspace(2.0, 199.0, 70), 8),
}
for name, (modes, (tz, tr), r, z, snap) in cases.items():
    K = synthetic_covariance(modes, tz, tr, snapshots=snap, seed=3)
    dms = DeviceModeSet(modes, K)
    ref = ob.ambiguity_surface(modes, K, r, z)
    for eng in ("simt", "umma"):
        e, pos = err(dms.bartlett_grid(r, z, engine=eng).cpu().numpy().astype(np.float64), ref)
        print(f"{name:42s} {eng:5s} max rel err {e:.2e} at (z {z[pos[0]]:.1f} m, r {r[pos[1]]:.2f} km)")
modes = pekeris_modes(SWELLEX_TONALS_13)
K = synthetic_covariance(modes, 60.0, 5.0, tilt_deg=1.5)
r, z, tilt = np.linspace(0.4, 9.7, 150), np.linspace(3.0, 195.0, 3), np.array([-3.5, 0.0, 1.5, 4.0, 2.5])
cand = np.array([[ri, zi, ti] for ti in tilt for zi in z for ri in r])
ref = ob.bartlett_points(modes, K, cand).reshape(len(tilt), len(z), len(r))
dms = DeviceModeSet(modes, K)
for eng in ("simt", "umma"):
    e, pos = err(dms.bartlett_grid(r, z, tilt, engine=eng).cpu().numpy().astype(np.float64), ref)
    print(f"{'tilt volume (13 tonals, 64 rec, 150x3x5)':42s} {eng:5s} max rel err {e:.2e} at {pos}")
