"""GPU debug harness for the tcgen05 grid kernel: UMMA vs SIMT vs oracle on small and full-size grids.
Each case runs in a subprocess under a timeout so a protocol bug traps/time-outs instead of hanging the box.

    python tools/umma_debug.py [case ...]         # driver: runs the cases (default: all) in subprocesses
    python tools/umma_debug.py case NAME          # one case in-process
"""
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def case(name: str) -> None:
    import torch
    from bogp_b200 import mfp
    from bogp_b200.modes import INV_TONALS_3, REC_Z_21, SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
    if name.startswith("small"):
        modes = pekeris_modes(INV_TONALS_3, REC_Z_21)
        K = synthetic_covariance(modes, 71.11, 1.07)
        r, z = np.linspace(0.1, 8.0, 200), np.linspace(1.0, 200.0, 40)
    elif name.startswith("rank8"):
        modes = pekeris_modes(SWELLEX_TONALS_13)
        K = synthetic_covariance(modes, 60.0, 5.0, snapshots=8, seed=3)
        r, z = np.linspace(0.1, 10.0, 300), np.linspace(1.0, 200.0, 70)
    else:
        modes = pekeris_modes(SWELLEX_TONALS_13)
        r, z = np.linspace(0.1, 10.0, 1000), np.linspace(1.0, 200.0, 1000)
        K = synthetic_covariance(modes, z[296], r[495])
    dms = mfp.DeviceModeSet(modes, K)
    print(name, "tonal_info", dms.tonal_info(), flush=True)
    kw = dict(atype="cbf_ml", multifreq_method="product") if name.endswith("ml") else {}
    simt = dms.bartlett_grid(r, z, engine="simt", **kw).cpu().numpy()
    t0 = time.time()
    umma = dms.bartlett_grid(r, z, engine="umma", **kw)
    torch.cuda.synchronize()
    umma = umma.cpu().numpy()
    print(name, "umma ran in", time.time() - t0, "s", flush=True)
    den = np.maximum(np.abs(simt), 0.05 * np.abs(simt).max())
    err = np.abs(umma - simt) / den
    print(name, "max rel err umma vs simt", float(np.nanmax(err)), "nan", int(np.isnan(umma).sum()),
          "argmax", np.unravel_index(np.argmax(umma), umma.shape), np.unravel_index(np.argmax(simt), simt.shape))
    if not np.nanmax(err) <= 1e-4:
        bad = np.argwhere(~(err <= 1e-4))
        print(name, "bad count", len(bad), "of", err.size, "first", bad[:8].tolist())
        print("umma", umma[:2, :6], "\nsimt", simt[:2, :6])
    if not name.startswith("full"):
        from oracle import bartlett as ob
        ref = ob.ambiguity_surface(modes, K, r, z, **kw)
        e2 = np.abs(umma - ref) / np.maximum(np.abs(ref), 0.05 * np.abs(ref).max())
        e3 = np.abs(simt - ref) / np.maximum(np.abs(ref), 0.05 * np.abs(ref).max())
        print(name, "max rel err vs oracle: umma", float(np.nanmax(e2)), "simt", float(np.nanmax(e3)))
    else:
        for _ in range(3):
            dms.bartlett_grid(r, z, engine="umma")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            dms.bartlett_grid(r, z, engine="umma")
        b.record()
        torch.cuda.synchronize()
        print(name, "umma ms/surface", a.elapsed_time(b) / 10)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "case":
        case(sys.argv[2])
        sys.exit(0)
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    cases = args or ["small", "small_ml", "rank8", "full"]
    swaps = ("0", "1") if "--swap" in sys.argv else ("0",)
    for swap in swaps:
        for c in cases:
            env = dict(os.environ, BOGP_UMMA_SWAP=swap)
            print(f"=== case {c} swap={swap}", flush=True)
            try:
                r = subprocess.run([sys.executable, __file__, "case", c], env=env, timeout=150, capture_output=True,
                                   text=True)
                print(r.stdout[-3000:], r.stderr[-1500:], "rc", r.returncode, flush=True)
            except subprocess.TimeoutExpired as e:
                print("TIMEOUT", (e.stdout or b"")[-2000:], flush=True)
                break
