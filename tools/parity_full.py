"""Whole BASELINE configs[1] surface (1000 x 1000 x 13 tonals) of every engine against the float64 oracle computed over a
process pool: un-floored and floored relative-error distribution, arg-max.  GPU box:  python tools/parity_full.py"""
import json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np


def main():
    from bogp_b200.mfp import DeviceModeSet
    from bogp_b200.modes import SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
    from oracle import bartlett as ob
    modes = pekeris_modes(SWELLEX_TONALS_13)
    r, z = np.linspace(0.1, 10.0, 1000), np.linspace(1.0, 200.0, 1000)
    K = synthetic_covariance(modes, z[296], r[495])
    t0 = time.perf_counter()
    ref = ob.ambiguity_surface_pool(modes, K, r, z)
    print(f"oracle: {time.perf_counter() - t0:.1f} s on {os.cpu_count()} cores", file=sys.stderr)
    dms = DeviceModeSet(modes, K)
    for tag, engine, env in (("simt", "simt", {}), ("umma_f16", "umma", {}), ("umma_tf32", "umma", {"BOGP_UMMA_KIND": "tf32"})):
        os.environ.update(env)
        got = dms.bartlett_grid(r, z, engine=engine).cpu().numpy()
        for k in env:
            os.environ.pop(k)
        print(json.dumps({"engine": tag, **ob.parity_stats(got, ref)}))


if __name__ == "__main__":
    main()
