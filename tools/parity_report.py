"""Parity margins of the Bartlett engines against the float64 oracle on the test configurations: floored
(|got - ref| / max(|ref|, 0.05 max|ref|), the tests' bar) AND un-floored (|got - ref| / |ref|: max, p99, share of nodes
above 1e-5).  Engines: CUDA cores, tcgen05 FP16-split (default), tcgen05 3xTF32 (BOGP_UMMA_KIND=tf32).
GPU box:  python tools/parity_report.py  > profiles/parity_report_r2.txt"""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from bogp_b200.mfp import DeviceModeSet
from bogp_b200.modes import INV_TONALS_3, REC_Z_21, SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
from oracle import bartlett as ob

ENGINES = (("simt", "simt", {}), ("umma f16", "umma", {}), ("umma tf32", "umma", {"BOGP_UMMA_KIND": "tf32"}))


def report(name, fn, ref):
    for tag, eng, env in ENGINES:
        os.environ.update(env)
        got = fn(eng)
        for k in env:
            os.environ.pop(k)
        st = ob.parity_stats(got, ref)
        print(f"{name:44s} {tag:9s} floored max {st['max_floored']:.2e} | un-floored max {st['max_unfloored']:.2e} p99 {st['p99_unfloored']:.2e} "
              f"share>1e-5 {st['frac_unfloored_gt_1e-5']:.1e} | min|ref| {st['min_ref']:.1e} | same arg-max {st['same_argmax']}")


def main():
    cases = {
        "cfg1 (3 tonals, 21 rec, 100x50)": (pekeris_modes(INV_TONALS_3, REC_Z_21), (71.11, 1.07), np.linspace(0.1, 8.0, 100), np.linspace(1.0, 200.0, 50), 0),
        "cfg2 small (13 tonals, 64 rec, 160x24)": (pekeris_modes(SWELLEX_TONALS_13), (60.0, 5.0), np.linspace(0.1, 10.0, 160), np.linspace(1.0, 200.0, 24), 0),
        "cfg2 rank 8 (300x70)": (pekeris_modes(SWELLEX_TONALS_13), (60.0, 5.0), np.linspace(0.3, 10.0, 300), np.linspace(2.0, 199.0, 70), 8),
    }
    for name, (modes, (tz, tr), r, z, snap) in cases.items():
        K = synthetic_covariance(modes, tz, tr, snapshots=snap, seed=3)
        dms = DeviceModeSet(modes, K)
        ref = ob.ambiguity_surface(modes, K, r, z)
        report(name, lambda eng: dms.bartlett_grid(r, z, engine=eng).cpu().numpy().astype(np.float64), ref)
    modes = pekeris_modes(SWELLEX_TONALS_13)
    r, z = np.linspace(0.1, 10.0, 1000), np.linspace(1.0, 200.0, 1000)
    K = synthetic_covariance(modes, z[296], r[495])
    dms = DeviceModeSet(modes, K)
    report("BASELINE configs[1], all 10^6 nodes", lambda eng: dms.bartlett_grid(r, z, engine=eng).cpu().numpy().astype(np.float64),
           ob.ambiguity_surface_pool(modes, K, r, z))
    K = synthetic_covariance(modes, 60.0, 5.0, tilt_deg=1.5)
    r, z, tilt = np.linspace(0.4, 9.7, 150), np.linspace(3.0, 195.0, 3), np.array([-3.5, 0.0, 1.5, 4.0, 2.5])
    cand = np.array([[ri, zi, ti] for ti in tilt for zi in z for ri in r])
    ref = ob.bartlett_points(modes, K, cand).reshape(len(tilt), len(z), len(r))
    dms = DeviceModeSet(modes, K)
    for tag, eng, env in ENGINES[:2]:
        got = dms.bartlett_grid(r, z, tilt, engine=eng).cpu().numpy().astype(np.float64)
        st = ob.parity_stats(got, ref)
        print(f"{'tilt volume (13 tonals, 64 rec, 150x3x5)':44s} {('umma tf32 (tilt engine)' if eng == 'umma' else tag):9s} floored max {st['max_floored']:.2e} | un-floored max "
              f"{st['max_unfloored']:.2e} p99 {st['p99_unfloored']:.2e} share>1e-5 {st['frac_unfloored_gt_1e-5']:.1e} | min|ref| {st['min_ref']:.1e} | same arg-max {st['same_argmax']}")


if __name__ == "__main__":
    main()
