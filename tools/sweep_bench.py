"""Time-step sweep timing: T surfaces of the reference's mfp.yaml shape (200 x 100 grid, 13 tonals,
64 receivers), one covariance per step.  Prints per-step wall time split into covariance set-up and surface."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bogp_b200 import sweep
from bogp_b200.mfp import MatchedFieldProcessor, ModeRunner, beamformer
from bogp_b200.modes import SWELLEX_TONALS_13, pekeris_modes, synthetic_covariance
modes = pekeris_modes(SWELLEX_TONALS_13)
T = 40
K = np.stack([synthetic_covariance(modes, 60.0 + 0.5 * t, 3.0 + 0.05 * t) for t in range(T)], axis=1)      # [F, T, N, N]
rvec, zvec = np.linspace(0.1, 10.0, 200), np.linspace(1.0, 200.0, 100)
mfp = MatchedFieldProcessor(runner=ModeRunner(modes), covariance_matrix=K[:, 0], freq=modes.freq, parameters={}, beamformer=beamformer)
sweep.ambiguity_sweep(mfp, K[:, :3], rvec, zvec)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = sweep.ambiguity_sweep(mfp, K, rvec, zvec, keep_surfaces=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
dms = mfp._device_modes()
t1 = time.perf_counter()
for t in range(T):
    dms.set_covariance(np.ascontiguousarray(K[:, t]))
tc = time.perf_counter() - t1
print(f"sweep: {T} steps in {dt*1e3:.1f} ms = {dt/T*1e3:.3f} ms/step; covariance set-up alone {tc/T*1e3:.3f} ms/step; "
      f"peaks follow the track: {[p[1] for p in res['peak'][:3]]} ... {[p[1] for p in res['peak'][-2:]]}")
