"""One env-batch launch (1e4 environments x 3 tonals x 21 receivers) for an ncu capture of envstream_kernel."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bogp_b200.envbatch import EnvBatch
from bogp_b200.modes import INV_TONALS_3, REC_Z_21, pekeris_modes, perturbed_pekeris_batch, synthetic_covariance
rng = np.random.default_rng(2)
base = perturbed_pekeris_batch(INV_TONALS_3, REC_Z_21, rng.uniform(200.0, 230.0, 200), rng.uniform(1600.0, 1700.0, 200))
ref = pekeris_modes(INV_TONALS_3, REC_Z_21)
Cn = 10_000
eb = EnvBatch([base[i % 200] for i in range(Cn)], rng.uniform(0.5, 6.0, Cn), rng.uniform(10.0, 190.0, Cn), synthetic_covariance(ref, 71.11, 1.07))
out = torch.empty(Cn, dtype=torch.float32, device="cuda")
for _ in range(3):
    eb.bartlett(out=out)
torch.cuda.synchronize()
print(float(out.max()))
