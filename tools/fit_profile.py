"""One fused GP fit (n=144, d=2, 100 AdamW steps) for an ncu capture of gp_fit_kernel."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bogp_b200.gp import FitConfig, SingleTaskGP, fit_gp_adamw
n, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (144, 2)
rng = np.random.default_rng(0)
X = rng.uniform(-1, 1, (n, d)); y = np.sin(3 * X[:, 0]) + 0.3 * np.cos(2 * X.sum(1)); y = (y - y.mean()) / y.std()
m = fit_gp_adamw(SingleTaskGP(X, y), FitConfig(steps=100))
torch.cuda.synchronize()
print(m.lengthscale.numpy(), m.outputscale, m.noise)
