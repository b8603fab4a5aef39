"""A few fused MC-qEI evaluations at BASELINE configs[2] sizes (n = 256, d = 3, q = 4, 64 restarts, 512 MC samples): the
command ncu wraps for the qei_kernel capture."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bogp_b200 import acquisition as A
from bogp_b200.gp import SingleTaskGP
rng = np.random.default_rng(0)
X = rng.uniform(-1, 1, (256, 3)); y = np.sin(3 * X[:, 0]) + 0.3 * np.cos(2 * X.sum(1)); y = (y - y.mean()) / y.std()
model = SingleTaskGP(X, y, lengthscale=[0.5] * 3, outputscale=1.0, noise=1e-4)
acq = A.qExpectedImprovement(model, float(y.max()) - 0.3, q=4, mc_samples=512, seed=0)
Xq = torch.as_tensor(rng.uniform(-1, 1, (64, 4, 3)), device="cuda").requires_grad_(True)
for _ in range(6):
    v = acq(Xq)
    torch.autograd.grad(v.sum(), Xq)
torch.cuda.synchronize()
print("ok")
