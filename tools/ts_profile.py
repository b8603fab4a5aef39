"""Thompson-sampling kernel chain (K2-TS): parity against the oracle at a small size, timing at the reference's size
(b = 5000 candidates, baxus.py:45).   python tools/ts_profile.py [b] [n] [d]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bogp_b200 import _lib  # noqa: E402
from bogp_b200.generation import MaxPosteriorSampling  # noqa: E402
from bogp_b200.gp import SingleTaskGP  # noqa: E402
from oracle import gp as ogp  # noqa: E402


def make(n, d, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(3 * X[:, 0]) + 0.3 * np.cos(2 * X.sum(1))
    y = (y - y.mean()) / y.std()
    ls = rng.uniform(0.3, 1.5, d)
    return X, y, ls


def parity(b=300, n=40, d=3, S=3):
    X, y, ls = make(n, d)
    rng = np.random.default_rng(1)
    Xc = rng.uniform(-1, 1, (b, d))
    z = rng.standard_normal((S, b))
    m = SingleTaskGP(X, y, lengthscale=ls, outputscale=1.3, noise=1e-4)
    ts = MaxPosteriorSampling(m)
    ts(Xc, num_samples=S, base_samples=z, return_samples=True)
    st = ogp.GPState(X, y, "matern52", ls, 1.3, 0.0, 1e-4)
    idx, f, jit = ogp.thompson_sample(st, Xc, z)
    fs = ts.last_samples.cpu()
    err = float((fs - f).abs().max() / f.abs().max())
    print(f"parity b={b} n={n} d={d}: idx gpu {ts.last_indices.tolist()} oracle {idx.tolist()} jitter {ts.last_jitter} / {jit} "
          f"max rel sample err {err:.2e}")


def perf(b=5000, n=500, d=12):
    X, y, ls = make(n, d)
    rng = np.random.default_rng(1)
    Xc = rng.uniform(-1, 1, (b, d))
    m = SingleTaskGP(X, y, lengthscale=ls, outputscale=1.3, noise=1e-4)
    ts = MaxPosteriorSampling(m)
    z = rng.standard_normal((1, b))
    for _ in range(2):
        ts(Xc, base_samples=z)
    torch.cuda.synchronize(); _lib.timing_read(); _lib.timing_enable(True)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ts(Xc, base_samples=z)
    torch.cuda.synchronize()
    call_ms = (time.perf_counter() - t0) / reps * 1e3
    _lib.timing_enable(False)
    kms, kn = _lib.timing_read()
    flop = 2 * (b ** 3 / 3 + b * b * n / 2 * 1.0 + b * n * n / 2)
    print(f"perf b={b} n={n} d={d}: call {call_ms:.2f} ms, covariance+factorisation {kms / max(kn, 1):.2f} ms "
          f"({flop / (kms / max(kn, 1) * 1e-3) / 1e12:.2f} TFLOP/s f64), jitter {ts.last_jitter}")
    t0 = time.perf_counter()
    st = ogp.GPState(X, y, "matern52", ls, 1.3, 0.0, 1e-4)
    idx, f, jit = ogp.thompson_sample(st, Xc, z)
    print(f"  oracle (torch CPU, {torch.get_num_threads()} threads): {(time.perf_counter() - t0) * 1e3:.0f} ms, idx {idx.tolist()} vs gpu {ts.last_indices.tolist()}, jitter {jit}")


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    torch.cuda.set_device(0)
    parity()
    parity(b=1000, n=137, d=20, S=1)
    perf(*a)
