"""Weak scaling of the two other sharded Bartlett paths on the GPUs of one box ("measure scaling on 10^7 as well"):
configs[3] -- the (tilt, depth, range) volume, whole tilt planes per rank (sharding.sharded_tilt_volume), 25 tilts = 2.5e6
candidates per rank -- and configs[4] -- 10^4 per-candidate environments per rank (sharding.sharded_candidates).  Device
time of a step (CUDA events) as the max over ranks, the 16-byte extremum exchange by NCCL inside the step.
  python tools/scale_paths.py            |  torchrun --nproc-per-node N tools/scale_paths.py"""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch, torch.distributed as dist
from bogp_b200 import mfp
from bogp_b200.envbatch import EnvBatch
from bogp_b200.modes import (INV_TONALS_3, REC_Z_21, SWELLEX_TONALS_13, pekeris_modes, perturbed_pekeris_batch, synthetic_covariance)
from bogp_b200.sharding import sharded_candidates, sharded_tilt_volume

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def timed(fn, reps=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# configs[3]: 1000 r x 100 z x (25 tilts per rank)
modes = pekeris_modes(SWELLEX_TONALS_13)
r, z = np.linspace(0.1, 10.0, 1000), np.linspace(1.0, 200.0, 100)
tilt = np.linspace(-4.0, 4.0, 25 * world)
dms = mfp.DeviceModeSet(modes, synthetic_covariance(modes, z[30], r[495], tilt_deg=tilt[7]))
res = {}
def tilt_step():
    res["tilt"] = sharded_tilt_volume(lambda lo, hi: dms.bartlett_grid(r, z, tilt[lo:hi]), len(tilt), len(z), len(r))[1]
ms = timed(tilt_step)
if rank == 0:
    C = len(r) * len(z) * len(tilt)
    print(json.dumps({"path": "configs[3] tilt volume, weak (25 tilt planes = 2.5e6 candidates per rank)", "n_gpus": world, "ms_per_step": ms,
                      "cand_freq_per_s": C * 13 / (ms * 1e-3), "argmax": list(res["tilt"][1]), "truth": [7, 30, 495]}))

# configs[4]: 10^4 environments per rank (200 distinct mode sets re-used, as tools/bench_paths.py does)
rng = np.random.default_rng(3)
base = perturbed_pekeris_batch(INV_TONALS_3, REC_Z_21, rng.uniform(200.0, 230.0, 200), rng.uniform(1600.0, 1700.0, 200))
truth = pekeris_modes(INV_TONALS_3, REC_Z_21)
K = synthetic_covariance(truth, 71.11, 1.07)
Cn = 10_000 * world
env = rng.integers(0, 200, Cn)
rr, zs = rng.uniform(0.9, 1.3, Cn), rng.uniform(60.0, 80.0, Cn)
from bogp_b200.sharding import slab
lo, hi = slab(Cn, rank, world)
eb = EnvBatch([base[e] for e in env[lo:hi]], rr[lo:hi], zs[lo:hi], K)
def env_step():
    res["env"] = sharded_candidates(lambda a, b: eb.bartlett(atype="cbf_ml", multifreq_method="product"), Cn, maximize=False)[1]
ms = timed(env_step)
if rank == 0:
    print(json.dumps({"path": "configs[4] per-candidate environments, weak (1e4 environments per rank)", "n_gpus": world, "ms_per_step": ms,
                      "cand_freq_per_s": Cn * 3 / (ms * 1e-3), "argmin": res["env"][1]}))
eb.close()
if world > 1:
    dist.destroy_process_group()
