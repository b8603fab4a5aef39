"""Roofline numbers for the other kernels of the path: the GP posterior/acquisition kernel (K2),
the scattered-candidate and tilted-array Bartlett kernels (K1-SIMT) and the per-environment kernel (K1-ENV).

    python tools/bench_paths.py [gp] [fit] [points] [tilt] [env] [ts] > profiles/paths_<tag>.jsonl

One JSON line per case.  Kernel time = CUDA events around the dominant kernel inside the library
(bogp_timing_enable/read) where it is instrumented, CUDA events around the call otherwise; 3 warm-up + 10 timed calls.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from bogp_b200 import _lib, mfp  # noqa: E402
from bogp_b200 import acquisition as A  # noqa: E402
from bogp_b200.gp import SingleTaskGP  # noqa: E402
from bogp_b200.modes import (INV_TONALS_3, REC_Z_21, SWELLEX_TONALS_13, pekeris_modes, perturbed_pekeris_batch,  # noqa: E402
                             synthetic_covariance)

HBM_GBS = 6547.5
try:
    HBM_GBS = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=10, warm=3, lib_timing=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    _lib.timing_read()
    _lib.timing_enable(lib_timing)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    _lib.timing_enable(False)
    kms, kn = _lib.timing_read()
    return a.elapsed_time(b) / reps, (kms / kn if kn else None)


def bench_gp():
    rng = np.random.default_rng(0)
    for n in (128, 256, 512):
        for d in (2, 3, 7):
            X = rng.uniform(-1, 1, (n, d))
            y = np.sin(3 * X[:, 0]) + 0.3 * np.cos(2 * X.sum(1))
            y = (y - y.mean()) / y.std()
            model = SingleTaskGP(X, y, lengthscale=[0.5] * d, outputscale=1.0, noise=1e-4)
            ei = A.ExpectedImprovement(model, float(y.max()))
            b = 4096
            Xs = torch.as_tensor(rng.uniform(-1, 1, (b, 1, d)), device="cuda")
            with torch.no_grad():
                call_ms, k_ms = timed(lambda: ei(Xs))
            Xg = Xs.clone().requires_grad_(True)

            def fwd_bwd():
                v = ei(Xg)
                torch.autograd.grad(v.sum(), Xg)
            callg_ms, kg_ms = timed(fwd_bwd)
            flop_f = n * (3 * d + 24) + n * n
            flop_b = flop_f + n * n + 6 * n * d
            print(json.dumps({"kernel": "gp_eval_kernel (K2)", "case": f"n={n} d={d} b={b} EI", "dtype": "f64",
                              "forward_kernel_ms": k_ms, "forward_call_ms": call_ms,
                              "forward_cand_per_s": b / (k_ms * 1e-3) if k_ms else None,
                              "forward_tflops": b * flop_f / (k_ms * 1e-3) / 1e12 if k_ms else None,
                              "value_grad_kernel_ms": kg_ms, "value_grad_call_ms": callg_ms,
                              "value_grad_tflops": b * flop_b / (kg_ms * 1e-3) / 1e12 if kg_ms else None,
                              "alg_flop_per_cand_forward": flop_f, "alg_flop_per_cand_value_grad": flop_b,
                              "alg_bytes_per_cand": 8 * d + 8, "hbm_GBs_forward": b * (8 * d + 8) / (k_ms * 1e-3) / 1e9 if k_ms else None,
                              "bound": "fp64 FMA (compute); HBM fraction reported for reference", "hbm_peak_GBs": HBM_GBS}), flush=True)


def bench_fit():
    """K2-FIT: 100 AdamW steps of the GP hyper-parameter fit, fused kernel (one launch) vs the torch/cuSOLVER loop."""
    from bogp_b200.gp import FitConfig, fit_gp_adamw, fit_gp_adamw_torch
    rng = np.random.default_rng(0)
    for n, d in ((64, 2), (144, 2), (144, 7), (256, 3), (512, 7), (1000, 7)):
        X = rng.uniform(-1, 1, (n, d))
        y = np.sin(3 * X[:, 0]) + 0.3 * np.cos(2 * X.sum(1))
        y = (y - y.mean()) / y.std()
        fit_gp_adamw(SingleTaskGP(X, y), FitConfig(steps=5), engine="fused")
        torch.cuda.synchronize(); _lib.timing_read(); _lib.timing_enable(True)
        t0 = time.perf_counter()
        m = fit_gp_adamw(SingleTaskGP(X, y), FitConfig(steps=100), engine="fused")
        call_ms = (time.perf_counter() - t0) * 1e3
        _lib.timing_enable(False)
        k_ms, _ = _lib.timing_read()
        fit_gp_adamw_torch(SingleTaskGP(X, y), FitConfig(steps=3)); torch.cuda.synchronize()
        t0 = time.perf_counter()
        t = fit_gp_adamw_torch(SingleTaskGP(X, y), FitConfig(steps=100)); torch.cuda.synchronize()
        torch_ms = (time.perf_counter() - t0) * 1e3
        flop = 100 * (n ** 3 / 3 + n ** 3 / 3 + n ** 3 / 3 + 2 * n * n * (3 * d + 25) / 2)      # chol + inverse + X^T X + 2 elementwise passes
        print(json.dumps({"kernel": "gp_fit_kernel (K2-FIT)", "case": f"n={n} d={d} Matern-5/2, 100 AdamW steps, " + ("one CTA" if n <= 144 else "cooperative grid"),
                          "dtype": "f64", "kernel_ms": k_ms, "call_ms": call_ms, "torch_cusolver_loop_ms": torch_ms,
                          "speedup_vs_torch_loop": torch_ms / call_ms, "ms_per_step": k_ms / 100, "alg_gflop": flop / 1e9,
                          "gflops": flop / (k_ms * 1e-3) / 1e9,
                          "lengthscale_rel_diff_vs_torch": float(np.max(np.abs(m.lengthscale.numpy() - t.lengthscale.numpy()) / t.lengthscale.numpy())),
                          "bound": "latency of a sequential factorisation on one SM (fp64); no launch or host round trip between steps"}), flush=True)


def bench_points():
    modes = pekeris_modes(SWELLEX_TONALS_13)
    K = synthetic_covariance(modes, 60.0, 5.0)
    dms = mfp.DeviceModeSet(modes, K)
    rng = np.random.default_rng(1)
    flop = dms.flops_per_candidate()
    for C in (4096, 1_000_000):
        cand = np.stack([rng.uniform(0.1, 10.0, C), rng.uniform(1.0, 200.0, C), np.zeros(C)], 1)
        cd = torch.as_tensor(cand, device="cuda")
        call_ms, k_ms = timed(lambda: dms.bartlett_points(cd))
        print(json.dumps({"kernel": "points_modespace_kernel (K1-SIMT)", "case": f"C={C} scattered, 13 tonals, no tilt",
                          "kernel_ms": k_ms, "call_ms": call_ms, "cand_freq_per_s": C * 13 / (k_ms * 1e-3),
                          "alg_tflops": C * flop / (k_ms * 1e-3) / 1e12, "bound": "fp32 FMA issue + sincos (CUDA cores)"}), flush=True)


def bench_tilt():
    modes = pekeris_modes(SWELLEX_TONALS_13)
    K = synthetic_covariance(modes, 60.0, 5.0, tilt_deg=1.0)
    dms = mfp.DeviceModeSet(modes, K)
    N, sumM = 64, sum(modes.n_modes)
    r, z = np.linspace(0.1, 10.0, 1000), np.linspace(1.0, 200.0, 100)
    import os
    for nt, engines in ((10, ("simt", "umma", "umma_tf32")), (100, ("simt", "umma", "umma_tf32"))):
        tilt = np.linspace(-4.0, 4.0, nt)
        C = len(r) * len(z) * len(tilt)
        out = torch.empty((len(tilt), len(z), len(r)), dtype=torch.float32, device="cuda")
        for engine_tag in engines:
            engine = "umma" if engine_tag.startswith("umma") else engine_tag
            if engine_tag == "umma_tf32":
                os.environ["BOGP_UMMA_KIND"] = "tf32"
            call_ms, k_ms = timed(lambda: dms.bartlett_grid(r, z, tilt, out=out, engine=engine), reps=3, warm=1)
            os.environ.pop("BOGP_UMMA_KIND", None)
            name = ("grid_tilt_kernel (K1-SIMT, receiver space, Theta_tau in shared memory)" if engine == "simt" else
                    "umma_grid_kernel<tilt> (K1-UMMA-TILT: tcgen05 " + ("3xTF32" if engine_tag == "umma_tf32" else "3xFP16-split") + ", Theta_tau as B operand, receiver-space epilogue)")
            flop = 8.0 * N * sumM          # complex Theta x complex a: 4 real MACs per (receiver, mode)
            print(json.dumps({"kernel": name, "engine": engine_tag,
                              "case": f"configs[3]: 1000 r x 100 z x {nt} tilt = {C:.0e} candidates, 13 tonals, 64 receivers",
                              "kernel_ms": k_ms, "call_ms": call_ms, "cand_freq_per_s": C * 13 / (k_ms * 1e-3),
                              "alg_tflops": C * flop / (k_ms * 1e-3) / 1e12, "alg_flop_per_cand": flop,
                              "tensor_frac_of_measured_3xtf32_peak": (C * flop / (k_ms * 1e-3) / 1e12) / 278.6 if engine == "umma" else None,
                              "bound": "tensor (3xTF32) with an fp32 receiver-space epilogue" if engine == "umma" else
                                       "fp32 FMA + shared-memory operand loads (CUDA cores)"}), flush=True)


def bench_env():
    import os
    from bogp_b200.envbatch import EnvBatch
    rng = np.random.default_rng(2)
    base = perturbed_pekeris_batch(INV_TONALS_3, REC_Z_21, rng.uniform(200.0, 230.0, 200), rng.uniform(1600.0, 1700.0, 200))
    ref = pekeris_modes(INV_TONALS_3, REC_Z_21)
    K = synthetic_covariance(ref, 71.11, 1.07)
    for Cn in (10_000, 40_000):
        t0 = time.time()
        mode_sets = [base[i % 200] for i in range(Cn)]      # 200 distinct environments tiled (host set-up cost)
        eb = EnvBatch(mode_sets, rng.uniform(0.5, 6.0, Cn), rng.uniform(10.0, 190.0, Cn), K)
        setup_s = time.time() - t0
        out = torch.empty(Cn, dtype=torch.float32, device="cuda")
        for name, env in (("envstream_kernel, 1 stage/warp", {"BOGP_ENV_STAGES": "1"}),
                          ("envstream_kernel, 2 stages/warp", {"BOGP_ENV_STAGES": "2"}),
                          ("envbatch_kernel (SIMT loads)", {"BOGP_ENV_SIMT": "1"})):
            for k in ("BOGP_ENV_STAGES", "BOGP_ENV_SIMT"):
                os.environ.pop(k, None)
            os.environ.update(env)
            call_ms, _ = timed(lambda: eb.bartlett(out=out), lib_timing=False)
            print(json.dumps({"kernel": f"K1-ENV {name}", "case": f"configs[4]: {Cn} environments x 3 tonals x 21 receivers, Mmax={eb.Mmax}",
                              "kernel_ms": call_ms, "cand_freq_per_s": Cn * 3 / (call_ms * 1e-3), "bytes_streamed": eb.bytes_streamed,
                              "achieved_GBs": eb.bytes_streamed / (call_ms * 1e-3) / 1e9, "hbm_peak_GBs": HBM_GBS,
                              "frac": eb.bytes_streamed / (call_ms * 1e-3) / 1e9 / HBM_GBS, "bound": "hbm", "host_setup_s": setup_s}), flush=True)
        for k in ("BOGP_ENV_STAGES", "BOGP_ENV_SIMT"):
            os.environ.pop(k, None)
        eb.close()


def bench_ts():
    """K2-TS: one joint posterior draw over the BAxUS candidate set (baxus.py:45: 2000-5000 candidates) and its arg-max."""
    from bogp_b200.generation import MaxPosteriorSampling
    from oracle import gp as ogp
    rng = np.random.default_rng(0)
    for b, n, d in ((2000, 100, 3), (5000, 200, 12), (5000, 500, 12), (5000, 500, 48)):
        X = rng.uniform(-1, 1, (n, d))
        y = np.sin(3 * X[:, 0]) + 0.3 * np.cos(2 * X.sum(1))
        y = (y - y.mean()) / y.std()
        ls = rng.uniform(0.3, 1.5, d) * (1.0 if d <= 12 else 2.5)
        model = SingleTaskGP(X, y, lengthscale=ls, outputscale=1.3, noise=1e-4)
        ts = MaxPosteriorSampling(model)
        Xc = rng.uniform(-1, 1, (b, d))
        z = rng.standard_normal((1, b))
        call_ms, k_ms = timed(lambda: ts(Xc, base_samples=z), reps=5, warm=2)
        t0 = time.perf_counter()
        idx, f, jit = ogp.thompson_sample(ogp.GPState(X, y, "matern52", ls, 1.3, 0.0, 1e-4), Xc, z)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        flop = 2.0 * (b ** 3 / 3.0 + b * b * n / 2.0)
        print(json.dumps({"kernel": "ts_gemm_nt chain (K2-TS: posterior covariance + blocked Cholesky, FP64 tensor cores)",
                          "case": f"b={b} candidates, n={n} observations, d={d}", "dtype": "f64",
                          "covariance_factorisation_ms": k_ms, "call_ms": call_ms,
                          "alg_gflop": flop / 1e9, "tflops": flop / (k_ms * 1e-3) / 1e12, "fp64_mma_peak_tflops": 37.1,
                          "frac_of_fp64_mma_peak": flop / (k_ms * 1e-3) / 1e12 / 37.1,
                          "oracle_torch_cpu_ms": cpu_ms, "cpu_threads": torch.get_num_threads(),
                          "same_pick_as_oracle": bool(int(ts.last_indices[0]) == int(idx[0])), "jitter": ts.last_jitter,
                          "bound": "FP64 tensor (mma.sync m8n8k4, measured peak tools/micro/fp64_peak.cu) early, the "
                                   "dependent column chain of the 64 x 64 diagonal blocks (~600 cycles per column) late"}),
              flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gp", "fit", "points", "tilt", "env", "ts"]
    torch.cuda.set_device(0)
    for w in which:
        {"gp": bench_gp, "fit": bench_fit, "points": bench_points, "tilt": bench_tilt, "env": bench_env, "ts": bench_ts}[w]()
