"""CPU tier: host-side logic of both boundaries and the multi-rank reduction (gloo, world_size 2)."""
import os
import pickle
from functools import partial

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bogp_b200 import mfp, sharding
from bogp_b200.modes import INV_TONALS_3, REC_Z_21, ModeSet, pekeris_modes


def test_modeset_contract_and_interpolation():
    m = pekeris_modes(INV_TONALS_3, REC_Z_21)
    assert m.n_freq == 3 and m.n_rec == 21 and all(M > 0 for M in m.n_modes)
    z = np.array([0.0, 0.1, 33.3, 216.9, 500.0, -3.0])
    got = m.phi_src(1, z)
    kz = (np.arange(1, m.n_modes[1] + 1) - 0.5) * np.pi / 217.0
    exact = np.sqrt(2 / 217.0) * np.sin(np.outer(np.clip(z, 0, 217.0), kz))
    np.testing.assert_allclose(got, exact, atol=3e-4)       # linear interpolation of a 0.25 m table (h^2 f''/8)
    with pytest.raises(ValueError):
        ModeSet([1.0], [np.ones(3, complex)], [np.ones((4, 2))], [np.ones((5, 3))], 0.0, 1.0, np.zeros(4))


def test_mfp_constructor_mirrors_reference_and_rejects_unfusable():
    m = pekeris_modes(INV_TONALS_3, REC_Z_21)
    K = np.zeros((3, 21, 21), complex)
    proc = mfp.MatchedFieldProcessor(runner=mfp.ModeRunner(m), covariance_matrix=K, freq=INV_TONALS_3,
                                     parameters={"tilt": 0.0}, beamformer=partial(mfp.beamformer, atype="cbf_ml"),
                                     multifreq_method="product")
    assert proc.atype == "cbf_ml" and proc.multifreq_method == "product"
    clone = pickle.loads(pickle.dumps(proc))                   # process pools / hydra-zen partials pickle it
    assert clone.freq == proc.freq and clone._dms is None
    with pytest.raises(TypeError):
        mfp.MatchedFieldProcessor(lambda p: None, K, INV_TONALS_3, {})
    with pytest.raises(TypeError):
        mfp.MatchedFieldProcessor(mfp.ModeRunner(m), K, INV_TONALS_3, {}, beamformer=lambda K, p: 0.0)
    with pytest.raises(ValueError):
        mfp.MatchedFieldProcessor(mfp.ModeRunner(m), K, INV_TONALS_3, {}, multifreq_method="median")
    with pytest.raises(ValueError):
        mfp.MatchedFieldProcessor(mfp.ModeRunner(m), K[:2], INV_TONALS_3, {})
    # per-tonal merge follows the reference: fixed | {freq, title} | search, or the formatter's signature
    seen = []

    def formatter(freq, title, fixed_parameters, search_parameters):
        seen.append((freq, title))
        return fixed_parameters | search_parameters

    proc2 = mfp.MatchedFieldProcessor(mfp.ModeRunner(m), K, INV_TONALS_3, {"src_z": 60.0}, parameter_formatter=formatter)
    merged = proc2._merged({"rec_r": 1.0})
    assert merged["src_z"] == 60.0 and seen[0] == (148.0, "148.0Hz") and len(seen) == 3


def test_slab_partition_is_balanced_and_complete():
    for n in (0, 1, 7, 1000, 1001):
        for w in (1, 2, 3, 8):
            spans = [sharding.slab(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_combine_extrema_tie_and_nan_rules():
    v = torch.tensor([0.5, float("nan"), 0.75, 0.75])
    i = torch.tensor([3, 10, 40, 20])
    assert sharding.combine_extrema(v, i) == (0.75, 20)           # first occurrence wins ties
    assert sharding.combine_extrema(v, i, maximize=False) == (0.5, 3)
    assert sharding.combine_extrema(torch.tensor([float("nan")]), torch.tensor([-1]))[1] == -1


def _worker(rank, world, port, surf, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = torch.from_numpy(surf)
        local, (val, pos), full = sharding.sharded_surface(lambda lo, hi: S[lo:hi].clone(), S.shape[0], S.shape[1],
                                                           gather_surface=True)
        lo, hi = sharding.slab(S.shape[0], rank, world)
        assert local.shape[0] == hi - lo
        q.put((rank, val, pos, bool(torch.equal(torch.nan_to_num(full, nan=-7.0), torch.nan_to_num(S, nan=-7.0)))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_argmax_gloo(world):
    """N > 1 path on CPU: row slabs per rank, all-gather of (max, index), identical answer on every rank."""
    rng = np.random.default_rng(5)
    surf = rng.random((11, 17)).astype(np.float32)
    surf[3, 4] = surf[9, 2] = 2.0                                # tie across ranks: the lower flat index wins
    surf[0, 0] = np.nan
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, surf, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, val, pos, same in res:
        assert val == 2.0 and pos == (3, 4) and same


def _sweep_worker(rank, world, port, n_steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bogp_b200 import sweep
        steps = list(sweep.time_steps_of(rank, world, n_steps))
        # stand-in for the per-rank ambiguity_sweep result: the peak of step t is (t / 10, (t, 2 t))
        local = {"timesteps": steps, "peak": [(t / 10.0, (t, 2 * t)) for t in steps]}
        q.put((rank, steps, sweep.gather_peaks(local)))
    finally:
        dist.destroy_process_group()


def test_time_step_sweep_partition_and_gather_gloo():
    """The sweep's multi-GPU axis on CPU (world 2, gloo): contiguous blocks of time steps per rank, one all-gather of
    the per-step peaks, the same merged dictionary on every rank (ref: the pool over time steps, mfp.py:166-177)."""
    world, n_steps = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sweep_worker, args=(r, world, port, n_steps, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    res.sort()
    assert res[0][1] == [0, 1, 2, 3] and res[1][1] == [4, 5, 6]
    want = {t: (t / 10.0, (t, 2 * t)) for t in range(n_steps)}
    for _, _, merged in res:
        assert merged == want and list(merged) == sorted(merged)


class _ToyAcq:
    """CPU stand-in for an acquisition function: a smooth multi-modal surface on [0, 1]^d with a known global maximum."""
    nonneg = True

    def __init__(self):
        self.calls = []

    def __call__(self, X):
        self.calls.append(int(X.shape[0]))
        x = X.reshape(X.shape[0], -1)
        peak = torch.tensor([0.7, 0.2, 0.55], dtype=x.dtype)[: x.shape[1]]
        return torch.exp(-8.0 * ((x - peak) ** 2).sum(-1)) + 0.3 * torch.exp(-30.0 * ((x - 0.1) ** 2).sum(-1))


def _optim_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bogp_b200.optim import optimize_acqf
        torch.manual_seed(100 + rank)                      # different local streams: the seed / restarts come from rank 0
        acq = _ToyAcq()
        bounds = torch.stack([torch.zeros(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64)])
        cand, val = optimize_acqf(acq, bounds, q=1, num_restarts=8, raw_samples=256, shard=True)
        allX, allv = optimize_acqf(acq, bounds, q=1, num_restarts=8, raw_samples=256, shard=True, return_best_only=False)
        q.put((rank, cand.numpy(), float(val), acq.calls[0], tuple(allX.shape), tuple(allv.shape)))
    finally:
        dist.destroy_process_group()


def test_optimize_acqf_sharded_raw_samples_and_restarts_gloo():
    """SURVEY 8e for K2 on CPU (world 2, gloo): each rank evaluates half of the raw samples and optimises half of the
    restarts, values and candidates are all-gathered, every rank returns the same best candidate -- the global maximum."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_optim_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=180) for _ in range(world))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, c0, v0, n0, sx0, sv0), (r1, c1, v1, n1, sx1, sv1) = res
    assert n0 == n1 == 128                                  # raw_samples / P evaluations per rank
    np.testing.assert_array_equal(c0, c1)
    assert v0 == v1 and sx0 == sx1 == (8, 1, 3) and sv0 == sv1 == (8,)
    np.testing.assert_allclose(c0.reshape(-1), [0.7, 0.2, 0.55], atol=1e-4)
    assert abs(v0 - 1.0) < 1e-3


def _cand_worker(rank, world, port, vals, vol, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bogp_b200.sharding import sharded_candidates, sharded_tilt_volume
        v = torch.as_tensor(vals)
        local, (val, idx), full = sharded_candidates(lambda lo, hi: v[lo:hi], len(vals), maximize=False, gather=True)
        V = torch.as_tensor(vol)
        planes, (vmax, pos) = sharded_tilt_volume(lambda lo, hi: V[lo:hi], *vol.shape)
        q.put((rank, len(local), val, idx, bool(torch.equal(full, v)), tuple(planes.shape), vmax, pos))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_candidates_and_tilt_volume_gloo(world):
    """configs[4] / configs[3] sharding on CPU: candidate slabs (arg-min: the inversion minimises cbf_ml) with the gathered
    values in global order, tilt planes with the (tilt, depth, range) index of the global maximum; every rank agrees."""
    rng = np.random.default_rng(11)
    vals = rng.random(23).astype(np.float32)
    vals[7] = vals[19] = -1.0                                     # tie across ranks: the lower index wins
    vol = rng.random((5, 4, 6)).astype(np.float32)
    vol[3, 2, 5] = 9.0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_cand_worker, args=(r, world, port, vals, vol, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    from bogp_b200.sharding import slab
    for rank, n_local, val, idx, same, pshape, vmax, pos in res:
        lo, hi = slab(23, rank, world)
        assert n_local == hi - lo and val == -1.0 and idx == 7 and same
        tl, th = slab(5, rank, world)
        assert pshape == (th - tl, 4, 6) and vmax == 9.0 and pos == (3, 2, 5)


def test_shim_exports_the_fit_protocol():
    """Every name the reference's hand-rolled loops import for the GP fit (ei.py:7-14, baxus.py:10-21, demo_1d.py) resolves
    under shim.install(), and the explicit kernel of baxus.py:306-319 selects the Matern-5/2 ARD model with Interval
    constraints and no priors.  No device call."""
    import torch
    from bogp_b200 import shim
    shim.install(force=True)
    try:
        import botorch
        import gpytorch
        from botorch.exceptions import ModelFittingError
        from botorch.fit import fit_gpytorch_mll
        from botorch.models import SingleTaskGP
        from gpytorch.constraints import Interval
        from gpytorch.kernels import MaternKernel, RBFKernel, ScaleKernel
        from gpytorch.likelihoods import GaussianLikelihood
        from gpytorch.mlls import ExactMarginalLogLikelihood
        assert issubclass(ModelFittingError, Exception) and callable(fit_gpytorch_mll)
        X, Y = torch.rand(12, 4, dtype=torch.double), torch.rand(12, dtype=torch.double)
        with botorch.settings.validate_input_scaling(False), gpytorch.settings.max_cholesky_size(2000):
            m = SingleTaskGP(X, Y, covar_module=ScaleKernel(MaternKernel(nu=2.5, ard_num_dims=4, lengthscale_constraint=Interval(0.005, 10)),
                                                            outputscale_constraint=Interval(0.05, 10)),
                             likelihood=GaussianLikelihood(noise_constraint=Interval(1e-8, 1e-3)))
        assert m.kind == "matern52" and m.noise_bounds == (1e-8, 1e-3)
        assert m.fit_defaults.lengthscale_bounds == (0.005, 10.0) and m.fit_defaults.outputscale_bounds == (0.05, 10.0)
        assert m.fit_defaults.lengthscale_prior is None and m.fit_defaults.outputscale_prior is None
        assert SingleTaskGP(X, Y, covar_module=ScaleKernel(RBFKernel(ard_num_dims=4))).kind == "rbf"
        mll = ExactMarginalLogLikelihood(m.likelihood, m)
        (raw,) = m.parameters()
        assert isinstance(raw, torch.nn.Parameter) and raw.shape == (7,) and float(raw.abs().sum()) == 0.0
        opt = torch.optim.AdamW([{"params": m.parameters()}], lr=0.1)
        assert opt.param_groups[0]["params"][0] is raw
        m.train(); m.likelihood.train()
        assert m(X) is None and mll.train() is mll
        import pytest
        with pytest.raises(NotImplementedError):
            MaternKernel(nu=1.5)
        # demo_1d.py:6-8,92-100,162: `from botorch import fit_gpytorch_mll`, InputDataWarning, state_dict warm start
        from botorch import fit_gpytorch_mll as top_level
        from botorch.exceptions import InputDataWarning
        assert top_level is fit_gpytorch_mll and issubclass(InputDataWarning, Warning)
        with torch.no_grad():
            raw.copy_(torch.linspace(-0.3, 0.3, 7, dtype=torch.double))
        sd = m.state_dict()
        m2 = SingleTaskGP(torch.rand(15, 4, dtype=torch.double), torch.rand(15, 1, dtype=torch.double),
                          covar_module=ScaleKernel(MaternKernel(nu=2.5, ard_num_dims=4, lengthscale_constraint=Interval(0.005, 10)),
                                                   outputscale_constraint=Interval(0.05, 10)))
        m2.load_state_dict(sd)
        assert torch.equal(m2.parameters()[0].detach(), raw.detach())
        sig = lambda v: 1.0 / (1.0 + np.exp(-v))                                    # noqa: E731
        np.testing.assert_allclose(m2.lengthscale.numpy(), 0.005 + (10 - 0.005) * sig(np.linspace(-0.3, 0.3, 7)[:4]), rtol=1e-12)
        assert m2.outputscale == pytest.approx(0.05 + (10 - 0.05) * sig(np.linspace(-0.3, 0.3, 7)[4]), rel=1e-12)
        assert m2.mean_const == pytest.approx(0.3)
        with pytest.raises(KeyError):
            SingleTaskGP(torch.rand(5, 2, dtype=torch.double), torch.rand(5, dtype=torch.double)).load_state_dict(sd)
        # baxus.py:116: the trust region is scaled by model.covar_module.base_kernel.lengthscale
        w = m2.covar_module.base_kernel.lengthscale.detach().view(-1)
        assert w.shape == (4,) and torch.equal(w, m2.lengthscale) and m2.covar_module.base_kernel.lengthscale.shape == (1, 4)
        assert float(m2.covar_module.outputscale) == pytest.approx(m2.outputscale)
    finally:
        shim.uninstall()


def test_grid_vector_cache_follows_values_not_objects():
    """DeviceModeSet._vec (the per-call host work of the grid entry points): the converted vector and its pointer are
    reused while the VALUES are the same -- a new array object with equal values hits, the same object modified in
    place misses -- and the km -> m scaling is applied once.  No device call."""
    import ctypes as C
    from bogp_b200 import mfp
    d = mfp.DeviceModeSet.__new__(mfp.DeviceModeSet)
    d._h, d._vec_cache, d._out_cache = C.c_void_p(), {}, None
    r = np.linspace(0.1, 10.0, 64)
    a1, p1 = d._vec("r", r, 1000.0)
    np.testing.assert_array_equal(a1, 1000.0 * r)
    assert a1.dtype == np.float64 and a1.flags.c_contiguous
    a2, p2 = d._vec("r", r.copy(), 1000.0)                       # other object, same values: hit
    assert a2 is a1 and p2 is p1
    a3, _ = d._vec("r", list(r), 1000.0)                         # a list with the same values: hit
    assert a3 is a1
    r[5] += 0.25                                                  # same object, modified in place: miss
    a4, p4 = d._vec("r", r, 1000.0)
    assert a4 is not a1 and a4[5] == 1000.0 * r[5]
    z = np.array([5.0, 6.0], dtype=np.float32)                    # other dtype, unscaled key
    az, _ = d._vec("z", z)
    assert az.dtype == np.float64 and az.tolist() == [5.0, 6.0]
    assert d._vec("r", r, 1000.0)[0] is a4                        # keys do not evict each other
    s0, _ = d._vec("z", 7.5)                                      # scalars become length-1 vectors
    assert s0.shape == (1,) and s0[0] == 7.5
    out = np.empty((2, 3), dtype=np.float32)
    assert d._out_ptr(out) is d._out_ptr(out)
    assert d._out_ptr(np.empty((2, 3), dtype=np.float32)) is not None
    d._h = C.c_void_p()                                           # nothing to destroy


def test_every_shimmed_import_of_the_reference_resolves():
    """tests/golden/reference_imports.json (tools/make_golden_imports.py: read with `ast` from the reference's hot-path
    scripts -- strategy / utils, the hydra-zen configs, the run scripts, the hand-rolled BO loops, demo_1d) lists every
    name they import from tritonoa, oao, ax, botorch and gpytorch: under shim.install() each one resolves.  The one file
    left out is optimization/gpmodels.py, which no other reference file imports: it SUBCLASSES GPyTorch's ExactGP
    (RBF-ARD surrogate), mirrored here by SingleTaskGP(covar_module="rbf") rather than by stand-in base classes."""
    import importlib
    import json
    from pathlib import Path
    from bogp_b200 import shim
    listed = json.loads((Path(__file__).parent / "golden" / "reference_imports.json").read_text())
    assert sum(len(v) for v in listed.values()) >= 90
    shim.install(force=True)
    try:
        missing = []
        for rel, items in listed.items():
            if rel == "optimization/gpmodels.py":
                continue
            for mod, name, line in items:
                try:
                    m = importlib.import_module(mod)
                except ImportError as exc:
                    missing.append((rel, line, mod, name, repr(exc)))
                    continue
                if name is not None and not hasattr(m, name):
                    missing.append((rel, line, mod, name))
        assert not missing, missing
    finally:
        shim.uninstall()

