"""the oao-compatible optimiser front-end, the strategy mirror and the dotted-path shims.

CPU tier: host logic with a closed-form objective (no CUDA call).  GPU tier: a localisation job as
ref/projects/swellex96_localization/data/run.py:22-49 runs it (Sobol warm-up -> GP/EI -> results.csv), with the
fused matched-field objective, and a grid search that is one surface launch.
"""
import sys

import numpy as np
import pytest

from bogp_b200 import oao, shim
from bogp_b200.strategy import GPEIStrategy, GridSearchStrategy, SobolStrategy


def _space():
    bounds = oao.SearchSpaceBounds(bounds=[
        oao.SearchParameterBounds("rec_r", -1.0, 1.0, relative=True, min_lower_bound=0.010, max_upper_bound=8.0),
        oao.SearchParameterBounds("src_z", -40.0, 40.0, relative=True, min_lower_bound=1.0, max_upper_bound=200.0)])
    return oao.format_search_space(bounds, {"rec_r": 1.07, "src_z": 71.11, "time_step": 250})


def test_format_search_space_clipping_matches_reference_logic():
    """ref/optimization/utils.py:72-98: relative bounds around the scenario, shifted (not shrunk) at the limits."""
    sp = _space()
    assert sp.names == ["rec_r", "src_z"]
    np.testing.assert_allclose(sp.bounds, [[0.07, 31.11], [2.07, 111.11]])
    b = oao.SearchSpaceBounds(bounds=[oao.SearchParameterBounds("rec_r", -1.0, 1.0, True, 0.010, 8.0),
                                      oao.SearchParameterBounds("src_z", -40.0, 40.0, True, 1.0, 200.0),
                                      oao.SearchParameterBounds("tilt", -4.0, 4.0)])
    sp = oao.format_search_space(b, {"rec_r": 0.5, "src_z": 190.0})
    np.testing.assert_allclose(sp.bounds, [[0.010, 120.0, -4.0], [2.010, 200.0, 4.0]])
    u = np.random.default_rng(0).random((5, 3))
    np.testing.assert_allclose(sp.to_unit(sp.from_unit(u)), u, atol=1e-14)


def test_format_search_space_matches_reference_code():
    """Against the outputs of the reference's OWN ``format_search_space`` (optimization/utils.py:72-98), executed from its
    source by tools/make_golden_frontend.py on the localisation / tilt search spaces and scenarios at and inside the
    absolute limits -- this row of the front-end is pinned, not restated."""
    import json
    from pathlib import Path
    g = json.loads((Path(__file__).parent / "golden" / "format_search_space.json").read_text())
    assert len(g["cases"]) >= 12
    for c in g["cases"]:
        bounds = oao.SearchSpaceBounds(bounds=[oao.SearchParameterBounds(**b) for b in c["bounds"]])
        sp = oao.format_search_space(bounds, c["scenario"])
        got = [(p.name, p.type, [float(v) for v in p.bounds]) for p in sp.parameters]
        want = [(p["name"], p["type"], [float(v) for v in p["bounds"]]) for p in c["parameters"]]
        assert got == want, (c["space"], c["scenario"])


def test_strategies_mirror_reference_signatures():
    s = GPEIStrategy(warmup_trials=8, warmup_parallelism=4, num_trials=3, max_parallelism=1, seed=7)()
    assert [st.model for st in s._steps] == [oao.Models.SOBOL, oao.Models.GPEI]
    assert s._steps[0].model_kwargs == {"seed": 7} and s.num_trials == 11
    assert SobolStrategy(num_trials=16, max_parallelism=4, seed=1)()._steps[0].num_trials == 16
    g = GridSearchStrategy(num_trials=5)
    assert g() is g and g.num_trials == 5


def test_sobol_and_grid_loops_and_results_schema(tmp_path):
    """Closed-form objective: trial bookkeeping, batch times, running best, the aggregate.py schema, client JSON."""
    truth = np.array([1.3, 80.0])

    def f(p):
        return float(np.exp(-((p["rec_r"] - truth[0]) / 0.5) ** 2 - ((p["src_z"] - truth[1]) / 20.0) ** 2))

    obj = oao.NoiselessFormattedObjective(f, name="bartlett", properties_kw={"minimize": False})
    assert obj({"rec_r": 1.3, "src_z": 80.0}) == {"bartlett": (1.0, 0.0)}
    opt = oao.QuasiRandom(obj, _space(), SobolStrategy(num_trials=40, max_parallelism=16, seed=3)())
    client = opt.run(name="sobol")
    assert len(client.trials) == 40 and len(opt.batch_execution_times) == 3          # 16 + 16 + 8
    df = oao.get_results(client, opt.batch_execution_times, minimize=False)
    assert list(df.columns) == ["generation_method", "rec_r", "src_z", "bartlett", "best_rec_r", "best_src_z",
                                "best_value", "time"]
    assert df.index.name == "trial_index" and (np.diff(df["best_value"]) >= 0).all()
    i = int(df["bartlett"].values.argmax())
    assert df["best_rec_r"].iloc[-1] == df["rec_r"].iloc[i] and df["time"].nunique() == 3
    # same seed -> same trials (Sobol engine is seeded like Ax's Models.SOBOL(seed))
    again = oao.QuasiRandom(obj, _space(), SobolStrategy(num_trials=40, max_parallelism=16, seed=3)()).run()
    assert [t["parameters"] for t in again.trials] == [t["parameters"] for t in client.trials]
    lo, hi = _space().bounds
    X = np.array([[t["parameters"]["rec_r"], t["parameters"]["src_z"]] for t in client.trials])
    assert (X >= lo).all() and (X <= hi).all()
    client.save_to_json_file(tmp_path / "client.json")
    import json
    snap = json.loads((tmp_path / "client.json").read_text())
    assert len(snap["trials"]) == 40 and snap["metric_name"] == "bartlett"
    # grid: num_trials points per dimension, end points included
    g = oao.GridSearch(obj, _space(), GridSearchStrategy(num_trials=7))
    gc = g.run()
    assert len(gc.trials) == 49 and len(g.batch_execution_times) == 1
    best, val = gc.get_best_parameters()
    assert abs(best["rec_r"] - truth[0]) < 0.35 and abs(best["src_z"] - truth[1]) < 14
    # minimisation flips the running best
    objm = oao.NoiselessFormattedObjective(lambda p: -f(p), name="ml", properties_kw={"minimize": True})
    cm = oao.QuasiRandom(objm, _space(), SobolStrategy(num_trials=8, max_parallelism=8, seed=3)())
    dfm = oao.get_results(cm.run(), cm.batch_execution_times, minimize=True)
    assert (np.diff(dfm["best_value"]) <= 0).all()


def test_gpei_loop_with_injected_backend():
    """The GPEI step: UnitX / StandardizeY transforms, q per batch, seeds -- with a stand-in candidate generator."""
    calls = []

    class Backend:
        def gp_ei_candidates(self, U, y, q, seed):
            calls.append((U.copy(), y.copy(), q, seed))
            return np.full((q, U.shape[1]), 0.5)

    obj = oao.NoiselessFormattedObjective(lambda p: p["rec_r"] + p["src_z"], "m", {"minimize": True})
    opt = oao.BayesianOptimization(obj, _space(), GPEIStrategy(6, 3, 5, 2, seed=11)(), backend=Backend())
    client = opt.run()
    assert len(client.trials) == 11 and [c[2] for c in calls] == [2, 2, 1]
    U, y, _, _ = calls[0]
    assert U.shape == (6, 2) and (U >= 0).all() and (U <= 1).all()
    assert abs(y.mean()) < 1e-12 and abs(y.std(ddof=1) - 1.0) < 1e-12
    # minimise -> the GP sees the negated, standardised objective: its arg-max is the smallest raw value
    raw = np.array([t["value"] for t in client.trials[:6]])
    assert int(np.argmax(y)) == int(np.argmin(raw))
    mid = _space().from_unit(np.full((1, 2), 0.5))[0]
    assert client.trials[-1]["parameters"] == {"rec_r": mid[0], "src_z": mid[1]}
    assert client.trials[-1]["generation_method"] == "GPEI"


def test_shims_resolve_reference_import_paths():
    """The dotted paths of conf/optimization/run.py:15-21,89-99, ei.py:8-18 and strategy.py:5-7."""
    done = shim.install()
    try:
        assert done["tritonoa"] and done["oao"]
        from tritonoa.sp.mfp import MatchedFieldProcessor
        from tritonoa.sp.beamforming import beamformer, covariance
        from tritonoa.at.models.kraken.runner import run_kraken
        from tritonoa.at.models.kraken.kraken import clean_up_kraken_files
        from oao.optimizer import BayesianOptimization, GridSearch, QuasiRandom
        from oao.results import get_results
        from oao.space import SearchParameter, SearchSpace, SearchSpaceBounds
        from oao.strategy import GridStrategy
        from ax.modelbridge.generation_strategy import GenerationStep, GenerationStrategy
        from ax.modelbridge.registry import Models
        from botorch.models import SingleTaskGP
        from botorch.acquisition import ExpectedImprovement, LogExpectedImprovement, UpperConfidenceBound
        from botorch.optim import optimize_acqf
        from gpytorch.constraints import Interval
        from gpytorch.likelihoods import GaussianLikelihood
        import bogp_b200.mfp as m, bogp_b200.gp as g
        assert MatchedFieldProcessor is m.MatchedFieldProcessor and run_kraken is m.run_kraken
        assert SingleTaskGP is g.SingleTaskGP and BayesianOptimization is oao.BayesianOptimization
        lik = GaussianLikelihood(noise_constraint=Interval(1e-8, 1e-3))
        gp = SingleTaskGP(np.random.default_rng(0).random((5, 2)), np.arange(5.0), likelihood=lik)
        assert gp.noise_bounds == (1e-8, 1e-3)
        clean_up_kraken_files("anywhere")
    finally:
        shim.uninstall()
    assert "tritonoa" not in sys.modules


# ------------------------------------------------------------------------------------------------ GPU tier
@pytest.mark.gpu
def test_localisation_job_end_to_end(cfg1, tmp_path):
    """data/run.py:22-49 with the fused objective: Sobol warm-up, GP + EI trials, results.csv, best near the truth."""
    from functools import partial
    from bogp_b200.mfp import MatchedFieldProcessor, ModeRunner, beamformer, simulate_covariance
    modes, (r0, z0) = cfg1["modes"], cfg1["truth"]
    runner = ModeRunner(modes)
    K = simulate_covariance(runner, {"rec_r": r0, "src_z": z0}, modes.freq)
    np.testing.assert_allclose(K, cfg1["K"], atol=1e-14)
    mfp = MatchedFieldProcessor(runner=runner, covariance_matrix=K, freq=modes.freq, parameters={},
                                beamformer=partial(beamformer, atype="cbf"))
    obj = oao.NoiselessFormattedObjective(mfp, name="bartlett", properties_kw={"minimize": False})
    space = oao.format_search_space(oao.SearchSpaceBounds(bounds=[
        oao.SearchParameterBounds("rec_r", -1.0, 1.0, True, 0.010, 8.0),
        oao.SearchParameterBounds("src_z", -40.0, 40.0, True, 1.0, 200.0)]), {"rec_r": r0, "src_z": z0})
    opt = oao.BayesianOptimization(obj, space, GPEIStrategy(warmup_trials=32, warmup_parallelism=16, num_trials=10,
                                                            max_parallelism=1, seed=2009)(), fit_steps=50)
    client = opt.run(name="gpei")
    df = oao.get_results(client, opt.batch_execution_times, minimize=False)
    df.to_csv(tmp_path / "results.csv")
    assert len(df) == 42 and (df["generation_method"].iloc[32:] == "GPEI").all()
    # the running best never regresses, and the GP trials exploit: they score higher on average than the warm-up
    # (the Bartlett surface is multi-modal with a ~50 m wide main lobe; 42 trials need not find it)
    assert df["best_value"].iloc[-1] >= df["best_value"].iloc[31]
    assert df["bartlett"].iloc[32:].mean() > df["bartlett"].iloc[:32].mean()
    assert ((df["rec_r"] >= space.bounds[0][0]) & (df["rec_r"] <= space.bounds[1][0])).all()
    # every recorded value is the fused objective at the recorded parameters
    i = 37
    p = {"rec_r": df["rec_r"].iloc[i], "src_z": df["src_z"].iloc[i]}
    assert mfp(p) == pytest.approx(df["bartlett"].iloc[i], rel=1e-6)
    # q = 4 batches (qgpei, conf/optimization/run.py:202-208)
    optq = oao.BayesianOptimization(obj, space, GPEIStrategy(16, 16, 8, 4, seed=2009)(), fit_steps=30, mc_samples=128,
                                    num_restarts=8, raw_samples=256)
    cq = optq.run()
    assert len(cq.trials) == 24 and len(optq.batch_execution_times) == 3


@pytest.mark.gpu
def test_grid_search_is_one_surface_launch(cfg1):
    from bogp_b200 import _lib
    from bogp_b200.mfp import MatchedFieldProcessor, ModeRunner, beamformer
    from oracle import bartlett as ob
    modes = cfg1["modes"]
    mfp = MatchedFieldProcessor(runner=ModeRunner(modes), covariance_matrix=cfg1["K"], freq=modes.freq, parameters={},
                                beamformer=beamformer)
    obj = oao.NoiselessFormattedObjective(mfp, name="bartlett", properties_kw={"minimize": False})
    space = oao.SearchSpace([oao.SearchParameter("rec_r", "range", [0.5, 2.0]),
                             oao.SearchParameter("src_z", "range", [40.0, 100.0])])
    mfp.surface([1.0], [50.0])                       # build the device handle outside the counted region
    n0 = _lib.kernel_launches()
    g = oao.GridSearch(obj, space, GridSearchStrategy(num_trials=24))
    client = g.run()
    assert _lib.kernel_launches() - n0 <= 3          # operand tables + surface kernel (+ nothing per trial)
    assert len(client.trials) == 576
    ref = ob.ambiguity_surface(modes, cfg1["K"], np.linspace(0.5, 2.0, 24), np.linspace(40.0, 100.0, 24))
    got = np.array([t["value"] for t in client.trials]).reshape(24, 24)
    np.testing.assert_allclose(got, ref, atol=2e-6)
    best, _ = client.get_best_parameters()
    jz, jr = np.unravel_index(ref.argmax(), ref.shape)
    assert best == {"src_z": np.linspace(40.0, 100.0, 24)[jz], "rec_r": np.linspace(0.5, 2.0, 24)[jr]}


# ------------------------------------------------------------------------------------------------ sweep (8f row 3)
def test_covariance_front_end_and_disk_formats(tmp_path):
    """acoustics.py:116-124 per-snapshot covariances, the <f>Hz/covariance.npy layout and mfp.py:111-121 loader."""
    from bogp_b200 import sweep
    rng = np.random.default_rng(0)
    p = rng.standard_normal((12, 5)) + 1j * rng.standard_normal((12, 5))
    K = sweep.snapshot_covariance(p)
    assert K.shape == (12, 5, 5)
    d = p[3] / np.linalg.norm(p[3])
    np.testing.assert_allclose(K[3], np.outer(d, d.conj()), atol=1e-15)
    np.testing.assert_allclose(np.trace(K, axis1=1, axis2=2).real, 1.0, atol=1e-14)
    K4 = sweep.snapshot_covariance(p, averaging=4)
    assert K4.shape == (3, 5, 5)
    np.testing.assert_allclose(K4[1], K[4:8].mean(0) / np.trace(K[4:8].mean(0)).real, atol=1e-15)
    assert np.linalg.matrix_rank(K4[1]) == 4
    freqs = [148.0, 235.0]
    for f in freqs:
        sweep.save_covariance(K * (1 + f / 1000), tmp_path, f)
    assert (tmp_path / "148.0Hz" / "covariance.npy").exists()
    L = sweep.load_covariance(tmp_path, freqs, num_segments=12, num_elements=5)
    assert L.shape == (2, 12, 5, 5) and L.dtype == complex
    np.testing.assert_allclose(L[1], K * 1.235)
    with pytest.raises(ValueError):
        sweep.load_covariance(tmp_path, freqs, num_segments=11)
    # time steps are split like depth rows: contiguous, balanced, complete
    got = [list(sweep.time_steps_of(r, 3, 350)) for r in range(3)]
    assert sum(got, []) == list(range(350)) and {len(g) for g in got} == {116, 117}


@pytest.mark.gpu
def test_time_step_sweep_matches_oracle_and_reference_files(cfg1, tmp_path):
    """mfp.py:131-216 for a short towed-source event: per-step covariances from synthetic snapshots, one fused
    surface per step, reference file names, peaks that follow the source, two 'ranks' covering all steps."""
    import pickle
    from bogp_b200 import sweep
    from bogp_b200.mfp import MatchedFieldProcessor, ModeRunner, beamformer
    from oracle import bartlett as ob
    from oracle import field as of
    modes = cfg1["modes"]
    rvec, zvec = np.linspace(0.5, 3.0, 90), np.linspace(20.0, 120.0, 41)
    track = [(1.0 + 0.25 * t, 60.0 + 5.0 * t) for t in range(6)]                 # (range km, depth m) per time step
    data = {f: np.stack([of.replica(modes, i, z, r)[:, 0] * np.exp(1j * t) for t, (r, z) in enumerate(track)])
            for i, f in enumerate(modes.freq)}
    for f in modes.freq:
        sweep.save_covariance(sweep.snapshot_covariance(data[f]), tmp_path / "acoustic", f)
    K = sweep.load_covariance(tmp_path / "acoustic", modes.freq)
    assert K.shape == (3, 6, 21, 21)
    mfp = MatchedFieldProcessor(runner=ModeRunner(modes), covariance_matrix=K[:, 0], freq=modes.freq, parameters={},
                                beamformer=beamformer)
    parts = [sweep.ambiguity_sweep(mfp, K, rvec, zvec, savepath=tmp_path / "mfp", rank=r, world=2) for r in range(2)]
    assert parts[0]["timesteps"] == [0, 1, 2] and parts[1]["timesteps"] == [3, 4, 5]
    with open(tmp_path / "mfp" / "grid_parameters.pkl", "rb") as fh:
        grid = pickle.load(fh)
    assert len(grid["rec_r"]) == 90 and len(grid["src_z"]) == 41
    for part in parts:
        for k, t in enumerate(part["timesteps"]):
            ref = ob.ambiguity_surface(modes, K[:, t], rvec, zvec)
            disk = np.load(tmp_path / "mfp" / f"surface_{t:03d}.npy")
            assert disk.dtype == np.float64 and disk.shape == (41, 90)
            np.testing.assert_allclose(disk, ref, atol=2e-6)
            np.testing.assert_array_equal(disk.astype(np.float32), part["surfaces"][k])
            val, (jz, jr) = part["peak"][k]
            assert (jz, jr) == np.unravel_index(ref.argmax(), ref.shape)
            assert abs(rvec[jr] - track[t][0]) < 0.03 and abs(zvec[jz] - track[t][1]) < 2.6
