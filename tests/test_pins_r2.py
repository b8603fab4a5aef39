"""CPU tier: set-up pieces of the hot path pinned against outputs of the reference's own source (tests/golden/pins_r2.npz,
generator tools/make_golden_pins.py), the `run_kraken` runner the unchanged configs build, and the formatter contract."""
import copy
import functools
from pathlib import Path

import numpy as np
import pytest
import torch

from bogp_b200 import baxus, mfp, shim, sweep
from bogp_b200.modes import INV_TONALS_3, REC_Z_21, modeset_from_tables, pekeris_modes, replica_field

GOLD = np.load(Path(__file__).parent / "golden" / "pins_r2.npz")


@pytest.fixture()
def provider():
    modes = pekeris_modes(INV_TONALS_3, REC_Z_21)
    mfp.set_mode_provider(modes)
    yield modes
    mfp.set_mode_provider(None)


def test_simulate_covariance_matches_reference_source(provider):
    """optimization/utils.py:23-32 executed from source over partial(run_kraken) gave the golden K; the fused front end
    (mfp.simulate_covariance) and the oracle must reproduce it."""
    from oracle import bartlett as ob
    from oracle import field as of
    for tag, params in (("vla", {"src_z": 71.11, "rec_r": 1.07}), ("tilt", {"src_z": 60.0, "rec_r": 3.2, "tilt": 1.5})):
        ref = GOLD[f"simcov_{tag}"]
        assert ref.shape == (3, 21, 21)
        got = mfp.simulate_covariance(functools.partial(mfp.run_kraken), params, INV_TONALS_3)
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-14)
        orc = ob.simulate_covariance(lambda q: of.replica(provider, list(provider.freq).index(q["freq"]), q["src_z"], q["rec_r"], q.get("tilt", 0.0))[:, 0],
                                     params, INV_TONALS_3)
        np.testing.assert_allclose(orc, ref, rtol=0, atol=1e-13)


def test_run_kraken_function_and_partials(provider):
    """`pbuilds(run_kraken)` instantiates to functools.partial(run_kraken): recognised by identity, never called by the
    objective; called with an env dict it returns the replica field of that one source position."""
    p = mfp.run_kraken({"src_z": 71.11, "rec_r": 1.07, "freq": 235.0})
    assert p.shape == (21,) and np.iscomplexobj(p)
    np.testing.assert_allclose(p, replica_field(provider, 1, 71.11, 1.07), rtol=1e-15)
    assert mfp.run_kraken({"src_z": 71.11, "rec_r": np.array([1.0, 2.0]), "freq": 148.0}).shape == (21, 2)
    K = GOLD["simcov_vla"]
    for runner in (mfp.run_kraken, functools.partial(mfp.run_kraken), functools.partial(functools.partial(mfp.run_kraken)),
                   mfp.ModeRunner(provider), functools.partial(mfp.ModeRunner(provider).__class__, provider)()):
        proc = mfp.MatchedFieldProcessor(runner=runner, covariance_matrix=K, freq=INV_TONALS_3, parameters={"tmpdir": "."})
        assert proc.runner.modes_for({}) is provider
    with pytest.raises(TypeError):
        mfp.MatchedFieldProcessor(runner=lambda q: 0, covariance_matrix=K, freq=INV_TONALS_3, parameters={})
    with pytest.raises(TypeError):      # bound arguments cannot be honoured by the fused kernels
        mfp.MatchedFieldProcessor(runner=functools.partial(mfp.run_kraken, {"x": 1}), covariance_matrix=K, freq=INV_TONALS_3, parameters={})
    with pytest.raises(KeyError):
        mfp.run_kraken({"src_z": 1.0, "rec_r": 1.0})
    mfp.set_mode_provider(None)
    with pytest.raises(RuntimeError, match="set_mode_provider"):
        mfp.MatchedFieldProcessor(runner=mfp.run_kraken, covariance_matrix=K, freq=INV_TONALS_3, parameters={})


def test_reference_config_wiring_under_shim(provider):
    """The wiring of projects/swellex96_localization/conf/optimization/run.py:55-99 with hydra-zen's zen_partial builds
    replaced by functools.partial (what `instantiate` hands back): runner -> simulated covariance -> objective."""
    done = shim.install(force=True)
    try:
        from tritonoa.at.models.kraken.runner import run_kraken
        from tritonoa.sp.beamforming import beamformer
        from tritonoa.sp.mfp import MatchedFieldProcessor
        from tritonoa.sp.processing import simulate_covariance
        assert run_kraken is mfp.run_kraken and done["tritonoa"]
        MFPConf = functools.partial(MatchedFieldProcessor)           # pbuilds(MatchedFieldProcessor, ...)
        RunnerConf = functools.partial(run_kraken)                   # pbuilds(run_kraken), instantiated
        K = simulate_covariance(runner=RunnerConf, parameters={"src_z": 71.11, "rec_r": 1.07}, freq=list(INV_TONALS_3))
        np.testing.assert_allclose(K, GOLD["simcov_vla"], rtol=0, atol=1e-14)
        objective = MFPConf(runner=RunnerConf, covariance_matrix=K, freq=list(INV_TONALS_3),
                            parameters={"rec_z": list(REC_Z_21)} | {"tmpdir": "."}, beamformer=functools.partial(beamformer, atype="cbf"))
        assert objective.atype == "cbf" and objective.runner.is_static
    finally:
        shim.uninstall()


def test_snapshot_covariance_matches_reference_source():
    """acoustics.py:116-124 executed from source on a seeded data.npy."""
    got = sweep.snapshot_covariance(GOLD["snap_data"])
    np.testing.assert_allclose(got, GOLD["snap_K"], rtol=0, atol=1e-15)


def test_baxus_embedding_matches_reference_source():
    """baxus.py:156-235 executed from source on fixed torch seeds: same embedding, same split, same observations."""
    n = 0
    while f"emb_{n}_args" in GOLD:
        seed, D, d = (int(v) for v in GOLD[f"emb_{n}_args"])
        torch.manual_seed(seed)
        S = baxus.embedding_matrix(D, d, torch.float64)
        np.testing.assert_array_equal(S.numpy(), GOLD[f"emb_{n}_S"])
        if f"emb_{n}_S2" in GOLD:
            X = torch.rand(5, d, dtype=torch.float64) * 2 - 1
            np.testing.assert_array_equal(X.numpy(), GOLD[f"emb_{n}_X"])
            S2, X2 = baxus.increase_embedding_and_observations(S, X, 3, torch.float64)
            np.testing.assert_array_equal(S2.numpy(), GOLD[f"emb_{n}_S2"])
            np.testing.assert_array_equal(X2.numpy(), GOLD[f"emb_{n}_X2"])
        n += 1
    assert n == 5


def test_mutating_formatter_never_sees_shared_dicts(provider):
    """The reference's formatter (swellex96_inv/data/bo/param_map.py) pops the search keys and edits
    fixed_parameters['layerdata'] in place: every tonal of every candidate must get fresh copies and the caller's
    dicts must come back untouched."""
    seen = []

    def formatter(freq, title, fixed_parameters, search_parameters):
        h_w = search_parameters.pop("h_w")                       # pops, like format_h_w's callers
        water = fixed_parameters["layerdata"][0]
        water["z"].append(h_w)                                   # in-place edit of the nested list
        fixed_parameters["rec_z"] = [z + 1.0 for z in fixed_parameters["rec_z"]]
        seen.append((freq, h_w, list(water["z"])))
        return fixed_parameters | {"freq": freq, "title": title} | search_parameters

    fixed = {"layerdata": [{"z": [0.0, 50.0, 100.0]}], "rec_z": [1.0, 2.0], "src_z": 60.0, "rec_r": 1.0}
    fixed0 = copy.deepcopy(fixed)
    proc = mfp.MatchedFieldProcessor(mfp.ModeRunner(provider), GOLD["simcov_vla"], INV_TONALS_3, fixed, parameter_formatter=formatter)
    cands = [{"h_w": 210.0, "rec_r": 1.5}, {"h_w": 225.0, "rec_r": 2.5}]
    cands0 = copy.deepcopy(cands)
    merged = [proc._merged(c) for c in cands]
    assert cands == cands0 and proc.parameters == fixed0 and fixed == fixed0
    assert merged[0]["layerdata"][0]["z"] == [0.0, 50.0, 100.0, 210.0]
    assert merged[1]["layerdata"][0]["z"] == [0.0, 50.0, 100.0, 225.0]
    assert all(z == [0.0, 50.0, 100.0, h] for _, h, z in seen) and len(seen) == 6


def test_modeset_from_tables_round_trip(provider):
    """Tables as a host normal-mode program returns them (k, z, phi per tonal) -> ModeSet with the same fields."""
    z_fine = np.arange(0.0, 217.0 + 1e-9, 0.25)
    ms = modeset_from_tables(provider.freq, provider.k, [z_fine] * 3, [provider.phi_src(f, z_fine) for f in range(3)],
                             REC_Z_21, z_pivot=provider.z_pivot)
    for f in range(3):
        np.testing.assert_allclose(ms.phi_src_tab[f], provider.phi_src_tab[f][: ms.nz_tab], atol=1e-12)
        np.testing.assert_allclose(ms.phi_rec[f], provider.phi_rec[f], atol=5e-4)      # receivers sit between mesh points (linear interpolation)
        np.testing.assert_allclose(replica_field(ms, f, 71.11, 1.07), replica_field(provider, f, 71.11, 1.07), rtol=2e-2, atol=1e-5)
