"""BAxUS host logic (CPU tier, against golden values produced by the reference's own
``BaxusState`` / ``update_state`` source -- tools/make_golden_baxus.py) and Thompson sampling on the device (GPU tier,
``bogp_gp_thompson`` through the C ABI against the oracle's joint posterior draw)."""
import json
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import gp as ogp

GOLD = Path(__file__).parent / "golden"


# ------------------------------------------------------------------------------------------------ CPU tier
def test_baxus_state_matches_reference_golden():
    from bogp_b200.baxus import BaxusState
    g = json.loads((GOLD / "baxus_state.json").read_text())
    assert len(g["states"]) >= 30
    for rec in g["states"]:
        st = BaxusState(dim=rec["dim"], eval_budget=rec["eval_budget"])
        assert (st.d_init, st.n_splits, st.target_dim) == (rec["d_init"], rec["n_splits"], rec["d_init"]), rec
        for row in rec["by_target_dim"]:
            st.target_dim = row["target_dim"]
            assert st.split_budget == row["split_budget"], (rec["dim"], row)
            assert st.failure_tolerance == row["failure_tolerance"], (rec["dim"], row)


def test_update_state_trajectory_matches_reference_golden():
    from bogp_b200.baxus import BaxusState, update_state
    g = json.loads((GOLD / "baxus_state.json").read_text())
    st = BaxusState(dim=206, eval_budget=400)
    restarts = 0
    for step in g["trajectory"]:
        st = update_state(st, torch.tensor([step["y"]], dtype=torch.float64))
        assert st.length == step["length"] and st.best_value == step["best_value"]
        assert (st.success_counter, st.failure_counter) == (step["success_counter"], step["failure_counter"])
        assert st.restart_triggered == step["restart_triggered"]
        if st.restart_triggered:
            restarts += 1
            st.restart_triggered = False
            st.target_dim = min(st.dim, st.target_dim * 3)
            st.length = st.length_init
            st.failure_counter = st.success_counter = 0
    assert restarts >= 1          # the scripted sequence collapses the trust region at least once


def test_loop_args():
    from bogp_b200.baxus import BAxUSLoopArgs
    a = BAxUSLoopArgs(true_dim=6)
    assert (a.dim, a.n_candidates, a.budget, a.n_init, a.num_restarts, a.raw_samples) == (206, 5000, 500, 100, 40, 1024)
    assert BAxUSLoopArgs(true_dim=6, dummy_dim=0).n_candidates == 2000
    assert BAxUSLoopArgs(true_dim=6, dummy_dim=10).n_candidates == 3200


@pytest.mark.parametrize("input_dim,target_dim", [(206, 3), (20, 7), (6, 6), (5, 9)])
def test_embedding_matrix(input_dim, target_dim):
    from bogp_b200.baxus import embedding_matrix
    torch.manual_seed(0)
    S = embedding_matrix(input_dim, target_dim)
    if target_dim >= input_dim:
        assert torch.equal(S, torch.eye(input_dim, dtype=torch.float64))
        return
    assert S.shape == (target_dim, input_dim)
    assert torch.all((S == 0) | (S.abs() == 1))
    assert torch.all((S != 0).sum(0) == 1)                     # each input dimension in exactly one bin
    sizes = (S != 0).sum(1)
    assert sizes.max() - sizes.min() <= 1 and sizes.min() >= 1  # near-equal bins
    assert (S == -1).any() and (S == 1).any()


def test_increase_embedding_keeps_embedded_points():
    from bogp_b200.baxus import EmbeddingExhausted, embedding_matrix, increase_embedding_and_observations
    torch.manual_seed(1)
    S = embedding_matrix(26, 3)
    X = 2 * torch.rand(11, 3, dtype=torch.float64) - 1
    S2, X2 = increase_embedding_and_observations(S, X, 3)
    assert S2.shape == (9, 26) and X2.shape == (11, 9)
    assert torch.all((S2 != 0).sum(0) == 1)
    assert torch.allclose(X2 @ S2, X @ S, atol=0, rtol=0)       # the evaluated points did not move
    assert torch.equal(X2[:, :3], X)
    # rows with 2 members split into 2; with one member the reference's loop ends the optimisation (TypeError)
    S3 = torch.zeros(2, 4, dtype=torch.float64)
    S3[0, [0, 1]] = 1.0
    S3[1, [2, 3]] = -1.0
    S4, X4 = increase_embedding_and_observations(S3, X[:, :2], 3)
    assert S4.shape == (4, 4) and torch.all((S4 != 0).sum(1) == 1)
    with pytest.raises(TypeError):
        increase_embedding_and_observations(S4, X4, 3)
    with pytest.raises(EmbeddingExhausted):
        increase_embedding_and_observations(S4, X4, 3)
    with pytest.raises(AssertionError):
        increase_embedding_and_observations(S, X[:, :2], 3)


def test_trust_region_and_thompson_candidates():
    """Trust region scaled by the length scales (volume-preserving weights, clipped to [-1, 1]) and the candidate set
    of the Thompson step: inside the region, equal to the incumbent except in a random non-empty coordinate subset."""
    from types import SimpleNamespace

    from bogp_b200.baxus import BaxusState, _trust_region, thompson_candidates
    torch.manual_seed(0)
    d = 40
    X = 2 * torch.rand(30, d, dtype=torch.float64) - 1
    Y = torch.randn(30, dtype=torch.float64)
    ls = torch.exp(torch.randn(d, dtype=torch.float64))
    state = BaxusState(dim=206, eval_budget=400)
    center, lb, ub = _trust_region(state, SimpleNamespace(lengthscale=ls), X, Y)
    assert torch.equal(center, X[int(torch.argmax(Y))])
    w = ls / ls.mean()
    w = w / torch.prod(w.pow(1.0 / d))
    assert abs(float(torch.prod(w.pow(1.0 / d))) - 1.0) < 1e-12
    assert torch.allclose(lb, torch.clamp(center - w * state.length, -1, 1)) and torch.allclose(ub, torch.clamp(center + w * state.length, -1, 1))
    assert torch.all(lb >= -1) and torch.all(ub <= 1) and torch.all(lb <= center) and torch.all(center <= ub)
    Xc = thompson_candidates(center, lb, ub, 3000)
    assert Xc.shape == (3000, d) and torch.all(Xc >= lb) and torch.all(Xc <= ub)
    moved = (Xc != center).sum(1)
    assert int(moved.min()) >= 1                                   # every candidate differs from the incumbent somewhere
    assert 15.0 < float(moved.double().mean()) < 25.0              # 20 coordinates on average when dim > 20
    assert int(moved.max()) < d                                    # and never all of them at dim = 40
    Xc5 = thompson_candidates(center[:5], lb[:5], ub[:5], 500)
    assert torch.all((Xc5 != center[:5]).sum(1) == 5)              # dim <= 20: every coordinate is perturbed


def test_oracle_thompson_sample_properties():
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (15, 2))
    y = np.sin(3 * X[:, 0]) + X[:, 1]
    st = ogp.GPState(X, y, "matern52", [0.5, 0.7], 1.0, 0.0, 1e-4)
    Xc = rng.uniform(-1, 1, (40, 2))
    idx, f, jit = ogp.thompson_sample(st, Xc, np.zeros((1, 40)))
    mu, _ = ogp.posterior(st, torch.as_tensor(Xc))
    assert int(idx[0]) == int(torch.argmax(mu)) and torch.allclose(f[0], mu)      # z = 0: the posterior mean
    idx3, f3, _ = ogp.thompson_sample(st, Xc, rng.standard_normal((3, 40)))
    assert len(set(idx3.tolist())) == 3                                             # without replacement
    # the draws have the posterior's covariance: f - mu = z L^T with L L^T = Sigma (checked through unit vectors)
    _, fe, _ = ogp.thompson_sample(st, Xc, np.eye(40))
    Lt = fe - mu                                   # row j = column j of L
    _, cov = ogp.posterior(st, torch.as_tensor(Xc))
    assert torch.allclose(Lt.T @ Lt, cov, atol=1e-9)


# ------------------------------------------------------------------------------------------------ GPU tier
def _problem(n, d, seed=0, noise=1e-4):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (n, d))
    y = np.sin(3 * X[:, 0]) + 0.3 * np.cos(2 * X.sum(1))
    y = (y - y.mean()) / y.std()
    ls = rng.uniform(0.3, 1.5, d)
    return X, y, ls, noise


@pytest.mark.gpu
@pytest.mark.parametrize("b,n,d,S,kind", [(300, 40, 3, 3, "matern52"), (64, 64, 2, 1, "rbf"), (1000, 137, 20, 2, "matern52"),
                                          (257, 9, 40, 1, "matern52"), (1, 5, 2, 1, "matern52")])
def test_thompson_matches_oracle(b, n, d, S, kind):
    from bogp_b200.generation import MaxPosteriorSampling
    from bogp_b200.gp import SingleTaskGP
    X, y, ls, noise = _problem(n, d)
    rng = np.random.default_rng(1)
    Xc = rng.uniform(-1, 1, (b, d))
    z = rng.standard_normal((S, b))
    model = SingleTaskGP(X, y, kind, lengthscale=ls, outputscale=1.3, noise=noise)
    ts = MaxPosteriorSampling(model, replacement=False)
    X_next = ts(torch.as_tensor(Xc), num_samples=S, base_samples=z, return_samples=True)
    idx, f, jit = ogp.thompson_sample(ogp.GPState(X, y, kind, ls, 1.3, 0.0, noise), Xc, z)
    assert ts.last_jitter == jit == 0.0
    assert ts.last_indices.tolist() == idx.tolist()                        # bit-identical picks
    assert X_next.shape == (S, d) and np.array_equal(X_next.numpy(), Xc[idx])
    fs = ts.last_samples.cpu()
    assert float((fs - f).abs().max()) <= 1e-9 * float(f.abs().max())       # float64 both sides (bar: 1e-5)
    assert torch.allclose(ts.last_values, f[torch.arange(S), torch.as_tensor(idx)], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_thompson_jitter_on_singular_covariance():
    """Duplicated candidates make Sigma exactly singular: the factorisation must fail, retry with jitter as
    psd_safe_cholesky does, and still pick what the oracle picks."""
    from bogp_b200.generation import MaxPosteriorSampling
    from bogp_b200.gp import SingleTaskGP
    X, y, ls, noise = _problem(30, 3)
    rng = np.random.default_rng(2)
    Xc = rng.uniform(-1, 1, (100, 3))
    Xc = np.concatenate([Xc, Xc[:50]])               # 50 exact duplicates
    z = rng.standard_normal((1, 150))
    model = SingleTaskGP(X, y, lengthscale=ls, outputscale=1.0, noise=noise)
    ts = MaxPosteriorSampling(model)
    ts(Xc, base_samples=z, return_samples=True)
    idx, f, jit = ogp.thompson_sample(ogp.GPState(X, y, "matern52", ls, 1.0, 0.0, noise), Xc, z)
    assert ts.last_jitter > 0.0 and jit > 0.0
    if ts.last_jitter == jit:
        assert float((ts.last_samples.cpu() - f).abs().max()) <= 1e-3 * float(f.abs().max())
    # the draw is only defined up to the jitter; the pick must be one of the oracle's near-maximal candidates
    assert float(f[0, ts.last_indices[0]]) >= float(f[0].max()) - 1e-3 * float(f[0].abs().max())


@pytest.mark.gpu
def test_thompson_rejects_bad_arguments():
    from bogp_b200 import _lib
    from bogp_b200.generation import MaxPosteriorSampling
    from bogp_b200.gp import SingleTaskGP
    X, y, ls, noise = _problem(10, 2)
    model = SingleTaskGP(X, y, lengthscale=ls, noise=noise)
    ts = MaxPosteriorSampling(model)
    with pytest.raises(ValueError):
        ts(np.zeros((5, 3)))
    with pytest.raises(ValueError):
        ts(np.zeros((5, 2)), num_samples=6)
    with pytest.raises(_lib.BogpError):
        ts(np.zeros((20000, 2)))
    with pytest.raises(NotImplementedError):
        MaxPosteriorSampling(model, replacement=True)
    # a model with more than 16 input dimensions samples fine but the acquisition kernels refuse it, loudly
    from bogp_b200.acquisition import ExpectedImprovement
    X, y, ls, noise = _problem(12, 24)
    big = SingleTaskGP(X, y, lengthscale=ls, noise=noise)
    assert MaxPosteriorSampling(big)(np.random.default_rng(0).uniform(-1, 1, (70, 24))).shape == (1, 24)
    with pytest.raises(_lib.BogpError):
        ExpectedImprovement(big, 0.0)(torch.zeros(3, 1, 24, dtype=torch.float64, device="cuda"))


@pytest.mark.gpu
def test_thompson_full_size_properties():
    """The reference's size (5000 candidates, baxus.py:45) against size-independent properties: z = 0 returns the
    arg-max of the posterior mean, and L L^T reproduces the marginal posterior variances (diag via unit vectors is too
    slow, so: Var[f_c] from the row norms of L through samples with one-hot z)."""
    from bogp_b200.generation import MaxPosteriorSampling
    from bogp_b200.gp import SingleTaskGP
    X, y, ls, noise = _problem(300, 12)
    rng = np.random.default_rng(3)
    Xc = rng.uniform(-1, 1, (5000, 12))
    model = SingleTaskGP(X, y, lengthscale=ls, outputscale=1.3, noise=noise)
    ts = MaxPosteriorSampling(model)
    ts(Xc, base_samples=np.zeros((1, 5000)), return_samples=True)
    post = model.posterior(torch.as_tensor(Xc).unsqueeze(1))
    mu = post.mean.reshape(-1).cpu()
    assert ts.last_jitter == 0.0
    assert torch.allclose(ts.last_samples.cpu()[0], mu, rtol=1e-9, atol=1e-9)
    assert int(ts.last_indices[0]) == int(torch.argmax(mu))
    # one-hot z_j = 1 gives f - mu = column j of L; j = 0: L[:, 0] = Sigma[:, 0] / sqrt(Sigma[0, 0])
    e0 = np.zeros((1, 5000)); e0[0, 0] = 1.0
    ts(Xc, base_samples=e0, return_samples=True)
    col0 = ts.last_samples.cpu()[0] - mu
    var0 = float(post.variance.reshape(-1)[0])
    assert abs(float(col0[0]) - math.sqrt(var0)) <= 1e-7 * math.sqrt(var0)
    _, cov = ogp.posterior(ogp.GPState(X, y, "matern52", ls, 1.3, 0.0, noise), torch.as_tensor(Xc[:200]))
    assert torch.allclose(col0[:200], cov[:, 0] / math.sqrt(float(cov[0, 0])), rtol=1e-6, atol=1e-9)
    # last row: ||L[b-1, :]||^2 = Sigma[b-1, b-1] -- through z = e_j for a few j is O(b); instead use all-ones z on the
    # last candidate only: f - mu = sum_j L[b-1, j]; compare with the oracle on a 64-candidate tail problem
    tail = Xc[-64:]
    zt = rng.standard_normal((1, 64))
    ts(tail, base_samples=zt, return_samples=True)
    idx, f, _ = ogp.thompson_sample(ogp.GPState(X, y, "matern52", ls, 1.3, 0.0, noise), tail, zt)
    assert ts.last_indices.tolist() == idx.tolist()


@pytest.mark.gpu
def test_bounded_fit_matches_oracle():
    """Interval-constrained fit (baxus.py:304-319) -- fused kernel and torch loop against the oracle's AdamW loop."""
    from bogp_b200.baxus import baxus_fit_config
    from bogp_b200.gp import SingleTaskGP, fit_gp_adamw
    X, y, _, _ = _problem(40, 3)
    cfg = baxus_fit_config(steps=60)
    ocfg = ogp.FitConfig(steps=60, lengthscale_prior=None, outputscale_prior=None, lengthscale_bounds=(0.005, 10.0),
                         outputscale_bounds=(0.05, 10.0))
    ref = ogp.fit_adamw(X, y, ocfg)
    for engine in ("fused", "torch"):
        m = fit_gp_adamw(SingleTaskGP(X, y), cfg, engine=engine)
        assert np.allclose(m.lengthscale.numpy(), ref.lengthscale.numpy(), rtol=2e-6), engine
        assert abs(m.outputscale - ref.outputscale) <= 2e-6 * ref.outputscale
        assert abs(m.noise - ref.noise) <= 2e-6 * ref.noise + 1e-12
        assert 0.005 < float(m.lengthscale.min()) and float(m.lengthscale.max()) < 10.0 and 0.05 < m.outputscale < 10.0


@pytest.mark.gpu
def test_baxus_loop_embedded_objective():
    """End to end: a 2-D bowl hidden in 30 dimensions.  BAxUS must beat its own Sobol warm-up, keep every point inside
    [-1, 1]^dim, and report one time per evaluation (the contract of loop(), baxus.py:262-425)."""
    from bogp_b200.baxus import BAxUSLoopArgs, loop
    calls = []

    def objective(x, true_dim=None):
        assert x.shape[1] == 30 and np.all(np.abs(x) <= 1.0)
        calls.append(len(x))
        xs = x[:, :true_dim]
        return list(((xs[:, 0] - 0.3) ** 2 + (xs[:, 1] + 0.5) ** 2 + 0.05 * np.sin(5 * xs[:, 0])).astype(float))

    torch.manual_seed(0)
    args = BAxUSLoopArgs(true_dim=2, budget=45, n_init=15, dummy_dim=28)
    X, Y, times = loop(objective, dim=args.dim, budget=args.budget, n_init=args.n_init, true_dim=2,
                       n_candidates=600, seed=3)
    assert X.shape == (45, 2) and Y.shape == (45,) and len(times) == 45
    assert calls[0] == 15 and all(c == 1 for c in calls[1:])
    assert float(Y[15:].min()) < float(Y[:15].min())          # the model-based phase improves on the warm-up
