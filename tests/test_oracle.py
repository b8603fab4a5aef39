"""CPU tier: the float64 oracle against independent implementations, in-repo formula pins and invariants.

The reference has no tests or golden vectors for this path; these are the authored pins (i)-(viii).
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import bartlett as ob
from oracle import field as of
from oracle import gp as ogp
from oracle import optimize as oopt

GOLD = Path(__file__).parent / "golden"


def test_truth_gives_unit_bartlett(cfg1):
    """(i) cbf = 1 and cbf_ml = 0 at the truth used to synthesise K (simulation mode, utils.py:23-32)."""
    m, K = cfg1["modes"], cfg1["K"]
    cand = np.array([[cfg1["truth"][0], cfg1["truth"][1], 0.0]])
    assert abs(ob.bartlett_points(m, K, cand)[0] - 1.0) < 1e-12
    assert abs(ob.bartlett_points(m, K, cand, atype="cbf_ml", multifreq_method="product")[0]) < 1e-12


def test_surface_matches_golden_and_bounds(cfg1):
    g = np.load(GOLD / "bartlett_cfg1.npz")
    surf = ob.ambiguity_surface(cfg1["modes"], cfg1["K"], cfg1["r"], cfg1["z"])
    np.testing.assert_allclose(surf, g["cbf_mean"], rtol=1e-12, atol=1e-14)
    assert surf.min() >= 0.0 and surf.max() <= 1.0 + 1e-12               # (ii)
    assert tuple(np.unravel_index(np.argmax(surf), surf.shape)) == tuple(g["argmax"])


def test_grid_equals_points_and_zero_tilt(cfg1):
    """(v) grid == points on grid nodes; (vi) zero tilt == vertical-array path."""
    m, K = cfg1["modes"], cfg1["K"]
    r, z = cfg1["r"][::9], cfg1["z"][::7]
    surf = ob.ambiguity_surface(m, K, r, z)
    cand = np.array([[ri, zi, 0.0] for zi in z for ri in r])
    np.testing.assert_allclose(ob.bartlett_points(m, K, cand).reshape(len(z), len(r)), surf, rtol=1e-12)
    p0 = of.replica(m, 1, 50.0, r, 0.0)
    dr0 = of.range_offsets(m.rec_z, 0.0, m.z_pivot)
    p1 = of.pressure_field(m.k[1], m.phi_rec[1], of.interp_phi_src(m.phi_src_tab[1], m.z0, m.dz, 50.0)[0],
                           1000.0 * r, dr0 + 1e-300)   # force the tilted code path with zero offsets
    np.testing.assert_allclose(p1, p0, rtol=1e-10)


def test_scale_invariance_and_single_tonal(cfg1):
    """(iii) invariance to complex scaling of the replica; (iv) mean == product when F = 1."""
    m, K = cfg1["modes"], cfg1["K"]
    p = of.replica(m, 0, 33.0, np.array([2.5, 4.0]))
    b = ob.beamformer(K[0], p)
    np.testing.assert_allclose(ob.beamformer(K[0], (0.3 - 2.0j) * p), b, rtol=1e-12)
    m1 = m.select([2])
    cand = np.array([[3.3, 80.0, 0.0]])
    assert ob.bartlett_points(m1, K[2:3], cand)[0] == pytest.approx(
        ob.bartlett_points(m1, K[2:3], cand, multifreq_method="product")[0], rel=1e-14)


@pytest.mark.parametrize("name,kind", [("m52_d2", "matern52"), ("m52_d7", "matern52"), ("rbf_d3", "rbf")])
def test_gp_posterior_and_acq_match_sklearn_scipy(name, kind):
    """Oracle posterior vs scikit-learn; EI / UCB / PI vs the in-repo SciPy formulas (bo_examples.py:78-84, :124)."""
    g = np.load(GOLD / "gp_sklearn.npz")
    gp = ogp.GPState(g[f"{name}_X"], g[f"{name}_y"], kind, g[f"{name}_ls"], float(g[f"{name}_s"]), 0.0,
                     float(g[f"{name}_noise"]))
    Xs = torch.as_tensor(g[f"{name}_Xs"]).unsqueeze(1)
    mu, var = ogp.posterior_mean_var(gp, Xs)
    np.testing.assert_allclose(mu.squeeze(-1).numpy(), g[f"{name}_mu"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(var.squeeze(-1).sqrt().numpy(), g[f"{name}_sd"], rtol=1e-5, atol=1e-8)
    best = float(g[f"{name}_best"])
    np.testing.assert_allclose(ogp.acq_value(gp, "ei", Xs, best).numpy(), g[f"{name}_ei"], rtol=1e-4, atol=1e-10)
    np.testing.assert_allclose(ogp.acq_value(gp, "ucb", Xs, beta=4.0).numpy(), g[f"{name}_ucb"], rtol=1e-6)
    np.testing.assert_allclose(ogp.acq_value(gp, "pi", Xs, best).numpy(), g[f"{name}_pi"], rtol=1e-4, atol=1e-10)
    logei = ogp.acq_value(gp, "logei", Xs, best).numpy()
    ok = g[f"{name}_ei"] > 1e-200
    np.testing.assert_allclose(logei[ok], np.log(g[f"{name}_ei"][ok]), rtol=1e-6, atol=1e-8)


def test_workload_constants_match_reference_files(cfg1):
    """Tonal sets, array geometry and the truth point of the synthetic configurations against the values read from the
    reference's own files (tools/make_golden_constants.py -> tests/golden/reference_constants.json)."""
    import json

    from bogp_b200 import modes as M
    g = json.loads((GOLD / "reference_constants.json").read_text())
    assert M.INV_TONALS_3 == g["inv_tonals"] and M.SWELLEX_TONALS_13 == g["swellex_tonals_13"]
    assert M.REC_Z_21 == g["rec_z_21"] and len(g["rec_z_21"]) == 21
    assert np.array_equal(M.rec_z_64(), np.asarray(g["rec_z_64"])) and g["rec_z_64_linspace"] == [94.125, 212.25, 64]
    assert (g["truth_r_km"], g["truth_src_z_m"]) == (1.07, 71.11)          # config 1's source (conftest cfg1)
    assert [p["name"] for p in g["inv_search_space"]] == ["rec_r", "src_z", "tilt", "h_w", "h_sed", "c_p_sed_top", "dc_p_sed"]
    assert cfg1["truth"] == (g["truth_r_km"], g["truth_src_z_m"])


def test_acquisition_formulas_match_reference_code():
    """EI / UCB against the outputs of the reference's OWN numpy functions (bo_examples.py:61-84, :107-124), executed
    from its source by tools/make_golden_acq.py -- these two rows of the oracle are pinned, not recalled."""
    g = np.load(GOLD / "acq_reference.npz")
    mu, sigma, best_f = torch.as_tensor(g["mu"]), torch.as_tensor(g["sigma"]), float(g["best_f"])
    ei = ogp.analytic_acq("ei", mu, sigma ** 2, best_f)
    assert float((ei - torch.as_tensor(g["ei"])).abs().max()) <= 1e-14
    ei_xi = ogp.analytic_acq("ei", mu, sigma ** 2, best_f + 0.05)              # xi shifts the incumbent
    assert float((ei_xi - torch.as_tensor(g["ei_xi005"])).abs().max()) <= 1e-14
    for kappa in (0.5, 1.0, 2.0):
        ucb = ogp.analytic_acq("ucb", mu, sigma ** 2, beta=kappa ** 2)          # botorch: mu + sqrt(beta) sigma
        assert float((ucb - torch.as_tensor(g[f"ucb_kappa_{kappa}"])).abs().max()) <= 1e-14
    # LogEI is the log of the same quantity wherever EI is representable
    ok = torch.as_tensor(g["ei"]) > 1e-200
    assert torch.allclose(ogp.analytic_acq("logei", mu, sigma ** 2, best_f)[ok], torch.log(torch.as_tensor(g["ei"]))[ok],
                          rtol=1e-9, atol=1e-9)


def test_gp_interpolates_and_grad_matches_fd():
    """(vii) posterior interpolates as noise -> 0, variance >= 0, EI >= 0; (viii) d acq/dX vs finite differences."""
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (20, 3))
    y = np.sin(2 * X[:, 0]) + X[:, 1] * X[:, 2]
    gp = ogp.GPState(X, y, "matern52", [0.6, 0.7, 0.8], 1.0, 0.0, 1e-10)
    mu, var = ogp.posterior_mean_var(gp, torch.as_tensor(X).unsqueeze(1))
    np.testing.assert_allclose(mu.squeeze(-1).numpy(), y, atol=1e-6)
    assert float(var.min()) >= 0.0
    gp = ogp.GPState(X, y, "matern52", [0.6, 0.7, 0.8], 1.0, 0.0, 1e-4)
    Xs = torch.as_tensor(rng.uniform(-1, 1, (6, 1, 3)))
    for kind in ("ei", "logei", "pi", "ucb"):
        v, g = ogp.acq_value_and_grad(gp, kind, Xs, best_f=float(y.max()), beta=2.0)
        if kind in ("ei", "pi"):
            assert float(v.min()) >= 0.0
        h = 1e-6
        for a in range(3):
            e = torch.zeros_like(Xs)
            e[..., a] = h
            fd = (ogp.acq_value(gp, kind, Xs + e, float(y.max()), 2.0) - ogp.acq_value(gp, kind, Xs - e, float(y.max()), 2.0)) / (2 * h)
            np.testing.assert_allclose(g[:, 0, a].numpy(), fd.numpy(), rtol=2e-5, atol=1e-9)


def test_qei_q1_tends_to_analytic_ei():
    rng = np.random.default_rng(1)
    X = rng.uniform(-1, 1, (24, 2))
    y = np.cos(3 * X[:, 0]) * X[:, 1]
    gp = ogp.GPState(X, y, "matern52", [0.5, 0.5], 1.0, 0.0, 1e-4)
    Xs = torch.as_tensor(rng.uniform(-1, 1, (16, 1, 2)))
    base = ogp.sobol_normal_samples(1, 4096, 0)
    qei = ogp.acq_value(gp, "qei", Xs, float(y.max()), base_samples=base)
    ei = ogp.acq_value(gp, "ei", Xs, float(y.max()))
    np.testing.assert_allclose(qei.numpy(), ei.numpy(), rtol=5e-2, atol=2e-4)


def test_oracle_bo_loop_finds_source(cfg1):
    """Config 1 end-to-end on the oracle: 2-D (range, depth) BO converges near the truth with a tiny budget."""
    m, K = cfg1["modes"], cfg1["K"]
    lo, hi = np.array([0.57, 51.0]), np.array([1.57, 91.0])     # +-0.5 km, +-20 m around the truth

    def objective(Xn):
        Xn = np.atleast_2d(Xn)
        phys = lo + (Xn + 1.0) * 0.5 * (hi - lo)
        cand = np.concatenate([phys, np.zeros((len(phys), 1))], axis=1)
        return -ob.bartlett_points(m, K, cand)            # minimise -B (ei.py:49 negates again)

    X, Y = oopt.bo_loop(objective, dim=2, budget=18, n_init=8, num_restarts=4, raw_samples=64, seed=0,
                        fit_cfg=ogp.FitConfig(steps=25))
    assert X.shape == (18, 2) and np.isfinite(Y.numpy()).all()
    assert float((-Y).max()) > 0.9 and float((-Y)[8:].max()) > float((-Y)[:8].max())
