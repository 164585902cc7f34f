"""Pin the CPU oracle against the reference's own golden artefacts.

Fixtures come from ``tests/golden/make_golden.py`` (reference test data:
``tests/data/input/models/*_model.zip``, ``tests/data/reference/train_colvars/*.csv``,
``tests/data/reference/traj_cluster/*.csv``) and from the reference's own
``statistics.cluster_data`` for KMeans.
"""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN


def _standardized(c1, cv):
    return oracle.standardize(c1["X"], c1[f"{cv}_features_norm_mean"], c1[f"{cv}_features_norm_range"])


def test_stats_match_golden_norms(c1):
    st = oracle.column_stats(c1["X"])
    mean, rng = oracle.prepare_normalization(st, "mean_std")
    for cv in ("pca", "tica", "htica"):
        np.testing.assert_allclose(mean, c1[f"{cv}_features_norm_mean"], rtol=0, atol=2e-7)
        # ddof=1 pinned: ddof=0 would differ by ~3e-3 relative
        np.testing.assert_allclose(rng, c1[f"{cv}_features_norm_range"], rtol=1e-6)


def test_normalization_modes():
    st = {"mean": np.array([1.0, 2.0]), "std": np.array([0.5, 0.0]),
          "min": np.array([0.0, 2.0]), "max": np.array([4.0, 2.0])}
    m, r = oracle.prepare_normalization(st, None)
    assert np.array_equal(m, [0, 0]) and np.array_equal(r, [1, 1])
    m, r = oracle.prepare_normalization(st, "mean_std")
    assert np.array_equal(m, [1, 2]) and np.array_equal(r, [0.5, 1.0])   # 0 range -> 1
    m, r = oracle.prepare_normalization(st, "min_max_range1")
    assert np.array_equal(m, [0, 2]) and np.array_equal(r, [4, 1])
    m, r = oracle.prepare_normalization(st, "min_max_range2")
    assert np.array_equal(m, [2, 2]) and np.array_equal(r, [2, 1])
    with pytest.raises(ValueError):
        oracle.prepare_normalization(st, "bogus")


@pytest.mark.parametrize("cv", ["pca", "tica", "htica"])
def test_projection_reproduces_golden_csv(c1, cv):
    """Golden weights + the reference's projection / normalisation formulae give the golden
    CSV at its 4 printed decimals (pins A4, A9, A10)."""
    P = oracle.project_normalized(c1["X"], c1[f"{cv}_features_norm_mean"],
                                  c1[f"{cv}_features_norm_range"], c1[f"{cv}_cv_weights"],
                                  c1[f"{cv}_cv_norm_mean"], c1[f"{cv}_cv_norm_range"])
    got = np.array([[float("%.4f" % v) for v in row] for row in P])
    mism = np.sum(got != c1[f"{cv}_csv"])
    assert mism <= 2, f"{mism} cells differ at 4 decimals"
    np.testing.assert_allclose(P, c1[f"{cv}_csv"], atol=1.01e-4)


@pytest.mark.parametrize("cv", ["pca", "tica", "htica"])
def test_cv_normalization_matches_golden(c1, cv):
    Z = _standardized(c1, cv)
    P = (Z @ c1[f"{cv}_cv_weights"]).astype(np.float32)
    cm, cr = oracle.cv_normalization(P)
    np.testing.assert_allclose(cm, c1[f"{cv}_cv_norm_mean"], atol=1e-6)
    np.testing.assert_allclose(cr, c1[f"{cv}_cv_norm_range"], rtol=1e-6)


def test_pca_weights_match_golden(c1):
    Z = _standardized(c1, "pca")
    _, W = oracle.pca(Z, 2)
    np.testing.assert_allclose(W, c1["pca_cv_weights"], atol=5e-6)


def test_tica_weights_match_golden(c1):
    """mlcolvar restatement pinned: N - lag pairs, mean of x_t, /M, symmetrised C_tau,
    reg 1e-6, unit-L2 columns, sign of row 0.  Residual 1e-4 is LAPACK noise on an
    ill-conditioned problem."""
    Z = _standardized(c1, "tica")
    evals, W = oracle.tica(Z, 1, 2)
    assert evals[0] > evals[1] > 0.9
    np.testing.assert_allclose(W, c1["tica_cv_weights"], atol=3e-4)
    # N - lag - 1 pairs would be off by ~0.2: make sure we are far from that
    assert np.abs(W - c1["tica_cv_weights"]).max() < 1e-3


def test_tica_from_sums_equals_direct(c1):
    Z = _standardized(c1, "tica")
    e1, W1 = oracle.tica(Z, 1, 2)
    e2, W2 = oracle.tica_from_sums(*oracle.lagged_sums(Z, 1), 2)
    np.testing.assert_allclose(e1, e2, rtol=1e-9)
    np.testing.assert_allclose(W1, W2, atol=1e-7)


def test_htica_weights_match_golden(c1):
    Z = _standardized(c1, "htica")
    assert len(oracle.htica_chunks(54, 10)) == 11          # 10 x 5 + 1 x 4 (torch.split by size)
    W, T1, V2 = oracle.htica(Z, 1, 10, 5, 2)
    assert T1.shape == (54, 54) and V2.shape == (54, 2)
    np.testing.assert_allclose(W, c1["htica_cv_weights"], atol=2e-4)


def test_htica_chunks_edge_cases():
    assert oracle.htica_chunks(4950, 10) == [(i * 495, (i + 1) * 495) for i in range(10)]
    assert oracle.htica_chunks(5, 10) == []
    assert oracle.htica_chunks(7, 3) == [(0, 2), (2, 4), (4, 6), (6, 7)]


@pytest.mark.parametrize("cv", ["pca", "tica", "htica"])
def test_find_centroids_matches_golden_cluster_csv(c1, cv):
    """Golden traj_cluster CSVs (hierarchical labels): centroids are cluster means
    (statistics.py:330-335) and find_centroids marks the closest sample (K3)."""
    Y = c1[f"{cv}_cluster_cv"]
    lab = c1[f"{cv}_cluster_label"]
    ks = np.unique(lab)
    cent = np.stack([Y[lab == k].mean(axis=0) for k in ks])
    idx = oracle.find_centroids(Y, cent)
    flags = np.zeros(len(Y), dtype=bool)
    flags[idx] = True
    assert np.array_equal(flags, c1[f"{cv}_cluster_centroid"])


@pytest.mark.parametrize("name", ["blobs_d2_k5", "blobs_d4_k10_grid", "blobs_d10_k40",
                                  "uniform_d3_k7_grid"])
def test_kmeans_restatement_matches_reference_cluster_data(kmeans_ref, name):
    """Lloyd restatement == the reference's statistics.cluster_data(initial_centroids=...)."""
    X = kmeans_ref[f"{name}_X"]
    res = oracle.kmeans_lloyd(X, kmeans_ref[f"{name}_init"])
    ref_labels = kmeans_ref[f"{name}_labels"]
    diff = np.flatnonzero(res["labels"] != ref_labels)
    # any disagreement must be an exact / near tie
    gap = res["second"][diff] - res["best"][diff]
    assert np.all(gap <= 1e-9 * np.maximum(res["best"][diff], 1e-30)), (len(diff), gap)
    if len(diff) == 0:
        np.testing.assert_allclose(res["centers"], kmeans_ref[f"{name}_centers"], rtol=1e-9, atol=1e-12)


def test_kmeans_matches_installed_sklearn_float64():
    from sklearn.cluster import KMeans
    rng = np.random.default_rng(3)
    X = np.round(rng.uniform(-1, 1, size=(2500, 2)), 3)      # coarse grid -> exact ties exist
    init = X[:6].copy()
    km = KMeans(n_clusters=6, random_state=0, init=init.copy(), n_init=1).fit(X.copy())
    res = oracle.kmeans_lloyd(X, init)
    diff = np.flatnonzero(res["labels"] != km.labels_)
    gap = res["second"][diff] - res["best"][diff]
    assert np.all(np.abs(gap) <= 1e-9), gap
    assert res["n_iter"] == km.n_iter_


def test_kmeans_empty_cluster_relocation():
    X = np.array([[0.0, 0.0], [0.1, 0.0], [5.0, 5.0], [5.1, 5.0], [9.0, 9.0]])
    init = np.array([[0.0, 0.0], [5.0, 5.0], [100.0, 100.0]])       # third centre starts empty
    from sklearn.cluster import KMeans
    km = KMeans(n_clusters=3, random_state=0, init=init.copy(), n_init=1).fit(X.copy())
    res = oracle.kmeans_lloyd(X, init)
    assert np.array_equal(res["labels"], km.labels_)
    np.testing.assert_allclose(res["centers"], km.cluster_centers_, rtol=1e-12)


def test_cluster_scores_match_sklearn(kmeans_ref):
    from sklearn.metrics import calinski_harabasz_score, davies_bouldin_score
    X = kmeans_ref["blobs_d4_k10_grid_X"]
    lab = kmeans_ref["blobs_d4_k10_grid_labels"]
    assert np.isclose(oracle.calinski_harabasz(X, lab), calinski_harabasz_score(X, lab), rtol=1e-10)
    assert np.isclose(oracle.davies_bouldin(X, lab), davies_bouldin_score(X, lab), rtol=1e-10)


def test_deeptica_loss_matches_torch_autograd_free_form():
    """A.4 restatement vs an independent torch float64 evaluation."""
    import torch
    rng = np.random.default_rng(0)
    f = rng.standard_normal((512, 3))
    g = 0.9 * f + 0.3 * rng.standard_normal((512, 3))
    loss, evals = oracle.deeptica_loss(f, g, reg=1e-6)
    ft = torch.tensor(f); gt = torch.tensor(g)
    mu = ft.mean(0)
    a = ft - mu; b = gt - mu
    C0 = a.T @ a / 512
    Ct = 0.5 * (a.T @ b + b.T @ a) / 512
    L = torch.linalg.cholesky(C0 + 1e-6 * torch.eye(3, dtype=torch.float64))
    Li = torch.linalg.inv(L)
    ev = torch.linalg.eigvalsh(Li @ Ct @ Li.T).flip(0)
    np.testing.assert_allclose(evals, ev.numpy(), rtol=1e-10)
    assert np.isclose(loss, -(ev ** 2).sum().item())


def test_reference_deeptica_model_reproduces_its_golden_projection(c1):
    """the reference's golden DeepTICA model (TorchScript `cv_weights.pt` inside
    `deep_tica_model.zip`, loadable with plain torch) maps the C1 features onto the golden
    `deep_tica_projected_trajectory.csv` to print precision -- the fixture that pins A12."""
    import io
    import zipfile
    import torch
    z = zipfile.ZipFile(os.path.join(GOLDEN, "deep_tica_model.zip"))
    meta = json.loads(z.read("model/metadata.json"))
    assert meta == {"cv_name": "deep_tica", "cv_dimension": 2}
    labels = z.read("model/features_labels.txt").decode().strip().split("\n")
    assert labels == [str(v) for v in c1["features"]]
    model = torch.jit.load(io.BytesIO(z.read("model/cv_weights.pt")), map_location="cpu")
    with torch.no_grad():
        P = model(torch.from_numpy(c1["X"].copy())).numpy()
    assert list(c1["deep_tica_csv_cols"]) == ["DeepTIC 1", "DeepTIC 2"]
    np.testing.assert_allclose(P, c1["deep_tica_csv"], atol=5.1e-5)


def test_float64_device_checker_equals_numpy_oracle():
    """oracle/float64_device.py (the torch float64 checker used at BASELINE sizes on the GPU box) is
    the same function as the numpy oracle that is pinned to the reference's golden artefacts."""
    import torch
    from conftest import synth_features
    from oracle import float64_device as f64
    X = synth_features(3000, 37, seed=4)
    st = oracle.column_stats(X)
    m, r = oracle.prepare_normalization(st, "mean_std")
    Z = oracle.standardize(X, m, r)
    Xt, mt, rt = torch.from_numpy(X), torch.from_numpy(m.astype(np.float32)), torch.from_numpy(r.astype(np.float32))
    lag = 7
    s = f64.lagged_sums(Xt, lag, mt, rt, chunk=700)
    ref = dict(zip(("S0", "St", "a", "b", "M"), oracle.lagged_sums(Z, lag)))
    for k in ("S0", "St", "a", "b"):
        np.testing.assert_allclose(s[k].numpy(), ref[k], rtol=1e-12, atol=1e-9)
    assert s["M"] == ref["M"]
    ev, V = f64.tica_from_sums(s["S0"], s["St"], s["a"], s["b"], s["M"], 3)
    ev_o, V_o = oracle.tica(Z, lag, 3)
    np.testing.assert_allclose(ev.numpy(), ev_o, rtol=1e-9)
    np.testing.assert_allclose(V.numpy(), V_o, atol=1e-8)
    W, T1, V2 = f64.htica(Xt, lag, mt, rt, 5, 3, 2, chunk=900)
    W_o, _, _ = oracle.htica(Z, lag, 5, 3, 2)
    np.testing.assert_allclose(W.numpy(), W_o, atol=1e-7)
    Pn, mn, mx = f64.project_normalized(Xt, mt, rt, V, chunk=1000)
    P = oracle.project(Z, V_o)
    cm, cr = oracle.cv_normalization(P)
    np.testing.assert_allclose(Pn.numpy(), (P - cm) / cr, atol=1e-7)
    assert f64.eigvec_error(V, -torch.from_numpy(V_o)) < 1e-8


def test_fes_restatement_reproduces_the_reference_legacy_output():
    """N4: oracle/fes_oracle.py (binned Gaussian KDE -> -kbT log) reproduces the FES the reference
    itself wrote for the bundled calpha_transitions example (mlcolvar compute_fes through KDEpy,
    bandwidth 0.025, 200 bins, 300 K) from the projected trajectory stored next to it."""
    from oracle import fes_oracle as fo
    g = dict(np.load(os.path.join(GOLDEN, "fes_legacy_pca.npz")))
    fes, grid, bounds, err = fo.compute_fes(g["X"], float(g["temperature"]), int(g["num_bins"]),
                                            [tuple(b) for b in g["bounds"]], float(g["bandwidth"]), 1, 1e-10)
    assert err is None and fes.shape == g["fes"].shape
    np.testing.assert_allclose(np.asarray(grid), g["grid"], atol=1e-12)
    low = g["fes"] < 20                      # KDEpy truncates the kernel where the density is negligible
    assert np.abs(fes - g["fes"])[low].max() < 2e-3
    assert np.unravel_index(fes.argmin(), fes.shape) == np.unravel_index(g["fes"].argmin(), g["fes"].shape)
    # the legacy bounds are the data range +/- 5 % (today's get_ranges uses 0.5 %, figures.py:441)
    lo, hi = g["X"].min(0), g["X"].max(0)
    np.testing.assert_allclose(g["bounds"], np.stack([lo - 0.05 * (hi - lo), hi + 0.05 * (hi - lo)], 1), rtol=1e-9)


def test_fes_block_average_properties():
    from oracle import fes_oracle as fo
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.normal(-0.5, 0.1, 4000), rng.normal(0.4, 0.2, 6000)])
    rng.shuffle(x)
    fes1, grid, b, e1 = fo.compute_fes(x, 300, 150, fo.get_ranges(x), 0.05, 1, 1e-10)
    fesb, _, _, eb = fo.compute_fes(x, 300, 150, fo.get_ranges(x), 0.05, 100, 1e-10)
    assert e1 is None and eb.shape == fesb.shape == (150,)
    core = fes1 < 3                                       # well-sampled region (100 frames per block:
    assert np.abs(fesb - fes1)[core].max() < 1.0          # averaging -log of noisy densities is biased in the tails)
    assert (eb[core] < 0.5).all() and (eb >= 0).all()
    # density integrates to one on the grid
    dens = fo.binned_density(x, b, 150, 0.05)
    assert abs(np.trapezoid(dens, grid) - 1.0) < 1e-3
