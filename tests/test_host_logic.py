"""CPU tests: the C-ABI library loads and exports what include/dcg.h declares, argument errors
are reported without a GPU, the FP64 eigen stage (device-agnostic torch) matches the oracle and
the reference's golden weights, and the host-side mirrors (schemas, colvars reader, model.zip,
Lloyd driver control flow) behave like the reference."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest
import torch

import oracle
from conftest import GOLDEN, ROOT, synth_features


# ------------------------------------------------------------------------------------------------
# boundary
# ------------------------------------------------------------------------------------------------
def test_library_exports_every_symbol_declared_in_header():
    from deep_cartograph_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "dcg.h")).read()
    declared = set(re.findall(r"\b(dcg_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dcg_version() == 100
    assert lib.dcg_error_string(0) == b"ok"
    assert b"workspace" in lib.dcg_error_string(-1002)


def test_argument_errors_without_gpu():
    from deep_cartograph_b200 import _lib
    lib = _lib.load()
    assert lib.dcg_colstats_f32(None, 10, 4, 4, None, None, None, None, None, 0, None) == -1000
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    assert lib.dcg_colstats_f32(p, 0, 4, 4, p, p, p, p, p, 64, None) == -1001
    assert lib.dcg_colstats_f32(p, 10, 4, 2, p, p, p, p, p, 64, None) == -1001           # ld < f
    assert lib.dcg_colstats_f32(p, 10, 4, 4, p, p, p, p, None, 0, None) == -1002
    assert lib.dcg_cov_lag_f32(p, 10, 4, 4, 10, None, None, 0, p, p, p, p, 0, p, 1 << 20, None) == -1001  # lag >= n
    assert lib.dcg_cov_lag_f32(p, 10, 4, 4, 1, None, None, 0, p, p, p, p, 7, p, 1 << 20, None) == -1005   # engine
    assert lib.dcg_project_f32(p, 10, 4, 4, None, None, p, 65, p, None, None, p, 1 << 20, None) == -1001
    assert lib.dcg_kmeans_step(p, 10, 40, 40, 4, p, 3, p, p, p, p, None, 1, None, p, 64, None) == -1001
    assert lib.dcg_kmeans_step(p, 10, 4, 4, 2, p, 3, p, p, p, p, None, 1, None, p, 64, None) == -1005
    # hTICA level-1 block projection: s <= 16, block <= 1018, 16-byte aligned rows
    assert lib.dcg_project_blocks_f32(p, 10, 8, 8, None, None, p, 4, 17, p, 34, p, 1 << 20, None) == -1001
    assert lib.dcg_project_blocks_f32(p, 10, 2000, 2000, None, None, p, 1019, 4, p, 8, p, 1 << 20, None) == -1001
    assert lib.dcg_project_blocks_f32(p, 10, 8, 8, None, None, p, 4, 2, p, 3, p, 1 << 20, None) == -1001   # p_ld too small
    assert lib.dcg_project_blocks_f32(p, 10, 6, 6, None, None, p, 3, 2, p, 4, p, 1 << 20, None) == -1003   # ld % 4 != 0
    assert lib.dcg_project_blocks_f32(p, 10, 8, 8, None, None, p, 4, 2, p, 4, None, 0, None) == -1002
    assert lib.dcg_project_blocks_workspace_bytes(1000, 4950, 495) == 10 * 8 * 8 * 32 * (16 + 4) * 4
    assert lib.dcg_colstats_workspace_bytes(1000, 10) > 0
    assert lib.dcg_cov_workspace_bytes(1000, 1000, 10, 0, 1) >= 256 + 26 * 32     # 16 S_tau + 10 S0 super-tile descriptors
    assert lib.dcg_ticacov_out_doubles(3) == 2 + 3 + 18 + 6


def test_ops_have_no_cpu_fallback():
    from deep_cartograph_b200 import ops
    X = torch.zeros(8, 4)
    for call in (lambda: ops.column_stats(X),
                 lambda: ops.standardize_(X, torch.zeros(4), torch.ones(4)),
                 lambda: ops.lagged_covariance(X, 1),
                 lambda: ops.project(X, torch.zeros(4, 2)),
                 lambda: ops.project_blocks(X, torch.zeros(4, 1), 2),
                 lambda: ops.kmeans_step(X, torch.zeros(2, 4, dtype=torch.float64), torch.zeros(8, dtype=torch.int32)),
                 lambda: ops.nearest_to_centers(X, torch.zeros(2, 4, dtype=torch.float64))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_product_package_does_not_import_oracle():
    import subprocess
    import sys
    code = ("import sys; import deep_cartograph_b200.tools, deep_cartograph_b200.modules.statistics, "
            "deep_cartograph_b200.modules.cv_learning.deep_tica; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "deep_cartograph_b200")):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), fn


# ------------------------------------------------------------------------------------------------
# FP64 eigen stage (torch.linalg; runs on CPU tensors here, on the device in production)
# ------------------------------------------------------------------------------------------------
def _sums_t(Z, lag):
    S0, St, a, b, M = oracle.lagged_sums(Z, lag)
    return torch.from_numpy(S0), torch.from_numpy(St), torch.from_numpy(a), torch.from_numpy(b), M


def test_tica_from_sums_matches_oracle_and_golden(c1):
    from deep_cartograph_b200 import linalg
    Z = oracle.standardize(c1["X"], c1["tica_features_norm_mean"], c1["tica_features_norm_range"])
    S0, St, a, b, M = _sums_t(Z, 1)
    evals, V = linalg.tica_from_sums(S0, St, a, b, M, 2)
    revals, rV = oracle.tica(Z, 1, 2)
    np.testing.assert_allclose(evals.numpy(), revals, rtol=1e-8)
    np.testing.assert_allclose(V.numpy(), rV, atol=1e-6)        # ill-conditioned: cond(C0) ~ 1e4
    np.testing.assert_allclose(V.numpy(), c1["tica_cv_weights"], atol=3e-4)


def test_pca_from_sums_matches_golden(c1):
    from deep_cartograph_b200 import linalg
    Z = oracle.standardize(c1["X"], c1["pca_features_norm_mean"], c1["pca_features_norm_range"]).astype(np.float64)
    evals, W = linalg.pca_from_sums(torch.from_numpy(Z.T @ Z), torch.from_numpy(Z.sum(0)), Z.shape[0], 2)
    np.testing.assert_allclose(W.numpy(), c1["pca_cv_weights"], atol=5e-6)
    from sklearn.decomposition import PCA
    pca = PCA(n_components=2).fit(Z)
    np.testing.assert_allclose(evals.numpy(), pca.explained_variance_, rtol=1e-9)


def test_htica_from_full_sums_matches_oracle_and_golden(c1):
    from deep_cartograph_b200 import linalg
    Z = oracle.standardize(c1["X"], c1["htica_features_norm_mean"], c1["htica_features_norm_range"])
    S0, St, a, b, M = _sums_t(Z, 1)
    W, T1, V2 = linalg.htica_from_full_sums(S0, St, a, b, M, 10, 5, 2)
    rW, rT1, rV2 = oracle.htica(Z, 1, 10, 5, 2)
    assert linalg.htica_chunks(54, 10) == oracle.htica_chunks(54, 10)
    np.testing.assert_allclose(T1.numpy(), rT1, atol=1e-7)
    np.testing.assert_allclose(W.numpy(), rW, atol=1e-6)
    np.testing.assert_allclose(W.numpy(), c1["htica_cv_weights"], atol=2e-4)


def test_htica_block_diagonal_level1_equals_full():
    from deep_cartograph_b200 import linalg
    Z = oracle.standardize(*(lambda X: (X, X.mean(0), X.std(0, ddof=1)))(synth_features(4000, 60, seed=4)))
    S0, St, a, b, M = _sums_t(Z, 5)
    chunks = linalg.htica_chunks(60, 4)
    T_full = linalg.htica_level1(S0, St, a, b, M, chunks, 3)
    mask = torch.zeros(60, 60, dtype=torch.bool)
    for s, e in chunks:
        mask[s:e, s:e] = True
    T_blk = linalg.htica_level1(S0 * mask, St * mask, a, b, M, chunks, 3)
    np.testing.assert_allclose(T_full.numpy(), T_blk.numpy(), atol=1e-12)


def _slow_mode_sums(F, n=6000, lag=5, slow=6, seed=0):
    g = np.random.default_rng(seed)
    rho = np.exp(-1.0 / (300.0 * 0.5 ** np.arange(slow)))
    z = np.zeros((n, slow))
    e = g.standard_normal((n, slow))
    for t in range(1, n):
        z[t] = rho * z[t - 1] + np.sqrt(1 - rho ** 2) * e[t]
    X = z @ g.standard_normal((slow, F)) + 0.5 * g.standard_normal((n, F))
    X = (X - X.mean(0)) / X.std(0)
    M = n - lag
    t = torch.from_numpy
    return t(X[:M].T @ X[:M]), t(X[:M].T @ X[lag:]), t(X[:M].sum(0)), t(X[lag:].sum(0)), M


@pytest.mark.parametrize("F,out", [(256, 3), (400, 5)])
def test_partial_eigensolver_matches_dense_route(F, out, monkeypatch):
    """The shift-and-invert iteration returns the eigenpairs of the dense Cholesky route
    (mlcolvar cholesky_eigh) to 2e-8, far inside the 1e-5 parity tolerance."""
    from deep_cartograph_b200 import linalg
    S0, St, a, b, M = _slow_mode_sums(F)
    monkeypatch.setattr(linalg, "_PARTIAL_MIN_F", 10 ** 9)
    e_dense, V_dense = linalg.tica_from_sums(S0, St, a, b, M, out)
    monkeypatch.setattr(linalg, "_PARTIAL_MIN_F", 192)
    fast0 = linalg.EIG_STATS["fast"]
    e_fast, V_fast = linalg.tica_from_sums(S0, St, a, b, M, out)
    assert linalg.EIG_STATS["fast"] == fast0 + 1
    np.testing.assert_allclose(e_fast.numpy(), e_dense.numpy(), rtol=1e-10)
    np.testing.assert_allclose(V_fast.numpy(), V_dense.numpy(), atol=2e-8)


def test_partial_eigensolver_batched_blocks_match_single_solves():
    """hTICA level 1 solves its equal-width blocks as one batch; every block must equal its own
    single solve."""
    from deep_cartograph_b200 import linalg
    S0, St, a, b, M = _slow_mode_sums(512, n=5000, slow=6, seed=4)
    chunks = linalg.htica_chunks(512, 2)
    assert chunks == [(0, 256), (256, 512)]
    fast0 = linalg.EIG_STATS["fast"]
    T1 = linalg.htica_level1(S0, St, a, b, M, chunks, 3)
    assert linalg.EIG_STATS["fast"] == fast0 + 2
    c = 0
    for (s0, e0) in chunks:
        _, Vb = linalg.tica_from_sums(S0[s0:e0, s0:e0], St[s0:e0, s0:e0], a[s0:e0], b[s0:e0], M, 3)
        np.testing.assert_allclose(T1[s0:e0, c:c + 3].numpy(), Vb.numpy(), atol=2e-8)
        assert float(T1[:s0, c:c + 3].abs().max()) == 0.0 if s0 else True
        c += 3


def test_partial_eigensolver_falls_back_on_flat_spectrum():
    """More eigenpairs requested than there are slow modes: the wanted eigenvalues sit in the
    noise bulk, the iteration gives up and the dense route answers (same result as always)."""
    from deep_cartograph_b200 import linalg
    S0, St, a, b, M = _slow_mode_sums(256, slow=2)
    dense0 = linalg.EIG_STATS["dense"]
    evals, V = linalg.tica_from_sums(S0, St, a, b, M, 12)
    assert linalg.EIG_STATS["dense"] == dense0 + 1
    C0 = (S0 / M - torch.outer(a / M, a / M)).numpy()
    Ct = (St / M - torch.outer(a / M, b / M)).numpy()
    from oracle import cv_oracle
    ref, _ = cv_oracle._cholesky_eigh(0.5 * (C0 + C0.T), 0.5 * (Ct + Ct.T), 1e-6, 12)
    np.testing.assert_allclose(evals.numpy(), ref, rtol=1e-9, atol=1e-12)


def test_restandardize_sums_is_exact_algebra():
    """Sums accumulated under provisional standardisation (m0, r0) map to the sums under the final
    (mean, range) by the FP64 affine correction the streamed loader relies on."""
    from deep_cartograph_b200 import linalg
    g = np.random.default_rng(4)
    n, f, lag = 500, 7, 3
    X = g.standard_normal((n, f)) * g.uniform(0.1, 2.0, f) + g.uniform(-3, 3, f)
    m0, r0 = X[::7].mean(0) + 0.1, X[::7].std(0) * 1.3
    mean, rng = X.mean(0), X.std(0, ddof=1)

    def sums(m, r):
        Z = (X - m) / r
        S0, St, a, b, M = oracle.lagged_sums(Z, lag)
        t = torch.from_numpy
        return {"S0": t(S0), "St": t(St), "a": t(a), "b": t(b), "M": M}

    t = torch.from_numpy
    got = linalg.restandardize_sums(sums(m0, r0), t(m0), t(r0), t(mean), t(rng))
    ref = sums(mean, rng)
    for k in ("S0", "St", "a", "b"):
        np.testing.assert_allclose(got[k].numpy(), ref[k].numpy(), rtol=1e-10, atol=1e-9)


def test_tica_failure_raises_for_calculator_to_catch():
    from deep_cartograph_b200 import linalg
    S0 = -torch.eye(3, dtype=torch.float64)
    with pytest.raises(RuntimeError):
        linalg.tica_from_sums(S0, S0, torch.zeros(3, dtype=torch.float64), torch.zeros(3, dtype=torch.float64), 10, 2)


# ------------------------------------------------------------------------------------------------
# host-side mirrors
# ------------------------------------------------------------------------------------------------
def test_colvars_reader_semantics(tmp_path):
    from deep_cartograph_b200.modules.plumed.colvars import create_dataframe_from_files
    feats = [l.strip() for l in open(os.path.join(GOLDEN, "peptide_c1_features.txt")) if l.strip()]
    path = os.path.join(GOLDEN, "peptide_c1.dat")
    df = create_dataframe_from_files([path], features_list=feats, file_label="traj_label")
    assert df.shape == (164, 55) and list(df.columns[:-1]) == feats
    assert df[feats].to_numpy().dtype == np.float32
    # start / stop / stride slice EACH file before concatenation; files become one series
    df2 = create_dataframe_from_files([path, path], features_list=feats[::-1], file_label="traj_label",
                                      start=4, stop=100, stride=3)
    assert df2.shape == (64, 55) and list(df2.columns[:-1]) == feats[::-1]
    assert df2["traj_label"].tolist() == [0] * 32 + [1] * 32
    np.testing.assert_array_equal(df2[feats].to_numpy()[:32], df[feats].to_numpy()[4:100:3])
    with pytest.raises(ValueError):
        create_dataframe_from_files([path], features_list=feats + ["nope"])
    # 'time' is dropped when no features_list is given
    df3 = create_dataframe_from_files(path)
    assert "time" not in df3.columns and df3.shape == (164, 54)


def test_binary_sidecar_matches_text_reader(tmp_path):
    """the memory-mapped float32 sidecar gives exactly the matrix, column order and
    file labels of the text reader for every slice / selection, and goes stale with its source."""
    import shutil
    from deep_cartograph_b200.modules.plumed import colvars as cio
    feats = [l.strip() for l in open(os.path.join(GOLDEN, "peptide_c1_features.txt")) if l.strip()]
    a = str(tmp_path / "a.dat"); b = str(tmp_path / "b.dat")
    shutil.copy(os.path.join(GOLDEN, "peptide_c1.dat"), a)
    shutil.copy(os.path.join(GOLDEN, "peptide_c1.dat"), b)
    assert not cio.has_fresh_sidecar(a)
    for p in (a, b):
        cio.write_sidecar(p, chunk_rows=50)            # several chunks
        assert cio.has_fresh_sidecar(p)
    for kw in ({}, {"start": 4, "stop": 100, "stride": 3}, {"start": 10}, {"stop": 7}, {"stride": 5}):
        for fl in (None, feats, feats[::-1], feats[3:20:2]):
            df = cio.create_dataframe_from_files([a, b], features_list=fl, file_label="traj_label", **kw)
            lab = df.pop("traj_label").to_numpy()
            X, names, labels = cio.create_matrix_from_sidecars([a, b], features_list=fl, chunk_rows=17, **kw)
            assert names == df.columns.tolist()
            assert X.dtype == np.float32 and np.array_equal(X, df.to_numpy(dtype=np.float32))
            assert np.array_equal(labels, lab)
    with pytest.raises(ValueError):
        cio.create_matrix_from_sidecars([a], features_list=feats + ["nope"])
    with pytest.raises(ValueError):
        cio.create_matrix_from_sidecars([a], start=1000)
    # preallocated (pinned-style) destination
    buf = np.zeros((400, len(feats)), dtype=np.float32)
    X, _, _ = cio.create_matrix_from_sidecars([a, b], features_list=feats, out=buf)
    assert X.shape == (328, len(feats)) and np.shares_memory(X, buf)
    got = {}

    def alloc(rows, feats):
        got["buf"] = np.full((rows, feats), -1.0, dtype=np.float32)
        return got["buf"]

    X2, _, _ = cio.create_matrix_from_sidecars([a, b], features_list=feats, allocator=alloc)
    assert np.shares_memory(X2, got["buf"]) and np.array_equal(X2, X)
    # touching the text file invalidates the sidecar
    with open(a, "a") as fh:
        fh.write("\n")
    assert not cio.has_fresh_sidecar(a) and cio.has_fresh_sidecar(b)


def test_schemas_accept_reference_yaml_and_defaults():
    from deep_cartograph_b200.yaml_schemas.train_colvars import TrainColvarsSchema
    from deep_cartograph_b200.yaml_schemas.traj_cluster import TrajClusterSchema
    cfg = TrainColvarsSchema(**{"cvs": ["pca", "tica", "deep_tica", "htica", "ae", "vae"],
                                "common": {"dimension": 2, "lag_time": 1, "features_normalization": "mean_std",
                                           "architecture": {"encoder": {"layers": [16, 8]}},
                                           "training": {"general": {"num_tries": 1}}, "bias": {"method": "x"}},
                                "figures": {"fes": {"compute": True}}, "tica": {"lag_time": 3}}).model_dump()
    assert cfg["common"]["num_subspaces"] == 10 and cfg["common"]["subspaces_dimension"] == 5
    assert cfg["common"]["tica_regularization"] == 1e-6 and cfg["tica"] == {"lag_time": 3}
    assert cfg["common"]["backend"]["cov_engine"] == "auto"
    d = TrajClusterSchema().model_dump()
    assert d["algorithm"] == "hierarchical" and d["search_interval"] == [3, 10] and d["n_init"] == 20
    with pytest.raises(Exception):
        TrajClusterSchema(algorithm="spectral")


def test_merge_configurations_and_zip_roundtrip(tmp_path):
    from deep_cartograph_b200.modules.common import merge_configurations, unzip_files, zip_files
    merged = merge_configurations({"a": 1, "n": {"x": 1, "y": 2}}, {"n": {"y": 3}, "b": 4})
    assert merged == {"a": 1, "n": {"x": 1, "y": 3}, "b": 4}
    model = tmp_path / "model"
    model.mkdir()
    (model / "metadata.json").write_text("{}")
    np.save(model / "cv_weights.npy", np.eye(2, dtype=np.float32))
    zip_files(str(tmp_path / "model.zip"), str(model))
    import zipfile
    assert sorted(zipfile.ZipFile(tmp_path / "model.zip").namelist()) == ["model/cv_weights.npy", "model/metadata.json"]
    unzip_files(str(tmp_path / "model.zip"), str(tmp_path / "out"))
    assert (tmp_path / "out" / "model" / "cv_weights.npy").exists()


def test_calculator_loads_reference_model_zip_layout(tmp_path, c1):
    """model.zip written with the reference's member names loads (no GPU needed for load)."""
    import json
    import zipfile
    from deep_cartograph_b200.modules.cv_learning import CVCalculator, TICACalculator
    z = tmp_path / "tica_model.zip"
    with zipfile.ZipFile(z, "w") as zf:
        zf.writestr("model/metadata.json", json.dumps({"cv_name": "tica", "cv_dimension": 2}))
        zf.writestr("model/features_labels.txt", "\n".join(str(s) for s in c1["features"]) + "\n")
        for name in ("cv_weights", "cv_norm_mean", "cv_norm_range", "features_norm_mean", "features_norm_range"):
            import io
            buf = io.BytesIO()
            np.save(buf, c1[f"tica_{name}"])
            zf.writestr(f"model/{name}.npy", buf.getvalue())
    calc = CVCalculator.load(str(z), str(tmp_path / "out"))
    assert isinstance(calc, TICACalculator) and calc.get_cv_type() == "linear"
    assert calc.cv.shape == (54, 2) and calc.get_labels() == ["TIC 1", "TIC 2"]
    assert calc.num_features == 54 and calc.get_cv_parameters()["weights"] is calc.cv
    with pytest.raises(FileNotFoundError):
        CVCalculator.load(str(tmp_path / "missing.zip"), str(tmp_path))


def test_prepare_normalization_modes_match_oracle():
    from deep_cartograph_b200.modules.cv_learning import PCACalculator
    st = {"mean": np.array([1.0, 2.0], np.float32), "std": np.array([0.5, 0.0], np.float32),
          "min": np.array([0.0, 2.0], np.float32), "max": np.array([4.0, 2.0], np.float32)}
    for mode in (None, "mean_std", "min_max_range1", "min_max_range2"):
        calc = PCACalculator(configuration={"features_normalization": mode, "dimension": 2}, output_path=".")
        calc.features_stats = st
        m, r = calc.prepare_normalization()
        rm, rr = oracle.prepare_normalization(st, mode)
        np.testing.assert_array_equal(m, rm)
        np.testing.assert_array_equal(r, rr)
    calc = PCACalculator(configuration={"features_normalization": "bogus"}, output_path=".")
    calc.features_stats = st
    with pytest.raises(ValueError):
        calc.prepare_normalization()


def test_lloyd_driver_control_flow_with_emulated_kernel(monkeypatch, kmeans_ref):
    """The host Lloyd driver (centring, tolerance, strict convergence, final E-step, relocation)
    against the reference's own KMeans output, with the device E-step emulated in numpy."""
    from deep_cartograph_b200 import ops
    from deep_cartograph_b200.modules.statistics import statistics

    def fake_step(Y, C, labels, update_sums=True, want_gap=False, absmax=None):
        lab, best, second = oracle.kmeans_assign(Y.numpy(), C.numpy())
        old = labels.numpy().copy()
        labels.copy_(torch.from_numpy(lab))
        k, d = C.shape
        sums = np.zeros((k, d)); np.add.at(sums, lab, Y.numpy().astype(np.float64))
        stats = torch.tensor([float((old != lab).sum()), float((best + (Y.numpy() ** 2).sum(1)).sum()),
                              float((second - best <= 0).sum())], dtype=torch.float64)
        return {"sums": torch.from_numpy(sums), "counts": torch.from_numpy(np.bincount(lab, minlength=k).astype(np.float64)),
                "stats": stats, "gap": None}

    def fake_update(C, sums, counts, info=None):
        # dcg_kmeans_update: centres in place unless a cluster is empty; info = [n_empty, shift]
        n_empty = int((counts == 0).sum())
        shift = 0.0
        if n_empty == 0:
            C_new = sums * (1.0 / counts).unsqueeze(1)
            shift = float(((C_new - C) ** 2).sum())
            C.copy_(C_new)
        return torch.tensor([float(n_empty), shift], dtype=torch.float64)

    def fake_iterate(Y, C, labels, work, absmax=None):
        # dcg_kmeans_iterate: work = [sums | counts | stats 3 | info 2]
        k, d = C.shape
        r = fake_step(Y, C, labels)
        info = fake_update(C, r["sums"], r["counts"])
        work[:k * d] = r["sums"].reshape(-1)
        work[k * d:k * d + k] = r["counts"]
        work[k * d + k:k * d + k + 3] = r["stats"]
        work[k * d + k + 3:k * d + k + 5] = info
        return {"sums": work[:k * d].view(k, d), "counts": work[k * d:k * d + k],
                "stats": work[k * d + k:k * d + k + 3], "info": work[k * d + k + 3:k * d + k + 5]}

    def fake_iterate_n(Y, C, labels, work, iters, tol, absmax=None):
        # dcg_kmeans_iterate_n: up to `iters` iterations, the convergence tests taken "on the device";
        # work = [sums | counts | stats 3 | info 2 | ctl 3], ctl = [stopped, iterations done, tol]
        k, d = C.shape
        o = k * d + k
        work[o + 5:o + 8] = torch.tensor([0.0, 0.0, tol], dtype=torch.float64)
        r = None
        for _ in range(iters):
            r = fake_iterate(Y, C, labels, work)
            work[o + 6] += 1.0
            if float(work[o + 3]) > 0 or float(work[o]) == 0.0 or float(work[o + 4]) <= tol:
                work[o + 5] = 1.0
                break
        return dict(r, ctl=work[o + 5:o + 8])

    monkeypatch.setattr(ops, "kmeans_step", fake_step)
    monkeypatch.setattr(ops, "kmeans_update_", fake_update)
    monkeypatch.setattr(ops, "kmeans_iterate_", fake_iterate)
    monkeypatch.setattr(ops, "kmeans_iterate_n_", fake_iterate_n)
    monkeypatch.setattr(ops, "kmeans_work", lambda k, d, device: torch.zeros(k * d + k + 8, dtype=torch.float64))
    for name in ("blobs_d2_k5", "blobs_d4_k10_grid", "uniform_d3_k7_grid"):
        X = torch.from_numpy(kmeans_ref[f"{name}_X"])
        res = statistics.kmeans_lloyd(X, torch.from_numpy(kmeans_ref[f"{name}_init"]))
        ref = oracle.kmeans_lloyd(kmeans_ref[f"{name}_X"], kmeans_ref[f"{name}_init"])
        assert res["n_iter"] == ref["n_iter"] and res["strict"] == ref["strict"]
        assert np.array_equal(res["labels"].numpy(), ref["labels"])
        np.testing.assert_allclose(res["centers"].numpy(), ref["centers"], rtol=1e-10, atol=1e-12)
    # empty-cluster relocation path
    X = torch.tensor([[0.0, 0.0], [0.1, 0.0], [5.0, 5.0], [5.1, 5.0], [9.0, 9.0]], dtype=torch.float64)
    init = torch.tensor([[0.0, 0.0], [5.0, 5.0], [100.0, 100.0]], dtype=torch.float64)
    res = statistics.kmeans_lloyd(X, init)
    ref = oracle.kmeans_lloyd(X.numpy(), init.numpy())
    assert np.array_equal(res["labels"].numpy(), ref["labels"])
    np.testing.assert_allclose(res["centers"].numpy(), ref["centers"], rtol=1e-12)


def test_cluster_data_rejects_out_of_scope_algorithms(caplog):
    """hdbscan / hierarchical are outside the hot path: reported the reference's way (error logged,
    sys.exit(1)), not as a traceback -- this is what traj_cluster(configuration={}) runs into, whose
    schema default is hierarchical (reference yaml_schemas/traj_cluster.py:25)."""
    from deep_cartograph_b200.modules.statistics import statistics
    for algo in ("hdbscan", "hierarchical"):
        with pytest.raises(SystemExit) as ex:
            statistics.cluster_data(np.zeros((4, 2)), {"algorithm": algo})
        assert ex.value.code == 1
    with pytest.raises(SystemExit):
        statistics.optimize_clustering(np.zeros((4, 2)), {"algorithm": "hierarchical"})
    assert "outside the B200 hot path" in caplog.text


def test_deeptica_covariance_backward_formula_cpu(monkeypatch):
    """Analytic backward of the fused covariance op vs torch autograd (kernel emulated on CPU)."""
    from deep_cartograph_b200 import ops
    from deep_cartograph_b200.modules.cv_learning import deep_tica

    from conftest import cpu_ticacov_sums, cpu_ticaloss
    monkeypatch.setattr(ops, "ticacov_sums", cpu_ticacov_sums)
    monkeypatch.setattr(ops, "ticaloss", cpu_ticaloss)
    torch.manual_seed(0)
    B, d = 300, 3
    f = torch.randn(B, d, requires_grad=True)
    g = (0.6 * f.detach() + 0.5 * torch.randn(B, d)).requires_grad_(True)
    w = torch.rand(B) + 0.5
    # (i) covariance op + torch.linalg autograd, (ii) the fused eigen-loss with its analytic gradient
    loss_r, evals_r = deep_tica.tica_loss_reference(f, g, w, w, reg=1e-6)
    loss_r.backward()
    gf_r, gg_r = f.grad.clone(), g.grad.clone()
    f.grad = None; g.grad = None
    loss, evals = deep_tica.tica_loss(f, g, w, w, reg=1e-6)
    assert abs(loss.item() - loss_r.item()) < 1e-10
    np.testing.assert_allclose(evals.numpy(), evals_r.detach().numpy(), atol=1e-10)
    loss.backward()
    f2 = f.detach().double().requires_grad_(True)
    g2 = g.detach().double().requires_grad_(True)
    wn = w.double() / w.double().sum()
    mu = (wn[:, None] * f2).sum(0)
    a, b = f2 - mu, g2 - mu
    C0 = (wn[:, None] * a).T @ a
    C0 = 0.5 * (C0 + C0.T)
    Ct = (wn[:, None] * a).T @ b
    Ct = 0.5 * (Ct + Ct.T)
    ev = deep_tica.reduced_eigenvalues(C0, Ct, 1e-6)
    ref = -(ev ** 2).sum()
    ref.backward()
    rl, rev = oracle.deeptica_loss(f.detach().numpy(), g.detach().numpy(), w.numpy(), w.numpy(), reg=1e-6)
    assert abs(loss.item() - rl) < 1e-9 and abs(ref.item() - rl) < 1e-9
    np.testing.assert_allclose(f.grad.numpy(), f2.grad.numpy(), atol=1e-6)
    np.testing.assert_allclose(g.grad.numpy(), g2.grad.numpy(), atol=1e-6)
    np.testing.assert_allclose(gf_r.numpy(), f2.grad.numpy(), atol=1e-6)
    np.testing.assert_allclose(gg_r.numpy(), g2.grad.numpy(), atol=1e-6)


def test_block_triangular_inverse_matches_trsm():
    """linalg._tri_inv_lower (diagonal blocks inverted as a batch, joined bottom-up with
    [[A, 0], [B, C]]^-1 = [[A^-1, 0], [-C^-1 B A^-1, C^-1]]) against a plain triangular solve: sizes below the
    blocking threshold, multiples of the block count, sizes that need an identity pad, batches."""
    from deep_cartograph_b200 import linalg
    g = torch.Generator().manual_seed(3)
    for nb, F in ((1, 100), (1, 256), (2, 300), (3, 495), (1, 513), (1, 1000)):
        A = torch.randn(nb, F, F, generator=g, dtype=torch.float64)
        A = A @ A.mT + F * torch.eye(F, dtype=torch.float64)
        L = torch.linalg.cholesky(A)
        ref = torch.linalg.solve_triangular(L, torch.eye(F, dtype=torch.float64).expand(nb, F, F), upper=False)
        got = linalg._tri_inv_lower(L)
        assert got.shape == ref.shape
        assert float((got - ref).abs().max() / ref.abs().max()) < 1e-13, (nb, F)
        assert float(torch.triu(got, diagonal=1).abs().max()) == 0.0          # stays lower triangular
