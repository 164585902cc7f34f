"""KMeans clustering in CV space on the B200 kernels.

Host-side mirror of the reference's ``modules/statistics/statistics.py`` for the KMeans path:
``cluster_data`` (:112-157), ``kmeans_clustering`` (:159-197), ``find_centroids`` (:337-379),
``optimize_clustering`` (:17-110, KMeans only, with the one-pass scores).  The Lloyd driver
reproduces scikit-learn's control flow (sklearn/cluster/_kmeans.py, ``algorithm='lloyd'``) so
that labels from fixed initial centroids are identical to the reference's:

  centre X by its column mean (:1486-1493); tol_eff = tol * mean(var(X, axis=0)) (:285-294);
  per iteration E-step (lowest index wins ties) + FP64 sums / counts; empty clusters relocated
  to the farthest points (_k_means_common.pyx:167-211); centres = sums * (1 / counts);
  stop when labels are unchanged ("strict") or sum ||dc||^2 <= tol_eff (:703-740); one more
  E-step when not strict (:742-754); centres returned un-centred.

HDBSCAN and agglomerative clustering are out of scope and raise.
"""
from __future__ import annotations

import logging
import sys
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from ... import ops
from ...parallel import FrameShards

logger = logging.getLogger(__name__)

# set by the last kmeans_clustering call: iterations, strict convergence, tie counts ...
last_kmeans_report: Dict = {}


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("deep_cartograph_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_device(features) -> torch.Tensor:
    if isinstance(features, torch.Tensor):
        t = features
    else:
        t = torch.from_numpy(np.ascontiguousarray(features))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    return t.to(_device()).contiguous()


def _kmeanspp_sklearn(Yc: torch.Tensor, k: int, rs: np.random.RandomState) -> torch.Tensor:
    """scikit-learn's greedy k-means++ (sklearn/cluster/_kmeans.py ``_kmeans_plusplus``, reached from
    reference statistics.py:189-195 with ``init='k-means++'``, ``random_state=0``) on the CENTRED frames
    ``Yc`` (KMeans.fit centres X first, _kmeans.py:1486-1493): the random draws come from the SAME numpy
    ``RandomState`` stream (first centre: ``choice(n, p=uniform)``; then ``2 + int(log k)`` uniform
    draws per centre against the cumulative D^2 potential), the O(n k) distance work runs on the device
    in FP64.  The centres equal scikit-learn's unless a draw lands within rounding of a cumulative-sum
    boundary."""
    n, d = Yc.shape
    Yd = Yc.to(torch.float64)
    n_trials = 2 + int(np.log(k))
    centers = torch.empty((k, d), dtype=torch.float64, device=Yc.device)
    if n <= 5_000_000:
        first = int(rs.choice(n, p=np.full(n, 1.0 / n)))
    else:      # the same draw without the n-vector: choice() searches u in cdf_i = (i + 1) / n -> floor(u n)
        first = min(n - 1, int(np.floor(rs.random_sample() * n)))
    centers[0] = Yd[first]
    closest = ((Yd - centers[0]) ** 2).sum(dim=1)
    pot = float(closest.sum().item())
    for c in range(1, k):
        rand_vals = torch.from_numpy(rs.uniform(size=n_trials) * pot).to(Yc.device)
        cand = torch.searchsorted(torch.cumsum(closest, 0), rand_vals).clamp_(max=n - 1)
        dist = torch.stack([((Yd - Yd[i]) ** 2).sum(dim=1) for i in cand])
        dist = torch.minimum(dist, closest[None, :])
        pots = dist.sum(dim=1)
        best = int(torch.argmin(pots).item())
        pot = float(pots[best].item())
        closest = dist[best]
        centers[c] = Yd[cand[best]]
    return centers


# Lloyd iterations enqueued per host read on one device (dcg_kmeans_iterate_n)
_LLOYD_BATCH = 4


def kmeans_lloyd(Y: torch.Tensor, init_centers: torch.Tensor, max_iter: int = 300, tol: float = 1e-4,
                 shards: Optional[FrameShards] = None, want_gap: bool = False) -> Dict:
    """Lloyd iterations on the device; ``Y`` is this rank's (frames x d) shard (float32/64),
    ``init_centers`` (k x d).  Returns dict(labels, centers, n_iter, strict, inertia, ties, gap)."""
    dev = Y.device
    n, d = Y.shape
    k = init_centers.shape[0]
    C = init_centers.to(dev, torch.float64).clone()

    # centre by the (global) column mean; tolerance from the (global) column variances
    st_n = torch.tensor([float(n)], dtype=torch.float64, device=dev)
    s1 = Y.sum(dim=0, dtype=torch.float64)
    if shards is not None:
        packed = shards.allreduce_sum_(torch.cat([st_n, s1]))
        st_n, s1 = packed[:1], packed[1:]
    n_tot = float(st_n.item())
    x_mean = s1 / n_tot
    Yc = (Y - x_mean.to(Y.dtype)).contiguous()
    s2 = (Yc.to(torch.float64) ** 2).sum(dim=0)
    sc = Yc.sum(dim=0, dtype=torch.float64)
    if shards is not None:
        packed = shards.allreduce_sum_(torch.cat([s2, sc]))
        s2, sc = packed[:d], packed[d:]
    var = s2 / n_tot - (sc / n_tot) ** 2
    tol_eff = float(var.mean().item()) * tol
    C = C - x_mean
    # bound on |Yc| of this shard: lets the E-step accumulate its partial sums in exact 64-bit
    # fixed point (native shared-memory integer adds) instead of FP64 compare-and-swap loops
    absmax = Yc.abs().amax().to(torch.float64).reshape(1) if n > 0 else None

    labels = torch.full((n,), -1, dtype=torch.int32, device=dev)
    strict = False
    n_iter = 0
    res = None
    C = C.contiguous()
    work = ops.kmeans_work(k, d, dev)
    it = 0
    while it < max_iter:
        if shards is None:
            # single device: up to _LLOYD_BATCH whole iterations (zero + E-step + FP64 sums + M-step finish each)
            # per library call; the convergence tests run on the device and the launches after the stopping
            # iteration are no-ops, so the state is the one a one-at-a-time loop would have stopped in
            res = ops.kmeans_iterate_n_(Yc, C, labels, work, min(_LLOYD_BATCH, max_iter - it), tol_eff, absmax=absmax)
            sums, counts = res["sums"], res["counts"]
            changed, _, _, n_empty, shift_tot, _, done, _ = work[k * d + k:k * d + k + 8].tolist()   # ONE host read
            it += int(done)
        else:
            # [sums | counts | stats] land in one buffer and are all-reduced in place
            res = ops.kmeans_step_packed_(Yc, C, labels, work, absmax=absmax)
            shards.allreduce_sum_(res["packed"])
            sums, counts = res["sums"], res["counts"]
            # M-step finish on the device (centres updated in place unless a cluster is empty);
            # ONE host read per iteration: [changed, inertia, ties, n_empty, shift]
            ops.kmeans_update_(C, sums, counts, info=work[k * d + k + 3:k * d + k + 5])
            changed, _, _, n_empty, shift_tot = work[k * d + k:k * d + k + 5].tolist()
            it += 1
        if n_empty > 0:
            empty = torch.nonzero(counts == 0).flatten()
            sums, counts = _relocate_empty(Yc, C, labels, sums, counts, empty, shards)
            C_new = C.clone()
            nz = counts > 0
            C_new[nz] = sums[nz] * (1.0 / counts[nz]).unsqueeze(1)
            shift_tot = float(((C_new - C) ** 2).sum().item())
            C = C_new
        n_iter = it
        if changed == 0:
            strict = True
            break
        if shift_tot <= tol_eff:
            break
    # final E-step with the final centres when not strictly converged (labels consistent with C)
    if not strict:
        res = ops.kmeans_step(Yc, C, labels, update_sums=False, want_gap=want_gap)
    elif want_gap:
        res = ops.kmeans_step(Yc, C, labels, update_sums=False, want_gap=True)
    stats = res["stats"]
    if shards is not None:
        stats = shards.allreduce_sum_(stats.clone())
    return {"labels": labels, "centers": C + x_mean, "n_iter": n_iter, "strict": strict,
            "inertia": float(stats[1].item()), "ties": int(stats[2].item()), "gap": res.get("gap"),
            "tol_eff": tol_eff}


def _relocate_empty(Yc, C, labels, sums, counts, empty, shards):
    """sklearn _relocate_empty_clusters_dense (sklearn/cluster/_k_means_common.pyx:167-211, reached from
    reference statistics.py:189-195): the n_empty frames farthest from their centres become the new
    centres of the empty clusters (in index order of the empty clusters), and leave their old ones.
    Labels are left untouched, as in sklearn.  Rare path, torch ops on the device.

    Frame-sharded: ``sums`` / ``counts`` are the already all-reduced (replicated) global ones; every
    rank offers its n_empty farthest frames (distance, old label, coordinates), the candidates are
    all-gathered, and every rank applies the same global top-n_empty -- ties go to the lower rank,
    then the lower local index, i.e. the lower GLOBAL frame index of the contiguous shards."""
    n_empty = int(empty.numel())
    n, d = Yc.shape
    dev = Yc.device
    dist = ((Yc.to(torch.float64) - C[labels.long()]) ** 2).sum(dim=1)
    m = min(n_empty, n)
    top = torch.topk(dist, m, largest=True, sorted=True) if m > 0 else None
    # candidates: [distance | old label | coordinates], best first; stable for equal distances
    cand = torch.full((n_empty, 2 + d), float("-inf"), dtype=torch.float64, device=dev)
    if m > 0:
        far = top.indices
        order = torch.argsort(far)                                  # equal distances: lower index first
        far = far[order][torch.argsort(-dist[far[order]], stable=True)]
        cand[:m, 0] = dist[far]
        cand[:m, 1] = labels[far].to(torch.float64)
        cand[:m, 2:] = Yc[far].to(torch.float64)
    if shards is not None and shards.world > 1:
        allc = torch.empty((shards.world * n_empty, 2 + d), dtype=torch.float64, device=dev)
        torch.distributed.all_gather_into_tensor(allc, cand.contiguous(), group=shards.group)
        # rank-major order + stable sort = ties to the lower global frame index
        pick = torch.argsort(-allc[:, 0], stable=True)[:n_empty]
        cand = allc[pick]
    sums = sums.clone()
    counts = counts.clone()
    rows = cand.tolist()
    for idx in range(n_empty):
        new_id = int(empty[idx].item())
        if rows[idx][0] == float("-inf"):
            break                                                   # fewer frames than empty clusters
        old_id = int(rows[idx][1])
        x = cand[idx, 2:]
        sums[old_id] -= x
        sums[new_id] = x
        counts[new_id] = 1.0
        counts[old_id] -= 1.0
    return sums, counts


def kmeans_clustering(feature_matrix, num_clusters: int, n_init: int,
                      initial_centroids=None) -> Tuple[np.ndarray, np.ndarray]:
    """Reference statistics.py:159-197: ``KMeans(n_clusters, random_state=0, init, n_init).fit_predict``.
    With ``initial_centroids`` the number of clusters comes from their shape and a single run is made (as
    sklearn does for an ndarray init); otherwise ``n_init`` k-means++ seedings drawn from ONE
    ``RandomState(0)`` stream as sklearn draws them, and the run with the lowest inertia wins (a later
    run replaces the best only if its inertia is lower and its clustering differs, _kmeans.py:1516-1530)."""
    global last_kmeans_report
    logger.debug("Clustering frames with kmeans...")
    Y = _to_device(feature_matrix)
    runs = []
    if initial_centroids is not None:
        init = torch.as_tensor(np.asarray(initial_centroids), dtype=torch.float64)
        num_clusters = init.shape[0]
        runs.append(kmeans_lloyd(Y, init))
    else:
        rs = np.random.RandomState(0)
        x_mean = Y.to(torch.float64).mean(dim=0)
        Yc = (Y - x_mean.to(Y.dtype)).contiguous()
        for _ in range(max(1, int(n_init))):
            runs.append(kmeans_lloyd(Y, _kmeanspp_sklearn(Yc, num_clusters, rs) + x_mean))
    logger.debug("Number of clusters: {}".format(num_clusters))
    best = runs[0]
    for r in runs[1:]:
        if r["inertia"] < best["inertia"]:
            best = r
    last_kmeans_report = {"n_iter": best["n_iter"], "strict": best["strict"], "ties": best["ties"],
                          "inertia": best["inertia"], "n_runs": len(runs)}
    if best["ties"]:
        logger.info(f"KMeans: {best['ties']} frames sit on an exact tie between two centres "
                    "(lowest centre index kept, as in scikit-learn)")
    return best["labels"].cpu().numpy(), best["centers"].cpu().numpy()


def cluster_data(features, settings: Dict, initial_centroids=None) -> Tuple[np.ndarray, np.ndarray]:
    """Reference statistics.py:112-157 (defaults filled in the same way)."""
    settings["algorithm"] = settings.get("algorithm", "kmeans")
    settings["num_clusters"] = settings.get("num_clusters", 10)
    settings["n_init"] = settings.get("n_init", 10)
    if settings["algorithm"] == "kmeans":
        return kmeans_clustering(features, settings["num_clusters"], settings["n_init"], initial_centroids)
    if settings["algorithm"] in ("hdbscan", "hierarchical"):
        _outside_hot_path(settings["algorithm"])
    raise Exception(f"clustering algorithm {settings['algorithm']} not implemented")


def cluster_scores(features, labels, centers=None) -> Dict[str, float]:
    """Calinski-Harabasz and Davies-Bouldin scores (sklearn definitions, used at reference
    statistics.py:73-74): counts and means of the members from FP64 reductions, the per-cluster
    dispersion sums from one pass of ``dcg_cluster_dispersion``."""
    Y = _to_device(features)
    lab = torch.as_tensor(np.asarray(labels), device=Y.device).long()
    n, d = Y.shape
    ids, inv = torch.unique(lab, return_inverse=True)
    k = ids.numel()
    Yd = Y.to(torch.float64)
    cnt = torch.zeros(k, dtype=torch.float64, device=Y.device).index_add_(0, inv, torch.ones(n, dtype=torch.float64, device=Y.device))
    cent = torch.zeros((k, d), dtype=torch.float64, device=Y.device).index_add_(0, inv, Yd) / cnt[:, None]
    ssq, sdist = ops.cluster_dispersion(Y, inv.to(torch.int32), cent)
    intra_disp = float(ssq.sum().item())
    mean = Yd.mean(dim=0)
    extra_disp = float((cnt * ((cent - mean) ** 2).sum(dim=1)).sum().item())
    ch = 1.0 if intra_disp == 0.0 else extra_disp * (n - k) / (intra_disp * (k - 1.0))
    sk = sdist / cnt
    D = torch.cdist(cent, cent)
    if torch.allclose(sk, torch.zeros_like(sk)) or torch.allclose(D, torch.zeros_like(D)):
        db = 0.0
    else:
        D = D.masked_fill(D == 0, float("inf"))
        db = float(((sk[:, None] + sk[None, :]) / D).max(dim=1).values.mean().item())
    return {"calinski_harabasz": float(ch), "davies_bouldin": db}


def silhouette(features, labels, sample_size: Optional[int] = None, seed: int = 0, chunk: int = 2048) -> float:
    """Mean silhouette coefficient (sklearn ``silhouette_score``, Euclidean; reference statistics.py:75)
    on the device in FP64: for every frame i, a_i = mean distance to the other members of its cluster,
    b_i = the smallest mean distance to the members of another cluster, s_i = (b_i - a_i) / max(a_i, b_i)
    (0 for singleton clusters).  O(N^2 d): ``sample_size`` evaluates it on a uniform random subsample
    of that many frames against each other -- sklearn's own ``sample_size`` estimator."""
    Y = _to_device(features).to(torch.float64)
    lab = torch.as_tensor(np.asarray(labels), device=Y.device).long()
    n = Y.shape[0]
    if sample_size is not None and sample_size < n:
        idx = torch.from_numpy(np.random.RandomState(seed).permutation(n)[:sample_size]).to(Y.device)
        Y, lab = Y[idx], lab[idx]
        n = sample_size
    ids, inv = torch.unique(lab, return_inverse=True)
    k = ids.numel()
    if not 2 <= k <= n - 1:
        raise ValueError("Number of labels is %d. Valid values are 2 to n_samples - 1 (inclusive)" % k)
    onehot = torch.zeros((n, k), dtype=torch.float64, device=Y.device)
    onehot[torch.arange(n, device=Y.device), inv] = 1.0
    cnt = onehot.sum(dim=0)
    total = torch.zeros((), dtype=torch.float64, device=Y.device)
    for s0 in range(0, n, chunk):
        e0 = min(n, s0 + chunk)
        Dc = torch.cdist(Y[s0:e0], Y) @ onehot                          # (chunk, k): summed distances to every cluster
        own = inv[s0:e0]
        rows = torch.arange(e0 - s0, device=Y.device)
        n_own = cnt[own]
        a = Dc[rows, own] / (n_own - 1).clamp(min=1.0)
        Dm = Dc / cnt[None, :]
        Dm[rows, own] = float("inf")
        b = Dm.min(dim=1).values
        sil = (b - a) / torch.maximum(a, b)
        sil = torch.where(n_own > 1, sil, torch.zeros_like(sil))
        total += torch.nan_to_num(sil).sum()
    return float((total / n).item())


def optimize_clustering(features, settings: Dict):
    """Reference statistics.py:17-110 for KMeans: every k of ``search_interval`` is clustered and scored
    with Calinski-Harabasz, Davies-Bouldin and the mean silhouette; each score is min-max normalised over
    the interval and the best k maximises (CH - DB + silhouette) / 3.  Like the reference it mutates
    ``settings['num_clusters']`` and ignores ``opt_num_clusters``.  The silhouette is exact up to
    ``settings['silhouette_max_samples']`` frames (default 50000, O(N^2) on the device); beyond that it is
    sklearn's ``sample_size`` estimator on that many frames (seed 0) and a warning says so."""
    if settings.get("algorithm", "kmeans") != "kmeans":
        _outside_hot_path(settings.get("algorithm"))
    lo, hi = settings.get("search_interval", [2, 15])
    ks = list(range(lo, hi + 1))
    feats = np.asarray(features)
    max_sil = int(settings.get("silhouette_max_samples", 50000))
    sample = None
    if feats.shape[0] > max_sil:
        sample = max_sil
        logger.warning(f"Silhouette score estimated on a random subsample of {max_sil} of {feats.shape[0]} frames "
                       "(scikit-learn's sample_size estimator); the reference evaluates all pairs")
    ch, db, sil, results = [], [], [], []
    for k in ks:
        settings["num_clusters"] = k
        labels, centroids = cluster_data(feats, settings)
        sc = cluster_scores(feats, labels)
        ch.append(sc["calinski_harabasz"])
        db.append(sc["davies_bouldin"])
        sil.append(silhouette(feats, labels, sample_size=sample))
        logger.debug(f"Calinski-Harabasz score: {round(ch[-1], 3)}")
        logger.debug(f"Davies-Bouldin score: {round(db[-1], 3)}")
        logger.debug(f"Average silhouette score: {round(sil[-1], 3)}")
        results.append((labels, centroids))

    def norm(v):
        v = np.asarray(v, dtype=np.float64)
        return (v - v.min()) / (v.max() - v.min())
    score = (norm(ch) - norm(db) + norm(sil)) / 3
    best = int(np.argmax(score))
    logger.info(f"Best number of clusters: {ks[best]}")
    labels, centroids = results[best]
    if len(centroids) == 0:
        logger.warning("No clusters found using the provided settings. Try different settings or a different algorithm")
    global last_optimize_report
    last_optimize_report = {"k": ks, "calinski_harabasz": ch, "davies_bouldin": db, "silhouette": sil,
                            "score": score.tolist(), "best_k": ks[best]}
    return labels, centroids


last_optimize_report: Dict = {}


def _outside_hot_path(algorithm):
    """hdbscan / hierarchical clustering are not part of the B200 hot path: say so
    the way the reference reports fatal configuration problems -- log an error and exit -- instead of a
    traceback."""
    logger.error(f"Clustering algorithm '{algorithm}' is outside the B200 hot path (only 'kmeans' is accelerated); "
                 "set traj_cluster.algorithm: kmeans or use the reference package for it. Exiting...")
    sys.exit(1)


def find_centroids(data: pd.DataFrame, centroids: np.ndarray, clustering_features: List[str]) -> pd.DataFrame:
    """Mark the sample closest to each centroid (reference statistics.py:337-379) in ONE pass
    over the data instead of k."""
    if len(centroids) == 0:
        logger.warning("No centroids found")
        return pd.DataFrame()
    if len(centroids[0]) != len(clustering_features):
        logger.error("  The dimension of the centroids is not the same as the dimension of the used features for clustering.\n")
        sys.exit(1)
    Y = _to_device(data.loc[:, clustering_features].to_numpy())
    C = torch.as_tensor(np.asarray(centroids), dtype=torch.float64, device=Y.device)
    idx = ops.nearest_to_centers(Y, C).cpu().numpy()
    data["centroid"] = False
    data.loc[data.index[idx], "centroid"] = True
    return data
