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


def _kmeanspp_init(Y: torch.Tensor, k: int, seed: int) -> torch.Tensor:
    """k-means++ seeding (greedy variant as in sklearn _kmeans_plusplus) on the device with a
    seeded torch generator.  NOT bit-identical to numpy's RandomState stream: parity with the
    reference is defined for fixed initial centroids (BASELINE.json north_star)."""
    n, d = Y.shape
    gen = torch.Generator(device=Y.device)
    gen.manual_seed(seed)
    Yd = Y.to(torch.float64)
    n_trials = 2 + int(np.log(k))
    centers = torch.empty((k, d), dtype=torch.float64, device=Y.device)
    first = int(torch.randint(n, (1,), generator=gen, device=Y.device).item())
    centers[0] = Yd[first]
    closest = ((Yd - centers[0]) ** 2).sum(dim=1)
    for c in range(1, k):
        pot = closest.sum()
        r = torch.rand(n_trials, generator=gen, device=Y.device, dtype=torch.float64) * pot
        cand = torch.searchsorted(torch.cumsum(closest, 0), r).clamp_(max=n - 1)
        dist = ((Yd[cand][:, None, :] - Yd[None, :, :]) ** 2).sum(dim=2) if n * n_trials * d < 5e8 else \
            torch.stack([((Yd - Yd[i]) ** 2).sum(dim=1) for i in cand])
        dist = torch.minimum(dist, closest[None, :])
        best = int(torch.argmin(dist.sum(dim=1)).item())
        centers[c] = Yd[cand[best]]
        closest = dist[best]
    return centers


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
    for it in range(max_iter):
        if shards is None:
            # single device: memset + E-step + FP64 sums + M-step finish in one library call
            res = ops.kmeans_iterate_(Yc, C, labels, work, absmax=absmax)
            sums, counts = res["sums"], res["counts"]
            changed, _, _, n_empty, shift_tot = work[k * d + k:k * d + k + 5].tolist()   # ONE host read
        else:
            # [sums | counts | stats] land in one buffer and are all-reduced in place
            res = ops.kmeans_step_packed_(Yc, C, labels, work, absmax=absmax)
            shards.allreduce_sum_(res["packed"])
            sums, counts = res["sums"], res["counts"]
            # M-step finish on the device (centres updated in place unless a cluster is empty);
            # ONE host read per iteration: [changed, inertia, ties, n_empty, shift]
            ops.kmeans_update_(C, sums, counts, info=work[k * d + k + 3:k * d + k + 5])
            changed, _, _, n_empty, shift_tot = work[k * d + k:k * d + k + 5].tolist()
        if n_empty > 0:
            empty = torch.nonzero(counts == 0).flatten()
            sums, counts = _relocate_empty(Yc, C, labels, sums, counts, empty, shards)
            C_new = C.clone()
            nz = counts > 0
            C_new[nz] = sums[nz] * (1.0 / counts[nz]).unsqueeze(1)
            shift_tot = float(((C_new - C) ** 2).sum().item())
            C = C_new
        n_iter = it + 1
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
    """Reference statistics.py:159-197.  With ``initial_centroids`` the number of clusters comes
    from their shape and a single run is made (as sklearn does for an ndarray init); otherwise
    ``n_init`` k-means++ seedings (seed 0, 1, ...) and the lowest inertia wins."""
    global last_kmeans_report
    logger.debug("Clustering frames with kmeans...")
    Y = _to_device(feature_matrix)
    runs = []
    if initial_centroids is not None:
        init = torch.as_tensor(np.asarray(initial_centroids), dtype=torch.float64)
        num_clusters = init.shape[0]
        runs.append(kmeans_lloyd(Y, init))
    else:
        for s in range(max(1, int(n_init))):
            runs.append(kmeans_lloyd(Y, _kmeanspp_init(Y, num_clusters, seed=s)))
    logger.debug("Number of clusters: {}".format(num_clusters))
    best = min(runs, key=lambda r: r["inertia"])
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
        raise NotImplementedError(
            f"clustering algorithm {settings['algorithm']} is outside the B200 hot path "
            "(SURVEY.md section 2); use algorithm: kmeans or the reference package")
    raise Exception(f"clustering algorithm {settings['algorithm']} not implemented")


def cluster_scores(features, labels, centers=None) -> Dict[str, float]:
    """Calinski-Harabasz and Davies-Bouldin scores (sklearn definitions, used at reference
    statistics.py:73-74) from per-cluster sums computed on the device in FP64."""
    Y = _to_device(features).to(torch.float64)
    lab = torch.as_tensor(np.asarray(labels), device=Y.device).long()
    n, d = Y.shape
    ids, inv = torch.unique(lab, return_inverse=True)
    k = ids.numel()
    cnt = torch.zeros(k, dtype=torch.float64, device=Y.device).index_add_(0, inv, torch.ones(n, dtype=torch.float64, device=Y.device))
    cent = torch.zeros((k, d), dtype=torch.float64, device=Y.device).index_add_(0, inv, Y) / cnt[:, None]
    diff = Y - cent[inv]
    sq = (diff ** 2).sum(dim=1)
    intra_disp = float(sq.sum().item())
    mean = Y.mean(dim=0)
    extra_disp = float((cnt * ((cent - mean) ** 2).sum(dim=1)).sum().item())
    ch = 1.0 if intra_disp == 0.0 else extra_disp * (n - k) / (intra_disp * (k - 1.0))
    s = torch.zeros(k, dtype=torch.float64, device=Y.device).index_add_(0, inv, sq.sqrt()) / cnt
    D = torch.cdist(cent, cent)
    if torch.allclose(s, torch.zeros_like(s)) or torch.allclose(D, torch.zeros_like(D)):
        db = 0.0
    else:
        D = D.masked_fill(D == 0, float("inf"))
        db = float(((s[:, None] + s[None, :]) / D).max(dim=1).values.mean().item())
    return {"calinski_harabasz": float(ch), "davies_bouldin": db}


def optimize_clustering(features, settings: Dict):
    """Reference statistics.py:17-110 for KMeans.  Scans ``search_interval`` and picks the best k
    by the min-max normalised (CH - DB [+ silhouette]) score.  The O(N^2) silhouette term is
    included (scikit-learn, on the host) only when N <= settings['silhouette_max_samples']
    (default 20000); beyond that it is dropped and the combination is (CH - DB) / 2."""
    if settings.get("algorithm", "kmeans") != "kmeans":
        raise NotImplementedError("only kmeans is accelerated; see cluster_data")
    lo, hi = settings.get("search_interval", [2, 15])
    ks = list(range(lo, hi + 1))
    feats = np.asarray(features)
    use_sil = feats.shape[0] <= int(settings.get("silhouette_max_samples", 20000))
    ch, db, sil, results = [], [], [], []
    for k in ks:
        settings["num_clusters"] = k
        labels, centroids = cluster_data(feats, settings)
        sc = cluster_scores(feats, labels)
        ch.append(sc["calinski_harabasz"])
        db.append(sc["davies_bouldin"])
        if use_sil:
            from sklearn.metrics import silhouette_score
            sil.append(silhouette_score(feats, labels))
        results.append((labels, centroids))

    def norm(v):
        v = np.asarray(v, dtype=np.float64)
        return (v - v.min()) / (v.max() - v.min())
    score = (norm(ch) - norm(db) + norm(sil)) / 3 if use_sil else (norm(ch) - norm(db)) / 2
    best = int(np.argmax(score))
    logger.info(f"Best number of clusters: {ks[best]}")
    labels, centroids = results[best]
    if len(centroids) == 0:
        logger.warning("No clusters found using the provided settings. Try different settings or a different algorithm")
    return labels, centroids


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
