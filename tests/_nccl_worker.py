"""torchrun worker of tests/test_multigpu_nccl.py: the frame-sharded hot path on N GPUs (NCCL) against
the same path on ONE GPU and against the float64 device checker.  Rank 0 holds the whole series too.

Checked: (1) merged column statistics; (2) in-place lag halo (P2P) + all-reduced S0 / St / a / b;
(3) TICA eigenvalues / eigenvectors / normalised projections of the sharded TICACalculator (resident
shards and shards streamed from pinned host memory, where only SOME ranks keep their speculative sums);
(4) hTICA on the block-diagonal path; (5) the sharded Lloyd driver: labels, centres, iteration count,
with an empty cluster relocated to a frame of another rank."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from deep_cartograph_b200 import linalg, ops
    from deep_cartograph_b200.modules.cv_learning.cv_calculator import HTICACalculator, TICACalculator
    from deep_cartograph_b200.modules.statistics import statistics
    from deep_cartograph_b200.parallel import FrameShards, shard_range
    from deep_cartograph_b200.synthetic import cluster_points, feature_matrix
    from oracle import float64_device as f64

    shards = FrameShards()
    n, f, lag, d = 120_001, 331, 7, 4                 # odd sizes: uneven shards, rows not 16-byte aligned
    s, e = shard_range(n, rank, world)
    Xloc = feature_matrix(n, f, s, e, dev, seed=3)
    out = {}

    # ---- TICA through the calculator, shards resident on the device
    cfg = {"dimension": d, "lag_time": lag, "features_normalization": "mean_std"}
    calc = TICACalculator(configuration=cfg, output_path=tempfile.mkdtemp())
    calc.load_training_tensor(Xloc, shards=shards)
    sums = calc._lagged_sums(lag)
    calc.create_output_folders(); calc.compute_cv(); calc.set_labels()
    P = calc.normalize_cv()
    # ---- the same from pinned host memory; odd ranks recompute, even ranks keep their speculative sums
    cfg2 = dict(cfg, backend={"speculative_sums": rank % 2 == 0, "h2d_chunk_bytes": 16 << 20})
    calc2 = TICACalculator(configuration=cfg2, output_path=tempfile.mkdtemp())
    calc2.load_training_tensor(Xloc.cpu().pin_memory(), shards=shards)
    kept = getattr(calc2, "_spec", None) is not None
    calc2.create_output_folders(); calc2.compute_cv()
    # ---- hTICA, block-diagonal level 1 + level-2 pass (the path C3 takes)
    cfg3 = dict(cfg, dimension=3, num_subspaces=6, subspaces_dimension=3, backend={"htica_full_gram_max_features": 0})
    calc3 = HTICACalculator(configuration=cfg3, output_path=tempfile.mkdtemp())
    calc3.load_training_tensor(Xloc, shards=shards)
    calc3.create_output_folders(); calc3.compute_cv()

    # ---- the peer-memory all-reduce (csrc/p2p.cu) against NCCL: sums and maxima, sizes from 1 element to a full
    #      slot, many exchanges in a row (both inbox parities, ranks running ahead of each other)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for it, nel in enumerate([1, 5, 403, 11_003, 16_384, 7, 7, 7, 7, 1000]):
        v = torch.randn(nel, generator=g, dtype=torch.float64, device=dev)
        for op, red in ((0, dist.ReduceOp.SUM), (1, dist.ReduceOp.MAX)):
            want = v.clone()
            dist.all_reduce(want, op=red)
            peer = shards._peer_allreduce(v)
            assert peer is not None, "peer-memory all-reduce unavailable on this box"
            got = peer.allreduce_(v.clone(), op)
            if op == 1:
                assert torch.equal(got, want), (it, nel)
            else:
                assert torch.allclose(got, want, rtol=1e-14, atol=1e-14), (it, nel)
            # every rank holds bitwise the same result
            chk = got.clone()
            dist.broadcast(chk, 0)
            assert torch.equal(chk, got), (it, nel, "ranks differ")
        if rank == it % world:
            torch.cuda._sleep(2_000_000)              # skew the ranks

    # ---- KMeans: same frames on every path (rank 0 draws them); cluster 5 starts empty
    k, dk, nk = 6, 3, 50_003
    Y = cluster_points(nk, dk, 5, dev, seed=11, dtype=torch.float64).round(decimals=4)
    Y[-1] = 3.0                                        # an outlier on the LAST rank: the relocation target
    dist.broadcast(Y, 0)
    init = torch.cat([Y[:5].clone(), torch.full((1, dk), 50.0, dtype=torch.float64, device=dev)])
    ks, ke = shard_range(nk, rank, world)
    km = statistics.kmeans_lloyd(Y[ks:ke].contiguous(), init, shards=shards)

    # gather the sharded results on rank 0
    def gather_rows(t):
        sizes = [shard_range(t_total, r, world) for r in range(world)]
        parts = [torch.empty((b - a,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) for a, b in sizes]
        dist.all_gather(parts, t.contiguous())
        return torch.cat(parts)
    t_total = n
    P_all = gather_rows(P)
    t_total = nk
    lab_all = gather_rows(km["labels"])
    kept_all = [None] * world
    dist.all_gather_object(kept_all, kept)

    if rank == 0:
        X = feature_matrix(n, f, 0, n, dev, seed=3)
        assert torch.equal(X[s:e], Xloc), "synthetic rows must not depend on the sharding"
        one = TICACalculator(configuration=cfg, output_path=tempfile.mkdtemp())
        one.load_training_tensor(X)
        s1 = one._lagged_sums(lag)
        one.create_output_folders(); one.compute_cv(); one.set_labels()
        P1 = one.normalize_cv()
        # (1) statistics
        for key in ("mean", "std", "min", "max"):
            a, b = torch.from_numpy(calc.features_stats[key]), torch.from_numpy(one.features_stats[key])
            assert torch.allclose(a, b, rtol=3e-7, atol=0), key
        assert calc.num_frames == n
        # (2) sums: sharded == single GPU up to the FP32 chunking of the contraction (each is within
        #     2e-6 of float64), both within 1e-5 of the float64 checker
        mean, rng = one._norm_on_device()
        ref = f64.lagged_sums(X, lag, mean, rng)
        scale = ref["S0"].abs().max()
        assert sums["M"] == s1["M"] == ref["M"] == n - lag
        # (the exact integer engine returns the SYMMETRIC part of St -- all that TICA uses, mlcolvar symmetrises
        #  C_tau -- so St is compared through its symmetric part whatever the engine)
        sym = lambda t: 0.5 * (t + t.T)
        for key in ("S0", "St"):
            assert ((sym(sums[key]) - sym(s1[key])).abs().max() / scale).item() < 4e-6, key
            assert ((sym(sums[key]) - sym(ref[key])).abs().max() / scale).item() < 1e-5, key
            assert ((sym(s1[key]) - sym(ref[key])).abs().max() / scale).item() < 1e-5, key
        for key in ("a", "b"):
            assert (sums[key] - ref[key]).abs().max().item() < 1e-6 * n, key
        # (3) TICA
        ev_ref, V_ref = f64.tica_from_sums(ref["S0"], ref["St"], ref["a"], ref["b"], ref["M"], d)
        Pn_ref, _, _ = f64.project_normalized(X, mean, rng, V_ref)
        for name, c in (("sharded", calc), ("sharded+streamed", calc2), ("single", one)):
            ev = torch.from_numpy(c.eigenvalues).to(dev)
            assert ((ev - ev_ref).abs() / ev_ref.abs()).max().item() < 1e-5, name
            err = f64.eigvec_error(torch.from_numpy(c.cv).to(dev), V_ref)
            out[f"evec_{name}"] = err
            assert err < 5e-5, (name, err)   # tightened to 1e-5 with the exact (integer) contraction engine
        assert (P_all.double() - Pn_ref).abs().max().item() < 1e-4
        assert (P_all - P1).abs().max().item() < 2e-5
        assert any(kept_all) and not all(kept_all), kept_all       # the mixed case really happened
        # (4) hTICA block path
        W_ref, _, _ = f64.htica(X, lag, mean, rng, 6, 3, 3)
        W = torch.from_numpy(calc3.cv).to(dev).double()
        sgn = torch.sign((W * W_ref).sum(0, keepdim=True))
        out["htica_W"] = float((W * sgn - W_ref).abs().max())
        assert out["htica_W"] < 5e-5
        # (5) KMeans
        km1 = statistics.kmeans_lloyd(Y, init)
        assert km["n_iter"] == km1["n_iter"]
        assert torch.equal(lab_all, km1["labels"])
        assert torch.allclose(km["centers"], km1["centers"], rtol=1e-12, atol=1e-12)
        assert int((km1["labels"] == 5).sum()) >= 1                # the empty cluster was relocated
        print("NCCL_WORKER_OK", world, out, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
