"""world_size-2 gloo tests (CPU) of the frame-sharding collectives: statistics merge,
lag halo, partial-sum all-reduce.  The per-rank kernels are emulated with the oracle so that only
the host-side sharding logic is under test here; the device kernels are covered by -m gpu."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, f, lag, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from conftest import synth_features
        from deep_cartograph_b200 import linalg
        from deep_cartograph_b200.parallel import FrameShards, shard_range
        X = synth_features(n, f, seed=1)
        s, e = shard_range(n, rank, world)
        shards = FrameShards()
        # ---- statistics: per-rank (n, mean, M2, min, max) -> Chan merge
        loc = oracle.column_stats(X[s:e])
        st = {"n": e - s, "mean": torch.from_numpy(loc["mean"]),
              "m2": torch.from_numpy(loc["std"] ** 2 * (e - s - 1)),
              "min": torch.from_numpy(loc["min"].astype(np.float32)),
              "max": torch.from_numpy(loc["max"].astype(np.float32))}
        merged = shards.merge_stats(st)
        ref = oracle.column_stats(X)
        assert merged["n"] == n
        np.testing.assert_allclose(merged["mean"].numpy(), ref["mean"], rtol=1e-12)
        np.testing.assert_allclose(np.sqrt(merged["m2"].numpy() / (n - 1)), ref["std"], rtol=1e-10)
        np.testing.assert_array_equal(merged["min"].numpy(), ref["min"].astype(np.float32))
        # ---- halo: received in place when the shard is a view of a buffer with spare rows
        mean, rng = oracle.prepare_normalization(ref, "mean_std")
        Z = oracle.standardize(X, mean, rng)
        buf = torch.zeros((e - s + lag, f), dtype=torch.float32)
        buf[:e - s] = torch.from_numpy(Z[s:e])
        Zh = shards.with_halo(buf[:e - s], lag)
        if rank + 1 < world:
            assert Zh.shape[0] == e - s + lag and Zh.data_ptr() == buf.data_ptr()
            np.testing.assert_array_equal(Zh[e - s:].numpy(), Z[e:e + lag])
        else:
            assert Zh.shape[0] == e - s
        # copy path (no spare rows)
        Zh2 = shards.with_halo(torch.from_numpy(Z[s:e].copy()), lag)
        np.testing.assert_array_equal(Zh2.numpy(), Zh.numpy())
        # ---- partial sums -> one fused FP64 all-reduce -> same TICA as the unsharded oracle
        S0, St, a, b, M = oracle.lagged_sums(Zh.numpy(), lag)
        tot = shards.allreduce_sums({"S0": torch.from_numpy(S0), "St": torch.from_numpy(St),
                                     "a": torch.from_numpy(a), "b": torch.from_numpy(b), "M": M})
        rS0, rSt, ra, rb, rM = oracle.lagged_sums(Z, lag)
        assert tot["M"] == rM == n - lag
        np.testing.assert_allclose(tot["S0"].numpy(), rS0, rtol=1e-11, atol=1e-8)
        np.testing.assert_allclose(tot["St"].numpy(), rSt, rtol=1e-11, atol=1e-8)
        # the same with the sums laid out as ops.lagged_covariance lays them out -- one flat buffer
        # [S0 | St | a | b | M], reduced in place -- and with a dict whose S0 no longer IS the slice of the
        # buffer (replaced by the caller, e.g. re-standardised speculative sums): packed as before
        flat = torch.cat([torch.from_numpy(S0).reshape(-1), torch.from_numpy(St).reshape(-1), torch.from_numpy(a),
                          torch.from_numpy(b), torch.tensor([float(M)], dtype=torch.float64)])
        views = {"S0": flat[:f * f].view(f, f), "St": flat[f * f:2 * f * f].view(f, f),
                 "a": flat[2 * f * f:2 * f * f + f], "b": flat[2 * f * f + f:2 * f * f + 2 * f], "M": M, "flat": flat}
        tot2 = shards.allreduce_sums(dict(views), m_total=n - lag)
        assert tot2["S0"].data_ptr() == flat.data_ptr() and tot2["M"] == n - lag
        for key in ("S0", "St", "a", "b"):
            np.testing.assert_array_equal(tot2[key].numpy(), tot[key].numpy())
        flat3 = torch.cat([torch.from_numpy(S0).reshape(-1), torch.from_numpy(St).reshape(-1), torch.from_numpy(a),
                           torch.from_numpy(b), torch.tensor([float(M)], dtype=torch.float64)])
        stale = {"S0": torch.from_numpy(S0).clone(), "St": flat3[f * f:2 * f * f].view(f, f),
                 "a": flat3[2 * f * f:2 * f * f + f], "b": flat3[2 * f * f + f:2 * f * f + 2 * f], "M": M, "flat": flat3}
        tot3 = shards.allreduce_sums(stale)
        for key in ("S0", "St", "a", "b"):
            np.testing.assert_array_equal(tot3[key].numpy(), tot[key].numpy())
        evals, V = linalg.tica_from_sums(tot["S0"], tot["St"], tot["a"], tot["b"], tot["M"], 3)
        revals, rV = oracle.tica(Z, lag, 3)
        np.testing.assert_allclose(evals.numpy(), revals, rtol=1e-9)
        np.testing.assert_allclose(V.numpy(), rV, atol=1e-8)
        # ---- min / max and plain sums
        mn, mx = shards.allreduce_minmax(torch.tensor([float(rank)]), torch.tensor([float(rank)]))
        assert mn.item() == 0 and mx.item() == world - 1
        t = shards.allreduce_sum_(torch.ones(3, dtype=torch.float64))
        assert torch.all(t == world)
        c = torch.full((2,), float(rank))
        shards.broadcast_(c, 0)
        assert torch.all(c == 0)
        out.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        import traceback
        out.put((rank, "FAIL: " + traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("n,f,lag", [(600, 12, 7), (257, 5, 1)])
def test_frame_sharding_world2(n, f, lag):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, f, lag, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    results = [out.get(timeout=5) for _ in range(world)]
    for p in procs:
        assert p.exitcode == 0, results
    assert sorted(r[1] for r in results) == ["ok"] * world, results


def _kmeans_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from deep_cartograph_b200 import ops
        from deep_cartograph_b200.modules.statistics import statistics
        from deep_cartograph_b200.parallel import FrameShards, shard_range

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
            n_empty = int((counts == 0).sum())
            shift = 0.0
            if n_empty == 0:
                C_new = sums * (1.0 / counts).unsqueeze(1)
                shift = float(((C_new - C) ** 2).sum())
                C.copy_(C_new)
            out_info = torch.tensor([float(n_empty), shift], dtype=torch.float64)
            if info is not None:
                info.copy_(out_info)
                return info
            return out_info

        def fake_step_packed(Y, C, labels, work, absmax=None):
            r = fake_step(Y, C, labels)
            k, d = C.shape
            o = k * d
            work[:o] = r["sums"].reshape(-1); work[o:o + k] = r["counts"]; work[o + k:o + k + 3] = r["stats"]
            return {"sums": work[:o].view(k, d), "counts": work[o:o + k], "stats": work[o + k:o + k + 3],
                    "packed": work[:o + k + 3]}

        ops.kmeans_step, ops.kmeans_update_, ops.kmeans_step_packed_ = fake_step, fake_update, fake_step_packed
        g = np.random.default_rng(3)
        k, d, n = 6, 3, 4001
        cent = g.uniform(-1, 1, size=(k, d))
        Y = np.round(cent[g.integers(0, k, size=n)] + 0.15 * g.standard_normal((n, d)), 4)   # CSV hand-off precision
        init = Y[:k].copy()
        s, e = shard_range(n, rank, world)
        res = statistics.kmeans_lloyd(torch.from_numpy(Y[s:e]), torch.from_numpy(init), shards=FrameShards())
        ref = oracle.kmeans_lloyd(Y, init)
        assert res["n_iter"] == ref["n_iter"], (res["n_iter"], ref["n_iter"])
        np.testing.assert_array_equal(res["labels"].numpy(), ref["labels"][s:e])
        np.testing.assert_allclose(res["centers"].numpy(), ref["centers"], rtol=1e-12, atol=1e-12)
        # ---- empty-cluster relocation across shards (sklearn _k_means_common.pyx:167-211): the third
        # initial centre attracts nobody; the frame farthest from its centre lives on the LAST rank
        # and must become the new centre on every rank
        Y2 = np.concatenate([g.normal(0.0, 0.05, size=(40, 2)), g.normal(5.0, 0.05, size=(40, 2)),
                             np.array([[9.0, 9.0]])])
        init2 = np.array([[0.0, 0.0], [5.0, 5.0], [100.0, 100.0]])
        s2, e2 = shard_range(len(Y2), rank, world)
        res2 = statistics.kmeans_lloyd(torch.from_numpy(Y2[s2:e2]), torch.from_numpy(init2), shards=FrameShards())
        ref2 = oracle.kmeans_lloyd(Y2, init2)
        assert res2["n_iter"] == ref2["n_iter"]
        np.testing.assert_array_equal(res2["labels"].numpy(), ref2["labels"][s2:e2])
        np.testing.assert_allclose(res2["centers"].numpy(), ref2["centers"], rtol=1e-12, atol=1e-12)
        assert np.bincount(ref2["labels"], minlength=3)[2] == 1          # the relocated centre kept its frame
        out.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        out.put((rank, "FAIL: " + traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_lloyd_driver_world2():
    """The Lloyd driver with frames sharded over two ranks (per-iteration all-reduce of
    [sums | counts | stats], centre update, convergence tests) reproduces the unsharded float64
    Lloyd: same iteration count, identical labels, same centres.  The device E-step is emulated."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_kmeans_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    results = [out.get(timeout=5) for _ in range(world)]
    for p in procs:
        assert p.exitcode == 0, results
    assert sorted(r[1] for r in results) == ["ok"] * world, results


def test_shard_range_partitions_all_frames():
    from deep_cartograph_b200.parallel import shard_range
    for n, w in [(10, 3), (1_000_000, 8), (7, 8)]:
        rs = [shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))


def _deeptica_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deep_cartograph_b200 import ops
        from deep_cartograph_b200.modules.cv_learning import deep_tica
        from deep_cartograph_b200.parallel import FrameShards, shard_range
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from conftest import cpu_ticacov_sums, cpu_ticaloss
        ops.ticacov_sums = cpu_ticacov_sums
        ops.ticaloss = cpu_ticaloss
        torch.manual_seed(11)
        B, F, d = 301, 12, 3
        # slowly varying inputs so that C_tau is well conditioned
        base = torch.cumsum(torch.randn(B + 5, F), dim=0) * 0.1
        x_t, x_lag = base[:B].float(), base[5:B + 5].float()
        net = torch.nn.Sequential(torch.nn.Linear(F, 8), torch.nn.Tanh(), torch.nn.Linear(8, d))
        ref = [p.detach().clone() for p in net.parameters()]
        # single process, the whole minibatch, through torch.linalg autograd (independent of the
        # analytic eigen-loss gradient used by tica_loss)
        loss_all, ev_all = deep_tica.tica_loss_reference(net(x_t), net(x_lag), reg=1e-6)
        loss_all.backward()
        g_all = [p.grad.detach().clone() for p in net.parameters()]
        for p in net.parameters():
            p.grad = None
        # sharded: this rank's part of the minibatch, sums all-reduced
        s, e = shard_range(B, rank, world)
        shards = FrameShards()
        loss_loc, ev_loc = deep_tica.tica_loss(net(x_t[s:e]), net(x_lag[s:e]), reg=1e-6, shards=shards)
        loss_loc.backward()
        deep_tica.allreduce_gradients_(net, shards)
        assert abs(float(loss_loc) - float(loss_all)) <= 1e-10 * abs(float(loss_all)), (float(loss_loc), float(loss_all))
        np.testing.assert_allclose(ev_loc.detach().numpy(), ev_all.detach().numpy(), rtol=1e-10)
        # the last bias has an analytically zero gradient (the covariances are mean-free): absolute
        # tolerance on the scale of the largest gradient
        gmax = max(float(ga.abs().max()) for ga in g_all)
        for p, ga, r in zip(net.parameters(), g_all, ref):
            assert torch.equal(p.detach(), r)
            np.testing.assert_allclose(p.grad.numpy(), ga.numpy(), rtol=2e-4, atol=2e-6 * gmax)
        out.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        out.put((rank, "FAIL: " + traceback.format_exc()))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_deeptica_loss_and_gradients_world2():
    """SURVEY 8e, DeepTICA "exact all-reduce": with the minibatch split over two ranks the loss,
    the eigenvalues and the (summed) parameter gradients equal those of the whole minibatch in one
    process.  The correlation-sum kernel is emulated on the CPU."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_deeptica_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    results = [out.get(timeout=5) for _ in range(world)]
    for p in procs:
        assert p.exitcode == 0, results
    assert sorted(r[1] for r in results) == ["ok"] * world, results
