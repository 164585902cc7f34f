"""Frame sharding across the GPUs of one box.

One process per GPU (``torchrun``); rank r owns the contiguous frame range
``[r*N/G, (r+1)*N/G)``.  Every pass of the hot path is a sum / min / max over independent
frames, so the data path needs exactly three exchanges, all issued through
``torch.distributed`` (NCCL over NVLink on the GPUs, gloo in the CPU tests):

  * statistics: all-gather of the per-rank (n, mean, M2, min, max) and a Chan merge (FP64);
  * lag-tau halo: each rank receives the first ``lag`` rows of the next rank (the pairs (t, t+lag)
    near a shard edge straddle it); the last rank has none;
  * partial sums: one FP64 SUM all-reduce of ``[S0 | St | a | b | M]`` per covariance pass and of
    ``[sums | counts | stats]`` per Lloyd iteration; min / max all-reduce of the CV range.
"""
from __future__ import annotations

import ctypes
import logging
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

logger = logging.getLogger(__name__)

# doubles per rank and exchange of the peer-memory all-reduce (KMeans k = 1000, d = 10 packs 11 003)
_P2P_SLOT_DOUBLES = 16384


class PeerAllReduce:
    """The small all-reduces of the sharded path as ONE kernel per GPU over NVLink peer memory
    (``csrc/p2p.cu``: push into every peer's inbox, flags, reduce in rank order) instead of an NCCL call
    each: at a few KB these are pure latency.  Built collectively (every rank of the group must construct
    it); inboxes are cudaMalloc allocations shared through CUDA IPC handles."""

    def __init__(self, group, rank: int, world: int, device: torch.device):
        from . import _lib
        self._lib = _lib
        lib = _lib.load()
        self.rank, self.world, self.device = rank, world, device
        self.slot = _P2P_SLOT_DOUBLES
        nbytes = lib.dcg_p2p_buffer_bytes(world, self.slot)
        if nbytes == 0:
            raise RuntimeError("peer all-reduce: world size out of range")
        own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(device):
            _lib.check("dcg_p2p_alloc", lib.dcg_p2p_alloc(nbytes, ctypes.byref(own), handle))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self._bases = (ctypes.c_void_p * world)()
            self._opened = []
            for r in range(world):
                if r == rank:
                    self._bases[r] = own.value
                    continue
                peer = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
                _lib.check("dcg_p2p_open", lib.dcg_p2p_open(buf, ctypes.byref(peer)))
                self._bases[r] = peer.value
                self._opened.append(peer.value)
        self._own = own.value
        self.seq = 0
        dist.barrier(group=group)            # every inbox is mapped (and zeroed) before the first exchange

    def allreduce_(self, t: torch.Tensor, op: int) -> torch.Tensor:
        """In-place SUM (op 0) / MAX (op 1) of a contiguous float64 CUDA tensor of at most ``slot`` elements."""
        self.seq += 1
        lib = self._lib.load()
        with torch.cuda.device(t.device):
            self._lib.check("dcg_p2p_allreduce_f64",
                            lib.dcg_p2p_allreduce_f64(t.data_ptr(), t.data_ptr(), t.numel(), op, self.rank, self.world,
                                                      self._bases, self.slot, self.seq,
                                                      torch.cuda.current_stream(t.device).cuda_stream))
        return t

    def fits(self, t: torch.Tensor) -> bool:
        return (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and 0 < t.numel() <= self.slot
                and t.device == self.device)


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows ``[start, stop)`` owned by ``rank``."""
    return (rank * n_total) // world, ((rank + 1) * n_total) // world


class FrameShards:
    """Collectives of the frame-sharded hot path for one process group."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._peer = None                    # PeerAllReduce, built on first use (collectively)
        self._peer_failed = False

    def _peer_allreduce(self, t: torch.Tensor):
        """The peer-memory all-reduce object when ``t`` qualifies (CUDA, NCCL group on one box, few KB), else None.
        ``DCG_P2P_ALLREDUCE=0`` keeps every exchange on NCCL.  The decision depends only on properties that are
        equal on all ranks (backend, world size, dtype, element count), so it is collective."""
        if (self._peer_failed or self.world < 2 or not t.is_cuda or t.dtype != torch.float64
                or t.numel() > _P2P_SLOT_DOUBLES or t.numel() == 0 or not t.is_contiguous()
                or os.environ.get("DCG_P2P_ALLREDUCE", "1") == "0" or dist.get_backend(self.group) != "nccl"):
            return None
        if self._peer is None:
            ok = torch.ones(1, dtype=torch.int32, device=t.device)
            try:
                self._peer = PeerAllReduce(self.group, self.rank, self.world, t.device)
            except Exception as exc:          # no IPC / peer access on this box: NCCL for everything
                logger.warning("peer-memory all-reduce unavailable (%s); using NCCL", exc)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)      # all ranks or none
            if int(ok.item()) == 0:
                self._peer, self._peer_failed = None, True
                return None
        return self._peer

    # ---- statistics ---------------------------------------------------------------------------
    def merge_stats(self, st: dict, n_total: Optional[int] = None) -> dict:
        """All-gather the per-rank column statistics and merge them on the device in one shot
        (FP64): N = sum n_r, mean = sum n_r mean_r / N, M2 = sum M2_r + sum n_r (mean_r - mean)^2,
        min / max over ranks.  One collective, one host read (the global frame count) -- none when
        the caller already knows ``n_total`` (repeated passes over the same shards)."""
        f = st["mean"].numel()
        dev = st["mean"].device
        if dev.type == "cuda":
            # one launch to pack, one collective, one launch to merge (the eager version below is ~25 small
            # kernels: 0.25 ms of the C2 step at every GPU count > 1)
            from . import ops
            packed = ops.stats_pack(st)
            flat = torch.empty(self.world * (1 + 4 * f), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(flat, packed, group=self.group)
            m = ops.stats_merge(flat, self.world, f)
            n_all = int(n_total) if n_total is not None else int(round(float(m["n"].item())))
            return {"n": n_all, "mean": m["mean"], "m2": m["m2"], "min": m["min"], "max": m["max"]}
        packed = torch.empty(1 + 4 * f, dtype=torch.float64, device=dev)
        packed[0] = float(st["n"])
        packed[1:1 + f] = st["mean"]
        packed[1 + f:1 + 2 * f] = st["m2"]
        packed[1 + 2 * f:1 + 3 * f] = st["min"].to(torch.float64)
        packed[1 + 3 * f:] = st["max"].to(torch.float64)
        flat = torch.empty(self.world * (1 + 4 * f), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(flat, packed, group=self.group)
        allp = flat.view(self.world, 1 + 4 * f)
        n_r = allp[:, 0:1]                                   # (world, 1); empty shards carry n = 0
        live = n_r > 0
        N = n_r.sum()
        mean_r = torch.where(live, allp[:, 1:1 + f], torch.zeros((), dtype=torch.float64, device=dev))
        mean = (n_r * mean_r).sum(dim=0) / N
        m2 = (torch.where(live, allp[:, 1 + f:1 + 2 * f], torch.zeros((), dtype=torch.float64, device=dev))
              + n_r * (mean_r - mean) ** 2).sum(dim=0)
        inf = torch.full((), float("inf"), dtype=torch.float64, device=dev)
        mn = torch.where(live, allp[:, 1 + 2 * f:1 + 3 * f], inf).min(dim=0).values.to(torch.float32)
        mx = torch.where(live, allp[:, 1 + 3 * f:], -inf).max(dim=0).values.to(torch.float32)
        n_all = int(n_total) if n_total is not None else int(round(float(N.item())))
        return {"n": n_all, "mean": mean, "m2": m2, "min": mn, "max": mx}

    # ---- halo ---------------------------------------------------------------------------------
    def with_halo(self, X: torch.Tensor, lag: int) -> torch.Tensor:
        """Own rows followed by the first ``lag`` rows of the next rank (none on the last rank).
        If ``X`` is a leading view of a buffer with ``lag`` spare rows, the halo is received in
        place (no copy of the shard)."""
        if lag == 0 or self.world == 1:
            return X
        n, f = X.shape
        if n < lag:
            raise ValueError(f"shard of {n} frames is shorter than the lag {lag}")
        # in place when X is the leading view of a 2-D buffer (possibly with padded rows) that has
        # `lag` spare rows after the shard
        base = X._base if X._base is not None else None
        room = (base is not None and base.dim() == 2 and base.shape[1] >= f and base.stride(1) == 1
                and base.stride(0) == X.stride(0) and base.data_ptr() == X.data_ptr()
                and base.shape[0] >= n + lag)
        has_next = self.rank + 1 < self.world
        if has_next:
            out = base[:n + lag, :f] if room else torch.cat([X, torch.empty((lag, f), dtype=X.dtype, device=X.device)])
            halo = out[n:n + lag]
        else:
            out = X
            halo = None
        head = X[:lag].contiguous()
        recv = None
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, head, self._global(self.rank - 1), group=self.group))
        if has_next:
            recv = halo if halo.is_contiguous() else torch.empty((lag, f), dtype=X.dtype, device=X.device)
            ops.append(dist.P2POp(dist.irecv, recv, self._global(self.rank + 1), group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if recv is not None and recv is not halo:
            halo.copy_(recv)
        return out

    def _global(self, group_rank: int) -> int:
        return dist.get_global_rank(self.group, group_rank) if self.group is not None else group_rank

    # ---- reductions ---------------------------------------------------------------------------
    def allreduce_sums(self, s: dict, m_total: Optional[int] = None) -> dict:
        """FP64 SUM all-reduce of one fused buffer [S0 | St | a | b | M].  ``m_total`` (the global
        number of lagged pairs, n_total - lag) saves the host read of the reduced M."""
        keys = [k for k in ("S0", "St", "a", "b") if s.get(k) is not None]
        dev = s["a"].device
        flat = s.get("flat")

        def _views_of_flat():
            # the dict's tensors must still BE the slices of `flat` (a caller may have replaced one, e.g. after
            # re-standardising speculative sums): otherwise pack as usual
            if flat is None or flat.numel() != sum(s[k].numel() for k in keys) + 1:
                return False
            o = 0
            for k in keys:
                if s[k].data_ptr() != flat.data_ptr() + 8 * o or not s[k].is_contiguous():
                    return False
                o += s[k].numel()
            return True

        if _views_of_flat():
            # ops.lagged_covariance laid the sums out as one buffer [S0 | St | a | b | M]: reduce it in place
            flat[-1:].fill_(float(s["M"]))
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            out = dict(s)
            out["M"] = int(m_total) if m_total is not None else int(round(float(flat[-1].item())))
            return out
        flat = torch.cat([s[k].reshape(-1) for k in keys] +
                         [torch.tensor([float(s["M"])], dtype=torch.float64, device=dev)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        out = dict(s)
        o = 0
        for k in keys:
            nel = s[k].numel()
            out[k] = flat[o:o + nel].view(s[k].shape)
            o += nel
        out["M"] = int(m_total) if m_total is not None else int(round(float(flat[o].item())))
        return out

    def allreduce_minmax(self, mn: torch.Tensor, mx: torch.Tensor):
        """Global per-column min and max in ONE collective: MAX over [-min | max]."""
        d = mn.numel()
        buf = torch.cat([-mn, mx])
        b64 = buf.to(torch.float64)
        peer = self._peer_allreduce(b64)
        if peer is not None:
            buf = peer.allreduce_(b64, 1).to(buf.dtype)      # float32 -> float64 -> float32 is exact
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=self.group)
        return -buf[:d], buf[d:]

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        peer = self._peer_allreduce(t)
        if peer is not None:
            return peer.allreduce_(t, 0)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def broadcast_(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        dist.broadcast(t, self._global(src), group=self.group)
        return t
