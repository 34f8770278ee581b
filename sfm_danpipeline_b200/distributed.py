"""Multi-GPU all-pairs matching: one process per GPU, torch.distributed for the plumbing.

The reference is a single-threaded CPU loop over image pairs (findBestPair,
/root/reference/src/Sfm.cpp:511-515); the pairs are independent, so the path shards by pair:

  1. rank 0 packs ``imagesDescriptors`` into its device blob (H2D once) and BROADCASTS the blob
     to every other rank's blob (NCCL over NVLink) -- every rank holds all descriptors;
  2. the N(N-1)/2 pairs are dealt to ranks by a deterministic cost-balanced rule
     (cost = rows_q * rows_t) that every rank evaluates identically -- no scheduling traffic;
  3. every rank matches its shard with the CUDA library (results stay on its device);
  4. per-pair counts and the packed cv::DMatch records are GATHERED to rank 0 (send/recv of
     ragged segments), which copies them to host memory once and indexes them per pair.

There is no collective inside the matching itself.  The same functions run on CPU tensors with
the gloo backend (tests/test_distributed_cpu.py drives them with the oracle standing in for
the CUDA matcher) so the host-side logic is covered without GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from ._lib import DMATCH_DTYPE


# --------------------------------------------------------------------------- sharding
def all_pairs(n_images: int) -> np.ndarray:
    """findBestPair's enumeration: every q<t, row-major, as an (n,2) int32 array."""
    q, t = np.triu_indices(n_images, 1)
    return np.stack([q, t], 1).astype(np.int32)


def shard_pairs(pairs: np.ndarray, rows, world_size: int) -> list[np.ndarray]:
    """Indices (into `pairs`) owned by each rank.

    Cost-sorted snake deal: pairs sorted by descending rows_q*rows_t (stable), dealt
    0..W-1, W-1..0, ...; each rank's list is then put back in ascending pair order so
    consecutive launches share a query image (L2 reuse).  Deterministic on every rank."""
    rows = np.asarray(rows, np.int64)
    cost = rows[pairs[:, 0]] * rows[pairs[:, 1]]
    order = np.argsort(-cost, kind="stable")
    pos = np.arange(len(order))
    lap, off = pos // world_size, pos % world_size
    owner = np.where(lap % 2 == 0, off, world_size - 1 - off)
    return [np.sort(order[owner == r]) for r in range(world_size)]


# --------------------------------------------------------------------------- broadcast
class _DevPtr:
    """Zero-copy view of a raw device allocation for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def device_bytes_as_tensor(ptr: int, nbytes: int, device: int) -> torch.Tensor:
    return torch.as_tensor(_DevPtr(ptr, nbytes), device=torch.device("cuda", device))


def broadcast_descriptors(matcher, descriptors, src: int = 0, group=None):
    """Step 1.  `descriptors` is only read on rank `src`.  Returns (rows, cols)."""
    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        matcher.set_descriptors(descriptors)
        meta[0] = (list(matcher.rows), int(matcher.cols))
    dist.broadcast_object_list(meta, src=src, group=group)
    rows, cols = meta[0]
    if rank != src:
        matcher.reserve_descriptors(rows, cols)
    ptr, nbytes = matcher.descriptor_blob()
    if nbytes:
        blob = device_bytes_as_tensor(ptr, nbytes, matcher.device)
        dist.broadcast(blob, src=src, group=group)
        torch.cuda.current_stream(matcher.device).synchronize()
    return rows, cols


# --------------------------------------------------------------------------- gather
_PINNED: dict = {}


def _to_host(t: torch.Tensor) -> np.ndarray:
    """Device -> host through a cached page-locked buffer (a fresh pinned allocation per call
    would cost more than the copy)."""
    if t.device.type != "cuda":
        return t.numpy()
    n = t.numel()
    buf = _PINNED.get(t.dtype)
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n + n // 4, 1 << 16), dtype=t.dtype, pin_memory=True)
        _PINNED[t.dtype] = buf
    view = buf[:n].view(t.shape)
    view.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return view.numpy()


@dataclass
class PairTable:
    """All-pairs result on the destination rank: pair i owns matches[offsets[i]:offsets[i]+counts[i]]."""
    pairs: np.ndarray    # (n,2) int32
    counts: np.ndarray   # (n,) int32
    offsets: np.ndarray  # (n,) int64
    matches: np.ndarray  # (total,) DMATCH_DTYPE

    def getMatching(self, idx_query: int, idx_train: int) -> np.ndarray:
        if not hasattr(self, "_index"):
            self._index = {(int(q), int(t)): i for i, (q, t) in enumerate(self.pairs)}
        i = self._index[(idx_query, idx_train)]
        return self.matches[self.offsets[i]: self.offsets[i] + self.counts[i]]


def gather_results(pairs: np.ndarray, shards: list[np.ndarray], local_counts: torch.Tensor,
                   local_matches: torch.Tensor, dst: int = 0, group=None):
    """Step 4.  local_counts: int32[len(shard)]; local_matches: int32[n_local, 4] (cv::DMatch records
    viewed as 4 x 32 bit), both on this rank's device (or CPU for gloo).  Returns a PairTable on
    `dst`, None elsewhere.  Segments are not re-ordered: the offsets table points into them."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = local_counts.device
    # tiny metadata: every rank's record total
    totals = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([local_matches.shape[0]], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(totals, mine, group=group) if dev.type == "cuda" else \
        dist.all_gather(list(totals.split(1)), mine, group=group)
    totals_h = totals.cpu().numpy()
    # every ragged segment moves in ONE grouped point-to-point launch (batch_isend_irecv): unbatched
    # send/recv pairs are serialised as separate collectives by the NCCL process group
    if rank != dst:
        ops = []
        if len(shards[rank]):
            ops.append(dist.P2POp(dist.isend, local_counts, dst, group))
        if totals_h[rank]:
            ops.append(dist.P2POp(dist.isend, local_matches, dst, group))
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return None
    base = np.concatenate([[0], np.cumsum(totals_h)]).astype(np.int64)
    all_matches = torch.empty((int(base[-1]), 4), dtype=torch.int32, device=dev)
    counts = np.zeros(len(pairs), np.int32)
    offsets = np.zeros(len(pairs), np.int64)
    cbase = np.concatenate([[0], np.cumsum([len(sh) for sh in shards])]).astype(np.int64)
    all_counts = torch.empty(int(cbase[-1]), dtype=torch.int32, device=dev)
    ops = []
    for r in range(world):
        if r == rank:
            all_counts[cbase[r]: cbase[r + 1]] = local_counts
            if totals_h[r]:
                all_matches[base[r]: base[r + 1]] = local_matches
        else:
            if len(shards[r]):
                ops.append(dist.P2POp(dist.irecv, all_counts[cbase[r]: cbase[r + 1]], r, group))
            if totals_h[r]:
                ops.append(dist.P2POp(dist.irecv, all_matches[base[r]: base[r + 1]], r, group))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    counts_h = all_counts.cpu().numpy()
    for r in range(world):
        c_h = counts_h[cbase[r]: cbase[r + 1]]
        if len(c_h):
            counts[shards[r]] = c_h
            offsets[shards[r]] = base[r] + np.concatenate([[0], np.cumsum(c_h[:-1], dtype=np.int64)])
    host = _to_host(all_matches).view(DMATCH_DTYPE).reshape(-1)  # the one device->host read (view of a reused pinned buffer)
    return PairTable(pairs, counts, offsets, host)


# --------------------------------------------------------------------------- one rank's shard
_SHARD_BUF: dict = {}


def match_shard(matcher, mine: np.ndarray, rows, fraction: float = 0.25):
    """Step 3: match this rank's pairs, results left in torch-owned device buffers.

    The worst case is one record per query row (16 B each: 20 GB per rank at cfg-5), so the buffer
    is first sized for `fraction` of that (synthetic and real data keep ~1/8) and the call is
    repeated with the full size if the library reports SFMM_ERANGE.  Buffers are cached per device.
    Returns (counts int32[len(mine)], matches int32[n,4], n)."""
    from ._lib import SFMM_ERANGE, SfmmError
    dev = torch.device("cuda", matcher.device)
    worst = int(np.asarray(rows, np.int64)[mine[:, 0]].sum()) if len(mine) else 0
    cap = min(worst, max(int(worst * fraction), 1 << 20))
    while True:
        key = (matcher.device,)
        buf = _SHARD_BUF.get(key)
        if buf is None or buf[0].numel() < max(len(mine), 1) or buf[1].shape[0] < max(cap, 1):
            buf = (torch.empty(max(len(mine), 1), dtype=torch.int32, device=dev),
                   torch.empty((max(cap, 1), 4), dtype=torch.int32, device=dev))
            _SHARD_BUF[key] = buf
        d_counts, d_matches = buf
        torch.cuda.current_stream(dev).synchronize()
        try:
            n = matcher.match_pairs_device(mine, d_counts.data_ptr(), d_matches.data_ptr(), cap)
            return d_counts[: len(mine)], d_matches[:n], n
        except SfmmError as e:
            if e.code != SFMM_ERANGE or cap >= worst:
                raise
            cap = worst


# --------------------------------------------------------------------------- the whole job
def match_all_pairs_distributed(matcher, descriptors, dst: int = 0, group=None, resident: bool = False):
    """Steps 1-4 on every rank of the default (NCCL) group.  `descriptors` is read on rank `dst`
    only; `resident=True` skips step 1 (descriptors already broadcast).  Returns
    (PairTable on dst | None, info dict)."""
    import os
    import time
    trace = os.environ.get("SFMM_TRACE") == "1"
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    t0 = time.perf_counter()
    if not resident:
        broadcast_descriptors(matcher, descriptors, dst, group)
    t1 = time.perf_counter()
    rows = matcher.rows
    pairs = all_pairs(len(rows))
    shards = shard_pairs(pairs, rows, world)
    mine = pairs[shards[rank]]
    t2 = time.perf_counter()
    d_counts, d_matches, n = match_shard(matcher, mine, rows)
    t3 = time.perf_counter()
    table = gather_results(pairs, shards, d_counts, d_matches, dst, group)
    t4 = time.perf_counter()
    info = {"pairs_local": len(mine), "matches_local": n,
            "ms": {"broadcast": 1e3 * (t1 - t0), "shard": 1e3 * (t2 - t1), "match": 1e3 * (t3 - t2), "gather": 1e3 * (t4 - t3)}}
    if trace and rank == dst:
        print("[sfmm trace]", info["ms"], flush=True)
    return table, info
