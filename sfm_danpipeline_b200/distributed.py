"""Multi-GPU all-pairs matching: one process per GPU, torch.distributed for the plumbing.

The reference is a single-threaded CPU loop over image pairs (findBestPair,
/root/reference/src/Sfm.cpp:511-515); the pairs are independent, so the path shards by pair:

  1. rank 0 packs ``imagesDescriptors`` into its device blob (H2D once) and BROADCASTS the blob
     to every other rank's blob (NCCL over NVLink) -- every rank holds all descriptors;
  2. the N(N-1)/2 pairs are dealt to ranks by a deterministic cost-balanced rule
     (cost = rows_q * rows_t) that every rank evaluates identically -- no scheduling traffic;
  3. every rank matches its shard with the CUDA library, chunk by chunk (results stay on its device);
  4. per-pair counts and the packed cv::DMatch records are GATHERED to rank 0 (send/recv of
     ragged segments) as every chunk finishes; rank 0 copies chunk k to host memory on a side
     stream while chunk k+1 is being matched (``match_and_gather``), and indexes the records per
     pair.  (``gather_results`` is the one-shot form of the same step.)

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


_SHARD_CACHE: dict = {}


def shard_pairs_cached(n_images: int, rows, world_size: int):
    """(all_pairs, shard_pairs) for a layout, computed once per (row counts, world size): the enumeration and the cost sort
    depend on nothing else, and on a job of ~20 000 pairs they cost ~1 ms -- not nothing next to a 70 ms step on 8 GPUs."""
    key = (int(n_images), tuple(int(r) for r in rows), int(world_size))
    hit = _SHARD_CACHE.get(key)
    if hit is None:
        if len(_SHARD_CACHE) > 8:
            _SHARD_CACHE.clear()
        pairs = all_pairs(n_images)
        hit = _SHARD_CACHE[key] = (pairs, shard_pairs(pairs, rows, world_size))
    return hit


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
    dev = torch.device("cuda", matcher.device)
    # the layout (image count, width, element type, rows per image) travels as two small tensors: no pickling on the path
    head = torch.zeros(3, dtype=torch.int64, device=dev)
    if rank == src:
        matcher.set_descriptors(descriptors)
        head = torch.tensor([len(matcher.rows), int(matcher.cols), int(bool(matcher.elem_u8))], dtype=torch.int64, device=dev)
    dist.broadcast(head, src=src, group=group)
    n_images, cols, elem_u8 = (int(x) for x in head.cpu().tolist())
    rows_t = torch.tensor(matcher.rows, dtype=torch.int32, device=dev) if rank == src else torch.empty(n_images, dtype=torch.int32, device=dev)
    if n_images:
        dist.broadcast(rows_t, src=src, group=group)
    rows = rows_t.cpu().tolist() if rank != src else list(matcher.rows)
    if rank != src:
        matcher.reserve_descriptors(rows, cols, bool(elem_u8))
    ptr, nbytes = matcher.descriptor_blob()
    if nbytes:
        blob = device_bytes_as_tensor(ptr, nbytes, matcher.device)
        dist.broadcast(blob, src=src, group=group)
        torch.cuda.current_stream(matcher.device).synchronize()
    return rows, cols


# --------------------------------------------------------------------------- gather
class _PinnedPool:
    """Page-locked int32 buffers for the gathered records.  A fresh pinned allocation costs more than the copy
    it serves, so buffers are recycled -- but never while a PairTable still views one: a table OWNS its buffer
    (``PairTable._keep``) and a finalizer hands it back when the table is garbage collected."""

    def __init__(self):
        self.free: list[torch.Tensor] = []

    def take(self, n_int32: int, pinned: bool) -> torch.Tensor:
        best = None
        for i, b in enumerate(self.free):
            if b.numel() >= n_int32 and b.is_pinned() == pinned and (best is None or b.numel() < self.free[best].numel()):
                best = i
        if best is not None:
            return self.free.pop(best)
        n = max(n_int32 + n_int32 // 4, 1 << 16)
        return torch.empty(n, dtype=torch.int32, pin_memory=pinned)

    def give(self, buf: torch.Tensor):
        self.free.append(buf)
        self.free.sort(key=lambda b: b.numel())
        del self.free[:-4]  # keep the four largest


_POOL = _PinnedPool()


@dataclass
class PairTable:
    """All-pairs result on the destination rank: pair i owns matches[offsets[i]:offsets[i]+counts[i]]."""
    pairs: np.ndarray    # (n,2) int32
    counts: np.ndarray   # (n,) int32
    offsets: np.ndarray  # (n,) int64
    matches: np.ndarray  # (total,) DMATCH_DTYPE -- a view of the buffer this table owns (valid as long as the table)
    _keep: object = None

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
    # the one device->host read, into a buffer the table owns
    buf = _POOL.take(all_matches.numel(), dev.type == "cuda")
    view = buf[: all_matches.numel()].view(all_matches.shape)
    view.copy_(all_matches, non_blocking=True)
    if dev.type == "cuda":
        torch.cuda.current_stream(dev).synchronize()
    return _own(PairTable(pairs, counts, offsets, view.numpy().view(DMATCH_DTYPE).reshape(-1)), buf)


def _own(table: "PairTable", buf: torch.Tensor) -> "PairTable":
    import weakref
    table._keep = buf
    weakref.finalize(table, _POOL.give, buf)
    return table


# --------------------------------------------------------------------------- pipelined match + gather
def chunk_bounds(shard_rows_q: np.ndarray, n_chunks: int) -> list[tuple[int, int]]:
    """Cut one rank's pair list (query-row count per pair, in shard order) into `n_chunks` contiguous ranges of
    about equal query rows (~ equal work).  Deterministic: every rank computes every other rank's cuts."""
    n = len(shard_rows_q)
    if n == 0:
        return [(0, 0)] * n_chunks
    csum = np.cumsum(shard_rows_q, dtype=np.int64)
    cuts = [0]
    for k in range(1, n_chunks):
        cuts.append(max(cuts[-1], int(np.searchsorted(csum, csum[-1] * k / n_chunks, side="left"))))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(n_chunks)]


def gather_chunks(pairs: np.ndarray, shards: list[np.ndarray], rows) -> int:
    """How many chunks the pipelined NCCL gather uses: one per ~16 Mi query rows of the busiest rank, between 2 (so that
    something overlaps) and 8; 1 for small jobs.  Measured at cfg-3 on 2 GPUs: a chunk boundary costs ~0.75 ms (the GPU idles
    while the host notices the end of the chunk, posts the sends and plans the next launch), the last chunk's device -> host
    copy is exposed -- 16 chunks cost 12 ms per step, the optimum is near sqrt(total copy time / boundary cost).  The same on every rank."""
    rows = np.asarray(rows, np.int64)
    busiest = max((int(rows[pairs[sh, 0]].sum()) for sh in shards if len(sh)), default=0)
    if busiest <= (4 << 20):
        return 1
    return int(min(8, max(2, -(-busiest // (16 << 20)))))


def match_and_gather(match_fn, pairs: np.ndarray, shards: list[np.ndarray], rows, dst: int = 0, group=None,
                     n_chunks: int | None = None):
    """Steps 3 + 4, pipelined: every rank matches its shard in `n_chunks` pieces; as soon as a piece is done its
    counts and records travel to `dst` (two point-to-point messages per rank: counts first, so that `dst` can size
    the second), and `dst` copies piece k to host memory on a side stream while piece k+1 is being matched.  What
    stays exposed at the end is the last piece only.

    match_fn(pairs_chunk, slot) -> (counts int32[n], records int32[m, 4]) on this rank's device (or CPU tensors under
    gloo); `slot` (0..2) names the output buffer set it may reuse -- a set is only reused three pieces later, when its
    sends / copies have long finished.  Returns a PairTable on `dst`, None elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    rows = np.asarray(rows, np.int64)
    K = n_chunks or gather_chunks(pairs, shards, rows)
    bounds = [chunk_bounds(rows[pairs[sh, 0]] if len(sh) else np.zeros(0, np.int64), K) for sh in shards]
    mine = pairs[shards[rank]]

    if rank != dst:
        inflight: list[list] = [[], [], []]
        for k in range(K):
            lo, hi = bounds[rank][k]
            slot = k % 3
            for w in inflight[slot]:
                w.wait()
            inflight[slot] = []
            if hi == lo:
                continue
            counts, recs = match_fn(mine[lo:hi], slot)
            inflight[slot].append(dist.isend(counts, dst, group=group))
            if recs.shape[0]:
                inflight[slot].append(dist.isend(recs, dst, group=group))
        for ws in inflight:
            for w in ws:
                w.wait()
        return None

    # ---- destination rank
    dev = None
    counts = np.zeros(len(pairs), np.int32)
    offsets = np.zeros(len(pairs), np.int64)
    worst = int(rows[pairs[:, 0]].sum()) if len(pairs) else 0
    cap = min(worst, max(worst // 4, 1 << 16))  # records; synthetic and real data keep ~1/8 of the query rows
    buf = None
    host = None  # int32 [cap, 4] view of buf
    used = 0
    side = None
    staging = [None, None, None]   # device buffers the other ranks' records land in, per slot
    events = [None, None, None]

    def ensure_host(n_records: int, pinned: bool):
        nonlocal buf, host, cap
        if host is not None and n_records <= cap:
            return
        if host is not None:  # grow: rare (more than a quarter of all query rows matched)
            if side is not None:
                side.synchronize()
            cap = max(n_records, min(worst, 2 * cap))
            nbuf = _POOL.take(cap * 4, pinned)
            nbuf[: used * 4].copy_(buf[: used * 4])
            _POOL.give(buf)
            buf = nbuf
        else:
            cap = max(cap, n_records)
            buf = _POOL.take(cap * 4, pinned)
        cap = buf.numel() // 4
        host = buf[: cap * 4].view(cap, 4)

    for k in range(K):
        slot = k % 3
        lo, hi = bounds[rank][k]
        own_counts = own_recs = None
        if events[slot] is not None:
            events[slot].synchronize()  # this slot's previous copies to the host (three pieces ago) have finished
        if hi > lo:
            own_counts, own_recs = match_fn(mine[lo:hi], slot)
            dev = own_counts.device
        if dev is None:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        is_cuda = dev.type == "cuda"
        if is_cuda and side is None:
            side = torch.cuda.Stream(dev)
        # phase 1: the other ranks' counts of piece k
        peers = [r for r in range(world) if r != rank and bounds[r][k][1] > bounds[r][k][0]]
        sizes = [bounds[r][k][1] - bounds[r][k][0] for r in peers]
        cbuf = torch.empty(sum(sizes), dtype=torch.int32, device=dev)
        cparts = list(cbuf.split(sizes)) if peers else []
        ops = [dist.P2POp(dist.irecv, cparts[i], r, group) for i, r in enumerate(peers)]
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        c_h = cbuf.cpu().numpy() if peers else np.zeros(0, np.int32)  # (synchronises: the sizes of phase 2)
        own_c_h = own_counts.cpu().numpy() if own_counts is not None else np.zeros(0, np.int32)
        totals = [int(x.sum()) for x in np.split(c_h, np.cumsum(sizes)[:-1])] if peers else []
        own_total = int(own_c_h.sum())
        ensure_host(used + own_total + sum(totals), is_cuda)
        # phase 2: their records, into this slot's staging buffer
        need = sum(totals)
        if staging[slot] is None or staging[slot].shape[0] < need:
            staging[slot] = torch.empty((max(need + need // 4, 1), 4), dtype=torch.int32, device=dev)
        sparts, o = [], 0
        for t in totals:
            sparts.append(staging[slot][o: o + t])
            o += t
        ops = [dist.P2POp(dist.irecv, sparts[i], r, group) for i, r in enumerate(peers) if totals[i]]
        works = dist.batch_isend_irecv(ops) if ops else []
        for w in works:
            w.wait()  # NCCL: the current stream waits, the host does not
        # directory of piece k: records are laid out in arrival order (own, then peers in rank order)
        def place(r, lo_r, c_arr, start):
            idx = shards[r][lo_r: lo_r + len(c_arr)]
            counts[idx] = c_arr
            offsets[idx] = start + np.concatenate([[0], np.cumsum(c_arr[:-1], dtype=np.int64)]) if len(c_arr) else start
        own_at = used
        if own_counts is not None:
            place(rank, lo, own_c_h, used)
        used += own_total
        peers_at = used
        for i, r in enumerate(peers):
            place(r, bounds[r][k][0], c_h[sum(sizes[:i]): sum(sizes[:i + 1])], used)
            used += totals[i]
        # device -> host of piece k on the side stream: overlaps the matching of piece k+1
        if is_cuda:
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                if own_total:
                    host[own_at: own_at + own_total].copy_(own_recs[:own_total], non_blocking=True)
                if need:
                    host[peers_at: peers_at + need].copy_(staging[slot][:need], non_blocking=True)
                events[slot] = torch.cuda.Event()
                events[slot].record(side)
        else:
            if own_total:
                host[own_at: own_at + own_total].copy_(own_recs[:own_total])
            if need:
                host[peers_at: peers_at + need].copy_(staging[slot][:need])
    if side is not None:
        side.synchronize()
    ensure_host(used, dev is not None and dev.type == "cuda")
    return _own(PairTable(pairs, counts, offsets, host[:used].numpy().view(DMATCH_DTYPE).reshape(-1)), buf)


# --------------------------------------------------------------------------- gather through a shared pinned host table
@dataclass
class ShardedPairTable:
    """All-pairs result on the destination rank when every rank wrote its records into its own shared-memory table
    (``match_and_share``): pair i lives in ``parts[owner[i]][offsets[i] : offsets[i] + counts[i]]``.  Same look-up interface
    as PairTable.  The parts are read-only maps of the other ranks' tables: valid until one of them matches again
    (``detach()`` copies everything into one private PairTable)."""
    pairs: np.ndarray    # (n,2) int32
    counts: np.ndarray   # (n,) int32
    offsets: np.ndarray  # (n,) int64, inside the owner's part
    owner: np.ndarray    # (n,) int16
    parts: list          # one DMATCH_DTYPE array per rank

    def getMatching(self, idx_query: int, idx_train: int) -> np.ndarray:
        if not hasattr(self, "_index"):
            self._index = {(int(q), int(t)): i for i, (q, t) in enumerate(self.pairs)}
        i = self._index[(idx_query, idx_train)]
        return self.parts[self.owner[i]][self.offsets[i]: self.offsets[i] + self.counts[i]]

    @property
    def n_matches(self) -> int:
        return int(sum(len(p) for p in self.parts))

    @property
    def nbytes(self) -> int:
        return int(sum(p.nbytes for p in self.parts))

    @property
    def matches(self) -> np.ndarray:
        return np.concatenate(self.parts) if self.parts else np.zeros(0, DMATCH_DTYPE)

    def detach(self) -> PairTable:
        base = np.concatenate([[0], np.cumsum([len(p) for p in self.parts])]).astype(np.int64)
        return PairTable(self.pairs, self.counts.copy(), self.offsets + base[self.owner], self.matches)


def assemble_shared(pairs: np.ndarray, shards: list[np.ndarray], metas: list[np.ndarray], prefix_base: str,
                    shm_dir: str = "/dev/shm") -> ShardedPairTable:
    """Destination side of match_and_share.  metas[r] = int32 [generation, n_lo, n_hi, counts of rank r's shard...]."""
    counts = np.zeros(len(pairs), np.int32)
    offsets = np.zeros(len(pairs), np.int64)
    owner = np.zeros(len(pairs), np.int16)
    parts = []
    for r, meta in enumerate(metas):
        gen, n = int(meta[0]), (int(meta[1]) & 0xFFFFFFFF) | (int(meta[2]) << 32)
        c = np.asarray(meta[3: 3 + len(shards[r])], np.int32)
        counts[shards[r]] = c
        offsets[shards[r]] = np.concatenate([[0], np.cumsum(c[:-1], dtype=np.int64)]) if len(c) else 0  # records lie in pair order
        owner[shards[r]] = r
        assert int(c.sum()) == n, (r, int(c.sum()), n)
        if n:
            parts.append(np.memmap(f"{shm_dir}{prefix_base}{r}.{gen}", dtype=DMATCH_DTYPE, mode="r", shape=(n,)))
        else:
            parts.append(np.zeros(0, DMATCH_DTYPE))
    return ShardedPairTable(pairs, counts, offsets, owner, parts)


def enable_shared_tables(matcher, group=None) -> str:
    """Once per matcher: every rank moves its match table into a shared-memory segment named after a job nonce that
    rank 0 draws and broadcasts.  Returns the common prefix ("/sfmm_<nonce>_"; rank r's table is "<prefix>r.<generation>")."""
    base = getattr(matcher, "_shm_prefix_base", None)
    if base:
        return base
    import os
    dev = torch.device("cuda", matcher.device)
    nonce = torch.tensor([int.from_bytes(os.urandom(7), "little")], dtype=torch.int64, device=dev)
    dist.broadcast(nonce, src=0, group=group)
    base = f"/sfmm_{int(nonce.item()):x}_"
    matcher.share_table(f"{base}{dist.get_rank(group)}")
    matcher._shm_prefix_base = base
    return base


def match_and_share(matcher, pairs: np.ndarray, shards: list[np.ndarray], dst: int = 0, group=None, rows=None):
    """Steps 3 + 4 without a funnel: every rank matches its shard through the library's own pipelined host path -- chunk k's
    records go device -> host over THIS GPU's PCIe link, straight into its (shared, page-locked) match table, while chunk k+1
    is being matched -- and `dst` maps the other ranks' tables.  The only traffic between ranks is one small NCCL gather of
    per-pair counts.  Returns a ShardedPairTable on `dst`, None elsewhere.

    If a rank cannot place its table in shared memory (a container with a small /dev/shm), every rank learns it through the
    same gather and the whole group falls back, for this call and the following ones, to the chunk-wise NCCL gather
    (match_and_gather; needs `rows`)."""
    from ._lib import SFMM_ENOMEM, SfmmError
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if getattr(matcher, "_shm_unavailable", False):
        return _gather_over_nccl(matcher, pairs, shards, rows, dst, group)
    base = enable_shared_tables(matcher, group)
    mine = pairs[shards[rank]]
    matcher.clear_results()
    failed = 0
    try:
        matcher.match_pairs(mine)
    except SfmmError as e:
        if e.code != SFMM_ENOMEM:
            raise
        failed = 1
    width = 4 + max(len(sh) for sh in shards)
    meta = np.zeros(width, np.int32)
    counts = np.zeros(0, np.int32)
    if not failed:
        _p, counts, _o, _m = matcher.result_table(copy=False)
        name, n = matcher.shared_table_info()
        meta[0], meta[2] = (int(name.rsplit(".", 1)[1]) if name else 0), n >> 32
        meta[1:2] = np.array([n & 0xFFFFFFFF], np.uint32).view(np.int32)
        meta[4: 4 + len(counts)] = counts
    meta[3] = failed
    dev = torch.device("cuda", matcher.device)
    t = torch.from_numpy(meta).to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)  # (everyone needs the failure flags; the payload is a few KB per rank)
    metas = [o.cpu().numpy() for o in out]
    if any(int(m[3]) for m in metas):
        matcher._shm_unavailable = True
        matcher.share_table(None)
        if rows is None:
            raise SfmmError(SFMM_ENOMEM, "shared-memory match tables do not fit on this host and no row counts were given for the NCCL fallback")
        return _gather_over_nccl(matcher, pairs, shards, rows, dst, group)
    if rank != dst:
        return None
    return assemble_shared(pairs, shards, [np.concatenate([m[:3], m[4:]]) for m in metas], base)


def _gather_over_nccl(matcher, pairs, shards, rows, dst, group):
    def match_fn(chunk, slot):
        c, m, _k = match_shard(matcher, chunk, rows, slot=slot)
        return c, m
    return match_and_gather(match_fn, pairs, shards, rows, dst, group)


# --------------------------------------------------------------------------- one rank's shard
_SHARD_BUF: dict = {}


def match_shard(matcher, mine: np.ndarray, rows, fraction: float = 0.25, slot: int = 0):
    """Step 3: match this rank's pairs, results left in torch-owned device buffers.

    The worst case is one record per query row (16 B each: 20 GB per rank at cfg-5), so the buffer
    is first sized for `fraction` of that (synthetic and real data keep ~1/8) and the call is
    repeated with the full size if the library reports SFMM_ERANGE.  Buffers are cached per (device, slot)
    and REUSED by the next call with the same slot: what is returned are views of them, valid until then.
    Returns (counts int32[len(mine)], matches int32[n,4], n)."""
    from ._lib import SFMM_ERANGE, SfmmError
    dev = torch.device("cuda", matcher.device)
    worst = int(np.asarray(rows, np.int64)[mine[:, 0]].sum()) if len(mine) else 0
    cap = min(worst, max(int(worst * fraction), 1 << 20))
    while True:
        key = (matcher.device, slot)
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
def match_all_pairs_distributed(matcher, descriptors, dst: int = 0, group=None, resident: bool = False, gather: str = "shared"):
    """Steps 1-4 on every rank of the default (NCCL) group.  `descriptors` is read on rank `dst`
    only; `resident=True` skips step 1 (descriptors already broadcast).  `gather` picks step 4:
      "shared"  every rank's records go device -> host over its own PCIe link into a shared page-locked table that `dst` maps
                (match_and_share; the default: no funnel, no GPU time spent on transfers),
      "nccl"    chunk-wise NCCL gather to `dst`'s device, copied to host there while the next chunk is matched (match_and_gather),
      "nccl-once" the whole shard first, one NCCL gather at the end (gather_results).
    Returns (PairTable / ShardedPairTable on dst | None, info dict)."""
    pipelined = gather == "nccl"
    import os
    import time
    trace = os.environ.get("SFMM_TRACE") == "1"
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    t0 = time.perf_counter()
    if not resident:
        broadcast_descriptors(matcher, descriptors, dst, group)
    t1 = time.perf_counter()
    rows = matcher.rows
    pairs, shards = shard_pairs_cached(len(rows), rows, world)
    mine = pairs[shards[rank]]
    t2 = time.perf_counter()
    n = 0
    if gather == "shared":
        table = match_and_share(matcher, pairs, shards, dst, group, rows=rows)
        n = 0 if getattr(matcher, "_shm_unavailable", False) else matcher.shared_table_info()[1]
        t3 = t4 = time.perf_counter()
    elif pipelined:
        def match_fn(chunk, slot):
            nonlocal n
            c, m, k = match_shard(matcher, chunk, rows, slot=slot)
            n += k
            return c, m
        table = match_and_gather(match_fn, pairs, shards, rows, dst, group)
        t3 = t4 = time.perf_counter()
    else:
        d_counts, d_matches, n = match_shard(matcher, mine, rows)
        t3 = time.perf_counter()
        table = gather_results(pairs, shards, d_counts, d_matches, dst, group)
        t4 = time.perf_counter()
    info = {"pairs_local": len(mine), "matches_local": n,
            "ms": {"broadcast": 1e3 * (t1 - t0), "shard": 1e3 * (t2 - t1), "match": 1e3 * (t3 - t2), "gather": 1e3 * (t4 - t3)}}
    if trace and rank == dst:
        print("[sfmm trace]", info["ms"], flush=True)
    return table, info
