"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: pair sharding and the ragged
gather to rank 0.  The oracle stands in for the CUDA matcher (this is tests/, the checker may be
used here); the functions under test are the ones bench.py runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from sfm_danpipeline_b200 import distributed as D
from sfm_danpipeline_b200 import synth


def test_all_pairs_enumeration_is_findbestpair_order():
    assert D.all_pairs(4).tolist() == [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]]
    assert D.all_pairs(1).shape == (0, 2)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shards_partition_the_pairs_and_balance_cost(world):
    rng = np.random.default_rng(world)
    rows = rng.integers(1, 20000, 40)
    pairs = D.all_pairs(40)
    shards = D.shard_pairs(pairs, rows, world)
    allidx = np.concatenate(shards)
    assert sorted(allidx.tolist()) == list(range(len(pairs)))  # disjoint cover
    assert all((np.diff(s) > 0).all() for s in shards)  # ascending inside a rank
    cost = rows[pairs[:, 0]].astype(np.int64) * rows[pairs[:, 1]]
    per = np.array([cost[s].sum() for s in shards], float)
    assert per.max() / per.mean() < 1.05
    again = D.shard_pairs(pairs, rows, world)
    assert all((a == b).all() for a, b in zip(shards, again))  # deterministic on every rank


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_match_fn(descs):
    """Stands in for the CUDA matcher: (pairs chunk, slot) -> (counts, records) as CPU tensors."""
    def fn(chunk, slot):
        res = [oracle.match_pair(descs[q], descs[t], 0) for q, t in chunk]
        counts = torch.tensor([len(r) for r in res], dtype=torch.int32)
        cat = np.concatenate(res) if res else np.zeros(0, oracle.DMATCH_DTYPE)
        return counts, torch.from_numpy(cat.view(np.int32).reshape(-1, 4).copy())
    return fn


def _worker(rank, world, port, rows_list, out_path, n_chunks=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        descs = synth.binary_images(len(rows_list), rows_list, seed=4)
        pairs = D.all_pairs(len(descs))
        shards = D.shard_pairs(pairs, rows_list, world)
        mine = pairs[shards[rank]]
        if n_chunks:  # the pipelined form: matched and gathered piece by piece
            table = D.match_and_gather(_oracle_match_fn(descs), pairs, shards, rows_list, dst=0, n_chunks=n_chunks)
        else:
            counts, matches = _oracle_match_fn(descs)(mine, 0)
            table = D.gather_results(pairs, shards, counts, matches, dst=0)
        if rank == 0:
            ok = True
            for i, (q, t) in enumerate(pairs):
                exp = oracle.match_pair(descs[q], descs[t], 0)
                got = table.getMatching(int(q), int(t))
                ok &= got.tobytes() == exp.tobytes()
            ok &= int(table.counts.sum()) == len(table.matches)
            open(out_path, "w").write("ok" if ok else "mismatch")
        else:
            assert table is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,rows", [(2, [120, 0, 300, 64, 1, 200]), (3, [50, 60])])
def test_gather_to_rank0_over_gloo(tmp_path, world, rows):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(world, _free_port(), rows, out), nprocs=world, join=True)
    assert open(out).read() == "ok"


@pytest.mark.parametrize("world,rows,chunks", [(2, [120, 0, 300, 64, 1, 200], 3), (3, [50, 60], 2), (3, [90, 70, 0, 33, 150], 5),
                                               (8, [60, 0, 90, 33, 1, 120, 75, 48, 80], 4)])
def test_pipelined_match_and_gather_over_gloo(tmp_path, world, rows, chunks):
    """match_and_gather: pieces with no pairs on some ranks, ranks without any pair, empty images -- world size up to 8."""
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(world, _free_port(), rows, out, chunks), nprocs=world, join=True)
    assert open(out).read() == "ok"


def test_chunk_bounds_cover_the_shard_on_every_rank():
    rng = np.random.default_rng(0)
    for n, k in [(0, 3), (1, 4), (7, 7), (100, 8), (5, 16)]:
        r = rng.integers(0, 5000, n)
        b = D.chunk_bounds(r, k)
        assert len(b) == k and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] and b[i][0] <= b[i][1] for i in range(k - 1))
    rows = rng.integers(1, 20000, 30)
    pairs = D.all_pairs(30)
    shards = D.shard_pairs(pairs, rows, 4)
    assert 1 <= D.gather_chunks(pairs, shards, rows) <= 16


def test_returned_tables_do_not_alias_each_other():
    """A PairTable owns its (pooled) buffer: a later gather must not overwrite an earlier table."""
    a = D._POOL.take(1000, False)
    ta = D._own(D.PairTable(np.zeros((0, 2), np.int32), np.zeros(0, np.int32), np.zeros(0, np.int64),
                            a[:8].numpy().view(oracle.DMATCH_DTYPE)), a)
    b = D._POOL.take(1000, False)
    assert b.data_ptr() != a.data_ptr()
    del ta
    import gc
    gc.collect()
    assert any(x.data_ptr() == a.data_ptr() for x in D._POOL.free)  # handed back only once the table is gone


def test_assemble_shared_tables_from_segments(tmp_path):
    """Destination side of the shared-table gather (match_and_share): directory + read-only maps of every rank's segment."""
    rows = [120, 0, 300, 64, 1, 200]
    descs = synth.binary_images(len(rows), rows, seed=4)
    pairs = D.all_pairs(len(rows))
    world = 3
    shards = D.shard_pairs(pairs, rows, world)
    base = "/sfmm_test_"
    metas = []
    width = 3 + max(len(sh) for sh in shards)
    for r in range(world):
        res = [oracle.match_pair(descs[q], descs[t], 0) for q, t in pairs[shards[r]]]
        cat = np.concatenate(res) if res else np.zeros(0, oracle.DMATCH_DTYPE)
        gen = 5 + r
        if len(cat):
            cat.tofile(str(tmp_path) + f"{base}{r}.{gen}")
        meta = np.zeros(width, np.int32)
        meta[0], meta[1], meta[2] = gen, len(cat), 0
        meta[3: 3 + len(res)] = [len(x) for x in res]
        metas.append(meta)
    table = D.assemble_shared(pairs, shards, metas, base, shm_dir=str(tmp_path))
    for q, t in pairs:
        assert np.asarray(table.getMatching(int(q), int(t))).tobytes() == oracle.match_pair(descs[q], descs[t], 0).tobytes()
    flat = table.detach()
    assert int(flat.counts.sum()) == len(flat.matches) == table.n_matches
    for q, t in pairs:
        assert flat.getMatching(int(q), int(t)).tobytes() == oracle.match_pair(descs[q], descs[t], 0).tobytes()
