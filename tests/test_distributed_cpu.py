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


def _worker(rank, world, port, rows_list, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        descs = synth.binary_images(len(rows_list), rows_list, seed=4)
        pairs = D.all_pairs(len(descs))
        shards = D.shard_pairs(pairs, rows_list, world)
        mine = pairs[shards[rank]]
        res = [oracle.match_pair(descs[q], descs[t], 0) for q, t in mine]
        counts = torch.tensor([len(r) for r in res], dtype=torch.int32)
        cat = np.concatenate(res) if res else np.zeros(0, oracle.DMATCH_DTYPE)
        matches = torch.from_numpy(cat.view(np.int32).reshape(-1, 4).copy())
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
