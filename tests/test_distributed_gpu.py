"""GPU test of the multi-rank host path over NCCL (world size = number of visible GPUs, 1 on the
driver's box): broadcast into the library's blob, sharded matching into torch-owned device buffers,
ragged gather to rank 0 -- compared with the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from sfm_danpipeline_b200 import synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    from sfm_danpipeline_b200 import Matcher, NORM_HAMMING
    from sfm_danpipeline_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rows = [900, 0, 1300, 257, 64, 2100]
        descs = synth.binary_images(len(rows), rows, seed=41)
        descs[1] = np.zeros((0, 61), np.uint8)
        with Matcher(NORM_HAMMING, device=rank) as m:
            ok = True
            for gather in ("shared", "nccl", "nccl-once"):  # every form of step 4 gives the same table
                for _ in range(2):  # twice: cached buffers, re-broadcast
                    table, info = D.match_all_pairs_distributed(m, descs if rank == 0 else None, 0, gather=gather)
                if rank == 0:
                    for (q, t) in synth.all_pairs(len(rows)):
                        ok &= np.asarray(table.getMatching(q, t)).tobytes() == oracle.match_pair(descs[q], descs[t], 0).tobytes()
                    ok &= int(table.counts.sum()) == len(table.matches)
                table = None
                dist.barrier()
            if rank == 0:
                open(out_path, "w").write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


def test_broadcast_shard_gather_over_nccl(tmp_path):
    world = torch.cuda.device_count()  # all visible GPUs (1 on the driver's test box, up to 8 under gpurun --gpus N)
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert open(out).read() == "ok"
