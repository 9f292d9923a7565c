"""CPU, world_size 2, gloo: the N>1 host logic of the hot path -- disjoint per-rank frames (partition = batch,
no data-path collective), MAX-over-ranks timing, whole-job rate, contiguous sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import dcf_b200 as dcf
    r, w = dcf.dist_util.init("gloo")
    assert (r, w) == (rank, world) and dcf.dist_util.env_rank_world() == (rank, world, rank)
    wl = dcf.synthetic.make_workload("tiny", seed=dcf.dist_util.rank_seed(100, rank))
    dcf.dist_util.barrier()
    ms, ms2 = dcf.dist_util.max_over_ranks([10.0 + 5.0 * rank, 3.0 - rank])
    lo, hi = dcf.dist_util.shard_frames(7, rank, world)
    out[rank] = dict(ms=ms, ms2=ms2, shard=(lo, hi), points_sum=float(wl["points"].sum()),
                     rate=dcf.dist_util.aggregate_rate(4, world, 10, ms))
    import torch.distributed as dist
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0]["ms"] == res[1]["ms"] == 15.0            # slowest rank's time on every rank
    assert res[0]["ms2"] == res[1]["ms2"] == 3.0
    assert res[0]["shard"] == (0, 4) and res[1]["shard"] == (4, 7)
    assert res[0]["points_sum"] != res[1]["points_sum"]    # disjoint frames per rank
    assert res[0]["rate"] == pytest.approx(4 * 2 * 10 / 0.015)


def test_single_process_is_a_noop(dcf):
    assert dcf.dist_util.max_over_ranks([1.5, 2.5]) == [1.5, 2.5]
    dcf.dist_util.barrier()
    assert dcf.dist_util.shard_frames(5, 0, 1) == (0, 5)
    assert [dcf.dist_util.shard_frames(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
