"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups exercise the partitioning
and the one real exchange step of the row-sharded mode (gather of compacted point lists)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sh = importlib.import_module("3dscan_b200.sharding")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_points(rank, cap=1000):
    rng = np.random.default_rng(100 + rank)
    n = [317, 0, 954, 12][rank % 4]
    pts = torch.from_numpy(rng.normal(size=(cap, 3)).astype(np.float32))
    return pts, n


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts, n = _rank_points(rank)
        want = torch.cat([_rank_points(r)[0][:_rank_points(r)[1]] for r in range(world)])
        got, counts = sh.gather_points(pts, n, dst=0)
        assert counts == [_rank_points(r)[1] for r in range(world)]
        if rank == 0:
            assert torch.equal(got, want)
        else:
            assert got is None
        # destination other than 0, preallocated output
        out = torch.empty((4000, 3)) if rank == world - 1 else None
        got, _ = sh.gather_points(pts, n, dst=world - 1, out=out)
        if rank == world - 1:
            assert torch.equal(got, want) and got.data_ptr() == out.data_ptr()
        # the aligned-staging route of the CUDA path (forced here), with the tensors sized exactly to
        # each rank's own count: a peer may send more rows than this rank holds
        got, _ = sh.gather_points(pts[:n].clone(), n, dst=0, aligned_staging=True)
        if rank == 0:
            assert torch.equal(got, want)
        allp, counts2 = sh.allgather_points(pts, n)
        assert counts2 == counts and torch.equal(allp, want)
        # frame-parallel: every scan is owned by exactly one rank
        mine = sh.scans_for_rank(11, rank, world)
        owned = [torch.zeros(11, dtype=torch.int64) for _ in range(world)]
        flags = torch.zeros(11, dtype=torch.int64)
        flags[mine] = 1
        dist.all_gather(owned, flags)
        assert torch.equal(torch.stack(owned).sum(0), torch.ones(11, dtype=torch.int64))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_points_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_row_blocks_partition_the_frame():
    for H in (1, 7, 3000, 6144):
        for world in (1, 2, 3, 4, 8):
            blocks = [sh.row_block(H, r, world) for r in range(world)]
            assert blocks[0][0] == 0
            for (a0, an), (b0, _) in zip(blocks, blocks[1:]):
                assert a0 + an == b0
            assert blocks[-1][0] + blocks[-1][1] == H
            sizes = [n for _, n in blocks]
            assert max(sizes) - min(sizes) <= 1
