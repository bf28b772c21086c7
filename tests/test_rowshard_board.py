"""CPU tests (3 processes, no GPU) of the row-shard gather protocol of include/scan3d_shard.h in its GPU-less mode:
counts -> base offsets -> pushes to the final raster offset -> slot reuse only after root's release -- with distinct
data per scan and no barrier between scans, i.e. the overlapped schedule bench.py runs on GPUs."""
import importlib
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _points(rank, scan, n):
    k = np.arange(n, dtype=np.float32)
    return np.stack([k + 1000 * rank, np.full(n, scan, np.float32), k * 0.5 + rank], axis=1)


def _count(rank, scan, world):
    return [0, 7, 1, 300, 50][(rank * 3 + scan) % 5] + (scan % 3 == 2 and rank == world - 1) * 11


def _worker(name, rank, world, scans, slots, slow_root, q):
    try:
        sh = importlib.import_module("3dscan_b200.sharding")
        g = sh.RowShardGroup(name, rank, world, -1, 4096, slots)
        ok = True
        for k in range(scans):
            s = k % slots
            mine = _points(rank, k, _count(rank, k, world))
            if rank == 0:
                g.output_host(s, len(mine))[:] = mine            # root's kernel writes its own points in place
            total, counts = g.gather_host(s, mine)
            assert counts == [_count(r, k, world) for r in range(world)], (k, counts)
            if rank == 0:
                if slow_root:
                    time.sleep(0.02)                             # a slow consumer: the others must not overwrite the cloud
                want = np.concatenate([_points(r, k, _count(r, k, world)) for r in range(world)])
                got = g.output_host(s, total).copy()
                ok = ok and total == len(want) and np.array_equal(got, want)
                g.release(s)
        g.close()
        q.put((rank, ok, None))
    except Exception as e:      # noqa: BLE001
        q.put((rank, False, repr(e)))


def _run(world, scans, slots, slow_root):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = "t%d_%d_%d_%d" % (os.getpid(), world, scans, int(slow_root))
    ps = [ctx.Process(target=_worker, args=(name, r, world, scans, slots, slow_root, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=30)
    for rank, ok, err in res:
        assert err is None, (rank, err)
        assert ok, rank


def test_rowshard_board_three_ranks_pipelined_slots():
    _run(world=3, scans=9, slots=2, slow_root=False)


def test_rowshard_board_slow_root_never_loses_a_cloud():
    _run(world=2, scans=7, slots=2, slow_root=True)


def test_rowshard_board_single_slot():
    _run(world=3, scans=4, slots=1, slow_root=True)
