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
    # same frame again with the exchange folded into the kernel: the other ranks' IO warps write their
    # points into rank 0's memory over NVLink (PeerPointSink), rank 0 only moves the blocks into place
    import bench
    sink = sh.PeerPointSink(ctx, rows * W, dst=0, slots=2)
    out = torch.empty((H * W, 3), dtype=torch.float32, device="cuda") if rank == 0 else None
    cnt = bench._wrap_device(torch, ctx.device_point_count(), (1,), "<i4")
    st = torch.cuda.Stream()                               # (stream handle 0 would mean "ctx-owned stream")
    torch.cuda.set_stream(st)
    ctx.set_stream(st.cuda_stream)
    stack_d = torch.from_numpy(np.ascontiguousarray(stack[:, row0:row0 + rows])).cuda()
    roi_d = torch.from_numpy(roi).cuda()
    ok2 = True
    for rep in range(3):                                   # slot 0, 1, 0: exercises block reuse
        if rank == 0:
            out.zero_()
        torch.cuda.synchronize(); dist.barrier()
        sink.begin(rep % 2)
        ctx.reconstruct_dev(stack_d.data_ptr(), roi_d.data_ptr())
        own = bench._wrap_device(torch, ctx.device_points(), (rows * W, 3), "<f4") if rank == 0 else None
        got2, counts2 = sink.finish(rep % 2, cnt, own, out)
        torch.cuda.synchronize()
        if rank == 0:
            same = sum(counts2) == n_ref and np.array_equal(got2.cpu().numpy().view(np.uint32), want.view(np.uint32))
            ok2 = ok2 and same
    if rank == 0:
        print("in-kernel NVLink point stream (PeerPointSink), 3 scans: identical to 1 GPU: %s" % ok2)
    ok = ok and ok2
    sink.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
