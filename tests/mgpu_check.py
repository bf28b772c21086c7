#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, NCCL): the row-sharded reconstruction gathered on rank 0 is byte-identical to
the single-GPU reconstruction of the same frame -- through the NCCL send/recv gather and through the C++ row-shard
group (include/scan3d_shard.h), the latter in exactly bench.py's overlapped schedule: SIX DISTINCT scans pipelined
over two contexts / two streams / two output slots with no barrier between them, every gathered cloud compared.
Usage:
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
N_SCANS = 6


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    W, H, PW, PH = 2048, 1537, 2048, 1536
    cal, ocal, _ = calibs(W / 1600.0, PW / 1280.0)
    full = s3.make_config(W, H, PW, PH, 8, 10, 10, 2, 2, 2)
    row0, rows = sh.row_block(H, rank, world)
    cfg = s3.make_config(W, rows, PW, PH, 8, 10, 10, 2, 2, 2, row0=row0, H_total=H)
    # distinct scans: every rank renders the same frames (deterministic), keeps its rows; rank 0 also keeps the
    # single-GPU result of every whole frame
    stacks, rois, want = [], [], []
    ref = s3.Scan3D(full, lr, cal) if rank == 0 else None
    for k in range(N_SCANS):
        stack, roi = s3.synth_stack(full, cal, s3.default_synth_params(seed=0x3D5CA9 + k, roi_fraction=0.75 - 0.07 * k))
        if rank == 0:
            ref.reconstruct(stack, roi)
            want.append(ref.points().copy())
        stacks.append(torch.from_numpy(np.ascontiguousarray(stack[:, row0:row0 + rows])).cuda())
        rois.append(torch.from_numpy(roi).cuda())
        del stack
    if ref is not None:
        ref.close()

    # ---- 1. NCCL gather (all-gather of the counts + send/recv), scan 0
    ctx = s3.Scan3D(cfg, lr, cal)
    ctx.reconstruct_dev(stacks[0].data_ptr(), rois[0].data_ptr())
    n = ctx.point_count()
    pts = torch.from_numpy(ctx.points()).cuda()
    got, counts = sh.gather_points(pts, n, dst=0)
    ok = True
    if rank == 0:
        ok = sum(counts) == len(want[0]) and np.array_equal(got.cpu().numpy().view(np.uint32), want[0].view(np.uint32))
        print("row-shard over %d GPUs, NCCL gather: %d points, counts %s, identical to 1 GPU: %s" % (world, sum(counts), counts, ok))
    ctx.close()

    # ---- 2. the row-shard group in bench.py's schedule: decode(k+1) is enqueued before gather(k); no barrier anywhere
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    ctxs = [s3.Scan3D(cfg, lr, cal, stream=st.cuda_stream) for st in streams]
    group = sh.RowShardGroup("chk%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.getppid()), rank, world, lr, H * W, slots=2)
    for i in range(2):
        group.bind(i, ctxs[i])

    def decode(k):
        ctxs[k & 1].reconstruct_dev(stacks[k].data_ptr(), rois[k].data_ptr())

    ok2 = True
    decode(0)
    for k in range(N_SCANS):
        if k + 1 < N_SCANS:
            decode(k + 1)
        total, counts = group.gather(k & 1, ctxs[k & 1])
        if rank == 0:
            cloud = s3.wrap_device(group.output_ptr(k & 1), (total, 3), "<f4")
            host = np.asarray(torch.as_tensor(cloud, device="cuda").cpu())
            same = total == len(want[k]) and np.array_equal(host.view(np.uint32), want[k].view(np.uint32))
            ok2 = ok2 and same
            if not same:
                print("  scan %d: %d points (want %d), counts %s: MISMATCH" % (k, total, len(want[k]), counts))
        group.release(k & 1)
    if rank == 0:
        print("row-shard group (shared-memory board + copy-engine pushes over NVLink), %d distinct pipelined scans: "
              "every cloud identical to 1 GPU: %s" % (N_SCANS, ok2))
    ok = ok and ok2
    torch.cuda.synchronize()
    group.close()
    for c in ctxs:
        c.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
