"""Diagnostics (not a test): where does the row-shard step spend its time?  torchrun, 2 GPUs."""
import os, sys, time, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):  # run from the repo root
    sys.path.insert(0, p)
import numpy as np
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
s3 = importlib.import_module("3dscan_b200")
sh = importlib.import_module("3dscan_b200.sharding")
from helpers import load_calib_c1, scaled_calib
import bench
W, Ht, N, M, fw = 8192, 6144, 8, 10, 8
_c = scaled_calib(load_calib_c1(), W / 1600.0, W / 1280.0)
cal = s3.make_calib(*[_c[k] for k in ("Kc", "dc", "Kp", "dp", "rc", "tc", "rp", "tp")])
row0, rows = sh.row_block(Ht, rank, world)
cfg = s3.make_config(W, rows, W, Ht, N, M, M, fw, fw, 2, row0=row0, H_total=Ht, flags=s3.FLAG_FAST_TRIANGULATION)
nf = s3.stack_planes(cfg)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
torch.cuda.set_stream(streams[0])
ctxs = [s3.Scan3D(cfg, lr, cal, stream=st.cuda_stream) for st in streams]
stack_h = torch.empty((nf, rows, W), dtype=torch.uint8, pin_memory=True)
roi_h = torch.empty((Ht, W), dtype=torch.uint8, pin_memory=True)
s3.synth_stack(cfg, cal, s3.default_synth_params(seed=0x3D5CA9), out=stack_h.numpy(), roi_out=roi_h.numpy(), threads=12)
stack_d, roi_d = stack_h.to("cuda"), roi_h.to("cuda")
outs = [torch.empty((Ht * W, 3), dtype=torch.float32, device="cuda") if rank == 0 else None for _ in ctxs]
srcs = [bench._wrap_device(torch, c.device_points(), (rows * W, 3), "<f4") for c in ctxs]
cnts = [bench._wrap_device(torch, c.device_point_count(), (1,), "<i4") for c in ctxs]
def decode(k): ctxs[k & 1].reconstruct_dev(stack_d.data_ptr(), roi_d.data_ptr())
def gather(k):
    with torch.cuda.stream(streams[k & 1]):
        sh.gather_points(srcs[k & 1], cnts[k & 1], dst=0, out=outs[k & 1])
def timed(name, fn, reps=10):
    for i in range(4): fn(i)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = time.time()
    for i in range(reps): fn(i)
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.time() - t) / reps
    if rank == 0: print("%-58s %.3f ms per scan" % (name, dt * 1e3), flush=True)
timed("decode only, one ctx", lambda i: decode(0))
timed("decode only, alternating ctxs/streams", lambda i: decode(i))
timed("decode + gather, one ctx (serial)", lambda i: (decode(0), gather(0)))
timed("gather only (same points again)", lambda i: gather(0))
timed("decode(i+1) enqueued, then gather(i)  [bench pipeline]", lambda i: (decode(i + 1), gather(i)))
dist.destroy_process_group()
