"""Multi-GPU partitioning of the reconstruction path (one process per GPU, torch.distributed).

Two schemes, both named by BASELINE.json's north star:

  * frame-parallel batches (configs[3]): scans are independent, scan i goes to rank i % world.
    No collective touches the data path (`scans_for_rank`).

  * one very large frame, row-sharded (configs[4]): rank r owns a contiguous block of rows of
    every frame of the stack (`row_block`).  Everything on the path is per pixel except the ROI
    mask recurrence (3/wrapped_phase.cpp:266-279), whose closed form needs ROI rows y-2..y+1; the
    ROI plane (1 B/pixel) is simply replicated and each ctx is created with (row0, H_total), so no
    halo exchange is needed.  Each rank's compacted point list is in raster order of its rows, so
    concatenating the lists in rank order gives the raster order of the full frame.  The one real
    exchange step has two implementations: `RowShardGroup` (the default on GPUs: binding of the C++ host
    module include/scan3d_shard.h -- counts on a shared-memory board, one copy-engine push per rank over
    NVLink straight to the final raster offset, no kernel) and `gather_points` (an all-gather of the
    per-rank counts followed by point-to-point sends into rank `dst`'s buffer: NCCL over NVLink on GPUs,
    the same code over gloo on CPU tensors in the tests).
"""
import torch
import torch.distributed as dist


def scans_for_rank(n_scans, rank, world):
    """Indices of the scans rank `rank` reconstructs in a frame-parallel batch."""
    return list(range(rank, n_scans, world))


def row_block(H_total, rank, world):
    """(row0, rows) of the contiguous row block owned by `rank`; blocks differ by at most 1 row."""
    base, extra = divmod(H_total, world)
    rows = base + (1 if rank < extra else 0)
    row0 = rank * base + min(rank, extra)
    return row0, rows


def gather_points(points, count, dst=0, group=None, out=None, aligned_staging=None, slot=0):
    """Concatenate the ranks' compacted point lists, in rank order, on rank `dst`.

    points: [capacity, C] tensor on this rank (only the first `count` rows are valid)
    count : int, number of valid rows on this rank
    Returns (all_points, counts) on rank `dst` ([sum(counts), C] tensor, list of ints) and
    (None, counts) elsewhere.  `out` may preallocate the destination on rank `dst`.
    aligned_staging: receive misaligned blocks through a staging block (default: on CUDA tensors).
    slot: callers that overlap two gathers on two streams pass different slots: the staging blocks are kept per
    (peer, slot), so the receive of one gather never lands in a block the other is still copying out of.
    """
    if not dist.is_initialized():
        return points[:int(count)], [int(count)]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = points.device
    if torch.is_tensor(count):              # device-resident count (no host sync before the exchange)
        c = count.reshape(1).to(torch.int64)
    else:
        c = torch.tensor([int(count)], dtype=torch.int64, device=dev)
    counts_all = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_all, c, group=group)
    counts = [int(v) for v in counts_all.tolist()]     # the one host sync: sizes of the sends
    count = counts[rank]
    offsets = [0]
    for n in counts:
        offsets.append(offsets[-1] + n)
    if world == 1:
        return points[:count], counts
    ops = []
    result = None
    if rank == dst:
        total = offsets[-1]
        result = out[:total] if out is not None else torch.empty((total,) + tuple(points.shape[1:]), dtype=points.dtype, device=dev)
        if counts[dst]:
            result[offsets[dst]:offsets[dst + 1]].copy_(points[:counts[dst]])
        row_bytes = points.element_size() * (points[0].numel() if points.dim() > 1 else 1)
        staged = []
        use_staging = dev.type == "cuda" if aligned_staging is None else bool(aligned_staging)
        for r in range(world):
            if r != dst and counts[r]:
                target = result[offsets[r]:offsets[r + 1]]
                if use_staging and (offsets[r] * row_bytes) % 16:
                    # NCCL moves 16 bytes per thread; a receive address that is only 4-byte aligned
                    # (12 B points at an arbitrary offset) halves the transfer rate (measured 207 vs
                    # 440 GB/s).  Receive into an aligned staging block, then one device copy.
                    stage = _staging((r, slot), counts[r], tuple(points.shape[1:]), points.dtype, dev)[:counts[r]]
                    staged.append((target, stage))
                    target = stage
                ops.append(dist.P2POp(dist.irecv, target, r, group))
    elif count:
        ops.append(dist.P2POp(dist.isend, points[:count].contiguous(), dst, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank == dst:
        for target, stage in staged:
            target.copy_(stage)
    return result, counts


_stage_cache = {}


def _staging(peer, rows, row_shape, dtype, dev):
    """Per-peer receive block of at least `rows` rows (kept between calls, grown when a peer sends more)."""
    key = (peer, tuple(row_shape), dtype, str(dev))
    cur = _stage_cache.get(key)
    if cur is None or cur.shape[0] < rows:
        cur = torch.empty((max(rows, 2 * (cur.shape[0] if cur is not None else 0)),) + tuple(row_shape), dtype=dtype, device=dev)
        _stage_cache[key] = cur
    return cur


def allgather_points(points, count, group=None):
    """Every rank receives the full, rank-ordered point list (padded all-gather)."""
    world = dist.get_world_size(group)
    dev = points.device
    c = torch.tensor([int(count)], dtype=torch.int64, device=dev)
    counts_t = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts_t, c, group=group)
    counts = [int(t.item()) for t in counts_t]
    cap = max(max(counts), 1)
    padded = torch.zeros((cap,) + tuple(points.shape[1:]), dtype=points.dtype, device=dev)
    padded[:count].copy_(points[:count])
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, counts)]), counts


class RowShardGroup:
    """Binding of include/scan3d_shard.h: the row-sharded mode's gather of every rank's compacted points on rank 0.

    The protocol lives in the C++ host module (3dscan_b200/csrc/scan3d_shard.cu): per-scan control words on a POSIX
    shared-memory board, root's output blocks shared through CUDA IPC, ONE copy-engine push per rank and scan over
    NVLink straight to the points' final raster offset -- no kernel, so the persistent reconstruction kernel of the
    next scan runs undisturbed.  A block is reused every `slots` scans; the ranks push into it only after root has
    released the previous cloud (`release`), so overlapped schedules are race-free by construction.

    name: identical on every rank, unique per job (e.g. derived from MASTER_PORT).  device < 0 selects the GPU-less
    mode (blocks in shared memory, points from host arrays): the CPU tests of the ordering logic use it."""

    def __init__(self, name, rank, world, device, capacity_points, slots=2):
        import ctypes as C
        import importlib
        s3 = importlib.import_module("3dscan_b200")
        self._C, self._s3 = C, s3
        self.L = s3.cuda_lib()
        L = self.L
        L.scan3d_shard_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]
        L.scan3d_shard_destroy.argtypes = [C.c_void_p]
        L.scan3d_shard_last_error.argtypes = [C.c_void_p]
        L.scan3d_shard_last_error.restype = C.c_char_p
        L.scan3d_shard_bind.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.scan3d_shard_gather.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.scan3d_shard_gather_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.scan3d_shard_output.argtypes = [C.c_void_p, C.c_int]
        L.scan3d_shard_output.restype = C.c_void_p
        L.scan3d_shard_release.argtypes = [C.c_void_p, C.c_int]
        self.rank, self.world, self.slots, self.capacity, self.device = rank, world, slots, int(capacity_points), device
        h = C.c_void_p()
        rc = L.scan3d_shard_create(str(name).encode(), rank, world, device, int(capacity_points), slots, C.byref(h))
        if rc:
            raise RuntimeError("scan3d_shard_create: " + (L.scan3d_shard_last_error(None) or b"").decode())
        self.h = h

    def _ck(self, rc, what):
        if rc:
            raise RuntimeError(what + ": " + (self.L.scan3d_shard_last_error(self.h) or b"").decode())

    def bind(self, slot, ctx):
        self._ck(self.L.scan3d_shard_bind(self.h, slot, ctx.h), "scan3d_shard_bind")

    def gather(self, slot, ctx):
        """After ctx.reconstruct_dev(...) of this rank's rows was enqueued.  Returns (total, counts)."""
        C = self._C
        total, counts = C.c_int64(), (C.c_int64 * self.world)()
        self._ck(self.L.scan3d_shard_gather(self.h, slot, ctx.h, C.byref(total), counts), "scan3d_shard_gather")
        return total.value, list(counts)

    def gather_host(self, slot, points):
        """GPU-less mode: points = this rank's float32 [n][3] array."""
        import numpy as np
        C = self._C
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        total, counts = C.c_int64(), (C.c_int64 * self.world)()
        self._ck(self.L.scan3d_shard_gather_host(self.h, slot, pts.ctypes.data_as(C.c_void_p), pts.shape[0], C.byref(total), counts),
                 "scan3d_shard_gather_host")
        return total.value, list(counts)

    def output_ptr(self, slot):
        return self.L.scan3d_shard_output(self.h, slot)

    def output_host(self, slot, n):
        """GPU-less mode, root: the first n gathered points of the slot as a numpy view."""
        import numpy as np
        C = self._C
        p = self.output_ptr(slot)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(int(n), 3))

    def release(self, slot):
        self._ck(self.L.scan3d_shard_release(self.h, slot), "scan3d_shard_release")

    def close(self):
        if self.h:
            self.L.scan3d_shard_destroy(self.h)
            self.h = None
