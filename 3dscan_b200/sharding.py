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
    exchange step is `gather_points`: an all-gather of the per-rank counts followed by
    point-to-point sends of the compacted points into rank `dst`'s buffer at the right offset
    (NCCL over NVLink on GPUs; the same code runs over gloo on CPU tensors in the tests).
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


def gather_points(points, count, dst=0, group=None, out=None, aligned_staging=None):
    """Concatenate the ranks' compacted point lists, in rank order, on rank `dst`.

    points: [capacity, C] tensor on this rank (only the first `count` rows are valid)
    count : int, number of valid rows on this rank
    Returns (all_points, counts) on rank `dst` ([sum(counts), C] tensor, list of ints) and
    (None, counts) elsewhere.  `out` may preallocate the destination on rank `dst`.
    aligned_staging: receive misaligned blocks through a staging block (default: on CUDA tensors).
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
                    stage = _staging(r, counts[r], tuple(points.shape[1:]), points.dtype, dev)[:counts[r]]
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


class PeerPointSink:
    """Row-sharded mode with the exchange folded into the reconstruction kernel.

    Rank `dst` owns `slots` staging blocks per other rank (`scan3d_peer_alloc`) and shares their
    CUDA IPC handles; every other rank maps its blocks (`scan3d_peer_open`) and hands the mapped
    pointer to its context
    (`scan3d_set_points_buffer`), so the fused kernel's IO warps stream the compacted points
    straight into `dst`'s memory over NVLink while the decode is still running -- no NCCL transfer
    afterwards, nothing that needs SMs beside the persistent kernel.  What is left per scan is
    `finish()`: an all-gather of the counts (the one host sync) and, on `dst`, one device copy per
    other rank that moves its block to its raster offset behind the lower ranks' points.

    Ordering: `finish(slot)` is stream-ordered after the rank's reconstruction; the all-gather
    completes on `dst` only after every rank's kernel has, and a kernel's peer writes are performed
    by the time it completes.  A block is written again `slots` scans later; `finish` makes the
    stream that runs the next all-gather wait for the copy-out of the block that scan will reuse.
    """

    def __init__(self, ctx, capacity_points, dst=0, group=None, slots=2):
        import importlib
        s3 = importlib.import_module("3dscan_b200")
        self._s3 = s3
        self.group, self.dst, self.slots = group, dst, slots
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.ctx = list(ctx) if isinstance(ctx, (list, tuple)) else [ctx] * slots
        self.device = torch.cuda.current_device()
        caps = [None] * self.world
        dist.all_gather_object(caps, int(capacity_points), group=group)
        self.caps = caps
        self.copied = [None] * slots          # events: block of slot s copied out on dst
        self.owned, self.opened, self.stage = [], [], {}
        payload = [None]
        if self.rank == dst:
            handles = {}
            for r in range(self.world):
                if r == dst:
                    continue
                blocks, hs = [], []
                for _ in range(slots):
                    ptr, h = s3.peer_alloc(self.device, caps[r] * 12)
                    self.owned.append(ptr)
                    blocks.append(s3.wrap_device(ptr, (caps[r], 3), "<f4"))
                    hs.append(h)
                self.stage[r], handles[r] = blocks, hs
            payload[0] = handles
        dist.broadcast_object_list(payload, src=dst, group=group)
        if self.rank != dst:
            for s, h in enumerate(payload[0][self.rank]):
                ptr = s3.peer_open(self.device, h)        # opened FROM this rank's device: lazy peer access
                self.opened.append(ptr)
                self.ctx[s].set_points_buffer(ptr, caps[self.rank])
        dist.barrier(group=group)

    def begin(self, slot):
        """Before enqueuing the reconstruction of a scan that uses block `slot` (needed when one
        context serves several slots; with one context per slot it changes nothing)."""
        if self.rank != self.dst:
            self.ctx[slot].set_points_buffer(self.opened[slot], self.caps[self.rank])

    def bind_output(self, slot, out):
        """dst == lowest rank only: its own points start at offset 0, so its context can write `out` directly."""
        if self.rank == self.dst and self.dst == 0:
            self.ctx[slot].set_points_buffer(out.data_ptr(), out.shape[0])

    def finish(self, slot, count, own_points, out):
        """All ranks, after their reconstruction of this scan was enqueued on the current stream.
        count: device-resident count tensor of this rank.  Returns (points, counts) on dst."""
        dev = count.device
        cur = torch.cuda.current_stream()
        nxt = (slot + 1) % self.slots
        if self.rank == self.dst and self.copied[nxt] is not None:
            cur.wait_event(self.copied[nxt])      # the scan that reuses block `nxt` starts after this all-gather
        counts_all = torch.zeros(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts_all, count.reshape(1).to(torch.int64), group=self.group)
        counts = [int(v) for v in counts_all.tolist()]
        if self.rank != self.dst:
            return None, counts
        offsets = [0]
        for n in counts:
            offsets.append(offsets[-1] + n)
        result = out[:offsets[-1]]
        own = result[offsets[self.dst]:offsets[self.dst + 1]]
        if counts[self.dst] and own.data_ptr() != own_points.data_ptr():
            own.copy_(own_points[:counts[self.dst]])
        for r, blocks in self.stage.items():
            if counts[r]:
                result[offsets[r]:offsets[r + 1]].copy_(blocks[slot][:counts[r]])
        ev = torch.cuda.Event()
        ev.record(cur)
        self.copied[slot] = ev
        return result, counts

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for c in set(self.ctx):
            try:
                c.set_points_buffer(None, 0)
            except Exception:
                pass
        for p in self.opened:
            self._s3.peer_close(self.device, p)
        dist.barrier(group=self.group)
        for p in self.owned:
            self._s3.peer_free(self.device, p)
        self.opened, self.owned, self.stage = [], [], {}
