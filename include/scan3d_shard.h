/*
 * scan3d_shard.h -- C ABI of the row-sharded mode (BASELINE.json configs[4]: ONE very large frame over the GPUs of
 * one node, one process per GPU).  No reference counterpart: the reference is single-process; the contract kept is
 * the order of its point cloud -- valid pixels in row-major raster order (8/save_point_cloud.cpp:85-136).
 *
 * Rank r owns a contiguous block of rows (every frame of the stack, rows [row0, row0 + H)) and a ctx created with
 * (row0, H_total); the ROI plane is replicated.  Each rank's compacted points are in raster order of its rows, so
 * the frame's cloud is the ranks' lists concatenated in rank order.  This module delivers that concatenation to
 * rank 0 ("root"):
 *
 *   - root owns `slots` output blocks of capacity_points x 12 bytes and shares them through CUDA IPC; root's own
 *     context writes its points straight into the block (they start at offset 0);
 *   - per scan every rank publishes its point count in a small POSIX shared-memory board; rank r > 0 then knows its
 *     base offset (the sum of the lower ranks' counts) and pushes its points with ONE copy-engine transfer over
 *     NVLink to their final place in root's block -- no kernel, no SM: the next scan's persistent reconstruction
 *     kernel (which leaves no SM to a collective's kernel) runs undisturbed on another stream;
 *   - a block is reused every `slots` scans: the ranks push scan k + slots into it only after root has released
 *     scan k's cloud (scan3d_shard_release) -- the board carries that edge, so the schedule is race-free by
 *     construction whatever the caller overlaps.
 *
 * Calls are collective: every rank calls create / gather / destroy in the same order with the same slot.
 * All entries return 0 or a negative scan3d_status; scan3d_shard_last_error gives the message.
 */
#ifndef SCAN3D_SHARD_H
#define SCAN3D_SHARD_H

#include "scan3d.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct scan3d_shard scan3d_shard;

/* name: the same string on every rank, unique per job and node (it names the shared-memory board).
 * capacity_points: W * H_total.  device: this rank's GPU.  Blocks until every rank has joined (60 s). */
int scan3d_shard_create(const char *name, int rank, int world, int device, int64_t capacity_points, int slots,
                        scan3d_shard **out);
int scan3d_shard_destroy(scan3d_shard *sh);
const char *scan3d_shard_last_error(const scan3d_shard *sh);   /* sh may be NULL: last create error */

/* Root only (no-op elsewhere): ctx writes the points of its next reconstructions straight into the slot's block.
 * Call before the reconstruction whose points go to that slot -- and enqueue that reconstruction only after the slot's
 * previous cloud has been released (root's own kernel is the first writer of the next cloud). */
int scan3d_shard_bind(scan3d_shard *sh, int slot, scan3d_ctx *ctx);

/* Every rank, after scan3d_reconstruct_dev(ctx, ...) of its row block was enqueued: waits for that reconstruction
 * (ctx's stream), exchanges the counts, pushes the points (rank > 0), and on root waits until every rank's points
 * have landed.  counts (optional) receives the world counts; *total_points their sum.  For overlap enqueue the
 * NEXT scan's reconstruction (another ctx, another stream) before calling this. */
int scan3d_shard_gather(scan3d_shard *sh, int slot, scan3d_ctx *ctx, int64_t *total_points, int64_t *counts);

/* The same protocol with no GPU anywhere (group created with device < 0: the blocks live in shared memory); the rank's
 * points come from host memory, root copies its own to scan3d_shard_output(slot) before the call.  For CPU-only tests
 * of the ordering logic (counts, base offsets, slot reuse); not a compute path. */
int scan3d_shard_gather_host(scan3d_shard *sh, int slot, const float *points_host, int64_t count, int64_t *total_points,
                             int64_t *counts);

/* Root: device pointer of the slot's gathered cloud, f32 [total][3] in raster order.  NULL elsewhere. */
void *scan3d_shard_output(scan3d_shard *sh, int slot);

/* Root: the slot's cloud has been consumed (everything the caller enqueued on ctx's stream so far has completed is
 * the caller's business: release AFTER synchronising with its consumer).  The ranks may then push the scan that
 * reuses the slot.  No-op elsewhere. */
int scan3d_shard_release(scan3d_shard *sh, int slot);

#ifdef __cplusplus
}
#endif
#endif
