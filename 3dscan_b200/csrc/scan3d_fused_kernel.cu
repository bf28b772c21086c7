// scan3d_fused_kernel.cu -- FIRST GENERATION of the fused kernel (selected with SCAN3D_FUSED_IMPL=6, kept as the
// measured baseline of DESIGN.md section 7; the default is k_fused7 in scan3d_fused_kernel7.cu) plus the work-list
// kernels both generations share.  ONE persistent kernel that takes the captured pattern
// stack to absolute phase maps, fringe orders, c_p_map, validity and the raster-ordered compacted
// point cloud, reading every input byte once and writing every contract output once.
//
// Replaces, per scan, the reference call sequence compute_wrapped_phase(0/1) -> unwrap_phase(0/1)
// -> compute_c_p_map() -> triangulate() -> save_point_cloud()'s gather
// (M_tech_project_console/m_tech_project_console.cpp:372-401).
//
// Shape of the kernel (sm_100a):
//   * persistent grid (SMs x CTAs/SM), static round-robin over LINEAR tiles of T = 128*CW pixels
//     of the row-major frame (tile order == raster order, which the compaction relies on; every
//     frame segment of a tile is one contiguous byte range even when it crosses a row end);
//   * warp-specialised CTA of CW+2 warps:
//       - PRODUCER warp: moves each tile's NF frame segments and a 4-row ROI window HBM -> shared
//         memory with cp.async.bulk (TMA bulk copies) completing on an mbarrier ring of S slots;
//       - CW CONSUMER warps: (1) SWAR integer phase on 4 consecutive pixels per thread straight
//         from shared memory (mask-recurrence closed form, phase-shift numerators/denominators,
//         Gray threshold, packed fringe orders); (2) FP64 phase per pixel in the reference's exact
//         IEEE operation order (atan2 -> +Pi -> +code*2Pi -> lrint correspondence -> undistorted
//         pixels -> 4x3 normal-equation solve); results are staged in the tile's own (now dead)
//         shared-memory slot.  One CTA-local barrier per tile; consumers never wait on other CTAs;
//       - EPILOGUE warp: drains each staged slot: TMA bulk stores of the plane outputs, then the
//         raster-order compaction of the tile's points (count -> decoupled look-back across tiles,
//         epoch-tagged so no per-launch reset -> scatter), then hands the slot back to the producer.
#include <stdio.h>
#include <stdlib.h>

#include "scan3d_fused_common.cuh"

namespace s3d {

constexpr int MAX_STAGES = 4;
constexpr int SMEM_FIXED = 3 * MAX_STAGES * 8 + ATAN_TAB_DOUBLES * 8 + MAX_STAGES * 4 + MAX_STAGES * 8 * 4 + 16;   // barriers + atan table + slot tile ids + per-slot warp counts + pad

struct TileGeom {
    int T;            // pixels per tile
    int roi_row;      // T + 2*halo
    int roi_bytes;    // 4 rows
    int out_unwv, out_codev, out_valid, out1_bytes, out_unwh, out_codeh, out_cp, out_x, out_cx, out2_bytes;
};
__host__ __device__ constexpr TileGeom geom_of(int T)
{
    TileGeom g{};
    g.T = T;
    g.roi_row = T + 2 * ROI_HALO;
    g.roi_bytes = 4 * g.roi_row;
    g.out_unwv = 0;
    g.out_codev = g.out_unwv + 4 * T;
    g.out_valid = g.out_codev + 2 * T;
    g.out1_bytes = g.out_valid + T;
    g.out_unwh = g.out1_bytes;
    g.out_codeh = g.out_unwh + 4 * T;
    g.out_cp = g.out_codeh + 2 * T;
    g.out_x = g.out_cp + 8 * T;
    g.out_cx = g.out_x + 12 * T;            // the tile's points, compacted in raster order
    g.out2_bytes = g.out_cx + 12 * T;
    return g;
}

static int num_frames(const scan3d_config& c)
{
    return c.dirs == 2 ? 2 * c.N + 2 * (c.M_v + c.M_h) : c.N + 2 * c.M_v;
}
static int stage_bytes_of(const scan3d_config& c, int T)
{
    const TileGeom g = geom_of(T);
    const int in_b = num_frames(c) * T;
    const int out_b = c.dirs == 2 ? g.out2_bytes : g.out1_bytes;
    return (in_b > out_b ? in_b : out_b) + g.roi_bytes;
}

struct FusedPlan {
    int cw, minb, stages, stage_bytes;
    size_t smem;
};

// Launch shape: consumer warps per CTA, CTAs per SM, pipeline slots.  Default (6 consumer warps,
// 2 CTAs/SM, 2 slots = 12 consumer warps per SM at 128 registers) picked from the sweep recorded in
// profiles/; SCAN3D_FUSED_CFG="cw,ctas_per_sm,stages" overrides it for tuning runs.
static bool plan_for(const scan3d_config& c, FusedPlan* out)
{
    if (c.W % 16 != 0) return false;
    if (!(c.N == 3 || c.N == 4 || c.N == 5 || c.N == 8)) return false;
    int cw = 6, minb = 2, stages = 2;
    if (const char* e = getenv("SCAN3D_FUSED_CFG")) {
        int a = 0, b = 0, s = 0;
        if (sscanf(e, "%d,%d,%d", &a, &b, &s) == 3) { cw = a; minb = b; stages = s; }
    }
    if (stages < 2) stages = 2;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    // instantiated launch shapes (consumer warps, CTAs/SM): the requested one first, then the
    // rest from most to least parallel; first one whose slots fit in shared memory wins
    const int shapes[4][2] = {{cw, minb}, {6, 2}, {4, 2}, {8, 1}};
    for (int i = 0; i < 4; i++) {
        const int w = shapes[i][0], b = shapes[i][1];
        if (!((w == 6 && b == 2) || (w == 4 && b == 2) || (w == 8 && b == 1))) continue;
        for (int st = stages; st >= 2; st--) {
            const int sb = stage_bytes_of(c, 128 * w);
            // slots + fixed area + (both directions) the epilogue's private copy of one tile's points
            const size_t smem = (size_t)st * sb + SMEM_FIXED + (c.dirs == 2 ? 12 * 128 * w : 0);
            if ((smem + 1024) * b <= (size_t)SMEM_MAX + 1024) {   // 1 KB per CTA is reserved by the runtime
                out->cw = w; out->minb = b; out->stages = st; out->stage_bytes = sb; out->smem = smem;
                return true;
            }
        }
    }
    return false;
}

bool fused_supported(const scan3d_config& c, int* stages_out, size_t* smem_out)
{
    FusedPlan p;
    if (!plan_for(c, &p)) return false;
    if (stages_out) *stages_out = p.stages;
    if (smem_out) *smem_out = p.smem;
    return true;
}

int fused_num_tiles(const scan3d_config& c)
{
    // upper bound over every tile size the planner can choose (smallest tile = 256 px)
    return (int)(((size_t)c.W * c.H + 255) / 256) + 1;
}

// ---- work list: tiles that contain at least one ROI pixel ---------------------------------------
// One warp per tile looks at the tile's ROI bytes.  Tiles without any selected pixel never enter
// the main kernel: their outputs (phase 0, fringe order -1, valid 0, c_p_map 0) are written right
// here and their 56 input frame segments are never read.  The flags are then compacted, in raster
// order, into the work list the persistent kernel walks in phase-aligned rounds.
__global__ void k_tile_flags(const uint8_t* __restrict__ roi, int T, int plane, int n_tiles, size_t roi_off,
                             uint8_t* __restrict__ flags, float* __restrict__ unw_v, float* __restrict__ unw_h,
                             int16_t* __restrict__ code_v, int16_t* __restrict__ code_h, uint8_t* __restrict__ valid,
                             int2* __restrict__ cpmap, int dirs)
{
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= n_tiles) return;
    const int p0 = t * T, wt = min(T, plane - p0);
    const uint4* r = reinterpret_cast<const uint4*>(roi + roi_off + p0);
    bool any = false;
    for (int i = lane; i < wt / 16; i += 32) {
        const uint4 v = r[i];
        any |= (v.x | v.y | v.z | v.w) != 0;
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) flags[t] = any ? 1 : 0;
    if (any) return;
    const uint4 z = make_uint4(0, 0, 0, 0), m1 = make_uint4(~0u, ~0u, ~0u, ~0u);
    for (int i = lane; i < wt / 4; i += 32) {       // 4 floats
        reinterpret_cast<uint4*>(unw_v + p0)[i] = z;
        if (dirs == 2) reinterpret_cast<uint4*>(unw_h + p0)[i] = z;
    }
    for (int i = lane; i < wt / 8; i += 32) {       // 8 int16 = -1
        reinterpret_cast<uint4*>(code_v + p0)[i] = m1;
        if (dirs == 2) reinterpret_cast<uint4*>(code_h + p0)[i] = m1;
    }
    for (int i = lane; i < wt / 16; i += 32) reinterpret_cast<uint4*>(valid + p0)[i] = z;
    if (dirs == 2)
        for (int i = lane; i < wt / 2; i += 32) reinterpret_cast<uint4*>(cpmap + p0)[i] = z;
}

// exclusive scan of the flags by one CTA -> list of non-empty tile ids (raster order) + its length.
// Each thread owns a contiguous chunk of flags, so one block-wide scan suffices.
__global__ void k_tile_list(const uint8_t* __restrict__ flags, int n_tiles, int* __restrict__ list,
                            int* __restrict__ n_list, uint32_t* __restrict__ d_count)
{
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int per = (n_tiles + 1023) / 1024;
    const int i0 = threadIdx.x * per, i1 = min(n_tiles, i0 + per);
    int cnt = 0;
    for (int i = i0; i < i1; i++) cnt += flags[i] != 0;
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        int v = wsum[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        wsum[lane] = v;   // inclusive over warps
    }
    __syncthreads();
    int pos = (w ? wsum[w - 1] : 0) + incl - cnt;
    for (int i = i0; i < i1; i++)
        if (flags[i]) list[pos++] = i;
    if (threadIdx.x == 0) {
        *n_list = wsum[31];
        n_list[n_tiles + 8] = 0;   // v7 dynamic scheduler: next work-list position
        *d_count = 0;   // overwritten by the fused kernel's last tile when there is any work
    }
}

cudaError_t launch_worklist(const FusedArgs& a, int T, int dirs, cudaStream_t st)
{
    k_tile_flags<<<(a.n_tiles + 7) / 8, 256, 0, st>>>(a.roi, T, a.W * a.H, a.n_tiles, (size_t)a.row0 * a.W, a.tile_flags,
                                                      a.unw_v, a.unw_h, a.code_v, a.code_h, a.valid, a.cpmap, dirs);
    k_tile_list<<<1, 1024, 0, st>>>(a.tile_flags, a.n_tiles, a.tile_list, a.n_list, a.d_count);
    return cudaGetLastError();
}

// ---- the kernel --------------------------------------------------------------------------------
// registers per thread for a launch shape: the whole register file split over MINB resident CTAs
constexpr int regs_for(int cw, int minb)
{
    const int r = 65536 / (minb * (cw + 2) * 32);
    return r > 255 ? 255 : (r / 8) * 8;
}

template <int N, int DIRS, int CW, int MINB, bool EXACT>
__global__ void __launch_bounds__((CW + 2) * 32) __maxnreg__(regs_for(CW, MINB))
k_fused(const __grid_constant__ FusedArgs a, const __grid_constant__ DeviceCalib cal, const int stages,
        const int stage_bytes)
{
    constexpr int T = 128 * CW;             // pixels per tile
    constexpr int NCONS = 32 * CW;          // consumer threads, 4 pixels each
    constexpr int WPF = T / 4;              // 32-bit words per frame segment
    constexpr TileGeom G = geom_of(T);

    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
    double* tab = reinterpret_cast<double*>(bars + 3 * MAX_STAGES);
    volatile int* slot_tile = reinterpret_cast<volatile int*>(tab + ATAN_TAB_DOUBLES);   // tile held by each slot (-1: no more work)
    volatile uint32_t* slot_cnt = reinterpret_cast<volatile uint32_t*>(const_cast<int*>(slot_tile) + MAX_STAGES);   // [slot][warp] valid points
    // epilogue-private buffer for one tile's compacted points (lets the slot go back to the
    // producer before the look-back has resolved); 16-byte aligned right after the fixed area
    float* xbuf = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes + SMEM_FIXED);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + MAX_STAGES),
                   bar_staged = smem_u32(bars + 2 * MAX_STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NF = DIRS == 2 ? 2 * N + 2 * (a.M_v + a.M_h) : N + 2 * a.M_v;
    const int W = a.W;
    const int plane = W * a.H;              // local pixels (validated < 2^31)
    constexpr bool fastdiv = true;   // the host verified the exact-quotient shortcut (else this kernel is not used)

    if (tid == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
            mbar_init(bar_staged + 8 * s, CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    for (int i = tid; i < ATAN_TAB_DOUBLES; i += (CW + 2) * 32) tab[i] = a.atan_tab[i];
    __syncthreads();

    // Static, phase-aligned schedule over the work list (tiles holding ROI pixels, raster order):
    // in round r the grid works on list entries [r*G, (r+1)*G), so the tiles a look-back has to wait
    // for are always being computed at the same time (a dynamic scheduler de-phases the CTAs and,
    // with a 2-slot pipeline, turns the in-order look-back into a convoy: measured 4x slower).
    // The producer publishes each slot's list position; -1 ends every role's loop.  The look-back
    // chain runs over list positions; `tile` below is the pixel tile the entry points to.
    const int n_work = *a.n_list;
    if (warp == CW) {
        // ================================ PRODUCER ================================
        const long long roi_total = (long long)W * a.H_total;
        for (int it = 0;; it++) {
            const int s = it % stages, ph = (it / stages) & 1;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            if (lane == 0) trace(a.trace, it, 0);
            const int pos = it * (int)gridDim.x + (int)blockIdx.x;
            if (pos >= n_work) {
                if (lane == 0) {
                    slot_tile[s] = -1;
                    mbar_arrive(bar_full + 8 * s);
                }
                break;
            }
            const int tile = a.tile_list[pos];
            const int p0 = tile * T, wt = min(T, plane - p0);
            // ROI window: 4 linear segments, one per row offset dr = -2..+1, each covering the
            // tile's pixels shifted by dr rows plus a 16-byte halo on both sides
            const long long gbase = (long long)a.row0 * W + p0 - ROI_HALO;
            long long seg0 = 0, seg1 = 0;
            uint32_t roi_tx = 0;
            if (lane < 4) {
                seg0 = max(gbase + (long long)(lane - 2) * W, 0LL);
                seg1 = min(gbase + (long long)(lane - 2) * W + wt + 2 * ROI_HALO, roi_total);
                if (seg1 > seg0) roi_tx = (uint32_t)(seg1 - seg0);
            }
            uint32_t roi_sum = roi_tx;
            roi_sum += __shfl_xor_sync(0xffffffffu, roi_sum, 1);
            roi_sum += __shfl_xor_sync(0xffffffffu, roi_sum, 2);
            roi_sum = __shfl_sync(0xffffffffu, roi_sum, 0);
            const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
            if (lane == 0) {
                slot_tile[s] = pos;
                mbar_expect_tx(bar_full + 8 * s, (uint32_t)NF * wt + roi_sum);
            }
            __syncwarp();
            const uint8_t* src = a.stack + p0;
            for (int f = lane; f < NF; f += 32)
                bulk_g2s(dst + f * T, src + (size_t)f * plane, (uint32_t)wt, bar_full + 8 * s);
            if (roi_tx)
                bulk_g2s(dst + (stage_bytes - G.roi_bytes) + lane * G.roi_row +
                             (uint32_t)(seg0 - (gbase + (long long)(lane - 2) * W)),
                         a.roi + seg0, roi_tx, bar_full + 8 * s);
            if (lane == 0) trace(a.trace, it, 1);
        }
        return;
    }

    if (warp == CW + 1) {
        // ================================ EPILOGUE ================================
        const unsigned long long tag = (unsigned long long)(a.epoch & 0x3fffffffu) << 34;
        for (int it = 0;; it++) {
            const int s = it % stages, ph = (it / stages) & 1;
            mbar_wait_epilogue(bar_staged + 8 * s, ph);
            if (lane == 0) trace(a.trace, it, 5);
            const int tile = slot_tile[s];           // position in the work list == look-back index
            if (tile < 0) break;
            const int p0 = a.tile_list[tile] * T, wt = min(T, plane - p0);
            uint8_t* stage = smem + (size_t)s * stage_bytes;
            if (lane == 0) {
                bulk_s2g(a.unw_v + p0, smem_u32(stage + G.out_unwv), 4u * wt);
                bulk_s2g(a.code_v + p0, smem_u32(stage + G.out_codev), 2u * wt);
                bulk_s2g(a.valid + p0, smem_u32(stage + G.out_valid), 1u * wt);
                if (DIRS == 2) {
                    bulk_s2g(a.unw_h + p0, smem_u32(stage + G.out_unwh), 4u * wt);
                    bulk_s2g(a.code_h + p0, smem_u32(stage + G.out_codeh), 2u * wt);
                    bulk_s2g(a.cpmap + p0, smem_u32(stage + G.out_cp), 8u * wt);
                }
                bulk_commit();
            }
            if (DIRS == 2) {
                // ---- tile total: the consumers left one count per warp ----
                const uint32_t* vw = reinterpret_cast<const uint32_t*>(stage + G.out_valid);
                uint32_t total = lane < CW ? slot_cnt[s * 8 + lane] : 0u;
#pragma unroll
                for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                // ---- publish the aggregate at once, take a private copy of the points and hand the
                //      slot back: the look-back below then overlaps the next tile's compute ----
                const bool last = tile == n_work - 1;
                const bool resolve = total != 0 || last || tile == 0;
                if (tile > 0 && resolve && lane == 0) st_state(a.tile_state + tile, tag | (1ull << 32) | total);
                if (total) {
                    const float4* cx4 = reinterpret_cast<const float4*>(stage + G.out_cx);
                    float4* xb4 = reinterpret_cast<float4*>(xbuf);
                    const int nv = (3 * (int)total + 3) >> 2;
                    for (int v = lane; v < nv; v += 32) xb4[v] = cx4[v];
                }
                __syncwarp();
                const bool extras = a.pix != nullptr || a.rgb != nullptr;   // they need the slot's flags later
                if (!extras && lane == 0) {
                    bulk_wait_read();                    // the TMA stores have read the slot
                    trace(a.trace, it, 7);
                    mbar_arrive(bar_empty + 8 * s);      // hand it back to the producer
                }
                // ---- decoupled look-back: exclusive prefix of the valid counts of tiles < tile ----
                uint32_t excl = 0;
                if (!resolve) {
                    // an empty tile has nothing to write: it never waits.  If its predecessor's
                    // inclusive prefix happens to be there it forwards it, else it posts aggregate 0.
                    if (lane == 0) {
                        const unsigned long long w = ld_state(a.tile_state + tile - 1);
                        const bool fwd = (w >> 34) == (tag >> 34) && ((w >> 32) & 3ull) == 2;
                        st_state(a.tile_state + tile, fwd ? w : (tag | (1ull << 32)));
                    }
                } else {
                    if (tile > 0) {
                        int look = tile - 1;
                        while (true) {
                            const int idx = look - lane;
                            unsigned long long w = tag | (2ull << 32);   // virtual tile < 0: prefix 0
                            if (idx >= 0) {
                                w = ld_state(a.tile_state + idx);
                                while ((w >> 34) != (tag >> 34) || ((w >> 32) & 3ull) == 0) {
                                    __nanosleep(200);
                                    w = ld_state(a.tile_state + idx);
                                }
                            }
                            const bool is_prefix = ((w >> 32) & 3ull) == 2;
                            const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
                            const int stop = pm ? __ffs(pm) - 1 : 31;     // nearest tile holding a prefix
                            uint32_t v = lane <= stop ? (uint32_t)w : 0;
#pragma unroll
                            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                            excl += v;
                            if (pm) break;
                            look -= 32;
                        }
                    }
                    if (lane == 0) {
                        st_state(a.tile_state + tile, tag | (2ull << 32) | (excl + total));
                        if (last) *a.d_count = excl + total;
                    }
                }
                if (lane == 0) trace(a.trace, it, 6);
                // ---- stream the tile's points out as one contiguous block at its offset
                //      (8/save_point_cloud.cpp:85-136) ----
                if (total) {
                    const float* cx = xbuf;
                    float* dst = a.pts + 3 * (size_t)excl;
                    const int n = 3 * (int)total;
                    // head so that the body is 16-byte aligned in global memory
                    const int head = min(n, (int)((4 - (((uintptr_t)dst >> 2) & 3)) & 3));
                    if (lane < head) dst[lane] = cx[lane];
                    const int nvec = (n - head) >> 2;
                    for (int v = lane; v < nvec; v += 32) {
                        const int i = head + 4 * v;
                        *reinterpret_cast<float4*>(dst + i) = make_float4(cx[i], cx[i + 1], cx[i + 2], cx[i + 3]);
                    }
                    const int tail0 = head + 4 * nvec;
                    if (lane < n - tail0) dst[tail0 + lane] = cx[tail0 + lane];
                }
                if (extras) {
                    // optional per-point extras (pixel index, texture colour): walk the valid flags
                    // that are still in the slot, then release it
                    if (total) {
                        uint32_t run = excl;
                        for (int k = 0; k < CW; k++) {
                            const int w = k * 32 + lane, lp = 4 * w;
                            const uint32_t f = lp < wt ? vw[w] : 0u;
                            const uint32_t cnt = __popc(f);
                            uint32_t incl = cnt;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                                if (lane >= o) incl += t;
                            }
                            uint32_t dstp = run + incl - cnt;
                            run += __shfl_sync(0xffffffffu, incl, 31);
                            for (int j = 0; j < 4; j++) {
                                if ((f >> (8 * j)) & 1u) {
                                    const size_t gp = (size_t)p0 + lp + j;
                                    if (a.pix) a.pix[dstp] = (uint32_t)((size_t)a.row0 * W + gp);
                                    if (a.rgb) {
                                        a.rgb[3 * (size_t)dstp + 0] = a.texture[3 * gp + 2];
                                        a.rgb[3 * (size_t)dstp + 1] = a.texture[3 * gp + 1];
                                        a.rgb[3 * (size_t)dstp + 2] = a.texture[3 * gp + 0];
                                    }
                                    dstp++;
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) {
                        bulk_wait_read();
                        trace(a.trace, it, 7);
                        mbar_arrive(bar_empty + 8 * s);
                    }
                }
                __syncwarp();   // xbuf is reused by the next tile
                continue;
            }
            __syncwarp();
            if (lane == 0) {
                bulk_wait_read();                    // the TMA stores have read the slot
                trace(a.trace, it, 7);
                mbar_arrive(bar_empty + 8 * s);      // hand it back to the producer
            }
        }
        if (lane == 0) bulk_wait_all();
        return;
    }

    // ================================ CONSUMERS ================================
    for (int it = 0;; it++) {
        const int s = it % stages, ph = (it / stages) & 1;
        mbar_wait_consumer(bar_full + 8 * s, ph);
        if (tid == 0) trace(a.trace, it, 2);
        const int tile = slot_tile[s];
        if (tile < 0) {                          // no more work: pass the end marker to the epilogue
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_staged + 8 * s);
            break;
        }
        const int p0 = a.tile_list[tile] * T, wt = min(T, plane - p0);
        const int lp0 = 4 * tid;                 // first of this thread's 4 pixels inside the tile
        const bool active = lp0 < wt;
        const int row = (p0 + lp0) / W;          // the 4 pixels share a row (W % 4 == 0)
        const int xt = (p0 + lp0) - row * W;
        const int y = a.row0 + row;
        uint8_t* stage = smem + (size_t)s * stage_bytes;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(stage);
        const uint8_t* sroi = stage + (stage_bytes - G.roi_bytes);

        // ---------------- integer phase: mask, fringe terms, Gray bits ----------------
        uint32_t mbits = 0;
        Terms Tv, Th;
        uint32_t gvA = 0, gvB = 0, ghA = 0, ghB = 0;
        if (active) {
            const bool window_ok = y >= 2 && y + 1 < a.H_total && xt >= 4 && xt + 7 < W;
            bool fast = false;
            if (window_ok) {
                uint32_t any_zero = 0;   // haszero(v): non-zero iff some byte of v is zero
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(sroi + r * G.roi_row + (lp0 + ROI_HALO - 4) + 4 * c);
                        any_zero |= (v - 0x01010101u) & ~v & 0x80808080u;
                    }
                if (any_zero == 0) { mbits = 0xf; fast = true; }
            }
            if (!fast) {
                const uint32_t centre = *reinterpret_cast<const uint32_t*>(sroi + 2 * G.roi_row + lp0 + ROI_HALO);
                if (centre != 0) {
#pragma unroll 1
                    for (int j = 0; j < 4; j++) {
                        const int x = xt + j;
                        // ROI flag of global pixel (gx, gy): row offset gy - y in -2..+1 selects the
                        // window row, column offset gx - x in -2..+2 stays inside the halo
                        auto inv = [&](int gx, int gy) {
                            return sroi[(gy - y + 2) * G.roi_row + (lp0 + j + (gx - x) + ROI_HALO)] == 0;
                        };
                        bool v = !inv(x, y);
                        const bool border = x == 0 || y == 0 || x == W - 1 || y == a.H_total - 1;
                        if (v && !border) v = !mask_trigger(x, y, W, a.H_total, inv);
                        mbits |= (v ? 1u : 0u) << j;
                    }
                }
            }
            if (mbits) {
                fringe_terms<N>(sw, 0, WPF, tid, Tv);
                gray_bits(sw, N, N + a.M_v, a.M_v, WPF, tid, gvA, gvB);
                if (DIRS == 2) {
                    const int fh = N + 2 * a.M_v;
                    fringe_terms<N>(sw, fh, WPF, tid, Th);
                    gray_bits(sw, fh + N, fh + N + a.M_h, a.M_h, WPF, tid, ghA, ghB);
                }
            }
        }
        // every consumer is done reading this slot's inputs: they are overwritten by the staged
        // outputs below
        cons_sync<NCONS>();
        if (tid == 0) trace(a.trace, it, 3);

        // ---------------- FP64 phase, one pixel at a time, results staged in the slot ----------------
        float* o_unwv = reinterpret_cast<float*>(stage + G.out_unwv);
        int16_t* o_codev = reinterpret_cast<int16_t*>(stage + G.out_codev);
        uint8_t* o_valid = stage + G.out_valid;
        float* o_unwh = reinterpret_cast<float*>(stage + G.out_unwh);
        int16_t* o_codeh = reinterpret_cast<int16_t*>(stage + G.out_codeh);
        int2* o_cp = reinterpret_cast<int2*>(stage + G.out_cp);
        float* o_x = reinterpret_cast<float*>(stage + G.out_x);
        uint32_t vbits = 0;
        if (active) {
            // (a) decode, straight-line and branch-free so that the 4 pixels (and, inside each,
            //     the two directions) interleave in the FP64 pipe.  Pixels outside the mask are
            //     computed too and discarded by selects (tiles without ROI pixels never get here).
            float r_unwv[4], r_unwh[4];
            int r_cv[4], r_ch[4];
            int2 r_cp[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int x = xt + j;
                const bool m = (mbits >> j) & 1u;
                const int cv = code_of(gvA, gvB, j, a.M_v);
                const float wv = add_pi(phase_of<N>(Tv, j, tab));                    // 4/phase_unwrap.cpp:290
                float unwv = (x == 0 || x == W - 1) ? 0.0f : unwrap_abs(wv, cv, fastdiv);  // :285, :291
                unwv = m ? unwv : 0.0f;
                bool v = m;
                r_cv[j] = m ? cv : -1;
                if (DIRS == 2) {
                    const int ch = code_of(ghA, ghB, j, a.M_h);
                    const float wh = add_pi(phase_of<N>(Th, j, tab));
                    float unwh = (y == 0 || y == a.H_total - 1) ? 0.0f : unwrap_abs(wh, ch, fastdiv);  // :304, :309
                    unwh = m ? unwh : 0.0f;
                    long long px = 0, py = 0;                                         // 5/compute_correspondance.cpp:648-675
                    const bool okx = correspond(unwv, a.fw_v, &px, fastdiv);
                    const bool oky = correspond(unwh, a.fw_h, &py, fastdiv);
                    // FE_INVALID on x rejects before y is computed (:650-655); on y after x is stored
                    r_cp[j].x = (m && okx) ? sat32(px) : 0;
                    r_cp[j].y = (m && okx && oky) ? sat32(py) : 0;
                    v = m && okx && oky && !(px > a.PW - 1 || py > a.PH - 1 || px < 0 || py < 0);
                    r_unwh[j] = unwh;
                    r_ch[j] = m ? ch : -1;
                }
                r_unwv[j] = unwv;
                vbits |= (v ? 1u : 0u) << j;
            }
            // the thread's 4 pixels are consecutive: one vector store per plane
            *reinterpret_cast<float4*>(o_unwv + lp0) = make_float4(r_unwv[0], r_unwv[1], r_unwv[2], r_unwv[3]);
            *reinterpret_cast<uint2*>(o_codev + lp0) =
                make_uint2((uint32_t)(r_cv[0] & 0xffff) | ((uint32_t)r_cv[1] << 16), (uint32_t)(r_cv[2] & 0xffff) | ((uint32_t)r_cv[3] << 16));
            *reinterpret_cast<uint32_t*>(o_valid + lp0) =
                (vbits & 1u) | ((vbits & 2u) << 7) | ((vbits & 4u) << 14) | ((vbits & 8u) << 21);
            if (DIRS == 2) {
                *reinterpret_cast<float4*>(o_unwh + lp0) = make_float4(r_unwh[0], r_unwh[1], r_unwh[2], r_unwh[3]);
                *reinterpret_cast<uint2*>(o_codeh + lp0) =
                    make_uint2((uint32_t)(r_ch[0] & 0xffff) | ((uint32_t)r_ch[1] << 16), (uint32_t)(r_ch[2] & 0xffff) | ((uint32_t)r_ch[3] << 16));
                *reinterpret_cast<int4*>(o_cp + lp0) = make_int4(r_cp[0].x, r_cp[0].y, r_cp[1].x, r_cp[1].y);
                *reinterpret_cast<int4*>(o_cp + lp0 + 2) = make_int4(r_cp[2].x, r_cp[2].y, r_cp[3].x, r_cp[3].y);
            }
            // (b) triangulation of the surviving pixels (7/triangulation.cpp:1230-1247)
            if (DIRS == 2) {
#pragma unroll 1
                for (int j = 0; j < 4; j++) {
                    if (!((vbits >> j) & 1u)) continue;
                    const int x = xt + j, lp = lp0 + j;
                    const int2 cp = o_cp[lp];
                    double uc, vc, up, vp, X[3];
                    if (a.cam_lut) {
                        const double2 t = a.cam_lut[(size_t)p0 + lp];
                        uc = t.x; vc = t.y;
                    } else {
                        undistorted_pixel_nodist(cal.Kc, cal.ifx_c, cal.ify_c, cal.cam_std != 0, (double)x, (double)y, &uc, &vc);
                    }
                    if (a.proj_lut) {
                        const double2 t = a.proj_lut[(size_t)cp.y * a.PW + cp.x];
                        up = t.x; vp = t.y;
                    } else {
                        undistorted_pixel_nodist(cal.Kp, cal.ifx_p, cal.ify_p, cal.proj_std != 0, (double)cp.x, (double)cp.y, &up, &vp);
                    }
                    if (EXACT) triangulate_point(cal.Ac, cal.Ap, uc, vc, up, vp, X);
                    else triangulate_point_fast(cal.Ac, cal.Ap, uc, vc, up, vp, X);
                    o_x[3 * lp + 0] = __double2float_rn(X[0]);                       // 8/save_point_cloud.cpp:94-96
                    o_x[3 * lp + 1] = __double2float_rn(X[1]);
                    o_x[3 * lp + 2] = __double2float_rn(X[2]);
                }
            }
        }
        if (DIRS == 2) {
            // ---- tile-local compaction: block scan of the valid counts, then every thread moves
            //      its points to their raster rank inside the slot ----
            const uint32_t cnt = __popc(vbits);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) slot_cnt[s * 8 + warp] = incl;
            cons_sync<NCONS>();
            uint32_t rank = incl - cnt;
#pragma unroll
            for (int w2 = 0; w2 < CW; w2++) rank += w2 < warp ? slot_cnt[s * 8 + w2] : 0u;
            float* cx = reinterpret_cast<float*>(stage + G.out_cx);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if ((vbits >> j) & 1u) {
                    const int lp = lp0 + j;
                    cx[3 * rank + 0] = o_x[3 * lp + 0];
                    cx[3 * rank + 1] = o_x[3 * lp + 1];
                    cx[3 * rank + 2] = o_x[3 * lp + 2];
                    rank++;
                }
            }
        }
        fence_async_smem();   // staged outputs -> visible to the async (TMA) proxy
        __syncwarp();
        if (tid == 0) trace(a.trace, it, 4);
        if (lane == 0) mbar_arrive(bar_staged + 8 * s);
    }
}

template <int N, int DIRS, int CW, int MINB, bool EXACT>
static cudaError_t launch_fused_t(const FusedArgs& a, const DeviceCalib& cal, int sm_count, const FusedPlan& p,
                                  cudaStream_t st)
{
    auto kern = k_fused<N, DIRS, CW, MINB, EXACT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (CW + 2) * 32, p.smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    if (per_sm > MINB) per_sm = MINB;
    // every CTA must be resident (the look-back chain waits on earlier tiles)
    const int grid = a.n_tiles < sm_count * per_sm ? a.n_tiles : sm_count * per_sm;
    e = launch_worklist(a, 128 * CW, DIRS, st);
    if (e != cudaSuccess) return e;
    kern<<<grid, (CW + 2) * 32, p.smem, st>>>(a, cal, p.stages, p.stage_bytes);
    return cudaGetLastError();
}

template <int N, int DIRS>
static cudaError_t launch_fused_nd(const FusedArgs& a, const DeviceCalib& cal, int sm_count, const FusedPlan& p,
                                   bool exact, cudaStream_t st)
{
    // DIRS == 1 has no triangulation: only the EXACT = true instance exists
#define S3D_CASE(CWV, MB)                                                                              \
    if (p.cw == CWV && p.minb == MB) {                                                                 \
        if (exact || DIRS == 1) return launch_fused_t<N, DIRS, CWV, MB, true>(a, cal, sm_count, p, st); \
        return launch_fused_t<N, DIRS, CWV, MB, (DIRS == 1)>(a, cal, sm_count, p, st);                  \
    }
    S3D_CASE(6, 2) S3D_CASE(4, 2) S3D_CASE(8, 1)
#undef S3D_CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_fused(const scan3d_config& c, const FusedArgs& a_in, const DeviceCalib& cal, int sm_count,
                         cudaStream_t st)
{
    FusedPlan p;
    if (!plan_for(c, &p)) return cudaErrorInvalidValue;
    if (((uintptr_t)a_in.stack & 15) || ((uintptr_t)a_in.roi & 15)) return cudaErrorMisalignedAddress;
    FusedArgs a = a_in;
    const int T = 128 * p.cw;
    a.tiles_per_row = 0;
    a.n_tiles = (int)(((size_t)c.W * c.H + T - 1) / T);
    const bool exact = !(c.flags & SCAN3D_FLAG_FAST_TRIANGULATION);
#define S3D_F(NN) (c.dirs == 2 ? launch_fused_nd<NN, 2>(a, cal, sm_count, p, exact, st) : launch_fused_nd<NN, 1>(a, cal, sm_count, p, exact, st))
    switch (c.N) {
        case 3: return S3D_F(3);
        case 4: return S3D_F(4);
        case 5: return S3D_F(5);
        case 8: return S3D_F(8);
    }
#undef S3D_F
    return cudaErrorInvalidValue;
}

}  // namespace s3d
