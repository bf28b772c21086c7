// scan3d_fused_kernel8.cu -- third cut of the single-pass kernel ("v8").  Same work and the same bit-for-bit
// results as k_fused7 (scan3d_fused_kernel7.cu); what changed is who waits for whom (round-2 ncu per-line
// profile of v7, profiles/r2_optimisation_log.md: 17 % of the consumer warps' time at the CTA-wide count
// barrier, 12 % on global loads, 2 extra launches per scan for the work list):
//
//   * ONE launch per scan.  There is no work-list pre-pass: work positions are tile indices drawn from a
//     global counter that is never reset (the host passes the counter's value at launch: every CTA overdraws
//     exactly once, so a launch advances it by n_tiles + grid); the IO warp looks at the tile's own ROI bytes
//     and, when no pixel is selected, writes the tile's constant outputs itself, publishes a zero count and
//     draws again -- the 56 frame segments of such a tile are never read;
//   * of a tile that does hold ROI pixels only the 128-pixel sub-tiles (one per consumer warp) with ROI
//     pixels are loaded: one 2-D tensor-map copy [frames] x [128 B] each;
//   * no barrier among the consumer warps.  Every warp owns a private point buffer (2 x 128 points): it
//     counts its own survivors, reports the count through an mbarrier, triangulates straight into its buffer
//     and -- two tiles later, when the IO warp has resolved the tile's place in the raster order
//     (decoupled look-back) and published one base offset per warp -- streams its own points out.  The
//     IO warp never touches a point;
//   * a warp whose 128 pixels hold no ROI pixel skips the FP64 phase altogether;
//   * the undistortion-table entry of the next surviving pixel is fetched while the current one is solved.
//
//   CTA = CW consumer warps + 1 IO warp; shared memory and register budget as in v7 (3 CTAs = 21 consumer
//   warps per SM at 80 registers for the 56-frame 12 MP configuration).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "scan3d_fused_common.cuh"

namespace s3d {

constexpr int ROI8_ROW = 128 + 2 * ROI_HALO;     // one warp's ROI window row: its 128 pixels + halo
constexpr int FIXED8 = 24 * 8 /*mbarriers: full[8] free[8] counted[2] prefix[2] (+4 spare)*/ + 64 /*cnts[2][8]*/ + 64 /*base[2][8]*/ +
                       32 /*pos ring[8]*/ + 32 /*sub-tile mask ring[8]*/;

static int num_frames8(const scan3d_config& c)
{
    return c.dirs == 2 ? 2 * c.N + 2 * (c.M_v + c.M_h) : c.N + 2 * c.M_v;
}
static size_t smem8(const scan3d_config& c, int cw)
{
    const int T = 128 * cw;
    return (size_t)num_frames8(c) * T + (size_t)cw * 4 * ROI8_ROW + (c.dirs == 2 ? 2 * 12 * T + 2 * cw * 16 : 0) + FIXED8;
}

struct Plan8 {
    int cw, minb;
    size_t smem;
};

static bool plan8(const scan3d_config& c, Plan8* out)
{
    if (c.W % 16 != 0) return false;
    if (!(c.N == 3 || c.N == 4 || c.N == 5 || c.N == 8)) return false;
    int minb0 = 3;
    if (const char* e = getenv("SCAN3D_FUSED_CFG")) {
        int a = 0, b = 0;
        if (sscanf(e, "%d,%d", &a, &b) == 2 && a == 7 && (b == 2 || b == 3)) minb0 = b;
    }
    for (int b = minb0; b >= 2; b--) {
        const size_t sm = smem8(c, 7);
        if ((sm + 1024) * b <= (size_t)SMEM_MAX + 1024) {
            out->cw = 7; out->minb = b; out->smem = sm;
            return true;
        }
    }
    return false;
}

bool fused8_supported(const scan3d_config& c)
{
    Plan8 p;
    return plan8(c, &p);
}

constexpr int regs8(int cw, int minb)
{
    // per SM sub-partition: ceil(resident warps / 4) warps share 16384 registers
    const int warps = minb * (cw + 1);
    const int r = 16384 / (((warps + 3) / 4) * 32);
    return r > 255 ? 255 : (r / 8) * 8;
}

template <int N, int DIRS, int CW, int MINB, bool EXACT>
__global__ void __launch_bounds__((CW + 1) * 32) __maxnreg__(regs8(CW, MINB))
k_fused8(const __grid_constant__ FusedArgs a, const __grid_constant__ DeviceCalib cal, const __grid_constant__ CUtensorMap stack_map)
{
    constexpr int T = 128 * CW;
    constexpr bool fastdiv = true;   // host verified (else this kernel is not used)
    static_assert(CW <= 8, "per-warp state is kept in 8-entry rows");

    extern __shared__ __align__(128) uint8_t smem[];
    const int NF = DIRS == 2 ? 2 * N + 2 * (a.M_v + a.M_h) : N + 2 * a.M_v;
    uint8_t* slot = smem;                                                          // [CW][NF][128]: one sub-slot per warp
    uint8_t* sroi = smem + (size_t)NF * T;                                         // [CW][4][ROI8_ROW]
    float* cxb = reinterpret_cast<float*>(sroi + CW * 4 * ROI8_ROW);                // [2][CW][3*128] points
    uint32_t* vbal = reinterpret_cast<uint32_t*>(cxb + (DIRS == 2 ? 2 * 3 * T : 0));   // [2][CW][4] ballots of the valid bits
    uint64_t* bars = reinterpret_cast<uint64_t*>(vbal + (DIRS == 2 ? 2 * CW * 4 : 0));
    volatile uint32_t* cnts = reinterpret_cast<volatile uint32_t*>(bars + 24);     // [2][8] survivors per warp
    volatile uint32_t* base = cnts + 16;                                           // [2][8] global point offset per warp
    volatile int* posr = reinterpret_cast<volatile int*>(const_cast<uint32_t*>(base) + 16);   // [8] work positions of the CTA's tiles (-1: end)
    volatile int* subr = posr + 8;                                                 // [8] their sub-tile masks
    const uint32_t bar_full = smem_u32(bars), bar_free = smem_u32(bars + 8), bar_counted = smem_u32(bars + 16),
                   bar_prefix = smem_u32(bars + 18);
    const double* __restrict__ tab = a.atan_tab;    // 72 doubles, read through L1 (the shared memory is full)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = a.W;
    const int plane = W * a.H;
    const int n_tiles = a.n_tiles;

    if (tid == 0) {
        for (int w = 0; w < CW; w++) {
            mbar_init(bar_full + 8 * w, 1);
            mbar_init(bar_free + 8 * w, 1);
        }
        mbar_init(bar_counted, CW);
        mbar_init(bar_counted + 8, CW);
        mbar_init(bar_prefix, 1);
        mbar_init(bar_prefix + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    __syncthreads();

    if (warp == CW) {
        // ================================ IO WARP ================================
        const unsigned long long tag = (unsigned long long)(a.epoch & 0x3fffffffu) << 34;
        const long long roi_total = (long long)W * a.H_total;
        int kw = 0;                              // lane w: sub-tiles handed to consumer warp w so far
        bool done_w = lane >= CW;                // lane w: warp w got its end marker
        int drawn = 0;                           // tiles (and the end marker) drawn by this CTA
        int n_real = 0;                          // ... of which real tiles
        bool end_drawn = false;
        int agg_it = 0;                          // next tile whose count gets published
        int epi_it = 0;                          // next tile whose place in the raster order gets resolved
        bool resolving = false;
        int look = 0;
        uint32_t excl = 0;
        uint32_t tot[2] = {0, 0}, woff[2] = {0, 0};
        int epos[2] = {0, 0};
        for (;;) {
            const bool all_done = __all_sync(0xffffffffu, done_w);
            if (all_done && !(DIRS == 2 && epi_it < n_real)) break;
            bool progressed = false;
            // ---- (1) consumer warps whose sub-slot is free: hand each the next tile's sub-tile ----
            unsigned ready = __ballot_sync(0xffffffffu, !done_w && mbar_test(bar_free + 8 * lane, (kw & 1) ^ 1));
            while (ready) {
                progressed = true;
                const int w = __ffs(ready) - 1;
                ready &= ready - 1;
                const int k = __shfl_sync(0xffffffffu, kw, w);
                if (k == drawn) {
                    // the first warp to get here draws the CTA's next tile: work positions until one with ROI pixels turns up
                    int pos = -1;
                    uint32_t sub = 0;
                    while (!end_drawn) {
                        pos = 0;
                        if (lane == 0) pos = (int)(atomicAdd(a.sched_ctr, 1u) - a.pos_base);
                        pos = __shfl_sync(0xffffffffu, pos, 0);
                        if (pos >= n_tiles) { pos = -1; end_drawn = true; break; }
                        const int p0 = pos * T, wt = min(T, plane - p0);
                        const uint8_t* r = a.roi + (size_t)a.row0 * W + p0;
                        // whoever draws position p pulls the ROI bytes of position p + grid into L2: that is about where
                        // the draws will be one tile period from now
                        if (lane < CW && pos + (int)gridDim.x < n_tiles && 128 * lane < plane - p0 - (int)gridDim.x * T)
                            prefetch_l2(r + (size_t)gridDim.x * T + 128 * lane);
                        sub = 0;
                        for (int o = 0; o < T; o += 512) {
                            const int off = o + 16 * lane;
                            uint4 v = make_uint4(0, 0, 0, 0);
                            if (off < wt) v = __ldg(reinterpret_cast<const uint4*>(r + off));
                            const unsigned m = __ballot_sync(0xffffffffu, (v.x | v.y | v.z | v.w) != 0);
                            // 8 lanes = 128 bytes = one sub-tile
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                if ((m >> (8 * q)) & 0xffu) sub |= 1u << (o / 128 + q);
                        }
                        // the first and the last tile always take the regular route (prefix seed / final count)
                        if (sub != 0 || pos == 0 || pos == n_tiles - 1) break;
                        // no selected pixel: constant outputs, zero count, next draw
                        const uint4 z = make_uint4(0, 0, 0, 0), m1 = make_uint4(~0u, ~0u, ~0u, ~0u);
                        for (int i = lane; i < wt / 4; i += 32) {
                            reinterpret_cast<uint4*>(a.unw_v + p0)[i] = z;
                            if (DIRS == 2) reinterpret_cast<uint4*>(a.unw_h + p0)[i] = z;
                        }
                        for (int i = lane; i < wt / 8; i += 32) {
                            reinterpret_cast<uint4*>(a.code_v + p0)[i] = m1;
                            if (DIRS == 2) reinterpret_cast<uint4*>(a.code_h + p0)[i] = m1;
                        }
                        for (int i = lane; i < wt / 16; i += 32) reinterpret_cast<uint4*>(a.valid + p0)[i] = z;
                        if (DIRS == 2) {
                            for (int i = lane; i < wt / 2; i += 32) reinterpret_cast<uint4*>(a.cpmap + p0)[i] = z;
                            if (lane == 0) {
                                // count 0 goes out at once; it is a prefix already when the predecessor's is known
                                const unsigned long long ws = ld_state(a.tile_state + pos - 1);
                                const bool pre = (ws >> 34) == (tag >> 34) && ((ws >> 32) & 3ull) == 2;
                                st_state(a.tile_state + pos, pre ? ws : (tag | (1ull << 32)));
                            }
                        }
                    }
                    if (lane == 0) {
                        posr[drawn & 7] = pos;
                        subr[drawn & 7] = (int)sub;
                    }
                    __syncwarp();
                    drawn++;
                    if (pos >= 0) n_real++;
                }
                const int pos = posr[k & 7];
                const bool has = pos >= 0 && ((subr[k & 7] >> w) & 1);
                const uint32_t bfull = bar_full + 8 * w;
                if (!has) {
                    // end marker, or a sub-tile without ROI pixels: nothing to load
                    if (lane == 0) mbar_arrive(bfull);
                } else {
                    const int p0w = pos * T + 128 * w, wtw = min(128, plane - p0w);
                    const long long gbase = (long long)a.row0 * W + p0w - ROI_HALO;
                    long long seg0 = 0, seg1 = 0;
                    uint32_t roi_tx = 0;
                    if (lane < 4) {
                        seg0 = max(gbase + (long long)(lane - 2) * W, 0LL);
                        seg1 = min(gbase + (long long)(lane - 2) * W + wtw + 2 * ROI_HALO, roi_total);
                        if (seg1 > seg0) roi_tx = (uint32_t)(seg1 - seg0);
                    }
                    uint32_t roi_sum = roi_tx;
                    roi_sum += __shfl_xor_sync(0xffffffffu, roi_sum, 1);
                    roi_sum += __shfl_xor_sync(0xffffffffu, roi_sum, 2);
                    roi_sum = __shfl_sync(0xffffffffu, roi_sum, 0);
                    if (lane == 0) mbar_expect_tx(bfull, (uint32_t)NF * (a.use_tmap ? 128u : (uint32_t)wtw) + roi_sum);
                    __syncwarp();
                    const uint32_t dst = smem_u32(slot) + w * NF * 128;
                    if (a.use_tmap) {
                        // [NF frames] x [128 B] as one 2-D tensor copy (64-bit elements; bytes past the end of a frame
                        // are zero-filled): lands as a dense [NF][128] block
                        if (lane == 0) tensor_g2s_2d(dst, &stack_map, p0w >> 3, 0, bfull);
                    } else {
                        const uint8_t* src = a.stack + p0w;
                        for (int f = lane; f < NF; f += 32) bulk_g2s(dst + f * 128, src + (size_t)f * plane, (uint32_t)wtw, bfull);
                    }
                    if (roi_tx)
                        bulk_g2s(smem_u32(sroi) + (w * 4 + lane) * ROI8_ROW + (uint32_t)(seg0 - (gbase + (long long)(lane - 2) * W)),
                                 a.roi + seg0, roi_tx, bfull);
                }
                if (lane == w) {
                    kw++;
                    done_w = pos < 0;
                }
            }
            if (DIRS == 2) {
                // ---- (2) every warp of a tile has reported its survivors: the tile's count goes out at once
                //      (its triangulation is still running; every later tile's look-back needs it) ----
                if (agg_it < n_real && agg_it < epi_it + 2 &&
                    __any_sync(0xffffffffu, mbar_try(bar_counted + 8 * (agg_it & 1), (agg_it >> 1) & 1))) {
                    progressed = true;
                    const int b = agg_it & 1;
                    const int pos = posr[agg_it & 7];
                    const uint32_t c = lane < CW ? cnts[b * 8 + lane] : 0u;
                    uint32_t incl = c;
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 7);
                    if (lane == 0) st_state(a.tile_state + pos, tag | ((pos ? 1ull : 2ull) << 32) | total);
                    tot[b] = total;
                    woff[b] = incl - c;
                    epos[b] = pos;
                    agg_it++;
                }
                // ---- (3) resumable decoupled look-back of the oldest counted tile; once its exclusive prefix is
                //      known every warp of the tile gets its base offset and streams its points out itself ----
                if (epi_it < agg_it) {
                    const int b = epi_it & 1;
                    if (!resolving) { resolving = true; excl = 0; look = epos[b] - 1; }
                    bool resolved = epos[b] == 0;
                    if (!resolved) {
#pragma unroll 1
                        for (int hop = 0; hop < 24; hop++) {
                            const int idx = look - lane;
                            unsigned long long ws = tag | (2ull << 32);   // virtual tile < 0: prefix 0
                            if (idx >= 0) ws = ld_state(a.tile_state + idx);
                            const bool okw = (ws >> 34) == (tag >> 34) && ((ws >> 32) & 3ull) != 0;
                            const bool is_prefix = okw && ((ws >> 32) & 3ull) == 2;
                            const unsigned rm = __ballot_sync(0xffffffffu, okw);
                            const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
                            // needed lanes: from the nearest tile up to the first known prefix
                            const int stop = pm ? __ffs(pm) - 1 : 31;
                            const unsigned need = stop == 31 ? 0xffffffffu : ((2u << stop) - 1u);
                            if ((rm & need) != need) break;            // a needed count is not out yet
                            progressed = true;
                            uint32_t v = lane <= stop ? (uint32_t)ws : 0;
#pragma unroll
                            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                            excl += v;
                            if (pm) { resolved = true; break; }
                            look -= 32;
                        }
                    }
                    if (resolved) {
                        progressed = true;
                        if (lane == 0) {
                            if (epos[b] != 0) st_state(a.tile_state + epos[b], tag | (2ull << 32) | (excl + tot[b]));
                            if (epos[b] == n_tiles - 1) *a.d_count = excl + tot[b];
                        }
                        if (lane < CW) base[b * 8 + lane] = excl + woff[b];
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_prefix + 8 * b);
                        epi_it++;
                        resolving = false;
                    }
                }
            }
            if (!progressed) __nanosleep(200);
        }
        return;
    }

    // ================================ CONSUMERS ================================
    // Every warp runs its own pipeline over the CTA's tile sequence: sub-slot, ROI window, point buffers and the
    // full / free barriers are the warp's own; the warps of a CTA only meet in the tile's count (bar_counted).
    const uint32_t my_full = bar_full + 8 * warp, my_free = bar_free + 8 * warp;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(slot + (size_t)warp * NF * 128);
    const uint8_t* sroi_w = sroi + warp * 4 * ROI8_ROW;

    // streams the warp's n_pts points of tile number k of this CTA (they sit in buffer k & 1) to their place
    // in the raster order: gb = global index of the warp's first point
    auto drain = [&](int k, uint32_t gb, uint32_t n_pts) {
        const int b = k & 1;
        const float* src = cxb + (b * CW + warp) * 384;
        float* dst = a.pts + 3 * (size_t)gb;
        const int n = 3 * (int)n_pts;
        for (int i = lane; i < n; i += 32) dst[i] = src[i];
        if (a.pix || a.rgb) {
            const uint32_t* vb4 = vbal + (b * CW + warp) * 4;
            const uint32_t vb = ((vb4[0] >> lane) & 1u) | (((vb4[1] >> lane) & 1u) << 1) | (((vb4[2] >> lane) & 1u) << 2) |
                                (((vb4[3] >> lane) & 1u) << 3);
            const int pix0 = posr[k & 7] * T + 4 * tid;
            const uint32_t c = __popc(vb);
            uint32_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            uint32_t dp = gb + incl - c;
            for (int j = 0; j < 4; j++)
                if ((vb >> j) & 1u) {
                    const size_t gp = (size_t)pix0 + j;
                    if (a.pix) a.pix[dp] = (uint32_t)((size_t)a.row0 * W + gp);
                    if (a.rgb) {
                        a.rgb[3 * (size_t)dp + 0] = a.texture[3 * gp + 2];
                        a.rgb[3 * (size_t)dp + 1] = a.texture[3 * gp + 1];
                        a.rgb[3 * (size_t)dp + 2] = a.texture[3 * gp + 0];
                    }
                    dp++;
                }
        }
    };

    int it = 0;
    for (;; it++) {
        if (!mbar_try(my_full, it & 1))
            while (!mbar_try(my_full, it & 1)) __nanosleep(64);
        const int pos = posr[it & 7];
        if (pos < 0) break;
        const bool loaded = (subr[it & 7] >> warp) & 1;       // this warp's 128 pixels hold ROI pixels and were loaded
        const int p0 = pos * T, wt = min(T, plane - p0);
        const int lp0 = 4 * tid, lpw = 4 * lane;
        const bool active = lp0 < wt;
        const int row = (p0 + lp0) / W;
        const int xt = (p0 + lp0) - row * W;
        const int y = a.row0 + row;

        // ---------------- integer phase: mask, fringe terms, Gray bits ----------------
        uint32_t mbits = 0;
        Terms Tv, Th;
        uint32_t gvA = 0, gvB = 0, ghA = 0, ghB = 0;
        if (active && loaded) {
            const bool window_ok = y >= 2 && y + 1 < a.H_total && xt >= 4 && xt + 7 < W;
            bool fast = false;
            if (window_ok) {
                uint32_t any_zero = 0;
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(sroi_w + r * ROI8_ROW + (lpw + ROI_HALO - 4) + 4 * c);
                        any_zero |= (v - 0x01010101u) & ~v & 0x80808080u;
                    }
                if (any_zero == 0) { mbits = 0xf; fast = true; }
            }
            if (!fast) {
                const uint32_t centre = *reinterpret_cast<const uint32_t*>(sroi_w + 2 * ROI8_ROW + lpw + ROI_HALO);
                if (centre != 0) {
#pragma unroll 1
                    for (int j = 0; j < 4; j++) {
                        const int x = xt + j;
                        auto inv = [&](int gx, int gy) {
                            return sroi_w[(gy - y + 2) * ROI8_ROW + (lpw + j + (gx - x) + ROI_HALO)] == 0;
                        };
                        bool v = !inv(x, y);
                        const bool border = x == 0 || y == 0 || x == W - 1 || y == a.H_total - 1;
                        if (v && !border) v = !mask_trigger(x, y, W, a.H_total, inv);
                        mbits |= (v ? 1u : 0u) << j;
                    }
                }
            }
            if (mbits) {
                fringe_terms<N>(sw, 0, 32, lane, Tv);
                gray_bits(sw, N, N + a.M_v, a.M_v, 32, lane, gvA, gvB);
                if (DIRS == 2) {
                    const int fh = N + 2 * a.M_v;
                    fringe_terms<N>(sw, fh, 32, lane, Th);
                    gray_bits(sw, fh + N, fh + N + a.M_h, a.M_h, 32, lane, ghA, ghB);
                }
            }
        }
        // the warp is done with its sub-slot: the next tile's sub-tile can come in
        __syncwarp();
        if (lane == 0) mbar_arrive(my_free);
        // ---------------- FP64 phase (registers only) ----------------
        uint32_t vbits = 0;
        int4 cp01 = make_int4(0, 0, 0, 0), cp23 = cp01;   // the 4 pixels' correspondences, for the triangulation below
        const size_t g = (size_t)p0 + lp0;
        if (__any_sync(0xffffffffu, mbits != 0)) {
            if (active) {
                // Two passes of 2 pixels: inside a pass everything is straight-line (2 pixels x 2
                // directions interleave in the FP64 pipe); the pass loop is rolled to keep the
                // consumer loop inside the instruction cache.
#pragma unroll 1
                for (int h = 0; h < 2; h++) {
                    float r_unwv[2], r_unwh[2];
                    int r_cv[2], r_ch[2];
                    int2 r_cp[2];
                    uint32_t vb = 0;
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int j = 2 * h + u;
                        const int x = xt + j;
                        const bool m = (mbits >> j) & 1u;
                        const int cv = code_of(gvA, gvB, j, a.M_v);
                        const float wv = add_pi(phase_of<N>(Tv, j, tab));                    // 4/phase_unwrap.cpp:290
                        float unwv = (x == 0 || x == W - 1) ? 0.0f : unwrap_abs(wv, cv, fastdiv);  // :285, :291
                        unwv = m ? unwv : 0.0f;
                        bool v = m;
                        r_cv[u] = m ? cv : -1;
                        if (DIRS == 2) {
                            const int ch = code_of(ghA, ghB, j, a.M_h);
                            const float wh = add_pi(phase_of<N>(Th, j, tab));
                            float unwh = (y == 0 || y == a.H_total - 1) ? 0.0f : unwrap_abs(wh, ch, fastdiv);  // :304, :309
                            unwh = m ? unwh : 0.0f;
                            int px, py;                                                       // 5/compute_correspondance.cpp:648-675
                            const bool okx = correspond32(unwv, a.fw_v, &px);
                            const bool oky = correspond32(unwh, a.fw_h, &py);
                            // FE_INVALID on x rejects before y is computed (:650-655); on y after x is stored
                            r_cp[u].x = (m && okx) ? px : 0;
                            r_cp[u].y = (m && okx && oky) ? py : 0;
                            // 0 <= p <= P-1 as one unsigned compare (saturated values fall outside as well)
                            v = m && okx && oky && (unsigned)px <= (unsigned)(a.PW - 1) && (unsigned)py <= (unsigned)(a.PH - 1);
                            r_unwh[u] = unwh;
                            r_ch[u] = m ? ch : -1;
                        }
                        r_unwv[u] = unwv;
                        vb |= (v ? 1u : 0u) << u;
                    }
                    // plane outputs of the 2 pixels: one (vector) store per plane
                    const size_t gh = g + 2 * h;
                    *reinterpret_cast<float2*>(a.unw_v + gh) = make_float2(r_unwv[0], r_unwv[1]);
                    *reinterpret_cast<uint32_t*>(a.code_v + gh) = (uint32_t)(r_cv[0] & 0xffff) | ((uint32_t)r_cv[1] << 16);
                    *reinterpret_cast<uint16_t*>(a.valid + gh) = (uint16_t)((vb & 1u) | ((vb & 2u) << 7));
                    if (DIRS == 2) {
                        *reinterpret_cast<float2*>(a.unw_h + gh) = make_float2(r_unwh[0], r_unwh[1]);
                        *reinterpret_cast<uint32_t*>(a.code_h + gh) = (uint32_t)(r_ch[0] & 0xffff) | ((uint32_t)r_ch[1] << 16);
                        const int4 c4 = make_int4(r_cp[0].x, r_cp[0].y, r_cp[1].x, r_cp[1].y);
                        *reinterpret_cast<int4*>(a.cpmap + gh) = c4;
                        if (h == 0) cp01 = c4; else cp23 = c4;
                        // the undistortion tables are gathered once per surviving pixel a few thousand
                        // cycles from now: pull the lines into L2 meanwhile (DRAM latency -> L2 latency)
                        if (vb) {
                            if (a.cam_lut) prefetch_l2(a.cam_lut + gh);
                            if (a.proj_lut) {
                                if (vb & 1u) prefetch_l2(a.proj_lut + (size_t)c4.y * a.PW + c4.x);
                                if (vb & 2u) prefetch_l2(a.proj_lut + (size_t)c4.w * a.PW + c4.z);
                            }
                        }
                    }
                    vbits |= vb << (2 * h);
                }
            }
        } else if (active) {
            // no ROI pixel among the warp's 128: constant outputs (phase 0, fringe order -1, invalid, c_p_map 0)
            *reinterpret_cast<float4*>(a.unw_v + g) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<uint2*>(a.code_v + g) = make_uint2(~0u, ~0u);
            *reinterpret_cast<uint32_t*>(a.valid + g) = 0u;
            if (DIRS == 2) {
                *reinterpret_cast<float4*>(a.unw_h + g) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<uint2*>(a.code_h + g) = make_uint2(~0u, ~0u);
                reinterpret_cast<int4*>(a.cpmap + g)[0] = make_int4(0, 0, 0, 0);
                reinterpret_cast<int4*>(a.cpmap + g)[1] = make_int4(0, 0, 0, 0);
            }
        }
        if (DIRS == 2) {
            // ---- the warp's survivors: count, report, then triangulate straight into the warp's buffer ----
            const int b = it & 1;
            const uint32_t cnt = __popc(vbits);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t wtotal = __shfl_sync(0xffffffffu, incl, 31);
            // buffer b still holds the warp's points of tile it-2: their base offset must be known by now
            uint32_t gb_prev = 0, n_prev = 0;
            if (it >= 2) {
                const uint32_t par = ((it - 2) >> 1) & 1;
                if (!mbar_try(bar_prefix + 8 * b, par))
                    while (!mbar_try(bar_prefix + 8 * b, par)) __nanosleep(64);
                gb_prev = base[b * 8 + warp];
                n_prev = cnts[b * 8 + warp];
            }
            __syncwarp();
            // (base[b] and cnts[b] are rewritten for this tile only after every warp has arrived here)
            if (lane == 0) {
                cnts[b * 8 + warp] = wtotal;
                mbar_arrive(bar_counted + 8 * b);
            }

            if (it >= 2) {
                drain(it - 2, gb_prev, n_prev);
                __syncwarp();
            }
            if (a.pix || a.rgb) {
                // the threads' valid bits as 4 ballots, for the pixel indices / colours that go with the points
                uint32_t* vb4 = vbal + (b * CW + warp) * 4;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t m = __ballot_sync(0xffffffffu, (vbits >> j) & 1u);
                    if (lane == 0) vb4[j] = m;
                }
                __syncwarp();
            }
            float* cx = cxb + (b * CW + warp) * 384;
            uint32_t rank = incl - cnt;
            // triangulation of the surviving pixels (7/triangulation.cpp:1230-1247); the camera table entry of the
            // next pixel is in flight while this one is solved
            double2 lut_next = make_double2(0.0, 0.0);
            if (a.cam_lut && vbits) lut_next = a.cam_lut[g];
#pragma unroll 1
            for (int j = 0; j < 4; j++) {
                const double2 lut_cur = lut_next;
                if (a.cam_lut && vbits && j < 3) lut_next = a.cam_lut[g + j + 1];
                if (!((vbits >> j) & 1u)) continue;
                const int x = xt + j;
                const int cpx = j == 0 ? cp01.x : j == 1 ? cp01.z : j == 2 ? cp23.x : cp23.z;
                const int cpy = j == 0 ? cp01.y : j == 1 ? cp01.w : j == 2 ? cp23.y : cp23.w;
                double uc, vc, up, vp, Xd[3];
                if (a.cam_lut) {
                    uc = lut_cur.x; vc = lut_cur.y;
                } else {
                    undistorted_pixel_nodist(cal.Kc, cal.ifx_c, cal.ify_c, cal.cam_std != 0, (double)x, (double)y, &uc, &vc);
                }
                if (a.proj_lut) {
                    const double2 t = a.proj_lut[(size_t)cpy * a.PW + cpx];
                    up = t.x; vp = t.y;
                } else {
                    undistorted_pixel_nodist(cal.Kp, cal.ifx_p, cal.ify_p, cal.proj_std != 0, (double)cpx, (double)cpy, &up, &vp);
                }
                if (EXACT) triangulate_point(cal.Ac, cal.Ap, uc, vc, up, vp, Xd);
                else triangulate_point_fast(cal.Ac, cal.Ap, uc, vc, up, vp, Xd);
                cx[3 * rank + 0] = __double2float_rn(Xd[0]);                       // 8/save_point_cloud.cpp:94-96
                cx[3 * rank + 1] = __double2float_rn(Xd[1]);
                cx[3 * rank + 2] = __double2float_rn(Xd[2]);
                rank++;
            }
            __syncwarp();
        }
    }
    if (DIRS == 2) {
        // the last two tiles of this CTA are still in the buffers
        for (int k = max(it - 2, 0); k < it; k++) {
            const int b = k & 1;
            const uint32_t par = (k >> 1) & 1;
            if (!mbar_try(bar_prefix + 8 * b, par))
                while (!mbar_try(bar_prefix + 8 * b, par)) __nanosleep(64);
            drain(k, base[b * 8 + warp], cnts[b * 8 + warp]);
        }
    }
}

// 2-D view of the capture stack for the tile loads: inner dimension = one frame as 64-bit words,
// outer = the frames; box = [16 words = 128 B] x [NF frames], landing in shared memory as a dense [NF][128] block.
static bool stack_tensor_map8(CUtensorMap* map, const uint8_t* stack, size_t plane, int NF)
{
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    static bool looked_up = false;
    if (!looked_up) {
        looked_up = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            encode = (encode_fn)fn;
    }
    if (!encode || NF > 256 || (plane & 15) || ((uintptr_t)stack & 15)) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)(plane / 8), (cuuint64_t)NF};
    const cuuint64_t gstride[1] = {(cuuint64_t)plane};
    const cuuint32_t box[2] = {16u, (cuuint32_t)NF};
    const cuuint32_t estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<uint8_t*>(stack), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int N, int DIRS, int CW, int MINB, bool EXACT>
static cudaError_t launch8_t(const FusedArgs& a, const DeviceCalib& cal, int sm_count, const Plan8& p, Fused8Cache* cache,
                             uint32_t* advance, cudaStream_t st)
{
    auto kern = k_fused8<N, DIRS, CW, MINB, EXACT>;
    static bool configured = false;     // per instantiation: attributes and occupancy are set once per process
    static int per_sm_cached = 0;
    static size_t smem_cached = 0;
    if (!configured || smem_cached != p.smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (CW + 1) * 32, p.smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        if (getenv("SCAN3D_DEBUG")) {
            cudaFuncAttributes fa;
            cudaFuncGetAttributes(&fa, kern);
            fprintf(stderr, "k_fused8<%d,%d,%d,%d,%d>: occupancy %d CTAs/SM, %d regs, %zu B dyn smem, local %zu B\n", N, DIRS, CW, MINB,
                    (int)EXACT, per_sm, fa.numRegs, p.smem, fa.localSizeBytes);
        }
        per_sm_cached = per_sm > MINB ? MINB : per_sm;
        smem_cached = p.smem;
        configured = true;
    }
    const int grid = a.n_tiles < sm_count * per_sm_cached ? a.n_tiles : sm_count * per_sm_cached;   // all CTAs resident
    FusedArgs a2 = a;
    const int NF = DIRS == 2 ? 2 * N + 2 * (a.M_v + a.M_h) : N + 2 * a.M_v;
    // the tensor map depends on the stack's address only: keep the last few (a ring of resident stacks is the common case)
    const CUtensorMap* map = nullptr;
    static const bool no_tmap = getenv("SCAN3D_NO_TMAP") != nullptr;
    if (!no_tmap) {
        for (int i = 0; i < Fused8Cache::SLOTS; i++)
            if (cache->stack[i] == a.stack) map = &cache->map[i];
        if (!map) {
            const int i = cache->next++ % Fused8Cache::SLOTS;
            cache->stack[i] = nullptr;
            if (stack_tensor_map8(&cache->map[i], a.stack, (size_t)a.W * a.H, NF)) {
                cache->stack[i] = a.stack;
                map = &cache->map[i];
            }
        }
    }
    alignas(64) CUtensorMap dummy;
    if (!map) { memset(&dummy, 0, sizeof(dummy)); map = &dummy; }
    a2.use_tmap = map != &dummy;
    kern<<<grid, (CW + 1) * 32, p.smem, st>>>(a2, cal, *map);
    *advance = (uint32_t)a.n_tiles + (uint32_t)grid;     // every CTA draws one position past the end
    return cudaGetLastError();
}

template <int N, int DIRS>
static cudaError_t launch8_nd(const FusedArgs& a, const DeviceCalib& cal, int sm_count, const Plan8& p, bool exact,
                              Fused8Cache* cache, uint32_t* advance, cudaStream_t st)
{
#define S3D_CASE8(MB)                                                                                          \
    if (p.minb == MB) {                                                                                        \
        if (exact || DIRS == 1) return launch8_t<N, DIRS, 7, MB, true>(a, cal, sm_count, p, cache, advance, st); \
        return launch8_t<N, DIRS, 7, MB, (DIRS == 1)>(a, cal, sm_count, p, cache, advance, st);                  \
    }
    S3D_CASE8(3) S3D_CASE8(2)
#undef S3D_CASE8
    return cudaErrorInvalidValue;
}

cudaError_t launch_fused8(const scan3d_config& c, const FusedArgs& a_in, const DeviceCalib& cal, int sm_count,
                          Fused8Cache* cache, uint32_t* advance, cudaStream_t st)
{
    Plan8 p;
    if (!plan8(c, &p)) return cudaErrorInvalidValue;
    if (((uintptr_t)a_in.stack & 15) || ((uintptr_t)a_in.roi & 15)) return cudaErrorMisalignedAddress;
    FusedArgs a = a_in;
    const int T = 128 * p.cw;
    a.tiles_per_row = 0;
    a.n_tiles = (int)(((size_t)c.W * c.H + T - 1) / T);
    a.dynamic = 1;
    const bool exact = !(c.flags & SCAN3D_FLAG_FAST_TRIANGULATION);
#define S3D_F8(NN) (c.dirs == 2 ? launch8_nd<NN, 2>(a, cal, sm_count, p, exact, cache, advance, st) : launch8_nd<NN, 1>(a, cal, sm_count, p, exact, cache, advance, st))
    switch (c.N) {
        case 3: return S3D_F8(3);
        case 4: return S3D_F8(4);
        case 5: return S3D_F8(5);
        case 8: return S3D_F8(8);
    }
#undef S3D_F8
    return cudaErrorInvalidValue;
}

}  // namespace s3d
