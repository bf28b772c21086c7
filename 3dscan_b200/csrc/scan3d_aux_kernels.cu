// scan3d_aux_kernels.cu -- sm_100a kernels for the steps either side of the reconstruction path
// (SURVEY.md 8 f2 / f4).  All are byte/element streaming work, bounded by HBM:
//
//   k_undistort_map   cv::undistort's fixed-point map, once per calibration and device
//                     (2/project_pattern.cpp:220: the reference rebuilds it for every captured frame)
//   k_remap_tiled     cv::remap of a whole captured stack through that map: the map entry of a pixel is read once
//                     and applied to every frame (F bytes in + F bytes out + 6 B map per pixel).  One CTA per
//                     8 x 256 output tile: a copy warp stages the tile's source box (the map is close to the
//                     identity) frame after frame into a ring of shared-memory stages (cp.async vectors, completion
//                     on mbarriers), 8 warps blend from there: 2 loads + 1 byte permute per pixel pair and source row,
//                     two-way dot products with doubled weights (scan3d_aux_math.h), ~7 instructions per output byte
//   k_remap_frames    the same through per-tap global gathers: any width, any distortion (fallback)
//   k_roi_fill        image_scissor's scan-line fill (m_tech_project_console.cpp:186-229)
//   k_register_points register_point_clouds' rigid transform (9/register_point_clouds.cpp:117-137)
#include "scan3d_internal.h"
#include "../common/scan3d_aux_math.h"

namespace s3d {

struct AuxCalib {
    double K[9];
    double d[5];
};

// One thread per row: initUndistortRectifyMap accumulates _x += ir[0] along the row, so a row is a
// sequential chain; rows are independent.  Once per calibration.
__global__ void __launch_bounds__(64) k_undistort_map(AuxCalib c, int W, int H, short2* __restrict__ map_xy,
                                                       uint16_t* __restrict__ map_frac)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= H) return;
    s3a::undistort_map_row(c.K, c.d, W, H, row, reinterpret_cast<int16_t*>(map_xy + (size_t)row * W),
                           map_frac + (size_t)row * W);
}

__device__ __forceinline__ int tap(const uint8_t* __restrict__ src, int W, int H, int x, int y)
{
    return ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) ? (int)__ldg(src + (size_t)y * W + x) : 0;
}

// VEC output pixels per thread (4 when W % 4 == 0: one 32-bit store per frame, 16 B + 8 B map loads).
template <int VEC>
__global__ void __launch_bounds__(256) k_remap_frames(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                       const short2* __restrict__ map_xy,
                                                       const uint16_t* __restrict__ map_frac, int W, int H, int n_frames)
{
    const size_t plane = (size_t)W * H;
    const size_t groups = plane / VEC;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
        const size_t p = g * VEC;
        short2 xy[VEC];
        int fr[VEC];
        if (VEC == 4) {
            const int4 m = __ldg(reinterpret_cast<const int4*>(map_xy + p));
            const int mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                xy[k].x = (short)(mm[k] & 0xffff);
                xy[k].y = (short)(mm[k] >> 16);
            }
            const uint2 f = __ldg(reinterpret_cast<const uint2*>(map_frac + p));
            fr[0] = f.x & 0xffff; fr[1] = f.x >> 16; fr[2] = f.y & 0xffff; fr[3] = f.y >> 16;
        } else {
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                xy[k] = map_xy[p + k];
                fr[k] = map_frac[p + k];
            }
        }
        for (int f = 0; f < n_frames; f++) {
            const uint8_t* s = src + (size_t)f * plane;
            uint32_t packed = 0;
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                const int x = xy[k].x, y = xy[k].y;
                const int v0 = tap(s, W, H, x, y), v1 = tap(s, W, H, x + 1, y);
                const int v2 = tap(s, W, H, x, y + 1), v3 = tap(s, W, H, x + 1, y + 1);
                packed |= (uint32_t)s3a::bilinear_u8(v0, v1, v2, v3, fr[k]) << (8 * k);
            }
            if (VEC == 4)
                *reinterpret_cast<uint32_t*>(dst + (size_t)f * plane + p) = packed;
            else
                dst[(size_t)f * plane + p] = (uint8_t)packed;
        }
    }
}

// ---- PTX helpers of the staged remap (mbarrier pipeline fed by bulk copies) -------------------------------------
__device__ __forceinline__ uint32_t rm_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rm_bar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void rm_bar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void rm_bar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "RM_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra RM_WAIT_%=;\n\t}"
        ::"r"(bar), "r"(parity)
        : "memory");
}

// the four words of a pair window: {base, base + 4} of the row and of the row below (immediate offsets: one address
// computation per window)
__device__ __forceinline__ void rm_window(uint32_t addr, uint32_t& t0, uint32_t& t1, uint32_t& b0, uint32_t& b1)
{
    static_assert(s3a::REMAP_BOX_W == 384, "the offsets below are the row pitch");
    asm volatile("ld.shared.u32 %0, [%4];\n\tld.shared.u32 %1, [%4+4];\n\tld.shared.u32 %2, [%4+384];\n\tld.shared.u32 %3, [%4+388];"
                 : "=r"(t0), "=r"(t1), "=r"(b0), "=r"(b1)
                 : "r"(addr)
                 : "memory");
}

// Measured alternatives (profiles/r2_optimisation_log.md section 5): one cp.async.bulk per box row instead of cp.async
// vectors 435 us (the copy engine serves ~1 small request per 35 cycles and SM); 4 / 8 stages 408 / 394 us; 2 / 4 CTAs
// per SM 584 / 418 us (4: 56 registers, spills).
constexpr int REMAP_STAGES = 6;
constexpr int REMAP_CTAS_PER_SM = 3;
constexpr int REMAP_CONSUMERS = 256;                                // 8 blending warps
constexpr int REMAP_THREADS = REMAP_CONSUMERS + 32;                 // + 1 copy warp
constexpr int REMAP_STAGE_BYTES = s3a::REMAP_BOX_H * s3a::REMAP_BOX_W + 128;   // + padding: a window's second word may lie past the last row
constexpr int REMAP_SMEM_BYTES = REMAP_STAGES * REMAP_STAGE_BYTES + 2 * REMAP_STAGES * 8 + 16;

// One CTA per output tile of REMAP_TILE_H x REMAP_TILE_W pixels and all the frames of the stack.  Warp 8 copies the
// tile's source box of frame after frame into a ring of shared-memory stages (16-byte cp.async vectors, one
// instruction per box row; the lanes' arrivals on the stage's "full" mbarrier fire when their copies have landed);
// warps 0..7 blend: thread t owns the 4-pixel groups (row t/64, columns 4*(t%64)..+3) and (row t/64 + 4, same
// columns), waits for the stage, cuts the taps of each pixel PAIR out of two aligned words per source row
// (scan3d_aux_math.h: remap_pair_window) and stores 4 output bytes per group; a warp hands the stage back through
// its "empty" mbarrier.  No CTA-wide barrier inside the frame loop.  W % 16 == 0.
__global__ void __launch_bounds__(REMAP_THREADS, REMAP_CTAS_PER_SM) k_remap_tiled(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                const short2* __restrict__ map_xy,
                                                                const uint16_t* __restrict__ map_frac, int W, int H, int n_frames)
{
    constexpr int NS = REMAP_STAGES;
    extern __shared__ __align__(128) uint8_t rm_smem_raw[];
    uint8_t* box = rm_smem_raw;                                                   // [NS][REMAP_STAGE_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(box + NS * REMAP_STAGE_BYTES);   // full[NS], empty[NS]
    int* ext = reinterpret_cast<int*>(bars + 2 * NS);                             // min sx, max sx, min sy, max sy over the tile
    const size_t plane = (size_t)W * H;
    const int tiles_x = (W + s3a::REMAP_TILE_W - 1) / s3a::REMAP_TILE_W;
    const int tx0 = (blockIdx.x % tiles_x) * s3a::REMAP_TILE_W, ty0 = (blockIdx.x / tiles_x) * s3a::REMAP_TILE_H;
    const int t = threadIdx.x;
    const bool consumer = t < REMAP_CONSUMERS;
    short2 xy[2][4];
    int fr[2][4];
    bool have[2] = {false, false};
    size_t pix[2] = {0, 0};
    int lo_x = 0x7fffffff, hi_x = -0x7fffffff, lo_y = 0x7fffffff, hi_y = -0x7fffffff;
    if (consumer) {
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const int x = tx0 + 4 * (t & 63), y = ty0 + (t >> 6) + 4 * g;
            have[g] = x < W && y < H;     // W % 4 == 0: a group is inside or outside as a whole
            pix[g] = (size_t)y * W + x;
            if (have[g]) {
                const int4 m = __ldg(reinterpret_cast<const int4*>(map_xy + pix[g]));
                const int mm[4] = {m.x, m.y, m.z, m.w};
                const uint2 f = __ldg(reinterpret_cast<const uint2*>(map_frac + pix[g]));
                fr[g][0] = f.x & 0xffff; fr[g][1] = f.x >> 16; fr[g][2] = f.y & 0xffff; fr[g][3] = f.y >> 16;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    xy[g][k].x = (short)(mm[k] & 0xffff);
                    xy[g][k].y = (short)(mm[k] >> 16);
                    lo_x = min(lo_x, (int)xy[g][k].x); hi_x = max(hi_x, (int)xy[g][k].x);
                    lo_y = min(lo_y, (int)xy[g][k].y); hi_y = max(hi_y, (int)xy[g][k].y);
                }
            }
        }
    }
    if (t == 0) {
        ext[0] = 0x7fffffff; ext[1] = -0x7fffffff; ext[2] = 0x7fffffff; ext[3] = -0x7fffffff;
        for (int s = 0; s < NS; s++) {
            rm_bar_init(rm_smem(bars + s), 32);                         // full: one arrival per lane of the copy warp
            rm_bar_init(rm_smem(bars + NS + s), REMAP_CONSUMERS / 32);  // empty: one arrival per blending warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo_x = min(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o)); hi_x = max(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
        lo_y = min(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o)); hi_y = max(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
    }
    if ((t & 31) == 0 && consumer) { atomicMin(&ext[0], lo_x); atomicMax(&ext[1], hi_x); atomicMin(&ext[2], lo_y); atomicMax(&ext[3], hi_y); }
    __syncthreads();
    const s3a::RemapBox b = s3a::remap_tile_box(ext[0], ext[1], ext[2], ext[3], W, H);

    if (!b.ok) {   // strong distortion (the box does not fit): per-tap gathers from global memory, as k_remap_frames
        if (!consumer) return;
        for (int f = 0; f < n_frames; f++) {
            const uint8_t* s = src + (size_t)f * plane;
#pragma unroll
            for (int g = 0; g < 2; g++) {
                if (!have[g]) continue;
                uint32_t packed = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int x = xy[g][k].x, y = xy[g][k].y;
                    packed |= (uint32_t)s3a::bilinear_u8(tap(s, W, H, x, y), tap(s, W, H, x + 1, y), tap(s, W, H, x, y + 1),
                                                         tap(s, W, H, x + 1, y + 1), fr[g][k]) << (8 * k);
                }
                *reinterpret_cast<uint32_t*>(dst + (size_t)f * plane + pix[g]) = packed;
            }
        }
        return;
    }

    // a box that reaches outside the image: the border bytes are zeros in every stage, written once (the copies
    // never touch them); inside the image every staged byte is overwritten per frame and needs no initial value
    if (b.x0 < 0 || b.y0 < 0 || b.x0 + b.w > W || b.y0 + b.rows > H) {
        for (int i = t; i < NS * REMAP_STAGE_BYTES / 16; i += REMAP_THREADS) reinterpret_cast<uint4*>(box)[i] = make_uint4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (n_frames <= 0) return;

    if (!consumer) {
        const int lane = t & 31;
        int s = 0;
        uint32_t ph = 1;      // parity of the "empty" phase that precedes the stage's first use: passes at once
        // ---------------- copy warp: lane c owns the box's 16-byte vector column c, one cp.async per box row ----------------
        // (rows and vector columns outside the image are never written: they hold the zero border)
        const int xv = b.x0 + 16 * lane;
        const bool lane_in = lane < (b.w >> 4) && xv >= 0 && xv + 16 <= W;
        const int r_lo = b.y0 < 0 ? -b.y0 : 0, r_hi = b.y0 + b.rows > H ? H - b.y0 : b.rows;
        const uint8_t* from = src + (long long)(b.y0 + r_lo) * W + (lane_in ? xv : 0);
        const uint32_t d0 = rm_smem(box) + (uint32_t)(r_lo * s3a::REMAP_BOX_W + 16 * lane);
        uint32_t sbase = d0;
        for (int f = 0; f < n_frames; f++) {
            const uint32_t full = rm_smem(bars + s), empty = rm_smem(bars + NS + s);
            rm_bar_wait(empty, ph);
            if (lane_in) {
                const uint8_t* g = from;
                uint32_t d = sbase;
#pragma unroll 4
                for (int r = r_lo; r < r_hi; r++) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
                    g += W;
                    d += s3a::REMAP_BOX_W;
                }
            }
            // the lane's arrival on "full" fires when its copies above have landed
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full) : "memory");
            from += plane;
            sbase += REMAP_STAGE_BYTES;
            if (++s == NS) { s = 0; ph ^= 1u; sbase = d0; }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        return;
    }

    // ---------------- blending warps ----------------
    // frame-independent per pixel pair: byte address of its first window word in stage 0, the permute selector, the
    // doubled weight pairs; per thread 4 pairs.  A pair that does not qualify keeps its second pixel's own address.
    uint32_t w_off[2][2], w_sel[2][2], own_off[2][2];
    uint32_t wA[2][4], wB[2][4];
    uint32_t irregular = 0;
#pragma unroll
    for (int g = 0; g < 2; g++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int off0 = have[g] ? s3a::remap_box_offset(b, xy[g][2 * j].x, xy[g][2 * j].y) : 0;
            const int off1 = have[g] ? s3a::remap_box_offset(b, xy[g][2 * j + 1].x, xy[g][2 * j + 1].y) : 0;
            int base;
            uint32_t sel;
            if (!s3a::remap_pair_window(off0, off1, &base, &sel)) irregular |= 1u << (2 * g + j);
            w_off[g][j] = (uint32_t)base;
            w_sel[g][j] = sel;
            own_off[g][j] = (uint32_t)(off1 & ~3) | ((uint32_t)(off1 & 3) << 30);   // box offsets are far below 2^30
            s3a::bilinear_weight_pairs_x2(fr[g][2 * j], &wA[g][2 * j], &wB[g][2 * j]);
            s3a::bilinear_weight_pairs_x2(fr[g][2 * j + 1], &wA[g][2 * j + 1], &wB[g][2 * j + 1]);
        }
    // pair slots in which SOME lane of the warp is irregular: the fix-up below is skipped warp-wide for the others
    uint32_t warp_irregular = irregular;
#pragma unroll
    for (int o = 16; o; o >>= 1) warp_irregular |= __shfl_xor_sync(0xffffffffu, warp_irregular, o);

    uint8_t* out0 = dst + pix[0];
    uint8_t* out1 = dst + pix[1];
    int s = 0;
    uint32_t ph = 0;
    const uint32_t box0 = rm_smem(box);
    uint32_t sb = box0;                            // shared-memory address of the current stage
    for (int f = 0; f < n_frames; f++) {
        rm_bar_wait(rm_smem(bars + s), ph);
#pragma unroll
        for (int g = 0; g < 2; g++) {
            if (!have[g]) continue;
            uint32_t acc[4];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                uint32_t t0, t1, b0, b1;
                rm_window(sb + w_off[g][j], t0, t1, b0, b1);
                const uint32_t top = s3a::permute_bytes(t0, t1, w_sel[g][j]);
                const uint32_t bot = s3a::permute_bytes(b0, b1, w_sel[g][j]);
                acc[2 * j] = s3a::blend_acc_x2(wA[g][2 * j], wB[g][2 * j], top, bot, false);
                acc[2 * j + 1] = s3a::blend_acc_x2(wA[g][2 * j + 1], wB[g][2 * j + 1], top, bot, true);
                if (warp_irregular & (1u << (2 * g + j))) {
                    if (irregular & (1u << (2 * g + j))) {
                        const uint32_t o = own_off[g][j] >> 30, osel = o | ((o + 1) << 4);
                        rm_window(sb + (own_off[g][j] & 0x3fffffffu), t0, t1, b0, b1);
                        acc[2 * j + 1] = s3a::blend_acc_x2(wA[g][2 * j + 1], wB[g][2 * j + 1], s3a::permute_bytes(t0, t1, osel),
                                                           s3a::permute_bytes(b0, b1, osel), false);
                    }
                }
            }
            *reinterpret_cast<uint32_t*>(g == 0 ? out0 : out1) = s3a::pack_acc_bytes(acc[0], acc[1], acc[2], acc[3]);
        }
        out0 += plane;
        out1 += plane;
        __syncwarp();
        if ((t & 31) == 0) rm_bar_arrive(rm_smem(bars + NS + s));
        sb += REMAP_STAGE_BYTES;
        if (++s == NS) { s = 0; ph ^= 1u; sb = box0; }
    }
}

// One CTA per row.  The reference's nested search (start pixel, next non-zero pixel, fill between,
// restart AT the end pixel) fills every zero pixel that lies strictly between the first and the last
// non-zero pixel of the row; outline pixels themselves stay unselected.
__global__ void __launch_bounds__(256) k_roi_fill(const uint8_t* outline, int W, uint8_t* __restrict__ roi,
                                                   uint8_t* filled)
{
    __shared__ int s_first, s_last;
    const int row = blockIdx.x;
    const uint8_t* o = outline + (size_t)row * W;
    if (threadIdx.x == 0) { s_first = 0x7fffffff; s_last = -1; }
    __syncthreads();
    int first = 0x7fffffff, last = -1;
    for (int c = threadIdx.x; c < W; c += blockDim.x)
        if (o[c] != 0) {
            first = min(first, c);
            last = max(last, c);
        }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        first = min(first, __shfl_xor_sync(0xffffffffu, first, off));
        last = max(last, __shfl_xor_sync(0xffffffffu, last, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&s_first, first);
        atomicMax(&s_last, last);
    }
    __syncthreads();
    first = s_first;
    last = s_last;
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        const uint8_t v = o[c];
        const bool in = v == 0 && c > first && c < last;
        roi[(size_t)row * W + c] = in ? 1 : 0;
        if (filled) filled[(size_t)row * W + c] = in ? 255 : v;
    }
}

struct RegArgs {
    float R[16];
    float tx, ty, tz;
};

__global__ void __launch_bounds__(256) k_register_points(const float* src, float* dst,
                                                          long long n, RegArgs a)
{
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        float x = src[3 * k], y = src[3 * k + 1], z = src[3 * k + 2];
        s3a::register_point(a.R, a.tx, a.ty, a.tz, x, y, z);
        dst[3 * k] = x;
        dst[3 * k + 1] = y;
        dst[3 * k + 2] = z;
    }
}

// ---- launchers ----------------------------------------------------------------------------
cudaError_t launch_undistort_map(const double K[9], const double d[5], int W, int H, short2* map_xy, uint16_t* map_frac,
                                 cudaStream_t st)
{
    AuxCalib c;
    for (int k = 0; k < 9; k++) c.K[k] = K[k];
    for (int k = 0; k < 5; k++) c.d[k] = d[k];
    k_undistort_map<<<(H + 63) / 64, 64, 0, st>>>(c, W, H, map_xy, map_frac);
    return cudaGetLastError();
}

cudaError_t launch_remap_frames(const uint8_t* src, uint8_t* dst, const short2* map_xy, const uint16_t* map_frac, int W,
                                int H, int n_frames, int sm_count, cudaStream_t st)
{
    const size_t plane = (size_t)W * H;
    // a single frame (the capture loop's per-image call) does not amortise the tile set-up: 48.6 us through the
    // per-tap gathers against 53.9 us staged at 12 MP; from two frames on the staged kernel wins (6 us per further frame
    // against 23 us)
    if (n_frames >= 2 && W % 16 == 0 && (((uintptr_t)src | (uintptr_t)dst) % 16 == 0)) {
        const int tiles = ((W + s3a::REMAP_TILE_W - 1) / s3a::REMAP_TILE_W) * ((H + s3a::REMAP_TILE_H - 1) / s3a::REMAP_TILE_H);
        static bool attr_set[64] = {};    // the attribute is per device
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            e = cudaFuncSetAttribute(k_remap_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, REMAP_SMEM_BYTES);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        k_remap_tiled<<<tiles, REMAP_THREADS, REMAP_SMEM_BYTES, st>>>(src, dst, map_xy, map_frac, W, H, n_frames);
        return cudaGetLastError();
    }
    const bool vec = (plane % 4 == 0) && (((uintptr_t)src | (uintptr_t)dst) % 4 == 0);
    const size_t groups = vec ? plane / 4 : plane;
    size_t blocks = (groups + 255) / 256;
    const size_t cap = (size_t)sm_count * 8;   // a multiple of the SM count; grid-stride above it
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (vec)
        k_remap_frames<4><<<(unsigned)blocks, 256, 0, st>>>(src, dst, map_xy, map_frac, W, H, n_frames);
    else
        k_remap_frames<1><<<(unsigned)blocks, 256, 0, st>>>(src, dst, map_xy, map_frac, W, H, n_frames);
    return cudaGetLastError();
}

cudaError_t launch_roi_fill(const uint8_t* outline, int W, int H, uint8_t* roi, uint8_t* filled, cudaStream_t st)
{
    k_roi_fill<<<H, 256, 0, st>>>(outline, W, roi, filled);
    return cudaGetLastError();
}

cudaError_t launch_register_points(const float* src, float* dst, long long n, const float R[16], float tx, float ty,
                                   float tz, int sm_count, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    RegArgs a;
    for (int k = 0; k < 16; k++) a.R[k] = R[k];
    a.tx = tx; a.ty = ty; a.tz = tz;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_register_points<<<(unsigned)blocks, 256, 0, st>>>(src, dst, n, a);
    return cudaGetLastError();
}

}  // namespace s3d
