// scan3d_aux_kernels.cu -- sm_100a kernels for the steps either side of the reconstruction path
// (SURVEY.md 8 f2 / f4).  All are byte/element streaming work, bounded by HBM:
//
//   k_undistort_map   cv::undistort's fixed-point map, once per calibration and device
//                     (2/project_pattern.cpp:220: the reference rebuilds it for every captured frame)
//   k_remap_tiled     cv::remap of a whole captured stack through that map: the map entry of a pixel is read once
//                     and applied to every frame (F bytes in + F bytes out + 6 B map per pixel).  One CTA per
//                     8 x 256 output tile; per frame the tile's source box (the map is close to the identity) is
//                     staged in shared memory with 16-byte cp.async copies, double-buffered over the frame loop,
//                     and blended from there: ~10 instructions per output byte (scan3d_aux_math.h)
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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifndef S3D_REMAP_STAGES
#define S3D_REMAP_STAGES 6
#endif
constexpr int REMAP_STAGES = S3D_REMAP_STAGES;
// One CTA (256 threads) per output tile of REMAP_TILE_H x REMAP_TILE_W pixels; thread t owns the 4-pixel groups
// (row t/64, columns 4*(t%64)..+3) and (row t/64 + 4, same columns).  W % 16 == 0.
__global__ void __launch_bounds__(256, 3) k_remap_tiled(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                      const short2* __restrict__ map_xy, const uint16_t* __restrict__ map_frac,
                                                      int W, int H, int n_frames)
{
    // ring of staged source boxes: REMAP_STAGES - 1 frames are in flight per CTA while one is blended (the kernel is
    // bound by the bytes it keeps in flight: 2 buffers gave 0.31 of the HBM roofline) (+ one vector per buffer: the
    // taps are cut out of word pairs)
    constexpr int NS = REMAP_STAGES;
    __shared__ __align__(16) uint8_t box[NS][s3a::REMAP_BOX_H * s3a::REMAP_BOX_W + 16];
    __shared__ int ext[4];   // min sx, max sx, min sy, max sy over the tile
    const size_t plane = (size_t)W * H;
    const int tiles_x = (W + s3a::REMAP_TILE_W - 1) / s3a::REMAP_TILE_W;
    const int tx0 = (blockIdx.x % tiles_x) * s3a::REMAP_TILE_W, ty0 = (blockIdx.x / tiles_x) * s3a::REMAP_TILE_H;
    const int t = threadIdx.x;
    short2 xy[2][4];
    int fr[2][4];
    bool have[2];
    size_t pix[2];
    int lo_x = 0x7fffffff, hi_x = -0x7fffffff, lo_y = 0x7fffffff, hi_y = -0x7fffffff;
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
    if (t == 0) { ext[0] = 0x7fffffff; ext[1] = -0x7fffffff; ext[2] = 0x7fffffff; ext[3] = -0x7fffffff; }
    __syncthreads();
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo_x = min(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o)); hi_x = max(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
        lo_y = min(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o)); hi_y = max(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
    }
    if ((t & 31) == 0) { atomicMin(&ext[0], lo_x); atomicMax(&ext[1], hi_x); atomicMin(&ext[2], lo_y); atomicMax(&ext[3], hi_y); }
    __syncthreads();
    const s3a::RemapBox b = s3a::remap_tile_box(ext[0], ext[1], ext[2], ext[3], W, H);

    if (!b.ok) {   // strong distortion (the box does not fit): per-tap gathers from global memory, as k_remap_frames
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

    // the box's 16-byte vectors, at most 2 per thread (16 rows x 18 vectors = 288): where each lands in the buffer and
    // where it comes from in a frame -- the same for every frame
    const int vec_per_row = b.w >> 4, n_vec = b.rows * vec_per_row;
    int v_dst[2], v_kind[2];          // kind: 0 = none, 1 = copy, 2 = outside the image (constant border: zeros)
    long long v_src[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
        const int v = t + 256 * q;
        v_kind[q] = 0; v_dst[q] = 0; v_src[q] = 0;
        if (v < n_vec) {
            const int r = v / vec_per_row, c = v - r * vec_per_row;
            v_dst[q] = r * s3a::REMAP_BOX_W + 16 * c;
            v_src[q] = (long long)(b.y0 + r) * W + (b.x0 + 16 * c);
            v_kind[q] = s3a::remap_box_vector_inside(b, r, c, W, H) ? 1 : 2;
        }
    }
    auto stage = [&](int f, int buf) {
        const uint8_t* s = src + (size_t)f * plane;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            if (v_kind[q] == 1) cp_async16(&box[buf][v_dst[q]], s + v_src[q]);
            else if (v_kind[q] == 2) *reinterpret_cast<uint4*>(&box[buf][v_dst[q]]) = make_uint4(0, 0, 0, 0);
        }
        cp_async_commit();
    };
    // frame-independent per pixel: offset of its first tap in the box, the 4 weights as two packed pairs
    int off[2][4];
    uint32_t wA[2][4], wB[2][4];
#pragma unroll
    for (int g = 0; g < 2; g++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            off[g][k] = have[g] ? s3a::remap_box_offset(b, xy[g][k].x, xy[g][k].y) : 0;
            s3a::bilinear_weight_pairs(fr[g][k], &wA[g][k], &wB[g][k]);
        }

    if (n_frames <= 0) return;
    // every iteration commits exactly one copy group (an empty one past the last frame), so "at most NS - 2 groups
    // pending" always means "frame f has landed"
    for (int f = 0; f < NS - 1; f++) {
        if (f < n_frames) stage(f, f);
        else cp_async_commit();
    }
    for (int f = 0; f < n_frames; f++) {
        cp_async_wait<NS - 2>();
        __syncthreads();               // frame f is visible to everybody, and everybody is done with frame f - 1 ...
        if (f + NS - 1 < n_frames) stage(f + NS - 1, (f + NS - 1) % NS);      // ... whose buffer this refills
        else cp_async_commit();
        const uint8_t* sb = box[f % NS];
#pragma unroll
        for (int g = 0; g < 2; g++) {
            if (!have[g]) continue;
            uint32_t packed = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
                packed |= s3a::bilinear_u8_pairs(wA[g][k], wB[g][k], s3a::box_taps(sb, off[g][k]),
                                                 s3a::box_taps(sb, off[g][k] + s3a::REMAP_BOX_W)) << (8 * k);
            *reinterpret_cast<uint32_t*>(dst + (size_t)f * plane + pix[g]) = packed;
        }
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
    if (W % 16 == 0 && (((uintptr_t)src | (uintptr_t)dst) % 16 == 0)) {
        const int tiles = ((W + s3a::REMAP_TILE_W - 1) / s3a::REMAP_TILE_W) * ((H + s3a::REMAP_TILE_H - 1) / s3a::REMAP_TILE_H);
        k_remap_tiled<<<tiles, 256, 0, st>>>(src, dst, map_xy, map_frac, W, H, n_frames);
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
