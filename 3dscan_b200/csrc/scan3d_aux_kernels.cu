// scan3d_aux_kernels.cu -- sm_100a kernels for the steps either side of the reconstruction path
// (SURVEY.md 8 f2 / f4).  All are byte/element streaming work, bounded by HBM:
//
//   k_undistort_map   cv::undistort's fixed-point map, once per calibration and device
//                     (2/project_pattern.cpp:220: the reference rebuilds it for every captured frame)
//   k_remap_frames    cv::remap of a whole captured stack through that map: the map entry of a pixel is
//                     read once and applied to every frame (F bytes in + F bytes out + 6 B map per pixel)
//   k_roi_fill        image_scissor's scan-line fill (m_tech_project_console.cpp:186-229)
//   k_register_points register_point_clouds' rigid transform (9/register_point_clouds.cpp:117-137)
#include "scan3d_internal.h"
#include "../common/scan3d_aux_math.h"

// Experimental variants of k_remap_frames (tools/build_variants.py; off in the default build):
//   S3D_VAR_REMAP_UNROLL=n  frame loop unrolled n times (loads of n frames in flight per thread)
//   S3D_VAR_REMAP_WINDOW    two aligned 8-byte loads per source row instead of 8 one-byte gathers when a thread's
//                           4 pixels allow it (scan3d_aux_math.h, "windowed gather"; 96 % of the groups at 12 MP)
#ifndef S3D_VAR_REMAP_UNROLL
#define S3D_VAR_REMAP_UNROLL 1
#endif
#ifndef S3D_VAR_REMAP_WINDOW
#define S3D_VAR_REMAP_WINDOW 0
#endif
#define S3D_PRAGMA_(x) _Pragma(#x)
#if S3D_VAR_REMAP_UNROLL > 1
#define S3D_UNROLL_N_(n) S3D_PRAGMA_(unroll n)
#define S3D_FRAME_UNROLL S3D_UNROLL_N_(S3D_VAR_REMAP_UNROLL)
#else
#define S3D_FRAME_UNROLL
#endif

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
#if S3D_VAR_REMAP_WINDOW
        if (VEC == 4) {
            int16_t xy16[8];
#pragma unroll
            for (int k = 0; k < 4; k++) { xy16[2 * k] = xy[k].x; xy16[2 * k + 1] = xy[k].y; }
            const s3a::RemapGroup grp = s3a::remap_group_prepare(xy16, W, H, (W & 7) == 0 && ((uintptr_t)src & 7) == 0);
            if (grp.fast) {
                S3D_FRAME_UNROLL
                for (int f = 0; f < n_frames; f++) {
                    const uint8_t* s = src + (size_t)f * plane + grp.base;
                    const uint2 a0 = __ldg(reinterpret_cast<const uint2*>(s)), a1 = __ldg(reinterpret_cast<const uint2*>(s + 8));
                    const uint2 b0 = __ldg(reinterpret_cast<const uint2*>(s + W)), b1 = __ldg(reinterpret_cast<const uint2*>(s + W + 8));
                    const uint32_t r0[4] = {a0.x, a0.y, a1.x, a1.y}, r1[4] = {b0.x, b0.y, b1.x, b1.y};
                    *reinterpret_cast<uint32_t*>(dst + (size_t)f * plane + p) = s3a::remap_group_blend(grp, r0, r1, fr);
                }
                continue;
            }
        }
#endif
        S3D_FRAME_UNROLL
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
