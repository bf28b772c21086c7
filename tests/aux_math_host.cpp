// Host build of 3dscan_b200/common/scan3d_aux_math.h (the very expressions the CUDA kernels of
// scan3d_aux_kernels.cu evaluate), so that the CPU test suite can compare them with the oracle
// without a GPU.  Built by tests/test_aux_math_host.py with g++ -O2 -ffp-contract=off.
#include <stddef.h>
#include <string.h>

#include "../3dscan_b200/common/scan3d_aux_math.h"

extern "C" {

void s3a_host_undistort_map(const double* K, const double* d, int W, int H, int16_t* xy, uint16_t* frac)
{
    for (int row = 0; row < H; row++)   // k_undistort_map: one thread per row
        s3a::undistort_map_row(K, d, W, H, row, xy + (size_t)row * W * 2, frac + (size_t)row * W);
}

static inline int tap(const uint8_t* src, int W, int H, int x, int y)
{
    return ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) ? (int)src[(size_t)y * W + x] : 0;
}

void s3a_host_remap(const uint8_t* src, int W, int H, const int16_t* xy, const uint16_t* frac, uint8_t* dst)
{
    for (size_t p = 0; p < (size_t)W * H; p++) {   // k_remap_frames: one thread per pixel (group)
        const int x = xy[2 * p], y = xy[2 * p + 1];
        dst[p] = s3a::bilinear_u8(tap(src, W, H, x, y), tap(src, W, H, x + 1, y), tap(src, W, H, x, y + 1),
                                  tap(src, W, H, x + 1, y + 1), frac[p]);
    }
}

// k_remap_tiled: per output tile the source box from the map's extremes, the box staged in a
// buffer with the kernel's pitch, taps gathered from the buffer; tiles whose box does not qualify take the per-tap path.
// Returns the number of tiles that were staged.
long long s3a_host_remap_tiled(const uint8_t* src, int W, int H, const int16_t* xy, const uint16_t* frac, uint8_t* dst)
{
    using namespace s3a;
    alignas(16) static uint8_t box[REMAP_BOX_H * REMAP_BOX_W + 16];
    long long staged = 0;
    for (int ty0 = 0; ty0 < H; ty0 += REMAP_TILE_H)
        for (int tx0 = 0; tx0 < W; tx0 += REMAP_TILE_W) {
            int lo_x = 0x7fffffff, hi_x = -0x7fffffff, lo_y = 0x7fffffff, hi_y = -0x7fffffff;
            for (int y = ty0; y < ty0 + REMAP_TILE_H && y < H; y++)
                for (int x = tx0; x < tx0 + REMAP_TILE_W && x < W; x++) {
                    const int sx = xy[2 * ((size_t)y * W + x)], sy = xy[2 * ((size_t)y * W + x) + 1];
                    lo_x = sx < lo_x ? sx : lo_x; hi_x = sx > hi_x ? sx : hi_x;
                    lo_y = sy < lo_y ? sy : lo_y; hi_y = sy > hi_y ? sy : hi_y;
                }
            const RemapBox b = remap_tile_box(lo_x, hi_x, lo_y, hi_y, W, H);
            if (b.ok) {
                staged++;
                memset(box, 0xAA, sizeof(box));      // anything not staged must never be read
                for (int r = 0; r < b.rows; r++)
                    for (int c = 0; c < b.w / 16; c++) {
                        uint8_t* d = box + r * REMAP_BOX_W + 16 * c;
                        if (remap_box_vector_inside(b, r, c, W, H)) memcpy(d, src + (long long)(b.y0 + r) * W + (b.x0 + 16 * c), 16);
                        else memset(d, 0, 16);
                    }
            }
            for (int y = ty0; y < ty0 + REMAP_TILE_H && y < H; y++)
                for (int x = tx0; x < tx0 + REMAP_TILE_W && x < W; x++) {
                    const size_t p = (size_t)y * W + x;
                    const int sx = xy[2 * p], sy = xy[2 * p + 1];
                    if (b.ok) {
                        // the kernel's blend: packed weight pairs, taps by funnel shift, two-way dot products
                        const int off = remap_box_offset(b, sx, sy);
                        uint32_t wA, wB;
                        bilinear_weight_pairs(frac[p], &wA, &wB);
                        dst[p] = (uint8_t)bilinear_u8_pairs(wA, wB, box_taps(box, off), box_taps(box, off + REMAP_BOX_W));
                    } else {
                        dst[p] = bilinear_u8(tap(src, W, H, sx, sy), tap(src, W, H, sx + 1, sy), tap(src, W, H, sx, sy + 1),
                                             tap(src, W, H, sx + 1, sy + 1), frac[p]);
                    }
                }
        }
    return staged;
}

void s3a_host_register_rotation(float theta_deg, float* R) { s3a::register_rotation(theta_deg, R); }

void s3a_host_register_points(float* xyz, long long n, float theta_deg, float tx, float ty, float tz)
{
    float R[16];
    s3a::register_rotation(theta_deg, R);
    for (long long k = 0; k < n; k++) s3a::register_point(R, tx, ty, tz, xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2]);
}

void s3a_host_roi_fill(const uint8_t* outline, int W, int H, uint8_t* roi, uint8_t* filled)
{
    for (int row = 0; row < H; row++) {   // k_roi_fill: first / last outline pixel of the row, then the closed form
        const uint8_t* o = outline + (size_t)row * W;
        int first = 0x7fffffff, last = -1;
        for (int c = 0; c < W; c++)
            if (o[c] != 0) {
                if (c < first) first = c;
                if (c > last) last = c;
            }
        for (int c = 0; c < W; c++) {
            const bool in = o[c] == 0 && c > first && c < last;
            roi[(size_t)row * W + c] = in ? 1 : 0;
            if (filled) filled[(size_t)row * W + c] = in ? 255 : o[c];
        }
    }
}
}
