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

// k_remap_tiled: per output tile the source box from the map's extremes, the box staged in a buffer with the
// kernel's pitch by one copy per row (what the copies do not cover is the zero border), the taps of every pixel PAIR
// cut out of two aligned words per row with the pair's selector (a pair that does not qualify: the second pixel from
// its own words), doubled weights, byte 2 of the accumulators packed.  Tiles whose box does not qualify take the
// per-tap path.  Returns the number of staged tiles; *irregular_pairs counts the pairs that took the fix-up.
long long s3a_host_remap_tiled(const uint8_t* src, int W, int H, const int16_t* xy, const uint16_t* frac, uint8_t* dst,
                               long long* irregular_pairs)
{
    using namespace s3a;
    alignas(16) static uint8_t box[REMAP_BOX_H * REMAP_BOX_W + 128];
    long long staged = 0, irr = 0;
    auto word = [&](int off) { uint32_t v; memcpy(&v, box + off, 4); return v; };
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
                const bool border = b.x0 < 0 || b.y0 < 0 || b.x0 + b.w > W || b.y0 + b.rows > H;
                memset(box, border ? 0 : 0xAA, sizeof(box));      // inside the image nothing unstaged may ever be used
                for (int r = 0; r < REMAP_BOX_H; r++) {
                    int d_off;
                    long long s_off;
                    const int n = remap_box_row_copy(b, r, W, H, &d_off, &s_off);
                    if (n) memcpy(box + d_off, src + s_off, (size_t)n);
                }
            }
            for (int y = ty0; y < ty0 + REMAP_TILE_H && y < H; y++)
                for (int x = tx0; x < tx0 + REMAP_TILE_W && x < W; x += 4) {     // W % 4 == 0 on this path
                    const size_t p = (size_t)y * W + x;
                    if (!b.ok) {
                        for (int k = 0; k < 4; k++) {
                            const int sx = xy[2 * (p + k)], sy = xy[2 * (p + k) + 1];
                            dst[p + k] = bilinear_u8(tap(src, W, H, sx, sy), tap(src, W, H, sx + 1, sy), tap(src, W, H, sx, sy + 1),
                                                     tap(src, W, H, sx + 1, sy + 1), frac[p + k]);
                        }
                        continue;
                    }
                    uint32_t acc[4];
                    for (int j = 0; j < 2; j++) {
                        const size_t q = p + 2 * j;
                        const int off0 = remap_box_offset(b, xy[2 * q], xy[2 * q + 1]);
                        const int off1 = remap_box_offset(b, xy[2 * (q + 1)], xy[2 * (q + 1) + 1]);
                        int base;
                        uint32_t sel, wA0, wB0, wA1, wB1;
                        const bool regular = remap_pair_window(off0, off1, &base, &sel);
                        bilinear_weight_pairs_x2(frac[q], &wA0, &wB0);
                        bilinear_weight_pairs_x2(frac[q + 1], &wA1, &wB1);
                        const uint32_t top = permute_bytes(word(base), word(base + 4), sel);
                        const uint32_t bot = permute_bytes(word(base + REMAP_BOX_W), word(base + REMAP_BOX_W + 4), sel);
                        acc[2 * j] = blend_acc_x2(wA0, wB0, top, bot, false);
                        acc[2 * j + 1] = blend_acc_x2(wA1, wB1, top, bot, true);
                        if (!regular) {
                            irr++;
                            const int own = off1 & ~3;
                            const uint32_t osel = remap_own_selector(off1);
                            acc[2 * j + 1] = blend_acc_x2(wA1, wB1, permute_bytes(word(own), word(own + 4), osel),
                                                          permute_bytes(word(own + REMAP_BOX_W), word(own + REMAP_BOX_W + 4), osel), false);
                        }
                    }
                    const uint32_t packed = pack_acc_bytes(acc[0], acc[1], acc[2], acc[3]);
                    memcpy(dst + p, &packed, 4);
                }
        }
    if (irregular_pairs) *irregular_pairs = irr;
    return staged;
}

// the doubled, saturated weight pairs give bilinear_u8's byte for every fraction and the given taps
int s3a_host_blend_x2_equals_reference(int v0, int v1, int v2, int v3)
{
    using namespace s3a;
    for (int frac = 0; frac < 1024; frac++) {
        uint32_t wA, wB;
        bilinear_weight_pairs_x2(frac, &wA, &wB);
        const uint32_t top = (uint32_t)v0 | ((uint32_t)v1 << 8) | ((uint32_t)v0 << 16) | ((uint32_t)v1 << 24);
        const uint32_t bot = (uint32_t)v2 | ((uint32_t)v3 << 8) | ((uint32_t)v2 << 16) | ((uint32_t)v3 << 24);
        const uint32_t lo = blend_acc_x2(wA, wB, top, bot, false), hi = blend_acc_x2(wA, wB, top, bot, true);
        const uint8_t want = bilinear_u8(v0, v1, v2, v3, frac);
        if (((lo >> 16) & 0xffu) != want || (lo >> 24) != 0 || hi != lo) return frac + 1;
    }
    return 0;
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
