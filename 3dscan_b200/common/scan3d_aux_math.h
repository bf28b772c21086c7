// scan3d_aux_math.h -- per-element arithmetic of the steps either side of the reconstruction path
// (SURVEY.md 8 f2 / f4), written once for the CUDA kernels (scan3d_aux_kernels.cu) and for a host
// build of the very same expressions that the CPU tests compare with the oracle
// (tests/test_aux_math_host.py compiles this header with g++ -ffp-contract=off).
//
//   cvUndistort2            2/project_pattern.cpp:220,234,372-427   -> s3a::undistort_map_row, s3a::bilinear_u8
//   register_point_clouds   9/register_point_clouds.cpp:93-137      -> s3a::register_rotation, s3a::register_point
//
// cvUndistort2 is OpenCV 2.4's cv::undistort: per stripe of max(1, 4096 / W) rows the camera matrix
// gets cy' = cy - y0 and is inverted (closed-form 3x3), initUndistortRectifyMap walks each row with
// running sums (_x += ir[0], ...) and stores a fixed-point map (5 fractional bits), and cv::remap
// blends 4 neighbours with 15-bit weights, constant (0) border.  Every double operation below is a
// separately rounded IEEE operation in OpenCV's order (the reference was an SSE2 build, no FMA).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define S3A_HD __host__ __device__ __forceinline__
#else
#define S3A_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define S3A_MUL(a, b) __dmul_rn((a), (b))
#define S3A_ADD(a, b) __dadd_rn((a), (b))
#define S3A_SUB(a, b) __dsub_rn((a), (b))
#define S3A_DIV(a, b) __ddiv_rn((a), (b))
#define S3A_FMUL(a, b) __fmul_rn((a), (b))
#define S3A_FADD(a, b) __fadd_rn((a), (b))
#define S3A_FSUB(a, b) __fsub_rn((a), (b))
#else  // host translation units are built with -ffp-contract=off
#define S3A_MUL(a, b) ((a) * (b))
#define S3A_ADD(a, b) ((a) + (b))
#define S3A_SUB(a, b) ((a) - (b))
#define S3A_DIV(a, b) ((a) / (b))
#define S3A_FMUL(a, b) ((float)(a) * (float)(b))
#define S3A_FADD(a, b) ((float)(a) + (float)(b))
#define S3A_FSUB(a, b) ((float)(a) - (float)(b))
#endif

namespace s3a {

// cv::invert of a 3x3 CV_64F matrix (DECOMP_LU uses the closed form); returns 0 when singular
// (OpenCV then zero-fills the result).
S3A_HD int invert3(const double* S, double* t)
{
#define S3A_DET2(a, b, c, d) S3A_SUB(S3A_MUL(S[a], S[b]), S3A_MUL(S[c], S[d]))
    double dd = S3A_ADD(S3A_SUB(S3A_MUL(S[0], S3A_DET2(4, 8, 5, 7)), S3A_MUL(S[1], S3A_DET2(3, 8, 5, 6))),
                        S3A_MUL(S[2], S3A_DET2(3, 7, 4, 6)));
    if (dd == 0.) {
        for (int k = 0; k < 9; k++) t[k] = 0.;
        return 0;
    }
    dd = S3A_DIV(1., dd);
    t[0] = S3A_MUL(S3A_DET2(4, 8, 5, 7), dd);
    t[1] = S3A_MUL(S3A_DET2(2, 7, 1, 8), dd);
    t[2] = S3A_MUL(S3A_DET2(1, 5, 2, 4), dd);
    t[3] = S3A_MUL(S3A_DET2(5, 6, 3, 8), dd);
    t[4] = S3A_MUL(S3A_DET2(0, 8, 2, 6), dd);
    t[5] = S3A_MUL(S3A_DET2(2, 3, 0, 5), dd);
    t[6] = S3A_MUL(S3A_DET2(3, 7, 4, 6), dd);
    t[7] = S3A_MUL(S3A_DET2(1, 6, 0, 7), dd);
    t[8] = S3A_MUL(S3A_DET2(0, 4, 1, 3), dd);
#undef S3A_DET2
    return 1;
}

// saturate_cast<int>(double) = cvRound: round half to even, x86 "integer indefinite" for NaN / overflow
S3A_HD int cv_round(double v)
{
    if (!(v == v) || v >= 2147483647.5 || v <= -2147483648.5) {
        if (v >= 2147483647.5) return 2147483647;  // saturate_cast clamps what it can represent
        return (int)0x80000000;
    }
#if defined(__CUDA_ARCH__)
    return __double2int_rn(v);
#else
    return (int)lrint(v);
#endif
}

// the stripe height cv::undistort works in
S3A_HD int undistort_stripe_rows(int W, int H)
{
    int s = (1 << 12) / (W > 1 ? W : 1);
    if (s < 1) s = 1;
    return s < H ? s : H;
}

// One row of cv::undistort's map.  K = camera matrix (row-major 3x3), d = (k1,k2,p1,p2,k3).
// m1 -> [W][2] int16 (source column, source row), m2 -> [W] uint16 (fy*32 + fx).
S3A_HD void undistort_map_row(const double* K, const double* d, int W, int H, int row, int16_t* m1, uint16_t* m2)
{
    const int stripe = undistort_stripe_rows(W, H);
    const int y0 = (row / stripe) * stripe, i = row - y0;
    const double fx = K[0], fy = K[4], u0 = K[2], v0 = K[5];
    const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
    double Ar[9], ir[9];
    for (int k = 0; k < 9; k++) Ar[k] = K[k];
    Ar[5] = S3A_SUB(v0, (double)y0);
    invert3(Ar, ir);
    double _x = S3A_ADD(S3A_MUL((double)i, ir[1]), ir[2]);
    double _y = S3A_ADD(S3A_MUL((double)i, ir[4]), ir[5]);
    double _w = S3A_ADD(S3A_MUL((double)i, ir[7]), ir[8]);
    for (int j = 0; j < W; j++) {
        const double w = S3A_DIV(1., _w), x = S3A_MUL(_x, w), y = S3A_MUL(_y, w);
        const double x2 = S3A_MUL(x, x), y2 = S3A_MUL(y, y);
        const double r2 = S3A_ADD(x2, y2), _2xy = S3A_MUL(S3A_MUL(2., x), y);
        // kr = (1 + ((k3 r2 + k2) r2 + k1) r2) / (1 + ((k6 r2 + k5) r2 + k4) r2) with k4..k6 = 0:
        // the denominator is exactly 1 and the division by it the identity
        const double kr = S3A_ADD(1., S3A_MUL(S3A_ADD(S3A_MUL(S3A_ADD(S3A_MUL(k3, r2), k2), r2), k1), r2));
        const double ux = S3A_ADD(S3A_ADD(S3A_MUL(x, kr), S3A_MUL(p1, _2xy)), S3A_MUL(p2, S3A_ADD(r2, S3A_MUL(2., x2))));
        const double vy = S3A_ADD(S3A_ADD(S3A_MUL(y, kr), S3A_MUL(p1, S3A_ADD(r2, S3A_MUL(2., y2)))), S3A_MUL(p2, _2xy));
        const double u = S3A_ADD(S3A_MUL(fx, ux), u0);
        const double v = S3A_ADD(S3A_MUL(fy, vy), v0);
        const int iu = cv_round(S3A_MUL(u, 32.)), iv = cv_round(S3A_MUL(v, 32.));
        m1[2 * j] = (int16_t)(iu >> 5);
        m1[2 * j + 1] = (int16_t)(iv >> 5);
        m2[j] = (uint16_t)((iv & 31) * 32 + (iu & 31));
        _x = S3A_ADD(_x, ir[0]);
        _y = S3A_ADD(_y, ir[3]);
        _w = S3A_ADD(_w, ir[6]);
    }
}

// cv::remap's fixed-point bilinear blend for 8-bit pixels: weights 32768 * {(1-fy)(1-fx), (1-fy)fx,
// fy(1-fx), fy fx} are exact integers, result = (sum + 2^14) >> 15.  Taps outside the image are 0
// (BORDER_CONSTANT), which also covers OpenCV's "completely outside" branch.
S3A_HD uint8_t bilinear_u8(int v0, int v1, int v2, int v3, int frac)
{
    const int ax = frac & 31, ay = (frac >> 5) & 31;
    const int s = v0 * (32 * (32 - ay) * (32 - ax)) + v1 * (32 * (32 - ay) * ax) + v2 * (32 * ay * (32 - ax)) +
                  v3 * (32 * ay * ax) + (1 << 14);
    return (uint8_t)(s >> 15);  // <= 255 because the weights sum to 2^15
}

// ---- tiled staging (k_remap_tiled) ------------------------------------------------------------------------
// A CTA owns an output tile of TILE_H x TILE_W pixels.  The source pixels its taps touch form a bounding box that
// is barely larger than the tile (the map is close to the identity); per frame that box is staged in shared
// memory with aligned 16-byte copies and the taps are gathered from there.  The box is frame-independent.
constexpr int REMAP_TILE_W = 256, REMAP_TILE_H = 8;      // output tile
// staging buffer: rows and row pitch in bytes.  The pitch is a multiple of 128 B, so the shared-memory bank of a
// staged byte does not depend on its row (with a 288-byte pitch a warp whose pixels straddle two source rows paid a
// two-way bank conflict on most loads: 40 % of the shared-memory wavefronts of the round-2 kernel)
constexpr int REMAP_BOX_W = 384, REMAP_BOX_H = 16;

struct RemapBox {
    int ok;         // 1: the box fits the buffer (it may reach outside the image: those bytes are staged as zeros,
                    //    which is cv::remap's constant border)
    int x0, y0;     // first staged source column (multiple of 16) and row
    int w, rows;    // staged bytes per row (multiple of 16, <= REMAP_BOX_W) and rows (<= REMAP_BOX_H)
};

// minx..maxx / miny..maxy: extremes of the map's integer source coordinates (sx, sy) over the tile's pixels
S3A_HD RemapBox remap_tile_box(int minx, int maxx, int miny, int maxy, int W, int H)
{
    RemapBox b;
    b.x0 = minx & ~15;
    b.y0 = miny;
    b.w = ((maxx + 1 - b.x0 + 1) + 15) & ~15;      // taps reach column maxx + 1
    b.rows = maxy + 1 - miny + 1;                   // ... and row maxy + 1
    b.ok = (W % 16 == 0) && b.w <= REMAP_BOX_W && b.rows <= REMAP_BOX_H && b.w > 0 && b.rows > 0;
    (void)H;
    return b;
}

// The part of staged row r that lies inside the image, as ONE aligned copy: *dst_off = offset in the staging buffer,
// *src_off = offset in a frame, returns the byte count (a multiple of 16; 0: the whole row is border).  x0 and W are
// multiples of 16, so the inside part starts and ends on 16-byte boundaries.  What it does not cover stays zero.
S3A_HD int remap_box_row_copy(const RemapBox& b, int r, int W, int H, int* dst_off, long long* src_off)
{
    const int y = b.y0 + r;
    const int xa = b.x0 < 0 ? 0 : b.x0, xb = b.x0 + b.w > W ? W : b.x0 + b.w;
    *dst_off = r * REMAP_BOX_W + (xa - b.x0);
    *src_off = (long long)y * W + xa;
    if (r >= b.rows || y < 0 || y >= H || xb <= xa) return 0;
    return xb - xa;
}

// offset of source pixel (sx, sy) inside the staging buffer (row pitch REMAP_BOX_W)
S3A_HD int remap_box_offset(const RemapBox& b, int sx, int sy) { return (sy - b.y0) * REMAP_BOX_W + (sx - b.x0); }

// The blend of one output pixel from the staged box.  Neighbouring output pixels (2j, 2j + 1) of a row nearly always
// read neighbouring source pixels of ONE row pair: both pixels' taps (x, x + 1) then lie inside the two aligned
// 32-bit words that hold the first pixel's first tap, and one byte permute with a frame-independent selector cuts
// all four out of them -- 2 shared-memory loads + 1 permute per pixel PAIR and row instead of 4 loads + 2 funnel
// shifts.  A pair that does not qualify (the second pixel on another source row, or more than 6 bytes further on)
// takes its second pixel from that pixel's own two words.
//   off0, off1: remap_box_offset of the two pixels.  *base = offset of the first word (multiple of 4),
//   *sel = permute selector giving [t0(x) t0(x+1) t1(x) t1(x+1)] out of {word(base), word(base + 4)};
//   returns true if the pair qualifies (else *sel's upper half selects nothing useful and must not be used).
S3A_HD bool remap_pair_window(int off0, int off1, int* base, uint32_t* sel)
{
    *base = off0 & ~3;
    const int o0 = off0 & 3, o1 = off1 - *base;
    const bool regular = o1 >= 0 && o1 <= 6;
    const uint32_t p1 = regular ? (uint32_t)o1 : 0u;
    *sel = (uint32_t)o0 | ((uint32_t)(o0 + 1) << 4) | (p1 << 8) | ((p1 + 1) << 12);
    return regular;
}
// selector of a single pixel's taps (low two bytes) out of its own two words
S3A_HD uint32_t remap_own_selector(int off) { return (uint32_t)(off & 3) | ((uint32_t)((off & 3) + 1) << 4); }

// bytes of the word pair {lo, hi} picked by four 4-bit indices (PRMT, default mode: index 0..7)
S3A_HD uint32_t permute_bytes(uint32_t lo, uint32_t hi, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(lo, hi, sel);
#else
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * k)) & 7))) & 0xffu) << (8 * k);
    return r;
#endif
}

// The four bilinear weights of cv::remap's table (bilinear_u8: products of two 5-bit fractions times 32, summing
// to 2^15) DOUBLED and packed as two pairs of unsigned 16-bit integers (wA = row y: w00 | w01 << 16, wB = row y + 1):
// with a sum of 2^16 the rounded result (S + 2^14) >> 15 = (2 S + 2^15) >> 16 is byte 2 of the accumulator, which a
// byte permute packs without shifts.  The only weight that does not fit, 2 * 32768 (fraction 0), is stored as 65535:
// 65535 t + 2^15 = 65536 t + (2^15 - t), and 0 < 2^15 - t < 2^16 for a byte t, so byte 2 is still exactly t.
S3A_HD void bilinear_weight_pairs_x2(int frac, uint32_t* wA, uint32_t* wB)
{
    const uint32_t ax = frac & 31, ay = (frac >> 5) & 31;
    uint32_t w00 = 64 * (32 - ay) * (32 - ax);
    if (w00 > 65535u) w00 = 65535u;
    *wA = w00 | ((64 * (32 - ay) * ax) << 16);
    *wB = (64 * ay * (32 - ax)) | ((64 * ay * ax) << 16);
}
// acc + w.lo * taps.byte0 + w.hi * taps.byte1  (IDP.2A.LO)  /  ... taps.byte2, taps.byte3  (IDP.2A.HI)
S3A_HD uint32_t dot2_u16_u8(uint32_t w, uint32_t taps, uint32_t acc)
{
#if defined(__CUDA_ARCH__)
    return __dp2a_lo(w, taps, acc);
#else
    return acc + (w & 0xffffu) * (taps & 0xffu) + (w >> 16) * ((taps >> 8) & 0xffu);
#endif
}
S3A_HD uint32_t dot2_u16_u8_hi(uint32_t w, uint32_t taps, uint32_t acc)
{
#if defined(__CUDA_ARCH__)
    return __dp2a_hi(w, taps, acc);
#else
    return acc + (w & 0xffffu) * ((taps >> 16) & 0xffu) + (w >> 16) * ((taps >> 24) & 0xffu);
#endif
}
// accumulators of the pixel in the low (hi = false) / high (hi = true) half of the tap words; the result is byte 2
S3A_HD uint32_t blend_acc_x2(uint32_t wA, uint32_t wB, uint32_t taps_row0, uint32_t taps_row1, bool hi)
{
    return hi ? dot2_u16_u8_hi(wB, taps_row1, dot2_u16_u8_hi(wA, taps_row0, 1u << 15))
              : dot2_u16_u8(wB, taps_row1, dot2_u16_u8(wA, taps_row0, 1u << 15));
}
// byte 2 of four accumulators -> one word of four output pixels
S3A_HD uint32_t pack_acc_bytes(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3)
{
    return permute_bytes(permute_bytes(a0, a1, 0x0062), permute_bytes(a2, a3, 0x0062), 0x5410);
}

// register_point_clouds' rotation about Y for a cloud captured at turntable angle theta (degrees,
// float): R(0,0) = R(2,2) = (float)cos(theta*Pi/180.0), R(0,2) = (float)(-1.0f*sin(...)),
// R(2,0) = (float)sin(...), with the reference's Pi = 22.0/7.0 (global_cv.h:62, expanded textually:
// theta*22.0/7.0/180.0).  Host only (libm, as in the reference).
inline void register_rotation(float theta_deg, float R[16])
{
    for (int k = 0; k < 16; k++) R[k] = 0.0f;  // cvCreateMat leaves them undefined; policy: 0
    const double a = theta_deg * 22.0 / 7.0 / 180.0;
    R[0] = (float)cos(a);
    R[2] = (float)(-1.0f * sin(a));
    R[8] = (float)sin(a);
    R[10] = (float)cos(a);
    R[5] = 1.0f;
    R[15] = 1.0f;
}

// One point of 9/register_point_clouds.cpp:117-137: float subtraction of the pivot, cvMatMul of the
// 4x4 float R with the 4x1 float point (OpenCV's unrolled len == 4 path: float products summed left
// to right), float addition of the pivot.
S3A_HD void register_point(const float* R, float tx, float ty, float tz, float& x, float& y, float& z)
{
    const float p0 = S3A_FSUB(x, tx), p1 = S3A_FSUB(y, ty), p2 = S3A_FSUB(z, tz), p3 = 1.0f;
    float q[3];
    for (int i = 0; i < 3; i++) {
        float t = S3A_FMUL(R[4 * i], p0);
        t = S3A_FADD(t, S3A_FMUL(R[4 * i + 1], p1));
        t = S3A_FADD(t, S3A_FMUL(R[4 * i + 2], p2));
        t = S3A_FADD(t, S3A_FMUL(R[4 * i + 3], p3));
        q[i] = t;
    }
    x = S3A_FADD(q[0], tx);
    y = S3A_FADD(q[1], ty);
    z = S3A_FADD(q[2], tz);
}

}  // namespace s3a
