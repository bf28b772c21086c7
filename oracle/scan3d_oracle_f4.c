/*
 * scan3d_oracle_f4.c -- CPU ORACLE (test infrastructure, NOT product code; see scan3d_oracle.h)
 * for the steps either side of the hot path:
 *
 *   capture-side undistortion   2/project_pattern.cpp:220,234,370-427  cvUndistort2(cap, undist_cap, K, d)
 *   ROI producer                M_tech_project_console/m_tech_project_console.cpp:186-229 (image_scissor's fill)
 *   turntable registration      9/register_point_clouds.cpp:83-148
 *
 * cvUndistort2 lives in OpenCV 2.4 (modules/imgproc/src/undistort.cpp: cv::undistort ->
 * initUndistortRectifyMap(CV_16SC2) per stripe -> remap(INTER_LINEAR, BORDER_CONSTANT), and
 * modules/imgproc/src/imgwarp.cpp: remapBilinear with the fixed-point BilinearTab_i), which is not
 * under /root/reference; its published algorithm is restated here.  No undistorted/original image
 * pair survives in the reference tree ("Original" folders are empty), so this piece is
 * "parity unpinned" by the reference and pinned against cv2 4.13's cv2.undistort instead
 * (tools/make_golden.py -> tests/golden/f4_kat.npz).  The registration's float cvMatMul (4x4 * 4x1,
 * OpenCV's unrolled len==4 path: float products summed left to right) is pinned against cv2.gemm the
 * same way.  The ROI fill is plain reference code and pinned by the stored i1.jpg (every row of the
 * filled image is one run).
 */
#include "scan3d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define Pi 22.0/7.0 /* PROJECT_GLOBAL/global_cv.h:62, textual */

/* cv::invert, 3x3 CV_64F, DECOMP_LU: closed form (modules/core/src/lapack.cpp). */
static int invert3(const double *S, double *t)
{
#define Sd(y, x) S[(y) * 3 + (x)]
    double d = Sd(0, 0) * (Sd(1, 1) * Sd(2, 2) - Sd(1, 2) * Sd(2, 1)) -
               Sd(0, 1) * (Sd(1, 0) * Sd(2, 2) - Sd(1, 2) * Sd(2, 0)) +
               Sd(0, 2) * (Sd(1, 0) * Sd(2, 1) - Sd(1, 1) * Sd(2, 0));
    if (d == 0.) {
        memset(t, 0, 9 * sizeof(double));
        return 0;
    }
    d = 1. / d;
    t[0] = (Sd(1, 1) * Sd(2, 2) - Sd(1, 2) * Sd(2, 1)) * d;
    t[1] = (Sd(0, 2) * Sd(2, 1) - Sd(0, 1) * Sd(2, 2)) * d;
    t[2] = (Sd(0, 1) * Sd(1, 2) - Sd(0, 2) * Sd(1, 1)) * d;
    t[3] = (Sd(1, 2) * Sd(2, 0) - Sd(1, 0) * Sd(2, 2)) * d;
    t[4] = (Sd(0, 0) * Sd(2, 2) - Sd(0, 2) * Sd(2, 0)) * d;
    t[5] = (Sd(0, 2) * Sd(1, 0) - Sd(0, 0) * Sd(1, 2)) * d;
    t[6] = (Sd(1, 0) * Sd(2, 1) - Sd(1, 1) * Sd(2, 0)) * d;
    t[7] = (Sd(0, 1) * Sd(2, 0) - Sd(0, 0) * Sd(2, 1)) * d;
    t[8] = (Sd(0, 0) * Sd(1, 1) - Sd(0, 1) * Sd(1, 0)) * d;
#undef Sd
    return 1;
}

static int cv_round(double v) /* saturate_cast<int>(double) = cvRound: half-to-even, saturating */
{
    if (!(v == v)) return (int)0x80000000;
    if (v >= 2147483647.5) return 2147483647;
    if (v <= -2147483648.5) return (int)0x80000000;
    return (int)lrint(v);
}

/* cv::undistort's map for the whole frame: for every stripe of stripe_size0 rows the new camera
 * matrix gets cy' = cy - y0, is inverted, and initUndistortRectifyMap walks each row with the
 * running sums _x += ir[0], _y += ir[3], _w += ir[6].  map_xy = [H][W][2] int16 (integer source
 * column,row), map_frac = [H][W] uint16 = fy5*32 + fx5 (INTER_BITS = 5). */
void o3d_undistort_map(const double K[9], const double d[5], int W, int H, int16_t *map_xy,
                       uint16_t *map_frac)
{
    int stripe0 = (1 << 12) / (W > 1 ? W : 1);
    if (stripe0 < 1) stripe0 = 1;
    if (stripe0 > H) stripe0 = H;
    const double fx = K[0], fy = K[4], u0 = K[2], v0 = K[5];
    const double k1 = d[0], k2 = d[1], p1 = d[2], p2 = d[3], k3 = d[4];
    const double k4 = 0., k5 = 0., k6 = 0.;
    for (int y0 = 0; y0 < H; y0 += stripe0) {
        const int rows = (H - y0) < stripe0 ? (H - y0) : stripe0;
        double Ar[9], ir[9];
        memcpy(Ar, K, sizeof(Ar));
        Ar[5] = v0 - y0;
        /* (Ar.colRange(0,3) * I).inv(DECOMP_LU): the product with the identity is exact */
        invert3(Ar, ir);
        for (int i = 0; i < rows; i++) {
            double _x = i * ir[1] + ir[2], _y = i * ir[4] + ir[5], _w = i * ir[7] + ir[8];
            int16_t *m1 = map_xy + ((size_t)(y0 + i) * W) * 2;
            uint16_t *m2 = map_frac + (size_t)(y0 + i) * W;
            for (int j = 0; j < W; j++, _x += ir[0], _y += ir[3], _w += ir[6]) {
                const double w = 1. / _w, x = _x * w, y = _y * w;
                const double x2 = x * x, y2 = y * y;
                const double r2 = x2 + y2, _2xy = 2 * x * y;
                const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
                const double u = fx * (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2)) + u0;
                const double v = fy * (y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy) + v0;
                const int iu = cv_round(u * 32), iv = cv_round(v * 32);
                m1[j * 2] = (int16_t)(iu >> 5);
                m1[j * 2 + 1] = (int16_t)(iv >> 5);
                m2[j] = (uint16_t)((iv & 31) * 32 + (iu & 31));
            }
        }
    }
}

/* cv::remap, 8UC1, INTER_LINEAR, BORDER_CONSTANT (value 0), fixed-point maps: weights are
 * BilinearTab_i = 32768 * {(1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx} (exact integers, their sum is
 * 32768 without the table's correction step), result = (sum + (1 << 14)) >> 15. */
void o3d_remap_bilinear(const uint8_t *src, int W, int H, const int16_t *map_xy,
                        const uint16_t *map_frac, uint8_t *dst)
{
    for (int r = 0; r < H; r++)
        for (int c = 0; c < W; c++) {
            const size_t p = (size_t)r * W + c;
            const int sx = map_xy[2 * p], sy = map_xy[2 * p + 1];
            const int f = map_frac[p] & 1023, ax = f & 31, ay = f >> 5;
            const int w0 = 32 * (32 - ay) * (32 - ax), w1 = 32 * (32 - ay) * ax;
            const int w2 = 32 * ay * (32 - ax), w3 = 32 * ay * ax;
            int v0, v1, v2, v3;
            if ((unsigned)sx < (unsigned)(W - 1 > 0 ? W - 1 : 0) && (unsigned)sy < (unsigned)(H - 1 > 0 ? H - 1 : 0)) {
                const uint8_t *S = src + (size_t)sy * W + sx;
                v0 = S[0]; v1 = S[1]; v2 = S[W]; v3 = S[W + 1];
            } else if (sx >= W || sx + 1 < 0 || sy >= H || sy + 1 < 0) {
                dst[p] = 0;
                continue;
            } else {
                const int x0ok = (unsigned)sx < (unsigned)W, x1ok = (unsigned)(sx + 1) < (unsigned)W;
                const int y0ok = (unsigned)sy < (unsigned)H, y1ok = (unsigned)(sy + 1) < (unsigned)H;
                v0 = (x0ok && y0ok) ? src[(size_t)sy * W + sx] : 0;
                v1 = (x1ok && y0ok) ? src[(size_t)sy * W + sx + 1] : 0;
                v2 = (x0ok && y1ok) ? src[(size_t)(sy + 1) * W + sx] : 0;
                v3 = (x1ok && y1ok) ? src[(size_t)(sy + 1) * W + sx + 1] : 0;
            }
            const int s = (v0 * w0 + v1 * w1 + v2 * w2 + v3 * w3 + (1 << 14)) >> 15;
            dst[p] = (uint8_t)(s < 0 ? 0 : s > 255 ? 255 : s);
        }
}

/* cvUndistort2(src, dst, K, d) on n_frames 8-bit single-channel images [n][H][W]
 * (2/project_pattern.cpp:220,234,380,393,416,427: every captured frame goes through it). */
void o3d_undistort_frames(const uint8_t *src, int n_frames, int W, int H, const double K[9],
                          const double d[5], uint8_t *dst)
{
    int16_t *mxy = (int16_t *)malloc((size_t)W * H * 2 * sizeof(int16_t));
    uint16_t *mf = (uint16_t *)malloc((size_t)W * H * sizeof(uint16_t));
    o3d_undistort_map(K, d, W, H, mxy, mf);
    for (int f = 0; f < n_frames; f++)
        o3d_remap_bilinear(src + (size_t)f * W * H, W, H, mxy, mf, dst + (size_t)f * W * H);
    free(mxy);
    free(mf);
}

/* image_scissor's fill (m_tech_project_console.cpp:186-229), literally: per row, from a non-zero
 * outline pixel p1 search the next non-zero pixel i, set everything strictly between to selected
 * (and to 255 in the outline image), then restart the search AT i.  outline is modified in place
 * like the reference's internal_image (what it saves as i1.jpg); roi = selected_region as u8. */
void o3d_roi_fill(uint8_t *outline, int W, int H, uint8_t *roi)
{
    memset(roi, 0, (size_t)W * H); /* :186-191 */
    for (int j = 0; j < H; j++) {
        int p1 = -1;
        for (int i = 0; i < W; i++) {
            if (outline[(size_t)j * W + i] != 0 && p1 == -1) { /* :200 start point */
                p1 = i;
                for (i = p1 + 1; i < W; i++) {                 /* :206 */
                    if (outline[(size_t)j * W + i] != 0) {
                        for (int h = p1 + 1; h < i; h++) {     /* :211-215 */
                            outline[(size_t)j * W + h] = (unsigned char)255;
                            roi[(size_t)j * W + h] = 1;
                        }
                        break;
                    }
                }
                p1 = -1;                                       /* :222 */
                i--;                                           /* :223 */
            }
        }
    }
}

/* register_point_clouds' per-cloud transform (9/register_point_clouds.cpp:92-127).  R is a 4x4
 * CV_32F matrix of which only (0,0),(0,2),(1,1),(2,0),(2,2),(3,3) are ever written (:34-35,
 * :93-100); cvCreateMat does not clear, policy here: the rest is 0.  theta is a float in degrees,
 * theta*Pi/180.0 and cos/sin are double, the store rounds to float.  Per point (:117-137): float
 * subtract of the pivot, cvMatMul(R, point, point) (float 4x4 * 4x1 through a temporary: OpenCV's
 * unrolled len == 4 path, float t = a0*b0 + a1*b1 + a2*b2 + a3*b3 left to right, then
 * (float)(t * 1.0)), float add of the pivot. */
void o3d_register_rotation(float theta_deg, float R[16])
{
    memset(R, 0, 16 * sizeof(float));
    R[0 * 4 + 0] = cos(theta_deg * Pi / 180.0);
    R[0 * 4 + 2] = -1.0f * sin(theta_deg * Pi / 180.0);
    R[2 * 4 + 0] = sin(theta_deg * Pi / 180.0);
    R[2 * 4 + 2] = cos(theta_deg * Pi / 180.0);
    R[1 * 4 + 1] = 1.0f;
    R[3 * 4 + 3] = 1.0f;
}

void o3d_register_points(float *xyz, int64_t n, float theta_deg, float tx, float ty, float tz)
{
    float R[16];
    o3d_register_rotation(theta_deg, R);
    for (int64_t k = 0; k < n; k++) {
        float p[4], q[4];
        p[0] = xyz[3 * k]; p[1] = xyz[3 * k + 1]; p[2] = xyz[3 * k + 2]; p[3] = 1.0f;
        p[0] -= tx; p[1] -= ty; p[2] -= tz;
        for (int i = 0; i < 4; i++) {
            float t = R[i * 4] * p[0];
            t = t + R[i * 4 + 1] * p[1];
            t = t + R[i * 4 + 2] * p[2];
            t = t + R[i * 4 + 3] * p[3];
            q[i] = (float)(t * 1.0);
        }
        q[0] += tx; q[1] += ty; q[2] += tz;
        xyz[3 * k] = q[0]; xyz[3 * k + 1] = q[1]; xyz[3 * k + 2] = q[2];
    }
}
