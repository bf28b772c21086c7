/*
 * scan3d_oracle.c -- CPU ORACLE (test infrastructure, NOT product code; see the header).
 *
 * Plain-C restatement of the reference's per-pixel reconstruction loops.  Every function
 * cites the reference file:line it follows (paths relative to /root/reference).
 * Compile with -ffp-contract=off: the reference was an SSE2 build (no FMA), and the GPU
 * path mirrors the same IEEE operation order with __dmul_rn/__dadd_rn.
 */
#include "scan3d_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* PROJECT_GLOBAL/global_cv.h:62 -- textual, unparenthesised macro.  Every use below keeps
 * the reference's token sequence so that the expansion associates identically. */
#define Pi 22.0/7.0

#define PAR_FOR _Pragma("omp parallel for schedule(static) num_threads(nth) if (nth > 1)")

/* Layout of the intermediate planes (valid maps, wrapped / unwrapped phases, fringe orders, intersection_points):
 * 0 = row-major [row][col] (the default, what every test and the product compare in); 1 = the reference's own
 * arrays, int (*)[H] etc. indexed [col][row] from row-outer loops (common_variables.h:12-21, e.g.
 * 3/wrapped_phase.cpp:164-183): every plane access then strides by H elements.  Same values either way; the switch
 * exists for the reference-faithful single-thread timing of BASELINE.md section 3 (o3d_reconstruct_colrow).
 * Input images, c_p_map ([W*H][2], row-major in the reference too, 5/...:648) and the look-up tables keep their
 * layout.  Not thread-safe across concurrent calls with different layouts (the timing leg is single-threaded). */
static int g_colrow = 0;
#define PL(row, col) (g_colrow ? (long)(col) * H + (row) : (long)(row) * W + (col))

static int clamp_threads(int threads)
{
#ifdef _OPENMP
    /* an explicit request may exceed OMP_NUM_THREADS (launchers such as torchrun export
     * OMP_NUM_THREADS=1), never the number of processors */
    if (threads <= 0) return omp_get_max_threads();
    int np = omp_get_num_procs();
    return threads > np ? np : threads;
#else
    (void)threads;
    return 1;
#endif
}

int o3d_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ stage 3 */

/* 3/wrapped_phase.cpp:78-82 (clear) + :106-115 (valid where selected_region==1).
 * The modulation test :84-104 is commented out in the reference and is not applied. */
void o3d_check_roi(const uint8_t *roi, int W, int H, int32_t *valid)
{
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) valid[PL(i, j)] = roi[(long)i * W + j] != 0 ? 1 : 0;
}

/* The same function exactly as committed (3/wrapped_phase.cpp:78-127): the map is cleared, then filled from
 * `selected_region[j][i] == 1` ONLY when number_of_patterns_fringe is 3 or 4 (:106); the 5-step block (:117-127)
 * is commented out, so a 5-step run of the reference ends with an all-zero valid map and an empty cloud.
 * o3d_check_roi above (any N, roi != 0) is the deliberate extension the product's default follows; this literal
 * form backs SCAN3D_FLAG_STRICT_REFERENCE. */
void o3d_check_roi_strict(const uint8_t *roi, int N, int W, int H, int32_t *valid)
{
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) valid[PL(i, j)] = 0;
    if (N == 4 || N == 3)
        for (int i = 0; i < H; i++)
            for (int j = 0; j < W; j++)
                if (roi[(long)i * W + j] == 1) valid[PL(i, j)] = 1;
}

/* Extension weights for N not in {3,4,5,8}: shifts delta_k = 2*pi*k/N (true pi), libm. */
/* check_I_mod_criteria, the branch the reference keeps commented out (3/wrapped_phase.cpp:84-104,
 * 3-step only): gamma = sqrtf(3 (I0-I2)^2 + (2 I1 - I0 - I2)^2) / (float)(I0+I1+I2), the pixel is
 * kept when gamma > 0.01 and it lies in the selected region.  Arithmetic as written there: the
 * differences are ints, the squares and the sum doubles (exact), sqrtf takes the sum as a float
 * (exact below 2^24), t1/t2 is a float division, the comparison promotes to double.  A black pixel
 * gives 0/0 = NaN, which fails the comparison. */
void o3d_check_I_mod_criteria(const uint8_t *fringe, const uint8_t *roi, int W, int H, int32_t *valid)
{
    const size_t n = (size_t)W * H;
    for (int row = 0; row < H; row++)
        for (int col = 0; col < W; col++) {
            const size_t i = (size_t)row * W + col;
            const int i0 = fringe[i], i1 = fringe[n + i], i2 = fringe[2 * n + i];
            const double d = i0 - i2;
            const double e = 2.0 * i1 - i0 - i2;
            const float t1 = sqrtf((float)(3.0 * (d * d) + e * e));
            const float t2 = (float)(i0 + i1 + i2);
            const float t3 = t1 / t2;
            valid[PL(row, col)] = (t3 > 0.01 && roi[i] != 0) ? 1 : 0;
        }
}

static void nstep_weights(int N, double *s, double *c)
{
    for (int k = 0; k < N; k++) {
        double a = 2.0 * 3.14159265358979323846 * (double)k / (double)N;
        s[k] = sin(a);
        c[k] = cos(a);
    }
}

/* 3/wrapped_phase.cpp:151-238 */
void o3d_wrapped_phase(const uint8_t *fringe, int N, int W, int H, const int32_t *valid,
                       float *wrapped, uint8_t *dbg, int threads)
{
    const int nth = clamp_threads(threads);
    const long plane = (long)W * H;
    double ws[64], wc[64];
    if (N != 3 && N != 4 && N != 5 && N != 8) nstep_weights(N, ws, wc);

    PAR_FOR
    for (int row = 0; row < H; row++) {
        for (int col = 0; col < W; col++) {
            const long p = (long)row * W + col;     /* in the input images */
            const long q = PL(row, col);            /* in the planes */
            if (valid[q] != 1) continue;
            float t1, t2, t3;
            if (N == 3) { /* :171-179 */
                t1 = (float)fringe[0 * plane + p] - (float)fringe[2 * plane + p];
                t2 = 2.0 * ((float)fringe[1 * plane + p]) - (float)fringe[0 * plane + p] -
                     (float)fringe[2 * plane + p];
                wrapped[q] = atan2(t1, t2); /* double atan2, stored as float */
                t3 = 128.0f + 127.0f * (wrapped[q] / (Pi));
            } else if (N == 4) { /* :195-201 */
                t1 = (float)fringe[3 * plane + p] - (float)fringe[1 * plane + p];
                t2 = (float)fringe[0 * plane + p] - (float)fringe[2 * plane + p];
                wrapped[q] = atan2(t1, t2);
                t3 = 127.0f + 128.0f * (wrapped[q] / (Pi));
            } else if (N == 5) { /* :217-222 (Hariharan), float atan2f */
                t1 = 2.0 * ((float)fringe[1 * plane + p] - (float)fringe[3 * plane + p]);
                t2 = 2.0 * (float)fringe[2 * plane + p] - (float)fringe[0 * plane + p] -
                     (float)fringe[4 * plane + p];
                wrapped[q] = atan2f(t1, t2);
                t3 = 127.0 + 128.0 * (wrapped[q] / (Pi));
            } else if (N == 8) {
                /* EXTENSION (no reference counterpart): 8-step, shifts k*pi/4, same phase
                 * origin as the reference's 4-step (phi = theta - pi).  Integer parts are
                 * exact; one double multiply by sqrt(1/2) and one double add each. */
                const int I0 = fringe[0 * plane + p], I1 = fringe[1 * plane + p],
                          I2 = fringe[2 * plane + p], I3 = fringe[3 * plane + p],
                          I4 = fringe[4 * plane + p], I5 = fringe[5 * plane + p],
                          I6 = fringe[6 * plane + p], I7 = fringe[7 * plane + p];
                const double r = 0.70710678118654752440;
                const double d1 = (double)(I6 - I2) + (double)(I5 + I7 - I1 - I3) * r;
                const double d2 = (double)(I0 - I4) + (double)(I1 + I7 - I3 - I5) * r;
                wrapped[q] = atan2(d1, d2);
                t3 = 127.0f + 128.0f * (wrapped[q] / (Pi));
            } else {
                /* EXTENSION: generic N-step, sequential double sums, k ascending. */
                double S = 0.0, C = 0.0;
                for (int k = 0; k < N; k++) {
                    const double I = (double)fringe[(long)k * plane + p];
                    S = S + I * ws[k];
                    C = C + I * wc[k];
                }
                wrapped[q] = atan2(0.0 - S, C); /* 0.0 - S: S == +0 must not become -0 */
                t3 = 127.0f + 128.0f * (wrapped[q] / (Pi));
            }
            if (dbg) dbg[p] = (unsigned char)(int)(t3);
        }
    }
}

/* ---- fdlibm atan2f restated (glibc sysdeps/ieee754/flt-32/e_atan2f.c + s_atanf.c) ----------
 * NOT used by o3d_wrapped_phase (which calls libm's atan2f like the reference); exported so the
 * tests can pin this operation sequence -- the one the CUDA path mirrors for the 5-step formula
 * -- against the libm of the machine the tests run on. */
static inline int32_t f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
static inline float i2f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
static float atanf_restated(float x)
{
    static const float atanhi[] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    static const float atanlo[] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    static const float aT[] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                               9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                               4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
    float w, s1, s2, z;
    int32_t ix, hx = f2i(x), id;
    ix = hx & 0x7fffffff;
    if (ix >= 0x4c000000) {
        if (ix > 0x7f800000) return x + x;
        return hx > 0 ? atanhi[3] + atanlo[3] : -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000) {
        if (ix < 0x31000000) return x;
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else { id = 3; x = -1.0f / x; }
        }
    }
    z = x * x;
    w = z * z;
    s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return hx < 0 ? -z : z;
}
float o3d_atan2f_restated(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    float z;
    int32_t k, m, hx = f2i(x), hy = f2i(y), ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return atanf_restated(y);
    m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) return m < 2 ? y : (m == 2 ? pi + tiny : -pi - tiny);
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    k = (iy - ix) >> 23;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = atanf_restated(fabsf(y / x));
    switch (m) {
        case 0: return z;
        case 1: return i2f(f2i(z) ^ (int32_t)0x80000000);
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}
/* count of (t1,t2) on the 5-step input grid where the restatement differs from libm atan2f */
long o3d_atan2f_restated_mismatches(void)
{
    long bad = 0;
    for (int a = -510; a <= 510; a += 2)
        for (int b = -510; b <= 510; b++) {
            const float r = atan2f((float)a, (float)b), q = o3d_atan2f_restated((float)a, (float)b);
            bad += f2i(r) != f2i(q);
        }
    return bad;
}

/* 3/wrapped_phase.cpp:266-279 (vertical) == :306-318 (horizontal): one raster pass,
 * "any pixel with an invalid, not-yet-visited neighbour becomes invalid and visited". */
void o3d_mask_recurrence(int32_t *valid, int W, int H, uint8_t *dbg)
{
    uint8_t *visited = (uint8_t *)calloc((size_t)W * H, 1);
#define V(u, y) valid[PL(y, u)]
#define S(u, y) visited[PL(y, u)]
    for (int y = 1; y < H - 1; y++)
        for (int u = 1; u < W - 1; u++) {
            if (((V(u - 1, y - 1) != 1) && !S(u - 1, y - 1)) ||
                ((V(u, y - 1) != 1) && !S(u, y - 1)) ||
                ((V(u + 1, y - 1) != 1) && !S(u + 1, y - 1)) ||
                ((V(u - 1, y) != 1) && !S(u - 1, y)) || ((V(u + 1, y) != 1) && !S(u + 1, y)) ||
                ((V(u - 1, y + 1) != 1) && !S(u - 1, y + 1)) ||
                ((V(u, y + 1) != 1) && !S(u, y + 1)) ||
                ((V(u + 1, y + 1) != 1) && !S(u + 1, y + 1))) {
                V(u, y) = 0;
                S(u, y) = 1;
                if (dbg) dbg[(long)y * W + u] = 0;
            }
        }
#undef V
#undef S
    free(visited);
}

/* Closed form of the recurrence above (derivation: SURVEY.md 8a row 3 / DESIGN.md). */
void o3d_mask_closed_form(const int32_t *valid0, int W, int H, int32_t *valid1)
{
#define INV(x, y) (valid0[(long)(y) * W + (x)] != 1)
#define BORDER(x, y) ((x) == 0 || (y) == 0 || (x) == W - 1 || (y) == H - 1)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const long p = (long)y * W + x;
            if (BORDER(x, y)) { valid1[p] = valid0[p]; continue; }
            static const int ex[4] = {-1, 0, 1, -1}, ey[4] = {-1, -1, -1, 0}; /* EARLY */
            static const int lx[4] = {1, -1, 0, 1}, ly[4] = {0, 1, 1, 1};     /* LATE  */
            int trig = 0;
            for (int k = 0; k < 4; k++) trig |= INV(x + lx[k], y + ly[k]);
            for (int k = 0; k < 4 && !trig; k++) {
                const int qx = x + ex[k], qy = y + ey[k];
                if (!INV(qx, qy)) continue;
                if (BORDER(qx, qy)) { trig = 1; break; }
                int E = 0;
                for (int j = 0; j < 4; j++) E |= INV(qx + lx[j], qy + ly[j]);
                for (int j = 0; j < 4; j++) {
                    const int rx = qx + ex[j], ry = qy + ey[j];
                    E |= INV(rx, ry) && BORDER(rx, ry);
                }
                if (!E) trig = 1;
            }
            valid1[p] = (valid0[p] == 1 && !trig) ? 1 : (trig ? 0 : valid0[p]);
        }
#undef INV
#undef BORDER
}

/* ------------------------------------------------------------------ stage 4 */

/* 4/phase_unwrap.cpp:134-275, Gray-coded branch (count==1, :162-202 / :229-266). */
void o3d_decode_gray(const uint8_t *gray, const uint8_t *inv, int M, int W, int H,
                     const int32_t *valid, int32_t *code, int threads)
{
    const int nth = clamp_threads(threads);
    const long plane = (long)W * H;
    PAR_FOR
    for (int row = 0; row < H; row++)
        for (int col = 0; col < W; col++) {
            const long p = (long)row * W + col, q = PL(row, col);
            code[q] = -1; /* :141-143 */
            if (valid[q] != 1) continue;
            int c = 0, Bprev = 0;
            for (int i = 0; i < M; i++) {
                /* :183 -- uchar - uchar is int arithmetic; tie (==0) decodes as 1 */
                const int G = ((int)gray[(long)i * plane + p] - (int)inv[(long)i * plane + p]) >= 0;
                const int B = (i == 0) ? G : (Bprev != G); /* :187-191 */
                c += B * (1 << (M - 1 - i));               /* :193, bit 0 = MSB */
                Bprev = B;
            }
            code[q] = c; /* no range reject (:196-200 is a no-op) */
        }
}

/* 4/phase_unwrap.cpp:278-316 */
void o3d_unwrap(int dir, float *wrapped, const int32_t *code, const int32_t *valid, int W,
                int H, float *unwrapped, int threads)
{
    const int nth = clamp_threads(threads);
    const int c0 = dir == 0 ? 1 : 0, c1 = dir == 0 ? W - 1 : W;
    const int r0 = dir == 0 ? 0 : 1, r1 = dir == 0 ? H : H - 1;
    PAR_FOR
    for (int row = r0; row < r1; row++)
        for (int col = c0; col < c1; col++) {
            const long p = PL(row, col);
            if (valid[p] != 1) continue;
            wrapped[p] += Pi;                              /* :290 / :308 */
            unwrapped[p] = wrapped[p] + code[p] * 2.0 * Pi; /* :291 / :309 */
        }
}

/* 4/phase_unwrap.cpp:330-336 / :349-355 */
void o3d_unwrapped_image(const float *unwrapped, const int32_t *valid, int W, int H,
                         int number_of_codes, uint8_t *img)
{
    float t;
    for (long p = 0; p < (long)W * H; p++)
        if (valid[p] == 1) {
            t = unwrapped[p] / (2.0 * Pi * number_of_codes);
            img[p] = (unsigned char)(int)(t * 255);
        }
}

/* ------------------------------------------------------------------ stage 5 */

/* 5/compute_correspondance.cpp:60-77 + :642-679.  FE_INVALID of lrint (NaN / inf / out of
 * long range) is restated as an explicit test so that the loop can run in parallel. */
static int lrint_checked(double v, int64_t *out)
{
    if (!(v == v) || v >= 9223372036854775808.0 || v < -9223372036854775808.0) return 0;
    *out = (int64_t)lrint(v); /* default rounding mode: half-to-even */
    return 1;
}

void o3d_compute_c_p_map(const float *unw_v, const float *unw_h, const int32_t *valid_v,
                         const int32_t *valid_h, int fw_v, int fw_h, int PW, int PH, int W,
                         int H, int64_t *cpmap, int32_t *valid, int threads)
{
    const int nth = clamp_threads(threads);
    PAR_FOR
    for (int r = 0; r < H; r++)
        for (int c = 0; c < W; c++) {
            const long p = (long)r * W + c;         /* c_p_map row */
            const long q = PL(r, c);                /* planes */
            valid[q] = (valid_v[q] == 1 && valid_h[q] == 1) ? 1 : 0; /* :60-77 */
            if (valid[q] != 1) continue;
            if (!lrint_checked(fw_v * (unw_v[q] / (2.0 * Pi)), &cpmap[2 * p + 0])) { /* :648 */
                valid[q] = 0;
                continue;
            }
            if (!lrint_checked(fw_h * (unw_h[q] / (2.0 * Pi)), &cpmap[2 * p + 1])) { /* :659 */
                valid[q] = 0;
                continue;
            }
            if (cpmap[2 * p] > (PW - 1) || cpmap[2 * p + 1] > (PH - 1) || cpmap[2 * p] < 0 ||
                cpmap[2 * p + 1] < 0) /* :671-675 */
                valid[q] = 0;
        }
}

/* ------------------------------------------------------------------ stage 6/7 */

/* cvRodrigues2, rotation vector -> matrix (OpenCV 2.4 modules/calib3d/src/calibration.cpp),
 * called at 7/triangulation.cpp:1072,1080 and 6/system_calibration.cpp:1489-1490. */
void o3d_rodrigues(const double rvec[3], double R[9])
{
    double rx = rvec[0], ry = rvec[1], rz = rvec[2];
    const double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if (theta < DBL_EPSILON) {
        for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double c = cos(theta), s = sin(theta), c1 = 1. - c;
    const double itheta = theta ? 1. / theta : 0.;
    rx *= itheta; ry *= itheta; rz *= itheta;
    const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    for (int k = 0; k < 9; k++) R[k] = c * I[k] + c1 * rrt[k] + s * r_x[k];
}

/* cvMatMul restated: D[m x n] = A[m x k] * B[k x n], sums k ascending from 0 (OpenCV
 * GEMMSingleMul / the unrolled small-matrix path give the same rounding sequence). */
static void matmul(const double *A, const double *B, double *D, int m, int k, int n)
{
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int t = 0; t < k; t++) s += A[i * k + t] * B[t * n + j];
            D[i * n + j] = s;
        }
}

/* 6/system_calibration.cpp:1489-1504 */
void o3d_compose_relative(const double rc[3], const double tc[3], const double rp[3],
                          const double tp[3], double R[9], double T[3])
{
    double Rc[9], Rp[9], RpT[9], RT[3];
    o3d_rodrigues(rc, Rc);
    o3d_rodrigues(rp, Rp);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) RpT[i * 3 + j] = Rp[j * 3 + i]; /* cvTranspose :1494 */
    matmul(Rc, RpT, R, 3, 3, 3);                                    /* :1495 */
    matmul(R, tp, RT, 3, 3, 1);                                     /* :1503 */
    for (int i = 0; i < 3; i++) T[i] = tc[i] - RT[i];               /* :1504 */
}

/* One point of cvUndistortPoints (OpenCV 2.4 modules/imgproc/src/undistort.cpp): no R, no P,
 * 5-coefficient model (k1,k2,p1,p2,k3) -> k[5..7] = 0 so the rational numerator is exactly 1;
 * 5 fixed iterations.  Returns NORMALISED coordinates. */
static inline void undistort_one(double u, double v, double fx, double fy, double ifx,
                                 double ify, double cx, double cy, const double k[5],
                                 double *ox, double *oy)
{
    double x, y, x0, y0;
    x0 = x = (u - cx) * ifx;
    y0 = y = (v - cy) * ify;
    (void)fx; (void)fy;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = 1. / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
        const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    *ox = x;
    *oy = y;
}

void o3d_undistort_points(const double *src_xy, int n, const double K[9], const double d[5],
                          double *dst_xy)
{
    const double fx = K[0], fy = K[4], ifx = 1. / fx, ify = 1. / fy, cx = K[2], cy = K[5];
    for (int i = 0; i < n; i++)
        undistort_one(src_xy[2 * i], src_xy[2 * i + 1], fx, fy, ifx, ify, cx, cy, d,
                      &dst_xy[2 * i], &dst_xy[2 * i + 1]);
}

/* 7/triangulation.cpp:262-307 (camera) / :269-276,352-378 (projector): undistort every pixel
 * centre, then cvMatMul(K, [xn;yn;1]) and divide by the third row.  Row index by integer
 * division (the reference's floorf((float)f/(float)W) is identical below 2^24 pixels). */
void o3d_undistort_lut(const double K[9], const double d[5], int W, int H, double *lut,
                       int threads)
{
    const int nth = clamp_threads(threads);
    const double fx = K[0], fy = K[4], ifx = 1. / fx, ify = 1. / fy, cx = K[2], cy = K[5];
    const long n = (long)W * H;
    PAR_FOR
    for (int row = 0; row < H; row++)
        for (int col = 0; col < W; col++) {
            double xn, yn;
            undistort_one((double)col, (double)row, fx, fy, ifx, ify, cx, cy, d, &xn, &yn);
            /* cvMatMul(K, P, P): s = 0; s += K[i][0]*xn; s += K[i][1]*yn; s += K[i][2]*1 */
            double m0 = 0.0, m1 = 0.0, m2 = 0.0;
            m0 += K[0] * xn; m0 += K[1] * yn; m0 += K[2] * 1.0;
            m1 += K[3] * xn; m1 += K[4] * yn; m1 += K[5] * 1.0;
            m2 += K[6] * xn; m2 += K[7] * yn; m2 += K[8] * 1.0;
            const long p = (long)row * W + col;
            lut[p] = m0 / m2;     /* :305-307: rows 0,1 are divided before row 2 divides itself */
            lut[n + p] = m1 / m2;
        }
}

/* 7/triangulation.cpp:1061-1126: A = K * [R|t] */
void o3d_compute_A(const double K[9], const double rvec[3], const double tvec[3], double A[12])
{
    double R[9], Rt[12];
    o3d_rodrigues(rvec, R);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) Rt[i * 4 + j] = R[i * 3 + j];
        Rt[i * 4 + 3] = tvec[i];
    }
    matmul(K, Rt, A, 3, 3, 4);
}

/* 7/triangulation.cpp:1134-1218: P (4x3), F (4x1), V = ((P^T P)^-1 P^T) F with cvInvert's
 * closed-form 3x3 (OpenCV 2.4 modules/core/src/lapack.cpp, n==3, CV_64F branch). */
void o3d_triangulate_point(const double Ac[12], const double Ap[12], double uc, double vc,
                           double up, double vp, double xyz[3])
{
    double P[12], F[4], Pt[12], S[9], Si[9], I2[12];
    for (int j = 0; j < 3; j++) {
        P[0 * 3 + j] = Ac[0 * 4 + j] - uc * Ac[2 * 4 + j]; /* :1152-1154 */
        P[1 * 3 + j] = Ac[1 * 4 + j] - vc * Ac[2 * 4 + j]; /* :1157-1159 */
        P[2 * 3 + j] = Ap[0 * 4 + j] - up * Ap[2 * 4 + j]; /* :1162-1164 */
        P[3 * 3 + j] = Ap[1 * 4 + j] - vp * Ap[2 * 4 + j]; /* :1166-1168 */
    }
    F[0] = Ac[11] * uc - Ac[3]; /* :1181 */
    F[1] = Ac[11] * vc - Ac[7]; /* :1182 */
    F[2] = Ap[11] * up - Ap[3]; /* :1187 */
    F[3] = Ap[11] * vp - Ap[7]; /* :1188 */
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 3; j++) Pt[j * 4 + i] = P[i * 3 + j]; /* cvTranspose :1202 */
    matmul(Pt, P, S, 3, 4, 3);                                    /* :1203 */
    {                                                             /* cvInvert :1204 */
#define Sd(y, x) S[(y) * 3 + (x)]
        double dd = Sd(0, 0) * (Sd(1, 1) * Sd(2, 2) - Sd(1, 2) * Sd(2, 1)) -
                    Sd(0, 1) * (Sd(1, 0) * Sd(2, 2) - Sd(1, 2) * Sd(2, 0)) +
                    Sd(0, 2) * (Sd(1, 0) * Sd(2, 1) - Sd(1, 1) * Sd(2, 0));
        if (dd != 0.) {
            dd = 1. / dd;
            Si[0] = (Sd(1, 1) * Sd(2, 2) - Sd(1, 2) * Sd(2, 1)) * dd;
            Si[1] = (Sd(0, 2) * Sd(2, 1) - Sd(0, 1) * Sd(2, 2)) * dd;
            Si[2] = (Sd(0, 1) * Sd(1, 2) - Sd(0, 2) * Sd(1, 1)) * dd;
            Si[3] = (Sd(1, 2) * Sd(2, 0) - Sd(1, 0) * Sd(2, 2)) * dd;
            Si[4] = (Sd(0, 0) * Sd(2, 2) - Sd(0, 2) * Sd(2, 0)) * dd;
            Si[5] = (Sd(0, 2) * Sd(1, 0) - Sd(0, 0) * Sd(1, 2)) * dd;
            Si[6] = (Sd(1, 0) * Sd(2, 1) - Sd(1, 1) * Sd(2, 0)) * dd;
            Si[7] = (Sd(0, 1) * Sd(2, 0) - Sd(0, 0) * Sd(2, 1)) * dd;
            Si[8] = (Sd(0, 0) * Sd(1, 1) - Sd(0, 1) * Sd(1, 0)) * dd;
        } else {
            for (int k = 0; k < 9; k++) Si[k] = 0.; /* cvInvert zero-fills a singular dst */
        }
#undef Sd
    }
    matmul(Si, Pt, I2, 3, 3, 4); /* :1205 */
    matmul(I2, F, xyz, 3, 4, 1); /* :1206 */
}

/* 7/triangulation.cpp:1223-1247 */
void o3d_triangulate(const double A_cam[12], const double A_proj[12], const double *cam_lut,
                     const double *proj_lut, const int64_t *cpmap, const int32_t *valid, int W,
                     int H, int PW, int PH, double *xyz, int threads)
{
    const int nth = clamp_threads(threads);
    const long n = (long)W * H, np = (long)PW * PH;
    PAR_FOR
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) {
            const long p = (long)i * W + j, q = PL(i, j);
            if (valid[q] != 1) continue;
            const int cx = (int)cpmap[2 * p], cy = (int)cpmap[2 * p + 1]; /* :1146-1147 */
            const long ql = (long)cy * PW + cx;
            o3d_triangulate_point(A_cam, A_proj, cam_lut[p], cam_lut[n + p], proj_lut[ql],
                                  proj_lut[np + ql], &xyz[3 * q]);
        }
}

/* ------------------------------------------------------------------ stage 8 */

/* 8/save_point_cloud.cpp:33-39 (count) + :85-136 (raster-order gather, (float) casts). */
int64_t o3d_compact(const double *xyz, const int32_t *valid, const uint8_t *texture, int W,
                    int H, float *out_xyz, uint8_t *out_rgb, uint32_t *out_pix)
{
    int64_t y = 0;
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) {
            const long p = (long)i * W + j, q = PL(i, j);
            if (valid[q] != 1) continue;
            if (out_rgb) {
                /* cvSplit(I1, blue, green, red): texture is BGR-interleaved */
                out_rgb[3 * y + 0] = texture ? texture[3 * p + 2] : 0;
                out_rgb[3 * y + 1] = texture ? texture[3 * p + 1] : 0;
                out_rgb[3 * y + 2] = texture ? texture[3 * p + 0] : 0;
            }
            if (out_xyz) {
                out_xyz[3 * y + 0] = (float)xyz[3 * q + 0];
                out_xyz[3 * y + 1] = (float)xyz[3 * q + 1];
                out_xyz[3 * y + 2] = (float)xyz[3 * q + 2];
            }
            if (out_pix) out_pix[y] = (uint32_t)p;
            y++;
        }
    return y;
}

/* ------------------------------------------------------------------ whole path */

void o3d_reconstruct(const o3d_config *cfg, const o3d_calib *cal, const uint8_t *fringe_v,
                     const uint8_t *gray_v, const uint8_t *inv_v, const uint8_t *fringe_h,
                     const uint8_t *gray_h, const uint8_t *inv_h, const uint8_t *roi,
                     o3d_outputs *out, int threads)
{
    o3d_reconstruct_ex(cfg, cal, fringe_v, gray_v, inv_v, fringe_h, gray_h, inv_h, roi, 0, out, threads);
}

void o3d_reconstruct_ex(const o3d_config *cfg, const o3d_calib *cal, const uint8_t *fringe_v,
                        const uint8_t *gray_v, const uint8_t *inv_v, const uint8_t *fringe_h,
                        const uint8_t *gray_h, const uint8_t *inv_h, const uint8_t *roi,
                        int modulation, o3d_outputs *out, int threads)
{
    const int W = cfg->W, H = cfg->H;
    const size_t n = (size_t)W * H;

    /* m_tech_project_console.cpp:372-384: wrapped(0), wrapped(1), unwrap(0), unwrap(1) */
    for (int dir = 0; dir < cfg->dirs; dir++) {
        int32_t *valid = dir == 0 ? out->valid_v : out->valid_h;
        float *wr = dir == 0 ? out->wrapped_v : out->wrapped_h;
        float *un = dir == 0 ? out->unwrapped_v : out->unwrapped_h;
        int32_t *code = dir == 0 ? out->code_v : out->code_h;
        const int M = dir == 0 ? cfg->M_v : cfg->M_h;
        memset(wr, 0, n * sizeof(float));
        memset(un, 0, n * sizeof(float));
        /* `modulation` is a bit set: 1 = the commented-out modulation criterion, 2 = check_I_mod_criteria as committed */
        if ((modulation & 1) && cfg->N == 3) o3d_check_I_mod_criteria(dir == 0 ? fringe_v : fringe_h, roi, W, H, valid);
        else if (modulation & 2) o3d_check_roi_strict(roi, cfg->N, W, H, valid);
        else o3d_check_roi(roi, W, H, valid);
        o3d_wrapped_phase(dir == 0 ? fringe_v : fringe_h, cfg->N, W, H, valid, wr, NULL, threads);
        o3d_mask_recurrence(valid, W, H, NULL);
        o3d_decode_gray(dir == 0 ? gray_v : gray_h, dir == 0 ? inv_v : inv_h, M, W, H, valid,
                        code, threads);
        o3d_unwrap(dir, wr, code, valid, W, H, un, threads);
    }
    out->count = 0;
    if (cfg->dirs < 2) return;

    memset(out->cpmap, 0, n * 2 * sizeof(int64_t));
    o3d_compute_c_p_map(out->unwrapped_v, out->unwrapped_h, out->valid_v, out->valid_h,
                        cfg->fw_v, cfg->fw_h, cfg->PW, cfg->PH, W, H, out->cpmap, out->valid,
                        threads);

    /* triangulate(): assign_3d_coordinates (full-frame LUTs, :228-439) + method 3 */
    const size_t np = (size_t)cfg->PW * cfg->PH;
    double *cam_lut = (double *)malloc(2 * n * sizeof(double));
    double *proj_lut = (double *)malloc(2 * np * sizeof(double));
    double A_cam[12], A_proj[12];
    o3d_undistort_lut(cal->Kc, cal->dc, W, H, cam_lut, threads);
    o3d_undistort_lut(cal->Kp, cal->dp, cfg->PW, cfg->PH, proj_lut, threads);
    o3d_compute_A(cal->Kc, cal->rc, cal->tc, A_cam);
    o3d_compute_A(cal->Kp, cal->rp, cal->tp, A_proj);
    memset(out->xyz, 0, n * 3 * sizeof(double));
    o3d_triangulate(A_cam, A_proj, cam_lut, proj_lut, out->cpmap, out->valid, W, H, cfg->PW,
                    cfg->PH, out->xyz, threads);
    out->count = o3d_compact(out->xyz, out->valid, NULL, W, H, out->pts, NULL, out->pix);
    free(cam_lut);
    free(proj_lut);
}

/* The reference's CPU pipeline as it lays its data out: every intermediate plane [col][row], row-outer loops, one
 * thread ("ref-faithful").  Outputs come back in that layout too: valid_*, wrapped_*,
 * unwrapped_*, code_*, valid are [W][H], xyz is [W][H][3]; cpmap, pts and pix as in o3d_reconstruct (row-major
 * c_p_map, raster-ordered points).  tests/test_oracle_golden.py checks it against o3d_reconstruct plane by plane. */
void o3d_reconstruct_colrow(const o3d_config *cfg, const o3d_calib *cal, const uint8_t *fringe_v,
                            const uint8_t *gray_v, const uint8_t *inv_v, const uint8_t *fringe_h,
                            const uint8_t *gray_h, const uint8_t *inv_h, const uint8_t *roi,
                            int modulation, o3d_outputs *out)
{
    g_colrow = 1;
    o3d_reconstruct_ex(cfg, cal, fringe_v, gray_v, inv_v, fringe_h, gray_h, inv_h, roi, modulation, out, 1);
    g_colrow = 0;
}
