// scan3d_synth.cpp -- synthetic "captured" pattern stacks for tests and benchmarks.
//
// The projected patterns follow the reference's stage-1 conventions (1/pattern_generator.cpp):
// fringe k of an N-step set is 127 + 128*cos((x/fw)*2*Pi - Pi + delta_k) with the reference's
// shifts (N=3: (k-1)*Pi/2, :302; N=4: k*Pi/2, :342; N=5: (k-2)*Pi/2, :381; extension N=8:
// k*pi/4), Gray bit i of stripe floor(x/fw) is B(i-1) xor B(i) with bit 0 the MSB (:83-100),
// inverse = 255 - pattern (:497).  A scene (plane + sphere) is rendered through the same
// pinhole + distortion model the triangulation stage assumes, so that decoding and
// triangulating the stack returns the scene.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/scan3d_host.h"
#include "../common/scan3d_pattern_profile.h"

namespace {

const double PI_REF = 22.0 / 7.0;
const double PI_TRUE = 3.14159265358979323846;

inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
// approximately N(0,1): sum of 4 uniforms, variance-normalised (deterministic, branch-free)
inline double gauss(uint64_t key)
{
    const uint64_t a = mix64(key), b = mix64(a);
    const double s = u01(a) + u01(a << 17 | a >> 47) + u01(b) + u01(b << 23 | b >> 41);
    return (s - 2.0) * 1.7320508075688772;
}

void rodrigues(const double r[3], double R[9])
{
    const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < 1e-300) {
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0);
        return;
    }
    const double c = cos(th), s = sin(th), c1 = 1 - c, x = r[0] / th, y = r[1] / th, z = r[2] / th;
    const double M[9] = {c + c1 * x * x,     c1 * x * y - s * z, c1 * x * z + s * y,
                         c1 * x * y + s * z, c + c1 * y * y,     c1 * y * z - s * x,
                         c1 * x * z - s * y, c1 * y * z + s * x, c + c1 * z * z};
    memcpy(R, M, sizeof(M));
}

// the same 5-iteration inverse the pipeline applies to pixel centres
void undistort_norm(const double K[9], const double k[5], double u, double v, double* ox, double* oy)
{
    const double x0 = (u - K[2]) / K[0], y0 = (v - K[5]) / K[4];
    double x = x0, y = y0;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double ic = 1. / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        const double dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
        const double dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
        x = (x0 - dx) * ic;
        y = (y0 - dy) * ic;
    }
    *ox = x;
    *oy = y;
}

// forward projection into the projector.  The pipeline treats undistort(pixel) as the ideal
// coordinate, so the ideal->pixel map must be its inverse; a few Newton-free fixed-point steps
// of the forward distortion model are accurate to well below a projector pixel.
void project(const double K[9], const double k[5], double xn, double yn, double* px, double* py)
{
    const double r2 = xn * xn + yn * yn;
    const double cd = 1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2;
    const double xd = xn * cd + 2 * k[2] * xn * yn + k[3] * (r2 + 2 * xn * xn);
    const double yd = yn * cd + k[2] * (r2 + 2 * yn * yn) + 2 * k[3] * xn * yn;
    *px = K[0] * xd + K[2];
    *py = K[4] * yd + K[5];
}

inline double fringe_shift(int N, int k, double pi)
{
    if (N == 3) return (k - 1) * (pi / 2.0);
    if (N == 4) return k * (pi / 2.0);
    if (N == 5) return (k - 2) * (pi / 2.0);
    return 2.0 * PI_TRUE * k / N;   // extension: equally spaced over the true period
}

using s3d_profile::gray_bit;

inline uint8_t clamp_u8(double v)
{
    const double r = floor(v + 0.5);
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

}  // namespace

extern "C" {

void scan3d_synth_default_params(scan3d_synth_params* p)
{
    memset(p, 0, sizeof(*p));
    p->sphere_c[0] = 60.0; p->sphere_c[1] = 40.0; p->sphere_c[2] = -30.0;
    p->sphere_r = 25.0;
    p->plane_z = 0.0;
    p->albedo_lo = 0.3; p->albedo_hi = 1.0;
    p->ambient_max = 20.0;
    p->noise_sigma = 1.5;
    p->roi_fraction = 0.75;
    p->true_pi = 1;
    p->projector_pixelated = 0;
    p->seed = 0x3D5CA9ull;
}

void scan3d_scale_calibration(const scan3d_calib* in, double cs, double ps, scan3d_calib* out)
{
    *out = *in;
    out->Kc[0] *= cs; out->Kc[4] *= cs; out->Kc[2] *= cs; out->Kc[5] *= cs;
    out->Kp[0] *= ps; out->Kp[4] *= ps; out->Kp[2] *= ps; out->Kp[5] *= ps;
}

int scan3d_synth_pattern_row(int kind, int n_or_m, int fw, int k, int length, uint8_t* out)
{
    if (!out || fw < 1 || length < 1) return SCAN3D_ERR_ARG;
    s3d_profile::pattern_profile(kind, n_or_m, fw, k, length, out);
    return SCAN3D_OK;
}

int scan3d_synth_stack(const scan3d_config* cfg, const scan3d_calib* cal, const scan3d_synth_params* p,
                       uint8_t* stack, uint8_t* roi_full, float* truth, int threads)
{
    if (!cfg || !cal || !p || !stack) return SCAN3D_ERR_ARG;
    const int W = cfg->W, H = cfg->H, Ht = cfg->H_total > 0 ? cfg->H_total : cfg->H, row0 = cfg->row0;
    const int N = cfg->N, D = cfg->dirs;
    const size_t plane = (size_t)W * H;
    const double pi = p->true_pi ? PI_TRUE : PI_REF;
#ifdef _OPENMP
    const int nth = threads > 0 ? threads : omp_get_max_threads();
#else
    const int nth = 1;
    (void)threads;
#endif
    double Rc[9], Rp[9];
    rodrigues(cal->rc, Rc);
    rodrigues(cal->rp, Rp);
    // camera centre in the world frame: C = -Rc^T tc
    double C[3];
    for (int i = 0; i < 3; i++) C[i] = -(Rc[0 * 3 + i] * cal->tc[0] + Rc[1 * 3 + i] * cal->tc[1] + Rc[2 * 3 + i] * cal->tc[2]);

    // ROI: centred ellipse of the requested area fraction (full frame, every rank identical)
    if (roi_full) {
        const double f = sqrt(p->roi_fraction / PI_TRUE);   // a = f*W, b = f*H  ->  pi*a*b = frac*W*H
        const double a = f * W, b = f * Ht, cx = 0.5 * (W - 1), cy = 0.5 * (Ht - 1);
#pragma omp parallel for num_threads(nth) schedule(static)
        for (int y = 0; y < Ht; y++)
            for (int x = 0; x < W; x++) {
                const double dx = (x - cx) / a, dy = (y - cy) / b;
                const bool in = dx * dx + dy * dy <= 1.0 && x >= 2 && y >= 2 && x < W - 2 && y < Ht - 2;
                roi_full[(size_t)y * W + x] = in ? 1 : 0;
            }
    }

    // plane offsets inside the stack
    size_t off_fr[2], off_g[2], off_i[2];
    {
        size_t o = 0;
        for (int d = 0; d < D; d++) {
            const int M = d == 0 ? cfg->M_v : cfg->M_h;
            off_fr[d] = o; o += (size_t)N * plane;
            off_g[d] = o;  o += (size_t)M * plane;
            off_i[d] = o;  o += (size_t)M * plane;
        }
    }

#pragma omp parallel for num_threads(nth) schedule(dynamic, 8)
    for (int yl = 0; yl < H; yl++) {
        const int y = row0 + yl;
        for (int x = 0; x < W; x++) {
            const size_t pidx = (size_t)yl * W + x;
            const uint64_t pkey = mix64(p->seed ^ ((uint64_t)y * 0x100000001B3ull + (uint64_t)x));
            // camera ray
            double xn, yn;
            undistort_norm(cal->Kc, cal->dc, (double)x, (double)y, &xn, &yn);
            double d[3];
            for (int i = 0; i < 3; i++) d[i] = Rc[0 * 3 + i] * xn + Rc[1 * 3 + i] * yn + Rc[2 * 3 + i];
            // nearest hit: sphere, else plane
            double lam = -1.0;
            if (p->sphere_r > 0) {
                double oc[3] = {C[0] - p->sphere_c[0], C[1] - p->sphere_c[1], C[2] - p->sphere_c[2]};
                const double A = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                const double B = 2 * (oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2]);
                const double Cq = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - p->sphere_r * p->sphere_r;
                const double disc = B * B - 4 * A * Cq;
                if (disc >= 0) {
                    const double l = (-B - sqrt(disc)) / (2 * A);
                    if (l > 0) lam = l;
                }
            }
            if (lam < 0 && d[2] != 0.0) {
                const double l = (p->plane_z - C[2]) / d[2];
                if (l > 0) lam = l;
            }
            bool lit = lam > 0;
            double X[3] = {0, 0, 0}, px = -1, py = -1;
            if (lit) {
                for (int i = 0; i < 3; i++) X[i] = C[i] + lam * d[i];
                double Xp[3];
                for (int i = 0; i < 3; i++) Xp[i] = Rp[i * 3 + 0] * X[0] + Rp[i * 3 + 1] * X[1] + Rp[i * 3 + 2] * X[2] + cal->tp[i];
                if (Xp[2] > 0) {
                    project(cal->Kp, cal->dp, Xp[0] / Xp[2], Xp[1] / Xp[2], &px, &py);
                    lit = px >= 0 && py >= 0 && px < cfg->PW && py < cfg->PH;
                } else {
                    lit = false;
                }
            }
            if (truth) {
                const float nanv = nanf("");
                truth[3 * pidx + 0] = lam > 0 ? (float)X[0] : nanv;
                truth[3 * pidx + 1] = lam > 0 ? (float)X[1] : nanv;
                truth[3 * pidx + 2] = lam > 0 ? (float)X[2] : nanv;
            }
            // smooth albedo + per-pixel ambient
            const double tex = 0.5 + 0.25 * sin(0.013 * x + 0.007 * y) + 0.25 * cos(0.011 * y - 0.005 * x);
            const double albedo = p->albedo_lo + (p->albedo_hi - p->albedo_lo) * tex;
            const double ambient = p->ambient_max * u01(mix64(pkey ^ 0xA5A5ull));
            int frame = 0;
            for (int dd = 0; dd < D; dd++) {
                const int M = dd == 0 ? cfg->M_v : cfg->M_h;
                const int fw = dd == 0 ? cfg->fw_v : cfg->fw_h;
                double q = dd == 0 ? px : py;                 // projector coordinate along the code axis
                if (p->projector_pixelated) q = floor(q);
                for (int k = 0; k < N; k++, frame++) {
                    double v = ambient;
                    if (lit) v += albedo * (127.0 + 128.0 * cos((q / fw) * 2.0 * pi - pi + fringe_shift(N, k, pi)));
                    v += p->noise_sigma * gauss(pkey + 0x1000ull * (frame + 1));
                    stack[off_fr[dd] + (size_t)k * plane + pidx] = clamp_u8(v);
                }
                const int stripe = lit ? (int)floor(q / fw) : 0;
                for (int i = 0; i < M; i++, frame += 2) {
                    const int g = lit ? gray_bit(stripe, i, M) : 0;
                    double v0 = ambient + (lit ? albedo * (g ? 255.0 : 0.0) : 0.0);
                    double v1 = ambient + (lit ? albedo * (g ? 0.0 : 255.0) : 0.0);
                    v0 += p->noise_sigma * gauss(pkey + 0x1000ull * (frame + 1));
                    v1 += p->noise_sigma * gauss(pkey + 0x1000ull * (frame + 2));
                    stack[off_g[dd] + (size_t)i * plane + pidx] = clamp_u8(v0);
                    stack[off_i[dd] + (size_t)i * plane + pidx] = clamp_u8(v1);
                }
            }
        }
    }
    return SCAN3D_OK;
}

}  // extern "C"
