// fused_math_host.cpp -- HOST build of the fused kernel's arithmetic, for the CPU test suite.
//
// Includes the product's own device headers (3dscan_b200/csrc/scan3d_math.cuh, scan3d_fused_math.cuh) through
// tests/cuda_host_shim.h and walks a whole captured stack the way k_fused7's consumer threads do: 4 consecutive
// pixels per "thread", SWAR integer phase, FP64 phase, correspondence, triangulation, raster-order compaction.
// What this checks without a GPU: every arithmetic function the kernels call, on whole scans, against the
// oracle and the reference's golden images.  What it cannot check: the kernels' pipeline (TMA, mbarriers,
// look-back) -- that is the GPU tests' job.  Built by tests/test_fused_math_host.py:
//   g++ -O2 -std=c++17 -ffp-contract=off -I/usr/local/cuda/include -shared -fPIC
#include "cuda_host_shim.h"

#include "../3dscan_b200/csrc/scan3d_fused_math.cuh"

#include <vector>

using namespace s3d;

namespace {

struct HostArgs {
    int W, H, PW, PH, N, M_v, M_h, fw_v, fw_h, dirs, exact;
    int mask_is_final;   // roi already is the post-recurrence mask (the reference's stored images give only that)
};

template <int N>
void run(const HostArgs& a, const DeviceCalib& cal, const uint8_t* stack, const uint8_t* roi, const double* tab,
         float* unw_v, float* unw_h, int16_t* code_v, int16_t* code_h, uint8_t* valid, int32_t* cpmap, float* pts,
         int64_t* count)
{
    const int W = a.W, H = a.H;
    const size_t plane = (size_t)W * H;
    const int NF = a.dirs == 2 ? 2 * N + 2 * (a.M_v + a.M_h) : N + 2 * a.M_v;
    const bool fastdiv = true;
    // the undistorted-pixel tables of k_undistort_lut, only for distorted devices (as scan3d_set_calibration)
    std::vector<double> cam_lut, proj_lut;
    if (a.dirs == 2 && cal.cam_distorted) {
        cam_lut.resize(2 * plane);
        for (size_t p = 0; p < plane; p++)
            undistorted_pixel(cal.Kc, cal.dc, (double)(p % W), (double)(p / W), &cam_lut[2 * p], &cam_lut[2 * p + 1]);
    }
    if (a.dirs == 2 && cal.proj_distorted) {
        const size_t pp = (size_t)a.PW * a.PH;
        proj_lut.resize(2 * pp);
        for (size_t p = 0; p < pp; p++)
            undistorted_pixel(cal.Kp, cal.dp, (double)(p % a.PW), (double)(p / a.PW), &proj_lut[2 * p], &proj_lut[2 * p + 1]);
    }
    int64_t n_pts = 0;
    std::vector<uint32_t> sw(NF);
    for (size_t g = 0; g < plane; g += 4) {           // one consumer thread's 4 pixels
        const int y = (int)(g / W), xt = (int)(g % W);
        // mask: ROI after the raster recurrence, closed form (k_fused7's slow path; the fast path is its special case)
        uint32_t mbits = 0;
        for (int j = 0; j < 4; j++) {
            const int x = xt + j;
            auto inv = [&](int gx, int gy) { return roi[(size_t)gy * W + gx] == 0; };
            bool v = !inv(x, y);
            const bool border = x == 0 || y == 0 || x == W - 1 || y == H - 1;
            if (v && !border && !a.mask_is_final) v = !mask_trigger(x, y, W, H, inv);
            mbits |= (v ? 1u : 0u) << j;
        }
        Terms Tv{}, Th{};
        uint32_t gvA = 0, gvB = 0, ghA = 0, ghB = 0;
        if (mbits) {
            for (int f = 0; f < NF; f++) memcpy(&sw[f], stack + (size_t)f * plane + g, 4);   // little-endian word, as in shared memory
            fringe_terms<N>(sw.data(), 0, 1, 0, Tv);
            gray_bits(sw.data(), N, N + a.M_v, a.M_v, 1, 0, gvA, gvB);
            if (a.dirs == 2) {
                const int fh = N + 2 * a.M_v;
                fringe_terms<N>(sw.data(), fh, 1, 0, Th);
                gray_bits(sw.data(), fh + N, fh + N + a.M_h, a.M_h, 1, 0, ghA, ghB);
            }
        }
        for (int j = 0; j < 4; j++) {
            // the kernel's pass structure: pairs (2h, 2h + 1), registers selected once per pair by h
            const int h = j >> 1, u = j & 1;
            const TermsPair Pv = terms_pair(Tv, h), Ph = terms_pair(Th, h);
            const uint32_t gvA_h = gvA >> (16 * h), gvB_h = gvB >> (16 * h), ghA_h = ghA >> (16 * h), ghB_h = ghB >> (16 * h);
            const int x = xt + j;
            const size_t p = g + j;
            const bool m = (mbits >> j) & 1u;
            const int cv = code_of_pair(gvA_h, gvB_h, u, a.M_v);
            if (cv != code_of(gvA, gvB, j, a.M_v)) abort();
            float unwv = 0.0f, unwh = 0.0f;
            if (m) {                                   // the kernel evaluates unconditionally and selects; same values
                const float wv = add_pi(phase_of_pair<N>(Pv, u, tab));
                unwv = (x == 0 || x == W - 1) ? 0.0f : unwrap_abs(wv, cv, fastdiv);
            }
            bool v = m;
            unw_v[p] = unwv;
            code_v[p] = (int16_t)(m ? cv : -1);
            if (a.dirs == 2) {
                const int ch = code_of_pair(ghA_h, ghB_h, u, a.M_h);
                if (m) {
                    const float wh = add_pi(phase_of_pair<N>(Ph, u, tab));
                    unwh = (y == 0 || y == H - 1) ? 0.0f : unwrap_abs(wh, ch, fastdiv);
                }
                int px = 0, py = 0;
                const bool okx = correspond32(unwv, (double)a.fw_v, &px);
                const bool oky = correspond32(unwh, (double)a.fw_h, &py);
                const int cpx = (m && okx) ? px : 0, cpy = (m && okx && oky) ? py : 0;
                v = m && okx && oky && (unsigned)px <= (unsigned)(a.PW - 1) && (unsigned)py <= (unsigned)(a.PH - 1);
                unw_h[p] = unwh;
                code_h[p] = (int16_t)(m ? ch : -1);
                cpmap[2 * p] = cpx;
                cpmap[2 * p + 1] = cpy;
                if (v) {
                    double uc, vc, up, vp, Xd[3];
                    if (!cam_lut.empty()) { uc = cam_lut[2 * p]; vc = cam_lut[2 * p + 1]; }
                    else undistorted_pixel_nodist(cal.Kc, cal.ifx_c, cal.ify_c, cal.cam_std != 0, (double)x, (double)y, &uc, &vc);
                    if (!proj_lut.empty()) {
                        const size_t q = (size_t)cpy * a.PW + cpx;
                        up = proj_lut[2 * q]; vp = proj_lut[2 * q + 1];
                    } else {
                        undistorted_pixel_nodist(cal.Kp, cal.ifx_p, cal.ify_p, cal.proj_std != 0, (double)cpx, (double)cpy, &up, &vp);
                    }
                    if (a.exact) triangulate_point(cal.Ac, cal.Ap, uc, vc, up, vp, Xd);
                    else triangulate_point_fast(cal.Ac, cal.Ap, uc, vc, up, vp, Xd);
                    pts[3 * n_pts + 0] = __double2float_rn(Xd[0]);
                    pts[3 * n_pts + 1] = __double2float_rn(Xd[1]);
                    pts[3 * n_pts + 2] = __double2float_rn(Xd[2]);
                    n_pts++;
                }
            }
            valid[p] = v ? 1 : 0;
        }
    }
    *count = n_pts;
}

}  // namespace

extern "C" {

// cfg = {W, H, PW, PH, N, M_v, M_h, fw_v, fw_h, dirs, exact, mask_is_final}; K/d as in scan3d_calib; A_cam / A_proj = K[R|t]
// (3x4 row-major, from the oracle's compute_A).  W % 4 == 0.
int s3d_host_fused_math(const int* cfg, const double* Kc, const double* dc, const double* Kp, const double* dp,
                        const double* A_cam, const double* A_proj, const uint8_t* stack,
                        const uint8_t* roi, float* unw_v, float* unw_h, int16_t* code_v, int16_t* code_h, uint8_t* valid,
                        int32_t* cpmap, float* pts, int64_t* count)
{
    double atan_tab[ATAN_TAB_DOUBLES];
    fill_atan_table(atan_tab);   // the library's own table builder (scan3d_math.cuh)
    HostArgs a{cfg[0], cfg[1], cfg[2], cfg[3], cfg[4], cfg[5], cfg[6], cfg[7], cfg[8], cfg[9], cfg[10], cfg[11]};
    if (a.W % 4 != 0 || a.W < 4) return -1;
    DeviceCalib cal{};
    memcpy(cal.Ac, A_cam, sizeof(cal.Ac));
    memcpy(cal.Ap, A_proj, sizeof(cal.Ap));
    memcpy(cal.Kc, Kc, sizeof(cal.Kc));
    memcpy(cal.dc, dc, sizeof(cal.dc));
    memcpy(cal.Kp, Kp, sizeof(cal.Kp));
    memcpy(cal.dp, dp, sizeof(cal.dp));
    cal.ifx_c = 1. / Kc[0]; cal.ify_c = 1. / Kc[4];
    cal.ifx_p = 1. / Kp[0]; cal.ify_p = 1. / Kp[4];
    auto std_form = [](const double* K) { return K[1] == 0.0 && K[3] == 0.0 && K[6] == 0.0 && K[7] == 0.0 && K[8] == 1.0; };
    cal.cam_std = std_form(Kc);
    cal.proj_std = std_form(Kp);
    cal.fast_div_ok = 1;
    for (int i = 0; i < 5; i++) {
        if (dc[i] != 0.0) cal.cam_distorted = 1;
        if (dp[i] != 0.0) cal.proj_distorted = 1;
    }
    switch (a.N) {
#define CASE(NN) case NN: run<NN>(a, cal, stack, roi, atan_tab, unw_v, unw_h, code_v, code_h, valid, cpmap, pts, count); return 0;
        CASE(3) CASE(4) CASE(5) CASE(8)
#undef CASE
    }
    return -2;
}

// (float)atan2(y, x) as the kernels evaluate it, for n pairs
void s3d_host_atan2_to_float(const double* y, const double* x, int n, float* out)
{
    double atan_tab[ATAN_TAB_DOUBLES];
    fill_atan_table(atan_tab);
    for (int i = 0; i < n; i++) out[i] = atan2_to_float(y[i], x[i], (float)y[i], (float)x[i], atan_tab);
}

// the stage kernels' wrapped phase (k_wrapped: wrapped_phase<N, false>; N = 0 is the generic-N extension with the
// sin / cos weights scan3d_create() computes) for n pixels; I = [N][n] samples
int s3d_host_stage_wrapped_phase(const uint8_t* I, int N, int n, float* out)
{
    double atan_tab[ATAN_TAB_DOUBLES], w[128];
    fill_atan_table(atan_tab);
    memset(w, 0, sizeof(w));
    for (int k = 0; k < N && k < 64; k++) {          // scan3d_create()
        const double a = 2.0 * 3.14159265358979323846 * (double)k / (double)N;
        w[k] = sin(a);
        w[64 + k] = cos(a);
    }
    if (N < 3 || N > 16) return -1;
    for (int p = 0; p < n; p++) {
        int v[16];
        for (int k = 0; k < N; k++) v[k] = I[(size_t)k * n + p];
        switch (N) {
            case 3: out[p] = wrapped_phase<3, false>(v, w, w + 64, N, atan_tab, atan_tab + 33); break;
            case 4: out[p] = wrapped_phase<4, false>(v, w, w + 64, N, atan_tab, atan_tab + 33); break;
            case 5: out[p] = wrapped_phase<5, false>(v, w, w + 64, N, atan_tab, atan_tab + 33); break;
            case 8: out[p] = wrapped_phase<8, false>(v, w, w + 64, N, atan_tab, atan_tab + 33); break;
            default: out[p] = wrapped_phase<0, false>(v, w, w + 64, N, atan_tab, atan_tab + 33); break;
        }
    }
    return 0;
}
}
