// scan3d_stages.cpp -- the reference's stage functions (same names, C++ linkage) as thin host code
// over the C ABI: file I/O and layout shims here, all per-pixel work in libscan3d.so.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/scan3d_compat.h"
#include "../../include/scan3d_host.h"

// ---- PROJECT_GLOBAL/common_variables.h:6-10,23-24 (same defaults) ----
int number_of_codes_vertical = 40, number_of_codes_horizontal = 23;
int number_of_patterns_binary_vertical = 6, number_of_patterns_binary_horizontal = 5;
int number_of_patterns_fringe = 3;
int fringe_width_pixels_vertical = 32, fringe_width_pixels_horizontal = 32;
int Camera_imagewidth = 1600, Camera_imageheight = 1200, Projector_imagewidth = 1280, Projector_imageheight = 720;

int *selected_region = nullptr;
int *valid_map_vertical = nullptr, *valid_map_horizontal = nullptr, *valid_map = nullptr;
int *code_vertical = nullptr, *code_horizontal = nullptr;
float *wrapped_phi_vertical = nullptr, *wrapped_phi_horizontal = nullptr;
float *unwrapped_phi_vertical = nullptr, *unwrapped_phi_horizontal = nullptr;
long int (*c_p_map)[2] = nullptr;
double *intersection_points = nullptr;
int scan3d_compat_export = 1;

namespace {
std::string g_root;
scan3d_ctx* g_ctx = nullptr;
int g_device = 0;
scan3d_config g_cfg{};

[[noreturn]] void die(const char* what, const char* detail)
{
    fprintf(stderr, "\nscan3d: %s: %s\n", what, detail ? detail : "");
    exit(EXIT_FAILURE);
}
void ck(int rc, const char* what)
{
    if (rc != SCAN3D_OK) die(what, g_ctx ? scan3d_last_error(g_ctx) : scan3d_last_error(nullptr));
}
size_t npix() { return (size_t)Camera_imagewidth * Camera_imageheight; }

// (re)create the context when the pattern configuration globals changed, like the reference
// which reads them afresh in every stage
void ensure_ctx()
{
    scan3d_config c{};
    c.W = Camera_imagewidth; c.H = Camera_imageheight; c.PW = Projector_imagewidth; c.PH = Projector_imageheight;
    c.N = number_of_patterns_fringe;
    c.M_v = number_of_patterns_binary_vertical; c.M_h = number_of_patterns_binary_horizontal;
    c.fw_v = fringe_width_pixels_vertical; c.fw_h = fringe_width_pixels_horizontal;
    c.dirs = 2; c.row0 = 0; c.H_total = c.H; c.flags = 0;
    if (g_ctx && memcmp(&c, &g_cfg, sizeof(c)) == 0) return;
    if (g_ctx) {
        // a configuration global changed between two stages: the planes of the old shape cannot be carried over
        // (the reference would go on with stale arrays of the old size here); say so instead of doing it silently
        fprintf(stderr, "scan3d compat: pattern / frame globals changed between stages -- the context is recreated, "
                        "the results of earlier stages are dropped (call the stages again from compute_wrapped_phase)\n");
        scan3d_destroy(g_ctx);
    }
    g_ctx = nullptr;
    ck(scan3d_create(&c, g_device, &g_ctx), "scan3d_create");
    g_cfg = c;
}

template <class T>
T* plane(T*& p, size_t n)
{
    if (!p) p = new T[n]();
    return p;
}

// row-major [H][W] (K values per pixel) <-> the reference's [col][row]: in 32 x 32 blocks, so that both sides stay in
// the cache (the plain double loop strides one side by a whole column: 5-10x slower at 1600 x 1200)
template <int K, class S, class D, class F>
void transpose_blocks(const S* src, D* dst, int rows, int cols, F conv)
{
    constexpr int B = 32;
    for (int r0 = 0; r0 < rows; r0 += B)
        for (int c0 = 0; c0 < cols; c0 += B) {
            const int r1 = r0 + B < rows ? r0 + B : rows, c1 = c0 + B < cols ? c0 + B : cols;
            for (int c = c0; c < c1; c++)
                for (int r = r0; r < r1; r++)
                    for (int k = 0; k < K; k++) dst[((size_t)c * rows + r) * K + k] = conv(src[((size_t)r * cols + c) * K + k]);
        }
}
template <class S, class D>
void to_col_row(const S* src, D* dst)
{
    transpose_blocks<1>(src, dst, Camera_imageheight, Camera_imagewidth, [](S v) { return (D)v; });
}

std::vector<uint8_t> roi_row_major()
{
    if (!selected_region) die("compute_wrapped_phase", "selected_region is not set (image_scissor() fills it in the reference)");
    const int W = Camera_imagewidth, H = Camera_imageheight;
    std::vector<uint8_t> roi(npix());
    // [col][row] -> row-major is the same block transpose with the roles of rows and columns exchanged
    transpose_blocks<1>(selected_region, roi.data(), W, H, [](int v) { return (uint8_t)(v == 1 ? 1 : 0); });
    return roi;
}

void load_planes(const std::string& dir, const char* prefix, int count, std::vector<uint8_t>& out)
{
    out.resize((size_t)count * npix());
    for (int i = 0; i < count; i++) {
        int w = 0, h = 0, rc = SCAN3D_ERR_IO;
        const std::string names[2] = {dir + prefix + "Captured_image_" + std::to_string(i) + ".bmp",
                                      dir + prefix + "Gray_captured_image_" + std::to_string(i) + ".bmp"};
        for (const std::string& n : names) {
            rc = scan3d_read_bmp8(n.c_str(), &w, &h, nullptr, 0);
            if (rc == SCAN3D_OK) {
                if (w != Camera_imagewidth || h != Camera_imageheight) die("captured image has the wrong size", n.c_str());
                rc = scan3d_read_bmp8(n.c_str(), &w, &h, out.data() + (size_t)i * npix(), (int64_t)npix());
                break;
            }
        }
        if (rc != SCAN3D_OK) die("cannot load captured image", names[0].c_str());
    }
}

void export_stage34(int d)
{
    if (!scan3d_compat_export) return;
    const size_t n = npix();
    std::vector<float> f(n);
    std::vector<uint8_t> m(n);
    std::vector<int32_t> c(n);
    if (scan3d_get_plane(g_ctx, SCAN3D_PLANE_WRAPPED_V + d, f.data()) == SCAN3D_OK)
        to_col_row(f.data(), plane(d ? wrapped_phi_horizontal : wrapped_phi_vertical, n));
    if (scan3d_get_plane(g_ctx, d ? 10 : SCAN3D_PLANE_MASK, m.data()) == SCAN3D_OK)
        to_col_row(m.data(), plane(d ? valid_map_horizontal : valid_map_vertical, n));
}
}  // namespace

int scan3d_compat_init(const char* root, int cam_w, int cam_h, int proj_w, int proj_h, int device)
{
    g_root = root ? root : ".";
    Camera_imagewidth = cam_w; Camera_imageheight = cam_h;
    Projector_imagewidth = proj_w; Projector_imageheight = proj_h;
    g_device = device;
    return SCAN3D_OK;
}

void scan3d_compat_shutdown()
{
    if (g_ctx) scan3d_destroy(g_ctx);
    g_ctx = nullptr;
}

scan3d_ctx* scan3d_compat_ctx() { return g_ctx; }

void load_matrices()
{
    ensure_ctx();
    scan3d_calib cal;
    if (scan3d_load_calibration(g_root.c_str(), &cal) != SCAN3D_OK) die("load_matrices", scan3d_host_last_error());
    ck(scan3d_set_calibration(g_ctx, &cal), "scan3d_set_calibration");
}

void compute_wrapped_phase(int pattern_type)
{
    ensure_ctx();
    const std::string dir = g_root + "/Captured_patterns/Fringe_patterns/" + (pattern_type ? "Horizontal" : "Vertical") + "/Undistorted/";
    std::vector<uint8_t> fr;
    load_planes(dir, "", number_of_patterns_fringe, fr);
    const std::vector<uint8_t> roi = roi_row_major();
    ck(scan3d_compute_wrapped_phase(g_ctx, pattern_type, fr.data(), roi.data()), "compute_wrapped_phase");
    export_stage34(pattern_type);
}

void unwrap_phase(int pattern_type)
{
    ensure_ctx();
    const int M = pattern_type ? number_of_patterns_binary_horizontal : number_of_patterns_binary_vertical;
    const std::string dir = g_root + "/Captured_patterns/Coded_patterns/Gray_coded/" + (pattern_type ? "Horizontal" : "Vertical") + "/Undistorted/";
    std::vector<uint8_t> g, gi;
    load_planes(dir, "", M, g);
    load_planes(dir, "inverse_", M, gi);
    ck(scan3d_unwrap_phase(g_ctx, pattern_type, g.data(), gi.data()), "unwrap_phase");
    if (!scan3d_compat_export) return;
    const size_t n = npix();
    std::vector<float> f(n);
    std::vector<int32_t> c(n);
    export_stage34(pattern_type);   // wrapped_phi now holds phi + Pi, like the reference
    ck(scan3d_get_plane(g_ctx, SCAN3D_PLANE_UNWRAPPED_V + pattern_type, f.data()), "get unwrapped");
    to_col_row(f.data(), plane(pattern_type ? unwrapped_phi_horizontal : unwrapped_phi_vertical, n));
    ck(scan3d_get_code_i32(g_ctx, pattern_type, c.data()), "get code");
    to_col_row(c.data(), plane(pattern_type ? code_horizontal : code_vertical, n));
}

void compute_c_p_map()
{
    ensure_ctx();
    ck(scan3d_compute_c_p_map(g_ctx), "compute_c_p_map");
    if (!scan3d_compat_export) return;
    const size_t n = npix();
    std::vector<uint8_t> v(n);
    ck(scan3d_get_plane(g_ctx, SCAN3D_PLANE_VALID, v.data()), "get valid");
    to_col_row(v.data(), plane(valid_map, n));
    if (!c_p_map) c_p_map = new long int[n][2]();
    std::vector<int64_t> cp(2 * n);
    ck(scan3d_get_cpmap_i64(g_ctx, cp.data()), "get c_p_map");
    for (size_t i = 0; i < n; i++) { c_p_map[i][0] = (long)cp[2 * i]; c_p_map[i][1] = (long)cp[2 * i + 1]; }
}

void triangulate()
{
    ensure_ctx();
    scan3d_calib probe;   // read_parameters(): the reference re-reads the matrices on every call
    if (scan3d_load_calibration(g_root.c_str(), &probe) == SCAN3D_OK) ck(scan3d_set_calibration(g_ctx, &probe), "scan3d_set_calibration");
    ck(scan3d_triangulate(g_ctx), "triangulate");
    if (!scan3d_compat_export) return;
    const size_t n = npix();
    std::vector<double> x(3 * n);
    ck(scan3d_get_plane(g_ctx, SCAN3D_PLANE_XYZ, x.data()), "get xyz");
    double* dst = plane(intersection_points, 3 * n);
    const int W = Camera_imagewidth, H = Camera_imageheight;
    transpose_blocks<3>(x.data(), dst, H, W, [](double v) { return v; });
}

void save_point_cloud(unsigned cloud_index)
{
    ensure_ctx();
    // texture.bmp, loaded in colour and split into blue / green / red as the reference does
    // (8/save_point_cloud.cpp:59-66,88-90); optional here: without it the points stay black
    int w = 0, h = 0;
    const std::string tex = g_root + "/Point_cloud/texture.bmp";
    if (scan3d_read_bmp_bgr(tex.c_str(), &w, &h, nullptr, 0) == SCAN3D_OK && w == Camera_imagewidth && h == Camera_imageheight) {
        std::vector<uint8_t> bgr(3 * npix());
        if (scan3d_read_bmp_bgr(tex.c_str(), &w, &h, bgr.data(), (int64_t)bgr.size()) != SCAN3D_OK)
            die("save_point_cloud", scan3d_host_last_error());
        ck(scan3d_set_texture(g_ctx, bgr.data()), "scan3d_set_texture");
    }
    int64_t n = 0;
    ck(scan3d_compact_points(g_ctx, &n), "save_point_cloud");
    char name[64];
    snprintf(name, sizeof(name), "/Point_cloud/point_cloud_%u.ply", cloud_index);
    ck(scan3d_write_ply(g_ctx, (g_root + name).c_str(), 0), "scan3d_write_ply");
    snprintf(name, sizeof(name), "/Point_cloud/point_cloud_%u.pcd", cloud_index);      // 8/save_point_cloud.cpp:211-212
    ck(scan3d_write_pcd(g_ctx, (g_root + name).c_str()), "scan3d_write_pcd");
    fprintf(stderr, "Saved %lld data points to %s\n", (long long)n, (g_root + name).c_str());
}

void generate_pattern()
{
    // save_pattern_images (1/pattern_generator.cpp:436-479): Fringe_patterns/{Vertical,Horizontal}/Pattern_k.bmp,
    // Coded_patterns/Gray_coded/{Vertical,Horizontal}/[inverse_]Pattern_j.bmp -- generated on the GPU
    ensure_ctx();
    const size_t pp = (size_t)g_cfg.PW * g_cfg.PH;
    for (int d = 0; d < 2; d++) {
        const int M = d == 0 ? g_cfg.M_v : g_cfg.M_h;
        const char* dir = d == 0 ? "Vertical" : "Horizontal";
        std::vector<uint8_t> img((size_t)scan3d_pattern_bytes(&g_cfg, d));
        ck(scan3d_generate_patterns(g_ctx, d, img.data()), "scan3d_generate_patterns");
        char name[160];
        auto save = [&](const char* fmt, int idx, size_t plane) {
            snprintf(name, sizeof(name), fmt, dir, idx);
            if (scan3d_write_bmp8((g_root + name).c_str(), g_cfg.PW, g_cfg.PH, img.data() + plane * pp) != SCAN3D_OK)
                die("generate_pattern", scan3d_host_last_error());
        };
        for (int k = 0; k < g_cfg.N; k++) save("/Generated_patterns/Fringe_patterns/%s/Pattern_%d.bmp", k, (size_t)k);
        for (int j = 0; j < M; j++) {
            save("/Generated_patterns/Coded_patterns/Gray_coded/%s/Pattern_%d.bmp", j, (size_t)(g_cfg.N + j));
            save("/Generated_patterns/Coded_patterns/Gray_coded/%s/inverse_Pattern_%d.bmp", j, (size_t)(g_cfg.N + M + j));
        }
    }
}

void register_point_clouds(unsigned num_point_clouds, float tx, float ty, float tz, float rot_step)
{
    ensure_ctx();
    // first pass: total size (9/register_point_clouds.cpp:63-68)
    std::vector<int64_t> sizes(num_point_clouds, 0);
    int64_t total = 0;
    char name[64];
    for (unsigned i = 0; i < num_point_clouds; i++) {
        snprintf(name, sizeof(name), "/Point_cloud/point_cloud_%u.ply", i);
        if (scan3d_read_ply_points((g_root + name).c_str(), nullptr, nullptr, 0, &sizes[i]) != SCAN3D_OK)
            die("register_point_clouds", scan3d_host_last_error());
        total += sizes[i];
    }
    fprintf(stderr, "\nregistred cloud width:%lld", (long long)total);
    std::vector<float> xyz((size_t)total * 3);
    std::vector<uint8_t> rgb((size_t)total * 3);
    float theta = 0.0f;                       // :78
    int64_t prev_last_point_id = 0;
    for (unsigned i = 0; i < num_point_clouds; i++) {
        snprintf(name, sizeof(name), "/Point_cloud/point_cloud_%u.ply", i);
        float* dst = xyz.data() + 3 * prev_last_point_id;
        int64_t n = 0;
        if (scan3d_read_ply_points((g_root + name).c_str(), dst, rgb.data() + 3 * prev_last_point_id, sizes[i], &n) != SCAN3D_OK)
            die("register_point_clouds", scan3d_host_last_error());
        ck(scan3d_register_points(g_ctx, dst, dst, n, theta, tx, ty, tz), "scan3d_register_points");   // :93-137
        fprintf(stderr, "\nRotating by :%f", theta);
        theta += rot_step;                    // :141
        prev_last_point_id += n;
    }
    if (scan3d_write_ply_points((g_root + "/Point_cloud/registered_point_cloud.ply").c_str(), xyz.data(), rgb.data(), total, 0) != SCAN3D_OK)
        die("register_point_clouds", scan3d_host_last_error());
}

void image_scissor_fill(const unsigned char* internal_image)
{
    ensure_ctx();
    std::vector<uint8_t> roi(npix()), filled(npix());
    ck(scan3d_roi_fill(g_ctx, internal_image, roi.data(), filled.data()), "scan3d_roi_fill");
    static std::vector<int> region;   // m_tech_project_console.cpp:183 allocates (and leaks) a new plane per call
    region.assign(npix(), 0);
    selected_region = region.data();
    const int W = Camera_imagewidth, H = Camera_imageheight;
    transpose_blocks<1>(roi.data(), selected_region, H, W, [](uint8_t v) { return (int)v; });
    if (scan3d_write_bmp8((g_root + "/i1.bmp").c_str(), W, H, filled.data()) != SCAN3D_OK)
        die("image_scissor_fill", scan3d_host_last_error());
}

void undistort_capture(const unsigned char* cap, unsigned char* undist_cap, int device_kind)
{
    ensure_ctx();
    ck(scan3d_undistort_frames(g_ctx, device_kind, cap, 1, undist_cap), "scan3d_undistort_frames");
}

void reconstruct_scan(unsigned cloud_index)
{
    ensure_ctx();
    std::vector<uint8_t> stack((size_t)scan3d_stack_bytes(&g_cfg));
    if (scan3d_load_captured_set(g_root.c_str(), &g_cfg, stack.data()) != SCAN3D_OK) die("reconstruct_scan", scan3d_host_last_error());
    const std::vector<uint8_t> roi = roi_row_major();
    int64_t n = 0;
    ck(scan3d_reconstruct(g_ctx, stack.data(), roi.data(), &n), "scan3d_reconstruct");
    char name[64];
    snprintf(name, sizeof(name), "/Point_cloud/point_cloud_%u.ply", cloud_index);
    ck(scan3d_write_ply(g_ctx, (g_root + name).c_str(), 0), "scan3d_write_ply");
}
