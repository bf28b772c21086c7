// m_tech_console.cpp -- the reference's main() (M_tech_project_console/m_tech_project_console.cpp:244-412) without
// its interactive parts, on top of libscan3d_compat.so: the same calls in the same order, every per-pixel stage on
// the B200.  Capture (camera, projector window) and the mouse lasso are replaced by files:
//
//   <root>/Captured_patterns/...                    the captured (already undistorted) images of every view, as the
//                                                   reference's capture stage leaves them (2/project_pattern.cpp)
//   <root>/i1_outline.bmp (optional)                the lasso outline image_scissor() would have recorded; without it
//                                                   the whole frame minus a 2-pixel border is selected
//
//   g++ -O2 -std=c++17 -I include examples/m_tech_console.cpp -L 3dscan_b200/lib -lscan3d_compat -lscan3d_host -lscan3d
//       (+ -Wl,-rpath,$PWD/3dscan_b200/lib -o m_tech_console)
//   ./m_tech_console <root> [n_scans] [rot_step_degrees] [tx ty tz]
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <string>
#include <vector>

#include "scan3d_compat.h"
#include "scan3d_host.h"

// SCAN3D_CONSOLE_TIMES=1: wall clock of every reference-named stage on stderr (file reads, the [col][row]
// transposes of the exported globals and the host<->device copies included: this is the drop-in path as linked)
static bool g_times = false;
template <class F>
static void timed(const char* name, F&& f)
{
    const auto t0 = std::chrono::steady_clock::now();
    f();
    if (g_times)
        fprintf(stderr, "  %-28s %9.2f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
}

int main(int argc, char** argv)
{
    g_times = getenv("SCAN3D_CONSOLE_TIMES") != nullptr;
    if (argc < 2) {
        fprintf(stderr, "usage: %s <M_tech_project_console directory> [n_scans] [rot_step] [tx ty tz]\n", argv[0]);
        return 2;
    }
    const char* root = argv[1];
    const unsigned n_scans = argc > 2 ? (unsigned)atoi(argv[2]) : 1;          // m_tech_project_console.cpp:269-283
    const float rot_step = argc > 3 ? (float)atof(argv[3]) : 0.0f;
    const float tx = argc > 6 ? (float)atof(argv[4]) : 0.0f, ty = argc > 6 ? (float)atof(argv[5]) : 0.0f,
                tz = argc > 6 ? (float)atof(argv[6]) : 0.0f;

    // PROJECT_GLOBAL/global_cv.h:49-53 and common_variables.h:6-10 (the reference's compile-time configuration)
    if (scan3d_compat_init(root, 1600, 1200, 1280, 720, 0) != 0) return 1;
    number_of_patterns_fringe = 3;
    number_of_patterns_binary_vertical = 6;
    number_of_patterns_binary_horizontal = 5;
    fringe_width_pixels_vertical = fringe_width_pixels_horizontal = 32;

    timed("generate_pattern", [] { generate_pattern(); });      // STEP-1 (:300)
    timed("load_matrices", [] { load_matrices(); });            // STEP-2: the stored calibration (:310-330 run the calibration itself)

    unsigned t = 0;
    for (unsigned scan = 0; scan < n_scans; scan++) {
        // STEP-3 (project_pattern: capture) is hardware; the captured set is read from <root> by the stages below
        // STEP-4 & 5: region selection (:365)
        const int W = Camera_imagewidth, H = Camera_imageheight;
        std::vector<unsigned char> outline((size_t)W * H, 0);
        int ow = 0, oh = 0;
        const std::string path = std::string(root) + "/i1_outline.bmp";
        if (!(scan3d_read_bmp8(path.c_str(), &ow, &oh, nullptr, 0) == 0 && ow == W && oh == H &&
              scan3d_read_bmp8(path.c_str(), &ow, &oh, outline.data(), (int64_t)outline.size()) == 0)) {
            for (int r = 1; r < H - 1; r++) outline[(size_t)r * W + 1] = outline[(size_t)r * W + W - 2] = 255;
        }
        timed("image_scissor_fill", [&] { image_scissor_fill(outline.data()); });

        printf("\nComputing wrapped phase...\n");
        timed("compute_wrapped_phase(0)", [] { compute_wrapped_phase(0); });             // :371
        timed("compute_wrapped_phase(1)", [] { compute_wrapped_phase(1); });             // :375
        timed("unwrap_phase(0)", [] { unwrap_phase(0); });                               // :379
        timed("unwrap_phase(1)", [] { unwrap_phase(1); });                               // :383
        printf("\nComputing correspondence...\n");
        timed("compute_c_p_map", [] { compute_c_p_map(); });                             // :388 (STEP-6)
        printf("\nTriangulating...\n");
        timed("triangulate", [] { triangulate(); });                                     // :394 (STEP-7)
        timed("save_point_cloud", [&] { save_point_cloud(t); });                         // :400 (STEP-8)
        t++;
    }
    if (n_scans > 1) {
        printf("\nMerging view %u...", t);
        timed("register_point_clouds", [&] { register_point_clouds(n_scans, tx, ty, tz, rot_step); });   // :408
    }
    scan3d_compat_shutdown();
    return 0;
}
