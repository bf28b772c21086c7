/*
 * scan3d_compat.h -- the reference's own stage interface (C++ linkage, same names, same
 * argument meaning, same globals) implemented on top of the C ABI in scan3d.h.
 *
 * A maintainer of pranavkantgaur/3dscan links libscan3d_compat.so instead of compiling
 * 3/wrapped_phase.cpp, 4/phase_unwrap.cpp, 5/compute_correspondance.cpp, 7/triangulation.cpp and
 * 8/save_point_cloud.cpp; main() (M_tech_project_console/m_tech_project_console.cpp:372-401)
 * keeps calling
 *     compute_wrapped_phase(0); compute_wrapped_phase(1); unwrap_phase(0); unwrap_phase(1);
 *     compute_c_p_map(); triangulate(); save_point_cloud(t);
 * exactly as declared in PROJECT_GLOBAL/intermodule_dependencies.h:10-25.
 *
 * Differences from the reference, all forced by its compile-time constants:
 *   - Camera_imagewidth/height and Projector_imagewidth/height (global_cv.h:49-53) are runtime
 *     values given to scan3d_compat_init(); the global planes are therefore flat arrays indexed
 *     [col * Camera_imageheight + row] (the reference's [col][row] layout) instead of
 *     pointer-to-array types.
 *   - the data root is not hard-coded to /home/pranav/Desktop/M_tech_project_console.
 *   - stages run on the GPU; the globals below are filled after every stage so that code reading
 *     them keeps working.  Set scan3d_compat_export = 0 to skip that copy.
 *   - fatal conditions print a message and exit(EXIT_FAILURE) (the reference printf()s and
 *     exit(0)s, or crashes on a NULL image).
 */
#ifndef SCAN3D_COMPAT_H
#define SCAN3D_COMPAT_H

#include "scan3d.h"

/* ---- PROJECT_GLOBAL/common_variables.h:6-24 ---- */
extern int number_of_codes_vertical, number_of_codes_horizontal;
extern int number_of_patterns_binary_vertical, number_of_patterns_binary_horizontal;
extern int number_of_patterns_fringe;
extern int fringe_width_pixels_vertical, fringe_width_pixels_horizontal;
extern int Camera_imagewidth, Camera_imageheight, Projector_imagewidth, Projector_imageheight;

/* planes in the reference's [col][row] order: element (col,row) at [col*Camera_imageheight+row] */
extern int *selected_region;                 /* INPUT: 1 = pixel selected (m_tech_project_console.cpp:183-229) */
extern int *valid_map_vertical, *valid_map_horizontal, *valid_map;
extern int *code_vertical, *code_horizontal;
extern float *wrapped_phi_vertical, *wrapped_phi_horizontal;
extern float *unwrapped_phi_vertical, *unwrapped_phi_horizontal;
extern long int (*c_p_map)[2];               /* [row*Camera_imagewidth+col][2], like the reference */
extern double *intersection_points;          /* [(col*Camera_imageheight+row)*3 + k] */
extern int scan3d_compat_export;             /* 1 (default): refresh the globals after each stage */

/* root = the "M_tech_project_console" directory (captured patterns, calibration XML, Point_cloud/) */
int scan3d_compat_init(const char *root, int cam_w, int cam_h, int proj_w, int proj_h, int device);
void scan3d_compat_shutdown();
scan3d_ctx *scan3d_compat_ctx();

/* ---- PROJECT_GLOBAL/intermodule_dependencies.h ---- */
void generate_pattern();                      /* 1/pattern_generator.cpp:513: writes <root>/Generated_patterns/... (:436-479;
                                                 the plain binary-coded set, which no later stage reads, is not written) */
void load_matrices();                         /* 6/system_calibration.cpp:1526 */
void compute_wrapped_phase(int pattern_type); /* 3/wrapped_phase.cpp:402 */
void unwrap_phase(int pattern_type);          /* 4/phase_unwrap.cpp:367 (declared int, defined void) */
void compute_c_p_map();                       /* 5/compute_correspondance.cpp:630 */
void triangulate();                           /* 7/triangulation.cpp:1444 */
void save_point_cloud(unsigned cloud_index);  /* 8/save_point_cloud.cpp:26 */
/* 9/register_point_clouds.cpp:22 (called at m_tech_project_console.cpp:408): reads
 * <root>/Point_cloud/point_cloud_<i>.ply for i < num_point_clouds, rotates cloud i by i*rot_step degrees
 * about Y through (tx,ty,tz) on the GPU, writes <root>/Point_cloud/registered_point_cloud.ply */
void register_point_clouds(unsigned num_point_clouds, float tx, float ty, float tz, float rot_step);
/* the non-interactive half of image_scissor() (m_tech_project_console.cpp:183-231): from the lasso
 * outline (internal_image, row-major u8 [Camera_imageheight][Camera_imagewidth], non-zero = drawn) to
 * the selected_region global; saves the filled outline as <root>/i1.bmp (the reference writes i1.jpg) */
void image_scissor_fill(const unsigned char *internal_image);
/* cvUndistort2(cap, undist_cap, K, d) of the capture loop (2/project_pattern.cpp:220,234,372-427) for one
 * 8-bit image: device_kind 0 = camera frame, 1 = projector pattern; needs load_matrices() */
void undistort_capture(const unsigned char *cap, unsigned char *undist_cap, int device_kind);
/* the whole sequence above in one fused pass (no reference counterpart) */
void reconstruct_scan(unsigned cloud_index);

#endif
