/*
 * scan3d_host.h -- C ABI of the C++ host side (no CUDA dependency): file formats of the
 * reference tree, the synthetic captured-stack generator used by the tests and the benchmark,
 * and a PLY writer for host-side point lists.
 *
 * Reference counterparts:
 *   scan3d_read_bmp8 / scan3d_load_captured_set   cvLoadImage(..., GRAYSCALE) at 3/wrapped_phase.cpp:44,
 *                                                 4/phase_unwrap.cpp:78-90,117-127
 *   scan3d_read_cv_matrix / scan3d_load_calibration   cvOpenFileStorage/cvReadByName at
 *                                                 6/system_calibration.cpp:1526-1554, 7/triangulation.cpp:152-168
 *   scan3d_synth_pattern_row                      1/pattern_generator.cpp:56-197,291-397,490-507
 *   scan3d_write_ply_points                       pcl::io::savePLYFile at 8/save_point_cloud.cpp:217
 *   scan3d_write_pcd_points                       pcl::io::savePCDFileASCII at 8/save_point_cloud.cpp:212
 *   scan3d_read_ply_points                        pcl::io::loadPLYFile at 9/register_point_clouds.cpp:66,87
 */
#ifndef SCAN3D_HOST_H
#define SCAN3D_HOST_H

#include <stdint.h>

#include "scan3d.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- files ---- */
/* 8-bit palettised or 24-bit BMP -> grey u8 [H][W] (top-down).  Pass buf = NULL to query the size. */
int scan3d_read_bmp8(const char *path, int *W, int *H, uint8_t *buf, int64_t buf_bytes);
/* cvLoadImage(path) with its default (colour) flag: buf = [H][W][3] B,G,R bytes, top row first.  8- and 24-bit
 * uncompressed BMP (8/save_point_cloud.cpp:59-66 loads Point_cloud/texture.bmp this way). */
int scan3d_read_bmp_bgr(const char *path, int *W, int *H, uint8_t *buf, int64_t buf_bytes);
int scan3d_write_bmp8(const char *path, int W, int H, const uint8_t *buf);
/* OpenCV XML <name type_id="opencv-matrix"> with <dt>d</dt>: reads rows*cols doubles. */
int scan3d_read_cv_matrix(const char *path, const char *name, int rows, int cols, double *out);
/* root = ".../M_tech_project_console": reads the 8 matrices load_matrices() reads. */
int scan3d_load_calibration(const char *root, scan3d_calib *cal);
/* root = ".../M_tech_project_console": fills the stack in scan3d_reconstruct() order from
 * Captured_patterns/{Fringe_patterns,Coded_patterns/Gray_coded}/{Vertical,Horizontal}/Undistorted. */
int scan3d_load_captured_set(const char *root, const scan3d_config *cfg, uint8_t *stack);
int scan3d_write_ply_points(const char *path, const float *xyz, const uint8_t *rgb, int64_t n, int binary);
/* ASCII PCD v0.7, FIELDS x y z rgb (rgb packed into a float as PCL 1.6's PointXYZRGB does) */
int scan3d_write_pcd_points(const char *path, const float *xyz, const uint8_t *rgb, int64_t n);
/* pcl::io::loadPLYFile's job in register_point_clouds (9/register_point_clouds.cpp:66,87): positions
 * (and uchar red/green/blue when present) of an ascii or binary_little_endian PLY.  xyz = NULL only
 * queries *n_out. */
int scan3d_read_ply_points(const char *path, float *xyz, uint8_t *rgb, int64_t capacity, int64_t *n_out);
const char *scan3d_host_last_error(void);

/* ---- synthetic captures ---- */
typedef struct {
    double sphere_c[3];     /* world mm */
    double sphere_r;        /* <= 0: no sphere */
    double plane_z;         /* background plane Z = plane_z in the world frame */
    double albedo_lo, albedo_hi;
    double ambient_max;     /* DN */
    double noise_sigma;     /* DN, Gaussian */
    double roi_fraction;    /* area fraction of the elliptical ROI (0.75) */
    int32_t true_pi;        /* 0: fringes use the reference's Pi = 22/7; 1: 3.14159... */
    int32_t projector_pixelated; /* 1: sample the pattern at floor(px) like a real projector */
    uint64_t seed;
} scan3d_synth_params;

void scan3d_synth_default_params(scan3d_synth_params *p);
/* Renders the captured stack ([NF][H][W], scan3d_reconstruct order) and the ROI for the rows
 * [cfg->row0, cfg->row0 + cfg->H) of a cfg->H_total frame.  truth_xyz ([H][W][3] f32, optional)
 * receives the scene point seen by each pixel (NaN when the ray hits nothing). */
int scan3d_synth_stack(const scan3d_config *cfg, const scan3d_calib *cal, const scan3d_synth_params *p,
                       uint8_t *stack, uint8_t *roi_full, float *truth_xyz, int threads);
/* One row (vertical patterns) / column (horizontal) of the PROJECTED pattern exactly as
 * 1/pattern_generator.cpp writes it: kind 0 = fringe k of an N-step set, 1 = Gray bit k of M,
 * 2 = inverse Gray bit k.  length = projector width or height. */
int scan3d_synth_pattern_row(int kind, int n_or_m, int fw, int k, int length, uint8_t *out);
/* Reference calibration scaled to another resolution. */
void scan3d_scale_calibration(const scan3d_calib *in, double cam_scale, double proj_scale, scan3d_calib *out);

#ifdef __cplusplus
}
#endif
#endif
