#!/usr/bin/env python
"""Freeze golden fixtures under tests/golden/ from the reference tree and from cv2.

Run HERE (the build container, where /root/reference is mounted read-only and cv2 4.13 is
importable); the GPU box has neither, so the outputs are committed:

  tests/golden/c1_crop.npz    320x384 crop of the reference's own captured scan (C1):
                              fringe/Gray/inverse-Gray u8 stacks for both directions, plus the
                              matching crops of the reference's stored Wrapped_phase_image.bmp
                              and Unwrapped_phase_*.bmp (the stage-3/4 golden outputs).
  tests/golden/c1_full.npz    the same scan uncropped (1600x1200, 28 frames + the four stored images), for the
                              full-frame GPU test; `make_golden.py c1full` makes only this.
  tests/golden/calib_c1.json  the 8 calibration matrices (stage-7 input) + Relative_geometry
                              KAT (6/system_calibration.cpp:1489-1504 output).
  tests/golden/opencv_kat.npz known answers computed with cv2 4.13 for the OpenCV arithmetic
                              on the path: undistortPointsIter(COUNT,5), Rodrigues,
                              transpose/gemm/invert/gemm/gemm triangulation chain.
  tests/golden/pattern_kat.npz first row / column of the reference's Generated_patterns
                              (stage-1 conventions used by the synthetic generator).
  tests/golden/f4_kat.npz     either side of the path: cv2.undistort of small random
                              images (pins the cvUndistort2 restatement), the float cv2.gemm chain of
                              register_point_clouds, and the per-row extent of the reference's stored
                              i1.jpg (image_scissor's filled outline).  `make_golden.py f4` makes only this.
"""
import json
import os
import re
import sys

import numpy as np

REF = "/root/reference/M_tech_project_console/"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

CROP_R0, CROP_R1 = 184, 504   # rows  (valid bbox of the scan: rows 201-824)
CROP_C0, CROP_C1 = 480, 864   # cols  (valid bbox: cols 492-1233)


def read_bmp8(path):
    """8-bit palettised (identity grey palette), bottom-up BMP -> [H][W] u8."""
    b = open(path, "rb").read()
    off = int.from_bytes(b[10:14], "little")
    W = int.from_bytes(b[18:22], "little", signed=True)
    H = int.from_bytes(b[22:26], "little", signed=True)
    assert int.from_bytes(b[28:30], "little") == 8, path
    pal = np.frombuffer(b[54:54 + 1024], np.uint8).reshape(256, 4)
    assert (pal[:, 0] == np.arange(256)).all() and (pal[:, 2] == np.arange(256)).all()
    stride = (W + 3) & ~3
    a = np.frombuffer(b[off:off + stride * abs(H)], np.uint8).reshape(abs(H), stride)[:, :W]
    return a[::-1].copy() if H > 0 else a.copy()


def read_xml(path):
    t = open(path).read()
    rows = int(re.search(r"<rows>(\d+)", t).group(1))
    cols = int(re.search(r"<cols>(\d+)", t).group(1))
    d = re.search(r"<data>(.*?)</data>", t, re.S).group(1).split()
    return np.array([float(x) for x in d]).reshape(rows, cols)


def load_c1_direction(name, M):
    base = REF + "Captured_patterns/"
    fr = np.stack([read_bmp8(f"{base}Fringe_patterns/{name}/Undistorted/Gray_captured_image_{i}.bmp")
                   for i in range(3)])
    g = np.stack([read_bmp8(f"{base}Coded_patterns/Gray_coded/{name}/Undistorted/Gray_captured_image_{i}.bmp")
                  for i in range(M)])
    gi = np.stack([read_bmp8(f"{base}Coded_patterns/Gray_coded/{name}/Undistorted/inverse_Gray_captured_image_{i}.bmp")
                   for i in range(M)])
    gw = read_bmp8(REF + f"Wrapped_phase_images/{name}/Wrapped_phase_image.bmp")
    gu = read_bmp8(REF + f"Unwrapped_phase_images/Gray_coded/{name}/Unwrapped_phase_{name.lower()}.bmp")
    return fr, g, gi, gw, gu


def load_calib():
    c = {}
    c["Kc"] = read_xml(REF + "Camera_calibration/Matrices/cam_intrinsic_mat.xml")
    c["dc"] = read_xml(REF + "Camera_calibration/Matrices/cam_distortion_vect.xml").ravel()
    c["Kp"] = read_xml(REF + "Projector_calibration/Matrices/proj_intrinsic_mat.xml")
    c["dp"] = read_xml(REF + "Projector_calibration/Matrices/proj_distortion_vect.xml").ravel()
    t = REF + "Triangulation/"
    c["rc"] = read_xml(t + "Camera_extrinsic_parametrs/world_to_cam_rot_vect.xml").ravel()
    c["tc"] = read_xml(t + "Camera_extrinsic_parametrs/world_to_cam_trans_vect.xml").ravel()
    c["rp"] = read_xml(t + "Projector_extrinsic_parametrs/world_to_proj_rot_vect.xml").ravel()
    c["tp"] = read_xml(t + "Projector_extrinsic_parametrs/world_to_proj_trans_vect.xml").ravel()
    c["rel_R"] = read_xml(t + "Relative_geometry/proj_cam_rot_mat.xml")
    c["rel_T"] = read_xml(t + "Relative_geometry/proj_cam_trans_vect.xml").ravel()
    return c


def make_f4():
    import cv2
    cal = load_calib()
    rng = np.random.default_rng(20261018)
    kat = {"cv2_version": np.array(cv2.__version__)}
    cases = [(320, 240, cal["Kc"] * np.array([[0.2], [0.2], [1.0]]), cal["dc"]),                    # C1 calibration, scaled
             (256, 96, cal["Kc"] * np.array([[0.16], [0.16], [1.0]]), cal["dc"] * 6),
             (200, 150, np.array([[180.0, 0.4, 97.3], [0, 184.0, 80.1], [0, 0, 1]]), np.array([-0.31, 0.12, -0.002, 0.001, -0.03])),
             (4100, 5, np.array([[3600.0, 0, 2050.5], [0, 3610.0, 2.0], [0, 0, 1]]), np.array([0.0813, -0.1102, 0.0013, -0.0007, 0.021])),
             (1280 // 4, 720 // 4, cal["Kp"] * np.array([[0.25], [0.25], [1.0]]), cal["dp"])]         # projector: no distortion
    kat["n_undistort"] = np.array(len(cases))
    for i, (W, H, K, d) in enumerate(cases):
        src = rng.integers(0, 256, (H, W), dtype=np.uint8)
        kat[f"und_K{i}"], kat[f"und_d{i}"], kat[f"und_src{i}"] = K, d, src
        kat[f"und_dst{i}"] = cv2.undistort(src, K, d)
    # 9/register_point_clouds.cpp:93-137 with cv2.gemm standing in for cvMatMul (same unrolled float path)
    theta = np.float32(36.0)
    a = float(theta) * 22.0 / 7.0 / 180.0      # the float promotes to double, as in C
    R = np.zeros((4, 4), np.float32)
    R[0, 0] = np.cos(a); R[0, 2] = -1.0 * np.sin(a); R[2, 0] = np.sin(a); R[2, 2] = np.cos(a)
    R[1, 1] = 1.0; R[3, 3] = 1.0
    pts = (rng.normal(size=(2048, 3)) * [120, 80, 300] + [10, -20, 900]).astype(np.float32)
    t = np.array([12.5, -3.25, 903.1], np.float32)
    out = np.empty_like(pts)
    for k in range(len(pts)):
        p = np.array([[pts[k, 0]], [pts[k, 1]], [pts[k, 2]], [1.0]], np.float32)
        p[0, 0] -= t[0]; p[1, 0] -= t[1]; p[2, 0] -= t[2]
        q = cv2.gemm(R, p, 1.0, None, 0.0)
        q[0, 0] += t[0]; q[1, 0] += t[1]; q[2, 0] += t[2]
        out[k] = q[:3, 0]
    kat["reg_theta"], kat["reg_t"], kat["reg_R"], kat["reg_pts"], kat["reg_out"] = np.array(theta), t, R, pts, out
    # image_scissor's filled outline as the reference stored it (JPEG -> threshold)
    m = cv2.imread(REF + "i1.jpg", 0) > 127
    cols = np.arange(m.shape[1])
    kat["i1_shape"] = np.array(m.shape)
    kat["i1_count"] = m.sum(1).astype(np.int32)
    kat["i1_first"] = np.where(m.any(1), np.where(m, cols, m.shape[1]).min(1), 0).astype(np.int32)
    kat["i1_last"] = np.where(m.any(1), np.where(m, cols, -1).max(1), -1).astype(np.int32)
    # cvLoadImage(..., CV_LOAD_IMAGE_GRAYSCALE) of a colour BMP (3/wrapped_phase.cpp:44, 4/phase_unwrap.cpp:78-90): the
    # file bytes of a small 24-bit BMP (odd width: padded rows) and the grey image cv2.imread returns for it
    import tempfile
    col = rng.integers(0, 256, (9, 33, 3), dtype=np.uint8)
    col[0, :16] = np.array([[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 1, 1], [254, 254, 254], [128, 127, 129]] * 2, np.uint8)
    path = os.path.join(tempfile.mkdtemp(), "colour.bmp")
    assert cv2.imwrite(path, col)
    kat["bmp24_bytes"] = np.frombuffer(open(path, "rb").read(), np.uint8)
    kat["bmp24_grey"] = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    np.savez_compressed(os.path.join(OUT, "f4_kat.npz"), **kat)


def make_c1_full():
    """tests/golden/c1_full.npz: the reference's whole 1600x1200 captured scan (28 grey frames, both directions)
    and its stored stage-3/4 images, losslessly packed -- the GPU box has no /root/reference."""
    d = {"config": np.array([3, 6, 5, 32, 32, 40, 23, 1280, 720])}  # N,M_v,M_h,fw_v,fw_h,codes_v,codes_h,PW,PH
    for name, M, key in (("Vertical", 6, "v"), ("Horizontal", 5, "h")):
        fr, g, gi, gw, gu = load_c1_direction(name, M)
        d[f"fringe_{key}"], d[f"gray_{key}"], d[f"inv_{key}"] = fr, g, gi
        d[f"golden_wrapped_{key}"], d[f"golden_unwrapped_{key}"] = gw, gu
    np.savez_compressed(os.path.join(OUT, "c1_full.npz"), **d)


def main():
    import cv2
    if len(sys.argv) > 1 and sys.argv[1] == "c1full":
        make_c1_full()
        print("c1_full.npz", os.path.getsize(os.path.join(OUT, "c1_full.npz")))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "f4":
        make_f4()
        print("f4_kat.npz", os.path.getsize(os.path.join(OUT, "f4_kat.npz")))
        return
    os.makedirs(OUT, exist_ok=True)
    sl = (slice(CROP_R0, CROP_R1), slice(CROP_C0, CROP_C1))

    # ---- C1 crop
    d = {"crop": np.array([CROP_R0, CROP_R1, CROP_C0, CROP_C1]),
         "config": np.array([3, 6, 5, 32, 32, 40, 23, 1280, 720])}  # N,M_v,M_h,fw_v,fw_h,codes_v,codes_h,PW,PH
    for name, M, key in (("Vertical", 6, "v"), ("Horizontal", 5, "h")):
        fr, g, gi, gw, gu = load_c1_direction(name, M)
        d[f"fringe_{key}"] = fr[(slice(None),) + sl]
        d[f"gray_{key}"] = g[(slice(None),) + sl]
        d[f"inv_{key}"] = gi[(slice(None),) + sl]
        d[f"golden_wrapped_{key}"] = gw[sl]
        d[f"golden_unwrapped_{key}"] = gu[sl]
    np.savez_compressed(os.path.join(OUT, "c1_crop.npz"), **d)

    # ---- calibration
    cal = load_calib()
    json.dump({k: [float(x) for x in np.asarray(v).ravel()] for k, v in cal.items()},
              open(os.path.join(OUT, "calib_c1.json"), "w"), indent=1)

    # ---- OpenCV known answers (cv2 4.13; the reference linked 2.4.0/2.4.1, same algorithms)
    rng = np.random.default_rng(20261017)
    kat = {"cv2_version": np.array(cv2.__version__)}
    pts = np.stack([rng.uniform(0, 1600, 4096), rng.uniform(0, 1200, 4096)], 1)
    kat["und_pts"] = pts
    dists = np.stack([cal["dc"], np.array([0.0813, -0.1102, 0.0013, -0.0007, 0.021]),
                      cal["dp"], np.array([-0.31, 0.12, -0.002, 0.001, -0.03])])
    kat["und_dists"] = dists
    kat["und_K"] = cal["Kc"]
    crit = (cv2.TERM_CRITERIA_COUNT, 5, 0)
    kat["und_out"] = np.stack([
        cv2.undistortPointsIter(pts.reshape(-1, 1, 2), cal["Kc"], dd, None, None, crit).reshape(-1, 2)
        for dd in dists])
    rvecs = np.vstack([cal["rc"], cal["rp"], rng.normal(size=(13, 3)), [[1e-20, 0, 0]]])
    kat["rod_in"] = rvecs
    kat["rod_out"] = np.stack([cv2.Rodrigues(r.reshape(3, 1))[0] for r in rvecs])
    Ac = cv2.gemm(cal["Kc"], np.hstack([cv2.Rodrigues(cal["rc"])[0], cal["tc"].reshape(3, 1)]), 1, None, 0)
    Ap = cv2.gemm(cal["Kp"], np.hstack([cv2.Rodrigues(cal["rp"])[0], cal["tp"].reshape(3, 1)]), 1, None, 0)
    kat["A_cam"], kat["A_proj"] = Ac, Ap
    n = 4096
    uv = np.stack([rng.uniform(0, 1600, n), rng.uniform(0, 1200, n),
                   rng.uniform(0, 1280, n), rng.uniform(0, 720, n)], 1)
    V = np.empty((n, 3))
    for i, (uc, vc, up, vp) in enumerate(uv):
        P = np.array([Ac[0, :3] - uc * Ac[2, :3], Ac[1, :3] - vc * Ac[2, :3],
                      Ap[0, :3] - up * Ap[2, :3], Ap[1, :3] - vp * Ap[2, :3]])
        F = np.array([[Ac[2, 3] * uc - Ac[0, 3]], [Ac[2, 3] * vc - Ac[1, 3]],
                      [Ap[2, 3] * up - Ap[0, 3]], [Ap[2, 3] * vp - Ap[1, 3]]])
        Pt = cv2.transpose(P)
        I1 = cv2.gemm(Pt, P, 1, None, 0)
        _, I1i = cv2.invert(I1)
        I2 = cv2.gemm(I1i, Pt, 1, None, 0)
        V[i] = cv2.gemm(I2, F, 1, None, 0).ravel()
    kat["tri_in"], kat["tri_out"] = uv, V
    np.savez_compressed(os.path.join(OUT, "opencv_kat.npz"), **kat)

    # ---- stage-1 pattern conventions
    g = REF + "Generated_patterns/"
    pk = {}
    pk["fringe_v_row0"] = np.stack([read_bmp8(g + f"Fringe_patterns/Vertical/Pattern_{i}.bmp")[0] for i in range(3)])
    pk["fringe_h_col0"] = np.stack([read_bmp8(g + f"Fringe_patterns/Horizontal/Pattern_{i}.bmp")[:, 0] for i in range(3)])
    gc = g + "Coded_patterns/Gray_coded/"
    for pre, key in (("", "gray"), ("inverse_", "inv")):
        pk[f"{key}_v_row0"] = np.stack([read_bmp8(gc + f"Vertical/{pre}Pattern_{i}.bmp")[0] for i in range(6)])
        pk[f"{key}_h_col0"] = np.stack([read_bmp8(gc + f"Horizontal/{pre}Pattern_{i}.bmp")[:, 0] for i in range(5)])
    # Pattern_3 / Pattern_4 are left over from earlier 4-step (fw 16) and 5-step (fw 32) runs on a
    # 1024x768 projector: they pin the 4- and 5-step expressions (k = 3 and k = 4)
    pk["fringe4_k3_v_row0"] = read_bmp8(g + "Fringe_patterns/Vertical/Pattern_3.bmp")[0]
    pk["fringe4_k3_h_col0"] = read_bmp8(g + "Fringe_patterns/Horizontal/Pattern_3.bmp")[:, 0]
    pk["fringe5_k4_v_row0"] = read_bmp8(g + "Fringe_patterns/Vertical/Pattern_4.bmp")[0]
    pk["fringe5_k4_h_col0"] = read_bmp8(g + "Fringe_patterns/Horizontal/Pattern_4.bmp")[:, 0]
    np.savez_compressed(os.path.join(OUT, "pattern_kat.npz"), **pk)
    make_f4()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
