#!/usr/bin/env python
"""Freeze golden fixtures under tests/golden/ from the reference tree and from cv2.

Run HERE (the build container, where /root/reference is mounted read-only and cv2 4.13 is
importable); the GPU box has neither, so the outputs are committed:

  tests/golden/c1_crop.npz    320x384 crop of the reference's own captured scan (C1):
                              fringe/Gray/inverse-Gray u8 stacks for both directions, plus the
                              matching crops of the reference's stored Wrapped_phase_image.bmp
                              and Unwrapped_phase_*.bmp (the stage-3/4 golden outputs).
  tests/golden/calib_c1.json  the 8 calibration matrices (stage-7 input) + Relative_geometry
                              KAT (6/system_calibration.cpp:1489-1504 output).
  tests/golden/opencv_kat.npz known answers computed with cv2 4.13 for the OpenCV arithmetic
                              on the path: undistortPointsIter(COUNT,5), Rodrigues,
                              transpose/gemm/invert/gemm/gemm triangulation chain.
  tests/golden/pattern_kat.npz first row / column of the reference's Generated_patterns
                              (stage-1 conventions used by the synthetic generator).
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


def main():
    import cv2
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
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
