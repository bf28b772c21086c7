"""ctypes binding of the CPU oracle (oracle/libscan3d_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package (3dscan_b200/).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libscan3d_oracle.so")
    deps = [os.path.join(_HERE, f) for f in ("scan3d_oracle.c", "scan3d_oracle_f4.c", "scan3d_oracle.h")]
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(f) > os.path.getmtime(so) for f in deps)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


class Config(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("W", "H", "PW", "PH", "N", "M_v", "M_h", "fw_v", "fw_h", "dirs")]


class Calib(C.Structure):
    _fields_ = [("Kc", C.c_double * 9), ("dc", C.c_double * 5), ("Kp", C.c_double * 9),
                ("dp", C.c_double * 5), ("rc", C.c_double * 3), ("tc", C.c_double * 3),
                ("rp", C.c_double * 3), ("tp", C.c_double * 3)]


class Outputs(C.Structure):
    _fields_ = [("wrapped_v", C.c_void_p), ("wrapped_h", C.c_void_p),
                ("unwrapped_v", C.c_void_p), ("unwrapped_h", C.c_void_p),
                ("code_v", C.c_void_p), ("code_h", C.c_void_p),
                ("valid_v", C.c_void_p), ("valid_h", C.c_void_p), ("valid", C.c_void_p),
                ("cpmap", C.c_void_p), ("xyz", C.c_void_p), ("pts", C.c_void_p),
                ("pix", C.c_void_p), ("count", C.c_int64)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.o3d_compact.restype = C.c_int64
        _LIB.o3d_max_threads.restype = C.c_int
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _d(a, n):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    assert a.size == n, (a.size, n)
    return a


def max_threads():
    return int(lib().o3d_max_threads())


def make_calib(Kc, dc, Kp, dp, rc, tc, rp, tp):
    cal = Calib()
    for name, val, n in (("Kc", Kc, 9), ("dc", dc, 5), ("Kp", Kp, 9), ("dp", dp, 5),
                         ("rc", rc, 3), ("tc", tc, 3), ("rp", rp, 3), ("tp", tp, 3)):
        arr = _d(val, n)
        getattr(cal, name)[:] = arr.tolist()
    return cal


# ---- stage-level wrappers (all planes row-major [H][W]) ---------------------------------

def check_roi(roi):
    H, W = roi.shape
    roi = np.ascontiguousarray(roi, np.uint8)
    valid = np.empty((H, W), np.int32)
    lib().o3d_check_roi(_p(roi), W, H, _p(valid))
    return valid


def check_I_mod_criteria(fringe, roi):
    """The reference's commented-out modulation criterion (3/wrapped_phase.cpp:84-104, 3-step)."""
    H, W = roi.shape
    fringe = np.ascontiguousarray(fringe, np.uint8)
    assert fringe.shape == (3, H, W)
    roi = np.ascontiguousarray(roi, np.uint8)
    valid = np.empty((H, W), np.int32)
    lib().o3d_check_I_mod_criteria(_p(fringe), _p(roi), W, H, _p(valid))
    return valid


def wrapped_phase(fringe, valid, threads=1, want_dbg=True):
    N, H, W = fringe.shape
    fringe = np.ascontiguousarray(fringe, np.uint8)
    valid = np.ascontiguousarray(valid, np.int32)
    wrapped = np.zeros((H, W), np.float32)
    dbg = np.zeros((H, W), np.uint8) if want_dbg else None
    lib().o3d_wrapped_phase(_p(fringe), N, W, H, _p(valid), _p(wrapped), _p(dbg), threads)
    return wrapped, dbg


def mask_recurrence(valid, dbg=None):
    valid = np.ascontiguousarray(valid, np.int32).copy()
    H, W = valid.shape
    lib().o3d_mask_recurrence(_p(valid), W, H, _p(dbg))
    return valid


def mask_closed_form(valid0):
    valid0 = np.ascontiguousarray(valid0, np.int32)
    H, W = valid0.shape
    out = np.empty((H, W), np.int32)
    lib().o3d_mask_closed_form(_p(valid0), W, H, _p(out))
    return out


def decode_gray(gray, inv, valid, threads=1):
    M, H, W = gray.shape
    gray = np.ascontiguousarray(gray, np.uint8)
    inv = np.ascontiguousarray(inv, np.uint8)
    valid = np.ascontiguousarray(valid, np.int32)
    code = np.empty((H, W), np.int32)
    lib().o3d_decode_gray(_p(gray), _p(inv), M, W, H, _p(valid), _p(code), threads)
    return code


def unwrap(direction, wrapped, code, valid, threads=1):
    """Returns (wrapped_after_plus_pi, unwrapped)."""
    wrapped = np.ascontiguousarray(wrapped, np.float32).copy()
    H, W = wrapped.shape
    code = np.ascontiguousarray(code, np.int32)
    valid = np.ascontiguousarray(valid, np.int32)
    unw = np.zeros((H, W), np.float32)
    lib().o3d_unwrap(direction, _p(wrapped), _p(code), _p(valid), W, H, _p(unw), threads)
    return wrapped, unw


def unwrapped_image(unw, valid, number_of_codes):
    H, W = unw.shape
    img = np.zeros((H, W), np.uint8)
    lib().o3d_unwrapped_image(_p(np.ascontiguousarray(unw, np.float32)),
                              _p(np.ascontiguousarray(valid, np.int32)), W, H,
                              number_of_codes, _p(img))
    return img


def compute_c_p_map(unw_v, unw_h, valid_v, valid_h, fw_v, fw_h, PW, PH, threads=1):
    H, W = unw_v.shape
    cp = np.zeros((H * W, 2), np.int64)
    valid = np.empty((H, W), np.int32)
    lib().o3d_compute_c_p_map(_p(np.ascontiguousarray(unw_v, np.float32)),
                              _p(np.ascontiguousarray(unw_h, np.float32)),
                              _p(np.ascontiguousarray(valid_v, np.int32)),
                              _p(np.ascontiguousarray(valid_h, np.int32)),
                              fw_v, fw_h, PW, PH, W, H, _p(cp), _p(valid), threads)
    return cp, valid


def rodrigues(rvec):
    R = np.empty(9, np.float64)
    lib().o3d_rodrigues(_p(_d(rvec, 3)), _p(R))
    return R.reshape(3, 3)


def compose_relative(rc, tc, rp, tp):
    R = np.empty(9, np.float64)
    T = np.empty(3, np.float64)
    lib().o3d_compose_relative(_p(_d(rc, 3)), _p(_d(tc, 3)), _p(_d(rp, 3)), _p(_d(tp, 3)),
                               _p(R), _p(T))
    return R.reshape(3, 3), T


def undistort_points(xy, K, d):
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    out = np.empty_like(xy)
    lib().o3d_undistort_points(_p(xy), xy.shape[0], _p(_d(K, 9)), _p(_d(d, 5)), _p(out))
    return out


def undistort_lut(K, d, W, H, threads=1):
    lut = np.empty((2, H, W), np.float64)
    lib().o3d_undistort_lut(_p(_d(K, 9)), _p(_d(d, 5)), W, H, _p(lut), threads)
    return lut


def compute_A(K, rvec, tvec):
    A = np.empty(12, np.float64)
    lib().o3d_compute_A(_p(_d(K, 9)), _p(_d(rvec, 3)), _p(_d(tvec, 3)), _p(A))
    return A.reshape(3, 4)


def triangulate_point(A_cam, A_proj, uc, vc, up, vp):
    out = np.empty(3, np.float64)
    lib().o3d_triangulate_point(_p(_d(A_cam, 12)), _p(_d(A_proj, 12)), C.c_double(uc),
                                C.c_double(vc), C.c_double(up), C.c_double(vp), _p(out))
    return out


def triangulate(A_cam, A_proj, cam_lut, proj_lut, cpmap, valid, PW, PH, threads=1):
    H, W = valid.shape
    xyz = np.zeros((H, W, 3), np.float64)
    lib().o3d_triangulate(_p(_d(A_cam, 12)), _p(_d(A_proj, 12)),
                          _p(np.ascontiguousarray(cam_lut, np.float64)),
                          _p(np.ascontiguousarray(proj_lut, np.float64)),
                          _p(np.ascontiguousarray(cpmap, np.int64)),
                          _p(np.ascontiguousarray(valid, np.int32)), W, H, PW, PH, _p(xyz),
                          threads)
    return xyz


def compact(xyz, valid, texture=None):
    H, W = valid.shape
    xyz = np.ascontiguousarray(xyz, np.float64)
    valid = np.ascontiguousarray(valid, np.int32)
    n = int(lib().o3d_compact(_p(xyz), _p(valid), None, W, H, None, None, None))
    pts = np.empty((n, 3), np.float32)
    rgb = np.empty((n, 3), np.uint8)
    pix = np.empty(n, np.uint32)
    tex = None if texture is None else np.ascontiguousarray(texture, np.uint8)
    lib().o3d_compact(_p(xyz), _p(valid), _p(tex), W, H, _p(pts), _p(rgb), _p(pix))
    return pts, rgb, pix


class Result:
    pass


def reconstruct(cfg, cal, fringe_v, gray_v, inv_v, fringe_h, gray_h, inv_h, roi, threads=1,
                want_xyz=True, modulation=False, strict=False, colrow=False):
    """cfg: dict with W,H,PW,PH,N,M_v,M_h,fw_v,fw_h,dirs.  Returns a Result of numpy planes.
    colrow=True: the reference-faithful leg (o3d_reconstruct_colrow: the reference's [col][row] plane layout, one
    thread); r.seconds is the time of the C call alone, the planes come back transposed to [H][W] for comparison."""
    c = Config(**{k: int(cfg[k]) for k in
                  ("W", "H", "PW", "PH", "N", "M_v", "M_h", "fw_v", "fw_h", "dirs")})
    H, W = c.H, c.W
    r = Result()
    r.wrapped_v = np.empty((H, W), np.float32)
    r.unwrapped_v = np.empty((H, W), np.float32)
    r.code_v = np.empty((H, W), np.int32)
    r.valid_v = np.empty((H, W), np.int32)
    two = c.dirs == 2
    r.wrapped_h = np.empty((H, W), np.float32) if two else None
    r.unwrapped_h = np.empty((H, W), np.float32) if two else None
    r.code_h = np.empty((H, W), np.int32) if two else None
    r.valid_h = np.empty((H, W), np.int32) if two else None
    r.valid = np.empty((H, W), np.int32) if two else None
    r.cpmap = np.empty((H * W, 2), np.int64) if two else None
    r.xyz = np.empty((H, W, 3), np.float64) if two else None
    r.pts = np.empty((H * W, 3), np.float32) if two else None
    r.pix = np.empty(H * W, np.uint32) if two else None
    o = Outputs(_p(r.wrapped_v), _p(r.wrapped_h), _p(r.unwrapped_v), _p(r.unwrapped_h),
                _p(r.code_v), _p(r.code_h), _p(r.valid_v), _p(r.valid_h), _p(r.valid),
                _p(r.cpmap), _p(r.xyz), _p(r.pts), _p(r.pix), 0)
    u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
    keep = [u8(x) for x in (fringe_v, gray_v, inv_v, fringe_h, gray_h, inv_h, roi)]
    import time as _time
    flags = int(bool(modulation)) | (2 if strict else 0)
    t0 = _time.perf_counter()
    if colrow:
        lib().o3d_reconstruct_colrow(C.byref(c), C.byref(cal), *[_p(k) for k in keep], flags, C.byref(o))
    else:
        lib().o3d_reconstruct_ex(C.byref(c), C.byref(cal), *[_p(k) for k in keep], flags, C.byref(o), int(threads))
    r.seconds = _time.perf_counter() - t0
    if colrow:      # [W][H] -> [H][W]
        for name in ("wrapped_v", "unwrapped_v", "code_v", "valid_v", "wrapped_h", "unwrapped_h", "code_h", "valid_h", "valid"):
            a = getattr(r, name)
            if a is not None:
                setattr(r, name, np.ascontiguousarray(a.reshape(W, H).T))
        if two:
            r.xyz = np.ascontiguousarray(r.xyz.reshape(W, H, 3).transpose(1, 0, 2))
    r.count = int(o.count)
    if two:
        r.pts = r.pts[:r.count]
        r.pix = r.pix[:r.count]
    return r


# ---- either side of the path (scan3d_oracle_f4.c) -------------------------------------------

def undistort_map(K, d, W, H):
    """cv::undistort's fixed-point map: (xy int16 [H][W][2], frac uint16 [H][W])."""
    mxy = np.empty((H, W, 2), np.int16)
    mf = np.empty((H, W), np.uint16)
    lib().o3d_undistort_map(_p(_d(K, 9)), _p(_d(d, 5)), W, H, _p(mxy), _p(mf))
    return mxy, mf


def undistort_frames(frames, K, d):
    """cvUndistort2 on [n][H][W] u8 frames."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n, H, W = frames.shape
    out = np.empty_like(frames)
    lib().o3d_undistort_frames(_p(frames), n, W, H, _p(_d(K, 9)), _p(_d(d, 5)), _p(out))
    return out


def roi_fill(outline):
    """image_scissor's fill: returns (roi u8 [H][W], outline after the in-place fill)."""
    o = np.ascontiguousarray(outline, np.uint8).copy()
    H, W = o.shape
    roi = np.empty((H, W), np.uint8)
    lib().o3d_roi_fill(_p(o), W, H, _p(roi))
    return roi, o


def register_rotation(theta_deg):
    R = np.empty(16, np.float32)
    lib().o3d_register_rotation(C.c_float(theta_deg), _p(R))
    return R.reshape(4, 4)


def register_points(xyz, theta_deg, tx, ty, tz):
    xyz = np.ascontiguousarray(xyz, np.float32).copy()
    lib().o3d_register_points(_p(xyz), C.c_int64(xyz.shape[0]), C.c_float(theta_deg), C.c_float(tx),
                              C.c_float(ty), C.c_float(tz))
    return xyz
