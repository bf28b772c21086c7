"""Helpers shared by the GPU parity tests: run the oracle and the CUDA path on one input and
compare every contract output."""
import importlib

import numpy as np

import oracle_ffi as o
from helpers import load_calib_c1, scaled_calib

s3 = importlib.import_module("3dscan_b200")


def calibs(cam_scale=1.0, proj_scale=1.0, dc=None, dp=None):
    c = scaled_calib(load_calib_c1(), cam_scale, proj_scale)
    if dc is not None:
        c["dc"] = np.asarray(dc, np.float64)
    if dp is not None:
        c["dp"] = np.asarray(dp, np.float64)
    args = [c[k] for k in ("Kc", "dc", "Kp", "dp", "rc", "tc", "rp", "tp")]
    return s3.make_calib(*args), o.make_calib(*args), c


def cfg_dict(cfg):
    return {k: getattr(cfg, k) for k in ("W", "H", "PW", "PH", "N", "M_v", "M_h", "fw_v", "fw_h", "dirs")}


def run_oracle(cfg, ocal, stack, roi, threads=0, modulation=False, strict=False, colrow=False):
    d = s3.split_stack(cfg, stack)
    return o.reconstruct(cfg_dict(cfg), ocal, d["fringe_v"], d["gray_v"], d["inv_v"],
                         d.get("fringe_h"), d.get("gray_h"), d.get("inv_h"), roi, threads=threads,
                         modulation=modulation, strict=strict, colrow=colrow)


def compare(cfg, ref, ctx, fused=True, report=None):
    """Asserts the parity bar: integer planes bit-exact, phase <= 1e-4 rad (and counts the
    non-identical floats), points <= 1e-5 relative.  Returns a dict of statistics."""
    st = {}
    mask = ref.valid_v == 1
    code_v = ctx.code_i32(0)
    assert np.array_equal(code_v, ref.code_v), "code_v differs"
    unw_v = ctx.plane(s3.PLANE_UNWRAPPED_V)
    st["unw_v_nonidentical"] = int((unw_v.view(np.uint32) != ref.unwrapped_v.view(np.uint32)).sum())
    assert np.abs(unw_v.astype(np.float64) - ref.unwrapped_v).max(initial=0) <= 1e-4
    if cfg.dirs == 1:
        m = ctx.plane(s3.PLANE_VALID if fused else s3.PLANE_MASK)
        assert np.array_equal(m.astype(np.int32), ref.valid_v), "mask differs"
        return st
    assert np.array_equal(ctx.code_i32(1), ref.code_h), "code_h differs"
    unw_h = ctx.plane(s3.PLANE_UNWRAPPED_H)
    st["unw_h_nonidentical"] = int((unw_h.view(np.uint32) != ref.unwrapped_h.view(np.uint32)).sum())
    assert np.abs(unw_h.astype(np.float64) - ref.unwrapped_h).max(initial=0) <= 1e-4
    valid = ctx.plane(s3.PLANE_VALID).astype(np.int32)
    cp = ctx.plane(s3.PLANE_CPMAP).reshape(-1, 2).astype(np.int64)
    st["cpmap_mismatch"] = int((cp != ref.cpmap).any(axis=1).sum())
    st["valid_mismatch"] = int((valid != ref.valid).sum())
    n = ctx.point_count()
    st["count"] = n
    pts = ctx.points()
    if st["unw_v_nonidentical"] == 0 and st["unw_h_nonidentical"] == 0:
        assert st["cpmap_mismatch"] == 0 and st["valid_mismatch"] == 0
        assert n == ref.count
        st["pts_nonidentical"] = int((pts.view(np.uint32) != ref.pts.view(np.uint32)).any(axis=1).sum())
        scale = np.maximum(np.abs(ref.pts).max(axis=1, keepdims=True), 1e-30)
        st["pts_max_rel"] = float((np.abs(pts - ref.pts) / scale).max(initial=0))
        assert st["pts_max_rel"] <= 1e-5
    else:
        # a phase that differs in the last float bit may move an lrint() correspondence: bound how many pixels
        # that can touch, and still hold every pixel it cannot have touched to the full bar
        assert st["cpmap_mismatch"] <= st["unw_v_nonidentical"] + st["unw_h_nonidentical"]
        assert st["valid_mismatch"] <= st["cpmap_mismatch"]
        touched = ((unw_v.view(np.uint32) != ref.unwrapped_v.view(np.uint32)) |
                   (unw_h.view(np.uint32) != ref.unwrapped_h.view(np.uint32))).reshape(-1)
        assert not ((cp != ref.cpmap).any(axis=1) & ~touched).any(), "c_p_map differs at a pixel whose phases are identical"
        assert not ((valid.reshape(-1) != ref.valid.reshape(-1)) & ~touched).any()
        # points: match them by pixel (both lists are in raster order of their valid pixels)
        g_pix = np.flatnonzero(valid.reshape(-1) == 1)
        r_pix = np.flatnonzero(ref.valid.reshape(-1) == 1)
        assert n == g_pix.size
        common = np.intersect1d(g_pix[~touched[g_pix]], r_pix[~touched[r_pix]])
        gi, ri = np.searchsorted(g_pix, common), np.searchsorted(r_pix, common)
        scale = np.maximum(np.abs(ref.pts[ri]).max(axis=1, keepdims=True), 1e-30)
        st["pts_max_rel"] = float((np.abs(pts[gi] - ref.pts[ri]) / scale).max(initial=0))
        st["pts_compared"] = int(common.size)
        assert st["pts_max_rel"] <= 1e-5
        assert abs(n - ref.count) <= st["valid_mismatch"]
    if report is not None:
        report.update(st)
    return st
