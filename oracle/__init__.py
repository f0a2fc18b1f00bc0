"""CPU oracle for the pyramidal-LK hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes wrapper over oracle/lk_oracle.c (see its header for what it restates and how it is pinned).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; nothing under 3dr_b200/ does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

BOX_AUTO_X86, BOX_TRUNC, BOX_SSE2 = 0, 1, 2
TERM_COUNT, TERM_EPS = 1, 2
USE_INITIAL_FLOW, GET_MIN_EIGENVALS = 4, 8

_lib = None


def build(force=False):
    """Compile oracle/lk_oracle.c with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("lk_oracle.c", "fast_oracle.c")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p, f32p, i16p, i32p = (ctypes.POINTER(t) for t in (ctypes.c_uint8, ctypes.c_float, ctypes.c_int16, ctypes.c_int32))
        c_int, c_long, c_double = ctypes.c_int, ctypes.c_long, ctypes.c_double
        L.orc_box_half.argtypes = [u8p, c_int, c_int, c_long, u8p, c_int]
        L.orc_pyrdown.argtypes = [u8p, c_int, c_int, c_long, u8p, c_long]
        L.orc_scharr.argtypes = [u8p, c_int, c_int, c_long, i16p, c_long]
        L.orc_lk_level_sizes.argtypes = [c_int, c_int, c_int, c_int, c_int, i32p, i32p]
        L.orc_calc_optical_flow_pyr_lk.argtypes = [u8p, c_long, u8p, c_long, c_int, c_int, f32p, f32p, u8p, f32p, c_int,
                                                   c_int, c_int, c_int, c_int, c_int, c_double, c_int, c_double, c_int,
                                                   i32p, u8p, f32p]
        L.orc_num_threads.restype = c_int
        i16p2 = ctypes.POINTER(ctypes.c_int16)
        L.orc_fast_detect.argtypes = [u8p, c_int, c_int, c_long, c_int, c_int, i16p2, c_int]
        L.orc_fast_score.argtypes = [u8p, c_long, i16p2, c_int, c_int, c_int, i32p]
        L.orc_fast_nonmax.argtypes = [i16p2, i32p, c_int, i32p]
        L.orc_shi_tomasi.argtypes = [u8p, c_int, c_int, c_long, c_int, c_int]
        L.orc_shi_tomasi.restype = ctypes.c_float
        L.orc_fast_detector_arc.argtypes = [ctypes.POINTER(u8p), c_int, c_int, c_int, c_int, c_int, c_double, u8p, c_int, i32p, i32p, f32p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _gray(img):
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 2
    if img.strides[1] != 1:
        img = np.ascontiguousarray(img)
    return img


def box_half(img, mode=BOX_AUTO_X86):
    """utils::reduce_to_half (src/utils.cpp:382-419). Returns (h/2, w/2) uint8, or raises on the
    shapes where the reference itself overruns its buffers."""
    img = _gray(img)
    h, w = img.shape
    out = np.zeros((h // 2, w // 2), np.uint8)
    rc = lib().orc_box_half(_p(img, ctypes.c_uint8), w, h, img.strides[0], _p(out, ctypes.c_uint8), mode)
    if rc != 0:
        raise ValueError("orc_box_half rc=%d" % rc)
    return out


def box_pyramid(img, n_levels=3, mode=BOX_AUTO_X86):
    """utils::create_img_pyramid (src/utils.cpp:421-430): [img, half, quarter, ...]."""
    pyr = [_gray(img)]
    for _ in range(1, n_levels):
        pyr.append(box_half(pyr[-1], mode))
    return pyr


def pyrdown(img):
    img = _gray(img)
    h, w = img.shape
    out = np.zeros(((h + 1) // 2, (w + 1) // 2), np.uint8)
    rc = lib().orc_pyrdown(_p(img, ctypes.c_uint8), w, h, img.strides[0], _p(out, ctypes.c_uint8), out.strides[0])
    assert rc == 0
    return out


def scharr(img):
    """calcScharrDeriv: (h, w, 2) int16, channel 0 = Ix, 1 = Iy."""
    img = _gray(img)
    h, w = img.shape
    out = np.zeros((h, w, 2), np.int16)
    rc = lib().orc_scharr(_p(img, ctypes.c_uint8), w, h, img.strides[0], _p(out, ctypes.c_int16), 2 * w)
    assert rc == 0
    return out


def lk_level_sizes(w, h, win, max_level):
    """[(w_l, h_l)] for l = 0..effective maxLevel (buildOpticalFlowPyramid's early stop)."""
    ws = np.zeros(32, np.int32)
    hs = np.zeros(32, np.int32)
    ml = lib().orc_lk_level_sizes(w, h, win[0], win[1], min(max_level, 31), _p(ws, ctypes.c_int32), _p(hs, ctypes.c_int32))
    return [(int(ws[l]), int(hs[l])) for l in range(ml + 1)]


def build_lk_pyramid(img, win=(21, 21), max_level=3, with_derivatives=False):
    """Gaussian pyramid (and Scharr derivatives) as calcOpticalFlowPyrLK builds them internally."""
    levels = [_gray(img)]
    h, w = levels[0].shape
    for (_w, _h) in lk_level_sizes(w, h, win, max_level)[1:]:
        levels.append(pyrdown(levels[-1]))
    if with_derivatives:
        return levels, [scharr(l) for l in levels]
    return levels


def calc_optical_flow_pyr_lk(prev, nxt, prev_pts, next_pts=None, win=(21, 21), max_level=3,
                             criteria=(TERM_COUNT | TERM_EPS, 30, 0.01), flags=0, min_eig_threshold=1e-4,
                             nthreads=0, want_err=True, trace=False):
    """cv::calcOpticalFlowPyrLK semantics. Returns (next_pts (N,2) f32, status (N,) u8, err (N,) f32[, trace])."""
    prev, nxt = _gray(prev), _gray(nxt)
    assert prev.shape == nxt.shape
    h, w = prev.shape
    pp = np.ascontiguousarray(np.asarray(prev_pts, np.float32).reshape(-1, 2))
    n = pp.shape[0]
    if flags & USE_INITIAL_FLOW:
        npts = np.ascontiguousarray(np.asarray(next_pts, np.float32).reshape(-1, 2)).copy()
        assert npts.shape[0] == n
    else:
        npts = np.zeros((n, 2), np.float32)
    status = np.zeros(n, np.uint8)
    err = np.zeros(n, np.float32)
    ti = np.zeros((n, 32), np.int32) if trace else None
    tc = np.zeros((n, 32), np.uint8) if trace else None
    tp = np.zeros((n, 32, 2), np.float32) if trace else None
    rc = lib().orc_calc_optical_flow_pyr_lk(
        _p(prev, ctypes.c_uint8), prev.strides[0], _p(nxt, ctypes.c_uint8), nxt.strides[0], w, h,
        _p(pp, ctypes.c_float), _p(npts, ctypes.c_float), _p(status, ctypes.c_uint8),
        _p(err, ctypes.c_float) if want_err else None, n, win[0], win[1], max_level,
        criteria[0], criteria[1], float(criteria[2]), flags, float(min_eig_threshold), nthreads,
        _p(ti, ctypes.c_int32) if trace else None, _p(tc, ctypes.c_uint8) if trace else None,
        _p(tp, ctypes.c_float) if trace else None)
    if rc < 0:
        raise ValueError("orc_calc_optical_flow_pyr_lk rc=%d" % rc)
    if trace:
        return npts, status, err, {"iters": ti, "code": tc, "pos": tp, "max_level": rc}
    return npts, status, err


def fast_detect(img, threshold=20, arc=10):
    """fast_corner_detect_<arc>: (n, 2) int16 corners (x, y) in raster order."""
    img = np.ascontiguousarray(_gray(img))
    h, w = img.shape
    xy = np.zeros((w * h, 2), np.int16)
    n = lib().orc_fast_detect(_p(img, ctypes.c_uint8), w, h, w, threshold, arc, _p(xy, ctypes.c_int16), w * h)
    return xy[:n].copy()


def fast_score(img, xy, threshold=20, arc=10):
    img = np.ascontiguousarray(_gray(img))
    xy = np.ascontiguousarray(xy, np.int16)
    sc = np.zeros(len(xy), np.int32)
    lib().orc_fast_score(_p(img, ctypes.c_uint8), img.shape[1], _p(xy, ctypes.c_int16), len(xy), threshold, arc, _p(sc, ctypes.c_int32))
    return sc


def fast_nonmax(xy, scores):
    xy = np.ascontiguousarray(xy, np.int16)
    scores = np.ascontiguousarray(scores, np.int32)
    keep = np.zeros(max(len(xy), 1), np.int32)
    n = lib().orc_fast_nonmax(_p(xy, ctypes.c_int16), _p(scores, ctypes.c_int32), len(xy), _p(keep, ctypes.c_int32))
    return keep[:n].copy()


def shi_tomasi(img, u, v):
    img = np.ascontiguousarray(_gray(img))
    return float(lib().orc_shi_tomasi(_p(img, ctypes.c_uint8), img.shape[1], img.shape[0], img.shape[1], int(u), int(v)))


def fast_detector(img, n_levels=3, cell_size=30, fast_threshold=20, detection_threshold=20.0, occupancy=None, box_mode=BOX_AUTO_X86,
                  arc=10):
    """FastDetector::detect on the Frame's box pyramid (src/features.cpp:43-98, src/frame.cpp:13-20).
    Returns (xy (n,2) int32 level-0 coordinates, level (n,), score (n,) float32) in grid-cell order."""
    pyr = [np.ascontiguousarray(p) for p in box_pyramid(img, n_levels, box_mode)]
    h, w = pyr[0].shape
    gc, gr = -(-w // cell_size), -(-h // cell_size)
    ptrs = (ctypes.POINTER(ctypes.c_uint8) * n_levels)(*[_p(p, ctypes.c_uint8) for p in pyr])
    xy = np.zeros((gc * gr, 2), np.int32)
    lv = np.zeros(gc * gr, np.int32)
    sc = np.zeros(gc * gr, np.float32)
    occ = np.ascontiguousarray(occupancy, np.uint8) if occupancy is not None else None
    n = lib().orc_fast_detector_arc(ptrs, w, h, n_levels, cell_size, fast_threshold, float(detection_threshold),
                                    _p(occ, ctypes.c_uint8) if occ is not None else None, int(arc), _p(xy, ctypes.c_int32),
                                    _p(lv, ctypes.c_int32), _p(sc, ctypes.c_float))
    assert n >= 0
    return xy[:n].copy(), lv[:n].copy(), sc[:n].copy()


def num_threads():
    return lib().orc_num_threads()
