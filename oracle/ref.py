"""oracle/_ref -- the reference's OWN functions on the hot path, compiled unmodified (oracle/build_ref.sh).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.

ctypes binding of oracle/_ref/libref3dr.so = /root/reference/src/utils.cpp:282-430 (shi_tomasi_score, halfSampleSSE2,
reduce_to_half, create_img_pyramid), src/initialization.cpp:171-249 (InitHelper::CheckFundamental) and
src/camera.cpp:25-41 (Pinhole::cam2world) between two stand-in headers (ref_stub_prefix.hpp / ref_stub_suffix.hpp).  It
pins the restatements in oracle/lk_oracle.c, oracle/fast_oracle.c and oracle/postfilter.py, and through them the CUDA
path, to the reference itself.  /root/reference does not exist on the GPU box: the library is built here (where the
reference is) and travels with the snapshot; `available()` says whether it is there.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libref3dr.so")
_lib = None


def build():
    """Run oracle/build_ref.sh (compiles only where /root/reference exists; keeps a prebuilt library otherwise)."""
    stubs = [os.path.join(_HERE, f) for f in ("ref_stub_prefix.hpp", "ref_stub_suffix.hpp", "build_ref.sh")]
    if os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(f) for f in stubs):
        return _LIB_PATH
    r = subprocess.run(["bash", os.path.join(_HERE, "build_ref.sh")], capture_output=True, text=True)
    if r.returncode != 0 and not os.path.exists(_LIB_PATH):
        raise RuntimeError("oracle/_ref could not be built: " + r.stderr[-2000:])
    return _LIB_PATH


def available():
    if os.path.exists(_LIB_PATH):
        return True
    try:
        build()
    except RuntimeError:
        return False
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p, f32p, f64p, i32p = (ctypes.POINTER(t) for t in (ctypes.c_uint8, ctypes.c_float, ctypes.c_double, ctypes.c_int32))
        c_int, c_double = ctypes.c_int, ctypes.c_double
        L.ref_box_pyramid.argtypes = [u8p, c_int, c_int, c_int, ctypes.POINTER(u8p)]
        L.ref_shi_tomasi.argtypes = [u8p, c_int, c_int, ctypes.c_long, i32p, c_int, f32p]
        L.ref_shi_tomasi.restype = None
        L.ref_check_fundamental.argtypes = [f32p, f32p, f32p, c_int, ctypes.c_float, u8p]
        L.ref_check_fundamental.restype = ctypes.c_float
        L.ref_cam2world.argtypes = [c_double, c_double, c_double, c_double, f64p, f64p, c_int, f64p]
        L.ref_cam2world.restype = None
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def box_pyramid(img, n_levels=3):
    """utils::create_img_pyramid (src/utils.cpp:421-430) as the Frame ctor calls it.  Returns [img, L1, L2, ...]; raises
    ValueError on the shapes where the reference's pointer walk leaves its buffers."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    outs, cw, ch = [], w, h
    for _ in range(1, n_levels):
        cw, ch = cw // 2, ch // 2
        outs.append(np.zeros((ch, cw), np.uint8))
    ptrs = (ctypes.POINTER(ctypes.c_uint8) * max(1, len(outs)))(*[_p(o, ctypes.c_uint8) for o in outs])
    rc = lib().ref_box_pyramid(_p(img, ctypes.c_uint8), w, h, n_levels, ptrs)
    if rc != 0:
        raise ValueError("ref_box_pyramid rc=%d" % rc)
    return [img] + outs


def shi_tomasi(img, uv):
    """utils::shi_tomasi_score (src/utils.cpp:282-321) at integer positions uv (n, 2) = (u, v)."""
    img = np.ascontiguousarray(img, np.uint8)
    uv = np.ascontiguousarray(uv, np.int32).reshape(-1, 2)
    out = np.zeros(len(uv), np.float32)
    lib().ref_shi_tomasi(_p(img, ctypes.c_uint8), img.shape[1], img.shape[0], img.strides[0], _p(uv, ctypes.c_int32), len(uv),
                         _p(out, ctypes.c_float))
    return out


def check_fundamental(F21, pts1, pts2, sigma=1.0):
    """InitHelper::CheckFundamental (src/initialization.cpp:171-249) for a stack of hypotheses -> (scores (H,), inliers (H, N))."""
    F = np.ascontiguousarray(F21, np.float32).reshape(-1, 9)
    p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
    scores = np.zeros(len(F), np.float32)
    inl = np.zeros((len(F), len(p1)), np.uint8)
    for i in range(len(F)):
        scores[i] = lib().ref_check_fundamental(_p(F[i], ctypes.c_float), _p(p1, ctypes.c_float), _p(p2, ctypes.c_float), len(p1),
                                                float(sigma), _p(inl[i], ctypes.c_uint8))
    return scores, inl


def cam2world(uv, fx, fy, cx, cy, dist=(0, 0, 0, 0, 0)):
    """Pinhole::cam2world(u, v) (src/camera.cpp:25-41): unit bearing vectors (n, 3) of pixels uv (n, 2), double."""
    uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2)
    d = np.ascontiguousarray(dist, np.float64)
    out = np.zeros((len(uv), 3), np.float64)
    lib().ref_cam2world(fx, fy, cx, cy, _p(d, ctypes.c_double), _p(uv, ctypes.c_double), len(uv), _p(out, ctypes.c_double))
    return out
