"""3dr_b200 -- B200-native pyramidal Lucas-Kanade path for kvmanohar22/3DR (ctypes binding of include/dr3lk.h).

The product is the CUDA library `3dr_b200/lib/libdr3lk.so` (hand-written sm_100a kernels behind a C ABI, built by
`3dr_b200/csrc/Makefile` / `__graft_entry__.build()`); this module only loads it and mirrors the reference's two call
surfaces for this path with numpy arrays in place of cv::Mat / std::vector:

  * `Context.calc_optical_flow_pyr_lk`  <->  cv::calcOpticalFlowPyrLK   (reference src/initialization.cpp:608-613)
  * `Context.box_pyramid`               <->  utils::create_img_pyramid  (reference src/utils.cpp:421-430)

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked), but creating a
`Context` without a CUDA device raises, and a missing library raises at import of `lib()`.

The package name starts with a digit, so import it with `importlib.import_module("3dr_b200")`.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DR3LK_LIB", os.path.join(_HERE, "lib", "libdr3lk.so"))  # override only for A/B kernel experiments

OK, E_ARG, E_SIZE, E_CUDA, E_UNSUPPORTED = 0, -1, -2, -3, -4
TERM_COUNT, TERM_EPS = 1, 2
USE_INITIAL_FLOW, GET_MIN_EIGENVALS = 4, 8
BOX_AUTO_X86, BOX_TRUNC, BOX_SSE2 = 0, 1, 2
MAX_LEVELS = 16

# every symbol include/dr3lk.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "dr3lk_create", "dr3lk_destroy", "dr3lk_last_error", "dr3lk_set_stream", "dr3lk_synchronize", "dr3lk_launch_count",
    "dr3lk_set_profiling", "dr3lk_profile_read", "dr3lk_debug_check_read", "dr3lk_debug_set_fast_arc", "dr3lk_host_alloc", "dr3lk_host_free", "dr3lk_host_register", "dr3lk_host_unregister", "dr3lk_box_pyramid", "dr3lk_box_pyramid_device",
    "dr3lk_calc_optical_flow_pyr_lk", "dr3lk_track_batch", "dr3lk_track_batch_host", "dr3lk_lk_level_sizes",
    "dr3lk_build_lk_pyramid", "dr3lk_pyramid_create", "dr3lk_pyramid_destroy", "dr3lk_pyramid_levels",
    "dr3lk_calc_optical_flow_pyr_lk_cached", "dr3lk_track_frame", "dr3lk_filter_tracks", "dr3lk_fast_detect", "dr3lk_score_fundamental",
    "dr3lk_multi_create", "dr3lk_multi_destroy", "dr3lk_multi_size", "dr3lk_multi_context", "dr3lk_multi_last_error", "dr3lk_shard_range",
    "dr3lk_multi_track_batch_host", "dr3lk_init_first_frame", "dr3lk_init_second_frame", "dr3lk_init_score_fundamental",
]


class Dr3lkError(RuntimeError):
    """Raised where the reference path would throw cv::Exception (or on CUDA failures)."""

    def __init__(self, code, msg):
        super().__init__("dr3lk error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load libdr3lk.so; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    c_int, c_size_t, c_double, c_void_p = ctypes.c_int, ctypes.c_size_t, ctypes.c_double, ctypes.c_void_p
    P = ctypes.POINTER
    L.dr3lk_create.argtypes = [P(c_void_p), c_int]
    L.dr3lk_destroy.argtypes = [c_void_p]
    L.dr3lk_destroy.restype = None
    L.dr3lk_last_error.argtypes = [c_void_p]
    L.dr3lk_last_error.restype = ctypes.c_char_p
    L.dr3lk_set_stream.argtypes = [c_void_p, c_void_p]
    L.dr3lk_synchronize.argtypes = [c_void_p]
    if hasattr(L, "dr3lk_debug_check_read"):
        L.dr3lk_debug_check_read.argtypes = [c_void_p, P(ctypes.c_uint64)]
    L.dr3lk_debug_set_fast_arc.argtypes = [c_void_p, c_int]
    L.dr3lk_launch_count.argtypes = [c_void_p]
    L.dr3lk_launch_count.restype = ctypes.c_uint64
    L.dr3lk_set_profiling.argtypes = [c_void_p, c_int]
    L.dr3lk_profile_read.argtypes = [c_void_p, P(ctypes.c_float), P(c_int), P(ctypes.c_float), P(c_int)]
    L.dr3lk_host_alloc.argtypes = [c_size_t]
    L.dr3lk_host_alloc.restype = c_void_p
    L.dr3lk_host_free.argtypes = [c_void_p]
    L.dr3lk_host_free.restype = None
    L.dr3lk_host_register.argtypes = [c_void_p, c_size_t]
    L.dr3lk_host_unregister.argtypes = [c_void_p]
    L.dr3lk_box_pyramid.argtypes = [c_void_p, c_void_p, c_int, c_int, c_size_t, c_int, P(c_void_p), c_int]
    L.dr3lk_box_pyramid_device.argtypes = [c_void_p, c_void_p, c_int, c_int, c_size_t, c_size_t, c_int, c_int, P(c_void_p), c_int]
    lk_tail = [c_int, c_int, c_int, c_int, c_int, c_double, c_int, c_double]  # win_w .. min_eig_threshold
    L.dr3lk_calc_optical_flow_pyr_lk.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_int, c_void_p,
                                                 c_void_p, c_void_p, c_void_p, c_int] + lk_tail
    L.dr3lk_track_batch.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_size_t, c_size_t, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p] + lk_tail
    L.dr3lk_track_batch_host.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_size_t, c_size_t, c_int, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int] + lk_tail
    if hasattr(L, "dr3lk_multi_create"):  # absent only in older builds loaded through DR3LK_LIB for A/B runs
        L.dr3lk_multi_create.argtypes = [P(c_void_p), P(c_int), c_int]
        L.dr3lk_multi_destroy.argtypes = [c_void_p]
        L.dr3lk_multi_destroy.restype = None
        L.dr3lk_multi_size.argtypes = [c_void_p]
        L.dr3lk_multi_context.argtypes = [c_void_p, c_int]
        L.dr3lk_multi_context.restype = c_void_p
        L.dr3lk_multi_last_error.argtypes = [c_void_p]
        L.dr3lk_multi_last_error.restype = ctypes.c_char_p
        L.dr3lk_shard_range.argtypes = [c_int, c_int, c_int, P(c_int), P(c_int)]
        L.dr3lk_shard_range.restype = None
        L.dr3lk_multi_track_batch_host.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_size_t, c_size_t, c_int, c_void_p,
                                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int] + lk_tail
    L.dr3lk_lk_level_sizes.argtypes = [c_int, c_int, c_int, c_int, c_int, P(c_int), P(c_int)]
    L.dr3lk_build_lk_pyramid.argtypes = [c_void_p, c_void_p, c_int, c_int, c_size_t, c_int, c_int, c_int, P(c_void_p),
                                         P(c_void_p), P(c_int)]
    L.dr3lk_pyramid_create.argtypes = [c_void_p, c_void_p, c_int, c_int, c_size_t, c_int, c_int, c_int, P(c_void_p)]
    L.dr3lk_pyramid_destroy.argtypes = [c_void_p]
    L.dr3lk_pyramid_destroy.restype = None
    L.dr3lk_pyramid_levels.argtypes = [c_void_p]
    L.dr3lk_calc_optical_flow_pyr_lk_cached.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int] + lk_tail
    if hasattr(L, "dr3lk_track_frame"):  # absent only in older builds loaded through DR3LK_LIB for A/B runs
        L.dr3lk_track_frame.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_int] + lk_tail + [c_int, P(c_void_p)]
    L.dr3lk_filter_tracks.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_double, c_double, c_double, c_double, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, P(c_int)]
    L.dr3lk_fast_detect.argtypes = [c_void_p, c_void_p, c_int, c_int, c_size_t, c_int, c_int, c_int, c_double, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, P(c_int)]
    L.dr3lk_score_fundamental.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, ctypes.c_float, c_void_p, c_void_p, P(c_int)]
    if hasattr(L, "dr3lk_init_first_frame"):  # absent only in older builds loaded through DR3LK_LIB for A/B runs
        L.dr3lk_init_first_frame.argtypes = [c_void_p, c_void_p, c_int, c_int, c_size_t, c_int, c_int, c_int, c_double, c_int, c_void_p, c_int, c_int,
                                             c_int, c_void_p, c_void_p, c_void_p, P(c_int), P(c_void_p)]
        L.dr3lk_init_second_frame.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_int] + lk_tail + [
            c_double, c_double, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, P(c_int)]
        L.dr3lk_init_score_fundamental.argtypes = [c_void_p, c_void_p, c_int, ctypes.c_float, c_void_p, c_void_p, P(c_int)]
    _lib = L
    return L


def lk_level_sizes(w, h, win=(21, 21), max_level=3):
    """[(w_l, h_l)] for l = 0..effective maxLevel (buildOpticalFlowPyramid's early stop)."""
    n = min(max_level, MAX_LEVELS - 1) + 1
    ws, hs = (ctypes.c_int * n)(), (ctypes.c_int * n)()
    ml = lib().dr3lk_lk_level_sizes(w, h, win[0], win[1], min(max_level, MAX_LEVELS - 1), ws, hs)
    return [(ws[l], hs[l]) for l in range(ml + 1)]


def _gray(img):
    img = np.asarray(img)
    if img.dtype != np.uint8 or img.ndim != 2:
        raise Dr3lkError(E_SIZE, "image must be 2-D uint8 (CV_8UC1)")
    if img.strides[1] != 1 or img.strides[0] < img.shape[1]:
        img = np.ascontiguousarray(img)
    return img


def host_register(arr):
    """Page-locks the memory of an existing C-contiguous numpy array (dr3lk_host_register); pair with host_unregister(arr)."""
    rc = lib().dr3lk_host_register(arr.ctypes.data, arr.nbytes)
    if rc != 0:
        raise Dr3lkError(rc, "dr3lk_host_register failed")


def host_unregister(arr):
    rc = lib().dr3lk_host_unregister(arr.ctypes.data)
    if rc != 0:
        raise Dr3lkError(rc, "dr3lk_host_unregister failed")


class PinnedArray:
    """numpy view over page-locked host memory from dr3lk_host_alloc (lets the host-batch path overlap copies)."""

    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._ptr = lib().dr3lk_host_alloc(max(nbytes, 1))
        if not self._ptr:
            raise Dr3lkError(E_CUDA, "dr3lk_host_alloc(%d) failed" % nbytes)
        buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._ptr:
            self.array = None
            lib().dr3lk_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One CUDA device + stream + scratch memory (dr3lk_ctx). Not thread-safe; use one per thread / GPU."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        rc = lib().dr3lk_create(ctypes.byref(self._h), device)
        if rc != OK:
            raise Dr3lkError(rc, lib().dr3lk_last_error(None).decode())
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().dr3lk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != OK:
            raise Dr3lkError(rc, lib().dr3lk_last_error(self._h).decode())

    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream), or None/0 for the own stream."""
        self._check(lib().dr3lk_set_stream(self._h, ctypes.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._check(lib().dr3lk_synchronize(self._h))

    def set_profiling(self, on):
        self._check(lib().dr3lk_set_profiling(self._h, int(bool(on))))

    def profile_read(self):
        """(lk_ms, lk_launches, pyramid_ms, pyramid_builds) since the last read; waits for the recorded events."""
        a, b, c, d = ctypes.c_float(), ctypes.c_int(), ctypes.c_float(), ctypes.c_int()
        self._check(lib().dr3lk_profile_read(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
        return a.value, b.value, c.value, d.value

    @property
    def launch_count(self):
        return int(lib().dr3lk_launch_count(self._h))

    def debug_check_read(self):
        """Checked build (DR3LK_LIB=.../libdr3lk_checked.so) only: (violations, first kind, first detail, checks executed)
        of the LK kernels' own bounds checks since the last read; raises E_UNSUPPORTED with the default library."""
        out = (ctypes.c_uint64 * 4)()
        self._check(lib().dr3lk_debug_check_read(self._h, out))
        return tuple(int(v) for v in out)

    def debug_set_fast_arc(self, arc):
        """Parity hook: 9 makes fast_detect / init_first_frame run OpenCV's FAST-9 segment test instead of the reference's FAST-10."""
        self._check(lib().dr3lk_debug_set_fast_arc(self._h, int(arc)))

    # ---- utils::create_img_pyramid ------------------------------------------------------------
    def box_pyramid(self, img, n_levels=3, mode=BOX_AUTO_X86):
        """Returns [img, level1, ...]; level 0 is the caller's array (shallow, as in the reference)."""
        img0 = _gray(img)
        h, w = img0.shape
        outs, lw, lh = [], w, h
        for _ in range(1, n_levels):
            lw, lh = lw // 2, lh // 2
            outs.append(np.zeros((max(lh, 0), max(lw, 0)), np.uint8))
        ptrs = (ctypes.c_void_p * max(len(outs), 1))(*[o.ctypes.data for o in outs])
        self._check(lib().dr3lk_box_pyramid(self._h, img0.ctypes.data, w, h, img0.strides[0], n_levels, ptrs, mode))
        return [img0] + outs

    def box_pyramid_device(self, img_ptr, w, h, pitch, image_stride, batch, out_ptrs, mode=BOX_AUTO_X86):
        n_levels = len(out_ptrs) + 1
        ptrs = (ctypes.c_void_p * max(len(out_ptrs), 1))(*out_ptrs)
        self._check(lib().dr3lk_box_pyramid_device(self._h, img_ptr, w, h, pitch, image_stride, batch, n_levels, ptrs, mode))

    # ---- the LK pyramids ------------------------------------------------------------------------
    def build_lk_pyramid(self, img, win=(21, 21), max_level=3, with_derivatives=False):
        img0 = _gray(img)
        h, w = img0.shape
        sizes = lk_level_sizes(w, h, win, max_level)
        levels = [np.zeros((lh, lw), np.uint8) for (lw, lh) in sizes]
        derivs = [np.zeros((lh, lw, 2), np.int16) for (lw, lh) in sizes] if with_derivatives else None
        lp = (ctypes.c_void_p * len(levels))(*[a.ctypes.data for a in levels])
        dp = (ctypes.c_void_p * len(levels))(*[a.ctypes.data for a in derivs]) if with_derivatives else None
        ml = ctypes.c_int(-1)
        self._check(lib().dr3lk_build_lk_pyramid(self._h, img0.ctypes.data, w, h, img0.strides[0], win[0], win[1], max_level,
                                                 lp, dp, ctypes.byref(ml)))
        assert ml.value == len(sizes) - 1
        return (levels, derivs) if with_derivatives else levels

    # ---- cv::calcOpticalFlowPyrLK ---------------------------------------------------------------
    def calc_optical_flow_pyr_lk(self, prev, nxt, prev_pts, next_pts=None, win=(21, 21), max_level=3,
                                 criteria=(TERM_COUNT | TERM_EPS, 30, 0.01), flags=0, min_eig_threshold=1e-4, want_err=True):
        """Same argument meaning as cv2.calcOpticalFlowPyrLK; returns (nextPts (N,2) f32, status (N,) u8, err (N,) f32|None)."""
        prev, nxt = _gray(prev), _gray(nxt)
        if prev.shape != nxt.shape:
            raise Dr3lkError(E_SIZE, "(-215:Assertion failed) prevPyr[level * lvlStep1].size() == nextPyr[level * lvlStep2].size()")
        h, w = prev.shape
        pp = np.ascontiguousarray(np.asarray(prev_pts, np.float32).reshape(-1, 2))
        n = pp.shape[0]
        if flags & USE_INITIAL_FLOW:
            if next_pts is None:
                raise Dr3lkError(E_ARG, "OPTFLOW_USE_INITIAL_FLOW needs nextPts")
            npts = np.ascontiguousarray(np.asarray(next_pts, np.float32).reshape(-1, 2)).copy()
            if npts.shape[0] != n:
                raise Dr3lkError(E_ARG, "(-215:Assertion failed) nextPtsMat.checkVector(2, CV_32F, true) == npoints")
        else:
            npts = np.zeros((n, 2), np.float32)
        status = np.zeros(n, np.uint8)
        err = np.zeros(n, np.float32) if want_err else None
        self._check(lib().dr3lk_calc_optical_flow_pyr_lk(
            self._h, prev.ctypes.data, prev.strides[0], nxt.ctypes.data, nxt.strides[0], w, h, pp.ctypes.data,
            npts.ctypes.data, status.ctypes.data, err.ctypes.data if want_err else None, n, win[0], win[1], max_level,
            criteria[0], criteria[1], float(criteria[2]), flags, float(min_eig_threshold)))
        return npts, status, err

    def calc_optical_flow_pyr_lk_cached(self, prev_pyr, next_pyr, prev_pts, next_pts=None, max_level=3,
                                        criteria=(TERM_COUNT | TERM_EPS, 30, 0.01), flags=0, min_eig_threshold=1e-4, want_err=True):
        """calc_optical_flow_pyr_lk on two prebuilt `Pyramid`s (no image upload, no pyramid build)."""
        pp = np.ascontiguousarray(np.asarray(prev_pts, np.float32).reshape(-1, 2))
        n = pp.shape[0]
        if flags & USE_INITIAL_FLOW:
            npts = np.ascontiguousarray(np.asarray(next_pts, np.float32).reshape(-1, 2)).copy()
            if npts.shape[0] != n:
                raise Dr3lkError(E_ARG, "(-215:Assertion failed) nextPtsMat.checkVector(2, CV_32F, true) == npoints")
        else:
            npts = np.zeros((n, 2), np.float32)
        status = np.zeros(n, np.uint8)
        err = np.zeros(n, np.float32) if want_err else None
        win = prev_pyr.win
        self._check(lib().dr3lk_calc_optical_flow_pyr_lk_cached(
            self._h, prev_pyr._h, next_pyr._h, pp.ctypes.data, npts.ctypes.data, status.ctypes.data,
            err.ctypes.data if want_err else None, n, win[0], win[1], max_level, criteria[0], criteria[1], float(criteria[2]), flags,
            float(min_eig_threshold)))
        return npts, status, err

    def track_frame(self, prev_pyr, next_img, prev_pts, next_pts=None, max_level=3, criteria=(TERM_COUNT | TERM_EPS, 30, 0.01),
                    flags=0, min_eig_threshold=1e-4, want_err=True, keep_next=0):
        """Streaming form: track from a cached previous-frame `Pyramid` into a NEW host image in one call (one upload, the
        new frame's pyramid, LK, one download).  keep_next: 0 discard the new frame's pyramid, 1 keep its Gaussian levels,
        2 keep it with derivatives (it can be the previous frame of the next call).  Returns (next_pts, status, err, Pyramid|None)."""
        img1 = _gray(next_img)
        if img1.shape != prev_pyr.shape:
            raise Dr3lkError(E_SIZE, "(-215:Assertion failed) prevImg.size() == nextImg.size()")
        pp = np.ascontiguousarray(np.asarray(prev_pts, np.float32).reshape(-1, 2))
        n = pp.shape[0]
        if flags & USE_INITIAL_FLOW:
            npts = np.ascontiguousarray(np.asarray(next_pts, np.float32).reshape(-1, 2)).copy()
            if npts.shape[0] != n:
                raise Dr3lkError(E_ARG, "(-215:Assertion failed) nextPtsMat.checkVector(2, CV_32F, true) == npoints")
        else:
            npts = np.zeros((n, 2), np.float32)
        status = np.zeros(n, np.uint8)
        err = np.zeros(n, np.float32) if want_err else None
        win = prev_pyr.win
        out = ctypes.c_void_p()
        self._check(lib().dr3lk_track_frame(
            self._h, prev_pyr._h, img1.ctypes.data, img1.strides[0], pp.ctypes.data, npts.ctypes.data, status.ctypes.data,
            err.ctypes.data if want_err else None, n, win[0], win[1], max_level, criteria[0], criteria[1], float(criteria[2]), flags,
            float(min_eig_threshold), keep_next, ctypes.byref(out) if keep_next else None))
        return npts, status, err, (Pyramid._wrap(self, out, win, prev_pyr.shape) if keep_next else None)

    def filter_tracks(self, ref_pts, cur_pts, status, fx=None, fy=None, cx=0.0, cy=0.0, dist=None):
        """Reference src/initialization.cpp:615-635: drop status == 0 (order kept), disparity norms and, when a pinhole
        (fx, fy, cx, cy[, dist = (d0..d4)]) is given, unit bearing vectors of the current points (Pinhole::cam2world,
        src/camera.cpp:25-41, both branches). Returns (ref, cur, disparity, bearing|None)."""
        r = np.ascontiguousarray(np.asarray(ref_pts, np.float32).reshape(-1, 2))
        c = np.ascontiguousarray(np.asarray(cur_pts, np.float32).reshape(-1, 2))
        st = np.ascontiguousarray(np.asarray(status, np.uint8))
        n = r.shape[0]
        assert c.shape[0] == n and st.shape[0] == n
        o_r, o_c = np.zeros((n, 2), np.float32), np.zeros((n, 2), np.float32)
        disp = np.zeros(n, np.float64)
        bear = np.zeros((n, 3), np.float64) if fx is not None else None
        k = ctypes.c_int(0)
        d5 = np.ascontiguousarray(dist, np.float64) if dist is not None else None
        assert d5 is None or d5.shape == (5,)
        self._check(lib().dr3lk_filter_tracks(self._h, r.ctypes.data, c.ctypes.data, st.ctypes.data, n, float(fx or 1.0), float(fy or 1.0),
                                              float(cx), float(cy), d5.ctypes.data if d5 is not None else None,
                                              o_r.ctypes.data, o_c.ctypes.data, disp.ctypes.data,
                                              bear.ctypes.data if bear is not None else None, ctypes.byref(k)))
        k = k.value
        return o_r[:k], o_c[:k], disp[:k], (bear[:k] if bear is not None else None)

    def fast_detect(self, img, n_levels=3, cell_size=30, fast_threshold=20, detection_threshold=20.0, occupancy=None,
                    box_mode=BOX_AUTO_X86):
        """FastDetector::detect of the reference (src/features.cpp:43-98): returns (xy (n,2) int32 level-0 pixels,
        level (n,) int32, score (n,) float32) in grid-cell order."""
        img0 = np.ascontiguousarray(_gray(img))
        h, w = img0.shape
        ncell = (-(-w // cell_size)) * (-(-h // cell_size))
        xy, lv, sc = np.zeros((ncell, 2), np.int32), np.zeros(ncell, np.int32), np.zeros(ncell, np.float32)
        occ = np.ascontiguousarray(occupancy, np.uint8) if occupancy is not None else None
        n = ctypes.c_int(0)
        self._check(lib().dr3lk_fast_detect(self._h, img0.ctypes.data, w, h, img0.strides[0], n_levels, cell_size, fast_threshold,
                                            float(detection_threshold), box_mode, occ.ctypes.data if occ is not None else None,
                                            xy.ctypes.data, lv.ctypes.data, sc.ctypes.data, ctypes.byref(n)))
        return xy[:n.value].copy(), lv[:n.value].copy(), sc[:n.value].copy()

    def score_fundamental(self, F21, pts1, pts2, sigma=1.0, want_inliers=True):
        """InitHelper::CheckFundamental (src/initialization.cpp:171-249) for a batch of hypotheses.
        F21 (H,3,3) f32, pts (N,2) f32 -> (scores (H,) f32, inliers (H,N) u8 | None, best index)."""
        F = np.ascontiguousarray(np.asarray(F21, np.float32).reshape(-1, 9))
        a = np.ascontiguousarray(np.asarray(pts1, np.float32).reshape(-1, 2))
        b = np.ascontiguousarray(np.asarray(pts2, np.float32).reshape(-1, 2))
        assert a.shape == b.shape
        H, n = F.shape[0], a.shape[0]
        sc = np.zeros(H, np.float32)
        inl = np.zeros((H, n), np.uint8) if want_inliers else None
        best = ctypes.c_int(-1)
        self._check(lib().dr3lk_score_fundamental(self._h, F.ctypes.data, H, a.ctypes.data, b.ctypes.data, n, float(sigma), sc.ctypes.data,
                                                  inl.ctypes.data if want_inliers else None, ctypes.byref(best)))
        return sc, inl, best.value

    # ---- the two-frame initialiser front end, device-resident between its steps (src/initialization.cpp:546-661) ----
    def init_first_frame(self, img, n_levels=3, cell_size=30, fast_threshold=20, detection_threshold=20.0, occupancy=None,
                         box_mode=BOX_AUTO_X86, win=(30, 30), max_level=4):
        """Init::process_first_frame: one upload; FastDetector::detect on the Frame's box pyramid AND the frame's LK pyramid.
        Returns (xy, level, score, Pyramid)."""
        img0 = np.ascontiguousarray(_gray(img))
        h, w = img0.shape
        ncell = (-(-w // cell_size)) * (-(-h // cell_size))
        xy, lv, sc = np.zeros((ncell, 2), np.int32), np.zeros(ncell, np.int32), np.zeros(ncell, np.float32)
        occ = np.ascontiguousarray(occupancy, np.uint8) if occupancy is not None else None
        n, out = ctypes.c_int(0), ctypes.c_void_p()
        self._check(lib().dr3lk_init_first_frame(self._h, img0.ctypes.data, w, h, img0.strides[0], n_levels, cell_size, fast_threshold,
                                                 float(detection_threshold), box_mode, occ.ctypes.data if occ is not None else None,
                                                 win[0], win[1], max_level, xy.ctypes.data, lv.ctypes.data, sc.ctypes.data, ctypes.byref(n),
                                                 ctypes.byref(out)))
        return xy[:n.value].copy(), lv[:n.value].copy(), sc[:n.value].copy(), Pyramid._wrap(self, out, win, (h, w))

    def init_second_frame(self, ref_pyr, cur_img, kps_ref, kps_cur=None, max_level=4, criteria=(TERM_COUNT | TERM_EPS, 1000, 1e-3),
                          flags=USE_INITIAL_FLOW, min_eig_threshold=1e-4, fx=None, fy=None, cx=0.0, cy=0.0, dist=None):
        """Init::process_second_frame up to the RANSAC: one upload, LK against the reference pyramid, erase-by-status, disparities,
        bearings; one download.  Defaults are the reference's literal parameters (src/initialization.cpp:593-613).
        Returns dict(ref, cur, disparity, bearing, status, err)."""
        img1 = _gray(cur_img)
        if img1.shape != ref_pyr.shape:
            raise Dr3lkError(E_SIZE, "(-215:Assertion failed) prevImg.size() == nextImg.size()")
        r = np.ascontiguousarray(np.asarray(kps_ref, np.float32).reshape(-1, 2))
        n = r.shape[0]
        c = np.ascontiguousarray(np.asarray(kps_cur, np.float32).reshape(-1, 2)) if kps_cur is not None else None
        o_r, o_c = np.zeros((n, 2), np.float32), np.zeros((n, 2), np.float32)
        disp = np.zeros(n, np.float64)
        bear = np.zeros((n, 3), np.float64) if fx is not None else None
        st, err = np.zeros(n, np.uint8), np.zeros(n, np.float32)
        d5 = np.ascontiguousarray(dist, np.float64) if dist is not None else None
        k = ctypes.c_int(0)
        win = ref_pyr.win
        self._check(lib().dr3lk_init_second_frame(
            self._h, ref_pyr._h, img1.ctypes.data, img1.strides[0], r.ctypes.data, c.ctypes.data if c is not None else None, n, win[0], win[1],
            max_level, criteria[0], criteria[1], float(criteria[2]), flags, float(min_eig_threshold), float(fx or 0.0), float(fy or 0.0),
            float(cx), float(cy), d5.ctypes.data if d5 is not None else None, o_r.ctypes.data, o_c.ctypes.data, disp.ctypes.data,
            bear.ctypes.data if bear is not None else None, st.ctypes.data, err.ctypes.data, ctypes.byref(k)))
        k = k.value
        return {"ref": o_r[:k], "cur": o_c[:k], "disparity": disp[:k], "bearing": bear[:k] if bear is not None else None, "status": st, "err": err}

    def init_score_fundamental(self, F21, n_kept, sigma=1.0, want_inliers=True):
        """CheckFundamental for a batch of hypotheses against the tracks the last init_second_frame left on the device."""
        F = np.ascontiguousarray(np.asarray(F21, np.float32).reshape(-1, 9))
        H = F.shape[0]
        sc = np.zeros(H, np.float32)
        inl = np.zeros((H, n_kept), np.uint8) if want_inliers else None
        best = ctypes.c_int(-1)
        self._check(lib().dr3lk_init_score_fundamental(self._h, F.ctypes.data, H, float(sigma), sc.ctypes.data,
                                                       inl.ctypes.data if want_inliers else None, ctypes.byref(best)))
        return sc, inl, best.value

    def track_batch(self, prev_ptr, next_ptr, w, h, pitch, image_stride, batch, prev_pts_ptr, next_pts_ptr, status_ptr,
                    err_ptr, pts_offset, stats_ptr=None, win=(21, 21), max_level=3,
                    criteria=(TERM_COUNT | TERM_EPS, 30, 0.01), flags=0, min_eig_threshold=1e-4):
        """Device-resident batch (integer device pointers, e.g. torch tensor.data_ptr()). Asynchronous."""
        offs = np.ascontiguousarray(np.asarray(pts_offset, np.int32))
        assert offs.shape[0] == batch + 1
        self._check(lib().dr3lk_track_batch(
            self._h, prev_ptr, next_ptr, w, h, pitch, image_stride, batch, prev_pts_ptr, next_pts_ptr, status_ptr,
            err_ptr or None, offs.ctypes.data, stats_ptr or None, win[0], win[1], max_level, criteria[0], criteria[1],
            float(criteria[2]), flags, float(min_eig_threshold)))

    def track_batch_host(self, prev, nxt, prev_pts, pts_offset, next_pts=None, win=(21, 21), max_level=3,
                         criteria=(TERM_COUNT | TERM_EPS, 30, 0.01), flags=0, min_eig_threshold=1e-4, want_err=True,
                         want_stats=False, chunk_pairs=0, out=None):
        """Host-buffer batch: prev/nxt (B,H,W) uint8 arrays (any row step), prev_pts (N,2) f32, pts_offset (B+1,) int32.
        `out` may hold pre-allocated (next_pts, status, err, stats) arrays (e.g. pinned)."""
        prev, nxt = np.asarray(prev), np.asarray(nxt)
        assert prev.dtype == np.uint8 and prev.ndim == 3 and prev.shape == nxt.shape and prev.strides == nxt.strides
        assert prev.strides[2] == 1
        B, h, w = prev.shape
        offs = np.ascontiguousarray(np.asarray(pts_offset, np.int32))
        pp = np.asarray(prev_pts, np.float32).reshape(-1, 2)
        assert pp.flags.c_contiguous
        n = pp.shape[0]
        if out is not None:
            npts, status, err, stats = out
        else:
            npts = np.zeros((n, 2), np.float32)
            status = np.zeros(n, np.uint8)
            err = np.zeros(n, np.float32) if want_err else None
            stats = np.zeros(n, np.uint32) if want_stats else None
        if flags & USE_INITIAL_FLOW:
            npts[...] = np.asarray(next_pts, np.float32).reshape(-1, 2)
        self._check(lib().dr3lk_track_batch_host(
            self._h, prev.ctypes.data, nxt.ctypes.data, w, h, prev.strides[1], prev.strides[0], B, pp.ctypes.data,
            npts.ctypes.data, status.ctypes.data, err.ctypes.data if err is not None else None, offs.ctypes.data,
            stats.ctypes.data if stats is not None else None, chunk_pairs, win[0], win[1], max_level, criteria[0],
            criteria[1], float(criteria[2]), flags, float(min_eig_threshold)))
        return npts, status, err, stats


class MultiContext:
    """One context + one worker thread per listed device (dr3lk_multi): the host-buffer batch sharded by frame pair over
    several GPUs from ONE process, results bit-identical to a single-device call."""

    def __init__(self, devices):
        devs = (ctypes.c_int * len(devices))(*devices)
        self._h = ctypes.c_void_p()
        rc = lib().dr3lk_multi_create(ctypes.byref(self._h), devs, len(devices))
        if rc != OK:
            raise Dr3lkError(rc, lib().dr3lk_multi_last_error(None).decode())
        self.devices = list(devices)

    def close(self):
        if getattr(self, "_h", None):
            lib().dr3lk_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def launch_count(self):
        return sum(int(lib().dr3lk_launch_count(lib().dr3lk_multi_context(self._h, i))) for i in range(len(self.devices)))

    def track_batch_host(self, prev, nxt, prev_pts, pts_offset, next_pts=None, win=(21, 21), max_level=3,
                         criteria=(TERM_COUNT | TERM_EPS, 30, 0.01), flags=0, min_eig_threshold=1e-4, want_err=True,
                         want_stats=False, chunk_pairs=0, out=None):
        """Same arguments and results as Context.track_batch_host."""
        prev, nxt = np.asarray(prev), np.asarray(nxt)
        assert prev.dtype == np.uint8 and prev.ndim == 3 and prev.shape == nxt.shape and prev.strides == nxt.strides
        assert prev.strides[2] == 1
        B, h, w = prev.shape
        offs = np.ascontiguousarray(np.asarray(pts_offset, np.int32))
        pp = np.asarray(prev_pts, np.float32).reshape(-1, 2)
        assert pp.flags.c_contiguous
        n = pp.shape[0]
        if out is not None:
            npts, status, err, stats = out
        else:
            npts = np.zeros((n, 2), np.float32)
            status = np.zeros(n, np.uint8)
            err = np.zeros(n, np.float32) if want_err else None
            stats = np.zeros(n, np.uint32) if want_stats else None
        if flags & USE_INITIAL_FLOW:
            npts[...] = np.asarray(next_pts, np.float32).reshape(-1, 2)
        rc = lib().dr3lk_multi_track_batch_host(
            self._h, prev.ctypes.data, nxt.ctypes.data, w, h, prev.strides[1], prev.strides[0], B, pp.ctypes.data,
            npts.ctypes.data, status.ctypes.data, err.ctypes.data if err is not None else None, offs.ctypes.data,
            stats.ctypes.data if stats is not None else None, chunk_pairs, win[0], win[1], max_level, criteria[0],
            criteria[1], float(criteria[2]), flags, float(min_eig_threshold))
        if rc != OK:
            raise Dr3lkError(rc, lib().dr3lk_multi_last_error(self._h).decode())
        return npts, status, err, stats


def shard_range(n_pairs, rank, world):
    """dr3lk_shard_range: the contiguous block of pairs rank `rank` of `world` owns."""
    lo, hi = ctypes.c_int(), ctypes.c_int()
    lib().dr3lk_shard_range(n_pairs, rank, world, ctypes.byref(lo), ctypes.byref(hi))
    return lo.value, hi.value


class Pyramid:
    """Device-resident LK pyramid of one frame (Gaussian levels + Scharr derivatives), SURVEY.md 8(f-2)."""

    def __init__(self, ctx, img, win=(21, 21), max_level=3):
        img0 = _gray(img)
        h, w = img0.shape
        self._h = ctypes.c_void_p()
        self.ctx, self.win, self.shape = ctx, tuple(win), (h, w)
        ctx._check(lib().dr3lk_pyramid_create(ctx._h, img0.ctypes.data, w, h, img0.strides[0], win[0], win[1], max_level,
                                              ctypes.byref(self._h)))

    @classmethod
    def _wrap(cls, ctx, handle, win, shape):
        """Adopt a pyramid handle returned by the library (Context.track_frame)."""
        self = cls.__new__(cls)
        self._h, self.ctx, self.win, self.shape = handle, ctx, tuple(win), tuple(shape)
        return self

    @property
    def levels(self):
        return lib().dr3lk_pyramid_levels(self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib().dr3lk_pyramid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decode_stats(stats):
    """stats words -> (iterations, template_levels, err_pass) arrays (see dr3lk_track_batch)."""
    s = np.asarray(stats, np.uint32)
    return (s & 0xFFFF).astype(np.int64), ((s >> 16) & 0xFF).astype(np.int64), ((s >> 24) & 1).astype(np.int64)


def algorithmic_bytes(stats, win):
    """B_feat = sum_levels(5T + it_l*T) + T*[err pass] + 21, T = (win_w+1)(win_h+1); summed over features."""
    it, tl, ep = decode_stats(stats)
    T = (win[0] + 1) * (win[1] + 1)
    return int((5 * T * tl + T * it + T * ep + 21).sum())
