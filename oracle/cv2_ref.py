"""The real reference arithmetic: OpenCV's CPU calcOpticalFlowPyrLK via the cv2 wheel -- TEST INFRASTRUCTURE.

This is the code path the reference calls at /root/reference/src/initialization.cpp:608-613
(cv::calcOpticalFlowPyrLK, OpenCV un-vendored; here opencv-python-headless 4.13.0).  Used to pin the C
restatement (oracle/lk_oracle.c), to generate tests/golden/, and as the `--impl reference` /
cpu_baseline arm of bench.py.  Never imported by 3dr_b200/.
"""
import numpy as np

try:
    import cv2
    HAVE_CV2 = True
    CV2_VERSION = cv2.__version__
except Exception:  # pragma: no cover
    cv2 = None
    HAVE_CV2 = False
    CV2_VERSION = None


def build_lk_pyramid(img, win=(21, 21), max_level=3, with_derivatives=False):
    ml, pyr = cv2.buildOpticalFlowPyramid(img, tuple(win), max_level, withDerivatives=with_derivatives)
    pyr = [np.ascontiguousarray(p) for p in pyr]
    if with_derivatives:
        return pyr[0::2], pyr[1::2]
    return pyr


def calc_optical_flow_pyr_lk(prev, nxt, prev_pts, next_pts=None, win=(21, 21), max_level=3,
                             criteria=(3, 30, 0.01), flags=0, min_eig_threshold=1e-4):
    pp = np.ascontiguousarray(np.asarray(prev_pts, np.float32).reshape(-1, 1, 2))
    if pp.shape[0] == 0:
        return np.zeros((0, 2), np.float32), np.zeros(0, np.uint8), np.zeros(0, np.float32)
    init = None
    if flags & 4:
        init = np.ascontiguousarray(np.asarray(next_pts, np.float32).reshape(-1, 1, 2)).copy()
    p1, st, err = cv2.calcOpticalFlowPyrLK(prev, nxt, pp, init, winSize=tuple(win), maxLevel=max_level,
                                           criteria=tuple(criteria), flags=flags, minEigThreshold=min_eig_threshold)
    return p1.reshape(-1, 2), st.reshape(-1), err.reshape(-1)


def fast_corners(img, threshold=20):
    """prevPts provider used for the C1/C2 workloads: cv2 FAST-9/16 with NMS."""
    det = cv2.FastFeatureDetector_create(threshold, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(img, None)
    pts = np.array([k.pt for k in kps], np.float32).reshape(-1, 2)
    resp = np.array([k.response for k in kps], np.float32)
    return pts, resp
