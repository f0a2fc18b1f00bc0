"""CPU restatement of the step right behind the LK call -- TEST INFRASTRUCTURE.

Follows /root/reference/src/initialization.cpp:615-635 (erase the points with status == 0 keeping order; disparity =
Vector2d(ref - cur).norm(); bearing = cam->cam2world(cur)) and /root/reference/src/camera.cpp:25-41 for an undistorted
Pinhole: ((u - cx)/fx, (v - cy)/fy, 1).normalized().  Parity unpinned: the reference cannot be compiled here (Eigen /
OpenCV absent) and has no test for this step; this restatement is plain IEEE double arithmetic in the order written.
"""
import numpy as np


def filter_tracks(ref_pts, cur_pts, status, fx=None, fy=None, cx=0.0, cy=0.0):
    ref = np.asarray(ref_pts, np.float32).reshape(-1, 2)
    cur = np.asarray(cur_pts, np.float32).reshape(-1, 2)
    keep = np.asarray(status) != 0
    r, c = ref[keep], cur[keep]
    d = (r - c).astype(np.float64)  # float subtraction (Point2f members), then promoted to double
    disp = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
    bear = None
    if fx is not None:
        x = (c[:, 0].astype(np.float64) - cx) / fx
        y = (c[:, 1].astype(np.float64) - cy) / fy
        nrm = np.sqrt((x * x + y * y) + 1.0)
        bear = np.stack([x / nrm, y / nrm, 1.0 / nrm], 1)
    return r, c, disp, bear


def check_fundamental(F21, pts1, pts2, sigma=1.0):
    """InitHelper::CheckFundamental (/root/reference/src/initialization.cpp:171-249) for a stack of hypotheses: fp32
    arithmetic in the order written there, the score accumulated match by match.  Returns (scores (H,), inliers (H,N)).
    Vectorised over the hypotheses only, so every hypothesis sees the reference's sequential accumulation order."""
    F = np.asarray(F21, np.float32).reshape(-1, 9)
    p1 = np.asarray(pts1, np.float32).reshape(-1, 2)
    p2 = np.asarray(pts2, np.float32).reshape(-1, 2)
    f11, f12, f13, f21, f22, f23, f31, f32, f33 = (F[:, i] for i in range(9))
    th, th_score = np.float32(3.841), np.float32(5.991)
    inv = np.float32(1.0 / float(np.float32(sigma) * np.float32(sigma)))
    score = np.zeros(F.shape[0], np.float32)
    inl = np.zeros((F.shape[0], p1.shape[0]), np.uint8)
    with np.errstate(all="ignore"):
        for i in range(p1.shape[0]):
            u1, v1, u2, v2 = p1[i, 0], p1[i, 1], p2[i, 0], p2[i, 1]
            a2 = f11 * u1 + f12 * v1 + f13
            b2 = f21 * u1 + f22 * v1 + f23
            c2 = f31 * u1 + f32 * v1 + f33
            num2 = a2 * u2 + b2 * v2 + c2
            chi1 = (num2 * num2 / (a2 * a2 + b2 * b2)) * inv
            ok1 = ~(chi1 > th)
            score = np.where(ok1, score + (th_score - chi1), score).astype(np.float32)
            a1 = f11 * u2 + f21 * v2 + f31
            b1 = f12 * u2 + f22 * v2 + f32
            c1 = f13 * u2 + f23 * v2 + f33
            num1 = a1 * u1 + b1 * v1 + c1
            chi2 = (num1 * num1 / (a1 * a1 + b1 * b1)) * inv
            ok2 = ~(chi2 > th)
            score = np.where(ok2, score + (th_score - chi2), score).astype(np.float32)
            inl[:, i] = ok1 & ok2
    return score, inl
