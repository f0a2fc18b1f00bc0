"""CPU restatement of the step right behind the LK call -- TEST INFRASTRUCTURE.

Follows /root/reference/src/initialization.cpp:615-635 (erase the points with status == 0 keeping order; disparity =
Vector2d(ref - cur).norm(); bearing = cam->cam2world(cur)) and /root/reference/src/camera.cpp:25-41 for an undistorted
Pinhole: ((u - cx)/fx, (v - cy)/fy, 1).normalized(), and for a distorted one (src/camera.cpp:32-40) cv::undistortPoints
on the float pixel with float K / D.  Pinned: tests/test_oracle_vs_ref.py checks these functions bit for bit against
oracle/_ref (the reference's cam2world and CheckFundamental compiled unmodified); `undistort_points` -- OpenCV code, not in
the reference tree -- against the cv2 wheel (tests/test_oracle.py, tests/golden/undistort.npz).  What stays a
restatement: Eigen's Vector3d::normalized() ((x*x + y*y) + z*z, Eigen >= 3.3 with SSE2 packets; Eigen is absent here).
"""
import numpy as np


def undistort_points(uv, fx, fy, cx, cy, dist):
    """cv::undistortPoints(src 32FC2, dst 32FC2, K float 3x3, D float 1x5), no R / P, default criteria (5 iterations), as
    called at /root/reference/src/camera.cpp:36: OpenCV's cvUndistortPointsInternal in double, result rounded to float.
    Returns normalised coordinates (n, 2) float32."""
    uv = np.asarray(uv, np.float32).reshape(-1, 2)
    fx, fy, cx, cy = (float(np.float32(v)) for v in (fx, fy, cx, cy))
    k = [float(np.float32(v)) for v in dist]
    ifx, ify = 1.0 / fx, 1.0 / fy
    u, v = uv[:, 0].astype(np.float64), uv[:, 1].astype(np.float64)
    x = (u - cx) * ifx
    y = (v - cy) * ify
    x0, y0 = x.copy(), y.copy()
    alive = np.ones(len(x), bool)  # points that have not hit the icdist < 0 exit
    with np.errstate(all="ignore"):
        for _ in range(5):
            r2 = x * x + y * y
            icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
            bad = alive & (icdist < 0)
            dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + 0 * r2 + 0 * r2 * r2
            dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + 0 * r2 + 0 * r2 * r2
            nx, ny = (x0 - dx) * icdist, (y0 - dy) * icdist
            upd = alive & ~bad
            x = np.where(upd, nx, np.where(bad, x0, x))
            y = np.where(upd, ny, np.where(bad, y0, y))
            alive &= ~bad
    return np.stack([x.astype(np.float32), y.astype(np.float32)], 1)


def filter_tracks(ref_pts, cur_pts, status, fx=None, fy=None, cx=0.0, cy=0.0, dist=None):
    ref = np.asarray(ref_pts, np.float32).reshape(-1, 2)
    cur = np.asarray(cur_pts, np.float32).reshape(-1, 2)
    keep = np.asarray(status) != 0
    r, c = ref[keep], cur[keep]
    d = (r - c).astype(np.float64)  # float subtraction (Point2f members), then promoted to double
    disp = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
    bear = None
    if fx is not None:
        if dist is not None and abs(dist[0]) > 1e-7:  # Pinhole::_distortion, src/camera.cpp:17
            n = undistort_points(c, fx, fy, cx, cy, dist).astype(np.float64)
            x, y = n[:, 0], n[:, 1]
        else:
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
