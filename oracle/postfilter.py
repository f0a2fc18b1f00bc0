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
