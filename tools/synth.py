"""Synthetic frame pairs for the throughput configs -- benchmark / test data only.

Recipe (numpy Generator(PCG64(seed = 1000 + pair))): sum of four unit-variance Gaussian-blurred white-noise fields
(sigma = 1, 2.5, 6, 16 px) on a (H+64)x(W+64) canvas, normalised, 128 + 48 z; the second frame is the canvas warped by
a small similarity about the centre (t ~ U(-6,6) px, rotation U(-0.5, 0.5) deg, scale 1 + U(-0.01, 0.01), bicubic,
REFLECT_101); both get independent N(0,1) noise, are rounded, clipped to uint8 and centre-cropped.
Corners: cv2.goodFeaturesToTrack(a, n_corners, 1e-4, min_dist).  Uses cv2 (available in this image); never imported by
the product package.
"""
import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def make_pair(seed, w=1241, h=376):
    """Returns (prev, next, M) with M the 2x3 affine map prev -> next coordinates (ground truth flow)."""
    assert cv2 is not None, "tools/synth.py needs cv2"
    rng = np.random.Generator(np.random.PCG64(seed))
    H, W = h + 64, w + 64
    z = np.zeros((H, W), np.float32)
    for sigma in (1.0, 2.5, 6.0, 16.0):
        f = rng.standard_normal((H, W)).astype(np.float32)
        f = cv2.GaussianBlur(f, (0, 0), sigma, borderType=cv2.BORDER_REFLECT_101)
        z += f / f.std()
    z = (z - z.mean()) / z.std()
    canvas = 128.0 + 48.0 * z
    tx, ty = rng.uniform(-6, 6, 2)
    ang = rng.uniform(-0.5, 0.5)
    sc = 1.0 + rng.uniform(-0.01, 0.01)
    M = cv2.getRotationMatrix2D((W / 2.0, H / 2.0), ang, sc)
    M[0, 2] += tx
    M[1, 2] += ty
    warped = cv2.warpAffine(canvas, M, (W, H), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REFLECT_101)
    a = canvas + rng.standard_normal((H, W)).astype(np.float32)
    b = warped + rng.standard_normal((H, W)).astype(np.float32)
    crop = (slice(32, 32 + h), slice(32, 32 + w))
    a8 = np.clip(np.rint(a[crop]), 0, 255).astype(np.uint8)
    b8 = np.clip(np.rint(b[crop]), 0, 255).astype(np.uint8)
    # ground truth in cropped coordinates: p' = M (p + 32) - 32
    Mc = M.copy()
    Mc[:, 2] = M[:, :2] @ np.array([32.0, 32.0]) + M[:, 2] - 32.0
    return np.ascontiguousarray(a8), np.ascontiguousarray(b8), Mc


def corners(img, n=8192, min_dist=3):
    pts = cv2.goodFeaturesToTrack(img, n, 1e-4, min_dist)
    return np.ascontiguousarray(pts.reshape(-1, 2).astype(np.float32))


def lattice(w, h, step=4, start=2):
    """C5: semi-dense lattice x = 2, 6, ...; y = 2, 6, ...  (310 x 94 = 29140 points at 1240/1241 x 376)."""
    xs, ys = np.meshgrid(np.arange(start, w, step, dtype=np.float32), np.arange(start, h, step, dtype=np.float32))
    return np.ascontiguousarray(np.stack([xs.ravel(), ys.ravel()], 1))


def make_batch(n_pairs, w=1241, h=376, n_corners=8192, seed0=1000, min_dist=3, points="corners"):
    """n_pairs distinct pairs. Returns prev (B,H,W) u8, next (B,H,W) u8, pts (N,2) f32, offsets (B+1,) i32."""
    prev = np.zeros((n_pairs, h, w), np.uint8)
    nxt = np.zeros((n_pairs, h, w), np.uint8)
    pts, offs = [], [0]
    for i in range(n_pairs):
        a, b, _ = make_pair(seed0 + i, w, h)
        prev[i], nxt[i] = a, b
        p = corners(a, n_corners, min_dist) if points == "corners" else lattice(w, h)
        pts.append(p)
        offs.append(offs[-1] + len(p))
    return prev, nxt, np.concatenate(pts).astype(np.float32), np.array(offs, np.int32)
