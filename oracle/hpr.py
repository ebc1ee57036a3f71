"""Oracle (TEST INFRASTRUCTURE): hidden point removal.

Restates open3d's `PointCloud.hidden_point_removal(camera, radius)` as used by
pointdreamer/ours_utils.py:204-225 (open3d is NOT vendored and not installable here: PARITY
UNPINNED; this is Katz, Tal & Basri's published algorithm as open3d implements it): spherical flip
p' = p - eye, p^ = p' + 2 (R - |p'|) p'/|p'|, append the eye (origin), float64 Qhull convex hull,
visible = hull vertices other than the appended origin."""
import numpy as np


def hidden_point_removal(points, eye, radius):
    """points (N,3) -> sorted indices of the visible points."""
    from scipy.spatial import ConvexHull
    p = np.asarray(points, dtype=np.float64) - np.asarray(eye, dtype=np.float64)[None]
    n = np.linalg.norm(p, axis=1, keepdims=True)
    flipped = p + 2.0 * (radius - n) * p / n
    pts = np.concatenate([flipped, np.zeros((1, 3))], 0)
    hull = ConvexHull(pts)
    vid = np.unique(hull.vertices)
    return vid[vid < points.shape[0]]


def point_validation_by_o3d(points, eye_positions, radius):
    """ours_utils.py:204-225 -> bool[V,N]."""
    eyes = np.asarray(eye_positions, dtype=np.float64).reshape(-1, 3)
    vis = np.zeros((eyes.shape[0], points.shape[0]), dtype=bool)
    for i, eye in enumerate(eyes):
        vis[i, hidden_point_removal(points, eye, radius)] = True
    return vis
