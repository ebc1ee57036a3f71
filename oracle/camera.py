"""Oracle (TEST INFRASTRUCTURE): camera rig and world -> NDC transform.

Restates
  * utils/camera_utils.py:86-102   fibonacci_sphere
  * utils/camera_utils.py:104-114  calculate_up_vector
  * utils/camera_utils.py:116-245  create_cameras (fibonacci_sphere distribution, distance 1.6,
                                   fov = pi/4, at = origin)
  * kaolin 0.15.0 `Camera.from_args(eye, at, up, fov, width, height).transform(points)`
    (call sites ours_utils.py:99, unproject.py:241).  kaolin is NOT vendored in the reference
    and not installable here: PARITY UNPINNED for this function.  The canonical arithmetic
    below is the published OpenGL look-at + pinhole model (near 1e-2, far 1e2).

Canonical fp32 arithmetic (every op a single IEEE binary32 operation, in this order):
    cx = ((r00*x + r01*y) + r02*z) + t0        (same for cy, cz with rows 1, 2)
    d  = -cz
    ndc_x = (cx * f) / d ;  ndc_y = (cy * f) / d ;  ndc_z = za - zb / d
with the 12 view-matrix entries, f = 1/tan(fov/2), za = (far+near)/(far-near) and
zb = 2*far*near/(far-near) computed in float64 on the host and rounded once to fp32.
"""
import math

import numpy as np

NEAR = 1e-2
FAR = 1e2


def fibonacci_sphere(samples, radius):
    """utils/camera_utils.py:86-102 (float64, same op order)."""
    points = []
    phi = math.pi * (3.0 - math.sqrt(5.0))
    for i in range(samples):
        y = 1 - (i / float(samples - 1)) * 2
        radius_y = math.sqrt(1 - y * y)
        theta = phi * i
        x = math.cos(theta) * radius_y * radius
        z = math.sin(theta) * radius_y * radius
        y = y * radius
        points.append((x, y, z))
    return np.array(points)


def calculate_up_vector(eye_position, target_position, world_up=None):
    """utils/camera_utils.py:104-114."""
    gaze_direction = target_position - eye_position
    if world_up is None:
        world_up = np.array([0, 1, 0])
    if np.allclose(np.cross(gaze_direction, world_up), 0):
        up_vector = np.array([0, 0, 1])
    else:
        side_vector = np.cross(gaze_direction, world_up)
        up_vector = np.cross(side_vector, gaze_direction)
        up_vector = up_vector / np.linalg.norm(up_vector)
    return up_vector


def view_params(eye, at, up, fov):
    """Look-at rotation/translation (float64 -> fp32) + pinhole constants.

    Returns a float32 array of 16 values: r00..r22 (row major), t0,t1,t2, f, za, zb, 0.
    """
    eye = np.asarray(eye, dtype=np.float64)
    at = np.asarray(at, dtype=np.float64)
    up = np.asarray(up, dtype=np.float64)
    backward = eye - at
    backward = backward / np.linalg.norm(backward)
    right = np.cross(up, backward)
    right = right / np.linalg.norm(right)
    up2 = np.cross(backward, right)
    R = np.stack([right, up2, backward], 0)
    t = -R @ eye
    f = 1.0 / math.tan(fov / 2.0)
    za = (FAR + NEAR) / (FAR - NEAR)
    zb = 2.0 * FAR * NEAR / (FAR - NEAR)
    out = np.zeros(16, dtype=np.float64)
    out[:9] = R.reshape(-1)
    out[9:12] = t
    out[12] = f
    out[13] = za
    out[14] = zb
    return out.astype(np.float32)


def transform(params, pts):
    """world (M,3) fp32 -> NDC (M,3) fp32 with the canonical op order."""
    p = np.asarray(params, dtype=np.float32)
    pts = np.asarray(pts, dtype=np.float32)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    cx = ((p[0] * x + p[1] * y) + p[2] * z) + p[9]
    cy = ((p[3] * x + p[4] * y) + p[5] * z) + p[10]
    cz = ((p[6] * x + p[7] * y) + p[8] * z) + p[11]
    d = -cz
    with np.errstate(divide="ignore", invalid="ignore"):
        nx = (cx * p[12]) / d
        ny = (cy * p[12]) / d
        nz = p[13] - p[14] / d
    return np.stack([nx, ny, nz], 1).astype(np.float32)


class Camera:
    """Stand-in for kaolin.render.camera.Camera with the attributes the path uses
    (`.transform`, `.width`, `.height`; ours_utils.py:99,142)."""

    def __init__(self, eye, at, up, fov, width, height):
        self.eye = np.asarray(eye, dtype=np.float64)
        self.at = np.asarray(at, dtype=np.float64)
        self.up = np.asarray(up, dtype=np.float64)
        self.fov = float(fov)
        self.width = int(width)
        self.height = int(height)
        self.params = view_params(eye, at, up, fov)

    def transform(self, pts):
        return transform(self.params, pts)


def create_cameras(num_views=8, distance=1.6, res=512):
    """utils/camera_utils.py:116-245 for distribution='fibonacci_sphere'.

    Returns (cameras, base_dirs[V,3] fp32, eye_positions (V,3) f64, up_dirs[V,3] fp32).
    """
    eye_positions = fibonacci_sphere(num_views, distance)
    cams = []
    base_dirs = np.zeros((num_views, 3), dtype=np.float32)
    up_dirs = np.zeros((num_views, 3), dtype=np.float32)
    fov = math.pi * 45 / 180
    for i, eye in enumerate(eye_positions):
        eye = np.array(eye)
        at = np.array([0, 0, 0])
        up = calculate_up_vector(eye, at)
        cams.append(Camera(eye, at, up, fov, res, res))
        base_dirs[i] = (eye - at).astype(np.float32)
        up_dirs[i] = np.asarray(up, dtype=np.float32)
    return cams, base_dirs, eye_positions, up_dirs
