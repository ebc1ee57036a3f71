"""Camera rig of the path (host side) — mirrors utils/camera_utils.py:86-245 of the reference.

`create_cameras` keeps the reference signature and return tuple
(cameras, base_dirs, eye_positions, up_dirs).  The camera objects expose what the hot path uses
of kaolin's Camera (`.transform`, `.width`, `.height`; ours_utils.py:99,142) plus `.params`, the
16 fp32 constants consumed by the CUDA kernels (layout in include/pdr.h:pdr_project).
Building the rig is float64 host arithmetic done once per run (reference: `prepare`,
demo.py:331-333); the per-point transform itself runs on the GPU.
"""
import math

import numpy as np
import torch

from . import _lib

NEAR = 1e-2
FAR = 1e2


def fibonacci_sphere(samples, radius):
    """utils/camera_utils.py:86-102."""
    points = []
    phi = math.pi * (3. - math.sqrt(5.))
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


def _view_params(eye, at, up, fov):
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
    out = np.zeros(16, dtype=np.float64)
    out[:9] = R.reshape(-1)
    out[9:12] = t
    out[12] = 1.0 / math.tan(fov / 2.0)
    out[13] = (FAR + NEAR) / (FAR - NEAR)
    out[14] = 2.0 * FAR * NEAR / (FAR - NEAR)
    return out.astype(np.float32)


class Camera:
    """Pinhole look-at camera with the subset of kaolin's Camera interface the path uses."""

    def __init__(self, eye, at, up, fov, width, height, device="cuda"):
        self.eye = np.asarray(eye, dtype=np.float64)
        self.fov = float(fov)
        self.width = int(width)
        self.height = int(height)
        self.device = torch.device(device)
        self.params_host = _view_params(eye, at, up, fov)
        self._params_dev = None

    @classmethod
    def from_args(cls, eye, at, up, fov, width, height, device="cuda", **kwargs):
        return cls(np.asarray(eye), np.asarray(at), np.asarray(up), fov, width, height, device)

    @property
    def params(self):
        if self._params_dev is None:
            self._params_dev = torch.from_numpy(self.params_host).to(self.device)
        return self._params_dev

    def transform(self, points):
        """world -> NDC on the GPU; (M,3) -> (M,3), (1,M,3) -> (1,M,3) like kaolin."""
        batched = points.dim() == 3
        p = points.reshape(-1, 3).float().contiguous()
        M = p.shape[0]
        dev = p.device
        pos = torch.empty(1, M, 4, device=dev)
        ws = torch.empty(4, dtype=torch.int32, device=dev)
        scratch = torch.empty(1, M, 2, device=dev)
        scratch2 = torch.empty(1, M, device=dev)
        cs = torch.empty(3, device=dev)
        import ctypes
        _lib.call("pdr_project", self.params.to(dev), p, M, p, M, 1, 0,
                  ctypes.c_double(0.0), ws, pos, scratch,
                  cs[:2], cs[2:], torch.empty(1, M, 2, device=dev),
                  scratch2)
        out = pos[0, :, :3].contiguous()
        return out.unsqueeze(0) if batched else out


def stack_params(cams, device):
    """[V,16] fp32 device tensor of the cameras' constants."""
    return torch.from_numpy(np.stack([c.params_host for c in cams])).to(device)


def create_cameras(num_views=8, distance=1.6, res=512, distribution='fibonacci_sphere',
                   device=torch.device('cuda'), vis=False):
    """utils/camera_utils.py:116-245 ('fibonacci_sphere' and 6-view 'self_defined' rigs).

    Returns (cameras, base_dirs[V,3] fp32 on `device`, eye_positions (V,3) numpy f64,
    up_dirs[V,3] fp32 on `device`)."""
    if distribution == 'fibonacci_sphere':
        eye_positions = fibonacci_sphere(num_views, distance)
    elif distribution == 'self_defined' and num_views == 6:
        eye_positions = distance * np.array([
            [0, 0, -1.0], [0, 0, 1.0], [0, -1.0, 0], [0, 1.0, 0], [-1.0, 0, 0], [1.0, 0, 0]])
    else:
        raise NotImplementedError(
            f"camera distribution {distribution!r} with {num_views} views is outside the hot "
            "path (configs/*.yaml all use fibonacci_sphere)")
    cameras = []
    base_dirs = torch.zeros((num_views, 3), dtype=torch.float)
    up_dirs = torch.zeros((num_views, 3), dtype=torch.float)
    fovy_angle = math.pi * 45 / 180
    for i, eye in enumerate(eye_positions):
        eye = np.array(eye)
        at = np.array([0, 0, 0])
        up = calculate_up_vector(eye, at)
        cameras.append(Camera.from_args(eye=eye, at=at, up=up, fov=fovy_angle, width=res,
                                        height=res, device=device))
        base_dirs[i] = torch.tensor(eye - at).float()
        up_dirs[i] = torch.tensor(up).float()
    return cameras, base_dirs.to(device), eye_positions, up_dirs.to(device)
