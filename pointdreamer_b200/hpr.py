"""Hidden point removal (ours_utils.py:204-225, open3d `hidden_point_removal`) on the GPU.

The reference copies the cloud to the host and runs a float64 Qhull convex hull per view; here
every point is tested for hull-vertex membership by a small fp64 LP on the device
(csrc/geom_hpr.cu).  No CPU path exists."""
import ctypes

import numpy as np
import torch

from . import _lib


def view_frames(eye_positions, at=None):
    """[V,12] float64: eye, ex, ey, ez per view; ez points from the eye to the scene centre."""
    eyes = np.asarray(eye_positions, dtype=np.float64).reshape(-1, 3)
    at = np.zeros(3) if at is None else np.asarray(at, dtype=np.float64)
    out = np.zeros((eyes.shape[0], 12), dtype=np.float64)
    for i, eye in enumerate(eyes):
        ez = at - eye
        ez = ez / np.linalg.norm(ez)
        a = np.array([1.0, 0.0, 0.0]) if abs(ez[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
        ex = np.cross(a, ez)
        ex = ex / np.linalg.norm(ex)
        ey = np.cross(ez, ex)
        out[i] = np.concatenate([eye, ex, ey, ez])
    return out


def hidden_point_removal(points, eye_positions, radius):
    """points [N,3] cuda tensor, eye_positions (V,3) array-like, radius -> bool[V,N] visibility."""
    dev = points.device
    pts = points.float().contiguous()
    N = pts.shape[0]
    frames = torch.from_numpy(view_frames(eye_positions)).to(dev)
    V = frames.shape[0]
    lib = _lib.load()
    lib.pdr_hidden_point_removal_workspace_bytes.restype = ctypes.c_size_t
    ws = torch.empty(lib.pdr_hidden_point_removal_workspace_bytes(V, N), dtype=torch.uint8,
                     device=dev)
    vis = torch.empty(V, N, dtype=torch.uint8, device=dev)
    _lib.call("pdr_hidden_point_removal", pts, N, V, frames, ctypes.c_double(float(radius)), ws,
              vis)
    return vis.view(torch.bool)
