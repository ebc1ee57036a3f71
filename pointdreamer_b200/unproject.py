"""UNPROJECT operators — same names/arguments/returns as the reference's
pointdreamer/unproject.py (unproject 201-425, dilate_atlas 480-504), run by libpdr.so."""
import ctypes

import torch

from . import _lib
from . import camera as _camera
from .ours_utils import nearest_fill, _u8


def _per_view(x, V, n, dev):
    """uv_centers / uv_scales arrive as tensors ([V,1,2] / [V,1,1]) or the scalars 0 / 2."""
    if torch.is_tensor(x):
        return x.reshape(V, n).float().contiguous()
    return torch.full((V, n), float(x), device=dev)


def unproject(inpainted_images, vertices, f_normals, view_img_res, cams, cam_res, base_dirs,
              gb_pos, mask, per_atlas_pixel_face_id, uv_centers, uv_scales, padding,
              inpaint_scale_factors, mesh_normalized_depths, edge_dilate_kernels, save_img_path,
              complete_unseen_by_projection=False):
    """unproject.py:201-425.  `save_img_path` (debug PNG triptychs, 459-474) is ignored.

    Returns (atlas_img[R,R,3] f32, shrinked_vis[V,R,R] bool, point_view_ids[P] int64,
             points_atlas_pixel_coord[P,2] int64, points[P,3] f32, atlas_painted_mask[R,R] bool).
    """
    dev = vertices.device
    R = mask.shape[1]
    V = len(cams)
    res = int(view_img_res)
    rescale = (uv_scales is not None and uv_centers is not None and
               inpaint_scale_factors is not None and padding is not None)
    params = _camera.stack_params(cams, dev)
    mask_u8 = _u8(mask[0, :, :, 0]).contiguous()
    face_id = per_atlas_pixel_face_id[0].to(torch.int64).contiguous()
    gb = gb_pos[0].float().contiguous()
    fn = f_normals.float().contiguous()
    F = fn.shape[0]
    kernels = [int(k) for k in edge_dilate_kernels]
    n_levels = len(kernels)
    karr = (ctypes.c_int * n_levels)(*kernels)

    lib = _lib.load()
    ws_counter = torch.zeros(1, dtype=torch.int32, device=dev)
    count = ctypes.c_int(0)
    _lib.call("pdr_mask_count", mask_u8, ctypes.c_size_t(R * R), ws_counter,
              ctypes.byref(count))
    P = count.value

    lib.pdr_unproject_workspace_bytes.restype = ctypes.c_size_t
    ws = torch.empty(lib.pdr_unproject_workspace_bytes(R, n_levels), dtype=torch.uint8, device=dev)
    atlas = torch.empty(R, R, 3, device=dev)
    shr = torch.empty(V, R, R, dtype=torch.uint8, device=dev)
    view_ids = torch.empty(P, dtype=torch.int64, device=dev)
    coords = torch.empty(P, 2, dtype=torch.int64, device=dev)
    points = torch.empty(P, 3, device=dev)
    painted = torch.empty(R, R, dtype=torch.uint8, device=dev)
    if rescale:
        centers = _per_view(uv_centers, V, 2, dev)
        scales = _per_view(uv_scales, V, 1, dev)
        sfs = inpaint_scale_factors.float().contiguous()
        pad = float(padding)
    else:
        centers = scales = sfs = None
        pad = 0.0
    _lib.call("pdr_unproject", inpainted_images.float().contiguous(), res,
              params, V, int(cam_res), base_dirs.float().contiguous(),
              gb, mask_u8, face_id, R, fn, F,
              centers, scales, ctypes.c_double(pad), 1 if rescale else 0,
              sfs, mesh_normalized_depths.float().contiguous(), karr, n_levels,
              1 if complete_unseen_by_projection else 0, ws, atlas,
              shr, view_ids, coords, points,
              painted)
    return atlas, shr.bool(), view_ids, coords, points, painted.bool()


def dilate_atlas(atlas_img, mask):
    """unproject.py:480-504: nearest-fill the chart gutters.  atlas [R,R,3], mask [1,R,R,1]."""
    known = mask[..., 0] != 0  # [1,R,R]
    return nearest_fill(atlas_img[None], known, channels_last=True)[0]
