"""PROJECT + INPAINT-dispatch operators — same names, arguments and return tuples as the
reference's pointdreamer/ours_utils.py, executed by libpdr.so's sm_100a kernels.

Tensor arguments are PyTorch CUDA tensors (the reference's convention); outputs are allocated
here and handed to the C ABI as raw device pointers.  There is no host synchronisation inside
these operators and no CPU/PyTorch fallback.
"""
import ctypes

import torch

from . import _lib
from . import camera as _camera


def _u8(t):
    return t.to(torch.uint8) if t.dtype != torch.uint8 else t


def get_rendered_hard_mask_and_face_idx_batch(cams, vertices, faces, points, glctx=None,
                                              rescale=True, padding=0.05):
    """ours_utils.py:93-150.  `glctx` is accepted and ignored (the rasteriser is our own).

    Returns (hard_masks[V,H,W] bool, face_idxs[V,H,W] int64, mesh_normalized_depths[V,H,W] f32,
             vertice_uvs[V,Vm,2], uv_centers[V,1,2], uv_scales[V,1,1], padding,
             point_uvs[V,N,2], point_depths[V,N]);
    with rescale=False uv_centers/uv_scales/padding are the scalars 0 / 2 / 0 like the reference.
    """
    dev = vertices.device
    V = len(cams)
    Vm, N = vertices.shape[0], points.shape[0]
    H, W = cams[0].height, cams[0].width
    if H != W:
        raise ValueError("square cameras only (the reference asserts width == height)")
    params = _camera.stack_params(cams, dev)
    verts = vertices.float().contiguous()
    pts = points.float().contiguous()
    pos = torch.empty(V, Vm, 4, device=dev)
    vuv = torch.empty(V, Vm, 2, device=dev)
    centers = torch.empty(V, 1, 2, device=dev)
    scales = torch.empty(V, 1, 1, device=dev)
    puv = torch.empty(V, N, 2, device=dev)
    pdepth = torch.empty(V, N, device=dev)
    ws = torch.empty(4 * V, dtype=torch.int32, device=dev)
    _lib.call("pdr_project", params, verts, Vm, pts, N, V,
              1 if rescale else 0, ctypes.c_double(float(padding)), ws, pos,
              vuv, centers, scales, puv, pdepth)
    hard_masks, face_idxs, depths, _ = rasterize(pos, faces, H, H)
    if rescale:
        return hard_masks, face_idxs, depths, vuv, centers, scales, padding, puv, pdepth
    return hard_masks, face_idxs, depths, vuv, 0, 2, 0, puv, pdepth


def rasterize(pos, faces, res, out_res):
    """Mesh z-buffer (stands in for nvdiffrast.torch.rasterize, ours_utils.py:142-147) plus the
    mask at `out_res` (demo.py:103-104).  Returns (mask_cam bool, face_idx int64, depth f32,
    mask_out bool)."""
    dev = pos.device
    V, Vm = pos.shape[0], pos.shape[1]
    f32 = faces.to(torch.int32).contiguous()
    F = f32.shape[0]
    keys = torch.empty(V * res * res, dtype=torch.int64, device=dev)
    depth = torch.empty(V, res, res, device=dev)
    face_idx = torch.empty(V, res, res, dtype=torch.int64, device=dev)
    mask_cam = torch.empty(V, res, res, dtype=torch.uint8, device=dev)
    mask_out = torch.empty(V, out_res, out_res, dtype=torch.uint8, device=dev)
    _lib.call("pdr_rasterize", pos.contiguous(), f32, V, Vm, F, res, out_res,
              keys, depth, face_idx, mask_cam,
              mask_out)
    return mask_cam.bool(), face_idx, depth, mask_out.bool()


def resize_hard_masks(hard_masks, res):
    """demo.py:103-104 `transforms.Resize((res,res))(mask.float()).bool()` for the 2x case
    (bilinear without antialias == OR of each 2x2 block)."""
    V, H, _ = hard_masks.shape
    if H == res:
        return hard_masks
    if H != 2 * res:
        raise NotImplementedError("cam_res must equal res or 2*res")
    out = torch.empty(V, res, res, dtype=torch.uint8, device=hard_masks.device)
    _lib.call("pdr_mask_half_any", _u8(hard_masks).contiguous(), V, H, out)
    return out.bool()


def get_point_validation_by_depth(cam_res, point_uvs, point_depths, mesh_depths, offset=0,
                                  vis=False):
    """ours_utils.py:153-202 -> (point_visibility[V,N] bool, point_pixels[V,N,2] int64)."""
    V, N, _ = point_uvs.shape
    dev = point_uvs.device
    visib = torch.empty(V, N, dtype=torch.uint8, device=dev)
    pix = torch.empty(V, N, 2, dtype=torch.int64, device=dev)
    _lib.call("pdr_point_visibility", point_uvs.contiguous(),
              point_depths.contiguous(), mesh_depths.contiguous(), V, N,
              int(cam_res), float(offset), int(cam_res), visib, pix,
              None)
    return visib.bool(), pix


def get_point_pixels(point_uvs, res):
    """demo.py:121-125: (point_uvs*res).long(), swap x/y, clip(0,res-1) -> [V,N,2] int64."""
    V, N, _ = point_uvs.shape
    pix = torch.empty(V, N, 2, dtype=torch.int64, device=point_uvs.device)
    _lib.call("pdr_point_visibility", point_uvs.contiguous(), None,
              None, V, N, int(res), 0.0, int(res), None, None,
              pix)
    return pix


def get_point_validation_by_o3d(points, eye_positions=None, hidden_point_removal_radius=None):
    """ours_utils.py:204-225 (open3d hidden_point_removal) -> bool[V,N], computed on the GPU."""
    from .hpr import hidden_point_removal
    return hidden_point_removal(points, eye_positions, hidden_point_removal_radius)


def get_sparse_images(point_pixels, colors, point_validation, hard_masks, save_path, view_num,
                      res, point_size, edge_point_size, mask_ratio_thresh):
    """ours_utils.py:848-882 -> (sparse_imgs, hard_mask0s, hard_mask2s [V,3,res,res] f32,
    scale_factors[V] f32).  PNG dumps (`save_path`) are host-side IO outside the hot path."""
    dev = point_pixels.device
    V, N = view_num, point_pixels.shape[1]
    sparse = torch.empty(V, 3, res, res, device=dev)
    m0 = torch.empty(V, 3, res, res, device=dev)
    m2 = torch.empty(V, 3, res, res, device=dev)
    scales = torch.empty(V, device=dev)
    lib = _lib.load()
    lib.pdr_sparse_images_workspace_bytes.restype = ctypes.c_size_t
    nbytes = lib.pdr_sparse_images_workspace_bytes(V, res)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.call("pdr_sparse_images", point_pixels.contiguous(),
              colors.float().contiguous(), _u8(point_validation).contiguous(),
              _u8(hard_masks).contiguous(), V, N, int(res), int(point_size),
              int(edge_point_size), ctypes.c_double(float(mask_ratio_thresh)), ws,
              sparse, m0, m2, scales)
    if save_path is not None:
        from .io_utils import save_sparse_pngs
        save_sparse_pngs(sparse, m0, m2, save_path)
    return sparse, m0, m2, scales


def naive_inpainting(img, no_need_inpaint_mask2, method='linear'):
    """ours_utils.py:610-643 for method='nearest' on the GPU (exact nearest-valid-pixel fill).
    img [C,H,W], mask [C,H,W] (channel 0 used) -> [C,H,W] tensor on the same device."""
    if method != 'nearest':
        raise NotImplementedError("only method='nearest' is on the hot path (configs/nearest.yaml)")
    return nearest_fill(img[None], no_need_inpaint_mask2[None, 0] != 0)[0]


def nearest_fill(imgs, known, channels_last=False):
    """imgs [B,C,H,W] (or [B,H,W,C]) f32, known [B,H,W] bool -> filled, same layout."""
    dev = imgs.device
    imgs = imgs.float().contiguous()
    if channels_last:
        B, H, W, C = imgs.shape
    else:
        B, C, H, W = imgs.shape
    out = torch.empty_like(imgs)
    lib = _lib.load()
    lib.pdr_nearest_fill_workspace_bytes.restype = ctypes.c_size_t
    ws = torch.empty(lib.pdr_nearest_fill_workspace_bytes(B, H, W), dtype=torch.uint8, device=dev)
    _lib.call("pdr_nearest_fill", imgs, _u8(known).contiguous(), B, C, H, W,
              1 if channels_last else 0, ws, out, None)
    return out


def get_inpainted_images(sparse_imgs, hard_mask0s, hard_mask2s, save_path, inpainter, view_num,
                         method='linear'):
    """ours_utils.py:884-951.  'DDNM_inpaint' runs all views as ONE batched, host-sync-free
    sampler (the reference loops views serially at 914-929); 'nearest' is the exact GPU fill."""
    if method == 'DDNM_inpaint':
        inpainted = inpainter.inpaint_batch(sparse_imgs, hard_mask2s[:, 0])
    elif method == 'nearest':
        inpainted = nearest_fill(sparse_imgs, hard_mask2s[:, 0] != 0)
    else:
        raise NotImplementedError(
            f"texture_gen_method {method!r}: only 'DDNM_inpaint' and 'nearest' are on the hot path")
    if save_path is not None:
        from .io_utils import save_inpainted_pngs
        save_inpainted_pngs(inpainted, hard_mask0s, save_path, rgba=(method == 'DDNM_inpaint'))
    return inpainted
