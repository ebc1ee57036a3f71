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
    """Masks cross the ABI as u8.  torch.bool has the same 1-byte storage holding 0/1, so a bool
    tensor is reinterpreted in place (no conversion kernel); other dtypes are converted."""
    if t.dtype == torch.uint8:
        return t
    if t.dtype == torch.bool:
        return t.view(torch.uint8)
    return t.to(torch.uint8)


def _as_bool(t):
    """u8 0/1 tensor written by a kernel -> the reference's bool dtype, zero copy."""
    return t.view(torch.bool)


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
    lib = _lib.load()
    lib.pdr_rasterize_workspace_bytes.restype = ctypes.c_size_t
    ws = torch.empty(lib.pdr_rasterize_workspace_bytes(V, F, res), dtype=torch.uint8, device=dev)
    depth = torch.empty(V, res, res, device=dev)
    face_idx = torch.empty(V, res, res, dtype=torch.int64, device=dev)
    mask_cam = torch.empty(V, res, res, dtype=torch.uint8, device=dev)
    mask_out = torch.empty(V, out_res, out_res, dtype=torch.uint8, device=dev)
    _lib.call("pdr_rasterize", pos.contiguous(), f32, V, Vm, F, res, out_res,
              ws, depth, face_idx, mask_cam,
              mask_out)
    return _as_bool(mask_cam), face_idx, depth, _as_bool(mask_out)


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
    return _as_bool(out)


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
    return _as_bool(visib), pix


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


# ------------------------------------------------------------------------------------------
# "next" row N1: optimize_color (ours_utils.py:1583-1785)
# ------------------------------------------------------------------------------------------
def interpolate(attr, pos, faces, face_idx, attr_faces, flip_y=False, want_mask=False):
    """nvdiffrast.torch.interpolate stand-in on top of `rasterize` (call sites
    extract_texture_map.py:60, ours_utils.py:1705).  attr [Na,C] (C = 2 or 3), pos [V,Vm,4] and
    faces [F,3] as given to `rasterize`, face_idx [V,res,res] from it, attr_faces [F,3].
    Returns out [V,res,res,C] f32 (and the u8->bool coverage mask in the same frame)."""
    dev = pos.device
    V, Vm = pos.shape[0], pos.shape[1]
    res = face_idx.shape[1]
    C = attr.shape[-1]
    out = torch.empty(V, res, res, C, device=dev)
    mask = torch.empty(V, res, res, dtype=torch.uint8, device=dev) if want_mask else None
    _lib.call("pdr_interpolate", pos.contiguous(), faces.to(torch.int32).contiguous(),
              face_idx.contiguous(), attr.float().contiguous(),
              attr_faces.to(torch.int32).contiguous(), V, Vm, res, C, 1 if flip_y else 0, out, mask)
    return (out, _as_bool(mask)) if want_mask else out


def face_normals(vertices, faces):
    """kal.ops.mesh.face_normals(index_vertices_by_faces(v, f), unit=True) (demo.py:421-422)."""
    F = faces.shape[0]
    out = torch.empty(F, 3, device=vertices.device)
    _lib.call("pdr_face_normals", vertices.float().contiguous(),
              faces.to(torch.int32).contiguous(), F, out)
    return out


def optimize_color(atlas_img, inpainted_imgs, vertices, faces, uvs, mesh_tex_idx, cams,
                   eye_positions=None, look_ats=None, up_dirs=None, uv_centers=None,
                   uv_scales=None, padding=0.0, inpaint_scale_factors=None, glctx=None,
                   shrinked_per_view_per_pixel_visibility=None, lr=5e-2, iterations=100,
                   print_every=10, res=1024, return_images=True):
    """ours_utils.py:1583-1785, same arguments (eye_positions / look_ats / up_dirs / glctx are
    accepted and unused, as in the reference's nvdiffrast branch; `res` is the reference's
    hard-coded render size 1024).  atlas_img [3,R,R] (permuted + flipped by the caller,
    demo.py:217) or None (random init, 1024^2).  Returns (atlas [1,3,R,R] f32,
    images [V,3,res,res] f64 of the last iteration or None).

    The whole optimisation runs in libpdr.so (csrc/texopt.cu): rasterise + interpolate the
    texture uv per view pixel once, build the texel-major contribution list once, then
    `iterations` x (forward signs, gradient gather + Adam).  torch.sort / nonzero are used once
    for the list (setup plumbing); the Adam scalars follow torch.optim.Adam + StepLR(15, 0.5)."""
    import math
    dev = vertices.device
    V = len(cams)
    if atlas_img is not None:
        atlas = atlas_img.detach().float().contiguous().clone()
    else:
        atlas = torch.rand((3, 1024, 1024), dtype=torch.float, device=dev)
    R = atlas.shape[2]
    if atlas.shape[1] != R:
        raise ValueError("square atlas expected")
    Vm = vertices.shape[0]
    params = _camera.stack_params(cams, dev)
    if uv_centers is None or not torch.is_tensor(uv_centers):
        raise NotImplementedError("optimize_color needs the crop parameters of the PROJECT stage "
                                  "(crop_img: True in every shipped config)")
    isf = (inpaint_scale_factors if inpaint_scale_factors is not None
           else torch.ones(V, device=dev)).float().contiguous()
    pos = torch.empty(V, Vm, 4, device=dev)
    _lib.call("pdr_project_fixed", params, vertices.float().contiguous(), Vm, V,
              ctypes.c_double(float(padding)), uv_centers.float().contiguous(),
              uv_scales.float().contiguous(), isf, pos)
    _, face_idx, _, _ = rasterize(pos, faces, res, res)
    uv_map, mask = interpolate(uvs, pos, faces, face_idx, mesh_tex_idx, flip_y=True, want_mask=True)
    del face_idx
    vis = shrinked_per_view_per_pixel_visibility
    r0 = inpainted_imgs.shape[-1]
    n_pix = V * res * res
    active = torch.empty(V, res, res, dtype=torch.uint8, device=dev)
    target = torch.empty(V, res, res, 3, device=dev)
    keys = torch.empty(n_pix * 4, dtype=torch.int64, device=dev)
    _lib.call("pdr_texopt_prepare", uv_map, _u8(mask).contiguous(),
              _u8(vis).contiguous() if vis is not None else None,
              inpainted_imgs.float().contiguous(), int(r0), V, int(res), int(R), active, target, keys)
    keys, _ = torch.sort(keys)
    n_valid = int((keys != torch.iinfo(torch.int64).max).sum().item())
    keys = keys[:n_valid].contiguous()
    entry_pix = torch.empty(max(n_valid, 1), dtype=torch.int32, device=dev)
    entry_w = torch.empty(max(n_valid, 1), dtype=torch.float64, device=dev)
    head = torch.zeros(max(n_valid, 1), dtype=torch.uint8, device=dev)
    _lib.call("pdr_texopt_build", keys, ctypes.c_longlong(n_valid), uv_map, int(R), entry_pix,
              entry_w, head)
    seg_start = torch.nonzero(head[:n_valid]).flatten()
    n_seg = int(seg_start.numel())
    seg_start = torch.cat([seg_start, torch.tensor([n_valid], dtype=torch.int64, device=dev)])
    signs = torch.zeros(n_pix * 4, dtype=torch.int8, device=dev)
    m = torch.zeros_like(atlas)
    v = torch.zeros_like(atlas)
    images = torch.empty(V, 3, res, res, dtype=torch.float64, device=dev) if return_images else None
    beta1, beta2, eps = 0.9, 0.999, 1e-8
    for it in range(iterations):
        last = it == iterations - 1
        _lib.call("pdr_texopt_forward", atlas, uv_map, active, target, V, int(res), int(R), signs,
                  images if last else None)
        step = it + 1
        lr_t = lr * (0.5 ** (it // 15))                      # StepLR(step_size=15, gamma=0.5)
        bc1 = 1 - beta1 ** step
        bc2 = 1 - beta2 ** step
        _lib.call("pdr_texopt_step", atlas, m, v, keys, seg_start, ctypes.c_longlong(n_seg),
                  entry_pix, entry_w, signs, V, int(res), int(R), float(1 - beta1), float(beta2),
                  float(1 - beta2), float(math.sqrt(bc2)), float(eps), float(-(lr_t / bc1)))
    return atlas.unsqueeze(0), images
