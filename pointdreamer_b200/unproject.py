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
              complete_unseen_by_projection=False, num_texels=None):
    """unproject.py:201-425.  `save_img_path` (debug PNG triptychs, 459-474) is ignored.
    num_texels: P = mask.sum() when the caller already knows it (the chart mask is static per
    xatlas_dict; `count_texels` computes it once) - then the call has no host synchronisation.

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
    P = int(num_texels) if num_texels is not None else count_texels(mask)

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
    return atlas, shr.view(torch.bool), view_ids, coords, points, painted.view(torch.bool)


def count_texels(mask):
    """P = number of chart texels of the atlas mask [1,R,R,1] (one device reduction + one 4-byte
    read back; the only host synchronisation of the UNPROJECT stage, done once per atlas)."""
    mask_u8 = _u8(mask[0, :, :, 0]).contiguous()
    ws_counter = torch.zeros(1, dtype=torch.int32, device=mask.device)
    count = ctypes.c_int(0)
    _lib.call("pdr_mask_count", mask_u8, ctypes.c_size_t(mask_u8.numel()), ws_counter,
              ctypes.byref(count))
    return count.value


def dilate_atlas(atlas_img, mask):
    """unproject.py:480-504: nearest-fill the chart gutters.  atlas [R,R,3], mask [1,R,R,1]."""
    known = mask[..., 0] != 0  # [1,R,R]
    return nearest_fill(atlas_img[None], known, channels_last=True)[0]


def paint_invisible_areas_by_neighbors(vertices, faces, uvs, face_uv_idx, to_inpaint_face_id,
                                       atlas_img, atlas_inpainted_mask, use_atlas=True):
    """unproject.py:93-196 ("next" row N2): subdivide the never-seen faces twice, give every vertex
    the colour of its atlas texel, propagate colours from coloured to never-coloured vertices by
    neighbour averaging until nothing changes (then the same number of smoothing rounds), write
    the vertex colours back and nearest-fill the rest of the atlas.

    atlas_img [R,R,3] f32, atlas_inpainted_mask [R,R] bool (neither is modified; the reference
    mutates both).  Returns the atlas [R,R,3] f32 (the reference returns the same values as
    float64), or (vertices, faces, vertex_colors) of the subdivided mesh when use_atlas=False."""
    from .mesh_utils import subdivide_with_uv
    dev = vertices.device
    R = atlas_inpainted_mask.shape[1]
    v, f = vertices.float(), faces.long()
    uv, fuv = uvs.float(), face_uv_idx.long()
    ids = to_inpaint_face_id.long()
    for _ in range(2):  # the reference reuses the ORIGINAL face ids on the renumbered faces
        v, f, uv, fuv = subdivide_with_uv(v, f, fuv, uv, face_index=ids)
    Vn, F = v.shape[0], f.shape[0]
    # kaolin adjacency_matrix as CSR: unique directed edges, neighbours ascending
    r = torch.roll(f, 1, dims=-1)
    key = torch.unique(torch.cat([(f << 32) | r, (r << 32) | f]).reshape(-1))
    rows = key >> 32
    rowptr = torch.zeros(Vn + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=Vn), 0)
    rowptr = rowptr.to(torch.int32)
    colidx = (key & 0xFFFFFFFF).to(torch.int32)

    atlas = atlas_img.float().contiguous().clone()
    mask = _u8(atlas_inpainted_mask).contiguous().clone()
    pix = torch.empty(Vn, dtype=torch.int64, device=dev)
    colors = torch.empty(Vn, 3, device=dev)
    count = torch.empty(Vn, device=dev)
    has = torch.empty(Vn, dtype=torch.uint8, device=dev)
    ws = torch.empty(Vn, dtype=torch.int32, device=dev)
    _lib.call("pdr_vertex_colors", f.to(torch.int32).contiguous(), fuv.to(torch.int32).contiguous(),
              F, uv.contiguous(), Vn, atlas, mask, R, ws, pix, colors, count, has)
    cur = (colors, count)
    nxt = (colors.clone(), count.clone())
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    total = int(has.sum().item())
    coloring_round, stage = 0, "uncolored"
    while stage == "uncolored" or coloring_round > 0:
        _lib.call("pdr_laplacian_round", rowptr, colidx, Vn, has, cur[0], cur[1], nxt[0], nxt[1],
                  counter)
        cur, nxt = nxt, cur
        new_total = int(counter.item())
        if new_total > total:
            total = new_total
            coloring_round += 1
        else:
            stage = "colored"
            coloring_round -= 1
        if coloring_round > 10000:
            break
    vert_colors = cur[0]
    if torch.isnan(vert_colors).any():
        raise RuntimeError("paint_invisible_areas_by_neighbors: NaN vertex colour")
    if not use_atlas:
        return v, f, vert_colors
    winner = torch.empty(R * R, dtype=torch.int32, device=dev)
    _lib.call("pdr_scatter_vertex_colors", pix, vert_colors, Vn, R, winner, atlas, mask)
    return nearest_fill(atlas[None], mask[None] != 0, channels_last=True)[0]
