"""Oracle (TEST INFRASTRUCTURE): UNPROJECT stage with Non-Border-First view selection.

Restates (reference file:line)
  * pointdreamer/unproject.py:201-425  unproject
  * pointdreamer/unproject.py:429-475  get_shrinked_per_view_per_pixel_visibility_torch
  * utils/utils_2d.py:799-827          detect_edges_in_gray_by_scharr_torch_batch
  * utils/utils_2d.py:833-845          dilate_torch_batch

The Scharr responses on {0,255} images are exact multiples of 127.5 whose smallest non-zero
value is 382.5, so both thresholds (>125 for the chart mask, >126.5 for per-view visibility)
reduce to "gx != 0 or gy != 0" in exact integer arithmetic.  Reflect padding
followed by a k x k max-pool equals a max over the window clamped to the image.
"""
import numpy as np

from . import camera as ocam
from .project import point_validation_by_depth

F32 = np.float32


def scharr_nonzero(img):
    """img [H,W] bool/int (0/1) -> bool: the zero-padded Scharr response (|gx|+|gy|)/2 != 0."""
    a = np.pad(img.astype(np.int64), 1)
    H, W = img.shape

    def s(dy, dx):
        return a[1 + dy:1 + dy + H, 1 + dx:1 + dx + W]

    gx = 3 * (s(-1, 1) - s(-1, -1)) + 10 * (s(0, 1) - s(0, -1)) + 3 * (s(1, 1) - s(1, -1))
    gy = 3 * (s(1, -1) - s(-1, -1)) + 10 * (s(1, 0) - s(-1, 0)) + 3 * (s(1, 1) - s(-1, 1))
    return (gx != 0) | (gy != 0)


def dilate_clamped(m, k):
    """dilate_torch_batch: reflect pad (k-1)//2 then max_pool2d(k) == clamped-window OR (odd k)."""
    if k % 2 != 1:
        raise ValueError("even dilation kernels change the output size in the reference")
    p = (k - 1) // 2
    H, W = m.shape
    # separable: rows then columns
    c = np.concatenate([np.zeros((H, 1), np.int64), np.cumsum(m.astype(np.int64), 1)], 1)
    lo = np.clip(np.arange(W) - p, 0, W)
    hi = np.clip(np.arange(W) + p + 1, 0, W)
    r = (c[:, hi] - c[:, lo]) > 0
    c = np.concatenate([np.zeros((1, W), np.int64), np.cumsum(r.astype(np.int64), 0)], 0)
    lo = np.clip(np.arange(H) - p, 0, H)
    hi = np.clip(np.arange(H) + p + 1, 0, H)
    return (c[hi, :] - c[lo, :]) > 0


def shrink_visibility(per_pixel_mask, vis, kernel_sizes):
    """unproject.py:429-458.  per_pixel_mask [R,R] bool, vis [R,R,V] bool ->
    [K,V,R,R] bool."""
    V = vis.shape[-1]
    if kernel_sizes[0] == 0:
        return vis.transpose(2, 0, 1)[None].copy()
    chart_edges = scharr_nonzero(per_pixel_mask)
    out = []
    view_edges = []
    for v in range(V):
        e = scharr_nonzero(vis[:, :, v]) & ~chart_edges
        view_edges.append(e)
    for k in kernel_sizes:
        per_view = []
        for v in range(V):
            border = dilate_clamped(view_edges[v], k)
            per_view.append(vis[:, :, v] & ~border)
        out.append(np.stack(per_view, 0))
    return np.stack(out, 0)


def softmax_rows(s):
    """torch.softmax(x, 1) in fp32: exp(x - max) / sum.  Canonical arithmetic (the library exp and
    the reduction order of torch's softmax are unpinned): exp evaluated in float64 and rounded
    once to fp32 (== the correctly rounded fp32 exp), the sum accumulated in fp32 in view order,
    one IEEE fp32 division - exactly what unproj_select_kernel does."""
    m = s.max(1, keepdims=True)
    e = np.exp((s - m).astype(F32).astype(np.float64)).astype(F32)
    tot = np.zeros((s.shape[0], 1), dtype=F32)
    for v in range(s.shape[1]):
        tot = (tot + e[:, v:v + 1]).astype(F32)
    return (e / tot).astype(F32)


def unproject(inpainted_images, f_normals, view_img_res, cam_params, cam_res, base_dirs, gb_pos,
              mask, per_atlas_pixel_face_id, uv_centers, uv_scales, padding, inpaint_scale_factors,
              mesh_normalized_depths, edge_dilate_kernels, complete_unseen_by_projection=False):
    """unproject.py:201-425.  numpy in/out with the reference's shapes:
    inpainted_images [V,3,res,res], gb_pos [1,R,R,3], mask [1,R,R,1] bool,
    per_atlas_pixel_face_id [1,R,R] int64.
    Returns (atlas_img[R,R,3], shrinked_vis[V,R,R] bool, point_view_ids[P] i64,
             points_atlas_pixel_coord[P,2] i64, points[P,3], atlas_painted_mask[R,R] bool)."""
    R = mask.shape[1]
    V = len(cam_params)
    ppm = mask[0, :, :, 0].astype(bool)
    face_id = per_atlas_pixel_face_id[0]
    coords = np.argwhere(ppm)  # row-major (row, col) == per_pixel_pixel_coord[mask]
    points = gb_pos[0][ppm].astype(F32)
    P = points.shape[0]

    tp = np.zeros((V, P, 3), dtype=F32)
    for i in range(V):
        tp[i] = ocam.transform(cam_params[i], points)
    depths = np.ascontiguousarray(tp[..., 2])
    uvs = tp[..., :2]
    if uv_scales is not None and uv_centers is not None and inpaint_scale_factors is not None \
            and padding is not None:
        uvs = (uvs - uv_centers) / uv_scales
        uvs_ns = uvs.copy()
        uvs = uvs * np.asarray(inpaint_scale_factors, dtype=F32)[:, None, None]
        pad_mul = F32(1 - 2 * padding)
        uvs = uvs * pad_mul
        uvs = uvs + F32(0.5)
        uvs_ns = uvs_ns * pad_mul
        uvs_ns = uvs_ns + F32(0.5)
    else:
        uvs = uvs * F32(0.5) + F32(0.5)
        uvs_ns = uvs
    vis, _ = point_validation_by_depth(cam_res, uvs_ns, depths, mesh_normalized_depths,
                                       offset=0.0001)  # [V,P]
    vis_atlas = np.zeros((R, R, V), dtype=bool)
    vis_atlas[ppm] = vis.T

    kernels = list(edge_dilate_kernels) * (R // 256)  # list repetition quirk, unproject.py:289
    per_kernel = shrink_visibility(ppm, vis_atlas, kernels)  # [K,V,R,R]

    normals = f_normals[face_id]  # -1 wraps to the last face, masked below (unproject.py:298)
    pn = normals[ppm].astype(F32)
    bd = np.asarray(base_dirs, dtype=F32)
    # canonical order of the 3-term dot product (torch matmul order is library-defined)
    sim = (pn[:, 0:1] * bd[None, :, 0] + pn[:, 1:2] * bd[None, :, 1]) + pn[:, 2:3] * bd[None, :, 2]

    pix = uvs * F32(view_img_res)
    pix = np.clip(pix, F32(0), F32(view_img_res - 1))
    pix = pix.astype(np.int64)
    pix = np.stack([pix[:, :, 1], pix[:, :, 0]], -1)  # V,P,(row,col)

    shr = per_kernel[0]
    cand = shr.transpose(1, 2, 0)[ppm]  # P,V
    cand = cand.copy()
    for i in range(1, len(edge_dilate_kernels)):
        left = cand.sum(1)
        nxt = per_kernel[i].transpose(1, 2, 0)[ppm]
        sel = left < 1
        cand[sel] = cand[sel] | nxt[sel]
        shr = per_kernel[i]
    if complete_unseen_by_projection:
        left = cand.sum(1)
        sel = left < 1
        cand[sel] = cand[sel] | vis.T[sel]

    w = softmax_rows(sim.astype(F32))
    w[~cand] = F32(-100)
    view_ids = w.argmax(1).astype(np.int64)  # first maximum
    if not complete_unseen_by_projection:
        view_ids[cand.sum(1) < 1] = -100

    atlas = np.zeros((R, R, 3), dtype=F32)
    painted = np.zeros((R, R), dtype=bool)
    for i in range(V):
        sel = view_ids == i
        img = inpainted_images[i][:, ::-1, :].transpose(1, 2, 0)  # flip rows, HWC
        atlas[coords[sel, 0], coords[sel, 1]] = img[pix[i][sel, 0], pix[i][sel, 1]]
        painted[coords[sel, 0], coords[sel, 1]] = True
    return atlas, shr, view_ids, coords.astype(np.int64), points, painted
