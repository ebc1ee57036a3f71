"""Oracle (TEST INFRASTRUCTURE): PROJECT stage — crop/rescale, mesh rasterisation, point
visibility, point splat, masks.

Restates (reference file:line under pointdreamer/ unless noted)
  * ours_utils.py:93-150   get_rendered_hard_mask_and_face_idx_batch
  * nvdiffrast.torch.rasterize (ours_utils.py:142) — NOT vendored, PARITY UNPINNED; the
    canonical rule is documented at `rasterize`
  * demo.py:103-104        512 -> 256 mask resize (bilinear, no antialias, != 0  == 2x2 any)
  * ours_utils.py:153-202  get_point_validation_by_depth
  * demo.py:121-125        point_pixels
  * ours_utils.py:456-495  paint_pixels   (deterministic winner = highest index)
  * ours_utils.py:497-532  get_forground_inner_edge_mask ('dilate')
  * ours_utils.py:954-1044 get_one_sparse_img
  * ours_utils.py:848-882  get_sparse_images
  * kaolin.metrics.pointcloud.sided_distance (ours_utils.py:1013) — NOT vendored, UNPINNED;
    canonical rule: exact squared distance on integer pixel coords, lowest index on ties.
"""
import numpy as np

from . import camera as ocam

F32 = np.float32
SUBPIX = 256  # fixed-point sub-pixel resolution of the canonical rasteriser


# ----------------------------------------------------------------------------------------
# P1: transform + crop/rescale           ours_utils.py:93-130
# ----------------------------------------------------------------------------------------
def project_vertices_points(cam_params, vertices, points, rescale=True, padding=0.05):
    """Returns dict with pos[V,Vm,4], vertice_uvs[V,Vm,2], uv_centers[V,1,2], uv_scales[V,1,1],
    padding, point_uvs[V,N,2], point_depths[V,N]  (all fp32)."""
    V = len(cam_params)
    Vm, N = vertices.shape[0], points.shape[0]
    pos = np.zeros((V, Vm, 4), dtype=F32)
    tp = np.zeros((V, N, 3), dtype=F32)
    for i in range(V):
        tv = ocam.transform(cam_params[i], vertices)
        tp[i] = ocam.transform(cam_params[i], points)
        pos[i, :, :3] = tv
        pos[i, :, 3] = 1.0
    if rescale:
        vuv = pos[:, :, :2]
        mn = vuv.min(1)[:, None, :]  # V,1,2
        mx = vuv.max(1)[:, None, :]
        uv_centers = (mn + mx) / F32(2)
        uv_scales = (mx - mn).max(2)[:, :, None]  # V,1,1
        pad_mul = F32(1 - 2 * padding)
        vuv = (vuv - uv_centers) / uv_scales
        vuv = vuv * pad_mul
        vuv = vuv + F32(0.5)
        vuv = np.clip(vuv, F32(0), F32(1))
        pos[:, :, :2] = vuv * F32(2) - F32(1)
        puv = tp[..., :2]
        puv = (puv - uv_centers) / uv_scales
        puv = puv * pad_mul
        puv = puv + F32(0.5)
        pdepth = tp[:, :, 2]
        pad_out = padding
    else:
        vuv = (pos[:, :, :2] + F32(1)) * F32(0.5)
        vuv = np.clip(vuv, F32(0), F32(1))
        puv = (tp[..., :2] + F32(1)) * F32(0.5)
        uv_centers = np.zeros((V, 1, 2), dtype=F32)
        uv_scales = np.full((V, 1, 1), 2, dtype=F32)
        pad_out = 0
        pdepth = tp[:, :, 2]
    return dict(pos=pos, vertice_uvs=vuv.astype(F32), uv_centers=uv_centers.astype(F32),
                uv_scales=uv_scales.astype(F32), padding=pad_out, point_uvs=puv.astype(F32),
                point_depths=np.ascontiguousarray(pdepth.astype(F32)))


# ----------------------------------------------------------------------------------------
# P2: canonical rasteriser (stands in for nvdiffrast.torch.rasterize)
# ----------------------------------------------------------------------------------------
def _snap(ndc, res):
    """NDC coordinate -> fixed-point pixel coordinate (1/256 px), round half up."""
    s = ((ndc + F32(1)) * F32(0.5)) * F32(res)
    return np.floor(s * F32(SUBPIX) + F32(0.5)).astype(np.int64)


def _edge_inclusive(dx, dy):
    """Fill rule for samples exactly on an edge (A->B = (dx,dy)), interior on the e>0 side:
    the edge owns the sample iff dy > 0 or (dy == 0 and dx < 0).  For a shared edge the two
    adjacent triangles traverse it in opposite directions, so exactly one of them owns it."""
    return (dy > 0) or (dy == 0 and dx < 0)


def rasterize(pos, faces, res, return_bary=False):
    """Canonical z-buffer rasteriser.

      * vertex xy snapped to a 1/256-pixel grid; sample point = pixel centre (col+0.5,row+0.5);
        row 0 is NDC y = -1 (nvdiffrast's bottom-up convention);
      * exact int64 edge functions; triangles of either orientation are drawn (no culling);
        zero-area triangles are skipped; samples on an edge follow `_edge_inclusive`;
      * depth = ((eA*zA + eB*zB) + eC*zC) / (eA+eB+eC) in fp32 (weights converted RN from
        int64); fragments with depth outside [-1, 1] are discarded;
      * nearest depth wins, equal depth -> lowest triangle index.

    pos [V,Vm,4] fp32 (w == 1), faces [F,3] int.  Returns depth[V,res,res] fp32 (0 empty),
    face_idx[V,res,res] int64 (-1 empty), mask[V,res,res] bool; with return_bary also the fp32
    barycentrics bary[V,res,res,2] = (eA/tot, eB/tot) of the winning triangle's vertices 0 and 1
    (nvdiffrast's rast[..., 0:2]; 0 where empty).
    """
    V = pos.shape[0]
    faces = np.asarray(faces, dtype=np.int64)
    depth = np.zeros((V, res, res), dtype=F32)
    fidx = -np.ones((V, res, res), dtype=np.int64)
    zbuf = np.full((V, res, res), np.inf, dtype=F32)
    bary = np.zeros((V, res, res, 2), dtype=F32) if return_bary else None
    for v in range(V):
        X = _snap(pos[v, :, 0], res)
        Y = _snap(pos[v, :, 1], res)
        Z = pos[v, :, 2].astype(F32)
        for f in range(faces.shape[0]):
            ia, ib, ic = faces[f]
            ax, ay, bx, by, cx, cy = X[ia], Y[ia], X[ib], Y[ib], X[ic], Y[ic]
            area = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)
            if area == 0:
                continue
            sgn = 1 if area > 0 else -1
            xmin = max((min(ax, bx, cx) - SUBPIX // 2 + SUBPIX - 1) // SUBPIX, 0)
            xmax = min((max(ax, bx, cx) - SUBPIX // 2) // SUBPIX, res - 1)
            ymin = max((min(ay, by, cy) - SUBPIX // 2 + SUBPIX - 1) // SUBPIX, 0)
            ymax = min((max(ay, by, cy) - SUBPIX // 2) // SUBPIX, res - 1)
            if xmin > xmax or ymin > ymax:
                continue
            px = (np.arange(xmin, xmax + 1, dtype=np.int64) * SUBPIX + SUBPIX // 2)[None, :]
            py = (np.arange(ymin, ymax + 1, dtype=np.int64) * SUBPIX + SUBPIX // 2)[:, None]
            # weight of A = edge(B,C,P), of B = edge(C,A,P), of C = edge(A,B,P)
            eA = sgn * ((cx - bx) * (py - by) - (cy - by) * (px - bx))
            eB = sgn * ((ax - cx) * (py - cy) - (ay - cy) * (px - cx))
            eC = sgn * ((bx - ax) * (py - ay) - (by - ay) * (px - ax))
            inA = (eA > 0) | ((eA == 0) & _edge_inclusive(sgn * (cx - bx), sgn * (cy - by)))
            inB = (eB > 0) | ((eB == 0) & _edge_inclusive(sgn * (ax - cx), sgn * (ay - cy)))
            inC = (eC > 0) | ((eC == 0) & _edge_inclusive(sgn * (bx - ax), sgn * (by - ay)))
            inside = inA & inB & inC
            if not inside.any():
                continue
            wa, wb, wc = eA.astype(F32), eB.astype(F32), eC.astype(F32)
            tot = (eA + eB + eC).astype(F32)
            z = ((wa * Z[ia] + wb * Z[ib]) + wc * Z[ic]) / tot
            ok = inside & (z >= F32(-1)) & (z <= F32(1))
            zb = zbuf[v, ymin:ymax + 1, xmin:xmax + 1]
            fb = fidx[v, ymin:ymax + 1, xmin:xmax + 1]
            win = ok & ((z < zb) | ((z == zb) & (f < fb)))
            zb[win] = z[win]
            fb[win] = f
            if return_bary:
                bb = bary[v, ymin:ymax + 1, xmin:xmax + 1]
                bb[win, 0] = (wa / tot)[win]
                bb[win, 1] = (wb / tot)[win]
    mask = fidx >= 0
    depth[mask] = zbuf[mask]
    if return_bary:
        return depth, fidx, mask, bary
    return depth, fidx, mask


def interpolate(bary, fidx, attr, attr_faces):
    """nvdiffrast.torch.interpolate (call sites models/get3d/extract_texture_map.py:60,
    ours_utils.py:1705), canonical fp32 rule: out = (u*a0 + v*a1) + ((1 - u) - v)*a2 with (u, v)
    the rasteriser's barycentrics; 0 where the pixel is empty.
    bary [V,H,W,2] fp32, fidx [V,H,W] int64, attr [Na,C] fp32, attr_faces [F,3] int."""
    attr = np.asarray(attr, dtype=F32)
    attr_faces = np.asarray(attr_faces, dtype=np.int64)
    hit = fidx >= 0
    f = np.where(hit, fidx, 0)
    a0, a1, a2 = attr[attr_faces[f, 0]], attr[attr_faces[f, 1]], attr[attr_faces[f, 2]]
    u, v = bary[..., 0:1], bary[..., 1:2]
    b2 = (F32(1) - u) - v
    out = (u * a0 + v * a1) + b2 * a2
    return np.where(hit[..., None], out, F32(0)).astype(F32)


def resize_mask_half_any(mask, res):
    """demo.py:103-104: bilinear (antialias=False) resize of the float mask followed by
    `.bool()`.  For an exact 2x reduction every output pixel averages its 2x2 block."""
    V, H, W = mask.shape
    if H == res:
        return mask.copy()
    assert H == 2 * res and W == 2 * res, "only the cam_res == 2*res case is restated exactly"
    m = mask.reshape(V, res, 2, res, 2)
    return m.any(axis=(2, 4))


# ----------------------------------------------------------------------------------------
# P4: depth visibility                   ours_utils.py:153-202
# ----------------------------------------------------------------------------------------
def point_validation_by_depth(cam_res, point_uvs, point_depths, mesh_depths, offset=0.0):
    pp = point_uvs * F32(cam_res)
    pp = np.clip(pp, F32(0), F32(cam_res - 1))
    pp = pp.astype(np.int64)  # trunc toward zero (values are >= 0)
    pix = np.stack([pp[:, :, 1], pp[:, :, 0]], -1)  # (row, col)
    V = point_uvs.shape[0]
    ref = mesh_depths[np.arange(V)[:, None], pix[:, :, 0], pix[:, :, 1]]
    vis = (point_depths - ref) <= F32(offset)
    return vis, pix


def point_pixels(point_uvs, res):
    """demo.py:121-125 (long() BEFORE clip)."""
    pp = (point_uvs * F32(res)).astype(np.int64)
    pp = np.stack([pp[:, :, 1], pp[:, :, 0]], -1)
    return np.clip(pp, 0, res - 1)


# ----------------------------------------------------------------------------------------
# P5..P8: splat + masks                  ours_utils.py:456-532, 848-1044
# ----------------------------------------------------------------------------------------
def paint_pixels(img, coords, colors, point_size):
    """ours_utils.py:456-495 with the deterministic index_put rule: the write with the highest
    flattened index wins."""
    C = img.shape[0]
    n = coords.shape[0]
    if np.isscalar(colors):
        colors = np.full((n, C), colors, dtype=F32)
    if point_size == 1:
        img[:, coords[:, 0], coords[:, 1]] = colors.T  # numpy: last duplicate wins
    else:
        s = point_size
        off = np.arange(-s + 1, s)
        gx, gy = np.meshgrid(off, off, indexing="ij")
        grid = np.stack([gx, gy], 2)[None] + coords[:, None, None, :]  # n,g,g,2
        cols = np.broadcast_to(colors[:, None, None, :], grid.shape[:3] + (C,))
        m = (grid[..., 0] >= 0) & (grid[..., 0] < img.shape[1]) & (grid[..., 1] >= 0) & \
            (grid[..., 1] < img.shape[2])
        g = grid[m]
        c = cols[m]
        img[:, g[:, 0], g[:, 1]] = c.T
    return img


def inner_edge_mask(fg):
    """ours_utils.py:519-522: maxpool3x3(~fg) & fg, pool padding = -inf (border is not bg)."""
    bg = ~fg
    H, W = fg.shape
    p = np.zeros((H + 2, W + 2), dtype=bool)
    p[1:-1, 1:-1] = bg
    dil = np.zeros_like(fg)
    for dy in range(3):
        for dx in range(3):
            dil |= p[dy:dy + H, dx:dx + W]
    return dil & fg


def bilinear_resize_any(mask, out_res):
    """transforms.Resize((out,out)) on a bool [1,H,W] tensor then `.bool()`
    (ours_utils.py:989-995; torchvision casts to fp32, bilinear, align_corners=False,
    antialias=False per SURVEY §8a P3, no rounding for bool outputs): a destination pixel is
    True iff a source pixel with non-zero interpolation weight is True."""
    H, W = mask.shape
    assert H == W
    scale = F32(H) / F32(out_res)  # area_pixel_compute_scale<float>
    dst = np.arange(out_res, dtype=F32)
    src = scale * (dst + F32(0.5)) - F32(0.5)
    src = np.maximum(src, F32(0))
    i0 = np.minimum(src.astype(np.int64), H - 1)
    i1 = np.minimum(i0 + 1, H - 1)
    lam = np.clip(src - i0.astype(F32), F32(0), F32(1))
    w1 = lam > 0  # weight of i1; weight of i0 = 1 - lam > 0 always (lam < 1)
    # value(y,x) = sum over the 2x2 footprint; non-zero iff any contributing source is set
    r0 = mask[i0][:, i0]
    r01 = mask[i0][:, i1] & w1[None, :]
    r10 = mask[i1][:, i0] & w1[:, None]
    r11 = mask[i1][:, i1] & w1[None, :] & w1[:, None]
    return r0 | r01 | r10 | r11


def nearest_valid_point(edge_px, valid_px):
    """canonical sided_distance: argmin of exact squared pixel distance, lowest index on ties."""
    if edge_px.shape[0] == 0:
        return np.zeros((0,), dtype=np.int64)
    if valid_px.shape[0] == 0:
        raise ValueError("view has no valid point (the reference fails here too)")
    out = np.empty(edge_px.shape[0], dtype=np.int64)
    vp = valid_px.astype(np.int64)
    for s in range(0, edge_px.shape[0], 512):
        e = edge_px[s:s + 512].astype(np.int64)
        d = ((e[:, None, :] - vp[None, :, :]) ** 2).sum(-1)
        out[s:s + 512] = d.argmin(1)  # first minimum = lowest index
    return out


def get_one_sparse_img(point_pixels_v, colors, valid, hard_mask, res, point_size,
                       edge_point_size, mask_ratio_thresh=0.82):
    """ours_utils.py:954-1044.  Returns sparse_img[3,res,res], hard_mask0, hard_mask2 (fp32),
    mask_ratio (fp32), scale_factor (fp32)."""
    fg_num = F32(hard_mask.sum())
    valid_num = int(valid.sum())
    mask_ratio = F32(1) - F32(valid_num) / fg_num
    pp = point_pixels_v
    if mask_ratio > mask_ratio_thresh:
        wanted = F32(valid_num) / F32(1 - mask_ratio_thresh)
        scale = wanted / fg_num
        uv = pp.astype(F32) / F32(res)
        uv = uv * F32(2) - F32(1)
        uv = uv * scale
        uv = (uv + F32(1)) * F32(0.5)
        ppf = uv * F32(res)
        ppf = np.clip(ppf, F32(0), F32(res - 1))
        pp = ppf.astype(np.int64)
        after = int(np.floor(F32(res) * scale))
        if (res - after) % 2 == 1:
            after += 1
        pad = int((res - after) / 2)
        small = bilinear_resize_any(hard_mask, after)
        hard_mask = np.zeros((res, res), dtype=bool)
        hard_mask[pad:pad + after, pad:pad + after] = small
    else:
        scale = F32(1)
    sparse = np.zeros((3, res, res), dtype=F32)
    vpix = pp[valid]
    vcol = colors[valid]
    sparse = paint_pixels(sparse, vpix, vcol, point_size)
    edge = inner_edge_mask(hard_mask)
    epix = np.argwhere(edge)
    idx = nearest_valid_point(epix, vpix)
    ecol = vcol[idx]
    sparse = paint_pixels(sparse, epix, ecol, edge_point_size)
    m0 = np.repeat(hard_mask[None].astype(F32), 3, 0)
    m2 = F32(1) - m0
    m2 = paint_pixels(m2, vpix, 1.0, point_size)
    m2 = paint_pixels(m2, epix, 1.0, edge_point_size)
    occupied = F32((m2[0] * hard_mask).sum())
    ratio_out = F32(1) - occupied / F32(hard_mask.sum())
    return sparse[:, ::-1].copy(), m0[:, ::-1].copy(), m2[:, ::-1].copy(), ratio_out, F32(scale)


def get_sparse_images(point_pixels_all, colors, point_validation, hard_masks, view_num, res,
                      point_size, edge_point_size, mask_ratio_thresh):
    """ours_utils.py:848-882 (save_path=None)."""
    sparse = np.zeros((view_num, 3, res, res), dtype=F32)
    m0s = np.zeros_like(sparse)
    m2s = np.zeros_like(sparse)
    scales = np.zeros((view_num,), dtype=F32)
    for i in range(view_num):
        s, m0, m2, _, sc = get_one_sparse_img(point_pixels_all[i], colors, point_validation[i],
                                              hard_masks[i], res, point_size, edge_point_size,
                                              mask_ratio_thresh)
        sparse[i] = s * m0
        m0s[i] = m0
        m2s[i] = m2
        scales[i] = sc
    return sparse, m0s, m2s, scales
