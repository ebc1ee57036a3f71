"""Oracle (TEST INFRASTRUCTURE): texture optimiser and atlas inputs ("next" rows N1 / N4).

Restates
  * pointdreamer/ours_utils.py:1583-1785  optimize_color
  * models/get3d/extract_texture_map.py:42-64  xatlas_uvmap_w_face_id (everything after
    xatlas.parametrize, which stays third-party: the parametrisation is an input here)
  * kaolin.ops.mesh.face_normals (demo.py:422)
on CPU with numpy + torch autograd.  Third-party arithmetic that is not vendored in the
reference (PARITY UNPINNED for these, canonical rules documented where they are defined):
  * kaolin 0.15.0 `render.mesh.texture_mapping(coords, maps, mode='bilinear')` - published
    algorithm: coords*2-1, y negated, F.grid_sample(align_corners=False, padding_mode='border');
  * nvdiffrast rasterize / interpolate - oracle/project.py:rasterize / interpolate;
  * kaolin Camera.transform - oracle/camera.py.
Pinned against the reference's own optimize_color / xatlas_uvmap_w_face_id executed under
oracle/ref_loader.py (tests/golden/make_golden_optimize.py -> optimize_small.npz).
"""
import numpy as np

from . import camera as ocam
from . import project as oproj

F32 = np.float32


def face_normals(vertices, faces):
    """kal.ops.mesh.face_normals(face_vertices, unit=True): normalised (v1-v0) x (v2-v0), fp32."""
    v = np.asarray(vertices, dtype=F32)
    f = np.asarray(faces, dtype=np.int64)
    e1 = v[f[:, 1]] - v[f[:, 0]]
    e2 = v[f[:, 2]] - v[f[:, 0]]
    nx = e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1]
    ny = e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2]
    nz = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    ln = np.sqrt((nx * nx + ny * ny) + nz * nz).astype(F32)
    d = np.maximum(ln, F32(1e-12))
    return np.stack([nx / d, ny / d, nz / d], 1).astype(F32)


def uvmap_w_face_id(mesh_v, mesh_pos_idx, uvs, mesh_tex_idx, resolution):
    """extract_texture_map.py:50-64: rasterise the mesh in UV space and interpolate the world
    position per texel.  Returns (gb_pos [1,R,R,3] f32, mask [1,R,R,1] bool,
    per_pixel_face_idx [1,R,R] int64)."""
    uvs = np.asarray(uvs, dtype=F32)
    uv_clip = uvs[None, ...] * F32(2.0) - F32(1.0)
    pos = np.concatenate([uv_clip, np.zeros_like(uv_clip[..., :1]), np.ones_like(uv_clip[..., :1])],
                         -1).astype(F32)
    _, fidx, mask, bary = oproj.rasterize(pos, mesh_tex_idx, resolution, return_bary=True)
    gb_pos = oproj.interpolate(bary, fidx, mesh_v, mesh_pos_idx)
    return gb_pos, mask[..., None], fidx


def view_uv_maps(cam_params, vertices, faces, uvs, mesh_tex_idx, uv_centers, uv_scales, padding,
                 inpaint_scale_factors, res):
    """ours_utils.py:1674-1716: per-view clip-space vertices (crop of the PROJECT stage times the
    inpainting scale factor), rasterise at `res`, interpolate the texture uv, flip vertically.
    Returns (uv_map [V,res,res,2] f32, mask [V,res,res] bool) in the flipped (image) frame."""
    V = len(cam_params)
    verts = np.asarray(vertices, dtype=F32)
    pos = np.zeros((V, verts.shape[0], 4), dtype=F32)
    for i in range(V):
        pos[i, :, :3] = ocam.transform(cam_params[i], verts)
        pos[i, :, 3] = 1.0
    c = np.asarray(uv_centers, dtype=F32).reshape(V, 1, 2)
    s = np.asarray(uv_scales, dtype=F32).reshape(V, 1, 1)
    isf = np.asarray(inpaint_scale_factors, dtype=F32).reshape(V, 1, 1)
    vuv = (pos[:, :, :2] - c) / s
    vuv = vuv * F32(1 - 2 * padding)
    vuv = vuv * isf
    vuv = vuv + F32(0.5)
    vuv = np.clip(vuv, F32(0), F32(1))
    pos[:, :, :2] = vuv * F32(2) - F32(1)
    _, fidx, mask, bary = oproj.rasterize(pos, faces, res, return_bary=True)
    uv_map = oproj.interpolate(bary, fidx, uvs, mesh_tex_idx)
    return uv_map[:, ::-1].copy(), mask[:, ::-1].copy()


def texture_mapping_bilinear(texture_coords, texture_maps):
    """kaolin 0.15.0 render.mesh.texture_mapping(mode='bilinear') (call site
    ours_utils.py:1721): [B,H,W,2] coords in [0,1], [B,C,R,R] maps -> [B,H,W,C]."""
    import torch
    B = texture_coords.shape[0]
    g = texture_coords.reshape(B, 1, -1, 2)
    g = g * 2.0 - 1.0
    g = torch.stack([g[..., 0], -g[..., 1]], -1)
    t = torch.nn.functional.grid_sample(texture_maps, g, mode="bilinear", align_corners=False,
                                        padding_mode="border")
    t = t.permute(0, 2, 3, 1)
    return t.reshape(*texture_coords.shape[:-1], texture_maps.shape[1])


def optimize_color(atlas_img, inpainted_imgs, uv_map, mask, shrinked_vis=None, lr=5e-2,
                   iterations=100, res=1024):
    """ours_utils.py:1607-1632, 1711-1782 given the rasterised uv_map / mask of view_uv_maps.
    atlas_img [3,R,R] f32 (already permuted + flipped by the caller, demo.py:217),
    inpainted_imgs [V,3,r0,r0] f32, shrinked_vis [V,R,R] bool or None.
    Returns (atlas [1,3,R,R] f32, images [V,3,res,res] f64) as numpy."""
    import torch
    atlas = torch.from_numpy(np.array(atlas_img, dtype=np.float32)).unsqueeze(0).requires_grad_()
    target = torch.from_numpy(np.ascontiguousarray(inpainted_imgs)).float()
    coords = torch.from_numpy(np.ascontiguousarray(uv_map))
    m = torch.from_numpy(np.ascontiguousarray(mask))[..., None]  # [V,res,res,1] bool
    V = coords.shape[0]
    R = atlas.shape[3]
    opt = torch.optim.Adam([atlas], lr=lr)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=15, gamma=0.5)
    tl = torch.clip((coords * R).long(), 0, R - 1)
    vis_mask = None
    if shrinked_vis is not None:
        sv = torch.from_numpy(np.ascontiguousarray(shrinked_vis))
        vis_mask = sv[torch.arange(V)[:, None, None], tl[..., 1], tl[..., 0]].unsqueeze(1)
    fg = m.permute(0, 3, 1, 2).repeat(1, 3, 1, 1).float()
    target = torch.nn.functional.interpolate(target, size=(res, res), mode="bilinear",
                                             align_corners=False, antialias=False)
    target = target * fg
    if vis_mask is not None:
        target = target * vis_mask.float()
    images = None
    for _ in range(iterations):
        opt.zero_grad()
        images = texture_mapping_bilinear(coords.double(), atlas.repeat(V, 1, 1, 1).double())
        images = torch.clamp(images * m, 0., 1.)
        images = torch.clamp(images, 0., 1.)
        images = images.permute(0, 3, 1, 2)
        images = images * fg
        if vis_mask is not None:
            images = images * vis_mask.float()
        loss = torch.mean(torch.abs(images - target))
        loss.backward()
        opt.step()
        sched.step()
    return atlas.detach().numpy(), images.detach().numpy()
