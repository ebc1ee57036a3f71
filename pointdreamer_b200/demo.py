"""The drop-in boundary: `colorize_one_mesh` with the reference's signature (demo.py:38-253),
plus `prepare`-style helpers to build its inputs from the reference's YAML configs.

The path is project -> inpaint -> unproject.  What the reference runs AFTER the path inside the
same function are "next" rows: `optimize_color` (optimize_from, N1) is built
(ours_utils.optimize_color -> csrc/texopt.cu) and so is `paint_invisible_areas_by_neighbors`
(complete_unseen_by: neighbor, N2; unproject.paint_invisible_areas_by_neighbors ->
csrc/neighbors.cu); both can be overridden with callables (`neighbor_fill=`, `optimize_color=`).
`paint_invisible_areas_by_optimize` (TextureField network) is out of scope.
"""
import os

import torch
import yaml

from . import camera as _camera
from . import ours_utils as _ou
from . import unproject as _un

# keys of configs/*.yaml consumed on the path (SURVEY §5 "config")
PATH_CONFIG_KEYS = ("view_num", "res", "cam_res", "refine_res", "point_validation_by_o3d",
                    "refine_point_validation_by_remove_abnormal_depth",
                    "hidden_point_removal_radius", "texture_gen_method", "point_size",
                    "edge_point_size", "crop_img", "crop_padding", "mask_ratio_thresh",
                    "optimize_from", "edge_dilate_kernels", "complete_unseen_by",
                    "xatlas_texture_res")

DEFAULT_CONFIG = dict(  # configs/default.yaml
    camera_distribution="fibonacci_sphere", cam_res=512, view_num=8, res=256, point_size=1,
    edge_point_size=1, point_validation_by_o3d=True, hidden_point_removal_radius=100,
    refine_point_validation_by_remove_abnormal_depth=False, refine_res=512, crop_img=True,
    crop_padding=0.05, mask_ratio_thresh=0.82, edge_dilate_kernels=[21], optimize_from="ours",
    xatlas_texture_res=1024, complete_unseen_by="neighbor", texture_gen_method="DDNM_inpaint")


def load_config(path):
    """demo.py:315-316: YAML -> dict (the reference wraps it in a Munch and splats it as **cfg)."""
    with open(path, "r") as f:
        return yaml.safe_load(f)


def prepare_cameras(cfg, device):
    """demo.py:331-353: camera rig dict consumed by colorize_one_mesh."""
    cams, base_dirs, eye_positions, up_dirs = _camera.create_cameras(
        num_views=cfg["view_num"], distribution=cfg.get("camera_distribution", "fibonacci_sphere"),
        distance=1.6, res=cfg["cam_res"], device=device)
    return dict(cams=cams, base_dirs=base_dirs, eye_positions=eye_positions, up_dirs=up_dirs,
                cam_RTs=None, cam_K=None)


def _num_texels(xatlas_dict):
    """mask.sum() of an atlas, computed once per xatlas_dict and kept in it (the chart mask is
    static per mesh, demo.py:430-448), so repeated colorize calls do not synchronise the host."""
    n = xatlas_dict.get('_pdr_num_texels')
    if n is None:
        n = _un.count_texels(xatlas_dict['mask'])
        xatlas_dict['_pdr_num_texels'] = n
    return n


def _after_path(atlas_img, atlas_painted_mask, shrinked_vis, inpainted_images, vertices, faces,
                uvs, mesh_tex_idx, mask, per_atlas_pixel_face_id, cams, eye_positions, up_dirs,
                uv_centers, uv_scales, padding, inpaint_scale_factors, glctx, complete_unseen_by,
                optimize_from, neighbor_fill=None, optimize_color=None, _mark=lambda name: None):
    """demo.py:180-246, what the reference runs AFTER project -> inpaint -> unproject inside
    colorize_one_mesh: unseen-texel completion ("next" row N2) and optimize_color (N1)."""
    if complete_unseen_by == 'unproject':
        atlas_img = _un.dilate_atlas(atlas_img, mask)
    elif complete_unseen_by == 'neighbor':
        if neighbor_fill is None:
            neighbor_fill = _un.paint_invisible_areas_by_neighbors
        to_inpaint_face_id = per_atlas_pixel_face_id[0][torch.logical_not(atlas_painted_mask)].unique()
        to_inpaint_face_id = to_inpaint_face_id[to_inpaint_face_id > -1]
        atlas_img = neighbor_fill(vertices, faces, uvs, mesh_tex_idx, to_inpaint_face_id,
                                  atlas_img, atlas_painted_mask, use_atlas=True)
    elif complete_unseen_by == 'optimize':
        raise NotImplementedError("complete_unseen_by='optimize' (TextureField) is out of scope")
    _mark("complete_unseen")
    # ---- demo.py:211-233: refine the atlas against the inpainted views ("next" row N1)
    if optimize_from is not None and optimize_from != 'None':
        if optimize_color is not None:  # caller-supplied replacement
            atlas_img = optimize_color(atlas_img, inpainted_images, shrinked_vis)
        else:
            atlas_in = atlas_img.permute(2, 0, 1).flip(1)  # [3,R,R]
            if optimize_from == 'scratch':
                init_atlas, vis = None, None
            elif optimize_from == 'naive':
                init_atlas, vis = atlas_in, None
            elif optimize_from == 'ours':
                init_atlas, vis = atlas_in, shrinked_vis
            else:
                raise ValueError(f"optimize_from={optimize_from!r}")
            atlas_opt, _ = _ou.optimize_color(
                init_atlas, inpainted_images, vertices, faces, uvs, mesh_tex_idx, cams,
                eye_positions, None, up_dirs, uv_centers, uv_scales, padding,
                inpaint_scale_factors, glctx, shrinked_per_view_per_pixel_visibility=vis,
                return_images=False)
            atlas_img = atlas_opt[0].flip(1).permute(1, 2, 0)  # [R,R,3]
    _mark("optimize_color")
    return atlas_img


def colorize_one_mesh(coords, colors, vertices, faces, f_normals, xatlas_dict, camera_info,
                      view_num, res, cam_res, refine_res, device, save_img_path,
                      point_validation_by_o3d, refine_point_validation_by_remove_abnormal_depth,
                      hidden_point_removal_radius, texture_gen_method, point_size, edge_point_size,
                      crop_img, crop_padding, mask_ratio_thresh, optimize_from, edge_dilate_kernels,
                      complete_unseen_by, inpainter, glctx, logger, xatlas_texture_res,
                      neighbor_fill=None, optimize_color=None, stage_events=None, **kwargs):
    """demo.py:38-253.  Returns (vertices, uvs, faces, mesh_tex_idx, atlas_img[R,R,3], mask).
    stage_events: optional list; (name, torch.cuda.Event) pairs are appended at the stage
    boundaries (bench.py's per-stage breakdown; no synchronisation is added)."""

    def _mark(name):
        if stage_events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            stage_events.append((name, ev))

    _mark("start")
    base_dirs = camera_info['base_dirs']
    cams = camera_info['cams']
    eye_positions = camera_info['eye_positions']
    uvs = xatlas_dict['uvs']
    mesh_tex_idx = xatlas_dict['mesh_tex_idx']
    gb_pos = xatlas_dict['gb_pos']
    mask = xatlas_dict['mask']
    per_atlas_pixel_face_id = xatlas_dict['per_atlas_pixel_face_id']

    with torch.no_grad():
        # ---- PROJECT (demo.py:93-129)
        (hard_masks, face_idxs, mesh_normalized_depths, vertice_uvs, uv_centers, uv_scales,
         padding, point_uvs, point_depths) = _ou.get_rendered_hard_mask_and_face_idx_batch(
            cams, vertices, faces, coords, glctx=glctx, rescale=crop_img, padding=crop_padding)
        if cam_res != res:
            hard_masks = _ou.resize_hard_masks(hard_masks, res)
        point_validation, _ = _ou.get_point_validation_by_depth(
            cam_res, point_uvs, point_depths, mesh_normalized_depths, offset=0.0001)
        if point_validation_by_o3d:
            point_validation2 = _ou.get_point_validation_by_o3d(coords, eye_positions,
                                                                hidden_point_removal_radius)
            point_validation = torch.logical_or(point_validation, point_validation2)
        if refine_point_validation_by_remove_abnormal_depth:
            raise NotImplementedError("refine_point_validation is off in every shipped config "
                                      "(configs/default.yaml:46) and outside the hot path")
        point_pixels = _ou.get_point_pixels(point_uvs, res)
        sparse_imgs, hard_mask0s, hard_mask2s, inpaint_scale_factors = _ou.get_sparse_images(
            point_pixels, colors, point_validation, hard_masks, save_img_path, view_num, res,
            point_size, edge_point_size, mask_ratio_thresh)

        _mark("project")
        # ---- INPAINT (demo.py:137-157), including the PNG cache of a previous run
        inpainted_images = None
        if save_img_path is not None:
            from .io_utils import load_inpainted_pngs
            cached = load_inpainted_pngs(save_img_path, view_num, res)
            if cached is not None:
                inpainted_images = torch.from_numpy(cached).to(device)
        if inpainted_images is None:
            inpainted_images = _ou.get_inpainted_images(
                sparse_imgs, hard_mask0s, hard_mask2s, save_img_path, inpainter, view_num,
                method=texture_gen_method)

        _mark("inpaint")
        # ---- UNPROJECT (demo.py:168-177)
        complete_unseen_by_projection = (complete_unseen_by == 'unproject')
        (atlas_img, shrinked_vis, point_view_ids, points_atlas_pixel_coord, points,
         atlas_painted_mask) = _un.unproject(
            inpainted_images, vertices, f_normals, res, cams, cam_res, base_dirs, gb_pos, mask,
            per_atlas_pixel_face_id, uv_centers, uv_scales, padding, inpaint_scale_factors,
            mesh_normalized_depths, edge_dilate_kernels, save_img_path,
            complete_unseen_by_projection, num_texels=_num_texels(xatlas_dict))

        _mark("unproject")
        atlas_img = _after_path(
            atlas_img, atlas_painted_mask, shrinked_vis, inpainted_images, vertices, faces, uvs,
            mesh_tex_idx, mask, per_atlas_pixel_face_id, cams, eye_positions,
            camera_info.get('up_dirs'), uv_centers, uv_scales, padding, inpaint_scale_factors,
            glctx, complete_unseen_by, optimize_from, neighbor_fill, optimize_color, _mark)
    return vertices, uvs, faces, mesh_tex_idx, atlas_img, mask


def colorize_batch(scenes, camera_info, cfg, inpainter, device):
    """BASELINE.json configs[3]: several shapes at once on one GPU.  PROJECT and UNPROJECT run per
    shape; all S*V diffusion chains run as ONE U-Net batch (chain s*V+v keeps the noise-stream slot
    the reference's serial loop over shapes and views would give it).  `scenes`: list of dicts with
    device tensors xyz, rgb, vertices, faces, f_normals, xatlas_dict.  Returns a list of atlases.
    The post-path steps cfg asks for (complete_unseen_by, optimize_from) run per shape exactly as
    in colorize_one_mesh, so the result equals shape-by-shape calls."""
    keys = {k: cfg[k] for k in PATH_CONFIG_KEYS}
    V, res, cam_res = keys["view_num"], keys["res"], keys["cam_res"]
    if keys["texture_gen_method"] != "DDNM_inpaint":
        return [colorize_one_mesh(sc["xyz"], sc["rgb"], sc["vertices"], sc["faces"],
                                  sc["f_normals"], sc["xatlas_dict"], camera_info, device=device,
                                  save_img_path=None, inpainter=inpainter, glctx=None, logger=None,
                                  **keys)[4] for sc in scenes]
    cams = camera_info['cams']
    staged = []
    with torch.no_grad():
        for sc in scenes:
            (hard_masks, _, depths, _, uv_centers, uv_scales, padding, point_uvs,
             point_depths) = _ou.get_rendered_hard_mask_and_face_idx_batch(
                cams, sc["vertices"], sc["faces"], sc["xyz"], glctx=None, rescale=keys["crop_img"],
                padding=keys["crop_padding"])
            if cam_res != res:
                hard_masks = _ou.resize_hard_masks(hard_masks, res)
            pv, _ = _ou.get_point_validation_by_depth(cam_res, point_uvs, point_depths, depths,
                                                      offset=0.0001)
            if keys["point_validation_by_o3d"]:
                pv = torch.logical_or(pv, _ou.get_point_validation_by_o3d(
                    sc["xyz"], camera_info['eye_positions'], keys["hidden_point_removal_radius"]))
            pp = _ou.get_point_pixels(point_uvs, res)
            sparse, m0, m2, scales = _ou.get_sparse_images(
                pp, sc["rgb"], pv, hard_masks, None, V, res, keys["point_size"],
                keys["edge_point_size"], keys["mask_ratio_thresh"])
            staged.append(dict(sparse=sparse, m2=m2, scales=scales, depths=depths,
                               uv_centers=uv_centers, uv_scales=uv_scales, padding=padding))
        inpainted = inpainter.inpaint_batch(torch.cat([s_["sparse"] for s_ in staged], 0),
                                            torch.cat([s_["m2"][:, 0] for s_ in staged], 0))
        atlases = []
        for i, (sc, st) in enumerate(zip(scenes, staged)):
            xa = sc["xatlas_dict"]
            views = inpainted[i * V:(i + 1) * V]
            atlas, shr, _, _, _, painted = _un.unproject(
                views, sc["vertices"], sc["f_normals"], res, cams, cam_res,
                camera_info['base_dirs'], xa["gb_pos"], xa["mask"], xa["per_atlas_pixel_face_id"],
                st["uv_centers"], st["uv_scales"], st["padding"], st["scales"], st["depths"],
                keys["edge_dilate_kernels"], None, keys["complete_unseen_by"] == 'unproject',
                num_texels=_num_texels(xa))
            atlases.append(_after_path(
                atlas, painted, shr, views, sc["vertices"], sc["faces"], xa.get("uvs"),
                xa.get("mesh_tex_idx"), xa["mask"], xa["per_atlas_pixel_face_id"], cams,
                camera_info['eye_positions'], camera_info.get('up_dirs'), st["uv_centers"],
                st["uv_scales"], st["padding"], st["scales"], None, keys["complete_unseen_by"],
                keys["optimize_from"]))
    return atlases


def colorize_from_host(scene, camera_info, cfg, inpainter, device):
    """End-to-end convenience used by bench.py's `e2e` leg: every input starts in (pinned) HOST
    memory, is copied to the device, run through colorize_one_mesh, and the atlas is copied back.
    Returns (atlas_host [R,R,3] float32 pinned tensor, h2d_bytes, d2h_bytes)."""
    xa = scene["xatlas_dict"]
    host = [scene["xyz"], scene["rgb"], scene["vertices"], scene["faces"], scene["f_normals"],
            xa["uvs"], xa["mesh_tex_idx"], xa["gb_pos"], xa["mask"], xa["per_atlas_pixel_face_id"]]
    h2d = sum(t.numel() * t.element_size() for t in host)
    d = [t.to(device, non_blocking=True) for t in host]
    xad = dict(uvs=d[5], mesh_tex_idx=d[6], gb_pos=d[7], mask=d[8], per_atlas_pixel_face_id=d[9])
    xad['_pdr_num_texels'] = int(xa["mask"].sum())  # counted on the host copy: no device sync
    keys = {k: cfg[k] for k in PATH_CONFIG_KEYS}
    out = colorize_one_mesh(d[0], d[1], d[2], d[3], d[4], xad, camera_info, device=device,
                            save_img_path=None, inpainter=inpainter, glctx=None, logger=None, **keys)
    atlas = out[4]
    atlas_host = torch.empty(atlas.shape, dtype=atlas.dtype, pin_memory=True)
    atlas_host.copy_(atlas, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return atlas_host, h2d, atlas.numel() * atlas.element_size()


# ------------------------------------------------------------------------------------------
# "next" row N3: the file-level flow of demo.py (recon_one_textured_mesh, 358-470) and its CLI
# ------------------------------------------------------------------------------------------
def recon_one_textured_mesh(cfg, inpainter, camera_info, pc_file, name, device, mesh_file=None,
                            logger=None):
    """demo.py:358-470 without the POCO/SPR geometry stage (out of scope): the
    untextured mesh must exist as `<pc_file minus .ply>_untextured_mesh.obj` (the reference's own
    load_exist_geo branch, demo.py:391-398) or be given as `mesh_file`.

    Writes the reference's output tree under cfg['output_path']/name:
      input_pc.ply, geo/xatlas_<R>.pth, others/{i}_{sparse,mask0,mask2,inpainted}.png,
      models/model_normalized.{obj,mtl,png}, others/atlas_wo_background.png.
    UV unwrapping: geo/xatlas_<R>.pth is reused when present (demo.py:430-438); otherwise the
    mesh's own `vt` / `f v/vt` records are taken as the parametrisation, or xatlas is called if it
    is installed."""
    from . import extract_texture_map as _etm
    from . import io_utils as _io
    out_root = os.path.join(cfg["output_path"], name)
    for sub in ("geo", "models", "others"):
        os.makedirs(os.path.join(out_root, sub), exist_ok=True)
    R = cfg["xatlas_texture_res"]
    xatlas_save_file = os.path.join(out_root, "geo", f"xatlas_{R}.pth")

    xyz_np, rgb_np = _io.read_ply_xyzrgb(pc_file)
    if len(xyz_np) > 30000:  # demo.py:371-374
        raise NotImplementedError(
            f"Point number > 30000! ({len(xyz_np)} points)({pc_file}) \n Please try uniformly "
            "subsampling the input point cloud first")
    xyz = torch.from_numpy(xyz_np).to(device)
    rgb = torch.from_numpy(rgb_np).float().to(device) / 255.0
    vmin, vmax = xyz.min(0)[0], xyz.max(0)[0]
    xyz = (xyz - (vmax + vmin) / 2.) / (vmax - vmin).max()
    _io.save_colored_pc_ply(xyz.cpu().numpy(), rgb.cpu().numpy(), os.path.join(out_root, "input_pc.ply"))

    geo = mesh_file or pc_file.replace(".ply", "_untextured_mesh.obj")
    if not os.path.exists(geo):
        raise NotImplementedError(
            f"no untextured mesh at {geo}: geometry reconstruction (POCO / SPR, demo.py:399-419) is "
            "outside this package; pass mesh_file=")
    v_np, uv_np, f_np, ft_np = _io.loadobjtex(geo)
    vertices = torch.from_numpy(v_np).to(device)
    faces = torch.from_numpy(f_np).to(device)
    vertices = (vertices - (vmax + vmin) / 2.) / (vmax - vmin).max()  # demo.py:396-397
    f_normals = _ou.face_normals(vertices, faces)

    xatlas_dict = None
    if os.path.exists(xatlas_save_file):
        try:
            xatlas_dict = {k: v.to(device) for k, v in torch.load(xatlas_save_file).items()}
        except Exception:
            xatlas_dict = None
    if xatlas_dict is None:
        par = (None, ft_np.astype("uint64"), uv_np) if uv_np is not None else None
        uvs, mesh_tex_idx, gb_pos, mask, face_id = _etm.xatlas_uvmap_w_face_id(
            None, vertices, faces, resolution=R, parametrization=par)
        xatlas_dict = {'uvs': uvs, 'mesh_tex_idx': mesh_tex_idx, 'gb_pos': gb_pos, 'mask': mask,
                       'per_atlas_pixel_face_id': face_id}
        torch.save({k: v for k, v in xatlas_dict.items() if torch.is_tensor(v)}, xatlas_save_file)

    keys = {k: cfg[k] for k in PATH_CONFIG_KEYS}
    vertices, uvs, faces, mesh_tex_idx, atlas_img, mask = colorize_one_mesh(
        xyz, rgb, vertices, faces, f_normals, xatlas_dict, camera_info, inpainter=inpainter,
        save_img_path=os.path.join(out_root, "others"), device=device, logger=logger, glctx=None,
        **keys)
    _io.save_textured_mesh(vertices, uvs, faces, mesh_tex_idx, atlas_img, mask, out_root)
    return out_root


def main(argv=None):
    """python -m pointdreamer_b200.demo --config configs/default.yaml --pc_file cloud.ply
    [--mesh_file mesh.obj] [--output_path output] [--ckpt 256x256_diffusion_uncond.pt]"""
    import argparse

    from .ddnm_inpainting import Inpainter
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("--config", default=None, help="one of the reference's configs/*.yaml")
    ap.add_argument("--pc_file", required=True)
    ap.add_argument("--mesh_file", default=None)
    ap.add_argument("--output_path", default=None)
    ap.add_argument("--ckpt", default="models/DDNM/256x256_diffusion_uncond.pt")
    ap.add_argument("--view_num", type=int, default=None)
    ap.add_argument("--texture_gen_method", default=None)
    args = ap.parse_args(argv)
    cfg = dict(DEFAULT_CONFIG)
    if args.config:
        cfg.update(load_config(args.config))
    for k in ("output_path", "view_num", "texture_gen_method"):
        if getattr(args, k) is not None:
            cfg[k] = getattr(args, k)
    cfg.setdefault("output_path", "output")
    device = torch.device("cuda:0")
    inpainter = Inpainter(device, ckpt_path=args.ckpt) \
        if cfg["texture_gen_method"] == "DDNM_inpaint" else None
    camera_info = prepare_cameras(cfg, device)
    name = os.path.basename(args.pc_file).split(".")[0]
    out = recon_one_textured_mesh(cfg, inpainter, camera_info, args.pc_file, name, device,
                                  mesh_file=args.mesh_file)
    print("wrote", out)


if __name__ == "__main__":
    main()
