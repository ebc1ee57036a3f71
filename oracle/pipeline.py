"""Oracle (TEST INFRASTRUCTURE): the whole geometry path with texture_gen_method='nearest'
(configs/nearest.yaml) as one function - project -> raster -> visibility (depth | HPR) -> splat ->
nearest fill -> unproject + NBF -> atlas dilation (demo.py:93-205), every boundary tensor returned.
Composition of the restatements in this package; used by the production-size oracle test and by
bench.py's `cpu_baseline` leg of BASELINE configs[0]."""
import numpy as np

from . import camera as ocam
from . import fill as ofill
from . import hpr as ohpr
from . import project as oproj
from . import unproject as ounproj


def run_path(cfg, sc):
    """cfg: view_num, res, cam_res, crop_padding, point_size, edge_point_size, mask_ratio_thresh,
    edge_dilate_kernels, complete_unseen_by_projection; sc: scene dict of numpy arrays
    (pointdreamer_b200.synthetic).  HPR is always OR-ed in (point_validation_by_o3d: True)."""
    V, res, cam_res = cfg["view_num"], cfg["res"], cfg["cam_res"]
    cams, base_dirs, eyes, _ = ocam.create_cameras(V, 1.6, cam_res)
    params = [c.params for c in cams]
    out = {}
    pr = oproj.project_vertices_points(params, sc["vertices"], sc["xyz"], True, cfg["crop_padding"])
    for k in ["point_uvs", "point_depths", "vertice_uvs", "uv_centers", "uv_scales"]:
        out[k] = pr[k]
    depth, fidx, mask = oproj.rasterize(pr["pos"], sc["faces"], cam_res)
    out.update(hard_masks_cam=mask, face_idxs=fidx, mesh_depths=depth)
    hm = oproj.resize_mask_half_any(mask, res)
    out["hard_masks"] = hm
    vis, pix = oproj.point_validation_by_depth(cam_res, pr["point_uvs"], pr["point_depths"], depth, 0.0001)
    out.update(point_validation=vis, point_pixels_cam=pix)
    vis2 = ohpr.point_validation_by_o3d(sc["xyz"], eyes, 100)
    out["point_validation_o3d"] = vis2
    pp = oproj.point_pixels(pr["point_uvs"], res)
    out["point_pixels"] = pp
    sparse, m0, m2, scales = oproj.get_sparse_images(
        pp, sc["rgb"], vis | vis2, hm, V, res, cfg["point_size"], cfg["edge_point_size"],
        cfg["mask_ratio_thresh"])
    out.update(sparse_imgs=sparse, hard_mask0s=m0, hard_mask2s=m2, scale_factors=scales)
    filled = np.stack([ofill.naive_inpainting_nearest(sparse[i], m2[i])[0] for i in range(V)])
    out["inpainted_nearest"] = filled
    xa = sc["xatlas_dict"]
    atlas, shr, view_ids, pcoord, points, painted = ounproj.unproject(
        filled, sc["f_normals"], res, params, cam_res, base_dirs, xa["gb_pos"], xa["mask"],
        xa["per_atlas_pixel_face_id"], pr["uv_centers"], pr["uv_scales"], cfg["crop_padding"],
        scales, depth, cfg["edge_dilate_kernels"], cfg["complete_unseen_by_projection"])
    out.update(atlas_img=atlas, shrinked_vis=shr, point_view_ids=view_ids,
               points_atlas_pixel_coord=pcoord, atlas_points=points, atlas_painted_mask=painted)
    out["atlas_dilated"] = ofill.dilate_atlas(atlas, xa["mask"])[0]
    return out
