"""Generate golden vectors by executing the REFERENCE's own source (YuQiao0303/PointDreamer,
/root/reference) through oracle/ref_loader.py (stub modules + the four third-party shims).

Run in the build container only:   python tests/golden/make_golden.py
Outputs: tests/golden/geom_case_*.npz   (inputs + every boundary tensor of the path)
         tests/golden/unet_case_*.npz   (see make_golden_unet.py)
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from pointdreamer_b200 import synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from make_golden_cases import CASES  # noqa: E402


def run_case(name, cfg):
    ou = ref_loader.load("pointdreamer.ours_utils")
    un = ref_loader.load("pointdreamer.unproject")
    cu = ref_loader.load("utils.camera_utils")
    from torchvision.transforms import transforms

    if cfg.get("scene") == "clock":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from proxy_mesh import clock_scene
        sc = clock_scene(os.path.join(HERE, "clock.ply"), atlas_res=cfg["atlas_res"])
    else:
        sc = synthetic.make_scene(cfg["n_points"], cfg["seed"], cfg["nu"], cfg["nv"],
                                  cfg["atlas_res"], charts=cfg["charts"])
    dev = torch.device("cpu")
    V, res, cam_res = cfg["view_num"], cfg["res"], cfg["cam_res"]
    cams, base_dirs, eye_positions, up_dirs = cu.create_cameras(
        num_views=V, distance=1.6, res=cam_res, distribution="fibonacci_sphere", device=dev)

    coords = torch.from_numpy(sc["xyz"])
    colors = torch.from_numpy(sc["rgb"])
    vertices = torch.from_numpy(sc["vertices"])
    faces = torch.from_numpy(sc["faces"])
    f_normals = torch.from_numpy(sc["f_normals"])
    xa = {k: torch.from_numpy(v) for k, v in sc["xatlas_dict"].items()}

    out = {}
    with torch.no_grad(), ref_loader.quiet():
        (hard_masks, face_idxs, depths, vertice_uvs, uv_centers, uv_scales, padding, point_uvs,
         point_depths) = ou.get_rendered_hard_mask_and_face_idx_batch(
            cams, vertices, faces, coords, glctx=None, rescale=cfg["crop_img"],
            padding=cfg["crop_padding"])
        out.update(hard_masks_cam=hard_masks.numpy(), face_idxs=face_idxs.numpy(),
                   mesh_depths=depths.numpy(), vertice_uvs=vertice_uvs.numpy(),
                   point_uvs=point_uvs.numpy(), point_depths=point_depths.numpy())
        if torch.is_tensor(uv_centers):
            out.update(uv_centers=uv_centers.numpy(), uv_scales=uv_scales.numpy())
        out["padding"] = np.float64(padding)
        # demo.py:103-104
        hm = transforms.Resize((res, res))(hard_masks.unsqueeze(1).float()).squeeze(1).bool()
        out["hard_masks"] = hm.numpy()
        # demo.py:107
        pv, pvpix = ou.get_point_validation_by_depth(cam_res, point_uvs, point_depths, depths,
                                                     offset=0.0001)
        out.update(point_validation=pv.numpy(), point_pixels_cam=pvpix.numpy())
        # HPR through the scipy/Qhull shim (ours_utils.py:204-225)
        pv2 = ou.get_point_validation_by_o3d(coords, eye_positions, 100)
        out["point_validation_o3d"] = pv2.numpy()
        if cfg.get("use_o3d"):
            pv = torch.logical_or(pv, pv2)  # demo.py:108-110
        # demo.py:121-125
        pp = (point_uvs * res).long()
        pp = torch.cat((pp[:, :, 1].unsqueeze(-1), pp[:, :, 0].unsqueeze(-1)), dim=-1)
        pp = pp.clip(0, res - 1)
        out["point_pixels"] = pp.numpy()
        sparse, m0, m2, scales = ou.get_sparse_images(
            pp, colors, pv, hm, None, V, res, cfg["point_size"], cfg["edge_point_size"],
            cfg["mask_ratio_thresh"])
        out.update(sparse_imgs=sparse.numpy(), hard_mask0s=m0.numpy(), hard_mask2s=m2.numpy(),
                   scale_factors=scales.numpy())
        # texture_gen_method == 'nearest' (ours_utils.py:930-941; scipy griddata)
        inpainted = ou.get_inpainted_images(sparse, m0, m2, None, None, V, method="nearest")
        out["inpainted_nearest"] = inpainted.numpy().astype(np.float32)
        inpainted = inpainted.float()
        with tempfile.TemporaryDirectory() as tmp:
            atlas, shr, view_ids, pcoord, points, painted = un.unproject(
                inpainted, vertices, f_normals, res, cams, cam_res, base_dirs, xa["gb_pos"],
                xa["mask"], xa["per_atlas_pixel_face_id"],
                uv_centers if torch.is_tensor(uv_centers) else uv_centers,
                uv_scales if torch.is_tensor(uv_scales) else uv_scales, padding, scales, depths,
                cfg["edge_dilate_kernels"], tmp, cfg["complete_unseen_by_projection"])
        out.update(atlas_img=atlas.numpy(), shrinked_vis=shr.numpy(),
                   point_view_ids=view_ids.numpy(), points_atlas_pixel_coord=pcoord.numpy(),
                   atlas_points=points.numpy(), atlas_painted_mask=painted.numpy())
        dil = un.dilate_atlas(atlas, xa["mask"])
        out["atlas_dilated"] = dil.numpy().astype(np.float32)

    out["base_dirs"] = base_dirs.numpy()
    out["eye_positions"] = np.asarray(eye_positions)
    out["up_dirs"] = up_dirs.numpy()
    out["cam_params"] = np.stack([c.params.numpy() for c in cams])
    path = os.path.join(HERE, f"geom_case_{name}.npz")
    packed = {}
    for k, v in out.items():
        v = np.asarray(v)
        if v.dtype == np.bool_:
            packed["bool:" + k] = np.packbits(v.reshape(-1))
            packed["shape:" + k] = np.asarray(v.shape)
        elif v.dtype == np.int64 and v.size and np.abs(v).max() < 2 ** 15:
            packed["i16:" + k] = v.astype(np.int16)
        else:
            packed[k] = v
    np.savez_compressed(path, **packed)
    print(name, "->", path, f"{os.path.getsize(path) / 1024:.0f} KiB",
          "valid/view:", pv.sum(1).tolist(), "scale:", scales.tolist(),
          "painted:", int(painted.sum()), "of", int(xa['mask'].sum()))


if __name__ == "__main__":
    only = sys.argv[1:] or list(CASES)
    for n in only:
        run_case(n, CASES[n])
