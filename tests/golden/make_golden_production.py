"""Production-size golden DIGESTS: the reference's OWN project / splat / nearest-fill / unproject /
NBF / dilate code (through oracle/ref_loader.py: stub modules + the third-party shims) on the
reference's five demo clouds at the sizes of configs/default.yaml:

    30 000 points, view_num 8, res 256, cam_res 512, atlas R = 1024, edge_dilate_kernels [21],
    point_validation_by_o3d True (HPR), texture_gen_method 'nearest' (no diffusion on the CPU).

The clouds are the reference's dataset/demo_data/{clock,cup,PaulFrankLunchBox,rolling_lion}.ply and
dataset/NBF_demo_data/2ce6_chair.ply (kept as test data under tests/golden/clouds/); the mesh +
atlas stand-in is the deterministic voxel-shell proxy of tests/proxy_mesh.py (the reference ships
no meshes, POCO weights and xatlas are unavailable).

A full-tensor fixture would be ~8 MB per cloud, so what is committed is the SHA-256 of the raw
bytes (C order, the reference's dtype) of every boundary tensor, plus shapes and a few counts:
tests/golden/production_digests.json.  The nearest fills are where scipy's cKDTree tie rule is
unpinned: for those the digest is taken of the canonical-rule fill (oracle/fill.py), after
checking here that the reference's own scipy result equals it on every pixel that has a unique
nearest source; the reference's `unproject` is then fed the canonical fill, so everything
downstream is tie-independent.

Run in the build container only:   python tests/golden/make_golden_production.py [cloud ...]
"""
import hashlib
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import fill as ofill  # noqa: E402
from oracle import ref_loader  # noqa: E402
from proxy_mesh import clock_scene  # noqa: E402

CLOUDS = ["clock", "cup", "PaulFrankLunchBox", "rolling_lion", "2ce6_chair"]
PROD = dict(view_num=8, res=256, cam_res=512, atlas_res=1024, voxel_grid=40, point_size=1,
            edge_point_size=1, crop_img=True, crop_padding=0.05, mask_ratio_thresh=0.82,
            edge_dilate_kernels=[21], complete_unseen_by_projection=True, use_o3d=True)
OUT = os.path.join(HERE, "production_digests.json")


def cloud_path(name):
    p = os.path.join(HERE, "clouds", name + ".ply")
    return p if os.path.exists(p) else os.path.join(HERE, name + ".ply")


def digest(a):
    a = np.ascontiguousarray(np.asarray(a))
    return dict(sha256=hashlib.sha256(a.tobytes()).hexdigest(), shape=list(a.shape),
                dtype=str(a.dtype))


def production_scene(name):
    return clock_scene(cloud_path(name), G=PROD["voxel_grid"], atlas_res=PROD["atlas_res"])


def run_cloud(name):
    cfg = PROD
    ou = ref_loader.load("pointdreamer.ours_utils")
    un = ref_loader.load("pointdreamer.unproject")
    cu = ref_loader.load("utils.camera_utils")
    from torchvision.transforms import transforms
    sc = production_scene(name)
    V, res, cam_res = cfg["view_num"], cfg["res"], cfg["cam_res"]
    cams, base_dirs, eye_positions, up_dirs = cu.create_cameras(
        num_views=V, distance=1.6, res=cam_res, distribution="fibonacci_sphere",
        device=torch.device("cpu"))
    coords, colors = torch.from_numpy(sc["xyz"]), torch.from_numpy(sc["rgb"])
    vertices, faces = torch.from_numpy(sc["vertices"]), torch.from_numpy(sc["faces"])
    f_normals = torch.from_numpy(sc["f_normals"])
    xa = {k: torch.from_numpy(v) for k, v in sc["xatlas_dict"].items()}
    d, info = {}, {}
    with torch.no_grad(), ref_loader.quiet():
        (hard_masks, face_idxs, depths, vertice_uvs, uv_centers, uv_scales, padding, point_uvs,
         point_depths) = ou.get_rendered_hard_mask_and_face_idx_batch(
            cams, vertices, faces, coords, glctx=None, rescale=True, padding=cfg["crop_padding"])
        for k, v in dict(hard_masks_cam=hard_masks, face_idxs=face_idxs, mesh_depths=depths,
                         vertice_uvs=vertice_uvs, point_uvs=point_uvs, point_depths=point_depths,
                         uv_centers=uv_centers, uv_scales=uv_scales).items():
            d[k] = digest(v.numpy())
        hm = transforms.Resize((res, res))(hard_masks.unsqueeze(1).float()).squeeze(1).bool()
        d["hard_masks"] = digest(hm.numpy())
        pv, pvpix = ou.get_point_validation_by_depth(cam_res, point_uvs, point_depths, depths,
                                                     offset=0.0001)
        d["point_validation"] = digest(pv.numpy())
        d["point_pixels_cam"] = digest(pvpix.numpy())
        pv2 = ou.get_point_validation_by_o3d(coords, eye_positions, 100)
        d["point_validation_o3d"] = digest(pv2.numpy())
        info["visible_by_depth"], info["visible_by_hpr"] = int(pv.sum()), int(pv2.sum())
        pv = torch.logical_or(pv, pv2)
        pp = (point_uvs * res).long()
        pp = torch.cat((pp[:, :, 1].unsqueeze(-1), pp[:, :, 0].unsqueeze(-1)), dim=-1).clip(0, res - 1)
        d["point_pixels"] = digest(pp.numpy())
        sparse, m0, m2, scales = ou.get_sparse_images(
            pp, colors, pv, hm, None, V, res, cfg["point_size"], cfg["edge_point_size"],
            cfg["mask_ratio_thresh"])
        for k, v in dict(sparse_imgs=sparse, hard_mask0s=m0, hard_mask2s=m2,
                         scale_factors=scales).items():
            d[k] = digest(v.numpy())
        info["scale_factors"] = [float(s) for s in scales]
        ref_fill = ou.get_inpainted_images(sparse, m0, m2, None, None, V, method="nearest")
        ref_fill = ref_fill.numpy().astype(np.float32)
        canon = np.empty_like(ref_fill)
        n_tie = 0
        for i in range(V):
            canon[i], tie = ofill.naive_inpainting_nearest(sparse[i].numpy(), m2[i].numpy())
            diff = (canon[i] != ref_fill[i]).any(0)
            assert not (diff & ~tie).any(), f"view {i}: scipy fill differs away from ties"
            n_tie += int(tie.sum())
        info["fill_tie_pixels"] = n_tie
        d["inpainted_nearest"] = digest(canon)
        with tempfile.TemporaryDirectory() as tmp:
            atlas, shr, view_ids, pcoord, points, painted = un.unproject(
                torch.from_numpy(canon), vertices, f_normals, res, cams, cam_res, base_dirs,
                xa["gb_pos"], xa["mask"], xa["per_atlas_pixel_face_id"], uv_centers, uv_scales,
                padding, scales, depths, cfg["edge_dilate_kernels"], tmp,
                cfg["complete_unseen_by_projection"])
        for k, v in dict(atlas_img=atlas, shrinked_vis=shr, point_view_ids=view_ids,
                         points_atlas_pixel_coord=pcoord, atlas_points=points,
                         atlas_painted_mask=painted).items():
            d[k] = digest(v.numpy())
        info["painted_texels"], info["chart_texels"] = int(painted.sum()), int(xa["mask"].sum())
        info["view_id_histogram"] = np.bincount(view_ids.numpy().clip(-1, V) + 1,
                                                minlength=V + 2).tolist()
        ref_dil = un.dilate_atlas(atlas, xa["mask"]).numpy().astype(np.float32)
        canon_dil, tie = ofill.dilate_atlas(atlas.numpy(), sc["xatlas_dict"]["mask"])
        diff = (canon_dil != ref_dil).any(-1)
        assert not (diff & ~tie).any(), "scipy atlas dilation differs away from ties"
        info["dilate_tie_texels"] = int(tie.sum())
        d["atlas_dilated"] = digest(canon_dil)
    info["n_points"], info["n_faces"] = int(coords.shape[0]), int(faces.shape[0])
    return dict(digests=d, info=info)


if __name__ == "__main__":
    names = sys.argv[1:] or CLOUDS
    out = json.load(open(OUT)) if os.path.exists(OUT) else dict(config=PROD, clouds={})
    out["config"] = PROD
    for n in names:
        t0 = time.time()
        out["clouds"][n] = run_cloud(n)
        print(n, f"{time.time() - t0:.0f} s", out["clouds"][n]["info"])
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
