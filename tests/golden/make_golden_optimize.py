"""Golden vectors for the "next" rows N1 (optimize_color) and N4 (xatlas_uvmap_w_face_id),
produced by executing the REFERENCE's own source through oracle/ref_loader.py.

Run in the build container only:   python tests/golden/make_golden_optimize.py
Output: tests/golden/optimize_small.npz

The reference hard-codes the render resolution (1024) inside optimize_color, so the fixture keeps
the mesh / atlas / view count small instead and stores the optimised atlas plus an 8x-strided
sample of the returned renders.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from pointdreamer_b200 import synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CFG = dict(n_points=1500, seed=5, nu=14, nv=12, atlas_res=96, charts=(2, 2), view_num=2, res=64,
           cam_res=128, iterations=6, padding=0.05)


def inputs(cfg=CFG):
    """Deterministic inputs shared by the generator and the tests."""
    sc = synthetic.make_scene(cfg["n_points"], cfg["seed"], cfg["nu"], cfg["nv"], cfg["atlas_res"],
                              charts=cfg["charts"])
    rng = np.random.default_rng(11)
    V, res, R = cfg["view_num"], cfg["res"], cfg["atlas_res"]
    yy, xx = np.meshgrid(np.arange(res), np.arange(res), indexing="ij")
    imgs = np.stack([np.stack([0.5 + 0.4 * np.sin(0.11 * xx * (c + 1) + 0.07 * yy * (v + 1) + c)
                               for c in range(3)]) for v in range(V)]).astype(np.float32)
    imgs = np.clip(imgs + rng.normal(0, 0.05, imgs.shape).astype(np.float32), 0, 1)
    atlas0 = rng.random((R, R, 3)).astype(np.float32)          # [R,R,3] as unproject returns it
    vis = rng.random((V, R, R)) < 0.9                           # shrinked visibility
    scale_factors = np.array([1.0, 0.85][:V] + [1.0] * max(0, V - 2), dtype=np.float32)
    return sc, imgs, atlas0, vis, scale_factors


def main():
    cfg = CFG
    ou = ref_loader.load("pointdreamer.ours_utils")
    cu = ref_loader.load("utils.camera_utils")
    etm = ref_loader.load("models.get3d.extract_texture_map")
    sc, imgs, atlas0, vis, scale_factors = inputs(cfg)
    dev = torch.device("cpu")
    V = cfg["view_num"]
    cams, base_dirs, eye_positions, up_dirs = cu.create_cameras(
        num_views=V, distance=1.6, res=cfg["cam_res"], distribution="fibonacci_sphere", device=dev)
    vertices = torch.from_numpy(sc["vertices"])
    faces = torch.from_numpy(sc["faces"])
    coords = torch.from_numpy(sc["xyz"])
    uvs = torch.from_numpy(sc["xatlas_dict"]["uvs"])
    mesh_tex_idx = torch.from_numpy(sc["xatlas_dict"]["mesh_tex_idx"])
    out = {}
    with ref_loader.quiet():
        # ---- N4: xatlas_uvmap_w_face_id with the parametrisation supplied ----
        ref_loader.set_xatlas_parametrization(
            np.arange(sc["vertices"].shape[0]), sc["xatlas_dict"]["mesh_tex_idx"].astype(np.uint64),
            sc["xatlas_dict"]["uvs"])
        with torch.no_grad():
            r_uvs, r_tex_idx, gb_pos, mask, face_id = etm.xatlas_uvmap_w_face_id(
                None, vertices, faces, resolution=cfg["atlas_res"])
        out.update(uvmap_gb_pos=gb_pos.numpy(), uvmap_mask=mask.numpy(),
                   uvmap_face_id=face_id.numpy())
        # ---- crop parameters of the PROJECT stage (inputs of optimize_color) ----
        with torch.no_grad():
            (_, _, _, _, uv_centers, uv_scales, padding, _, _) = \
                ou.get_rendered_hard_mask_and_face_idx_batch(cams, vertices, faces, coords,
                                                             glctx=None, rescale=True,
                                                             padding=cfg["padding"])
        out.update(uv_centers=uv_centers.numpy(), uv_scales=uv_scales.numpy())
        # ---- N1: optimize_color exactly as demo.py:211-233 calls it ----
        atlas_in = torch.from_numpy(atlas0).permute(2, 0, 1).flip(1)
        eye_t = torch.tensor(eye_positions).float()
        look_ats = torch.zeros((len(eye_positions), 3))
        with ref_loader.cuda_literals_to_cpu():
            atlas_out, images = ou.optimize_color(
                atlas_in, torch.from_numpy(imgs), vertices, faces, uvs, mesh_tex_idx, cams, eye_t,
                look_ats, up_dirs, uv_centers, uv_scales, padding, torch.from_numpy(scale_factors),
                None, shrinked_per_view_per_pixel_visibility=torch.from_numpy(vis),
                iterations=cfg["iterations"])
    out.update(atlas_out=atlas_out.detach().numpy(),
               images_s8=images.detach().numpy()[:, :, ::8, ::8].astype(np.float64),
               images_sum=np.float64(images.detach().double().sum().item()))
    # ---- second case: optimize_from == 'naive' (no visibility mask), demo.py:221-223 ----
    with ref_loader.quiet(), ref_loader.cuda_literals_to_cpu():
        atlas_in = torch.from_numpy(atlas0.copy()).permute(2, 0, 1).flip(1)
        atlas_nv, images_nv = ou.optimize_color(
            atlas_in, torch.from_numpy(imgs), vertices, faces, uvs, mesh_tex_idx, cams, eye_t,
            look_ats, up_dirs, uv_centers, uv_scales, padding, torch.from_numpy(scale_factors),
            None, shrinked_per_view_per_pixel_visibility=None, iterations=4)
    out.update(atlas_out_novis=atlas_nv.detach().numpy(),
               images_sum_novis=np.float64(images_nv.detach().double().sum().item()))
    path = os.path.join(HERE, "optimize_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
