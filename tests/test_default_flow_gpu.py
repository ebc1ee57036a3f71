"""configs/default.yaml's full post-path flow through colorize_one_mesh on the GPU:
project -> inpaint -> unproject -> paint_invisible_areas_by_neighbors (complete_unseen_by:
neighbor) -> optimize_color (optimize_from: ours), against the oracle chain.
texture_gen_method='nearest' keeps the inpainting exact so the comparison is tight; the DDNM
variant is covered by test_ddnm_gpu / smoke()."""
import numpy as np
import pytest
import torch

from oracle import camera as ocam, fill as ofill, hpr as ohpr, neighbors as onb
from oracle import optimize as oopt, project as oproj, unproject as ounproj

pytestmark = pytest.mark.gpu


def test_default_yaml_flow_matches_oracle(cuda):
    from pointdreamer_b200 import demo, synthetic
    V, res, cam_res, R = 2, 64, 128, 256  # R >= 256: unproject.py:289 repeats the kernel list R//256 times
    cfg = dict(demo.DEFAULT_CONFIG, view_num=V, res=res, cam_res=cam_res, xatlas_texture_res=R,
               texture_gen_method="nearest", edge_dilate_kernels=[5])
    assert cfg["complete_unseen_by"] == "neighbor" and cfg["optimize_from"] == "ours"
    sc = synthetic.make_scene(4000, seed=2, nu=20, nv=16, atlas_res=R, charts=(2, 2))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    xa = {k: t(v) for k, v in sc["xatlas_dict"].items()}
    cam_info = demo.prepare_cameras(cfg, cuda)
    keys = {k: cfg[k] for k in demo.PATH_CONFIG_KEYS}
    out = demo.colorize_one_mesh(t(sc["xyz"]), t(sc["rgb"]), t(sc["vertices"]), t(sc["faces"]),
                                 t(sc["f_normals"]), xa, cam_info, device=cuda, save_img_path=None,
                                 inpainter=None, glctx=None, logger=None, **keys)
    atlas = out[4].cpu().numpy()
    assert atlas.shape == (R, R, 3) and np.isfinite(atlas).all()

    # ---- oracle chain ----
    cams, base_dirs, eyes, _ = ocam.create_cameras(V, 1.6, cam_res)
    params = [c.params for c in cams]
    pr = oproj.project_vertices_points(params, sc["vertices"], sc["xyz"], True, 0.05)
    depth, fidx, mask = oproj.rasterize(pr["pos"], sc["faces"], cam_res)
    hm = oproj.resize_mask_half_any(mask, res)
    vis, _ = oproj.point_validation_by_depth(cam_res, pr["point_uvs"], pr["point_depths"], depth,
                                             0.0001)
    vis = vis | ohpr.point_validation_by_o3d(sc["xyz"], eyes, cfg["hidden_point_removal_radius"])
    pp = oproj.point_pixels(pr["point_uvs"], res)
    sparse, m0, m2, scales = oproj.get_sparse_images(pp, sc["rgb"], vis, hm, V, res, 1, 1, 0.82)
    inpainted = np.stack([ofill.naive_inpainting_nearest(sparse[v], m2[v])[0] for v in range(V)])
    xad = sc["xatlas_dict"]
    a0, shr, view_ids, pcoord, _, painted = ounproj.unproject(
        inpainted, sc["f_normals"], res, params, cam_res, base_dirs, xad["gb_pos"], xad["mask"],
        xad["per_atlas_pixel_face_id"], pr["uv_centers"], pr["uv_scales"], 0.05, scales, depth,
        cfg["edge_dilate_kernels"], False)
    ids = np.unique(xad["per_atlas_pixel_face_id"][0][~painted])
    ids = ids[ids > -1]
    a1, tie, _ = onb.paint_invisible_areas_by_neighbors(sc["vertices"], sc["faces"], xad["uvs"],
                                                        xad["mesh_tex_idx"], ids, a0, painted)
    uv_map, vmask = oopt.view_uv_maps(params, sc["vertices"], sc["faces"], xad["uvs"],
                                      xad["mesh_tex_idx"], pr["uv_centers"], pr["uv_scales"], 0.05,
                                      scales, 1024)
    a_in = np.ascontiguousarray(a1.transpose(2, 0, 1)[:, ::-1])
    a2, _ = oopt.optimize_color(a_in, inpainted, uv_map, vmask, shrinked_vis=shr, iterations=100,
                                res=1024)
    ref = a2[0][:, ::-1].transpose(1, 2, 0)
    err = np.abs(atlas - ref)
    mse = float((err.astype(np.float64) ** 2).mean())
    psnr = 10 * np.log10(1.0 / max(mse, 1e-30))
    print(f"default.yaml flow vs oracle: max abs {err.max():.3e}, PSNR {psnr:.1f} dB, "
          f"texels off by >1e-3: {(err.max(-1) > 1e-3).mean():.2e}")
    # observed on B200: max abs 2.5e-3, PSNR 88.3 dB, 2.6e-4 of the texels beyond 1e-3 (the L1 loss has a
    # sign() in its gradient: rounding-level differences flip signs where render == target)
    assert err.max() < 3.8e-3 and psnr > 86.0
    assert (err.max(-1) > 1e-3).mean() < 3.9e-4
    # the metric of BASELINE.json: texture PSNR on the 8-bit atlas image (psnr_ssmi.py:23-42)
    from pointdreamer_b200 import metrics
    p8 = metrics.calculate_psnr(metrics.atlas_to_uint8(atlas), metrics.atlas_to_uint8(ref))
    print(f"8-bit texture PSNR, CUDA flow vs oracle flow: {p8:.1f} dB")
    assert p8 > 70.0  # observed 72.0 dB
