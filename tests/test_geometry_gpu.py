"""CUDA geometry path (through the C ABI / the reference-shaped Python operators) against
(a) golden vectors produced by the reference's own source and (b) the numpy oracle on fresh
seeded inputs.  Integer / mask / index outputs must be bit-exact."""
import numpy as np
import pytest
import torch

from golden_util import load_geom_case

pytestmark = pytest.mark.gpu


def _t(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _run_pipeline(cfg, sc, dev, inpainted_override=None):
    from pointdreamer_b200 import camera, ours_utils, unproject as unproj
    V, res, cam_res = cfg["view_num"], cfg["res"], cfg["cam_res"]
    cams, base_dirs, eyes, ups = camera.create_cameras(V, 1.6, cam_res, device=dev)
    coords, colors = _t(sc["xyz"], dev), _t(sc["rgb"], dev)
    vertices, faces, f_normals = _t(sc["vertices"], dev), _t(sc["faces"], dev), _t(sc["f_normals"], dev)
    xa = {k: _t(v, dev) for k, v in sc["xatlas_dict"].items()}
    out = {}
    (hm_cam, face_idxs, depths, vuv, uv_centers, uv_scales, padding, puv, pdepth) = \
        ours_utils.get_rendered_hard_mask_and_face_idx_batch(
            cams, vertices, faces, coords, glctx=None, rescale=cfg["crop_img"],
            padding=cfg["crop_padding"])
    out.update(hard_masks_cam=hm_cam, face_idxs=face_idxs, mesh_depths=depths, vertice_uvs=vuv,
               point_uvs=puv, point_depths=pdepth)
    if cfg["crop_img"]:
        out.update(uv_centers=uv_centers, uv_scales=uv_scales)
    hm = ours_utils.resize_hard_masks(hm_cam, res)
    out["hard_masks"] = hm
    pv, pix_cam = ours_utils.get_point_validation_by_depth(cam_res, puv, pdepth, depths, offset=0.0001)
    out.update(point_validation=pv, point_pixels_cam=pix_cam)
    if cfg.get("use_o3d"):  # demo.py:108-110
        pv2 = ours_utils.get_point_validation_by_o3d(coords, eyes, 100)
        out["point_validation_o3d"] = pv2
        pv = torch.logical_or(pv, pv2)
    pp = ours_utils.get_point_pixels(puv, res)
    out["point_pixels"] = pp
    sparse, m0, m2, scales = ours_utils.get_sparse_images(
        pp, colors, pv, hm, None, V, res, cfg["point_size"], cfg["edge_point_size"],
        cfg["mask_ratio_thresh"])
    out.update(sparse_imgs=sparse, hard_mask0s=m0, hard_mask2s=m2, scale_factors=scales)
    inpainted = ours_utils.get_inpainted_images(sparse, m0, m2, None, None, V, method="nearest")
    out["inpainted_nearest"] = inpainted
    src = inpainted if inpainted_override is None else _t(inpainted_override, dev)
    atlas, shr, view_ids, pcoord, points, painted = unproj.unproject(
        src, vertices, f_normals, res, cams, cam_res, base_dirs, xa["gb_pos"], xa["mask"],
        xa["per_atlas_pixel_face_id"], uv_centers, uv_scales, padding, scales, depths,
        cfg["edge_dilate_kernels"], None, cfg["complete_unseen_by_projection"])
    out.update(atlas_img=atlas, shrinked_vis=shr, point_view_ids=view_ids,
               points_atlas_pixel_coord=pcoord, atlas_points=points, atlas_painted_mask=painted)
    out["atlas_dilated"] = unproj.dilate_atlas(atlas, xa["mask"])
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


EXACT = ["hard_masks_cam", "face_idxs", "mesh_depths", "vertice_uvs", "point_uvs", "point_depths",
         "uv_centers", "uv_scales", "hard_masks", "point_validation", "point_pixels_cam",
         "point_pixels", "sparse_imgs", "hard_mask0s", "hard_mask2s", "scale_factors",
         "shrinked_vis", "points_atlas_pixel_coord", "atlas_points", "point_validation_o3d"]


@pytest.mark.parametrize("name", ["a", "b", "c", "clock"])
def test_geometry_vs_reference_golden(cuda, name):
    from oracle import fill as ofill
    cfg, sc, g = load_geom_case(name)
    # feed the golden's inpainted views to unproject so tie pixels of the fill don't cascade
    got = _run_pipeline(cfg, sc, cuda, inpainted_override=g["inpainted_nearest"])
    for k in EXACT:
        if k not in g or k not in got:
            continue
        a, b = got[k], g[k]
        assert a.shape == tuple(b.shape), (k, a.shape, b.shape)
        nbad = int((a != b).sum())
        assert nbad == 0, f"{k}: {nbad} mismatching elements of {a.size}"
    # nearest fill vs scipy away from ties; exact vs the oracle rule everywhere
    for i in range(cfg["view_num"]):
        o, tie = ofill.naive_inpainting_nearest(g["sparse_imgs"][i], g["hard_mask2s"][i])
        assert np.array_equal(got["inpainted_nearest"][i], o)
        diff = (got["inpainted_nearest"][i] != g["inpainted_nearest"][i]).any(0)
        assert not (diff & ~tie).any()
    mism = got["point_view_ids"] != g["point_view_ids"]
    assert int(mism.sum()) == 0, f"case {name}: {int(mism.sum())} view-id mismatches of {mism.size}"
    assert np.array_equal(got["atlas_img"], g["atlas_img"])
    assert np.array_equal(got["atlas_painted_mask"], g["atlas_painted_mask"])
    o, tie = ofill.dilate_atlas(got["atlas_img"], sc["xatlas_dict"]["mask"])
    assert np.array_equal(got["atlas_dilated"], o)


@pytest.mark.parametrize("seed,n_points,V,res,cam_res,R", [
    (11, 5000, 4, 128, 256, 256),
    (12, 30000, 8, 256, 512, 512),
])
def test_geometry_vs_oracle(cuda, seed, n_points, V, res, cam_res, R):
    """Fresh seeded scene at (closer to) production size, CUDA vs the numpy oracle."""
    from oracle import camera as ocam, project as oproj, unproject as ounproj, fill as ofill
    from pointdreamer_b200 import synthetic
    nu = 40 if n_points < 10000 else 72
    cfg = dict(n_points=n_points, seed=seed, nu=nu, nv=nu, atlas_res=R, charts=(3, 3), view_num=V,
               res=res, cam_res=cam_res, point_size=1, edge_point_size=1, crop_img=True,
               crop_padding=0.05, mask_ratio_thresh=0.82, edge_dilate_kernels=[21 * R // 1024 | 1],
               complete_unseen_by_projection=False)
    sc = synthetic.make_scene(n_points, seed, nu, nu, R, charts=(3, 3))
    got = _run_pipeline(cfg, sc, cuda)
    cams, base_dirs, _, _ = ocam.create_cameras(V, 1.6, cam_res)
    params = [c.params for c in cams]
    pr = oproj.project_vertices_points(params, sc["vertices"], sc["xyz"], True, 0.05)
    for k in ["point_uvs", "point_depths", "vertice_uvs", "uv_centers", "uv_scales"]:
        assert np.array_equal(got[k], pr[k]), k
    depth, fidx, mask = oproj.rasterize(pr["pos"], sc["faces"], cam_res)
    assert np.array_equal(got["hard_masks_cam"], mask)
    assert np.array_equal(got["face_idxs"], fidx)
    assert np.array_equal(got["mesh_depths"], depth)
    hm = oproj.resize_mask_half_any(mask, res)
    assert np.array_equal(got["hard_masks"], hm)
    vis, pix = oproj.point_validation_by_depth(cam_res, pr["point_uvs"], pr["point_depths"], depth, 0.0001)
    assert np.array_equal(got["point_validation"], vis)
    pp = oproj.point_pixels(pr["point_uvs"], res)
    assert np.array_equal(got["point_pixels"], pp)
    sparse, m0, m2, scales = oproj.get_sparse_images(pp, sc["rgb"], vis, hm, V, res, 1, 1, 0.82)
    assert np.array_equal(got["scale_factors"], scales)
    assert np.array_equal(got["hard_mask0s"], m0)
    assert np.array_equal(got["hard_mask2s"], m2)
    assert np.array_equal(got["sparse_imgs"], sparse)
    for i in range(V):
        o, _ = ofill.naive_inpainting_nearest(sparse[i], m2[i])
        assert np.array_equal(got["inpainted_nearest"][i], o)
    xa = sc["xatlas_dict"]
    atlas, shr, view_ids, pcoord, points, painted = ounproj.unproject(
        got["inpainted_nearest"], sc["f_normals"], res, params, cam_res, base_dirs, xa["gb_pos"],
        xa["mask"], xa["per_atlas_pixel_face_id"], pr["uv_centers"], pr["uv_scales"], 0.05, scales,
        depth, cfg["edge_dilate_kernels"], False)
    assert np.array_equal(got["shrinked_vis"], shr)
    assert np.array_equal(got["points_atlas_pixel_coord"], pcoord)
    assert np.array_equal(got["atlas_points"], points)
    mism = got["point_view_ids"] != view_ids
    assert int(mism.sum()) == 0, f"{int(mism.sum())} view-id mismatches of {mism.size}"
    assert np.array_equal(got["atlas_painted_mask"], painted)
    assert np.array_equal(got["atlas_img"], atlas)


def test_unproject_without_crop_parameters(cuda):
    """unproject.py:262-264: with uv_centers / uv_scales / inpaint_scale_factors / padding = None the
    texel uv is plain ndc*0.5+0.5 and the crop arrays are never touched (NULL pointers in the ABI)."""
    from oracle import camera as ocam, unproject as ounproj
    from pointdreamer_b200 import camera, unproject as unproj
    cfg, sc, g = load_geom_case("c")
    V, res, cam_res = cfg["view_num"], cfg["res"], cfg["cam_res"]
    cams, base_dirs, _, _ = camera.create_cameras(V, 1.6, cam_res, device=cuda)
    xa = {k: _t(v, cuda) for k, v in sc["xatlas_dict"].items()}
    views = _t(g["inpainted_nearest"], cuda)
    depths = _t(g["mesh_depths"], cuda)
    got = unproj.unproject(views, _t(sc["vertices"], cuda), _t(sc["f_normals"], cuda), res, cams,
                           cam_res, base_dirs, xa["gb_pos"], xa["mask"],
                           xa["per_atlas_pixel_face_id"], None, None, None, None, depths,
                           cfg["edge_dilate_kernels"], None, False)
    torch.cuda.synchronize()
    ocams, obase, _, _ = ocam.create_cameras(V, 1.6, cam_res)
    xan = sc["xatlas_dict"]
    want = ounproj.unproject(g["inpainted_nearest"], sc["f_normals"], res, [c.params for c in ocams],
                             cam_res, obase, xan["gb_pos"], xan["mask"],
                             xan["per_atlas_pixel_face_id"], None, None, None, None,
                             g["mesh_depths"], cfg["edge_dilate_kernels"], False)
    for a, b, name in zip(got, want, ["atlas", "shrinked_vis", "view_ids", "coords", "points", "painted"]):
        assert np.array_equal(a.cpu().numpy(), b), name
