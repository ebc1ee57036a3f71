"""Oracle (oracle/optimize.py) vs the fixture produced by the reference's own optimize_color /
xatlas_uvmap_w_face_id (tests/golden/make_golden_optimize.py).  CPU only."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from make_golden_optimize import CFG, inputs  # noqa: E402
from oracle import camera as ocam  # noqa: E402
from oracle import optimize as oopt  # noqa: E402


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(os.path.join(HERE, "golden", "optimize_small.npz")))


def test_uvmap_w_face_id_matches_reference(golden):
    sc, *_ = inputs()
    xa = sc["xatlas_dict"]
    gb_pos, mask, face_id = oopt.uvmap_w_face_id(sc["vertices"], sc["faces"], xa["uvs"],
                                                 xa["mesh_tex_idx"], CFG["atlas_res"])
    assert np.array_equal(mask, golden["uvmap_mask"])
    assert np.array_equal(face_id, golden["uvmap_face_id"])
    assert np.array_equal(gb_pos, golden["uvmap_gb_pos"])
    # sanity against the analytic atlas of the synthetic scene (different rasterisation rule at
    # chart borders, same surface): interior texels agree on the face and nearly on the position
    both = mask[..., 0] & xa["mask"][..., 0]
    assert both.sum() > 0.9 * xa["mask"].sum()
    same_face = (face_id == xa["per_atlas_pixel_face_id"])[both].mean()
    assert same_face > 0.97
    assert np.abs(gb_pos - xa["gb_pos"])[both].max() < 0.05


def test_face_normals():
    sc, *_ = inputs()
    n = oopt.face_normals(sc["vertices"], sc["faces"])
    assert np.abs(np.linalg.norm(n, axis=1) - 1).max() < 1e-5
    assert np.abs(n - sc["f_normals"]).max() < 1e-4


def test_optimize_color_matches_reference(golden):
    sc, imgs, atlas0, vis, scale_factors = inputs()
    xa = sc["xatlas_dict"]
    cams, _, _, _ = ocam.create_cameras(CFG["view_num"], 1.6, CFG["cam_res"])
    params = [c.params for c in cams]
    uv_map, mask = oopt.view_uv_maps(params, sc["vertices"], sc["faces"], xa["uvs"],
                                     xa["mesh_tex_idx"], golden["uv_centers"], golden["uv_scales"],
                                     CFG["padding"], scale_factors, 1024)
    atlas_in = np.ascontiguousarray(atlas0.transpose(2, 0, 1)[:, ::-1])
    atlas, images = oopt.optimize_color(atlas_in, imgs, uv_map, mask, shrinked_vis=vis,
                                        iterations=CFG["iterations"], res=1024)
    # same torch ops on the same machine: identical
    assert np.abs(atlas - golden["atlas_out"]).max() == 0.0
    assert np.abs(images[:, :, ::8, ::8] - golden["images_s8"]).max() == 0.0
    assert abs(images.sum() - golden["images_sum"]) <= 1e-9 * abs(golden["images_sum"])
    # the optimiser did move the covered texels
    assert np.abs(atlas[0] - atlas_in).max() > 0.05


def test_optimize_color_without_visibility_matches_reference(golden):
    """optimize_from == 'naive': no shrinked-visibility mask (demo.py:221-223), 4 iterations."""
    sc, imgs, atlas0, vis, scale_factors = inputs()
    xa = sc["xatlas_dict"]
    cams, _, _, _ = ocam.create_cameras(CFG["view_num"], 1.6, CFG["cam_res"])
    uv_map, mask = oopt.view_uv_maps([c.params for c in cams], sc["vertices"], sc["faces"],
                                     xa["uvs"], xa["mesh_tex_idx"], golden["uv_centers"],
                                     golden["uv_scales"], CFG["padding"], scale_factors, 1024)
    atlas_in = np.ascontiguousarray(atlas0.transpose(2, 0, 1)[:, ::-1])
    atlas, images = oopt.optimize_color(atlas_in, imgs, uv_map, mask, shrinked_vis=None,
                                        iterations=4, res=1024)
    assert np.abs(atlas - golden["atlas_out_novis"]).max() == 0.0
    assert abs(images.sum() - golden["images_sum_novis"]) <= 1e-9 * abs(golden["images_sum_novis"])
