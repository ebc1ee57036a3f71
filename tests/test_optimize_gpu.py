"""CUDA texture optimiser / atlas inputs ("next" rows N1, N4) vs the oracle and the fixture made
by the reference's own optimize_color / xatlas_uvmap_w_face_id.  Through the C ABI."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from make_golden_optimize import CFG, inputs  # noqa: E402
from oracle import camera as ocam  # noqa: E402
from oracle import optimize as oopt  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(os.path.join(HERE, "golden", "optimize_small.npz")))


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_uvmap_w_face_id_bit_exact(cuda, golden):
    from pointdreamer_b200 import extract_texture_map as etm
    sc, *_ = inputs()
    xa = sc["xatlas_dict"]
    par = (np.arange(sc["vertices"].shape[0]), xa["mesh_tex_idx"].astype(np.uint64), xa["uvs"])
    uvs, tex_idx, gb_pos, mask, face_id = etm.xatlas_uvmap_w_face_id(
        None, _t(sc["vertices"], cuda), _t(sc["faces"], cuda), CFG["atlas_res"],
        parametrization=par)
    assert np.array_equal(mask.cpu().numpy(), golden["uvmap_mask"])
    assert np.array_equal(face_id.cpu().numpy(), golden["uvmap_face_id"])
    assert np.array_equal(gb_pos.cpu().numpy(), golden["uvmap_gb_pos"])
    assert np.array_equal(uvs.cpu().numpy(), xa["uvs"])
    assert np.array_equal(tex_idx.cpu().numpy(), xa["mesh_tex_idx"])


def test_face_normals_bit_exact(cuda):
    from pointdreamer_b200 import ours_utils as ou
    sc, *_ = inputs()
    n = ou.face_normals(_t(sc["vertices"], cuda), _t(sc["faces"], cuda)).cpu().numpy()
    assert np.array_equal(n, oopt.face_normals(sc["vertices"], sc["faces"]))


def _run_cuda(cuda, golden, iterations, res, vis=True, want_images=True):
    from pointdreamer_b200 import camera, ours_utils as ou
    sc, imgs, atlas0, vis_np, scale_factors = inputs()
    xa = sc["xatlas_dict"]
    cams, _, _, _ = camera.create_cameras(CFG["view_num"], 1.6, CFG["cam_res"], device=cuda)
    atlas_in = _t(atlas0, cuda).permute(2, 0, 1).flip(1)
    return ou.optimize_color(
        atlas_in, _t(imgs, cuda), _t(sc["vertices"], cuda), _t(sc["faces"], cuda),
        _t(xa["uvs"], cuda), _t(xa["mesh_tex_idx"], cuda), cams, None, None, None,
        _t(golden["uv_centers"], cuda), _t(golden["uv_scales"], cuda), CFG["padding"],
        _t(scale_factors, cuda), None,
        shrinked_per_view_per_pixel_visibility=_t(vis_np, cuda) if vis else None,
        iterations=iterations, res=res, return_images=want_images)


def _run_oracle(golden, iterations, res, vis=True):
    sc, imgs, atlas0, vis_np, scale_factors = inputs()
    xa = sc["xatlas_dict"]
    cams, _, _, _ = ocam.create_cameras(CFG["view_num"], 1.6, CFG["cam_res"])
    uv_map, mask = oopt.view_uv_maps([c.params for c in cams], sc["vertices"], sc["faces"],
                                     xa["uvs"], xa["mesh_tex_idx"], golden["uv_centers"],
                                     golden["uv_scales"], CFG["padding"], scale_factors, res)
    atlas_in = np.ascontiguousarray(atlas0.transpose(2, 0, 1)[:, ::-1])
    return oopt.optimize_color(atlas_in, imgs, uv_map, mask, shrinked_vis=vis_np if vis else None,
                               iterations=iterations, res=res), uv_map, mask


def test_view_uv_map_bit_exact(cuda, golden):
    """rasterise + interpolate + flip at 256^2: identical to the oracle's uv_map / mask."""
    from pointdreamer_b200 import _lib, camera, ours_utils as ou
    import ctypes
    sc, _, _, _, scale_factors = inputs()
    xa = sc["xatlas_dict"]
    res = 256
    cams, _, _, _ = camera.create_cameras(CFG["view_num"], 1.6, CFG["cam_res"], device=cuda)
    V, Vm = CFG["view_num"], sc["vertices"].shape[0]
    pos = torch.empty(V, Vm, 4, device=cuda)
    _lib.call("pdr_project_fixed", camera.stack_params(cams, cuda), _t(sc["vertices"], cuda), Vm, V,
              ctypes.c_double(CFG["padding"]), _t(golden["uv_centers"], cuda).contiguous(),
              _t(golden["uv_scales"], cuda).contiguous(), _t(scale_factors, cuda), pos)
    _, fidx, _, _ = ou.rasterize(pos, _t(sc["faces"], cuda), res, res)
    uv, m = ou.interpolate(_t(xa["uvs"], cuda), pos, _t(sc["faces"], cuda), fidx,
                           _t(xa["mesh_tex_idx"], cuda), flip_y=True, want_mask=True)
    ocams, _, _, _ = ocam.create_cameras(V, 1.6, CFG["cam_res"])
    uv_o, m_o = oopt.view_uv_maps([c.params for c in ocams], sc["vertices"], sc["faces"],
                                  xa["uvs"], xa["mesh_tex_idx"], golden["uv_centers"],
                                  golden["uv_scales"], CFG["padding"], scale_factors, res)
    assert np.array_equal(m.cpu().numpy(), m_o)
    assert np.array_equal(uv.cpu().numpy(), uv_o)


def test_optimize_color_vs_reference_fixture(cuda, golden):
    """Same call as the fixture (the reference's hard-coded 1024^2 renders, 6 iterations).
    Tolerance: 1e-3 abs per channel on the atlas (north-star RGB bound); observed error printed."""
    atlas, images = _run_cuda(cuda, golden, CFG["iterations"], 1024)
    a = atlas.cpu().numpy()
    err = np.abs(a - golden["atlas_out"])
    print("optimize_color vs reference fixture: max abs", err.max(), "mean abs", err.mean())
    assert a.shape == golden["atlas_out"].shape
    assert err.max() < 1e-3
    img = images.cpu().numpy()
    assert img.dtype == np.float64
    ierr = np.abs(img[:, :, ::8, ::8] - golden["images_s8"])
    assert ierr.max() < 1e-3
    assert abs(img.sum() - golden["images_sum"]) < 1e-6 * abs(golden["images_sum"])


@pytest.mark.parametrize("vis", [True, False])
def test_optimize_color_100_iterations_vs_oracle(cuda, golden, vis):
    """Full schedule (100 Adam iterations, StepLR) at 192^2 renders against the oracle."""
    atlas, images = _run_cuda(cuda, golden, 100, 192, vis=vis)
    (a_o, img_o), _, _ = _run_oracle(golden, 100, 192, vis=vis)
    err = np.abs(atlas.cpu().numpy() - a_o)
    mse = float((err.astype(np.float64) ** 2).mean())
    psnr = 10 * np.log10(1.0 / max(mse, 1e-30))
    print(f"100 iterations (vis={vis}): atlas max abs {err.max():.3e}, PSNR {psnr:.1f} dB, "
          f"texels off by >1e-3: {(err > 1e-3).mean():.2e}")
    # the L1 loss has a sign() in its gradient: an fp64-rounding-level difference can flip a sign
    # where render == target, so a few texels may drift by a fraction of one lr step
    # observed on B200: max abs 5.1e-3 / 6.9e-3, PSNR 82.4 / 82.7 dB, 8.0e-4 / 7.6e-4 beyond 1e-3
    assert err.max() < 1.05e-2 and psnr > 80.5
    assert (err > 1e-3).mean() < 1.2e-3
    assert np.abs(images.cpu().numpy() - img_o).max() < 2e-2


def test_optimize_color_deterministic(cuda, golden):
    a1, _ = _run_cuda(cuda, golden, 12, 256, want_images=False)
    a2, _ = _run_cuda(cuda, golden, 12, 256, want_images=False)
    assert torch.equal(a1, a2)


def test_optimize_color_without_visibility_vs_reference_fixture(cuda, golden):
    """optimize_from == 'naive' (no visibility mask), 4 iterations at the reference's 1024^2."""
    atlas, _ = _run_cuda(cuda, golden, 4, 1024, vis=False, want_images=False)
    err = np.abs(atlas.cpu().numpy() - golden["atlas_out_novis"])
    print("optimize_color (no vis) vs reference fixture: max abs", err.max())
    assert err.max() < 1e-3
