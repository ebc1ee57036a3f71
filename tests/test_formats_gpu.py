"""File-level flow ("next" row N3): PLY in, the reference's output tree out, through
recon_one_textured_mesh on the GPU; plus the 8-bit atlas quantisation against numpy."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_atlas_to_u8_matches_reference_quantisation(cuda):
    from pointdreamer_b200 import _lib
    R = 64
    g = torch.Generator().manual_seed(0)
    atlas = (torch.rand(R, R, 3, generator=g) * 1.4 - 0.2)
    mask = torch.rand(R, R, generator=g) < 0.6
    rgb = torch.empty(R, R, 3, dtype=torch.uint8, device=cuda)
    rgba = torch.empty(R, R, 4, dtype=torch.uint8, device=cuda)
    _lib.call("pdr_atlas_to_u8", atlas.to(cuda), mask.to(torch.uint8).to(cuda), R, rgb, rgba)
    img = np.asarray(atlas.numpy(), dtype=np.float32)
    img = (img - 0) * (255 / (1 - 0))                       # demo.py:283-285
    img = img.clip(0, 255).astype(np.uint8)                  # demo.py:296
    assert np.array_equal(rgb.cpu().numpy(), img[::-1])      # demo.py:299
    cat_mask = (mask.long() * 255).numpy().astype(np.uint8)[..., None]
    assert np.array_equal(rgba.cpu().numpy(), np.concatenate([img, cat_mask], -1)[::-1])


def test_recon_one_textured_mesh_output_tree(cuda, tmp_path):
    from PIL import Image
    from pointdreamer_b200 import demo, io_utils, synthetic
    R = 256
    sc = synthetic.make_scene(5000, seed=4, nu=20, nv=16, atlas_res=R, charts=(2, 2))
    pc = str(tmp_path / "shape.ply")
    io_utils.save_colored_pc_ply(sc["xyz"], sc["rgb"], pc)
    xyz_back, rgb_back = io_utils.read_ply_xyzrgb(pc)
    assert np.array_equal(xyz_back, sc["xyz"])
    assert np.array_equal(rgb_back, (sc["rgb"] * 255).astype(np.uint8))
    xa = sc["xatlas_dict"]
    mesh = str(tmp_path / "shape_untextured_mesh.obj")
    io_utils.savemeshtes2(sc["vertices"], xa["uvs"], sc["faces"], xa["mesh_tex_idx"], mesh)
    v, uv, f, ft = io_utils.loadobjtex(mesh)
    assert np.array_equal(f, sc["faces"]) and np.array_equal(ft, xa["mesh_tex_idx"])
    assert np.abs(v - sc["vertices"]).max() < 1e-6 and np.abs(uv - xa["uvs"]).max() < 1e-6

    cfg = dict(demo.DEFAULT_CONFIG, view_num=2, res=64, cam_res=128, xatlas_texture_res=R,
               texture_gen_method="nearest", edge_dilate_kernels=[5], output_path=str(tmp_path / "out"))
    cam = demo.prepare_cameras(cfg, cuda)
    root = demo.recon_one_textured_mesh(cfg, None, cam, pc, "shape", cuda)
    for rel in ("input_pc.ply", f"geo/xatlas_{R}.pth", "models/model_normalized.obj",
                "models/model_normalized.mtl", "models/model_normalized.png",
                "others/atlas_wo_background.png", "others/0_sparse.png", "others/0_mask0.png",
                "others/0_mask2.png", "others/1_inpainted.png"):
        assert os.path.exists(os.path.join(root, rel)), rel
    png = np.asarray(Image.open(os.path.join(root, "models", "model_normalized.png")))
    assert png.shape == (R, R, 3) and png.dtype == np.uint8 and png.std() > 5
    rgba = np.asarray(Image.open(os.path.join(root, "others", "atlas_wo_background.png")))
    assert rgba.shape == (R, R, 4) and set(np.unique(rgba[..., 3])) <= {0, 255}
    with open(os.path.join(root, "models", "model_normalized.mtl")) as fh:
        assert fh.read().endswith("map_Kd model_normalized.png\n")
    # second run: the cached xatlas dict and the cached {i}_inpainted.png are reused (demo.py:138-147,
    # 430-438) and give the same texture up to the 8-bit PNG round trip of the views
    root2 = demo.recon_one_textured_mesh(cfg, None, cam, pc, "shape", cuda)
    png2 = np.asarray(Image.open(os.path.join(root2, "models", "model_normalized.png")))
    assert np.abs(png2.astype(int) - png.astype(int)).mean() < 2.0
