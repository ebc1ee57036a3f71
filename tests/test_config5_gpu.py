"""BASELINE.json configs[4] (SURVEY §8d "config 5"): the stress configuration the reference cannot
run as written — dense noisy 100k-point cloud, view_num=16, 512^2 inpainting (ADM 512 preset,
channel_mult 0.5,1,1,2,2,4,4), cam_res 1024, atlas 2048.  No reference parity is possible; this
checks self-consistency of the whole path at that scale (a 2-step chain keeps it short)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_stress_config_runs_end_to_end(cuda):
    from pointdreamer_b200 import demo, synthetic
    from pointdreamer_b200.ddnm_inpainting import DEFAULT_DDNM_CONFIG, Inpainter
    from pointdreamer_b200.unet import DEFAULT_MODEL_CONFIG, channel_mult_for, random_state_dict
    model_cfg = dict(DEFAULT_MODEL_CONFIG, image_size=512, channel_mult=channel_mult_for(512))
    V, res, cam_res, R = 16, 512, 1024, 2048
    cfg = dict(demo.DEFAULT_CONFIG, view_num=V, res=res, cam_res=cam_res, xatlas_texture_res=R,
               complete_unseen_by="unproject", optimize_from=None)
    sc = synthetic.make_scene(100000, seed=9, noise_std=0.005, atlas_res=R)  # generate_1.py:72 noise
    sd = random_state_dict(model_cfg, seed=3, device=cuda)
    inp = Inpainter(cuda, state_dict=sd, model_config=model_cfg,
                    ddnm_config=dict(DEFAULT_DDNM_CONFIG, T_sampling=2), seed=42, offset=0)
    del sd
    cam = demo.prepare_cameras(cfg, cuda)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    xa = {k: t(v) for k, v in sc["xatlas_dict"].items()}
    keys = {k: cfg[k] for k in demo.PATH_CONFIG_KEYS}
    out = demo.colorize_one_mesh(t(sc["xyz"]), t(sc["rgb"]), t(sc["vertices"]), t(sc["faces"]),
                                 t(sc["f_normals"]), xa, cam, device=cuda, save_img_path=None,
                                 inpainter=inp, glctx=None, logger=None, **keys)
    torch.cuda.synchronize()
    atlas = out[4]
    assert atlas.shape == (R, R, 3)
    assert torch.isfinite(atlas).all()
    assert float(atlas.min()) >= 0.0 and float(atlas.max()) <= 1.0
    m = xa["mask"][0, :, :, 0]
    assert (atlas[m].sum(-1) > 0).float().mean() > 0.9  # charts are painted
    print("stress config: arena", inp.model.workspace_bytes / 1e9, "GB; atlas mean",
          float(atlas[m].mean()))
