"""oracle/unet.py against the golden produced by the reference's own UNetModel (CPU)."""
import os

import numpy as np
import torch

from golden_util import GOLDEN_DIR
from oracle import unet as ounet

SMALL = dict(image_size=64, in_channels=3, model_channels=64, out_channels=6, num_res_blocks=1,
             attention_resolutions="32,16,8", channel_mult=(1, 2, 3, 4), num_head_channels=64,
             num_heads=4, use_scale_shift_norm=True, resblock_updown=True, use_fp16=True,
             use_new_attention_order=False)


def test_unet_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN_DIR, "unet_small.npz"))
    sd = ounet.synthetic_state_dict(SMALL, seed=1234)
    wsum = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(wsum - float(g["weight_abs_sum"])) < 1e-6 * wsum, "synthetic weights drifted"
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["t"])
    y = ounet.UNetOracle(sd, SMALL, emulate_fp16=False).forward(x, t).numpy()
    assert np.abs(y - g["y_fp32"]).max() <= 1e-5
    y16 = ounet.UNetOracle(sd, SMALL, emulate_fp16=True).forward(x, t).numpy()
    # fp16 storage emulation stays within fp16 noise of the fp32 reference
    assert np.abs(y16 - g["y_fp32"]).max() <= 1e-2
    if "y_fp16" in g.files:
        assert np.abs(y16 - g["y_fp16"]).max() <= 1e-2


def test_default_spec_counts():
    """Block table of SURVEY §3.3 / Appendix D (enumerated from the reference model)."""
    spec = ounet.build_spec(ounet.DEFAULT_CONFIG)
    flat = [l for blk in spec["input"] + [spec["middle"]] + spec["output"] for l in blk]
    assert len(spec["input"]) == 18 and len(spec["output"]) == 18
    assert sum(1 for l in flat if l[0] == "res") == 42
    assert sum(1 for l in flat if l[0] == "attn") == 16
    sd = None
