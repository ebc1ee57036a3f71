"""Golden vectors for the ADM U-Net: run the REFERENCE's own UNetModel
(/root/reference/models/DDNM/guided_diffusion/unet.py via script_util.create_model) on CPU with
the seeded synthetic weights of oracle.unet.synthetic_state_dict and store input/output.

Run in the build container only:   python tests/golden/make_golden_unet.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_loader  # noqa: E402
from oracle import unet as ounet  # noqa: E402

SMALL = dict(image_size=64, in_channels=3, model_channels=64, out_channels=6, num_res_blocks=1,
             attention_resolutions="32,16,8", channel_mult=(1, 2, 3, 4), num_head_channels=64,
             num_heads=4, use_scale_shift_norm=True, resblock_updown=True, use_fp16=True,
             use_new_attention_order=False)


def ref_model(cfg, sd, fp16):
    su = ref_loader.load("models.DDNM.guided_diffusion.script_util")
    model = su.create_model(
        image_size=cfg["image_size"], num_channels=cfg["model_channels"],
        num_res_blocks=cfg["num_res_blocks"],
        channel_mult=",".join(str(m) for m in cfg["channel_mult"]),
        learn_sigma=True, class_cond=False, use_checkpoint=False,
        attention_resolutions=cfg["attention_resolutions"], num_heads=cfg["num_heads"],
        num_head_channels=cfg["num_head_channels"], num_heads_upsample=-1,
        use_scale_shift_norm=True, dropout=0.0, resblock_updown=True, use_fp16=fp16,
        use_new_attention_order=False)
    missing, unexpected = model.load_state_dict(sd, strict=True), None
    if fp16:
        model.convert_to_fp16()
    model.eval()
    return model


def main():
    torch.manual_seed(0)
    cfg = SMALL
    sd = ounet.synthetic_state_dict(cfg, seed=1234)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 3, cfg["image_size"], cfg["image_size"], generator=g)
    t = torch.tensor([990.0, 370.0])
    with torch.no_grad():
        y32 = ref_model(cfg, sd, fp16=False)(x, t)
        try:
            y16 = ref_model(cfg, sd, fp16=True)(x, t).float()
        except Exception as e:  # CPU half kernels missing
            print("reference fp16-on-CPU unavailable:", repr(e)[:200])
            y16 = None
    out = dict(x=x.numpy(), t=t.numpy(), y_fp32=y32.numpy())
    if y16 is not None:
        out["y_fp16"] = y16.numpy()
    wsum = float(sum(v.double().abs().sum() for v in sd.values()))
    out["weight_abs_sum"] = np.float64(wsum)
    np.savez_compressed(os.path.join(HERE, "unet_small.npz"), **out)
    print("y32 abs max", float(y32.abs().max()), "std", float(y32.std()))
    o32 = ounet.UNetOracle(sd, cfg, emulate_fp16=False).forward(x, t)
    print("oracle(fp32) vs reference fp32: max abs diff", float((o32 - y32).abs().max()))
    o16 = ounet.UNetOracle(sd, cfg, emulate_fp16=True).forward(x, t)
    print("oracle(fp16-emulated) vs reference fp32: max abs diff", float((o16 - y32).abs().max()))
    if y16 is not None:
        print("oracle(fp16-emulated) vs reference fp16-on-CPU: max abs diff",
              float((o16 - y16).abs().max()), " ref16 vs ref32:", float((y16 - y32).abs().max()))


if __name__ == "__main__":
    main()
