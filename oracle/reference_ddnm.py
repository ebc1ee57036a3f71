"""Oracle (TEST INFRASTRUCTURE): drive the REFERENCE's own, unmodified DDNM sampler and UNetModel.

Everything numerical here is the reference's code, imported through oracle/ref_loader.py from
/root/reference (build container) or its verbatim copy under baseline/_ref/ (GPU box):
  * `Diffusion` / `simplified_ddnm_inpainting`   models/DDNM/guided_diffusion/diffusion.py:80-113, 459-570
  * `create_model` -> `UNetModel`                 script_util.py:130-185, unet.py:396-664
  * `Inpainter.inpaint`                           models/DDNM/ddnm_inpainting.py:29-44
  * the per-view serial loop                      pointdreamer/ours_utils.py:914-929
What this file adds is only what the reference leaves implicit: synthetic weights instead of
the (absent) checkpoint, an explicit seed of the global CUDA generator at DDNM entry, and an
optional recorder around the model call.  Used by tests/test_ddnm_reference_chain_gpu.py and by
bench.py's `gpu_baseline` leg (the reference's execution style on the same B200) - never by the
product path.
"""
import sys

import torch

from . import ref_loader


def _load(name):
    """ref_loader.load switches torch.use_deterministic_algorithms(True) on for the CPU geometry
    goldens (index_put winner).  The reference's DDNM stage runs stock PyTorch - and
    on CUDA the deterministic mode would make its attention einsum (cuBLAS) raise - so the switch
    is put back to the reference's own setting here."""
    mod = ref_loader.load(name)
    torch.use_deterministic_algorithms(False)
    return mod


def _munch(d):
    ref_loader.install_stubs()
    return sys.modules["munch"].Munch.fromDict(d)


def build_model(model_cfg, state_dict, device, fp16=True):
    """The reference's UNetModel for `model_cfg` (keys of pointdreamer_b200.unet
    DEFAULT_MODEL_CONFIG) with `state_dict`, converted like Diffusion.get_model does
    (diffusion.py:436-455: create_model, convert_to_fp16, load_state_dict, eval)."""
    su = _load("models.DDNM.guided_diffusion.script_util")
    model = su.create_model(
        image_size=model_cfg["image_size"], num_channels=model_cfg["model_channels"],
        num_res_blocks=model_cfg["num_res_blocks"],
        channel_mult=",".join(str(m) for m in model_cfg["channel_mult"]),
        learn_sigma=True, class_cond=False, use_checkpoint=False,
        attention_resolutions=model_cfg["attention_resolutions"], num_heads=model_cfg["num_heads"],
        num_head_channels=model_cfg["num_head_channels"], num_heads_upsample=-1,
        use_scale_shift_norm=True, dropout=0.0, resblock_updown=True, use_fp16=fp16,
        use_new_attention_order=False)
    if fp16:
        model.convert_to_fp16()
    model.load_state_dict({k: v.detach().to("cpu") for k, v in state_dict.items()}, strict=True)
    model.to(device)
    model.eval()
    return model


def build_runner(device, image_size=256, T_sampling=100):
    """The reference's `Diffusion` with models/DDNM/configs/imagenet_256.yml's values
    (ddnm_inpainting.py:18-26) for the given image size / number of sampling steps."""
    diff = _load("models.DDNM.guided_diffusion.diffusion")
    cfg = _munch(dict(
        data=dict(dataset="ImageNet", image_size=image_size, channels=3, logit_transform=False,
                  uniform_dequantization=False, gaussian_dequantization=False, random_flip=True,
                  rescaled=True),
        model=dict(type="openai", var_type="fixedsmall"),
        diffusion=dict(beta_schedule="linear", beta_start=0.0001, beta_end=0.02,
                       num_diffusion_timesteps=1000),
        sampling=dict(batch_size=1),
        time_travel=dict(T_sampling=T_sampling, travel_length=1, travel_repeat=1)))
    args = _munch(dict(sigma_y=0, eta=0.85, seed=1234))
    return diff.Diffusion(args, cfg, device=torch.device(device))


class Recorder(torch.nn.Module):
    """Wraps the model handed to simplified_ddnm_inpainting: keeps every x_t it is called with
    and every eps it returns (first 3 channels, what the sampler uses, diffusion.py:529-530)."""

    def __init__(self, model, keep=True):
        super().__init__()
        self.model = model
        self.keep = keep
        self.xs, self.ets = [], []

    def forward(self, x, t):
        y = self.model(x, t)
        if self.keep:
            self.xs.append(x.detach().float().clone())
            self.ets.append(y[:, :3].detach().float().clone())
        return y


def reference_inpainter(runner, model):
    """A reference `Inpainter` (ddnm_inpainting.py:15-44) around an existing runner + model; its
    __init__ (config file, checkpoint download) is skipped, `inpaint` is the reference's."""
    mod = _load("models.DDNM.ddnm_inpainting")
    inp = object.__new__(mod.Inpainter)
    inp.runner = runner
    inp.model = model
    return inp


def run_views(inpainter, sparse_imgs, hard_mask2s, seed=42):
    """ours_utils.py:914-929: the reference's serial per-view loop over `inpainter.inpaint`, with
    the global CUDA generator seeded to (seed, offset 0) at entry (the reference's effective
    seed is kiui.seed_everything(42), demo.py:34).  sparse_imgs, hard_mask2s [V,3,S,S] on the
    GPU -> [V,3,S,S] fp32."""
    torch.cuda.manual_seed(seed)
    outs = []
    with ref_loader.quiet():
        for i in range(sparse_imgs.shape[0]):
            o = inpainter.inpaint(masked_imgs=sparse_imgs[i].permute(1, 2, 0).unsqueeze(0),
                                  masks=hard_mask2s[i].permute(1, 2, 0).unsqueeze(0))[0]
            outs.append(o.float())
    return torch.stack(outs)
