"""Golden vectors for the DDNM sampler: run the REFERENCE's own
Diffusion.simplified_ddnm_inpainting (/root/reference/models/DDNM/guided_diffusion/diffusion.py)
on CPU with the reference's UNetModel (small config, seeded synthetic weights, fp32).

The reference hard-codes `.to('cuda')` (diffusion.py:525,559); for this CPU run torch.Tensor.to
is wrapped so that a 'cuda' target is a no-op.  Noise comes from torch's global CPU generator
seeded with SEED right before the call, exactly in the reference's draw order.

Run in the build container only:   python tests/golden/make_golden_ddnm.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from oracle import ref_loader  # noqa: E402
from oracle import unet as ounet  # noqa: E402
from make_golden_unet import SMALL, ref_model  # noqa: E402

SEED = 1234
T_SAMPLING = 10


def main():
    diff = ref_loader.load("models.DDNM.guided_diffusion.diffusion")
    Munch = sys.modules["munch"].Munch
    cfg = Munch.fromDict(dict(
        data=dict(dataset="ImageNet", image_size=SMALL["image_size"], channels=3,
                  logit_transform=False, uniform_dequantization=False,
                  gaussian_dequantization=False, random_flip=True, rescaled=True),
        model=dict(type="openai", var_type="fixedsmall"),
        diffusion=dict(beta_schedule="linear", beta_start=0.0001, beta_end=0.02,
                       num_diffusion_timesteps=1000),
        sampling=dict(batch_size=1),
        time_travel=dict(T_sampling=T_SAMPLING, travel_length=1, travel_repeat=1)))
    args = Munch.fromDict(dict(sigma_y=0, eta=0.85, seed=1234))
    runner = diff.Diffusion(args, cfg, device=torch.device("cpu"))
    sd = ounet.synthetic_state_dict(SMALL, seed=1234)
    model = ref_model(SMALL, sd, fp16=False)

    S = SMALL["image_size"]
    g = torch.Generator().manual_seed(5)
    V = 2
    imgs = torch.rand(V, 3, S, S, generator=g)
    masks = (torch.rand(V, S, S, generator=g) < 0.3).float()
    sparse = imgs * masks[:, None]

    orig_to = torch.Tensor.to

    def to_patched(self, *a, **k):
        if a and isinstance(a[0], str) and a[0].startswith("cuda"):
            return self
        return orig_to(self, *a, **k)

    torch.Tensor.to = to_patched
    try:
        torch.manual_seed(SEED)
        outs = []
        with torch.no_grad():
            for v in range(V):  # the reference's serial per-view loop (ours_utils.py:916-923)
                o = runner.simplified_ddnm_inpainting(
                    model, sparse[v][None, None], masks[v][None])  # [1,1,3,S,S], [1,S,S]
                outs.append(o[0, 0])
    finally:
        torch.Tensor.to = orig_to
    out = torch.stack(outs).numpy()

    # reference schedule / alpha values for the default 100-step configuration
    times = diff.get_schedule_jump(100, 1, 1)
    betas = torch.from_numpy(diff.get_beta_schedule("linear", beta_start=0.0001, beta_end=0.02,
                                                    num_diffusion_timesteps=1000)).float()
    alphas = torch.stack([diff.compute_alpha(betas, torch.tensor([t * 10 if t >= 0 else -1]))
                          .reshape(()) for t in times]).numpy()
    np.savez_compressed(os.path.join(HERE, "ddnm_small.npz"), sparse=sparse.numpy(),
                        masks=masks.numpy(), out=out, times100=np.asarray(times),
                        alphas100=alphas, seed=np.int64(SEED), T_sampling=np.int64(T_SAMPLING))
    print("out range", out.min(), out.max(), "known-pixel max err",
          float(np.abs(out - sparse.numpy())[np.broadcast_to(masks[:, None].numpy() > 0, out.shape)].max()))

    # check the oracle right away
    from oracle import ddnm as oddnm
    o32 = ounet.UNetOracle(sd, SMALL, emulate_fp16=False)
    fn = lambda x, t: o32.forward(torch.from_numpy(x), torch.from_numpy(t)).numpy()
    noise = oddnm.torch_cpu_noise_stream(SEED, V, T_SAMPLING + 1, (3, S, S))
    mine = oddnm.sample(fn, sparse.numpy(), masks.numpy(), noise, T_sampling=T_SAMPLING)
    print("oracle vs reference sampler: max abs diff", float(np.abs(mine - out).max()))


if __name__ == "__main__":
    main()
