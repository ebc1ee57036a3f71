"""DDNM inpainting — mirrors models/DDNM/ddnm_inpainting.py (Inpainter, 15-44) and
models/DDNM/guided_diffusion/diffusion.py (Diffusion.__init__ 80-113, get_model 435-457,
simplified_ddnm_inpainting 459-570) on top of the native engine in libpdr.so.

Differences from the reference are operational, not numerical: all chains of a call run as ONE
batch through the U-Net, the whole T_sampling-step loop is a single C-ABI call without host
synchronisation (the reference moves x to the CPU and back every step, diffusion.py:554-555),
and the noise stream position is explicit (`seed`, `offset`) instead of "whatever the global
CUDA generator is at" (SURVEY §8a D1, Appendix C).
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from .unet import DEFAULT_MODEL_CONFIG, UNetEngine, random_state_dict

# models/DDNM/configs/imagenet_256.yml + ddnm_inpainting.py:20-24
DEFAULT_DDNM_CONFIG = dict(beta_schedule="linear", beta_start=0.0001, beta_end=0.02,
                           num_diffusion_timesteps=1000, T_sampling=100, travel_length=1,
                           travel_repeat=1, sigma_y=0.0, eta=0.85)


def get_schedule_jump(T_sampling, travel_length, travel_repeat):
    """diffusion.py:770-791 (RePaint schedule)."""
    jumps = {}
    for j in range(0, T_sampling - travel_length, travel_length):
        jumps[j] = travel_repeat - 1
    t = T_sampling
    ts = []
    while t >= 1:
        t = t - 1
        ts.append(t)
        if jumps.get(t, 0) > 0:
            jumps[t] = jumps[t] - 1
            for _ in range(travel_length):
                t = t + 1
                ts.append(t)
    ts.append(-1)
    return ts


def step_table(cfg):
    """Timesteps and the 7 fp32 coefficients of every reverse step (diffusion.py:515-552),
    computed on the host once: (sqrt(1-at), sqrt(at), sqrt(at_next), gamma_t, c1, c2, lambda_t)."""
    if cfg["beta_schedule"] != "linear" or cfg["travel_repeat"] != 1:
        raise NotImplementedError("only the linear schedule without time travel is on the path")
    f32 = np.float32
    n = cfg["num_diffusion_timesteps"]
    betas = np.linspace(cfg["beta_start"], cfg["beta_end"], n, dtype=np.float64).astype(f32)
    one_minus = (f32(1) - np.concatenate([np.zeros(1, f32), betas])).astype(f32)
    acp = np.cumprod(one_minus.astype(np.float64)).astype(f32)  # compute_alpha, index t+1
    skip = n // cfg["T_sampling"]
    times = get_schedule_jump(cfg["T_sampling"], cfg["travel_length"], cfg["travel_repeat"])
    sigma_y = f32(2 * cfg["sigma_y"])
    eta = cfg["eta"]
    ts, coefs = [], []
    for i, j in zip(times[:-1], times[1:]):
        i, j = i * skip, j * skip
        if j < 0:
            j = -1
        at, at_next = acp[i + 1], acp[j + 1]
        sigma_t = np.sqrt(f32(1) - at_next * at_next, dtype=f32)
        if sigma_t >= at_next * sigma_y:
            lambda_t = f32(1.0)
            gamma_t = np.sqrt(sigma_t * sigma_t - (at_next * sigma_y) * (at_next * sigma_y), dtype=f32)
        else:
            lambda_t = f32(sigma_t / (at_next * sigma_y))
            gamma_t = f32(0.0)
        c1 = np.sqrt(f32(1) - at_next, dtype=f32) * f32(eta)
        c2 = np.sqrt(f32(1) - at_next, dtype=f32) * f32((1 - eta ** 2) ** 0.5)
        ts.append(f32(i))
        coefs.append([np.sqrt(f32(1) - at, dtype=f32), np.sqrt(at, dtype=f32),
                      np.sqrt(at_next, dtype=f32), gamma_t, c1, c2, lambda_t])
    return np.asarray(ts, dtype=f32), np.ascontiguousarray(np.asarray(coefs, dtype=f32))


class Inpainter:
    """`Inpainter(device).inpaint(masked_imgs, masks)` like the reference, plus `inpaint_batch`.

    state_dict: weights under the reference's parameter names; default: load `ckpt_path`
    (models/DDNM/256x256_diffusion_uncond.pt).  A missing checkpoint is an error, as in the
    reference (diffusion.py:435-457 downloads or fails); seeded random-init weights of the same
    architecture are used only on the explicit opt-in `allow_random_weights=True` (benchmarks
    and tests: there is no network for the checkpoint here).  `synthetic_weights` tells which.
    seed / offset: position of torch's Philox stream at DDNM entry (the reference uses the
    global CUDA generator, effectively kiui.seed_everything(42), demo.py:34)."""

    def __init__(self, device, state_dict=None, model_config=None, ddnm_config=None, seed=42,
                 offset=0, ckpt_path='models/DDNM/256x256_diffusion_uncond.pt',
                 allow_random_weights=False):
        self.device = torch.device(device)
        self.model_config = dict(DEFAULT_MODEL_CONFIG if model_config is None else model_config)
        self.ddnm_config = dict(DEFAULT_DDNM_CONFIG if ddnm_config is None else ddnm_config)
        self.synthetic_weights = False
        if state_dict is None:
            if os.path.exists(ckpt_path):
                state_dict = torch.load(ckpt_path, map_location="cpu")
            elif allow_random_weights:
                state_dict = random_state_dict(self.model_config, seed=1234, device=self.device)
                self.synthetic_weights = True
            else:
                raise FileNotFoundError(
                    f"DDNM checkpoint not found at {os.path.abspath(ckpt_path)!r} (the reference "
                    "loads models/DDNM/256x256_diffusion_uncond.pt, diffusion.py:435-457); pass "
                    "ckpt_path=, state_dict=, or allow_random_weights=True for seeded random-init "
                    "weights of the same architecture")
        self.model = UNetEngine(state_dict, self.model_config, device=self.device)
        self.ts, self.coefs = step_table(self.ddnm_config)
        self.seed = int(seed)
        self.offset = int(offset)
        self.chains_done = 0  # advances like the reference's global generator would
        self._t_dev = {}

    def _t_table(self, V):
        if V not in self._t_dev:
            t = torch.from_numpy(self.ts)[:, None].repeat(1, V).contiguous().to(self.device)
            self._t_dev[V] = t
        return self._t_dev[V]

    def inpaint_batch(self, masked_imgs, masks, chain0=None):
        """masked_imgs [V,3,S,S] fp32 in [0,1]; masks [V,S,S] (1 = known) -> [V,3,S,S] in [0,1].
        Chain v reproduces the reference's (chain0+v)-th serial `inpaint` call bit for bit in
        its noise stream."""
        V, _, S, _ = masked_imgs.shape
        if S != self.model_config["image_size"]:
            raise ValueError(f"x_T is {self.model_config['image_size']}^2 (diffusion.py:493-499); "
                             f"got {S}^2 inputs")
        if chain0 is None:
            chain0 = self.chains_done
            self.chains_done += V
        self.model.plan(V)
        dev = self.device
        steps = len(self.ts)
        x = torch.empty(V, 3, S, S, device=dev)
        y = torch.empty_like(x)
        et = torch.empty_like(x)
        out = torch.empty_like(x)
        coef = self.coefs.ctypes.data_as(ctypes.c_void_p)
        _lib.call("pdr_ddnm_sample", self.model.handle, masked_imgs.float().contiguous(),
                  masks.float().contiguous(), V, steps, coef, self._t_table(V),
                  ctypes.c_ulonglong(self.seed), ctypes.c_ulonglong(self.offset),
                  ctypes.c_ulonglong(steps + 1), int(chain0), x, y, et, out)
        return out

    def inpaint(self, masked_imgs, masks):
        """ddnm_inpainting.py:29-44: masked_imgs [1,H,W,3], masks [1,H,W,3] -> [1,3,H,W]."""
        imgs = masked_imgs.permute(0, 3, 1, 2)
        return self.inpaint_batch(imgs, masks[:, :, :, 0])
