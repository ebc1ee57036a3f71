"""Oracle (TEST INFRASTRUCTURE): functional restatement of the ADM U-Net forward pass.

Restates models/DDNM/guided_diffusion/unet.py (UNetModel 396-664, ResBlock 143-256,
AttentionBlock 259-305, QKVAttentionLegacy 328-358, Upsample 92-110, Downsample 113-140),
nn.py (GroupNorm32 17-19, timestep_embedding 103-121) and script_util.py:130-185
(create_model) as plain CPU torch ops over a state_dict that uses the reference's parameter
names, so the real `256x256_diffusion_uncond.pt` checkpoint loads unchanged.

Precision contract (`emulate_fp16=True`, the reference's `use_fp16: true` torso,
unet.py:619-625 / fp16_util.py:15-22): conv weights/biases and torso activations are rounded
to fp16 at every point the reference stores an fp16 tensor, arithmetic in between is fp32
(what cuDNN/ATen do with fp32 accumulation); GroupNorm, the time embedding MLP, the ResBlock
`emb_layers`, the softmax and the `out` head run in fp32 exactly like the reference.
Pinned against the reference's own UNetModel run on CPU (tests/golden/make_golden_unet.py).
"""
import math

import torch
import torch.nn.functional as F

DEFAULT_CONFIG = dict(  # models/DDNM/configs/imagenet_256.yml:14-33
    image_size=256, in_channels=3, model_channels=256, out_channels=6, num_res_blocks=2,
    attention_resolutions="32,16,8", channel_mult=(1, 1, 2, 2, 4, 4), num_head_channels=64,
    num_heads=4, use_scale_shift_norm=True, resblock_updown=True, use_fp16=True,
    use_new_attention_order=False)


def channel_mult_for(image_size):
    """script_util.py:149-160."""
    return {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4),
            64: (1, 2, 3, 4)}[image_size]


def build_spec(cfg):
    """Block structure exactly as UNetModel.__init__ (unet.py:441-617) builds it.

    Returns dict(input=[...], middle=[...], output=[...]) where each entry is a list of layer
    descriptors ('conv', cin, cout) | ('res', cin, cout, up, down) | ('attn', ch, heads)."""
    mc = cfg["model_channels"]
    mult = cfg["channel_mult"]
    nrb = cfg["num_res_blocks"]
    attn_ds = [cfg["image_size"] // int(r) for r in cfg["attention_resolutions"].split(",")]
    hc = cfg["num_head_channels"]

    def heads(ch):
        return cfg["num_heads"] if hc == -1 else ch // hc

    ch = int(mult[0] * mc)
    inp = [[("conv", cfg["in_channels"], ch)]]
    chans = [ch]
    ds = 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", ch, int(m * mc), False, False)]
            ch = int(m * mc)
            if ds in attn_ds:
                layers.append(("attn", ch, heads(ch)))
            inp.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([("res", ch, ch, False, True)])
            chans.append(ch)
            ds *= 2
    mid = [("res", ch, ch, False, False), ("attn", ch, heads(ch)), ("res", ch, ch, False, False)]
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, int(mc * m), False, False)]
            ch = int(mc * m)
            if ds in attn_ds:
                layers.append(("attn", ch, heads(ch)))
            if level and i == nrb:
                layers.append(("res", ch, ch, True, False))
                ds //= 2
            out.append(layers)
    return dict(input=inp, middle=mid, output=out, final_ch=ch)


def _h(x, emulate):
    return x.half().float() if emulate else x


def timestep_embedding(timesteps, dim, max_period=10000):
    """nn.py:103-121."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def group_norm32(x, w, b, emulate):
    """GroupNorm32 (nn.py:17-19): fp32 statistics and affine, cast back to the input dtype."""
    return _h(F.group_norm(x, 32, w.float(), b.float(), eps=1e-5), emulate)


def silu16(x, emulate):
    return _h(F.silu(x), emulate)


class UNetOracle:
    def __init__(self, state_dict, cfg=None, emulate_fp16=True):
        self.cfg = dict(DEFAULT_CONFIG if cfg is None else cfg)
        self.spec = build_spec(self.cfg)
        self.emulate = emulate_fp16
        self.sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}

    # weights of the fp16 torso are stored in fp16 (convert_module_to_f16 only touches convs)
    def _cw(self, name):
        return _h(self.sd[name + ".weight"], self.emulate), _h(self.sd[name + ".bias"], self.emulate)

    def _conv(self, x, name, pad):
        w, b = self._cw(name)
        if w.dim() == 3:
            w = w[..., None]
        return _h(F.conv2d(x, w, b, padding=pad), self.emulate)

    def _res(self, x, emb, p, cin, cout, up, down):
        e = self.emulate
        h = silu16(group_norm32(x, self.sd[p + ".in_layers.0.weight"],
                                self.sd[p + ".in_layers.0.bias"], e), e)
        if up:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        elif down:
            h = _h(F.avg_pool2d(h, 2), e)
            x = _h(F.avg_pool2d(x, 2), e)
        h = self._conv(h, p + ".in_layers.2", 1)
        emb_out = F.linear(F.silu(emb), self.sd[p + ".emb_layers.1.weight"],
                           self.sd[p + ".emb_layers.1.bias"])
        emb_out = _h(emb_out, e)[..., None, None]
        scale, shift = torch.chunk(emb_out, 2, dim=1)
        h = group_norm32(h, self.sd[p + ".out_layers.0.weight"], self.sd[p + ".out_layers.0.bias"], e)
        # fp16 tensor arithmetic: every op rounds to fp16 (unet.py:248-252)
        h = _h(_h(h * _h(1 + scale, e), e) + shift, e)
        h = silu16(h, e)
        h = self._conv(h, p + ".out_layers.3", 1)
        if cin != cout:
            x = self._conv(x, p + ".skip_connection", 0)
        return _h(x + h, e)

    def _attn(self, x, p, ch, heads):
        e = self.emulate
        b, c, hh, ww = x.shape
        xf = x.reshape(b, c, -1)
        n = group_norm32(xf, self.sd[p + ".norm.weight"], self.sd[p + ".norm.bias"], e)
        wq, bq = self._cw(p + ".qkv")
        qkv = _h(F.conv1d(n, wq, bq), e)
        length = qkv.shape[-1]
        dh = c // heads
        q, k, v = qkv.reshape(b * heads, dh * 3, length).split(dh, dim=1)
        scale = 1 / math.sqrt(math.sqrt(dh))
        w = _h(torch.einsum("bct,bcs->bts", _h(q * scale, e), _h(k * scale, e)), e)
        w = _h(torch.softmax(w.float(), dim=-1), e)
        a = _h(torch.einsum("bts,bcs->bct", w, v), e).reshape(b, -1, length)
        wp, bp = self._cw(p + ".proj_out")
        hproj = _h(F.conv1d(a, wp, bp), e)
        return _h(xf + hproj, e).reshape(b, c, hh, ww)

    def _run(self, layers, prefix, h, emb):
        for j, l in enumerate(layers):
            p = f"{prefix}.{j}"
            if l[0] == "conv":
                h = self._conv(h, p, 1)
            elif l[0] == "res":
                h = self._res(h, emb, p, l[1], l[2], l[3], l[4])
            else:
                h = self._attn(h, p, l[1], l[2])
        return h

    @torch.no_grad()
    def forward(self, x, timesteps):
        """x [B,3,H,W] fp32, timesteps [B] -> [B,out_channels,H,W] fp32 (unet.py:635-664)."""
        sd = self.sd
        mc = self.cfg["model_channels"]
        emb = timestep_embedding(timesteps, mc)
        emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
        emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
        hs = []
        h = _h(x.float(), self.emulate)
        for i, layers in enumerate(self.spec["input"]):
            h = self._run(layers, f"input_blocks.{i}", h, emb)
            hs.append(h)
        h = self._run(self.spec["middle"], "middle_block", h, emb)
        for i, layers in enumerate(self.spec["output"]):
            h = torch.cat([h, hs.pop()], dim=1)
            h = self._run(layers, f"output_blocks.{i}", h, emb)
        h = h.float()
        h = F.silu(F.group_norm(h, 32, sd["out.0.weight"], sd["out.0.bias"], eps=1e-5))
        return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


def synthetic_state_dict(cfg=None, seed=1234, zero_scale=0.5):
    """Seeded synthetic weights with the reference's parameter names and shapes.

    The pretrained checkpoint is absent and the default init zeroes every ResBlock
    output conv, attention proj_out and the final conv (`zero_module`, unet.py:210-212, 294,
    616), which would make all parity checks vacuous, so those are re-randomised too."""
    cfg = dict(DEFAULT_CONFIG if cfg is None else cfg)
    spec = build_spec(cfg)
    g = torch.Generator().manual_seed(seed)
    mc = cfg["model_channels"]
    ted = mc * 4
    sd = {}

    def rnd(*shape, scale):
        return torch.randn(*shape, generator=g) * scale

    def lin(name, cin, cout):
        sd[name + ".weight"] = rnd(cout, cin, scale=1.0 / math.sqrt(cin))
        sd[name + ".bias"] = rnd(cout, scale=0.02)

    def conv(name, cin, cout, k, gain=1.0, dims=2):
        shape = (cout, cin, k, k) if dims == 2 else (cout, cin, k)
        sd[name + ".weight"] = rnd(*shape, scale=gain / math.sqrt(cin * k ** dims))
        sd[name + ".bias"] = rnd(cout, scale=0.02)

    def gn(name, ch):
        sd[name + ".weight"] = 1.0 + rnd(ch, scale=0.1)
        sd[name + ".bias"] = rnd(ch, scale=0.05)

    lin("time_embed.0", mc, ted)
    lin("time_embed.2", ted, ted)

    def block(prefix, layers):
        for j, l in enumerate(layers):
            p = f"{prefix}.{j}"
            if l[0] == "conv":
                conv(p, l[1], l[2], 3)
            elif l[0] == "res":
                cin, cout = l[1], l[2]
                gn(p + ".in_layers.0", cin)
                conv(p + ".in_layers.2", cin, cout, 3)
                lin(p + ".emb_layers.1", ted, 2 * cout)
                sd[p + ".emb_layers.1.weight"] *= 0.3
                gn(p + ".out_layers.0", cout)
                conv(p + ".out_layers.3", cout, cout, 3, gain=zero_scale)
                if cin != cout:
                    conv(p + ".skip_connection", cin, cout, 1)
            else:
                ch = l[1]
                gn(p + ".norm", ch)
                conv(p + ".qkv", ch, 3 * ch, 1, dims=1)
                conv(p + ".proj_out", ch, ch, 1, gain=zero_scale, dims=1)

    for i, layers in enumerate(spec["input"]):
        block(f"input_blocks.{i}", layers)
    block("middle_block", spec["middle"])
    for i, layers in enumerate(spec["output"]):
        block(f"output_blocks.{i}", layers)
    gn("out.0", spec["final_ch"])
    conv("out.2", spec["final_ch"], cfg["out_channels"], 3, gain=zero_scale)
    return sd
