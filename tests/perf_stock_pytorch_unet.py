"""Baseline measurement (NOT a pytest file): the reference's U-Net execution style - stock PyTorch
fp16 tensors + cuDNN/ATen kernels, SURVEY §8d "(ii) stock-PyTorch-on-B200" - timed on the same GPU
as the native engine.  It restates UNetModel.forward with REAL half tensors (convert_to_fp16 torso,
GroupNorm32 through fp32, fp32 time embedding / out head) over the oracle's block spec and the same
synthetic weights, checks its output against the native engine, and prints one JSON line.

  python tests/perf_stock_pytorch_unet.py [batch=1] [n=20]

batch 1 is how the reference runs (one view at a time, ours_utils.py:914-929); batch 8 is what a
user could do by hand.  This is a baseline only: nothing in pointdreamer_b200/ uses it.
"""
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import unet as ounet  # noqa: E402


class StockHalfUNet:
    def __init__(self, sd, cfg, dev, channels_last=False):
        self.cfg, self.spec, self.dev, self.cl = cfg, ounet.build_spec(cfg), dev, channels_last
        self.p = {}
        for k, v in sd.items():
            v = v.to(dev)
            torso_conv = v.dim() >= 3 or (k.endswith(".bias") and (k[:-5] + ".weight") in sd
                                          and sd[k[:-5] + ".weight"].dim() >= 3)
            if torso_conv and not k.startswith("out."):
                v = v.half()                       # convert_module_to_f16 touches convs only
                if v.dim() == 3:
                    v = v[..., None]
                if v.dim() == 4 and channels_last:
                    v = v.contiguous(memory_format=torch.channels_last)
            self.p[k] = v

    def gn(self, x, name):
        return F.group_norm(x.float(), 32, self.p[name + ".weight"], self.p[name + ".bias"],
                            eps=1e-5).type(x.dtype)

    def conv(self, x, name, pad):
        return F.conv2d(x, self.p[name + ".weight"], self.p[name + ".bias"], padding=pad)

    def res(self, x, emb, p, cin, cout, up, down):
        h = F.silu(self.gn(x, p + ".in_layers.0"))
        if up:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        elif down:
            h = F.avg_pool2d(h, 2)
            x = F.avg_pool2d(x, 2)
        h = self.conv(h, p + ".in_layers.2", 1)
        e = F.linear(F.silu(emb), self.p[p + ".emb_layers.1.weight"],
                     self.p[p + ".emb_layers.1.bias"]).type(h.dtype)[..., None, None]
        scale, shift = torch.chunk(e, 2, dim=1)
        h = self.gn(h, p + ".out_layers.0") * (1 + scale) + shift
        h = self.conv(F.silu(h), p + ".out_layers.3", 1)
        if cin != cout:
            x = self.conv(x, p + ".skip_connection", 0)
        return x + h

    def attn(self, x, p, ch, heads):
        b, c, hh, ww = x.shape
        xf = x.reshape(b, c, -1)
        n = self.gn(xf, p + ".norm")
        qkv = F.conv1d(n, self.p[p + ".qkv.weight"][..., 0], self.p[p + ".qkv.bias"])
        length = qkv.shape[-1]
        dh = c // heads
        q, k, v = qkv.reshape(b * heads, dh * 3, length).split(dh, dim=1)
        s = 1 / math.sqrt(math.sqrt(dh))
        w = torch.einsum("bct,bcs->bts", q * s, k * s)
        w = torch.softmax(w.float(), dim=-1).type(w.dtype)
        a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, length)
        hp = F.conv1d(a, self.p[p + ".proj_out.weight"][..., 0], self.p[p + ".proj_out.bias"])
        return (xf + hp).reshape(b, c, hh, ww)

    def run(self, layers, prefix, h, emb):
        for j, l in enumerate(layers):
            p = f"{prefix}.{j}"
            if l[0] == "conv":
                h = self.conv(h, p, 1)
            elif l[0] == "res":
                h = self.res(h, emb, p, l[1], l[2], l[3], l[4])
            else:
                h = self.attn(h, p, l[1], l[2])
        return h

    @torch.no_grad()
    def forward(self, x, t):
        mc = self.cfg["model_channels"]
        emb = ounet.timestep_embedding(t.cpu(), mc).to(self.dev)
        emb = F.linear(emb, self.p["time_embed.0.weight"], self.p["time_embed.0.bias"])
        emb = F.linear(F.silu(emb), self.p["time_embed.2.weight"], self.p["time_embed.2.bias"])
        hs = []
        h = x.half()
        if self.cl:
            h = h.contiguous(memory_format=torch.channels_last)
        for i, layers in enumerate(self.spec["input"]):
            h = self.run(layers, f"input_blocks.{i}", h, emb)
            hs.append(h)
        h = self.run(self.spec["middle"], "middle_block", h, emb)
        for i, layers in enumerate(self.spec["output"]):
            h = self.run(layers, f"output_blocks.{i}", torch.cat([h, hs.pop()], dim=1), emb)
        h = h.float()
        h = F.silu(F.group_norm(h, 32, self.p["out.0.weight"], self.p["out.0.bias"], eps=1e-5))
        return F.conv2d(h, self.p["out.2.weight"], self.p["out.2.bias"], padding=1)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    from pointdreamer_b200.unet import UNetEngine, random_state_dict
    cfg = dict(ounet.DEFAULT_CONFIG)
    sd = random_state_dict(cfg, 1234, dev)
    x = torch.randn(B, 3, 256, 256, device=dev)
    t = torch.full((B,), 500.0, device=dev)
    res = {}
    ref_out = None
    for name, cl in (("nchw", False), ("channels_last", True)):
        m = StockHalfUNet(sd, cfg, dev, channels_last=cl)
        for _ in range(3):
            y = m.forward(x, t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            y = m.forward(x, t)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / n
        ref_out = y
        del m
    eng = UNetEngine(sd, cfg, device=dev)
    eng.plan(B)
    for _ in range(3):
        z = eng(x, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        z = eng(x, t)
    e1.record()
    torch.cuda.synchronize()
    ours = e0.elapsed_time(e1) / n
    best = min(res.values())
    print(json.dumps({
        "what": "ADM 256x256 U-Net forward, fp16 torso, synthetic weights, one B200",
        "batch": B, "stock_pytorch_ms": res, "native_engine_ms": ours,
        "speedup_vs_best_stock": best / ours,
        "max_abs_diff_native_vs_stock": float((z - ref_out).abs().max()),
        "stock_shapes_per_s_if_800_forwards_batch1": (1000.0 / (800 * best)) if B == 1 else None,
        "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}))


if __name__ == "__main__":
    main()
