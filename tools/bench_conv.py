"""Micro-benchmark of the tcgen05 conv kernel on the U-Net's dominant layer shapes (GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointdreamer_b200 import _lib

dev = torch.device("cuda:0")
shapes = [  # B,H,W,C1,C2,Cout,taps
    (8, 64, 64, 256, 0, 512, 9),
    (8, 64, 64, 1024, 0, 512, 9),
    (8, 32, 32, 1024, 0, 512, 9),
    (8, 16, 16, 2048, 0, 1024, 9),
    (8, 256, 256, 256, 0, 256, 9),
    (8, 256, 256, 256, 256, 256, 9),
    (8, 128, 128, 256, 0, 256, 9),
    (8, 128, 128, 512, 0, 512, 9),
    (8, 64, 64, 512, 0, 512, 9),
    (8, 32, 32, 512, 0, 512, 9),
    (8, 32, 32, 1024, 0, 1024, 9),
    (8, 16, 16, 1024, 0, 1024, 9),
    (8, 8, 8, 1024, 0, 1024, 9),
    (8, 256, 256, 256, 256, 256, 1),
    (8, 32, 32, 512, 0, 1536, 1),
]
res = []
for (B, H, W, C1, C2, Cout, taps) in shapes:
    for bn in (64, 128, 256, 512):
        if Cout % (256 if bn == 512 else bn):
            continue
        x1 = torch.randn(B, H, W, C1, device=dev).half()
        x2 = torch.randn(B, H, W, C2, device=dev).half() if C2 else None
        w = (torch.randn(Cout, taps * (C1 + C2), device=dev) * 0.02).half()
        b = torch.randn(Cout, device=dev)
        out = torch.empty(B, H, W, Cout, device=dev, dtype=torch.float16)
        args = (x1, x2, w, b, None, out,
                B, H, W, C1, C2, Cout, taps, bn)
        for _ in range(3):
            _lib.call("pdr_conv_tc", *args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            _lib.call("pdr_conv_tc", *args)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2.0 * B * H * W * Cout * taps * (C1 + C2)
        r = dict(shape=[B, H, W, C1, C2, Cout, taps], bn=bn, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1))
        print(json.dumps(r), flush=True)
        res.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_conv.json", "w"), indent=1)
