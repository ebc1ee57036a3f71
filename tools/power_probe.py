"""Power / clock draw of the two kernel classes that make up a U-Net forward, each looped alone for a
few seconds on the B200 (nvidia-smi sampled every 100 ms): the dominant tcgen05 conv (256->256 3x3
at 256^2, batch 8) and the GroupNorm+SiLU apply pass over the same tensor.  Backs the energy
argument of DESIGN.md section 4 (why fusing the GroupNorm transform into the power-capped conv is
time-neutral): the conv sits at the 1 kW cap with reduced clocks, the apply pass does not."""
import json
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointdreamer_b200 import _lib

dev = torch.device("cuda:0")
B, H, W, C = 8, 256, 256, 256
x = torch.randn(B, H, W, C, device=dev).half()
w = (torch.randn(C, 9 * C, device=dev) * 0.02).half()
bias = torch.randn(C, device=dev)
out = torch.empty(B, H, W, C, device=dev, dtype=torch.float16)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
lib = _lib.load()
ws = torch.empty(B * 4096 * 2 * C, device=dev)
stats = torch.empty(B * 64, device=dev)


def conv():
    _lib.call("pdr_conv_tc", x, None, w, bias, None, out, B, H, W, C, 0, C, 9, 0)


def gn():
    _lib.call("pdr_group_norm", x, None, B, H, W, C, 0, gamma, beta, None, 0, 0, 1, 0, ws, stats, out)


def sample(fn, seconds):
    rows = []
    p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw",
                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    th = threading.Thread(target=lambda: [rows.append(l.strip()) for l in p.stdout], daemon=True)
    th.start()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    p.terminate()
    vals = [r.split(",") for r in rows if "," in r]
    vals = vals[len(vals) // 3:]  # steady state
    mhz = sorted(float(v[0]) for v in vals)
    pw = sorted(float(v[1]) for v in vals)
    return dict(launches=n, us_per_launch=e0.elapsed_time(e1) * 1e3 / n, sm_mhz_median=mhz[len(mhz) // 2],
                power_w_median=pw[len(pw) // 2], power_w_max=pw[-1])


res = dict(conv_256x256_256to256_b8=sample(conv, 4.0), gn_apply_same_tensor=sample(gn, 4.0))
c, g = res["conv_256x256_256to256_b8"], res["gn_apply_same_tensor"]
c["tflops"] = 2.0 * B * H * W * C * 9 * C / c["us_per_launch"] / 1e6
g["gbs"] = 2.0 * B * H * W * C * 2 / g["us_per_launch"] / 1e3
c["joule_per_launch"] = c["power_w_median"] * c["us_per_launch"] * 1e-6
g["joule_per_launch"] = g["power_w_median"] * g["us_per_launch"] * 1e-6
print(json.dumps(res, indent=1))
