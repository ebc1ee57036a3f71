"""Full-size ADM U-Net (imagenet_256.yml) on the native engine: forward and DDNM chain timing."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointdreamer_b200.unet import UNetEngine, random_state_dict, DEFAULT_MODEL_CONFIG
from pointdreamer_b200.ddnm_inpainting import Inpainter
from pointdreamer_b200 import _lib

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
t0 = time.time()
sd = random_state_dict(DEFAULT_MODEL_CONFIG, 1234, dev)
torch.cuda.synchronize()
print("weights", time.time() - t0, "s", torch.cuda.memory_allocated() / 1e9, "GB")
inp = Inpainter(dev, state_dict=sd)
del sd
eng = inp.model
x = torch.randn(B, 3, 256, 256, device=dev)
t = torch.full((B,), 500.0, device=dev)
eng.plan(B)
print("workspace GB", eng.workspace_bytes / 1e9, "alloc GB", torch.cuda.memory_allocated() / 1e9)
l0 = _lib.launch_count()
y = eng(x, t)
torch.cuda.synchronize()
print("launches per forward", _lib.launch_count() - l0, "out std", y.std().item(), "finite", torch.isfinite(y).all().item())
if os.environ.get("PDR_QUICK"):
    sys.exit(0)
for _ in range(2):
    eng(x, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    eng(x, t, n_out=3)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps(dict(B=B, fwd_ms=ms, tflops=2.2397 * B / ms * 1e3 / 1e3)))
sparse = torch.rand(B, 3, 256, 256, device=dev)
mask = (torch.rand(B, 256, 256, device=dev) < 0.3).float()
sparse = sparse * mask[:, None]
torch.cuda.synchronize()
t0 = time.time()
out = inp.inpaint_batch(sparse, mask)
torch.cuda.synchronize()
dt = time.time() - t0
print(json.dumps(dict(chain_s=dt, per_step_ms=dt * 10, known_err=float((out - sparse)[mask[:, None].expand_as(out) > 0].abs().max()))))
