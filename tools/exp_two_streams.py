"""Experiment (result: no gain, 0.998x; the 4-engine variant HUNG the GPU box once - always run under
`timeout`, and do not pass an argument > 2): do two half-batch DDNM samplers on two CUDA streams (tensor-bound convs of one
overlapping the HBM-bound GroupNorm / attention of the other) beat one full-batch sampler?"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointdreamer_b200.ddnm_inpainting import Inpainter
from pointdreamer_b200.unet import random_state_dict, DEFAULT_MODEL_CONFIG

dev = torch.device("cuda:0")
V = 8
sd = random_state_dict(DEFAULT_MODEL_CONFIG, 1234, dev)
full = Inpainter(dev, state_dict=sd)
g = torch.Generator(device=dev).manual_seed(0)
mask = (torch.rand(V, 256, 256, device=dev, generator=g) < 0.3).float()
sparse = torch.rand(V, 3, 256, 256, device=dev, generator=g) * mask[:, None]

def timed(fn, n=2):
    fn(); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.time() - t0) / n, out

t_full, out_full = timed(lambda: full.inpaint_batch(sparse, mask, chain0=0))
print(json.dumps({"one_stream_batch8_s": t_full}), flush=True)

nsplit = int(sys.argv[1]) if len(sys.argv) > 1 else 2
parts = [Inpainter(dev, state_dict=sd) for _ in range(nsplit)]
streams = [torch.cuda.Stream() for _ in range(nsplit)]
per = V // nsplit

def split_run():
    outs = []
    cur = torch.cuda.current_stream()
    for i, (p, s) in enumerate(zip(parts, streams)):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs.append(p.inpaint_batch(sparse[i * per:(i + 1) * per], mask[i * per:(i + 1) * per],
                                        chain0=i * per))
    for s in streams:
        cur.wait_stream(s)
    return torch.cat(outs)

t_split, out_split = timed(split_run)
print(json.dumps({"streams": nsplit, "split_s": t_split, "speedup": t_full / t_split,
                  "identical": bool(torch.equal(out_full, out_split))}))
