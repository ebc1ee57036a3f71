"""Full-size chain parity report (GPU box): the reference's own sampler + UNetModel (baseline/_ref)
next to the CUDA path for the 100-step, 256^2, 552.8M-parameter chain; writes the JSON that
profiles/r02_chain_parity.json is a copy of.

    python tools/chain_parity_report.py [--views 8] [--steps 100] [--out gpurun_out/chain_parity.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from chain_parity import study  # noqa: E402
from pointdreamer_b200.unet import DEFAULT_MODEL_CONFIG  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=8)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "chain_parity.json"))
args = ap.parse_args()
dev = torch.device("cuda:0")
res = study(dev, dict(DEFAULT_MODEL_CONFIG), args.views, args.steps)
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(res, open(args.out, "w"), indent=1)
brief = {k: v for k, v in res.items() if not isinstance(v, (list, dict)) or k.startswith("final")}
brief["drift_ours_every10"] = res["drift_ours"][::10] + res["drift_ours"][-1:]
for k in ("drift_ref16_cudnn_benchmark", "drift_ref32"):
    if k in res:
        brief[k + "_every10"] = res[k][::10] + res[k][-1:]
tf = res["teacher_forced_ours_vs_ref16"]
brief["teacher_forced_ours_max"] = max(tf)
brief["teacher_forced_ours_every10"] = tf[::10]
brief["teacher_forced_ref32"] = res.get("teacher_forced_ref32_vs_ref16")
print(json.dumps(brief, indent=1))
