import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from golden_util import load_geom_case
import test_geometry_gpu as T
dev = torch.device("cuda:0")
for name in ["a", "b", "c"]:
    cfg, sc, g = load_geom_case(name)
    got = T._run_pipeline(cfg, sc, dev, inpainted_override=g["inpainted_nearest"])
    print("==== case", name)
    for k in T.EXACT + ["inpainted_nearest", "point_view_ids", "atlas_img", "atlas_painted_mask", "atlas_dilated"]:
        if k in g and k in got:
            a, b = got[k], g[k]
            print(f"{k:28s} mismatches {int((a != b).sum())} / {a.size}")
    s, gs = got["sparse_imgs"], g["sparse_imgs"]
    for v in range(cfg["view_num"]):
        bad = (s[v] != gs[v]).any(0)
        ys, xs = np.nonzero(bad)
        print("view", v, "bad px", bad.sum(), "nonzero got", (s[v] != 0).any(0).sum(), "nonzero gold", (gs[v] != 0).any(0).sum())
        for y, x in list(zip(ys, xs))[:4]:
            print("   px", y, x, "got", s[v, :, y, x], "gold", gs[v, :, y, x], "m2 got/gold", got["hard_mask2s"][v, 0, y, x], g["hard_mask2s"][v, 0, y, x])
