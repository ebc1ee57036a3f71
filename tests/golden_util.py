"""Loader for the fixtures written by tests/golden/make_golden.py."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(path):
    z = np.load(path)
    out = {}
    for k in z.files:
        if k.startswith("bool:"):
            name = k[5:]
            shape = tuple(z["shape:" + name])
            n = int(np.prod(shape))
            out[name] = np.unpackbits(z[k])[:n].astype(bool).reshape(shape)
        elif k.startswith("shape:"):
            continue
        elif k.startswith("i16:"):
            out[k[4:]] = z[k].astype(np.int64)
        else:
            out[k] = z[k]
    return out


def load_geom_case(name):
    """Returns (cfg, scene, golden) for geometry case `name` ('a', 'b', 'c')."""
    import sys
    sys.path.insert(0, GOLDEN_DIR)
    from make_golden_cases import CASES
    from pointdreamer_b200 import synthetic
    cfg = CASES[name]
    if cfg.get("scene") == "clock":
        from proxy_mesh import clock_scene
        sc = clock_scene(os.path.join(GOLDEN_DIR, "clock.ply"), atlas_res=cfg["atlas_res"])
    else:
        sc = synthetic.make_scene(cfg["n_points"], cfg["seed"], cfg["nu"], cfg["nv"],
                                  cfg["atlas_res"], charts=cfg["charts"])
    return cfg, sc, load_npz(os.path.join(GOLDEN_DIR, f"geom_case_{name}.npz"))
