"""Golden vectors for the "next" row N2 (paint_invisible_areas_by_neighbors), produced by
executing the REFERENCE's own function through oracle/ref_loader.py.

Run in the build container only:   python tests/golden/make_golden_neighbors.py
Output: tests/golden/neighbors_small.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from pointdreamer_b200 import synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CFG = dict(n_points=500, seed=7, nu=14, nv=12, atlas_res=96, charts=(2, 2))


def inputs(cfg=CFG, unseen_below=-0.12):
    """Scene + an atlas painted everywhere except where the surface lies below a plane (the
    part no camera saw) and a few random holes."""
    sc = synthetic.make_scene(cfg["n_points"], cfg["seed"], cfg["nu"], cfg["nv"], cfg["atlas_res"],
                              charts=cfg["charts"])
    xa = sc["xatlas_dict"]
    R = cfg["atlas_res"]
    rng = np.random.default_rng(3)
    mask = xa["mask"][0, :, :, 0]
    painted = mask & (xa["gb_pos"][0, :, :, 1] > unseen_below) & (rng.random((R, R)) < 0.97)
    atlas = (rng.random((R, R, 3)).astype(np.float32)) * painted[..., None]
    face_id = xa["per_atlas_pixel_face_id"][0]
    ids = np.unique(face_id[~painted])       # demo.py:178-179
    ids = ids[ids > -1]
    return sc, atlas.astype(np.float32), painted, ids


def run_reference(unseen_below):
    un = ref_loader.load("pointdreamer.unproject")
    sc, atlas, painted, ids = inputs(unseen_below=unseen_below)
    xa = sc["xatlas_dict"]
    with ref_loader.quiet():
        out = un.paint_invisible_areas_by_neighbors(
            torch.from_numpy(sc["vertices"]), torch.from_numpy(sc["faces"]),
            torch.from_numpy(xa["uvs"]), torch.from_numpy(xa["mesh_tex_idx"]),
            torch.from_numpy(ids), torch.from_numpy(atlas.copy()),
            torch.from_numpy(painted.copy()), use_atlas=True)
    return out.numpy().astype(np.float32), len(ids)


def main():
    out, n = run_reference(-0.12)
    out2, n2 = run_reference(0.05)   # a larger never-seen region: more colouring rounds
    path = os.path.join(HERE, "neighbors_small.npz")
    np.savez_compressed(path, atlas_out=out, n_to_inpaint=np.int64(n), atlas_out_b=out2,
                        n_to_inpaint_b=np.int64(n2))
    print("wrote", path, out.shape, "faces to inpaint:", n, n2)


if __name__ == "__main__":
    main()
