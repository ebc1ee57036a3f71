"""torchrun --nproc-per-node G tools/verify_view_sharding.py
One shape, its V diffusion chains sharded by view over G GPUs (pointdreamer_b200.dist
.inpaint_views_sharded: chain v keeps its slot of the noise stream), ONE NCCL all-gather, then every
rank unprojects.  Rank 0 also runs all V chains alone and checks that the sharded result is
bit-identical (SURVEY §8e "sharded output equals the single-GPU output bit-for-bit")."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pointdreamer_b200 import demo, synthetic
from pointdreamer_b200 import dist as pdist
from pointdreamer_b200 import ours_utils as ou
from pointdreamer_b200.ddnm_inpainting import DEFAULT_DDNM_CONFIG, Inpainter

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
cfg = dict(demo.DEFAULT_CONFIG, complete_unseen_by="unproject", optimize_from=None)
V, res, cam_res = cfg["view_num"], cfg["res"], cfg["cam_res"]
sc = synthetic.make_scene(30000, seed=0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
xyz, rgb, vertices, faces = t(sc["xyz"]), t(sc["rgb"]), t(sc["vertices"]), t(sc["faces"])
cam = demo.prepare_cameras(cfg, dev)
inp = Inpainter(dev, ddnm_config=dict(DEFAULT_DDNM_CONFIG, T_sampling=T), seed=42, offset=0, allow_random_weights=True)
(hm, _, depths, _, uvc, uvs, pad, puv, pdep) = ou.get_rendered_hard_mask_and_face_idx_batch(
    cam["cams"], vertices, faces, xyz)
hm = ou.resize_hard_masks(hm, res)
pv, _ = ou.get_point_validation_by_depth(cam_res, puv, pdep, depths, offset=0.0001)
pv = pv | ou.get_point_validation_by_o3d(xyz, cam["eye_positions"], cfg["hidden_point_removal_radius"])
sparse, m0, m2, scales = ou.get_sparse_images(ou.get_point_pixels(puv, res), rgb, pv, hm, None, V, res, 1, 1, 0.82)
torch.cuda.synchronize()
dist.barrier()
t0 = time.time()
sharded = pdist.inpaint_views_sharded(inp, sparse, m2[:, 0])
torch.cuda.synchronize()
dt = time.time() - t0
ok = None
if rank == 0:
    alone = inp.inpaint_batch(sparse, m2[:, 0], chain0=0)
    ok = bool(torch.equal(alone, sharded))
same = torch.tensor([float(sharded.double().sum().item())], device=dev, dtype=torch.float64)
allsum = [torch.zeros_like(same) for _ in range(world)]
dist.all_gather(allsum, same)
if rank == 0:
    print(json.dumps({"gpus": world, "views": V, "T_sampling": T, "sharded_equals_single_gpu_bitwise": ok,
                      "all_ranks_hold_the_same_views": bool(all(float(a) == float(allsum[0]) for a in allsum)),
                      "sharded_inpaint_s": dt}))
dist.destroy_process_group()
