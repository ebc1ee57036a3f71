"""Time optimize_color (N1) at the reference's size: 8 views, 1024^2 renders, 1024^2 atlas."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pointdreamer_b200 import camera, ours_utils as ou, synthetic, _lib

dev = torch.device("cuda:0")
V, R = 8, 1024
sc = synthetic.make_scene(30000, seed=0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cams, base_dirs, eyes, ups = camera.create_cameras(V, 1.6, 512, device=dev)
vertices, faces, xyz = t(sc["vertices"]), t(sc["faces"]), t(sc["xyz"])
xa = {k: t(v) for k, v in sc["xatlas_dict"].items()}
(_, _, _, _, uvc, uvs_, padding, _, _) = ou.get_rendered_hard_mask_and_face_idx_batch(cams, vertices, faces, xyz)
g = torch.Generator(device="cpu").manual_seed(0)
imgs = torch.rand(V, 3, 256, 256, generator=g).to(dev)
atlas = torch.rand(3, R, R, generator=g).to(dev)
vis = (torch.rand(V, R, R, generator=g) < 0.9).to(dev)
isf = torch.ones(V, device=dev)
for rep in range(3):
    torch.cuda.synchronize(); l0 = _lib.launch_count(); t0 = time.time()
    a, im = ou.optimize_color(atlas, imgs, vertices, faces, xa["uvs"], xa["mesh_tex_idx"], cams, None, None, None,
                              uvc, uvs_, padding, isf, None, shrinked_per_view_per_pixel_visibility=vis,
                              return_images=False)
    torch.cuda.synchronize()
    print(f"optimize_color 8x1024^2, R=1024, 100 iterations: {(time.time()-t0)*1e3:.1f} ms, {_lib.launch_count()-l0} launches")
