"""GPU hidden point removal vs the Qhull oracle and the golden produced through the reference's
get_point_validation_by_o3d (with the open3d shim)."""
import numpy as np
import pytest
import torch

from golden_util import load_geom_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_hpr_vs_reference_golden(cuda, name):
    from pointdreamer_b200 import ours_utils
    cfg, sc, g = load_geom_case(name)
    pts = torch.from_numpy(sc["xyz"]).to(cuda)
    vis = ours_utils.get_point_validation_by_o3d(pts, g["eye_positions"], 100).cpu().numpy()
    ref = g["point_validation_o3d"]
    mism = int((vis != ref).sum())
    print(f"case {name}: visible {int(ref.sum())}/{ref.size}, mismatches {mism}")
    assert mism == 0


def test_hpr_full_size_vs_oracle(cuda):
    from oracle import camera as ocam, hpr as ohpr
    from pointdreamer_b200 import ours_utils, synthetic
    xyz, _, _ = synthetic.make_cloud(30000, seed=5)
    _, _, eyes, _ = ocam.create_cameras(8, 1.6, 512)
    pts = torch.from_numpy(xyz).to(cuda)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ours_utils.get_point_validation_by_o3d(pts, eyes, 100)
    e0.record()
    vis = ours_utils.get_point_validation_by_o3d(pts, eyes, 100)
    e1.record()
    torch.cuda.synchronize()
    vis = vis.cpu().numpy()
    ref = ohpr.point_validation_by_o3d(xyz, eyes, 100)
    mism = int((vis != ref).sum())
    print(f"30k x 8 views: visible {int(ref.sum())}, mismatches {mism}, {e0.elapsed_time(e1):.2f} ms")
    assert mism == 0
