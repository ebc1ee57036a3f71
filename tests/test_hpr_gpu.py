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


def _unit_ball_cloud(rng, n, scale=0.45):
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return (d * (scale * rng.uniform(0.2, 1.0, (n, 1)) ** (1 / 3))).astype(np.float32)


@pytest.mark.parametrize("n", [4, 5, 17, 33, 64, 65, 200, 1000])
def test_hpr_small_clouds_vs_oracle(cuda, n):
    """Edge sizes of the ordering code: fewer points than one warp / one tile of survivors, no
    survivor besides the extremes, a single tile."""
    from oracle import camera as ocam, hpr as ohpr
    from pointdreamer_b200 import ours_utils
    xyz = _unit_ball_cloud(np.random.default_rng(100 + n), n)
    _, _, eyes, _ = ocam.create_cameras(4, 1.6, 512)
    vis = ours_utils.get_point_validation_by_o3d(torch.from_numpy(xyz).to(cuda), eyes, 100).cpu().numpy()
    ref = ohpr.point_validation_by_o3d(xyz, eyes, 100)
    assert int((vis != ref).sum()) == 0, (n, int(ref.sum()))


def test_hpr_clustered_cloud_vs_oracle(cuda):
    """A cloud whose bounding box is spanned by a few outliers: almost every point falls into the same
    cell of both grids (one extreme block, one huge Morton bucket: the rank-by-index path)."""
    from oracle import camera as ocam, hpr as ohpr
    from pointdreamer_b200 import ours_utils
    rng = np.random.default_rng(7)
    core = _unit_ball_cloud(rng, 6000, scale=0.01) + np.float32([0.1, -0.05, 0.02])
    outliers = _unit_ball_cloud(rng, 12, scale=0.45)
    xyz = np.concatenate([core, outliers]).astype(np.float32)
    xyz = xyz[rng.permutation(xyz.shape[0])]
    _, _, eyes, _ = ocam.create_cameras(3, 1.6, 512)
    vis = ours_utils.get_point_validation_by_o3d(torch.from_numpy(xyz).to(cuda), eyes, 100).cpu().numpy()
    ref = ohpr.point_validation_by_o3d(xyz, eyes, 100)
    print(f"clustered: visible {int(ref.sum())}/{ref.size}")
    assert int((vis != ref).sum()) == 0


def test_hpr_large_noisy_cloud_vs_oracle(cuda):
    """configs[4]-sized input: 100k points with surface noise, 16 views."""
    from oracle import camera as ocam, hpr as ohpr
    from pointdreamer_b200 import ours_utils, synthetic
    xyz, _, _ = synthetic.make_cloud(100000, seed=11)
    xyz = (xyz + np.random.default_rng(3).normal(0, 2e-3, xyz.shape)).astype(np.float32)
    _, _, eyes, _ = ocam.create_cameras(16, 1.6, 512)
    vis = ours_utils.get_point_validation_by_o3d(torch.from_numpy(xyz).to(cuda), eyes, 100).cpu().numpy()
    ref = ohpr.point_validation_by_o3d(xyz, eyes[:4], 100)
    assert int((vis[:4] != ref).sum()) == 0
