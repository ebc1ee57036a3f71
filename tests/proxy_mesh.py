"""Test harness alias: the deterministic proxy mesh + atlas for a real point cloud lives in
pointdreamer_b200/synthetic.py (bench.py uses it for BASELINE configs[0] / [2] too)."""
from pointdreamer_b200.synthetic import proxy_scene_from_ply as clock_scene  # noqa: F401
from pointdreamer_b200.synthetic import quad_atlas, voxel_shell  # noqa: F401
