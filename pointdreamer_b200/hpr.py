"""Hidden point removal (ours_utils.py:204-225, open3d `hidden_point_removal`).

NOT BUILT YET: the reference runs Katz' HPR as a float64 Qhull convex hull on
the CPU; the B200 version needs an exact GPU convex-hull membership kernel.  There is
deliberately no CPU fallback — asking for it raises, and `point_validation_by_o3d: False`
selects the depth-only visibility the reference ORs it with (demo.py:107-112).
"""


def hidden_point_removal(points, eye_positions, radius):
    raise NotImplementedError(
        "point_validation_by_o3d=True needs the GPU hidden-point-removal kernel, which is not "
        "built yet; run with point_validation_by_o3d=False (depth-only visibility). "
        "No CPU fallback is provided on purpose.")
