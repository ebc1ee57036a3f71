"""Configurations of the golden geometry cases (shared by the generator and the tests)."""
CASES = {
    # name: dict(config)
    "a": dict(n_points=3000, seed=1, nu=24, nv=24, atlas_res=256, charts=(2, 2), view_num=3,
              res=64, cam_res=128, point_size=1, edge_point_size=1, crop_img=True,
              crop_padding=0.05, mask_ratio_thresh=0.82, edge_dilate_kernels=[5],
              complete_unseen_by_projection=False),
    # sparse cloud -> mask_ratio > thresh -> shrink-and-pad branch; NBF off; fallback to all views
    "b": dict(n_points=250, seed=2, nu=16, nv=16, atlas_res=256, charts=(2, 2), view_num=2,
              res=64, cam_res=128, point_size=1, edge_point_size=1, crop_img=True,
              crop_padding=0.05, mask_ratio_thresh=0.82, edge_dilate_kernels=[0],
              complete_unseen_by_projection=True),
    # window splats, multi-level NBF, no crop
    "c": dict(n_points=2000, seed=3, nu=20, nv=20, atlas_res=256, charts=(3, 3), view_num=4,
              res=64, cam_res=128, point_size=2, edge_point_size=2, crop_img=False,
              crop_padding=0.05, mask_ratio_thresh=0.82, edge_dilate_kernels=[7, 3],
              complete_unseen_by_projection=False),
    # BASELINE.json configs[0]: dataset/demo_data/clock.ply, view_num=2, texture_gen_method='nearest'
    # (configs/nearest.yaml), HPR on; proxy voxel-shell mesh + quad atlas (tests/proxy_mesh.py);
    # reduced raster/atlas resolutions keep the fixture small
    "clock": dict(scene="clock", view_num=2, res=128, cam_res=256, atlas_res=256, point_size=1,
                  edge_point_size=1, crop_img=True, crop_padding=0.05, mask_ratio_thresh=0.82,
                  edge_dilate_kernels=[21], complete_unseen_by_projection=True, use_o3d=True),
}
