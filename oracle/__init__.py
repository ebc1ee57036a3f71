"""CPU oracle for the project -> DDNM-inpaint -> unproject path.

TEST INFRASTRUCTURE ONLY.  This package is a plain numpy / CPU-torch restatement of the
reference algorithm (YuQiao0303/PointDreamer @ 6fa8552).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import it, and only as the checker (or as the timed CPU baseline) — never as part of
the product path in `pointdreamer_b200/`.

Parity status ("pinned" = checked against outputs of the reference's own source executed in
the build container through `oracle/ref_loader.py`; fixtures + generator in `tests/golden/`):

  pinned   : ours_utils.get_rendered_hard_mask_and_face_idx_batch (crop/rescale arithmetic),
             get_point_validation_by_depth, paint_pixels, get_forground_inner_edge_mask,
             get_one_sparse_img, get_sparse_images, naive_inpainting('nearest') away from
             ties, unproject + NBF (Scharr / dilate), dilate_atlas, the ADM U-Net forward,
             the DDNM schedule and step arithmetic; the "next" rows optimize_color,
             xatlas_uvmap_w_face_id (after xatlas.parametrize), subdivide_with_uv,
             compute_vertex_only_uv_mask, paint_invisible_areas_by_neighbors (optimize.py,
             neighbors.py; fixtures optimize_small.npz / neighbors_small.npz).
  UNPINNED : third-party arithmetic that is not vendored in the reference and not
             installable here — kaolin Camera.transform, nvdiffrast.rasterize fill rule,
             kaolin sided_distance tie rule, open3d hidden_point_removal, scipy cKDTree
             tie rule, the pretrained ADM weights; nvdiffrast.interpolate, kaolin
             texture_mapping / uniform_laplacian / face_normals, trimesh unique_rows /
             faces_to_edges (restated from their published algorithms), xatlas.parametrize
             (its result is an input).  The oracle fixes one canonical rule for
             each (documented next to the code) and the goldens were produced with shims
             that implement that same rule.

fp32 conventions: every floating-point expression on the geometry side is written as a
sequence of single IEEE-754 binary32 operations (numpy float32 elementwise ops, no FMA, no
matmul) in exactly the order the CUDA kernels use, so integer results derived from them are
bit-exact.
"""
