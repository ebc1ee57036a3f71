// Host launchers of the geometry kernels (geom_project.cu, geom_splat.cu, geom_fill.cu,
// geom_unproject.cu).  Pointers are device pointers; see include/pdr.h for layouts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace pdr {

int project_launch(const float* cams, const float* vertices, int Vm, const float* points, int N,
                   int V, int rescale, double padding, int* ws_minmax, float* pos,
                   float* vertice_uvs, float* uv_centers, float* uv_scales, float* point_uvs,
                   float* point_depths, cudaStream_t stream);

int rasterize_launch(const float* pos, const int* faces, int V, int Vm, int F, int res,
                     int out_res, unsigned long long* ws_keys, float* depth, long long* face_idx,
                     uint8_t* mask_cam, uint8_t* mask_out, cudaStream_t stream);

int mask_half_any_launch(const uint8_t* in, int V, int res_in, uint8_t* out, cudaStream_t stream);

int point_visibility_launch(const float* puv, const float* pdepth, const float* mesh_depths,
                            int V, int N, int cam_res, float offset, int res, uint8_t* vis,
                            long long* pix_cam, long long* pix_res, cudaStream_t stream);

size_t sparse_images_workspace_bytes(int V, int res);
int sparse_images_launch(const long long* point_pixels, const float* colors, const uint8_t* valid,
                         const uint8_t* hard_masks, int V, int N, int res, int point_size,
                         int edge_point_size, double mask_ratio_thresh, void* workspace,
                         float* sparse, float* m0, float* m2, float* scale_factors,
                         cudaStream_t stream);

size_t nearest_fill_workspace_bytes(int B, int H, int W);
int nearest_fill_launch(const float* img, const uint8_t* known, int B, int C, int H, int W,
                        int channels_last, void* workspace, float* out, int* src_index,
                        cudaStream_t stream);

size_t unproject_workspace_bytes(int R, int n_levels);
int unproject_launch(const float* images, int res, const float* cams, int V, int cam_res,
                     const float* base_dirs, const float* gb_pos, const uint8_t* mask,
                     const long long* face_id, int R, const float* f_normals, int F,
                     const float* uv_centers, const float* uv_scales, double padding, int rescale,
                     const float* scale_factors, const float* mesh_depths, const int* kernels_host,
                     int n_levels, int n_kernels_total, int complete_unseen, void* workspace,
                     float* atlas, uint8_t* shrinked_vis, long long* point_view_ids,
                     long long* point_coords, float* points, uint8_t* painted,
                     cudaStream_t stream);

size_t hpr_workspace_bytes(int V, int N);
int hpr_launch(const float* points, int N, int V, const double* frames_dev, double radius,
               void* workspace, uint8_t* vis, cudaStream_t stream);

int mask_count_sync(const uint8_t* mask, size_t n, int* ws_counter, int* out_host,
                    cudaStream_t stream);

}  // namespace pdr
