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

size_t rasterize_workspace_bytes(int V, int F, int res);
int rasterize_launch(const float* pos, const int* faces, int V, int Vm, int F, int res,
                     int out_res, void* workspace, float* depth, long long* face_idx,
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

int interpolate_launch(const float* pos, const int* faces, const long long* face_idx,
                       const float* attr, const int* attr_faces, int V, int Vm, int res, int C,
                       int flip_y, float* out, uint8_t* mask_out, cudaStream_t stream);
int face_normals_launch(const float* verts, const int* faces, int F, float* out,
                        cudaStream_t stream);
int project_fixed_launch(const float* cams, const float* vertices, int Vm, int V, double padding,
                         const float* centers, const float* scales, const float* inpaint_scales,
                         float* pos, cudaStream_t stream);

int texopt_prepare_launch(const float* uv_map, const uint8_t* mask, const uint8_t* vis,
                          const float* inpainted, int r0, int V, int res, int R, uint8_t* active,
                          float* target, long long* keys, cudaStream_t stream);
int texopt_build_launch(const long long* sorted_keys, long long n_valid, const float* uv_map, int R,
                        unsigned int* entry_pix, double* entry_w, uint8_t* head,
                        cudaStream_t stream);
int texopt_forward_launch(const float* atlas, const float* uv_map, const uint8_t* active,
                          const float* target, int V, int res, int R, signed char* signs,
                          double* images, cudaStream_t stream);
int texopt_step_launch(float* atlas, float* m, float* v, const long long* sorted_keys,
                       const long long* seg_start, long long n_seg, const unsigned int* entry_pix,
                       const double* entry_w, const signed char* signs, int V, int res, int R,
                       float lerp_w, float beta2, float one_m_beta2, float bc2_sqrt, float eps,
                       float neg_step_size, cudaStream_t stream);

int vertex_colors_launch(const int* faces, const int* face_uv_idx, int F, const float* uvs, int Vn,
                         const float* atlas, const uint8_t* mask, int R, int* ws_uv_idx,
                         long long* pix, float* colors, float* count, uint8_t* has_color,
                         cudaStream_t stream);
int laplacian_round_launch(const int* rowptr, const int* colidx, int Vn, const uint8_t* fixed,
                           const float* colors_in, const float* count_in, float* colors_out,
                           float* count_out, int* colored_total, cudaStream_t stream);
int scatter_vertex_colors_launch(const long long* pix, const float* colors, int Vn, int R,
                                 int* ws_winner, float* atlas, uint8_t* mask,
                                 cudaStream_t stream);

int atlas_to_u8_launch(const float* atlas, const uint8_t* mask, int R, uint8_t* rgb, uint8_t* rgba,
                       cudaStream_t stream);

int mask_count_sync(const uint8_t* mask, size_t n, int* ws_counter, int* out_host,
                    cudaStream_t stream);

}  // namespace pdr
