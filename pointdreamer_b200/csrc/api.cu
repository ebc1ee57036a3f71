// extern "C" surface of libpdr.so (declared in include/pdr.h).
#include "pdr.h"
#include "common.cuh"
#include "conv_tc.h"
#include "geom.h"
#include "unet_engine.h"
#include "unet_ops.h"

using namespace pdr;

extern "C" {

int pdr_version(void) { return PDR_VERSION; }
const char* pdr_last_error(void) { return get_error(); }
unsigned long long pdr_launch_count(void) { return g_launch_count; }

int pdr_conv_tc(const void* x1, const void* x2, const void* w, const float* bias,
                const void* residual, void* out, int B, int H, int W, int C1, int C2, int Cout,
                int taps, int bn, void* stream) {
  PDR_CHECK_ARG(x1 && w && out, "pdr_conv_tc: null pointer");
  PDR_CHECK_ARG(B > 0 && H > 0 && W > 0, "pdr_conv_tc: empty shape");
  PDR_CHECK_ARG(Cout % 64 == 0, "pdr_conv_tc: Cout (%d) must be a multiple of 64", Cout);
  if (bn == 0) bn = conv_tc_pick_bn(B, H, W, Cout, taps);
  ConvTensorMap ma1, ma2, mw;
  const int halo = conv_tc_halo_ok(H, W, taps) ? 1 : 0;
  PDR_TRY(conv_tc_make_act_map(&ma1, x1, B, H, W, C1, halo));
  if (C2 > 0) {
    PDR_CHECK_ARG(x2 != nullptr, "pdr_conv_tc: x2 is null but C2=%d", C2);
    PDR_TRY(conv_tc_make_act_map(&ma2, x2, B, H, W, C2, halo));
  }
  PDR_TRY(conv_tc_make_weight_map(&mw, w, Cout, taps * (C1 + C2), bn == 512 ? 128 : bn));
  ConvTensorMap mo;
  if (!halo) PDR_TRY(conv_tc_make_act_map(&mo, out, B, H, W, Cout, 0));
  return conv_tc_launch(&ma1, C2 > 0 ? &ma2 : nullptr, &mw, bn, B, H, W, C1, C2, Cout, taps, bias,
                        (const __half*)residual, (__half*)out, nullptr, (cudaStream_t)stream, 0.f,
                        nullptr, nullptr, 0, 0, 1, nullptr, halo, nullptr, 0, halo ? nullptr : &mo);
}

int pdr_conv_tc_skip(const void* x, const void* w, const float* bias, const void* s1,
                     const void* s2, void* out, int B, int H, int W, int C, int S1, int S2,
                     int Cout, int bn, void* stream) {
  PDR_CHECK_ARG(x && w && s1 && out, "pdr_conv_tc_skip: null pointer");
  PDR_CHECK_ARG(B > 0 && H > 0 && W > 0 && Cout % 64 == 0, "pdr_conv_tc_skip: bad shape");
  if (bn == 0) bn = conv_tc_pick_bn(B, H, W, Cout, 9);
  ConvTensorMap ma, ms1, ms2, mw;
  const int halo = conv_tc_halo_ok(H, W, 9) ? 1 : 0;
  PDR_TRY(conv_tc_make_act_map(&ma, x, B, H, W, C, halo));
  PDR_TRY(conv_tc_make_act_map(&ms1, s1, B, H, W, S1, halo));
  if (S2 > 0) {
    PDR_CHECK_ARG(s2 != nullptr, "pdr_conv_tc_skip: s2 is null but S2=%d", S2);
    PDR_TRY(conv_tc_make_act_map(&ms2, s2, B, H, W, S2, halo));
  }
  PDR_TRY(conv_tc_make_weight_map(&mw, w, Cout, 9 * C + S1 + S2, bn == 512 ? 128 : bn));
  return conv_tc_launch(&ma, nullptr, &mw, bn, B, H, W, C, 0, Cout, 9, bias, nullptr, (__half*)out,
                        nullptr, (cudaStream_t)stream, 0.f, &ms1, S2 > 0 ? &ms2 : nullptr, S1, S2, 1,
                        nullptr, halo);
}

int pdr_linear(const float* in, const float* W, const float* bias, int B, int K, int N,
               int mode_in, float* out, void* out_fp16, void* stream) {
  PDR_CHECK_ARG(in && W && (out || out_fp16), "pdr_linear: null pointer");
  return linear_launch(in, W, bias, B, K, N, mode_in, out, (__half*)out_fp16,
                       (cudaStream_t)stream);
}
int pdr_stem_conv(const float* x, const void* w, const float* bias, int B, int H, int W, int C,
                  void* out, void* stream) {
  PDR_CHECK_ARG(x && w && bias && out, "pdr_stem_conv: null pointer");
  return stem_conv_launch(x, (const __half*)w, bias, B, H, W, C, (__half*)out,
                          (cudaStream_t)stream);
}
int pdr_group_norm(const void* x1, const void* x2, int B, int H, int W, int C1, int C2,
                   const float* gamma, const float* beta, const void* film, int film_stride,
                   int film_off, int silu, int resample, float* ws, float* stats, void* out,
                   void* stream) {
  PDR_CHECK_ARG(x1 && gamma && beta && ws && stats && out, "pdr_group_norm: null pointer");
  PDR_TRY(gn_stats_launch((const __half*)x1, (const __half*)x2, B, H * W, C1, C2, ws, stats,
                          (cudaStream_t)stream));
  return gn_apply_launch((const __half*)x1, (const __half*)x2, B, H, W, C1, C2, stats, nullptr,
                         nullptr, gamma, beta,
                         (const __half*)film, film_stride, film_off, silu, resample, (__half*)out,
                         (cudaStream_t)stream);
}
int pdr_resample(const void* x, int B, int H, int W, int C, int mode, void* out, void* stream) {
  PDR_CHECK_ARG(x && out, "pdr_resample: null pointer");
  return resample_launch((const __half*)x, B, H, W, C, mode, (__half*)out, (cudaStream_t)stream);
}
int pdr_attention(const void* qkv, int B, int T, int heads, void* out, void* stream) {
  PDR_CHECK_ARG(qkv && out, "pdr_attention: null pointer");
  return attention_launch((const __half*)qkv, B, T, heads, 0, (__half*)out, (cudaStream_t)stream);
}
int pdr_attention_prescaled(const void* qkv, int B, int T, int heads, void* out, void* stream) {
  PDR_CHECK_ARG(qkv && out, "pdr_attention_prescaled: null pointer");
  return attention_launch((const __half*)qkv, B, T, heads, 1, (__half*)out, (cudaStream_t)stream);
}
int pdr_unet_head(const void* h, const float* gamma, const float* beta, const float* w,
                  const float* bias, int B, int H, int W, int C, int n_out, float* ws,
                  float* stats, float* out, void* stream) {
  PDR_CHECK_ARG(h && gamma && beta && w && bias && ws && stats && out, "pdr_unet_head: null");
  PDR_TRY(gn_stats_launch((const __half*)h, nullptr, B, H * W, C, 0, ws, stats,
                          (cudaStream_t)stream));
  return head_launch((const __half*)h, stats, gamma, beta, w, bias, B, H, W, C, n_out, out, n_out,
                     (cudaStream_t)stream);
}

int pdr_unet_create(const PdrUnetConfig* cfg, void** handle) { return unet_create(cfg, handle); }
int pdr_unet_destroy(void* handle) { return unet_destroy(handle); }
int pdr_unet_set_param(void* handle, const char* name, const void* ptr, size_t bytes) {
  return unet_set_param(handle, name, ptr, bytes);
}
int pdr_unet_workspace_bytes(void* handle, int B, size_t* bytes) {
  return unet_workspace_bytes(handle, B, bytes);
}
int pdr_unet_plan(void* handle, int B, void* workspace, size_t bytes) {
  return unet_plan(handle, B, workspace, bytes);
}
int pdr_unet_forward(void* handle, const float* x, const float* t, float* out, int n_out,
                     void* stream) {
  return unet_forward_serialized(handle, x, t, out, n_out, (cudaStream_t)stream);
}

int pdr_unet_profile_begin(void* handle, int every, int max_forwards) {
  return unet_profile_begin(handle, every, max_forwards);
}
int pdr_unet_profile_end(void* handle, double* ms, double* flops, long long* launches,
                         long long* forwards) {
  return unet_profile_end(handle, ms, flops, launches, forwards);
}

int pdr_randn_like_torch(float* out, long long numel, unsigned long long seed,
                         unsigned long long offset, void* stream) {
  PDR_CHECK_ARG(out && numel > 0, "pdr_randn_like_torch: bad argument");
  return randn_like_torch_launch(out, numel, seed, offset, (cudaStream_t)stream);
}
unsigned long long pdr_randn_offset_increment(long long numel) {
  long long T;
  unsigned long long inc;
  philox_launch_geometry(numel, &T, &inc);
  return inc;
}
int pdr_ddnm_sample(void* unet, const float* sparse, const float* mask, int V, int steps,
                    const float* coef_host, const float* t_dev, unsigned long long seed,
                    unsigned long long offset_base, unsigned long long draws_per_chain,
                    int chain0, float* x, float* y, float* et, float* out, void* stream) {
  return ddnm_sample(unet, sparse, mask, V, steps, coef_host, t_dev, seed, offset_base,
                     draws_per_chain, chain0, x, y, et, out, (cudaStream_t)stream);
}
int pdr_ddnm_prepare(const float* sparse, const float* mask, int V, int S,
                     unsigned long long seed, unsigned long long offset_base,
                     unsigned long long draws_per_chain, int chain0, float* y, float* x,
                     void* stream) {
  PDR_CHECK_ARG(sparse && mask && y && x, "pdr_ddnm_prepare: null pointer");
  return ddnm_prepare_launch(sparse, mask, V, 3, S, S, seed, offset_base, draws_per_chain, chain0,
                             y, x, (cudaStream_t)stream);
}
int pdr_ddnm_step(float* x, const float* et, int et_channels, const float* y, const float* mask,
                  int V, int S, const float* c, unsigned long long seed,
                  unsigned long long offset_base, unsigned long long draws_per_chain, int chain0,
                  int draw_index, void* stream) {
  PDR_CHECK_ARG(x && et && y && mask && c, "pdr_ddnm_step: null pointer");
  DdnmStepCoef k;
  k.sqrt_1m_at = c[0], k.sqrt_at = c[1], k.sqrt_at_next = c[2], k.gamma_t = c[3];
  k.c1 = c[4], k.c2 = c[5], k.lambda_t = c[6];
  return ddnm_step_launch(x, et, et_channels, y, mask, V, 3, S, S, k, seed, offset_base,
                          draws_per_chain, chain0, draw_index, (cudaStream_t)stream);
}
int pdr_ddnm_final(const float* x, long long n, float* out, void* stream) {
  PDR_CHECK_ARG(x && out, "pdr_ddnm_final: null pointer");
  return ddnm_final_launch(x, n, out, (cudaStream_t)stream);
}

int pdr_project(const float* cam_params, const float* vertices, int Vm, const float* points,
                int N, int V, int rescale, double padding, int* ws_minmax, float* pos,
                float* vertice_uvs, float* uv_centers, float* uv_scales, float* point_uvs,
                float* point_depths, void* stream) {
  PDR_CHECK_ARG(cam_params && vertices && points && ws_minmax && pos && vertice_uvs &&
                    uv_centers && uv_scales && point_uvs && point_depths,
                "pdr_project: null pointer");
  return project_launch(cam_params, vertices, Vm, points, N, V, rescale, padding, ws_minmax, pos,
                        vertice_uvs, uv_centers, uv_scales, point_uvs, point_depths,
                        (cudaStream_t)stream);
}

size_t pdr_rasterize_workspace_bytes(int V, int F, int res) {
  return rasterize_workspace_bytes(V, F, res);
}
int pdr_rasterize(const float* pos, const int* faces, int V, int Vm, int F, int res, int out_res,
                  void* workspace, float* depth, long long* face_idx, uint8_t* mask_cam,
                  uint8_t* mask_out, void* stream) {
  PDR_CHECK_ARG(pos && faces && workspace && depth && face_idx && mask_cam && mask_out,
                "pdr_rasterize: null pointer");
  return rasterize_launch(pos, faces, V, Vm, F, res, out_res, workspace, depth, face_idx,
                          mask_cam, mask_out, (cudaStream_t)stream);
}

int pdr_mask_half_any(const uint8_t* mask_in, int V, int res_in, uint8_t* mask_out, void* stream) {
  PDR_CHECK_ARG(mask_in && mask_out, "pdr_mask_half_any: null pointer");
  return mask_half_any_launch(mask_in, V, res_in, mask_out, (cudaStream_t)stream);
}

int pdr_point_visibility(const float* point_uvs, const float* point_depths,
                         const float* mesh_depths, int V, int N, int cam_res, float offset,
                         int res, uint8_t* vis, long long* pix_cam, long long* pix_res,
                         void* stream) {
  PDR_CHECK_ARG(point_uvs != nullptr, "pdr_point_visibility: null pointer");
  return point_visibility_launch(point_uvs, point_depths, mesh_depths, V, N, cam_res, offset, res,
                                 vis, pix_cam, pix_res, (cudaStream_t)stream);
}

size_t pdr_sparse_images_workspace_bytes(int V, int res) {
  return sparse_images_workspace_bytes(V, res);
}
int pdr_sparse_images(const long long* point_pixels, const float* colors, const uint8_t* valid,
                      const uint8_t* hard_masks, int V, int N, int res, int point_size,
                      int edge_point_size, double mask_ratio_thresh, void* workspace,
                      float* sparse, float* hard_mask0, float* hard_mask2, float* scale_factors,
                      void* stream) {
  PDR_CHECK_ARG(point_pixels && colors && valid && hard_masks && workspace && sparse &&
                    hard_mask0 && hard_mask2 && scale_factors,
                "pdr_sparse_images: null pointer");
  return sparse_images_launch(point_pixels, colors, valid, hard_masks, V, N, res, point_size,
                              edge_point_size, mask_ratio_thresh, workspace, sparse, hard_mask0,
                              hard_mask2, scale_factors, (cudaStream_t)stream);
}

size_t pdr_nearest_fill_workspace_bytes(int B, int H, int W) {
  return nearest_fill_workspace_bytes(B, H, W);
}
int pdr_nearest_fill(const float* img, const uint8_t* known, int B, int C, int H, int W,
                     int channels_last, void* workspace, float* out, int* src_index,
                     void* stream) {
  PDR_CHECK_ARG(img && known && workspace && out, "pdr_nearest_fill: null pointer");
  return nearest_fill_launch(img, known, B, C, H, W, channels_last, workspace, out, src_index,
                             (cudaStream_t)stream);
}

size_t pdr_unproject_workspace_bytes(int R, int n_levels) {
  return unproject_workspace_bytes(R, n_levels);
}
int pdr_unproject(const float* images, int res, const float* cam_params, int V, int cam_res,
                  const float* base_dirs, const float* gb_pos, const uint8_t* mask,
                  const long long* face_id, int R, const float* f_normals, int F,
                  const float* uv_centers, const float* uv_scales, double padding, int rescale,
                  const float* scale_factors, const float* mesh_depths, const int* kernels_host,
                  int n_levels, int complete_unseen, void* workspace, float* atlas,
                  uint8_t* shrinked_vis, long long* point_view_ids, long long* point_coords,
                  float* points, uint8_t* painted, void* stream) {
  PDR_CHECK_ARG(images && cam_params && base_dirs && gb_pos && mask && face_id && f_normals &&
                    mesh_depths && kernels_host && workspace && atlas && painted,
                "pdr_unproject: null pointer");
  PDR_CHECK_ARG(!rescale || (uv_centers && uv_scales && scale_factors),
                "pdr_unproject: rescale requested without uv_centers/uv_scales/scale_factors");
  return unproject_launch(images, res, cam_params, V, cam_res, base_dirs, gb_pos, mask, face_id, R,
                          f_normals, F, uv_centers, uv_scales, padding, rescale, scale_factors,
                          mesh_depths, kernels_host, n_levels, n_levels, complete_unseen,
                          workspace, atlas, shrinked_vis, point_view_ids, point_coords, points,
                          painted, (cudaStream_t)stream);
}
size_t pdr_hidden_point_removal_workspace_bytes(int V, int N) { return hpr_workspace_bytes(V, N); }
int pdr_hidden_point_removal(const float* points, int N, int V, const double* frames,
                             double radius, void* workspace, uint8_t* vis, void* stream) {
  PDR_CHECK_ARG(points && frames && workspace && vis, "pdr_hidden_point_removal: null pointer");
  PDR_CHECK_ARG(radius > 0.0, "pdr_hidden_point_removal: radius must be positive");
  return hpr_launch(points, N, V, frames, radius, workspace, vis, (cudaStream_t)stream);
}
int pdr_mask_count(const uint8_t* mask, size_t n, int* ws_counter, int* out_host, void* stream) {
  PDR_CHECK_ARG(mask && ws_counter && out_host, "pdr_mask_count: null pointer");
  return mask_count_sync(mask, n, ws_counter, out_host, (cudaStream_t)stream);
}

int pdr_interpolate(const float* pos, const int* faces, const long long* face_idx,
                    const float* attr, const int* attr_faces, int V, int Vm, int res, int C,
                    int flip_y, float* out, uint8_t* mask_out, void* stream) {
  PDR_CHECK_ARG(pos && faces && face_idx && attr && attr_faces && out,
                "pdr_interpolate: null pointer");
  return interpolate_launch(pos, faces, face_idx, attr, attr_faces, V, Vm, res, C, flip_y, out,
                            mask_out, (cudaStream_t)stream);
}
int pdr_face_normals(const float* vertices, const int* faces, int F, float* out, void* stream) {
  PDR_CHECK_ARG(vertices && faces && out, "pdr_face_normals: null pointer");
  return face_normals_launch(vertices, faces, F, out, (cudaStream_t)stream);
}
int pdr_project_fixed(const float* cam_params, const float* vertices, int Vm, int V,
                      double padding, const float* uv_centers, const float* uv_scales,
                      const float* inpaint_scales, float* pos, void* stream) {
  PDR_CHECK_ARG(cam_params && vertices && uv_centers && uv_scales && inpaint_scales && pos,
                "pdr_project_fixed: null pointer");
  return project_fixed_launch(cam_params, vertices, Vm, V, padding, uv_centers, uv_scales,
                              inpaint_scales, pos, (cudaStream_t)stream);
}
int pdr_texopt_prepare(const float* uv_map, const uint8_t* mask, const uint8_t* vis,
                       const float* inpainted, int r0, int V, int res, int R, uint8_t* active,
                       float* target, long long* keys, void* stream) {
  PDR_CHECK_ARG(uv_map && mask && inpainted && active && target && keys,
                "pdr_texopt_prepare: null pointer");
  return texopt_prepare_launch(uv_map, mask, vis, inpainted, r0, V, res, R, active, target, keys,
                               (cudaStream_t)stream);
}
int pdr_texopt_build(const long long* sorted_keys, long long n_valid, const float* uv_map, int R,
                     unsigned int* entry_pix, double* entry_w, uint8_t* head, void* stream) {
  PDR_CHECK_ARG(sorted_keys && uv_map && entry_pix && entry_w && head,
                "pdr_texopt_build: null pointer");
  return texopt_build_launch(sorted_keys, n_valid, uv_map, R, entry_pix, entry_w, head,
                             (cudaStream_t)stream);
}
int pdr_texopt_forward(const float* atlas, const float* uv_map, const uint8_t* active,
                       const float* target, int V, int res, int R, signed char* signs,
                       double* images, void* stream) {
  PDR_CHECK_ARG(atlas && uv_map && active && target && signs, "pdr_texopt_forward: null pointer");
  return texopt_forward_launch(atlas, uv_map, active, target, V, res, R, signs, images,
                               (cudaStream_t)stream);
}
int pdr_texopt_step(float* atlas, float* m, float* v, const long long* sorted_keys,
                    const long long* seg_start, long long n_seg, const unsigned int* entry_pix,
                    const double* entry_w, const signed char* signs, int V, int res, int R,
                    float lerp_w, float beta2, float one_minus_beta2, float bc2_sqrt, float eps,
                    float neg_step_size, void* stream) {
  PDR_CHECK_ARG(atlas && m && v && sorted_keys && seg_start && entry_pix && entry_w && signs,
                "pdr_texopt_step: null pointer");
  return texopt_step_launch(atlas, m, v, sorted_keys, seg_start, n_seg, entry_pix, entry_w, signs,
                            V, res, R, lerp_w, beta2, one_minus_beta2, bc2_sqrt, eps,
                            neg_step_size, (cudaStream_t)stream);
}

int pdr_vertex_colors(const int* faces, const int* face_uv_idx, int F, const float* uvs, int Vn,
                      const float* atlas, const uint8_t* mask, int R, int* ws_uv_idx,
                      long long* pix, float* colors, float* count, uint8_t* has_color,
                      void* stream) {
  PDR_CHECK_ARG(faces && face_uv_idx && uvs && atlas && mask && ws_uv_idx && pix && colors &&
                    count && has_color,
                "pdr_vertex_colors: null pointer");
  return vertex_colors_launch(faces, face_uv_idx, F, uvs, Vn, atlas, mask, R, ws_uv_idx, pix,
                              colors, count, has_color, (cudaStream_t)stream);
}
int pdr_laplacian_round(const int* rowptr, const int* colidx, int Vn, const uint8_t* fixed,
                        const float* colors_in, const float* count_in, float* colors_out,
                        float* count_out, int* colored_total, void* stream) {
  PDR_CHECK_ARG(rowptr && colidx && fixed && colors_in && count_in && colors_out && count_out &&
                    colored_total,
                "pdr_laplacian_round: null pointer");
  return laplacian_round_launch(rowptr, colidx, Vn, fixed, colors_in, count_in, colors_out,
                                count_out, colored_total, (cudaStream_t)stream);
}
int pdr_scatter_vertex_colors(const long long* pix, const float* colors, int Vn, int R,
                              int* ws_winner, float* atlas, uint8_t* mask, void* stream) {
  PDR_CHECK_ARG(pix && colors && ws_winner && atlas && mask,
                "pdr_scatter_vertex_colors: null pointer");
  return scatter_vertex_colors_launch(pix, colors, Vn, R, ws_winner, atlas, mask,
                                      (cudaStream_t)stream);
}

int pdr_atlas_to_u8(const float* atlas, const uint8_t* mask, int R, uint8_t* rgb, uint8_t* rgba,
                    void* stream) {
  PDR_CHECK_ARG(atlas && rgb, "pdr_atlas_to_u8: null pointer");
  return atlas_to_u8_launch(atlas, mask, R, rgb, rgba, (cudaStream_t)stream);
}

}  // extern "C"
