/* pdr.h — C ABI of libpdr.so, the B200-native (sm_100a) implementation of PointDreamer's
 * project -> DDNM-inpaint -> unproject hot path.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - layouts are the contiguous layouts of the reference's PyTorch tensors (stated per call);
 *     masks cross the ABI as uint8 (0/1);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no call synchronises
 *     unless documented;
 *   - return 0 = ok, < 0 = argument/shape error, > 0 = cudaError_t; pdr_last_error() describes it;
 *   - the library owns no global mutable state besides the last-error string, a launch counter
 *     and handles created by pdr_*_create().
 *
 * Each entry point cites the reference interface it replaces (file:line under the reference
 * repository YuQiao0303/PointDreamer @ 6fa8552).
 */
#ifndef PDR_H
#define PDR_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDR_VERSION 100 /* 0.1.0 */

/* ------------------------------------------------------------------ core ---------------- */
int pdr_version(void);
const char* pdr_last_error(void);
/* number of kernels this library has launched since load (bench.py "gpu_launches") */
unsigned long long pdr_launch_count(void);

/* ------------------------------------------------------------ U-Net layers -------------- */
/* fp16 NHWC convolution on the tcgen05 tensor cores (3x3 pad 1 when taps==9, 1x1 when taps==1).
 * Replaces nn.Conv2d / nn.Conv1d(k=1) in the fp16 torso of the ADM U-Net
 * (models/DDNM/guided_diffusion/unet.py:176-222, 291-294; nn.py:22-32).
 *   x1 [B,H,W,C1] fp16, x2 [B,H,W,C2] fp16 or NULL (channel concat, unet.py:660-662)
 *   w  [Cout][taps*(C1+C2)] fp16, K index = tap*(C1+C2)+c, tap = ky*3+kx
 *   bias [Cout] fp32 or NULL; residual [B,H,W,Cout] fp16 or NULL; out [B,H,W,Cout] fp16
 *   bn: N tile, 64 / 128 / 256 (must divide Cout), 512 = 2-CTA kernel (cta_group::2, SM pairs share
 *   the weight tile; needs Cout % 256 == 0 and an even number of 128-pixel tiles), 0 = auto.
 *   3x3 convs on maps with W % 8 == 0 and H % 16 == 0 run the halo kernel (all nine taps of a
 *   64-channel chunk read one 10x18-pixel shared-memory tile); every variant accumulates K in the
 *   same order, so the output bits do not depend on bn.
 *   C1, C2, Cout % 64 == 0. */
int pdr_conv_tc(const void* x1, const void* x2, const void* w, const float* bias,
                const void* residual, void* out, int B, int H, int W, int C1, int C2, int Cout,
                int taps, int bn, void* stream);

/* ResBlock tail in one GEMM: out = conv3x3(x) + conv1x1(cat(s1,s2)) + bias, i.e. out_layers.3 and
 * skip_connection of a channel-changing ResBlock (unet.py:206-222, 256) accumulated together
 * (one fp16 rounding of the sum; the reference rounds each branch and the sum).
 *   x [B,H,W,C] fp16 ; s1 [B,H,W,S1], s2 [B,H,W,S2] or NULL fp16 ;
 *   w [Cout][9*C + S1 + S2] fp16 (3x3 weights tap-major, then the 1x1 weights) ;
 *   bias [Cout] fp32 (sum of both biases) ; out [B,H,W,Cout] fp16 ; bn as in pdr_conv_tc */
int pdr_conv_tc_skip(const void* x, const void* w, const float* bias, const void* s1,
                     const void* s2, void* out, int B, int H, int W, int C, int S1, int S2,
                     int Cout, int bn, void* stream);

/* Individual non-GEMM U-Net kernels (exported for unit parity tests; the engine below calls the
 * same launchers).  NHWC fp16 activations, fp32 parameters.
 *   pdr_linear: out[b][n] = bias[n] + sum_k f(in[b][k]) W[n][k]; mode_in 0 id, 1 SiLU,
 *               2 in = timestep_embedding(t[b], K)  (nn.py:103-121; unet.py:450-455, 199-205)
 *   pdr_stem_conv: x [B,3,H,W] fp32 -> fp16 -> conv3x3 -> [B,H,W,C] fp16   (unet.py:482-484,655)
 *   pdr_group_norm: GroupNorm32(32 groups) of cat(x1,x2) (+FiLM (1+scale),shift from
 *               film[b][film_off + {c, C+c}]) (+SiLU) (+resample 1=avgpool2, 2=nearest x2)
 *               (nn.py:17-19, unet.py:236-252);  ws: float[B*slabs*2*C], stats: float[B*64]
 *   pdr_resample: avg-pool 2 (mode 1) / nearest x2 (mode 2)                  (unet.py:92-140)
 *   pdr_attention: QKVAttentionLegacy on qkv [B,T,3C] -> [B,T,C], head dim 64 (unet.py:328-358)
 *   pdr_unet_head: GN + SiLU + conv3x3(C -> n_out) in fp32, NCHW fp32 output  (unet.py:613-617)*/
int pdr_linear(const float* in, const float* W, const float* bias, int B, int K, int N,
               int mode_in, float* out, void* out_fp16, void* stream);
int pdr_stem_conv(const float* x, const void* w, const float* bias, int B, int H, int W, int C,
                  void* out, void* stream);
int pdr_group_norm(const void* x1, const void* x2, int B, int H, int W, int C1, int C2,
                   const float* gamma, const float* beta, const void* film, int film_stride,
                   int film_off, int silu, int resample, float* ws, float* stats, void* out,
                   void* stream);
int pdr_resample(const void* x, int B, int H, int W, int C, int mode, void* out, void* stream);
int pdr_attention(const void* qkv, int B, int T, int heads, void* out, void* stream);
/* the same with q and k ALREADY multiplied by 64^-1/4 and rounded to fp16 (what the engine's qkv
 * projection emits); sequences with T % 128 == 0 run on tcgen05 (S and O in tensor memory) */
int pdr_attention_prescaled(const void* qkv, int B, int T, int heads, void* out, void* stream);
int pdr_unet_head(const void* h, const float* gamma, const float* beta, const float* w,
                  const float* bias, int B, int H, int W, int C, int n_out, float* ws,
                  float* stats, float* out, void* stream);

/* --------------------------------------------------------- U-Net engine + DDNM ---------- */
/* Structure of the ADM U-Net, the arguments of script_util.create_model (script_util.py:130-185)
 * as resolved from models/DDNM/configs/imagenet_256.yml. channel_mult is stored doubled so the
 * 512-pixel preset's 0.5 is representable. */
typedef struct PdrUnetConfig {
  int image_size, in_channels, model_channels, out_channels, num_res_blocks;
  int n_mult, channel_mult_x2[8];
  int n_attn_ds, attn_ds[8]; /* downsample factors that carry attention */
  int num_head_channels;
} PdrUnetConfig;

/* Replaces Diffusion.get_model / create_model + convert_to_fp16 (diffusion.py:435-457).
 * Parameters are registered under the reference's state_dict names:
 *   conv weights  fp16 [Cout][taps*Cin] (tap-major: K = (ky*3+kx)*Cin + c), biases fp32;
 *   GroupNorm / Linear / out.2 parameters fp32 in PyTorch layout;
 *   "emb_all.weight" [sum 2*Cout, 4*mc] / "emb_all.bias": every ResBlock's emb_layers.1
 *   concatenated in module order (input_blocks, middle_block, output_blocks);
 *   "<resblock>.out_layers.3_skip.weight" fp16 [Cout][9*Cout + Cin] / ".bias" fp32 for every
 *   ResBlock with a skip_connection: out_layers.3's [Cout][9*Cout] followed along K by the 1x1
 *   skip weights [Cout][Cin], biases summed - the skip branch runs inside out_layers.3's GEMM;
 *   optional "input_blocks.0.0_tc.weight" fp16 [C][64] (the stem's [C][27] zero padded) +
 *   "input_blocks.0.0_tc.bias": runs the stem as a tensor-core GEMM over 3x3 patches. */
int pdr_unet_create(const PdrUnetConfig* cfg, void** handle);
int pdr_unet_destroy(void* handle);
int pdr_unet_set_param(void* handle, const char* name, const void* ptr, size_t bytes);
int pdr_unet_workspace_bytes(void* handle, int B, size_t* bytes);
int pdr_unet_plan(void* handle, int B, void* workspace, size_t bytes);
/* UNetModel.forward (unet.py:635-664): x [B,3,S,S] fp32, t [B] fp32 -> out [B,n_out,S,S] fp32 */
int pdr_unet_forward(void* handle, const float* x, const float* t, float* out, int n_out,
                     void* stream);

/* Live per-kernel-class timing of the engine (CUDA events on the launching stream around every
 * launch of each `every`-th forward).  pdr_unet_profile_end synchronises, fills the
 * PDR_OP_CLASSES-long arrays (milliseconds, algorithmic FLOPs of the tensor-core launches,
 * launch counts) and stops profiling. */
enum PdrOpClass { PDR_OP_CONV_TC = 0, PDR_OP_GN_STATS, PDR_OP_GN_APPLY, PDR_OP_RESAMPLE,
                  PDR_OP_ATTENTION, PDR_OP_LINEAR, PDR_OP_STEM, PDR_OP_HEAD, PDR_OP_CLASSES };
int pdr_unet_profile_begin(void* handle, int every, int max_forwards);
int pdr_unet_profile_end(void* handle, double* ms, double* flops, long long* launches,
                         long long* forwards);

/* torch.randn-compatible standard normals (Philox4x32-10 + Box-Muller, torch's thread mapping):
 * out[numel] == torch.randn(numel, device='cuda') drawn with (seed, philox offset). */
int pdr_randn_like_torch(float* out, long long numel, unsigned long long seed,
                         unsigned long long offset, void* stream);
/* philox offset increment of one torch.randn(numel) call on this device */
unsigned long long pdr_randn_offset_increment(long long numel);

/* Diffusion.simplified_ddnm_inpainting (diffusion.py:459-570) for V chains at once.
 *   sparse [V,3,S,S] fp32 in [0,1]; mask [V,S,S] fp32 (1 = known)
 *   coef_host: HOST float[steps][7] = sqrt(1-at), sqrt(at), sqrt(at_next), gamma_t, c1, c2, lambda_t
 *   t_dev: float[steps][V] (timestep fed to the U-Net at every step)
 *   noise of chain v, draw d comes from torch's global-generator stream at draw index
 *   (chain0 + v)*draws_per_chain + d  (d = 0: x_T, d = 1+s: step s), see SURVEY Appendix C
 *   x, y: float[V*3*S*S] scratch; et: float[V*3*S*S] scratch; out [V,3,S,S] fp32 in [0,1] */
int pdr_ddnm_sample(void* unet, const float* sparse, const float* mask, int V, int steps,
                    const float* coef_host, const float* t_dev, unsigned long long seed,
                    unsigned long long offset_base, unsigned long long draws_per_chain,
                    int chain0, float* x, float* y, float* et, float* out, void* stream);
/* single pieces of the sampler (exported for parity tests) */
int pdr_ddnm_prepare(const float* sparse, const float* mask, int V, int S,
                     unsigned long long seed, unsigned long long offset_base,
                     unsigned long long draws_per_chain, int chain0, float* y, float* x,
                     void* stream);
int pdr_ddnm_step(float* x, const float* et, int et_channels, const float* y, const float* mask,
                  int V, int S, const float* coef7_host, unsigned long long seed,
                  unsigned long long offset_base, unsigned long long draws_per_chain, int chain0,
                  int draw_index, void* stream);
int pdr_ddnm_final(const float* x, long long n, float* out, void* stream);

/* --------------------------------------------------------------- PROJECT --------------- */
/* Camera transform + crop/rescale of mesh vertices and cloud points for all V views.
 * Replaces ours_utils.py:93-130 (get_rendered_hard_mask_and_face_idx_batch up to the
 * rasterize call) incl. kaolin Camera.transform (ours_utils.py:99; canonical arithmetic in
 * oracle/camera.py).
 *   cam_params [V,16] fp32 (r00..r22, t0..t2, f, za, zb, 0); vertices [Vm,3]; points [N,3]
 *   ws_minmax: int[4*V] scratch
 *   pos [V,Vm,4] (rescaled NDC xy, NDC z, 1) ; vertice_uvs [V,Vm,2] ; uv_centers [V,2] ;
 *   uv_scales [V] ; point_uvs [V,N,2] ; point_depths [V,N]                     (all fp32) */
int pdr_project(const float* cam_params, const float* vertices, int Vm, const float* points,
                int N, int V, int rescale, double padding, int* ws_minmax, float* pos,
                float* vertice_uvs, float* uv_centers, float* uv_scales, float* point_uvs,
                float* point_depths, void* stream);

/* Z-buffer rasteriser of the mesh for all views.  Replaces nvdiffrast.torch.rasterize as
 * used at ours_utils.py:142-147 plus the 512->256 mask resize of demo.py:103-104.
 * Tile-binned: triangles are binned into 32 x 8 pixel tiles, one CTA per tile walks its bin from
 * shared memory with one thread per pixel (no atomics on the z-buffer).
 *   pos [V,Vm,4] fp32 ; faces [F,3] int32 ;
 *   workspace: pdr_rasterize_workspace_bytes(V, F, res) bytes, 16-byte aligned
 *   depth [V,res,res] fp32 (0 empty) ; face_idx [V,res,res] int64 (-1 empty) ;
 *   mask_cam [V,res,res] u8 ; mask_out [V,out_res,out_res] u8 (out_res == res or res/2) */
size_t pdr_rasterize_workspace_bytes(int V, int F, int res);
int pdr_rasterize(const float* pos, const int* faces, int V, int Vm, int F, int res, int out_res,
                  void* workspace, float* depth, long long* face_idx, uint8_t* mask_cam,
                  uint8_t* mask_out, void* stream);

/* 2x mask reduction.  Replaces demo.py:103-104 (torchvision Resize, bilinear without antialias,
 * then .bool()  ==  OR of each 2x2 block).  mask_in [V,res_in,res_in] u8 -> [V,res_in/2,res_in/2] */
int pdr_mask_half_any(const uint8_t* mask_in, int V, int res_in, uint8_t* mask_out, void* stream);

/* Depth visibility + pixel quantisation.  Replaces ours_utils.py:153-202
 * (get_point_validation_by_depth) and demo.py:121-125 (point_pixels at `res`).
 * Any of vis / pix_cam / pix_res may be NULL.
 *   point_uvs [V,N,2] ; point_depths [V,N] ; mesh_depths [V,cam_res,cam_res]
 *   vis [V,N] u8 ; pix_cam [V,N,2] int64 (row,col at cam_res) ; pix_res [V,N,2] int64 */
int pdr_point_visibility(const float* point_uvs, const float* point_depths,
                         const float* mesh_depths, int V, int N, int cam_res, float offset,
                         int res, uint8_t* vis, long long* pix_cam, long long* pix_res,
                         void* stream);

/* Hidden point removal for all views.  Replaces ours_utils.py:204-225 get_point_validation_by_o3d
 * (open3d PointCloud.hidden_point_removal(eye, radius): spherical flip + float64 Qhull, visible =
 * hull vertices).  Hull-vertex membership is decided per point by a 2-variable LP in fp64 after a
 * projective map that sends the eye to infinity (see csrc/geom_hpr.cu).
 *   points [N,3] fp32 ; frames [V,12] fp64 DEVICE = eye(3), ex(3), ey(3), ez(3) with ez the unit
 *   vector from the eye towards the scene and (ex, ey, ez) orthonormal ; radius: the HPR radius
 *   workspace: pdr_hidden_point_removal_workspace_bytes(V,N) bytes ; vis [V,N] u8 */
size_t pdr_hidden_point_removal_workspace_bytes(int V, int N);
int pdr_hidden_point_removal(const float* points, int N, int V, const double* frames,
                             double radius, void* workspace, uint8_t* vis, void* stream);

/* Sparse view images + hole masks.  Replaces ours_utils.py:848-882 get_sparse_images
 * (get_one_sparse_img 954-1044, paint_pixels 456-495, inner edge mask 497-532, kaolin
 * sided_distance 1013).
 *   point_pixels [V,N,2] int64 (row,col) ; colors [N,3] fp32 ; valid [V,N] u8 ;
 *   hard_masks [V,res,res] u8 ; workspace: pdr_sparse_images_workspace_bytes(V,res) bytes
 *   sparse / hard_mask0 / hard_mask2 [V,3,res,res] fp32 ; scale_factors [V] fp32 */
size_t pdr_sparse_images_workspace_bytes(int V, int res);
int pdr_sparse_images(const long long* point_pixels, const float* colors, const uint8_t* valid,
                      const uint8_t* hard_masks, int V, int N, int res, int point_size,
                      int edge_point_size, double mask_ratio_thresh, void* workspace,
                      float* sparse, float* hard_mask0, float* hard_mask2, float* scale_factors,
                      void* stream);

/* Exact nearest-valid-pixel fill.  Replaces ours_utils.py:610-643 naive_inpainting('nearest')
 * (scipy griddata) and unproject.py:480-504 dilate_atlas.
 *   img [B,C,H,W] fp32 (channels_last=0) or [B,H,W,C] (channels_last=1) ; known [B,H,W] u8 ;
 *   out same layout as img ; src_index [B,H,W] int32 (linear index of the source) or NULL */
size_t pdr_nearest_fill_workspace_bytes(int B, int H, int W);
int pdr_nearest_fill(const float* img, const uint8_t* known, int B, int C, int H, int W,
                     int channels_last, void* workspace, float* out, int* src_index,
                     void* stream);

/* ------------------------------------------------------------- UNPROJECT --------------- */
/* Back-projection of the inpainted views into the UV atlas with Non-Border-First selection.
 * Replaces unproject.py:201-425 (unproject), :429-475 (NBF shrink), utils_2d.py:799-845.
 *   images [V,3,res,res] fp32 ; cam_params [V,16] ; base_dirs [V,3] ; gb_pos [R,R,3] ;
 *   mask [R,R] u8 ; face_id [R,R] int64 ; f_normals [F,3] ; uv_centers [V,2] ; uv_scales [V] ;
 *   scale_factors [V] ; mesh_depths [V,cam_res,cam_res] ;
 *   kernels_host: HOST int[n_levels] = edge_dilate_kernels ; rescale = 0 only when the
 *   reference's uv_centers/uv_scales/padding/scale_factors would be None (unproject.py:260)
 *   atlas [R,R,3] fp32 ; shrinked_vis [V,R,R] u8 ; point_view_ids [P] int64 ;
 *   point_coords [P,2] int64 ; points [P,3] fp32 ; painted [R,R] u8  (P = count of mask) */
size_t pdr_unproject_workspace_bytes(int R, int n_levels);
int pdr_unproject(const float* images, int res, const float* cam_params, int V, int cam_res,
                  const float* base_dirs, const float* gb_pos, const uint8_t* mask,
                  const long long* face_id, int R, const float* f_normals, int F,
                  const float* uv_centers, const float* uv_scales, double padding, int rescale,
                  const float* scale_factors, const float* mesh_depths, const int* kernels_host,
                  int n_levels, int complete_unseen, void* workspace, float* atlas,
                  uint8_t* shrinked_vis, long long* point_view_ids, long long* point_coords,
                  float* points, uint8_t* painted, void* stream);
/* number of set bytes in mask[n]; synchronises `stream`. ws_counter: int[1] device scratch */
int pdr_mask_count(const uint8_t* mask, size_t n, int* ws_counter, int* out_host, void* stream);

/* ------------------------------------- "next" rows: atlas inputs (N4), optimiser (N1) --- */
/* Attribute interpolation over a rasterised mesh.  Replaces nvdiffrast.torch.interpolate as used
 * at models/get3d/extract_texture_map.py:60 (gb_pos per atlas texel) and ours_utils.py:1697
 * (texture uv per view pixel).  Canonical rule (nvdiffrast's scheme): fp32 barycentrics u = eA/(eA+eB+eC),
 * v = eB/(eA+eB+eC) from the rasteriser's exact integer edge functions, then
 * (u*a0 + v*a1) + ((1-u)-v)*a2 (oracle/project.py:interpolate).
 *   pos [V,Vm,4] fp32 and faces [F,3] int32: what pdr_rasterize was called with ;
 *   face_idx [V,res,res] int64 from pdr_rasterize ; attr [Na,C] fp32 (C = 2 or 3) ;
 *   attr_faces [F,3] int32 (attribute index triple per face) ; flip_y != 0: output row y is
 *   raster row res-1-y (torch.flip(..,[1]), ours_utils.py:1702-1705)
 *   out [V,res,res,C] fp32 (0 where empty) ; mask_out [V,res,res] u8 or NULL (same flip) */
int pdr_interpolate(const float* pos, const int* faces, const long long* face_idx,
                    const float* attr, const int* attr_faces, int V, int Vm, int res, int C,
                    int flip_y, float* out, uint8_t* mask_out, void* stream);
/* Unit face normals.  Replaces kal.ops.mesh.face_normals(face_vertices, unit=True), demo.py:422.
 *   vertices [Vm,3] fp32 ; faces [F,3] int32 ; out [F,3] fp32 */
int pdr_face_normals(const float* vertices, const int* faces, int F, float* out, void* stream);
/* Per-view clip-space vertices of optimize_color (ours_utils.py:1676-1693): camera transform,
 * then ((uv - c)/s) * (1 - 2*padding) * inpaint_scale + 0.5, clip [0,1], *2-1.
 *   cam_params [V,16] ; vertices [Vm,3] ; uv_centers [V,2] ; uv_scales [V] ; inpaint_scales [V]
 *   pos [V,Vm,4] fp32 (x, y, NDC z, 1) */
int pdr_project_fixed(const float* cam_params, const float* vertices, int Vm, int V,
                      double padding, const float* uv_centers, const float* uv_scales,
                      const float* inpaint_scales, float* pos, void* stream);

/* Texture optimiser.  Replaces optimize_color, ours_utils.py:1583-1785 (100 Adam iterations of
 * an L1 fit of the atlas' float64 bilinear renders to the inpainted views; kaolin
 * texture_mapping == F.grid_sample(align_corners=False, padding_mode='border'), y reversed).
 * The atlas is the [3,R,R] planar fp32 parameter in the frame optimize_color receives it
 * (demo.py:217: permute(2,0,1).flip(1)).  Call order: prepare -> (host: sort keys, count keys
 * != INT64_MAX) -> build -> (host: seg_start = indices where head == 1, then n_valid appended)
 * -> iterations x (forward, step).
 *   prepare: uv_map [V,res,res,2] fp32 and mask [V,res,res] u8 in the FLIPPED frame
 *            (pdr_interpolate with flip_y=1) ; vis [V,R,R] u8 shrinked visibility or NULL ;
 *            inpainted [V,3,r0,r0] fp32 (bilinearly resized to res like transforms.Resize) ->
 *            active [V,res,res] u8 ; target [V,res,res,3] fp32 ; keys int64[V*res*res*4]
 *            (texel << 32 | (pixel*4 + corner), INT64_MAX = no contribution)
 *   build:   sorted keys -> entry_pix u32[n_valid], entry_w f64[n_valid], head u8[n_valid]
 *   forward: signs int8[V*res*res*4] (zero-initialised by the caller) ; images f64 [V,3,res,res]
 *            or NULL (the reference's second return value)
 *   step:    gradient gather + one Adam update of the touched texels; m, v: fp32 [3,R,R] state;
 *            lerp_w = 1-beta1, bc2_sqrt = sqrt(1-beta2^t), neg_step_size = -lr_t/(1-beta1^t)
 *            (torch.optim.Adam's foreach formulas) */
int pdr_texopt_prepare(const float* uv_map, const uint8_t* mask, const uint8_t* vis,
                       const float* inpainted, int r0, int V, int res, int R, uint8_t* active,
                       float* target, long long* keys, void* stream);
int pdr_texopt_build(const long long* sorted_keys, long long n_valid, const float* uv_map, int R,
                     unsigned int* entry_pix, double* entry_w, uint8_t* head, void* stream);
int pdr_texopt_forward(const float* atlas, const float* uv_map, const uint8_t* active,
                       const float* target, int V, int res, int R, signed char* signs,
                       double* images, void* stream);
int pdr_texopt_step(float* atlas, float* m, float* v, const long long* sorted_keys,
                    const long long* seg_start, long long n_seg, const unsigned int* entry_pix,
                    const double* entry_w, const signed char* signs, int V, int res, int R,
                    float lerp_w, float beta2, float one_minus_beta2, float bc2_sqrt, float eps,
                    float neg_step_size, void* stream);

/* ----------------------- "next" row N2: never-seen texels from mesh neighbours ---------- */
/* Pieces of paint_invisible_areas_by_neighbors (unproject.py:93-196, use_atlas=True) after the
 * host-side mesh subdivision (utils/mesh_utils.py:7-114).
 *   vertex_colors (unproject.py:117-134): faces, face_uv_idx [F,3] int32 ; uvs [Nu,2] fp32 ;
 *     atlas [R,R,3] fp32 ; mask [R,R] u8 (painted) ; ws_uv_idx int[Vn] scratch ->
 *     pix int64[Vn] (row*R+col of each vertex' texel), colors [Vn,3], count [Vn] fp32 (1 = has
 *     colour), has_color u8[Vn].  A seam vertex takes its largest uv index.
 *   laplacian_round (unproject.py:160-163 with kaolin's uniform Laplacian as a CSR adjacency,
 *     rowptr int[Vn+1], colidx int[nnz], neighbours ascending): every vertex with fixed == 0 gets
 *     the count-weighted mean of its neighbours (or keeps its value); colors_out/count_out of
 *     fixed vertices are not written (initialise both buffers alike); colored_total: device
 *     int = number of vertices with a colour after the round.
 *   scatter_vertex_colors (unproject.py:181-182): atlas[pix[v]] = colors[v], mask = 1; a texel
 *     shared by several vertices takes the highest vertex index; ws_winner int[R*R] scratch. */
int pdr_vertex_colors(const int* faces, const int* face_uv_idx, int F, const float* uvs, int Vn,
                      const float* atlas, const uint8_t* mask, int R, int* ws_uv_idx,
                      long long* pix, float* colors, float* count, uint8_t* has_color,
                      void* stream);
int pdr_laplacian_round(const int* rowptr, const int* colidx, int Vn, const uint8_t* fixed,
                        const float* colors_in, const float* count_in, float* colors_out,
                        float* count_out, int* colored_total, void* stream);
int pdr_scatter_vertex_colors(const long long* pix, const float* colors, int Vn, int R,
                              int* ws_winner, float* atlas, uint8_t* mask, void* stream);

/* ----------------------------------- "next" row N3: output formats ---------------------- */
/* 8-bit atlas exactly as save_textured_mesh quantises it (demo.py:283-301): img*255 in fp32,
 * clip [0,255], truncate, rows flipped.  atlas [R,R,3] fp32 ; mask [R,R] u8 or NULL ->
 * rgb u8 [R,R,3] (model_normalized.png) ; rgba u8 [R,R,4] or NULL (atlas_wo_background.png,
 * alpha = mask*255). */
int pdr_atlas_to_u8(const float* atlas, const uint8_t* mask, int R, uint8_t* rgb, uint8_t* rgba,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDR_H */
