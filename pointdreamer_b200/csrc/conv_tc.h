// Host interface of the tcgen05 implicit-GEMM convolution (conv_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace pdr {

// opaque CUtensorMap (128 bytes, 64-byte aligned)
struct alignas(64) ConvTensorMap {
  uint8_t bytes[128];
};

// 128-output-pixel tile of an H x W feature map: bw*bh*bb == 128
void conv_tc_pick_box(int B, int H, int W, int* bw, int* bh, int* bb);
// N tile that keeps all SMs busy for this layer: 64 / 128 / 256, or 512 = the 2-CTA kernel
// (cta_group::2, 256-wide tile per SM pair; its weight map uses a 128-row box)
int conv_tc_pick_bn(int B, int H, int W, int Cout, int taps = 1);

// NHWC fp16 activation [B,H,W,C], C % 64 == 0
// halo != 0: the box is the 10 x 18 pixel halo of an 8 x 16 output tile (conv_halo_kernel)
int conv_tc_make_act_map(ConvTensorMap* out, const void* ptr, int B, int H, int W, int C,
                         int halo = 0);
// 3x3 convs on maps with W % 8 == 0 and H % 16 == 0 run the halo kernel: all nine taps of a
// 64-channel chunk read ONE shared-memory tile (activation maps must be encoded with halo = 1)
bool conv_tc_halo_ok(int H, int W, int taps);
// generic 2-D fp16 map (SWIZZLE_128B): matrix [outer][inner] with `inner` contiguous, box
// {box_inner (64 = 128 bytes), box_outer}; used by the tcgen05 attention for the [B*T, 3C] qkv view
int conv_tc_make_map_2d(ConvTensorMap* out, const void* ptr, unsigned long long inner,
                        unsigned long long outer, int box_inner, int box_outer);
// fp16 weights [Cout][K] with K = taps*Cin ordered (tap, channel)
int conv_tc_make_weight_map(ConvTensorMap* out, const void* ptr, int Cout, int K, int BN);

// out[B,H,W,Cout] = conv(cat(A1,A2)) + bias (+ residual); taps = 9 (3x3 pad 1) or 1 (1x1).
// Fused 1x1 skip branch (ResBlock skip_connection, unet.py:219-222,256): with S1 (+S2) > 0 the
// GEMM's K dimension is extended by the channels of cat(s1, s2) read at the centre tap, so
// out = conv3x3(cat(A1,A2)) + conv1x1(cat(s1,s2)) + bias in ONE accumulation; the weight matrix
// is [Cout][taps*(C1+C2) + S1+S2] (skip weights appended along K) and bias = both biases summed.
// qk_scale != 0 (qkv projections only): the q and k channels of the legacy head-major layout
// (c % 192 < 128) are stored as fp16(fp16(acc + bias) * qk_scale), i.e. the attention's
// `q * scale`, `k * scale` (unet.py:349-351) is applied here
int conv_tc_launch(const ConvTensorMap* a1, const ConvTensorMap* a2, const ConvTensorMap* w,
                   int BN, int B, int H, int W, int C1, int C2, int Cout, int taps,
                   const float* bias, const __half* residual, __half* out, float* stats_partial,
                   cudaStream_t stream, float qk_scale = 0.f, const ConvTensorMap* s1 = nullptr,
                   const ConvTensorMap* s2 = nullptr, int S1 = 0, int S2 = 0, int ksplit = 1,
                   float* splitk_ws = nullptr, int halo = 0, const float4* gn_coeff = nullptr,
                   int gn_film = 0, const ConvTensorMap* omap = nullptr);
// omap (plain kernels only): tensor map of `out` made by conv_tc_make_act_map(out, B, H, W, Cout);
// the epilogue then stages 128-pixel x 64-channel blocks in shared memory and writes them with TMA
// stores instead of per-thread 16-byte stores
// gn_coeff (halo kernel only): the A maps point at the RAW tensor and GroupNorm32 (+FiLM when
// gn_film) + SiLU is applied to each halo tile in shared memory with the constants of
// gn_coeff_launch (unet_ops.h): float4 (ga, gb, fs, fsh) per (image, channel of cat(A1, A2))
// split-K for layers with too few tiles to fill the GPU: returns the number of K ranges (1 = off)
// for a layer with `num_k` 64-channel K steps and may change *bn; the workspace holds the fp32
// partial tiles (conv_tc_split_workspace_bytes) and a finishing kernel applies the epilogue
int conv_tc_pick_split(int B, int H, int W, int Cout, int num_k, int* bn);
size_t conv_tc_split_workspace_bytes(int B, int H, int W, int Cout, int ksplit);
// number of (m-tile, epilogue-warp) partial rows written per image when stats_partial is used,
// or 0 when the fused statistics are not available for this spatial size
int conv_tc_stats_rows_per_image(int H, int W);

}  // namespace pdr
