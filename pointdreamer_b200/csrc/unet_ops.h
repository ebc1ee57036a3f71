// Host launchers of the non-GEMM U-Net kernels (unet_ops.cu, attention.cu) and the DDNM sampler
// kernels (ddnm.cu).  All pointers are device pointers.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace pdr {

// out[b][n] = bias[n] + sum_k f(in[b][k]) W[n][k]; mode_in 0 identity, 1 SiLU, 2 timestep embedding
// skip: optional device flag; the launch does nothing when *skip >= 0 (embedding cache hit)
int linear_launch(const float* in, const float* W, const float* bias, int B, int K, int N,
                  int mode_in, float* out, __half* out16, cudaStream_t stream,
                  const int* skip = nullptr);

// Timestep-embedding cache.  The FiLM vectors of all ResBlocks (emb16 [B][etot] fp16) are a pure
// function of (weights, t); a sampler visits the same timesteps for every shape.  The cache lives in
// device memory and is looked up ON the device (no host knowledge of t, graph-capturable):
//   lookup: all B timesteps equal and cached -> meta.hit = slot, else -1 (and meta.store = a free
//           slot or -1);  the three embedding linears are launched with skip = &meta.hit;
//   finish: hit -> broadcast the cached row to the B rows of emb16; store -> keep row 0.
// Cached rows are the bits a recomputation would produce (every row is computed independently).
static constexpr int EMB_CACHE_SLOTS = 128;
struct EmbCacheMeta {
  float t[EMB_CACHE_SLOTS];
  int n, hit, store, pad;
};
int emb_cache_lookup_launch(const float* t, int B, EmbCacheMeta* meta, int enabled,
                            cudaStream_t stream);
int emb_cache_finish_launch(__half* emb16, int B, int etot, __half* cache, EmbCacheMeta* meta,
                            const float* t, cudaStream_t stream);

int stem_conv_launch(const float* x, const __half* w, const float* bias, int B, int H, int W,
                     int C, __half* out, cudaStream_t stream);

// x [B,3,H,W] fp32 -> [B,H,W,64] fp16 3x3 patches (27 taps*channels, zero padded to 64)
int stem_im2col_launch(const float* x, int B, int H, int W, __half* out, cudaStream_t stream);

int gn_stats_slabs(int B, int HW);
// ws_partial: float[B * slabs * 2 * (C1+C2)];  stats: float[B*32*2] (mean, rstd)
int gn_stats_launch(const __half* x1, const __half* x2, int B, int HW, int C1, int C2,
                    float* ws_partial, float* stats, cudaStream_t stream);
// statistics fused into the producing tcgen05 conv: partial rows -> per-8-channel fp64 sums ->
// (mean, rstd) of a (possibly concatenated) tensor
int sums8_reduce_launch(const float* partial, int B, int R, int C, double* sums8,
                        cudaStream_t stream);
int gn_finalize_sums_launch(const double* s1, const double* s2, int B, int HW, int C1, int C2,
                            float* stats, cudaStream_t stream);
int gn_apply_launch(const __half* x1, const __half* x2, int B, int H, int W, int C1, int C2,
                    const float* stats, const double* sums1, const double* sums2,
                    const float* gamma, const float* beta,
                    const __half* film, int film_stride, int film_off, int silu, int resample,
                    __half* out, cudaStream_t stream, __half* raw_out = nullptr);
// constants of a GroupNorm32 (+FiLM) applied inside the consuming conv: coeff float4[B][C1+C2]
int gn_coeff_launch(int B, int H, int W, int C1, int C2, const float* stats, const double* sums1,
                    const double* sums2, const float* gamma, const float* beta, const __half* film,
                    int film_stride, int film_off, float4* coeff, cudaStream_t stream);
int resample_launch(const __half* x, int B, int H, int W, int C, int mode, __half* out,
                    cudaStream_t stream);
int head_launch(const __half* h, const float* stats, const float* gamma, const float* beta,
                const float* w, const float* bias, int B, int H, int W, int C, int n_out,
                float* out, int out_channels_total, cudaStream_t stream);

// qkv [B,T,3C] (legacy head-major layout) -> a [B,T,C]; head dim 64
// prescaled != 0: q and k were already multiplied by 64^-1/4 (fp16) by the qkv conv's epilogue
// tcgen05 version (attention_tc.cu): T % 128 == 0 and q, k already scaled
bool attention_tc_ok(int T, int prescaled);
int attention_tc_launch(const __half* qkv, int B, int T, int heads, __half* out, cudaStream_t stream);
int attention_launch(const __half* qkv, int B, int T, int heads, int prescaled, __half* out,
                     cudaStream_t stream);

// ---- DDNM sampler ----
struct DdnmStepCoef {
  float sqrt_1m_at, sqrt_at, sqrt_at_next, gamma_t, c1, c2, lambda_t;
};
// threads-per-draw / philox counter increment of torch.randn for `numel` on this device
void philox_launch_geometry(long long numel, long long* threads, unsigned long long* counter_inc);
int ddnm_prepare_launch(const float* sparse, const float* mask, int V, int C, int H, int W,
                        unsigned long long seed, unsigned long long offset_base,
                        unsigned long long draws_per_chain, int chain0, float* y, float* x,
                        cudaStream_t stream);
int ddnm_step_launch(float* x, const float* et, int et_channels, const float* y, const float* mask,
                     int V, int C, int H, int W, DdnmStepCoef coef, unsigned long long seed,
                     unsigned long long offset_base, unsigned long long draws_per_chain,
                     int chain0, int draw_index, cudaStream_t stream);
int ddnm_final_launch(const float* x, long long n, float* out, cudaStream_t stream);
int randn_like_torch_launch(float* out, long long numel, unsigned long long seed,
                            unsigned long long offset, cudaStream_t stream);

}  // namespace pdr
