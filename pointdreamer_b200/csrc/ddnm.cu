// DDNM inpainting sampler kernels (K10): x_T / y preparation, the fused per-step update with
// in-kernel Philox noise that reproduces torch.randn's stream bit for bit, final transform.
//
// Reference: models/DDNM/guided_diffusion/diffusion.py:459-570 (simplified_ddnm_inpainting),
// models/DDNM/datasets/__init__.py:208-234 (data_transform / inverse_data_transform).
// Every elementwise expression keeps the reference's op order with one rounding per torch op
// (__fmul_rn/__fadd_rn/__fdiv_rn are never contracted).  The file itself is compiled with the
// default -fmad=true because cuRAND's Box-Muller must contract exactly as it does inside
// PyTorch's own build for the noise to be bit-identical to torch.randn.
//
// Noise: the reference draws torch.randn(1,3,H,W) once per chain for x_T and randn_like once
// per step, from the global CUDA generator, chain after chain.  Draw k of
// that stream uses Philox offset base + k*counter_inc; element i of a draw is component
// (i / T) % 4 of curand_normal4 from subsequence (i % T), T = threads of torch's launch
// (ATen/native/cuda/DistributionTemplates.h: block 256, grid = min(SMs*blocks/SM, ceil(n/256)),
// unroll 4).  A batched sampler therefore reads chain c's step-s noise at draw index
// c*draws_per_chain + 1 + s and is bit-identical to the serial reference.
#include <curand_kernel.h>
#include "common.cuh"
#include "unet_ops.h"

namespace pdr {

void philox_launch_geometry(long long numel, long long* threads, unsigned long long* counter_inc) {
  int dev = 0, sms = 148, tpsm = 2048;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&tpsm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
  const long long block = 256, unroll = 4;
  long long grid = (numel + block - 1) / block;
  const long long cap = (long long)sms * (tpsm / block);
  if (grid > cap) grid = cap;
  *threads = grid * block;
  *counter_inc = (unsigned long long)(((numel - 1) / (block * grid * unroll) + 1) * 4);
}

__device__ __forceinline__ float torch_randn_element(unsigned long long seed,
                                                     unsigned long long offset, long long i,
                                                     long long T) {
  const long long q = i / T, idx = i - q * T;
  const long long it = q >> 2;
  const int ii = (int)(q & 3);
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)idx, offset + 4ull * (unsigned long long)it, &st);
  const float4 r = curand_normal4(&st);
  return ii == 0 ? r.x : (ii == 1 ? r.y : (ii == 2 ? r.z : r.w));
}

__global__ void randn_like_torch_kernel(float* out, long long numel, unsigned long long seed,
                                        unsigned long long offset, long long T) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < numel) out[i] = torch_randn_element(seed, offset, i, T);
}

int randn_like_torch_launch(float* out, long long numel, unsigned long long seed,
                            unsigned long long offset, cudaStream_t stream) {
  long long T;
  unsigned long long inc;
  philox_launch_geometry(numel, &T, &inc);
  randn_like_torch_kernel<<<cdiv(numel, 256), 256, 0, stream>>>(out, numel, seed, offset, T);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// y = A(data_transform(x_orig)) = (2*x - 1) * mask ;  x_T = randn        (diffusion.py:478-499)
__global__ void ddnm_prepare_kernel(const float* __restrict__ sparse,
                                    const float* __restrict__ mask, int V, int C, int HW,
                                    unsigned long long seed, unsigned long long offset_base,
                                    unsigned long long counter_inc,
                                    unsigned long long draws_per_chain, int chain0, long long T,
                                    float* __restrict__ y, float* __restrict__ x) {
  const long long per = (long long)C * HW;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)V * per) return;
  const int v = (int)(i / per);
  const long long e = i - (long long)v * per;
  const int p = (int)(e % HW);
  const float xo = __fsub_rn(__fmul_rn(2.0f, sparse[i]), 1.0f);
  y[i] = __fmul_rn(xo, mask[(size_t)v * HW + p]);
  const unsigned long long draw = (unsigned long long)(chain0 + v) * draws_per_chain;
  x[i] = torch_randn_element(seed, offset_base + draw * counter_inc, e, T);
}

int ddnm_prepare_launch(const float* sparse, const float* mask, int V, int C, int H, int W,
                        unsigned long long seed, unsigned long long offset_base,
                        unsigned long long draws_per_chain, int chain0, float* y, float* x,
                        cudaStream_t stream) {
  long long T;
  unsigned long long inc;
  philox_launch_geometry((long long)C * H * W, &T, &inc);
  const long long n = (long long)V * C * H * W;
  ddnm_prepare_kernel<<<cdiv(n, 256), 256, 0, stream>>>(sparse, mask, V, C, H * W, seed,
                                                       offset_base, inc, draws_per_chain, chain0,
                                                       T, y, x);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// one reverse step for all chains (diffusion.py:520-552), x updated in place
__global__ void ddnm_step_kernel(float* __restrict__ x, const float* __restrict__ et,
                                 int et_channels, const float* __restrict__ y,
                                 const float* __restrict__ mask, int V, int C, int HW,
                                 DdnmStepCoef k, unsigned long long seed,
                                 unsigned long long offset_base, unsigned long long counter_inc,
                                 unsigned long long draws_per_chain, int chain0, int draw_index,
                                 long long T) {
  const long long per = (long long)C * HW;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)V * per) return;
  const int v = (int)(i / per);
  const long long e = i - (long long)v * per;
  const int c = (int)(e / HW), p = (int)(e % HW);
  const float xt = x[i];
  const float eps = et[((size_t)v * et_channels + c) * HW + p];  // et[:, :3]
  const float mk = mask[(size_t)v * HW + p];
  // Eq. 12: x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()
  const float x0 = __fdiv_rn(__fsub_rn(xt, __fmul_rn(eps, k.sqrt_1m_at)), k.sqrt_at);
  // Eq. 17: x0_t_hat = x0_t - lambda_t * Ap(A(x0_t) - y)
  const float x0_hat =
      __fsub_rn(x0, __fmul_rn(k.lambda_t, __fmul_rn(__fsub_rn(__fmul_rn(x0, mk), y[i]), mk)));
  const unsigned long long draw =
      (unsigned long long)(chain0 + v) * draws_per_chain + (unsigned long long)draw_index;
  const float noise = torch_randn_element(seed, offset_base + draw * counter_inc, e, T);
  // xt_next = at_next.sqrt() * x0_t_hat + gamma_t * (c1 * randn_like(x0_t) + c2 * et)
  x[i] = __fadd_rn(__fmul_rn(k.sqrt_at_next, x0_hat),
                   __fmul_rn(k.gamma_t, __fadd_rn(__fmul_rn(k.c1, noise), __fmul_rn(k.c2, eps))));
}

int ddnm_step_launch(float* x, const float* et, int et_channels, const float* y, const float* mask,
                     int V, int C, int H, int W, DdnmStepCoef coef, unsigned long long seed,
                     unsigned long long offset_base, unsigned long long draws_per_chain,
                     int chain0, int draw_index, cudaStream_t stream) {
  long long T;
  unsigned long long inc;
  philox_launch_geometry((long long)C * H * W, &T, &inc);
  const long long n = (long long)V * C * H * W;
  ddnm_step_kernel<<<cdiv(n, 256), 256, 0, stream>>>(x, et, et_channels, y, mask, V, C, H * W, coef,
                                                    seed, offset_base, inc, draws_per_chain,
                                                    chain0, draw_index, T);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// inverse_data_transform: clamp((x + 1) / 2, 0, 1)
__global__ void ddnm_final_kernel(const float* __restrict__ x, long long n,
                                  float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fminf(fmaxf(__fdiv_rn(__fadd_rn(x[i], 1.0f), 2.0f), 0.0f), 1.0f);
}

int ddnm_final_launch(const float* x, long long n, float* out, cudaStream_t stream) {
  ddnm_final_kernel<<<cdiv(n, 256), 256, 0, stream>>>(x, n, out);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
