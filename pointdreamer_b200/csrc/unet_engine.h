// Host interface of the native U-Net runtime and DDNM sampler (unet_engine.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "pdr.h"

namespace pdr {

int unet_create(const PdrUnetConfig* cfg, void** handle);
int unet_destroy(void* handle);
int unet_set_param(void* handle, const char* name, const void* ptr, size_t bytes);
int unet_workspace_bytes(void* handle, int B, size_t* bytes);
int unet_plan(void* handle, int B, void* workspace, size_t bytes);
// out: [B, n_out, S, S] fp32 (first n_out of the model's output channels)
int unet_forward(void* handle, const float* x, const float* t, float* out, int n_out,
                 cudaStream_t stream);
// the public entry: the same, ordered after the previous submission of any engine on this device
int unet_forward_serialized(void* handle, const float* x, const float* t, float* out, int n_out,
                            cudaStream_t stream);
int ddnm_sample(void* handle, const float* sparse, const float* mask, int V, int steps,
                const float* coef_host, const float* t_dev, unsigned long long seed,
                unsigned long long offset_base, unsigned long long draws_per_chain, int chain0,
                float* x, float* y, float* et, float* out, cudaStream_t stream);
int unet_profile_begin(void* handle, int every, int max_forwards);
int unet_profile_end(void* handle, double* ms, double* flops, long long* launches,
                     long long* forwards);
int unet_planned_batch(void* handle);
int unet_image_size(void* handle);
int unet_out_channels(void* handle);

}  // namespace pdr
