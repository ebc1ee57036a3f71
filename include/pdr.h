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
 *   bn: N tile, 128 / 256 / 0 (auto).  C1, C2 % 64 == 0, Cout % 128 == 0. */
int pdr_conv_tc(const void* x1, const void* x2, const void* w, const float* bias,
                const void* residual, void* out, int B, int H, int W, int C1, int C2, int Cout,
                int taps, int bn, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDR_H */
