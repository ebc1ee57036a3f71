// Experiment: can ONE shared-memory halo tile (10 x 18 pixels x 64 channels, SWIZZLE_128B, K-major)
// feed all nine taps of a 3x3 convolution through UMMA descriptors whose start address is shifted
// by whole 128-byte rows and whose 8-row groups are 1280 B apart (SBO = 10 pixels)?
// D = A * I (B = identity over 64 channels), so D[m][n] must equal H[(m/8+dy)*10 + m%8+dx][n].
// Tries the descriptor's 3-bit "base offset" field = 0 and = (start >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I pointdreamer_b200/csrc -I include \
//        tools/experiments/halo_desc_test.cu -o gpurun_out/halo_desc_test && gpurun_out/halo_desc_test
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
using namespace pdr;

static constexpr int HW = 10, HH = 18, ROWS = HW * HH;  // 180 halo rows of 128 B

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128) halo_test(int off_rows, int use_base_off, float* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __half* H = (__half*)smem;                      // 180 rows * 128 B = 23040 B (pad to 23552)
  __half* Bm = (__half*)(smem + 24576);            // 64 rows * 128 B
  uint64_t* bar = (uint64_t*)(smem + 24576 + 8192);
  uint32_t* tptr = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x;
  // SWIZZLE_128B as TMA writes it: 16-byte chunk index XOR (row & 7), rows 128 B apart
  for (int i = tid; i < ROWS * 64; i += 128) {
    const int r = i / 64, c = i % 64;
    const int chunk = (c / 8) ^ (r & 7);
    H[r * 64 + chunk * 8 + (c % 8)] = __float2half((float)((r * 64 + c) % 2048));
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    const int n = i / 64, k = i % 64;
    const int chunk = (k / 8) ^ (n & 7);
    Bm[n * 64 + chunk * 8 + (k % 8)] = __float2half(n == k ? 1.f : 0.f);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  if (tid < 32) {
    tmem_alloc(tptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(H) + off_rows * 128, b0 = smem_u32(Bm);
    const uint32_t idesc = make_idesc_f16(128, 64);
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t aa = a0 + kk * 32;
      const uint64_t ad = desc_sw128(aa, HW * 128, use_base_off ? (aa >> 7) & 7 : 0);
      const uint64_t bd = desc_sw128(b0 + kk * 32, 1024, 0);
      umma_f16(tmem, ad, bd, idesc, kk != 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + ch * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 64 * 4);
  const int smem = 24576 + 8192 + 64 + 1024;
  cudaFuncSetAttribute(halo_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> h(128 * 64);
  int all_ok[2] = {1, 1};
  for (int ub = 0; ub < 2; ++ub)
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx) {
        halo_test<<<1, 128, smem>>>(dy * HW + dx, ub, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("base_off=%d tap(%d,%d): CUDA error %s\n", ub, dy, dx, cudaGetErrorString(e));
          return 1;
        }
        cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            const int r = (m / 8 + dy) * HW + (m % 8 + dx);
            const float want = (float)((r * 64 + n) % 2048);
            if (h[m * 64 + n] != want) {
              if (first < 0) first = m * 64 + n;
              ++bad;
            }
          }
        if (bad) all_ok[ub] = 0;
        printf("base_off=%s tap(dy=%d,dx=%d) start+%4d B: %s (%d of 8192 wrong", ub ? "(addr>>7)&7" : "0",
               dy, dx, (dy * HW + dx) * 128, bad ? "MISMATCH" : "ok", bad);
        if (bad) printf("; first m=%d n=%d got %.0f", first / 64, first % 64, h[first]);
        printf(")\n");
      }
  printf("RESULT base_off=0: %s ; base_off=(addr>>7)&7: %s\n", all_ok[0] ? "ALL TAPS OK" : "fails",
         all_ok[1] ? "ALL TAPS OK" : "fails");
  return 0;
}
