// TMA-fed tcgen05 implicit-GEMM convolution (3x3 "same" or 1x1, stride 1) for sm_100a.
//
// Replaces the cuDNN fp16 Conv2d/Conv1d calls made by the reference U-Net
// (models/DDNM/guided_diffusion/unet.py:176-222 ResBlock convs, :291-294 attention qkv/proj,
//  fp16 torso per unet.py:619-625 / fp16_util.py:15-22).
//
// GEMM view:  D[M = B*H*W, N = Cout] = A[M, K = taps*Cin] * Wt[N, K]^T  (+bias, +residual)
//   * activations are NHWC fp16; the A operand of every (tap, 64-channel chunk) is ONE 4-D TMA box
//     {64 ch, bw, bh, bb} (bw*bh*bb = 128 output pixels) whose (x, y) coordinates are shifted by
//     the tap offset: TMA zero-fills out-of-bounds elements, which IS the conv's zero padding,
//     so no im2col buffer ever exists;
//   * weights are [Cout][taps*Cin] fp16 (K contiguous), one 2-D TMA box {64, BN} per chunk;
//   * both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes;
//   * accumulators live in TMEM (2 x BN fp32 columns, double buffered) so the epilogue of tile i
//     overlaps the MMAs of tile i+1;  persistent CTAs (one per SM) walk the tile list.
//   * A may come from TWO tensors (channel concat of the decoder's skip connection,
//     unet.py:660-662) so the concat is never materialised.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..7 = epilogue (TMEM -> registers -> +bias (+residual) -> fp16 -> global).  The halo kernel
// has two TMA producers (warp 0: weight tiles, warp 3: halo tiles) and 8 more warps that apply a
// fused GroupNorm to the halo tile in shared memory.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "conv_tc.h"
#include "gn_math.cuh"

namespace pdr {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;  // 64 fp16 = 128 B = one swizzle row
static constexpr int UMMA_K = 16;
static constexpr int NUM_THREADS = 256;

struct ConvArgs {
  int B, H, W;
  int C1, C2;  // channels of the two A sources (C2 may be 0)
  int S1, S2;  // fused 1x1 skip branch: channels of its (up to two) sources, 0 = none.  Its
               // weights are K columns taps*(C1+C2) .. taps*(C1+C2)+S1+S2-1 of the weight matrix
  int Cout;
  int taps;  // 9 (3x3, pad 1) or 1 (1x1)
  int bw, bh, bb;
  int tiles_x, tiles_y, tiles_b, tiles_n;
  const float* bias;        // [Cout] or null
  const __half* residual;   // [B,H,W,Cout] or null
  __half* out;              // [B,H,W,Cout]
  // optional fused GroupNorm statistics: per (m-tile, epilogue warp, 8-channel chunk) sum and
  // sum of squares of the fp16 outputs, [tiles_m][4][Cout/8][2] floats (null = off)
  float* stats_partial;
  float qk_scale;  // != 0: channels with c % 192 < 128 are re-rounded after a multiply by it
  // split-K (1-CTA kernel only): the K loop is cut into `ksplit` equal ranges that run as
  // separate work items; each writes its raw fp32 accumulators to splitk_ws
  // [ksplit][tiles_m*128][Cout] and splitk_finish_kernel applies the epilogue.  1 = off.
  int ksplit;
  float* splitk_ws;
  int halo;  // 1: conv_halo_kernel (8 x 16 pixel tiles, activation maps encoded with the halo box)
  // halo kernel only: GroupNorm32 (+FiLM) + SiLU of the conv INPUT applied in shared memory by the
  // transform warps (the A maps then point at the raw tensor): float4 (ga, gb, fs, fsh) per
  // (image, input channel of cat(A1, A2)); null = the input is already activated
  const float4* gn_coeff;
  int gn_film;
  // plain kernels: the epilogue stages each 128-pixel x 64-channel block of the output tile in
  // shared memory (128-byte swizzle) and writes it with ONE TMA store (cp.async.bulk.tensor,
  // UTMASTG) instead of 16-byte row-strided stores from every thread
  int tma_store;
};

static constexpr int OUT_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB, one TMA-store box

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_STAGE_OFFSET = STAGES * STAGE_BYTES;  // 2 x [128 px][64 ch] fp16
  static constexpr int BAR_OFFSET = OUT_STAGE_OFFSET + 2 * OUT_STAGE_BYTES;
  // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem ptr
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16;
};

// One output tile of the epilogue for the calling warp (TMEM lane quadrant q): TMEM -> registers ->
// +bias (+residual) -> fp16 -> global, plus the optional deterministic GroupNorm partial sums.
// TMA-store state of the epilogue warps of one CTA: staging buffers + how many 64-channel blocks
// have been issued so far (same value in all 128 epilogue threads)
struct OutStage {
  uint8_t* buf = nullptr;        // 2 x OUT_STAGE_BYTES, 1024-byte aligned; null = direct stores
  const CUtensorMap* map = nullptr;
  uint32_t blocks = 0;
};

template <int BN>
__device__ __forceinline__ void epilogue_tile(const ConvArgs& args, uint32_t tmem_acc, int m_lin,
                                              int n_tile, int q, int lane, OutStage* os = nullptr) {
  const int r = q * 32 + lane;  // row of the tile == TMEM lane
  const int pw = r % args.bw;
  const int ph = (r / args.bw) % args.bh;
  const int pb = r / (args.bw * args.bh);
  int m_tile = m_lin;
  const int tx = m_tile % args.tiles_x;
  m_tile /= args.tiles_x;
  const int ty = m_tile % args.tiles_y;
  const int tb = m_tile / args.tiles_y;
  const int x = tx * args.bw + pw, y = ty * args.bh + ph, b = tb * args.bb + pb;
  const int n0 = n_tile * BN;
  const bool valid = (b < args.B) && (y < args.H) && (x < args.W);
  const size_t row_off = (((size_t)b * args.H + y) * args.W + x) * (size_t)args.Cout + n0;
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
  // the residual of chunk ch+1 is requested before chunk ch is processed, so its memory latency
  // hides behind one chunk of TMEM reads / conversions / stores instead of stalling every chunk
  const bool has_res = valid && args.residual != nullptr;
  uint4 rnext[4];
  if (has_res) {
    const uint4* rp = (const uint4*)(args.residual + row_off);
#pragma unroll
    for (int j = 0; j < 4; ++j) rnext[j] = __ldg(rp + j);
  }
#pragma unroll 1
  for (int ch = 0; ch < BN / 32; ++ch) {
    uint32_t v[32];
    float s8[4] = {0.f, 0.f, 0.f, 0.f}, q8[4] = {0.f, 0.f, 0.f, 0.f};
    uint4 rcur[4];
    if (has_res) {
#pragma unroll
      for (int j = 0; j < 4; ++j) rcur[j] = rnext[j];
      if (ch + 1 < BN / 32) {
        const uint4* rp = (const uint4*)(args.residual + row_off + (ch + 1) * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) rnext[j] = __ldg(rp + j);
      }
    }
    tmem_ld_32x32(taddr + ch * 32, v);
    tmem_ld_wait();
    if (valid) {
      const int nb = n0 + ch * 32;
      __align__(16) __half o[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (args.bias) bv = __ldg((const float4*)(args.bias + nb + j));
        o[j + 0] = __float2half_rn(__uint_as_float(v[j + 0]) + bv.x);
        o[j + 1] = __float2half_rn(__uint_as_float(v[j + 1]) + bv.y);
        o[j + 2] = __float2half_rn(__uint_as_float(v[j + 2]) + bv.z);
        o[j + 3] = __float2half_rn(__uint_as_float(v[j + 3]) + bv.w);
      }
      if (args.qk_scale != 0.f && (nb % 192) < 128) {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = __float2half_rn(__half2float(o[j]) * args.qk_scale);
      }
      if (args.residual) {
        // reference adds two fp16 tensors (unet.py:256, :305): fp32 add, one more rounding
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 rv = rcur[j];
          const __half* rh = (const __half*)&rv;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            o[j * 8 + e] =
                __float2half_rn(__half2float(o[j * 8 + e]) + __half2float(rh[e]));
        }
      }
      if (os == nullptr) {
        uint4* op = (uint4*)(args.out + row_off + ch * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) op[j] = ((const uint4*)o)[j];
      } else {
        // this thread's 64 bytes of row r of the staged block: 16-byte chunk c of a row lives at
        // (c ^ (r & 7)) - the layout a SWIZZLE_128B tensor map expects
        uint8_t* dst = os->buf + (os->blocks & 1) * OUT_STAGE_BYTES + r * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *(uint4*)(dst + ((((ch & 1) * 4 + j) ^ (r & 7)) << 4)) = ((const uint4*)o)[j];
      }
      if (args.stats_partial) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float s = 0.f, q2 = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float f = __half2float(o[g * 8 + j]);
            s += f;
            q2 += f * f;
          }
          s8[g] = s;
          q8[g] = q2;
        }
      }
    }
    if (args.stats_partial) {
      // reduce the 8 per-row values (4 chunk sums, 4 chunk sums of squares) over the warp's 32
      // rows with a halving butterfly (9 shuffles, fixed order -> deterministic): after the
      // xor-16/8/4 steps every lane owns ONE of the 8 values, xor-2/1 finish it.
      float w[8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        w[g] = valid ? s8[g] : 0.f;
        w[4 + g] = valid ? q8[g] : 0.f;
      }
      float x4[4], x2[2], x1;
      {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float send = hi ? w[i] : w[i + 4];
          const float keep = hi ? w[i + 4] : w[i];
          x4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
      }
      {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float send = hi ? x4[i] : x4[i + 2];
          const float keep = hi ? x4[i + 2] : x4[i];
          x2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
      }
      {
        const bool hi = lane & 4;
        const float send = hi ? x2[0] : x2[1];
        const float keep = hi ? x2[1] : x2[0];
        x1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      x1 += __shfl_xor_sync(0xffffffffu, x1, 2);
      x1 += __shfl_xor_sync(0xffffffffu, x1, 1);
      if ((lane & 3) == 0) {
        const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        const int g = idx & 3, is_q = idx >> 2;
                    args.stats_partial[(((size_t)m_lin * 4 + q) * (args.Cout / 8) + (n0 + ch * 32) / 8 + g) * 2 +
                           is_q] = x1;
      }
    }
    if (os != nullptr && (ch & 1)) {
      // both 32-channel halves of the block are staged: make the generic-proxy writes visible to
      // the async proxy, gather the four epilogue warps, one thread issues the store
      fence_proxy_async();
      named_bar_sync(1, 128);
      if (q == 0 && lane == 0) {
        tma_store_4d(os->map, os->buf + (os->blocks & 1) * OUT_STAGE_BYTES, n0 + (ch >> 1) * 64,
                     tx * args.bw, ty * args.bh, tb * args.bb);
        bulk_commit_group();
        bulk_wait_group_read<1>();  // the OTHER buffer (issued one block ago) has been read
      }
      ++os->blocks;
      named_bar_sync(1, 128);  // ... so everybody may overwrite it
    }
  }
}

// split-K work item: raw fp32 accumulators of this K range -> workspace (row = TMEM lane)
template <int BN>
__device__ __forceinline__ void epilogue_partial(const ConvArgs& args, uint32_t tmem_acc, int m_lin,
                                                 int n_tile, int ks, int q, int lane) {
  const int tiles_m = args.tiles_b * args.tiles_y * args.tiles_x;
  const size_t row = ((size_t)ks * tiles_m + m_lin) * BLOCK_M + q * 32 + lane;
  float* dst = args.splitk_ws + row * args.Cout + n_tile * BN;
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
  for (int ch = 0; ch < BN / 32; ++ch) {
    uint32_t v[32];
    tmem_ld_32x32(taddr + ch * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *(uint4*)(dst + ch * 32 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
}

// sum of the K ranges (fixed order) + bias (+ residual) -> fp16, the same rounding points as
// epilogue_tile; thread = 8 consecutive output channels of one tile row
__global__ void splitk_finish_kernel(const ConvArgs args) {
  const int tiles_m = args.tiles_b * args.tiles_y * args.tiles_x;
  const int c8n = args.Cout / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)tiles_m * BLOCK_M * c8n) return;
  const int c0 = (int)(i % c8n) * 8;
  const long long row = i / c8n;
  const int r = (int)(row % BLOCK_M);
  int m_tile = (int)(row / BLOCK_M);
  const int pw = r % args.bw, ph = (r / args.bw) % args.bh, pb = r / (args.bw * args.bh);
  const int tx = m_tile % args.tiles_x;
  m_tile /= args.tiles_x;
  const int ty = m_tile % args.tiles_y;
  const int tb = m_tile / args.tiles_y;
  const int x = tx * args.bw + pw, y = ty * args.bh + ph, b = tb * args.bb + pb;
  if (b >= args.B || y >= args.H || x >= args.W) return;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int ks = 0; ks < args.ksplit; ++ks) {
    const float* p = args.splitk_ws + ((size_t)ks * tiles_m * BLOCK_M + row) * args.Cout + c0;
    const float4 a0 = *(const float4*)p, a1 = *(const float4*)(p + 4);
    acc[0] += a0.x, acc[1] += a0.y, acc[2] += a0.z, acc[3] += a0.w;
    acc[4] += a1.x, acc[5] += a1.y, acc[6] += a1.z, acc[7] += a1.w;
  }
  const size_t off = (((size_t)b * args.H + y) * args.W + x) * (size_t)args.Cout + c0;
  __align__(16) __half o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    o[j] = __float2half_rn(acc[j] + (args.bias ? args.bias[c0 + j] : 0.f));
  if (args.residual) {
    const uint4 rv = __ldg((const uint4*)(args.residual + off));
    const __half* rh = (const __half*)&rv;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = __float2half_rn(__half2float(o[j]) + __half2float(rh[j]));
  }
  *(uint4*)(args.out + off) = *(const uint4*)o;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmS1,
               const __grid_constant__ CUtensorMap tmS2, const __grid_constant__ CUtensorMap tmO,
               const ConvArgs args) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16-B aligned by the ABI: realign to 1024 B for SWIZZLE_128B
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

  uint64_t* full_bar = (uint64_t*)(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int Ctot = args.C1 + args.C2;
  const int kchunks_per_tap = Ctot / BLOCK_K;
  const int num_k_main = args.taps * kchunks_per_tap;
  const int num_k = num_k_main + (args.S1 + args.S2) / BLOCK_K;  // + fused 1x1 skip branch
  const int tiles_m = args.tiles_b * args.tiles_y * args.tiles_x;
  const int num_tiles = tiles_m * args.tiles_n;
  const int ksplit = args.ksplit;          // K ranges per tile (1 = the whole K loop)
  const int k_per = num_k / ksplit;        // exact (checked by the launcher)
  const int num_work = num_tiles * ksplit;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    if (args.C2 > 0) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    if (args.tma_store) tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================================================== TMA producer ====
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
        const int tile = work / ksplit;
        const int k_begin = (work - tile * ksplit) * k_per, k_end = k_begin + k_per;
        const int n_tile = tile % args.tiles_n;
        int m_tile = tile / args.tiles_n;
        const int tx = m_tile % args.tiles_x;
        m_tile /= args.tiles_x;
        const int ty = m_tile % args.tiles_y;
        const int tb = m_tile / args.tiles_y;
        const int x0 = tx * args.bw, y0 = ty * args.bh, b0 = tb * args.bb;
        const int n0 = n_tile * BN;
        for (int k = k_begin; k < k_end; ++k) {
          // chunk-major K order (all taps of a 64-channel chunk, then the next chunk): the same
          // order as the halo kernel, so every variant accumulates a given output identically
          const int kc = k / args.taps;
          const int tap = k - kc * args.taps;
          int dy = 0, dx = 0;
          if (args.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          if (k < num_k_main) {
            const int c = kc * BLOCK_K;
            if (c < args.C1)
              tma_load_4d(sa, &tmA1, &full_bar[stage], c, x0 + dx, y0 + dy, b0);
            else
              tma_load_4d(sa, &tmA2, &full_bar[stage], c - args.C1, x0 + dx, y0 + dy, b0);
            tma_load_2d(sb, &tmB, &full_bar[stage], tap * Ctot + c, n0);
          } else {  // 1x1 skip branch: centre tap of the block input, its own weight columns
            const int c = (k - num_k_main) * BLOCK_K;
            if (c < args.S1)
              tma_load_4d(sa, &tmS1, &full_bar[stage], c, x0, y0, b0);
            else
              tma_load_4d(sa, &tmS2, &full_bar[stage], c - args.S1, x0, y0, b0);
            tma_load_2d(sb, &tmB, &full_bar[stage], args.taps * Ctot + c, n0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================================================= MMA issuer ====
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16(BLOCK_M, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int k = 0; k < k_per; ++k) {
          mbar_wait(&full_bar[stage], phase);  // TMA bytes landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            const uint64_t adesc = make_smem_desc_sw128(sa + kk * UMMA_K * 2);
            const uint64_t bdesc = make_smem_desc_sw128(sb + kk * UMMA_K * 2);
            umma_f16(tmem_d, adesc, bdesc, idesc, (k | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ========================================================= epilogue ====
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    OutStage os;
    const bool tma_out = args.tma_store && ksplit == 1;
    if (tma_out) {
      os.buf = smem + L::OUT_STAGE_OFFSET;
      os.map = &tmO;
    }
    for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
      const int tile = work / ksplit;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (ksplit > 1)
        epilogue_partial<BN>(args, tmem_base + (uint32_t)(acc * BN), tile / args.tiles_n,
                             tile % args.tiles_n, work - tile * ksplit, q, lane);
      else
        epilogue_tile<BN>(args, tmem_base + (uint32_t)(acc * BN), tile / args.tiles_n,
                          tile % args.tiles_n, q, lane, tma_out ? &os : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (tma_out && q == 0 && lane == 0) bulk_wait_group<0>();  // stores complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}


// ------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a cluster of two CTAs (one SM pair) computes a 256 x 256 output
// tile.  Each CTA stages its own 128-row A tile and HALF of the 256-row weight tile; the leader's
// tcgen05.mma reads both halves, so shared-memory traffic per FLOP drops by a third and each SM's
// TMEM holds its own 128 x 256 accumulator.  Only the leader CTA issues MMAs; both run TMA
// producers and epilogues.
template <int STAGES>
struct SmemLayout2 {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = 128 * BLOCK_K * 2;  // this CTA's half of the 256-row B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_STAGE_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = OUT_STAGE_OFFSET + 2 * OUT_STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16;
};

template <int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmS1,
                const __grid_constant__ CUtensorMap tmS2, const __grid_constant__ CUtensorMap tmO,
                const ConvArgs args) {
  constexpr int BN = 256;
  using L = SmemLayout2<STAGES>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const int Ctot = args.C1 + args.C2;
  const int kchunks_per_tap = Ctot / BLOCK_K;
  const int num_k_main = args.taps * kchunks_per_tap;
  const int num_k = num_k_main + (args.S1 + args.S2) / BLOCK_K;  // + fused 1x1 skip branch
  const int tiles_m = args.tiles_b * args.tiles_y * args.tiles_x;  // even (checked on the host)
  const int num_pairs = (tiles_m / 2) * args.tiles_n;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    if (args.C2 > 0) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      // leader only: its arrive.expect_tx covers the bytes of BOTH CTAs.  The peer never runs a
      // phase ahead (its slots are released by the leader's multicast commit), so its TMA bytes
      // can only land in the phase the leader is about to arm.
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);  // multicast tcgen05.commit of the leader
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);   // multicast tcgen05.commit
      mbar_init(&tmem_empty[i], 8);  // 4 epilogue warps of each CTA (used in the leader only)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_ptr_smem, 2 * BN);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ============================================ TMA producer (both CTAs) ====
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pair = cluster_id; pair < num_pairs; pair += num_clusters) {
        const int n_tile = pair % args.tiles_n;
        int m_tile = (pair / args.tiles_n) * 2 + (int)rank;
        const int tx = m_tile % args.tiles_x;
        m_tile /= args.tiles_x;
        const int ty = m_tile % args.tiles_y;
        const int tb = m_tile / args.tiles_y;
        const int x0 = tx * args.bw, y0 = ty * args.bh, b0 = tb * args.bb;
        const int n0 = n_tile * BN + (int)rank * 128;  // this CTA's half of the weight rows
        for (int k = 0; k < num_k; ++k) {
          // chunk-major K order (all taps of a 64-channel chunk, then the next chunk): the same
          // order as the halo kernel, so every variant accumulates a given output identically
          const int kc = k / args.taps;
          const int tap = k - kc * args.taps;
          int dy = 0, dx = 0;
          if (args.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);  // bytes of both CTAs
          if (k < num_k_main) {
            const int c = kc * BLOCK_K;
            if (c < args.C1)
              tma2_load_4d(sa, &tmA1, &full_bar[stage], c, x0 + dx, y0 + dy, b0);
            else
              tma2_load_4d(sa, &tmA2, &full_bar[stage], c - args.C1, x0 + dx, y0 + dy, b0);
            tma2_load_2d(sb, &tmB, &full_bar[stage], tap * Ctot + c, n0);
          } else {  // 1x1 skip branch
            const int c = (k - num_k_main) * BLOCK_K;
            if (c < args.S1)
              tma2_load_4d(sa, &tmS1, &full_bar[stage], c, x0, y0, b0);
            else
              tma2_load_4d(sa, &tmS2, &full_bar[stage], c - args.S1, x0, y0, b0);
            tma2_load_2d(sb, &tmB, &full_bar[stage], args.taps * Ctot + c, n0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && leader) {
    // ================================================ MMA issuer (leader) ====
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16(2 * BLOCK_M, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int pair = cluster_id; pair < num_pairs; pair += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int k = 0; k < num_k; ++k) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            const uint64_t adesc = make_smem_desc_sw128(sa + kk * UMMA_K * 2);
            const uint64_t bdesc = make_smem_desc_sw128(sb + kk * UMMA_K * 2);
            umma2_f16(tmem_d, adesc, bdesc, idesc, (k | kk) != 0 ? 1u : 0u);
          }
          umma2_commit_multicast(&empty_bar[stage]);  // frees the slot in both CTAs
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma2_commit_multicast(&tmem_full[acc]);  // both CTAs' epilogues
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ============================================== epilogue (both CTAs) ====
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    OutStage os;
    if (args.tma_store) {
      os.buf = smem + L::OUT_STAGE_OFFSET;
      os.map = &tmO;
    }
    for (int pair = cluster_id; pair < num_pairs; pair += num_clusters) {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN>(args, tmem_base + (uint32_t)(acc * BN),
                        (pair / args.tiles_n) * 2 + (int)rank, pair % args.tiles_n, q, lane,
                        args.tma_store ? &os : nullptr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);  // the leader's barrier
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (args.tma_store && q == 0 && lane == 0) bulk_wait_group<0>();  // stores complete before exit
  }

  tc_fence_before();
  cluster_sync_all();  // nobody may exit (or free TMEM) while the peer still uses its smem/TMEM
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 2 * BN);
  }
}


// ------------------------------------------------------------------------------------------
// Halo variant for 3x3 convolutions on maps of at least 16 x 8: the A operand of ALL NINE taps of
// a 64-channel chunk comes from ONE shared-memory tile.  The output tile is 8 pixels wide and 16
// tall; TMA loads its 10 x 18 halo (zero filled outside the image = the conv padding) as 180 rows
// of 128 B (SWIZZLE_128B).  An 8-pixel output row is one 8-row core-matrix group, so tap (dy,dx) is
// the SAME tile seen through a descriptor whose start address is shifted by (dy*10+dx) rows and
// whose groups are 10 rows (1280 B) apart - the swizzle is a function of the absolute shared
// memory address, so whole-row shifts keep it consistent (tools/experiments/halo_desc_test.cu).
// Per k-step a CTA now fills 23 KB / 9 + its weight tile instead of 16 KB + its weight tile:
// L2 -> shared memory traffic drops by 40 % (2-CTA) to 46 % (1-CTA, BN = 128), which lifts the
// 1-CTA tiles off the shared-memory fill limit.  Two rings: A (one slot per chunk) and B (one
// slot per tap).  TWO = cta_group::2 pair (BN = 256, each CTA stages half of the weight tile).
static constexpr int HALO_W = 10, HALO_H = 18;
static constexpr int HALO_BYTES = HALO_W * HALO_H * BLOCK_K * 2;            // 23040
static constexpr int HALO_STAGE = (HALO_BYTES + 1023) / 1024 * 1024;        // 23552

__device__ __forceinline__ uint64_t make_smem_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <bool TWO, int BN, int AST, int BST>
struct SmemLayoutHalo {
  static constexpr int B_ROWS = TWO ? 128 : BN;  // weight rows staged by this CTA
  static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int B_OFFSET = AST * HALO_STAGE;
  static constexpr int BAR_OFFSET = B_OFFSET + BST * B_BYTES;
  // a_full[AST], a_empty[AST], a_ready[AST], b_full[BST], b_empty[BST], tmem_full[2],
  // tmem_empty[2], tmem ptr
  static constexpr int TOTAL = BAR_OFFSET + (3 * AST + 2 * BST + 4) * 8 + 16;
};

static constexpr int HALO_TWARPS = 8;                         // transform warps (fused GroupNorm)
static constexpr int HALO_THREADS = 256 + 32 * HALO_TWARPS;  // 8 warps as in the plain kernels + them

template <bool TWO, int BN, int AST, int BST>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmS1,
                 const __grid_constant__ CUtensorMap tmS2, const ConvArgs args) {
  using L = SmemLayoutHalo<TWO, BN, AST, BST>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = (uint64_t*)(smem + L::BAR_OFFSET);
  uint64_t* a_empty = a_full + AST;
  uint64_t* a_ready = a_empty + AST;  // fused GroupNorm: the transform warps' "tile activated"
  uint64_t* b_full = a_ready + AST;
  uint64_t* b_empty = b_full + BST;
  uint64_t* tmem_full = b_empty + BST;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  const int Ctot = args.C1 + args.C2;
  const int chunks_main = Ctot / BLOCK_K;
  const int chunks_all = chunks_main + (args.S1 + args.S2) / BLOCK_K;  // + fused 1x1 skip branch
  const int tiles_m = args.tiles_b * args.tiles_y * args.tiles_x;
  // work items: 1-CTA = (m tile, n tile); 2-CTA = (pair of m tiles, n tile) per cluster
  const int num_work = (TWO ? tiles_m / 2 : tiles_m) * args.tiles_n;
  const int worker = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_workers = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr uint32_t TXMUL = TWO ? 2u : 1u;  // the leader's barrier counts the bytes of both CTAs
  // fused GroupNorm: every CTA's halo lands on its OWN a_full (its transform warps wait there), the
  // MMA issuer waits on a_ready, which the transform warps of both CTAs of a pair arrive on
  const bool fused = args.gn_coeff != nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    if (args.C2 > 0) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < AST; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&a_ready[i], TWO ? 2 * HALO_TWARPS : HALO_TWARPS);  // one arrive per transform warp
    }
    for (int i = 0; i < BST; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], TWO ? 8 : 4);  // epilogue warps (of both CTAs in 2-CTA mode)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (TWO) {
      tmem_alloc2(tmem_ptr_smem, 2 * BN);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_ptr_smem, 2 * BN);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (TWO) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ======================================= TMA producer: weight tiles (B) ====
    // The two rings have their own producers: were the halo (A) loads issued by this loop too,
    // between the tap loads, the halo of chunk c+1 could only be requested once every tap of
    // chunk c had a free slot, i.e. ONE chunk ahead of the MMAs whatever the depth of the A
    // ring - too late when the transform warps still have to activate the tile (fused GroupNorm).
    if (elect_one_sync()) {
      int bs = 0;
      uint32_t bph = 0;
      for (int work = worker; work < num_work; work += num_workers) {
        const int n_tile = work % args.tiles_n;
        const int n0 = n_tile * BN + (TWO ? (int)rank * 128 : 0);
        for (int ch = 0; ch < chunks_all; ++ch) {
          const bool main_chunk = ch < chunks_main;
          const int ntaps = main_chunk ? 9 : 1;
          for (int t = 0; t < ntaps; ++t) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            uint8_t* sb = smem + L::B_OFFSET + bs * L::B_BYTES;
            if (leader) mbar_expect_tx(&b_full[bs], TXMUL * L::B_BYTES);
            const int kcoord = main_chunk ? t * Ctot + ch * BLOCK_K
                                          : 9 * Ctot + (ch - chunks_main) * BLOCK_K;
            if (TWO)
              tma2_load_2d(sb, &tmB, &b_full[bs], kcoord, n0);
            else
              tma_load_2d(sb, &tmB, &b_full[bs], kcoord, n0);
            if (++bs == BST) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ========================================= TMA producer: halo tiles (A) ====
    if (elect_one_sync()) {
      int as = 0;
      uint32_t aph = 0;
      for (int work = worker; work < num_work; work += num_workers) {
        int m_tile = TWO ? (work / args.tiles_n) * 2 + (int)rank : work / args.tiles_n;
        const int tx = m_tile % args.tiles_x;
        m_tile /= args.tiles_x;
        const int ty = m_tile % args.tiles_y;
        const int tb = m_tile / args.tiles_y;
        const int x0 = tx * 8 - 1, y0 = ty * 16 - 1;  // halo origin (may be -1: zero fill)
        for (int ch = 0; ch < chunks_all; ++ch) {
          const bool main_chunk = ch < chunks_main;
          mbar_wait(&a_empty[as], aph ^ 1);
          uint8_t* sa = smem + as * HALO_STAGE;
          if (fused)
            mbar_expect_tx(&a_full[as], HALO_BYTES);
          else if (leader)
            mbar_expect_tx(&a_full[as], TXMUL * HALO_BYTES);
          int c = (main_chunk ? ch : ch - chunks_main) * BLOCK_K;
          const CUtensorMap* tm;
          if (main_chunk) {
            tm = c < args.C1 ? &tmA1 : &tmA2;
            if (c >= args.C1) c -= args.C1;
          } else {
            tm = c < args.S1 ? &tmS1 : &tmS2;
            if (c >= args.S1) c -= args.S1;
          }
          if (TWO && !fused)
            tma2_load_4d(sa, tm, &a_full[as], c, x0, y0, tb);
          else
            tma_load_4d(sa, tm, &a_full[as], c, x0, y0, tb);
          if (++as == AST) {
            as = 0;
            aph ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && leader) {
    // ======================================================= MMA issuer ====
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_f16(TWO ? 2 * BLOCK_M : BLOCK_M, BN);
      int as = 0, bs = 0, acc = 0;
      uint32_t aph = 0, bph = 0, acc_phase = 0;
      for (int work = worker; work < num_work; work += num_workers) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        uint32_t first = 1;
        for (int ch = 0; ch < chunks_all; ++ch) {
          mbar_wait(fused ? &a_ready[as] : &a_full[as], aph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + as * HALO_STAGE);
          const int ntaps = ch < chunks_main ? 9 : 1;
          for (int t = 0; t < ntaps; ++t) {
            const int tap = ntaps == 9 ? t : 4;  // the skip branch reads the centre tap
            const uint32_t a_tap = sa + (uint32_t)((tap / 3) * HALO_W + tap % 3) * 128u;
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const uint32_t sb = smem_u32(smem + L::B_OFFSET + bs * L::B_BYTES);
#pragma unroll
            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
              const uint64_t adesc = make_smem_desc_sw128_sbo(a_tap + kk * UMMA_K * 2, HALO_W * 128);
              const uint64_t bdesc = make_smem_desc_sw128(sb + kk * UMMA_K * 2);
              if (TWO)
                umma2_f16(tmem_d, adesc, bdesc, idesc, first ? 0u : 1u);
              else
                umma_f16(tmem_d, adesc, bdesc, idesc, first ? 0u : 1u);
              first = 0;
            }
            if (TWO) umma2_commit_multicast(&b_empty[bs]); else umma_commit(&b_empty[bs]);
            if (++bs == BST) {
              bs = 0;
              bph ^= 1;
            }
          }
          if (TWO) umma2_commit_multicast(&a_empty[as]); else umma_commit(&a_empty[as]);
          if (++as == AST) {
            as = 0;
            aph ^= 1;
          }
        }
        if (TWO) umma2_commit_multicast(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // ==================================== transform warps (fused GroupNorm) ====
    // GroupNorm32 (+FiLM) + SiLU of the raw halo tile, in place, once per chunk: thread = fixed
    // 8-channel group (its 8 x (ga, gb, fs, fsh) live in registers for the chunk), rows strided
    // by 16; pixels outside the image stay zero (the conv pads the ACTIVATED tensor).
    if (fused) {
      constexpr int RSTEP = 4 * HALO_TWARPS;                      // rows between a thread's rows
      constexpr int NROWS = (HALO_W * HALO_H + RSTEP - 1) / RSTEP;
      static_assert(RSTEP % 8 == 0, "the swizzle term (row & 7) must be constant per thread");
      const int tt = threadIdx.x - 256;          // 0 .. 32*HALO_TWARPS-1
      const int lc = tt & 7, row0 = tt >> 3;     // logical 16-byte chunk, first halo row
      const uint32_t toff = (uint32_t)row0 * 128u + (uint32_t)((lc ^ (row0 & 7)) << 4);
      const bool film = args.gn_film != 0;
      int as = 0;
      uint32_t aph = 0;
      for (int work = worker; work < num_work; work += num_workers) {
        int m_tile = TWO ? (work / args.tiles_n) * 2 + (int)rank : work / args.tiles_n;
        const int tx = m_tile % args.tiles_x;
        m_tile /= args.tiles_x;
        const int ty = m_tile % args.tiles_y;
        const int tb = m_tile / args.tiles_y;
        const int x0 = tx * 8 - 1, y0 = ty * 16 - 1;
        uint32_t vmask = 0;  // which of this thread's rows are pixels inside the image
#pragma unroll
        for (int i = 0; i < NROWS; ++i) {
          const int r = row0 + i * RSTEP;
          const int hy = r / HALO_W, hx = r - hy * HALO_W;
          const int gy = y0 + hy, gx = x0 + hx;
          if (r < HALO_W * HALO_H && gy >= 0 && gy < args.H && gx >= 0 && gx < args.W)
            vmask |= 1u << i;
        }
        for (int ch = 0; ch < chunks_all; ++ch) {
          float2 ga[4], gb[4], fsh[4];
          __half2 fs[4];
          if (ch < chunks_main) {  // constants of this thread's 8 channels (L1-resident table)
            const float4* cp = args.gn_coeff + (size_t)tb * Ctot + ch * BLOCK_K + lc * 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 c0 = __ldg(cp + 2 * j), c1 = __ldg(cp + 2 * j + 1);
              ga[j] = make_float2(c0.x, c1.x);
              gb[j] = make_float2(c0.y, c1.y);
              fs[j] = __floats2half2_rn(c0.z, c1.z);  // fp16 values: exact
              fsh[j] = make_float2(c0.w, c1.w);
            }
          }
          mbar_wait(&a_full[as], aph);
          if (ch < chunks_main) {
            // branch-free over the thread's rows: all loads first, 24 independent half2 chains,
            // predicated stores (rows outside the image keep TMA's zeros; a row index past the
            // tile only happens in the last step and stays inside the padded stage)
            uint8_t* base = smem + as * HALO_STAGE + toff;
            uint4 v[NROWS];
#pragma unroll
            for (int i = 0; i < NROWS; ++i) v[i] = *(const uint4*)(base + i * RSTEP * 128);
#pragma unroll
            for (int i = 0; i < NROWS; ++i) {
              __half2* h = (__half2*)&v[i];
              if (film) {
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = gn_apply_two<true>(h[j], ga[j], gb[j], fs[j], fsh[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = gn_apply_two<false>(h[j], ga[j], gb[j], fs[j], fsh[j]);
              }
            }
#pragma unroll
            for (int i = 0; i < NROWS; ++i)
              if ((vmask >> i) & 1u) *(uint4*)(base + i * RSTEP * 128) = v[i];
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's reads
          __syncwarp();
          if (lane == 0) {
            if (TWO) mbar_arrive_cluster(&a_ready[as], 0); else mbar_arrive(&a_ready[as]);
          }
          if (++as == AST) {
            as = 0;
            aph ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ========================================================= epilogue ====
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int work = worker; work < num_work; work += num_workers) {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m_lin = TWO ? (work / args.tiles_n) * 2 + (int)rank : work / args.tiles_n;
      epilogue_tile<BN>(args, tmem_base + (uint32_t)(acc * BN), m_lin, work % args.tiles_n, q, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (TWO) mbar_arrive_cluster(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  if (TWO) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (TWO) tmem_dealloc2(tmem_base, 2 * BN); else tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------ host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  // resolved through the runtime so the library has no link-time dependency on libcuda
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

void conv_tc_pick_box(int B, int H, int W, int* bw, int* bh, int* bb) {
  int w = W < 16 ? W : 16;
  int h = 128 / w;
  if (h > H) h = H;
  int b = 128 / (w * h);
  *bw = w;
  *bh = h;
  *bb = b;
  (void)B;
}

bool conv_tc_halo_ok(int H, int W, int taps) {
  static const bool disabled = getenv("PDR_NO_HALO") != nullptr;  // A/B switch (profiles/README.md)
  return !disabled && taps == 9 && W % 8 == 0 && H % 16 == 0;
}

int conv_tc_make_act_map(ConvTensorMap* out, const void* ptr, int B, int H, int W, int C,
                         int halo) {
  EncodeTiledFn enc = get_encode_fn();
  PDR_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  PDR_CHECK_ARG(C % BLOCK_K == 0, "activation channels (%d) must be a multiple of 64", C);
  PDR_CHECK_ARG(((uintptr_t)ptr & 15) == 0, "activation pointer must be 16-byte aligned");
  int bw, bh, bb;
  conv_tc_pick_box(B, H, W, &bw, &bh, &bb);
  if (halo) {
    PDR_CHECK_ARG(W % 8 == 0 && H % 16 == 0, "halo tiles need W %% 8 == 0 and H %% 16 == 0");
    bw = HALO_W, bh = HALO_H, bb = 1;  // the 10 x 18 halo of an 8 x 16 output tile
  }
  PDR_CHECK_ARG(halo || (bw * bh * bb == 128 && W % bw == 0 && H % bh == 0),
                "unsupported spatial size %dx%d for the 128-pixel tile", H, W);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)ptr, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PDR_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(act) failed: CUresult %d", (int)r);
  return 0;
}

int conv_tc_make_weight_map(ConvTensorMap* out, const void* ptr, int Cout, int K, int BN) {
  EncodeTiledFn enc = get_encode_fn();
  PDR_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  PDR_CHECK_ARG(K % BLOCK_K == 0 && Cout % BN == 0, "weight shape [%d,%d] not tileable by %d",
                Cout, K, BN);
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PDR_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight) failed: CUresult %d", (int)r);
  return 0;
}

int conv_tc_make_map_2d(ConvTensorMap* out, const void* ptr, unsigned long long inner,
                        unsigned long long outer, int box_inner, int box_outer) {
  EncodeTiledFn enc = get_encode_fn();
  PDR_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  PDR_CHECK_ARG(box_inner * 2 == 128 && box_outer >= 1 && box_outer <= 256 && inner % 8 == 0,
                "2-D map: the box must be 128 bytes wide and at most 256 rows");
  PDR_CHECK_ARG(((uintptr_t)ptr & 15) == 0, "2-D map: pointer must be 16-byte aligned");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)inner * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PDR_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: CUresult %d", (int)r);
  return 0;
}

int conv_tc_stats_rows_per_image(int H, int W) {
  int bw, bh, bb;
  conv_tc_pick_box(1, H, W, &bw, &bh, &bb);
  if (bb != 1 || bw * bh != 128 || W % bw || H % bh) return 0;
  return (H / bh) * (W / bw) * 4;
}

int conv_tc_pick_bn(int B, int H, int W, int Cout, int taps) {
  // Returns the N tile with the lowest modelled time: 64 / 128 / 256 (1-CTA) or 512 = the 2-CTA
  // kernel (256-wide tile per SM pair).  Time = waves over the 148 SMs (74 pairs) x cost of one
  // k-step of the tile; K is common to all candidates.  k-step costs (us), fitted to the measured
  // layer tables in profiles/README.md:
  //  * plain kernels: max(MMA, TMA) with MMA = bn * 1.078 ns (15.2 TFLOP/s per SM) and TMA =
  //    (128 + bn) * 128 B at ~150 GB/s per SM, the shared-memory fill rate one CTA sustains from
  //    L2 (bn = 64 / 128 / 256 reach 43 / 65 / 80 % of the tensor peak); SM pair 0.285;
  //  * halo kernels (3x3, maps >= 16 x 8): the activation tile is filled once per nine k-steps, so
  //    every tile is MMA-bound; measured efficiencies 53 / 80 / 88 % (1-CTA) and 96 % (pair).
  int bw, bh, bb;
  conv_tc_pick_box(B, H, W, &bw, &bh, &bb);
  const bool halo = conv_tc_halo_ok(H, W, taps);
  if (halo) bb = 1, bw = 8, bh = 16;
  const long long tiles_m = (long long)cdiv(B, bb) * (H / bh) * (W / bw);
  const int sms = num_sms();
  auto kstep = [halo](int bn) {
    if (halo) return bn == 64 ? 0.130 : (bn == 128 ? 0.1725 : 0.3136);
    const double mma = bn * 1.078e-3, tma = (128 + bn) * 0.853e-3;
    return mma > tma ? mma : tma;
  };
  int best = 0;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 64; bn >>= 1) {
    if (Cout % bn != 0) continue;
    const long long tiles = tiles_m * (Cout / bn);
    const double cost = (double)cdiv(tiles, sms) * kstep(bn);
    if (cost < best_cost * 0.98) {  // prefer the larger tile on a tie (less L2 traffic)
      best_cost = cost;
      best = bn;
    }
  }
  if (Cout % 256 == 0 && tiles_m % 2 == 0) {
    const long long pairs = (tiles_m / 2) * (Cout / 256);
    const double cost = (double)cdiv(pairs, sms / 2) * (halo ? 0.2875 : 0.285);
    if (cost < best_cost * 0.98) best = 512;
  }
  return best;
}

int conv_tc_pick_split(int B, int H, int W, int Cout, int num_k, int* bn) {
  // Layers with fewer tiles than half the SMs (the 8x8 maps) are bound by how fast ONE CTA can
  // fill shared memory for its whole K loop; cutting K into ranges spreads that over idle SMs.
  // Same cost model as conv_tc_pick_bn plus ~4 us for the finishing pass.  Returns ksplit and
  // may raise *bn (a larger tile re-reads less of the activations).
  if (*bn == 512) return 1;
  int bw, bh, bb;
  conv_tc_pick_box(B, H, W, &bw, &bh, &bb);
  const long long tiles_m = (long long)cdiv(B, bb) * (H / bh) * (W / bw);
  const int sms = num_sms();
  auto kstep = [](int n) {
    const double mma = n * 1.078e-3, tma = (128 + n) * 0.853e-3;
    return mma > tma ? mma : tma;
  };
  const long long base_tiles = tiles_m * (Cout / *bn);
  if (base_tiles * 2 > sms) return 1;
  const double base = (double)cdiv(base_tiles, sms) * num_k * kstep(*bn);
  double best = base * 0.8;  // must win clearly
  int best_ks = 1, best_bn = *bn;
  for (int n = 64; n <= 128; n <<= 1) {
    if (Cout % n != 0) continue;
    const long long tiles = tiles_m * (Cout / n);
    for (int ks = 2; ks <= 4; ++ks) {
      if (num_k % ks != 0 || num_k / ks < 16 || tiles * ks > sms) continue;
      const double cost = (double)(num_k / ks) * kstep(n) + 4.0;
      if (cost < best) best = cost, best_ks = ks, best_bn = n;
    }
  }
  *bn = best_bn;
  return best_ks;
}

size_t conv_tc_split_workspace_bytes(int B, int H, int W, int Cout, int ksplit) {
  int bw, bh, bb;
  conv_tc_pick_box(B, H, W, &bw, &bh, &bb);
  const size_t tiles_m = (size_t)cdiv(B, bb) * (H / bh) * (W / bw);
  return ksplit <= 1 ? 0 : (size_t)ksplit * tiles_m * BLOCK_M * Cout * sizeof(float);
}

template <int BN, int STAGES>
static int launch_impl(const ConvTensorMap* a1, const ConvTensorMap* a2, const ConvTensorMap* w,
                       const ConvTensorMap* s1, const ConvTensorMap* s2, const ConvTensorMap* o,
                       const ConvArgs& args, cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  constexpr int smem_bytes = L::TOTAL + 1024;  // +1024 for the manual realignment
  static bool configured = false;
  if (!configured) {
    PDR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  const int tiles = args.tiles_b * args.tiles_y * args.tiles_x * args.tiles_n * args.ksplit;
  int grid = tiles < num_sms() ? tiles : num_sms();
  conv_tc_kernel<BN, STAGES><<<grid, NUM_THREADS, smem_bytes, stream>>>(
      *(const CUtensorMap*)a1, *(const CUtensorMap*)(a2 ? a2 : a1), *(const CUtensorMap*)w,
      *(const CUtensorMap*)(s1 ? s1 : a1), *(const CUtensorMap*)(s2 ? s2 : (s1 ? s1 : a1)),
      *(const CUtensorMap*)(o ? o : a1), args);
  PDR_COUNT_LAUNCH();
  if (args.ksplit > 1) {
    const long long n = (long long)args.tiles_b * args.tiles_y * args.tiles_x * BLOCK_M *
                        (args.Cout / 8);
    splitk_finish_kernel<<<cdiv(n, 256), 256, 0, stream>>>(args);
    PDR_COUNT_LAUNCH();
  }
  PDR_LAUNCH_CHECK();
  return 0;
}

template <int STAGES>
static int launch_impl2(const ConvTensorMap* a1, const ConvTensorMap* a2, const ConvTensorMap* w,
                        const ConvTensorMap* s1, const ConvTensorMap* s2, const ConvTensorMap* o,
                        const ConvArgs& args, cudaStream_t stream) {
  using L = SmemLayout2<STAGES>;
  constexpr int smem_bytes = L::TOTAL + 1024;
  static bool configured = false;
  if (!configured) {
    PDR_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<STAGES>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  const int pairs = (args.tiles_b * args.tiles_y * args.tiles_x / 2) * args.tiles_n;
  int clusters = pairs < num_sms() / 2 ? pairs : num_sms() / 2;
  conv_tc2_kernel<STAGES><<<2 * clusters, NUM_THREADS, smem_bytes, stream>>>(
      *(const CUtensorMap*)a1, *(const CUtensorMap*)(a2 ? a2 : a1), *(const CUtensorMap*)w,
      *(const CUtensorMap*)(s1 ? s1 : a1), *(const CUtensorMap*)(s2 ? s2 : (s1 ? s1 : a1)),
      *(const CUtensorMap*)(o ? o : a1), args);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

template <bool TWO, int BN, int AST, int BST>
static int launch_halo(const ConvTensorMap* a1, const ConvTensorMap* a2, const ConvTensorMap* w,
                       const ConvTensorMap* s1, const ConvTensorMap* s2, const ConvArgs& args,
                       cudaStream_t stream) {
  using L = SmemLayoutHalo<TWO, BN, AST, BST>;
  constexpr int smem_bytes = L::TOTAL + 1024;
  static_assert(smem_bytes <= 227 * 1024, "halo kernel shared memory budget");
  auto kern = conv_halo_kernel<TWO, BN, AST, BST>;
  static bool configured = false;
  if (!configured) {
    PDR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    if (TWO) PDR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 0));
    configured = true;
  }
  const int tiles_m = args.tiles_b * args.tiles_y * args.tiles_x;
  const int work = (TWO ? tiles_m / 2 : tiles_m) * args.tiles_n;
  const int units = TWO ? num_sms() / 2 : num_sms();
  const int workers = work < units ? work : units;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(TWO ? 2 * workers : workers);
  cfg.blockDim = dim3(HALO_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TWO ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PDR_CUDA(cudaLaunchKernelEx(&cfg, kern, *(const CUtensorMap*)a1,
                              *(const CUtensorMap*)(a2 ? a2 : a1), *(const CUtensorMap*)w,
                              *(const CUtensorMap*)(s1 ? s1 : a1),
                              *(const CUtensorMap*)(s2 ? s2 : (s1 ? s1 : a1)), args));
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

// BN == 512 selects the 2-CTA kernel (256-wide N tile per SM pair; the weight map must have been
// encoded with a 128-row box)
int conv_tc_launch(const ConvTensorMap* a1, const ConvTensorMap* a2, const ConvTensorMap* w,
                   int BN, int B, int H, int W, int C1, int C2, int Cout, int taps,
                   const float* bias, const __half* residual, __half* out, float* stats_partial,
                   cudaStream_t stream, float qk_scale, const ConvTensorMap* s1,
                   const ConvTensorMap* s2, int S1, int S2, int ksplit, float* splitk_ws,
                   int halo, const float4* gn_coeff, int gn_film, const ConvTensorMap* omap) {
  PDR_CHECK_ARG(gn_coeff == nullptr || halo, "a fused GroupNorm needs the halo kernel");
  static const bool no_tma_store = getenv("PDR_NO_TMA_STORE") != nullptr;  // A/B switch
  PDR_CHECK_ARG(!halo || (conv_tc_halo_ok(H, W, taps) && ksplit == 1),
                "halo mode needs a 3x3 conv on a map with W %% 8 == 0, H %% 16 == 0 and no split-K");
  PDR_CHECK_ARG(ksplit >= 1 && (ksplit == 1 || (splitk_ws != nullptr && BN != 512 &&
                                                stats_partial == nullptr && qk_scale == 0.f)),
                "split-K needs a workspace, a 1-CTA tile, and no fused statistics / qk scaling");
  PDR_CHECK_ARG((taps * ((C1 + C2) / BLOCK_K) + (S1 + S2) / BLOCK_K) % ksplit == 0,
                "split-K: %d does not divide the number of K steps", ksplit);
  PDR_CHECK_ARG(S1 >= 0 && S2 >= 0 && S1 % BLOCK_K == 0 && S2 % BLOCK_K == 0 &&
                    (S1 == 0 || s1 != nullptr) && (S2 == 0 || (s2 != nullptr && S1 > 0)),
                "fused skip branch: channels must be multiples of 64 with their tensor maps");
  PDR_CHECK_ARG(taps == 9 || taps == 1, "taps must be 9 or 1 (got %d)", taps);
  PDR_CHECK_ARG(C1 > 0 && C1 % BLOCK_K == 0 && C2 % BLOCK_K == 0, "C1/C2 must be multiples of 64");
  PDR_CHECK_ARG(BN == 64 || BN == 128 || BN == 256 || BN == 512, "BN must be 64, 128, 256 or 512");
  const bool two_cta = BN == 512;
  if (two_cta) BN = 256;
  PDR_CHECK_ARG(Cout % BN == 0, "Cout (%d) must be a multiple of BN (%d)", Cout, BN);
  PDR_CHECK_ARG(C2 == 0 || a2 != nullptr, "second A tensor map missing");
  ConvArgs args;
  args.B = B;
  args.H = H;
  args.W = W;
  args.C1 = C1;
  args.C2 = C2;
  args.S1 = S1;
  args.S2 = S2;
  args.Cout = Cout;
  args.taps = taps;
  conv_tc_pick_box(B, H, W, &args.bw, &args.bh, &args.bb);
  if (halo) args.bw = 8, args.bh = 16, args.bb = 1;
  args.halo = halo;
  args.gn_coeff = gn_coeff;
  args.gn_film = gn_film;
  PDR_CHECK_ARG(args.bw * args.bh * args.bb == 128 && W % args.bw == 0 && H % args.bh == 0,
                "unsupported spatial size %dx%d", H, W);
  args.tiles_x = W / args.bw;
  args.tiles_y = H / args.bh;
  args.tiles_b = cdiv(B, args.bb);
  args.tiles_n = Cout / BN;
  args.bias = bias;
  args.residual = residual;
  args.out = out;
  args.stats_partial = stats_partial;
  args.qk_scale = qk_scale;
  args.ksplit = ksplit;
  args.splitk_ws = splitk_ws;
  args.tma_store = (omap != nullptr && !halo && ksplit == 1 && !no_tma_store) ? 1 : 0;
  PDR_CHECK_ARG(qk_scale == 0.f || (Cout % 192 == 0 && !residual && !stats_partial),
                "qk_scale is for qkv projections (Cout %% 192 == 0, no residual, no statistics)");
  PDR_CHECK_ARG(!stats_partial || (args.bb == 1 && Cout % 32 == 0),
                "fused GroupNorm statistics need one image per tile (H*W >= 128)");
  if (two_cta) {
    PDR_CHECK_ARG((args.tiles_b * args.tiles_y * args.tiles_x) % 2 == 0,
                  "2-CTA conv needs an even number of 128-pixel tiles");
    if (halo) return launch_halo<true, 256, 3, 9>(a1, a2, w, s1, s2, args, stream);
    return launch_impl2<6>(a1, a2, w, s1, s2, omap, args, stream);
  }
  if (halo) {
    if (BN == 256) return launch_halo<false, 256, 2, 5>(a1, a2, w, s1, s2, args, stream);
    if (BN == 128) return launch_halo<false, 128, 3, 9>(a1, a2, w, s1, s2, args, stream);
    return launch_halo<false, 64, 3, 12>(a1, a2, w, s1, s2, args, stream);
  }
  if (BN == 256) return launch_impl<256, 4>(a1, a2, w, s1, s2, omap, args, stream);
  if (BN == 128) return launch_impl<128, 6>(a1, a2, w, s1, s2, omap, args, stream);
  return launch_impl<64, 8>(a1, a2, w, s1, s2, omap, args, stream);
}

}  // namespace pdr
