// Self-attention of the ADM U-Net's AttentionBlock (QKVAttentionLegacy) on the 5th-generation tensor
// cores: S = Q K^T and O = P V as tcgen05.mma with the accumulators in tensor memory.
//
// Reference: models/DDNM/guided_diffusion/unet.py:299-305, 337-354.  Same contract as
// attention_kernel (attention.cu): qkv [B, T, 3C] fp16 in the legacy head-major channel order
// c = head*192 + {q: 0..63, k: 64..127, v: 128..191}, q and k already multiplied by 64^-1/4 (the qkv
// projection's epilogue does it), the logits are an fp16 tensor, softmax runs in fp32 and is cast to
// fp16, the weighted sum is an fp16 tensor.  Because the normalised probabilities are rounded to
// fp16 BEFORE the weighted sum, row max and row sum must be final before any P is formed: two
// passes over the keys, S recomputed in the second (a 128 x 128 x 64 MMA costs ~0.1 us).
//
// One CTA = 128 queries of one (image, head); thread = query row (TMEM lane) for the softmax.
//   warp 0   TMA producer: Q once, K chunks (128 keys, both passes, 2 stages), V chunks (pass 2)
//   warp 1   MMA issuer:   S[128 x 128] = Q K_c^T (K = 64: four tcgen05.mma), pass 2 also
//                          O[128 x 64] += P_c V_c (K = 128 keys: eight tcgen05.mma; V_c is read as an
//                          MN-major B operand straight from its [key][d] tile - no transpose)
//   warp 2   TMEM allocator (256 columns: S 128, O 64)
//   warps 4-11 softmax (two warps per TMEM lane quadrant, each owning 64 of the chunk's 128 key
//            columns): tcgen05.ld their row of S, round to fp16; pass 1: online max / sum (the two
//            halves of a row are merged once, through shared memory); pass 2: P = fp16(exp(s - m) / l)
//            written to shared memory in the K-major 128-byte-swizzled layout the MMA reads (a thread's
//            64 columns are exactly one k-tile row); at the end O -> fp16 -> global.
// 98 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM overlap each other's MMA,
// TMA and softmax phases, so the per-CTA pipeline itself is strictly sequential (one S buffer, one P
// buffer) and easy to reason about.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "conv_tc.h"
#include "unet_ops.h"

namespace pdr {

static constexpr int AT_Q = 128;        // queries per CTA
static constexpr int AT_KC = 128;       // keys per chunk
static constexpr int AT_D = 64;         // head dim
static constexpr int AT_TILE = 16384;   // 128 rows x 128 B
static constexpr int AT_OFF_Q = 0;
static constexpr int AT_OFF_K = AT_TILE;              // 2 stages
static constexpr int AT_OFF_V = 3 * AT_TILE;          // 1 stage
static constexpr int AT_OFF_P = 4 * AT_TILE;          // 128 x 128 fp16 = 2 k-tiles of 16 KB
static constexpr int AT_OFF_BAR = 6 * AT_TILE;
static constexpr int AT_OFF_ML = AT_OFF_BAR + 16 * 8 + 16;  // float2[2][128]: (m, l) of the two column halves
static constexpr int AT_SMEM = AT_OFF_ML + 2 * 128 * 8;
static constexpr int AT_THREADS = 384;
static constexpr float AT_LOG2E = 1.4426950408889634f;

// 2^x, one MUFU.EX2 (the library exp2f adds range fix-ups; arguments here are <= 0 and results
// that underflow are meant to be 0)
__device__ __forceinline__ float at_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MN-major, 128-byte-swizzled shared-memory matrix descriptor: rows are K (here: keys), 128 B apart,
// 64 contiguous MN elements (here: the head dim) per row; 8-row groups 1024 B apart (SBO).  One
// swizzle atom covers the whole MN extent (64), so the leading-dimension offset is never used.
__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(AT_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, int T, int heads,
                    __half* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + AT_OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* s_empty = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* p_empty = bars + 10;
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_ptr_smem = (uint32_t*)(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int nc = T / AT_KC;
  const int row0 = b * T;                 // first row of this image in the [B*T, 3C] view
  const int col_q = head * 3 * AT_D, col_k = col_q + AT_D, col_v = col_q + 2 * AT_D;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 8);  // one arrive per softmax warp
    mbar_init(p_full, 8);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

  if (warp == 0) {
    // ===================================================== TMA producer ====
    if (elect_one_sync()) {
      mbar_expect_tx(q_full, AT_TILE);
      tma_load_2d(smem + AT_OFF_Q, &tmQKV, q_full, col_q, row0 + qt * AT_Q);
      for (int i = 0; i < 2 * nc; ++i) {
        const int c = i < nc ? i : i - nc, st = i & 1;
        mbar_wait(&k_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], AT_TILE);
        tma_load_2d(smem + AT_OFF_K + st * AT_TILE, &tmQKV, &k_full[st], col_k, row0 + c * AT_KC);
        if (i >= nc) {
          mbar_wait(v_empty, (c & 1) ^ 1);
          mbar_expect_tx(v_full, AT_TILE);
          tma_load_2d(smem + AT_OFF_V, &tmQKV, v_full, col_v, row0 + c * AT_KC);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================================================= MMA issuer ====
    if (elect_one_sync()) {
      constexpr uint32_t idesc_s = make_idesc_f16(AT_Q, AT_KC);             // K-major A and B
      constexpr uint32_t idesc_o = make_idesc_f16(AT_Q, AT_D) | (1u << 16);  // B (= V) is MN-major
      const uint32_t sq = smem_u32(smem + AT_OFF_Q), sp = smem_u32(smem + AT_OFF_P);
      const uint32_t sv = smem_u32(smem + AT_OFF_V);
      mbar_wait(q_full, 0);
      for (int i = 0; i < 2 * nc; ++i) {
        const int st = i & 1;
        mbar_wait(&k_full[st], (i >> 1) & 1);
        mbar_wait(s_empty, (i & 1) ^ 1);  // the softmax warps have read the previous S
        tc_fence_after();
        const uint32_t sk = smem_u32(smem + AT_OFF_K + st * AT_TILE);
#pragma unroll
        for (int kk = 0; kk < AT_D / 16; ++kk)
          umma_f16(tmem_s, make_smem_desc_sw128(sq + kk * 32), make_smem_desc_sw128(sk + kk * 32),
                   idesc_s, kk != 0 ? 1u : 0u);
        umma_commit(&k_empty[st]);
        umma_commit(s_full);
        if (i >= nc) {
          const int c = i - nc;
          mbar_wait(p_full, c & 1);
          mbar_wait(v_full, c & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < AT_KC / 16; ++kk)
            umma_f16(tmem_o, make_smem_desc_sw128(sp + (kk >> 2) * AT_TILE + (kk & 3) * 32),
                     make_smem_desc_sw128_mn(sv + kk * 2048), idesc_o, (c | kk) != 0 ? 1u : 0u);
          umma_commit(v_empty);
          umma_commit(p_empty);
        }
      }
      umma_commit(o_full);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ========================================================== softmax ====
    const int q = warp & 3;           // TMEM lane quadrant of this warp
    const int hf = (warp - 4) >> 2;   // which 64 of the chunk's 128 key columns
    const int r = q * 32 + lane;      // query row of this thread == TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float m = -INFINITY, l = 0.f;
    uint32_t v[32];
    // ---- pass 1: row max and sum of exp over the fp16-rounded logits (this thread's columns) ----
    for (int i = 0; i < nc; ++i) {
      mbar_wait(s_full, i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int sc = 2 * hf; sc < 2 * hf + 2; ++sc) {
        tmem_ld_32x32(tmem_s + lane_off + sc * 32, v);
        tmem_ld_wait();
        float tm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 s2 = __half22float2(
              __floats2half2_rn(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
          v[j] = __float_as_uint(s2.x);
          v[j + 1] = __float_as_uint(s2.y);
          tm = fmaxf(tm, fmaxf(s2.x, s2.y));
        }
        const float mn = fmaxf(m, tm);
        const float mnl = -mn * AT_LOG2E;  // exp(s - mn) = 2^(s*log2e - mn*log2e): one FFMA + one MUFU
        float ts = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) ts += at_ex2(fmaf(__uint_as_float(v[j]), AT_LOG2E, mnl));
        l = l * at_ex2((m - mn) * AT_LOG2E) + ts;
        m = mn;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
    }
    // merge the two column halves of every row (the sum is commutative: both threads get the same bits)
    {
      float2* ml = (float2*)(smem + AT_OFF_ML);
      ml[hf * 128 + r] = make_float2(m, l);
      named_bar_sync(1, 256);
      const float2 o = ml[(hf ^ 1) * 128 + r];
      const float mm = fmaxf(m, o.x);
      const float la = l * at_ex2((m - mm) * AT_LOG2E), lb = o.y * at_ex2((o.x - mm) * AT_LOG2E);
      l = hf == 0 ? la + lb : lb + la;
      m = mm;
    }
    const float inv_l = 1.f / l;
    const float ml2 = -m * AT_LOG2E;
    // ---- pass 2: P = fp16(softmax) -> shared memory (A operand of P V) ----
    uint8_t* prow = smem + AT_OFF_P + hf * AT_TILE + r * 128;  // this thread's 64 keys = k-tile hf
    for (int c = 0; c < nc; ++c) {
      const int i = nc + c;
      mbar_wait(s_full, i & 1);
      mbar_wait(p_empty, (c & 1) ^ 1);  // the MMAs that read the previous P have retired
      tc_fence_after();
#pragma unroll 1
      for (int s2i = 0; s2i < 2; ++s2i) {
        tmem_ld_32x32(tmem_s + lane_off + (2 * hf + s2i) * 32, v);
        tmem_ld_wait();
        uint32_t p[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 s2 = __half22float2(
              __floats2half2_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])));
          const __half2 h = __floats2half2_rn(at_ex2(fmaf(s2.x, AT_LOG2E, ml2)) * inv_l,
                                              at_ex2(fmaf(s2.y, AT_LOG2E, ml2)) * inv_l);
          p[j] = *(const uint32_t*)&h;
        }
        // keys 32 s2i .. 32 s2i + 31 of this thread's k-tile row: 16-byte chunks s2i * 4 + 0..3
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *(uint4*)(prow + (((s2i * 4 + j) ^ (r & 7)) << 4)) =
              make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
      }
      tc_fence_before();
      fence_proxy_async();  // generic-proxy writes of P -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(s_empty);
        mbar_arrive(p_full);
      }
    }
    // ---- O -> fp16 -> global (32 of the 64 head-dim columns per thread) ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int C = heads * AT_D;
    __half* orow = out + ((size_t)(row0 + qt * AT_Q + r)) * C + head * AT_D + hf * 32;
    {
      tmem_ld_32x32(tmem_o + lane_off + hf * 32, v);
      tmem_ld_wait();
      __align__(16) __half o[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = __float2half_rn(__uint_as_float(v[j]));
#pragma unroll
      for (int j = 0; j < 4; ++j) ((uint4*)orow)[j] = ((const uint4*)o)[j];
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

bool attention_tc_ok(int T, int prescaled) {
  static const bool disabled = getenv("PDR_NO_TC_ATTENTION") != nullptr;  // A/B switch
  return !disabled && prescaled && T % AT_KC == 0 && T >= AT_KC;
}

int attention_tc_launch(const __half* qkv, int B, int T, int heads, __half* out,
                        cudaStream_t stream) {
  PDR_CHECK_ARG(T % AT_KC == 0 && heads >= 1 && B >= 1, "attention (tcgen05): bad shape");
  ConvTensorMap tm;
  // qkv viewed as a [B*T, 3C] fp16 matrix; one box = 128 rows x 64 columns (one of q / k / v of a head)
  PDR_TRY(conv_tc_make_map_2d(&tm, qkv, (unsigned long long)heads * 3 * AT_D,
                              (unsigned long long)B * T, AT_D, 128));
  static bool configured = false;
  if (!configured) {
    PDR_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  AT_SMEM + 1024));
    configured = true;
  }
  attention_tc_kernel<<<dim3(T / AT_Q, heads, B), AT_THREADS, AT_SMEM + 1024, stream>>>(
      *(const CUtensorMap*)&tm, T, heads, out);
  PDR_COUNT_LAUNCH();
  PDR_LAUNCH_CHECK();
  return 0;
}

}  // namespace pdr
