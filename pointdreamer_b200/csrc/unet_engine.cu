// Native runtime of the ADM U-Net forward pass and the DDNM sampling loop.
//
// Reference: models/DDNM/guided_diffusion/unet.py:396-664 (UNetModel), script_util.py:130-185
// (create_model), diffusion.py:459-570 (simplified_ddnm_inpainting).  The engine rebuilds the
// reference's module structure from the same config, resolves parameters by the reference's
// state_dict names, plans every activation into one caller-provided arena (static first-fit
// planner, no allocation at run time), pre-encodes all TMA tensor maps and then replays a flat
// list of kernel launches: one C-ABI call per forward / per 100-step chain, no host
// synchronisation, no Python in the loop.
#include <functional>
#include <map>
#include <string>
#include <vector>
#include "common.cuh"
#include "conv_tc.h"
#include "unet_ops.h"
#include <stdlib.h>
#include "unet_engine.h"

namespace pdr {

struct Param {
  const void* ptr;
  size_t bytes;
};

struct Act {  // NHWC fp16 activation in the arena
  size_t off = 0;
  int B = 0, H = 0, W = 0, C = 0;
  bool skip = false;  // owned by the skip stack
  bool has_sums = false;  // per-8-channel GroupNorm sums were produced by the conv epilogue
  size_t sums_off = 0;    // double[B][C/8][2]
  size_t bytes() const { return (size_t)B * H * W * C * 2; }
};

class Planner {
 public:
  size_t alloc(size_t bytes) {
    bytes = (bytes + 1023) & ~(size_t)1023;
    for (size_t i = 0; i < blocks_.size(); ++i) {
      if (blocks_[i].free && blocks_[i].size >= bytes) {
        if (blocks_[i].size > bytes) {
          Block rest{blocks_[i].off + bytes, blocks_[i].size - bytes, true};
          blocks_[i].size = bytes;
          blocks_.insert(blocks_.begin() + i + 1, rest);
        }
        blocks_[i].free = false;
        return blocks_[i].off;
      }
    }
    Block nb{top_, bytes, false};
    if (!blocks_.empty() && blocks_.back().free) {  // grow the trailing free block
      nb.off = blocks_.back().off;
      blocks_.pop_back();
    }
    blocks_.push_back(nb);
    top_ = nb.off + bytes;
    if (top_ > high_) high_ = top_;
    return nb.off;
  }
  void release(size_t off) {
    for (size_t i = 0; i < blocks_.size(); ++i) {
      if (blocks_[i].off == off && !blocks_[i].free) {
        blocks_[i].free = true;
        if (i + 1 < blocks_.size() && blocks_[i + 1].free) {
          blocks_[i].size += blocks_[i + 1].size;
          blocks_.erase(blocks_.begin() + i + 1);
        }
        if (i > 0 && blocks_[i - 1].free) {
          blocks_[i - 1].size += blocks_[i].size;
          blocks_.erase(blocks_.begin() + i);
        }
        return;
      }
    }
  }
  size_t high_water() const { return high_; }

 private:
  struct Block {
    size_t off, size;
    bool free;
  };
  std::vector<Block> blocks_;
  size_t top_ = 0, high_ = 0;
};

struct ConvOp {
  ConvTensorMap a1, a2, w, s1, s2, o;
};

class UnetEngine {
 public:
  PdrUnetConfig cfg;
  std::map<std::string, Param> params;
  // plan
  int planned_B = 0;
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  using OpFn = std::function<int(const float*, const float*, float*, int, cudaStream_t)>;
  struct Op {
    OpFn fn;
    int cls;       // PdrOpClass
    double flops;  // algorithmic FLOPs of one launch (tensor-core convs only)
  };
  struct OpList {
    std::vector<Op> v;
    int cur_cls = 0;
    double cur_flops = 0.0;
    void push_back(OpFn f) { v.push_back(Op{std::move(f), cur_cls, cur_flops}); cur_flops = 0.0; }
    void clear() { v.clear(); }
  } ops;
  // profiling: CUDA events around every op of sampled forwards
  bool profiling = false;
  int profile_every = 1, forward_counter = 0;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_op;  // op index of each recorded (start, stop) pair
  size_t ev_used = 0;
  double prof_ms[PDR_OP_CLASSES] = {0}, prof_flops[PDR_OP_CLASSES] = {0};
  long long prof_launches[PDR_OP_CLASSES] = {0};
  long long prof_forwards = 0;
  // CUDA graph of one forward (sampler path): removes ~400 launch gaps per step
  cudaGraphExec_t graph_exec = nullptr;
  const float* graph_x = nullptr;
  float* graph_out = nullptr;
  int graph_nout = 0;
  bool graph_ok = true, warmed = false;
  unsigned long long graph_launches = 0;
  void drop_graph() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    graph_exec = nullptr;
    warmed = false;
  }
  std::vector<ConvOp*> conv_ops;
  std::string err;

  ~UnetEngine() {
    clear_plan();
    for (auto ev : ev_pool) cudaEventDestroy(ev);
  }
  void clear_plan() {
    drop_graph();
    ev_used = 0;
    ev_op.clear();
    for (auto* c : conv_ops) delete c;
    conv_ops.clear();
    ops.clear();
    planned_B = 0;
  }

  int time_embed_dim() const { return cfg.model_channels * 4; }

  const Param* find(const std::string& name, size_t expect_bytes) {
    auto it = params.find(name);
    if (it == params.end()) {
      set_error("U-Net parameter '%s' was not provided", name.c_str());
      return nullptr;
    }
    if (expect_bytes && it->second.bytes != expect_bytes) {
      set_error("U-Net parameter '%s' has %zu bytes, expected %zu", name.c_str(),
                it->second.bytes, expect_bytes);
      return nullptr;
    }
    return &it->second;
  }

  bool attn_at(int ds) const {
    for (int i = 0; i < cfg.n_attn_ds; ++i)
      if (cfg.attn_ds[i] == ds) return true;
    return false;
  }

  // ---- plan ---------------------------------------------------------------------------
  // dry == true only measures the arena.
  int plan(int B, void* workspace, size_t ws_bytes, bool dry, size_t* need);

 private:
  Planner pl_;
  bool dry_ = true;
  int B_ = 0;
  size_t off_stats_[2] = {0, 0}, off_gnws_ = 0, off_e1_ = 0, off_emb_ = 0, off_emb16_ = 0;
  size_t off_embcache_ = 0, off_embmeta_ = 0;  // timestep-embedding cache (unet_ops.h)
  size_t off_partial_ = 0, partial_bytes_ = 0, partial_reserved_ = 0;
  // GroupNorm (+FiLM) + SiLU fused into the halo conv's transform warps (PDR_FUSED_GN, read at every
  // plan): 2 = default, the convs with ONE N tile (Cout <= 256: the 256^2 and 128^2 layers, where
  // the activated tensor is largest and the transformed halo is not recomputed per N tile);
  // 1 = every eligible conv; 0 = separate gn_apply pass everywhere.  All modes produce the same
  // bits.  A/B in one call (profiles/r02j_*): 2018.5 ms (0) / 1976 ms (1) / 1971 ms (2) per shape.
  int fuse_gn_ = 0;  // 0 off, 1 every eligible conv, 2 only convs with ONE N tile (Cout <= 256)

 public:
  size_t off_tcur_ = 0;  // float[B]: the timestep the captured graph reads
  float* tcur() const { return reinterpret_cast<float*>(arena + off_tcur_); }

 private:
  int emb_total_ = 0;
  int emb_cursor_ = 0;

  template <class T>
  T* P(size_t off) const {
    return reinterpret_cast<T*>(arena + off);
  }

  Act new_act(int H, int W, int C) {
    Act a;
    a.B = B_, a.H = H, a.W = W, a.C = C;
    a.off = pl_.alloc(a.bytes());
    return a;
  }
  void drop(const Act& a) {
    if (a.skip) return;
    pl_.release(a.off);
    if (a.has_sums) pl_.release(a.sums_off);
  }

  int add_gn(const Act& x1, const Act* x2, const std::string& pname, int which,
             const __half* /*unused*/, bool for_head = false) {
    const int C = x1.C + (x2 ? x2->C : 0);
    const Param* g = find(pname + ".weight", (size_t)C * 4);
    const Param* b = find(pname + ".bias", (size_t)C * 4);
    if (!g || !b) return -1;
    if (dry_) return 0;
    const __half* p1 = P<__half>(x1.off);
    const __half* p2 = x2 ? P<__half>(x2->off) : nullptr;
    const int HW = x1.H * x1.W, C1 = x1.C, C2 = x2 ? x2->C : 0, Bn = B_;
    float* ws = P<float>(off_gnws_);
    float* st = P<float>(off_stats_[which]);
    ops.cur_cls = PDR_OP_GN_STATS;
    if (!for_head && x1.has_sums && (!x2 || x2->has_sums) && (C1 + C2) % 256 == 0)
      return 0;  // gn_apply derives mean/rstd from the conv epilogue's sums itself
    if (x1.has_sums && (!x2 || x2->has_sums) && (C1 + C2) % 256 == 0) {
      const double* s1 = P<double>(x1.sums_off);
      const double* s2 = x2 ? P<double>(x2->sums_off) : nullptr;
      ops.push_back([=](const float*, const float*, float*, int, cudaStream_t s) {
        return gn_finalize_sums_launch(s1, s2, Bn, HW, C1, C2, st, s);
      });
      return 0;
    }
    ops.push_back([=](const float*, const float*, float*, int, cudaStream_t s) {
      return gn_stats_launch(p1, p2, Bn, HW, C1, C2, ws, st, s);
    });
    return 0;
  }

  // raw: resampling blocks only - the resampled RAW input (x_upd of ResBlock._forward,
  // unet.py:241) is written by the same kernel from the same read (no separate resample pass)
  int add_gn_apply(const Act& x1, const Act* x2, const std::string& pname, int which, bool film,
                   int film_off, bool silu, int resample, const Act& out, const Act* raw = nullptr) {
    if (dry_) return 0;
    const int C = x1.C + (x2 ? x2->C : 0);
    const float* g = (const float*)find(pname + ".weight", (size_t)C * 4)->ptr;
    const float* b = (const float*)find(pname + ".bias", (size_t)C * 4)->ptr;
    const __half* p1 = P<__half>(x1.off);
    const __half* p2 = x2 ? P<__half>(x2->off) : nullptr;
    const int H = x1.H, W = x1.W, C1 = x1.C, C2 = x2 ? x2->C : 0, Bn = B_;
    const float* st = P<float>(off_stats_[which]);
    const bool from_sums = x1.has_sums && (!x2 || x2->has_sums) && C % 256 == 0;
    const double* s1 = from_sums ? P<double>(x1.sums_off) : nullptr;
    const double* s2 = from_sums && x2 ? P<double>(x2->sums_off) : nullptr;
    const __half* fl = film ? P<__half>(off_emb16_) : nullptr;
    const int fstride = emb_total_;
    __half* o = P<__half>(out.off);
    __half* ro = raw ? P<__half>(raw->off) : nullptr;
    ops.cur_cls = PDR_OP_GN_APPLY;
    ops.push_back([=](const float*, const float*, float*, int, cudaStream_t s) {
      return gn_apply_launch(p1, p2, Bn, H, W, C1, C2, st, s1, s2, g, b, fl, fstride, film_off,
                             silu ? 1 : 0, resample, o, s, ro);
    });
    return 0;
  }

  // GroupNorm applied INSIDE the consuming conv (halo kernel): only the per-(image, channel)
  // constants are materialised.  Returns the arena offset of the float4[B][C] table in *coeff_off.
  int add_gn_coeff(const Act& x1, const Act* x2, const std::string& pname, int which, bool film,
                   int film_off, size_t* coeff_off) {
    const int C = x1.C + (x2 ? x2->C : 0);
    *coeff_off = pl_.alloc((size_t)B_ * C * sizeof(float4));
    if (dry_) return 0;
    const float* g = (const float*)find(pname + ".weight", (size_t)C * 4)->ptr;
    const float* b = (const float*)find(pname + ".bias", (size_t)C * 4)->ptr;
    const int H = x1.H, W = x1.W, C1 = x1.C, C2 = x2 ? x2->C : 0, Bn = B_;
    const float* st = P<float>(off_stats_[which]);
    const bool from_sums = x1.has_sums && (!x2 || x2->has_sums) && C % 256 == 0;
    const double* s1 = from_sums ? P<double>(x1.sums_off) : nullptr;
    const double* s2 = from_sums && x2 ? P<double>(x2->sums_off) : nullptr;
    const __half* fl = film ? P<__half>(off_emb16_) : nullptr;
    const int fstride = emb_total_;
    float4* co = P<float4>(*coeff_off);
    ops.cur_cls = PDR_OP_GN_APPLY;
    ops.push_back([=](const float*, const float*, float*, int, cudaStream_t s) {
      return gn_coeff_launch(Bn, H, W, C1, C2, from_sums ? nullptr : st, s1, s2, g, b, fl, fstride,
                             film_off, co, s);
    });
    return 0;
  }

  // want_sums: also produce the GroupNorm sums of `out` in the epilogue (when the tile allows)
  int add_conv(const Act& x1, const Act* x2, const std::string& pname, int taps, const Act* res,
               Act& out, bool want_sums = false, float qk_scale = 0.f, const Act* skip1 = nullptr,
               const Act* skip2 = nullptr, const size_t* gn_coeff_off = nullptr, int gn_film = 0) {
    const int stat_rows = want_sums && out.C % 256 == 0 ? conv_tc_stats_rows_per_image(out.H, out.W) : 0;
    if (stat_rows > 0) {
      out.has_sums = true;
      out.sums_off = pl_.alloc((size_t)B_ * (out.C / 8) * 2 * sizeof(double));
      const size_t need = (size_t)B_ * stat_rows * (out.C / 8) * 2 * sizeof(float);
      if (need > partial_bytes_) partial_bytes_ = need;
    }
    const int Cin = x1.C + (x2 ? x2->C : 0);
    const int Cs1 = skip1 ? skip1->C : 0, Cs2 = skip2 ? skip2->C : 0;  // fused 1x1 skip branch
    const Param* w = find(pname + ".weight", (size_t)out.C * (taps * Cin + Cs1 + Cs2) * 2);
    const Param* b = find(pname + ".bias", (size_t)out.C * 4);
    if (!w || !b) return -1;
    if (out.C % 64 != 0 || x1.C % 64 != 0 || (x2 && x2->C % 64 != 0)) {
      set_error("conv '%s': channels %d+%d -> %d not supported by the tcgen05 tile", pname.c_str(),
                x1.C, x2 ? x2->C : 0, out.C);
      return -1;
    }
    // tile + split-K choice; the split-K partial tiles share the statistics scratch (a conv has
    // one or the other, and its consumer kernel follows it on the stream)
    int bn = conv_tc_pick_bn(B_, out.H, out.W, out.C, taps);
    int ksplit = 1;
    if (stat_rows == 0 && qk_scale == 0.f && taps == 9 && !gn_coeff_off) {
      // decided for a FIXED reference batch of 8: splitting changes the fp32 summation order, and
      // a chain's bits must not depend on the batch it happens to run in
      ksplit = conv_tc_pick_split(8, out.H, out.W, out.C, taps * (Cin / 64) + (Cs1 + Cs2) / 64, &bn);
      const size_t need = conv_tc_split_workspace_bytes(B_, out.H, out.W, out.C, ksplit);
      if (need > partial_bytes_) partial_bytes_ = need;
    }
    if (dry_) return 0;
    ConvOp* c = new ConvOp();
    conv_ops.push_back(c);
    const int halo = ksplit == 1 && conv_tc_halo_ok(out.H, out.W, taps) ? 1 : 0;
    if (gn_coeff_off && !halo) {
      set_error("internal: fused GroupNorm requested for a conv that does not run the halo kernel");
      return -1;
    }
    const float4* coeff = gn_coeff_off ? P<float4>(*gn_coeff_off) : nullptr;
    PDR_TRY(conv_tc_make_act_map(&c->a1, P<__half>(x1.off), B_, x1.H, x1.W, x1.C, halo));
    if (x2)
      PDR_TRY(conv_tc_make_act_map(&c->a2, P<__half>(x2->off), B_, x2->H, x2->W, x2->C, halo));
    PDR_TRY(conv_tc_make_weight_map(&c->w, w->ptr, out.C, taps * Cin + Cs1 + Cs2,
                                    bn == 512 ? 128 : bn));
    if (skip1)
      PDR_TRY(conv_tc_make_act_map(&c->s1, P<__half>(skip1->off), B_, skip1->H, skip1->W, Cs1,
                                   halo));
    if (skip2)
      PDR_TRY(conv_tc_make_act_map(&c->s2, P<__half>(skip2->off), B_, skip2->H, skip2->W, Cs2,
                                   halo));
    const bool tma_out = !halo && ksplit == 1;  // plain kernels write their tiles with TMA stores
    if (tma_out) PDR_TRY(conv_tc_make_act_map(&c->o, P<__half>(out.off), B_, out.H, out.W, out.C, 0));
    const bool hs1 = skip1 != nullptr, hs2 = skip2 != nullptr;
    const int H = out.H, W = out.W, C1 = x1.C, C2 = x2 ? x2->C : 0, Co = out.C, Bn = B_;
    const float* bias = (const float*)b->ptr;
    const __half* r = res ? P<__half>(res->off) : nullptr;
    __half* o = P<__half>(out.off);
    const bool has2 = x2 != nullptr;
    float* partial = stat_rows > 0 ? P<float>(off_partial_) : nullptr;
    float* split_ws = ksplit > 1 ? P<float>(off_partial_) : nullptr;
    ops.cur_cls = PDR_OP_CONV_TC;
    ops.cur_flops = 2.0 * Bn * H * W * (double)Co * (taps * (C1 + C2) + Cs1 + Cs2);
    ops.push_back([=](const float*, const float*, float*, int, cudaStream_t s) {
      return conv_tc_launch(&c->a1, has2 ? &c->a2 : nullptr, &c->w, bn, Bn, H, W, C1, C2, Co, taps,
                            bias, r, o, partial, s, qk_scale, hs1 ? &c->s1 : nullptr,
                            hs2 ? &c->s2 : nullptr, Cs1, Cs2, ksplit, split_ws, halo, coeff,
                            gn_film, tma_out ? &c->o : nullptr);
    });
    if (stat_rows > 0) {
      double* sums = P<double>(out.sums_off);
      ops.cur_cls = PDR_OP_GN_STATS;
      ops.push_back([=](const float*, const float*, float*, int, cudaStream_t s) {
        return sums8_reduce_launch(partial, Bn, stat_rows, Co, sums, s);
      });
    }
    return 0;
  }

  int res_block(const std::string& p, const Act& x1, const Act* x2, int Cout, bool up, bool down,
                Act* result) {
    const int Cin = x1.C + (x2 ? x2->C : 0);
    const int resample = down ? 1 : (up ? 2 : 0);
    const int Ho = down ? x1.H / 2 : (up ? x1.H * 2 : x1.H);
    const int Wo = down ? x1.W / 2 : (up ? x1.W * 2 : x1.W);
    const int film_off = emb_cursor_;
    emb_cursor_ += 2 * Cout;
    // in_layers: GN + SiLU (+ resample) -> conv
    PDR_TRY(add_gn(x1, x2, p + ".in_layers.0", 0, nullptr));
    // GroupNorm + SiLU of the block input applied inside in_layers.2's halo kernel (no activated
    // copy of the input in HBM) when nothing is resampled in between
    // (mode 2: only where the transformed halo is not recomputed per N tile - the transform's
    // instruction energy is what the fusion costs)
    const bool fuse_here = fuse_gn_ == 1 || (fuse_gn_ == 2 && Cout <= 256);
    const bool fuse1 = fuse_here && !resample && conv_tc_halo_ok(Ho, Wo, 9);
    size_t coeff1 = 0;
    Act a1, xr;
    bool have_xr = false;
    if (resample) {
      if (x2) {
        set_error("resampling ResBlock with a concatenated input is not part of the model");
        return -1;
      }
      xr = new_act(Ho, Wo, Cin);
      have_xr = true;
    }
    if (fuse1) {
      PDR_TRY(add_gn_coeff(x1, x2, p + ".in_layers.0", 0, false, 0, &coeff1));
    } else {
      a1 = new_act(Ho, Wo, Cin);
      PDR_TRY(add_gn_apply(x1, x2, p + ".in_layers.0", 0, false, 0, true, resample, a1,
                           have_xr ? &xr : nullptr));
    }
    Act h1 = new_act(Ho, Wo, Cout);
    if (fuse1) {
      PDR_TRY(add_conv(x1, x2, p + ".in_layers.2", 9, nullptr, h1, true, 0.f, nullptr, nullptr,
                       &coeff1, 0));
      pl_.release(coeff1);
    } else {
      PDR_TRY(add_conv(a1, nullptr, p + ".in_layers.2", 9, nullptr, h1, true));
      drop(a1);
    }
    // out_layers: GN * (1+scale) + shift, SiLU, conv (+ skip)
    PDR_TRY(add_gn(h1, nullptr, p + ".out_layers.0", 1, nullptr));
    const bool fuse2 = fuse_here && conv_tc_halo_ok(Ho, Wo, 9);
    size_t coeff2 = 0;
    Act a2;
    if (fuse2) {
      PDR_TRY(add_gn_coeff(h1, nullptr, p + ".out_layers.0", 1, true, film_off, &coeff2));
      a2 = h1;  // out_layers.3 reads the raw conv output and activates it tile by tile
    } else {
      a2 = new_act(Ho, Wo, Cout);
      PDR_TRY(add_gn_apply(h1, nullptr, p + ".out_layers.0", 1, true, film_off, true, 0, a2));
      drop(h1);
    }
    Act out = new_act(Ho, Wo, Cout);
    if (Cin != Cout) {
      // channel-changing block: skip_connection (1x1) rides in the K loop of out_layers.3 -
      // "<p>.out_layers.3_skip" = [w3x3 | w1x1] along K, biases summed (set by the host side)
      const Act& sx = have_xr ? xr : x1;
      PDR_TRY(add_conv(a2, nullptr, p + ".out_layers.3_skip", 9, nullptr, out, true, 0.f, &sx,
                       have_xr ? nullptr : x2, fuse2 ? &coeff2 : nullptr, 1));
    } else {
      Act skip = have_xr ? xr : x1;
      PDR_TRY(add_conv(a2, nullptr, p + ".out_layers.3", 9, &skip, out, true, 0.f, nullptr, nullptr,
                       fuse2 ? &coeff2 : nullptr, 1));
    }
    if (fuse2) pl_.release(coeff2);
    drop(a2);
    if (have_xr) drop(xr);
    *result = out;
    return 0;
  }

  int attn_block(const std::string& p, const Act& x, Act* result) {
    const int C = x.C;
    const int dh = cfg.num_head_channels;
    if (dh != 64) {
      set_error("attention head dim %d unsupported (the model uses 64)", dh);
      return -1;
    }
    const int heads = C / dh;
    PDR_TRY(add_gn(x, nullptr, p + ".norm", 0, nullptr));
    Act n = new_act(x.H, x.W, C);
    PDR_TRY(add_gn_apply(x, nullptr, p + ".norm", 0, false, 0, false, 0, n));
    Act qkv = new_act(x.H, x.W, 3 * C);
    // q*scale and k*scale (unet.py:349-351) are applied by the projection's epilogue
    PDR_TRY(add_conv(n, nullptr, p + ".qkv", 1, nullptr, qkv, false, 0.35355339059327373f));
    drop(n);
    Act a = new_act(x.H, x.W, C);
    if (!dry_) {
      const __half* q = P<__half>(qkv.off);
      __half* o = P<__half>(a.off);
      const int T = x.H * x.W, Bn = B_;
      ops.cur_cls = PDR_OP_ATTENTION;
      ops.cur_flops = 4.0 * Bn * heads * (double)T * T * 64;
      ops.push_back([=](const float*, const float*, float*, int, cudaStream_t s) {
        return attention_launch(q, Bn, T, heads, 1, o, s);
      });
    }
    drop(qkv);
    Act out = new_act(x.H, x.W, C);
    PDR_TRY(add_conv(a, nullptr, p + ".proj_out", 1, &x, out, true));
    drop(a);
    *result = out;
    return 0;
  }
};

int UnetEngine::plan(int B, void* workspace, size_t ws_bytes, bool dry, size_t* need) {
  clear_plan();
  pl_ = Planner();
  dry_ = dry;
  fuse_gn_ = getenv("PDR_FUSED_GN") ? atoi(getenv("PDR_FUSED_GN")) : 2;
  if (fuse_gn_ < 0 || fuse_gn_ > 2) fuse_gn_ = 2;
  B_ = B;
  arena = (uint8_t*)workspace;
  arena_bytes = ws_bytes;
  emb_cursor_ = 0;
  const int mc = cfg.model_channels, ted = time_embed_dim();
  const int S = cfg.image_size;

  // total width of all emb_layers outputs (2*Cout per ResBlock), in construction order
  {
    int total = 0, ch = (int)(cfg.channel_mult_x2[0] * mc / 2), ds = 1;
    std::vector<int> chans{ch};
    for (int level = 0; level < cfg.n_mult; ++level) {
      const int co = (int)(cfg.channel_mult_x2[level] * mc / 2);
      for (int i = 0; i < cfg.num_res_blocks; ++i) {
        total += 2 * co;
        ch = co;
        chans.push_back(ch);
      }
      if (level != cfg.n_mult - 1) {
        total += 2 * ch;
        chans.push_back(ch);
        ds *= 2;
      }
    }
    total += 2 * ch * 2;  // middle block: two ResBlocks
    for (int level = cfg.n_mult - 1; level >= 0; --level) {
      const int co = (int)(cfg.channel_mult_x2[level] * mc / 2);
      for (int i = 0; i <= cfg.num_res_blocks; ++i) {
        total += 2 * co;
        ch = co;
        if (level && i == cfg.num_res_blocks) total += 2 * ch;
      }
    }
    emb_total_ = total;
  }

  // persistent scratch
  off_stats_[0] = pl_.alloc((size_t)B * 64 * 4);
  off_stats_[1] = pl_.alloc((size_t)B * 64 * 4);
  {
    size_t mx = 0;
    for (int r = S; r >= 1; r /= 2) {
      const size_t slabs = gn_stats_slabs(B, r * r);
      mx = std::max(mx, slabs);
    }
    off_gnws_ = pl_.alloc((size_t)B * mx * 2 * 4096 * 4);
  }
  // scratch of the conv epilogue's statistics rows (size known from the dry pass)
  off_partial_ = pl_.alloc(partial_reserved_ > 0 ? partial_reserved_ : 1024);
  partial_bytes_ = 0;
  off_tcur_ = pl_.alloc((size_t)B * 4);
  off_e1_ = pl_.alloc((size_t)B * ted * 4);
  off_emb_ = pl_.alloc((size_t)B * ted * 4);
  off_emb16_ = pl_.alloc((size_t)B * emb_total_ * 2);
  off_embcache_ = pl_.alloc((size_t)EMB_CACHE_SLOTS * emb_total_ * 2);
  off_embmeta_ = pl_.alloc(sizeof(EmbCacheMeta));
  if (!dry_) PDR_CUDA(cudaMemset(arena + off_embmeta_, 0, sizeof(EmbCacheMeta)));  // empty cache

  // ---- time embedding + all emb_layers (fp32) ----
  {
    const Param* w0 = find("time_embed.0.weight", (size_t)ted * mc * 4);
    const Param* b0 = find("time_embed.0.bias", (size_t)ted * 4);
    const Param* w2 = find("time_embed.2.weight", (size_t)ted * ted * 4);
    const Param* b2 = find("time_embed.2.bias", (size_t)ted * 4);
    const Param* we = find("emb_all.weight", (size_t)emb_total_ * ted * 4);
    const Param* be = find("emb_all.bias", (size_t)emb_total_ * 4);
    if (!w0 || !b0 || !w2 || !b2 || !we || !be) return -1;
    if (!dry_) {
      float* e1 = P<float>(off_e1_);
      float* emb = P<float>(off_emb_);
      __half* e16 = P<__half>(off_emb16_);
      const int etot = emb_total_;
      __half* cache = P<__half>(off_embcache_);
      EmbCacheMeta* meta = P<EmbCacheMeta>(off_embmeta_);
      // PDR_NO_EMB_CACHE (read at every plan): A/B switch, every forward recomputes the embeddings
      const int cache_on = getenv("PDR_NO_EMB_CACHE") == nullptr;
      ops.cur_cls = PDR_OP_LINEAR;
      ops.push_back([=](const float*, const float* t, float*, int, cudaStream_t s) {
        PDR_TRY(emb_cache_lookup_launch(t, B, meta, cache_on, s));
        PDR_TRY(linear_launch(t, (const float*)w0->ptr, (const float*)b0->ptr, B, mc, ted, 2, e1,
                              nullptr, s, &meta->hit));
        PDR_TRY(linear_launch(e1, (const float*)w2->ptr, (const float*)b2->ptr, B, ted, ted, 1,
                              emb, nullptr, s, &meta->hit));
        PDR_TRY(linear_launch(emb, (const float*)we->ptr, (const float*)be->ptr, B, ted, etot, 1,
                              nullptr, e16, s, &meta->hit));
        return emb_cache_finish_launch(e16, B, etot, cache, meta, t, s);
      });
    }
  }

  // ---- input blocks ----
  std::vector<Act> hs;
  int ch = (int)(cfg.channel_mult_x2[0] * mc / 2);
  Act h = new_act(S, S, ch);
  {
    const Param* w = find("input_blocks.0.0.weight", (size_t)ch * 27 * 2);
    const Param* b = find("input_blocks.0.0.bias", (size_t)ch * 4);
    if (!w || !b) return -1;
    const bool tc_stem = params.count("input_blocks.0.0_tc.weight") && ch % 64 == 0;
    if (tc_stem) {
      // stem as a GEMM on the tensor cores: 27 -> 64 zero-padded patches, [ch][64] weights
      Act patches = new_act(S, S, 64);
      if (!dry_) {
        __half* pp = P<__half>(patches.off);
        ops.cur_cls = PDR_OP_STEM;
        ops.push_back([=](const float* x, const float*, float*, int, cudaStream_t s) {
          return stem_im2col_launch(x, B, S, S, pp, s);
        });
      }
      PDR_TRY(add_conv(patches, nullptr, "input_blocks.0.0_tc", 1, nullptr, h, true));
      drop(patches);
    } else if (!dry_) {
      __half* o = P<__half>(h.off);
      const int Cc = ch;
      ops.cur_cls = PDR_OP_STEM;
      ops.push_back([=](const float* x, const float*, float*, int, cudaStream_t s) {
        return stem_conv_launch(x, (const __half*)w->ptr, (const float*)b->ptr, B, S, S, Cc, o, s);
      });
    }
  }
  h.skip = true;
  hs.push_back(h);
  int ds = 1, blk = 1;
  for (int level = 0; level < cfg.n_mult; ++level) {
    const int co = (int)(cfg.channel_mult_x2[level] * mc / 2);
    for (int i = 0; i < cfg.num_res_blocks; ++i) {
      const std::string p = "input_blocks." + std::to_string(blk);
      Act r;
      PDR_TRY(res_block(p + ".0", h, nullptr, co, false, false, &r));
      if (attn_at(ds)) {
        Act a;
        PDR_TRY(attn_block(p + ".1", r, &a));
        drop(r);
        r = a;
      }
      h = r;
      h.skip = true;
      hs.push_back(h);
      ++blk;
    }
    if (level != cfg.n_mult - 1) {
      const std::string p = "input_blocks." + std::to_string(blk);
      Act r;
      PDR_TRY(res_block(p + ".0", h, nullptr, h.C, false, true, &r));
      h = r;
      h.skip = true;
      hs.push_back(h);
      ++blk;
      ds *= 2;
    }
  }
  // ---- middle ----
  {
    Act r0, a, r1;
    PDR_TRY(res_block("middle_block.0", h, nullptr, h.C, false, false, &r0));
    PDR_TRY(attn_block("middle_block.1", r0, &a));
    drop(r0);
    PDR_TRY(res_block("middle_block.2", a, nullptr, a.C, false, false, &r1));
    drop(a);
    h = r1;  // not a skip
  }
  // ---- output blocks ----
  blk = 0;
  for (int level = cfg.n_mult - 1; level >= 0; --level) {
    const int co = (int)(cfg.channel_mult_x2[level] * mc / 2);
    for (int i = 0; i <= cfg.num_res_blocks; ++i) {
      const std::string p = "output_blocks." + std::to_string(blk);
      Act skip = hs.back();
      hs.pop_back();
      Act r;
      PDR_TRY(res_block(p + ".0", h, &skip, co, false, false, &r));
      drop(h);
      skip.skip = false;
      drop(skip);
      int sub = 1;
      if (attn_at(ds)) {
        Act a;
        PDR_TRY(attn_block(p + "." + std::to_string(sub), r, &a));
        drop(r);
        r = a;
        ++sub;
      }
      if (level && i == cfg.num_res_blocks) {
        Act u;
        PDR_TRY(res_block(p + "." + std::to_string(sub), r, nullptr, r.C, true, false, &u));
        drop(r);
        r = u;
        ds /= 2;
      }
      h = r;
      ++blk;
    }
  }
  // ---- head ----
  {
    PDR_TRY(add_gn(h, nullptr, "out.0", 0, nullptr, true));
    const Param* g = find("out.0.weight", (size_t)h.C * 4);
    const Param* b = find("out.0.bias", (size_t)h.C * 4);
    const Param* w = find("out.2.weight", (size_t)cfg.out_channels * h.C * 9 * 4);
    const Param* bb = find("out.2.bias", (size_t)cfg.out_channels * 4);
    if (!g || !b || !w || !bb) return -1;
    if (!dry_) {
      const __half* hp = P<__half>(h.off);
      const float* st = P<float>(off_stats_[0]);
      const int Cc = h.C;
      ops.cur_cls = PDR_OP_HEAD;
      ops.push_back([=](const float*, const float*, float* out, int n_out, cudaStream_t s) {
        return head_launch(hp, st, (const float*)g->ptr, (const float*)b->ptr,
                           (const float*)w->ptr, (const float*)bb->ptr, B, S, S, Cc, n_out, out,
                           n_out, s);
      });
    }
    drop(h);
  }
  if (emb_cursor_ != emb_total_) {
    set_error("internal: emb_layers width mismatch (%d vs %d)", emb_cursor_, emb_total_);
    return -1;
  }
  if (dry_) {
    // the scratch was a placeholder in the dry pass: account for its real size
    partial_reserved_ = (partial_bytes_ + 1023) & ~(size_t)1023;
    if (need) *need = pl_.high_water() + partial_reserved_;
    return 0;
  }
  if (partial_bytes_ > partial_reserved_) {
    set_error("internal: statistics scratch grew between the dry and the real plan");
    return -1;
  }
  if (need) *need = pl_.high_water();
  if (!dry_) {
    if (pl_.high_water() > ws_bytes) {
      set_error("U-Net workspace too small: need %zu bytes, got %zu", pl_.high_water(), ws_bytes);
      clear_plan();
      return -1;
    }
    planned_B = B;
  }
  return 0;
}

// ------------------------------------------------------------------------------ C surface --
int unet_create(const PdrUnetConfig* cfg, void** handle) {
  PDR_CHECK_ARG(cfg && handle, "unet_create: null argument");
  PDR_CHECK_ARG(cfg->n_mult >= 1 && cfg->n_mult <= 8 && cfg->n_attn_ds >= 0 && cfg->n_attn_ds <= 8,
                "unet_create: bad config");
  PDR_CHECK_ARG(cfg->in_channels == 3, "unet_create: in_channels must be 3");
  UnetEngine* e = new UnetEngine();
  e->cfg = *cfg;
  *handle = e;
  return 0;
}
int unet_destroy(void* handle) {
  delete (UnetEngine*)handle;
  return 0;
}
int unet_set_param(void* handle, const char* name, const void* ptr, size_t bytes) {
  PDR_CHECK_ARG(handle && name && ptr, "unet_set_param: null argument");
  UnetEngine* e = (UnetEngine*)handle;
  e->params[name] = Param{ptr, bytes};
  e->clear_plan();
  return 0;
}
int unet_workspace_bytes(void* handle, int B, size_t* bytes) {
  PDR_CHECK_ARG(handle && bytes && B > 0, "unet_workspace_bytes: bad argument");
  return ((UnetEngine*)handle)->plan(B, nullptr, 0, true, bytes);
}
int unet_plan(void* handle, int B, void* workspace, size_t bytes) {
  PDR_CHECK_ARG(handle && workspace && B > 0, "unet_plan: bad argument");
  PDR_CHECK_ARG(((uintptr_t)workspace & 1023) == 0, "unet_plan: workspace must be 1 KiB aligned");
  size_t need = 0;
  PDR_TRY(((UnetEngine*)handle)->plan(B, nullptr, 0, true, &need));  // sizes the scratch
  return ((UnetEngine*)handle)->plan(B, workspace, bytes, false, nullptr);
}
int unet_forward(void* handle, const float* x, const float* t, float* out, int n_out,
                 cudaStream_t stream) {
  PDR_CHECK_ARG(handle && x && t && out, "unet_forward: null argument");
  UnetEngine* e = (UnetEngine*)handle;
  PDR_CHECK_ARG(e->planned_B > 0, "unet_forward: call pdr_unet_plan first");
  PDR_CHECK_ARG(n_out >= 1 && n_out <= e->cfg.out_channels, "unet_forward: bad n_out");
  const bool sample = e->profiling && (e->forward_counter++ % e->profile_every == 0) &&
                      e->ev_used + 2 * e->ops.v.size() <= e->ev_pool.size();
  if (!sample) {
    for (auto& op : e->ops.v) PDR_TRY(op.fn(x, t, out, n_out, stream));
    return 0;
  }
  for (size_t i = 0; i < e->ops.v.size(); ++i) {
    PDR_CUDA(cudaEventRecord(e->ev_pool[e->ev_used], stream));
    PDR_TRY(e->ops.v[i].fn(x, t, out, n_out, stream));
    PDR_CUDA(cudaEventRecord(e->ev_pool[e->ev_used + 1], stream));
    e->ev_op.push_back((int)i);
    e->ev_used += 2;
  }
  e->prof_forwards++;
  return 0;
}

// Sampling profiler: time every op of each `every`-th forward with CUDA events on the launching
// stream (bench.py's live roofline measurement).  max_forwards bounds the event pool.
int unet_profile_begin(void* handle, int every, int max_forwards) {
  PDR_CHECK_ARG(handle && every >= 1 && max_forwards >= 1, "unet_profile_begin: bad argument");
  UnetEngine* e = (UnetEngine*)handle;
  PDR_CHECK_ARG(e->planned_B > 0, "unet_profile_begin: plan first");
  const size_t want = 2 * e->ops.v.size() * (size_t)max_forwards;
  while (e->ev_pool.size() < want) {
    cudaEvent_t ev;
    PDR_CUDA(cudaEventCreate(&ev));
    e->ev_pool.push_back(ev);
  }
  e->profiling = true;
  e->profile_every = every;
  e->forward_counter = 0;
  e->ev_used = 0;
  e->ev_op.clear();
  e->prof_forwards = 0;
  for (int c = 0; c < PDR_OP_CLASSES; ++c) e->prof_ms[c] = e->prof_flops[c] = 0, e->prof_launches[c] = 0;
  return 0;
}
// Synchronises the events, accumulates per-class totals and stops profiling.
int unet_profile_end(void* handle, double* ms, double* flops, long long* launches,
                     long long* forwards) {
  PDR_CHECK_ARG(handle && ms && flops && launches && forwards, "unet_profile_end: null argument");
  UnetEngine* e = (UnetEngine*)handle;
  for (size_t k = 0; k < e->ev_op.size(); ++k) {
    PDR_CUDA(cudaEventSynchronize(e->ev_pool[2 * k + 1]));
    float t = 0.f;
    PDR_CUDA(cudaEventElapsedTime(&t, e->ev_pool[2 * k], e->ev_pool[2 * k + 1]));
    const auto& op = e->ops.v[e->ev_op[k]];
    e->prof_ms[op.cls] += t;
    e->prof_flops[op.cls] += op.flops;
    e->prof_launches[op.cls] += 1;
  }
  for (int c = 0; c < PDR_OP_CLASSES; ++c) {
    ms[c] = e->prof_ms[c];
    flops[c] = e->prof_flops[c];
    launches[c] = e->prof_launches[c];
  }
  *forwards = e->prof_forwards;
  e->profiling = false;
  e->ev_used = 0;
  e->ev_op.clear();
  return 0;
}

// One forward replayed from a CUDA graph (x / out / n_out fixed, t read from the engine's t_cur
// buffer).  The first call after planning runs eagerly (one-time kernel attribute setup), the
// second captures; any capture failure permanently falls back to eager launches.
static int unet_forward_graphed(UnetEngine* e, const float* x, float* out, int n_out,
                                cudaStream_t stream) {
  if (!e->graph_ok) return unet_forward(e, x, e->tcur(), out, n_out, stream);
  if (e->profiling) {
    // sampled forwards run eagerly with events around every launch; the others replay the graph
    if (e->forward_counter % e->profile_every == 0)
      return unet_forward(e, x, e->tcur(), out, n_out, stream);
    e->forward_counter++;
  }
  if (!e->warmed) {
    e->warmed = true;
    return unet_forward(e, x, e->tcur(), out, n_out, stream);
  }
  if (!e->graph_exec || e->graph_x != x || e->graph_out != out || e->graph_nout != n_out) {
    if (e->graph_exec) cudaGraphExecDestroy(e->graph_exec);
    e->graph_exec = nullptr;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      e->graph_ok = false;
      return unet_forward(e, x, e->tcur(), out, n_out, stream);
    }
    const bool was_profiling = e->profiling;
    e->profiling = false;
    const unsigned long long l0 = g_launch_count;
    const int rc = unet_forward(e, x, e->tcur(), out, n_out, stream);
    e->graph_launches = g_launch_count - l0;
    g_launch_count = l0;  // captured, not executed
    e->profiling = was_profiling;
    const cudaError_t ce = cudaStreamEndCapture(stream, &graph);
    if (rc != 0 || ce != cudaSuccess || graph == nullptr ||
        cudaGraphInstantiate(&e->graph_exec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      e->graph_exec = nullptr;
      e->graph_ok = false;
      return unet_forward(e, x, e->tcur(), out, n_out, stream);
    }
    cudaGraphDestroy(graph);
    e->graph_x = x;
    e->graph_out = out;
    e->graph_nout = n_out;
  }
  PDR_CUDA(cudaGraphLaunch(e->graph_exec, stream));
  g_launch_count += e->graph_launches;  // kernels of ours executed by the replay
  return 0;
}

// Engines of one process are serialised on the device: work submitted through the public entry
// points (pdr_unet_forward, pdr_ddnm_sample) waits for the previous submission of ANY engine on ANY
// stream of the same device.  The tcgen05 conv kernels are persistent, allocate all of an SM's
// tensor memory and (2-CTA variants) need both SMs of a pair; in round 1 four engines driven from
// four streams hung a box once.  Running them concurrently cannot be faster anyway - one engine
// already holds the GPU at its 1 kW power limit (profiles/r02f_power_probe.json) - so the library
// orders them instead of leaving the interleaving to the hardware scheduler.
struct DeviceTurn {
  cudaEvent_t ev = nullptr;
  bool recorded = false;
};
static DeviceTurn g_turn[64];
static DeviceTurn* device_turn() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  DeviceTurn* t = &g_turn[dev];
  if (!t->ev && cudaEventCreateWithFlags(&t->ev, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return t;
}
static void turn_begin(cudaStream_t stream) {
  DeviceTurn* t = device_turn();
  if (t && t->recorded) cudaStreamWaitEvent(stream, t->ev, 0);
}
static void turn_end(cudaStream_t stream) {
  DeviceTurn* t = device_turn();
  if (t && cudaEventRecord(t->ev, stream) == cudaSuccess) t->recorded = true;
}

int unet_forward_serialized(void* handle, const float* x, const float* t, float* out, int n_out,
                            cudaStream_t stream) {
  turn_begin(stream);
  const int rc = unet_forward(handle, x, t, out, n_out, stream);
  turn_end(stream);
  return rc;
}

// DDNM chain for V views at once (diffusion.py:459-570): prepare, `steps` x (U-Net + fused update),
// final transform.  coef_host: [steps][7] floats (DdnmStepCoef order); t_dev: [steps][V] device.
int ddnm_sample(void* handle, const float* sparse, const float* mask, int V, int steps,
                const float* coef_host, const float* t_dev, unsigned long long seed,
                unsigned long long offset_base, unsigned long long draws_per_chain, int chain0,
                float* x, float* y, float* et, float* out, cudaStream_t stream) {
  PDR_CHECK_ARG(handle && sparse && mask && coef_host && t_dev && x && y && et && out,
                "ddnm_sample: null argument");
  UnetEngine* e = (UnetEngine*)handle;
  PDR_CHECK_ARG(e->planned_B == V, "ddnm_sample: engine planned for batch %d, got %d chains",
                e->planned_B, V);
  PDR_CHECK_ARG(draws_per_chain >= (unsigned long long)steps + 1,
                "ddnm_sample: draws_per_chain must be >= steps + 1");
  const int S = e->cfg.image_size;
  turn_begin(stream);
  struct TurnEnd {
    cudaStream_t s;
    ~TurnEnd() { turn_end(s); }
  } turn_guard{stream};
  PDR_TRY(ddnm_prepare_launch(sparse, mask, V, 3, S, S, seed, offset_base, draws_per_chain, chain0,
                              y, x, stream));
  for (int s = 0; s < steps; ++s) {
    PDR_CUDA(cudaMemcpyAsync(e->tcur(), t_dev + (size_t)s * V, (size_t)V * sizeof(float),
                             cudaMemcpyDeviceToDevice, stream));
    PDR_TRY(unet_forward_graphed(e, x, et, 3, stream));
    DdnmStepCoef k;
    const float* c = coef_host + (size_t)s * 7;
    k.sqrt_1m_at = c[0], k.sqrt_at = c[1], k.sqrt_at_next = c[2], k.gamma_t = c[3];
    k.c1 = c[4], k.c2 = c[5], k.lambda_t = c[6];
    PDR_TRY(ddnm_step_launch(x, et, 3, y, mask, V, 3, S, S, k, seed, offset_base, draws_per_chain,
                             chain0, 1 + s, stream));
  }
  return ddnm_final_launch(x, (long long)V * 3 * S * S, out, stream);
}

int unet_planned_batch(void* handle) { return handle ? ((UnetEngine*)handle)->planned_B : 0; }
int unet_image_size(void* handle) { return handle ? ((UnetEngine*)handle)->cfg.image_size : 0; }
int unet_out_channels(void* handle) { return handle ? ((UnetEngine*)handle)->cfg.out_channels : 0; }

}  // namespace pdr
