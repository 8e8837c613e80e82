// Prepared-weight records and the block walk shared by the two denoiser executors.
#pragma once
#include "model.cuh"

namespace echo {

struct ConvW {   // repacked [cout][taps][cin] (K index = tap*cin + c); linear layers have taps == 1
  const float* w = nullptr;
  const __nv_bfloat16* wb = nullptr;
  const __nv_bfloat16* wb_lo = nullptr;   // split-precision mode: wb = bf16(w), wb_lo = bf16(w - wb)
  const float* b = nullptr;
  int cout = 0, cin = 0, taps = 1;
};
struct NormW {
  const float* g = nullptr;
  const float* b = nullptr;
  int c = 0;
};
struct ResW {
  int cin = 0, cout = 0;
  NormW n1, n2;
  ConvW c1, c2, skip;
  bool has_skip = false;
  int emb_off = 0;   // column offset of this block's emb_layers output in the stacked emb projection
};
struct AttnW {
  int C = 0, heads = 0, dh = 0;
  NormW norm, ln1, ln3;
  ConvW proj_in, proj_out;
  ConvW qkv;        // [3C, C] = [to_q; to_k; to_v] of attn1 (bias-free)
  ConvW qkv_pad;    // bf16 mode: same, every head zero-padded to attention_pad_dh(dh) rows -> [3*heads*dhp, C]
  ConvW attn1_out;  // [C, C] + bias
  ConvW attn2_out;  // [C, C] + bias, applied to attn2.to_v(context) (one context token: softmax == 1)
  ConvW v_only;     // attn1.to_v alone [C, C] (layout branch: one token, self-attention == to_out(to_v(x)))
  ConvW ff1, ff2;   // [8C, C], [C, 4C]
  ConvW ff1_geglu;  // bf16 mode: ff1 rows permuted per 256-row tile to [128 a | 128 g] for the fused GEGLU epilogue
  int v2_off = 0;   // column offset of attn2.to_v(context) in the stacked projection
};
struct BlockW {
  enum Kind { CONV_IN, RES, DOWN } kind = RES;
  ResW res;
  bool attn = false;
  AttnW at;
  bool up = false;
  ConvW conv;   // CONV_IN / DOWN op / Upsample conv
  ConvW up_fold;  // bf16 mode, Upsample conv: weights folded per output phase (fold_upsample_weight), taps = 48
  int ds = 1;
};

struct UNetPlan {
  std::vector<BlockW> in_blocks, out_blocks;
  ResW mid0, mid2;
  AttnW mid_at;
  NormW out_norm;
  ConvW out_conv;
  ConvW stem_pad;           // bf16 mode: stem conv with cin zero-padded 3 -> 16 so it runs on the tcgen05 kernel
  ConvW out_conv_pad;       // bf16 mode: out conv zero-padded to 32 output channels so it runs on the tcgen05 kernel
  ConvW time0, time2;       // time_embed.0 / .2
  ConvW emb_stack;          // all ResBlock emb_layers.1 stacked [sum cout, 4*mc]
  ConvW v2_stack;           // all attn2.to_v stacked [sum C, context_dim] (bias-free)
  int emb_total = 0, v2_total = 0;
  int n_res = 0, n_attn = 0;
};

struct UNetCfg {
  int dims = 3;   // 1 (layout) or 3 (shape)
  int in_channels = 0, out_channels = 0, model_channels = 0;
  std::vector<int> channel_mult, attention_resolutions;
  int num_res_blocks = 2, num_heads = 8, context_dim = 1280;
  bool want_bf16 = false;
  bool want_x3 = false;   // hi/lo bf16 copies of the trunk's contraction weights (ECHO_PREC_X3), no layout-changing repacks
};

// Walks the reference constructor order (openai_model_3d.py:563-728 / denoise_net.py:553-713) and prepares
// every trunk weight: conv kernels repacked tap-major (centre tap only for the length-1 layout convs), emb_layers
// and attn2.to_v stacked, optional bf16 copies.
void build_unet_plan(const WeightMap& wm, const UNetCfg& cfg, DevPool& pool, UNetPlan& plan, cudaStream_t s);

}  // namespace echo
