// Weight preparation for the two UNets: walks the reference's constructor order and repacks each tensor once.
#include "unet.cuh"

#include <algorithm>

namespace echo {
namespace {

struct Prep {
  const WeightMap& wm;
  const UNetCfg& cfg;
  DevPool& pool;
  cudaStream_t s;

  const __nv_bfloat16* to_bf16(const float* w, size_t n) {
    if (!cfg.want_bf16) return nullptr;
    __nv_bfloat16* o = pool.alloc_n<__nv_bfloat16>(n);
    convert(w, F32, o, BF16, (int64_t)n, s);
    return o;
  }
  // split-precision copies of a contraction weight
  void to_x3(ConvW& c, size_t n) {
    if (!cfg.want_x3 || n % 4) return;
    __nv_bfloat16* hi = pool.alloc_n<__nv_bfloat16>(n);
    __nv_bfloat16* lo = pool.alloc_n<__nv_bfloat16>(n);
    split_bf16(c.w, (int64_t)n, hi, lo, s);
    c.wb = hi;
    c.wb_lo = lo;
  }
  const float* copy_vec(const std::string& name, int n) {
    const WView& v = wm.get(name, {n});
    float* o = pool.alloc_n<float>(n);
    ECHO_CUDA(cudaMemcpyAsync(o, v.p, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
    return o;
  }
  NormW norm(const std::string& p, int c) {
    NormW n;
    n.c = c;
    n.g = copy_vec(p + ".weight", c);
    n.b = copy_vec(p + ".bias", c);
    return n;
  }
  // conv_nd(dims, cin, cout, k): 3-D kernels are repacked tap-major; 1-D kernels act on length-1 signals with
  // padding k/2, so only the centre tap ever multiplies data.
  ConvW conv(const std::string& p, int cin, int cout, int k) {
    ConvW c;
    c.cin = cin;
    c.cout = cout;
    const size_t n_el = (size_t)cout * cin;
    if (cfg.dims == 3) {
      const WView& v = wm.get(p + ".weight", {cout, cin, k, k, k});
      c.taps = k * k * k;
      float* o = pool.alloc_n<float>(n_el * c.taps);
      if (c.taps == 1) ECHO_CUDA(cudaMemcpyAsync(o, v.p, sizeof(float) * n_el, cudaMemcpyDeviceToDevice, s));
      else repack_conv_weight(v.p, cout, cin, c.taps, o, s);
      c.w = o;
    } else {
      const WView& v = wm.get(p + ".weight", {cout, cin, k});
      c.taps = 1;
      float* o = pool.alloc_n<float>(n_el);
      conv1d_center_tap(v.p, cout, cin, k, o, s);
      c.w = o;
    }
    c.wb = to_bf16(c.w, n_el * c.taps);
    to_x3(c, n_el * c.taps);
    c.b = copy_vec(p + ".bias", cout);
    return c;
  }
  ConvW linear(const std::string& p, int cin, int cout, bool bias) {
    ConvW c;
    c.cin = cin;
    c.cout = cout;
    c.taps = 1;
    const WView& v = wm.get(p + ".weight", {cout, cin});
    float* o = pool.alloc_n<float>((size_t)cout * cin);
    ECHO_CUDA(cudaMemcpyAsync(o, v.p, sizeof(float) * cout * cin, cudaMemcpyDeviceToDevice, s));
    c.w = o;
    c.wb = to_bf16(o, (size_t)cout * cin);
    if (cfg.dims == 3) to_x3(c, (size_t)cout * cin);
    if (bias) c.b = copy_vec(p + ".bias", cout);
    return c;
  }
  // rows of several [cout_i, K] matrices stacked into one [sum, K]
  ConvW stack(const std::vector<std::pair<std::string, int>>& items, int K, bool bias, bool x3 = false) {
    int total = 0;
    for (auto& it : items) total += it.second;
    ConvW c;
    c.cin = K;
    c.cout = total;
    float* o = pool.alloc_n<float>((size_t)total * K);
    float* b = bias ? pool.alloc_n<float>(total) : nullptr;
    int off = 0;
    for (auto& it : items) {
      const WView& v = wm.get(it.first + ".weight", {it.second, K});
      ECHO_CUDA(cudaMemcpyAsync(o + (size_t)off * K, v.p, sizeof(float) * it.second * K, cudaMemcpyDeviceToDevice, s));
      if (bias) {
        const WView& bv = wm.get(it.first + ".bias", {it.second});
        ECHO_CUDA(cudaMemcpyAsync(b + off, bv.p, sizeof(float) * it.second, cudaMemcpyDeviceToDevice, s));
      }
      off += it.second;
    }
    c.w = o;
    c.b = b;
    c.wb = to_bf16(o, (size_t)total * K);
    if (x3) to_x3(c, (size_t)total * K);   // the stacked attn1 q/k/v (token-wise); emb / attn2.to_v stacks stay few-row fp32
    return c;
  }

  std::vector<std::pair<std::string, int>> emb_items, v2_items;
  int emb_total = 0, v2_total = 0;

  ResW res(const std::string& p, int cin, int cout) {
    ResW r;
    r.cin = cin;
    r.cout = cout;
    r.n1 = norm(p + "in_layers.0", cin);
    r.c1 = conv(p + "in_layers.2", cin, cout, 3);
    r.emb_off = emb_total;
    emb_items.push_back({p + "emb_layers.1", cout});
    emb_total += cout;
    r.n2 = norm(p + "out_layers.0", cout);
    r.c2 = conv(p + "out_layers.3", cout, cout, 3);
    r.has_skip = cin != cout;
    if (r.has_skip) r.skip = conv(p + "skip_connection", cin, cout, 1);
    return r;
  }

  AttnW attn(const std::string& p, int ch) {
    AttnW a;
    a.heads = cfg.num_heads;
    a.dh = ch / cfg.num_heads;
    a.C = a.heads * a.dh;
    ECHO_CHECK(a.C == ch, "attention: channels %d not divisible by heads %d", ch, cfg.num_heads);
    ECHO_CHECK(!wm.has(p + "transformer_blocks.1.norm1.weight"), "transformer_depth > 1 is not supported");
    const int C = a.C;
    a.norm = norm(p + "norm", ch);
    a.proj_in = conv(p + "proj_in", ch, C, 1);
    const std::string b = p + "transformer_blocks.0.";
    a.ln1 = norm(b + "norm1", C);
    a.ln3 = norm(b + "norm3", C);
    if (cfg.dims == 3) {
      a.qkv = stack({{b + "attn1.to_q", C}, {b + "attn1.to_k", C}, {b + "attn1.to_v", C}}, C, false, true);
      const int dhp = attention_pad_dh(a.dh);
      if (cfg.want_bf16 && dhp) {   // head-padded copy for the flash kernel: row (m*heads + h)*dhp + d <- row m*C + h*dh + d
        const size_t rows = (size_t)3 * a.heads * dhp;
        float* o = pool.alloc_n<float>(rows * C);
        ECHO_CUDA(cudaMemsetAsync(o, 0, rows * C * sizeof(float), s));
        for (int m = 0; m < 3; ++m)
          for (int h = 0; h < a.heads; ++h)
            ECHO_CUDA(cudaMemcpyAsync(o + ((size_t)(m * a.heads + h) * dhp) * C, a.qkv.w + ((size_t)m * C + (size_t)h * a.dh) * C,
                                      sizeof(float) * a.dh * C, cudaMemcpyDeviceToDevice, s));
        a.qkv_pad.w = o;
        a.qkv_pad.cin = C;
        a.qkv_pad.cout = (int)rows;
        a.qkv_pad.taps = 1;
        a.qkv_pad.wb = to_bf16(o, rows * C);
      }
    } else {
      a.v_only = linear(b + "attn1.to_v", C, C, false);
    }
    a.attn1_out = linear(b + "attn1.to_out.0", C, C, true);
    a.attn2_out = linear(b + "attn2.to_out.0", C, C, true);
    a.v2_off = v2_total;
    v2_items.push_back({b + "attn2.to_v", C});
    v2_total += C;
    a.ff1 = linear(b + "ff.net.0.proj", C, 8 * C, true);
    a.ff2 = linear(b + "ff.net.2", 4 * C, C, true);
    if (cfg.want_bf16 && cfg.dims == 3 && (4 * C) % 128 == 0) {
      const int F = 4 * C, tiles = F / 128;
      float* o = pool.alloc_n<float>((size_t)2 * F * C);
      float* bo = pool.alloc_n<float>(2 * F);
      for (int t = 0; t < tiles; ++t) {
        ECHO_CUDA(cudaMemcpyAsync(o + (size_t)(t * 256) * C, a.ff1.w + (size_t)(t * 128) * C, sizeof(float) * 128 * C, cudaMemcpyDeviceToDevice, s));
        ECHO_CUDA(cudaMemcpyAsync(o + (size_t)(t * 256 + 128) * C, a.ff1.w + (size_t)(F + t * 128) * C, sizeof(float) * 128 * C, cudaMemcpyDeviceToDevice, s));
        ECHO_CUDA(cudaMemcpyAsync(bo + t * 256, a.ff1.b + t * 128, sizeof(float) * 128, cudaMemcpyDeviceToDevice, s));
        ECHO_CUDA(cudaMemcpyAsync(bo + t * 256 + 128, a.ff1.b + F + t * 128, sizeof(float) * 128, cudaMemcpyDeviceToDevice, s));
      }
      a.ff1_geglu = a.ff1;
      a.ff1_geglu.w = o;
      a.ff1_geglu.b = bo;
      a.ff1_geglu.wb = to_bf16(o, (size_t)2 * F * C);
    }
    a.proj_out = conv(p + "proj_out", C, ch, 1);
    return a;
  }
};

}  // namespace

void build_unet_plan(const WeightMap& wm, const UNetCfg& cfg, DevPool& pool, UNetPlan& plan, cudaStream_t s) {
  Prep P{wm, cfg, pool, s};
  const int mc = cfg.model_channels, emb = 4 * mc;
  auto has_attn = [&](int ds) {
    return std::find(cfg.attention_resolutions.begin(), cfg.attention_resolutions.end(), ds) != cfg.attention_resolutions.end();
  };
  plan.time0 = P.linear("time_embed.0", mc, emb, true);
  plan.time2 = P.linear("time_embed.2", emb, emb, true);

  std::vector<int> chans;
  int ch = mc, ds = 1;
  {
    BlockW b;
    b.kind = BlockW::CONV_IN;
    b.conv = P.conv("input_blocks.0.0", cfg.in_channels, mc, 3);
    plan.in_blocks.push_back(b);
    chans.push_back(mc);
  }
  const int L = (int)cfg.channel_mult.size();
  for (int level = 0; level < L; ++level) {
    const int mult = cfg.channel_mult[level];
    for (int r = 0; r < cfg.num_res_blocks; ++r) {
      const std::string p = "input_blocks." + std::to_string(plan.in_blocks.size()) + ".";
      BlockW b;
      b.kind = BlockW::RES;
      b.ds = ds;
      b.res = P.res(p + "0.", ch, mult * mc);
      ch = mult * mc;
      b.attn = has_attn(ds);
      if (b.attn) b.at = P.attn(p + "1.", ch);
      plan.in_blocks.push_back(b);
      chans.push_back(ch);
    }
    if (level != L - 1) {
      const std::string p = "input_blocks." + std::to_string(plan.in_blocks.size()) + ".";
      BlockW b;
      b.kind = BlockW::DOWN;
      b.ds = ds;
      b.conv = P.conv(p + "0.op", ch, ch, 3);
      plan.in_blocks.push_back(b);
      chans.push_back(ch);
      ds *= 2;
    }
  }
  plan.mid0 = P.res("middle_block.0.", ch, ch);
  plan.mid_at = P.attn("middle_block.1.", ch);
  plan.mid2 = P.res("middle_block.2.", ch, ch);
  for (int level = L - 1; level >= 0; --level) {
    const int mult = cfg.channel_mult[level];
    for (int i = 0; i <= cfg.num_res_blocks; ++i) {
      const int ich = chans.back();
      chans.pop_back();
      const std::string p = "output_blocks." + std::to_string(plan.out_blocks.size()) + ".";
      BlockW b;
      b.kind = BlockW::RES;
      b.ds = ds;
      b.res = P.res(p + "0.", ch + ich, mc * mult);
      ch = mc * mult;
      int k = 1;
      b.attn = has_attn(ds);
      if (b.attn) {
        b.at = P.attn(p + "1.", ch);
        k = 2;
      }
      if (level && i == cfg.num_res_blocks) {
        b.up = true;
        b.conv = P.conv(p + std::to_string(k) + ".conv", ch, ch, 3);
        if (cfg.want_bf16 && cfg.dims == 3) {   // fold the nearest upsample into the conv: 12 taps per output phase
          const size_t n_src = (size_t)ch * 27 * ch, n_dst = (size_t)ch * 48 * ch;
          std::vector<float> hsrc(n_src), hdst(n_dst);
          ECHO_CUDA(cudaStreamSynchronize(s));
          ECHO_CUDA(cudaMemcpy(hsrc.data(), b.conv.w, n_src * sizeof(float), cudaMemcpyDeviceToHost));
          fold_upsample_weight(hsrc.data(), ch, ch, hdst.data());
          float* o = pool.alloc_n<float>(n_dst);
          ECHO_CUDA(cudaMemcpy(o, hdst.data(), n_dst * sizeof(float), cudaMemcpyHostToDevice));
          b.up_fold = b.conv;
          b.up_fold.taps = 48;
          b.up_fold.w = o;
          b.up_fold.wb = P.to_bf16(o, n_dst);
        }
        ds /= 2;
      }
      plan.out_blocks.push_back(b);
    }
  }
  if (cfg.want_bf16 && cfg.dims == 3 && cfg.in_channels < 16) {
    // stem: pad the input channels to one UMMA K step (16); the padded activation channels are zero as well
    const ConvW& c = plan.in_blocks[0].conv;
    const size_t n_src = (size_t)c.cout * c.taps * c.cin, n_dst = (size_t)c.cout * c.taps * 16;
    std::vector<float> hsrc(n_src), hdst(n_dst, 0.f);
    ECHO_CUDA(cudaStreamSynchronize(s));
    ECHO_CUDA(cudaMemcpy(hsrc.data(), c.w, n_src * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < (size_t)c.cout * c.taps; ++r)
      for (int ch = 0; ch < c.cin; ++ch) hdst[r * 16 + ch] = hsrc[r * c.cin + ch];
    float* o = pool.alloc_n<float>(n_dst);
    ECHO_CUDA(cudaMemcpy(o, hdst.data(), n_dst * sizeof(float), cudaMemcpyHostToDevice));
    plan.stem_pad = c;
    plan.stem_pad.cin = 16;
    plan.stem_pad.w = o;
    plan.stem_pad.wb = P.to_bf16(o, n_dst);
  }
  plan.out_norm = P.norm("out.0", mc);
  plan.out_conv = P.conv("out.2", mc, cfg.out_channels, 3);
  if (cfg.want_bf16 && cfg.dims == 3 && cfg.out_channels < 32) {
    // a 3-column GEMM wastes the tensor core far less than a scalar reduction wastes the SM: pad cout to 32 (zero rows)
    const size_t row = (size_t)plan.out_conv.taps * mc;
    float* o = pool.alloc_n<float>(32 * row);
    float* b = pool.alloc_n<float>(32);
    ECHO_CUDA(cudaMemsetAsync(o, 0, 32 * row * sizeof(float), s));
    ECHO_CUDA(cudaMemsetAsync(b, 0, 32 * sizeof(float), s));
    ECHO_CUDA(cudaMemcpyAsync(o, plan.out_conv.w, cfg.out_channels * row * sizeof(float), cudaMemcpyDeviceToDevice, s));
    ECHO_CUDA(cudaMemcpyAsync(b, plan.out_conv.b, cfg.out_channels * sizeof(float), cudaMemcpyDeviceToDevice, s));
    plan.out_conv_pad = plan.out_conv;
    plan.out_conv_pad.cout = 32;
    plan.out_conv_pad.w = o;
    plan.out_conv_pad.b = b;
    plan.out_conv_pad.wb = P.to_bf16(o, 32 * row);
  }
  plan.emb_stack = P.stack(P.emb_items, emb, true);
  plan.v2_stack = P.stack(P.v2_items, cfg.context_dim, false);
  plan.emb_total = P.emb_total;
  plan.v2_total = P.v2_total;
  plan.n_res = (int)P.emb_items.size();
  plan.n_attn = (int)P.v2_items.size();
}

}  // namespace echo
