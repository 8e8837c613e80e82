// Shape-branch denoiser step: UNet3DModel.forward (openai_model_3d.py:816-863) + the DDIM update
// (samplers/ddim.py:246-261) as a fixed sequence of kernels on one stream.
//
// Data layout in HBM: activations are channels-last (obj, d, h, w, c) so that (a) every GroupNorm / LayerNorm /
// GEGLU pass streams 16-byte channel vectors, (b) a 3x3x3 tap of the implicit GEMM reads a contiguous channel
// run (TMA box in the tcgen05 path), (c) "tokens x channels" for the transformer blocks is the same memory, no
// rearrange.  The reference's NCDHW latent is converted once on entry and once on exit (3 channels).
// Per-object vectors (time embedding, the 17 emb_layers outputs, the 11 attn2 contributions, the echo GCN) are
// computed once per step by few-row kernels and consumed as `rowvec` epilogue terms of the big contractions.
#include "unet.cuh"

#include <math.h>

namespace echo {
bool conv3d_small_cout_supported(int cin, int cout, int taps);
void conv3d_small_cout(const Act& x, const float* wt, const float* bias, int cout, float* out, cudaStream_t s);
}  // namespace echo

using namespace echo;

struct echo_shape {
  echo_shape_desc_t d;
  DevPool pool;
  UNetPlan plan;
  Gcn gcn;
  Arena arena;        // trunk temporaries (main stream)
  Arena side_arena;   // shape_embeddings temporaries (side stream: must not alias the trunk's stack)
  cudaStream_t side = nullptr;
  cudaStream_t side2 = nullptr;   // ResBlock 1x1 skip convolutions: independent of the GN -> conv -> GN chain, they fill its tails
  cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_codes = nullptr;
  bool dry = false;
  int prec = ECHO_PREC_FP32;
  bool batch_stats = false;   // shape_code_graph_cov normalises with the statistics of the batch (model.train() forward values)
  DT adt = F32;   // trunk activation dtype
  // prepared small weights
  ConvW se_conv0, se_conv2, se_lin, time_emb_lin;
  const float* pred_table = nullptr;
  int pred_rows = 0;   // rows of pred_embeddings (16)
  const float* freqs = nullptr;
  // schedule
  std::vector<float> h_coef;
  std::vector<int32_t> h_ts;
  float* d_coef = nullptr;
  int32_t* d_ts = nullptr;     // the DDIM timesteps on the device
  int* d_index = nullptr;      // the step's DDIM index for ECHO_INDEX_FROM_DEVICE calls (echo_shape_set_index)
  // persistent per-step buffers
  float *temb = nullptr, *e1 = nullptr, *emb = nullptr, *emb_act = nullptr, *node = nullptr, *pred = nullptr, *latent = nullptr, *codes = nullptr;
  float *embout = nullptr, *v2 = nullptr, *a2vec = nullptr;
  int64_t* t_dev = nullptr;
  int a2_total = 0;
  std::vector<int> a2_off;

  // ---- helpers (all respect `dry`) ----
  Act new_act_in(Arena& A, int n, int dd, int h, int w, int c, DT dt) {
    Act a;
    a.n = n; a.d = dd; a.h = h; a.w = w; a.c = c; a.dt = dt;
    a.p = A.alloc(a.bytes());
    return a;
  }
  Act new_act(int n, int dd, int h, int w, int c, DT dt) { return new_act_in(arena, n, dd, h, w, c, dt); }
  // activation whose producer (a tcgen05 GEMM) also emits the column partials the consuming GroupNorm needs
  Act new_act_cs(int n, int dd, int h, int w, int c, DT dt, bool up2 = false) {
    Act a = new_act(n, dd, h, w, c, dt);
    if (prec == ECHO_PREC_BF16 && dt == BF16 && gemm_tc_colsum_row_floats(c) && (dry || tc_available())) {
      GemmArgs g;
      g.od = dd; g.oh = h; g.ow = w; g.up2 = up2 ? 1 : 0;
      a.colsum_rows = gemm_tc_colsum_rows_per_obj(g);
      a.colsum = arena.alloc_n<float>((size_t)n * a.colsum_rows * gemm_tc_colsum_row_floats(c));
    }
    return a;
  }
  void lin(const float* X, int64_t ldx, int M, const ConvW& w, float* Y, int64_t ldy, int in_act, int act, cudaStream_t s) {
    if (dry) return;
    LinArgs a;
    a.X = X; a.ldx = ldx; a.M = M; a.K = w.cin; a.nout = w.cout; a.W = w.w; a.bias = w.b; a.Y = Y; a.ldy = ldy;
    a.in_act = in_act; a.act = act;
    linear_auto(a, s);
  }
  // implicit-GEMM conv / linear over a channels-last activation
  void contract(const Act& x, const ConvW& w, int k, int stride_hw, const float* rowvec, int64_t ld_rowvec, const Act* res,
                const Act& out, cudaStream_t s) {
    void* scratch = nullptr;
    void* splitk_ws = nullptr;
    if (stride_hw == 2 && prec == ECHO_PREC_BF16 && x.dt == BF16) scratch = arena.alloc(x.bytes());
    // split-precision mode: the fp32 input as hi + lo bf16 halves (every voxel-major contraction whose shape the tcgen05 kernel takes)
    const bool x3 = prec == ECHO_PREC_X3 && w.wb_lo && x.dt == F32 && out.dt == F32 && x.c % 16 == 0 && w.cout % 32 == 0;
    __nv_bfloat16 *a_hi = nullptr, *a_lo = nullptr;
    if (x3) {
      const size_t half = (size_t)x.rows() * x.c * sizeof(__nv_bfloat16);
      a_hi = (__nv_bfloat16*)arena.alloc(half);
      a_lo = (__nv_bfloat16*)arena.alloc(half);
      if (stride_hw == 2) scratch = arena.alloc(2 * half);   // space-to-depth copies of both halves
    }
    if (k == 3 && stride_hw == 1 && prec == ECHO_PREC_BF16 && x.dt == BF16 && out.dt == BF16 && (!res || res->dt == BF16)) {
      GemmArgs q;   // coarse levels: offer a workspace so the plan may split the taps over more CTAs
      q.n = x.n; q.od = out.d; q.oh = out.h; q.ow = out.w; q.kd = q.kh = q.kw = 3; q.sh = q.sw = 1; q.cout = w.cout;
      const size_t wsb = gemm_tc_splitk_ws_bytes(q);
      if (wsb) splitk_ws = arena.alloc(wsb);
    }
    if (dry) return;
    GemmArgs g;
    g.scratch = scratch;
    g.A = x.p; g.a_dt = x.dt; g.n = x.n; g.d = x.d; g.h = x.h; g.w = x.w; g.cin = x.c; g.lda = x.c;
    g.od = out.d; g.oh = out.h; g.ow = out.w;
    g.kd = g.kh = g.kw = k; g.sd = 1; g.sh = g.sw = stride_hw; g.pd = g.ph = g.pw = k / 2;
    g.W = w.w; g.w_dt = F32; g.w_stride_n = (int64_t)w.taps * w.cin; g.cout = w.cout;
    g.bias = w.b; g.rowvec = rowvec; g.ld_rowvec = ld_rowvec;
    if (res) { g.res = res->p; g.res_dt = res->dt; g.ld_res = res->c; }
    g.out = out.p; g.out_dt = out.dt; g.ldo = out.c;
    ECHO_CHECK(w.cin == x.c && w.cout == out.c && w.taps == k * k * k, "contract: weight/activation mismatch (cin %d vs %d, cout %d vs %d)",
               w.cin, x.c, w.cout, out.c);
    if (prec == ECHO_PREC_BF16 && w.wb && x.dt == BF16 && tc_available()) {
      GemmArgs t = g;
      t.W = w.wb; t.w_dt = BF16;
      t.colsum = out.colsum;
      t.splitk_ws = splitk_ws;
      if (gemm_tc_supported(t)) { gemm_tc(t, s); return; }
    }
    if (x3) {
      GemmArgs t = g;
      t.A = a_hi; t.A_lo = a_lo; t.a_dt = BF16;
      t.W = w.wb; t.W_lo = w.wb_lo; t.w_dt = BF16;
      ECHO_CHECK(gemm_tc_supported(t), "contract: split-precision contraction cin=%d cout=%d k=%d stride=%d not supported by the tcgen05 kernel",
                 x.c, w.cout, k, stride_hw);
      split_bf16((const float*)x.p, x.rows() * x.c, a_hi, a_lo, s);
      gemm_tc(t, s);
      return;
    }
    ECHO_CHECK(!out.colsum, "contract: column statistics were requested but the contraction left the tcgen05 path");
    gemm_simt(g, s);
  }
  Act gn(const Act& x, const NormW& nw, float eps, bool silu, DT odt, cudaStream_t s) {
    Act o = new_act(x.n, x.d, x.h, x.w, x.c, odt);
    if (x.colsum && gn_apply_cs_supported(x, nullptr, o)) {   // statistics from the producer's column partials: one kernel
      if (!dry) gn_apply_cs(x, nullptr, nw.g, nw.b, 32, eps, silu, o, nullptr, s);
      return o;
    }
    float* stats = arena.alloc_n<float>((size_t)x.n * 32 * 2);
    float* partial = arena.alloc_n<float>(gn_partial_floats(x, 32));
    if (!dry) {
      gn_stats(x, 32, eps, stats, partial, s);
      gn_apply(x, stats, nw.g, nw.b, 32, silu, o, s);
    }
    return o;
  }

  // ResBlock._forward (openai_model_3d.py:294-314).  `ca`/`cb` != null: x is the (not yet written) channel concat
  // [ca | cb] of a skip connection; the first GroupNorm reads both halves in place and writes x on the side.
  Act res_block(const Act& x, const ResW& r, int n_local, cudaStream_t s, const Act* ca = nullptr, const Act* cb = nullptr) {
    Act out = new_act_cs(x.n, x.d, x.h, x.w, r.cout, adt);
    const size_t m = arena.mark();
    Act a1;
    if (ca) {
      a1 = new_act(x.n, x.d, x.h, x.w, x.c, adt);
      if (!dry) gn_apply_cs(*ca, cb, r.n1.g, r.n1.b, 32, 1e-5f, true, a1, &x, s);
    } else {
      a1 = gn(x, r.n1, 1e-5f, true, adt, s);
    }
    // 1x1 skip conv (Cin != Cout): independent of the GN -> conv -> GN chain below, so it goes to its own stream right
    // after the first GroupNorm (which completes x when x is a fused concat) and runs in the tails of conv1 / GN2
    Act sk;
    bool forked = false;
    if (r.has_skip) {
      static const bool no_side2 = getenv("ECHO_NO_SIDE2") != nullptr;
      sk = new_act(x.n, x.d, x.h, x.w, r.cout, adt);
      forked = !dry && !no_side2 && side2;
      if (forked) {
        ECHO_CUDA(cudaEventRecord(ev_fork2, s));
        ECHO_CUDA(cudaStreamWaitEvent(side2, ev_fork2, 0));
        contract(x, r.skip, 1, 1, nullptr, 0, nullptr, sk, side2);
        ECHO_CUDA(cudaEventRecord(ev_join2, side2));
      } else {
        contract(x, r.skip, 1, 1, nullptr, 0, nullptr, sk, s);
      }
    }
    Act h1 = new_act_cs(x.n, x.d, x.h, x.w, r.cout, adt);
    contract(a1, r.c1, 3, 1, embout + r.emb_off, plan.emb_total, nullptr, h1, s);
    Act a2 = gn(h1, r.n2, 1e-5f, true, adt, s);
    if (forked) ECHO_CUDA(cudaStreamWaitEvent(s, ev_join2, 0));
    contract(a2, r.c2, 3, 1, nullptr, 0, r.has_skip ? &sk : &x, out, s);
    arena.release(m);
    (void)n_local;
    return out;
  }

  // SpatialTransformer3D + BasicTransformerBlock (attention.py:334-351, 237-245)
  Act transformer(const Act& x, const AttnW& a, int attn_index, cudaStream_t s) {
    Act out = new_act_cs(x.n, x.d, x.h, x.w, x.c, adt);
    const size_t m = arena.mark();
    const int C = a.C, tokens = (int)x.voxels();
    const int64_t rows = x.rows();
    Act xn = gn(x, a.norm, 1e-6f, false, adt, s);
    Act t0 = new_act(x.n, x.d, x.h, x.w, C, adt);
    contract(xn, a.proj_in, 1, 1, nullptr, 0, nullptr, t0, s);
    // self-attention
    Act l1 = new_act(x.n, x.d, x.h, x.w, C, adt);
    if (!dry) layer_norm(t0.p, t0.dt, rows, C, a.ln1.g, a.ln1.b, 1e-5f, l1.p, l1.dt, s);
    Act o = new_act(x.n, x.d, x.h, x.w, C, adt);
    if (prec == ECHO_PREC_BF16 && adt == BF16 && a.qkv_pad.wb && attention_pad_dh(a.dh) == 64 && (dry || attention_tc_supported(tokens, a.dh))) {
      // tcgen05 attention: q,k projected into [rows, 2*heads*64]; v projected straight into V^T per (object, head)
      const int hd = a.heads * 64;
      ConvW wqk = a.qkv_pad, wv = a.qkv_pad;
      wqk.cout = 2 * hd;
      wv.cout = hd;
      wv.w = a.qkv_pad.w + (size_t)2 * hd * C;
      wv.wb = a.qkv_pad.wb + (size_t)2 * hd * C;
      Act qk = new_act(x.n, x.d, x.h, x.w, 2 * hd, BF16);
      contract(l1, wqk, 1, 1, nullptr, 0, nullptr, qk, s);
      __nv_bfloat16* vt = (__nv_bfloat16*)arena.alloc((size_t)x.n * hd * tokens * sizeof(__nv_bfloat16));
      if (!dry) {
        GemmArgs g;
        g.A = l1.p; g.a_dt = BF16; g.n = x.n; g.d = x.d; g.h = x.h; g.w = x.w; g.cin = C; g.lda = C;
        g.od = x.d; g.oh = x.h; g.ow = x.w;
        g.W = wv.wb; g.w_dt = BF16; g.w_stride_n = C; g.cout = hd;
        g.out = vt; g.out_dt = BF16; g.ldo = hd; g.out_t = 1;
        ECHO_CHECK(gemm_tc_supported(g), "attention: V^T projection not supported by the tcgen05 kernel");
        gemm_tc(g, s);
        attention_tc((const __nv_bfloat16*)qk.p, vt, x.n, tokens, a.heads, a.dh, (__nv_bfloat16*)o.p, s);
      }
    } else if (prec == ECHO_PREC_BF16 && adt == BF16 && a.qkv_pad.wb && attention_bf16_supported(tokens, a.dh)) {
      Act qkv = new_act(x.n, x.d, x.h, x.w, a.qkv_pad.cout, BF16);
      contract(l1, a.qkv_pad, 1, 1, nullptr, 0, nullptr, qkv, s);
      if (!dry) attention_bf16((const __nv_bfloat16*)qkv.p, x.n, tokens, a.heads, a.dh, (__nv_bfloat16*)o.p, s);
    } else {
      Act qkv = new_act(x.n, x.d, x.h, x.w, 3 * C, F32);
      contract(l1, a.qkv, 1, 1, nullptr, 0, nullptr, qkv, s);
      float* ws = arena.alloc_n<float>(attention_f32_ws_floats(x.n, tokens, a.heads));
      if (adt == F32) {
        if (!dry) attention_f32((const float*)qkv.p, x.n, tokens, a.heads, a.dh, ws, (float*)o.p, s);
      } else {
        float* of = arena.alloc_n<float>((size_t)rows * C);
        if (!dry) {
          attention_f32((const float*)qkv.p, x.n, tokens, a.heads, a.dh, ws, of, s);
          convert(of, F32, o.p, o.dt, rows * C, s);
        }
      }
    }
    // t1 = to_out(attn1) + t0  +  [attn2: to_out(to_v(context)) broadcast over tokens — one context token]
    Act t1 = new_act(x.n, x.d, x.h, x.w, C, adt);
    contract(o, a.attn1_out, 1, 1, a2vec + a2_off[attn_index], a2_total, &t0, t1, s);
    // feed-forward (GEGLU)
    Act l3 = new_act(x.n, x.d, x.h, x.w, C, adt);
    if (!dry) layer_norm(t1.p, t1.dt, rows, C, a.ln3.g, a.ln3.b, 1e-5f, l3.p, l3.dt, s);
    Act gg = new_act(x.n, x.d, x.h, x.w, 4 * C, adt);
    bool fused = false;
    if (prec == ECHO_PREC_BF16 && adt == BF16 && a.ff1_geglu.wb && tc_available()) {
      fused = true;
      if (!dry) {   // ff1 GEMM with the GEGLU applied in its epilogue: f1 (rows x 8C) is never written
        GemmArgs g;
        g.A = l3.p; g.a_dt = BF16; g.n = x.n; g.d = x.d; g.h = x.h; g.w = x.w; g.cin = C; g.lda = C;
        g.od = x.d; g.oh = x.h; g.ow = x.w;
        g.W = a.ff1_geglu.wb; g.w_dt = BF16; g.w_stride_n = C; g.cout = 8 * C; g.bias = a.ff1_geglu.b;
        g.out = gg.p; g.out_dt = BF16; g.ldo = 4 * C; g.epi = 1;
        if (gemm_tc_supported(g)) gemm_tc(g, s);
        else fused = false;
      }
    }
    if (!fused) {
      Act f1 = new_act(x.n, x.d, x.h, x.w, 8 * C, adt);
      contract(l3, a.ff1, 1, 1, nullptr, 0, nullptr, f1, s);
      if (!dry) geglu(f1.p, f1.dt, rows, 4 * C, gg.p, gg.dt, s);
    }
    Act t2 = new_act(x.n, x.d, x.h, x.w, C, adt);
    contract(gg, a.ff2, 1, 1, nullptr, 0, &t1, t2, s);
    contract(t2, a.proj_out, 1, 1, nullptr, 0, &x, out, s);
    arena.release(m);
    return out;
  }

  // shape_embeddings stack on local objects (openai_model_3d.py:757-764): x_cl (n,16,16,16,3) f32 -> codes (n,64)
  void embed(const Act& xcl, float* codes_out, cudaStream_t s, size_t base_mark = 0) {
    Arena& A = side_arena;
    A.release(base_mark);
    Act c0 = new_act_in(A, xcl.n, xcl.d, xcl.h, xcl.w, 32, F32);
    contract(xcl, se_conv0, 3, 1, nullptr, 0, nullptr, c0, s);
    Act p0 = new_act_in(A, xcl.n, xcl.d / 2, xcl.h / 2, xcl.w / 2, 32, F32);
    if (!dry) maxpool3d(c0, 2, 2, p0, s);
    Act c1 = new_act_in(A, p0.n, p0.d, p0.h, p0.w, 64, F32);
    contract(p0, se_conv2, 3, 1, nullptr, 0, nullptr, c1, s);
    Act p1 = new_act_in(A, c1.n, (c1.d - 2) / 4 + 1, (c1.h - 2) / 4 + 1, (c1.w - 2) / 4 + 1, 64, F32);
    if (!dry) maxpool3d(c1, 2, 4, p1, s);
    float* flat = A.alloc_n<float>((size_t)p1.rows() * 64);
    if (!dry) flatten_ncdhw(p1, flat, s);
    ECHO_CHECK((int)p1.voxels() * 64 == se_lin.cin, "shape_embeddings: flatten width %d != %d", (int)p1.voxels() * 64, se_lin.cin);
    lin(flat, se_lin.cin, xcl.n, se_lin, codes_out, d.gconv_dim, 0, 0, s);
  }

  // everything after the codes are known.  out: e_t (ddim_index < 0) or x_prev, NCDHW f32, local objects.
  void trunk(const echo_graph* g, const float* x_local, const Act& xcl, int obj_begin, int n_local, const float* codes_all,
             const float* uc_all, const int64_t* t_all, int ddim_index, float* out_local, cudaStream_t s,
             cudaStream_t codes_stream = nullptr) {
    const int N = g->n_nodes, T = g->n_triples, mc = d.model_channels, E = 4 * mc, ctx = d.context_dim, gd = d.gconv_dim;
    const int nd = ctx + gd + (d.enable_t_emb ? gd : 0);
    // timestep embedding + time MLP for ALL nodes (the GCN needs every node's t_emb)
    if (!dry) timestep_embedding_tab(t_all, freqs, N, mc, temb, s);
    lin(temb, mc, N, plan.time0, e1, E, 0, 2, s);
    lin(e1, E, N, plan.time2, emb, E, 0, 0, s);
    // ---- fork: the echo chain (shape codes -> node features -> 5-layer GCN -> attn2 vectors) is only needed by the
    //      first SpatialTransformer (input block 4), so it runs on a side stream under the first ~1 ms of the trunk ----
    cudaStream_t q = s;
    if (!dry) {
      ECHO_CUDA(cudaEventRecord(ev_fork, s));
      ECHO_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
      if (codes_stream && codes_stream != s) {   // the all-gathered codes arrive on another stream: only the echo chain waits
        ECHO_CUDA(cudaEventRecord(ev_codes, codes_stream));
        ECHO_CUDA(cudaStreamWaitEvent(side, ev_codes, 0));
      }
      q = side;
    }
    if (!codes_all) {   // unsharded: embed here (sharded callers all-gathered the codes between embed and trunk)
      embed(xcl, codes, q);
      codes_all = codes;
    }
    // node features [obj_embed | code | t_emb] (openai_model_3d.py:808-812)
    if (!dry) {
      copy_cols(uc_all, ctx, N, ctx, node, nd, q);
      copy_cols(codes_all, gd, N, gd, node + ctx, nd, q);
    }
    if (d.enable_t_emb) lin(emb, E, N, time_emb_lin, node + ctx + gd, nd, 0, 0, q);
    if (!dry) {
      if (T > 0) embedding_rows(pred_table, 2 * gd, g->triples, 3, 1, T, pred, 2 * gd, q);
      gcn.forward(g, node, pred, latent, nullptr, q, batch_stats);
    }
    const float* lat_loc = latent + (size_t)obj_begin * ctx;
    {
      ConvW v = plan.v2_stack;
      lin(lat_loc, ctx, n_local, v, v2, plan.v2_total, 0, 0, q);
    }
    {
      int ai = 0;
      auto a2 = [&](const AttnW& a) {
        lin(v2 + a.v2_off, plan.v2_total, n_local, a.attn2_out, a2vec + a2_off[ai], a2_total, 0, 0, q);
        ++ai;
      };
      for (auto& b : plan.in_blocks) if (b.attn) a2(b.at);
      a2(plan.mid_at);
      for (auto& b : plan.out_blocks) if (b.attn) a2(b.at);
    }
    bool joined = dry;
    if (!dry) ECHO_CUDA(cudaEventRecord(ev_join, side));
    auto join = [&]() {
      if (!joined) { ECHO_CUDA(cudaStreamWaitEvent(s, ev_join, 0)); joined = true; }
    };
    // per-object time-embedding projections of the local objects (main stream: the first ResBlock needs them)
    const float* emb_loc = emb + (size_t)obj_begin * E;
    if (!dry) silu_f32(emb_loc, emb_act, (int64_t)n_local * E, s);   // shared by all 17 emb_layers
    lin(emb_act, E, n_local, plan.emb_stack, embout, plan.emb_total, 0, 0, s);
    // ---- UNet trunk ----
    const size_t m0 = arena.mark();
    std::vector<Act> hs;
    Act h = xcl;
    int ai = 0;
    for (auto& b : plan.in_blocks) {
      if (b.kind == BlockW::CONV_IN) {
        if (prec == ECHO_PREC_BF16 && plan.stem_pad.wb && adt == BF16) {
          // stem on the tensor cores: latent re-laid as bf16 channels-last with 3 -> 16 zero-padded channels
          Act x16 = new_act(h.n, h.d, h.h, h.w, 16, BF16);
          if (!dry) ncdhw_to_cl_pad16(x_local, h.n, d.in_channels, h.voxels(), (__nv_bfloat16*)x16.p, s);
          Act o = new_act_cs(h.n, h.d, h.h, h.w, b.conv.cout, adt);
          contract(x16, plan.stem_pad, 3, 1, nullptr, 0, nullptr, o, s);
          h = o;
        } else {
          Act o = new_act(h.n, h.d, h.h, h.w, b.conv.cout, adt);   // fp32 parity mode: SIMT path (K = 81)
          contract(h, b.conv, 3, 1, nullptr, 0, nullptr, o, s);
          h = o;
        }
      } else if (b.kind == BlockW::RES) {
        h = res_block(h, b.res, n_local, s);
        if (b.attn) { join(); h = transformer(h, b.at, ai++, s); }
      } else {   // Downsample: Conv3d k3 stride (1,2,2) pad 1 (openai_model_3d.py:188-192)
        Act o = new_act_cs(h.n, h.d, (h.h + 2 - 3) / 2 + 1, (h.w + 2 - 3) / 2 + 1, b.conv.cout, adt);
        contract(h, b.conv, 3, 2, nullptr, 0, nullptr, o, s);
        h = o;
      }
      hs.push_back(h);
    }
    h = res_block(h, plan.mid0, n_local, s);
    join();
    h = transformer(h, plan.mid_at, ai++, s);
    h = res_block(h, plan.mid2, n_local, s);
    for (auto& b : plan.out_blocks) {
      Act sk = hs.back();
      hs.pop_back();
      Act cat = new_act(h.n, h.d, h.h, h.w, h.c + sk.c, adt);
      if (h.colsum && gn_apply_cs_supported(h, &sk, cat)) {   // concat fused into the ResBlock's first GroupNorm
        h = res_block(cat, b.res, n_local, s, &h, &sk);
      } else {
        if (!dry) concat_channels(h, sk, cat, s);
        h = res_block(cat, b.res, n_local, s);
      }
      if (b.attn) h = transformer(h, b.at, ai++, s);
      if (b.up) {   // nearest x(1,2,2) then Conv3d k3 (openai_model_3d.py:150-157)
        if (prec == ECHO_PREC_BF16 && adt == BF16 && b.up_fold.wb) {
          // upsample folded into the conv: four output phases, 12 taps each, on the low-resolution tensor
          Act o = new_act_cs(h.n, h.d, h.h * 2, h.w * 2, b.conv.cout, adt, true);
          if (!dry) {
            GemmArgs g;
            g.A = h.p; g.a_dt = BF16; g.n = h.n; g.d = h.d; g.h = h.h; g.w = h.w; g.cin = h.c; g.lda = h.c;
            g.od = o.d; g.oh = o.h; g.ow = o.w; g.up2 = 1;
            g.kd = g.kh = g.kw = 3; g.pd = g.ph = g.pw = 1;
            g.W = b.up_fold.wb; g.w_dt = BF16; g.w_stride_n = (int64_t)48 * h.c; g.cout = b.conv.cout; g.bias = b.up_fold.b;
            g.out = o.p; g.out_dt = BF16; g.ldo = o.c; g.colsum = o.colsum;
            ECHO_CHECK(gemm_tc_supported(g), "upsample conv: folded contraction not supported");
            gemm_tc(g, s);
          }
          h = o;
        } else {
          Act up = new_act(h.n, h.d, h.h * 2, h.w * 2, h.c, adt);
          if (!dry) upsample_hw2(h, up, s);
          Act o = new_act_cs(up.n, up.d, up.h, up.w, b.conv.cout, adt);
          contract(up, b.conv, 3, 1, nullptr, 0, nullptr, o, s);
          h = o;
        }
      }
    }
    Act hn = gn(h, plan.out_norm, 1e-5f, true, adt, s);
    const bool pad_out = prec == ECHO_PREC_BF16 && plan.out_conv_pad.wb && hn.dt == BF16;
    Act e = new_act(h.n, h.d, h.h, h.w, pad_out ? plan.out_conv_pad.cout : d.out_channels, F32);
    if (pad_out) {
      contract(hn, plan.out_conv_pad, 3, 1, nullptr, 0, nullptr, e, s);        // tcgen05, 32 padded output channels
    } else if (conv3d_small_cout_supported(hn.c, d.out_channels, plan.out_conv.taps)) {
      if (!dry) conv3d_small_cout(hn, plan.out_conv.w, plan.out_conv.b, d.out_channels, (float*)e.p, s);
    } else {
      contract(hn, plan.out_conv, 3, 1, nullptr, 0, nullptr, e, s);
    }
    if (!dry) {
      if (ddim_index == -1) cl_to_ncdhw(e.p, F32, e.n, d.out_channels, e.voxels(), e.c, out_local, s);
      else if (ddim_index == ECHO_INDEX_FROM_DEVICE)
        ddim_update(x_local, e.p, F32, true, e.n, d.out_channels, e.voxels(), e.c, d_coef, out_local, s, d_index);
      else ddim_update(x_local, e.p, F32, true, e.n, d.out_channels, e.voxels(), e.c, d_coef + 4 * ddim_index, out_local, s);
    }
    arena.release(m0);
  }

  // shape_embeddings on local objects with every temporary (incl. the channels-last copy) in the side arena, so it may
  // run on another stream while the trunk of the same step is already using the main arena
  void embed_entry(const float* x_local, int n_local, float* codes_out, cudaStream_t s) {
    side_arena.release(0);
    const int L = d.latent_size;
    Act xcl = new_act_in(side_arena, n_local, L, L, L, d.in_channels, F32);
    if (!dry) ncdhw_to_cl(x_local, n_local, d.in_channels, (int64_t)L * L * L, xcl.p, F32, s);
    embed(xcl, codes_out, s, side_arena.mark());
  }

  void run(const echo_graph* g, const float* x_local, int obj_begin, int n_local, const float* codes_all, const float* uc_all,
           const int64_t* t_all, int ddim_index, float* out_local, cudaStream_t s, cudaStream_t codes_stream = nullptr) {
    ECHO_CHECK(g && g->n_nodes <= d.max_nodes && g->n_triples <= d.max_triples, "shape: graph exceeds handle capacity");
    // nn.Embedding would raise an index error (denoise_net.py:764 / openai_model_3d.py:807)
    ECHO_CHECK(g->n_triples == 0 || (g->p_min >= 0 && g->p_max < pred_rows), "shape: predicate ids [%lld, %lld] outside pred_embeddings (%d rows)",
               (long long)g->p_min, (long long)g->p_max, pred_rows);
    ECHO_CHECK(n_local >= 0 && n_local <= d.max_local_nodes && obj_begin >= 0 && obj_begin + n_local <= g->n_nodes,
               "shape: bad local range [%d, %d) of %d nodes (capacity %d)", obj_begin, obj_begin + n_local, g->n_nodes, d.max_local_nodes);
    ECHO_CHECK(ddim_index < (int)h_ts.size() && ddim_index >= ECHO_INDEX_FROM_DEVICE, "shape: ddim_index %d out of range", ddim_index);
    arena.release(0);
    const int L = d.latent_size;
    Act xcl = new_act(n_local, L, L, L, d.in_channels, F32);
    if (!dry && n_local > 0) ncdhw_to_cl(x_local, n_local, d.in_channels, (int64_t)L * L * L, xcl.p, F32, s);
    if (!codes_all) ECHO_CHECK(n_local == g->n_nodes && obj_begin == 0, "shape: codes of all nodes are required when the trunk is sharded");
    if (n_local == 0) return;
    trunk(g, x_local, xcl, obj_begin, n_local, codes_all, uc_all, t_all, ddim_index, out_local, s, codes_stream);
  }
};

namespace echo {

// make_beta_schedule('linear') = linspace(sqrt(ls), sqrt(le), T)^2 in float64 (ldm_diffusion_util.py:44-47);
// alphas_cumprod registered as fp32 (echo2shape.py:185-190); DDIM: c = T // S, timesteps range(0, T, c) + 1
// (ldm_diffusion_util.py:70-79); a_prev[0] = ac[0] (:88); sigma = 0 for eta = 0.
// coef = n x 4 [sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)], ts = the n DDIM timesteps.  Host-only (also exported
// as echo_debug_ddim_schedule for CPU tests).
void ddim_schedule(int T, int S, float linear_start, float linear_end, std::vector<float>& coef, std::vector<int32_t>& ts) {
  ECHO_CHECK(T > 0 && S > 0 && S <= T, "shape: bad schedule T=%d S=%d", T, S);
  std::vector<float> ac(T);
  const double a = sqrt((double)linear_start), b = sqrt((double)linear_end);
  double cp = 1.0;
  for (int i = 0; i < T; ++i) {
    // torch.linspace(a, b, T, float64): start + i*step for the first half, end - (T-1-i)*step for the second
    const double step = (b - a) / (double)(T - 1);
    const double v = (i < T / 2) ? a + step * i : b - step * (T - 1 - i);
    const double beta = v * v;
    cp *= (1.0 - beta);
    ac[i] = (float)cp;
  }
  const int c = T / S;
  ts.clear();
  for (int t = 0; t < T; t += c) ts.push_back(t + 1);
  const int n = (int)ts.size();
  coef.assign((size_t)n * 4, 0.f);
  for (int i = 0; i < n; ++i) {
    ECHO_CHECK(ts[i] < T, "shape: DDIM timestep %d out of range for %d-step schedule (S must divide T)", ts[i], T);
    const float a_t = ac[ts[i]];
    const float a_prev = i == 0 ? ac[0] : ac[ts[i - 1]];
    // samplers/ddim.py:246-249,252-261: torch.full(fp32) of numpy values, then fp32 tensor arithmetic
    coef[4 * i + 0] = sqrtf(a_t);
    coef[4 * i + 1] = sqrtf(1.0f - a_t);   // np.sqrt(1. - ddim_alphas): ddim_alphas is float32 numpy, 1. - x stays float32
    coef[4 * i + 2] = sqrtf(a_prev);
    coef[4 * i + 3] = sqrtf(1.0f - a_prev - 0.0f);
  }
}

static void make_ddim_schedule(echo_shape* h) {
  ddim_schedule(h->d.timesteps, h->d.ddim_steps, h->d.linear_start, h->d.linear_end, h->h_coef, h->h_ts);
  h->d_coef = h->pool.upload(h->h_coef);
  h->d_ts = h->pool.alloc_n<int32_t>(h->h_ts.size());
  ECHO_CUDA(cudaMemcpy(h->d_ts, h->h_ts.data(), sizeof(int32_t) * h->h_ts.size(), cudaMemcpyHostToDevice));
  h->d_index = h->pool.alloc_n<int>(4);
  ECHO_CUDA(cudaMemset(h->d_index, 0, sizeof(int) * 4));
}

echo_shape* shape_create(const echo_shape_desc_t* desc, const echo_weight_t* weights, int n_weights) {
  ECHO_CHECK(desc, "shape: null desc");
  echo_shape* h = new echo_shape();
  try {
    h->d = *desc;
    const echo_shape_desc_t& d = h->d;
    ECHO_CHECK(d.max_nodes > 0 && d.max_local_nodes > 0 && d.max_local_nodes <= d.max_nodes, "shape: bad capacities");
    ECHO_CHECK(d.model_channels % 32 == 0 && d.num_levels >= 1 && d.num_levels <= 8, "shape: bad config");
    h->prec = d.precision;
    ECHO_CHECK(h->prec == ECHO_PREC_FP32 || h->prec == ECHO_PREC_BF16 || h->prec == ECHO_PREC_X3, "shape: unknown precision %d", h->prec);
    if (h->prec != ECHO_PREC_FP32 && !tc_available())
      fail(ECHO_ERR_UNSUPPORTED, "shape: ECHO_PREC_BF16 / ECHO_PREC_X3 need the sm_100a tcgen05 kernels on a B200-class device");
    h->adt = h->prec == ECHO_PREC_BF16 ? BF16 : F32;
    WeightMap wm;
    wm.load(weights, n_weights);
    cudaStream_t s = 0;
    UNetCfg cfg;
    cfg.dims = 3;
    cfg.in_channels = d.in_channels; cfg.out_channels = d.out_channels; cfg.model_channels = d.model_channels;
    cfg.channel_mult.assign(d.channel_mult, d.channel_mult + d.num_levels);
    cfg.attention_resolutions.assign(d.attention_resolutions, d.attention_resolutions + d.num_attention_resolutions);
    cfg.num_res_blocks = d.num_res_blocks; cfg.num_heads = d.num_heads; cfg.context_dim = d.context_dim;
    cfg.want_bf16 = h->prec == ECHO_PREC_BF16;
    cfg.want_x3 = h->prec == ECHO_PREC_X3;
    build_unet_plan(wm, cfg, h->pool, h->plan, s);

    const int mc = d.model_channels, E = 4 * mc, gd = d.gconv_dim, ctx = d.context_dim;
    // shape_embeddings (openai_model_3d.py:757-764) and friends; these stay fp32 in every mode (tiny)
    UNetCfg c32 = cfg;
    c32.want_bf16 = false;
    {
      // reuse the conv repacker through a throw-away plan-less Prep: replicate minimal logic here
      auto conv3 = [&](const std::string& p, int cin, int cout) {
        ConvW c;
        c.cin = cin; c.cout = cout; c.taps = 27;
        const WView& v = wm.get(p + ".weight", {cout, cin, 3, 3, 3});
        float* o = h->pool.alloc_n<float>((size_t)cout * cin * 27);
        repack_conv_weight(v.p, cout, cin, 27, o, s);
        c.w = o;
        const WView& bv = wm.get(p + ".bias", {cout});
        float* bo = h->pool.alloc_n<float>(cout);
        ECHO_CUDA(cudaMemcpyAsync(bo, bv.p, sizeof(float) * cout, cudaMemcpyDeviceToDevice, s));
        c.b = bo;
        return c;
      };
      auto linear = [&](const std::string& p, int cin, int cout) {
        ConvW c;
        c.cin = cin; c.cout = cout; c.taps = 1;
        const WView& v = wm.get(p + ".weight", {cout, cin});
        float* o = h->pool.alloc_n<float>((size_t)cout * cin);
        ECHO_CUDA(cudaMemcpyAsync(o, v.p, sizeof(float) * cout * cin, cudaMemcpyDeviceToDevice, s));
        c.w = o;
        const WView& bv = wm.get(p + ".bias", {cout});
        float* bo = h->pool.alloc_n<float>(cout);
        ECHO_CUDA(cudaMemcpyAsync(bo, bv.p, sizeof(float) * cout, cudaMemcpyDeviceToDevice, s));
        c.b = bo;
        return c;
      };
      h->se_conv0 = conv3("shape_embeddings.0", d.in_channels, 32);
      h->se_conv2 = conv3("shape_embeddings.2", 32, 64);
      const int Lp = ((d.latent_size / 2) - 2) / 4 + 1;
      h->se_lin = linear("shape_embeddings.5", 64 * Lp * Lp * Lp, gd);
      if (d.enable_t_emb) h->time_emb_lin = linear("shape_time_emb", E, gd);
      const WView& pt = wm.get("pred_embeddings.weight");
      ECHO_CHECK(pt.shape.size() == 2 && pt.shape[1] == 2 * gd, "pred_embeddings: bad shape");
      float* o = h->pool.alloc_n<float>(pt.numel());
      ECHO_CUDA(cudaMemcpyAsync(o, pt.p, sizeof(float) * pt.numel(), cudaMemcpyDeviceToDevice, s));
      h->pred_table = o;
      h->pred_rows = (int)pt.shape[0];
    }
    // frequency table of timestep_embedding: passed by the host (computed the way torch computes it) or rebuilt
    {
      const int half = mc / 2;
      float* f = h->pool.alloc_n<float>(half);
      if (wm.has("__timestep_freqs")) {
        const WView& v = wm.get("__timestep_freqs", {half});
        ECHO_CUDA(cudaMemcpyAsync(f, v.p, sizeof(float) * half, cudaMemcpyDeviceToDevice, s));
      } else {
        std::vector<float> hf(half);
        for (int i = 0; i < half; ++i) hf[i] = expf(-logf(10000.f) * (float)i / (float)half);
        ECHO_CUDA(cudaMemcpy(f, hf.data(), sizeof(float) * half, cudaMemcpyHostToDevice));
      }
      h->freqs = f;
    }
    // echo GCN (openai_model_3d.py:766-782)
    echo_gcn_desc_t gdsc = {};
    gdsc.input_dim_obj = ctx + gd + (d.enable_t_emb ? gd : 0);
    gdsc.input_dim_pred = 2 * gd;
    gdsc.num_layers = 5;
    gdsc.hidden_dim = 4 * gd;
    gdsc.output_dim = ctx;
    gdsc.max_nodes = d.max_nodes;
    gdsc.max_triples = d.max_triples > 0 ? d.max_triples : 1;
    gdsc.bn_eps = 1e-5f;
    gdsc.keep_train_weights = d.keep_train_weights;
    h->gcn.create(wm, "shape_code_graph_cov.", gdsc, h->pool);
    make_ddim_schedule(h);

    const size_t N = d.max_nodes, T = gdsc.max_triples, NL = d.max_local_nodes;
    h->temb = h->pool.alloc_n<float>(N * mc);
    h->e1 = h->pool.alloc_n<float>(N * E);
    h->emb = h->pool.alloc_n<float>(N * E);
    h->emb_act = h->pool.alloc_n<float>(NL * E);
    h->node = h->pool.alloc_n<float>(N * gdsc.input_dim_obj);
    h->pred = h->pool.alloc_n<float>(T * 2 * gd);
    h->latent = h->pool.alloc_n<float>(N * ctx);
    h->codes = h->pool.alloc_n<float>(N * gd);
    h->embout = h->pool.alloc_n<float>(NL * h->plan.emb_total);
    h->v2 = h->pool.alloc_n<float>(NL * h->plan.v2_total);
    h->a2_total = h->plan.v2_total;
    h->a2vec = h->pool.alloc_n<float>(NL * h->a2_total);
    h->t_dev = h->pool.alloc_n<int64_t>(N);
    {
      int off = 0;
      auto add = [&](const AttnW& a) { h->a2_off.push_back(off); off += a.C; };
      for (auto& b : h->plan.in_blocks) if (b.attn) add(b.at);
      add(h->plan.mid_at);
      for (auto& b : h->plan.out_blocks) if (b.attn) add(b.at);
    }
    ECHO_CUDA(cudaStreamSynchronize(s));
    // size the workspace with a dry run at full capacity
    {
      echo_graph fake;
      fake.n_nodes = d.max_nodes;
      fake.n_triples = 0;
      h->dry = true;
      h->arena.base = nullptr;
      h->arena.cap = ~size_t(0) >> 1;
      h->arena.off = h->arena.high = 0;
      h->side_arena.base = nullptr;
      h->side_arena.cap = ~size_t(0) >> 1;
      h->side_arena.off = h->side_arena.high = 0;
      if (d.max_local_nodes == d.max_nodes) h->run(&fake, nullptr, 0, d.max_nodes, nullptr, nullptr, nullptr, -1, nullptr, s);
      else {
        h->run(&fake, nullptr, 0, d.max_local_nodes, (const float*)8, nullptr, nullptr, -1, nullptr, s);
        // embed-only path
        h->embed_entry(nullptr, d.max_local_nodes, nullptr, s);
      }
      h->dry = false;
      const size_t need = h->arena.high + (size_t(1) << 20);
      h->arena.init(need);
      const size_t need_side = h->side_arena.high + (size_t(1) << 20);
      h->side_arena.init(need_side);
    }
    ECHO_CUDA(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    ECHO_CUDA(cudaStreamCreateWithFlags(&h->side2, cudaStreamNonBlocking));
    ECHO_CUDA(cudaEventCreateWithFlags(&h->ev_fork2, cudaEventDisableTiming));
    ECHO_CUDA(cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming));
    ECHO_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    ECHO_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    ECHO_CUDA(cudaEventCreateWithFlags(&h->ev_codes, cudaEventDisableTiming));
    return h;
  } catch (...) {
    h->arena.destroy();
    h->side_arena.destroy();
    h->pool.destroy();
    delete h;
    throw;
  }
}

void shape_destroy(echo_shape* h) {
  if (!h) return;
  if (h->side) cudaStreamDestroy(h->side);
  if (h->side2) cudaStreamDestroy(h->side2);
  if (h->ev_fork2) cudaEventDestroy(h->ev_fork2);
  if (h->ev_join2) cudaEventDestroy(h->ev_join2);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_codes) cudaEventDestroy(h->ev_codes);
  h->arena.destroy();
  h->side_arena.destroy();
  h->pool.destroy();
  delete h;
}

void shape_forward(echo_shape* h, const echo_graph* g, const float* x, const float* uc, const int64_t* t, float* out, cudaStream_t s) {
  h->run(g, x, 0, g ? g->n_nodes : 0, nullptr, uc, t, -1, out, s);
}

// the timesteps of a sampler step: ddim_timesteps[index] for every node, index from the host or from the device slot
static void fill_step_timesteps(echo_shape* h, int n_nodes, int ddim_index, cudaStream_t s) {
  if (ddim_index == ECHO_INDEX_FROM_DEVICE) fill_i64_from_slot(h->t_dev, n_nodes, h->d_ts, h->d_index, s);
  else fill_i64(h->t_dev, n_nodes, h->h_ts[ddim_index], s);
}

void shape_set_index(echo_shape* h, int ddim_index, cudaStream_t s) {
  ECHO_CHECK(ddim_index >= 0 && ddim_index < (int)h->h_ts.size(), "shape_set_index: bad ddim_index %d", ddim_index);
  set_i32(h->d_index, ddim_index, s);
}

void shape_step(echo_shape* h, const echo_graph* g, const float* x, const float* uc, int ddim_index, float* x_prev, cudaStream_t s) {
  ECHO_CHECK(g && (ddim_index == ECHO_INDEX_FROM_DEVICE || (ddim_index >= 0 && ddim_index < (int)h->h_ts.size())), "shape_step: bad ddim_index %d",
             ddim_index);
  ECHO_CHECK(!h->batch_stats, "shape_step: the sampler step runs on running statistics; switch echo_shape_set_batch_stats off first");
  fill_step_timesteps(h, g->n_nodes, ddim_index, s);
  h->run(g, x, 0, g->n_nodes, nullptr, uc, h->t_dev, ddim_index, x_prev, s);
}

void shape_set_batch_stats(echo_shape* h, bool on) {
  ECHO_CHECK(h, "shape_set_batch_stats: null handle");
  ECHO_CHECK(!on || h->d.keep_train_weights, "shape_set_batch_stats: the handle was created without keep_train_weights");
  h->batch_stats = on;
}

void shape_embed(echo_shape* h, const float* x_local, int n_local, float* codes_out, cudaStream_t s) {
  ECHO_CHECK(n_local >= 0 && n_local <= h->d.max_local_nodes, "shape_embed: n_local %d exceeds capacity", n_local);
  if (n_local == 0) return;
  h->embed_entry(x_local, n_local, codes_out, s);
}

void shape_trunk(echo_shape* h, const echo_graph* g, const float* x_local, int obj_begin, int n_local, const float* codes_all,
                 const float* uc_all, const int64_t* t_all, int ddim_index, float* out_local, cudaStream_t s, cudaStream_t codes_stream) {
  ECHO_CHECK(codes_all, "shape_trunk: codes_all is required");
  const int64_t* t_use = t_all;
  if (!t_all) {
    ECHO_CHECK(ddim_index >= 0 || ddim_index == ECHO_INDEX_FROM_DEVICE, "shape_trunk: timesteps_all or ddim_index required");
    ECHO_CHECK(ddim_index < (int)h->h_ts.size(), "shape_trunk: bad ddim_index %d", ddim_index);
    fill_step_timesteps(h, g->n_nodes, ddim_index, s);
    t_use = h->t_dev;
  }
  h->run(g, x_local, obj_begin, n_local, codes_all, uc_all, t_use, ddim_index, out_local, s, codes_stream);
}

}  // namespace echo

namespace echo {
const float* shape_latent(const echo_shape* h) { return h->latent; }
int shape_context_dim(const echo_shape* h) { return h->d.context_dim; }
void shape_tables(const echo_shape* h, const std::vector<float>** c, const std::vector<int32_t>** t) {
  *c = &h->h_coef;
  *t = &h->h_ts;
}
}  // namespace echo
