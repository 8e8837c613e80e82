// Layout-branch denoiser step: UNet1DModel.forward (denoise_net.py:773-806) + the DDPM ancestral update
// (diffusion_ddpm.py:220-264, 296-309).
//
// The layout "1-D UNet" runs on signals of length 1 (box_t (N,8) -> (N,8,1), denoise_net.py:788,796): objects are
// the batch dimension.  So every Conv1d(k=3, p=1[, stride 2]) is a matrix product with its centre tap, Up/Down
// sampling keeps length 1, and self-/cross-attention over one token is to_out(to_v(.)) (softmax of one logit == 1).
// What remains is a chain of ~150 few-row contractions (rows = nodes) bound by streaming ~115 M live weights from
// HBM once per step, plus GroupNorm/LayerNorm/GEGLU over (N, C) rows; the only cross-object coupling is the echo
// GCN.  Activations are fp32 row-major (N, C) throughout; ECHO_PREC_BF16 streams bf16 copies of the weights.
#include "layout_mk.cuh"
#include "unet.cuh"

#include <math.h>
#include <stdlib.h>

#include <deque>

using namespace echo;

struct echo_layout {
  echo_layout_desc_t d;
  DevPool pool;
  UNetPlan plan;
  Gcn gcn;
  Arena arena;
  int prec = ECHO_PREC_FP32;
  bool batch_stats = false;   // box_graph_cov normalises with the statistics of the batch (model.train() forward values)
  float* lin_scratch[2] = {nullptr, nullptr};   // handles for more than 64 nodes: rows of a Linear's prologue output (linear_auto)
  size_t lin_scratch_floats = 0;
  ConvW box_emb, time_emb_lin;
  const float* pred_table = nullptr;
  int pred_rows = 0;   // rows of pred_embeddings (16)
  const float* freqs = nullptr;
  std::vector<float> h_tab;
  float* d_tab = nullptr;
  float *temb = nullptr, *e1 = nullptr, *emb = nullptr, *emb_act = nullptr, *node = nullptr, *pred = nullptr, *latent = nullptr;
  float *embout = nullptr, *a2vec = nullptr, *eps = nullptr;
  int64_t* t_dev = nullptr;
  // CUDA-graph replay of one DDPM iteration (layout_step)
  float *sx = nullptr, *sobj = nullptr, *snoise = nullptr, *sout = nullptr;
  int* t_slot = nullptr;
  cudaGraphExec_t gexec = nullptr;
  uint64_t gkey_id = 0;
  int gkey_n = -1;
  int64_t graph_launches = 0;
  bool graph_failed = false;
  std::vector<int> a2_off;
  int a2_total = 0;
  // One token per object makes attention the identity on V (softmax over one key == 1), so each attention is two Linears
  // with nothing in between: they are multiplied once at creation.  attn1: to_out . to_v  (C x C per block);
  // attn2: [to_out_i . to_v_i] of all blocks stacked (a2_total x context_dim) -- one launch instead of 1 + 11.
  std::vector<ConvW> attn1_fused;   // by attention index
  ConvW attn2_fused;
  // the stacked emb_layers projection (46 M weights, the largest layer of the step) depends on the time embedding only,
  // so it runs on a second stream beside the GCN chain and joins at the first ResBlock (a parallel branch of the graph)
  cudaStream_t side = nullptr;
  cudaStream_t cap = nullptr;   // capture origin of the replayed graph: independent of the caller's stream (legacy stream 0 cannot capture)
  int64_t graph_replays = 0;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int N = 0;   // rows of the current call

  // ---- persistent executor (layout_mk.cu): the step recorded as a program of stages, one cooperative kernel per iteration ----
  bool mk_ok = false;
  int mk_ctas = 0;
  int mk_N = -1, mk_T = -1;
  std::vector<MkOp> mk_ops;
  std::vector<MkStage> mk_stages;
  MkOp* d_mk_ops = nullptr;
  MkStage* d_mk_stages = nullptr;
  std::vector<MkFetch> mk_fetch;       // every CTA's weight fetches in consumption order (mk_build_fetch)
  std::vector<int> mk_fetch_off;
  // the time path is a pure function of t and the weights: emb = time_embed(timestep_embedding(t)), the 22 stacked emb_layers
  // projections of SiLU(emb) and box_time_emb(emb) are tabulated once per handle for every t of the schedule (same kernels as
  // forward()), which takes 3 stages, 22 background ops and 26 % of the weight bytes out of every step
  float* ttab_emb = nullptr;    // [time_num][emb_total]
  float* ttab_node = nullptr;   // [time_num][gconv_dim]
  bool ttab_on = false;
  void build_time_table(cudaStream_t s);
  // proj_out(t1 + ff2(g) + b2) = [P | P F2] [t1 ; g] + (P b2 + bP): one contraction instead of two dependent ones
  std::vector<ConvW> ffo_fused;
  std::vector<MkStageLite> mk_stages_lite;
  MkStageLite* d_mk_stages_lite = nullptr;
  MkFetch* d_mk_fetch = nullptr;
  int* d_mk_fetch_off = nullptr;
  static constexpr int MK_MAX_FETCH = 1 << 16;
  unsigned* d_mk_counters = nullptr;   // [MK_MAX_STAGES] stage counters, [MK_MAX_BG] background counters, [1] epoch
  std::vector<float*> mk_bufs;         // activation buffers of the program, allocated once at capacity in program order
  size_t mk_buf_i = 0;
  int64_t mk_steps = 0;
  static constexpr int MK_MAX_BG = 64, MK_MAX_OPS = 1024;

  float* buf(int C) { return arena.alloc_n<float>((size_t)N * C); }
  float* mkb(int C) {
    if (mk_buf_i < mk_bufs.size()) return mk_bufs[mk_buf_i++];
    float* b = pool.alloc_n<float>((size_t)d.max_nodes * C);
    mk_bufs.push_back(b);
    ++mk_buf_i;
    return b;
  }
  void build_mk(int Nn, int T);
  void step_mk(const echo_graph* g, const float* x_t, const float* obj_embed, int t, const float* noise, float* x_prev, cudaStream_t s);

  // input of a fused Linear: (possibly concatenated) rows + the elementwise op that precedes the Linear in the network
  struct In {
    const float* X = nullptr; int C = 0; int64_t ld = 0;   // ld = 0: rows are dense (C, or 2C for the GEGLU input)
    const float* X2 = nullptr; int C2 = 0;
    int pro = PRO_NONE; const NormW* nw = nullptr; float eps = 1e-5f; bool silu = false;
    int width() const { return C + C2; }
  };
  static In plain(const float* X, int C) { In i; i.X = X; i.C = C; return i; }

  void lin(const In& in, const ConvW& w, float* Y, int64_t ldy, const float* res, int64_t ld_res, int act, cudaStream_t s,
           const float* res2 = nullptr, int64_t ld_res2 = 0) {
    LinArgs a;
    a.X = in.X; a.ldx = in.ld ? in.ld : (in.pro == PRO_GEGLU ? 2 * in.C : in.C); a.M = N; a.K = w.cin; a.nout = w.cout; a.bias = w.b; a.Y = Y; a.ldy = ldy;
    if (in.X2) { a.X2 = in.X2; a.ldx2 = in.C2; a.K1 = in.C; }
    a.res = res; a.ld_res = ld_res; a.res2 = res2; a.ld_res2 = ld_res2; a.act = act;
    a.pro = in.pro;
    if (in.nw) { a.gamma = in.nw->g; a.beta = in.nw->b; }
    a.eps = in.eps; a.pro_act = in.silu ? 1 : 0;
    if (in.pro == PRO_GN) a.cpg = in.width() / 32;
    ECHO_CHECK((in.pro == PRO_GEGLU ? in.C : in.width()) == w.cin, "layout: linear input width %d != %d", in.width(), w.cin);
    // fp32 weights in every precision mode: a layout layer is a latency chain, not a bandwidth problem (461 MB per step
    // against 1.6 ms), and the fp32 loader is the shorter instruction stream -- measured 1.60 ms/step against 1.70 with
    // bf16 weights; it also keeps the layout branch on the 1e-3 parity contract in ECHO_PREC_BF16
    a.W = w.w; a.w_dt = F32;
    if (lin_scratch[0]) {   // collated batches: the prologue's output rows, one buffer per stream the forward uses
      a.scratch = lin_scratch[s == side ? 1 : 0];
      a.scratch_floats = lin_scratch_floats;
    }
    linear_auto(a, s);
  }
  void lin(const float* X, int64_t ldx, const ConvW& w, float* Y, int64_t ldy, const float* res, int64_t ld_res, int in_act, int act,
           cudaStream_t s) {
    In i = plain(X, w.cin);
    i.ld = ldx;
    if (in_act == 1) i.pro = PRO_SILU;
    lin(i, w, Y, ldy, res, ld_res, act, s);
  }

  // ResBlock._forward on length-1 signals (denoise_net.py:293-313): GroupNorm+SiLU are prologues of the two Linears
  float* res_block(const In& x, const ResW& r, cudaStream_t s) {
    float* out = buf(r.cout);
    const size_t m = arena.mark();
    In a1 = x; a1.pro = PRO_GN; a1.nw = &r.n1; a1.eps = 1e-5f; a1.silu = true;
    float* h1 = buf(r.cout);
    lin(a1, r.c1, h1, r.cout, embout + r.emb_off, plan.emb_total, 0, s);
    const float* skip = x.X;
    if (r.has_skip) {
      float* sk = buf(r.cout);
      lin(x, r.skip, sk, r.cout, nullptr, 0, 0, s);
      skip = sk;
    } else {
      ECHO_CHECK(!x.X2, "layout: identity skip over a concatenated input");
    }
    In a2 = plain(h1, r.cout); a2.pro = PRO_GN; a2.nw = &r.n2; a2.eps = 1e-5f; a2.silu = true;
    lin(a2, r.c2, out, r.cout, skip, r.cout, 0, s);
    arena.release(m);
    return out;
  }

  // SpatialTransformer1D with one token per object (attention.py:353-396, 222-245): 6 fused Linears
  float* transformer(const float* x, const AttnW& a, int ai, cudaStream_t s) {
    const int C = a.C;
    float* out = buf(C);
    const size_t m = arena.mark();
    In xn = plain(x, C); xn.pro = PRO_GN; xn.nw = &a.norm; xn.eps = 1e-6f;            // Normalize (eps 1e-6) -> proj_in
    float* t0 = buf(C);
    lin(xn, a.proj_in, t0, C, nullptr, 0, 0, s);
    In l1 = plain(t0, C); l1.pro = PRO_LN; l1.nw = &a.ln1;                             // norm1 -> attn1 = to_out(to_v(.)), one matrix
    float* t1 = buf(C);                                                                 // ... + x + [attn2: to_out(to_v(context))]
    lin(l1, attn1_fused[ai], t1, C, t0, C, 0, s, a2vec + a2_off[ai], a2_total);
    In l3 = plain(t1, C); l3.pro = PRO_LN; l3.nw = &a.ln3;                             // norm3 -> GEGLU proj
    float* f1 = buf(8 * C);
    lin(l3, a.ff1, f1, 8 * C, nullptr, 0, 0, s);
    In gg = plain(f1, 4 * C); gg.pro = PRO_GEGLU;                                       // a * gelu(g) -> ff.net.2, + x
    float* t2 = buf(C);
    lin(gg, a.ff2, t2, C, t1, C, 0, s);
    lin(plain(t2, C), a.proj_out, out, C, x, C, 0, s);
    arena.release(m);
    return out;
  }

  void forward(const echo_graph* g, const float* box_t, const float* obj_embed, const int64_t* t, float* eps_out, cudaStream_t s) {
    ECHO_CHECK(g && g->n_nodes <= d.max_nodes && g->n_triples <= d.max_triples, "layout: graph exceeds handle capacity");
    // nn.Embedding would raise an index error (denoise_net.py:764 / openai_model_3d.py:807)
    ECHO_CHECK(g->n_triples == 0 || (g->p_min >= 0 && g->p_max < pred_rows), "layout: predicate ids [%lld, %lld] outside pred_embeddings (%d rows)",
               (long long)g->p_min, (long long)g->p_max, pred_rows);
    N = g->n_nodes;
    if (N == 0) return;
    const int T = g->n_triples, mc = d.model_channels, E = 4 * mc, gd = d.gconv_dim, od = d.obj_embed_dim;
    const int nd = od + gd + (d.enable_t_emb ? gd : 0);
    arena.release(0);
    timestep_embedding_tab(t, freqs, N, mc, temb, s);
    lin(temb, mc, plan.time0, e1, E, nullptr, 0, 0, 2, s);
    lin(e1, E, plan.time2, emb, E, nullptr, 0, 0, 0, s);
    // box_messsage_passing (denoise_net.py:758-771): node = [obj_embed | box_embeddings(box_t) | box_time_emb(emb)]
    copy_cols(obj_embed, od, N, od, node, nd, s);
    {
      const int p = prec;
      prec = ECHO_PREC_FP32;   // tiny layers stay fp32
      lin(box_t, d.in_channels, box_emb, node + od, nd, nullptr, 0, 0, 0, s);
      if (d.enable_t_emb) lin(emb, E, time_emb_lin, node + od + gd, nd, nullptr, 0, 0, 0, s);
      prec = p;
    }
    // all 22 emb_layers share SiLU(emb): activate once instead of in every warp of the stacked projection; fork it
    static const bool no_side = getenv("ECHO_NO_LAYOUT_SIDE") != nullptr;
    const bool fork = side && !no_side;
    cudaStream_t q = s;
    if (fork) {
      ECHO_CUDA(cudaEventRecord(ev_fork, s));
      ECHO_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
      q = side;
    }
    silu_f32(emb, emb_act, (int64_t)N * E, q);
    lin(emb_act, E, plan.emb_stack, embout, plan.emb_total, nullptr, 0, 0, 0, q);
    if (fork) ECHO_CUDA(cudaEventRecord(ev_join, side));
    if (T > 0) embedding_rows(pred_table, 2 * gd, g->triples, 3, 1, T, pred, 2 * gd, s);
    gcn.forward(g, node, pred, latent, nullptr, s, batch_stats);
    if (fork) ECHO_CUDA(cudaStreamWaitEvent(s, ev_join, 0));
    lin(latent, d.context_dim, attn2_fused, a2vec, a2_total, nullptr, 0, 0, 0, s);   // all 11 attn2 vectors at once
    std::vector<std::pair<const float*, int>> hs;
    const float* h = nullptr;
    int hc = 0, ai = 0;
    for (auto& b : plan.in_blocks) {
      if (b.kind == BlockW::CONV_IN) {
        float* o = buf(b.conv.cout);
        lin(box_t, d.in_channels, b.conv, o, b.conv.cout, nullptr, 0, 0, 0, s);
        h = o; hc = b.conv.cout;
      } else if (b.kind == BlockW::RES) {
        h = res_block(plain(h, hc), b.res, s); hc = b.res.cout;
        if (b.attn) h = transformer(h, b.at, ai++, s);
      } else {   // Downsample: Conv1d k3 stride 2 pad 1 on length 1 -> centre tap (denoise_net.py:172-198)
        float* o = buf(b.conv.cout);
        lin(h, hc, b.conv, o, b.conv.cout, nullptr, 0, 0, 0, s);
        h = o; hc = b.conv.cout;
      }
      hs.push_back({h, hc});
    }
    h = res_block(plain(h, hc), plan.mid0, s);
    h = transformer(h, plan.mid_at, ai++, s);
    h = res_block(plain(h, plan.mid0.cout), plan.mid2, s);
    hc = plan.mid2.cout;
    for (auto& b : plan.out_blocks) {
      auto sk = hs.back();
      hs.pop_back();
      In cat = plain(h, hc);   // th.cat([h, hs.pop()], dim=1) is never materialised: the fused Linears read both sources
      cat.X2 = sk.first; cat.C2 = sk.second;
      h = res_block(cat, b.res, s); hc = b.res.cout;
      if (b.attn) h = transformer(h, b.at, ai++, s);
      if (b.up) {   // Upsample: scale_factor 1 (denoise_net.py:154) then Conv1d k3 -> centre tap
        float* o = buf(b.conv.cout);
        lin(h, hc, b.conv, o, b.conv.cout, nullptr, 0, 0, 0, s);
        h = o; hc = b.conv.cout;
      }
    }
    {
      const int p = prec;
      prec = ECHO_PREC_FP32;
      In hn = plain(h, hc); hn.pro = PRO_GN; hn.nw = &plan.out_norm; hn.eps = 1e-5f; hn.silu = true;
      lin(hn, plan.out_conv, eps_out, d.out_channels, nullptr, 0, 0, s);
      prec = p;
    }
  }
};

// ---- the step as a program for the persistent executor -------------------------------------------------------------------
// Same network walk as forward() above (same fused weights, same prologue/epilogue assignment), with three differences that the
// single kernel makes possible: every row of the time path is the same (one timestep per step), so time_embed / emb_layers /
// box_time_emb are computed for ONE row; the GraphTripleConv gather-combine and mean pooling are prologues of the Linear that
// consumes them; the posterior update is the epilogue of the out conv.
// Tabulate the time path for every t of the schedule with the launchers of forward() (rows = timesteps, in chunks of the
// few-row kernel's 64 rows): the persistent executor reads row t instead of recomputing it in every step.
void echo_layout::build_time_table(cudaStream_t s) {
  const int TT = d.time_num, mc = d.model_channels, E = 4 * mc, gd = d.gconv_dim, et = plan.emb_total;
  ttab_emb = pool.alloc_n<float>((size_t)TT * et);
  ttab_node = pool.alloc_n<float>((size_t)TT * (d.enable_t_emb ? gd : 1));
  float *tmp = nullptr;
  const int CH = 64;
  ECHO_CUDA(cudaMalloc(&tmp, sizeof(float) * CH * (size_t)(mc + 3 * E) + sizeof(int64_t) * CH));
  float *tb = tmp, *x1 = tb + (size_t)CH * mc, *em = x1 + (size_t)CH * E, *ea = em + (size_t)CH * E;
  int64_t* ti = reinterpret_cast<int64_t*>(ea + (size_t)CH * E);
  const int saveN = N;
  try {
    for (int t0 = 0; t0 < TT; t0 += CH) {
      const int n = TT - t0 < CH ? TT - t0 : CH;
      std::vector<int64_t> hv(n);
      for (int i = 0; i < n; ++i) hv[i] = t0 + i;
      ECHO_CUDA(cudaMemcpyAsync(ti, hv.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice, s));
      ECHO_CUDA(cudaStreamSynchronize(s));   // hv dies with the iteration
      N = n;
      timestep_embedding_tab(ti, freqs, n, mc, tb, s);
      lin(tb, mc, plan.time0, x1, E, nullptr, 0, 0, 2, s);
      lin(x1, E, plan.time2, em, E, nullptr, 0, 0, 0, s);
      if (d.enable_t_emb) lin(em, E, time_emb_lin, ttab_node + (size_t)t0 * gd, gd, nullptr, 0, 0, 0, s);
      silu_f32(em, ea, (int64_t)n * E, s);
      lin(ea, E, plan.emb_stack, ttab_emb + (size_t)t0 * et, et, nullptr, 0, 0, 0, s);
    }
    ECHO_CUDA(cudaStreamSynchronize(s));
  } catch (...) {
    N = saveN;
    cudaFree(tmp);
    throw;
  }
  N = saveN;
  cudaFree(tmp);
  ttab_on = true;
}

void echo_layout::build_mk(int Nn, int T) {
  mk_ops.clear();
  mk_stages.clear();
  mk_buf_i = 0;
  const int mc = d.model_channels, E = 4 * mc, gd = d.gconv_dim, od = d.obj_embed_dim, ctx = d.context_dim;
  const int nd = od + gd + (d.enable_t_emb ? gd : 0);
  std::vector<MkOp> A, B;
  int bg_wait = -1, bg_arrive = -1;
  struct Chunk { MkOp op; int id; };
  std::deque<Chunk> pending;   // emb_layers projections waiting for a stage to ride in the shadow of
  auto flush = [&]() {
    if (A.empty() && B.empty()) return;
    if ((int)mk_stages.size() >= 2 && !pending.empty()) {   // emb exists after stage 1 (time_embed.2)
      B.push_back(pending.front().op);
      bg_arrive = pending.front().id;
      pending.pop_front();
    }
    MkStage st;
    memset(&st, 0, sizeof(st));
    st.op_begin = (int)mk_ops.size(); st.n_a = (int)A.size(); st.n_b = (int)B.size();
    st.bg_wait = bg_wait; st.bg_arrive = B.empty() ? -1 : bg_arrive;
    ECHO_CHECK(st.n_a + st.n_b <= MK_MAX_STAGE_OPS, "layout program: too many ops in one stage");
    // Balance the stage: a stage whose ops add up to a little more than one unit per CTA makes a few CTAs run two units while the
    // rest wait at the barrier.  Fatten the units (more features each) of the op with the most units until the stage fits one
    // round, as long as its weight slot allows.
    {
      std::vector<MkOp*> all;
      std::vector<int> mfu;
      for (auto* v : {&A, &B}) for (auto& o : *v) { all.push_back(&o); mfu.push_back(4); }
      auto total = [&]() { int t = 0; for (auto* o : all) t += o->units; return t; };
      for (auto* o : all) mk_plan_op(*o, mk_ctas, 4);
      static const bool no_balance = getenv("ECHO_MK_NO_BALANCE") != nullptr;
      for (int it = 0; it < 16 && !no_balance && total() > mk_ctas && total() <= 3 * mk_ctas; ++it) {
        int best = -1;
        for (size_t i = 0; i < all.size(); ++i)
          if (all[i]->type == MK_T_LIN && all[i]->rclass == 16 && (best < 0 || all[i]->units > all[best]->units)) {
            MkOp trial = *all[i];
            mk_plan_op(trial, mk_ctas, mfu[i] * 2);
            if (trial.units < all[i]->units) best = (int)i;
          }
        if (best < 0) break;
        mfu[best] *= 2;
        mk_plan_op(*all[best], mk_ctas, mfu[best]);
      }
    }
    int ub = 0;
    for (auto* v : {&A, &B})
      for (auto& o : *v) {
        o.unit_begin = ub % mk_ctas;   // the CTA that runs the op's first unit
        ub += o.units;
        mk_ops.push_back(o);
      }
    mk_stages.push_back(st);
    A.clear(); B.clear();
    bg_wait = -1; bg_arrive = -1;
  };
  auto base_op = [&](int M, int K, int nout, const float* X, int64_t ldx, const float* W, const float* bias, float* Y, int64_t ldy) {
    MkOp o;
    memset(&o, 0, sizeof(o));
    o.type = MK_T_LIN; o.M = M; o.K = K; o.nout = nout; o.X = X; o.ldx = ldx; o.W = W; o.bias = bias; o.Y = Y; o.ldy = ldy;
    o.eps = 1e-5f;
    return o;
  };
  auto lin_in = [&](const In& in, const ConvW& w, float* Y, int64_t ldy, const float* res, int64_t ld_res, int act) {
    ECHO_CHECK((in.pro == PRO_GEGLU ? in.C : in.width()) == w.cin, "layout program: linear input width %d != %d", in.width(), w.cin);
    MkOp o = base_op(Nn, w.cin, w.cout, in.X, in.ld ? in.ld : (in.pro == PRO_GEGLU ? 2 * in.C : in.C), w.w, w.b, Y, ldy);
    if (in.X2) { o.X2 = in.X2; o.ldx2 = in.C2; o.K1 = in.C; }
    ECHO_CHECK(in.pro != PRO_GEGLU, "layout program: GEGLU is the epilogue of ff1, not a prologue");
    o.pro = in.pro == PRO_GN ? MK_GN : in.pro == PRO_LN ? MK_LN : in.pro == PRO_SILU ? MK_SILU : MK_NONE;
    if (in.nw) { o.gamma = in.nw->g; o.beta = in.nw->b; }
    o.eps = in.eps; o.pro_act = in.silu ? 1 : 0;
    if (in.pro == PRO_GN) o.cpg = in.width() / 32;
    o.res = res; o.ld_res = ld_res; o.act = act;
    return o;
  };

  // ---- stage 0: time_embed.0 on the sinusoidal row (without the time table), node features that do not need the time MLP,
  //      predicate rows, conv_in ----
  {
    if (!ttab_on) {
      MkOp o = base_op(1, mc, E, nullptr, mc, plan.time0.w, plan.time0.b, e1, E);
      o.pro = MK_TEMB; o.act = 2;
      A.push_back(o);
    } else if (d.enable_t_emb) {   // box_time_emb(emb(t)) from the table -> every node row
      MkOp c;
      memset(&c, 0, sizeof(c));
      c.type = MK_T_COPY; c.M = Nn; c.K = gd; c.x_ext = MK_EXT_TNODE; c.ldx = 0; c.Y = node + od + gd; c.ldy = nd;
      A.push_back(c);
    }
    MkOp c;
    memset(&c, 0, sizeof(c));
    c.type = MK_T_COPY; c.M = Nn; c.K = od; c.x_ext = MK_EXT_OBJ; c.ldx = od; c.Y = node; c.ldy = nd;
    A.push_back(c);
    MkOp b = base_op(Nn, d.in_channels, gd, nullptr, d.in_channels, box_emb.w, box_emb.b, node + od, nd);
    b.x_ext = MK_EXT_XT;
    A.push_back(b);
    if (T > 0) {
      MkOp e;
      memset(&e, 0, sizeof(e));
      e.type = MK_T_EMBROWS; e.M = T; e.K = 2 * gd; e.aux0 = pred_table; e.Y = pred; e.ldy = 2 * gd;
      A.push_back(e);
    }
  }
  float* h0 = nullptr;
  {
    const BlockW& b0 = plan.in_blocks[0];
    ECHO_CHECK(b0.kind == BlockW::CONV_IN, "layout program: first block must be the input conv");
    h0 = mkb(b0.conv.cout);
    MkOp o = base_op(Nn, d.in_channels, b0.conv.cout, nullptr, d.in_channels, b0.conv.w, b0.conv.b, h0, b0.conv.cout);
    o.x_ext = MK_EXT_XT;
    A.push_back(o);
  }
  flush();
  // ---- stage 1: time_embed.2 ----
  if (!ttab_on) {
    A.push_back(base_op(1, E, E, e1, E, plan.time2.w, plan.time2.b, emb, E));
    flush();
  }
  // the emb_layers projections of all ResBlocks, one background op each, in consumption order (plan.emb_stack rows)
  std::vector<std::pair<int, int>> res_chunks;   // (emb_off, cout) per ResBlock in forward order
  {
    auto add = [&](const ResW& r) { res_chunks.push_back({r.emb_off, r.cout}); };
    for (auto& b : plan.in_blocks) if (b.kind == BlockW::RES) add(b.res);
    add(plan.mid0); add(plan.mid2);
    for (auto& b : plan.out_blocks) add(b.res);
    ECHO_CHECK((int)res_chunks.size() <= MK_MAX_BG, "layout program: too many ResBlocks");
    for (size_t j = 0; j < res_chunks.size() && !ttab_on; ++j) {
      const int off = res_chunks[j].first, co = res_chunks[j].second;
      MkOp o = base_op(1, E, co, emb, E, plan.emb_stack.w + (size_t)off * E, plan.emb_stack.b + off, embout + off, plan.emb_total);
      o.pro = MK_SILU;
      pending.push_back({o, (int)j});
    }
  }
  int res_counter = 0;   // index of the next ResBlock in forward order == its background counter
  // ---- stage 2: box_time_emb -> every node row ----
  if (d.enable_t_emb && !ttab_on) {
    MkOp o = base_op(1, E, gd, emb, E, time_emb_lin.w, time_emb_lin.b, node + od + gd, nd);
    o.bcast_rows = Nn;
    A.push_back(o);
    flush();
  }
  // ---- echo GCN: 4 stages per layer ----
  {
    const int H = gcn.H, dp = gcn.dp;
    const float* cur_obj = node;
    const float* cur_pred = pred;
    for (size_t li = 0; li < gcn.layers.size(); ++li) {
      const GcnLayer& L = gcn.layers[li];
      const bool lastl = li + 1 == gcn.layers.size();
      float* nobj = lastl ? latent : gcn.obj_pp[li & 1];
      float* npred = gcn.pred_pp[li & 1];
      A.push_back(base_op(Nn, L.din, 2 * H, cur_obj, L.din, L.w_so.w, nullptr, gcn.pso, 2 * H));
      if (T > 0) {
        MkOp o = base_op(T, dp, H, cur_pred, dp, L.w_p.w, nullptr, gcn.pp, H);
        A.push_back(o);
      }
      if (L.residual) A.push_back(base_op(Nn, L.din, L.dout, cur_obj, L.din, L.proj.w, L.proj.b, gcn.proj, L.dout));
      flush();
      if (T > 0) {
        MkOp o = base_op(T, H, 2 * H + dp, gcn.pso, 2 * H, L.w2.w, L.w2.b, gcn.t2, 2 * H + dp);
        o.pro = MK_EDGE; o.aux0 = gcn.pp; o.aux1 = L.b1; o.act = 1;
        A.push_back(o);
        flush();
      }
      {
        MkOp o = base_op(Nn, H, H, gcn.t2, 2 * H + dp, L.w3.w, L.w3.b, gcn.n1, H);
        o.pro = MK_POOL; o.aux_i = H + dp; o.act = 1;
        A.push_back(o);
      }
      if (T > 0 && !lastl) {   // the last layer's predicate output is never read (denoise_net.py:766)
        if (L.residual) {
          MkOp o = base_op(T, dp, dp, cur_pred, dp, L.projp.w, L.projp.b, npred, dp);
          o.res = gcn.t2 + H; o.ld_res = 2 * H + dp;
          A.push_back(o);
        } else {
          MkOp c;
          memset(&c, 0, sizeof(c));
          c.type = MK_T_COPY; c.M = T; c.K = dp; c.X = gcn.t2 + H; c.ldx = 2 * H + dp; c.Y = npred; c.ldy = dp;
          A.push_back(c);
        }
      }
      flush();
      {
        MkOp o = base_op(Nn, H, L.dout, gcn.n1, H, L.w4.w, L.w4.b, nobj, L.dout);
        o.act = 1;
        if (L.residual) { o.res = gcn.proj; o.ld_res = L.dout; }
        A.push_back(o);
      }
      flush();
      cur_obj = nobj;
      cur_pred = npred;
    }
  }
  // ---- all 11 attn2 = to_out(to_v(latent)) vectors ----
  A.push_back(base_op(Nn, ctx, a2_total, latent, ctx, attn2_fused.w, attn2_fused.b, a2vec, a2_total));
  flush();
  // ---- trunk ----
  auto res_block = [&](const In& x, const ResW& r) -> float* {
    float* out = mkb(r.cout);
    float* h1 = mkb(r.cout);
    In a1 = x; a1.pro = PRO_GN; a1.nw = &r.n1; a1.eps = 1e-5f; a1.silu = true;
    if (ttab_on) {
      MkOp o = lin_in(a1, r.c1, h1, r.cout, nullptr, 0, 0);
      o.res_ext = 1; o.aux_i = r.emb_off;   // + emb_layers(SiLU(emb(t))): one table row for all nodes
      A.push_back(o);
    } else {
      bg_wait = res_counter++;
      A.push_back(lin_in(a1, r.c1, h1, r.cout, embout + r.emb_off, 0 /* one row for all nodes */, 0));
    }
    const float* skip = x.X;
    if (r.has_skip) {
      float* sk = mkb(r.cout);
      A.push_back(lin_in(x, r.skip, sk, r.cout, nullptr, 0, 0));
      skip = sk;
    } else {
      ECHO_CHECK(!x.X2, "layout: identity skip over a concatenated input");
    }
    flush();
    In a2 = plain(h1, r.cout); a2.pro = PRO_GN; a2.nw = &r.n2; a2.eps = 1e-5f; a2.silu = true;
    A.push_back(lin_in(a2, r.c2, out, r.cout, skip, r.cout, 0));
    flush();
    return out;
  };
  auto transformer = [&](const float* x, const AttnW& at, int ai) -> float* {
    const int C = at.C;
    float* out = mkb(C);
    float* t0 = mkb(C);
    float* t1 = mkb(C);
    float* f1 = mkb(4 * C);
    float* t2 = mkb(C);
    In xn = plain(x, C); xn.pro = PRO_GN; xn.nw = &at.norm; xn.eps = 1e-6f;
    A.push_back(lin_in(xn, at.proj_in, t0, C, nullptr, 0, 0));
    flush();
    In l1 = plain(t0, C); l1.pro = PRO_LN; l1.nw = &at.ln1;
    {
      MkOp o = lin_in(l1, attn1_fused[ai], t1, C, t0, C, 0);
      o.res2 = a2vec + a2_off[ai]; o.ld_res2 = a2_total;
      A.push_back(o);
    }
    flush();
    In l3 = plain(t1, C); l3.pro = PRO_LN; l3.nw = &at.ln3;
    {   // ff.net.0 (GEGLU): value rows [0, 4C) and gate rows [4C, 8C) of the projection meet in the epilogue; f1 is (N, 4C)
      MkOp o = lin_in(l3, at.ff1, f1, 4 * C, nullptr, 0, 0);
      o.nout = 4 * C;
      o.epi = MK_EPI_GEGLU;
      A.push_back(o);
    }
    flush();
    if (!ffo_fused.empty()) {   // out = x + [P | P F2] [t1 ; g] + (P b2 + bP)
      In cat = plain(t1, C);
      cat.X2 = f1; cat.C2 = 4 * C;
      A.push_back(lin_in(cat, ffo_fused[ai], out, C, x, C, 0));
      flush();
      (void)t2;
    } else {
      A.push_back(lin_in(plain(f1, 4 * C), at.ff2, t2, C, t1, C, 0));
      flush();
      A.push_back(lin_in(plain(t2, C), at.proj_out, out, C, x, C, 0));
      flush();
    }
    return out;
  };
  std::vector<std::pair<const float*, int>> hs;
  const float* h = h0;
  int hc = plan.in_blocks[0].conv.cout, ai = 0;
  hs.push_back({h, hc});
  for (size_t bi = 1; bi < plan.in_blocks.size(); ++bi) {
    const BlockW& b = plan.in_blocks[bi];
    if (b.kind == BlockW::RES) {
      h = res_block(plain(h, hc), b.res); hc = b.res.cout;
      if (b.attn) h = transformer(h, b.at, ai++);
    } else {
      float* o = mkb(b.conv.cout);
      A.push_back(lin_in(plain(h, hc), b.conv, o, b.conv.cout, nullptr, 0, 0));
      flush();
      h = o; hc = b.conv.cout;
    }
    hs.push_back({h, hc});
  }
  h = res_block(plain(h, hc), plan.mid0);
  h = transformer(h, plan.mid_at, ai++);
  h = res_block(plain(h, plan.mid0.cout), plan.mid2);
  hc = plan.mid2.cout;
  for (auto& b : plan.out_blocks) {
    auto sk = hs.back();
    hs.pop_back();
    In cat = plain(h, hc);
    cat.X2 = sk.first; cat.C2 = sk.second;
    h = res_block(cat, b.res); hc = b.res.cout;
    if (b.attn) h = transformer(h, b.at, ai++);
    if (b.up) {
      float* o = mkb(b.conv.cout);
      A.push_back(lin_in(plain(h, hc), b.conv, o, b.conv.cout, nullptr, 0, 0));
      flush();
      h = o; hc = b.conv.cout;
    }
  }
  {
    In hn = plain(h, hc); hn.pro = PRO_GN; hn.nw = &plan.out_norm; hn.eps = 1e-5f; hn.silu = true;
    MkOp o = lin_in(hn, plan.out_conv, nullptr, d.out_channels, nullptr, 0, 0);
    o.epi = MK_EPI_DDPM; o.y_ext = MK_EXT_XPREV;
    A.push_back(o);
    flush();
  }
  ECHO_CHECK(pending.empty(), "layout program: %d emb_layers projections were never scheduled", (int)pending.size());
  ECHO_CHECK((int)mk_stages.size() <= MK_MAX_STAGES && (int)mk_ops.size() <= MK_MAX_OPS, "layout program too long (%d stages, %d ops)",
             (int)mk_stages.size(), (int)mk_ops.size());
  mk_stages_lite.clear();
  for (const MkStage& st : mk_stages) {
    ECHO_CHECK(st.op_begin < 65536 && st.bg_wait < 128 && st.bg_arrive < 128, "layout program: stage record out of range");
    MkStageLite l;
    l.op_begin = (unsigned short)st.op_begin; l.n_a = (unsigned char)st.n_a; l.n_b = (unsigned char)st.n_b;
    l.bg_wait = (signed char)st.bg_wait; l.bg_arrive = (signed char)st.bg_arrive; l.pad = 0;
    mk_stages_lite.push_back(l);
  }
  mk_build_fetch(mk_ops, mk_stages, mk_ctas, mk_fetch, mk_fetch_off);
  ECHO_CHECK((int)mk_fetch.size() <= MK_MAX_FETCH, "layout program: %d weight fetches exceed the table", (int)mk_fetch.size());
  mk_N = Nn;
  mk_T = T;
}

void echo_layout::step_mk(const echo_graph* g, const float* x_t, const float* obj_embed, int t, const float* noise, float* x_prev,
                          cudaStream_t s) {
  const int Nn = g->n_nodes, T = g->n_triples;
  static const bool no_ttab = getenv("ECHO_MK_NO_TTAB") != nullptr;
  if (!ttab_on && !no_ttab && !ttab_emb) build_time_table(s);
  if (Nn != mk_N || T != mk_T) {
    build_mk(Nn, T);
    // a new program renumbers the stages: restart the counters (stream-ordered behind the previous program's last launch)
    ECHO_CUDA(cudaMemsetAsync(d_mk_counters, 0, sizeof(unsigned) * (MK_MAX_STAGES + MK_MAX_BG + 2), s));
    ECHO_CUDA(cudaMemcpyAsync(d_mk_ops, mk_ops.data(), sizeof(MkOp) * mk_ops.size(), cudaMemcpyHostToDevice, s));
    ECHO_CUDA(cudaMemcpyAsync(d_mk_stages, mk_stages.data(), sizeof(MkStage) * mk_stages.size(), cudaMemcpyHostToDevice, s));
    ECHO_CUDA(cudaMemcpyAsync(d_mk_stages_lite, mk_stages_lite.data(), sizeof(MkStageLite) * mk_stages_lite.size(), cudaMemcpyHostToDevice, s));
    ECHO_CUDA(cudaMemcpyAsync(d_mk_fetch, mk_fetch.data(), sizeof(MkFetch) * mk_fetch.size(), cudaMemcpyHostToDevice, s));
    ECHO_CUDA(cudaMemcpyAsync(d_mk_fetch_off, mk_fetch_off.data(), sizeof(int) * mk_fetch_off.size(), cudaMemcpyHostToDevice, s));
    ECHO_CUDA(cudaStreamSynchronize(s));   // the host vectors may be rebuilt before an asynchronous copy would have read them
  }
  MkArgs a;
  memset(&a, 0, sizeof(a));
  a.ops = d_mk_ops; a.stages = d_mk_stages; a.n_stages = (int)mk_stages.size();
  a.fetch = d_mk_fetch; a.fetch_off = d_mk_fetch_off; a.stages_lite = d_mk_stages_lite;
  static const int max_stages = getenv("ECHO_MK_MAX_STAGES") ? atoi(getenv("ECHO_MK_MAX_STAGES")) : 0;   // debugging: run a prefix
  if (max_stages > 0 && max_stages < a.n_stages) a.n_stages = max_stages;
  a.bar = d_mk_counters; a.bg = d_mk_counters + MK_MAX_STAGES; a.epoch = d_mk_counters + MK_MAX_STAGES + MK_MAX_BG;
  a.err = a.epoch + 1;
  a.x_t = x_t; a.obj_embed = obj_embed; a.noise = noise; a.x_prev = x_prev;
  if (ttab_on) {
    a.emb_row = ttab_emb + (size_t)t * plan.emb_total;
    a.tnode_row = ttab_node + (size_t)t * d.gconv_dim;
  }
  a.t = t; a.tab = d_tab; a.T = d.time_num; a.freqs = freqs; a.temb_dim = d.model_channels;
  a.s_idx = g->s_idx; a.o_idx = g->o_idx; a.node_off = g->node_off; a.node_items = g->node_items;
  a.triples = (const long long*)g->triples;
  a.H = gcn.H;
  static const bool evict_first = !getenv("ECHO_MK_NO_EVICT_FIRST");
  static const bool fenced = getenv("ECHO_MK_FENCE") != nullptr;   // A/B: __threadfence + atomicAdd instead of red.release
  a.flags = (evict_first ? 1 : 0) | (fenced ? 2 : 0);
  static const char* timeline = getenv("ECHO_MK_TIMELINE");   // diagnostics: per-CTA per-stage SM clocks of every step -> file (last step wins)
  long long* d_dbg = nullptr;
  const size_t dbg_n = (size_t)mk_ctas * a.n_stages * 12;
  if (timeline) {
    ECHO_CUDA(cudaMalloc(&d_dbg, dbg_n * sizeof(long long)));
    ECHO_CUDA(cudaMemsetAsync(d_dbg, 0, dbg_n * sizeof(long long), s));
    a.dbg = d_dbg;
  }
  mk_launch(a, mk_ctas, s);
  ++mk_steps;
  if (timeline) {
    std::vector<long long> hd(dbg_n);
    ECHO_CUDA(cudaStreamSynchronize(s));
    ECHO_CUDA(cudaMemcpy(hd.data(), d_dbg, dbg_n * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(d_dbg);
    if (FILE* f = fopen(timeline, "wb")) {
      const long long hdr[2] = {mk_ctas, a.n_stages};
      fwrite(hdr, sizeof(long long), 2, f);
      fwrite(hd.data(), sizeof(long long), dbg_n, f);
      // per stage: foreground / background op counts and the shape of the first op
      for (int i = 0; i < a.n_stages; ++i) {
        const MkStage& st = mk_stages[i];
        const MkOp& o = mk_ops[st.op_begin];
        const long long rec[8] = {st.n_a, st.n_b, o.type, o.M, o.K, o.nout, o.pro, o.units};
        fwrite(rec, sizeof(long long), 8, f);
      }
      fclose(f);
    }
  }
}

namespace echo {

// get_betas('linear') = np.linspace(b0, b1, T) float64 (diffusion_ddpm.py:38-40); tables as fp32 torch ops (:133-162).
// tab = 5 x T: sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1, posterior_mean_coef2,
// posterior_log_variance_clipped.  Host-only (also exported as echo_debug_ddpm_tables for CPU tests).
void ddpm_tables(int T, float beta_start, float beta_end, std::vector<float>& tab) {
  ECHO_CHECK(T >= 1, "layout: time_num must be >= 1");
  std::vector<double> betas(T);
  const double b0 = (double)beta_start, b1 = (double)beta_end;
  const double step = T > 1 ? (b1 - b0) / (double)(T - 1) : 0.0;
  for (int i = 0; i < T; ++i) betas[i] = (double)i * step + b0;
  if (T > 1) betas[T - 1] = b1;
  std::vector<float> ac(T), acp(T), b32(T), a32(T);
  double cp = 1.0;
  for (int i = 0; i < T; ++i) {
    cp *= (1.0 - betas[i]);
    ac[i] = (float)cp;
    b32[i] = (float)betas[i];
    a32[i] = (float)(1.0 - betas[i]);
  }
  for (int i = 0; i < T; ++i) acp[i] = i == 0 ? 1.0f : ac[i - 1];
  tab.assign((size_t)5 * T, 0.f);
  for (int i = 0; i < T; ++i) {
    const float one_m = 1.0f - ac[i];
    tab[0 * T + i] = sqrtf(1.0f / ac[i]);
    tab[1 * T + i] = sqrtf(1.0f / ac[i] - 1.0f);
    tab[2 * T + i] = b32[i] * sqrtf(acp[i]) / one_m;
    tab[3 * T + i] = (1.0f - acp[i]) * sqrtf(a32[i]) / one_m;
    const float pv = b32[i] * (1.0f - acp[i]) / one_m;
    tab[4 * T + i] = logf(fmaxf(pv, 1e-20f));
  }
}

static void make_ddpm_tables(echo_layout* h) {
  ddpm_tables(h->d.time_num, h->d.beta_start, h->d.beta_end, h->h_tab);
  h->d_tab = h->pool.upload(h->h_tab);
}

echo_layout* layout_create(const echo_layout_desc_t* desc, const echo_weight_t* weights, int n_weights) {
  ECHO_CHECK(desc, "layout: null desc");
  echo_layout* h = new echo_layout();
  try {
    h->d = *desc;
    const echo_layout_desc_t& d = h->d;
    ECHO_CHECK(d.max_nodes > 0 && d.model_channels % 32 == 0 && d.num_levels >= 1 && d.num_levels <= 8, "layout: bad config");
    ECHO_CHECK(d.in_channels % 4 == 0 && d.obj_embed_dim % 4 == 0 && d.gconv_dim % 4 == 0, "layout: channel counts must be multiples of 4");
    h->prec = d.precision;
    ECHO_CHECK(h->prec == ECHO_PREC_FP32 || h->prec == ECHO_PREC_BF16, "layout: unknown precision %d", h->prec);
    WeightMap wm;
    wm.load(weights, n_weights);
    cudaStream_t s = 0;
    UNetCfg cfg;
    cfg.dims = 1;
    cfg.in_channels = d.in_channels; cfg.out_channels = d.out_channels; cfg.model_channels = d.model_channels;
    cfg.channel_mult.assign(d.channel_mult, d.channel_mult + d.num_levels);
    cfg.attention_resolutions.assign(d.attention_resolutions, d.attention_resolutions + d.num_attention_resolutions);
    cfg.num_res_blocks = d.num_res_blocks; cfg.num_heads = d.num_heads; cfg.context_dim = d.context_dim;
    cfg.want_bf16 = false;   // see lin(): the layout branch streams fp32 weights in both modes
    build_unet_plan(wm, cfg, h->pool, h->plan, s);
    const int mc = d.model_channels, E = 4 * mc, gd = d.gconv_dim;
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
    h->box_emb = linear("box_embeddings", d.in_channels, gd);
    if (d.enable_t_emb) h->time_emb_lin = linear("box_time_emb", E, gd);
    {
      const WView& pt = wm.get("pred_embeddings.weight");
      ECHO_CHECK(pt.shape.size() == 2 && pt.shape[1] == 2 * gd, "pred_embeddings: bad shape");
      float* o = h->pool.alloc_n<float>(pt.numel());
      ECHO_CUDA(cudaMemcpyAsync(o, pt.p, sizeof(float) * pt.numel(), cudaMemcpyDeviceToDevice, s));
      h->pred_table = o;
      h->pred_rows = (int)pt.shape[0];
    }
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
    echo_gcn_desc_t gdsc = {};
    gdsc.input_dim_obj = d.obj_embed_dim + gd + (d.enable_t_emb ? gd : 0);   // denoise_net.py:725-727
    gdsc.input_dim_pred = 2 * gd;
    gdsc.num_layers = 5;
    gdsc.hidden_dim = 4 * gd;
    gdsc.output_dim = d.context_dim;
    gdsc.max_nodes = d.max_nodes;
    gdsc.max_triples = d.max_triples > 0 ? d.max_triples : 1;
    gdsc.bn_eps = 1e-5f;
    gdsc.keep_train_weights = d.keep_train_weights;
    h->gcn.create(wm, "box_graph_cov.", gdsc, h->pool);
    make_ddpm_tables(h);
    const size_t N = d.max_nodes, T = gdsc.max_triples;
    if (N > 64) {   // widest prologue input of this network: the GEGLU product / the skip concatenations / the stacked embeddings
      h->lin_scratch_floats = N * (4096 + 8192);   // + partial tiles of a split reduction (sgemm_x3_auto)
      for (int i = 0; i < 2; ++i) h->lin_scratch[i] = h->pool.alloc_n<float>(h->lin_scratch_floats);
    }
    h->temb = h->pool.alloc_n<float>(N * mc);
    h->e1 = h->pool.alloc_n<float>(N * E);
    h->emb = h->pool.alloc_n<float>(N * E);
    h->emb_act = h->pool.alloc_n<float>(N * E);
    h->node = h->pool.alloc_n<float>(N * gdsc.input_dim_obj);
    h->pred = h->pool.alloc_n<float>(T * 2 * gd);
    h->latent = h->pool.alloc_n<float>(N * d.context_dim);
    h->embout = h->pool.alloc_n<float>(N * h->plan.emb_total);
    h->a2_total = h->plan.v2_total;
    h->a2vec = h->pool.alloc_n<float>(N * h->a2_total);
    h->eps = h->pool.alloc_n<float>(N * d.out_channels);
    h->t_dev = h->pool.alloc_n<int64_t>(N);
    h->sx = h->pool.alloc_n<float>(N * d.out_channels);
    h->snoise = h->pool.alloc_n<float>(N * d.out_channels);
    h->sout = h->pool.alloc_n<float>(N * d.out_channels);
    h->sobj = h->pool.alloc_n<float>(N * d.obj_embed_dim);
    h->t_slot = h->pool.alloc_n<int>(4);
    {
      int off = 0;
      auto add = [&](const AttnW& a) { h->a2_off.push_back(off); off += a.C; };
      for (auto& b : h->plan.in_blocks) if (b.attn) add(b.at);
      add(h->plan.mid_at);
      for (auto& b : h->plan.out_blocks) if (b.attn) add(b.at);
    }
    ECHO_CUDA(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    ECHO_CUDA(cudaStreamCreateWithFlags(&h->cap, cudaStreamNonBlocking));
    ECHO_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    ECHO_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    {   // fused attention matrices (see the member comments): products in fp32 on the device, bf16 copies in bf16 mode
      cudaStream_t s0 = 0;
      auto matmul = [&](const float* A, int M, int J, const float* B, int K, float* Cout) {   // C[M,K] = A[M,J] B[J,K]
        GemmArgs g;
        g.A = A; g.n = 1; g.w = M; g.ow = M; g.cin = J; g.lda = J;
        g.W = B; g.w_stride_n = 1; g.w_stride_k = K; g.cout = K;
        g.out = Cout; g.ldo = K;
        gemm_simt(g, s0);
      };
      auto to_bf16 = [&](const float* w, size_t n) {
        __nv_bfloat16* o = h->pool.alloc_n<__nv_bfloat16>(n);
        convert(w, F32, o, BF16, (int64_t)n, s0);
        return (const __nv_bfloat16*)o;
      };
      const bool bf = false;   // fp32 weights in both modes, see lin()
      const int ctx = d.context_dim;
      float* w2 = h->pool.alloc_n<float>((size_t)h->a2_total * ctx);
      float* b2 = h->pool.alloc_n<float>(h->a2_total);
      int ai = 0;
      auto fuse = [&](const AttnW& a) {
        const int C = a.C;
        ConvW f = a.attn1_out;                                              // bias of to_out; cin = cout = C
        float* w1 = h->pool.alloc_n<float>((size_t)C * C);
        matmul(a.attn1_out.w, C, C, a.v_only.w, C, w1);                     // to_v has no bias
        f.w = w1;
        f.wb = bf ? to_bf16(w1, (size_t)C * C) : nullptr;
        h->attn1_fused.push_back(f);
        const int off = h->a2_off[ai++];
        matmul(a.attn2_out.w, C, C, h->plan.v2_stack.w + (size_t)a.v2_off * ctx, ctx, w2 + (size_t)off * ctx);
        ECHO_CUDA(cudaMemcpyAsync(b2 + off, a.attn2_out.b, sizeof(float) * C, cudaMemcpyDeviceToDevice, s0));
      };
      for (auto& b : h->plan.in_blocks) if (b.attn) fuse(b.at);
      fuse(h->plan.mid_at);
      for (auto& b : h->plan.out_blocks) if (b.attn) fuse(b.at);
      {   // [P | P F2] and P b2 + bP per transformer block (persistent executor only)
        static const bool no_ffo = getenv("ECHO_MK_NO_FFO") != nullptr;
        auto fuse_ffo = [&](const AttnW& a) {
          const int C = a.C, K = 5 * C;
          float* w = h->pool.alloc_n<float>((size_t)C * K);
          float* pf = h->pool.alloc_n<float>((size_t)C * 4 * C);
          float* bb = h->pool.alloc_n<float>(C);
          matmul(a.proj_out.w, C, C, a.ff2.w, 4 * C, pf);                                   // P F2: [C, 4C]
          ECHO_CUDA(cudaMemcpy2DAsync(w, sizeof(float) * K, a.proj_out.w, sizeof(float) * C, sizeof(float) * C, C, cudaMemcpyDeviceToDevice, s0));
          ECHO_CUDA(cudaMemcpy2DAsync(w + C, sizeof(float) * K, pf, sizeof(float) * 4 * C, sizeof(float) * 4 * C, C, cudaMemcpyDeviceToDevice, s0));
          {   // bias: P b2 + bP  (b2 as a [C,1] matrix)
            GemmArgs g;
            g.A = a.proj_out.w; g.n = 1; g.w = C; g.ow = C; g.cin = C; g.lda = C;
            g.W = a.ff2.b; g.w_stride_n = 1; g.w_stride_k = 1; g.cout = 1;
            g.out = bb; g.ldo = 1;
            gemm_simt(g, s0);
            add_rowvec(bb, F32, 1, C, a.proj_out.b, C, 1, s0);
          }
          ConvW f;
          f.w = w; f.b = bb; f.cin = K; f.cout = C; f.taps = 1;
          h->ffo_fused.push_back(f);
        };
        if (!no_ffo) {
          for (auto& b : h->plan.in_blocks) if (b.attn) fuse_ffo(b.at);
          fuse_ffo(h->plan.mid_at);
          for (auto& b : h->plan.out_blocks) if (b.attn) fuse_ffo(b.at);
        }
      }
      h->attn2_fused.cin = ctx; h->attn2_fused.cout = h->a2_total; h->attn2_fused.taps = 1;
      h->attn2_fused.w = w2; h->attn2_fused.b = b2;
      h->attn2_fused.wb = bf ? to_bf16(w2, (size_t)h->a2_total * ctx) : nullptr;
      ECHO_CUDA(cudaStreamSynchronize(s0));
    }
    {   // persistent executor (layout_mk.cu)
      static const bool no_mk = getenv("ECHO_NO_MK") != nullptr;
      h->mk_ok = !no_mk && mk_available(&h->mk_ctas) && d.in_channels == d.out_channels && d.model_channels % 2 == 0;
      if (h->mk_ok) {
        h->d_mk_ops = (MkOp*)h->pool.alloc(sizeof(MkOp) * echo_layout::MK_MAX_OPS);
        h->d_mk_stages = (MkStage*)h->pool.alloc(sizeof(MkStage) * MK_MAX_STAGES);
        h->d_mk_stages_lite = (MkStageLite*)h->pool.alloc(sizeof(MkStageLite) * MK_MAX_STAGES);
        h->d_mk_fetch = (MkFetch*)h->pool.alloc(sizeof(MkFetch) * echo_layout::MK_MAX_FETCH);
        h->d_mk_fetch_off = (int*)h->pool.alloc(sizeof(int) * (h->mk_ctas + 1));
        h->d_mk_counters = (unsigned*)h->pool.alloc(sizeof(unsigned) * (MK_MAX_STAGES + echo_layout::MK_MAX_BG + 2));
        ECHO_CUDA(cudaMemset(h->d_mk_counters, 0, sizeof(unsigned) * (MK_MAX_STAGES + echo_layout::MK_MAX_BG + 2)));
      }
    }
    // workspace: every block output is (N, <= 2*mc*max_mult) fp32; ~60 live buffers is a generous bound
    int maxmult = 1;
    for (int i = 0; i < d.num_levels; ++i) maxmult = d.channel_mult[i] > maxmult ? d.channel_mult[i] : maxmult;
    const size_t per = (size_t)N * mc * maxmult * sizeof(float);
    const size_t nblk = h->plan.in_blocks.size() + h->plan.out_blocks.size() + 4;
    h->arena.init(per * (nblk * 5 + 64) + (size_t(1) << 20));
    ECHO_CUDA(cudaStreamSynchronize(s));
    return h;
  } catch (...) {
    h->arena.destroy();
    h->pool.destroy();
    delete h;
    throw;
  }
}

void layout_destroy(echo_layout* h) {
  if (!h) return;
  if (h->side) cudaStreamDestroy(h->side);
  if (h->cap) cudaStreamDestroy(h->cap);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  h->arena.destroy();
  h->pool.destroy();
  delete h;
}

void layout_forward(echo_layout* h, const echo_graph* g, const float* box_t, const float* obj_embed, const int64_t* t, float* eps_out,
                    cudaStream_t s) {
  h->forward(g, box_t, obj_embed, t, eps_out, s);
}

// ---- one DDPM iteration as a replayed CUDA graph ------------------------------------------------------------------
// The layout step is ~320 dependent launches of microsecond kernels: issued one by one it is bound by launch latency,
// not by the GPU.  The step is therefore captured once per (scene graph, node count) into a CUDA graph whose kernels
// read the timestep from a device slot and the inputs from handle-owned staging buffers; an iteration is then
// 1 staging kernel + 1 graph launch + 1 copy-out kernel.  ECHO_NO_GRAPH=1 disables it (per-kernel profiling).
namespace {
__global__ void layout_stage_in_kernel(const float* __restrict__ x, const float* __restrict__ obj, const float* __restrict__ noise,
                                       int nx, int nobj, float* __restrict__ sx, float* __restrict__ sobj, float* __restrict__ snoise,
                                       int* __restrict__ t_slot, int t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *t_slot = t;
  if (i < nx) { sx[i] = x[i]; snoise[i] = noise[i]; }
  if (i < nobj) sobj[i] = obj[i];
}
__global__ void layout_fill_t_kernel(const int* __restrict__ t_slot, int64_t* __restrict__ t_dev, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t_dev[i] = *t_slot;
}
// same arithmetic as ddpm_update_kernel (elem.cu), timestep read from the device slot
__global__ void layout_ddpm_update_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                                          const float* __restrict__ tab, int T, const int* __restrict__ t_slot, int count,
                                          float* __restrict__ out) {
  const int t = *t_slot;
  const float a = tab[t], b = tab[T + t], c1 = tab[2 * T + t], c2 = tab[3 * T + t], lv = tab[4 * T + t];
  const float sig = (t == 0 ? 0.f : 1.f) * expf(0.5f * lv);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    const float x0 = __fsub_rn(__fmul_rn(a, x[i]), __fmul_rn(b, eps[i]));
    const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, x[i]));
    out[i] = __fadd_rn(mean, __fmul_rn(sig, noise[i]));
  }
}
__global__ void layout_copy_out_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
}  // namespace

void layout_set_batch_stats(echo_layout* h, bool on) {
  ECHO_CHECK(h, "layout_set_batch_stats: null handle");
  ECHO_CHECK(!on || h->d.keep_train_weights, "layout_set_batch_stats: the handle was created without keep_train_weights");
  h->batch_stats = on;
}

int g_layout_mode = 0;   // echo_debug_set_layout_mode: 0 = automatic, 1 = never the persistent executor (per-layer kernels / graph replay)
void set_layout_mode(int m) { g_layout_mode = m; }
// {persistent-executor steps, graph replays, program stages, program ops, kernels inside the replayed graph, executor CTAs;
//  negative CTA count: the executor's barrier watchdog fired (synchronises the device to read the flag)}
void layout_info(const echo_layout* h, int64_t* out6) {
  unsigned err = 0;
  if (h->mk_ok) ECHO_CUDA(cudaMemcpy(&err, h->d_mk_counters + MK_MAX_STAGES + echo_layout::MK_MAX_BG + 1, sizeof(err), cudaMemcpyDeviceToHost));
  out6[0] = h->mk_steps; out6[1] = h->graph_replays; out6[2] = (int64_t)h->mk_stages.size(); out6[3] = (int64_t)h->mk_ops.size();
  out6[4] = h->graph_launches; out6[5] = h->mk_ok ? (err ? -h->mk_ctas : h->mk_ctas) : 0;
}

void layout_step(echo_layout* h, const echo_graph* g, const float* x_t, const float* obj_embed, int t, const float* noise, float* x_prev,
                 cudaStream_t s) {
  ECHO_CHECK(t >= 0 && t < h->d.time_num, "layout_step: t=%d outside [0, %d)", t, h->d.time_num);
  ECHO_CHECK(g, "layout_step: null graph");
  ECHO_CHECK(!h->batch_stats, "layout_step: the sampler step runs on running statistics; switch echo_layout_set_batch_stats off first");
  const int N = g->n_nodes, nx = N * h->d.out_channels, nobj = N * h->d.obj_embed_dim;
  if (N == 0) return;
  ECHO_CHECK(N <= h->d.max_nodes && h->d.in_channels == h->d.out_channels, "layout_step: graph exceeds handle capacity");
  // the persistent executor: one cooperative kernel per iteration (few-row graphs; larger batches keep the per-layer kernels)
  if (h->mk_ok && g_layout_mode != 1 && N <= 64 && g->n_triples <= 512) {
    ECHO_CHECK(g->n_triples == 0 || (g->p_min >= 0 && g->p_max < h->pred_rows), "layout: predicate ids [%lld, %lld] outside pred_embeddings (%d rows)",
               (long long)g->p_min, (long long)g->p_max, h->pred_rows);
    ECHO_CHECK(g->n_triples <= h->d.max_triples, "layout_step: graph exceeds handle capacity");
    h->step_mk(g, x_t, obj_embed, t, noise, x_prev, s);
    return;
  }
  static const bool no_graph = getenv("ECHO_NO_GRAPH") != nullptr;
  if (no_graph || h->graph_failed) {
    fill_i64(h->t_dev, N, t, s);
    h->forward(g, x_t, obj_embed, h->t_dev, h->eps, s);
    ddpm_update(x_t, h->eps, noise, h->d_tab, h->d.time_num, t, (int64_t)nx, x_prev, s);
    return;
  }
  const int mx = nx > nobj ? nx : nobj;
  layout_stage_in_kernel<<<cdiv(mx, 256), 256, 0, s>>>(x_t, obj_embed, noise, nx, nobj, h->sx, h->sobj, h->snoise, h->t_slot, t);
  ECHO_LAUNCH_CHECK();
  if (!h->gexec || h->gkey_id != g->id || h->gkey_n != N) {
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    cudaGraph_t graph = nullptr;
    const int64_t before = g_launches;
    cudaError_t e = cudaStreamBeginCapture(h->cap, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      try {
        layout_fill_t_kernel<<<cdiv(N, 128), 128, 0, h->cap>>>(h->t_slot, h->t_dev, N);
        ECHO_LAUNCH_CHECK();
        h->forward(g, h->sx, h->sobj, h->t_dev, h->eps, h->cap);
        layout_ddpm_update_kernel<<<cdiv(nx, 128), 128, 0, h->cap>>>(h->sx, h->eps, h->snoise, h->d_tab, h->d.time_num, h->t_slot, nx, h->sout);
        ECHO_LAUNCH_CHECK();
      } catch (...) {
        cudaStreamEndCapture(h->cap, &graph);
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        throw;
      }
      e = cudaStreamEndCapture(h->cap, &graph);
    }
    if (e == cudaSuccess && graph) e = cudaGraphInstantiate(&h->gexec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess || !h->gexec) {   // capture / instantiation failed: run the kernels directly (retried on the next graph)
      cudaGetLastError();
      h->gexec = nullptr;
      h->gkey_id = 0;
      fill_i64(h->t_dev, N, t, s);
      h->forward(g, x_t, obj_embed, h->t_dev, h->eps, s);
      ddpm_update(x_t, h->eps, noise, h->d_tab, h->d.time_num, t, (int64_t)nx, x_prev, s);
      return;
    }
    h->graph_launches = g_launches - before;
    g_launches = before;
    h->gkey_id = g->id;
    h->gkey_n = N;
  }
  ECHO_CUDA(cudaGraphLaunch(h->gexec, s));
  ++h->graph_replays;
  count_launch((int)h->graph_launches);
  layout_copy_out_kernel<<<cdiv(nx, 128), 128, 0, s>>>(h->sout, x_prev, nx);
  ECHO_LAUNCH_CHECK();
}

}  // namespace echo

namespace echo {
const std::vector<float>& layout_table(const echo_layout* h) { return h->h_tab; }
}  // namespace echo
