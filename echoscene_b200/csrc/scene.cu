// Once-per-scene encoders of Sg2ScDiffModel.sample: what runs between the dataset's tensors and the two
// diffusion chains.  Reference: model/EchoScene.py
//   init_encoder :143-157   obj_embed  = [CLIP text feature | obj_embeddings_ec(objs)]            (N, add + 2*gd)
//                           pred_embed = [CLIP relation feature | pred_embeddings_ec(p)]           (T, add + 2*gd)
//                           latent_obj = gconv_net_ec(obj_embed, pred_embed, edges)                (N, add + 2*gd)
//   sample       :393-397   change flag = zeros(N, gd) (np.zeros per node -> torch -> .cuda() in the reference)
//   manipulate   :181-195   latent = gconv_net_manipulation([latent_obj | change | obj_embed], pred_embed, edges)
//   rel_s_mlp    :97-100    make_mlp([feat, 960, 1280], 'batch', norelu=True) = Linear, BatchNorm1d, ReLU, Linear
//                :405-410   uc_s = rel_s_mlp(obj_embed), c_s = rel_s_mlp(latent)
//
// Everything is a composition of the launchers the denoiser steps already use (row copies, embedding rows, the
// GraphTripleConvNet executor of gcn.cu, the few-row linear); eval BatchNorm is folded into its Linear at create time.
// One call = ~150 asynchronous launches on the caller's stream: no host synchronisation, no per-node host->device
// copies (the reference does N of them, :393-397), capturable in a CUDA graph.
#include "model.cuh"

#include <algorithm>

struct echo_scene {
  echo::DevPool pool;
  echo_scene_desc_t d;
  int feat = 0, emb = 0, add = 0, gd = 0, din_mani = 0;
  echo::Gcn ec, mani;
  const float* obj_tab = nullptr;
  const float* pred_tab = nullptr;
  const float* pred_tab_man = nullptr;   // pred_embeddings_man_dc when desc.manipulate_pred_dc (layout-only model), else null
  bool has_rel_s = false;
  echo::Mat rel0, rel3;          // rel_s_mlp.0 (+ BatchNorm rel_s_mlp.1 folded), rel_s_mlp.3
  echo::Mat rel0_raw;            // batch-statistics mode: rel_s_mlp.0 unfolded + BatchNorm1d scale / shift
  const float *rel_bn_g = nullptr, *rel_bn_b = nullptr;
  bool has_train = false, batch_stats = false;
  float bn_eps = 1e-5f;
  // workspace (max_nodes / max_triples rows)
  float *obj_embed = nullptr, *pred_embed = nullptr, *latent_obj = nullptr, *mani_in = nullptr, *latent = nullptr, *rel_h = nullptr;
  float* pred_embed_man = nullptr;       // predicate embeddings of the manipulate stage when they come from pred_tab_man
};

namespace echo {
namespace {

// out[r, :] = [feat[r, :add] | table[idx[r*stride + off], :emb]]   (torch.cat([text_feat, embedding(objs)], dim=1))
void embed_rows(const echo_scene* h, const float* table, const int64_t* idx, int64_t stride, int64_t off, const float* feat,
                int64_t rows, float* out, cudaStream_t s) {
  if (rows == 0) return;
  if (h->add > 0) copy_cols(feat, h->add, rows, h->add, out, h->feat, s);
  embedding_rows(table, h->emb, idx, stride, off, rows, out + h->add, h->feat, s);
}

void check_graph(const echo_scene* h, const echo_graph* g, const char* what) {
  ECHO_CHECK(h && g, "%s: null handle", what);
  ECHO_CHECK(g->n_nodes > 0, "%s: empty scene", what);
  ECHO_CHECK(g->n_nodes <= h->d.max_nodes && g->n_triples <= h->d.max_triples, "%s: graph (%d nodes, %d triples) exceeds handle capacity (%d, %d)",
             what, g->n_nodes, g->n_triples, h->d.max_nodes, h->d.max_triples);
  // nn.Embedding raises an index error for these (EchoScene.py:150)
  ECHO_CHECK(g->n_triples == 0 || (g->p_min >= 0 && g->p_max < h->d.num_preds), "%s: predicate ids [%lld, %lld] outside pred_embeddings_ec (%d rows)",
             what, (long long)g->p_min, (long long)g->p_max, h->d.num_preds);
}

}  // namespace

echo_scene* scene_create(const echo_scene_desc_t* desc, const echo_weight_t* weights, int n_weights) {
  ECHO_CHECK(desc, "scene_create: null desc");
  const echo_scene_desc_t& d = *desc;
  ECHO_CHECK(d.gconv_dim > 0 && d.gconv_dim % 4 == 0 && d.add_dim >= 0 && d.add_dim % 4 == 0 && d.num_objs > 0 && d.num_preds > 0 &&
                 d.num_layers > 0 && d.max_nodes > 0,
             "scene_create: bad dims (gconv_dim and add_dim must be multiples of 4)");
  echo_scene* h = new echo_scene();
  try {
    h->d = d;
    if (h->d.max_triples < 1) h->d.max_triples = 1;
    h->gd = d.gconv_dim;
    h->emb = 2 * d.gconv_dim;
    h->add = d.add_dim;
    h->feat = h->emb + h->add;                       // out_dim_ini_encoder == out_dim_manipulator (EchoScene.py:40-44)
    h->din_mani = h->feat + h->gd + h->feat;         // latent | change flag | embedding + CLIP (EchoScene.py:78-86)
    const float eps = d.bn_eps > 0 ? d.bn_eps : 1e-5f;
    cudaStream_t s = 0;
    WeightMap wm;
    wm.load(weights, n_weights);
    auto table = [&](const char* key, int rows) {
      const WView& t = wm.get(key, {rows, h->emb});
      float* o = h->pool.alloc_n<float>(t.numel());
      ECHO_CUDA(cudaMemcpyAsync(o, t.p, sizeof(float) * t.numel(), cudaMemcpyDeviceToDevice, s));
      return (const float*)o;
    };
    h->obj_tab = table("obj_embeddings_ec.weight", d.num_objs);
    h->pred_tab = table("pred_embeddings_ec.weight", d.num_preds);
    if (d.manipulate_pred_dc) h->pred_tab_man = table("pred_embeddings_man_dc.weight", d.num_preds);   // EchoLayout.py:154
    echo_gcn_desc_t g = {};
    g.input_dim_obj = h->feat; g.input_dim_pred = h->feat; g.num_layers = d.num_layers; g.hidden_dim = 4 * h->gd;
    g.output_dim = h->feat; g.max_nodes = h->d.max_nodes; g.max_triples = h->d.max_triples; g.bn_eps = eps;
    g.keep_train_weights = d.keep_train_weights;
    h->bn_eps = eps;
    h->ec.create(wm, "gconv_net_ec.", g, h->pool);
    g.input_dim_obj = h->din_mani;
    g.num_layers = std::min(d.num_layers, 5);        // EchoScene.py:84
    h->mani.create(wm, "gconv_net_manipulation.", g, h->pool);
    h->has_rel_s = wm.has("rel_s_mlp.0.weight");     // absent in the layout-only model (EchoLayout.py)
    if (h->has_rel_s) {
      ECHO_CHECK(d.rel_s_hidden > 0 && d.context_dim > 0, "scene_create: rel_s_mlp dims missing");
      h->rel0 = folded_linear(wm, "rel_s_mlp.0", "rel_s_mlp.1", d.rel_s_hidden, h->feat, eps, h->pool, s);
      h->rel3 = plain_linear(wm, "rel_s_mlp.3", d.context_dim, d.rel_s_hidden, h->pool, s);
      h->rel_h = h->pool.alloc_n<float>((size_t)h->d.max_nodes * d.rel_s_hidden);
      if (d.keep_train_weights && wm.has("rel_s_mlp.1.running_mean")) {
        h->rel0_raw = plain_linear(wm, "rel_s_mlp.0", d.rel_s_hidden, h->feat, h->pool, s);
        auto vec = [&](const char* name) {
          const WView& v = wm.get(name, {d.rel_s_hidden});
          float* o = h->pool.alloc_n<float>(d.rel_s_hidden);
          ECHO_CUDA(cudaMemcpyAsync(o, v.p, sizeof(float) * d.rel_s_hidden, cudaMemcpyDeviceToDevice, s));
          return (const float*)o;
        };
        h->rel_bn_g = vec("rel_s_mlp.1.weight");
        h->rel_bn_b = vec("rel_s_mlp.1.bias");
      }
    }
    h->has_train = d.keep_train_weights != 0;
    const size_t N = h->d.max_nodes, T = h->d.max_triples;
    h->obj_embed = h->pool.alloc_n<float>(N * h->feat);
    h->pred_embed = h->pool.alloc_n<float>(T * h->feat);
    h->latent_obj = h->pool.alloc_n<float>(N * h->feat);
    h->mani_in = h->pool.alloc_n<float>(N * h->din_mani);
    h->latent = h->pool.alloc_n<float>(N * h->feat);
    if (h->pred_tab_man) h->pred_embed_man = h->pool.alloc_n<float>(T * h->feat);
    ECHO_CUDA(cudaStreamSynchronize(s));
  } catch (...) {
    h->pool.destroy();
    delete h;
    throw;
  }
  return h;
}

void scene_destroy(echo_scene* h) {
  if (!h) return;
  h->pool.destroy();
  delete h;
}

// init_encoder (EchoScene.py:143-157).  Outputs may be null (the handle's own buffers are used then).
void scene_init_encoder(echo_scene* h, const echo_graph* g, const int64_t* objs, const float* text_feat, const float* rel_feat,
                        float* obj_embed_out, float* pred_embed_out, float* latent_obj_out, cudaStream_t s) {
  check_graph(h, g, "scene_init_encoder");
  const int N = g->n_nodes, T = g->n_triples;
  ECHO_CHECK(objs && (h->add == 0 || (text_feat && (rel_feat || T == 0))), "scene_init_encoder: null input");
  float* oe = obj_embed_out ? obj_embed_out : h->obj_embed;
  float* pe = pred_embed_out ? pred_embed_out : h->pred_embed;
  float* lo = latent_obj_out ? latent_obj_out : h->latent_obj;
  embed_rows(h, h->obj_tab, objs, 1, 0, text_feat, N, oe, s);
  embed_rows(h, h->pred_tab, g->triples, 3, 1, rel_feat, T, pe, s);
  h->ec.forward(g, oe, pe, lo, nullptr, s, h->batch_stats);
}

// manipulate (EchoScene.py:181-195): latent_f (N, feat + gd) = [latent | change flag].
void scene_manipulate(echo_scene* h, const echo_graph* g, const float* latent_f, const int64_t* objs, const float* text_feat,
                      const float* rel_feat, float* latent_out, float* obj_embed_out, float* pred_embed_out, cudaStream_t s) {
  check_graph(h, g, "scene_manipulate");
  const int N = g->n_nodes, T = g->n_triples;
  ECHO_CHECK(latent_f && objs && (h->add == 0 || (text_feat && (rel_feat || T == 0))), "scene_manipulate: null input");
  float* oe = obj_embed_out ? obj_embed_out : h->obj_embed;
  float* pe = pred_embed_out ? pred_embed_out : h->pred_embed;
  float* lt = latent_out ? latent_out : h->latent;
  embed_rows(h, h->obj_tab, objs, 1, 0, text_feat, N, oe, s);
  embed_rows(h, h->pred_tab_man ? h->pred_tab_man : h->pred_tab, g->triples, 3, 1, rel_feat, T, pe, s);
  const int lf = h->feat + h->gd;
  copy_cols(latent_f, lf, N, lf, h->mani_in, h->din_mani, s);                  // torch.cat([latent_f, obj_embed], dim=1), :192
  copy_cols(oe, h->feat, N, h->feat, h->mani_in + lf, h->din_mani, s);
  h->mani.forward(g, h->mani_in, pe, lt, nullptr, s, h->batch_stats);
}

// rel_s_mlp (EchoScene.py:97-100): x (rows, feat) -> out (rows, context_dim)
void scene_rel_s(echo_scene* h, const float* x, int rows, float* out, cudaStream_t s) {
  ECHO_CHECK(h, "scene_rel_s: null handle");
  ECHO_CHECK(h->has_rel_s, "scene_rel_s: the state_dict has no rel_s_mlp (layout-only model)");
  ECHO_CHECK(rows >= 0 && rows <= h->d.max_nodes, "scene_rel_s: %d rows exceed handle capacity %d", rows, h->d.max_nodes);
  if (rows == 0) return;
  ECHO_CHECK(x && out, "scene_rel_s: null argument");
  LinArgs a;
  const bool bs = h->batch_stats && h->rel_bn_g;   // Linear -> BatchNorm1d on the statistics of these rows -> ReLU
  a.X = x; a.ldx = h->feat; a.M = rows; a.K = h->feat; a.nout = h->rel0.nout; a.W = bs ? h->rel0_raw.w : h->rel0.w;
  a.bias = bs ? h->rel0_raw.b : h->rel0.b; a.act = bs ? 0 : 1;
  a.Y = h->rel_h; a.ldy = h->rel0.nout;
  linear_auto(a, s);
  if (bs) bn_rows_train(h->rel_h, h->rel0.nout, rows, h->rel0.nout, h->rel_bn_g, h->rel_bn_b, h->bn_eps, s);
  a = LinArgs();
  a.X = h->rel_h; a.ldx = h->rel0.nout; a.M = rows; a.K = h->rel0.nout; a.nout = h->rel3.nout; a.W = h->rel3.w; a.bias = h->rel3.b;
  a.Y = out; a.ldy = h->rel3.nout;
  linear_auto(a, s);
}

// The encoder stage of Sg2ScDiffModel.sample (EchoScene.py:388-410) in one call.  `change` (N, gd) or null = zeros.
void scene_encode(echo_scene* h, const echo_graph* g, const int64_t* objs, const float* text_feat, const float* rel_feat,
                  const float* change, float* obj_embed_out, float* latent_out, float* uc_s_out, float* c_s_out, cudaStream_t s) {
  check_graph(h, g, "scene_encode");
  const int N = g->n_nodes, T = g->n_triples;
  ECHO_CHECK(objs && (h->add == 0 || (text_feat && (rel_feat || T == 0))), "scene_encode: null input");
  float* oe = obj_embed_out ? obj_embed_out : h->obj_embed;
  float* lt = latent_out ? latent_out : h->latent;
  // dec and enc graph are the same scene here, so the embeddings of init_encoder and manipulate coincide (except the
  // predicate table of the layout-only model, below)
  embed_rows(h, h->obj_tab, objs, 1, 0, text_feat, N, oe, s);
  embed_rows(h, h->pred_tab, g->triples, 3, 1, rel_feat, T, h->pred_embed, s);
  h->ec.forward(g, oe, h->pred_embed, h->latent_obj, nullptr, s, h->batch_stats);
  if (change) copy_cols(change, h->gd, N, h->gd, h->mani_in + h->feat, h->din_mani, s);
  else ECHO_CUDA(cudaMemsetAsync(h->mani_in, 0, sizeof(float) * (size_t)N * h->din_mani, s));   // the zero change flag, :393-397
  copy_cols(h->latent_obj, h->feat, N, h->feat, h->mani_in, h->din_mani, s);
  copy_cols(oe, h->feat, N, h->feat, h->mani_in + h->feat + h->gd, h->din_mani, s);
  const float* pe_man = h->pred_embed;
  if (h->pred_tab_man) {   // the layout-only model embeds the predicates of this stage with its own table
    embed_rows(h, h->pred_tab_man, g->triples, 3, 1, rel_feat, T, h->pred_embed_man, s);
    pe_man = h->pred_embed_man;
  }
  h->mani.forward(g, h->mani_in, pe_man, lt, nullptr, s, h->batch_stats);
  if (uc_s_out) scene_rel_s(h, oe, N, uc_s_out, s);
  if (c_s_out) scene_rel_s(h, lt, N, c_s_out, s);
}

void scene_set_batch_stats(echo_scene* h, bool on) {
  ECHO_CHECK(h, "scene_set_batch_stats: null handle");
  ECHO_CHECK(!on || h->has_train, "scene_set_batch_stats: the handle was created without keep_train_weights");
  h->batch_stats = on;
}

}  // namespace echo
