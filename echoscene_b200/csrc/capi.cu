// extern "C" boundary of libechoscene_b200 (include/echoscene_b200.h).  Exceptions never cross it: every entry
// point returns a code and leaves the message in thread-local storage.
#include <vector>
#include "unet.cuh"

#include <atomic>
#include <functional>
#include <algorithm>

struct echo_layout;
struct echo_shape;
struct echo_scene;
struct echo_gcn_train;

namespace echo {
void set_tc_mode(int);
const char* last_error();
echo_layout* layout_create(const echo_layout_desc_t*, const echo_weight_t*, int);
void layout_destroy(echo_layout*);
void layout_forward(echo_layout*, const echo_graph*, const float*, const float*, const int64_t*, float*, cudaStream_t);
void layout_step(echo_layout*, const echo_graph*, const float*, const float*, int, const float*, float*, cudaStream_t);
const std::vector<float>& layout_table(const echo_layout*);
void set_layout_mode(int);
void layout_info(const echo_layout*, int64_t*);
echo_shape* shape_create(const echo_shape_desc_t*, const echo_weight_t*, int);
void shape_destroy(echo_shape*);
void shape_forward(echo_shape*, const echo_graph*, const float*, const float*, const int64_t*, float*, cudaStream_t);
void shape_set_index(echo_shape*, int, cudaStream_t);
void shape_set_batch_stats(echo_shape*, bool);
size_t mesh_workspace_bytes(int R);
void mesh_marching_cubes(const float*, int, float, float*, int64_t, int*, int64_t, int*, void*, size_t, cudaStream_t);
echo_gcn_train* gcn_train_create(const echo_gcn_desc_t*, const echo_weight_t*, int, const echo_weight_t*, int);
void gcn_train_forward(echo_gcn_train*, const echo_graph*, const float*, const float*, float*, float*, cudaStream_t);
void gcn_train_backward(echo_gcn_train*, const echo_graph*, const float*, const float*, float*, float*, cudaStream_t);
void gcn_train_destroy(echo_gcn_train*);
void layout_set_batch_stats(echo_layout*, bool);
void scene_set_batch_stats(echo_scene*, bool);
echo_optimizer* optimizer_create(const echo_opt_tensor_t*, int);
void optimizer_destroy(echo_optimizer*);
void optimizer_step(echo_optimizer*, int64_t, double, double, double, double, double, double, cudaStream_t);
void optimizer_set_tensors(echo_optimizer*, const echo_opt_tensor_t*, int, cudaStream_t);
void optimizer_info(const echo_optimizer*, int64_t*, float*, cudaStream_t);
void shape_step(echo_shape*, const echo_graph*, const float*, const float*, int, float*, cudaStream_t);
void shape_embed(echo_shape*, const float*, int, float*, cudaStream_t);
void shape_trunk(echo_shape*, const echo_graph*, const float*, int, int, const float*, const float*, const int64_t*, int, float*,
                 cudaStream_t, cudaStream_t);
const float* shape_latent(const echo_shape*);
echo_vqvae* vqvae_create(const echo_vqvae_desc_t*, const echo_weight_t*, int);
void vqvae_destroy(echo_vqvae*);
void vqvae_decode(echo_vqvae*, const float*, int, float*, int*, cudaStream_t);
echo_vqvae* vqvae_encoder_create(const echo_vqvae_desc_t*, const echo_weight_t*, int);
void vqvae_encode(echo_vqvae*, const float*, int, float*, cudaStream_t);
int shape_context_dim(const echo_shape*);
void ddpm_tables(int, float, float, std::vector<float>&);
void ddim_schedule(int, int, float, float, std::vector<float>&, std::vector<int32_t>&);
echo_scene* scene_create(const echo_scene_desc_t*, const echo_weight_t*, int);
void scene_destroy(echo_scene*);
void scene_init_encoder(echo_scene*, const echo_graph*, const int64_t*, const float*, const float*, float*, float*, float*, cudaStream_t);
void scene_manipulate(echo_scene*, const echo_graph*, const float*, const int64_t*, const float*, const float*, float*, float*, float*,
                      cudaStream_t);
void scene_rel_s(echo_scene*, const float*, int, float*, cudaStream_t);
void scene_encode(echo_scene*, const echo_graph*, const int64_t*, const float*, const float*, const float*, float*, float*, float*,
                  float*, cudaStream_t);
void shape_tables(const echo_shape*, const std::vector<float>**, const std::vector<int32_t>**);
bool conv3d_small_cout_supported(int cin, int cout, int taps);
void conv3d_small_cout(const Act& x, const float* wt, const float* bias, int cout, float* out, cudaStream_t s);

static int guard(const std::function<void()>& f) {
  try {
    f();
    return ECHO_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return ECHO_ERR_INVALID;
  } catch (...) {
    set_last_error("unknown error");
    return ECHO_ERR_INVALID;
  }
}

// scratch for the single-operator entry points (tests / profiling only): grown on demand, freed at exit
struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
  void* get(size_t bytes) {
    if (bytes > cap) {
      if (p) { cudaDeviceSynchronize(); cudaFree(p); }
      ECHO_CUDA(cudaMalloc(&p, bytes));
      cap = bytes;
    }
    return p;
  }
};
static Scratch g_scratch[4];

// Host side of echo_graph_create: subject / object indices, CSR of the incident (triple, role) items per node and the
// predicate id range.  Item order per node = the order the reference's scatter_add visits them (graph.py:176-177): subject
// roles by ascending triple index, then object roles.  Throws for node indices outside [0, N) (graph.py:146-147 would raise).
static void build_csr(const int64_t* h, int T, int N, std::vector<int>& si, std::vector<int>& oi, std::vector<int>& off,
                      std::vector<int>& items, int64_t& p_lo, int64_t& p_hi) {
  si.assign(T, 0);
  oi.assign(T, 0);
  off.assign((size_t)N + 1, 0);
  items.assign((size_t)2 * T, 0);
  p_lo = 0;
  p_hi = -1;
  if (T) p_lo = p_hi = h[1];
  for (int t = 0; t < T; ++t) {
    const int64_t a = h[3 * t], b = h[3 * t + 2];
    ECHO_CHECK(a >= 0 && a < N && b >= 0 && b < N, "graph_create: triple %d has node index out of range [0, %d)", t, N);
    si[t] = (int)a;
    oi[t] = (int)b;
    p_lo = std::min(p_lo, h[3 * t + 1]);
    p_hi = std::max(p_hi, h[3 * t + 1]);
    off[a + 1]++;
    off[b + 1]++;
  }
  for (int n = 0; n < N; ++n) off[n + 1] += off[n];
  std::vector<int> fill(off.begin(), off.end() - 1);
  for (int t = 0; t < T; ++t) items[fill[si[t]]++] = t * 2 + 0;
  for (int t = 0; t < T; ++t) items[fill[oi[t]]++] = t * 2 + 1;
}

}  // namespace echo

using namespace echo;

extern "C" {

int echo_version(void) { return ECHO_ABI_VERSION; }
const char* echo_last_error(void) { return echo::last_error(); }
int echo_has_tcgen05(void) {
  int r = 0;
  guard([&] { r = tc_available() ? 1 : 0; });
  return r;
}
void echo_debug_set_tc_mode(int mode) { echo::set_tc_mode(mode); }
void echo_debug_set_layout_mode(int mode) { echo::set_layout_mode(mode); }
int echo_debug_layout_info(const echo_layout_t* h, int64_t* out6) {
  return guard([&] {
    ECHO_CHECK(h && out6, "layout_info: null argument");
    echo::layout_info(h, out6);
  });
}
void echo_debug_probe_begin(int64_t rows, int32_t cin, int32_t cout, int32_t ksize) { echo::tc_probe_begin(rows, cin, cout, ksize); }
int32_t echo_debug_probe_end(double* avg_ms) { return echo::tc_probe_end(avg_ms); }
int echo_debug_fold_upsample_weight(const float* w_host, int32_t cout, int32_t cin, int32_t up_depth, float* out_host) {
  return guard([&] {
    ECHO_CHECK(w_host && out_host && cout > 0 && cin > 0, "fold_upsample_weight: bad arguments");
    fold_upsample_weight(w_host, cout, cin, out_host, up_depth != 0);
  });
}
void echo_debug_tc_plan(int32_t n, int32_t d, int32_t h, int32_t w, int32_t cin, int32_t cout, int32_t ksize, int32_t epi, int32_t up2,
                        int32_t allow_splitk, int32_t sms, int32_t* out4) {
  echo::tc_plan_describe(n, d, h, w, cin, cout, ksize, epi, up2, allow_splitk, sms, out4);
}
void echo_debug_probe_timeline(void* buf_dev) { echo::tc_probe_timeline((unsigned long long*)buf_dev); }
int64_t echo_launch_count(void) { return g_launches; }
void echo_launch_count_reset(void) { g_launches = 0; }

int echo_debug_ddpm_tables(int32_t time_num, float beta_start, float beta_end, float* host_out) {
  return guard([&] {
    ECHO_CHECK(host_out, "debug_ddpm_tables: null output");
    std::vector<float> tab;
    ddpm_tables(time_num, beta_start, beta_end, tab);
    memcpy(host_out, tab.data(), tab.size() * sizeof(float));
  });
}
int echo_debug_ddim_schedule(int32_t timesteps, int32_t ddim_steps, float linear_start, float linear_end, int32_t capacity,
                             float* host_coef_out, int32_t* host_timesteps_out, int32_t* n_out) {
  return guard([&] {
    ECHO_CHECK(host_coef_out && host_timesteps_out && n_out, "debug_ddim_schedule: null output");
    std::vector<float> coef;
    std::vector<int32_t> ts;
    ddim_schedule(timesteps, ddim_steps, linear_start, linear_end, coef, ts);
    ECHO_CHECK((int)ts.size() <= capacity, "debug_ddim_schedule: %d steps exceed the caller's capacity %d", (int)ts.size(), capacity);
    memcpy(host_coef_out, coef.data(), coef.size() * sizeof(float));
    memcpy(host_timesteps_out, ts.data(), ts.size() * sizeof(int32_t));
    *n_out = (int32_t)ts.size();
  });
}

int echo_debug_graph_csr(const int64_t* triples_host, int32_t T, int32_t N, int32_t* node_off_out, int32_t* node_items_out,
                         int64_t* pred_range_out) {
  return guard([&] {
    ECHO_CHECK(N >= 0 && T >= 0 && (triples_host || T == 0) && node_off_out && (node_items_out || T == 0), "debug_graph_csr: bad arguments");
    std::vector<int> si, oi, off, items;
    int64_t lo = 0, hi = -1;
    build_csr(triples_host, T, N, si, oi, off, items, lo, hi);
    for (int n = 0; n <= N; ++n) node_off_out[n] = off[n];
    for (size_t i = 0; i < items.size(); ++i) node_items_out[i] = items[i];
    if (pred_range_out) { pred_range_out[0] = lo; pred_range_out[1] = hi; }
  });
}

int echo_graph_create(echo_graph_t** out, const int64_t* triples_dev, int32_t T, int32_t N, void* stream) {
  return guard([&] {
    ECHO_CHECK(out && N >= 0 && T >= 0 && (triples_dev || T == 0), "graph_create: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<int64_t> h((size_t)T * 3);
    if (T) {
      ECHO_CUDA(cudaMemcpyAsync(h.data(), triples_dev, sizeof(int64_t) * 3 * T, cudaMemcpyDeviceToHost, s));
      ECHO_CUDA(cudaStreamSynchronize(s));
    }
    std::vector<int> si, oi, off, items;
    int64_t p_lo = 0, p_hi = -1;
    build_csr(h.data(), T, N, si, oi, off, items, p_lo, p_hi);
    static std::atomic<uint64_t> next_id{1};
    echo_graph* g = new echo_graph();
    g->id = next_id++;
    g->n_nodes = N;
    g->p_min = p_lo;
    g->p_max = p_hi;
    g->n_triples = T;
    auto up = [&](const void* src, size_t bytes) {
      void* d = nullptr;
      ECHO_CUDA(cudaMalloc(&d, bytes ? bytes : 16));
      if (bytes) ECHO_CUDA(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
      return d;
    };
    try {
      g->triples = (int64_t*)up(h.data(), sizeof(int64_t) * 3 * T);
      g->s_idx = (int*)up(si.data(), sizeof(int) * T);
      g->o_idx = (int*)up(oi.data(), sizeof(int) * T);
      g->node_off = (int*)up(off.data(), sizeof(int) * (N + 1));
      g->node_items = (int*)up(items.data(), sizeof(int) * 2 * T);
    } catch (...) {
      echo_graph_destroy(g);  // frees whatever was uploaded (unset members are null)
      throw;
    }
    *out = g;
  });
}

void echo_graph_destroy(echo_graph_t* g) {
  if (!g) return;
  cudaFree(g->triples);
  cudaFree(g->s_idx);
  cudaFree(g->o_idx);
  cudaFree(g->node_off);
  cudaFree(g->node_items);
  delete g;
}

int echo_gather_rows(const float* obj, const int64_t* idx, int64_t n_idx, int64_t n_rows, int64_t dim, float* out, void* stream) {
  return guard([&] {
    ECHO_CHECK(obj && (out || n_idx == 0) && (idx || n_idx == 0) && dim > 0 && n_rows >= 0, "gather_rows: bad arguments");
    gather_rows(obj, idx, n_idx, n_rows, dim, out, (cudaStream_t)stream);
  });
}

int echo_gcn_create(echo_gcn_t** out, const echo_gcn_desc_t* desc, const echo_weight_t* weights, int32_t n_weights) {
  return guard([&] {
    ECHO_CHECK(out && desc, "gcn_create: null argument");
    echo_gcn* h = new echo_gcn();
    try {
      WeightMap wm;
      wm.load(weights, n_weights);
      echo_gcn_desc_t d = *desc;
      if (d.max_triples < 1) d.max_triples = 1;
      h->net.create(wm, "", d, h->pool);
    } catch (...) {
      h->pool.destroy();
      delete h;
      throw;
    }
    *out = h;
  });
}

int echo_gcn_forward(echo_gcn_t* h, const echo_graph_t* g, const float* obj, const float* pred, float* obj_out, float* pred_out,
                     void* stream) {
  return guard([&] {
    ECHO_CHECK(h && g && obj && (pred || g->n_triples == 0) && obj_out, "gcn_forward: null argument");
    h->net.forward(g, obj, pred, obj_out, pred_out, (cudaStream_t)stream);
  });
}
int echo_gcn_forward_train(echo_gcn_t* h, const echo_graph_t* g, const float* obj, const float* pred, float* obj_out, float* pred_out,
                           void* stream) {
  return guard([&] {
    ECHO_CHECK(h && g && obj && (pred || g->n_triples == 0) && obj_out, "gcn_forward_train: null argument");
    h->net.forward(g, obj, pred, obj_out, pred_out, (cudaStream_t)stream, true);
  });
}

void echo_gcn_destroy(echo_gcn_t* h) {
  if (!h) return;
  h->pool.destroy();
  delete h;
}

int echo_scene_create(echo_scene_t** out, const echo_scene_desc_t* desc, const echo_weight_t* weights, int32_t n_weights) {
  return guard([&] {
    ECHO_CHECK(out, "scene_create: null out");
    *out = scene_create(desc, weights, n_weights);
  });
}
int echo_scene_init_encoder(echo_scene_t* h, const echo_graph_t* g, const int64_t* objs, const float* text_feat, const float* rel_feat,
                            float* obj_embed_out, float* pred_embed_out, float* latent_obj_out, void* stream) {
  return guard([&] { scene_init_encoder(h, g, objs, text_feat, rel_feat, obj_embed_out, pred_embed_out, latent_obj_out, (cudaStream_t)stream); });
}
int echo_scene_manipulate(echo_scene_t* h, const echo_graph_t* g, const float* latent_f, const int64_t* objs, const float* text_feat,
                          const float* rel_feat, float* latent_out, float* obj_embed_out, float* pred_embed_out, void* stream) {
  return guard([&] {
    scene_manipulate(h, g, latent_f, objs, text_feat, rel_feat, latent_out, obj_embed_out, pred_embed_out, (cudaStream_t)stream);
  });
}
int echo_scene_rel_s(echo_scene_t* h, const float* x, int32_t rows, float* out, void* stream) {
  return guard([&] { scene_rel_s(h, x, rows, out, (cudaStream_t)stream); });
}
int echo_scene_encode(echo_scene_t* h, const echo_graph_t* g, const int64_t* objs, const float* text_feat, const float* rel_feat,
                      const float* change, float* obj_embed_out, float* latent_out, float* uc_s_out, float* c_s_out, void* stream) {
  return guard([&] {
    scene_encode(h, g, objs, text_feat, rel_feat, change, obj_embed_out, latent_out, uc_s_out, c_s_out, (cudaStream_t)stream);
  });
}
void echo_scene_destroy(echo_scene_t* h) { scene_destroy(h); }

int echo_layout_create(echo_layout_t** out, const echo_layout_desc_t* desc, const echo_weight_t* weights, int32_t n_weights) {
  return guard([&] {
    ECHO_CHECK(out, "layout_create: null out");
    *out = layout_create(desc, weights, n_weights);
  });
}
int echo_layout_forward(echo_layout_t* h, const echo_graph_t* g, const float* box_t, const float* obj_embed, const int64_t* t,
                        float* eps_out, void* stream) {
  return guard([&] {
    ECHO_CHECK(h && g && box_t && obj_embed && t && eps_out, "layout_forward: null argument");
    layout_forward(h, g, box_t, obj_embed, t, eps_out, (cudaStream_t)stream);
  });
}
int echo_layout_step(echo_layout_t* h, const echo_graph_t* g, const float* x_t, const float* obj_embed, int32_t t, const float* noise,
                     float* x_prev, void* stream) {
  return guard([&] {
    ECHO_CHECK(h && g && x_t && obj_embed && noise && x_prev, "layout_step: null argument");
    layout_step(h, g, x_t, obj_embed, t, noise, x_prev, (cudaStream_t)stream);
  });
}
void echo_layout_destroy(echo_layout_t* h) { layout_destroy(h); }
int echo_layout_schedule(const echo_layout_t* h, float* host_out) {
  return guard([&] {
    ECHO_CHECK(h && host_out, "layout_schedule: null argument");
    const std::vector<float>& t = layout_table(h);
    std::copy(t.begin(), t.end(), host_out);
  });
}

int echo_shape_create(echo_shape_t** out, const echo_shape_desc_t* desc, const echo_weight_t* weights, int32_t n_weights) {
  return guard([&] {
    ECHO_CHECK(out, "shape_create: null out");
    *out = shape_create(desc, weights, n_weights);
  });
}
int echo_shape_forward(echo_shape_t* h, const echo_graph_t* g, const float* x, const float* obj_embed, const int64_t* t, float* eps_out,
                       void* stream) {
  return guard([&] {
    ECHO_CHECK(h && g && x && obj_embed && t && eps_out, "shape_forward: null argument");
    shape_forward(h, g, x, obj_embed, t, eps_out, (cudaStream_t)stream);
  });
}
int echo_shape_step(echo_shape_t* h, const echo_graph_t* g, const float* x_t, const float* obj_embed, int32_t ddim_index, float* x_prev,
                    void* stream) {
  return guard([&] {
    ECHO_CHECK(h && g && x_t && obj_embed && x_prev, "shape_step: null argument");
    shape_step(h, g, x_t, obj_embed, ddim_index, x_prev, (cudaStream_t)stream);
  });
}
int echo_train_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac, const float* sqrt_1mac, int64_t rows,
                        int64_t row_len, float* out, void* stream) {
  return guard([&] {
    ECHO_CHECK(rows >= 0 && row_len >= 0, "q_sample: negative size");
    ECHO_CHECK(rows * row_len == 0 || (x0 && noise && t && sqrt_ac && sqrt_1mac && out), "q_sample: null argument");
    q_sample(x0, noise, t, sqrt_ac, sqrt_1mac, rows, row_len, out, (cudaStream_t)stream);
  });
}
int echo_train_mse_rows(const float* pred, const float* target, int64_t rows, int64_t row_len, const int32_t* ranges_host, int32_t n_ranges,
                        float* out, void* stream) {
  return guard([&] {
    ECHO_CHECK(rows >= 0 && row_len > 0 && n_ranges > 0 && n_ranges <= 16 && ranges_host, "mse_rows: bad arguments");
    ECHO_CHECK(rows == 0 || (pred && target && out), "mse_rows: null argument");
    for (int k = 0; k < n_ranges; ++k)
      ECHO_CHECK(ranges_host[2 * k] >= 0 && ranges_host[2 * k] < ranges_host[2 * k + 1] && ranges_host[2 * k + 1] <= row_len,
                 "mse_rows: range %d = [%d, %d) outside a row of %lld", k, ranges_host[2 * k], ranges_host[2 * k + 1], (long long)row_len);
    static thread_local int* d_ranges = nullptr;   // 32 ints, reused: the copy is stream-ordered in front of the kernel
    if (!d_ranges) ECHO_CUDA(cudaMalloc(&d_ranges, sizeof(int) * 32));
    ECHO_CUDA(cudaMemcpyAsync(d_ranges, ranges_host, sizeof(int) * 2 * n_ranges, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    mse_rows(pred, target, rows, row_len, d_ranges, n_ranges, out, (cudaStream_t)stream);
  });
}
int echo_optimizer_create(echo_optimizer_t** out, const echo_opt_tensor_t* tensors, int32_t n_tensors) {
  return guard([&] {
    ECHO_CHECK(out, "optimizer_create: null argument");
    *out = optimizer_create(tensors, n_tensors);
  });
}
int echo_optimizer_set_tensors(echo_optimizer_t* h, const echo_opt_tensor_t* tensors, int32_t n_tensors, void* stream) {
  return guard([&] {
    ECHO_CHECK(h, "optimizer_set_tensors: null handle");
    optimizer_set_tensors(h, tensors, n_tensors, (cudaStream_t)stream);
  });
}
int echo_optimizer_step(echo_optimizer_t* h, int64_t step, double lr, double beta1, double beta2, double eps, double weight_decay,
                        double clip_max_norm, void* stream) {
  return guard([&] {
    ECHO_CHECK(h, "optimizer_step: null handle");
    optimizer_step(h, step, lr, beta1, beta2, eps, weight_decay, clip_max_norm, (cudaStream_t)stream);
  });
}
int echo_optimizer_info(const echo_optimizer_t* h, int64_t* out4, float* clip2, void* stream) {
  return guard([&] {
    ECHO_CHECK(h && out4, "optimizer_info: null argument");
    optimizer_info(h, out4, clip2, (cudaStream_t)stream);
  });
}
void echo_optimizer_destroy(echo_optimizer_t* h) { optimizer_destroy(h); }
int echo_metrics_validate_constraints(const int64_t* triples, int64_t n_triples, const float* boxes, int64_t n_nodes, int32_t box_dim,
                                      const int32_t* keep, int32_t changes_mode, const int32_t* rel_of_pred_host, int32_t n_preds, int32_t strict,
                                      float overlap_threshold, int8_t* out_rel, int8_t* out_ok, void* stream) {
  return guard([&] {
    ECHO_CHECK(n_triples >= 0 && n_nodes >= 0 && (box_dim == 6 || box_dim == 7) && n_preds > 0 && n_preds <= 64 && rel_of_pred_host,
               "validate_constraints: bad arguments (box_dim %d, %d predicates)", box_dim, n_preds);
    ECHO_CHECK(n_triples == 0 || (triples && boxes && out_rel && out_ok), "validate_constraints: null argument");
    for (int i = 0; i < n_preds; ++i)
      ECHO_CHECK(rel_of_pred_host[i] >= -1 && rel_of_pred_host[i] <= ECHO_REL_SYMMETRICAL_TO, "validate_constraints: relation code %d of predicate %d",
                 rel_of_pred_host[i], i);
    static thread_local int32_t* d_rel = nullptr;   // 64 ints, reused; the copy is stream-ordered in front of the kernel
    if (!d_rel) ECHO_CUDA(cudaMalloc(&d_rel, sizeof(int32_t) * 64));
    ECHO_CUDA(cudaMemcpyAsync(d_rel, rel_of_pred_host, sizeof(int32_t) * n_preds, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    validate_constraints(triples, n_triples, boxes, n_nodes, box_dim, keep, changes_mode != 0, d_rel, n_preds, strict != 0, overlap_threshold,
                         out_rel, out_ok, (cudaStream_t)stream);
  });
}
int echo_gcn_train_create(echo_gcn_train_t** out, const echo_gcn_desc_t* desc, const echo_weight_t* params, int32_t n_params,
                          const echo_weight_t* grads, int32_t n_grads) {
  return guard([&] {
    ECHO_CHECK(out, "gcn_train_create: null out");
    *out = gcn_train_create(desc, params, n_params, grads, n_grads);
  });
}
int echo_gcn_train_forward(echo_gcn_train_t* h, const echo_graph_t* g, const float* obj_vecs, const float* pred_vecs, float* obj_out,
                           float* pred_out, void* stream) {
  return guard([&] { gcn_train_forward(h, g, obj_vecs, pred_vecs, obj_out, pred_out, (cudaStream_t)stream); });
}
int echo_gcn_train_backward(echo_gcn_train_t* h, const echo_graph_t* g, const float* d_obj_out, const float* d_pred_out, float* d_obj_in,
                            float* d_pred_in, void* stream) {
  return guard([&] { gcn_train_backward(h, g, d_obj_out, d_pred_out, d_obj_in, d_pred_in, (cudaStream_t)stream); });
}
void echo_gcn_train_destroy(echo_gcn_train_t* h) { gcn_train_destroy(h); }
int64_t echo_mesh_workspace_bytes(int32_t resolution) { return resolution >= 2 && resolution <= 512 ? (int64_t)mesh_workspace_bytes(resolution) : -1; }
int echo_mesh_marching_cubes(const float* sdf, int32_t resolution, float level, float* verts, int64_t max_verts, int32_t* faces,
                             int64_t max_faces, int32_t* counts, void* workspace, int64_t workspace_bytes, void* stream) {
  return guard([&] {
    mesh_marching_cubes(sdf, resolution, level, verts, max_verts, faces, max_faces, counts, workspace,
                        workspace_bytes > 0 ? (size_t)workspace_bytes : 0, (cudaStream_t)stream);
  });
}
int echo_layout_set_batch_stats(echo_layout_t* h, int32_t on) {
  return guard([&] { layout_set_batch_stats(h, on != 0); });
}
int echo_shape_set_batch_stats(echo_shape_t* h, int32_t on) {
  return guard([&] { shape_set_batch_stats(h, on != 0); });
}
int echo_scene_set_batch_stats(echo_scene_t* h, int32_t on) {
  return guard([&] { scene_set_batch_stats(h, on != 0); });
}
int echo_shape_set_index(echo_shape_t* h, int32_t ddim_index, void* stream) {
  return guard([&] {
    ECHO_CHECK(h, "shape_set_index: null handle");
    shape_set_index(h, ddim_index, (cudaStream_t)stream);
  });
}
int echo_shape_embed(echo_shape_t* h, const float* x_local, int32_t n_local, float* codes_out, void* stream) {
  return guard([&] {
    ECHO_CHECK(h && (x_local || n_local == 0) && (codes_out || n_local == 0), "shape_embed: null argument");
    shape_embed(h, x_local, n_local, codes_out, (cudaStream_t)stream);
  });
}
int echo_shape_trunk(echo_shape_t* h, const echo_graph_t* g, const float* x_local, int32_t obj_begin, int32_t n_local,
                     const float* codes_all, const float* obj_embed_all, const int64_t* t_all, int32_t ddim_index, float* out_local,
                     void* stream) {
  return guard([&] {
    ECHO_CHECK(h && g && codes_all && obj_embed_all && (n_local == 0 || (x_local && out_local)), "shape_trunk: null argument");
    shape_trunk(h, g, x_local, obj_begin, n_local, codes_all, obj_embed_all, t_all, ddim_index, out_local, (cudaStream_t)stream, nullptr);
  });
}
int echo_shape_trunk_async(echo_shape_t* h, const echo_graph_t* g, const float* x_local, int32_t obj_begin, int32_t n_local,
                           const float* codes_all, const float* obj_embed_all, const int64_t* t_all, int32_t ddim_index,
                           float* out_local, void* codes_stream, void* stream) {
  return guard([&] {
    ECHO_CHECK(h && g && codes_all && obj_embed_all && (n_local == 0 || (x_local && out_local)), "shape_trunk_async: null argument");
    shape_trunk(h, g, x_local, obj_begin, n_local, codes_all, obj_embed_all, t_all, ddim_index, out_local, (cudaStream_t)stream,
                (cudaStream_t)codes_stream);
  });
}
int echo_shape_latent(const echo_shape_t* h, int32_t n_nodes, float* out_dev, void* stream) {
  return guard([&] {
    ECHO_CHECK(h && out_dev && n_nodes >= 0, "shape_latent: bad arguments");
    ECHO_CUDA(cudaMemcpyAsync(out_dev, shape_latent(h), sizeof(float) * (size_t)n_nodes * shape_context_dim(h), cudaMemcpyDeviceToDevice,
                              (cudaStream_t)stream));
  });
}
void echo_shape_destroy(echo_shape_t* h) { shape_destroy(h); }

int echo_vqvae_create(echo_vqvae_t** out, const echo_vqvae_desc_t* desc, const echo_weight_t* weights, int32_t n_weights) {
  return guard([&] {
    ECHO_CHECK(out && desc && (weights || n_weights == 0), "vqvae_create: null argument");
    *out = vqvae_create(desc, weights, n_weights);
  });
}
int echo_vqvae_decode(echo_vqvae_t* h, const float* latents, int32_t n, float* sdf_out, int32_t* indices_out, void* stream) {
  return guard([&] {
    ECHO_CHECK(h && (n == 0 || (latents && sdf_out)), "vqvae_decode: null argument");
    vqvae_decode(h, latents, n, sdf_out, indices_out, (cudaStream_t)stream);
  });
}
int echo_vqvae_encoder_create(echo_vqvae_t** out, const echo_vqvae_desc_t* desc, const echo_weight_t* weights, int32_t n_weights) {
  return guard([&] {
    ECHO_CHECK(out && desc && (weights || n_weights == 0), "vqvae_encoder_create: null argument");
    *out = vqvae_encoder_create(desc, weights, n_weights);
  });
}
int echo_vqvae_encode(echo_vqvae_t* h, const float* sdf, int32_t n, float* latents_out, void* stream) {
  return guard([&] {
    ECHO_CHECK(h && (n == 0 || (sdf && latents_out)), "vqvae_encode: null argument");
    vqvae_encode(h, sdf, n, latents_out, (cudaStream_t)stream);
  });
}
void echo_vqvae_destroy(echo_vqvae_t* h) { vqvae_destroy(h); }
int echo_shape_schedule(const echo_shape_t* h, float* host_coef_out, int32_t* host_ts_out) {
  return guard([&] {
    ECHO_CHECK(h, "shape_schedule: null handle");
    const std::vector<float>* c;
    const std::vector<int32_t>* t;
    shape_tables(h, &c, &t);
    if (host_coef_out) std::copy(c->begin(), c->end(), host_coef_out);
    if (host_ts_out) std::copy(t->begin(), t->end(), host_ts_out);
  });
}

// ---- single operators ---------------------------------------------------------------------------------------------
int echo_op_conv3d(const float* x, int32_t n, int32_t d, int32_t h, int32_t w, int32_t cin, const float* weight, const float* bias,
                   int32_t cout, int32_t ksize, int32_t stride_d, int32_t stride_hw, float* out, int32_t precision, void* stream) {
  return guard([&] {
    ECHO_CHECK(x && weight && out && (ksize == 1 || ksize == 3) && stride_d == 1 && (stride_hw == 1 || stride_hw == 2),
               "op_conv3d: unsupported arguments");
    cudaStream_t s = (cudaStream_t)stream;
    const int taps = ksize * ksize * ksize;
    const size_t wn = (size_t)cout * cin * taps;
    float* wr = (float*)g_scratch[0].get(wn * sizeof(float));
    if (taps == 1) ECHO_CUDA(cudaMemcpyAsync(wr, weight, wn * sizeof(float), cudaMemcpyDeviceToDevice, s));
    else repack_conv_weight(weight, cout, cin, taps, wr, s);
    const int pad = ksize / 2;
    GemmArgs g;
    g.n = n; g.d = d; g.h = h; g.w = w; g.cin = cin; g.lda = cin;
    g.od = d; g.oh = (h + 2 * pad - ksize) / stride_hw + 1; g.ow = (w + 2 * pad - ksize) / stride_hw + 1;
    g.kd = g.kh = g.kw = ksize; g.sh = g.sw = stride_hw; g.pd = g.ph = g.pw = pad;
    g.w_stride_n = (int64_t)taps * cin; g.cout = cout; g.bias = bias; g.out = out; g.ldo = cout;
    const int64_t rows_in = (int64_t)n * d * h * w, rows_out = g.rows_out();
    if (precision == ECHO_PREC_BF16) {
      ECHO_CHECK(tc_available(), "op_conv3d: bf16 precision needs the tcgen05 kernels");
      __nv_bfloat16* xb = (__nv_bfloat16*)g_scratch[1].get((size_t)rows_in * cin * 2);
      __nv_bfloat16* wb = (__nv_bfloat16*)g_scratch[2].get(wn * 2);
      __nv_bfloat16* ob = (__nv_bfloat16*)g_scratch[3].get((size_t)rows_out * cout * 2);
      convert(x, F32, xb, BF16, rows_in * cin, s);
      convert(wr, F32, wb, BF16, (int64_t)wn, s);
      g.A = xb; g.a_dt = BF16; g.W = wb; g.w_dt = BF16; g.out = ob; g.out_dt = BF16;
      static Scratch s2d;
      if (stride_hw == 2) g.scratch = s2d.get((size_t)rows_in * cin * 2);
      static Scratch skws;
      if (const size_t wsb = gemm_tc_splitk_ws_bytes(g)) g.splitk_ws = skws.get(wsb);
      ECHO_CHECK(gemm_tc_supported(g), "op_conv3d: shape not supported by the tcgen05 kernel");
      gemm_tc(g, s);
      convert(ob, BF16, out, F32, rows_out * cout, s);
    } else if (ksize == 3 && stride_hw == 1 && conv3d_small_cout_supported(cin, cout, taps)) {
      Act xa;
      xa.p = (void*)x; xa.dt = F32; xa.n = n; xa.d = d; xa.h = h; xa.w = w; xa.c = cin;
      conv3d_small_cout(xa, wr, bias, cout, out, s);
    } else {
      g.A = x; g.W = wr;
      gemm_simt(g, s);
    }
  });
}

static void op_upconv3d_impl(const float* x, int32_t n, int32_t d, int32_t h, int32_t w, int32_t cin, const float* weight, const float* bias,
                            int32_t cout, float* out, int32_t precision, void* stream, bool up_depth);

int echo_op_upconv3d(const float* x, int32_t n, int32_t d, int32_t h, int32_t w, int32_t cin, const float* weight, const float* bias,
                     int32_t cout, float* out, int32_t precision, void* stream) {
  return guard([&] { op_upconv3d_impl(x, n, d, h, w, cin, weight, bias, cout, out, precision, stream, false); });
}

int echo_op_upconv3d_x2(const float* x, int32_t n, int32_t d, int32_t h, int32_t w, int32_t cin, const float* weight, const float* bias,
                        int32_t cout, float* out, int32_t precision, void* stream) {
  return guard([&] { op_upconv3d_impl(x, n, d, h, w, cin, weight, bias, cout, out, precision, stream, true); });
}

static void op_upconv3d_impl(const float* x, int32_t n, int32_t d, int32_t h, int32_t w, int32_t cin, const float* weight, const float* bias,
                            int32_t cout, float* out, int32_t precision, void* stream, bool up_depth) {
  {
    ECHO_CHECK(x && weight && out, "op_upconv3d: bad arguments");
    ECHO_CHECK(precision == ECHO_PREC_BF16 && tc_available(), "op_upconv3d: the folded upsample conv is a tcgen05 (ECHO_PREC_BF16) kernel");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t wn = (size_t)cout * cin * 27, fn = (size_t)cout * cin * (up_depth ? 64 : 48);
    float* wr = (float*)g_scratch[0].get(wn * sizeof(float));
    repack_conv_weight(weight, cout, cin, 27, wr, s);
    std::vector<float> hsrc(wn), hdst(fn);
    ECHO_CUDA(cudaMemcpyAsync(hsrc.data(), wr, wn * sizeof(float), cudaMemcpyDeviceToHost, s));
    ECHO_CUDA(cudaStreamSynchronize(s));
    fold_upsample_weight(hsrc.data(), cout, cin, hdst.data(), up_depth);
    static Scratch fold32;
    float* wf = (float*)fold32.get(fn * sizeof(float));
    ECHO_CUDA(cudaMemcpyAsync(wf, hdst.data(), fn * sizeof(float), cudaMemcpyHostToDevice, s));
    const int64_t rows_in = (int64_t)n * d * h * w, rows_out = rows_in * (up_depth ? 8 : 4);
    __nv_bfloat16* xb = (__nv_bfloat16*)g_scratch[1].get((size_t)rows_in * cin * 2);
    __nv_bfloat16* wb = (__nv_bfloat16*)g_scratch[2].get(fn * 2);
    __nv_bfloat16* ob = (__nv_bfloat16*)g_scratch[3].get((size_t)rows_out * cout * 2);
    convert(x, F32, xb, BF16, rows_in * cin, s);
    convert(wf, F32, wb, BF16, (int64_t)fn, s);
    GemmArgs g;
    g.A = xb; g.a_dt = BF16; g.n = n; g.d = d; g.h = h; g.w = w; g.cin = cin; g.lda = cin;
    g.od = up_depth ? 2 * d : d; g.oh = 2 * h; g.ow = 2 * w; g.up2 = up_depth ? 2 : 1;
    g.kd = g.kh = g.kw = 3; g.pd = g.ph = g.pw = 1;
    g.W = wb; g.w_dt = BF16; g.w_stride_n = (int64_t)(up_depth ? 64 : 48) * cin; g.cout = cout; g.bias = bias;
    g.out = ob; g.out_dt = BF16; g.ldo = cout;
    ECHO_CHECK(gemm_tc_supported(g), "op_upconv3d: shape not supported by the tcgen05 kernel");
    gemm_tc(g, s);
    convert(ob, BF16, out, F32, rows_out * cout, s);
    ECHO_CUDA(cudaStreamSynchronize(s));   // hsrc / hdst are stack-scoped host staging
  }
}

int echo_op_linear(const float* x, int64_t rows, int32_t cin, const float* weight, const float* bias, int32_t cout, float* out,
                   int32_t precision, void* stream) {
  return guard([&] {
    ECHO_CHECK(x && weight && out && rows >= 0, "op_linear: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == ECHO_PREC_BF16) {
      ECHO_CHECK(tc_available(), "op_linear: bf16 precision needs the tcgen05 kernels");
      __nv_bfloat16* xb = (__nv_bfloat16*)g_scratch[1].get((size_t)rows * cin * 2);
      __nv_bfloat16* wb = (__nv_bfloat16*)g_scratch[2].get((size_t)cout * cin * 2);
      __nv_bfloat16* ob = (__nv_bfloat16*)g_scratch[3].get((size_t)rows * cout * 2);
      convert(x, F32, xb, BF16, rows * cin, s);
      convert(weight, F32, wb, BF16, (int64_t)cout * cin, s);
      GemmArgs g;
      g.A = xb; g.a_dt = BF16; g.n = 1; g.w = (int)rows; g.ow = (int)rows; g.cin = cin; g.lda = cin;
      g.W = wb; g.w_dt = BF16; g.w_stride_n = cin; g.cout = cout; g.bias = bias; g.out = ob; g.out_dt = BF16; g.ldo = cout;
      ECHO_CHECK(gemm_tc_supported(g), "op_linear: shape not supported by the tcgen05 kernel");
      gemm_tc(g, s);
      convert(ob, BF16, out, F32, rows * cout, s);
      return;
    }
    LinArgs a;
    a.X = x; a.ldx = cin; a.M = (int)rows; a.K = cin; a.nout = cout; a.W = weight; a.bias = bias; a.Y = out; a.ldy = cout;
    linear_auto(a, s);
  });
}

int echo_op_group_norm(const float* x, int32_t n, int64_t voxels, int32_t c, int32_t groups, const float* gamma, const float* beta,
                       float eps, int32_t silu, float* out, void* stream) {
  return guard([&] {
    ECHO_CHECK(x && out && gamma && beta && groups == 32, "op_group_norm: bad arguments (groups must be 32)");
    cudaStream_t s = (cudaStream_t)stream;
    if (voxels == 1) {
      gn_rows(x, n, c, groups, gamma, beta, eps, silu != 0, out, s);
      return;
    }
    Act xa, oa;
    xa.p = (void*)x; xa.n = n; xa.d = 1; xa.h = 1; xa.w = (int)voxels; xa.c = c;
    oa = xa;
    oa.p = out;
    float* ws = (float*)g_scratch[0].get((gn_partial_floats(xa, groups) + (size_t)n * groups * 2) * sizeof(float));
    float* stats = ws + gn_partial_floats(xa, groups);
    gn_stats(xa, groups, eps, stats, ws, s);
    gn_apply(xa, stats, gamma, beta, groups, silu != 0, oa, s);
  });
}

int echo_op_layer_norm(const float* x, int64_t rows, int32_t c, const float* gamma, const float* beta, float eps, float* out,
                       void* stream) {
  return guard([&] {
    ECHO_CHECK(x && out && gamma && beta, "op_layer_norm: null argument");
    layer_norm(x, F32, rows, c, gamma, beta, eps, out, F32, (cudaStream_t)stream);
  });
}

int echo_op_attention(const float* qkv, int32_t n, int32_t tokens, int32_t heads, int32_t dh, float* out, int32_t precision,
                      void* stream) {
  return guard([&] {
    ECHO_CHECK(qkv && out, "op_attention: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int C = heads * dh;
    const int64_t rows = (int64_t)n * tokens;
    if (precision == ECHO_PREC_BF16) {
      ECHO_CHECK(tc_available(), "op_attention: bf16 precision needs sm_100a");
      ECHO_CHECK(attention_bf16_supported(tokens, dh), "op_attention: tokens=%d dh=%d not supported by the flash kernel", tokens, dh);
      const int dhp = attention_pad_dh(dh);
      __nv_bfloat16* ob = (__nv_bfloat16*)g_scratch[3].get((size_t)rows * C * 2);
      if (dhp == 64 && attention_tc_supported(tokens, dh)) {   // the tcgen05 kernel (what the shape step runs at 1024 tokens)
        __nv_bfloat16* qk = (__nv_bfloat16*)g_scratch[1].get((size_t)rows * 2 * heads * 64 * 2);
        __nv_bfloat16* vt = (__nv_bfloat16*)g_scratch[2].get((size_t)rows * heads * 64 * 2);
        split_qkv_tc(qkv, n, tokens, heads, dh, qk, vt, s);
        attention_tc(qk, vt, n, tokens, heads, dh, ob, s);
        convert(ob, BF16, out, F32, rows * C, s);
        ECHO_CUDA(cudaStreamSynchronize(s));
        ECHO_CHECK(attention_tc_error() == 0, "op_attention: tcgen05 kernel found a misaligned shared-memory window");
        return;
      }
      __nv_bfloat16* qb = (__nv_bfloat16*)g_scratch[1].get((size_t)rows * 3 * heads * dhp * 2);
      pad_qkv(qkv, rows, heads, dh, dhp, qb, s);
      attention_bf16(qb, n, tokens, heads, dh, ob, s);
      convert(ob, BF16, out, F32, rows * C, s);
      return;
    }
    float* ws = (float*)g_scratch[0].get(attention_f32_ws_floats(n, tokens, heads) * sizeof(float));
    attention_f32(qkv, n, tokens, heads, dh, ws, out, s);
  });
}

}  // extern "C"
