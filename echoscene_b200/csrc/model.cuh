// Host-side model structures behind the C ABI: named weights, the scene-graph CSR, the GraphTripleConvNet
// executor and the two denoiser-step executors.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "ops.cuh"

namespace echo {

struct WView {
  const float* p = nullptr;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

struct WeightMap {
  std::unordered_map<std::string, WView> m;
  void load(const echo_weight_t* w, int n);
  bool has(const std::string& k) const { return m.find(k) != m.end(); }
  const WView& get(const std::string& k) const;
  const WView& get(const std::string& k, std::initializer_list<int64_t> shape) const;   // shape-checked
};

// device memory owned by a handle (prepared weights, tables); grows in 64 MiB slabs
struct DevPool {
  std::vector<void*> slabs;
  char* cur = nullptr;
  size_t left = 0, total = 0;
  void* alloc(size_t bytes);
  template <class T>
  T* alloc_n(size_t n) { return (T*)alloc(n * sizeof(T)); }
  float* upload(const std::vector<float>& h);
  void destroy();
};

// a prepared [nout, K] matrix (+ bias), fp32 and optionally bf16
struct Mat {
  const float* w = nullptr;
  const __nv_bfloat16* wb = nullptr;
  const float* b = nullptr;
  int nout = 0, K = 0;
};

}  // namespace echo

// ---- opaque C handles ---------------------------------------------------------------------------------------------
struct echo_graph {
  uint64_t id = 0;              // unique per created graph (keys cached CUDA graphs; pointers can be recycled)
  int n_nodes = 0, n_triples = 0;
  int64_t* triples = nullptr;   // (T,3) int64 device copy
  int* s_idx = nullptr;         // (T)
  int* o_idx = nullptr;         // (T)
  int* node_off = nullptr;      // (N+1) CSR offsets
  int* node_items = nullptr;    // (2T) item = t*2 + role(0 = subject, 1 = object), subjects first, ascending t
  int64_t p_min = 0, p_max = -1;   // range of the predicate ids (host-side, from graph_create): embedding lookups check it
};

namespace echo {

struct GcnLayer {
  int din = 0, dp = 0, H = 0, dout = 0;
  Mat w_so;    // [2H, din]   net1.0 (BN folded), subject rows then object rows, no bias
  Mat w_p;     // [H, dp]     net1.0 predicate columns, no bias
  const float* b1 = nullptr;  // [H] folded bias of net1.0
  Mat w2;      // [2H+dp, H]  net1.3 (BN folded)
  Mat w3;      // [H, H]      net2.0
  Mat w4;      // [dout, H]   net2.3
  Mat proj;    // [dout, din] linear_projection
  Mat projp;   // [dp, dp]    linear_projection_pred
  bool residual = true;
  // batch-statistics mode (echo_gcn_forward_train): the same four Linears UNFOLDED + their BatchNorm1d scale / shift
  bool has_train = false;
  Mat t_so, t_p, t2, t3, t4;          // raw weights, same row layout as the folded ones (t_so / t_p: no bias)
  const float* t_b1 = nullptr;        // raw bias of net1.0
  const float *bn_g[4] = {nullptr, nullptr, nullptr, nullptr}, *bn_b[4] = {nullptr, nullptr, nullptr, nullptr};
};

struct Gcn {
  std::vector<GcnLayer> layers;
  int max_nodes = 0, max_triples = 0, max_d = 0, dp = 0, H = 0;
  float *pso = nullptr, *pp = nullptr, *h1 = nullptr, *t2 = nullptr, *pooled = nullptr, *n1 = nullptr, *proj = nullptr;
  float *obj_pp[2] = {nullptr, nullptr}, *pred_pp[2] = {nullptr, nullptr};
  void create(const WeightMap& wm, const std::string& prefix, const echo_gcn_desc_t& d, DevPool& pool);
  // obj [N, din0], pred [T, dp] -> obj_out [N, dout_last], pred_out [T, dp] (either may alias internal buffers)
  void forward(const echo_graph* g, const float* obj, const float* pred, float* obj_out, float* pred_out, cudaStream_t s,
               bool batch_stats = false);
  float bn_eps = 1e-5f;
};

void linear_auto(const LinArgs& a, cudaStream_t s);
// BatchNorm1d on the statistics of the `rows` rows of x (biased variance) + ReLU, in place (gcn.cu)
void bn_rows_train(float* x, int64_t ld, int rows, int C, const float* gamma, const float* beta, float eps, cudaStream_t s);
// Linear (+ eval BatchNorm1d folded in when `bn`.running_mean exists) / plain Linear copied into the handle's pool (gcn.cu)
Mat folded_linear(const WeightMap& wm, const std::string& lin, const std::string& bn, int nout, int K, float eps, DevPool& pool,
                  cudaStream_t s);
Mat plain_linear(const WeightMap& wm, const std::string& lin, int nout, int K, DevPool& pool, cudaStream_t s);

}  // namespace echo

struct echo_gcn {
  echo::DevPool pool;
  echo::Gcn net;
};
