// GraphTripleConv "echo" message passing (model/graph.py:124-211) as B200 kernels.
//
// Reference per layer: gather obj[s], obj[o] (T x Din each) -> cat with pred -> net1 (Linear, BN, ReLU) x2 ->
// split -> scatter_add into nodes -> /degree -> net2 -> + residual projections.
//
// Here (eval mode, BatchNorm folded into the Linear at create time):
//   1. PROJECT NODES FIRST: [Ps | Po] = obj @ [W1_s ; W1_o]^T  (N rows instead of T; net1's first Linear is
//      linear in the concat, so W1 [s|p|o] columns are applied before the gather),  Pp = pred @ W1_p^T.
//   2. edge_combine: one WARP PER EDGE reads rows Ps[s[t]], Po[o[t]], Pp[t] (H = 256 floats = two coalesced
//      16-byte loads per lane each), adds the folded bias, ReLU -> h1[t].  The gather is a pure row copy of
//      projected features; indices come from the CSR built once per graph (edges are constant over the chain).
//   3. t2 = ReLU(h1 @ W2'^T + b2')  (T x (2H+Dp)).
//   4. node_pool: one warp-group per node walks its CSR item list in a FIXED order (all subject roles by
//      ascending t, then object roles) and averages — deterministic, no float atomics, same summation order as
//      the reference's CPU scatter_add.
//   5. net2 (two few-row linears) + residual projections (graph.py:205-209).
#include "model.cuh"

#include <algorithm>
#include <cstdlib>

namespace echo {
namespace {

__global__ void edge_combine_kernel(const float* __restrict__ pso, const float* __restrict__ pp, const float* __restrict__ b1,
                                    const int* __restrict__ s_idx, const int* __restrict__ o_idx, int T, int H,
                                    float* __restrict__ h1, float floor_) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= T) return;
  const float4* ps = reinterpret_cast<const float4*>(pso + (int64_t)s_idx[t] * 2 * H);
  const float4* po = reinterpret_cast<const float4*>(pso + (int64_t)o_idx[t] * 2 * H + H);
  const float4* pq = reinterpret_cast<const float4*>(pp + (int64_t)t * H);
  const float4* bb = reinterpret_cast<const float4*>(b1);
  float4* out = reinterpret_cast<float4*>(h1 + (int64_t)t * H);
  for (int q = lane; q < H / 4; q += 32) {
    const float4 a = __ldg(ps + q), b = __ldg(pq + q), c = __ldg(po + q), d = __ldg(bb + q);
    float4 r;
    r.x = fmaxf(((a.x + b.x) + c.x) + d.x, floor_);   // floor_ = 0: ReLU; -inf: the pre-activation (batch-statistics mode)
    r.y = fmaxf(((a.y + b.y) + c.y) + d.y, floor_);
    r.z = fmaxf(((a.z + b.z) + c.z) + d.z, floor_);
    r.w = fmaxf(((a.w + b.w) + c.w) + d.w, floor_);
    out[q] = r;
  }
}

// The same with the projected node table [N][2H] staged in shared memory first (scene-sized graphs: N * 2H floats fit): every node
// row is read from HBM / L2 once per block, coalesced, instead of once per incident edge; a warp then combines one edge from
// shared memory.  Identical arithmetic and association, so the two kernels are interchangeable bit for bit.
__global__ void edge_combine_staged_kernel(const float* __restrict__ pso, const float* __restrict__ pp, const float* __restrict__ b1,
                                           const int* __restrict__ s_idx, const int* __restrict__ o_idx, int T, int N, int H,
                                           float* __restrict__ h1, float floor_) {
  extern __shared__ float4 node_s[];   // [N][2H / 4]
  const int row4 = 2 * H / 4;
  for (int i = threadIdx.x; i < N * row4; i += blockDim.x) node_s[i] = __ldg(reinterpret_cast<const float4*>(pso) + i);
  __syncthreads();
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int t = blockIdx.x * warps + (threadIdx.x >> 5); t < T; t += gridDim.x * warps) {
    const float4* ps = node_s + (size_t)s_idx[t] * row4;
    const float4* po = node_s + (size_t)o_idx[t] * row4 + H / 4;
    const float4* pq = reinterpret_cast<const float4*>(pp + (int64_t)t * H);
    const float4* bb = reinterpret_cast<const float4*>(b1);
    float4* out = reinterpret_cast<float4*>(h1 + (int64_t)t * H);
    for (int q = lane; q < H / 4; q += 32) {
      const float4 a = ps[q], b = __ldg(pq + q), c = po[q], d = __ldg(bb + q);
      float4 r;
      r.x = fmaxf(((a.x + b.x) + c.x) + d.x, floor_);
      r.y = fmaxf(((a.y + b.y) + c.y) + d.y, floor_);
      r.z = fmaxf(((a.z + b.z) + c.z) + d.z, floor_);
      r.w = fmaxf(((a.w + b.w) + c.w) + d.w, floor_);
      out[q] = r;
    }
  }
}

// BatchNorm1d on the statistics of the batch (the rows of the call), then ReLU, in place: y = (x - mean) rsqrt(var + eps) g + b with
// the BIASED variance, as torch normalises in training mode (model/layers.py:29-30 under model.train()).  Block = 32 columns x 8
// row lanes, two passes over the rows (mean, then centred squares), fixed summation order.
__global__ void bn_rows_train_kernel(float* __restrict__ x, int64_t ld, int rows, int C, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = threadIdx.y;
  const bool in = c < C;
  float s = 0.f;
  if (in) for (int r = r0; r < rows; r += 8) s += x[(int64_t)r * ld + c];
  red[r0][threadIdx.x] = s;
  __syncthreads();
  float mean = 0.f;
  for (int k = 0; k < 8; ++k) mean += red[k][threadIdx.x];
  mean /= (float)rows;
  __syncthreads();
  float q = 0.f;
  if (in) for (int r = r0; r < rows; r += 8) { const float d = x[(int64_t)r * ld + c] - mean; q = fmaf(d, d, q); }
  red[r0][threadIdx.x] = q;
  __syncthreads();
  float var = 0.f;
  for (int k = 0; k < 8; ++k) var += red[k][threadIdx.x];
  var /= (float)rows;
  if (!in) return;
  const float sc = rsqrtf(var + eps) * gamma[c], sh = beta[c];
  for (int r = r0; r < rows; r += 8) {
    const int64_t i = (int64_t)r * ld + c;
    x[i] = fmaxf((x[i] - mean) * sc + sh, 0.f);
  }
}

// pooled[n, :] = (sum over items of t2[t, role ? H+Dp : 0 ...]) / max(count, 1)     (graph.py:161-199)
__global__ void node_pool_kernel(const float* __restrict__ t2, int ld, int H, int off_o, const int* __restrict__ node_off,
                                 const int* __restrict__ node_items, int N, float* __restrict__ pooled) {
  const int n = blockIdx.x;
  if (n >= N) return;
  const int beg = node_off[n], end = node_off[n + 1];
  const float cnt = fmaxf((float)(end - beg), 1.f);
  for (int q = threadIdx.x; q < H / 4; q += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = beg; i < end; ++i) {
      const int item = node_items[i], t = item >> 1, role = item & 1;
      const float4 v = __ldg(reinterpret_cast<const float4*>(t2 + (int64_t)t * ld + (role ? off_o : 0)) + q);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x /= cnt; acc.y /= cnt; acc.z /= cnt; acc.w /= cnt;
    reinterpret_cast<float4*>(pooled + (int64_t)n * H)[q] = acc;
  }
}

}  // namespace

Mat folded_linear(const WeightMap& wm, const std::string& lin, const std::string& bn, int nout, int K, float eps, DevPool& pool,
                  cudaStream_t s) {
  const WView& w = wm.get(lin + ".weight", {nout, K});
  const WView& b = wm.get(lin + ".bias", {nout});
  Mat m;
  m.nout = nout;
  m.K = K;
  float* wo = pool.alloc_n<float>((size_t)nout * K);
  float* bo = pool.alloc_n<float>(nout);
  if (wm.has(bn + ".running_mean")) {
    fold_bn(w.p, b.p, wm.get(bn + ".weight", {nout}).p, wm.get(bn + ".bias", {nout}).p, wm.get(bn + ".running_mean", {nout}).p,
            wm.get(bn + ".running_var", {nout}).p, eps, nout, K, wo, bo, s);
  } else {
    ECHO_CUDA(cudaMemcpyAsync(wo, w.p, sizeof(float) * nout * K, cudaMemcpyDeviceToDevice, s));
    ECHO_CUDA(cudaMemcpyAsync(bo, b.p, sizeof(float) * nout, cudaMemcpyDeviceToDevice, s));
  }
  m.w = wo;
  m.b = bo;
  return m;
}

Mat plain_linear(const WeightMap& wm, const std::string& lin, int nout, int K, DevPool& pool, cudaStream_t s) {
  const WView& w = wm.get(lin + ".weight", {nout, K});
  const WView& b = wm.get(lin + ".bias", {nout});
  Mat m;
  m.nout = nout;
  m.K = K;
  float* wo = pool.alloc_n<float>((size_t)nout * K);
  float* bo = pool.alloc_n<float>(nout);
  ECHO_CUDA(cudaMemcpyAsync(wo, w.p, sizeof(float) * nout * K, cudaMemcpyDeviceToDevice, s));
  ECHO_CUDA(cudaMemcpyAsync(bo, b.p, sizeof(float) * nout, cudaMemcpyDeviceToDevice, s));
  m.w = wo;
  m.b = bo;
  return m;
}

void bn_rows_train(float* x, int64_t ld, int rows, int C, const float* gamma, const float* beta, float eps, cudaStream_t s) {
  if (rows == 0 || C == 0) return;
  bn_rows_train_kernel<<<cdiv(C, 32), dim3(32, 8), 0, s>>>(x, ld, rows, C, gamma, beta, eps);
  ECHO_LAUNCH_CHECK();
}

// Few-row kernel (HBM-bound on the weights, linear.cu) up to 64 rows.  Above -- collated batches: hundreds of nodes, thousands of
// edges -- the contraction is compute-bound and runs on the tensor cores at fp32 grade (3 x TF32, sgemm_x3.cu); a prologue
// (GroupNorm / LayerNorm / GEGLU / SiLU / concat) is first materialised into the caller's scratch rows.  Without scratch, or with
// bf16 weights, the previous routes remain: the few-row kernel at any M for prologue layers, the SIMT GEMM for plain ones.
void linear_auto(const LinArgs& a, cudaStream_t s) {
  if (a.M <= 64) {
    linear_rows(a, s);
    return;
  }
  const bool plain = a.in_act == 0 && a.pro == PRO_NONE && !a.X2;
  static const bool no_x3 = getenv("ECHO_NO_X3_LINEAR") != nullptr;
  if (a.w_dt == F32 && !no_x3 && (plain || (a.scratch && a.scratch_floats >= (size_t)a.M * a.K))) {
    SgemmX3Args g;
    g.A = plain ? a.X : a.scratch; g.sam = plain ? a.ldx : a.K; g.sak = 1;
    g.B = (const float*)a.W; g.sbk = 1; g.sbn = a.ldw ? a.ldw : a.K;
    g.C = a.Y; g.ldc = a.ldy; g.M = a.M; g.N = a.nout; g.K = a.K;
    g.bias = a.bias; g.act = a.act; g.res = a.res; g.ld_res = a.ld_res; g.res2 = a.res2; g.ld_res2 = a.ld_res2;
    if (sgemm_x3_supported(g)) {
      if (!plain) linear_prologue(a, a.scratch, s);
      const size_t used = plain ? 0 : ((size_t)a.M * a.K + 3) / 4 * 4;
      sgemm_x3_auto(g, a.scratch ? a.scratch + used : nullptr, a.scratch ? a.scratch_floats - used : 0, s);
      return;
    }
  }
  if (!plain || a.act == 2 || a.res2) {
    linear_rows(a, s);
    return;
  }
  GemmArgs g;
  g.A = a.X; g.n = 1; g.w = a.M; g.ow = a.M; g.cin = a.K; g.lda = a.ldx;
  g.W = a.W; g.w_dt = a.w_dt; g.w_stride_n = a.ldw ? a.ldw : a.K; g.cout = a.nout;
  g.bias = a.bias; g.act = a.act;
  g.out = a.Y; g.ldo = a.ldy;
  if (a.res && a.act == 0) {
    g.res = a.res; g.ld_res = a.ld_res;
    gemm_simt(g, s);
  } else {
    gemm_simt(g, s);
    // the residual is added AFTER the activation (graph.py:203-206): second pass
    if (a.res) add_rowvec(a.Y, F32, a.M, a.nout, a.res, a.ld_res, 1, s);
  }
}

void Gcn::create(const WeightMap& wm, const std::string& prefix, const echo_gcn_desc_t& d, DevPool& pool) {
  ECHO_CHECK(d.num_layers > 0 && d.hidden_dim % 4 == 0 && d.input_dim_obj % 4 == 0 && d.input_dim_pred % 4 == 0,
             "gcn: dims must be multiples of 4");
  cudaStream_t s = 0;
  max_nodes = d.max_nodes;
  max_triples = d.max_triples;
  dp = d.input_dim_pred;
  H = d.hidden_dim;
  const float eps = d.bn_eps > 0 ? d.bn_eps : 1e-5f;
  bn_eps = eps;
  max_d = d.input_dim_obj;
  layers.clear();
  for (int i = 0; i < d.num_layers; ++i) {
    GcnLayer L;
    L.din = d.input_dim_obj;
    L.dp = dp;
    L.H = H;
    const bool last = d.output_dim > 0 && i >= d.num_layers - 1;   // graph.py:239-243
    L.dout = last ? d.output_dim : d.input_dim_obj;
    max_d = std::max(max_d, L.dout);
    const std::string p = prefix + "gconvs." + std::to_string(i) + ".";
    const int k1 = 2 * L.din + dp;
    const bool bn = wm.has(p + "net1.1.running_mean");
    const int st = bn ? 3 : 2;   // build_mlp index layout (layers.py:21-38)
    Mat w1 = folded_linear(wm, p + "net1.0", p + "net1.1", H, k1, eps, pool, s);
    // split net1.0 by input block: [subject | predicate | object]
    float* wso = pool.alloc_n<float>((size_t)2 * H * L.din);
    copy_cols(w1.w, k1, H, L.din, wso, L.din, s);
    copy_cols(w1.w + L.din + dp, k1, H, L.din, wso + (size_t)H * L.din, L.din, s);
    float* wp = pool.alloc_n<float>((size_t)H * dp);
    copy_cols(w1.w + L.din, k1, H, dp, wp, dp, s);
    L.w_so.w = wso; L.w_so.nout = 2 * H; L.w_so.K = L.din;
    L.w_p.w = wp; L.w_p.nout = H; L.w_p.K = dp;
    L.b1 = w1.b;
    L.w2 = folded_linear(wm, p + "net1." + std::to_string(st), p + "net1." + std::to_string(st + 1), 2 * H + dp, H, eps, pool, s);
    L.w3 = folded_linear(wm, p + "net2.0", p + "net2.1", H, H, eps, pool, s);
    L.w4 = folded_linear(wm, p + "net2." + std::to_string(st), p + "net2." + std::to_string(st + 1), L.dout, H, eps, pool, s);
    if (bn && d.keep_train_weights) {   // batch-statistics mode: the unfolded Linears and the BatchNorm1d scale / shift
      auto vec = [&](const std::string& name, int n) {
        const WView& v = wm.get(name, {n});
        float* o = pool.alloc_n<float>(n);
        ECHO_CUDA(cudaMemcpyAsync(o, v.p, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
        return (const float*)o;
      };
      Mat r1 = plain_linear(wm, p + "net1.0", H, k1, pool, s);
      float* rso = pool.alloc_n<float>((size_t)2 * H * L.din);
      copy_cols(r1.w, k1, H, L.din, rso, L.din, s);
      copy_cols(r1.w + L.din + dp, k1, H, L.din, rso + (size_t)H * L.din, L.din, s);
      float* rp = pool.alloc_n<float>((size_t)H * dp);
      copy_cols(r1.w + L.din, k1, H, dp, rp, dp, s);
      L.t_so.w = rso; L.t_so.nout = 2 * H; L.t_so.K = L.din;
      L.t_p.w = rp; L.t_p.nout = H; L.t_p.K = dp;
      L.t_b1 = r1.b;
      L.t2 = plain_linear(wm, p + "net1.3", 2 * H + dp, H, pool, s);
      L.t3 = plain_linear(wm, p + "net2.0", H, H, pool, s);
      L.t4 = plain_linear(wm, p + "net2.3", L.dout, H, pool, s);
      const char* bns[4] = {"net1.1", "net1.4", "net2.1", "net2.4"};
      const int widths[4] = {H, 2 * H + dp, H, L.dout};
      for (int k = 0; k < 4; ++k) {
        L.bn_g[k] = vec(p + bns[k] + ".weight", widths[k]);
        L.bn_b[k] = vec(p + bns[k] + ".bias", widths[k]);
      }
      L.has_train = true;
    }
    L.residual = wm.has(p + "linear_projection.weight");
    if (L.residual) {
      L.proj = plain_linear(wm, p + "linear_projection", L.dout, L.din, pool, s);
      L.projp = plain_linear(wm, p + "linear_projection_pred", dp, dp, pool, s);
    }
    if (i + 1 < d.num_layers) ECHO_CHECK(L.dout == d.input_dim_obj, "gcn: inner layer width mismatch");
    layers.push_back(L);
  }
  const size_t N = max_nodes, T = max_triples;
  pso = pool.alloc_n<float>(N * 2 * H);
  pp = pool.alloc_n<float>(T * H);
  h1 = pool.alloc_n<float>(T * H);
  t2 = pool.alloc_n<float>(T * (2 * H + dp));
  pooled = pool.alloc_n<float>(N * H);
  n1 = pool.alloc_n<float>(N * H);
  proj = pool.alloc_n<float>(N * max_d);
  for (int i = 0; i < 2; ++i) {
    obj_pp[i] = pool.alloc_n<float>(N * max_d);
    pred_pp[i] = pool.alloc_n<float>(T * dp);
  }
  ECHO_CUDA(cudaStreamSynchronize(s));
}

void Gcn::forward(const echo_graph* g, const float* obj, const float* pred, float* obj_out, float* pred_out, cudaStream_t s,
                  bool batch_stats) {
  ECHO_CHECK(g, "gcn: null graph");
  const int N = g->n_nodes, T = g->n_triples;
  if (batch_stats)
    for (auto& L : layers)
      ECHO_CHECK(L.has_train, "gcn: batch-statistics forward needs a handle created with keep_train_weights (and BatchNorm1d MLPs)");
  // BatchNorm1d over the rows of the call + ReLU, in place (batch-statistics mode only)
  auto bn_relu = [&](float* x, int64_t ld, int rows, int C, const float* gm, const float* bt) { bn_rows_train(x, ld, rows, C, gm, bt, bn_eps, s); };
  ECHO_CHECK(N <= max_nodes && T <= max_triples, "gcn: graph (%d nodes, %d triples) exceeds handle capacity (%d, %d)", N, T,
             max_nodes, max_triples);
  const float* cur_obj = obj;
  const float* cur_pred = pred;
  for (size_t li = 0; li < layers.size(); ++li) {
    const GcnLayer& L = layers[li];
    const bool lastl = li + 1 == layers.size();
    float* nobj = lastl && obj_out ? obj_out : obj_pp[li & 1];
    float* npred = lastl && pred_out ? pred_out : pred_pp[li & 1];
    LinArgs a;
    // 1. node / predicate projections of net1.0
    const bool bs = batch_stats;
    a.X = cur_obj; a.ldx = L.din; a.M = N; a.K = L.din; a.nout = 2 * H; a.W = bs ? L.t_so.w : L.w_so.w; a.Y = pso; a.ldy = 2 * H;
    linear_auto(a, s);
    if (T > 0) {
      a = LinArgs();
      a.X = cur_pred; a.ldx = dp; a.M = T; a.K = dp; a.nout = H; a.W = bs ? L.t_p.w : L.w_p.w; a.Y = pp; a.ldy = H;
      linear_auto(a, s);
      // 2. warp-per-edge gather + combine (+ ReLU; batch statistics: the pre-activation, BatchNorm1d over the T rows, then ReLU)
      const size_t table = (size_t)N * 2 * H * sizeof(float);
      const float floor_ = bs ? -INFINITY : 0.f;
      const float* b1 = bs ? L.t_b1 : L.b1;
      if (table <= 48 * 1024) {   // a scene-sized graph: node features staged in shared memory, 16 edges per block
        edge_combine_staged_kernel<<<cdiv(T, 16), 512, table, s>>>(pso, pp, b1, g->s_idx, g->o_idx, T, N, H, h1, floor_);
      } else {
        edge_combine_kernel<<<cdiv((int64_t)T * 32, 256), 256, 0, s>>>(pso, pp, b1, g->s_idx, g->o_idx, T, H, h1, floor_);
      }
      ECHO_LAUNCH_CHECK();
      if (bs) bn_relu(h1, H, T, H, L.bn_g[0], L.bn_b[0]);
      // 3. second Linear of net1
      a = LinArgs();
      a.X = h1; a.ldx = H; a.M = T; a.K = H; a.nout = 2 * H + dp; a.W = bs ? L.t2.w : L.w2.w; a.bias = bs ? L.t2.b : L.w2.b;
      a.act = bs ? 0 : 1; a.Y = t2;
      a.ldy = 2 * H + dp;
      linear_auto(a, s);
      if (bs) bn_relu(t2, 2 * H + dp, T, 2 * H + dp, L.bn_g[1], L.bn_b[1]);
    }
    // 4. deterministic segmented mean
    node_pool_kernel<<<N, 64, 0, s>>>(t2, 2 * H + dp, H, H + dp, g->node_off, g->node_items, N, pooled);
    ECHO_LAUNCH_CHECK();
    // new_p = t2[:, H:H+dp] + linear_projection_pred(pred)
    if (T > 0) {
      if (L.residual) {
        a = LinArgs();
        a.X = cur_pred; a.ldx = dp; a.M = T; a.K = dp; a.nout = dp; a.W = L.projp.w; a.bias = L.projp.b;
        a.res = t2 + H; a.ld_res = 2 * H + dp; a.Y = npred; a.ldy = dp;
        linear_auto(a, s);
      } else {
        copy_cols(t2 + H, 2 * H + dp, T, dp, npred, dp, s);
      }
    }
    // 5. net2 + residual
    a = LinArgs();
    a.X = pooled; a.ldx = H; a.M = N; a.K = H; a.nout = H; a.W = bs ? L.t3.w : L.w3.w; a.bias = bs ? L.t3.b : L.w3.b; a.act = bs ? 0 : 1;
    a.Y = n1; a.ldy = H;
    linear_auto(a, s);
    if (bs) bn_relu(n1, H, N, H, L.bn_g[2], L.bn_b[2]);
    const float* resid = nullptr;
    if (L.residual) {
      a = LinArgs();
      a.X = cur_obj; a.ldx = L.din; a.M = N; a.K = L.din; a.nout = L.dout; a.W = L.proj.w; a.bias = L.proj.b; a.Y = proj;
      a.ldy = L.dout;
      linear_auto(a, s);
      resid = proj;
    }
    a = LinArgs();
    if (bs) {   // Linear -> BatchNorm1d(batch) -> ReLU, then the residual projection on top (graph.py:203-206)
      a.X = n1; a.ldx = H; a.M = N; a.K = H; a.nout = L.dout; a.W = L.t4.w; a.bias = L.t4.b; a.act = 0; a.Y = nobj; a.ldy = L.dout;
      linear_auto(a, s);
      bn_relu(nobj, L.dout, N, L.dout, L.bn_g[3], L.bn_b[3]);
      if (resid) add_rowvec(nobj, F32, N, L.dout, resid, L.dout, 1, s);
    } else {
      a.X = n1; a.ldx = H; a.M = N; a.K = H; a.nout = L.dout; a.W = L.w4.w; a.bias = L.w4.b; a.act = 1; a.res = resid;
      a.ld_res = L.dout; a.Y = nobj; a.ldy = L.dout;
      linear_auto(a, s);
    }
    cur_obj = nobj;
    cur_pred = npred;
  }
}

}  // namespace echo

