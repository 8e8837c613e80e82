// GraphTripleConvNet TRAINING executor: forward on batch statistics that keeps what the backward pass needs, and the backward pass
// (the first backward of DESIGN section 7's plan: `linear_rows` / GCN).  What the reference gets from autograd over
// model/graph.py:124-211 + model/layers.py:21-38 under model.train() is written out here per layer:
//
//   forward   Tin = [obj[s] | pred | obj[o]]                       (T x (2 din + dp))      gather_tin_kernel
//             z1 = Tin W1^T + b1,  a1 = ReLU(BN(z1))               (T x H)                 linear_auto + bn_relu_save_kernel
//             z2 = a1 W2^T + b2,   a2 = ReLU(BN(z2))               (T x (2H + dp))
//             pooled = segmented mean of a2[:, :H] / a2[:, H+dp:]  (N x H)                 node_pool_train_kernel
//             z3 = pooled W3^T + b3, a3 = ReLU(BN(z3)); z4 = a3 W4^T + b4, a4 = ReLU(BN(z4))
//             obj' = a4 + proj(obj),  pred' = a2[:, H:H+dp] + proj_p(pred)
//   backward  the same chain in reverse: BatchNorm1d(batch) + ReLU backward (bn_relu_bwd_kernel), dX = dY W (linear_dgrad_kernel),
//             dW += dY^T X, db += colsum(dY) (linear_wgrad_kernel, colsum_acc_kernel), the pooling's and the gather's adjoints as
//             per-edge / per-node kernels over the graph's CSR (deterministic: fixed summation order, no float atomics).
//
// The handle reads the caller's parameter tensors IN PLACE (no copies: an optimizer step between two iterations needs no rebuild) and
// ACCUMULATES into the caller's gradient tensors (+=, as autograd does into .grad).  BatchNorm1d's running statistics and
// num_batches_tracked are updated by the forward as torch does (momentum 0.1, unbiased variance).  Sizes here are tens to thousands
// of rows against weights of a few MB: every kernel is HBM / L2 bound on the weights and gradients, fp32 SIMT.
#include "model.cuh"

#include <algorithm>
#include <string>
#include <vector>

namespace echo {
namespace {

struct TLin {   // one nn.Linear: parameters read in place, gradients accumulated in place
  const float* w = nullptr;
  const float* b = nullptr;
  float* gw = nullptr;
  float* gb = nullptr;
  int nout = 0, K = 0;
};
struct TBn {    // one nn.BatchNorm1d
  const float* g = nullptr;
  const float* b = nullptr;
  float* gg = nullptr;
  float* gb = nullptr;
  float* rmean = nullptr;     // running statistics, updated by the forward (may be null)
  float* rvar = nullptr;
  long long* nbt = nullptr;   // num_batches_tracked (may be null)
  float* mean = nullptr;      // batch statistics of the last forward (handle-owned)
  float* rstd = nullptr;
  int C = 0;
};
struct TLayer {
  int din = 0, dout = 0;
  bool residual = false;
  TLin l1, l2, l3, l4, proj, projp;
  TBn bn[4];
  // saved by the forward
  float *x_obj = nullptr, *x_pred = nullptr, *tin = nullptr, *z1 = nullptr, *a1 = nullptr, *z2 = nullptr, *a2 = nullptr;
  float *pooled = nullptr, *z3 = nullptr, *a3 = nullptr, *z4 = nullptr;
};

__global__ void gather_tin_kernel(const float* __restrict__ obj, const float* __restrict__ pred, const int* __restrict__ s_idx,
                                  const int* __restrict__ o_idx, int T, int din, int dp, float* __restrict__ tin) {
  const int t = blockIdx.x;
  if (t >= T) return;
  const int k1 = 2 * din + dp;
  const float4* a = reinterpret_cast<const float4*>(obj + (int64_t)s_idx[t] * din);
  const float4* b = reinterpret_cast<const float4*>(pred + (int64_t)t * dp);
  const float4* c = reinterpret_cast<const float4*>(obj + (int64_t)o_idx[t] * din);
  float4* out = reinterpret_cast<float4*>(tin + (int64_t)t * k1);
  const int n1 = din / 4, n2 = dp / 4;
  for (int q = threadIdx.x; q < 2 * n1 + n2; q += blockDim.x)
    out[q] = q < n1 ? __ldg(a + q) : (q < n1 + n2 ? __ldg(b + (q - n1)) : __ldg(c + (q - n1 - n2)));
}

// BatchNorm1d on the statistics of the batch + ReLU, out of place; the arithmetic of bn_rows_train_kernel (gcn.cu: two passes, biased
// variance) over more row lanes.  Also: the batch statistics kept for the backward pass, and torch's running-statistics update
// (momentum, UNBIASED variance; torch/nn/modules/batchnorm.py).
constexpr int BN_LANES = 32;   // row lanes per block of 32 columns (1024 threads): batches of thousands of triples stay short

__global__ void __launch_bounds__(32 * BN_LANES) bn_relu_save_kernel(const float* __restrict__ z, int64_t ldz, int rows, int C,
                                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                      float eps, float* __restrict__ a, int64_t lda,
                                                                      float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                      float* __restrict__ rmean, float* __restrict__ rvar, float momentum) {
  __shared__ float red[BN_LANES][33];
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = threadIdx.y;
  const bool in = c < C;
  float s = 0.f;
  if (in) for (int r = r0; r < rows; r += BN_LANES) s += z[(int64_t)r * ldz + c];
  red[r0][threadIdx.x] = s;
  __syncthreads();
  float mean = 0.f;
  for (int k = 0; k < BN_LANES; ++k) mean += red[k][threadIdx.x];
  mean /= (float)rows;
  __syncthreads();
  float q = 0.f;
  if (in) for (int r = r0; r < rows; r += BN_LANES) { const float d = z[(int64_t)r * ldz + c] - mean; q = fmaf(d, d, q); }
  red[r0][threadIdx.x] = q;
  __syncthreads();
  float var = 0.f;
  for (int k = 0; k < BN_LANES; ++k) var += red[k][threadIdx.x];
  const float ssq = var;
  var /= (float)rows;
  if (!in) return;
  const float rstd = rsqrtf(var + eps);
  const float sc = rstd * gamma[c], sh = beta[c];
  for (int r = r0; r < rows; r += BN_LANES) a[(int64_t)r * lda + c] = fmaxf((z[(int64_t)r * ldz + c] - mean) * sc + sh, 0.f);
  if (r0 == 0) {
    mean_out[c] = mean;
    rstd_out[c] = rstd;
    if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
    if (rvar) rvar[c] = (1.f - momentum) * rvar[c] + momentum * (ssq / (float)(rows - 1));
  }
}

__global__ void bump_counter_kernel(long long* a, long long* b, long long* c, long long* d) {
  if (a) *a += 1;
  if (b) *b += 1;
  if (c) *c += 1;
  if (d) *d += 1;
}

// Backward of a = ReLU(BN_batch(z)): with xhat = (z - mean) rstd, g = da where the forward's output was positive,
//   dbeta = sum g,  dgamma = sum g xhat,  dz = gamma rstd (g - dbeta / R - xhat dgamma / R).
// dz may alias da.  Block = 32 columns x 32 row lanes, fixed summation order.
__global__ void __launch_bounds__(32 * BN_LANES) bn_relu_bwd_kernel(const float* da, int64_t ldd, const float* __restrict__ z, int64_t ldz,
                                                                     int rows, int C, const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, const float* __restrict__ mean_,
                                                                     const float* __restrict__ rstd_, float* dz, int64_t ldo,
                                                                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[2][BN_LANES][33];
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = threadIdx.y;
  const bool in = c < C;
  float mean = 0.f, rstd = 0.f, sc = 0.f, sh = 0.f;
  if (in) { mean = mean_[c]; rstd = rstd_[c]; sc = rstd * gamma[c]; sh = beta[c]; }
  float s1 = 0.f, s2 = 0.f;
  if (in)
    for (int r = r0; r < rows; r += BN_LANES) {
      const float zc = z[(int64_t)r * ldz + c] - mean;
      const float g = (zc * sc + sh) > 0.f ? da[(int64_t)r * ldd + c] : 0.f;
      s1 += g;
      s2 = fmaf(g, zc * rstd, s2);
    }
  red[0][r0][threadIdx.x] = s1;
  red[1][r0][threadIdx.x] = s2;
  __syncthreads();
  float db = 0.f, dg = 0.f;
  for (int k = 0; k < BN_LANES; ++k) { db += red[0][k][threadIdx.x]; dg += red[1][k][threadIdx.x]; }
  if (!in) return;
  const float inv = 1.f / (float)rows, gr = gamma[c] * rstd;
  for (int r = r0; r < rows; r += BN_LANES) {
    const float zc = z[(int64_t)r * ldz + c] - mean;
    const float g = (zc * sc + sh) > 0.f ? da[(int64_t)r * ldd + c] : 0.f;
    dz[(int64_t)r * ldo + c] = gr * (g - db * inv - (zc * rstd) * (dg * inv));
  }
  if (r0 == 0) {
    if (dgamma) dgamma[c] += dg;
    if (dbeta) dbeta[c] += db;
  }
}

// ---- the same two operations over MANY rows (collated batches: thousands of triples), in two stages so that the grid covers the GPU:
// row blocks of BN_RB rows reduce to partials (Chan's pairwise form for the variance: per-block mean and centred sum of squares),
// then every block merges the partials of its columns in block order -- deterministic -- and applies the result to its own rows.
constexpr int BN_RB = 256;

__global__ void __launch_bounds__(1024) bn_stats_partial_kernel(const float* __restrict__ z, int64_t ldz, int rows, int C,
                                                                float* __restrict__ pmean, float* __restrict__ pm2) {
  __shared__ float red[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = threadIdx.y, rb = blockIdx.y;
  const int rbeg = rb * BN_RB, rend = min(rows, rbeg + BN_RB), cnt = rend - rbeg;
  const bool in = c < C;
  float s = 0.f;
  if (in) for (int r = rbeg + r0; r < rend; r += 32) s += z[(int64_t)r * ldz + c];
  red[r0][threadIdx.x] = s;
  __syncthreads();
  float mean = 0.f;
  for (int k = 0; k < 32; ++k) mean += red[k][threadIdx.x];
  mean /= (float)cnt;
  __syncthreads();
  float q = 0.f;
  if (in) for (int r = rbeg + r0; r < rend; r += 32) { const float d = z[(int64_t)r * ldz + c] - mean; q = fmaf(d, d, q); }
  red[r0][threadIdx.x] = q;
  __syncthreads();
  if (r0 == 0 && in) {
    float m2 = 0.f;
    for (int k = 0; k < 32; ++k) m2 += red[k][threadIdx.x];
    pmean[(int64_t)rb * C + c] = mean;
    pm2[(int64_t)rb * C + c] = m2;
  }
}

__global__ void __launch_bounds__(1024) bn_relu_apply_kernel(const float* __restrict__ z, int64_t ldz, int rows, int C,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                             const float* __restrict__ pmean, const float* __restrict__ pm2, int nrb,
                                                             float* __restrict__ a, int64_t lda, float* __restrict__ mean_out,
                                                             float* __restrict__ rstd_out, float* __restrict__ rmean,
                                                             float* __restrict__ rvar, float momentum) {
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = threadIdx.y, rb = blockIdx.y;
  if (c >= C) return;
  float mean = 0.f;
  for (int b = 0; b < nrb; ++b) mean += (float)(min(rows, (b + 1) * BN_RB) - b * BN_RB) * pmean[(int64_t)b * C + c];
  mean /= (float)rows;
  float m2 = 0.f;
  for (int b = 0; b < nrb; ++b) {
    const float d = pmean[(int64_t)b * C + c] - mean;
    m2 += pm2[(int64_t)b * C + c] + (float)(min(rows, (b + 1) * BN_RB) - b * BN_RB) * d * d;
  }
  const float rstd = rsqrtf(m2 / (float)rows + eps);
  const float sc = rstd * gamma[c], sh = beta[c];
  const int rbeg = rb * BN_RB, rend = min(rows, rbeg + BN_RB);
  for (int r = rbeg + r0; r < rend; r += 32) a[(int64_t)r * lda + c] = fmaxf((z[(int64_t)r * ldz + c] - mean) * sc + sh, 0.f);
  if (rb == 0 && r0 == 0) {
    mean_out[c] = mean;
    rstd_out[c] = rstd;
    if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
    if (rvar) rvar[c] = (1.f - momentum) * rvar[c] + momentum * (m2 / (float)(rows - 1));
  }
}

__global__ void __launch_bounds__(1024) bn_bwd_partial_kernel(const float* __restrict__ da, int64_t ldd, const float* __restrict__ z,
                                                              int64_t ldz, int rows, int C, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ mean_,
                                                              const float* __restrict__ rstd_, float* __restrict__ p1,
                                                              float* __restrict__ p2) {
  __shared__ float red[2][32][33];
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = threadIdx.y, rb = blockIdx.y;
  const int rbeg = rb * BN_RB, rend = min(rows, rbeg + BN_RB);
  const bool in = c < C;
  float s1 = 0.f, s2 = 0.f;
  if (in) {
    const float mean = mean_[c], rstd = rstd_[c], sc = rstd * gamma[c], sh = beta[c];
    for (int r = rbeg + r0; r < rend; r += 32) {
      const float zc = z[(int64_t)r * ldz + c] - mean;
      const float g = (zc * sc + sh) > 0.f ? da[(int64_t)r * ldd + c] : 0.f;
      s1 += g;
      s2 = fmaf(g, zc * rstd, s2);
    }
  }
  red[0][r0][threadIdx.x] = s1;
  red[1][r0][threadIdx.x] = s2;
  __syncthreads();
  if (r0 == 0 && in) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < 32; ++k) { a += red[0][k][threadIdx.x]; b += red[1][k][threadIdx.x]; }
    p1[(int64_t)rb * C + c] = a;
    p2[(int64_t)rb * C + c] = b;
  }
}

__global__ void __launch_bounds__(1024) bn_bwd_apply_kernel(const float* da, int64_t ldd, const float* __restrict__ z, int64_t ldz, int rows,
                                                            int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ mean_, const float* __restrict__ rstd_,
                                                            const float* __restrict__ p1, const float* __restrict__ p2, int nrb, float* dz,
                                                            int64_t ldo, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * 32 + threadIdx.x, r0 = threadIdx.y, rb = blockIdx.y;
  if (c >= C) return;
  float db = 0.f, dg = 0.f;
  for (int b = 0; b < nrb; ++b) { db += p1[(int64_t)b * C + c]; dg += p2[(int64_t)b * C + c]; }
  const float mean = mean_[c], rstd = rstd_[c], sc = rstd * gamma[c], sh = beta[c];
  const float inv = 1.f / (float)rows, gr = gamma[c] * rstd;
  const int rbeg = rb * BN_RB, rend = min(rows, rbeg + BN_RB);
  for (int r = rbeg + r0; r < rend; r += 32) {
    const float zc = z[(int64_t)r * ldz + c] - mean;
    const float g = (zc * sc + sh) > 0.f ? da[(int64_t)r * ldd + c] : 0.f;
    dz[(int64_t)r * ldo + c] = gr * (g - db * inv - (zc * rstd) * (dg * inv));
  }
  if (rb == 0 && r0 == 0) {
    if (dgamma) dgamma[c] += dg;
    if (dbeta) dbeta[c] += db;
  }
}

// dX[m, k] (+)= sum_n dY[m, n] W[n, k]        W [N, K] row-major as nn.Linear stores it: coalesced over k, no transposed copy.
// Block: 16 rows x 128 columns of dX; dY tile staged in shared memory.
constexpr int DG_TM = 16, DG_TK = 128, DG_TN = 64;
__global__ void __launch_bounds__(DG_TK) linear_dgrad_kernel(const float* __restrict__ dY, int64_t ldy, const float* __restrict__ W,
                                                             int M, int N, int K, float* __restrict__ dX, int64_t ldx, int accumulate) {
  __shared__ float ys[DG_TM][DG_TN];
  const int k = blockIdx.x * DG_TK + threadIdx.x, m0 = blockIdx.y * DG_TM;
  float acc[DG_TM];
#pragma unroll
  for (int i = 0; i < DG_TM; ++i) acc[i] = 0.f;
  for (int n0 = 0; n0 < N; n0 += DG_TN) {
    for (int i = threadIdx.x; i < DG_TM * DG_TN; i += DG_TK) {
      const int mi = i / DG_TN, ni = i % DG_TN;
      ys[mi][ni] = (m0 + mi < M && n0 + ni < N) ? dY[(int64_t)(m0 + mi) * ldy + n0 + ni] : 0.f;
    }
    __syncthreads();
    if (k < K) {
      const int nend = min(DG_TN, N - n0);
      for (int ni = 0; ni < nend; ++ni) {
        const float w = __ldg(W + (int64_t)(n0 + ni) * K + k);
#pragma unroll
        for (int i = 0; i < DG_TM; ++i) acc[i] = fmaf(ys[i][ni], w, acc[i]);
      }
    }
    __syncthreads();
  }
  if (k >= K) return;
#pragma unroll
  for (int i = 0; i < DG_TM; ++i)
    if (m0 + i < M) {
      float* o = dX + (int64_t)(m0 + i) * ldx + k;
      *o = accumulate ? *o + acc[i] : acc[i];
    }
}

// dW[n, k] += sum_m dY[m, n] X[m, k]          block = 32 x 32 tile of dW, 256 threads x 4 outputs, rows in ascending order.
__global__ void __launch_bounds__(256) linear_wgrad_kernel(const float* __restrict__ dY, int64_t ldy, const float* __restrict__ X,
                                                           int64_t ldx, int M, int N, int K, float* __restrict__ dW) {
  __shared__ float ys[32][33], xs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // tx: k within the tile; ty: 4 n's each
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int m0 = 0; m0 < M; m0 += 32) {
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
      const int mi = i >> 5, ci = i & 31;
      ys[mi][ci] = (m0 + mi < M && n0 + ci < N) ? dY[(int64_t)(m0 + mi) * ldy + n0 + ci] : 0.f;
      xs[mi][ci] = (m0 + mi < M && k0 + ci < K) ? X[(int64_t)(m0 + mi) * ldx + k0 + ci] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int mi = 0; mi < 32; ++mi) {
      const float x = xs[mi][tx];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(ys[mi][ty * 4 + j], x, acc[j]);
    }
    __syncthreads();
  }
  if (k0 + tx < K)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + ty * 4 + j < N) dW[(int64_t)(n0 + ty * 4 + j) * K + k0 + tx] += acc[j];
}

__global__ void __launch_bounds__(1024) colsum_acc_kernel(const float* __restrict__ dY, int64_t ldy, int M, int N, float* __restrict__ db) {
  __shared__ float red[32][33];
  const int n = blockIdx.x * 32 + threadIdx.x, r0 = threadIdx.y;
  float s = 0.f;
  if (n < N) for (int m = r0; m < M; m += 32) s += dY[(int64_t)m * ldy + n];
  red[r0][threadIdx.x] = s;
  __syncthreads();
  if (r0 == 0 && n < N) {
    float t = 0.f;
    for (int k = 0; k < 32; ++k) t += red[k][threadIdx.x];
    db[n] += t;
  }
}

// dW += sum over the row chunks of a split wgrad, in chunk order (deterministic)
__global__ void splitk_acc_kernel(const float* __restrict__ ws, int64_t n, int splits, float* __restrict__ dW) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = 0.f;
  for (int b = 0; b < splits; ++b) t += ws[(int64_t)b * n + i];
  dW[i] += t;
}

// the segmented mean of the forward (graph.py:161-199), as node_pool_kernel in gcn.cu
__global__ void node_pool_train_kernel(const float* __restrict__ a2, int ld, int H, int off_o, const int* __restrict__ node_off,
                                       const int* __restrict__ node_items, int N, float* __restrict__ pooled) {
  const int n = blockIdx.x;
  if (n >= N) return;
  const int beg = node_off[n], end = node_off[n + 1];
  const float cnt = fmaxf((float)(end - beg), 1.f);
  for (int q = threadIdx.x; q < H / 4; q += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = beg; i < end; ++i) {
      const int item = node_items[i], t = item >> 1, role = item & 1;
      const float4 v = __ldg(reinterpret_cast<const float4*>(a2 + (int64_t)t * ld + (role ? off_o : 0)) + q);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x /= cnt; acc.y /= cnt; acc.z /= cnt; acc.w /= cnt;
    reinterpret_cast<float4*>(pooled + (int64_t)n * H)[q] = acc;
  }
}

// adjoint of [new_s | new_p | new_o] -> (pooled, pred'): da2[t] = [dpooled[s_t] / cnt(s_t) | dPred'[t] | dpooled[o_t] / cnt(o_t)]
__global__ void pool_bwd_kernel(const float* __restrict__ dpooled, const float* __restrict__ dpred, const int* __restrict__ s_idx,
                                const int* __restrict__ o_idx, const int* __restrict__ node_off, int T, int H, int dp,
                                float* __restrict__ da2) {
  const int t = blockIdx.x;
  if (t >= T) return;
  const int s = s_idx[t], o = o_idx[t], W = 2 * H + dp;
  const float cs = fmaxf((float)(node_off[s + 1] - node_off[s]), 1.f), co = fmaxf((float)(node_off[o + 1] - node_off[o]), 1.f);
  float* out = da2 + (int64_t)t * W;
  for (int q = threadIdx.x; q < W; q += blockDim.x) {
    float v;
    if (q < H) v = dpooled[(int64_t)s * H + q] / cs;
    else if (q < H + dp) v = dpred[(int64_t)t * dp + (q - H)];
    else v = dpooled[(int64_t)o * H + (q - H - dp)] / co;
    out[q] = v;
  }
}

// adjoint of the gather: dObj[n] += sum over the node's CSR items of dTin[t, subject or object block]   (fixed item order)
__global__ void gather_bwd_obj_kernel(const float* __restrict__ dtin, int k1, int din, int dp, const int* __restrict__ node_off,
                                      const int* __restrict__ node_items, int N, float* __restrict__ dobj) {
  const int n = blockIdx.x;
  if (n >= N) return;
  const int beg = node_off[n], end = node_off[n + 1];
  for (int q = threadIdx.x; q < din; q += blockDim.x) {
    float acc = 0.f;
    for (int i = beg; i < end; ++i) {
      const int item = node_items[i], t = item >> 1, role = item & 1;
      acc += dtin[(int64_t)t * k1 + (role ? din + dp : 0) + q];
    }
    dobj[(int64_t)n * din + q] += acc;
  }
}

__global__ void add_cols_kernel(const float* __restrict__ src, int64_t lds, int rows, int cols, float* __restrict__ dst, int64_t ldd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  dst[(int64_t)r * ldd + c] += src[(int64_t)r * lds + c];
}

struct Table {   // name -> (pointer, dtype, numel) of a caller-owned tensor table
  struct E { void* p; int dtype; int64_t numel; };
  std::unordered_map<std::string, E> m;
  void load(const echo_weight_t* w, int n) {
    for (int i = 0; i < n; ++i) {
      ECHO_CHECK(w[i].name && w[i].ndim >= 0 && w[i].ndim <= 6, "gcn_train: bad table entry %d", i);
      int64_t ne = 1;
      for (int k = 0; k < w[i].ndim; ++k) ne *= w[i].shape[k];
      m[w[i].name] = E{const_cast<void*>(w[i].data), w[i].dtype, ne};
    }
  }
  float* f32(const std::string& k, int64_t numel, bool required) const {
    auto it = m.find(k);
    if (it == m.end() || !it->second.p) {
      ECHO_CHECK(!required, "gcn_train: missing tensor '%s'", k.c_str());
      return nullptr;
    }
    ECHO_CHECK(it->second.dtype == 0 && it->second.numel == numel, "gcn_train: '%s' must be float32 with %lld elements", k.c_str(),
               (long long)numel);
    return (float*)it->second.p;
  }
  long long* i64(const std::string& k) const {
    auto it = m.find(k);
    if (it == m.end() || !it->second.p || it->second.dtype != 1) return nullptr;
    return (long long*)it->second.p;
  }
};

}  // namespace
}  // namespace echo

struct echo_gcn_train {
  echo::DevPool pool;
  std::vector<echo::TLayer> layers;
  int max_nodes = 0, max_triples = 0, dp = 0, H = 0, max_d = 0, max_k1 = 0;
  float eps = 1e-5f, momentum = 0.1f;
  float *a4 = nullptr, *dbuf_t = nullptr, *dtin = nullptr, *dbuf_n = nullptr, *dpooled = nullptr;
  float* bn_part = nullptr;   // per-row-block partials of the two-stage BatchNorm kernels
  float* ws = nullptr;        // partial weight gradients of a split reduction
  size_t ws_floats = 0;
  float *dobj_pp[2] = {nullptr, nullptr}, *dpred_pp[2] = {nullptr, nullptr};
  int saved_nodes = -1, saved_triples = -1;   // what the last forward saved (the backward must be called on the same graph)
  uint64_t saved_graph = 0;
};

namespace echo {

echo_gcn_train* gcn_train_create(const echo_gcn_desc_t* desc, const echo_weight_t* params, int n_params, const echo_weight_t* grads,
                                 int n_grads) {
  ECHO_CHECK(desc, "gcn_train_create: null descriptor");
  const echo_gcn_desc_t& d = *desc;
  ECHO_CHECK(d.num_layers > 0 && d.hidden_dim % 4 == 0 && d.input_dim_obj % 4 == 0 && d.input_dim_pred % 4 == 0 &&
                 (d.output_dim <= 0 || d.output_dim % 4 == 0), "gcn_train: dims must be multiples of 4");
  ECHO_CHECK(d.max_nodes > 0 && d.max_triples > 0, "gcn_train: max_nodes / max_triples must be positive");
  Table P, G;
  P.load(params, n_params);
  G.load(grads, n_grads);
  auto h = new echo_gcn_train();
  try {
    h->max_nodes = d.max_nodes; h->max_triples = d.max_triples; h->dp = d.input_dim_pred; h->H = d.hidden_dim;
    h->eps = d.bn_eps > 0 ? d.bn_eps : 1e-5f;
    const int H = h->H, dp = h->dp;
    const size_t N = d.max_nodes, T = d.max_triples;
    h->max_d = d.input_dim_obj;
    for (int i = 0; i < d.num_layers; ++i) {
      TLayer L;
      L.din = d.input_dim_obj;
      const bool last = d.output_dim > 0 && i >= d.num_layers - 1;   // graph.py:239-243
      L.dout = last ? d.output_dim : d.input_dim_obj;
      if (i + 1 < d.num_layers) ECHO_CHECK(L.dout == d.input_dim_obj, "gcn_train: inner layer width mismatch");
      h->max_d = std::max(h->max_d, L.dout);
      const std::string p = "gconvs." + std::to_string(i) + ".";
      ECHO_CHECK(P.m.count(p + "net1.1.weight"), "gcn_train: the training executor needs BatchNorm1d MLPs (mlp_normalization='batch')");
      const int k1 = 2 * L.din + dp;
      h->max_k1 = std::max(h->max_k1, k1);
      auto lin = [&](const std::string& name, int nout, int K) {
        TLin l;
        l.nout = nout; l.K = K;
        l.w = P.f32(name + ".weight", (int64_t)nout * K, true);
        l.b = P.f32(name + ".bias", nout, true);
        l.gw = G.f32(name + ".weight", (int64_t)nout * K, true);
        l.gb = G.f32(name + ".bias", nout, true);
        return l;
      };
      auto bn = [&](const std::string& name, int C) {
        TBn b;
        b.C = C;
        b.g = P.f32(name + ".weight", C, true);
        b.b = P.f32(name + ".bias", C, true);
        b.gg = G.f32(name + ".weight", C, true);
        b.gb = G.f32(name + ".bias", C, true);
        b.rmean = P.f32(name + ".running_mean", C, false);
        b.rvar = P.f32(name + ".running_var", C, false);
        b.nbt = P.i64(name + ".num_batches_tracked");
        b.mean = h->pool.alloc_n<float>(C);
        b.rstd = h->pool.alloc_n<float>(C);
        return b;
      };
      L.l1 = lin(p + "net1.0", H, k1);
      L.bn[0] = bn(p + "net1.1", H);
      L.l2 = lin(p + "net1.3", 2 * H + dp, H);
      L.bn[1] = bn(p + "net1.4", 2 * H + dp);
      L.l3 = lin(p + "net2.0", H, H);
      L.bn[2] = bn(p + "net2.1", H);
      L.l4 = lin(p + "net2.3", L.dout, H);
      L.bn[3] = bn(p + "net2.4", L.dout);
      L.residual = P.m.count(p + "linear_projection.weight") != 0;
      if (L.residual) {
        L.proj = lin(p + "linear_projection", L.dout, L.din);
        L.projp = lin(p + "linear_projection_pred", dp, dp);
      }
      L.x_obj = h->pool.alloc_n<float>(N * L.din);
      L.x_pred = h->pool.alloc_n<float>(T * dp);
      L.tin = h->pool.alloc_n<float>(T * k1);
      L.z1 = h->pool.alloc_n<float>(T * H);
      L.a1 = h->pool.alloc_n<float>(T * H);
      L.z2 = h->pool.alloc_n<float>(T * (2 * H + dp));
      L.a2 = h->pool.alloc_n<float>(T * (2 * H + dp));
      L.pooled = h->pool.alloc_n<float>(N * H);
      L.z3 = h->pool.alloc_n<float>(N * H);
      L.a3 = h->pool.alloc_n<float>(N * H);
      L.z4 = h->pool.alloc_n<float>(N * L.dout);
      h->layers.push_back(L);
    }
    h->a4 = h->pool.alloc_n<float>(N * h->max_d);
    h->dbuf_t = h->pool.alloc_n<float>(T * (2 * H + dp));
    h->dtin = h->pool.alloc_n<float>(T * std::max(h->max_k1, H));
    h->dbuf_n = h->pool.alloc_n<float>(N * std::max(h->max_d, H));
    h->dpooled = h->pool.alloc_n<float>(N * H);
    {
      const size_t maxc = std::max<size_t>(2 * H + dp, h->max_d);
      h->bn_part = h->pool.alloc_n<float>(2 * (size_t)cdiv(std::max(N, T), BN_RB) * maxc + 2 * maxc);
    }
    if (std::max(N, T) > 64) {   // collated batches: room for up to 16 row chunks of the largest weight gradient
      size_t wmax = 0;
      for (auto& L : h->layers)
        for (const TLin* l : {&L.l1, &L.l2, &L.l3, &L.l4, &L.proj, &L.projp}) wmax = std::max(wmax, (size_t)l->nout * l->K);
      h->ws_floats = 16 * wmax;
      h->ws = h->pool.alloc_n<float>(h->ws_floats);
    }
    for (int i = 0; i < 2; ++i) {
      h->dobj_pp[i] = h->pool.alloc_n<float>(N * h->max_d);
      h->dpred_pp[i] = h->pool.alloc_n<float>(T * dp);
    }
  } catch (...) {
    h->pool.destroy();
    delete h;
    throw;
  }
  return h;
}

void gcn_train_destroy(echo_gcn_train* h) {
  if (!h) return;
  h->pool.destroy();
  delete h;
}

// Y = X W^T + b (+ res): up to 64 rows the few-row kernel of the sampling path (linear.cu, HBM-bound on W); beyond, the 3 x TF32
// tensor-core GEMM (sgemm_x3.cu).
static void lin_fwd(const float* X, int64_t ldx, int M, const TLin& l, float* Y, int64_t ldy, const float* res, int64_t ld_res, float* ws,
                    size_t ws_floats, cudaStream_t s) {
  if (M == 0) return;
  if (M > 64) {
    SgemmX3Args g;
    g.A = X; g.sam = ldx; g.sak = 1; g.B = l.w; g.sbk = 1; g.sbn = l.K; g.C = Y; g.ldc = ldy; g.M = M; g.N = l.nout; g.K = l.K;
    g.bias = l.b; g.res = res; g.ld_res = ld_res;
    if (sgemm_x3_supported(g)) {
      sgemm_x3_auto(g, ws, ws_floats, s);
      return;
    }
  }
  LinArgs a;
  a.X = X; a.ldx = ldx; a.M = M; a.K = l.K; a.nout = l.nout; a.W = l.w; a.bias = l.b; a.Y = Y; a.ldy = ldy; a.res = res; a.ld_res = ld_res;
  linear_auto(a, s);
}

static void bn_fwd(const float* z, int rows, TBn& b, float* a, float eps, float momentum, float* part, cudaStream_t s) {
  ECHO_CHECK(rows > 1, "gcn_train: BatchNorm1d on batch statistics needs more than one row (torch raises here too), got %d", rows);
  if (rows <= 2 * BN_RB) {
    bn_relu_save_kernel<<<cdiv(b.C, 32), dim3(32, BN_LANES), 0, s>>>(z, b.C, rows, b.C, b.g, b.b, eps, a, b.C, b.mean, b.rstd, b.rmean,
                                                                      b.rvar, momentum);
  } else {
    const int nrb = cdiv(rows, BN_RB);
    float *pmean = part, *pm2 = part + (size_t)nrb * b.C;
    bn_stats_partial_kernel<<<dim3(cdiv(b.C, 32), nrb), dim3(32, 32), 0, s>>>(z, b.C, rows, b.C, pmean, pm2);
    ECHO_LAUNCH_CHECK();
    bn_relu_apply_kernel<<<dim3(cdiv(b.C, 32), nrb), dim3(32, 32), 0, s>>>(z, b.C, rows, b.C, b.g, b.b, eps, pmean, pm2, nrb, a, b.C, b.mean,
                                                                            b.rstd, b.rmean, b.rvar, momentum);
  }
  ECHO_LAUNCH_CHECK();
}

void gcn_train_forward(echo_gcn_train* h, const echo_graph* g, const float* obj, const float* pred, float* obj_out, float* pred_out,
                       cudaStream_t s) {
  ECHO_CHECK(h && g && obj && obj_out, "gcn_train_forward: null argument");
  const int N = g->n_nodes, T = g->n_triples, H = h->H, dp = h->dp, W2 = 2 * H + dp;
  ECHO_CHECK(N <= h->max_nodes && T <= h->max_triples, "gcn_train: graph (%d nodes, %d triples) exceeds handle capacity (%d, %d)", N, T,
             h->max_nodes, h->max_triples);
  ECHO_CHECK(T > 0 && pred && pred_out, "gcn_train_forward: a graph without triples has no net1 batch to normalise");
  const float* cur_obj = obj;
  const float* cur_pred = pred;
  for (size_t li = 0; li < h->layers.size(); ++li) {
    TLayer& L = h->layers[li];
    const bool lastl = li + 1 == h->layers.size();
    const int k1 = 2 * L.din + dp;
    if (cur_obj != L.x_obj) ECHO_CUDA(cudaMemcpyAsync(L.x_obj, cur_obj, sizeof(float) * N * L.din, cudaMemcpyDeviceToDevice, s));
    if (cur_pred != L.x_pred) ECHO_CUDA(cudaMemcpyAsync(L.x_pred, cur_pred, sizeof(float) * T * dp, cudaMemcpyDeviceToDevice, s));
    gather_tin_kernel<<<T, 128, 0, s>>>(L.x_obj, L.x_pred, g->s_idx, g->o_idx, T, L.din, dp, L.tin);
    ECHO_LAUNCH_CHECK();
    lin_fwd(L.tin, k1, T, L.l1, L.z1, H, nullptr, 0, h->ws, h->ws_floats, s);
    bn_fwd(L.z1, T, L.bn[0], L.a1, h->eps, h->momentum, h->bn_part, s);
    lin_fwd(L.a1, H, T, L.l2, L.z2, W2, nullptr, 0, h->ws, h->ws_floats, s);
    bn_fwd(L.z2, T, L.bn[1], L.a2, h->eps, h->momentum, h->bn_part, s);
    node_pool_train_kernel<<<N, 64, 0, s>>>(L.a2, W2, H, H + dp, g->node_off, g->node_items, N, L.pooled);
    ECHO_LAUNCH_CHECK();
    lin_fwd(L.pooled, H, N, L.l3, L.z3, H, nullptr, 0, h->ws, h->ws_floats, s);
    bn_fwd(L.z3, N, L.bn[2], L.a3, h->eps, h->momentum, h->bn_part, s);
    lin_fwd(L.a3, H, N, L.l4, L.z4, L.dout, nullptr, 0, h->ws, h->ws_floats, s);
    float* nobj = lastl ? obj_out : h->layers[li + 1].x_obj;
    float* npred = lastl ? pred_out : h->layers[li + 1].x_pred;
    if (L.residual) {   // graph.py:205-209
      bn_fwd(L.z4, N, L.bn[3], h->a4, h->eps, h->momentum, h->bn_part, s);
      lin_fwd(L.x_obj, L.din, N, L.proj, nobj, L.dout, h->a4, L.dout, h->ws, h->ws_floats, s);
      lin_fwd(L.x_pred, dp, T, L.projp, npred, dp, L.a2 + H, W2, h->ws, h->ws_floats, s);
    } else {
      bn_fwd(L.z4, N, L.bn[3], nobj, h->eps, h->momentum, h->bn_part, s);
      copy_cols(L.a2 + H, W2, T, dp, npred, dp, s);
    }
    bump_counter_kernel<<<1, 1, 0, s>>>(L.bn[0].nbt, L.bn[1].nbt, L.bn[2].nbt, L.bn[3].nbt);
    ECHO_LAUNCH_CHECK();
    cur_obj = nobj;
    cur_pred = npred;
  }
  h->saved_nodes = N; h->saved_triples = T; h->saved_graph = g->id;
}

// dW += dY^T X, db += colsum(dY), dX = dY W.  Up to 64 rows (one scene): the per-element kernels above, whose grids grow with the weight.
// Beyond (collated batches): the 3 x TF32 tensor-core GEMM (sgemm_x3.cu) with transposed operand strides; the weight gradient's
// reduction over rows is split into chunks over blockIdx.z so that small weights still fill the GPU, and the partial sums are added
// in chunk order.
static void lin_bwd(const float* dY, int64_t ldy, const float* X, int64_t ldx, int M, const TLin& l, float* dX, int64_t ldxo, float* ws,
                    size_t ws_floats, cudaStream_t s) {
  if (M == 0) return;
  colsum_acc_kernel<<<cdiv(l.nout, 32), dim3(32, 32), 0, s>>>(dY, ldy, M, l.nout, l.gb);
  ECHO_LAUNCH_CHECK();
  if (M <= 64) {
    linear_wgrad_kernel<<<dim3(cdiv(l.K, 32), cdiv(l.nout, 32)), 256, 0, s>>>(dY, ldy, X, ldx, M, l.nout, l.K, l.gw);
    ECHO_LAUNCH_CHECK();
    if (dX) {
      linear_dgrad_kernel<<<dim3(cdiv(l.K, DG_TK), cdiv(M, DG_TM)), DG_TK, 0, s>>>(dY, ldy, l.w, M, l.nout, l.K, dX, ldxo, 0);
      ECHO_LAUNCH_CHECK();
    }
    return;
  }
  // wgrad: dW[n_out, k_in] += sum_r dY[r, n_out] X[r, k_in]   (A = dY read transposed); the reduction over rows in `splits` chunks
  const int64_t wn = (int64_t)l.nout * l.K;
  const int tiles = cdiv(l.nout, 128) * cdiv(l.K, 64);
  int splits = std::max(1, std::min({cdiv(2 * 148, tiles), M / 128, (int)(ws_floats / (size_t)wn), 32}));
  const int chunk = (cdiv(M, splits) + 15) & ~15;       // equal chunks of whole k-tiles; the last one is the shorter remainder
  splits = cdiv(M, chunk);
  SgemmX3Args g;
  g.A = dY; g.sam = 1; g.sak = ldy; g.B = X; g.sbk = ldx; g.sbn = 1; g.M = l.nout; g.N = l.K; g.K = M; g.ldc = l.K;
  if (splits == 1) {
    g.C = l.gw; g.res = l.gw; g.ld_res = l.K;            // accumulate in place
  } else {
    g.C = ws; g.splits = splits; g.chunk = chunk; g.c_bs = wn;
  }
  SgemmX3Args d;   // dgrad: dX[r, k_in] = sum_n dY[r, n] W[n, k_in]
  d.A = dY; d.sam = ldy; d.sak = 1; d.B = l.w; d.sbk = l.K; d.sbn = 1; d.C = dX; d.ldc = ldxo; d.M = M; d.N = l.K; d.K = l.nout;
  ECHO_CHECK(sgemm_x3_supported(g) && (!dX || sgemm_x3_supported(d)), "gcn_train: backward GEMM operands outside sgemm_x3's contract "
             "(widths must be multiples of 4)");
  sgemm_x3(g, s);
  if (splits > 1) {
    splitk_acc_kernel<<<cdiv(wn, 256), 256, 0, s>>>(ws, wn, splits, l.gw);
    ECHO_LAUNCH_CHECK();
  }
  if (dX) sgemm_x3_auto(d, ws, ws_floats, s);   // (the workspace is reused in stream order: splitk_acc has consumed it by then)
}

static void bn_bwd(const float* da, int64_t ldd, const float* z, int rows, const TBn& b, float* dz, float* part, cudaStream_t s) {
  if (rows <= 2 * BN_RB) {
    bn_relu_bwd_kernel<<<cdiv(b.C, 32), dim3(32, BN_LANES), 0, s>>>(da, ldd, z, b.C, rows, b.C, b.g, b.b, b.mean, b.rstd, dz, b.C, b.gg, b.gb);
  } else {
    const int nrb = cdiv(rows, BN_RB);
    float *p1 = part, *p2 = part + (size_t)nrb * b.C;
    bn_bwd_partial_kernel<<<dim3(cdiv(b.C, 32), nrb), dim3(32, 32), 0, s>>>(da, ldd, z, b.C, rows, b.C, b.g, b.b, b.mean, b.rstd, p1, p2);
    ECHO_LAUNCH_CHECK();
    bn_bwd_apply_kernel<<<dim3(cdiv(b.C, 32), nrb), dim3(32, 32), 0, s>>>(da, ldd, z, b.C, rows, b.C, b.g, b.b, b.mean, b.rstd, p1, p2, nrb,
                                                                           dz, b.C, b.gg, b.gb);
  }
  ECHO_LAUNCH_CHECK();
}

void gcn_train_backward(echo_gcn_train* h, const echo_graph* g, const float* d_obj_out, const float* d_pred_out, float* d_obj_in,
                        float* d_pred_in, cudaStream_t s) {
  ECHO_CHECK(h && g && d_obj_out, "gcn_train_backward: null argument");
  const int N = g->n_nodes, T = g->n_triples, H = h->H, dp = h->dp, W2 = 2 * H + dp;
  ECHO_CHECK(h->saved_graph == g->id && h->saved_nodes == N && h->saved_triples == T,
             "gcn_train_backward: call echo_gcn_train_forward on this graph first (the backward reads what it saved)");
  const float* dO = d_obj_out;
  const float* dP = d_pred_out;   // null: the predicate output does not reach the loss (zero cotangent)
  for (int li = (int)h->layers.size() - 1; li >= 0; --li) {
    TLayer& L = h->layers[li];
    const int k1 = 2 * L.din + dp;
    float* dobj = li == 0 ? d_obj_in : h->dobj_pp[li & 1];
    float* dpred = li == 0 ? d_pred_in : h->dpred_pp[li & 1];
    float* dobj_w = dobj ? dobj : h->dobj_pp[li & 1];
    const bool want_in = li > 0 || d_obj_in != nullptr;    // layer 0: the input gradients are optional (leaf inputs of the caller)
    const bool want_pin = li > 0 || d_pred_in != nullptr;
    if (!dP) {   // materialise the zero cotangent of the last layer's predicate output
      ECHO_CUDA(cudaMemsetAsync(h->dpred_pp[(li + 1) & 1], 0, sizeof(float) * T * dp, s));
      dP = h->dpred_pp[(li + 1) & 1];
    }
    // residual projections (graph.py:205-209): they carry dO / dP straight to the layer's inputs
    if (L.residual) {
      lin_bwd(dO, L.dout, L.x_obj, L.din, N, L.proj, want_in ? dobj_w : nullptr, L.din, h->ws, h->ws_floats, s);
      lin_bwd(dP, dp, L.x_pred, dp, T, L.projp, want_pin ? dpred : nullptr, dp, h->ws, h->ws_floats, s);
    } else {
      if (want_in) ECHO_CUDA(cudaMemsetAsync(dobj_w, 0, sizeof(float) * N * L.din, s));
      if (want_pin) ECHO_CUDA(cudaMemsetAsync(dpred, 0, sizeof(float) * T * dp, s));
    }
    // net2
    bn_bwd(dO, L.dout, L.z4, N, L.bn[3], h->dbuf_n, h->bn_part, s);                                  // dz4 (N x dout)
    lin_bwd(h->dbuf_n, L.dout, L.a3, H, N, L.l4, h->dpooled, H, h->ws, h->ws_floats, s);               // da3 (N x H) in dpooled
    bn_bwd(h->dpooled, H, L.z3, N, L.bn[2], h->dpooled, h->bn_part, s);                              // dz3 in place
    lin_bwd(h->dpooled, H, L.pooled, H, N, L.l3, h->dbuf_n, H, h->ws, h->ws_floats, s);                // dpooled (N x H) in dbuf_n
    // pooling + predicate split
    pool_bwd_kernel<<<T, 128, 0, s>>>(h->dbuf_n, dP, g->s_idx, g->o_idx, g->node_off, T, H, dp, h->dbuf_t);   // da2 (T x W2)
    ECHO_LAUNCH_CHECK();
    // net1
    bn_bwd(h->dbuf_t, W2, L.z2, T, L.bn[1], h->dbuf_t, h->bn_part, s);                               // dz2 in place
    lin_bwd(h->dbuf_t, W2, L.a1, H, T, L.l2, h->dtin, H, h->ws, h->ws_floats, s);                      // da1 (T x H) in dtin
    bn_bwd(h->dtin, H, L.z1, T, L.bn[0], h->dtin, h->bn_part, s);                                    // dz1 in place
    // dz1 is needed as dY while dTin is written: move it out of the way
    ECHO_CUDA(cudaMemcpyAsync(h->dbuf_t, h->dtin, sizeof(float) * T * H, cudaMemcpyDeviceToDevice, s));
    const bool need_tin = want_in || want_pin;
    lin_bwd(h->dbuf_t, H, L.tin, k1, T, L.l1, need_tin ? h->dtin : nullptr, k1, h->ws, h->ws_floats, s);   // dTin (T x k1)
    if (want_in) {
      gather_bwd_obj_kernel<<<N, 128, 0, s>>>(h->dtin, k1, L.din, dp, g->node_off, g->node_items, N, dobj_w);
      ECHO_LAUNCH_CHECK();
    }
    if (want_pin) {
      add_cols_kernel<<<cdiv((int64_t)T * dp, 256), 256, 0, s>>>(h->dtin + L.din, k1, T, dp, dpred, dp);
      ECHO_LAUNCH_CHECK();
    }
    dO = dobj_w;
    dP = dpred;
  }
}

}  // namespace echo
